#!/bin/bash
# Measurements that were prepared on CPU after this round's GPU minutes ran out; one gpurun call takes them all
# (about 3 minutes of box time):  gpurun --timeout 420 -- 'bash tools/pending_measurements.sh'
mkdir -p gpurun_out
nproc > gpurun_out/pending_host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/pending_host.txt
# 1. reader throughput on the GPU box's host: BGZF and ordinary gzip (csrc/pgzip.h) against the one-thread zlib path
timeout 150 python tools/probe_bgzf.py 3000000 /tmp > gpurun_out/pending_reader.txt 2>&1; tail -9 gpurun_out/pending_reader.txt
HULK_B200_PGZ_TIMING=1 timeout 60 python tools/probe_bgzf.py 1000000 /tmp 2>&1 | grep "pgzip phases" | tail -2 >> gpurun_out/pending_reader.txt
# 2. SURVEY 8(d)'s realism variant (reads from a 100 Mbp genome) next to the headline workload
timeout 200 python bench.py --steps 50 --warmup 5 --reads genome --no-cpu-baseline > gpurun_out/pending_bench_genome.json 2> gpurun_out/pending_bench_genome.err
tail -c 600 gpurun_out/pending_bench_genome.json; tail -3 gpurun_out/pending_bench_genome.err
# 3. the bench line as the driver will take it (cpu_baseline now also single-threaded, host info)
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/pending_bench.json 2> gpurun_out/pending_bench.err
tail -c 900 gpurun_out/pending_bench.json; tail -3 gpurun_out/pending_bench.err
