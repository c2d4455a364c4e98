#!/bin/bash
# Round 2, eighteenth GPU pass (1 GPU): what was left unmeasured -- whole jobs C3 / C4-share with the table draw inside, the
# front end's wall clock on a FASTQ file (table draw on host / device, packed transport), readers on this host, genome variant.
mkdir -p gpurun_out /tmp/cli
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 300 python tools/run_c3_c4.py > gpurun_out/r02k_c3_c4_whole_jobs.jsonl 2> gpurun_out/r02k_c3_c4.err; echo "c3/c4 rc=$?"; cat gpurun_out/r02k_c3_c4_whole_jobs.jsonl | cut -c1-500; tail -2 gpurun_out/r02k_c3_c4.err
timeout 200 python bench.py --steps 50 --warmup 5 --reads genome --no-cpu-baseline > gpurun_out/r02k_bench_genome.json 2> gpurun_out/r02k_bench_genome.err; echo "genome rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02k_bench_genome.json").read().strip().splitlines()[-1])
print("genome: value %.0f M/s %.4f ms/step, e2e %.0f M/s, minimizers/read %.2f"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["n_minimizers"]/(d["steps"]*100000)), {k[:9]:round(v,4) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
PY
python - <<'PY'
import sys, os, time
sys.path.insert(0, ".")
import numpy as np, hulk_b200
n = 4_000_000
t0 = time.time()
reads = hulk_b200.synthetic_reads(n, 150, seed=1)
rec = np.empty((n, 4 + 10 + 1 + 150 + 3 + 150 + 1), dtype=np.uint8)      # "@r" + 10 digits + "\n" + seq + "\n+\n" + qual + "\n"
rec[:, 0] = ord("@"); rec[:, 1] = ord("r"); rec[:, 2] = ord("e"); rec[:, 3] = ord("a")
idx = np.arange(n)
for d in range(10):
    rec[:, 4 + 9 - d] = ord("0") + (idx // 10 ** d) % 10
rec[:, 14] = 10
rec[:, 15:165] = reads
rec[:, 165] = 10; rec[:, 166] = ord("+"); rec[:, 167] = 10
rec[:, 168:318] = ord("I")
rec[:, 318] = 10
rec.tofile("/tmp/cli/r.fq")
print("wrote %d reads in %.1fs, %.0f MB" % (n, time.time() - t0, os.path.getsize("/tmp/cli/r.fq") / 1e6))
PY
run_cli() { tag="$1"; shift; env "$@" bash -c '{ time hulk_b200/bin/hulk sketch -f /tmp/cli/r.fq -s 512 -i 100000 -o /tmp/cli/out > /tmp/cli/log.txt ; } 2> /tmp/cli/time.txt'; echo "== $tag: $(grep real /tmp/cli/time.txt | tr '\n' ' ') $(grep -c 'reached interval' /tmp/cli/log.txt) intervals; md5 $(grep -o '"md5sum": "[0-9a-f]*"' /tmp/cli/out.json | head -1)"; }
{ run_cli "host table draw, ASCII transport" X=1
run_cli "table draw on the device" HULK_B200_CWS_DEVICE=1
run_cli "table draw on the device, packed transport" HULK_B200_CWS_DEVICE=1 HULK_B200_PACK_INPUT=1
run_cli "table draw on the device, 2nd run (file cached)" HULK_B200_CWS_DEVICE=1; } > gpurun_out/r02k_cli_wall.txt 2>&1; cat gpurun_out/r02k_cli_wall.txt
timeout 150 python tools/probe_bgzf.py 3000000 /tmp > gpurun_out/r02k_reader.txt 2>&1; tail -9 gpurun_out/r02k_reader.txt
