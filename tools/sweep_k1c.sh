for c in 5 4 3 2; do HULK_B200_K1_CTAS=$c python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/bench_k1c$c.log 2>&1; done
HULK_B200_K1_CTAS=3 python tools/probe_timeline.py > gpurun_out/timeline_k1c3.txt 2>&1
