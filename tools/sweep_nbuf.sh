for nb in 2 3 4; do HULK_B200_NBUF=$nb python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nb$nb.log 2>&1; done
