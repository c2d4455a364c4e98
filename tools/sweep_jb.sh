for b in 4 2; do HULK_B200_JUMP_BATCH=$b python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/bench_jb$b.log 2>&1; done
