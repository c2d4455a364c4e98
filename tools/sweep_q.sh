python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/bench_queue.log 2>&1
HULK_B200_K1_FUSED=1 python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1
HULK_B200_NBUF=2 python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/bench_queue_nb2.log 2>&1
