# BASELINE config C3 shape on one B200: k=31, s=1024, concept drift 0.02 (22.7 GB float64 tables + 3.8 GB K32)
python bench.py --k 31 --s 1024 --decay 0.02 --interval 100000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err
tail -c 400 gpurun_out/bench_c3.err
