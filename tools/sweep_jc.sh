for j in 2 3 4 5 6; do HULK_B200_JUMP_CTAS=$j python bench.py --steps 80 --warmup 6 --no-cpu-baseline > gpurun_out/jc_$j.log 2>&1; done
