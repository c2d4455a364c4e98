for k1 in 3 4 5; do for nb in 3 4; do for k3 in 1 2; do
  HULK_B200_K1_CTAS=$k1 HULK_B200_NBUF=$nb HULK_B200_K3_CTAS=$k3 python bench.py --steps 60 --warmup 6 --no-cpu-baseline > gpurun_out/cfg_${k1}_${nb}_${k3}.log 2>&1
done; done; done
