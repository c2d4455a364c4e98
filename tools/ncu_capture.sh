#!/bin/bash
# Runs on the GPU box (via gpurun): ncu launch list of a short bench run + one `--set full` capture of
# each hot kernel.  Reports land in gpurun_out/; summaries are extracted here (CPU box) by
# tools/ncu_summarise.py and committed under profiles/.
#   TAG=r01 bash tools/ncu_capture.sh
TAG=${TAG:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
# every launch of our kernels with its device time (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[123]_' -c 400 \
    --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
# one full capture of each hot kernel (second step of the timed region)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k1_minimizer_histogram|k1_jump_queue|k2_cms_update|k3_filter|k3_resolve' -s 20 -c 5 \
    -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
