#!/bin/bash
# Round 2, twenty-third GPU pass (1 GPU): one-word scan for k = 9, 11 (parity), then BASELINE's C5 grid with the round-2 kernels.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "minimizer or histogram or no_decay or intervals or drift or packed" > gpurun_out/pytest_k11.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_k11.log
python - <<'PY'
# k = 9 as well (no other test uses it): spectrum against the oracle on reads with N, lower case and ragged lengths
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, hulk_b200 as hb
from oracle import oracle as O
from conftest import random_reads
for k in (9, 11):
    reads = random_reads(20000, 40, seed=k, n_frac=0.01, lower_frac=0.2, ragged=200)
    bases, offs = O.pack_reads(reads)
    h, nm = O.count_reads(k, 9, k ** 4, bases, offs)
    with hb.HistoSketch(k, 9, 4) as hs:
        hs.add_reads(bases, offs)
        assert (hs.histogram() == h.astype(np.uint32)).all() and hs.stats()["n_minimizers"] == nm
    print("k=%d spectrum ok (%d minimizers)" % (k, nm))
PY
bash tools/sweep_c5.sh > gpurun_out/r02o_c5_sweep.txt 2>&1; cat gpurun_out/r02o_c5_sweep.txt
