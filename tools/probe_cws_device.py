#!/usr/bin/env python3
"""Wall clock of HistoSketch.newCWS: drawn on the device against the host generator (all cores), C2 and a C3-sized slice."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hulk_b200 as hb
for k, s in [(21, 512), (31, 256), (31, 1024)]:
    D = k ** 4
    with hb.HistoSketch(k, 9, s, 1.0) as hs:
        t0 = time.perf_counter()
        hs.generate_tables_device()
        t1 = time.perf_counter()
        print("k=%d s=%d (%.0f M elements, %.1f GB of tables): device draw %.3f s" % (k, s, s * D / 1e6, 24.0 * s * D / 1e9, t1 - t0), flush=True)
    if s * D <= 250e6:
        t0 = time.perf_counter()
        r, c, b = hb.new_cws(s, D)
        t1 = time.perf_counter()
        print("   host generator (%d cpus): %.3f s (+ upload of %.1f GB)" % (len(os.sched_getaffinity(0)), t1 - t0, 24.0 * s * D / 1e9), flush=True)
        del r, c, b
