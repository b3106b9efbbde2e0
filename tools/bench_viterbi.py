#!/usr/bin/env python
"""Stage micro-benchmark: decode block only (k_viterbi) on config-5-shaped frames with random soft bits.
Prints frames/s, ACS/s and the per-launch device time measured with CUDA events on the ctx stream."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_pkg  # noqa: E402

pkg = load_pkg()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T, cr, total = 12220, 3, 14664
rng = np.random.default_rng(0)
base = rng.normal(0, 2, (64, total)).astype(np.float32)
llr = np.tile(base, (n // 64, 1)).reshape(-1)
fr = np.zeros(n, pkg.FRAME_DTYPE)
fr["cr"], fr["trellis"], fr["total"], fr["format"], fr["len"], fr["mcs"] = cr, T, total, 2, 1504, 7
fr["llr_off"] = np.arange(n, dtype=np.int64) * total
rx = pkg.Receiver(device=0)
rx.timing(True)
for it in range(4):
    t0 = time.time()
    rx.decode(llr, fr, pdu_stride=1600)
    wall = time.time() - t0
    tm = rx.timing_read(reset=True)["viterbi"]
    ms = tm[0] / max(tm[1], 1)
    print("iter %d: viterbi kernel %.3f ms for %d frames -> %.0f frames/s, %.3e ACS/s, %.2f GB/s LLR read  (wall %.2f s)"
          % (it, ms, n, n / ms * 1e3, n * T * 64.0 / ms * 1e3, n * total * 4 / ms / 1e6, wall))
