#!/usr/bin/env python
"""Per-kernel device times of the 2x2 path (BASELINE config 4: HT MCS8-15, 564-byte MPDUs, 50 000 frames, 30 dB) through
c8b_rx_batch2 (host buffers).  usage: python tools/bench_mimo.py [frames_per_mcs]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402

pkg = load_pkg()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 6250
g = np.load(os.path.join(ROOT, "tests", "golden", "frames_564.npz"))
dev = torch.device("cuda", 0)
frames = [(g["h%d_0" % m], g["h%d_1" % m]) for m in range(8, 16)]
(a, b), off, ln, kind = pkg.synth.make_items(torch, dev, frames, [per] * 8, snr_db=30.0, seed=4, rms=0.1875)
ha, hb = a.cpu().numpy(), b.cpu().numpy()
rx = pkg.Receiver(device=0, overlap=False)
rx.rx_batch2(ha, hb, off, ln, pdu_stride=640)
rx.timing(True)
rx.timing_read(reset=True)
t0 = time.perf_counter()
fr, pdu = rx.rx_batch2(ha, hb, off, ln, pdu_stride=640)
dt = time.perf_counter() - t0
st = rx.timing_read(reset=True)
rx.close()
n = len(off)
dev_ms = sum(v[0] for v in st.values())
print("config 4: %d frames, %d samples per antenna, %d decoded; host-buffer call %.1f ms (%.2f M frames/s, %.2f G samples/s per antenna)" %
      (n, ha.size, int((fr["npdu"] == 1).sum()), 1e3 * dt, n / dt / 1e6, ha.size / dt / 1e9))
print("device ms per stage (launches):", {k: (round(v[0], 3), v[1]) for k, v in st.items()}, "sum %.2f ms = %.2f M frames/s kernel-only" % (dev_ms, n / dev_ms / 1e3))
