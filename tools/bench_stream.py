#!/usr/bin/env python
"""Throughput of the live-stream session (c8b_stream_push) on a long capture: the reference's 25-frame demo capture
(tests/golden/frames_siso.npz items 1..25, 30 dB AWGN) repeated, pushed in general_work-sized pieces.  Real time for one
20 MHz channel is 20 M samples/s.  usage: python tools/bench_stream.py [push_samples ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402

pkg = load_pkg()
g = np.load(os.path.join(ROOT, "tests", "golden", "frames_siso.npz"))
offs = g["offs"]
one = np.ascontiguousarray(g["iq"][offs[1]:offs[26]])
reps = 64
x = np.tile(one, reps)
rng = np.random.default_rng(1)
s = 0.1875 / np.sqrt(2 * 10 ** 3.0)
x = (x + s * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
try:                                                                # a radio driver's DMA ring is page-locked: so is the source here
    import torch
    xt = torch.from_numpy(x).pin_memory()
    x = xt.numpy()
except Exception:
    pass
rx = pkg.Receiver(device=0, max_frames=512, chunk_items=1)
results = {}
for push in [int(a) for a in sys.argv[1:]] or [16384, 65536, 262144, 1048576]:
    for rep in range(2):                                            # first pass warms the scratch allocations
        rx.stream_begin(1, 1 << 22)
        rx.timing(True)
        rx.timing_read(reset=True)
        nf = npdu = 0
        t0 = time.perf_counter()
        for k in range(0, x.size, push):
            fr, base, pdu = rx.stream_push(x[k:k + push], flush=k + push >= x.size, frames_cap=1024)
            nf += fr.size
            npdu += int(fr["npdu"].sum())
        dt = time.perf_counter() - t0
        stages = rx.timing_read(reset=True)
    print("push %8d samples: %7.1f M samples/s (%.1fx real time), %d frames, %d PDUs of %d sent, %.2f ms per push" %
          (push, x.size / dt / 1e6, x.size / dt / 20e6, nf, npdu, 25 * reps, 1e3 * dt / ((x.size + push - 1) // push)))
    results[str(push)] = {"samples_per_s": x.size / dt, "frames": nf, "pdus": npdu, "sent": 25 * reps}
    print("          device ms per stage (passes):", {k: (round(v[0], 2), v[1]) for k, v in stages.items()}, "wall %.1f ms" % (1e3 * dt))
rx.close()
print("RESULT " + __import__("json").dumps(results))
