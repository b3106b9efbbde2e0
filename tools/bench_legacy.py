#!/usr/bin/env python
"""Per-kernel device times of BASELINE config 2 (11a MCS0-7, 564-byte MPDUs, 30 dB) through c8b_rx_batch_dev.
usage: python tools/bench_legacy.py [frames_per_mcs]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402

pkg = load_pkg()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = np.load(os.path.join(ROOT, "tests", "golden", "frames_564.npz"))
dev = torch.device("cuda", 0)
(iq,), off, ln, kind = pkg.synth.make_items(torch, dev, [g["l%d" % m] for m in range(8)], [per] * 8, snr_db=30.0, seed=2)
rx = pkg.Receiver(device=0, overlap=False)
rx.rx_batch_dev(iq.data_ptr(), off, ln, pdu_stride=640)
rx.timing(True)
rx.timing_read(reset=True)
fr, pdu = rx.rx_batch_dev(iq.data_ptr(), off, ln, pdu_stride=640)
st = rx.timing_read(reset=True)
rx.close()
n = len(off)
dev_ms = sum(v[0] for v in st.values())
print("config 2: %d frames, %d samples, %d decoded" % (n, int(ln.sum()), int((fr["npdu"] == 1).sum())))
ok = fr["status"] == 0
alg = float((fr["nsym"][ok].astype(np.int64) * (640 + 4 * fr["ncbps"][ok].astype(np.int64))).sum())     # k_demod: 640 B in + 4 nCBPS out per symbol
print("k_demod: %.2f GB algorithmic in %.3f ms = %.0f GB/s" % (alg / 1e9, st["demod"][0], alg / st["demod"][0] / 1e6))
print("device ms per stage (launches):", {k: (round(v[0], 3), v[1]) for k, v in st.items()},
      "sum %.2f ms = %.2f M frames/s, %.2f G samples/s kernel-only" % (dev_ms, n / dev_ms / 1e3, ln.sum() / dev_ms / 1e6))
