#!/usr/bin/env python
"""BASELINE config 4 (HT 2x2 SU-MIMO, MCS8-15, 564-byte MPDUs, 50 000 frames, 30 dB) with UNIQUE traffic: every frame carries its
own random MPDU, modulated on the device by the two-stream transmit synthesiser (c8b_tx_batch2_dev: stream k on antenna k,
the identity channel of tools/performance/gr_sumimo.py:70-78, independent AWGN per antenna), received by c8b_rx_batch2_dev and
every decoded MPDU compared with the bytes that were sent.  Prints the per-kernel device times.
usage: python tools/bench_mimo.py [frames_per_mcs] [mmse]"""
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
mmse = int(sys.argv[2]) if len(sys.argv) > 2 else 0
MPDU, GAP = 564, 300
dev = torch.device("cuda", 0)
rx = pkg.Receiver(device=0, overlap=False, mmse=mmse)
n = 8 * per
mcs = np.repeat(np.arange(8, 16, dtype=np.int32), per)
d = np.zeros(n, pkg.TXFRAME_DTYPE)
d["format"], d["mcs"], d["psdu_len"] = 1, mcs, MPDU
d["psdu_off"] = np.arange(n, dtype=np.int64) * MPDU
ns = np.array([rx.L.c8b_tx_nsamp(1, int(m), MPDU) for m in range(8, 16)], np.int64)
item = np.repeat(ns + 2 * GAP, per)
off = np.concatenate([[0], np.cumsum(item)[:-1]]).astype(np.int64)
d["out_off"] = off + GAP
d["cfo_hz"] = np.random.default_rng(4).uniform(-100e3, 100e3, n).astype(np.float32)
total = int(item.sum())
psdu = torch.zeros(n * MPDU + 16, dtype=torch.uint8, device=dev)
a = torch.zeros(total, dtype=torch.complex64, device=dev)
b = torch.zeros(total, dtype=torch.complex64, device=dev)
torch.cuda.synchronize()
rx.tx_random_psdu_dev(psdu.data_ptr(), n * MPDU, d, seed=0x80211)
t0 = time.perf_counter()
rx.tx_batch2_dev(psdu.data_ptr(), n * MPDU, d, a.data_ptr(), b.data_ptr(), total)
rx.sync()
t_tx = time.perf_counter() - t0
sigma = 0.1875 / np.sqrt(2.0 * 10 ** 3.0)                          # per antenna: LTF rms 0.1875 at multiplier 12 sqrt 2, 30 dB
for x, seed in ((a, 13579), (b, 24680)):                           # tools/performance/gr_sumimo.py:62-63
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    v = torch.view_as_real(x)
    for s0 in range(0, total, 1 << 26):
        e0 = min(total, s0 + (1 << 26))
        v[s0:e0] += torch.randn((e0 - s0, 2), generator=gen, device=dev) * sigma
torch.cuda.synchronize()
ln = item.astype(np.int32)
rx.rx_batch2_dev(a.data_ptr(), b.data_ptr(), off, ln, pdu_stride=640)
rx.timing(True)
rx.timing_read(reset=True)
t0 = time.perf_counter()
fr, pdu = rx.rx_batch2_dev(a.data_ptr(), b.data_ptr(), off, ln, pdu_stride=640)
dt = time.perf_counter() - t0
st = rx.timing_read(reset=True)
rx.close()
sent = psdu[:n * MPDU].view(n, MPDU).cpu().numpy()
ok = (fr["status"] == 0) & (fr["npdu"] == 1) & (fr["nss"] == 2)
same = ok & (pdu[:, 3:3 + MPDU] == sent).all(axis=1)
dev_ms = sum(v[0] for v in st.values())
print("config 4, unique frames: %d frames (%d samples per antenna, synthesised in %.1f ms), %d decoded, %d byte-identical to what was sent; "
      "device-resident call %.1f ms" % (n, total, 1e3 * t_tx, int(ok.sum()), int(same.sum()), 1e3 * dt))
print("device ms per stage (launches):", {k: (round(v[0], 3), v[1]) for k, v in st.items()},
      "sum %.2f ms = %.2f M frames/s, %.2f G samples/s per antenna kernel-only%s" % (dev_ms, n / dev_ms / 1e3, total / dev_ms / 1e6, " (MMSE)" if mmse else ""))
