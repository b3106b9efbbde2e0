#!/usr/bin/env python
"""Parity at scale: run the GPU chain on N config-5 items (same generator as bench.py), then re-run the
oracle on (a) every item the GPU did not decode to exactly one 1500-byte MPDU and (b) a random sample of
decoded ones, and compare status / PDU bytes.  Run on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import oracle_lib as ol  # noqa: E402
from __graft_entry__ import load_pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
pkg = load_pkg()
dev = torch.device("cuda", 0)
iq, mpdus = bench.make_batch_device(torch, dev, n, seed=0)
rx = pkg.Receiver(device=0, chunk_items=32768)
off = np.arange(n, dtype=np.int64) * bench.ITEM
ln = np.full(n, bench.ITEM, np.int32)
fr, pdu = rx.rx_batch_dev(iq.data_ptr(), off, ln, pdu_stride=bench.PDU_STRIDE)
good = (fr["status"] == 0) & (fr["npdu"] == 1) & (fr["pdu_bytes"] == 1504)
bad = np.nonzero(~good)[0]
rng = np.random.default_rng(0)
sample = rng.choice(np.nonzero(good)[0], size=min(512, int(good.sum())), replace=False)
idx = np.concatenate([bad, sample])
print("GPU: %d/%d decoded; checking %d failed + %d decoded items against the oracle" % (good.sum(), n, bad.size, sample.size))
h = iq.view(n, bench.ITEM)[torch.from_numpy(idx).to(dev)].cpu().numpy()
mism = 0
for k, i in enumerate(idx):
    fo, llr, po = ol.rx_item(np.ascontiguousarray(h[k]), max_frames=1)
    same = fo[0]["status"] == fr[i]["status"] and fo[0]["npdu"] == fr[i]["npdu"] and po.size == fr[i]["pdu_bytes"] and bytes(po) == bytes(pdu[i, :po.size])
    for key in ("sync_idx", "trig_idx", "format", "mcs", "len", "nsym", "trellis"):
        same = same and fo[0][key] == fr[i][key]
    if not same:
        mism += 1
        print("MISMATCH item %d: gpu status %d npdu %d sync %d | oracle status %d npdu %d sync %d" %
              (i, fr[i]["status"], fr[i]["npdu"], fr[i]["sync_idx"], fo[0]["status"], fo[0]["npdu"], fo[0]["sync_idx"]))
st, cnt = np.unique(fr["status"][bad], return_counts=True)
print("failed-item GPU statuses:", dict(zip(st.tolist(), cnt.tolist())))
print("RESULT: %d mismatches vs oracle out of %d checked" % (mism, idx.size))
