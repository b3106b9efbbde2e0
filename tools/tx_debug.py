#!/usr/bin/env python
"""per-slot error of the transmit synthesiser against the generator's golden waveforms (debug aid)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
g = np.load(ROOT + "/tests/golden/frames_siso.npz"); t = np.load(ROOT + "/tests/golden/frames_tx.npz")
o = np.cumsum(np.r_[0, t["psdu_len"]]); ps = [bytes(t["psdu"][o[i]:o[i + 1]]) for i in range(len(t["psdu_len"]))]
fmt = g["meta"][:, 0].astype(np.int32); mcs = g["meta"][:, 1].astype(np.int32); cfo = g["meta"][:, 2].astype(np.float32)
rx = pkg.Receiver(device=0)
iq, offs = rx.tx_batch(ps, fmt, mcs, gap=t["gap"], cfo=cfo)
sel = [int(a) for a in sys.argv[1:]] or range(len(ps))
for i in sel:
    gp = int(t["gap"][i]); a = iq[offs[i] + gp:offs[i + 1] - gp]; b = g["iq"][offs[i] + gp:offs[i + 1] - gp]
    ns = a.size // 80
    e = np.abs(a - b).reshape(ns, 80)
    bad = [(s, round(float(e[s].max()), 4), int(e[s].argmax())) for s in range(ns) if e[s].max() > 1e-5]
    print(i, "fmt", fmt[i], "mcs", mcs[i], "cfo", cfo[i], "slots", ns, "bad", bad[:14], "..." if len(bad) > 14 else "")
