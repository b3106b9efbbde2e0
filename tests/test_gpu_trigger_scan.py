"""The trigger scan of the batched path (bitmap words, bulk updates between the two samples where the FSM changes course,
traceless words stepped over, idle stretches skipped -- walk_word / k_trig_scan in csrc/k_frontend_w.cu) against a plain
per-sample model of lib/trigger_impl.cc:59-117 + sync's hold-off (lib/sync_impl.cc:94,141-146), on adversarial preac
sequences: plateaus of every length around 21, gaps of every length around 80 / 111, runs across word boundaries, ties of
the running maximum, NaNs, values exactly at the threshold."""
import numpy as np
import pytest

from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu
T = np.float32(0.3)


def model(p, start=0):
    """events (trig, latch, safe, stall) and the restart point at the end, sample by sample"""
    n = p.size
    nP = fP = fE = cd = 0
    conj = np.float32(0.0)
    latch, skip, safe, out, done = -1, 0, start, [], False
    for i in range(start, n):
        if nP == 0 and fP == 0 and i >= skip:
            safe = i
        ac, fl = p[i], 0
        if ac > T:
            nP += 1
            if ac > conj:
                conj = ac
                fl |= 2
            if nP > 20 and fP + fE == 0:
                fP, fE, cd = 1, 1, 80
        else:
            nP, fE, conj = 0, 0, np.float32(0.0)
        if fP:
            cd -= 1
            if cd == 0:
                fP = 0
                fl |= 1
        if fl == 0 or i < skip:
            continue
        if fl & 1:
            stall = int(n - i < 240)
            out.append((i, latch, safe, stall))
            if stall:
                done = True
                break
            skip = i + 111
        elif fl & 2:
            latch = i
    if not done and nP == 0 and fP == 0 and n >= skip:
        safe = n
    return out, safe


def adversarial(rng, n):
    p = np.zeros(n, np.float32)
    i = 0
    while i < n:
        kind = rng.integers(0, 10)
        if kind < 3:                                   # quiet stretch, sometimes with NaNs and threshold-equal values
            L = int(rng.choice([1, 2, 3, 31, 32, 33, 64, 79, 80, 81, 110, 111, 112, 200, 500]))
            v = rng.uniform(0.0, 0.3, L).astype(np.float32)
            v[rng.random(L) < 0.1] = np.nan
            v[rng.random(L) < 0.1] = T
        elif kind < 6:                                 # sporadic short runs in noise
            L = int(rng.integers(20, 300))
            v = np.where(rng.random(L) < 0.25, rng.uniform(0.31, 0.6, L), rng.uniform(0.0, 0.29, L)).astype(np.float32)
        else:                                          # plateau
            L = int(rng.choice([1, 5, 19, 20, 21, 22, 40, 99, 100, 101, 102, 130, 160, 180, 260]))
            v = rng.uniform(0.31, 0.9, L).astype(np.float32)
            if rng.random() < 0.5:                     # ties of the maximum, quantised values
                v = (np.round(v * 8) / 8 + 0.01).astype(np.float32)
            if rng.random() < 0.3:                     # a dip of one or two samples inside
                k = int(rng.integers(0, L))
                v[k:k + int(rng.integers(1, 3))] = 0.1
        v = v[: n - i]
        p[i:i + v.size] = v
        i += v.size
    return p


def test_trigger_scan_matches_the_per_sample_model():
    pkg = load_pkg()
    rx = pkg.Receiver(device=0)
    rng = np.random.default_rng(2024)
    nev = 0
    for case in range(1200):
        n = int(rng.choice([240, 333, 1000, 1024, 2047, 4096, 6000]))
        p = adversarial(rng, n)
        start = int(rng.choice([0, 0, 0, 17, 32, 64, 95]))
        want, wsafe = model(p, start)
        ev, safe_end = rx.trigger_events(p, start)
        got = [tuple(int(x) for x in e) for e in ev]
        assert got == want, (case, n, start, got[:6], want[:6])
        assert safe_end == wsafe, (case, n, start, safe_end, wsafe)
        nev += len(want)
    assert nev > 400
    # the reference generator's STF itself: preac of a real capture through the staged presiso
    g = np.load(__file__.rsplit("/", 1)[0] + "/golden/frames_siso.npz")
    x = g["iq"][g["offs"][1]:g["offs"][12]]
    x = (x + 0.02 * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
    pre, _ = rx.presiso(x)
    want, wsafe = model(pre)
    ev, safe_end = rx.trigger_events(pre)
    assert [tuple(int(v) for v in e) for e in ev] == want and len(want) >= 11 and safe_end == wsafe
    rx.close()
