"""The seven blocks one scheduler call at a time (gr-ieee80211_b200/csrc/blocks.h = the state machines c8b_blk_work runs),
over the host build of the per-frame routines: whatever sizes the scheduler calls them with, trigger / sync / signal / demod
must put the flags, tags and frames where ONE pass of the batch detect over the whole capture (detect_item, checked against
the oracle in test_host_logic.py) finds them.  Soft bits and PDU bytes are placeholders here -- the GPU test
(test_gpu_blocks.py) checks those."""
import numpy as np
import pytest

import hostsim_lib as hs
import oracle_lib as ol
from __graft_entry__ import load_pkg


def _presiso(x):
    O = ol.oracle()
    n = x.size
    preac, preconj = np.zeros(n, np.float32), np.zeros(2 * n, np.float32)
    O.orx_presiso(ol.c2f(x), n, preac, preconj)
    return preac, preconj.view(np.complex64)


def _batch(pkg, x, preac, maxf=40):
    H = hs.lib()
    f = np.zeros(maxf, pkg.FRAME_DTYPE)
    chan = np.zeros(128 * maxf, np.float32)
    H.hs_detect(ol.c2f(x), preac, x.size, 0, maxf, f.ctypes.data, chan)
    keep = (f["status"] != 9) & (f["nsamp"] > 0)
    f, chan = f[keep], chan.reshape(maxf, 128)[keep]
    for k in range(f.size):
        if f[k]["status"] == 0:
            hinv = np.zeros(128, np.float32)
            H.hs_header(ol.c2f(x), f[k:k + 1].ctypes.data, chan[k], 0, hinv)
    return f, chan


@pytest.mark.parametrize("seed,max_call", [(1, 4096), (2, 700), (3, 8192)])
def test_chain_equals_batch_detect(golden, seed, max_call):
    pkg = load_pkg()
    g = golden["frames_siso"]
    x = np.ascontiguousarray(g["iq"][g["offs"][0]:g["offs"][12]])
    rng = np.random.default_rng(seed)
    x = (x + 0.003 * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
    preac, preconj = _presiso(x)
    want, wchan = _batch(pkg, x, preac)
    assert want.size == 12 and np.all(want["status"] == 0)

    ch = pkg.blocks.Chain(nant=1, backend=hs.HostBackend(), seed=seed, max_call=max_call)
    msgs = ch.run(preac, preconj, x)
    ch.close()

    # trigger: the flag stream of one pass of the FSM over the whole array
    trig = np.concatenate(ch.trace["trigger"])
    ref = np.zeros(x.size, np.uint8)
    hs.lib().hs_trigger(preac, x.size, ref)
    assert trig.size == x.size and np.array_equal(trig, ref)
    # sync: one flag + tag per frame at the batch path's sync index
    sync = np.concatenate(ch.trace["sync"])
    assert np.array_equal(np.flatnonzero(sync), np.array([o for o, _ in ch.tags["sync"]]))
    got_sync = {o: t for o, t in ch.tags["sync"]}
    for f in want:
        t = got_sync[int(f["sync_idx"])]
        assert t["f"]["rad"] == f["rad"] and t["f"]["snr"] == f["snr"] and t["f"]["rssi"] == f["rssi"]
    # signal: frames back to back (nsamp + 320 each), tag at the first copied sample, the batch path's L-SIG fields and channel
    off = 0
    sig = np.concatenate(ch.trace["signal"])
    assert len(ch.tags["signal"]) == want.size
    for k, (o, t) in enumerate(ch.tags["signal"]):
        f = want[k]
        assert o == off and t["seq"] == k + 1 and t["nvec"] == 64
        for key in ("l_mcs", "l_len", "nsamp", "rad", "snr", "rssi"):
            assert t["f"][key] == f[key], (k, key)
        assert np.array_equal(t["vec"][:128], wchan[k])
        n = int(f["nsamp"])
        s0 = int(f["sync_idx"]) + 224
        ph = (np.arange(n, dtype=np.float32) + np.float32(224)) * np.float32(f["rad"])
        rot = x[s0:s0 + n] * np.exp(1j * ph.astype(np.float64))
        assert np.allclose(sig[off:off + n], rot, atol=2e-6)
        assert not sig[off + n:off + n + 320].any()
        off += n + 320
    assert sig.size == off
    # demod: tag at the first soft bit of each frame, `total` floats per frame
    off = 0
    assert len(ch.tags["demod"]) == want.size
    for k, (o, t) in enumerate(ch.tags["demod"]):
        f = want[k]
        assert o == off
        for key in ("format", "mcs", "len", "cr", "ampdu", "trellis", "total", "nss", "nsym"):
            assert t["f"][key] == f[key], (k, key, t["f"][key], f[key])
        assert t["f"]["snr"] == f["snr"] and abs(float(t["f"]["cfo_hz"]) - float(f["cfo_hz"])) < 1e-3
        off += int(f["total"])
    assert sum(a.size for a in ch.trace["llr"]) == off
    # decode: one (placeholder) PDU per frame, in order
    assert len(msgs) == want.size
    for m, f in zip(msgs, want):
        assert m[0] == f["format"] and (m[1] | (m[2] << 8)) == f["len"] and m[-1] == f["mcs"] and len(m) == f["len"] + 4


def test_chain_noise_and_junk_do_not_stall(golden):
    """noise, a periodic-16 burst (false triggers, L-SIG failures) and a frame cut off at the end of the capture: every block
    keeps consuming; a frame whose samples never arrive stays pending like in the reference."""
    pkg = load_pkg()
    g = golden["frames_siso"]
    rng = np.random.default_rng(9)
    t = np.tile((rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex64), 60)
    junk = np.concatenate([np.zeros(300, np.complex64), 0.2 * t, np.zeros(900, np.complex64)])
    fr = g["iq"][g["offs"][3]:g["offs"][4]]
    x = np.concatenate([junk, fr, junk, fr[:1500]]).astype(np.complex64)
    x = (x + 0.002 * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
    preac, preconj = _presiso(x)
    want, _ = _batch(pkg, x, preac)
    ch = pkg.blocks.Chain(nant=1, backend=hs.HostBackend(), seed=4, max_call=1024)
    msgs = ch.run(preac, preconj, x)
    ch.close()
    assert np.concatenate(ch.trace["trigger"]).size == x.size
    full = [f for f in want if f["status"] == 0 and f["sync_idx"] + 224 + f["nsamp"] + 320 <= x.size]
    assert len(msgs) == len(full) >= 1
    got = [(int(t["f"]["l_mcs"]), int(t["f"]["l_len"])) for _, t in ch.tags["signal"]]
    assert got[:len(full)] == [(int(f["l_mcs"]), int(f["l_len"])) for f in full]
