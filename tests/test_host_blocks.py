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


def _tags(pkg, *recs):
    t = np.zeros(len(recs), pkg._cabi.TAG_DTYPE)
    for k, (idx, fields, nvec) in enumerate(recs):
        t[k]["idx"], t[k]["nvec"] = idx, nvec
        for name, v in fields.items():
            t[k]["f"][name] = v
    return t


def test_block_error_paths_follow_the_reference(capfd):
    """what the blocks do off the happy path (lib/signal_impl.cc:96-100, lib/demod_impl.cc:72-103, lib/decode_impl.cc:93-127,141-157)"""
    pkg = load_pkg()
    B = pkg.blocks
    be = hs.HostBackend()
    none = np.zeros(0, pkg._cabi.TAG_DTYPE)
    # signal: a sync flag without a tag is reported, skipped (consumed up to and including it) and the block keeps searching
    sg = B.Block(B.SIGNAL, be)
    sync = np.zeros(50, np.uint8)
    sync[5] = 1
    c, outs, tg, msg = sg.work(50, [sync, np.zeros(50, np.complex64)], none)
    assert (c, outs[0].size, tg.size) == (6, 0, 0)
    assert "ieee80211 signal, error: input sync with no tag." in capfd.readouterr().out
    # ... and an L-SIG that fails its checks costs 80 samples (S_DEMOD -> S_TRIGGER, :156-160)
    sync = np.zeros(400, np.uint8)
    sync[0] = 1
    c, outs, tg, msg = sg.work(400, [sync, np.zeros(400, np.complex64)], _tags(pkg, (0, {"rad": 0.0, "snr": 10.0, "rssi": 1.0}, 0)))
    assert (c, outs[0].size, tg.size) == (80, 0, 0)
    sg.close()
    # demod: no tag at the first item -> waits (consumes nothing, like DEMOD_S_RDTAG)
    dm = B.Block(B.DEMOD, be)
    c, outs, tg, msg = dm.work(100, [np.zeros(100, np.complex64)], none)
    assert (c, outs[0].size) == (0, 0)
    dm.close()
    # decode: length over DECODE_B_MAX -> the frame's soft bits are swallowed, nothing is published
    dc = B.Block(B.DECODE, be)
    big = {"format": 0, "len": 5000, "total": 300, "cr": 0, "mcs": 0, "ampdu": 0, "trellis": 40022}
    c, _, tg, msg = dc.work(0, [np.zeros(100, np.float32)], _tags(pkg, (0, big, 0)))
    assert (c, msg) == (0, b"")
    c1, _, _, m1 = dc.work(0, [np.zeros(100, np.float32)], none)
    c2, _, _, m2 = dc.work(0, [np.zeros(400, np.float32)], none)
    assert (c1, c2, m1, m2) == (100, 200, b"", b"")
    # ... an NDP tag (trellis 0) publishes the channel report [20][0][4][128 x (re, im)] and swallows 1024 floats
    ndp = _tags(pkg, (0, {"format": 2, "len": 0, "total": 1024, "trellis": 0}, 128))
    ndp[0]["vec"][:] = np.arange(256, dtype=np.float32)
    c, _, _, msg = dc.work(0, [np.zeros(2000, np.float32)], ndp)
    assert c == 0 and len(msg) == 1027 and msg[:3] == bytes([20, 0, 4])
    assert np.array_equal(np.frombuffer(msg[3:], np.float32), np.arange(256, dtype=np.float32))
    c, _, _, msg = dc.work(0, [np.zeros(2000, np.float32)], none)
    assert (c, msg) == (1024, b"")
    # ... and back in IDLE it waits for the next tag
    c, _, _, msg = dc.work(0, [np.zeros(50, np.float32)], none)
    assert (c, msg) == (0, b"")
    dc.close()


def test_chain_2x2_equals_batch_detect(golden):
    """signal2 / demod2 (lib/signal2_impl.cc:63-212, lib/demod2_impl.cc:58-348): antenna 0 drives detection, both antennas are copied"""
    pkg = load_pkg()
    H = hs.lib()
    g = golden["frames_mimo"]
    a, b = np.ascontiguousarray(g["iq0"]), np.ascontiguousarray(g["iq1"])
    preac, preconj = _presiso(a)
    maxf = 32
    f = np.zeros(maxf, pkg.FRAME_DTYPE)
    chan = np.zeros(128 * maxf, np.float32)
    H.hs_detect(ol.c2f(a), preac, a.size, 0, maxf, f.ctypes.data, chan)
    keep = (f["status"] != 9) & (f["nsamp"] > 0)
    want, wchan = f[keep], chan.reshape(maxf, 128)[keep]
    for k in range(want.size):
        hinv, w2 = np.zeros(128, np.float32), np.zeros(528, np.float32)
        H.hs_header2(ol.c2f(a), ol.c2f(b), want[k:k + 1].ctypes.data, wchan[k], hinv, w2)
    assert want.size >= 16

    ch = pkg.blocks.Chain(nant=2, backend=hs.HostBackend(), seed=5, max_call=3000)
    msgs = ch.run(preac, preconj, a, b)
    ch.close()
    s0, s1 = np.concatenate(ch.trace["signal"]), np.concatenate(ch.trace["signal1"])
    off = 0
    assert len(ch.tags["signal"]) == want.size
    for k, (o, t) in enumerate(ch.tags["signal"]):
        fr = want[k]
        n, st = int(fr["nsamp"]), int(fr["sync_idx"]) + 224
        assert o == off and (t["f"]["l_mcs"], t["f"]["l_len"], t["f"]["nsamp"]) == (fr["l_mcs"], fr["l_len"], fr["nsamp"])
        ph = ((np.arange(n, dtype=np.float32) + np.float32(224)) * np.float32(fr["rad"])).astype(np.float64)
        assert np.allclose(s0[off:off + n], a[st:st + n] * np.exp(1j * ph), atol=2e-6)
        assert np.allclose(s1[off:off + n], b[st:st + n] * np.exp(1j * ph), atol=2e-6)
        off += n + 320
    assert s0.size == s1.size == off
    dem = want[want["status"] == 0]
    assert len(ch.tags["demod"]) == dem.size == len(msgs)
    for (o, t), fr in zip(ch.tags["demod"], dem):
        for key in ("format", "mcs", "len", "cr", "nss", "trellis", "total"):
            assert t["f"][key] == fr[key], (key, t["f"][key], fr[key])
    assert sum(x.size for x in ch.trace["llr"]) == int(dem["total"].sum())
