"""The seven receive blocks one scheduler call at a time (c8b_blk_work: what the gr::block shells of gr/lib call from
general_work) wired like examples/rx.grc / rx2.grc and driven with random call sizes: the flag streams, tags, soft bits and
published PDUs must be the ones the oracle finds in ONE pass over the whole capture."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


def _noisy(x, snr, seed):
    if snr is None:
        return np.ascontiguousarray(x)
    rng = np.random.default_rng(seed)
    s = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
    return (x + s * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)


def _check_tags(ch, fo, nant):
    ok = fo[(fo["status"] != 9) & (fo["nsamp"] > 0)]
    # sync: flag + tag at the oracle's sync index of every frame the signal block accepted (other flags: L-SIG failures)
    sync = np.concatenate(ch.trace["sync"])
    offs = [o for o, _ in ch.tags["sync"]]
    assert np.array_equal(np.flatnonzero(sync), np.array(offs))
    at = {o: t for o, t in ch.tags["sync"]}
    for f in ok:
        t = at[int(f["sync_idx"])]
        assert abs(float(t["f"]["rad"]) - float(f["rad"])) <= 1e-6
        assert np.allclose(t["f"]["snr"], f["snr"], rtol=1e-4, equal_nan=True) and np.allclose(t["f"]["rssi"], f["rssi"], rtol=1e-4)
    # signal: frames back to back, nsamp + 320 items each, tag at the first one
    assert len(ch.tags["signal"]) == ok.size
    off = 0
    for k, (o, t) in enumerate(ch.tags["signal"]):
        f = ok[k]
        assert o == off and t["seq"] == k + 1 and t["nvec"] == 64
        for key in ("l_mcs", "l_len", "nsamp"):
            assert t["f"][key] == f[key], (k, key)
        assert abs(float(t["f"]["cfo_hz"]) - float(f["cfo_hz"])) <= 4.0
        off += int(f["nsamp"]) + 320
    assert sum(a.size for a in ch.trace["signal"]) == off
    if nant == 2:
        assert sum(a.size for a in ch.trace["signal1"]) == off
    # demod: one tag per frame that reaches decode, `total` soft bits each
    dem = ok[ok["status"] == 0]
    assert len(ch.tags["demod"]) == dem.size
    off = 0
    for k, (o, t) in enumerate(ch.tags["demod"]):
        f = dem[k]
        assert o == off
        for key in ("format", "mcs", "len", "cr", "ampdu", "trellis", "total"):
            assert t["f"][key] == f[key], (k, key, t["f"][key], f[key])
        off += int(f["total"])
    return off


@pytest.mark.parametrize("snr,max_call,seed", [(None, 4096, 1), (25.0, 4096, 2), (25.0, 600, 3)])
def test_chain_siso_equals_oracle(golden, snr, max_call, seed):
    pkg = load_pkg()
    g = golden["frames_siso"]
    hi = 31 if max_call > 1000 else 10
    x = _noisy(g["iq"][:g["offs"][hi]], snr, seed)
    fo, lo, po = ol.rx_item(x, max_frames=40)
    want = pkg.blocks.split_messages(bytes(po))
    assert len(want) >= hi - 1

    rx = pkg.Receiver(device=0)
    preac, preconj = rx.presiso(x)
    trig_ref = rx.trigger(preac)
    rx.close()
    ch = pkg.blocks.Chain(nant=1, seed=seed, max_call=max_call)
    try:
        msgs = ch.run(preac, preconj, x)
    finally:
        ch.close()
    assert np.array_equal(np.concatenate(ch.trace["trigger"]), trig_ref)
    nllr = _check_tags(ch, fo, 1)
    llr = np.concatenate(ch.trace["llr"])
    assert llr.size == nllr == lo.size
    err = np.abs(llr - lo) / np.maximum(1.0, np.abs(lo))
    assert err.max() <= 1e-4, float(err.max())                    # north-star LLR tolerance
    assert msgs == want                                           # published PDUs: byte for byte, in order


@pytest.mark.parametrize("snr", [None, 30.0])
def test_chain_2x2_equals_oracle(golden, snr):
    pkg = load_pkg()
    g = golden["frames_mimo"]
    a, b = _noisy(g["iq0"], snr, 13579), _noisy(g["iq1"], snr, 24680)
    fo, lo, po = ol.rx_item2(a, b, max_frames=32)
    want = pkg.blocks.split_messages(bytes(po))
    assert len(want) >= 16
    rx = pkg.Receiver(device=0)
    preac, preconj = rx.presiso(a)
    rx.close()
    ch = pkg.blocks.Chain(nant=2, seed=7, max_call=4096)
    try:
        msgs = ch.run(preac, preconj, a, b)
    finally:
        ch.close()
    nllr = _check_tags(ch, fo, 2)
    llr = np.concatenate(ch.trace["llr"])
    assert llr.size == nllr == lo.size
    err = np.abs(llr - lo) / np.maximum(1.0, np.abs(lo))
    assert err.max() <= 1e-4, float(err.max())
    assert msgs == want


def test_chain_mu_and_ndp_equals_batch(golden):
    """demod(mupos, mugid) and the NDP channel report through the blocks = the batch path (pinned to the oracle in test_gpu_mu)."""
    pkg = load_pkg()
    g = golden["frames_mu"]
    x = np.ascontiguousarray(g["iq"])
    rx = pkg.Receiver(device=0, max_frames=16, mupos=0, mugid=2, chunk_items=1)
    fr, pdu = rx.rx_batch(x, [0], [x.size])
    preac, preconj = rx.presiso(x)
    rx.close()
    want = []
    for k in range(fr.size):
        if fr[k]["status"] != 9 and fr[k]["pdu_bytes"] > 0:
            want += pkg.blocks.split_messages(bytes(pdu[k, :fr[k]["pdu_bytes"]]))
    assert any(m[0] == 20 and len(m) == 1027 for m in want) and any(m[0] == 2 for m in want)
    ch = pkg.blocks.Chain(nant=1, mupos=0, mugid=2, seed=11, max_call=2048)
    try:
        msgs = ch.run(preac, preconj, x)
    finally:
        ch.close()
    assert len(msgs) == len(want)
    for m, w in zip(msgs, want):
        if w[0] == 20:
            assert m[:3] == w[:3] and len(m) == 1027
            u, v = np.frombuffer(m[3:], np.float32), np.frombuffer(w[3:], np.float32)
            assert np.max(np.abs(u - v)) <= 1e-5 * max(1.0, float(np.max(np.abs(v))))
        else:
            assert m == w
    ndp = [t for _, t in ch.tags["demod"] if t["nvec"] == 128]
    assert len(ndp) >= 1 and all(t["f"]["total"] == 1024 and t["f"]["trellis"] == 0 for t in ndp)


def test_blk_argument_errors_and_ports():
    pkg = load_pkg()
    L = pkg._cabi.lib()
    nin, nout = C.c_int(0), C.c_int(0)
    ib, ob = (C.c_int * 3)(), (C.c_int * 2)()
    assert L.c8b_blk_ports(pkg.blocks.SYNC, C.byref(nin), C.byref(nout), ib, ob) == 0
    assert (nin.value, nout.value, list(ib), list(ob)) == (3, 1, [1, 8, 8], [1, 0])
    assert L.c8b_blk_ports(7, None, None, None, None) == -3
    assert L.c8b_blk_forecast(pkg.blocks.DECODE, 100) == 260 and L.c8b_blk_forecast(pkg.blocks.DEMOD, 100) == 100
    h = C.c_void_p()
    assert L.c8b_blk_create(None, 9, C.byref(h)) == -3
    assert L.c8b_blk_create(None, pkg.blocks.TRIGGER, C.byref(h)) == 0
    c, p = C.c_int(0), C.c_int(0)
    assert L.c8b_blk_work(h, 4, None, None, None, None, 0, C.byref(c), C.byref(p), None, 0, None, None, 0, None) == -3
    # a trigger call with nothing to do is not an error
    x = np.zeros(4, np.float32)
    o = np.zeros(4, np.uint8)
    ni = (C.c_int * 1)(0)
    ip, op = (C.c_void_p * 1)(x.ctypes.data), (C.c_void_p * 1)(o.ctypes.data)
    assert L.c8b_blk_work(h, 4, ni, ip, op, None, 0, C.byref(c), C.byref(p), None, 0, None, None, 0, None) == 0
    assert (c.value, p.value) == (0, 0)
    L.c8b_blk_destroy(h)
