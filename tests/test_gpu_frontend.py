"""GPU front end (k_presiso / k_trigger / k_detect through the C ABI) vs the oracle."""
import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[0, 1], ids=["warp_frontend", "thread_frontend"])
def rx(request):
    """every test runs against both front ends: warp-cooperative (k_detect_w / k_header_w) and one thread per item"""
    r = load_pkg().Receiver(device=0, frontend_mode=request.param)
    yield r
    r.close()


def _noisy(g, snr_db, seed=7):
    iq = g["iq"].copy()
    if snr_db is not None:
        rng = np.random.default_rng(seed)
        s = 0.1875 / np.sqrt(2 * 10 ** (snr_db / 10))
        iq = (iq + s * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))).astype(np.complex64)
    return iq


@pytest.mark.parametrize("snr", [None, 30.0])
def test_presiso_bit_exact(rx, golden, snr):
    """same summation tree as the oracle -> identical bits, tile boundaries included (item of 37k samples)"""
    g = golden["frames_siso"]
    x = np.ascontiguousarray(_noisy(g, snr)[: g["offs"][3]])
    O = ol.oracle()
    n = x.size
    pa, pc = np.zeros(n, np.float32), np.zeros(2 * n, np.float32)
    O.orx_presiso(ol.c2f(x), n, pa, pc)
    ga, gc = rx.presiso(x)
    assert np.array_equal(ga.view(np.uint32), pa.view(np.uint32)) or np.array_equal(ga, pa, equal_nan=True)
    assert np.array_equal(gc.view(np.float32), pc, equal_nan=True)
    # trigger FSM on the GPU
    tg = rx.trigger(pa)
    to = np.zeros(n, np.uint8)
    O.orx_trigger(pa, n, to, np.zeros(5, np.int32))
    assert np.array_equal(tg, to) and (to & 1).sum() >= 3


def test_presiso_short_and_odd_lengths(rx):
    rng = np.random.default_rng(3)
    O = ol.oracle()
    for n in (1, 15, 16, 17, 63, 64, 65, 1023, 1024, 1025, 2049):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        pa, pc = np.zeros(n, np.float32), np.zeros(2 * n, np.float32)
        O.orx_presiso(ol.c2f(x), n, pa, pc)
        ga, gc = rx.presiso(x)
        assert np.array_equal(ga, pa, equal_nan=True), n
        assert np.array_equal(gc.view(np.float32), pc, equal_nan=True), n


@pytest.mark.parametrize("snr", [None, 30.0, 10.0])
def test_detect_matches_oracle(rx, golden, snr):
    g = golden["frames_siso"]
    iq = _noisy(g, snr)
    offs = g["offs"]
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    fr, chan = rx.detect(iq, off, ln)
    O = ol.oracle()
    for i in range(len(off)):
        fo, _, _ = ol.rx_item(iq[offs[i]:offs[i + 1]], max_frames=1)
        if fo[0]["nsamp"] == 0:
            assert fr[i]["status"] == fo[0]["status"]
            continue
        for k in ("trig_idx", "sync_idx", "l_mcs", "l_len", "nsamp"):
            assert fr[i][k] == fo[0][k], (i, k, fr[i][k], fo[0][k])
        assert abs(float(fr[i]["rad"]) - float(fo[0]["rad"])) <= 1e-6            # SURVEY 8d gate 3
        assert np.allclose(fr[i]["snr"], fo[0]["snr"], rtol=1e-4, equal_nan=True)
        assert np.allclose(fr[i]["rssi"], fo[0]["rssi"], rtol=1e-5)
        # legacy channel = tag "chan" of lib/signal_impl.cc:146-152
        h, llr48, bits = np.zeros(128, np.float32), np.zeros(48, np.float32), np.zeros(24, np.uint8)
        import ctypes as C
        m, l, ns = C.c_int(), C.c_int(), C.c_int()
        x = np.ascontiguousarray(iq[offs[i] + fo[0]["sync_idx"]: offs[i + 1]])
        O.orx_signal(ol.c2f(x), float(fo[0]["rad"]), h, llr48, bits, C.byref(m), C.byref(l), C.byref(ns))
        hh = h.view(np.complex64)
        assert np.max(np.abs(chan[i] - hh)) <= 1e-5 * np.max(np.abs(hh))


def test_detect_ragged_items(rx, golden):
    """empty, tiny, truncated and noise-only items in one batch"""
    g = golden["frames_siso"]
    x = g["iq"][g["offs"][0]:g["offs"][1]]
    rng = np.random.default_rng(9)
    parts = [x[:0], x[:64], x[:1450], x[:1700], x[:2500], x, (0.01 * (rng.standard_normal(5000) + 1j * rng.standard_normal(5000))).astype(np.complex64)]
    iq = np.concatenate(parts).astype(np.complex64)
    ln = np.array([p.size for p in parts], np.int32)
    off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64)
    fr, _ = rx.detect(iq, off, ln)
    for i, p in enumerate(parts):
        fo, _, _ = ol.rx_item(p if p.size else np.zeros(0, np.complex64), max_frames=1) if p.size else (None, None, None)
        want = fo[0]["status"] if fo is not None else 1
        assert fr[i]["status"] == want, (i, fr[i]["status"], want)
