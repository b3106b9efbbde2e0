"""2x2 SU-MIMO path on the GPU (signal2 + demod2: k_header2, k_demod2 through c8b_demod2 / c8b_rx_batch2)
vs the oracle's restatement of lib/signal2_impl.cc + lib/demod2_impl.cc, on the reference generator's frames
(HT MCS8-15, VHT 2SS MCS0-8; BASELINE config 4 family)."""
import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu
HDR = ("status", "format", "mcs", "len", "cr", "ampdu", "nss", "nsym", "nsymsamp", "ncbps", "ndbps", "trellis", "total", "data_off")


@pytest.fixture(scope="module")
def rx():
    r = load_pkg().Receiver(device=0)
    yield r
    r.close()


def _ants(g, snr):
    a, b = g["iq0"].copy(), g["iq1"].copy()
    if snr is not None:
        r0, r1 = np.random.default_rng(13579), np.random.default_rng(24680)       # tools/performance/gr_sumimo.py:62-63
        s = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
        a = (a + s * (r0.standard_normal(a.size) + 1j * r0.standard_normal(a.size))).astype(np.complex64)
        b = (b + s * (r1.standard_normal(b.size) + 1j * r1.standard_normal(b.size))).astype(np.complex64)
    return a, b


@pytest.mark.parametrize("snr", [None, 30.0])
def test_demod2_llrs_match_oracle(rx, golden, snr):
    g = golden["frames_mimo"]
    a, b = _ants(g, snr)
    offs = g["offs"]
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    fr, chan = rx.detect(a, off, ln)
    stride = 64 * 832
    fr2, llr = rx.demod2(a, b, off, ln, fr, chan, stride)
    worst = 0.0
    for i in range(len(off)):
        fo, lo, _ = ol.rx_item2(a[offs[i]:offs[i + 1]], b[offs[i]:offs[i + 1]], max_frames=1)
        for k in HDR:
            assert fr2[i][k] == fo[0][k], (i, k, fr2[i][k], fo[0][k])
        n = int(fo[0]["total"])
        err = np.abs(llr[i, :n] - lo[:n]) / np.maximum(1.0, np.abs(lo[:n]))
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-4, (i, int(np.argmax(err)), float(err.max()))
    print("worst relative LLR error (2x2) %.3g" % worst)


@pytest.mark.parametrize("snr", [None, 30.0, 20.0])
def test_rx_batch2_pdus_match_oracle(golden, snr):
    pkg = load_pkg()
    g = golden["frames_mimo"]
    a, b = _ants(g, snr)
    offs = g["offs"]
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    rx = pkg.Receiver(device=0, chunk_items=5)
    fr, pdu = rx.rx_batch2(a, b, off, ln)
    rx.close()
    el = g["exp_len"]
    eo = np.cumsum(np.r_[0, el])
    for i in range(len(off)):
        fo, _, po = ol.rx_item2(a[offs[i]:offs[i + 1]], b[offs[i]:offs[i + 1]], max_frames=1)
        assert fr[i]["status"] == fo[0]["status"] and fr[i]["npdu"] == fo[0]["npdu"] and fr[i]["pdu_bytes"] == po.size, (i, fr[i]["status"], fo[0]["status"])
        assert bytes(pdu[i, :po.size]) == bytes(po), i
        if snr is None or snr >= 30:
            assert fr[i]["npdu"] == 1 and bytes(pdu[i, 3:3 + el[i]]) == bytes(g["exp_mpdu"][eo[i]:eo[i + 1]]), i


def test_siso_frames_through_the_2x2_block(golden):
    """1-stream frames (legacy / HT / VHT) fed to the 2-antenna block: demod2 uses antenna 0 only (lib/demod2_impl.cc:473-497)"""
    pkg = load_pkg()
    g = golden["frames_siso"]
    offs = g["offs"]
    a = g["iq"]
    rng = np.random.default_rng(3)
    b = (0.01 * (rng.standard_normal(a.size) + 1j * rng.standard_normal(a.size))).astype(np.complex64)
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    rx = pkg.Receiver(device=0)
    fr, pdu = rx.rx_batch2(a, b, off, ln)
    rx.close()
    for i in range(len(off)):
        fo, _, po = ol.rx_item2(a[offs[i]:offs[i + 1]], b[offs[i]:offs[i + 1]], max_frames=1)
        assert fr[i]["status"] == fo[0]["status"] == 0 and fr[i]["nss"] == 1
        assert fr[i]["pdu_bytes"] == po.size and bytes(pdu[i, :po.size]) == bytes(po), i


@pytest.mark.parametrize("snr", [None, 25.0])
def test_header2_warp_kernel_equals_the_thread_kernel(golden, snr):
    """k_header2_w (one warp per frame, frontend_mode 0) against k_header2 (one thread per frame, frontend_mode 1): same
    header fields, same drop codes, soft bits within the LLR gate, PDUs byte for byte -- 2-stream and 1-stream frames"""
    pkg = load_pkg()
    res = []
    for mode in (0, 1):
        rx = pkg.Receiver(device=0, frontend_mode=mode)
        out = []
        for g, two in ((golden["frames_mimo"], True), (golden["frames_siso"], False)):
            if two:
                a, b = _ants(g, snr)
            else:
                a = g["iq"].copy()
                b = (0.01 * np.random.default_rng(3).standard_normal(a.size)).astype(np.complex64)
            offs = g["offs"]
            off, ln = offs[:-1], np.diff(offs).astype(np.int32)
            fr, chan = rx.detect(a, off, ln)
            fr2, llr = rx.demod2(a, b, off, ln, fr, chan, 64 * 832)
            frb, pdu = rx.rx_batch2(a, b, off, ln)
            out.append((fr2, llr, frb, pdu))
        rx.close()
        res.append(out)
    for (f0, l0, b0, p0), (f1, l1, b1, p1) in zip(*res):
        for i in range(f0.size):
            for k in HDR:
                assert f0[i][k] == f1[i][k], (i, k, f0[i][k], f1[i][k])
            n = int(f1[i]["total"]) if f1[i]["status"] == 0 else 0
            err = np.abs(l0[i, :n] - l1[i, :n]) / np.maximum(1.0, np.abs(l1[i, :n]))
            assert n == 0 or err.max() <= 1e-4, (i, float(err.max()))
            for key in ("sssnr0", "sssnr1"):                     # (the SIG-B SNR of a noiseless frame is rounding noise, ~140 dB)
                if abs(float(f1[i][key])) < 60.0:
                    assert abs(float(f0[i][key]) - float(f1[i][key])) <= 1e-2 * max(1.0, abs(float(f1[i][key]))), (i, key)
            assert b0[i]["status"] == b1[i]["status"] and b0[i]["pdu_bytes"] == b1[i]["pdu_bytes"]
            assert bytes(p0[i, :b0[i]["pdu_bytes"]]) == bytes(p1[i, :b1[i]["pdu_bytes"]])


def test_mmse_equaliser_option(golden):
    """cfg.mmse (north_star "LS/MMSE"): (H^H H + sigma^2 I)^-1 H^H instead of the reference's zero forcing
    (lib/demod2_impl.cc:410-429).  Off by default (every parity test above runs the reference's form); on a correlated 2x2
    channel at 15 dB, where zero forcing amplifies the noise, MMSE must decode at least as many frames; on a noiseless
    capture (sync's snr tag is NaN) it IS zero forcing."""
    pkg = load_pkg()
    g = golden["frames_mimo"]
    offs = g["offs"]
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    x0, x1 = g["iq0"], g["iq1"]
    y0 = (x0 + 0.85 * x1).astype(np.complex64)                   # H = [[1, 0.85], [0.85, 1]]: condition number 12
    y1 = (0.85 * x0 + x1).astype(np.complex64)
    ok = {}
    for mmse in (0, 1):
        rx = pkg.Receiver(device=0, mmse=mmse)
        fr, pdu = rx.rx_batch2(y0, y1, off, ln)
        clean = [bytes(pdu[i, :fr[i]["pdu_bytes"]]) for i in range(off.size)]
        n = 0
        for trial in range(6):
            r0, r1 = np.random.default_rng(100 + trial), np.random.default_rng(200 + trial)
            s = 0.1875 * 1.3 / np.sqrt(2 * 10 ** 1.5)
            a = (y0 + s * (r0.standard_normal(y0.size) + 1j * r0.standard_normal(y0.size))).astype(np.complex64)
            b = (y1 + s * (r1.standard_normal(y1.size) + 1j * r1.standard_normal(y1.size))).astype(np.complex64)
            fr, _ = rx.rx_batch2(a, b, off, ln)
            n += int((fr["npdu"] >= 1).sum())
        rx.close()
        ok[mmse] = (n, clean)
    assert ok[0][1] == ok[1][1] and all(len(c) > 0 for c in ok[0][1])      # noiseless: identical records, every frame decodes
    print("frames decoded at 15 dB on the correlated channel: ZF %d, MMSE %d of %d" % (ok[0][0], ok[1][0], 6 * off.size))
    assert ok[1][0] >= ok[0][0] and ok[1][0] > 0
