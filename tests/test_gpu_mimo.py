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
