"""The oracle's whole chain on waveforms made by the reference's own generator (tools/phy80211.py):
every frame must come back as its own MPDU with CRC-32 pass -- the end-to-end known-answer test the
reference offers for this path (recipe tools/pktGenExample.py:173-199)."""
import numpy as np
import pytest

import oracle_lib as ol


def _items(g):
    iq, offs = g["iq"], g["offs"]
    el = g["exp_len"]
    eo = np.cumsum(np.r_[0, el])
    for i in range(len(offs) - 1):
        yield i, iq[offs[i]:offs[i + 1]], bytes(g["exp_mpdu"][eo[i]:eo[i + 1]]), g["meta"][i]


def test_config1_legacy_mcs0_plumbing(golden):
    i, x, mpdu, meta = next(_items(golden["frames_siso"]))
    fr, llr, pdu = ol.rx_item(x)
    f = fr[0]
    assert f["status"] == 0
    assert (f["format"], f["mcs"], f["len"], f["nsym"], f["nsamp"], f["trellis"], f["total"]) == (0, 0, 94, 33, 2640, 774, 1584)
    recs = ol.split_pdus(pdu)
    assert len(recs) == 1
    assert recs[0] == bytes([0, 94, 0]) + mpdu + bytes([0])      # [fmt][len lo][len hi][MPDU][mcs]
    assert mpdu[:10].hex() == "08016e00f469d5800fa0"               # SURVEY 8c known answer


def test_all_formats_clean(golden):
    for i, x, mpdu, meta in _items(golden["frames_siso"]):
        fr, llr, pdu = ol.rx_item(x)
        f = fr[0]
        assert f["status"] == 0, (i, meta, f["status"])
        assert (f["format"], f["mcs"]) == (int(meta[0]), int(meta[1]))
        recs = ol.split_pdus(pdu)
        assert len(recs) >= 1 and recs[0][3:-1] == mpdu and recs[0][0] == int(meta[0]) and recs[0][-1] == int(meta[1]), (i, meta)
        # CFO estimate: rad = -2 pi f / 20e6 (compensation sign, lib/sync_impl.cc:181-196)
        assert abs(f["rad"] + 2 * np.pi * meta[2] / 20e6) < 2e-4


def test_awgn_30db(golden):
    rng = np.random.default_rng(13579)
    sigma = np.sqrt(0.1875 ** 2 / 10 ** 3.0 / 2)                   # tools/performance/perf_siso.py:92 at 30 dB
    ok = 0
    items = list(_items(golden["frames_siso"]))
    for i, x, mpdu, meta in items:
        y = (x + sigma * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
        fr, llr, pdu = ol.rx_item(y)
        recs = ol.split_pdus(pdu)
        ok += int(len(recs) >= 1 and recs[0][3:-1] == mpdu)
    assert ok == len(items)


def test_two_subframe_ampdu_quirk(golden):
    """lib/decode_impl.cc:336,346-351,415: tmpLen is OR-accumulated across subframes, so only the first
    subframe of an A-MPDU is delivered.  The oracle reproduces that."""
    g = golden["frames_siso"]
    x = g["iq"][g["offs"][-2]:g["offs"][-1]]
    fr, llr, pdu = ol.rx_item(x)
    assert fr[0]["npdu"] == 1


def test_noise_only_and_empty():
    rng = np.random.default_rng(1)
    x = (0.01 * (rng.standard_normal(5000) + 1j * rng.standard_normal(5000))).astype(np.complex64)
    fr, llr, pdu = ol.rx_item(x)
    assert fr[0]["status"] == 1 and pdu.size == 0
    fr, llr, pdu = ol.rx_item(np.zeros(64, np.complex64))
    assert fr[0]["status"] == 1


def test_truncated_frame(golden):
    i, x, mpdu, meta = next(_items(golden["frames_siso"]))
    fr, llr, pdu = ol.rx_item(x[:2500])
    assert fr[0]["status"] == 4 and pdu.size == 0


def test_two_frames_in_one_item(golden):
    items = list(_items(golden["frames_siso"]))
    x = np.concatenate([items[3][1], items[20][1]])
    fr, llr, pdu = ol.rx_item(x, max_frames=4)
    recs = ol.split_pdus(pdu)
    assert len(fr) == 2 and len(recs) == 2
    assert recs[0][3:-1] == items[3][2] and recs[1][3:-1] == items[20][2]


def test_batch_matches_item_loop(golden):
    import ctypes as C
    g = golden["frames_siso"]
    offs = g["offs"]
    n = len(offs) - 1
    lens = np.diff(offs).astype(np.int32)
    frames = np.zeros(n, ol.FRAME_DTYPE)
    stride = 4400
    pdu = np.zeros(n * stride, np.uint8)
    ol.oracle().orx_rx_batch(ol.c2f(g["iq"]), np.ascontiguousarray(offs[:-1]), lens, n, 4, frames.ctypes.data, pdu, stride)
    for i, x, mpdu, meta in _items(g):
        fr, llr, p1 = ol.rx_item(x, max_frames=1)
        assert frames[i]["pdu_bytes"] == p1.size
        assert bytes(pdu[i * stride: i * stride + p1.size]) == bytes(p1)


# ---------------------------------------------------------------- 2x2 (signal2 + demod2) ----------
def _mimo_items(g):
    offs, el = g["offs"], g["exp_len"]
    eo = np.cumsum(np.r_[0, el])
    for i in range(len(offs) - 1):
        yield i, g["iq0"][offs[i]:offs[i + 1]], g["iq1"][offs[i]:offs[i + 1]], bytes(g["exp_mpdu"][eo[i]:eo[i + 1]]), g["meta"][i]


@pytest.mark.parametrize("snr", [None, 30.0])
def test_mimo_2x2_all_mcs(golden, snr):
    """HT MCS8-15 and VHT 2SS MCS0-8 from the reference generator (tools/pktGenExample.py:206-217), identity channel
    as tools/performance/gr_sumimo.py:70-78, independent noise per antenna (seeds 13579 / 24680)"""
    r0, r1 = np.random.default_rng(13579), np.random.default_rng(24680)
    for i, a, b, mpdu, meta in _mimo_items(golden["frames_mimo"]):
        if snr is not None:
            s = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
            a = (a + s * (r0.standard_normal(a.size) + 1j * r0.standard_normal(a.size))).astype(np.complex64)
            b = (b + s * (r1.standard_normal(b.size) + 1j * r1.standard_normal(b.size))).astype(np.complex64)
        fr, llr, pdu = ol.rx_item2(a, b)
        f = fr[0]
        assert f["status"] == 0, (i, meta, f["status"])
        assert (f["format"], f["mcs"], f["nss"]) == (int(meta[0]), int(meta[1]), 2), (i, f["format"], f["mcs"], f["nss"])
        recs = ol.split_pdus(pdu)
        assert len(recs) >= 1 and recs[0][3:-1] == mpdu and recs[0][-1] == int(meta[1]), (i, meta)
        assert abs(f["rad"] + 2 * np.pi * meta[2] / 20e6) < 2e-4


def test_mimo_tables_vs_reference(golden):
    if not ol.have_ref():
        pytest.skip("oracle/_ref not built")
    R = ol.ref()
    f = np.zeros(128, np.float32)
    assert R.ref_table_f(b"LTF_NL_28_F_FLOAT2", f) == 64
    l = np.zeros(64, np.float32)
    ol.oracle().orx_ltf(1, l)
    assert np.array_equal(f[:64], l * 0.5)
    for name, want in ((b"PILOT_HT_2_1", [1, 1, -1, -1]), (b"PILOT_HT_2_2", [1, -1, -1, 1]), (b"PILOT_VHT", [1, 1, 1, -1]), (b"PILOT_HT_1", [1, 1, 1, -1])):
        assert R.ref_table_f(name, f) == 4 and list(f[:4]) == want


def _mu_items(g):
    offs = g["offs"]
    el = g["exp_len"]
    eo = np.cumsum(np.r_[0, el])
    for i in range(len(offs) - 1):
        kind, mcs, par = g["meta"][i]
        yield i, int(kind), int(mcs), par, np.ascontiguousarray(g["iq"][offs[i]:offs[i + 1]]), bytes(g["exp_mpdu"][eo[i]:eo[i + 1]])


def test_mu_mimo_user_positions(golden):
    """2-user VHT MU-MIMO frames of the reference generator (tools/phy80211.py genAmpduMu, zero-forcing precoder of a flat
    2x2 channel): each station decodes its own A-MPDU with demod(mupos, mugid = 2) (lib/demod_impl.cc:347-380)"""
    O = ol.oracle()
    try:
        n = 0
        for i, kind, mcs, par, x, mpdu in _mu_items(golden["frames_mu"]):
            if kind != 4:
                continue
            O.orx_set_mupos(int(par))
            fo, _, po = ol.rx_item(x, max_frames=2)
            assert len(fo) == 1 and fo[0]["status"] == 0 and (fo[0]["format"], fo[0]["mcs"], fo[0]["nss"]) == (2, mcs, 1)
            recs = ol.split_pdus(po)
            assert len(recs) == 1 and recs[0][3:-1] == mpdu and recs[0][0] == 2 and recs[0][-1] == mcs
            # the wrong user position estimates the other user's channel: the frame must not pass
            O.orx_set_mupos(1 - int(par))
            fw, _, pw = ol.rx_item(x, max_frames=2)
            assert pw.size == 0 or ol.split_pdus(pw)[0][3:-1] != mpdu
            n += 1
        assert n == 4
    finally:
        O.orx_set_mupos(0)


def test_ndp_channel_report(golden):
    """VHT NDP (empty A-MPDU, 2 transmit streams, one receive antenna): demod tags the two VHT-LTFs (mu2x1chan,
    lib/demod_impl.cc:238-249), decode publishes [20][len lo][len hi][128 x (re, im) float32] (lib/decode_impl.cc:100-121).
    With P = [[1, -1], [1, 1]] the LTF spectra separate the two transmit streams: |F1 - F2| / |F1 + F2| = |h0| / |h1|."""
    g = golden["frames_mu"]
    H = g["chan"]
    n = 0
    for i, kind, mcs, par, x, _ in _mu_items(g):
        if kind != 3:
            continue
        fo, _, po = ol.rx_item(x, max_frames=2)
        f = fo[0]
        assert len(fo) == 1 and f["status"] == 7 and (f["format"], f["nss"], f["nsym"], f["len"], f["total"]) == (2, 2, 0, 0, 1024)
        assert (f["npdu"], f["pdu_bytes"]) == (1, 1027) and bytes(po[:3]) == bytes([20, 0, 4])
        assert abs(f["cfo_hz"] + par) < 300.0                          # the receiver's correction = minus the applied CFO
        c = np.frombuffer(bytes(po[3:]), np.float32).view(np.complex64)
        F1, F2 = np.fft.fft(c[:64]), np.fft.fft(c[64:])
        used = [k for k in range(64) if not (k == 0 or 29 <= k <= 35 or k in (7, 21, 43, 57))]      # pilots carry no P matrix
        r = np.abs(F1 - F2)[used] / np.abs(F1 + F2)[used]
        assert np.allclose(r, abs(H[0, 0]) / abs(H[0, 1]), rtol=2e-3), (r.min(), r.max())
        n += 1
    assert n == 2
