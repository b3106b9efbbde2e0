"""Host logic: the product's per-frame routines (phy_serial.cuh, run one-thread-per-frame by k_detect /
k_header on the GPU), compiled for the host, against the oracle on the reference generator's frames."""
import ctypes as C

import numpy as np
import pytest

import hostsim_lib as hs
import oracle_lib as ol
from __graft_entry__ import load_pkg

DET = ("status", "trig_idx", "sync_idx", "l_mcs", "l_len", "nsamp")
HDR = ("status", "format", "mcs", "len", "cr", "ampdu", "nss", "nsym", "nsymsamp", "ncbps", "ndbps", "trellis", "total", "data_off")


def _items(g, noise=0.0, seed=3):
    iq, offs = g["iq"], g["offs"]
    rng = np.random.default_rng(seed)
    for i in range(len(offs) - 1):
        x = iq[offs[i]:offs[i + 1]]
        if noise:
            x = (x + noise * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
        yield i, np.ascontiguousarray(x)


def _detect(x, maxf=1, bitmap=False):
    pkg = load_pkg()
    H, O = hs.lib(), ol.oracle()
    xf = ol.c2f(x)
    n = x.size
    preac, preconj = np.zeros(n, np.float32), np.zeros(2 * n, np.float32)
    O.orx_presiso(xf, n, preac, preconj)
    f = np.zeros(maxf, pkg.FRAME_DTYPE)
    chan = np.zeros(128 * maxf, np.float32)
    H.hs_detect(xf, preac, n, -1 if bitmap else 0, maxf, f.ctypes.data, chan)
    return f, chan, preac, preconj.view(np.complex64)


@pytest.mark.parametrize("noise", [0.0, 0.1875 / np.sqrt(2 * 10 ** 3.0), 0.1875 / np.sqrt(2 * 10 ** 1.0)])
def test_detect_and_header_match_oracle(golden, noise):
    H = hs.lib()
    for i, x in _items(golden["frames_siso"], noise):
        fo, llr, pdu = ol.rx_item(x, max_frames=1)
        f, chan, _, _ = _detect(x)
        want_det = fo[0]["status"] if fo[0]["status"] in (1, 2, 3, 4) and fo[0]["nsamp"] == 0 else 0
        if fo[0]["nsamp"] == 0:                        # no frame record: only the drop code is comparable
            assert f[0]["status"] == fo[0]["status"], (i, f[0]["status"], fo[0]["status"])
            continue
        for k in DET[1:]:
            assert f[0][k] == fo[0][k], (i, k, f[0][k], fo[0][k])
        # float tags: the routines evaluate cos/sin/atan2 in double where the reference calls cosf/sinf/atan2f
        assert abs(float(f[0]["rad"]) - float(fo[0]["rad"])) <= 1e-8, (i, f[0]["rad"], fo[0]["rad"])
        assert abs(float(f[0]["cfo_hz"]) - float(fo[0]["cfo_hz"])) <= 0.05
        for k in ("snr", "rssi"):
            assert np.allclose(f[0][k], fo[0][k], rtol=1e-5, atol=0, equal_nan=True), (i, k, f[0][k], fo[0][k])
        if f[0]["status"] != 0:
            assert fo[0]["status"] == 4
            continue
        hinv = np.zeros(128, np.float32)
        H.hs_header(ol.c2f(x), f.ctypes.data, chan, 0, hinv)
        for k in HDR:
            assert f[0][k] == fo[0][k], (i, k, f[0][k], fo[0][k])
        # per-stream SNR of a noiseless frame is pure rounding noise (~120 dB): compare loosely
        assert abs(float(f[0]["sssnr0"]) - float(fo[0]["sssnr0"])) <= (0.5 if fo[0]["sssnr0"] > 60 else 0.01)


def test_truncated_noise_and_empty_items(golden):
    g = golden["frames_siso"]
    x = np.ascontiguousarray(g["iq"][g["offs"][0]:g["offs"][1]])
    for cut in (64, 200, 1000, 1450, 1500, 1700, 2500, x.size - 1):
        fo, _, _ = ol.rx_item(x[:cut], max_frames=1)
        f, _, _, _ = _detect(np.ascontiguousarray(x[:cut]))
        assert f[0]["status"] == fo[0]["status"], (cut, f[0]["status"], fo[0]["status"])
    rng = np.random.default_rng(5)
    for amp in (0.01, 0.3):
        z = (amp * (rng.standard_normal(6000) + 1j * rng.standard_normal(6000))).astype(np.complex64)
        fo, _, _ = ol.rx_item(z, max_frames=1)
        f, _, _, _ = _detect(z)
        assert f[0]["status"] == fo[0]["status"]
    # periodic-16 junk: triggers fire, LTF autocorrelation may or may not pass, L-SIG fails
    t = np.tile((rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex64), 60)
    z = np.concatenate([np.zeros(300, np.complex64), 0.2 * t, np.zeros(900, np.complex64)]).astype(np.complex64)
    z += (0.002 * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))).astype(np.complex64)
    fo, _, _ = ol.rx_item(z, max_frames=1)
    f, _, _, _ = _detect(z)
    assert f[0]["status"] == fo[0]["status"]


def test_latch_value_equals_presiso_output(golden):
    H = hs.lib()
    g = golden["frames_siso"]
    x = np.ascontiguousarray(g["iq"][g["offs"][5]:g["offs"][6]])
    _, _, preac, preconj = _detect(x)
    out = np.zeros(2, np.float32)
    for i in (0, 3, 15, 16, 17, 31, 47, 48, 63, 64, 300, 450, 500, 1234):
        H.hs_conj_at(ol.c2f(x), i, out)
        assert out[0] == preconj[i].real and out[1] == preconj[i].imag, i


def test_trigger_sigviterbi_fft_vs_oracle():
    H, O = hs.lib(), ol.oracle()
    rng = np.random.default_rng(11)
    ac = np.clip(rng.normal(0.3, 0.25, 20000), 0, 1).astype(np.float32)
    ac[5000:5300] = 0.7
    ac[9000:9030] = 0.5
    a, b = np.zeros(ac.size, np.uint8), np.zeros(ac.size, np.uint8)
    H.hs_trigger(ac, ac.size, a)
    O.orx_trigger(ac, ac.size, b, np.zeros(5, np.int32))
    assert np.array_equal(a, b) and (a & 1).sum() > 0
    for T in (24, 26, 48):
        for trial in range(100):
            llr = rng.normal(0, 1, 2 * T).astype(np.float32)
            if trial % 2:
                llr = (np.round(llr * 2) / 2).astype(np.float32)
            x, y = np.zeros(T, np.uint8), np.zeros(T, np.uint8)
            H.hs_sig_viterbi(llr, x, T)
            O.orx_sig_viterbi(llr, y, T)
            assert np.array_equal(x, y)
    for _ in range(20):
        z = ol.c2f((rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64))
        p, q = np.zeros(128, np.float32), np.zeros(128, np.float32)
        H.hs_fft64(z, p)
        O.orx_fft64(z, q)
        assert np.array_equal(p, q)
    for _ in range(300):
        bits = rng.integers(0, 2, 34).astype(np.uint8)
        crc = rng.integers(0, 2, 8).astype(np.uint8)
        assert H.hs_crc8(bits, 34, crc) == O.orx_crc8_check(bits, 34, crc)


@pytest.mark.parametrize("snr", [None, 30.0])
def test_header2_matches_oracle(golden, snr):
    """2-antenna header states (demod_header2) vs the oracle's demod2 on the reference generator's 2x2 frames"""
    pkg = load_pkg()
    H = hs.lib()
    g = golden["frames_mimo"]
    offs = g["offs"]
    r0, r1 = np.random.default_rng(13579), np.random.default_rng(24680)
    for i in range(len(offs) - 1):
        a, b = g["iq0"][offs[i]:offs[i + 1]], g["iq1"][offs[i]:offs[i + 1]]
        if snr is not None:
            s = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
            a = (a + s * (r0.standard_normal(a.size) + 1j * r0.standard_normal(a.size))).astype(np.complex64)
            b = (b + s * (r1.standard_normal(b.size) + 1j * r1.standard_normal(b.size))).astype(np.complex64)
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        fo, llr, pdu = ol.rx_item2(a, b, max_frames=1)
        f, chan, _, _ = _detect(a)
        for k in DET[1:]:
            assert f[0][k] == fo[0][k], (i, k, f[0][k], fo[0][k])
        hinv, w2 = np.zeros(128, np.float32), np.zeros(2 * 264, np.float32)
        H.hs_header2(ol.c2f(a), ol.c2f(b), f.ctypes.data, chan, hinv, w2)
        for k in HDR:
            assert f[0][k] == fo[0][k], (i, k, f[0][k], fo[0][k])
        assert f[0]["nss"] == 2
        if fo[0]["format"] == 2:
            for k in ("sssnr0", "sssnr1"):
                assert abs(float(f[0][k]) - float(fo[0][k])) <= (0.5 if fo[0][k] > 60 else 0.01), (i, k, f[0][k], fo[0][k])
        # identity channel: the folded zero-forcing matrix is ~ diag(1/h) per antenna
        w = w2.view(np.complex64)[:256].reshape(64, 4)
        used = [k for k in range(64) if not (k == 0 or 29 <= k <= 35)]
        assert np.all(np.abs(w[used, 1]) < 0.2 * np.abs(w[used, 0])) and np.all(np.abs(w[used, 2]) < 0.2 * np.abs(w[used, 3]))


def test_multi_frame_capture_matches_oracle(golden):
    """a capture holding many frames back to back (the reference demo: tools/pktGenExample.py writes 25 frames into one
    .bin): frames must be found in stream order with the flags inside a copied frame swallowed"""
    g = golden["frames_siso"]
    offs = g["offs"]
    x = np.ascontiguousarray(g["iq"][offs[1]:offs[12]])          # 11 frames, 400-sample gaps
    fo, _, _ = ol.rx_item(x, max_frames=16)
    f, chan, _, _ = _detect(x, maxf=16)
    assert len(fo) == 11
    for k in range(16):
        if k < len(fo):
            for key in DET:
                assert f[k][key] == fo[k][key], (k, key, f[k][key], fo[k][key])
        else:
            assert f[k]["status"] == 9
    # fewer records than frames: the first ones, in order
    f4, _, _, _ = _detect(x, maxf=4)
    for k in range(4):
        assert f4[k]["sync_idx"] == fo[k]["sync_idx"]


def test_bitmap_scan_is_exact(golden):
    """skipping 32 idle samples at a time (threshold bitmap from k_presiso) must not change a single field"""
    g = golden["frames_siso"]
    offs = g["offs"]
    rng = np.random.default_rng(8)
    caps = [np.ascontiguousarray(g["iq"][offs[1]:offs[12]]), np.ascontiguousarray(g["iq"][offs[20]:offs[21]][:1777])]
    z = np.ascontiguousarray(g["iq"][offs[5]:offs[9]])
    caps.append((z + 0.03 * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))).astype(np.complex64))
    t = np.tile((rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex64), 40)
    caps.append(np.concatenate([np.zeros(333, np.complex64), 0.2 * t, np.zeros(1000, np.complex64), 0.1 * t, np.zeros(500, np.complex64)]).astype(np.complex64))
    for x in caps:
        a, ca, _, _ = _detect(x, maxf=16)
        b, cb, _, _ = _detect(x, maxf=16, bitmap=True)
        assert a.tobytes() == b.tobytes() and np.array_equal(ca, cb)


def _stream_detect(x, pushes, maxf=8, cap=1 << 20, bitmap=True):
    """the window logic of c8b_stream_push (ctx.cu: stream_process) on the host-compiled detect routine: returns the
    decided frames with absolute indices"""
    pkg = load_pkg()
    H, O = hs.lib(), ol.oracle()
    win = np.zeros(0, np.complex64)
    base, frm, pos_abs, out, done = 0, 0, 0, [], 0

    def process(flush):
        """returns the scan's stall code (2: frame records used up)"""
        nonlocal win, base, frm, pos_abs
        n = win.size
        if (n <= frm and not flush) or n == 0:
            return 0
        xf = ol.c2f(np.ascontiguousarray(win))
        preac, preconj = np.zeros(n, np.float32), np.zeros(2 * n, np.float32)
        O.orx_presiso(xf, n, preac, preconj)
        f = np.zeros(maxf, pkg.FRAME_DTYPE)
        chan = np.zeros(128 * maxf, np.float32)
        sc = np.zeros(8, np.int32)
        sc[0], sc[1], sc[2] = frm, pos_abs - base, 1 if flush else 0
        H.hs_detect_scan(xf, preac, n, maxf, f.ctypes.data, chan, sc, 1 if bitmap else 0)
        safe, pos, nf = int(sc[3]), int(sc[4]), int(sc[5])
        if flush:
            nf = 0
            while nf < maxf and f[nf]["status"] != 9 and f[nf]["nsamp"] > 0:
                nf += 1
        for k in range(nf):
            out.append((base + int(f[k]["trig_idx"]), base + int(f[k]["sync_idx"]), int(f[k]["status"]), int(f[k]["l_mcs"]), int(f[k]["l_len"])))
        if flush:
            return 0
        keep = max(0, safe - 64)
        pos_abs = base + pos
        if keep == 0 and n >= cap:
            base, win, frm = base + n, win[:0], 0
            pos_abs = max(pos_abs, base)
            return 0
        win, base = win[keep:], base + keep
        frm = safe - keep
        return int(sc[6])

    for i, p in enumerate(pushes):
        last = i == len(pushes) - 1
        piece = x[done:done + p]
        done += p
        while True:
            room = cap - win.size
            take = min(room, piece.size)
            win = np.concatenate([win, piece[:take]])
            piece = piece[take:]
            while True:
                before = base + frm
                st = process(False)
                if st != 2 or base + frm == before:
                    break
            if piece.size == 0:
                if last:
                    process(True)
                break
    return out


def test_stream_windows_find_the_frames_of_one_pass(golden):
    """a capture pushed in arbitrary pieces through the window / restart-point logic yields exactly the frames one pass
    over the whole capture yields (indices, L-SIG fields), for clean and noisy captures, with and without the bitmap scan"""
    g = golden["frames_siso"]
    offs = g["offs"]
    rng = np.random.default_rng(21)
    x0 = np.ascontiguousarray(g["iq"][offs[1]:offs[12]])
    noisy = (x0 + 0.02 * (rng.standard_normal(x0.size) + 1j * rng.standard_normal(x0.size))).astype(np.complex64)
    for x in (x0, noisy):
        ref, _, _, _ = _detect(x, maxf=32)
        want = [(int(f["trig_idx"]), int(f["sync_idx"]), int(f["status"]), int(f["l_mcs"]), int(f["l_len"])) for f in ref if f["status"] != 9 and f["nsamp"] > 0]
        assert len(want) == 11
        for trial in range(6):
            cuts = np.sort(rng.integers(1, x.size, size=[3, 9, 40, 1, 200, 17][trial]))
            pushes = np.diff(np.concatenate([[0], cuts, [x.size]])).tolist()
            got = _stream_detect(x, pushes, maxf=[8, 3, 2, 8, 4, 16][trial], bitmap=trial % 2 == 0)
            assert got == want, (trial, got, want)
    # a capture cut in the middle of a frame at the end of the stream: flush reports the truncated frame like the batch pass
    xt = x0[:x0.size - 3000]
    ref, _, _, _ = _detect(xt, maxf=32)
    want = [(int(f["trig_idx"]), int(f["sync_idx"]), int(f["status"]), int(f["l_mcs"]), int(f["l_len"])) for f in ref if f["status"] != 9 and f["nsamp"] > 0]
    assert want[-1][2] == 4
    assert _stream_detect(xt, [5000, 7000, xt.size - 12000], maxf=8) == want
    # tiny windows: every window that fills without a decidable point is dropped, never a hang
    got = _stream_detect(x0, [x0.size], maxf=8, cap=4096)
    assert isinstance(got, list)
