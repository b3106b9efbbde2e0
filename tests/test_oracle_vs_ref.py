"""Pin the oracle: every function of oracle_rx.cc that also exists in the reference's
lib/cloud80211phy.cc must agree BIT-FOR-BIT with the unmodified reference compiled into
oracle/_ref/libc8p_ref.so.  Where _ref is absent (it always ships with gpurun, but a fresh clone
has no built artefacts) the same checks run against tests/golden/ref_vectors.npz, which was
produced from the reference by tests/golden/make_golden.py."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libc8p_ref.so not built (reference sources absent)")
RNG = np.random.default_rng(7)


def _crandn(n, s=1.0):
    return (RNG.normal(0, s, n) + 1j * RNG.normal(0, s, n)).astype(np.complex64)


# ---------------------------------------------------------------- tables (formula vs reference) ---
def test_tables_vs_golden(golden):
    g, O = golden["ref_vectors"], ol.oracle()
    buf = np.zeros(512, np.int32)
    names = {1: "Bpsk", 2: "Qpsk", 4: "16Qam", 6: "64Qam", 8: "256Qam"}
    for nb in (1, 2, 4, 6):
        n = O.orx_deint_map(0, nb, 1, buf)
        assert np.array_equal(buf[:n], g["tab_mapDeintLegacy" + names[nb]])
    for nb in (1, 2, 4, 6, 8):
        n = O.orx_deint_map(1, nb, 1, buf)
        assert np.array_equal(buf[:n], g["tab_mapDeintNonlegacy" + names[nb]])
        n = O.orx_deint_map(1, nb, 2, buf)
        assert np.array_equal(buf[:n], g["tab_deintNL2_%d" % nb])
    n = O.orx_deint_map(1, 1, 1, buf)
    assert np.array_equal(buf[:n], g["tab_mapDeintVhtSigB20"])
    f = np.zeros(127, np.float32)
    O.orx_pilot_polarity(f)
    assert np.array_equal(f, g["tab_PILOT_P"])
    l = np.zeros(64, np.float32)
    for kind, name in ((0, "LTF_L_26_F_FLOAT"), (1, "LTF_NL_28_F_FLOAT"), (2, "LTF_NL_28_F_FLOAT_VHT22")):
        O.orx_ltf(kind, l)
        assert np.array_equal(l, g["tab_" + name]), name
    nx, op = np.zeros(128, np.int32), np.zeros(128, np.int32)
    O.orx_trellis_tables(nx, op)
    assert np.array_equal(nx, g["tab_SV_STATE_NEXT"]) and np.array_equal(op, g["tab_SV_STATE_OUTPUT"])


# ---------------------------------------------------------------- Viterbi -------------------------
def test_viterbi_vs_golden(golden):
    g, O = golden["ref_vectors"], ol.oracle()
    for k in list(range(5)) + ["tie"]:
        llr, want = g["vit_llr_%s" % k], g["vit_bits_%s" % k]
        got = np.zeros(want.size, np.uint8)
        O.orx_viterbi(np.ascontiguousarray(llr), 0, want.size, got)
        assert np.array_equal(got, want), k
        if want.size <= 48:
            got2 = np.zeros(want.size, np.uint8)
            O.orx_sig_viterbi(np.ascontiguousarray(llr), got2, want.size)
            assert np.array_equal(got2, want)


@needs_ref
@pytest.mark.parametrize("tl", [24, 26, 48])
def test_sig_viterbi_vs_ref(tl):
    O, R = ol.oracle(), ol.ref()
    for trial in range(200):
        llr = RNG.normal(0, 1, 2 * tl).astype(np.float32)
        if trial % 3 == 0:
            llr = (np.round(llr * 2) / 2).astype(np.float32)   # force ties
        a, b = np.zeros(tl, np.uint8), np.zeros(tl, np.uint8)
        O.orx_sig_viterbi(llr, a, tl)
        R.ref_sig_viterbi(llr, b, tl)
        assert np.array_equal(a, b)


@needs_ref
def test_long_viterbi_vs_ref_ties_and_noise():
    O, R = ol.oracle(), ol.ref()
    for tl, q in ((777, 0), (1500, 4), (5000, 0), (12220, 2)):
        llr = RNG.normal(0, 2, 2 * tl).astype(np.float32)
        if q:
            llr = (np.round(llr * q) / q).astype(np.float32)
        a, b = np.zeros(tl, np.uint8), np.zeros(tl, np.uint8)
        O.orx_viterbi(llr, 0, tl, a)
        R.ref_sv_decode(llr, b, tl)
        assert np.array_equal(a, b)


# ---------------------------------------------------------------- SIG demod / checks / parsers ----
@needs_ref
def test_lsig_and_nlsig_demod_vs_ref():
    O, R = ol.oracle(), ol.ref()
    for _ in range(50):
        s1, s2, sg = _crandn(64), _crandn(64), _crandn(64)
        h1, h2 = np.zeros(128, np.float32), np.zeros(128, np.float32)
        l1, l2 = np.zeros(48, np.float32), np.zeros(48, np.float32)
        O.orx_lsig_demod(ol.c2f(s1), ol.c2f(s2), ol.c2f(sg), h1, l1)
        R.ref_lsig_demod(ol.c2f(s1), ol.c2f(s2), ol.c2f(sg), h2, l2)
        assert np.array_equal(h1, h2) and np.array_equal(l1, l2)
        h = ol.c2f(_crandn(64)).copy()
        a1, a2, b1, b2 = (np.zeros(96, np.float32) for _ in range(4))
        O.orx_nlsig_demod(ol.c2f(s1), ol.c2f(s2), h, a1, a2)
        R.ref_nlsig_demod(ol.c2f(s1), ol.c2f(s2), h, b1, b2)
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


def test_lsig_demod_vs_golden(golden):
    g, O = golden["ref_vectors"], ol.oracle()
    h, l = np.zeros(128, np.float32), np.zeros(48, np.float32)
    O.orx_lsig_demod(ol.c2f(g["lsig_s1"]), ol.c2f(g["lsig_s2"]), ol.c2f(g["lsig_sig"]), h, l)
    assert np.array_equal(h.view(np.complex64), g["lsig_h"]) and np.array_equal(l, g["lsig_llr"])
    a, b = np.zeros(96, np.float32), np.zeros(96, np.float32)
    O.orx_nlsig_demod(ol.c2f(g["lsig_s1"]), ol.c2f(g["lsig_sig"]), ol.c2f(g["nlsig_h"]), a, b)
    assert np.array_equal(a, g["nlsig_llrht"]) and np.array_equal(b, g["nlsig_llrvht"])


@needs_ref
def test_checks_and_parsers_vs_ref():
    O, R = ol.oracle(), ol.ref()
    nmod = R.ref_mod_nints()
    assert nmod == 65
    for trial in range(3000):
        b = RNG.integers(0, 2, 48).astype(np.uint8)
        # bias towards passing patterns so that the parsers see valid fields too
        if trial % 2:
            b[3], b[4] = 1, 0
            b[17] = b[:17].sum() & 1
        m1, l1, d1 = C.c_int(), C.c_int(), C.c_int()
        m2, l2, d2 = C.c_int(), C.c_int(), C.c_int()
        r1 = O.orx_check_legacy(b[:24].copy(), C.byref(m1), C.byref(l1), C.byref(d1))
        r2 = R.ref_check_legacy(b[:24].copy(), C.byref(m2), C.byref(l2), C.byref(d2))
        assert r1 == r2
        if r1:
            assert (m1.value, l1.value, d1.value) == (m2.value, l2.value, d2.value)
            a, c = np.zeros(17, np.int32), np.zeros(65, np.int32)
            O.orx_parse_l(m1.value, l1.value, a)
            R.ref_parse_l(m2.value, l2.value, c)
            assert np.array_equal(a, c[:17])
    for trial in range(3000):
        b = RNG.integers(0, 2, 48).astype(np.uint8)
        if trial % 2:   # make the CRC and reserved bits right so the parsers run on valid words
            b[2] = b[23] = b[33] = 1
            b[26] = 1
            if trial % 4 == 1:
                b[0] = b[1] = 0
                b[4:10] = 0 if trial % 8 == 1 else 1            # SU group ids 0 / 63
                b[31] = 0                                         # mcs <= 7 (valid index range also covers 8,9)
            else:
                b[5] = b[6] = b[7] = b[28] = b[29] = b[30] = b[32] = b[33] = 0
                b[26] = 1
            crc = np.zeros(8, np.uint8)
            R.ref_crc8_gen(b[:34].copy(), 34, crc)
            b[34:42] = crc
        assert O.orx_crc8_check(b[:34].copy(), 34, b[34:42].copy()) == R.ref_crc8_check(b[:34].copy(), 34, b[34:42].copy())
        assert O.orx_check_ht(b.copy()) == R.ref_check_ht(b.copy())
        assert O.orx_check_vhta(b.copy()) == R.ref_check_vhta(b.copy())
        if R.ref_check_ht(b.copy()):
            a, c = np.zeros(17, np.int32), np.zeros(65, np.int32)
            O.orx_parse_ht(b.copy(), a)
            R.ref_parse_ht(b.copy(), c)
            assert np.array_equal(a, c[:17])
        if R.ref_check_vhta(b.copy()):
            a, c = np.zeros(17, np.int32), np.zeros(65, np.int32)
            O.orx_parse_vhta(b.copy(), a)
            R.ref_parse_vhta(b.copy(), c)
            su = c[1] == 0
            if su and c[9] <= 9:
                assert np.array_equal(a, c[:17]), (a, c[:17])
                sb = RNG.integers(0, 2, 26).astype(np.uint8)
                if trial % 3:
                    sb[17:20] = 1
                O.orx_parse_vhtb(sb.copy(), a)
                R.ref_parse_vhtb(sb.copy(), c)
                assert np.array_equal(a, c[:17])


# ---------------------------------------------------------------- LLR demap -----------------------
def test_qam_to_llr_vs_golden(golden):
    g, O = golden["ref_vectors"], ol.oracle()
    for mod, nsd in ((0, 48), (2, 48), (3, 48), (4, 48), (0, 52), (2, 52), (3, 52), (4, 52), (5, 52)):
        q = g["llr_in_%d_%d" % (mod, nsd)]
        want = g["llr_out_%d_%d" % (mod, nsd)]
        got = np.zeros(want.size, np.float32)
        O.orx_qam_to_llr(ol.c2f(q).copy(), got, mod, nsd)
        assert np.array_equal(got, want), (mod, nsd)


@needs_ref
def test_bcc_crc8_scrambler_vs_ref():
    O, R = ol.oracle(), ol.ref()
    # descrambler inverts the reference TX scrambler for every seed (lib/decode_impl.cc:304-323 vs c8p.cc:2594-2606)
    for seed in range(1, 128):
        bits = RNG.integers(0, 2, 300).astype(np.uint8)
        bits[:7] = 0                                              # service field starts with 7 zero bits
        sc = np.zeros(300, np.uint8)
        R.ref_scramble(bits, sc, 300, seed)
        back = np.zeros(300, np.uint8)
        O.orx_descramble(sc, 300, back)
        assert np.array_equal(back, bits), seed
