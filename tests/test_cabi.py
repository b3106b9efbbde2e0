"""The C-ABI library loads and exports every symbol include/c80211b200.h declares; the LUT blob built
by formula equals the reference's transcribed tables (no compute calls: runs without a GPU)."""
import os
import re

import numpy as np

from __graft_entry__ import ROOT, build, load_pkg


def test_build_and_symbols():
    build()
    pkg = load_pkg()
    L = pkg._cabi.lib()
    hdr = open(os.path.join(ROOT, "include", "c80211b200.h")).read()
    declared = set(re.findall(r"\b(c8b_[a-z0-9_]+)\s*\(", hdr))
    bound = {s[0] for s in pkg._cabi.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(L, name), name
    assert L.c8b_abi_version() == 1


def test_no_device_fails_loudly():
    import ctypes as C
    pkg = load_pkg()
    L = pkg._cabi.lib()
    if L.c8b_device_count() > 0:
        return
    h = C.c_void_p()
    assert L.c8b_create(None, C.byref(h)) == -1          # C8B_ERR_NO_DEVICE: no CPU fallback
    assert b"no CPU path" in L.c8b_last_error(None)
    try:
        pkg.Receiver()
    except pkg.C8bError:
        pass
    else:
        raise AssertionError("Receiver() must raise without a GPU")


def test_lut_blob_matches_reference_tables(golden):
    pkg = load_pkg()
    blob = pkg.lut_blob()
    g = golden["ref_vectors"]
    # layout of struct c8b_lut (csrc/lut.h)
    o = 16
    def take(n, dt):
        nonlocal o
        a = np.frombuffer(blob, dt, n, o)
        o += n * np.dtype(dt).itemsize
        return a
    ltfL, ltfNL, ltfNL22 = take(64, "<f4"), take(64, "<f4"), take(64, "<f4")
    pilotP = take(128, "<f4")
    twr, twi = take(64, "<f4"), take(64, "<f4")
    twdr, twdi = take(32, "<f8"), take(32, "<f8")
    deintL = take(4 * 288, "<u2").reshape(4, 288)
    deintNL = take(2 * 5 * 416, "<u2").reshape(2, 5, 416)
    sigDemap = take(64, "i1")
    bmClass = take(32, "u1")
    binL, binNL = take(64, "u1"), take(64, "u1")
    crc = take(256, "<u4")
    crcz = take(6 * 32, "<u4").reshape(6, 32)
    pair01 = take(2, "<f4")
    o = (o + 15) & ~15                                   # alignas(16)
    demap = take(9 * 64, "<u2").reshape(9, 8, 8)
    demap2 = take(5 * 2 * 64, "<u2").reshape(5, 2, 8, 8)
    tw8 = take(128, "<f4").reshape(4, 8, 2, 2)             # [k1 / 2][j][k1 % 2][re, im]
    assert o == blob.size and pair01.tolist() == [0.0, 1.0]
    # k_demod's per-thread tables restate deintL / deintNL[0] / binToData* / tw*: entry (B, R) of bin j + 8*k2 puts soft bit
    # h*s + c at B + N_COL*((c + R) mod s) + h*N_COL*s
    for mode in range(9):
        leg = mode < 4
        nb = (1, 2, 4, 6)[mode] if leg else (1, 2, 4, 6, 8)[mode - 4]
        s_, ncol = max(nb // 2, 1), 16 if leg else 13
        mp = deintL[mode] if leg else deintNL[0, mode - 4]
        b2d = binL if leg else binNL
        for j in range(8):
            for k2 in range(8):
                e, d = int(demap[mode, j, k2]), int(b2d[j + 8 * k2])
                assert (e == 0xFFFF) == (d == 255)
                if d != 255:
                    B, R = e & 511, e >> 9
                    got = [B + ncol * ((c + R) % s_) + h * ncol * s_ for h in range(nb // s_) for c in range(s_)]
                    assert got == mp[d * nb:(d + 1) * nb].tolist(), (mode, j, k2)
    # two-stream symbols: deinterleaver of stream a + stream parser (c = a*s + 2*s*(k // s) + k % s)
    for m, nb in enumerate((1, 2, 4, 6, 8)):
        s_ = max(nb // 2, 1)
        for a in range(2):
            for j in range(8):
                for k2 in range(8):
                    e, d = int(demap2[m, a, j, k2]), int(binNL[j + 8 * k2])
                    assert (e == 0xFFFF) == (d == 255)
                    if d != 255:
                        P0, R, beta = e & 1023, (e >> 10) & 3, e >> 12
                        got = [P0 + 13 * ((c + R) % s_) + s_ * ((beta + 13 * ((c + R) % s_)) // s_) + 26 * s_ * h
                               for h in range(nb // s_) for c in range(s_)]
                        k = deintNL[a, m, d * nb:(d + 1) * nb].astype(int)
                        assert got == (a * s_ + 2 * s_ * (k // s_) + k % s_).tolist(), (m, a, j, k2)
    for j in range(8):
        for k1 in range(8):
            assert tw8[k1 >> 1, j, k1 & 1, 0] == twr[(j * k1) & 63] and tw8[k1 >> 1, j, k1 & 1, 1] == twi[(j * k1) & 63]
    assert np.array_equal(ltfL, g["tab_LTF_L_26_F_FLOAT"]) and np.array_equal(ltfNL, g["tab_LTF_NL_28_F_FLOAT"])
    assert np.array_equal(ltfNL22, g["tab_LTF_NL_28_F_FLOAT_VHT22"])
    assert np.array_equal(pilotP[:127], g["tab_PILOT_P"])
    names = {1: "Bpsk", 2: "Qpsk", 4: "16Qam", 6: "64Qam", 8: "256Qam"}
    for m, nb in enumerate((1, 2, 4, 6)):
        assert np.array_equal(deintL[m, :48 * nb], g["tab_mapDeintLegacy" + names[nb]])
    for m, nb in enumerate((1, 2, 4, 6, 8)):
        assert np.array_equal(deintNL[0, m, :52 * nb], g["tab_mapDeintNonlegacy" + names[nb]])
        assert np.array_equal(deintNL[1, m, :52 * nb], g["tab_deintNL2_%d" % nb])
    # trellis: class of 2k --0--> k from the reference's SV_STATE_OUTPUT[state*2 + input]
    assert np.array_equal(bmClass, g["tab_SV_STATE_OUTPUT"][0::2][0::2])
    nxt = g["tab_SV_STATE_NEXT"]
    for k in range(32):
        assert nxt[(2 * k) * 2] == k and nxt[(2 * k + 1) * 2] == k and nxt[(2 * k) * 2 + 1] == k + 32
        o2 = g["tab_SV_STATE_OUTPUT"]
        c = o2[(2 * k) * 2]
        assert o2[(2 * k + 1) * 2] == c ^ 3 and o2[(2 * k) * 2 + 1] == c ^ 3 and o2[(2 * k + 1) * 2 + 1] == c
    import zlib
    assert crc[1] == 0x77073096 and zlib.crc32(b"123456789") == 0xCBF43926
    # crcZ[p] advances the raw register by 64*2^p zero bytes: crc(m + zeros) follows from crc(m) linearly
    def raw(data, c=0xFFFFFFFF):
        for b in data:
            c = int(crc[(c ^ b) & 0xFF]) ^ (c >> 8)
        return c
    c0 = raw(b"hello world")
    for p_ in range(6):
        want = raw(bytes(64 << p_), c0)
        got = 0
        for i in range(32):
            if (c0 >> i) & 1:
                got ^= int(crcz[p_, i])
        assert got == want, p_
    w = np.exp(-2j * np.pi * np.arange(64) / 64)
    assert np.allclose(twr, w.real, atol=1e-7) and np.allclose(twi, w.imag, atol=1e-7)
    assert sigDemap[0] == -1 and (sigDemap >= 0).sum() == 48 and (binL < 255).sum() == 48 and (binNL < 255).sum() == 52
    assert binNL[1] == 26 and binL[1] == 24 and binNL[36] == 0 and binL[38] == 0


def test_tx_udp_parse_follows_pktgen():
    """MAC -> PHY datagram [format][mcs][nss][len16 LE][PSDU] (lib/pktgen_impl.cc:57-70 msgRead, :96-118 pktPop; writer
    tools/phy80211.py:1126-1137 genPktGrData): host-side parser of c8b_tx_from_udp, no GPU needed"""
    import ctypes as C
    import struct
    pkg = load_pkg()
    L = pkg._cabi.lib()
    f = np.zeros(1, pkg.TXFRAME_DTYPE)
    body = C.c_void_p()

    def parse(b):
        buf = (C.c_ubyte * max(len(b), 1)).from_buffer_copy(bytes(b) if b else b"\0")
        rc = L.c8b_tx_udp_parse(buf, len(b), f.ctypes.data, C.byref(body))
        return rc, (body.value - C.addressof(buf) if rc > 0 else None)

    mpdu = bytes(range(100))
    rc, o = parse(struct.pack("<BBBH", 2, 7, 1, len(mpdu)) + mpdu)          # genPktGrData(mpdu, VHT MCS7 one stream)
    assert rc == 1 and o == 5 and (f[0]["format"], f[0]["mcs"], f[0]["psdu_len"]) == (2, 7, 100)
    rc, _ = parse(struct.pack("<BBBH", 0, 3, 1, 40) + bytes(60))             # longer datagram than len: accepted (pktPop checks >=)
    assert rc == 1 and f[0]["psdu_len"] == 40
    assert parse(b"\x00\x00\x01\x10")[0] < 0                                 # under 5 bytes (msgRead :65-67)
    assert parse(struct.pack("<BBBH", 0, 0, 1, 50) + bytes(49))[0] < 0       # shorter than its len field (pktPop :108-111)
    assert parse(struct.pack("<BBBH", 0, 0, 1, 4096) + bytes(4096))[0] < 0   # len > 4095
    assert parse(struct.pack("<BBBH", 3, 0, 1, 10) + bytes(10))[0] < 0       # C8P_F_VHT_MU: two users per datagram
    rc, _ = parse(struct.pack("<BBBH", 1, 11, 2, 10) + bytes(10))            # HT MCS11, two streams (c8b_tx_batch2)
    assert rc == 2 and f[0]["mcs"] == 11
    rc, _ = parse(struct.pack("<BBBH", 2, 5, 2, 12) + bytes(12))             # VHT MCS5 x 2 streams: descriptor mcs 16 + 5
    assert rc == 2 and f[0]["mcs"] == 21
    assert parse(struct.pack("<BBBH", 1, 3, 2, 10) + bytes(10))[0] < 0       # HT MCS3 is one stream
    assert parse(struct.pack("<BBBH", 2, 5, 3, 12) + bytes(12))[0] < 0       # three streams
    assert parse(struct.pack("<BBBH", 0, 9, 1, 10) + bytes(10))[0] < 0       # no legacy MCS 9


def test_tx_udp_parse_mu_and_bfq():
    """the MU demo's datagrams (tools/phy80211.py:1139-1171 genPktGrDataMu / genPktGrBfQ; lib/pktgen_impl.cc:101-113,
    lib/modulation2_impl.cc:117-121), host-side parsers, no GPU needed"""
    import ctypes as C
    import struct
    pkg = load_pkg()
    L = pkg._cabi.lib()
    f = np.zeros(1, pkg.TXMU_DTYPE)
    p0, p1 = C.c_void_p(), C.c_void_p()

    def parse(b):
        buf = (C.c_ubyte * max(len(b), 1)).from_buffer_copy(bytes(b))
        rc = L.c8b_tx_udp_parse_mu(buf, len(b), f.ctypes.data, C.byref(p0), C.byref(p1))
        return rc, ((p0.value - C.addressof(buf), p1.value - C.addressof(buf)) if rc == 0 else None)

    a0, a1 = bytes(range(100)), bytes(range(192))
    hdr = lambda m0, n0, l0, m1, n1, l1, g: struct.pack("<BBBHBBHB", 3, m0, n0, l0, m1, n1, l1, g)
    rc, o = parse(hdr(4, 1, len(a0), 2, 1, len(a1), 2) + a0 + a1)            # genPktGrDataMu
    assert rc == 0 and o == (10, 110)
    assert f[0]["mcs"].tolist() == [4, 2] and f[0]["psdu_len"].tolist() == [100, 192] and f[0]["group_id"] == 2
    assert parse(hdr(4, 1, 100, 2, 1, 192, 2) + a0 + a1[:-1])[0] < 0          # shorter than the two length fields say (pktPop :127-130)
    assert parse(hdr(4, 1, 2048, 2, 1, 2048, 2) + bytes(4096))[0] < 0         # more than 4095 bytes in all
    assert parse(hdr(4, 2, 100, 2, 1, 192, 2) + a0 + a1)[0] < 0               # two streams for one user: not synthesised
    assert parse(hdr(9, 1, 100, 2, 1, 192, 2) + a0 + a1)[0] < 0               # no VHT MCS 9 at 20 MHz
    assert parse(hdr(4, 1, 100, 2, 1, 192, 0) + a0 + a1)[0] < 0               # group id 0 is single user
    assert parse(hdr(4, 1, 98, 2, 1, 192, 2) + a0[:98] + a1)[0] < 0           # an A-MPDU is a multiple of 4 bytes
    assert parse(struct.pack("<BBBH", 2, 0, 1, 8) + bytes(8))[0] < 0          # not an MU datagram
    rng = np.random.default_rng(5)
    q = (rng.normal(0, 1, (64, 2, 2)) + 1j * rng.normal(0, 1, (64, 2, 2))).astype(np.complex64)
    pkt = bytes([10]) + b"".join(struct.pack("<ff", float(q[k, i, j].real), float(q[k, i, j].imag)) for k in range(64) for i in range(2) for j in range(2))
    out = np.zeros(512, np.float32)
    buf = (C.c_ubyte * len(pkt)).from_buffer_copy(pkt)
    assert L.c8b_tx_udp_parse_bfq(buf, len(pkt), out.ctypes.data) == 0 and np.array_equal(out.view(np.complex64).reshape(64, 2, 2), q)
    assert L.c8b_tx_udp_parse_bfq(buf, len(pkt) - 1, out.ctypes.data) < 0     # lib/modulation2_impl.cc:117: exactly 2049 bytes
