#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE ITSELF.  Runs only in the dev
container (needs /root/reference); the GPU box and the test-suite only read the .npz files.

  frames_siso.npz     waveforms made by the reference's own generator tools/phy80211.py following
                      the recipe of tools/pktGenExample.py:173-199 (payload "1234567890"x3, L MCS0-7,
                      HT MCS0-7, VHT MCS0-8, multiplier 12, scrambler seed 93) + the MPDU each must
                      decode to.  Config 1 of BASELINE.json is item 0 (gapLen 1200 as in the recipe).
  frames_bench.npz    16 VHT MCS7 frames carrying random 1500-byte MPDUs (A-MPDU 1504 B, 47 symbols,
                      4560 samples): the unique frames bench.py replicates for config 5.
  ref_vectors.npz     input/output pairs of the unmodified lib/cloud80211phy.cc functions
                      (through oracle/_ref/libc8p_ref.so) so GPU tests can compare with the
                      reference's own outputs where /root/reference does not exist.

  frames_mu.npz       VHT NDP (2 transmit streams, one receive antenna) and 2-user MU-MIMO frames seen at each user
                      position (tools/cmu_ap.py recipe with a flat 2x2 channel and its zero-forcing precoder).

  frames_mu_tx.npz    2-user MU-MIMO frames as TRANSMITTED (both antennas), their A-MPDUs and the per-subcarrier spatial
                      mapping matrices: the MU side of the transmit synthesiser.

  frames_tx.npz       the PSDU bytes behind frames_siso / frames_bench (inputs of genFromMpdu / genFromAmpdu) for the
                      transmit synthesiser test.

usage: python tests/golden/make_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
STUB = "/tmp/_mpl_stub"
os.makedirs(os.path.join(STUB, "matplotlib"), exist_ok=True)
open(os.path.join(STUB, "matplotlib", "__init__.py"), "w").close()
with open(os.path.join(STUB, "matplotlib", "pyplot.py"), "w") as f:
    f.write("def __getattr__(n):\n    return lambda *a, **k: None\n")
sys.path.insert(0, STUB)
sys.path.insert(0, "/root/reference/tools")

with contextlib.redirect_stdout(io.StringIO()):
    import mac80211  # noqa: E402
    import phy80211  # noqa: E402
    import phy80211header as p8h  # noqa: E402
    import pktGenExample as pge  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def gen(phy, fmt, mcs, pkt, gap, cfo=0.0, mult=12.0, nsts=1):
    mod = p8h.modulation(phyFormat=fmt, mcs=mcs, bw=p8h.BW.BW20, nSTS=nsts, shortGi=False)
    if fmt == p8h.F.VHT:
        quiet(phy.genFromAmpdu, pkt, mod, vhtPartialAid=0, vhtGroupId=0)
    else:
        quiet(phy.genFromMpdu, pkt, mod)
    ss = quiet(phy.genFinalSig, multiplier=mult, cfoHz=cfo, num=1, gap=True, gapLen=gap)
    return [np.asarray(s, dtype=np.complex64) for s in ss]


def mac_mpdu(payload):
    return quiet(pge.genMac80211UdpMPDU, payload)


def mac_ampdu(payloads):
    return quiet(pge.genMac80211UdpAmpduVht, payloads)


def ampdu_split(ampdu):
    """MPDUs inside a VHT A-MPDU built by tools/mac80211.py:333-360 (delimiter: EOF, rsvd, len[12:14], len[0:12], CRC8, 0x4E)."""
    out, i, ampdu = [], 0, bytes(ampdu)
    while i + 4 <= len(ampdu):
        d0, d1 = ampdu[i], ampdu[i + 1]
        ln = ((d0 >> 4) & 0xF) | (d1 << 4) | (((d0 >> 2) & 3) << 12)
        out.append(ampdu[i + 4: i + 4 + ln])
        i += 4 + (ln + 3) // 4 * 4
    return out


def frames_siso():
    phy = phy80211.phy80211(ifDebug=False)
    payload = "123456789012345678901234567890"
    mpdu = mac_mpdu(payload)
    ampdu = mac_ampdu([payload])
    items, meta, exp = [], [], []
    # config 1: exactly the reference recipe (gap 1200)
    items.append(gen(phy, p8h.F.L, 0, mpdu, 1200)[0]); meta.append((0, 0, 0.0)); exp.append(bytes(mpdu))
    for fmt, code, rng, pkt in ((p8h.F.L, 0, range(0, 8), mpdu), (p8h.F.HT, 1, range(0, 8), mpdu), (p8h.F.VHT, 2, range(0, 9), ampdu)):
        for mcs in rng:
            items.append(gen(phy, fmt, mcs, pkt, 400)[0]); meta.append((code, mcs, 0.0))
            exp.append(ampdu_split(pkt)[0] if code == 2 else bytes(mpdu))
    # CFO cases (generator applies the CFO itself): +/- 50 kHz, 233 kHz (tools/performance/perf_wime.py:134-137)
    for fmt, code, mcs, pkt, cfo in ((p8h.F.L, 0, 4, mpdu, 50e3), (p8h.F.HT, 1, 7, mpdu, -50e3), (p8h.F.VHT, 2, 8, ampdu, 100e3),
                                     (p8h.F.VHT, 2, 7, ampdu, -233e3)):
        items.append(gen(phy, fmt, mcs, pkt, 400, cfo=cfo)[0]); meta.append((code, mcs, cfo))
        exp.append(ampdu_split(pkt)[0] if code == 2 else bytes(mpdu))
    # a 2-subframe VHT A-MPDU (exercises the de-aggregation walk, lib/decode_impl.cc:337-431)
    ampdu2 = mac_ampdu([payload, "This is packet for station 001"])
    items.append(gen(phy, p8h.F.VHT, 5, ampdu2, 400)[0]); meta.append((2, 5, 0.0)); exp.append(ampdu_split(ampdu2)[0])
    offs = np.cumsum([0] + [len(x) for x in items]).astype(np.int64)
    iq = np.concatenate(items).astype(np.complex64)
    explen = np.array([len(e) for e in exp], np.int32)
    expbuf = np.frombuffer(b"".join(exp), np.uint8)
    np.savez_compressed(os.path.join(HERE, "frames_siso.npz"), iq=iq, offs=offs, meta=np.array(meta, np.float64), exp_len=explen,
                        exp_mpdu=expbuf, ampdu2_second=np.frombuffer(ampdu_split(ampdu2)[1], np.uint8))
    print("frames_siso: %d items, %d samples" % (len(items), iq.size))


def frames_bench():
    phy = phy80211.phy80211(ifDebug=False)
    rng = np.random.default_rng(80211)
    frames, mpdus = [], []
    for i in range(16):
        # 1500-byte MPDU: 26-byte QoS 802.11 header + 8 LLC + 20 IPv4 + 8 UDP + payload + 4 FCS = 1500 -> payload 1434
        # (the generator takes str payloads: latin-1 keeps one byte per char)
        payload = bytes(rng.integers(1, 128, 1434, dtype=np.uint8)).decode("latin-1")
        ampdu = mac_ampdu([payload])
        mpdu = ampdu_split(ampdu)[0]
        assert len(mpdu) == 1500 and len(ampdu) == 1504, (len(mpdu), len(ampdu))
        s = gen(phy, p8h.F.VHT, 7, ampdu, 0)[0]
        assert s.size == 4560, s.size
        frames.append(s); mpdus.append(np.frombuffer(bytes(mpdu), np.uint8))
    np.savez_compressed(os.path.join(HERE, "frames_bench.npz"), iq=np.stack(frames).astype(np.complex64), mpdu=np.stack(mpdus))
    print("frames_bench: 16 x", frames[0].size)


def frames_mimo():
    """2x2 SU-MIMO frames (tools/pktGenExample.py:206-217, tools/performance/perf_sumimo.py:139-148): HT MCS8-15 and
    VHT 2SS MCS0-8, multiplier 12*sqrt(2), stream k -> antenna k (identity channel as tools/performance/gr_sumimo.py:70-78)."""
    phy = phy80211.phy80211(ifDebug=False)
    payload = "123456789012345678901234567890"
    mpdu = mac_mpdu(payload)
    ampdu = mac_ampdu([payload])
    a0, a1, meta, exp = [], [], [], []
    for fmt, code, rng, pkt in ((p8h.F.HT, 1, range(8, 16), mpdu), (p8h.F.VHT, 2, range(0, 9), ampdu)):
        for mcs in rng:
            ss = gen(phy, fmt, mcs, pkt, 400, mult=12.0 * np.sqrt(2), nsts=2)
            assert len(ss) == 2 and ss[0].size == ss[1].size
            a0.append(ss[0]); a1.append(ss[1]); meta.append((code, mcs, 0.0))
            exp.append(ampdu_split(pkt)[0] if code == 2 else bytes(mpdu))
    # one frame with CFO
    ss = gen(phy, p8h.F.HT, 11, mpdu, 400, cfo=75e3, mult=12.0 * np.sqrt(2), nsts=2)
    a0.append(ss[0]); a1.append(ss[1]); meta.append((1, 11, 75e3)); exp.append(bytes(mpdu))
    offs = np.cumsum([0] + [len(x) for x in a0]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "frames_mimo.npz"), iq0=np.concatenate(a0).astype(np.complex64),
                        iq1=np.concatenate(a1).astype(np.complex64), offs=offs, meta=np.array(meta, np.float64),
                        exp_len=np.array([len(e) for e in exp], np.int32), exp_mpdu=np.frombuffer(b"".join(exp), np.uint8))
    print("frames_mimo: %d items, %d samples per antenna" % (len(a0), offs[-1]))


def frames_564():
    """BASELINE configs 2-4 units: 564-byte MPDUs (500 random UDP bytes, tools/performance/perf_siso.py:126,132) at every MCS:
    legacy 0-7, VHT 1SS 0-8 (A-MPDU of one MPDU), HT 2x2 8-15 (two antennas).  No gaps: bench/tests add them."""
    phy = phy80211.phy80211(ifDebug=False)
    rng = np.random.default_rng(80211)
    payload = bytes(rng.integers(1, 128, 500, dtype=np.uint8)).decode("latin-1")
    mpdu = mac_mpdu(payload)
    ampdu = mac_ampdu([payload])
    assert len(mpdu) == 564, len(mpdu)
    out = {"mpdu": np.frombuffer(bytes(mpdu), np.uint8), "vht_mpdu": np.frombuffer(ampdu_split(ampdu)[0], np.uint8)}
    for mcs in range(8):
        out["l%d" % mcs] = gen(phy, p8h.F.L, mcs, mpdu, 0)[0]
    for mcs in range(9):
        out["v%d" % mcs] = gen(phy, p8h.F.VHT, mcs, ampdu, 0)[0]
    for mcs in range(8, 16):
        ss = gen(phy, p8h.F.HT, mcs, mpdu, 0, mult=12.0 * np.sqrt(2), nsts=2)
        out["h%d_0" % mcs], out["h%d_1" % mcs] = ss[0], ss[1]
    np.savez_compressed(os.path.join(HERE, "frames_564.npz"), **out)
    print("frames_564:", {k: v.size for k, v in out.items()})


def frames_sgi():
    """HT / VHT frames whose SIG field announces short GI.  The reference TX has no short-GI waveform (lib/cloud80211phy.cc:2489,
    tools/phy80211.py builds 80-sample symbols regardless) but its RX honours the bit (nSymSamp = 72, c8p.cc:898-902,1155-1159):
    such frames demodulate on a 72-sample raster and fail the CRC -- a parity case for the nSymSamp = 72 path."""
    phy = phy80211.phy80211(ifDebug=False)
    payload = "123456789012345678901234567890"
    mpdu, ampdu = mac_mpdu(payload), mac_ampdu([payload])
    items = []
    for fmt, pkt, mcs in ((p8h.F.HT, mpdu, 5), (p8h.F.VHT, ampdu, 7), (p8h.F.HT, mpdu, 0)):
        mod = p8h.modulation(phyFormat=fmt, mcs=mcs, bw=p8h.BW.BW20, nSTS=1, shortGi=True)
        if fmt == p8h.F.VHT:
            quiet(phy.genFromAmpdu, pkt, mod, vhtPartialAid=0, vhtGroupId=0)
        else:
            quiet(phy.genFromMpdu, pkt, mod)
        ss = quiet(phy.genFinalSig, multiplier=12.0, cfoHz=0.0, num=1, gap=True, gapLen=400)
        items.append(np.asarray(ss[0], dtype=np.complex64))
    offs = np.cumsum([0] + [len(x) for x in items]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "frames_sgi.npz"), iq=np.concatenate(items).astype(np.complex64), offs=offs)
    print("frames_sgi:", [len(x) for x in items])


def frames_sgi_true():
    """Short-GI frames the reference's receiver DECODES.  The generator writes the SIG fields for short GI but builds 80-sample
    symbols (frames_sgi: every symbol after the first is off the 72-sample raster, CRC fails).  The reference's demod reads the
    window [8, 72) of every symbol -- 8 samples into the cyclic prefix, the same offset its channel estimate has from the LTF
    (C8P_SYM_SAMP_SHIFT) -- so the waveform it decodes on a 72-sample raster is the FIRST 72 samples of each 80-sample DATA
    symbol (a transmitter's own short-GI symbol [8-sample prefix][body] would sit 8 samples off that estimate).  These
    frames pass the CRC: the well-conditioned parity case for nSymSamp = 72 on every symbol.  HT MCS 0 / 5 / 7, VHT MCS 3 / 7."""
    phy = phy80211.phy80211(ifDebug=False)
    rng = np.random.default_rng(72)
    items, mpdus = [], []
    for fmt, mcs, nbytes in ((p8h.F.HT, 5, 94), (p8h.F.VHT, 7, 94), (p8h.F.HT, 0, 60), (p8h.F.HT, 7, 700), (p8h.F.VHT, 3, 400)):
        body = bytes(rng.integers(0, 256, nbytes - 4, dtype=np.uint8))
        import zlib
        mp = body + (zlib.crc32(body) & 0xffffffff).to_bytes(4, "little")
        mod = p8h.modulation(phyFormat=fmt, mcs=mcs, bw=p8h.BW.BW20, nSTS=1, shortGi=True)
        if fmt == p8h.F.VHT:
            quiet(phy.genFromAmpdu, mac80211.genAmpduVHT([mp]), mod, vhtPartialAid=0, vhtGroupId=0)
        else:
            quiet(phy.genFromMpdu, mp, mod)
        x = np.asarray(quiet(phy.genFinalSig, multiplier=12.0, cfoHz=0.0, num=1, gap=True, gapLen=400)[0], dtype=np.complex64)
        # frame = 400 zeros + preamble + data + 400 zeros; data starts after L-STF/L-LTF/L-SIG (400) + SIG (160) + STF (80) + LTF (80) [+ SIG-B (80)]
        pre = 400 + 400 + 160 + 80 + 80 + (80 if fmt == p8h.F.VHT else 0)
        nsym = (x.size - 400 - pre) // 80
        assert pre + nsym * 80 + 400 == x.size, (x.size, pre, nsym)
        data = x[pre:pre + nsym * 80].reshape(nsym, 80)[:, :72]
        items.append(np.concatenate([x[:pre], data.reshape(-1), x[pre + nsym * 80:]]).astype(np.complex64))
        mpdus.append(np.frombuffer(mp, np.uint8))
    offs = np.cumsum([0] + [len(v) for v in items]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "frames_sgi_true.npz"), iq=np.concatenate(items), offs=offs,
                        exp_mpdu=np.concatenate(mpdus), exp_len=np.array([m.size for m in mpdus], np.int32))
    print("frames_sgi_true:", [len(v) for v in items])


def frames_mu():
    """VHT sounding and MU-MIMO as one station antenna sees them (tools/cmu_ap.py:64-200 is the reference's recipe):
      * NDP: genFromAmpdu with an empty A-MPDU, nSTS = 2 -> two transmit streams; the station receives h0*tx0 + h1*tx1.
        demod reports the two VHT-LTFs ("mu2x1chan", lib/demod_impl.cc:238-249), decode publishes the channel report.
      * MU-MIMO: genAmpduMu for two users, group id 2, zero-forcing precoder Q = H^H (H H^H)^-1 for a flat 2x2 channel H
        (rows = users); user u receives sum_t H[u, t] * tx_t and decodes with demod(mupos = u, mugid = 2)."""
    phy = phy80211.phy80211(ifDebug=False)
    H = np.array([[1.0, 0.5 * np.exp(0.9j)], [0.6 * np.exp(-0.4j), 0.9 * np.exp(2.0j)]])
    items, meta, exp = [], [], []
    for cfo in (0.0, 40e3):
        mod = p8h.modulation(phyFormat=p8h.F.VHT, mcs=0, bw=p8h.BW.BW20, nSTS=2, shortGi=False)
        quiet(phy.genFromAmpdu, b"", mod, vhtPartialAid=0, vhtGroupId=0)
        ss = quiet(phy.genFinalSig, multiplier=12.0 * np.sqrt(2), cfoHz=cfo, num=1, gap=True, gapLen=400)
        ss = [np.asarray(x, np.complex128) for x in ss]
        items.append((H[0, 0] * ss[0] + H[0, 1] * ss[1]).astype(np.complex64)); meta.append((3, 0, cfo)); exp.append(b"")
    Q = H.conj().T @ np.linalg.inv(H @ H.conj().T)
    Q = Q / np.linalg.norm(Q) * np.sqrt(2)
    bfQ = [Q.copy() for _ in range(64)]
    pk = [mac_ampdu(["1234567 packet for station 000"]), mac_ampdu(["7654321 packet for station 111"])]
    for mcs0, mcs1 in ((0, 0), (4, 2)):
        quiet(phy.genAmpduMu, nUser=2, bfQ=bfQ, groupId=2, ampdu0=pk[0], mod0=p8h.modulation(p8h.F.VHT, mcs0, p8h.BW.BW20, 1, False),
              ampdu1=pk[1], mod1=p8h.modulation(p8h.F.VHT, mcs1, p8h.BW.BW20, 1, False))
        ss = quiet(phy.genFinalSig, multiplier=18.0, cfoHz=0.0, num=1, gap=True, gapLen=400)
        ss = [np.asarray(x, np.complex128) for x in ss]
        for u in range(2):
            items.append((H[u, 0] * ss[0] + H[u, 1] * ss[1]).astype(np.complex64)); meta.append((4, (mcs0, mcs1)[u], float(u)))
            exp.append(ampdu_split(pk[u])[0])
    offs = np.cumsum([0] + [len(x) for x in items]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "frames_mu.npz"), iq=np.concatenate(items).astype(np.complex64), offs=offs,
                        meta=np.array(meta, np.float64), exp_len=np.array([len(e) for e in exp], np.int32),
                        exp_mpdu=np.frombuffer(b"".join(exp), np.uint8), chan=H.astype(np.complex64))
    print("frames_mu: %d items, %d samples; meta = (3 NDP | 4 MU, mcs, cfo | user position)" % (len(items), offs[-1]))


def frames_mu_tx():
    """Two-user VHT MU-MIMO frames as the access point TRANSMITS them (genAmpduMu, tools/phy80211.py:180-221): both antenna
    waveforms, the A-MPDUs of the two users and the per-subcarrier spatial-mapping matrices -- for the MU side of the transmit
    synthesiser (c8b_tx_mu_batch).  Q differs from subcarrier to subcarrier so its indexing is pinned too."""
    phy = phy80211.phy80211(ifDebug=False)
    rng = np.random.default_rng(8021111)
    bfQ = []
    for k in range(64):
        a = rng.normal(0, 1, (2, 2)) + 1j * rng.normal(0, 1, (2, 2))
        a = a / np.linalg.norm(a) * np.sqrt(2)
        bfQ.append(a)
    pk = [mac_ampdu(["1234567 packet for station 000"]), mac_ampdu(["7654321 packet for station 111", "and a second subframe"])]
    iq0, iq1, meta = [], [], []
    for mcs0, mcs1, cfo in ((0, 0, 0.0), (4, 2, 0.0), (7, 8, 35e3), (1, 5, 0.0)):
        quiet(phy.genAmpduMu, nUser=2, bfQ=bfQ, groupId=2, ampdu0=pk[0], mod0=p8h.modulation(p8h.F.VHT, mcs0, p8h.BW.BW20, 1, False),
              ampdu1=pk[1], mod1=p8h.modulation(p8h.F.VHT, mcs1, p8h.BW.BW20, 1, False))
        ss = quiet(phy.genFinalSig, multiplier=18.0, cfoHz=cfo, num=1, gap=True, gapLen=400)
        iq0.append(np.asarray(ss[0], np.complex64)); iq1.append(np.asarray(ss[1], np.complex64)); meta.append((mcs0, mcs1, cfo))
    offs = np.cumsum([0] + [len(x) for x in iq0]).astype(np.int64)
    # the same two A-MPDUs through a zero-forcing precoder and the flat channel of frames_mu, as each station receives them:
    # what the reference's receive chain makes of user 1's TWO-subframe A-MPDU (tests/test_ref_chain.py)
    H = np.array([[1.0, 0.5 * np.exp(0.9j)], [0.6 * np.exp(-0.4j), 0.9 * np.exp(2.0j)]])
    Qz = H.conj().T @ np.linalg.inv(H @ H.conj().T)
    Qz = Qz / np.linalg.norm(Qz) * np.sqrt(2)
    quiet(phy.genAmpduMu, nUser=2, bfQ=[Qz.copy() for _ in range(64)], groupId=2, ampdu0=pk[0], mod0=p8h.modulation(p8h.F.VHT, 0, p8h.BW.BW20, 1, False),
          ampdu1=pk[1], mod1=p8h.modulation(p8h.F.VHT, 0, p8h.BW.BW20, 1, False))
    ss = [np.asarray(x, np.complex128) for x in quiet(phy.genFinalSig, multiplier=18.0, cfoHz=0.0, num=1, gap=True, gapLen=400)]
    zf = [(H[u, 0] * ss[0] + H[u, 1] * ss[1]).astype(np.complex64) for u in range(2)]
    np.savez_compressed(os.path.join(HERE, "frames_mu_tx.npz"), iq0=np.concatenate(iq0), iq1=np.concatenate(iq1), offs=offs,
                        meta=np.array(meta, np.float64), q=np.array(bfQ).astype(np.complex64),
                        ampdu0=np.frombuffer(bytes(pk[0]), np.uint8), ampdu1=np.frombuffer(bytes(pk[1]), np.uint8),
                        zf_rx0=zf[0], zf_rx1=zf[1])
    print("frames_mu_tx: %d frames, %d samples per antenna, A-MPDUs of %d / %d bytes" % (len(iq0), offs[-1], len(pk[0]), len(pk[1])))


def frames_tx():
    """The PSDUs (MPDU for legacy / HT, A-MPDU for VHT) behind the waveforms of frames_siso.npz and frames_bench.npz, in the
    same order, so the transmit synthesiser (c8b_tx_batch) can be compared with the generator's samples."""
    payload = "123456789012345678901234567890"
    mpdu, ampdu = bytes(mac_mpdu(payload)), bytes(mac_ampdu([payload]))
    ampdu2 = bytes(mac_ampdu([payload, "This is packet for station 001"]))
    ps = [mpdu] + [mpdu] * 8 + [mpdu] * 8 + [ampdu] * 9 + [mpdu, mpdu, ampdu, ampdu] + [ampdu2]
    gaps = [1200] + [400] * 30
    g = np.load(os.path.join(HERE, "frames_siso.npz"))
    assert len(ps) == len(g["offs"]) - 1
    rng = np.random.default_rng(80211)
    bench = []
    for i in range(16):
        pl = bytes(rng.integers(1, 128, 1434, dtype=np.uint8)).decode("latin-1")
        bench.append(bytes(mac_ampdu([pl])))
    np.savez_compressed(os.path.join(HERE, "frames_tx.npz"), psdu=np.frombuffer(b"".join(ps), np.uint8),
                        psdu_len=np.array([len(p) for p in ps], np.int32), gap=np.array(gaps, np.int32),
                        bench_psdu=np.frombuffer(b"".join(bench), np.uint8), bench_len=np.array([len(b) for b in bench], np.int32))
    print("frames_tx: %d PSDUs (%d bytes), %d bench A-MPDUs" % (len(ps), sum(len(p) for p in ps), len(bench)))


def ref_vectors():
    import oracle_lib as ol
    R = ol.ref()
    rng = np.random.default_rng(20211)
    out = {}
    # tables (cloud80211phy.h:151-195)
    for name, n in (("mapDeintLegacyBpsk", 48), ("mapDeintLegacyQpsk", 96), ("mapDeintLegacy16Qam", 192), ("mapDeintLegacy64Qam", 288),
                    ("mapDeintNonlegacyBpsk", 52), ("mapDeintNonlegacyQpsk", 104), ("mapDeintNonlegacy16Qam", 208),
                    ("mapDeintNonlegacy64Qam", 312), ("mapDeintNonlegacy256Qam", 416), ("mapDeintVhtSigB20", 52),
                    ("SV_STATE_NEXT", 128), ("SV_STATE_OUTPUT", 128)):
        buf = np.zeros(512, np.int32)
        assert R.ref_table_i(name.encode(), buf) == n
        out["tab_" + name] = buf[:n].copy()
    for name, n in (("LTF_L_26_F_FLOAT", 64), ("LTF_NL_28_F_FLOAT", 64), ("LTF_NL_28_F_FLOAT_VHT22", 64), ("PILOT_P", 127)):
        buf = np.zeros(128, np.float32)
        assert R.ref_table_f(name.encode(), buf) == n
        out["tab_" + name] = buf[:n].copy()
    # 2-stream deinterleave maps are file-static in the reference: recover them by probing procSymDeintNL2SS2
    for nb in (1, 2, 4, 6, 8):
        n = 52 * nb
        o = np.full(n, -1, np.float32)
        R.ref_deint_nl(np.arange(n, dtype=np.float32), o, n, 2)
        m = np.zeros(n, np.int32)
        m[o.astype(np.int32)] = np.arange(n, dtype=np.int32)    # out[map[i]] = i  ->  map[i] = position holding i
        out["tab_deintNL2_%d" % nb] = m
    # soft Viterbi on noisy codewords, rate 1/2 (SV_Decode_Sig == decode block's ACS, c8p.cc:1890-1999)
    for k, tl in enumerate((24, 48, 200, 1000, 4000)):
        bits = rng.integers(0, 2, tl).astype(np.uint8); bits[-6:] = 0
        coded = np.zeros(2 * tl, np.uint8); R.ref_bcc(bits, coded, tl)
        llr = ((coded.astype(np.float32) * 2 - 1) * 1.0 + rng.normal(0, 0.9, 2 * tl)).astype(np.float32)
        dec = np.zeros(tl, np.uint8); R.ref_sv_decode(llr, dec, tl)
        out["vit_llr_%d" % k] = llr; out["vit_bits_%d" % k] = dec
    # pure-noise input: exercises ties / survivor order, not just the easy path
    llr = rng.normal(0, 1.0, 2 * 3000).astype(np.float32)
    llr = np.round(llr * 4) / 4          # quantised -> many exact metric ties
    dec = np.zeros(3000, np.uint8); R.ref_sv_decode(llr.astype(np.float32), dec, 3000)
    out["vit_llr_tie"] = llr.astype(np.float32); out["vit_bits_tie"] = dec
    # LLR demap (c8p.cc:2090-2148)
    for mod, nsd in ((0, 48), (2, 48), (3, 48), (4, 48), (0, 52), (2, 52), (3, 52), (4, 52), (5, 52)):
        q = (rng.normal(0, 0.8, nsd) + 1j * rng.normal(0, 0.8, nsd)).astype(np.complex64)
        nb = {0: 1, 2: 2, 3: 4, 4: 6, 5: 8}[mod]
        o = np.zeros(nsd * nb, np.float32)
        R.ref_qam_to_llr(ol.c2f(q).copy(), o, mod, nsd)
        out["llr_in_%d_%d" % (mod, nsd)] = q; out["llr_out_%d_%d" % (mod, nsd)] = o
    # L-SIG / HT-SIG / VHT-SIG-A demod on random tones (c8p.cc:609-648)
    s1 = (rng.normal(0, 1, 64) + 1j * rng.normal(0, 1, 64)).astype(np.complex64)
    s2 = (s1 + 0.05 * (rng.normal(0, 1, 64) + 1j * rng.normal(0, 1, 64))).astype(np.complex64)
    sg = (rng.normal(0, 1, 64) + 1j * rng.normal(0, 1, 64)).astype(np.complex64)
    h = np.zeros(128, np.float32); l48 = np.zeros(48, np.float32)
    R.ref_lsig_demod(ol.c2f(s1), ol.c2f(s2), ol.c2f(sg), h, l48)
    out["lsig_s1"], out["lsig_s2"], out["lsig_sig"], out["lsig_h"], out["lsig_llr"] = s1, s2, sg, h.view(np.complex64).copy(), l48
    lht = np.zeros(96, np.float32); lvht = np.zeros(96, np.float32)
    hh = h.copy(); hh[hh == 0] = 1.0
    R.ref_nlsig_demod(ol.c2f(s1), ol.c2f(sg), hh, lht, lvht)
    out["nlsig_h"], out["nlsig_llrht"], out["nlsig_llrvht"] = hh.view(np.complex64).copy(), lht, lvht
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("ref_vectors: %d arrays" % len(out))


if __name__ == "__main__":
    which = sys.argv[1:] or ["siso", "bench", "ref", "mimo", "564", "sgi", "sgi_true", "mu", "mu_tx", "tx"]
    if "mu" in which:
        frames_mu()
    if "mu_tx" in which:
        frames_mu_tx()
    if "tx" in which:
        frames_tx()
    if "siso" in which:
        frames_siso()
    if "bench" in which:
        frames_bench()
    if "ref" in which:
        ref_vectors()
    if "mimo" in which:
        frames_mimo()
    if "564" in which:
        frames_564()
    if "sgi" in which:
        frames_sgi()
    if "sgi_true" in which:
        frames_sgi_true()
