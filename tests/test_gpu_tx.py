"""Transmit synthesiser (c8b_tx_batch) against the reference's own generator: the waveforms of
tools/phy80211.py (genFromMpdu / genFromAmpdu + genFinalSig) stored in tests/golden/frames_siso.npz and frames_bench.npz,
regenerated on the GPU from the PSDU bytes (tests/golden/frames_tx.npz), must agree sample for sample; and the loop
TX -> RX must return the MPDUs."""
import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu
HERE = __import__("os").path.dirname(__import__("os").path.abspath(__file__))


def _tx_golden():
    return np.load(HERE + "/golden/frames_tx.npz")


def _psdus(t, key="psdu", lkey="psdu_len"):
    o = np.cumsum(np.r_[0, t[lkey]])
    return [bytes(t[key][o[i]:o[i + 1]]) for i in range(len(t[lkey]))]


def test_waveforms_equal_the_generator(golden):
    pkg = load_pkg()
    g, t = golden["frames_siso"], _tx_golden()
    ps = _psdus(t)
    fmt = g["meta"][:, 0].astype(np.int32)
    mcs = g["meta"][:, 1].astype(np.int32)
    cfo = g["meta"][:, 2].astype(np.float32)
    rx = pkg.Receiver(device=0)
    iq, offs = rx.tx_batch(ps, fmt, mcs, gap=t["gap"], cfo=cfo, multiplier=12.0, seed=93)
    rx.close()
    assert np.array_equal(offs, g["offs"])
    ref = g["iq"]
    peak = float(np.max(np.abs(ref)))
    for i in range(len(ps)):
        a, b = iq[offs[i]:offs[i + 1]], ref[offs[i]:offs[i + 1]]
        err = float(np.max(np.abs(a - b)))
        # float32 synthesis vs the generator's float64 rounded to float32 (CFO frames: phase of up to 7000 samples in float32)
        assert err <= (2e-6 if cfo[i] == 0 else 2e-5) * peak, (i, fmt[i], mcs[i], cfo[i], err)
        assert np.all(a[:t["gap"][i]] == 0) and np.all(a[-t["gap"][i]:] == 0)


def test_bench_frames_and_loopback(golden):
    """config-5 frames (VHT MCS7, 1504-byte A-MPDU): same samples as the generator's, and TX -> RX returns every MPDU"""
    pkg = load_pkg()
    gb, t = golden["frames_bench"], _tx_golden()
    ps = _psdus(t, "bench_psdu", "bench_len")
    rx = pkg.Receiver(device=0)
    iq, offs = rx.tx_batch(ps, 2, 7, gap=0)
    assert offs[1] == 4560
    got = iq.reshape(16, 4560)
    assert float(np.max(np.abs(got - gb["iq"]))) <= 2e-6 * float(np.max(np.abs(gb["iq"])))
    # loop back with gaps and noise
    rng = np.random.default_rng(3)
    iq, offs = rx.tx_batch(ps, 2, 7, gap=400, cfo=rng.uniform(-1e5, 1e5, 16))
    s = 0.1875 / np.sqrt(2 * 10 ** 3.0)
    x = (iq + s * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))).astype(np.complex64)
    fr, pdu = rx.rx_batch(x, offs[:-1], np.diff(offs).astype(np.int32))
    rx.close()
    for i in range(16):
        assert fr[i]["status"] == 0 and fr[i]["npdu"] == 1 and (fr[i]["format"], fr[i]["mcs"], fr[i]["len"]) == (2, 7, 1504)
        assert bytes(pdu[i, 3:1503]) == bytes(gb["mpdu"][i]) == ps[i][4:1504]


@pytest.mark.parametrize("fmt,mcss", [(0, range(8)), (1, range(8)), (2, range(9))])
def test_loopback_every_rate_and_length(fmt, mcss):
    """random PSDU lengths at every MCS: what the synthesiser emits, the receiver (and the oracle) decodes"""
    pkg = load_pkg()
    rng = np.random.default_rng(100 + fmt)
    rx = pkg.Receiver(device=0)
    ps, ms = [], []
    for mcs in mcss:
        for ln in (rng.integers(30, 200), rng.integers(200, 1600), 4 * rng.integers(400, 900)):
            body = bytes(rng.integers(0, 256, int(ln), dtype=np.uint8))
            crc = __import__("zlib").crc32(body) & 0xffffffff
            mpdu = body + crc.to_bytes(4, "little")
            if fmt == 2:                                          # one-MPDU A-MPDU: delimiter (tools/mac80211.py:333-360) + pad to 4
                n = len(mpdu)
                d0 = ((n & 0xf) << 4) | (((n >> 12) & 3) << 2) | 1   # EOF = 1 (single MPDU), reserved, len[12:14], len[0:4]
                d1 = (n >> 4) & 0xff
                bits = [(d0 >> k) & 1 for k in range(8)] + [(d1 >> k) & 1 for k in range(8)]
                c = [1] * 8
                for b in bits:
                    f = b ^ c[7]
                    c = [f, f ^ c[0], f ^ c[1], c[2], c[3], c[4], c[5], c[6]]
                crc8 = sum((1 - c[7 - k]) << k for k in range(8))
                mpdu_padded = mpdu + bytes((-n) % 4)
                ps.append(bytes([d0, d1, crc8, 0x4E]) + mpdu_padded)
            else:
                ps.append(mpdu)
            ms.append(mcs)
    iq, offs = rx.tx_batch(ps, fmt, ms, gap=300)
    s = 0.1875 / np.sqrt(2 * 10 ** 3.5)
    x = (iq + s * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))).astype(np.complex64)
    fr, pdu = rx.rx_batch(x, offs[:-1], np.diff(offs).astype(np.int32))
    rx.close()
    for i, p in enumerate(ps):
        want = p[4:4 + (len(p) - 4)] if fmt == 2 else p
        assert fr[i]["status"] == 0 and fr[i]["npdu"] == 1 and fr[i]["format"] == fmt and fr[i]["mcs"] == ms[i], (i, ms[i], len(p), fr[i]["status"], fr[i]["npdu"])
        nb = int(fr[i]["pdu_bytes"])
        got = bytes(pdu[i, 3:nb - 1])
        assert got == (want[:len(got)] if fmt == 2 else want), (i, ms[i])
        fo, _, po = ol.rx_item(x[offs[i]:offs[i + 1]], max_frames=1)
        assert bytes(po) == bytes(pdu[i, :nb])


def test_tx_argument_errors():
    pkg = load_pkg()
    rx = pkg.Receiver(device=0)
    assert rx.L.c8b_tx_nsamp(0, 0, 100) == (5 + 35) * 80 and rx.L.c8b_tx_nsamp(2, 7, 1504) == 4560
    assert rx.L.c8b_tx_nsamp(2, 9, 100) < 0 and rx.L.c8b_tx_nsamp(0, 8, 100) < 0 and rx.L.c8b_tx_nsamp(1, 0, 0) < 0 and rx.L.c8b_tx_nsamp(0, 0, 4096) < 0
    with pytest.raises(ValueError):
        rx.tx_batch([b"x" * 50], 3, 0)
    with pytest.raises(RuntimeError):
        rx.tx_batch([b"x" * 50], 0, 0, seed=0)
    rx.close()


def test_device_traffic_closed_loop():
    """c8b_tx_random_psdu_dev + c8b_tx_batch_dev + c8b_rx_batch_dev: 3000 frames with their own random MPDUs (all three formats,
    every MCS, mixed lengths) never leave the device until the decoded bytes are compared with what was sent"""
    import torch
    pkg = load_pkg()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    n = 3000
    d = np.zeros(n, pkg.TXFRAME_DTYPE)
    d["format"] = rng.integers(0, 3, n)
    d["mcs"] = np.where(d["format"] == 2, rng.integers(0, 9, n), rng.integers(0, 8, n))
    d["psdu_len"] = 4 * rng.integers(10, 380, n)
    d["psdu_off"] = np.concatenate([[0], np.cumsum(d["psdu_len"])[:-1]])
    d["cfo_hz"] = rng.uniform(-8e4, 8e4, n).astype(np.float32)
    rx = pkg.Receiver(device=0)
    ns = np.array([rx.L.c8b_tx_nsamp(int(f), int(m), int(l)) for f, m, l in zip(d["format"], d["mcs"], d["psdu_len"])], np.int64)
    item = ns + 600
    off = np.concatenate([[0], np.cumsum(item)[:-1]]).astype(np.int64)
    d["out_off"] = off + 300
    nb, nsamp = int(d["psdu_len"].sum()), int(item.sum())
    psdu = torch.zeros(nb + 16, dtype=torch.uint8, device=dev)
    iq = torch.zeros(nsamp, dtype=torch.complex64, device=dev)
    torch.cuda.synchronize()
    rx.tx_random_psdu_dev(psdu.data_ptr(), nb, d, seed=7)
    rx.tx_batch_dev(psdu.data_ptr(), nb, d, iq.data_ptr(), nsamp)
    rx.sync()
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    iq += torch.view_as_complex(torch.randn((nsamp, 2), generator=g, device=dev) * (0.1875 / np.sqrt(2 * 10 ** 3.3)))
    fr, pdu = rx.rx_batch_dev(iq.data_ptr(), off, item.astype(np.int32))
    rx.close()
    sent = psdu.cpu().numpy()
    x = iq.cpu().numpy()
    lost = []
    for i in range(n):
        f = fr[i]
        vht = d["format"][i] == 2
        want = bytes(sent[d["psdu_off"][i] + (4 if vht else 0): d["psdu_off"][i] + d["psdu_len"][i]])
        assert __import__("zlib").crc32(want[:-4]) & 0xffffffff == int.from_bytes(want[-4:], "little")
        if f["status"] == 0 and f["npdu"] == 1:
            assert f["format"] == d["format"][i] and f["mcs"] == d["mcs"][i] and bytes(pdu[i, 3:f["pdu_bytes"] - 1]) == want, i
        else:                                                      # lost to the noise: the oracle must lose it the same way
            fo, _, po = ol.rx_item(x[off[i]:off[i] + item[i]], max_frames=1)
            assert (f["status"], f["npdu"]) == (fo[0]["status"], fo[0]["npdu"]), (i, d[i], f["status"], fo[0]["status"])
            lost.append((i, int(f["status"])))
    assert len(lost) <= n // 200, lost


def test_udp_in_udp_out_loop():
    """both directions of the UDP PDU framing: MAC -> PHY datagrams [format][mcs][nss][len16][PSDU]
    (lib/pktgen_impl.cc:57-70, tools/phy80211.py genPktGrData) through c8b_tx_from_udp, the waveform through the receive
    chain, and PHY -> MAC records [format][len16][MPDU][mcs] (lib/decode_impl.cc:512-516) carrying the same bytes"""
    import struct
    import zlib
    pkg = load_pkg()
    rng = np.random.default_rng(11)

    def mpdu(n):
        b = bytes(rng.integers(0, 256, n - 4, dtype=np.uint8))
        return b + struct.pack("<I", zlib.crc32(b))

    def ampdu(m):                                                # tools/mac80211.py:333-360, one subframe, EOF set
        ln = len(m)
        hdr = bytearray(4)
        hdr[0] = 1 | ((ln >> 12 & 1) << 2) | ((ln >> 13 & 1) << 3) | ((ln & 0xF) << 4)
        hdr[1] = (ln >> 4) & 0xFF
        c = 0xFF
        for bit in range(16):
            b = (hdr[bit // 8] >> (bit % 8)) & 1
            top = (c >> 7) & 1
            c = (c << 1) & 0xFF
            if top ^ b:
                c ^= 0x07
        c ^= 0xFF
        hdr[2] = int("{:08b}".format(c)[::-1], 2)
        hdr[3] = 0x4E
        out = bytes(hdr) + m
        return out + bytes((-len(out)) % 4)

    sent, dgrams = [], []
    for fmt, mcs, n in [(0, 0, 60), (0, 7, 300), (1, 3, 500), (1, 7, 1200), (2, 0, 100), (2, 5, 700), (2, 8, 1500)]:
        m = mpdu(n)
        psdu = ampdu(m) if fmt == 2 else m
        sent.append((fmt, mcs, m))
        dgrams.append(struct.pack("<BBBH", fmt, mcs, 1, len(psdu)) + psdu)
    dgrams.insert(3, b"\x01\x02")                                # junk datagrams are skipped, like pktgen does
    dgrams.insert(5, struct.pack("<BBBH", 3, 0, 1, 4) + bytes(4))
    rx = pkg.Receiver(device=0, max_frames=16)
    iq, desc = rx.tx_from_udp(dgrams, gap=500)
    assert [int(d["psdu_len"]) for d in desc].count(-1) == 2 and iq.size > 0
    fr, pdu = rx.rx_batch(iq, [0], [iq.size])
    rx.close()
    recs = []
    for k in range(16):
        if fr[k]["npdu"] > 0:
            recs += pkg.blocks.split_messages(bytes(pdu[k, :fr[k]["pdu_bytes"]]))
    assert len(recs) == len(sent)
    for r, (fmt, mcs, m) in zip(recs, sent):
        assert r == bytes([fmt]) + struct.pack("<H", len(m)) + m + bytes([mcs])


def test_two_stream_waveforms_equal_the_generator(golden):
    """c8b_tx_batch2 against tools/phy80211.py with nSTS = 2 (the recipe of tools/pktGenExample.py:206-217; golden
    frames_mimo.npz: HT MCS8-15, VHT 2SS MCS0-8, one CFO case): both antennas sample for sample"""
    pkg = load_pkg()
    g, t = golden["frames_mimo"], _tx_golden()
    ps = _psdus(t)
    mpdu, ampdu = ps[0], ps[17]                                  # the MPDU / A-MPDU every frame of frames_mimo carries
    fmt = g["meta"][:, 0].astype(np.int32)
    mcs = g["meta"][:, 1].astype(np.int32)
    cfo = g["meta"][:, 2].astype(np.float32)
    code = np.where(fmt == 2, mcs + 16, mcs)
    rx = pkg.Receiver(device=0)
    iq0, iq1, offs = rx.tx_batch2([ampdu if f == 2 else mpdu for f in fmt], fmt, code, gap=400, cfo=cfo)
    rx.close()
    assert np.array_equal(offs, g["offs"])
    peak = float(np.max(np.abs(g["iq0"])))
    for i in range(len(fmt)):
        for a, (got, ref) in enumerate(((iq0, g["iq0"]), (iq1, g["iq1"]))):
            err = float(np.max(np.abs(got[offs[i]:offs[i + 1]] - ref[offs[i]:offs[i + 1]])))
            assert err <= (2e-6 if cfo[i] == 0 else 2e-5) * peak, (i, a, fmt[i], mcs[i], err / peak)


@pytest.mark.parametrize("fmt,codes", [(1, range(8, 16)), (2, range(16, 25))])
def test_two_stream_loopback_random_traffic(fmt, codes):
    """unique two-stream frames (random lengths) through c8b_tx_batch2 -> c8b_rx_batch2: every MPDU comes back; the oracle's
    2x2 chain agrees"""
    pkg = load_pkg()
    rng = np.random.default_rng(300 + fmt)
    rx = pkg.Receiver(device=0)
    ps, ms, want = [], [], []
    for c in codes:
        for ln in (int(rng.integers(40, 300)), 4 * int(rng.integers(150, 380))):
            body = bytes(rng.integers(0, 256, ln - 4, dtype=np.uint8))
            m = body + (__import__("zlib").crc32(body) & 0xffffffff).to_bytes(4, "little")
            if fmt == 2:
                n = len(m)
                d0 = ((n & 0xf) << 4) | (((n >> 12) & 3) << 2) | 1
                ps.append(bytes([d0, (n >> 4) & 0xff, 0, 0x4E]) + m + bytes((-n) % 4))
            else:
                ps.append(m)
            ms.append(c)
            want.append(m)
    a, b, offs = rx.tx_batch2(ps, fmt, ms, gap=300)
    s = 0.1875 * np.sqrt(2) / np.sqrt(2 * 10 ** 3.2)
    a = (a + s * (rng.standard_normal(a.size) + 1j * rng.standard_normal(a.size))).astype(np.complex64)
    b = (b + s * (rng.standard_normal(b.size) + 1j * rng.standard_normal(b.size))).astype(np.complex64)
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    fr, pdu = rx.rx_batch2(a, b, off, ln)
    rx.close()
    for i, m in enumerate(want):
        assert fr[i]["status"] == 0 and fr[i]["nss"] == 2 and fr[i]["npdu"] == 1, (i, ms[i], fr[i]["status"], fr[i]["npdu"])
        assert bytes(pdu[i, 3:3 + len(m)]) == m, (i, ms[i])
        if i % 5 == 0:
            _, _, po = ol.rx_item2(a[offs[i]:offs[i + 1]], b[offs[i]:offs[i + 1]], max_frames=1)
            assert bytes(po) == bytes(pdu[i, :po.size])


def _mu_golden():
    return np.load(HERE + "/golden/frames_mu_tx.npz")


def test_mu_waveforms_equal_the_generator():
    """c8b_tx_mu_batch against tools/phy80211.py genAmpduMu + genFinalSig (golden frames_mu_tx.npz: two users, MCS pairs
    (0,0) (4,2) (7,8) (1,5), different A-MPDU lengths, a spatial mapping matrix that differs on every subcarrier, one CFO
    case): both antennas sample for sample"""
    pkg = load_pkg()
    g = _mu_golden()
    a0, a1 = bytes(g["ampdu0"]), bytes(g["ampdu1"])
    mcs = [(int(m[0]), int(m[1])) for m in g["meta"]]
    cfo = g["meta"][:, 2].astype(np.float32)
    rx = pkg.Receiver(device=0)
    iq0, iq1, offs = rx.tx_mu_batch([(a0, a1)] * len(mcs), mcs, g["q"], group_id=2, gap=400, cfo=cfo, multiplier=18.0)
    rx.close()
    assert np.array_equal(offs, g["offs"])
    peak = float(np.max(np.abs(g["iq0"])))
    for i in range(len(mcs)):
        for a, (got, ref) in enumerate(((iq0, g["iq0"]), (iq1, g["iq1"]))):
            err = float(np.max(np.abs(got[offs[i]:offs[i + 1]] - ref[offs[i]:offs[i + 1]])))
            assert err <= (2e-6 if cfo[i] == 0 else 2e-5) * peak, (i, a, mcs[i], err / peak)


def test_mu_loopback_each_station_gets_its_frame():
    """access point -> flat 2x2 channel -> the two stations (tools/cmu_ap.py's recipe): frames precoded with the zero-forcing
    Q = H^H (H H^H)^-1 reach station u as its own stream; demod(mupos = u, mugid) + decode return user u's first MPDU, and
    exactly the PDU bytes of the oracle's chain on the same samples.  (Of a two-subframe A-MPDU the reference's MU receive
    path publishes the first subframe only -- its own blocks and the oracle agree on that, tests/test_ref_chain.py.)"""
    pkg = load_pkg()
    g = _mu_golden()
    a = [bytes(g["ampdu0"]), bytes(g["ampdu1"])]
    H = np.array([[1.0, 0.5 * np.exp(0.9j)], [0.6 * np.exp(-0.4j), 0.9 * np.exp(2.0j)]])
    Q = H.conj().T @ np.linalg.inv(H @ H.conj().T)
    Q = Q / np.linalg.norm(Q) * np.sqrt(2)
    q = np.broadcast_to(Q.astype(np.complex64), (64, 2, 2)).copy()
    pairs = [(0, 0), (3, 1), (5, 7), (8, 4), (2, 6)]
    tx = pkg.Receiver(device=0)
    iq0, iq1, offs = tx.tx_mu_batch([(a[0], a[1])] * len(pairs), pairs, q, group_id=2, gap=400, multiplier=18.0)
    tx.close()
    rng = np.random.default_rng(77)
    O = ol.oracle()
    try:
        for u in range(2):
            y = (H[u, 0] * iq0.astype(np.complex128) + H[u, 1] * iq1.astype(np.complex128))
            y = (y + 1e-3 * (rng.standard_normal(y.size) + 1j * rng.standard_normal(y.size))).astype(np.complex64)
            rx = pkg.Receiver(device=0, mupos=u, mugid=2)
            fr, pdu = rx.rx_batch(y, offs[:-1], np.diff(offs).astype(np.int32), pdu_stride=4400)
            rx.close()
            p = a[u]
            n0 = (p[0] >> 4) | (p[1] << 4) | (((p[0] >> 2) & 3) << 12)      # first subframe of the station's A-MPDU
            first = p[4:4 + n0]
            O.orx_set_mupos(u)
            for i in range(len(pairs)):
                fo, _, po = ol.rx_item(y[offs[i]:offs[i + 1]], max_frames=1)
                assert fr[i]["status"] == 0 and fr[i]["format"] == 2 and fr[i]["mcs"] == pairs[i][u], (u, i, fr[i]["status"], fr[i]["mcs"])
                nb = int(fr[i]["pdu_bytes"])
                assert nb == fo[0]["pdu_bytes"] and bytes(pdu[i, :nb]) == bytes(po[:nb]), (u, i)
                recs = pkg.rx.split_pdus(pdu[i, :nb])
                assert len(recs) >= 1 and recs[0][3:-1] == first, (u, i, len(recs))
    finally:
        O.orx_set_mupos(0)


def _one_mpdu_ampdu(mpdu):
    """delimiter (tools/mac80211.py:333-360: EOF, reserved, len[12:14], len[0:12], CRC-8, 0x4E) + MPDU + pad to 4"""
    n = len(mpdu)
    d0 = ((n & 0xf) << 4) | (((n >> 12) & 3) << 2) | 1
    d1 = (n >> 4) & 0xff
    bits = [(d0 >> k) & 1 for k in range(8)] + [(d1 >> k) & 1 for k in range(8)]
    c = [1] * 8
    for b in bits:
        f = b ^ c[7]
        c = [f, f ^ c[0], f ^ c[1], c[2], c[3], c[4], c[5], c[6]]
    crc8 = sum((1 - c[7 - k]) << k for k in range(8))
    return bytes([d0, d1, crc8, 0x4E]) + mpdu + bytes((-n) % 4)


@pytest.mark.parametrize("fmt,mcs,ln", [(0, 0, 4095), (0, 7, 4095), (1, 0, 4095), (1, 7, 4095), (2, 0, 4088), (2, 8, 4088), (0, 0, 4094), (1, 3, 4093)])
def test_maximum_length_frames(fmt, mcs, ln):
    """the longest frames the formats carry (4095-byte PSDUs: 1366 symbols at the lowest legacy rate, 109 600 samples):
    synthesised, received, and compared with the oracle record for record -- the LLR stride, the symbol grid of the demod
    kernels and the decode scratch at their upper ends, and the reference's behaviour at its 4095-byte limit"""
    pkg = load_pkg()
    rng = np.random.default_rng(ln + 10 * mcs + fmt)
    body = bytes(rng.integers(0, 256, ln - 4, dtype=np.uint8))
    mpdu = body + (__import__("zlib").crc32(body) & 0xffffffff).to_bytes(4, "little")
    psdu = _one_mpdu_ampdu(mpdu) if fmt == 2 else mpdu
    rx = pkg.Receiver(device=0)
    iq, offs = rx.tx_batch([psdu], fmt, [mcs], gap=300)
    s = 0.1875 / np.sqrt(2 * 10 ** 3.5)
    x = (iq + s * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))).astype(np.complex64)
    fr, pdu = rx.rx_batch(x, offs[:-1], np.diff(offs).astype(np.int32), pdu_stride=4400)
    rx.close()
    fo, _, po = ol.rx_item(x, max_frames=1)
    for k in ("status", "format", "mcs", "len", "nsym", "total", "npdu", "pdu_bytes"):
        assert fr[0][k] == fo[0][k], (k, fr[0][k], fo[0][k])
    nb = int(fr[0]["pdu_bytes"])
    assert bytes(pdu[0, :nb]) == bytes(po[:nb])
    if fr[0]["npdu"] == 1:
        assert bytes(pdu[0, 3:nb - 1]) == mpdu
    assert fr[0]["status"] == 0 and fr[0]["format"] == fmt and fr[0]["mcs"] == mcs


@pytest.mark.parametrize("fmt,code,ln", [(1, 8, 4095), (1, 15, 4095), (2, 16, 4088), (2, 24, 4088)])
def test_maximum_length_two_stream_frames(fmt, code, ln):
    """the same at two streams (HT MCS8 / 15, VHT 2SS MCS0 / 8) through c8b_tx_batch2 -> c8b_rx_batch2 and the oracle's 2x2 chain"""
    pkg = load_pkg()
    rng = np.random.default_rng(ln + code)
    body = bytes(rng.integers(0, 256, ln - 4, dtype=np.uint8))
    mpdu = body + (__import__("zlib").crc32(body) & 0xffffffff).to_bytes(4, "little")
    psdu = _one_mpdu_ampdu(mpdu) if fmt == 2 else mpdu
    rx = pkg.Receiver(device=0)
    iq0, iq1, offs = rx.tx_batch2([psdu], fmt, [code], gap=300)
    s = 0.1875 / np.sqrt(2 * 10 ** 3.5)
    x0 = (iq0 + s * (rng.standard_normal(iq0.size) + 1j * rng.standard_normal(iq0.size))).astype(np.complex64)
    x1 = (iq1 + s * (rng.standard_normal(iq1.size) + 1j * rng.standard_normal(iq1.size))).astype(np.complex64)
    fr, pdu = rx.rx_batch2(x0, x1, offs[:-1], np.diff(offs).astype(np.int32), pdu_stride=4400)
    rx.close()
    fo, _, po = ol.rx_item2(x0, x1, max_frames=1)
    for k in ("status", "format", "mcs", "len", "nss", "nsym", "total", "npdu", "pdu_bytes"):
        assert fr[0][k] == fo[0][k], (k, fr[0][k], fo[0][k])
    nb = int(fr[0]["pdu_bytes"])
    assert bytes(pdu[0, :nb]) == bytes(po[:nb])
    if (fmt, code) == (2, 24):
        # 53 symbols x 624 data bits = 33072 trellis steps: over the decode block's 32782 limit (lib/decode_impl.cc:93-97), so the
        # reference drops the frame after the header -- and so do the oracle and the GPU
        assert fr[0]["status"] == 6 and fr[0]["npdu"] == 0
    else:
        assert fr[0]["status"] == 0 and fr[0]["nss"] == 2 and fr[0]["npdu"] == 1 and bytes(pdu[0, 3:nb - 1]) == mpdu


@pytest.mark.parametrize("cfo", [-232e3, -150e3, 232e3])
def test_loopback_at_the_carrier_offset_limit(cfo):
    """+-232 kHz (40 ppm between two 5.8 GHz radios, the standard's worst case): sync's estimate, the rotation fused into the
    demod kernels and the oracle's agree frame for frame at every format"""
    pkg = load_pkg()
    rng = np.random.default_rng(int(abs(cfo)) + (cfo > 0))
    ps, fm, ms = [], [], []
    for fmt, mcs in ((0, 0), (0, 4), (0, 7), (1, 2), (1, 7), (2, 0), (2, 5), (2, 8)):
        body = bytes(rng.integers(0, 256, int(rng.integers(60, 900)), dtype=np.uint8))
        mpdu = body + (__import__("zlib").crc32(body) & 0xffffffff).to_bytes(4, "little")
        ps.append(_one_mpdu_ampdu(mpdu) if fmt == 2 else mpdu)
        fm.append(fmt)
        ms.append(mcs)
    rx = pkg.Receiver(device=0)
    iq, offs = rx.tx_batch(ps, fm, ms, gap=300, cfo=[cfo] * len(ps))
    s = 0.1875 / np.sqrt(2 * 10 ** 3.2)
    x = (iq + s * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))).astype(np.complex64)
    fr, pdu = rx.rx_batch(x, offs[:-1], np.diff(offs).astype(np.int32))
    rx.close()
    for i in range(len(ps)):
        fo, _, po = ol.rx_item(x[offs[i]:offs[i + 1]], max_frames=1)
        for k in ("status", "sync_idx", "format", "mcs", "len", "npdu", "pdu_bytes"):
            assert fr[i][k] == fo[0][k], (i, k, fr[i][k], fo[0][k])
        assert abs(fr[i]["cfo_hz"] - fo[0]["cfo_hz"]) <= 1e-6 * 20e6 / (2 * np.pi) + 1e-3 * abs(fo[0]["cfo_hz"]) * 1e-3
        nb = int(fr[i]["pdu_bytes"])
        assert bytes(pdu[i, :nb]) == bytes(po[:nb]), i
        assert fr[i]["status"] == 0 and fr[i]["npdu"] == 1 and abs(fr[i]["cfo_hz"] + cfo) < 3e3, (i, fr[i]["cfo_hz"])
