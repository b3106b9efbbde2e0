"""GPU decode block (k_viterbi.cu through the C ABI) vs the oracle: bit-exact scrambled bits and PDU
records for every code rate, ties included (SURVEY 8d parity gate 1)."""
import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 2], ids=["warp_per_frame", "thread_per_frame"])
def rx(request):
    """every decode test runs against both kernels: k_viterbi (decode_mode 1) and k_viterbi_tp (decode_mode 2)"""
    pkg = load_pkg()
    r = pkg.Receiver(device=0, decode_mode=request.param)
    yield r
    r.close()


def _frames(pkg, n):
    return np.zeros(n, pkg.FRAME_DTYPE)


def _nllr(cr, T):
    per = {0: (2, 1), 1: (3, 2), 2: (4, 3), 3: (6, 5)}[cr]
    # soft bits consumed by T steps of the puncture pattern (lib/cloud80211phy.cc:1857-1860)
    pat = {0: [2], 1: [2, 1], 2: [2, 1, 1], 3: [2, 1, 1, 1, 1]}[cr]
    return sum(pat[t % len(pat)] for t in range(T))


@pytest.mark.parametrize("quant", [0, 2, 1])
def test_viterbi_bits_all_rates_random(rx, quant):
    """pure-noise soft bits (quantised -> many exact metric ties): scrambled bits must equal the oracle's"""
    pkg = load_pkg()
    rng = np.random.default_rng(100 + quant)
    lens = [1, 2, 5, 6, 7, 24, 29, 30, 31, 59, 60, 61, 149, 150, 151, 299, 300, 301, 777, 1000, 4534, 12220, 32782]
    metas, llrs, off = [], [], 0
    for cr in range(4):
        for T in lens:
            n = _nllr(cr, T)
            x = rng.normal(0, 2, n).astype(np.float32)
            if quant:
                x = (np.round(x * quant) / quant).astype(np.float32)
            metas.append((cr, T, n, off))
            llrs.append(x)
            off += n
    llr = np.concatenate(llrs)
    fr = _frames(pkg, len(metas))
    for i, (cr, T, n, o) in enumerate(metas):
        fr[i]["cr"], fr[i]["trellis"], fr[i]["total"], fr[i]["llr_off"] = cr, T, n, o
        fr[i]["format"], fr[i]["len"], fr[i]["ampdu"] = 0, 14, 1      # no PDU path: ampdu set on a non-VHT frame
    out, pdu, scram = rx.decode(llr, fr, pdu_stride=64, want_scram=True)
    O = ol.oracle()
    for i, (cr, T, n, o) in enumerate(metas):
        want = np.zeros(T, np.uint8)
        O.orx_viterbi(np.ascontiguousarray(llr[o:o + n]), cr, T, want)
        assert np.array_equal(scram[i, :T], want), (cr, T, int(np.argmax(scram[i, :T] != want)))
        assert out[i]["npdu"] == 0


def test_viterbi_vs_reference_golden(rx, golden):
    """soft bits / decoded bits produced by the UNMODIFIED reference (SV_Decode_Sig via oracle/_ref)"""
    pkg = load_pkg()
    g = golden["ref_vectors"]
    keys = list(range(5)) + ["tie"]
    llr = np.concatenate([g["vit_llr_%s" % k] for k in keys])
    fr = _frames(pkg, len(keys))
    o = 0
    for i, k in enumerate(keys):
        T = g["vit_bits_%s" % k].size
        fr[i]["cr"], fr[i]["trellis"], fr[i]["total"], fr[i]["llr_off"], fr[i]["ampdu"], fr[i]["len"] = 0, T, 2 * T, o, 1, 14
        o += 2 * T
    out, pdu, scram = rx.decode(llr, fr, pdu_stride=64, want_scram=True)
    for i, k in enumerate(keys):
        want = g["vit_bits_%s" % k]
        assert np.array_equal(scram[i, :want.size], want), k


def _oracle_frames(golden, noise=0.0, seed=1):
    g = golden["frames_siso"]
    iq, offs = g["iq"], g["offs"]
    rng = np.random.default_rng(seed)
    res = []
    for i in range(len(offs) - 1):
        x = iq[offs[i]:offs[i + 1]]
        if noise:
            x = (x + noise * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
        fr, llr, pdu = ol.rx_item(x, max_frames=1)
        res.append((fr[0], llr, pdu))
    return res


@pytest.mark.parametrize("noise", [0.0, 0.1875 / np.sqrt(2 * 10 ** 3.0), 0.1875 / np.sqrt(2 * 10 ** 1.2)])
def test_decode_pdus_from_oracle_llrs(rx, golden, noise):
    """identical LLR input (the oracle's demod output for the reference generator's frames, clean, 30 dB,
    12 dB) -> identical PDU records: descramble, A-MPDU walk (incl. the tmpLen quirk), CRC-32"""
    pkg = load_pkg()
    res = _oracle_frames(golden, noise)
    fr = _frames(pkg, len(res))
    llrs, o = [], 0
    for i, (f, llr, pdu) in enumerate(res):
        for k in ("status", "format", "mcs", "len", "cr", "ampdu", "trellis", "total"):
            fr[i][k] = f[k]
        fr[i]["llr_off"] = o
        llrs.append(llr)
        o += llr.size
    out, pdu, scram = rx.decode(np.concatenate(llrs), fr, pdu_stride=4400, want_scram=True)
    O = ol.oracle()
    nok = 0
    for i, (f, llr, opdu) in enumerate(res):
        if f["status"] != 0:
            assert out[i]["npdu"] == 0
            continue
        want = np.zeros(f["trellis"], np.uint8)
        O.orx_viterbi(np.ascontiguousarray(llr), int(f["cr"]), int(f["trellis"]), want)
        assert np.array_equal(scram[i, :f["trellis"]], want), i
        assert out[i]["npdu"] == f["npdu"] and out[i]["pdu_bytes"] == opdu.size, (i, out[i]["npdu"], f["npdu"])
        assert bytes(pdu[i, :opdu.size]) == bytes(opdu), i
        nok += int(f["npdu"] > 0)
    if noise < 0.01:
        assert nok == len(res)


def test_decode_range_and_bad_status(rx):
    pkg = load_pkg()
    fr = _frames(pkg, 3)
    llr = np.zeros(100, np.float32)
    fr[0]["trellis"], fr[0]["len"], fr[0]["total"] = 40000, 100, 100          # lib/decode_impl.cc:93-97
    fr[1]["trellis"], fr[1]["len"], fr[1]["total"] = 40, 5000, 80
    fr[2]["status"], fr[2]["trellis"], fr[2]["total"] = 3, 40, 80
    out, pdu, _ = rx.decode(llr, fr, pdu_stride=64)
    assert out[0]["status"] == 6 and out[1]["status"] == 6 and out[2]["status"] == 3
    assert out["npdu"].sum() == 0


def _encode_legacy_psdu(psdu, seed=93):
    """SERVICE(16 zero bits) + PSDU (LSB first) + 6 tail zeros -> scramble (x^7+x^4+1) -> BCC r=1/2 -> +-4 soft bits"""
    bits = np.concatenate([np.zeros(16, np.uint8), np.unpackbits(np.frombuffer(psdu, np.uint8), bitorder="little"), np.zeros(6, np.uint8)])
    st, out = seed, np.zeros(bits.size, np.uint8)
    for i, b in enumerate(bits):
        fb = ((st >> 6) ^ (st >> 3)) & 1
        out[i] = b ^ fb
        st = ((st << 1) & 0x7E) | fb
    out[-6:] = 0                                                   # tail bits are zeroed after scrambling
    reg, coded = 0, np.zeros(2 * out.size, np.float32)
    for i, b in enumerate(out):
        reg = ((reg << 1) & 0x7E) | int(b)
        coded[2 * i] = 4.0 if bin(reg & 0o155).count("1") & 1 else -4.0
        coded[2 * i + 1] = 4.0 if bin(reg & 0o117).count("1") & 1 else -4.0
    return coded


def test_crc32_and_assemble_all_lengths(rx):
    """legacy PSDUs of many lengths with a valid FCS must all be published, byte for byte (lane-parallel CRC-32:
    head segments of 1..64 bytes, 1..64 segments); a corrupted one must not"""
    import zlib
    pkg = load_pkg()
    rng = np.random.default_rng(4)
    lens = [14, 15, 16, 63, 64, 65, 66, 127, 128, 129, 191, 192, 193, 1000, 1500, 2047, 2048, 2049, 4031, 4032, 4033, 4095]
    llrs, metas, o = [], [], 0
    for k, n in enumerate(lens + [300]):
        body = bytes(rng.integers(0, 256, n - 4, dtype=np.uint8))
        psdu = body + zlib.crc32(body).to_bytes(4, "little")
        if k == len(lens):
            psdu = bytes([psdu[0] ^ 1]) + psdu[1:]                 # broken FCS
        c = _encode_legacy_psdu(psdu, seed=1 + 5 * k)
        llrs.append(c); metas.append((n, psdu, o)); o += c.size
    fr = _frames(pkg, len(metas))
    for i, (n, psdu, off) in enumerate(metas):
        fr[i]["format"], fr[i]["mcs"], fr[i]["len"], fr[i]["cr"], fr[i]["ampdu"] = 0, 3, n, 0, 0
        fr[i]["trellis"], fr[i]["total"], fr[i]["llr_off"] = 8 * n + 22, 2 * (8 * n + 22), off
    llr = np.concatenate(llrs)
    out, pdu, _ = rx.decode(llr, fr, pdu_stride=4400)
    O = ol.oracle()
    import ctypes as C
    for i, (n, psdu, off) in enumerate(metas):
        of = np.zeros(1, ol.FRAME_DTYPE)
        for k in ("format", "mcs", "len", "cr", "ampdu", "trellis", "total"):
            of[0][k] = fr[i][k]
        buf, used = np.zeros(4400, np.uint8), C.c_int(0)
        np_o = O.orx_decode(np.ascontiguousarray(llr[off:off + fr[i]["total"]]), of.ctypes.data_as(C.POINTER(ol.OrxFrame)), buf, 4400, C.byref(used), None)
        assert out[i]["npdu"] == np_o and out[i]["pdu_bytes"] == used.value, (n, out[i]["npdu"], np_o)
        assert bytes(pdu[i, :used.value]) == bytes(buf[:used.value]), n
        if i < len(lens):
            assert out[i]["npdu"] == 1 and bytes(pdu[i, :n + 4]) == bytes([0, n & 255, n >> 8]) + psdu + bytes([3]), n
        else:
            assert out[i]["npdu"] == 0
