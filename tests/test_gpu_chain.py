"""Whole chain on the GPU (c8b_rx_batch, host buffers in, PDUs out) vs the oracle: the set of published
PDUs must be identical byte for byte and in order (SURVEY 8d gate 4)."""
import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


def _run_both(pkg, iq, off, ln, **kw):
    rx = pkg.Receiver(device=0, **kw)
    fr, pdu = rx.rx_batch(iq, off, ln)
    rx.close()
    n = len(off)
    fo = np.zeros(n, ol.FRAME_DTYPE)
    stride = 4400
    po = np.zeros(n * stride, np.uint8)
    ol.oracle().orx_rx_batch(ol.c2f(iq), np.ascontiguousarray(off, np.int64), np.ascontiguousarray(ln, np.int32), n, 8, fo.ctypes.data, po, stride)
    return fr, pdu, fo, po.reshape(n, stride)


def _compare(fr, pdu, fo, po):
    for i in range(fr.size):
        assert fr[i]["status"] == fo[i]["status"], (i, fr[i]["status"], fo[i]["status"])
        assert fr[i]["npdu"] == fo[i]["npdu"] and fr[i]["pdu_bytes"] == fo[i]["pdu_bytes"], (i, fr[i]["npdu"], fo[i]["npdu"])
        nb = int(fo[i]["pdu_bytes"])
        assert bytes(pdu[i, :nb]) == bytes(po[i, :nb]), i


def test_config1_plumbing(golden):
    """BASELINE config 1: tools/pktGenExample.py Legacy MCS0 frame -> PDU [0][len][MPDU][0]"""
    pkg = load_pkg()
    g = golden["frames_siso"]
    x = g["iq"][g["offs"][0]:g["offs"][1]]
    rx = pkg.Receiver(device=0)
    fr, pdu = rx.rx_batch(x, [0], [x.size])
    rx.close()
    mpdu = bytes(g["exp_mpdu"][:g["exp_len"][0]])
    assert fr[0]["status"] == 0 and fr[0]["npdu"] == 1 and (fr[0]["format"], fr[0]["mcs"], fr[0]["len"]) == (0, 0, 94)
    assert bytes(pdu[0, :fr[0]["pdu_bytes"]]) == bytes([0, 94, 0]) + mpdu + bytes([0])
    assert mpdu[:10].hex() == "08016e00f469d5800fa0"


@pytest.mark.parametrize("snr", [None, 30.0, 12.0])
@pytest.mark.parametrize("chunk", [16384, 7])
def test_all_formats_vs_oracle(golden, snr, chunk):
    pkg = load_pkg()
    g = golden["frames_siso"]
    iq = g["iq"].copy()
    if snr is not None:
        rng = np.random.default_rng(23)
        s = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
        iq = (iq + s * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))).astype(np.complex64)
    offs = g["offs"]
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    fr, pdu, fo, po = _run_both(pkg, iq, off, ln, chunk_items=chunk)
    _compare(fr, pdu, fo, po)
    if snr is None or snr >= 30:
        el = g["exp_len"]
        eo = np.cumsum(np.r_[0, el])
        for i in range(len(off)):
            assert fr[i]["npdu"] >= 1 and bytes(pdu[i, 3:3 + el[i]]) == bytes(g["exp_mpdu"][eo[i]:eo[i + 1]]), i


def test_ragged_batch_vs_oracle(golden):
    pkg = load_pkg()
    g = golden["frames_siso"]
    x = g["iq"][g["offs"][0]:g["offs"][1]]
    y = g["iq"][g["offs"][20]:g["offs"][21]]
    rng = np.random.default_rng(9)
    parts = [x[:64], x[:1450], x[:1700], x[:2500], x, y, y[:900], (0.01 * (rng.standard_normal(5000) + 1j * rng.standard_normal(5000))).astype(np.complex64),
             np.concatenate([x, y])]
    iq = np.concatenate(parts).astype(np.complex64)
    ln = np.array([p.size for p in parts], np.int32)
    off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64)
    fr, pdu, fo, po = _run_both(pkg, iq, off, ln, chunk_items=4)
    _compare(fr, pdu, fo, po)


def test_bench_frames_config5_shape(golden):
    """16 VHT MCS7 1500-byte frames (config 5 units), 400-sample gaps, 30 dB: every MPDU comes back"""
    pkg = load_pkg()
    g = golden["frames_bench"]
    fx, mp = g["iq"], g["mpdu"]
    rng = np.random.default_rng(5)
    items = []
    for k in range(64):
        f = fx[k % 16]
        z = np.concatenate([np.zeros(200, np.complex64), f, np.zeros(200, np.complex64)])
        s = 0.1875 / np.sqrt(2 * 10 ** 3.0)
        items.append((z + s * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))).astype(np.complex64))
    iq = np.concatenate(items)
    ln = np.full(64, items[0].size, np.int32)
    off = (np.arange(64) * items[0].size).astype(np.int64)
    fr, pdu, fo, po = _run_both(pkg, iq, off, ln, chunk_items=24)
    _compare(fr, pdu, fo, po)
    for k in range(64):
        assert fr[k]["npdu"] == 1 and (fr[k]["format"], fr[k]["mcs"], fr[k]["nsym"], fr[k]["trellis"], fr[k]["total"]) == (2, 7, 47, 12220, 14664)
        assert bytes(pdu[k, 3:1503]) == bytes(mp[k % 16])


def test_long_capture_many_frames_in_stream_order(golden):
    """the reference demo: tools/pktGenExample.py:183-199 writes L MCS0-7, HT MCS0-7, VHT MCS0-8 into ONE capture that
    examples/rx.grc decodes frame after frame.  Same here: one item, max_frames records, PDUs in stream order."""
    pkg = load_pkg()
    g = golden["frames_siso"]
    offs = g["offs"]
    x = np.ascontiguousarray(g["iq"][offs[1]:offs[26]])              # 25 frames (the demo set), 400-sample gaps
    rx = pkg.Receiver(device=0, max_frames=32)
    fr, pdu = rx.rx_batch(x, [0], [x.size])
    rx.close()
    fo, _, po = ol.rx_item(x, max_frames=32)
    recs = ol.split_pdus(po)
    assert len(fo) == 25 and len(recs) == 25
    el = g["exp_len"]
    eo = np.cumsum(np.r_[0, el])
    for k in range(32):
        if k >= 25:
            assert fr[k]["status"] == 9                                  # C8B_ST_EMPTY
            continue
        for key in ("status", "sync_idx", "trig_idx", "format", "mcs", "len", "nsym", "trellis", "npdu", "pdu_bytes"):
            assert fr[k][key] == fo[k][key], (k, key, fr[k][key], fo[k][key])
        assert bytes(pdu[k, :fr[k]["pdu_bytes"]]) == recs[k]
        assert recs[k][3:-1] == bytes(g["exp_mpdu"][eo[k + 1]:eo[k + 2]])


def test_two_items_each_with_several_frames(golden):
    pkg = load_pkg()
    g = golden["frames_siso"]
    offs = g["offs"]
    a = g["iq"][offs[3]:offs[7]]
    b = g["iq"][offs[20]:offs[23]]
    iq = np.concatenate([a, b]).astype(np.complex64)
    rx = pkg.Receiver(device=0, max_frames=5, chunk_items=1)
    fr, pdu = rx.rx_batch(iq, [0, a.size], [a.size, b.size])
    rx.close()
    for it, x in enumerate((a, b)):
        fo, _, po = ol.rx_item(np.ascontiguousarray(x), max_frames=5)
        recs = ol.split_pdus(po)
        for k in range(5):
            s = it * 5 + k
            if k < len(fo):
                assert fr[s]["status"] == fo[k]["status"] and fr[s]["sync_idx"] == fo[k]["sync_idx"] and fr[s]["item"] == it
                assert bytes(pdu[s, :fr[s]["pdu_bytes"]]) == recs[k]
            else:
                assert fr[s]["status"] == 9


def test_sc16_ingest_equals_fc32_on_the_widened_capture(golden):
    """c8b_rx_batch_sc16: the capture as interleaved int16 (UHD sc16) -- widened on the device by x * (1 / 32768), exactly the
    floats the fc32 entry point gets from the same quantised capture: frame records and PDUs byte for byte, multi-chunk"""
    pkg = load_pkg()
    g = golden["frames_siso"]
    rng = np.random.default_rng(5)
    s = 0.1875 / np.sqrt(2 * 10 ** 2.8)
    x = g["iq"] + s * (rng.standard_normal(g["iq"].size) + 1j * rng.standard_normal(g["iq"].size))
    q = np.empty((x.size, 2), np.int16)
    q[:, 0] = np.clip(np.round(x.real * 32768 * 0.9), -32768, 32767)
    q[:, 1] = np.clip(np.round(x.imag * 32768 * 0.9), -32768, 32767)
    wide = (q[:, 0].astype(np.float32) * np.float32(1 / 32768) + 1j * (q[:, 1].astype(np.float32) * np.float32(1 / 32768))).astype(np.complex64)
    offs = g["offs"]
    off, ln = offs[:-1], np.diff(offs).astype(np.int32)
    for chunk in (0, 7):
        rx = pkg.Receiver(device=0, chunk_items=chunk)
        fa, pa = rx.rx_batch(wide, off, ln)
        fb, pb = rx.rx_batch_sc16(q, off, ln)
        fc, pc = rx.rx_batch_sc16(q, off, ln)                   # again: the item table is already resident
        rx.close()
        assert fa.tobytes() == fb.tobytes() == fc.tobytes() and np.array_equal(pa, pb) and np.array_equal(pb, pc)
        assert int((fa["npdu"] == 1).sum()) >= 30


@pytest.mark.parametrize("frontend_mode,nitems", [(0, 10), (0, 70), (1, 10)])
def test_colliding_frames_vs_oracle(golden, frontend_mode, nitems):
    """a second frame starting inside the first, a third close behind (tests/oracle_lib.py colliding_captures; the reference's
    own blocks equal the oracle on the same recipe, tests/test_ref_chain.py): the trigger inside a payload, sync's hold-off and
    signal's swallow rule on all three detect kernels -- the segment-parallel scan (few items), the warp-per-item kernel
    (> 64 items) and the one-thread kernel -- and through the stream session in pieces"""
    pkg = load_pkg()
    caps = ol.colliding_captures(golden["frames_siso"], np.random.default_rng(4242 + nitems), n=nitems)
    off = np.cumsum([0] + [c.size for c in caps]).astype(np.int64)
    iq = np.concatenate(caps)
    MF = 8
    rx = pkg.Receiver(device=0, max_frames=MF, frontend_mode=frontend_mode)
    fr, pdu = rx.rx_batch(iq, off[:-1], np.diff(off).astype(np.int32), pdu_stride=4400)
    rx.close()
    nf = 0
    for it, x in enumerate(caps[:24]):
        fo, _, po = ol.rx_item(x, max_frames=MF)
        recs = ol.split_pdus(po)
        r = 0
        for k in range(MF):
            s = it * MF + k
            if k >= len(fo):
                assert fr[s]["status"] == 9 or (k == 0 and len(fo) == 0), (it, k, fr[s]["status"])
                continue
            for key in ("status", "sync_idx", "trig_idx", "format", "mcs", "len", "nsym", "npdu", "pdu_bytes"):
                assert fr[s][key] == fo[k][key], (it, k, key, fr[s][key], fo[k][key])
            got = ol.split_pdus(bytes(pdu[s, :fr[s]["pdu_bytes"]]))
            assert got == recs[r:r + len(got)], (it, k)
            r += len(got)
            nf += 1
    assert nf >= min(nitems, 24)
    if nitems == 10 and frontend_mode == 0:                                  # the same captures as a live stream, pushed in pieces
        rx = pkg.Receiver(device=0, max_frames=64, chunk_items=1)
        for it, x in enumerate(caps[:6]):
            fo, _, po = ol.rx_item(x, max_frames=16)
            rx.stream_begin(1, 1 << 16)
            got, rng = [], np.random.default_rng(it)
            k = 0
            while k < x.size:
                n = int(rng.integers(500, 4000))
                f2, base, p2 = rx.stream_push(x[k:k + n], flush=k + n >= x.size, frames_cap=64)
                for q in range(f2.size):
                    got.append((int(base[q]) + int(f2[q]["sync_idx"]), int(f2[q]["status"]), bytes(p2[q, :f2[q]["pdu_bytes"]])))
                k += n
            want = []
            recs, r = ol.split_pdus(po), 0
            for f in fo:
                npd = int(f["npdu"])
                want.append((int(f["sync_idx"]), int(f["status"]), b"".join(recs[r:r + npd])))
                r += npd
            assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in want], (it, got[:3], want[:3])
            assert [c for _, _, c in got] == [c for _, _, c in want], it
        rx.close()
