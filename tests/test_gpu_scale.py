"""BASELINE.json configs 2-4 at their full sizes, checked through size-independent properties:
 * round trip: every item must come back as the MPDU that was transmitted (a small loss is allowed at 30 dB for the
   densest constellations; a PUBLISHED PDU may never differ from the transmitted one -- CRC-32 guards that);
 * the GPU and the oracle agree on every item the GPU failed plus a random sample of decoded ones."""
import numpy as np
import pytest
import torch

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


def _check(pkg, fr, pdu, kind, fmts, mcss, mpdus, min_ok, name):
    n = fr.size
    good = np.zeros(n, bool)
    wrong = 0
    for i in range(n):
        if fr[i]["status"] == 0 and fr[i]["npdu"] >= 1:
            m = mpdus[kind[i]]
            rec = bytes(pdu[i, :len(m) + 4])
            if rec == bytes([fmts[kind[i]], len(m) & 255, len(m) >> 8]) + m + bytes([mcss[kind[i]]]):
                good[i] = True
            else:
                wrong += 1
    assert wrong == 0, "%s: %d published PDUs differ from what was sent" % (name, wrong)
    frac = good.mean()
    print("%s: %d/%d items decoded (%.4f)" % (name, good.sum(), n, frac))
    assert frac >= min_ok, (name, frac)
    return good


def _oracle_agrees(fr, pdu, good, items_of, nsample=48, two=False):
    rng = np.random.default_rng(0)
    bad = np.nonzero(~good)[0][:64]
    idx = np.concatenate([bad, rng.choice(np.nonzero(good)[0], size=min(nsample, int(good.sum())), replace=False)])
    for i in idx:
        x = items_of(int(i))
        fo, _, po = (ol.rx_item2(x[0], x[1], max_frames=1) if two else ol.rx_item(x, max_frames=1))
        assert fo[0]["status"] == fr[i]["status"] and fo[0]["npdu"] == fr[i]["npdu"] and po.size == fr[i]["pdu_bytes"], (int(i), fo[0]["status"], fr[i]["status"])
        assert bytes(po) == bytes(pdu[i, :po.size]), int(i)


def test_config2_legacy_mcs0_7_10k_frames():
    pkg = load_pkg()
    g = np.load(__file__.rsplit("/", 1)[0] + "/golden/frames_564.npz")
    dev = torch.device("cuda", 0)
    frames = [g["l%d" % m] for m in range(8)]
    (iq,), off, ln, kind = pkg.synth.make_items(torch, dev, frames, [1250] * 8, snr_db=30.0, seed=2)
    rx = pkg.Receiver(device=0, chunk_items=4096)
    fr, pdu = rx.rx_batch_dev(iq.data_ptr(), off, ln, pdu_stride=640)
    rx.close()
    mp = bytes(g["mpdu"])
    good = _check(pkg, fr, pdu, kind, [0] * 8, list(range(8)), [mp] * 8, 0.999, "config 2")
    h = iq.cpu().numpy()
    _oracle_agrees(fr, pdu, good, lambda i: h[off[i]:off[i] + ln[i]])


def test_config3_vht_mcs0_8_100k_frames():
    pkg = load_pkg()
    g = np.load(__file__.rsplit("/", 1)[0] + "/golden/frames_564.npz")
    dev = torch.device("cuda", 0)
    frames = [g["v%d" % m] for m in range(9)]
    counts = [11112] * 8 + [11104]
    (iq,), off, ln, kind = pkg.synth.make_items(torch, dev, frames, counts, snr_db=30.0, seed=3)
    rx = pkg.Receiver(device=0, chunk_items=16384)
    fr, pdu = rx.rx_batch_dev(iq.data_ptr(), off, ln, pdu_stride=640)
    rx.close()
    mp = bytes(g["vht_mpdu"])
    good = _check(pkg, fr, pdu, kind, [2] * 9, list(range(9)), [mp] * 9, 0.995, "config 3")
    for m in range(8):                                   # up to 64-QAM 5/6 nothing may be lost at 30 dB
        assert good[kind == m].mean() >= 0.999, m
    sel = np.concatenate([np.nonzero(~good)[0][:64], np.random.default_rng(1).choice(fr.size, 64, replace=False)])
    sel_t = {int(i): iq[int(off[i]): int(off[i] + ln[i])].cpu().numpy() for i in sel}
    rng_good = np.zeros(fr.size, bool)
    rng_good[[i for i in sel_t if good[i]]] = True
    for i, x in sel_t.items():
        fo, _, po = ol.rx_item(x, max_frames=1)
        assert fo[0]["status"] == fr[i]["status"] and fo[0]["npdu"] == fr[i]["npdu"] and bytes(po) == bytes(pdu[i, :po.size]), i


def test_config3_mcs9_llr_only_case():
    """VHT MCS9 cannot be generated at 20 MHz / 1 SS (tools/phy80211header.py:373-376); the C++ still handles the index
    (256-QAM 5/6, lib/cloud80211phy.cc:1275-1299): exercise demap-free decode with nDBPS 346 soft bits -> bit-exact vs oracle"""
    pkg = load_pkg()
    rng = np.random.default_rng(9)
    T, cr = 346 * 12, 3
    total = 416 * 12
    llr = (np.round(rng.normal(0, 2, (4, total)) * 4) / 4).astype(np.float32)
    fr = np.zeros(4, pkg.FRAME_DTYPE)
    fr["cr"], fr["trellis"], fr["total"], fr["format"], fr["mcs"], fr["len"], fr["ampdu"] = cr, T, total, 2, 9, 500, 1
    fr["llr_off"] = np.arange(4) * total
    rx = pkg.Receiver(device=0)
    out, pdu, scram = rx.decode(llr, fr, pdu_stride=640, want_scram=True)
    rx.close()
    for i in range(4):
        want = np.zeros(T, np.uint8)
        ol.oracle().orx_viterbi(np.ascontiguousarray(llr[i]), cr, T, want)
        assert np.array_equal(scram[i, :T], want)


def test_config4_ht_2x2_mcs8_15_50k_frames():
    pkg = load_pkg()
    g = np.load(__file__.rsplit("/", 1)[0] + "/golden/frames_564.npz")
    dev = torch.device("cuda", 0)
    frames = [(g["h%d_0" % m], g["h%d_1" % m]) for m in range(8, 16)]
    (a, b), off, ln, kind = pkg.synth.make_items(torch, dev, frames, [6250] * 8, snr_db=30.0, seed=4, rms=0.1875)
    ha, hb = a.cpu().numpy(), b.cpu().numpy()
    rx = pkg.Receiver(device=0, chunk_items=8192)
    fr, pdu = rx.rx_batch2(ha, hb, off, ln, pdu_stride=640)
    frd, pdud = rx.rx_batch2_dev(a.data_ptr(), b.data_ptr(), off, ln, pdu_stride=640)      # same capture, device resident
    rx.close()
    assert frd.tobytes() == fr.tobytes() and np.array_equal(pdud, pdu)
    mp = bytes(g["mpdu"])
    good = _check(pkg, fr, pdu, kind, [1] * 8, list(range(8, 16)), [mp] * 8, 0.995, "config 4")
    _oracle_agrees(fr, pdu, good, lambda i: (ha[off[i]:off[i] + ln[i]], hb[off[i]:off[i] + ln[i]]), two=True)


def test_config5_vht_mcs7_1500B_shard():
    """config 5 units (VHT MCS7, 1500-byte MPDU, 4960-sample items): two full pipeline chunks; every item the GPU does not
    decode must fail identically in the oracle, a sample of decoded ones must match byte for byte"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    pkg = load_pkg()
    dev = torch.device("cuda", 0)
    n = 2 * 56832
    iq, mpdus = bench.make_batch_device(torch, dev, n, seed=5)
    off = np.arange(n, dtype=np.int64) * bench.ITEM
    ln = np.full(n, bench.ITEM, np.int32)
    rx = pkg.Receiver(device=0)
    fr, pdu = rx.rx_batch_dev(iq.data_ptr(), off, ln, pdu_stride=bench.PDU_STRIDE)
    rx.close()
    good = (fr["status"] == 0) & (fr["npdu"] == 1) & (fr["pdu_bytes"] == 1504)
    assert good.mean() > 0.9995
    sel = np.concatenate([np.nonzero(~good)[0], np.random.default_rng(2).choice(np.nonzero(good)[0], 96, replace=False)])
    h = iq.view(n, bench.ITEM)[torch.from_numpy(sel).to(dev)].cpu().numpy()
    for k, i in enumerate(sel):
        fo, _, po = ol.rx_item(np.ascontiguousarray(h[k]), max_frames=1)
        assert fo[0]["status"] == fr[i]["status"] and fo[0]["npdu"] == fr[i]["npdu"] and fo[0]["sync_idx"] == fr[i]["sync_idx"], int(i)
        assert bytes(po) == bytes(pdu[i, :po.size]), int(i)
        if good[i]:
            assert bytes(pdu[i, 3:1503]) == bytes(mpdus[int(i) % 16])


def test_config5_the_bench_batch_failures_fail_in_the_reference_too():
    """The batch bench.py times (configs[4]: 1 048 576 unique VHT MCS7 1500-byte frames made by the transmit synthesiser, AWGN
    30 dB, rank-0 seed): EVERY frame the GPU does not decode to its own MPDU is re-run through the oracle and through the
    reference's own blocks (oracle/_ref, when built) and must fail there the same way -- same drop code, same records; every
    decoded frame equals the bytes that were sent (checked on the device, as bench.py does)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    pkg = load_pkg()
    dev = torch.device("cuda", 0)
    n = 1 << 20
    rx = pkg.Receiver(device=0)
    iq, psdu = bench.make_batch_tx(torch, dev, pkg, rx, n, seed=0)
    off = np.arange(n, dtype=np.int64) * bench.ITEM
    ln = np.full(n, bench.ITEM, np.int32)
    d_frames = torch.zeros(n * pkg.FRAME_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_pdu = torch.zeros(n * bench.PDU_STRIDE, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    rx.rx_batch_dev_async(iq.data_ptr(), off, ln, d_frames.data_ptr(), d_pdu.data_ptr(), bench.PDU_STRIDE)
    rx.sync()
    rx.close()
    fr = np.frombuffer(d_frames.cpu().numpy().tobytes(), dtype=pkg.FRAME_DTYPE)
    good = (fr["status"] == 0) & (fr["npdu"] == 1) & (fr["pdu_bytes"] == bench.MPDU_LEN + 4)
    same = torch.zeros(n, dtype=torch.bool, device=dev)
    for b in range(0, n, 65536):
        e = min(n, b + 65536)
        same[b:e] = (d_pdu.view(n, bench.PDU_STRIDE)[b:e, 3:3 + bench.MPDU_LEN] == psdu[b:e, 4:4 + bench.MPDU_LEN]).all(dim=1)
    same = same.cpu().numpy()
    assert np.array_equal(same & good, good)                       # every decoded frame carries the MPDU that was sent
    bad = np.nonzero(~good)[0]
    assert 0 < bad.size < n * 5e-4, bad.size
    h = iq.view(n, bench.ITEM)[torch.from_numpy(bad).to(dev)].cpu().numpy()
    hp = d_pdu.view(n, bench.PDU_STRIDE)[torch.from_numpy(bad).to(dev)].cpu().numpy()
    ref = ol.have_refchain()
    for k, i in enumerate(bad):
        x = np.ascontiguousarray(h[k])
        fo, _, po = ol.rx_item(x, max_frames=1)
        assert fo[0]["status"] == fr[i]["status"] and fo[0]["npdu"] == fr[i]["npdu"] and po.size == fr[i]["pdu_bytes"], (int(i), fo[0]["status"], fr[i]["status"])
        assert bytes(po) == bytes(hp[k, :po.size]), int(i)
        if ref:
            c = ol.RefChain()
            c.run(x, record=False)
            assert c.messages() == ol.split_pdus(po), int(i)       # the reference's blocks publish exactly the oracle's records (none, mostly)
            c.close()
    print("config 5, 1M frames: %d not decoded; oracle%s agree on every one" % (bad.size, " and reference blocks" if ref else ""))
