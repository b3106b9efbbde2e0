"""Chain-level pin of the oracle to the reference ITSELF: the reference's seven receive blocks (lib/trigger_impl.cc,
sync_impl.cc, signal_impl.cc, signal2_impl.cc, demod_impl.cc, demod2_impl.cc, decode_impl.cc + cloud80211phy.cc), compiled
UNMODIFIED against the miniature GNU Radio runtime of tests/gr_mock/include into oracle/_ref/libgr80211_ref.so
(oracle/Makefile, oracle/ref_chain.cc = the scheduler), run over whole captures -- and everything they make visible
(trigger flags, sync flags and tags, signal's tags and CFO-corrected copy, demod's tags and soft bits, decode's messages) is
compared BIT FOR BIT with the restatement oracle/oracle_rx.cc, which every GPU parity test is written against.  State
machines, consume / produce accounting, tag offsets, the S_COPY swallow rule, CLEAN, range checks, the NDP report and the
A-MPDU walk are thereby pinned to the reference's own code, for any scheduler call sizes.  CPU only."""
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.skipif(not ol.have_refchain(), reason="oracle/_ref/libgr80211_ref.so not built (needs /root/reference)")
HERE = os.path.dirname(os.path.abspath(__file__))


def _noisy(x, snr, seed, amp=0.1875):
    rng = np.random.default_rng(seed)
    s = amp / np.sqrt(2 * 10 ** (snr / 10))
    return (x + s * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)


def _check(bad, info):
    assert not bad, (bad[:6], info)


def test_reference_sources_are_compiled_in_place():
    """the recipe names the reference's files where they lie; nothing of them is copied into the repo"""
    mk = open(os.path.join(ol.ORACLE_DIR, "Makefile")).read()
    assert "$(REF)/lib/%_impl.cc" in mk and "REF ?= /root/reference" in mk
    for blk in ("trigger", "sync", "signal", "signal2", "demod", "demod2", "decode"):
        assert not os.path.exists(os.path.join(ol.ORACLE_DIR, blk + "_impl.cc"))
    L = ol.refchain_lib()
    for sym in ("refchain_create", "refchain_run", "refchain_stream", "refchain_tag", "refchain_msg", "refchain_bench"):
        assert hasattr(L, sym)


@pytest.mark.parametrize("snr,seed,max_call", [(None, 0, 0), (30, 1, 4096), (20, 2, 900), (12, 3, 1500), (8, 0, 1000), (5, 5, 2000), (2, 6, 3000), (0, 7, 0)])
def test_siso_capture(golden, snr, seed, max_call):
    """31 back-to-back frames (L MCS0-7, HT MCS0-7, VHT MCS0-8, CFO cases, 2-subframe A-MPDU) of the reference's generator"""
    x = golden["frames_siso"]["iq"]
    y = x if snr is None else _noisy(x, snr, 100 + seed)
    bad, info = ol.chain_vs_oracle(y, seed=seed, max_call=max_call)
    _check(bad, info)
    assert info["frames"] >= (16 if snr is None or snr >= 2 else 4)
    if snr is None or snr >= 30:
        assert info["messages"] == 31 and info["soft_bits"] == 42712


def test_siso_truncated_and_piecewise(golden):
    x = _noisy(golden["frames_siso"]["iq"], 25, 1)
    rng = np.random.default_rng(5)
    seen = set()
    for cut in rng.integers(300, x.size, 16):
        pieces = sorted(int(v) for v in rng.integers(1, cut, 3))
        bad, info = ol.chain_vs_oracle(x[:cut], seed=int(cut), max_call=3000, pieces=pieces)
        _check(bad, info)
        seen.add(info["statuses"][-1] if info["statuses"] else -1)
    assert 4 in seen and 0 in seen          # frames cut short (S_COPY never finishes) and whole ones


def test_colliding_frames(golden):
    rng = np.random.default_rng(4242)
    nf = 0
    for k, x in enumerate(ol.colliding_captures(golden["frames_siso"], rng)):
        bad, info = ol.chain_vs_oracle(x, seed=k, max_call=2500 if k & 1 else 0, max_frames=16)
        _check(bad, info)
        nf += info["frames"]
    assert nf >= 10


@pytest.mark.parametrize("snr,seed,max_call", [(None, 0, 0), (25, 1, 0), (10, 3, 1700)])
def test_2x2_capture(golden, snr, seed, max_call):
    """signal2 / demod2: HT MCS8-15 and VHT 2SS frames, both antennas"""
    g = golden["frames_mimo"]
    a, b = g["iq0"], g["iq1"]
    if snr is not None:
        a, b = _noisy(a, snr, 13579), _noisy(b, snr, 24680)
    bad, info = ol.chain_vs_oracle(a, b, seed=seed, max_call=max_call)
    _check(bad, info)
    assert info["frames"] == 18 and (snr == 10 or info["messages"] == 18)


@pytest.mark.parametrize("mupos", [0, 1])
def test_mu_mimo_and_ndp(golden, mupos):
    """demod(mupos, mugid): user-position channel estimate on 2-user frames; NDP -> 1024 floats + channel report message"""
    bad, info = ol.chain_vs_oracle(golden["frames_mu"]["iq"], mupos=mupos, mugid=2)
    _check(bad, info)
    assert info["statuses"][:2] == [7, 7] and info["messages"] == 4


@pytest.mark.parametrize("mupos", [0, 1])
def test_mu_mimo_two_subframe_ampdu(mupos):
    """a 2-user frame whose user 1 carries a TWO-subframe A-MPDU (golden frames_mu_tx.npz zf_rx*: the generator's frame through a
    zero-forcing precoder and a flat channel): the reference's blocks publish ONE message per station -- in the MU receive path
    the walk over the A-MPDU ends after the first subframe -- and the oracle restates exactly that"""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_mu_tx.npz"))
    bad, info = ol.chain_vs_oracle(g["zf_rx%d" % mupos], mupos=mupos, mugid=2, max_frames=4)
    _check(bad, info)
    assert info["frames"] == 1 and info["statuses"] == [0] and info["messages"] == 1


def test_short_gi_every_symbol():
    """72-sample raster: every symbol's soft bits of the reference's demod equal the oracle's"""
    g = np.load(os.path.join(HERE, "golden", "frames_sgi.npz"))
    for snr in (None, 30):
        x = g["iq"] if snr is None else _noisy(g["iq"], snr, 9)
        bad, info = ol.chain_vs_oracle(x)
        _check(bad, info)
        assert info["frames"] == 3 and info["soft_bits"] == 4056


def test_short_gi_frames_that_decode():
    """frames_sgi_true: the 72-sample-raster waveform the reference's receiver decodes (first 72 samples of every 80-sample DATA
    symbol, see make_golden.frames_sgi_true): oracle == reference blocks on every symbol, and the HT MCS0 / MCS5 and VHT MCS3
    frames come back as their MPDUs"""
    g = np.load(os.path.join(HERE, "golden", "frames_sgi_true.npz"))
    for snr in (None, 30):
        x = g["iq"] if snr is None else _noisy(g["iq"], snr, 10)
        bad, info = ol.chain_vs_oracle(x, seed=4, max_call=1800)
        _check(bad, info)
        assert info["frames"] == 5 and info["soft_bits"] == 17056 and info["messages"] == 3
    eo = np.cumsum(np.r_[0, g["exp_len"]])
    offs = g["offs"]
    for i in (0, 2, 4):
        fo, _, po = ol.rx_item(g["iq"][offs[i]:offs[i + 1]], max_frames=1)
        assert fo[0]["nsymsamp"] == 72 and ol.split_pdus(po)[0][3:3 + g["exp_len"][i]] == bytes(g["exp_mpdu"][eo[i]:eo[i + 1]])


def test_564_byte_frames_all_rates():
    """config 2 / 3 / 4 frame sizes: L MCS0-7, VHT MCS0-8 (SISO), HT MCS8-15 (2x2), 30 dB"""
    g = np.load(os.path.join(HERE, "golden", "frames_564.npz"))
    z = np.zeros(600, np.complex64)
    siso = np.concatenate([np.concatenate([z, g[k], z]) for k in ["l%d" % i for i in range(8)] + ["v%d" % i for i in range(9)]])
    bad, info = ol.chain_vs_oracle(_noisy(siso, 30, 4), seed=11, max_call=8192)
    _check(bad, info)
    assert info["frames"] == 17 and info["messages"] == 17
    a = np.concatenate([np.concatenate([z, g["h%d_0" % i], z]) for i in range(8, 16)])
    b = np.concatenate([np.concatenate([z, g["h%d_1" % i], z]) for i in range(8, 16)])
    bad, info = ol.chain_vs_oracle(_noisy(a, 30, 13579, 0.1875 * np.sqrt(2)), _noisy(b, 30, 24680, 0.1875 * np.sqrt(2)))
    _check(bad, info)
    assert info["frames"] == 8 and info["messages"] == 8


def test_bench_frames(golden):
    """config 5 frames (VHT MCS7, 1500-byte MPDU, 47 symbols) as one stream with 400-sample gaps"""
    iq = golden["frames_bench"]["iq"]
    z = np.zeros(400, np.complex64)
    x = np.concatenate([np.concatenate([fr, z]) for fr in iq[:8]] + [z])
    bad, info = ol.chain_vs_oracle(_noisy(x, 30, 8))
    _check(bad, info)
    assert info["frames"] == 8 and info["messages"] == 8 and info["soft_bits"] == 8 * 14664


def test_junk_inputs():
    rng = np.random.default_rng(1)
    noise = (0.05 * (rng.standard_normal(40000) + 1j * rng.standard_normal(40000))).astype(np.complex64)
    for x in (noise, np.zeros(5000, np.complex64), noise[:100], noise[:0],
              np.tile(np.exp(2j * np.pi * np.arange(16) / 16).astype(np.complex64), 2000)):     # a periodic tone: plateau forever
        bad, info = ol.chain_vs_oracle(x, seed=3, max_call=2048)
        _check(bad, info)
        assert info["messages"] == 0
