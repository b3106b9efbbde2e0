"""MU-MIMO user positions (demod(mupos, mugid), lib/demod_impl.cc:347-380) and the VHT NDP channel report (tag mu2x1chan
lib/demod_impl.cc:238-249, blob lib/decode_impl.cc:100-121) on the GPU path vs the oracle, on frames made by the reference's
own generator (tests/golden/frames_mu.npz)."""
import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


def _items(g, snr=None):
    offs = g["offs"]
    iq = g["iq"]
    if snr is not None:
        rng = np.random.default_rng(31)
        s = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
        iq = (iq + s * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))).astype(np.complex64)
    return iq, offs[:-1], np.diff(offs).astype(np.int32)


@pytest.mark.parametrize("frontend_mode", [0, 1])
@pytest.mark.parametrize("decode_mode", [1, 2])
@pytest.mark.parametrize("snr", [None, 30.0])
def test_mu_and_ndp_vs_oracle(golden, snr, decode_mode, frontend_mode):
    pkg = load_pkg()
    g = golden["frames_mu"]
    iq, off, ln = _items(g, snr)
    O = ol.oracle()
    el = g["exp_len"]
    eo = np.cumsum(np.r_[0, el])
    try:
        for mupos in (0, 1):
            rx = pkg.Receiver(device=0, mupos=mupos, mugid=2, decode_mode=decode_mode, frontend_mode=frontend_mode)
            fr, pdu = rx.rx_batch(iq, off, ln)
            rx.close()
            O.orx_set_mupos(mupos)
            for i in range(len(off)):
                kind, mcs, par = g["meta"][i]
                fo, _, po = ol.rx_item(iq[off[i]:off[i] + ln[i]], max_frames=1)
                f = fr[i]
                if kind == 4 and int(par) != mupos:
                    # the other station's frame: zero-forcing nulls it here, the channel estimate is numerical noise and the
                    # SIG-B bits are arbitrary -- the only requirement is that nothing is published
                    assert f["npdu"] == 0 and fo[0]["npdu"] == 0
                    continue
                for k in ("status", "format", "mcs", "len", "nss", "nsym", "total", "npdu", "pdu_bytes"):
                    assert f[k] == fo[0][k], (mupos, i, k, f[k], fo[0][k])
                nb = int(f["pdu_bytes"])
                if kind == 3:                                          # NDP: header bytes exact, samples to float tolerance
                    assert f["status"] == 7 and nb == 1027 and bytes(pdu[i, :3]) == bytes(po[:3]) == bytes([20, 0, 4])
                    a = np.frombuffer(bytes(pdu[i, 3:1027]), np.float32)
                    b = np.frombuffer(bytes(po[3:1027]), np.float32)
                    assert np.max(np.abs(a - b)) <= 1e-5 * max(1.0, float(np.max(np.abs(b))))
                else:
                    assert bytes(pdu[i, :nb]) == bytes(po[:nb]), (mupos, i)
                    if int(par) == mupos:                              # this station's frame decodes to its MPDU
                        assert f["status"] == 0 and bytes(pdu[i, 3:nb - 1]) == bytes(g["exp_mpdu"][eo[i]:eo[i + 1]])
    finally:
        O.orx_set_mupos(0)


def test_ndp_staged_and_flowgraph(golden):
    """c8b_demod leaves the tag content at the frame's place in the LLR arena, c8b_decode turns it into the blob; the
    flowgraph mirror publishes it on the message port and tags it on demod's output"""
    pkg = load_pkg()
    g = golden["frames_mu"]
    iq, off, ln = _items(g)
    rx = pkg.Receiver(device=0)
    fr, chan = rx.detect(iq, off[:2], ln[:2])
    stride = 4096
    fr2, llr = rx.demod(iq, off[:2], ln[:2], fr, chan, stride)
    assert list(fr2["status"]) == [7, 7] and list(fr2["total"]) == [1024, 1024]
    fr3, pdu, _ = rx.decode(llr.reshape(-1), fr2, pdu_stride=1100)
    for i in range(2):
        assert fr3[i]["npdu"] == 1 and fr3[i]["pdu_bytes"] == 1027
        assert np.array_equal(np.frombuffer(bytes(pdu[i, 3:1027]), np.float32), llr[i, :256])
    rx.close()
    tb = pkg.flowgraph.rx_top_block(nant=1, ifdebug=True, printer=None, max_frames=4)
    x = iq[off[0]:off[0] + ln[0]]
    tb.run(x)
    assert len(tb.decode.out) == 1 and len(tb.decode.out[0]) == 1027 and tb.decode.out[0][0] == 20
    t = tb.demod.tags[0]
    assert t["total"] == 1024 and t["mu2x1chan"].size == 128 and tb.decode.debug_lines == []
    tb.close()
