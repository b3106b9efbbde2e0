"""The reference-API mirror (gr-ieee80211_b200/flowgraph.py): block names / ctor args, tags, PDU messages, UDP
datagrams and decode's debug line, exercised the way the reference's own harness does
(tools/performance/perf_siso.py:105-118 scrapes the last debug line; README demo decodes a 25-frame capture)."""
import os
import re
import socket

import numpy as np
import pytest

from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


def test_rx_grc_demo_capture(golden, tmp_path):
    pkg = load_pkg()
    fg = pkg.flowgraph
    g = golden["frames_siso"]
    offs = g["offs"]
    path = os.path.join(tmp_path, "sig80211GenMultipleSiso_1x1_0.bin")
    fg.write_bin(path, g["iq"][offs[1]:offs[26]])                 # the README demo set: L 0-7, HT 0-7, VHT 0-8
    srv = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    srv.bind(("127.0.0.1", 0))
    srv.settimeout(5)
    lines = []
    tb = fg.rx_top_block(nant=1, ifdebug=True, udp=srv.getsockname(), printer=lines.append)
    assert (tb.demod.mupos, tb.demod.mugid) == (0, 2) and tb.decode.d_debug
    fr = tb.run(path)
    tb.close()
    assert len(fr) == 25 and len(tb.decode.out) == 25
    el = g["exp_len"]
    eo = np.cumsum(np.r_[0, el])
    for k in range(25):
        dgram, _ = srv.recvfrom(65536)                              # UDP PDU framing [fmt][len lo][len hi][MPDU][mcs]
        mp = bytes(g["exp_mpdu"][eo[k + 1]:eo[k + 2]])
        fmt, mcs = int(g["meta"][k + 1][0]), int(g["meta"][k + 1][1])
        assert dgram == bytes([fmt, len(mp) & 255, len(mp) >> 8]) + mp + bytes([mcs])
    # the scrape of tools/performance/perf_siso.py:105-118: last line of each format, split on ',' and ':'
    def last(name):
        return [ln for ln in lines if "crc32" in ln and name in ln][-1]
    leg = re.split(r"[,:]", last("legacy"))
    assert leg[0] == "ieee80211 decode" and leg[1] == " legacy crc32 correct" and leg[2] == " total" and int(leg[3]) == 8
    assert [int(leg[5 + 2 * i]) for i in range(8)] == [1] * 8
    vht = re.split(r"[,:]", last("vht"))
    assert int(vht[3]) == 25 and [int(vht[5 + 2 * i]) for i in range(10)] == [1] * 9 + [0]
    assert "sssnr0" in last("vht") and "sssnr1" in last("vht") and "sssnr0" not in last(", ht crc32")
    assert re.search(r",cfo:-?\d+\.\d{6},snr:", last(", ht crc32"))
    # tags with the reference's keys
    assert set(tb.sync.tags[0]) == {"rad", "snr", "rssi"}
    assert set(tb.signal.tags[0]) == {"cfo", "snr", "rssi", "seq", "mcs", "len", "nsamp", "chan"} and tb.signal.tags[3]["seq"] == 3
    assert tb.signal.tags[0]["chan"].shape == (64,)
    assert {"format", "mcs", "len", "cr", "ampdu", "trellis", "total"} <= set(tb.demod.tags[0])
    assert "sssnr0" in tb.demod.tags[-1] and "sssnr0" not in tb.demod.tags[0]


def test_rx2_grc_capture(golden, tmp_path):
    pkg = load_pkg()
    fg = pkg.flowgraph
    g = golden["frames_mimo"]
    p0, p1 = os.path.join(tmp_path, "mimo_0.bin"), os.path.join(tmp_path, "mimo_1.bin")
    fg.write_bin(p0, g["iq0"])
    fg.write_bin(p1, g["iq1"])
    tb = fg.rx_top_block(nant=2, ifdebug=False, printer=None)
    fr = tb.run(p0, p1)
    tb.close()
    assert len(fr) == 18 and len(tb.decode.out) == 18 and tb.decode.d_nPktCorrect == 0     # counters only move with ifdebug
    assert [int(r[0]) for r in tb.decode.out] == [1] * 8 + [2] * 9 + [1]
    assert "sssnr1" in tb.demod.tags[10]


def test_work_calls_equal_run(golden):
    """the capture delivered in general_work-sized pieces publishes the same PDUs, tags and debug lines as run()"""
    pkg = load_pkg()
    fg = pkg.flowgraph
    g = golden["frames_siso"]
    offs = g["offs"]
    x = np.ascontiguousarray(g["iq"][offs[1]:offs[26]])
    a, b = [], []
    tb = fg.rx_top_block(nant=1, ifdebug=True, printer=a.append)
    tb.run(x)
    out_run, tags_run = list(tb.decode.out), repr([sorted(t.items()) for t in tb.demod.tags])      # repr: noiseless snr is nan
    sync_run = list(tb.sync.offsets)
    tb.close()
    tb = fg.rx_top_block(nant=1, ifdebug=True, printer=b.append, max_frames=8)
    for k in range(0, x.size, 4096):                                  # 4096 items per call, like a GNU Radio buffer
        tb.work(x[k:k + 4096], flush=k + 4096 >= x.size)
    assert tb.decode.out == out_run and len(out_run) == 25
    assert repr([sorted(t.items()) for t in tb.demod.tags]) == tags_run
    assert tb.sync.offsets == sync_run and len(sync_run) == 25
    assert a == b
    tb.close()


def test_run_reports_used_up_frame_records(golden):
    """run() over a capture with more frames than max_frames keeps the first max_frames frames -- and says so"""
    pkg = load_pkg()
    fg = pkg.flowgraph
    g = golden["frames_siso"]
    offs = g["offs"]
    x = np.ascontiguousarray(g["iq"][offs[1]:offs[26]])
    lines = []
    tb = fg.rx_top_block(nant=1, ifdebug=False, printer=lines.append, max_frames=8)
    fr = tb.run(x)
    assert len(fr) == 8 and tb.truncated and any("frame records" in s for s in lines)
    tb.close()
    tb = fg.rx_top_block(nant=1, ifdebug=False, printer=lines.append, max_frames=64)
    fr = tb.run(x)
    assert len(fr) == 25 and not tb.truncated
    tb.close()
