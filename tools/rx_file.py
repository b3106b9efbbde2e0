#!/usr/bin/env python
"""Headless equivalent of examples/rx.grc / rx2.grc on a capture file (cf. tools/performance/gr_siso.py):
  python tools/rx_file.py sig.bin                 # SISO
  python tools/rx_file.py sig_0.bin sig_1.bin     # 2x2
  python tools/rx_file.py --sc16 capture.sc16     # interleaved int16 I/Q (UHD sc16) instead of fc32
prints decode's debug lines and sends every PDU to 127.0.0.1:9527 (tools/macExampleGrRx.py:29-43 listens there)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_pkg  # noqa: E402

pkg = load_pkg()
files = [a for a in sys.argv[1:] if not a.startswith("-")]
if not files:
    raise SystemExit(__doc__)
tb = pkg.flowgraph.rx_top_block(nant=len(files), ifdebug=True, udp=("127.0.0.1", 9527))
sc16 = "--sc16" in sys.argv
fr = tb.run(*[pkg.flowgraph.read_bin(f, sc16=sc16) for f in files])
print("frames: %d, PDUs published: %d" % (len(fr), len(tb.decode.out)))
tb.close()
