"""Host-side mirror of the reference's receive flowgraph API (GNU Radio is absent from this image, so these are
plain Python objects, not gr::blocks): the same block names and constructor arguments as
`gnuradio.ieee80211` (grc/ieee80211_*.block.yml: trigger(), sync(), signal(), signal2(), demod(mupos, mugid),
demod2(), decode(ifdebug)), wired like examples/rx.grc:753-767 / rx2.grc:676-692, fed from the reference's fc32
`.bin` captures (tools/phy80211.py:1063-1090) and emitting what the reference emits:

  * stream tags per frame with the reference's keys (sync: rad/snr/rssi, lib/sync_impl.cc:124-136;
    signal: cfo/snr/rssi/seq/mcs/len/nsamp/chan, lib/signal_impl.cc:135-152;
    demod: format/mcs/len/cr/ampdu/trellis/total/sssnr0[/sssnr1], lib/demod_impl.cc:224-263),
  * PDUs `[fmt][len lo][len hi][MPDU][mcs]` on the decode block's message port `out` (list, callback and/or UDP
    datagrams to 127.0.0.1:9527 like network.socket_pdu in rx.grc:758),
  * decode(ifdebug=True)'s debug lines and per-MCS counters, byte-compatible with lib/decode_impl.cc:377-411,
    456-509 (tools/performance/perf_siso.py:105-118 scrapes them).

All signal processing happens in libc80211b200.so on the GPU; a capture is handed over as ONE item with
`max_frames` records, so frames come back in stream order exactly as the block chain would produce them."""
import socket

import numpy as np

from .rx import Receiver, split_pdus

FORMAT_NAME = {0: "legacy", 1: "ht", 2: "vht"}


def read_bin(path, sc16=False):
    """fc32 interleaved capture, as written by tools/phy80211.py genMultiSigBinFile / blocks.file_source(gr_complex); sc16 = a
    UHD-style capture of interleaved int16 I/Q (`uhd_rx_cfile --wire sc16 -s`), widened the way UHD does: x / 32768"""
    if sc16:
        return (np.fromfile(path, dtype="<i2").astype(np.float32) * np.float32(1.0 / 32768.0)).view(np.complex64)
    return np.fromfile(path, dtype=np.complex64)


def write_bin(path, iq):
    np.asarray(iq, dtype=np.complex64).tofile(path)


class _Block:
    def __init__(self, name):
        self.name = name
        self.tags = []          # one dict per frame, keys as in the reference
        self.offsets = []       # absolute stream item each tag set sits on (add_item_tag's offset)


class trigger(_Block):
    def __init__(self):
        super().__init__("trigger")


class sync(_Block):
    def __init__(self):
        super().__init__("sync")


class signal(_Block):
    def __init__(self):
        super().__init__("signal")
        self.seq = 0            # d_nPktSeq: wraps at 1e9 (lib/signal_impl.cc:130-134)


class signal2(signal):
    def __init__(self):
        super().__init__()
        self.name = "signal2"


class demod(_Block):
    def __init__(self, mupos=0, mugid=2):
        super().__init__("demod")
        self.mupos, self.mugid = int(mupos), int(mugid)


class demod2(_Block):
    def __init__(self):
        super().__init__("demod2")


class decode(_Block):
    """decode(ifdebug): message port `out` -> self.out (list of bytes), optional callback, optional UDP client."""

    def __init__(self, ifdebug=False, udp=None, on_pdu=None, printer=print):
        super().__init__("decode")
        self.d_debug = bool(ifdebug)
        self.d_nPktCorrect = 0
        self.d_legacyMcsCount = [0] * 8
        self.d_htMcsCount = [0] * 8
        self.d_vhtMcsCount = [0] * 10
        self.out, self.on_pdu, self.printer = [], on_pdu, printer
        self.debug_lines = []
        self.sock, self.udp = None, udp
        if udp is not None:
            self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)

    def _say(self, s):
        self.debug_lines.append(s)
        if self.printer:
            self.printer(s)

    def _line(self, fmt, ok, f):
        if fmt == 2:
            s = "ieee80211 decode, vht crc32 %s, total:%d" % ("correct" if ok else "wrong", self.d_nPktCorrect)
            s += "".join(",%d:%d" % (i, c) for i, c in enumerate(self.d_vhtMcsCount))
            s += ",cfo:%f,snr:%f,rssi:%f,sssnr0:%f,sssnr1:%f" % (f["cfo_hz"], f["snr"], f["rssi"], f["sssnr0"], f["sssnr1"])
        else:
            cnt = self.d_legacyMcsCount if fmt == 0 else self.d_htMcsCount
            s = "ieee80211 decode, %s crc32 %s, total:%d" % ("legacy" if fmt == 0 else "ht", "correct" if ok else "wrong", self.d_nPktCorrect)
            s += "".join(",%d:%d" % (i, c) for i, c in enumerate(cnt))
            s += ",cfo:%f,snr:%f,rssi:%f" % (f["cfo_hz"], f["snr"], f["rssi"])
        self._say(s)

    def handle(self, f, pdu_bytes):
        """one frame record + its PDU area -> counters, debug lines, messages (lib/decode_impl.cc:325-520)"""
        fmt, mcs = int(f["format"]), int(f["mcs"])
        recs = split_pdus(pdu_bytes)
        if not recs:
            if self.d_debug and not (fmt == 1 and f["ampdu"]):
                self._line(fmt, False, f)
            return
        for r in recs:
            if self.d_debug:                                   # counters only advance with ifdebug (:393-399,482-494)
                self.d_nPktCorrect += 1
                if fmt == 2:
                    if 0 <= mcs < 10:
                        self.d_vhtMcsCount[mcs] += 1
                elif fmt == 0:
                    self.d_legacyMcsCount[mcs % 8] += 1
                else:
                    self.d_htMcsCount[mcs % 8] += 1
                self._line(fmt, True, f)
            self.publish(r)

    def publish(self, r):
        """message port `out`"""
        self.out.append(r)
        if self.on_pdu:
            self.on_pdu(r)
        if self.sock is not None:
            self.sock.sendto(r, self.udp)


def _tags_of(f, chan, seq):
    sy = {"rad": float(f["rad"]), "snr": float(f["snr"]), "rssi": float(f["rssi"])}
    sg = {"cfo": float(f["cfo_hz"]), "snr": float(f["snr"]), "rssi": float(f["rssi"]), "seq": seq, "mcs": int(f["l_mcs"]),
          "len": int(f["l_len"]), "nsamp": int(f["nsamp"])}
    if chan is not None:                                        # signal -> demod tag; internal to the fused path when streaming
        sg["chan"] = np.asarray(chan, np.complex64)
    dm = {"cfo": float(f["cfo_hz"]), "snr": float(f["snr"]), "rssi": float(f["rssi"]), "format": int(f["format"]), "mcs": int(f["mcs"]),
          "len": int(f["len"]), "cr": int(f["cr"]), "ampdu": int(f["ampdu"]), "trellis": int(f["trellis"]), "total": int(f["total"])}
    if f["format"] == 2:
        dm["sssnr0"] = float(f["sssnr0"])
        if f["nss"] == 2:
            dm["sssnr1"] = float(f["sssnr1"])
    return sy, sg, dm


class rx_top_block:
    """examples/rx.grc (nant=1: presiso -> trigger -> sync -> signal -> demod -> decode) or rx2.grc (nant=2: signal2 ->
    demod2).  run(capture) = tb.run() on a file_source: processes the whole capture, fills the blocks' tags, publishes PDUs."""

    def __init__(self, nant=1, ifdebug=False, mupos=0, mugid=2, udp=None, on_pdu=None, max_frames=None, device=0, printer=print, blob=None):
        self.nant = nant
        self.trigger, self.sync = trigger(), sync()
        self.signal = signal() if nant == 1 else signal2()
        self.demod = demod(mupos, mugid) if nant == 1 else demod2()
        self.decode = decode(ifdebug, udp=udp, on_pdu=on_pdu, printer=printer)
        # frame records per pass.  None: sized by the capture in run() (a frame is at least 400 samples, at most 4096 records) and
        # 512 per window pass in work() -- the library reserves soft-bit scratch per record, so a fixed 4096 costs gigabytes
        self.max_frames = max_frames
        self._rx, self._rx_mf = None, None
        self._rx_args = dict(device=device, chunk_items=1, mupos=mupos, mugid=mugid, blob=blob)
        self.frames = None
        self.truncated = False
        self._streaming = False

    def _receiver(self, max_frames):
        if self._rx is None or self._rx_mf != max_frames:
            if self._rx is not None:
                self._rx.close()
            self._rx, self._rx_mf = Receiver(max_frames=max_frames, **self._rx_args), max_frames
        return self._rx

    @property
    def rx(self):
        return self._receiver(self._rx_mf or self.max_frames or 512)

    def close(self):
        if self._rx is not None:
            self._rx.close()
            self._rx = None

    def run(self, capture0, capture1=None, pdu_stride=4400):
        x0 = read_bin(capture0) if isinstance(capture0, str) else np.asarray(capture0, np.complex64)
        if self.nant == 2:
            x1 = read_bin(capture1) if isinstance(capture1, str) else np.asarray(capture1, np.complex64)
            n = min(x0.size, x1.size)
            mf = self.max_frames or min(4096, n // 400 + 1)
            rx = self._receiver(mf)
            fr, pdu = rx.rx_batch2(x0[:n], x1[:n], [0], [n], pdu_stride=pdu_stride)
            _, chan = rx.detect(x0[:n], [0], [n])
        else:
            mf = self.max_frames or min(4096, x0.size // 400 + 1)
            rx = self._receiver(mf)
            fr, pdu = rx.rx_batch(x0, [0], [x0.size], pdu_stride=pdu_stride)
            _, chan = rx.detect(x0, [0], [x0.size])
        keep = fr["status"] != 9                                # C8B_ST_EMPTY
        self.frames = fr[keep]
        # the scan of an item stops when its frame records are used: say so instead of losing the rest silently
        self.truncated = bool(mf > 1 and keep.sum() >= mf)
        if self.truncated:
            self.decode.printer("ieee80211 rx: all %d frame records of the capture are used, later frames were not examined "
                                "(raise max_frames, or feed the capture through work())" % mf)
        for k in np.nonzero(keep)[0]:
            self._publish(fr[k], chan[k], pdu[k], 0)
        return self.frames

    def _publish(self, f, chan, pdu, base):
        if f["nsamp"] == 0:                                     # no frame accepted by L-SIG in the capture
            return
        sy, sg, dm = _tags_of(f, chan, self.signal.seq)
        self.sync.offsets.append(int(base) + int(f["sync_idx"]))           # nitems_written(0) + idx  (lib/sync_impl.cc:124-136)
        self.signal.offsets.append(int(base) + int(f["sync_idx"]))
        self.signal.seq = (self.signal.seq + 1) % 1000000000
        self.sync.tags.append(sy)
        self.signal.tags.append(sg)
        if f["status"] == 7:                                    # VHT NDP: tag mu2x1chan, total 1024 (lib/demod_impl.cc:238-249)
            dm["total"] = 1024
            if f["pdu_bytes"] == 1027:
                dm["mu2x1chan"] = np.frombuffer(bytes(pdu[3:1027]), np.float32).view(np.complex64).copy()
        if f["status"] in (0, 6, 7):                            # reached WRTAG (decode may still reject: DECODE_RANGE)
            self.demod.tags.append(dm)
        if f["status"] == 0:
            self.decode.handle(f, pdu[:f["pdu_bytes"]])
        elif f["status"] == 7 and f["pdu_bytes"] == 1027:       # channel report blob, no counters / debug line (decode_impl.cc:100-121)
            self.decode.publish(bytes(pdu[:1027]))

    def work(self, x0, x1=None, flush=False, window=0):
        """the scheduler's general_work calls: feed the next piece of the capture (any size); frames are published as soon
        as they are decidable, identically to run() over the whole capture.  flush=True ends the stream."""
        mf = self.max_frames or 512
        rx = self._receiver(mf) if not self._streaming else self._rx
        if not self._streaming:
            rx.stream_begin(self.nant, window)
            self._streaming = True
        fr, base, pdu = rx.stream_push(np.asarray(x0, np.complex64), None if x1 is None else np.asarray(x1, np.complex64),
                                       flush=flush, frames_cap=max(mf, 64) + len(x0) // 320)    # a frame is at least 400 samples
        for k in range(fr.size):
            self._publish(fr[k], None, pdu[k], base[k])
        if flush:
            self._streaming = False
        return fr, base
