"""Receiver: the batched, device-side counterpart of the reference flowgraph examples/rx.grc:753-767
(file source -> presiso -> trigger -> sync -> signal -> demod -> decode -> PDUs), plus the staged
entry points mirroring one reference block each (used by the parity tests and for profiling)."""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import C8bCfg, C8bError, FRAME_DTYPE, TXFRAME_DTYPE, TXMU_DTYPE, ptr


def lut_blob():
    """The lookup-table blob (uint8 array), built by formula on the host (csrc/lut.cc)."""
    L = _cabi.lib()
    n = L.c8b_lut_size()
    buf = np.zeros(n, np.uint8)
    rc = L.c8b_lut_blob(ptr(buf), n)
    if rc:
        raise C8bError("c8b_lut_blob failed: %d" % rc)
    return buf


def _producer_sync():
    """Device-pointer entry points run on the context's own (non-blocking) stream: if the caller produced the buffer with
    torch, wait for torch's current stream first (the library knows nothing about the producer's stream)."""
    import sys
    t = sys.modules.get("torch")
    if t is not None and t.cuda.is_available():
        t.cuda.current_stream().synchronize()


def _c2f(x):
    x = np.asarray(x)
    if x.dtype == np.complex64:
        return np.ascontiguousarray(x).view(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def split_pdus(buf):
    """PDU area -> list of records [fmt][len lo][len hi][MPDU][mcs] (lib/decode_impl.cc:359-361,414-419)."""
    out, i, buf = [], 0, bytes(buf)
    while i + 3 <= len(buf):
        ln = buf[i + 1] | (buf[i + 2] << 8)
        out.append(buf[i: i + ln + 4])
        i += ln + 4
    return out


class Receiver:
    """One context = one GPU + one stream.  `blob`: LUT blob to load (default: build locally; multi-GPU
    runs pass the blob broadcast from rank 0)."""

    def __init__(self, device=0, chunk_items=0, max_item_len=0, max_frames=1, mupos=0, mugid=0, blob=None, overlap=True, decode_mode=0, frontend_mode=0, mmse=0):
        self.L = _cabi.lib()
        if self.L.c8b_device_count() <= 0:
            raise C8bError("no CUDA device visible: gr-ieee80211_b200 has no CPU path")
        cfg = C8bCfg(device=device, chunk_items=chunk_items, max_item_len=max_item_len, max_frames=max_frames, mupos=mupos, mugid=mugid, no_overlap=0 if overlap else 1, decode_mode=decode_mode, frontend_mode=frontend_mode, mmse=mmse)
        self.max_frames = max(1, int(max_frames))
        h = C.c_void_p()
        rc = self.L.c8b_create(C.byref(cfg), C.byref(h))
        if rc:
            raise C8bError("c8b_create: %d %s" % (rc, (self.L.c8b_last_error(None) or b"").decode()))
        self.h = h
        blob = lut_blob() if blob is None else np.ascontiguousarray(blob, dtype=np.uint8)
        self._ck(self.L.c8b_lut_load(self.h, ptr(blob), blob.size), "c8b_lut_load")

    def close(self):
        if getattr(self, "h", None):
            self.L.c8b_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc:
            raise C8bError("%s: %d %s" % (what, rc, (self.L.c8b_last_error(self.h) or b"").decode()))

    @property
    def stream(self):
        return self.L.c8b_stream(self.h)

    def load_lut_device(self, dev_ptr, nbytes):
        self._ck(self.L.c8b_lut_load_dev(self.h, C.c_void_p(dev_ptr), nbytes), "c8b_lut_load_dev")

    def sync(self):
        self._ck(self.L.c8b_sync(self.h), "c8b_sync")

    # ---- timing -------------------------------------------------------------------------------
    def timing(self, on=True):
        self._ck(self.L.c8b_timing_enable(self.h, int(on)), "c8b_timing_enable")

    def timing_read(self, reset=True):
        ms = np.zeros(len(_cabi.K_NAMES), np.float64)
        n = np.zeros(len(_cabi.K_NAMES), np.int64)
        self._ck(self.L.c8b_timing_read(self.h, ptr(ms), ptr(n), int(reset)), "c8b_timing_read")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(_cabi.K_NAMES)}

    @staticmethod
    def _items(off, length, nsamples):
        """item table as the C ABI wants it, checked against the capture it indexes: the C ABI has no capture size argument,
        so an item past the end of the array would be a host out-of-bounds read"""
        off = np.ascontiguousarray(off, np.int64)
        length = np.ascontiguousarray(length, np.int32)
        if off.shape != length.shape or off.ndim != 1:
            raise ValueError("off / length: two 1-D arrays of one size")
        if off.size and (off.min() < 0 or length.min() < 0 or (off + length).max() > nsamples):
            raise ValueError("item outside the capture (%d samples)" % nsamples)
        return off, length

    # ---- whole chain --------------------------------------------------------------------------
    def rx_batch(self, iq, off, length, pdu_stride=4400):
        """iq: complex64 host array; item i = iq[off[i]:off[i]+length[i]].  Returns (frames, pdu)."""
        iqf = _c2f(iq)
        off, length = self._items(off, length, iqf.size // 2)
        n, ns = off.size, off.size * self.max_frames
        frames = np.zeros(ns, FRAME_DTYPE)
        pdu = np.zeros(ns * pdu_stride, np.uint8)
        self._ck(self.L.c8b_rx_batch(self.h, ptr(iqf), ptr(off), ptr(length), n, ptr(frames), ptr(pdu), pdu_stride), "c8b_rx_batch")
        return frames, pdu.reshape(ns, pdu_stride)

    def rx_batch_sc16(self, iq16, off, length, pdu_stride=4400):
        """iq16: int16 array of interleaved (I, Q) pairs (UHD sc16 wire format), shape [n, 2] or [2 n]; off / length in samples"""
        q = np.ascontiguousarray(iq16, np.int16).reshape(-1)
        off, length = self._items(off, length, q.size // 2)
        n, ns = off.size, off.size * self.max_frames
        frames = np.zeros(ns, FRAME_DTYPE)
        pdu = np.zeros(ns * pdu_stride, np.uint8)
        self._ck(self.L.c8b_rx_batch_sc16(self.h, ptr(q), ptr(off), ptr(length), n, ptr(frames), ptr(pdu), pdu_stride), "c8b_rx_batch_sc16")
        return frames, pdu.reshape(ns, pdu_stride)

    def rx_batch2(self, iq0, iq1, off, length, pdu_stride=4400):
        """2x2: antenna 0 drives detection, both antennas are demodulated (signal2 + demod2)."""
        a, b = _c2f(iq0), _c2f(iq1)
        if a.size != b.size:
            raise ValueError("the two antennas' captures differ in length")
        off, length = self._items(off, length, a.size // 2)
        n, ns = off.size, off.size * self.max_frames
        frames = np.zeros(ns, FRAME_DTYPE)
        pdu = np.zeros(ns * pdu_stride, np.uint8)
        self._ck(self.L.c8b_rx_batch2(self.h, ptr(a), ptr(b), ptr(off), ptr(length), n, ptr(frames), ptr(pdu), pdu_stride), "c8b_rx_batch2")
        return frames, pdu.reshape(ns, pdu_stride)

    # ---- live stream ---------------------------------------------------------------------------
    def stream_begin(self, nant=1, window=0):
        """start a session: the capture arrives in pieces (a block shell's general_work calls); needs max_frames >= 2"""
        self._ck(self.L.c8b_stream_begin(self.h, nant, window), "c8b_stream_begin")
        self._stream_nant = nant

    def stream_push(self, iq0, iq1=None, flush=False, frames_cap=None, pdu_stride=4400):
        """append samples; returns (frames, base, pdu): the frames that became decidable, the absolute stream index
        their trig_idx / sync_idx are relative to, and their PDU records"""
        a = _c2f(iq0)
        b = _c2f(iq1) if iq1 is not None else None
        n = a.size // 2
        if frames_cap is None:
            frames_cap = n // 400 + 4 * self.max_frames + 64     # more than one push can release (see the header)
        frames = np.zeros(frames_cap, FRAME_DTYPE)
        base = np.zeros(frames_cap, np.int64)
        pdu = np.zeros(frames_cap * pdu_stride, np.uint8)
        nf = C.c_int(0)
        self._ck(self.L.c8b_stream_push(self.h, ptr(a) if n else None, ptr(b) if (b is not None and n) else None, n, 1 if flush else 0,
                                        ptr(frames), frames_cap, C.byref(nf), ptr(base), ptr(pdu), pdu_stride), "c8b_stream_push")
        k = nf.value
        return frames[:k].copy(), base[:k].copy(), pdu.reshape(frames_cap, pdu_stride)[:k].copy()

    def stream_state(self):
        b, f, o = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.L.c8b_stream_state(self.h, C.byref(b), C.byref(f), C.byref(o)), "c8b_stream_state")
        return {"base": b.value, "fill": f.value, "overruns": o.value}

    # ---- transmit synthesiser -------------------------------------------------------------------
    def tx_layout(self, psdus, fmt, mcs, gap=400, cfo=None):
        """descriptors for one item per PSDU laid out [gap zeros][frame][gap zeros] (genFinalSig gap=True): returns
        (descriptor array, PSDU arena, item offsets)"""
        n = len(psdus)
        fmt = np.broadcast_to(np.asarray(fmt, np.int32), (n,))
        mcs = np.broadcast_to(np.asarray(mcs, np.int32), (n,))
        gap = np.broadcast_to(np.asarray(gap, np.int64), (n,))
        cfo = np.zeros(n, np.float32) if cfo is None else np.broadcast_to(np.asarray(cfo, np.float32), (n,))
        d = np.zeros(n, TXFRAME_DTYPE)
        offs = np.zeros(n + 1, np.int64)
        po = 0
        for i, p in enumerate(psdus):
            ns = self.L.c8b_tx_nsamp(int(fmt[i]), int(mcs[i]), len(p))
            if ns < 0:
                raise ValueError("unsupported format / mcs / length: %d %d %d" % (fmt[i], mcs[i], len(p)))
            d[i] = (fmt[i], mcs[i], po, len(p), cfo[i], offs[i] + gap[i])
            offs[i + 1] = offs[i] + 2 * gap[i] + ns
            po += len(p)
        arena = np.frombuffer(b"".join(bytes(p) for p in psdus) + b"\0", np.uint8).copy()
        return d, arena, offs

    def tx_batch(self, psdus, fmt, mcs, gap=400, cfo=None, multiplier=12.0, seed=93):
        """waveforms of tools/phy80211.py genFromMpdu / genFromAmpdu + genFinalSig(multiplier, cfoHz, gap): (iq, item offsets)"""
        d, arena, offs = self.tx_layout(psdus, fmt, mcs, gap, cfo)
        iq = np.zeros(int(offs[-1]), np.complex64)
        self._ck(self.L.c8b_tx_batch(self.h, ptr(arena), arena.size, ptr(d), d.size, multiplier, seed, ptr(iq.view(np.float32)), iq.size), "c8b_tx_batch")
        return iq, offs

    def tx_batch2(self, psdus, fmt, mcs, gap=400, cfo=None, multiplier=12.0 * 2 ** 0.5, seed=93):
        """two antennas: HT mcs 8-15 / VHT mcs 16 + MCS are two-stream frames (stream k on antenna k); returns (iq0, iq1, item offsets)"""
        d, arena, offs = self.tx_layout(psdus, fmt, mcs, gap, cfo)
        iq0 = np.zeros(int(offs[-1]), np.complex64)
        iq1 = np.zeros(int(offs[-1]), np.complex64)
        self._ck(self.L.c8b_tx_batch2(self.h, ptr(arena), arena.size, ptr(d), d.size, multiplier, seed, ptr(iq0.view(np.float32)),
                                      ptr(iq1.view(np.float32)), iq0.size), "c8b_tx_batch2")
        return iq0, iq1, offs

    def tx_batch2_dev(self, d_psdu_ptr, psdu_bytes, desc, d_iq0_ptr, d_iq1_ptr, iq_samples, multiplier=12.0 * 2 ** 0.5, seed=93):
        _producer_sync()
        self._ck(self.L.c8b_tx_batch2_dev(self.h, C.c_void_p(d_psdu_ptr), psdu_bytes, ptr(desc), desc.size, multiplier, seed,
                                          C.c_void_p(d_iq0_ptr), C.c_void_p(d_iq1_ptr), iq_samples), "c8b_tx_batch2_dev")

    def tx_mu_batch(self, ampdus, mcs, q, group_id=2, gap=400, cfo=None, multiplier=18.0, seed=93):
        """two-user VHT MU-MIMO frames (tools/phy80211.py genAmpduMu + genFinalSig): ampdus = [(A-MPDU of user 0, of user 1), ...],
        mcs = [(mcs0, mcs1), ...], q = complex array (64, 2, 2) -- or (nq, 64, 2, 2) with one set per frame -- of spatial mapping
        matrices, subcarriers -32 .. 31, [antenna, user].  Returns (iq0, iq1, item offsets)."""
        n = len(ampdus)
        q = np.ascontiguousarray(np.asarray(q, np.complex64).reshape(-1, 64, 2, 2))
        cfo = np.zeros(n, np.float32) if cfo is None else np.broadcast_to(np.asarray(cfo, np.float32), (n,))
        d = np.zeros(n, TXMU_DTYPE)
        offs = np.zeros(n + 1, np.int64)
        po, parts = 0, []
        for i, (a0, a1) in enumerate(ampdus):
            ns = self.L.c8b_tx_mu_nsamp(int(mcs[i][0]), len(a0), int(mcs[i][1]), len(a1))
            if ns < 0:
                raise ValueError("unsupported MU frame: mcs %s, A-MPDU lengths %d / %d" % (mcs[i], len(a0), len(a1)))
            d[i]["mcs"] = mcs[i]
            d[i]["psdu_len"] = (len(a0), len(a1))
            d[i]["psdu_off"] = (po, po + len(a0))
            d[i]["group_id"], d[i]["cfo_hz"], d[i]["out_off"], d[i]["q_index"] = group_id, cfo[i], offs[i] + gap, i if q.shape[0] == n and n > 1 else 0
            offs[i + 1] = offs[i] + 2 * gap + ns
            po += len(a0) + len(a1)
            parts += [bytes(a0), bytes(a1)]
        arena = np.frombuffer(b"".join(parts) + b"\0", np.uint8).copy()
        iq0 = np.zeros(int(offs[-1]), np.complex64)
        iq1 = np.zeros(int(offs[-1]), np.complex64)
        self._ck(self.L.c8b_tx_mu_batch(self.h, ptr(arena), arena.size, ptr(d), d.size, ptr(q.view(np.float32)), q.shape[0], multiplier, seed,
                                        ptr(iq0.view(np.float32)), ptr(iq1.view(np.float32)), iq0.size), "c8b_tx_mu_batch")
        return iq0, iq1, offs

    def tx_from_udp(self, datagrams, gap=400, multiplier=12.0, seed=93):
        """MAC -> PHY datagrams [format][mcs][nss][len16 LE][PSDU] (lib/pktgen_impl.cc:57-70, tools/phy80211.py genPktGrData) ->
        (iq, descriptors): one frame per accepted datagram, `gap` zeros around each; a skipped datagram has psdu_len -1"""
        blob = np.frombuffer(b"".join(bytes(d) for d in datagrams) + b"\0", np.uint8).copy()
        ln = np.array([len(d) for d in datagrams], np.int32)
        off = np.concatenate([[0], np.cumsum(ln[:-1])]).astype(np.int64) if len(datagrams) else np.zeros(0, np.int64)
        cap = int(sum(2 * gap + 80 * (8 + (len(d) * 8 + 22) // 24 + 2) for d in datagrams)) + gap + 1024
        iq = np.zeros(cap, np.complex64)
        desc = np.zeros(max(len(datagrams), 1), TXFRAME_DTYPE)
        used = C.c_int64(0)
        rc = self.L.c8b_tx_from_udp(self.h, ptr(blob), ptr(off), ptr(ln), len(datagrams), gap, multiplier, seed, ptr(iq.view(np.float32)),
                                    iq.size, C.byref(used), ptr(desc))
        if rc < 0:
            self._ck(rc, "c8b_tx_from_udp")
        return iq[:used.value].copy(), desc[:len(datagrams)]

    def tx_random_psdu_dev(self, d_psdu_ptr, psdu_bytes, desc, seed=1):
        _producer_sync()
        self._ck(self.L.c8b_tx_random_psdu_dev(self.h, C.c_void_p(d_psdu_ptr), psdu_bytes, ptr(desc), desc.size, seed), "c8b_tx_random_psdu_dev")

    def tx_batch_dev(self, d_psdu_ptr, psdu_bytes, desc, d_iq_ptr, iq_samples, multiplier=12.0, seed=93):
        _producer_sync()
        self._ck(self.L.c8b_tx_batch_dev(self.h, C.c_void_p(d_psdu_ptr), psdu_bytes, ptr(desc), desc.size, multiplier, seed,
                                         C.c_void_p(d_iq_ptr), iq_samples), "c8b_tx_batch_dev")

    def rx_batch_dev(self, d_iq_ptr, off, length, pdu_stride=4400, frames=None, pdu=None):
        off = np.ascontiguousarray(off, np.int64)
        length = np.ascontiguousarray(length, np.int32)
        n, ns = off.size, off.size * self.max_frames
        frames = np.zeros(ns, FRAME_DTYPE) if frames is None else frames
        pdu = np.zeros(ns * pdu_stride, np.uint8) if pdu is None else pdu
        _producer_sync()
        self._ck(self.L.c8b_rx_batch_dev(self.h, C.c_void_p(d_iq_ptr), ptr(off), ptr(length), n, ptr(frames), ptr(pdu), pdu_stride),
                 "c8b_rx_batch_dev")
        return frames, pdu.reshape(ns, pdu_stride)

    def rx_batch2_dev(self, d_iq0_ptr, d_iq1_ptr, off, length, pdu_stride=4400):
        off = np.ascontiguousarray(off, np.int64)
        length = np.ascontiguousarray(length, np.int32)
        n, ns = off.size, off.size * self.max_frames
        frames = np.zeros(ns, FRAME_DTYPE)
        pdu = np.zeros(ns * pdu_stride, np.uint8)
        _producer_sync()
        self._ck(self.L.c8b_rx_batch2_dev(self.h, C.c_void_p(d_iq0_ptr), C.c_void_p(d_iq1_ptr), ptr(off), ptr(length), n, ptr(frames), ptr(pdu),
                                          pdu_stride), "c8b_rx_batch2_dev")
        return frames, pdu.reshape(ns, pdu_stride)

    def rx_batch_dev_async(self, d_iq_ptr, off, length, d_frames_ptr, d_pdu_ptr, pdu_stride=4400):
        off = np.ascontiguousarray(off, np.int64)
        length = np.ascontiguousarray(length, np.int32)
        _producer_sync()
        self._ck(self.L.c8b_rx_batch_dev_async(self.h, C.c_void_p(d_iq_ptr), ptr(off), ptr(length), off.size, C.c_void_p(d_frames_ptr),
                                               C.c_void_p(d_pdu_ptr), pdu_stride), "c8b_rx_batch_dev_async")

    # ---- staged -------------------------------------------------------------------------------
    def presiso(self, iq, want_conj=True):
        iqf = _c2f(iq)
        n = iqf.size // 2
        preac = np.zeros(n, np.float32)
        preconj = np.zeros(2 * n, np.float32) if want_conj else None
        self._ck(self.L.c8b_presiso(self.h, ptr(iqf), n, ptr(preac), ptr(preconj)), "c8b_presiso")
        return preac, (preconj.view(np.complex64) if want_conj else None)

    def trigger(self, preac):
        preac = np.ascontiguousarray(preac, np.float32)
        out = np.zeros(preac.size, np.uint8)
        self._ck(self.L.c8b_trigger(self.h, ptr(preac), preac.size, ptr(out)), "c8b_trigger")
        return out

    def trigger_events(self, preac, start=0, cap=4096):
        """the warp trigger scan on a preac array: (events [n, 4] = trig, latch, safe, stall; safe_end)"""
        p = np.ascontiguousarray(preac, np.float32)
        ev = np.zeros((cap, 4), np.int32)
        cnt, se = C.c_int(0), C.c_int(0)
        self._ck(self.L.c8b_trigger_events(self.h, ptr(p), p.size, start, ptr(ev), cap, C.byref(cnt), C.byref(se)), "c8b_trigger_events")
        return ev[:cnt.value].copy(), se.value

    def detect(self, iq, off, length):
        iqf = _c2f(iq)
        off, length = self._items(off, length, iqf.size // 2)
        n, ns = off.size, off.size * self.max_frames
        frames = np.zeros(ns, FRAME_DTYPE)
        chan = np.zeros(ns * 128, np.float32)
        self._ck(self.L.c8b_detect(self.h, ptr(iqf), ptr(off), ptr(length), n, ptr(frames), ptr(chan)), "c8b_detect")
        return frames, chan.view(np.complex64).reshape(ns, 64)

    def demod(self, iq, off, length, frames, chan, llr_stride):
        iqf = _c2f(iq)
        off, length = self._items(off, length, iqf.size // 2)
        n, ns = off.size, off.size * self.max_frames
        frames = np.ascontiguousarray(frames).copy()
        assert frames.size == ns
        chanf = _c2f(chan)
        llr = np.zeros(ns * llr_stride, np.float32)
        self._ck(self.L.c8b_demod(self.h, ptr(iqf), ptr(off), ptr(length), n, ptr(frames), ptr(chanf), ptr(llr), llr_stride), "c8b_demod")
        return frames, llr.reshape(ns, llr_stride)

    def demod2(self, iq0, iq1, off, length, frames, chan, llr_stride):
        a, b = _c2f(iq0), _c2f(iq1)
        if a.size != b.size:
            raise ValueError("the two antennas' captures differ in length")
        off, length = self._items(off, length, a.size // 2)
        n, ns = off.size, off.size * self.max_frames
        frames = np.ascontiguousarray(frames).copy()
        assert frames.size == ns
        chanf = _c2f(chan)
        llr = np.zeros(ns * llr_stride, np.float32)
        self._ck(self.L.c8b_demod2(self.h, ptr(a), ptr(b), ptr(off), ptr(length), n, ptr(frames), ptr(chanf), ptr(llr), llr_stride), "c8b_demod2")
        return frames, llr.reshape(ns, llr_stride)

    def decode(self, llr, frames, pdu_stride=4400, want_scram=False, scram_stride=0):
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1)
        frames = np.ascontiguousarray(frames).copy()
        n = frames.size
        pdu = np.zeros(n * pdu_stride, np.uint8)
        scram = None
        if want_scram:
            scram_stride = scram_stride or int(max(1, frames["trellis"].max()))
            scram = np.zeros(n * scram_stride, np.uint8)
        self._ck(self.L.c8b_decode(self.h, ptr(llr), llr.size, ptr(frames), n, ptr(pdu), pdu_stride, ptr(scram), scram_stride), "c8b_decode")
        return frames, pdu.reshape(n, pdu_stride), (scram.reshape(n, scram_stride) if want_scram else None)
