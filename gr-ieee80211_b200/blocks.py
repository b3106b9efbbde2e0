"""The seven receive blocks one scheduler call at a time (c8b_blk_* of include/c80211b200.h) and a miniature of the GNU
Radio runtime to drive them: stream buffers with read / write counters, tags at absolute item offsets, fan-out, random
call sizes.  `Chain` wires them like examples/rx.grc:753-767 (nant = 1) or examples/rx2.grc:676-692 (nant = 2):

    preac ─ trigger ─┐
    preconj ─────────┤ sync ─┐
    sig ─────────────┘       ├ signal[2] ─ demod[2] ─ decode ─> PDU messages
    sig [, sig1] ────────────┘

This is what the compiled gr::block shells (gr/lib/*_impl.cc) do through the same entry points; here the "scheduler" is a
loop, which is all the parity tests need (tests/test_gpu_blocks.py; tests/test_host_logic.py runs the same state machines
over the host build of the per-frame routines)."""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import C8bCfg, C8bError, TAG_DTYPE, ptr

TRIGGER, SYNC, SIGNAL, SIGNAL2, DEMOD, DEMOD2, DECODE = range(7)
NAMES = ("trigger", "sync", "signal", "signal2", "demod", "demod2", "decode")
_IN_DT = {TRIGGER: ["<f4"], SYNC: ["u1", "<c8", "<c8"], SIGNAL: ["u1", "<c8"], SIGNAL2: ["u1", "<c8", "<c8"], DEMOD: ["<c8"],
          DEMOD2: ["<c8", "<c8"], DECODE: ["<f4"]}
_OUT_DT = {TRIGGER: ["u1"], SYNC: ["u1"], SIGNAL: ["<c8"], SIGNAL2: ["<c8", "<c8"], DEMOD: ["<f4"], DEMOD2: ["<f4"], DECODE: []}


class LibBackend:
    """c8b_blk_* of libc80211b200.so (the product)."""

    def __init__(self, device=0):
        self.L = _cabi.lib()
        self.device = device

    def create(self, kind, mupos=0, mugid=0):
        if self.L.c8b_device_count() <= 0:
            raise C8bError("no CUDA device visible: gr-ieee80211_b200 has no CPU path")
        cfg = C8bCfg(device=self.device, mupos=mupos, mugid=mugid)
        h = C.c_void_p()
        rc = self.L.c8b_blk_create(C.byref(cfg), kind, C.byref(h))
        if rc:
            raise C8bError("c8b_blk_create(%s): %d %s" % (NAMES[kind], rc, self.L.c8b_blk_last_error(None).decode()))
        return h

    def destroy(self, h):
        self.L.c8b_blk_destroy(h)

    def forecast(self, kind, noutput):
        return self.L.c8b_blk_forecast(kind, noutput)

    def work(self, h, *args):
        rc = self.L.c8b_blk_work(h, *args)
        if rc:
            raise C8bError("c8b_blk_work: %d %s" % (rc, self.L.c8b_blk_last_error(h).decode()))


class Block:
    """One block instance: work() = one general_work() call."""

    def __init__(self, kind, backend, mupos=0, mugid=0):
        self.kind, self.be = kind, backend
        self.h = backend.create(kind, mupos, mugid)
        self.in_dt, self.out_dt = _IN_DT[kind], _OUT_DT[kind]
        self.calls = 0

    def close(self):
        if self.h is not None:
            self.be.destroy(self.h)
            self.h = None

    def forecast(self, noutput):
        return self.be.forecast(self.kind, noutput)

    def work(self, noutput, ins, in_tags):
        """ins: one array per input port (the items available); in_tags: TAG_DTYPE array, idx relative to ins[0].
        Returns (consumed, outs[:produced], out_tags, message bytes)."""
        nin = len(self.in_dt)
        ins = [np.ascontiguousarray(a, dtype=dt) for a, dt in zip(ins, self.in_dt)]
        ninput = (C.c_int * nin)(*[a.size for a in ins])
        inp = (C.c_void_p * nin)(*[a.ctypes.data for a in ins])
        outs = [np.zeros(max(noutput, 1), dt) for dt in self.out_dt]
        outp = (C.c_void_p * max(len(outs), 1))(*[a.ctypes.data for a in outs])
        in_tags = np.ascontiguousarray(in_tags, dtype=TAG_DTYPE)
        out_tags = np.zeros(8, TAG_DTYPE)
        msg = np.zeros(1 << 16, np.uint8)
        consumed, produced, ntags, nmsg = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self.be.work(self.h, noutput, ninput, inp, outp, ptr(in_tags) if in_tags.size else None, in_tags.size, C.byref(consumed),
                     C.byref(produced), ptr(out_tags), out_tags.size, C.byref(ntags), ptr(msg), msg.size, C.byref(nmsg))
        self.calls += 1
        return consumed.value, [a[:produced.value] for a in outs], out_tags[:ntags.value].copy(), bytes(msg[:nmsg.value])


class Stream:
    """One edge of the flowgraph: the items not yet consumed, absolute counters, tags at absolute offsets."""

    def __init__(self, dt):
        self.buf = np.zeros(0, dt)
        self.nread = 0          # nitems_read of the consumer
        self.nwritten = 0       # nitems_written of the producer
        self.tags = []          # (absolute offset, TAG_DTYPE record)

    def push(self, items, tags=()):
        for off, t in tags:
            self.tags.append((self.nwritten + off, t))
        self.buf = np.concatenate([self.buf, np.asarray(items, self.buf.dtype)])
        self.nwritten += len(items)

    def avail(self):
        return self.buf.size

    def window(self, n):
        tg = [(o - self.nread, t) for o, t in self.tags if self.nread <= o < self.nread + n]
        return self.buf[:n], tg

    def consume(self, n):
        self.buf = self.buf[n:]
        self.nread += n
        self.tags = [(o, t) for o, t in self.tags if o >= self.nread]


def split_messages(buf):
    """decode's message bytes -> list of PDU records; an NDP report is [20][0][4] + 1024 bytes (n = len + 3)."""
    out, i = [], 0
    while i + 3 <= len(buf):
        ln = buf[i + 1] | (buf[i + 2] << 8)
        n = ln + 3 if buf[i] == 20 else ln + 4
        out.append(buf[i:i + n])
        i += n
    return out


class Chain:
    """rx.grc / rx2.grc as blocks + streams.  run() feeds whole arrays and plays the scheduler with random call sizes."""

    def __init__(self, nant=1, backend=None, mupos=0, mugid=0, seed=0, max_call=4096):
        be = backend if backend is not None else LibBackend()
        self.nant = nant
        self.rng = np.random.default_rng(seed)
        self.max_call = max_call
        kinds = [TRIGGER, SYNC, SIGNAL2 if nant == 2 else SIGNAL, DEMOD2 if nant == 2 else DEMOD, DECODE]
        self.blocks = [Block(k, be, mupos, mugid) for k in kinds]
        trig, sync, sig, dem, dec = self.blocks
        S = Stream
        self.src = {"preac": [S("<f4")], "preconj": [S("<c8")], "sig": [S("<c8"), S("<c8")], "sig1": [S("<c8")] if nant == 2 else []}
        e_ts, e_ss = S("u1"), S("u1")
        e_sd = [S("<c8") for _ in range(nant)]
        e_dd = S("<f4")
        self.ins = {id(trig): [self.src["preac"][0]], id(sync): [e_ts, self.src["preconj"][0], self.src["sig"][0]],
                    id(sig): [e_ss, self.src["sig"][1]] + self.src["sig1"], id(dem): e_sd, id(dec): [e_dd]}
        self.outs = {id(trig): [e_ts], id(sync): [e_ss], id(sig): e_sd, id(dem): [e_dd], id(dec): []}
        self.trace = {n: [] for n in ("trigger", "sync", "signal", "signal1", "llr")}   # whole output streams
        self.tags = {n: [] for n in ("sync", "signal", "demod", "decode")}                         # (absolute offset, record)
        self.messages = []

    def close(self):
        for b in self.blocks:
            b.close()

    def _call(self, b, big):
        ins = self.ins[id(b)]
        avail = [s.avail() for s in ins]
        cap = self.max_call if big else int(self.rng.integers(1, self.max_call + 1))
        if b.kind == DECODE:
            noutput, nin = cap, [min(avail[0], b.forecast(cap))]
        else:
            noutput = min([cap] + avail)           # forecast is 1:1: the scheduler shrinks noutput to what the inputs allow
            extra = 0 if big else int(self.rng.integers(0, 64))
            nin = [min(a, noutput + extra) for a in avail]
            if b.kind in (DEMOD, DEMOD2) and noutput == 0:
                noutput = cap                      # a block with pending output is also called when only output space changed
        if noutput <= 0 and max(nin) <= 0:
            return False
        wins = [s.window(n) for s, n in zip(ins, nin)]
        tg = np.zeros(len(wins[0][1]), TAG_DTYPE)
        for k, (o, t) in enumerate(wins[0][1]):
            tg[k] = t
            tg[k]["port"], tg[k]["idx"] = 0, o
        consumed, outs, otags, msg = b.work(noutput, [w[0] for w in wins], tg)
        assert 0 <= consumed <= min(nin) and all(len(o) <= noutput for o in outs)
        for s in ins:
            s.consume(consumed)
        name = NAMES[b.kind].rstrip("2")
        dst = self.outs[id(b)]
        base = dst[0].nwritten if dst else 0
        for q, o in enumerate(outs):
            dst[q].push(o, [(int(t["idx"]), t.copy()) for t in otags] if q == 0 else ())
            self.trace[{"demod": "llr"}.get(name, name) + ("1" if q else "")].append(o.copy())
        for t in otags:
            self.tags[name].append((base + int(t["idx"]), t.copy()))
        if msg:
            self.messages += split_messages(msg)
        return consumed > 0 or any(len(o) for o in outs) or len(otags) > 0 or bool(msg)

    def run(self, preac, preconj, sig, sig1=None):
        """Push the source arrays (presiso's outputs + the capture) and call the blocks until nothing moves."""
        self.src["preac"][0].push(preac)
        self.src["preconj"][0].push(preconj)
        for s in self.src["sig"]:
            s.push(sig)
        if self.nant == 2:
            self.src["sig1"][0].push(sig1)
        idle = 0
        while idle < 2:
            moved = False
            for b in self.blocks:
                for _ in range(int(self.rng.integers(1, 4))):
                    moved |= self._call(b, big=idle > 0)
            idle = 0 if moved else idle + 1
        return self.messages
