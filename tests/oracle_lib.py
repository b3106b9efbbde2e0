"""ctypes doors onto the CPU oracle (oracle/liboracle_rx.so) and, when present, the compiled
unmodified reference (oracle/_ref/libc8p_ref.so).  TEST INFRASTRUCTURE: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_rx.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libc8p_ref.so")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C")


class OrxFrame(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("item", C.c_int32), ("trig_idx", C.c_int32), ("sync_idx", C.c_int32),
        ("rad", C.c_float), ("snr", C.c_float), ("rssi", C.c_float), ("cfo_hz", C.c_float),
        ("l_mcs", C.c_int32), ("l_len", C.c_int32), ("nsamp", C.c_int32),
        ("format", C.c_int32), ("mcs", C.c_int32), ("len", C.c_int32), ("cr", C.c_int32), ("ampdu", C.c_int32),
        ("nss", C.c_int32), ("nsym", C.c_int32), ("nsymsamp", C.c_int32), ("ncbps", C.c_int32), ("ndbps", C.c_int32),
        ("trellis", C.c_int32), ("total", C.c_int32), ("data_off", C.c_int32),
        ("sssnr0", C.c_float), ("sssnr1", C.c_float),
        ("llr_off", C.c_int64), ("pdu_off", C.c_int64), ("npdu", C.c_int32), ("pdu_bytes", C.c_int32),
    ]


FRAME_DTYPE = np.dtype([(n, {C.c_int32: "<i4", C.c_float: "<f4", C.c_int64: "<i8"}[t]) for n, t in OrxFrame._fields_], align=True)
assert FRAME_DTYPE.itemsize == C.sizeof(OrxFrame)


def build_oracle(force=False):
    """Compile oracle/liboracle_rx.so (and _ref when the reference sources are mounted)."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "oracle_rx.cc")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle_rx.so"])
    if os.path.exists("/root/reference/lib/cloud80211phy.cc") and (force or not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orx_deint_map.argtypes = [C.c_int, C.c_int, C.c_int, i32p]
        L.orx_pilot_polarity.argtypes = [f32p]
        L.orx_ltf.argtypes = [C.c_int, f32p]
        L.orx_trellis_tables.argtypes = [i32p, i32p]
        L.orx_lsig_demap.argtypes = [i32p]
        L.orx_fft64.argtypes = [f32p, f32p]
        L.orx_presiso.argtypes = [f32p, C.c_int, f32p, f32p]
        L.orx_trigger.argtypes = [f32p, C.c_int, u8p, i32p]
        L.orx_sync.argtypes = [f32p, C.c_float, C.c_float] + [C.POINTER(C.c_int)] + [C.POINTER(C.c_float)] * 3 + [f32p]
        L.orx_signal.argtypes = [f32p, C.c_float, f32p, f32p, u8p] + [C.POINTER(C.c_int)] * 3
        L.orx_lsig_demod.argtypes = [f32p] * 5
        L.orx_nlsig_demod.argtypes = [f32p] * 5
        L.orx_sig_viterbi.argtypes = [f32p, u8p, C.c_int]
        L.orx_crc8_check.argtypes = [u8p, C.c_int, u8p]
        L.orx_check_legacy.argtypes = [u8p] + [C.POINTER(C.c_int)] * 3
        L.orx_check_ht.argtypes = [u8p]
        L.orx_check_vhta.argtypes = [u8p]
        L.orx_parse_l.argtypes = [C.c_int, C.c_int, i32p]
        L.orx_parse_ht.argtypes = [u8p, i32p]
        L.orx_parse_vhta.argtypes = [u8p, i32p]
        L.orx_parse_vhtb.argtypes = [u8p, i32p]
        L.orx_demod.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p, C.POINTER(OrxFrame), f32p, C.c_int]
        L.orx_qam_to_llr.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.orx_viterbi.argtypes = [f32p, C.c_int, C.c_int, u8p]
        L.orx_descramble.argtypes = [u8p, C.c_int, u8p]
        L.orx_decode.argtypes = [f32p, C.POINTER(OrxFrame), u8p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        L.orx_crc32.argtypes = [u8p, C.c_int]
        L.orx_crc32.restype = C.c_uint32
        L.orx_set_mupos.argtypes = [C.c_int]
        L.orx_set_mupos.restype = None
        L.orx_rx_item.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                  u8p, C.c_int64, C.POINTER(C.c_int64)]
        L.orx_rx_item2.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                   u8p, C.c_int64, C.POINTER(C.c_int64)]
        L.orx_demod2.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, f32p, C.POINTER(OrxFrame), f32p, C.c_int]
        L.orx_rx_batch.argtypes = [f32p, i64p, i32p, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int64]
        _oracle = L
    return _oracle


def have_ref():
    build_oracle()
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        build_oracle()
        L = C.CDLL(REF_SO)
        L.ref_lsig_demod.argtypes = [f32p] * 5
        L.ref_nlsig_demod.argtypes = [f32p] * 5
        L.ref_sig_viterbi.argtypes = [f32p, u8p, C.c_int]
        L.ref_sv_decode.argtypes = [f32p, u8p, C.c_int]
        L.ref_check_legacy.argtypes = [u8p] + [C.POINTER(C.c_int)] * 3
        L.ref_check_ht.argtypes = [u8p]
        L.ref_check_vhta.argtypes = [u8p]
        L.ref_crc8_check.argtypes = [u8p, C.c_int, u8p]
        L.ref_crc8_gen.argtypes = [u8p, C.c_int, u8p]
        L.ref_parse_l.argtypes = [C.c_int, C.c_int, i32p]
        L.ref_parse_ht.argtypes = [u8p, i32p]
        L.ref_parse_vhta.argtypes = [u8p, i32p]
        L.ref_parse_vhtb.argtypes = [u8p, i32p]
        L.ref_qam_to_llr.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.ref_deint_l.argtypes = [f32p, f32p, C.c_int]
        L.ref_deint_nl.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.ref_depas.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int]
        L.ref_bcc.argtypes = [u8p, u8p, C.c_int]
        L.ref_intl_vhtb20.argtypes = [u8p, u8p]
        L.ref_scramble.argtypes = [u8p, u8p, C.c_int, C.c_int]
        L.ref_table_i.argtypes = [C.c_char_p, i32p]
        L.ref_table_f.argtypes = [C.c_char_p, f32p]
        _ref = L
    return _ref


# the 17 SU fields of the reference's c8p_mod, in declaration order (cloud80211phy.h:61-79)
MOD_FIELDS = ["format", "sumu", "ampdu", "nSym", "nSymSamp", "nSD", "nSP", "nSS", "nLTF", "mcs", "len", "mod", "cr",
              "nBPSCS", "nDBPS", "nCBPS", "nCBPSS"]


def c2f(x):
    """complex64 array -> interleaved float32 view (contiguous)."""
    return np.ascontiguousarray(np.asarray(x, dtype=np.complex64)).view(np.float32)


def rx_item(iq, item=0, max_frames=4, want_llr=True):
    """Run the whole oracle chain on one item.  Returns (frames ndarray, llr ndarray, pdu bytes ndarray)."""
    L = oracle()
    iqf = c2f(iq)
    n = iqf.size // 2
    frames = np.zeros(max_frames, dtype=FRAME_DTYPE)
    llr_cap = max(1, (n // 80 + 2) * 832) if want_llr else 0
    llr = np.zeros(max(llr_cap, 1), np.float32)
    pdu = np.zeros(max(4096, n), np.uint8)
    lu, pu = C.c_int64(0), C.c_int64(0)
    nf = L.orx_rx_item(iqf, n, item, max_frames, frames.ctypes.data, llr.ctypes.data if want_llr else None, llr_cap, C.byref(lu),
                       pdu, pdu.size, C.byref(pu))
    return frames[:nf], llr[: lu.value], pdu[: pu.value]


def rx_item2(iq0, iq1, item=0, max_frames=4, want_llr=True):
    """2x2 oracle chain (signal2 + demod2) on one item: antenna 0 drives detection."""
    L = oracle()
    a, b = c2f(iq0), c2f(iq1)
    n = a.size // 2
    frames = np.zeros(max_frames, dtype=FRAME_DTYPE)
    llr_cap = max(1, (n // 80 + 2) * 832) if want_llr else 0
    llr = np.zeros(max(llr_cap, 1), np.float32)
    pdu = np.zeros(max(4096, n), np.uint8)
    lu, pu = C.c_int64(0), C.c_int64(0)
    nf = L.orx_rx_item2(a, b, n, item, max_frames, frames.ctypes.data, llr.ctypes.data if want_llr else None, llr_cap, C.byref(lu),
                        pdu, pdu.size, C.byref(pu))
    return frames[:nf], llr[: lu.value], pdu[: pu.value]


def split_pdus(buf):
    """Split a PDU arena slice into records [fmt][len lo][len hi][MPDU][mcs] (lib/decode_impl.cc:359-361,414-419)."""
    out, i = [], 0
    buf = bytes(buf)
    while i + 3 <= len(buf):
        ln = buf[i + 1] | (buf[i + 2] << 8)
        n = ln + 3 if buf[i] == 20 else ln + 4     # NDP channel report: [20][len][128 x (re, im)], no trailing mcs byte
        out.append(buf[i: i + n])
        i += n
    return out


# ---- the reference's own blocks under the mock scheduler (oracle/_ref/libgr80211_ref.so, built by oracle/Makefile) ----
REFCHAIN_SO = os.path.join(ORACLE_DIR, "_ref", "libgr80211_ref.so")
_refchain = None


def have_refchain():
    build_oracle()
    if os.path.exists("/root/reference/lib/demod_impl.cc") and not os.path.exists(REFCHAIN_SO):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
    return os.path.exists(REFCHAIN_SO)


def refchain_lib():
    global _refchain
    if _refchain is None:
        assert have_refchain(), "oracle/_ref/libgr80211_ref.so is missing (built only where /root/reference is mounted)"
        L = C.CDLL(REFCHAIN_SO)
        L.refchain_create.argtypes = [C.c_int] * 4
        L.refchain_create.restype = C.c_void_p
        L.refchain_destroy.argtypes = [C.c_void_p]
        L.refchain_destroy.restype = None
        L.refchain_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_uint, C.c_int, C.c_int, C.c_int]
        L.refchain_stream.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.refchain_stream.restype = C.c_long
        L.refchain_ntags.argtypes = [C.c_void_p, C.c_int]
        L.refchain_ntags.restype = C.c_long
        L.refchain_tag.argtypes = [C.c_void_p, C.c_int, C.c_long, C.POINTER(C.c_ulonglong), C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                   C.POINTER(C.c_double), C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int)]
        L.refchain_nmsgs.argtypes = [C.c_void_p]
        L.refchain_nmsgs.restype = C.c_long
        L.refchain_msg.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.POINTER(C.c_ubyte))]
        L.refchain_msg.restype = C.c_long
        L.refchain_calls.argtypes = [C.c_void_p]
        L.refchain_calls.restype = C.c_ulonglong
        L.refchain_backlog.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refchain_backlog.restype = C.c_long
        L.refchain_bench.argtypes = [f32p, i64p, i32p, C.c_int, C.c_int, i64p]
        _refchain = L
    return _refchain


REF_BLOCK_NAMES = ["trigger", "sync", "signal", "demod", "decode"]


class RefChain:
    """trigger -> sync -> signal[2] -> demod[2] -> decode built from the reference's unmodified *_impl.cc."""

    def __init__(self, nant=1, mupos=0, mugid=0, debug=False):
        self.L = refchain_lib()
        self.nant = nant
        self.h = self.L.refchain_create(nant, mupos, mugid, int(debug))
        assert self.h

    def close(self):
        if self.h:
            self.L.refchain_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def run(self, iq0, iq1=None, seed=0, max_call=0, record=True, flush=True, presiso=None):
        """Feed a capture piece (presiso from the oracle unless given) and run until nothing moves."""
        a = c2f(iq0)
        n = a.size // 2
        if presiso is None:
            preac = np.zeros(max(n, 1), np.float32)
            preconj = np.zeros(2 * max(n, 1), np.float32)
            if n:
                oracle().orx_presiso(a, n, preac, preconj)
        else:
            preac = np.ascontiguousarray(presiso[0], np.float32)
            preconj = c2f(presiso[1])
        b = c2f(iq1) if iq1 is not None else None
        self.L.refchain_run(self.h, preac.ctypes.data, preconj.ctypes.data, a.ctypes.data, b.ctypes.data if b is not None else None,
                            n, seed, max_call, int(record), int(flush))

    def stream(self, which):
        """0 trigger flags, 1 sync flags, 2 / 3 signal's output per antenna, 4 demod's soft bits"""
        p = C.c_void_p()
        n = self.L.refchain_stream(self.h, which, C.byref(p))
        dt = [np.uint8, np.uint8, np.complex64, np.complex64, np.float32][which]
        if n == 0:
            return np.zeros(0, dt)
        return np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(p.value), dtype=dt).copy()

    def tags(self, block):
        """{absolute offset: {key: value}} of everything block (index or name) attached, in stream order"""
        bi = REF_BLOCK_NAMES.index(block) if isinstance(block, str) else block
        out = {}
        off, typ, val, ncv = C.c_ulonglong(), C.c_int(), C.c_double(), C.c_int()
        cv = C.POINTER(C.c_float)()
        key = C.create_string_buffer(64)
        for i in range(self.L.refchain_ntags(self.h, bi)):
            assert self.L.refchain_tag(self.h, bi, i, C.byref(off), key, 64, C.byref(typ), C.byref(val), C.byref(cv), C.byref(ncv)) == 0
            if typ.value == 0:
                v = int(val.value)
            elif typ.value == 1:
                v = np.float32(val.value)
            else:
                v = np.ctypeslib.as_array(cv, shape=(2 * ncv.value,)).copy().view(np.complex64)
            out.setdefault(int(off.value), {})[key.value.decode()] = v
        return out

    def messages(self):
        out = []
        p = C.POINTER(C.c_ubyte)()
        for i in range(self.L.refchain_nmsgs(self.h)):
            n = self.L.refchain_msg(self.h, i, C.byref(p))
            assert n >= 0
            out.append(bytes(bytearray(p[:n])))
        return out

    def calls(self):
        return int(self.L.refchain_calls(self.h))


def _same_f32(a, b):
    a, b = np.float32(a), np.float32(b)
    return (np.isnan(a) and np.isnan(b)) or a.tobytes() == b.tobytes()


def chain_vs_oracle(iq0, iq1=None, mupos=0, mugid=0, max_frames=64, seed=0, max_call=0, pieces=None):
    """Run the reference's blocks (RefChain) and the restated oracle (orx_rx_item[2]) over one capture and compare everything
    the blocks make visible: trigger flags, sync flags + tags, signal tags + output samples, demod tags + soft bits, decode's
    messages -- bit for bit.  `pieces`: split points for feeding the reference chain in several run() calls.  Returns
    (list of mismatch strings, summary dict)."""
    nant = 2 if iq1 is not None else 1
    x0 = np.ascontiguousarray(iq0, np.complex64)
    x1 = np.ascontiguousarray(iq1, np.complex64) if iq1 is not None else None
    n = x0.size
    L = oracle()
    L.orx_set_mupos(mupos)
    try:
        fo, llr, po = (rx_item2(x0, x1, max_frames=max_frames) if nant == 2 else rx_item(x0, max_frames=max_frames))
    finally:
        L.orx_set_mupos(0)
    preac = np.zeros(max(n, 1), np.float32)
    preconj = np.zeros(2 * max(n, 1), np.float32)
    if n:
        L.orx_presiso(c2f(x0), n, preac, preconj)
    c = RefChain(nant, mupos, mugid)
    cuts = [0] + sorted(pieces or []) + [n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        c.run(x0[a:b], x1[a:b] if x1 is not None else None, seed=seed, max_call=max_call, flush=(b == n),
              presiso=(preac[a:b], preconj.view(np.complex64)[a:b]))
    bad = []
    # trigger: the flags are a pure function of preac
    trig = np.zeros(max(n, 1), np.uint8)
    st = np.zeros(5, np.int32)
    if n:
        L.orx_trigger(preac, n, trig, st)
    rt = c.stream(0)
    if rt.size != n or not np.array_equal(rt, trig[:n]):
        bad.append("trigger flags differ")
    sync_flags = np.flatnonzero(c.stream(1))
    stags, gtags, dtags = c.tags("sync"), c.tags("signal"), c.tags("demod")
    if sorted(stags) != list(sync_flags):
        bad.append("sync tags are not at the sync flags")
    frames = [f for f in fo if f["nsamp"] > 0]
    if len(frames) == max_frames:
        bad.append("max_frames too small for this capture")
    sig_out = [c.stream(2)] + ([c.stream(3)] if nant == 2 else [])
    soft = c.stream(4)
    goff = doff = 0
    nllr = 0
    for k, f in enumerate(frames):
        t = stags.get(int(f["sync_idx"]))
        if t is None or set(t) != {"rad", "snr", "rssi"} or not all(_same_f32(t[q], f[q]) for q in ("rad", "snr", "rssi")):
            bad.append("frame %d: sync tag %r vs oracle (%r, %r, %r)" % (k, t, f["rad"], f["snr"], f["rssi"]))
        t = gtags.get(goff)
        if t is None or set(t) != {"cfo", "snr", "rssi", "seq", "mcs", "len", "nsamp", "chan"}:
            bad.append("frame %d: no signal tag at %d" % (k, goff))
            break
        if (t["mcs"], t["len"], t["nsamp"], t["seq"]) != (f["l_mcs"], f["l_len"], f["nsamp"], k + 1) or not _same_f32(t["cfo"], f["cfo_hz"]) \
                or not _same_f32(t["snr"], f["snr"]) or not _same_f32(t["rssi"], f["rssi"]) or t["chan"].size != 64:
            bad.append("frame %d: signal tag values %r" % (k, {q: v for q, v in t.items() if q != "chan"}))
        if f["status"] == 4:       # ORX_E_TRUNC: S_COPY never finishes; whatever demod made of the partial copy is not compared
            break
        # signal's output: the CFO-corrected copy (pad samples are whatever the buffer held; zeros in the harness)
        start = int(f["sync_idx"]) + 224
        ph = (np.arange(f["nsamp"], dtype=np.float32) + np.float32(224)) * np.float32(f["rad"])
        for a, xs in enumerate([x0, x1][:nant]):
            want = (xs[start:start + f["nsamp"]] * (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)).astype(np.complex64)
            got = sig_out[a][goff:goff + f["nsamp"]]
            # cosf / sinf of libm vs numpy's float32 cos / sin may differ in the last place: compare to 2 ulp of the sample
            if got.size != want.size or np.abs(got - want).max(initial=0) > 4e-7 * max(1e-9, np.abs(want).max(initial=0)):
                bad.append("frame %d: signal output, antenna %d" % (k, a))
        goff += int(f["nsamp"]) + 320
        if f["status"] == 5:       # ORX_E_FORMAT: demod went to CLEAN without a tag (the tag count is checked below)
            continue
        t = dtags.get(doff)
        if t is None:
            bad.append("frame %d (status %d): no demod tag at %d" % (k, f["status"], doff))
            break
        if f["status"] == 7:       # NDP: 1024 floats (unwritten) + the channel report
            if t.get("total") != 1024 or t.get("trellis") != 0 or "mu2x1chan" not in t:
                bad.append("frame %d: NDP tag %r" % (k, sorted(t)))
            doff += 1024
            continue
        keys = {"cfo", "snr", "rssi", "format", "mcs", "len", "cr", "ampdu", "trellis", "total"}
        if f["format"] == 2:
            keys |= {"sssnr0"} | ({"sssnr1"} if nant == 2 else set())
        if set(t) != keys:
            bad.append("frame %d: demod tag keys %r" % (k, sorted(set(t) ^ keys)))
        else:
            for q in ("format", "mcs", "len", "cr", "ampdu", "trellis", "total"):
                if t[q] != f[q]:
                    bad.append("frame %d: demod tag %s = %r, oracle %r" % (k, q, t[q], f[q]))
            for q, fq in (("cfo", "cfo_hz"), ("snr", "snr"), ("rssi", "rssi"), ("sssnr0", "sssnr0"), ("sssnr1", "sssnr1")):
                if q in t and not _same_f32(t[q], f[fq]):
                    bad.append("frame %d: demod tag %s = %r, oracle %r" % (k, q, t[q], f[fq]))
        tot = int(f["total"])
        got, want = soft[doff:doff + tot], llr[int(f["llr_off"]):int(f["llr_off"]) + tot]
        if got.size != tot or got.tobytes() != want.tobytes():
            nd = int(np.count_nonzero(got != want)) if got.size == want.size else -1
            bad.append("frame %d: soft bits differ (%d of %d)" % (k, nd, tot))
        nllr += tot
        doff += tot
    if len(gtags) != len(frames):
        bad.append("signal tagged %d frames, oracle has %d" % (len(gtags), len(frames)))
    ntag = sum(1 for f in frames if f["status"] in (0, 6, 7))
    if len(dtags) != ntag and not any(f["status"] == 4 for f in frames):
        bad.append("demod tagged %d frames, oracle has %d" % (len(dtags), ntag))
    msgs, want = c.messages(), split_pdus(po)
    if msgs != want:
        bad.append("messages differ: %d vs oracle %d" % (len(msgs), len(want)))
    info = dict(frames=len(frames), messages=len(msgs), soft_bits=nllr, calls=c.calls(), statuses=[int(f["status"]) for f in frames])
    c.close()
    return bad, info


def colliding_captures(g, rng, n=10, snr=28):
    """captures in which a second frame starts INSIDE the first (its preamble over the payload, over the SIG fields, or right
    behind the preamble) and a third follows closely: what trigger / sync hold-off / signal's S_COPY swallow make of it"""
    iq, offs = g["iq"], g["offs"]
    out = []
    for _ in range(n):
        a, b, c = (int(v) for v in rng.integers(1, len(offs) - 1, 3))
        fa, fb, fc = (iq[offs[k]:offs[k + 1]] for k in (a, b, c))
        d1 = int(rng.choice([200, 330, 420, 700, 1500, fa.size // 2]))
        d2 = d1 + int(rng.integers(300, fb.size + 400))
        x = np.zeros(max(fa.size, d1 + fb.size, d2 + fc.size) + 200, np.complex64)
        x[:fa.size] += fa
        x[d1:d1 + fb.size] += np.complex64(rng.uniform(0.3, 1.5)) * fb
        x[d2:d2 + fc.size] += np.complex64(rng.uniform(0.3, 1.5)) * fc
        sg = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
        r2 = np.random.default_rng(int(rng.integers(1 << 30)))
        out.append((x + sg * (r2.standard_normal(x.size) + 1j * r2.standard_normal(x.size))).astype(np.complex64))
    return out
