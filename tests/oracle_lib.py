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
        out.append(buf[i: i + ln + 4])
        i += ln + 4
    return out
