"""ctypes binding of libc80211b200.so (include/c80211b200.h).  This is the reference-side stub a
maintainer would write for a Python host (INTEGRATION.md shows the C++ gr::block equivalent).

There is no CPU fallback: loading fails loudly when the library is missing, and creating a context
fails with C8B_ERR_NO_DEVICE when no CUDA device is visible."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# C8B_LIB: another build of the same library (kernel experiments: tools/build_variant.sh); default = the in-tree product
LIB_PATH = os.environ.get("C8B_LIB") or os.path.join(HERE, "lib", "libc80211b200.so")

K_NAMES = ("presiso", "detect", "header", "demod", "viterbi")

ST_OK, ST_NO_TRIGGER, ST_SYNC, ST_LSIG, ST_TRUNC, ST_FORMAT, ST_DECODE_RANGE, ST_NDP, ST_OVERFLOW, ST_EMPTY = range(10)


class C8bFrame(C.Structure):
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


FRAME_DTYPE = np.dtype([(n, {C.c_int32: "<i4", C.c_float: "<f4", C.c_int64: "<i8"}[t]) for n, t in C8bFrame._fields_], align=True)
assert FRAME_DTYPE.itemsize == C.sizeof(C8bFrame)


# c8b_tag: one stream tag group of the per-block entry points (c8b_blk_work)
TAG_DTYPE = np.dtype([("port", "<i4"), ("idx", "<i4"), ("nvec", "<i4"), ("seq", "<i4"), ("f", FRAME_DTYPE), ("vec", "<f4", (256,))], align=True)
assert TAG_DTYPE.itemsize == 16 + C.sizeof(C8bFrame) + 1024 and TAG_DTYPE.fields["f"][1] == 16


class C8bTxFrame(C.Structure):
    _fields_ = [("format", C.c_int32), ("mcs", C.c_int32), ("psdu_off", C.c_int64), ("psdu_len", C.c_int32), ("cfo_hz", C.c_float),
                ("out_off", C.c_int64)]


TXFRAME_DTYPE = np.dtype([("format", "<i4"), ("mcs", "<i4"), ("psdu_off", "<i8"), ("psdu_len", "<i4"), ("cfo_hz", "<f4"), ("out_off", "<i8")],
                         align=True)
assert TXFRAME_DTYPE.itemsize == C.sizeof(C8bTxFrame)


class C8bTxMu(C.Structure):
    _fields_ = [("mcs", C.c_int32 * 2), ("psdu_len", C.c_int32 * 2), ("psdu_off", C.c_int64 * 2), ("group_id", C.c_int32), ("cfo_hz", C.c_float),
                ("out_off", C.c_int64), ("q_index", C.c_int64)]


TXMU_DTYPE = np.dtype([("mcs", "<i4", (2,)), ("psdu_len", "<i4", (2,)), ("psdu_off", "<i8", (2,)), ("group_id", "<i4"), ("cfo_hz", "<f4"),
                       ("out_off", "<i8"), ("q_index", "<i8")], align=True)
assert TXMU_DTYPE.itemsize == C.sizeof(C8bTxMu)


class C8bCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("chunk_items", C.c_int32), ("max_item_len", C.c_int32), ("max_frames", C.c_int32),
                ("mupos", C.c_int32), ("mugid", C.c_int32), ("no_overlap", C.c_int32), ("decode_mode", C.c_int32), ("frontend_mode", C.c_int32), ("mmse", C.c_int32), ("reserved", C.c_int32 * 4)]


# every symbol include/c80211b200.h declares: (name, restype, argtypes)
_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
SYMBOLS = [
    ("c8b_abi_version", _i, []),
    ("c8b_device_count", _i, []),
    ("c8b_create", _i, [C.POINTER(C8bCfg), C.POINTER(_vp)]),
    ("c8b_destroy", None, [_vp]),
    ("c8b_last_error", C.c_char_p, [_vp]),
    ("c8b_stream", _vp, [_vp]),
    ("c8b_lut_size", _sz, []),
    ("c8b_lut_blob", _i, [_vp, _sz]),
    ("c8b_lut_load", _i, [_vp, _vp, _sz]),
    ("c8b_lut_load_dev", _i, [_vp, _vp, _sz]),
    ("c8b_rx_batch", _i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i64]),
    ("c8b_rx_batch_sc16", _i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i64]),
    ("c8b_rx_batch2", _i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i64]),
    ("c8b_rx_batch_dev", _i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i64]),
    ("c8b_rx_batch_dev_async", _i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i64]),
    ("c8b_rx_batch2_dev", _i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i64]),
    ("c8b_sync", _i, [_vp]),
    ("c8b_stream_begin", _i, [_vp, _i, _i64]),
    ("c8b_stream_push", _i, [_vp, _vp, _vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _i64]),
    ("c8b_stream_state", _i, [_vp, _vp, _vp, _vp]),
    ("c8b_tx_nsamp", _i, [_i, _i, _i]),
    ("c8b_tx_batch", _i, [_vp, _vp, _i64, _vp, _i, C.c_float, _i, _vp, _i64]),
    ("c8b_tx_batch_dev", _i, [_vp, _vp, _i64, _vp, _i, C.c_float, _i, _vp, _i64]),
    ("c8b_tx_batch2", _i, [_vp, _vp, _i64, _vp, _i, C.c_float, _i, _vp, _vp, _i64]),
    ("c8b_tx_batch2_dev", _i, [_vp, _vp, _i64, _vp, _i, C.c_float, _i, _vp, _vp, _i64]),
    ("c8b_tx_mu_nsamp", _i, [_i, _i, _i, _i]),
    ("c8b_tx_udp_parse_mu", _i, [_vp, _i, _vp, _vp, _vp]),
    ("c8b_tx_udp_parse_bfq", _i, [_vp, _i, _vp]),
    ("c8b_tx_mu_batch", _i, [_vp, _vp, _i64, _vp, _i, _vp, _i, C.c_float, _i, _vp, _vp, _i64]),
    ("c8b_tx_mu_batch_dev", _i, [_vp, _vp, _i64, _vp, _i, _vp, _i, C.c_float, _i, _vp, _vp, _i64]),
    ("c8b_tx_random_psdu_dev", _i, [_vp, _vp, _i64, _vp, _i, C.c_uint64]),
    ("c8b_tx_udp_parse", _i, [_vp, _i, _vp, _vp]),
    ("c8b_tx_from_udp", _i, [_vp, _vp, _vp, _vp, _i, _i, C.c_float, _i, _vp, _i64, _vp, _vp]),
    ("c8b_timing_enable", _i, [_vp, _i]),
    ("c8b_timing_read", _i, [_vp, _vp, _vp, _i]),
    ("c8b_presiso", _i, [_vp, _vp, _i64, _vp, _vp]),
    ("c8b_trigger", _i, [_vp, _vp, _i64, _vp]),
    ("c8b_trigger_events", _i, [_vp, _vp, _i64, _i, _vp, _i, _vp, _vp]),
    ("c8b_detect", _i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    ("c8b_demod", _i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i64]),
    ("c8b_demod2", _i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i64]),
    ("c8b_decode", _i, [_vp, _vp, _i64, _vp, _i, _vp, _i64, _vp, _i64]),
    ("c8b_blk_create", _i, [C.POINTER(C8bCfg), _i, C.POINTER(_vp)]),
    ("c8b_blk_destroy", None, [_vp]),
    ("c8b_blk_last_error", C.c_char_p, [_vp]),
    ("c8b_blk_ports", _i, [_i, _vp, _vp, _vp, _vp]),
    ("c8b_blk_forecast", _i, [_i, _i]),
    ("c8b_blk_work", _i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
]

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libc80211b200.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C gr-ieee80211_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)          # AttributeError = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class C8bError(RuntimeError):
    pass


def ptr(a):
    """numpy array (C contiguous) or None or int -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)
