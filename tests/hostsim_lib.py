"""TEST INFRASTRUCTURE: ctypes door onto tests/hostsim/libhostsim.so = the product's per-frame device
routines (gr-ieee80211_b200/csrc/phy_serial.cuh) compiled for the host."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostsim", "libhostsim.so")
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "hostsim")])
        L = C.CDLL(SO)
        L.hs_fft64.argtypes = [f32p, f32p]
        L.hs_conj_at.argtypes = [f32p, C.c_int, f32p]
        L.hs_trigger.argtypes = [f32p, C.c_int, u8p]
        L.hs_sync.argtypes = [f32p, C.c_float, C.c_float, C.POINTER(C.c_int)] + [C.POINTER(C.c_float)] * 3
        L.hs_sig_viterbi.argtypes = [f32p, u8p, C.c_int]
        L.hs_crc8.argtypes = [u8p, C.c_int, u8p]
        L.hs_detect.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, f32p]
        L.hs_detect_scan.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_void_p, f32p, np.ctypeslib.ndpointer(np.int32, flags="C"), C.c_int]
        L.hs_header.argtypes = [f32p, C.c_void_p, f32p, C.c_int, f32p]
        L.hs_header2.argtypes = [f32p, f32p, C.c_void_p, f32p, f32p, f32p]
        _lib = L
    return _lib


class HostBackend:
    """Backend of gr-ieee80211_b200/blocks.py over the host build of the block state machines (hs_blk_*): the per-frame
    arithmetic is the host build of phy_serial.cuh, soft bits are zeros and PDUs placeholders (see hostsim.cc)."""

    def __init__(self):
        self.L = lib()
        vp, i = C.c_void_p, C.c_int
        self.L.hs_blk_create.restype = vp
        self.L.hs_blk_create.argtypes = [i, i]
        self.L.hs_blk_destroy.argtypes = [vp]
        self.L.hs_blk_work.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, i, vp]

    def create(self, kind, mupos=0, mugid=0):
        return C.c_void_p(self.L.hs_blk_create(kind, mupos))

    def destroy(self, h):
        self.L.hs_blk_destroy(h)

    def forecast(self, kind, noutput):
        return noutput + 160 if kind == 6 else noutput

    def work(self, h, *args):
        rc = self.L.hs_blk_work(h, *args)
        assert rc == 0, rc
