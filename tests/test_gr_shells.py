"""The gr::block shells (gr-ieee80211_b200/gr/lib/rx_blocks_impl.cc) compiled against a miniature of the GNU Radio runtime
(tests/gr_mock: gr::block / io_signature / pmt with GNU Radio 3.10's names, signatures and access levels -- and the
reference's own public headers include/gnuradio/ieee80211/*.h when /root/reference is mounted) and driven by a mock
scheduler (tests/gr_mock/run_chain.cc).  CPU: they build, link against the product library and refuse to run without a
GPU.  GPU: the messages on decode's port and the stream tags equal the oracle's one-pass results."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import ROOT, load_pkg

MOCK = os.path.join(ROOT, "tests", "gr_mock")
EXE = os.path.join(MOCK, "build", "run_chain")
EXE_RX = os.path.join(MOCK, "build", "run_rx")
EXE_REF = os.path.join(MOCK, "build", "run_chain_ref")     # the same scheduler over the reference's unmodified blocks (CPU)


def _build():
    # the product library is built by __graft_entry__.build() (never rebuilt here: this process may have it mapped)
    assert os.path.exists(os.path.join(ROOT, "gr-ieee80211_b200", "lib", "libc80211b200.so")), "run __graft_entry__.build() first"
    subprocess.check_call(["make", "-s", "-C", MOCK])
    assert os.path.exists(EXE)


def test_shells_build_and_fail_loudly_without_gpu(tmp_path):
    _build()
    pkg = load_pkg()
    if pkg._cabi.lib().c8b_device_count() > 0:
        return
    for exe in (EXE, EXE_RX):
        r = subprocess.run([exe, "1", "0", "0", "1", "4096", "0", str(tmp_path), str(tmp_path / "out.txt")], capture_output=True, text=True)
        assert r.returncode == 4 and "no CPU path" in r.stderr      # make() throws: no GPU, no block


def test_rx_pybind_binding_compiles():
    """gr/python/ieee80211/bindings/rx_python.cc against pybind11 + the mock runtime (syntax / types only)"""
    try:
        import pybind11
    except ImportError:
        pytest.skip("pybind11 not installed")
    import sysconfig
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include()]
    if os.path.isdir("/root/reference/include"):
        inc.append("-I/root/reference/include")
    inc += ["-I" + os.path.join(MOCK, "include"), "-I" + os.path.join(ROOT, "gr-ieee80211_b200", "gr", "include"), "-I" + os.path.join(ROOT, "include")]
    src = os.path.join(ROOT, "gr-ieee80211_b200", "gr", "python", "ieee80211", "bindings", "rx_python.cc")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-fsyntax-only"] + inc + [src])


def _run(tmp_path, nant, x, x1=None, mupos=0, mugid=0, seed=1, max_call=4096, debug=0, exe=EXE, presiso=None):
    if presiso is None:
        pkg = load_pkg()
        rx = pkg.Receiver(device=0)
        preac, preconj = rx.presiso(x)
        rx.close()
    else:
        preac, preconj = presiso
    preac.tofile(tmp_path / "preac.f32")
    preconj.astype(np.complex64).tofile(tmp_path / "preconj.c64")
    x.astype(np.complex64).tofile(tmp_path / "sig.c64")
    if x1 is not None:
        x1.astype(np.complex64).tofile(tmp_path / "sig1.c64")
    out = tmp_path / "out.txt"
    r = subprocess.run([exe, str(nant), str(mupos), str(mugid), str(seed), str(max_call), str(debug), str(tmp_path), str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    msgs, tags = [], {}
    for line in open(out):
        w = line.split()
        if w[0] == "MSG":
            msgs.append(bytes.fromhex(w[1]))
        else:
            tags.setdefault((w[1], int(w[2])), {})[w[3]] = w[4]
    return msgs, tags, r.stdout


def _noisy(x, snr, seed):
    rng = np.random.default_rng(seed)
    s = 0.1875 / np.sqrt(2 * 10 ** (snr / 10))
    return (x + s * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)


@pytest.mark.gpu
@pytest.mark.parametrize("max_call", [4096, 900])
def test_shells_siso_equal_oracle(golden, tmp_path, max_call):
    _build()
    pkg = load_pkg()
    g = golden["frames_siso"]
    x = _noisy(g["iq"], 28.0, 21)
    fo, _, po = ol.rx_item(x, max_frames=40)
    want = pkg.blocks.split_messages(bytes(po))
    msgs, tags, stdout = _run(tmp_path, 1, x, seed=max_call, max_call=max_call, debug=1)
    assert msgs == want and len(want) >= 30
    print([ln for ln in stdout.splitlines() if ln.startswith("run_chain:")])
    ok = fo[(fo["status"] != 9) & (fo["nsamp"] > 0)]
    soff = doff = 0
    for f in ok:
        t = tags[("sync", int(f["sync_idx"]))]
        assert set(t) == {"rad", "snr", "rssi"} and abs(float(t["rad"]) - float(f["rad"])) <= 1e-6
        t = tags[("signal", soff)]
        assert set(t) == {"cfo", "snr", "rssi", "seq", "mcs", "len", "nsamp", "chan"}
        assert (int(t["mcs"]), int(t["len"]), int(t["nsamp"])) == (int(f["l_mcs"]), int(f["l_len"]), int(f["nsamp"])) and t["chan"].startswith("c32[64]")
        soff += int(f["nsamp"]) + 320
        if f["status"] == 0:
            t = tags[("demod", doff)]
            keys = {"cfo", "snr", "rssi", "format", "mcs", "len", "cr", "ampdu", "trellis", "total"} | ({"sssnr0"} if f["format"] == 2 else set())
            assert set(t) == keys, (set(t) ^ keys)
            for k in ("format", "mcs", "len", "cr", "ampdu", "trellis", "total"):
                assert int(t[k]) == int(f[k]), (k, t[k], f[k])
            doff += int(f["total"])
    # decode(ifdebug = True): one "crc32 correct" line per published MPDU, counters as lib/decode_impl.cc:377-411
    lines = [s for s in stdout.splitlines() if s.startswith("ieee80211 decode, ")]
    good = [s for s in lines if " crc32 correct, " in s]
    assert len(good) == len(want) and good[-1].split("total:")[1].split(",")[0] == str(len(want))


@pytest.mark.gpu
def test_shells_2x2_equal_oracle(golden, tmp_path):
    _build()
    pkg = load_pkg()
    g = golden["frames_mimo"]
    a, b = _noisy(g["iq0"], 30.0, 13579), _noisy(g["iq1"], 30.0, 24680)
    _, _, po = ol.rx_item2(a, b, max_frames=32)
    want = pkg.blocks.split_messages(bytes(po))
    msgs, tags, _ = _run(tmp_path, 2, a, b, seed=5)
    assert msgs == want and len(want) >= 16
    assert any("sssnr1" in t for (blk, _), t in tags.items() if blk == "demod2")


@pytest.mark.gpu
@pytest.mark.parametrize("nant,max_call", [(1, 8192), (1, 1500), (2, 8192)])
def test_rx_sink_block_equals_oracle(golden, tmp_path, nant, max_call):
    """gr::ieee80211::rx (gr/lib/rx_impl.cc): the chain as one sink block on the live-stream session, fed in random pieces"""
    _build()
    pkg = load_pkg()
    if nant == 1:
        x, x1 = _noisy(golden["frames_siso"]["iq"], 28.0, 33), None
        _, _, po = ol.rx_item(x, max_frames=40)
    else:
        g = golden["frames_mimo"]
        x, x1 = _noisy(g["iq0"], 30.0, 13579), _noisy(g["iq1"], 30.0, 24680)
        _, _, po = ol.rx_item2(x, x1, max_frames=32)
    want = pkg.blocks.split_messages(bytes(po))
    msgs, _, stdout = _run(tmp_path, nant, x, x1, mugid=2, seed=max_call, max_call=max_call, debug=1, exe=EXE_RX)
    assert msgs == want and len(want) >= 16
    print([ln for ln in stdout.splitlines() if ln.startswith("run_rx:")])
    good = [s for s in stdout.splitlines() if s.startswith("ieee80211 decode, ") and " crc32 correct, " in s]
    assert len(good) == len(want)


def _oracle_presiso(x):
    n = x.size
    preac, preconj = np.zeros(n, np.float32), np.zeros(2 * n, np.float32)
    ol.oracle().orx_presiso(ol.c2f(x), n, preac, preconj)
    return preac, preconj.view(np.complex64)


@pytest.mark.skipif(not os.path.exists(EXE_REF) and not os.path.isdir("/root/reference/lib"), reason="reference block objects not built")
def test_reference_blocks_under_the_mock_scheduler(golden, tmp_path):
    """CPU: run_chain_ref = tests/gr_mock/run_chain.cc linked with the reference's own *_impl.cc objects; its messages are
    the oracle's for two different call schedules (the dump the shells are compared with below)"""
    _build()
    x = _noisy(golden["frames_siso"]["iq"], 28.0, 21)
    _, _, po = ol.rx_item(x, max_frames=40)
    want = ol.split_pdus(po)
    for seed, max_call in ((1, 4096), (7, 900)):
        msgs, tags, _ = _run(tmp_path, 1, x, seed=seed, max_call=max_call, exe=EXE_REF, presiso=_oracle_presiso(x))
        assert msgs == want and len(want) >= 30
        assert sum(1 for (blk, _) in tags if blk == "signal") == 31 and sum(1 for (blk, _) in tags if blk == "demod") == 31


def _same_tags(got, want):
    """shell tags vs reference-block tags: same blocks / absolute offsets / keys; integers equal, floats to the parity gates"""
    assert sorted(got) == sorted(want), (sorted(set(got) ^ set(want))[:8])
    for k in want:
        assert set(got[k]) == set(want[k]), (k, set(got[k]) ^ set(want[k]))
        for key, w in want[k].items():
            g = got[k][key]
            if w.startswith("c32["):
                assert g.split(":")[0] == w.split(":")[0]
                a, b = float(g.split(":")[1]), float(w.split(":")[1])
                assert abs(a - b) <= 2e-4 * max(abs(b), 1e-3), (k, key, g, w)
            elif key in ("rad", "snr", "rssi", "cfo", "sssnr0", "sssnr1"):
                a, b = float(g), float(w)
                if np.isnan(b):
                    assert np.isnan(a)
                elif key == "rad":
                    assert abs(a - b) <= 1e-6
                elif key == "cfo":
                    assert abs(a - b) <= 1e-6 * 3183098.9
                else:
                    assert abs(a - b) <= 1e-3 * max(1.0, abs(b)), (k, key, g, w)
            else:
                assert int(g) == int(w), (k, key, g, w)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "demod_impl.o")), reason="reference block objects not built")
@pytest.mark.parametrize("nant,seed,max_call", [(1, 3, 4096), (1, 9, 1100), (2, 5, 4096)])
def test_shells_equal_reference_blocks(golden, tmp_path, nant, seed, max_call):
    """A/B under the same mock scheduler, same seed: the compiled shells (CUDA) against the reference's own unmodified blocks
    (CPU) -- identical messages, and every stream tag at the identical absolute offset with the same keys and values"""
    _build()
    if nant == 1:
        x, x1 = _noisy(golden["frames_siso"]["iq"], 28.0, 21), None
    else:
        g = golden["frames_mimo"]
        x, x1 = _noisy(g["iq0"], 30.0, 13579), _noisy(g["iq1"], 30.0, 24680)
    ps = _oracle_presiso(x)
    (tmp_path / "ref").mkdir()
    (tmp_path / "gpu").mkdir()
    rm, rt, _ = _run(tmp_path / "ref", nant, x, x1, seed=seed, max_call=max_call, exe=EXE_REF, presiso=ps)
    gm, gt, _ = _run(tmp_path / "gpu", nant, x, x1, seed=seed, max_call=max_call, exe=EXE, presiso=ps)
    assert gm == rm and len(rm) >= 16
    _same_tags(gt, rt)
