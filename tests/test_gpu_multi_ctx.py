"""Several contexts in ONE process: the header promises that different c8b_ctx / c8b_blk are independent (one GNU Radio
flowgraph = one thread per block, possibly one GPU per block).  Kernel attributes (the > 48 KB dynamic shared memory of the
decode kernels) are per DEVICE and are set per context at c8b_create -- not once per process."""
import threading

import numpy as np
import pytest

import oracle_lib as ol
from __graft_entry__ import load_pkg

pytestmark = pytest.mark.gpu


def _items(golden):
    g = golden["frames_siso"]
    offs = g["offs"]
    return g["iq"], offs[:-1], np.diff(offs).astype(np.int32)


def _expected(iq, off, ln):
    return [bytes(ol.rx_item(iq[o:o + n], max_frames=1)[2]) for o, n in zip(off, ln)]


@pytest.mark.parametrize("decode_mode", [1, 2])
def test_two_devices_one_process(golden, decode_mode):
    """a context per GPU, both decode kernels (65 KB / 19 KB of dynamic shared memory): the second device must get its own
    opt-in -- with a process-wide guard its first decode launch failed"""
    pkg = load_pkg()
    ndev = pkg._cabi.lib().c8b_device_count()
    if ndev < 2:
        pytest.skip("one GPU visible")
    iq, off, ln = _items(golden)
    want = _expected(iq, off, ln)
    rxs = [pkg.Receiver(device=d, decode_mode=decode_mode) for d in range(min(ndev, 4))]
    try:
        for rx in rxs + rxs[::-1]:
            fr, pdu = rx.rx_batch(iq, off, ln)
            assert [bytes(pdu[i, :fr[i]["pdu_bytes"]]) for i in range(off.size)] == want
    finally:
        for rx in rxs:
            rx.close()


def test_contexts_from_concurrent_threads(golden):
    """four host threads, each with its own context (created concurrently, used concurrently, both decode kernels)"""
    pkg = load_pkg()
    ndev = pkg._cabi.lib().c8b_device_count()
    iq, off, ln = _items(golden)
    want = _expected(iq, off, ln)
    errs = []

    def worker(k):
        try:
            rx = pkg.Receiver(device=k % ndev, decode_mode=1 + k % 2)
            for _ in range(3):
                fr, pdu = rx.rx_batch(iq, off, ln)
                assert [bytes(pdu[i, :fr[i]["pdu_bytes"]]) for i in range(off.size)] == want
            rx.close()
        except Exception as e:                                   # noqa: BLE001 -- reported by the main thread
            errs.append((k, repr(e)))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


def test_two_block_chains_from_two_threads(golden):
    """two complete seven-block chains (c8b_blk_*), one per thread -- GNU Radio's thread-per-block model with two flowgraphs"""
    pkg = load_pkg()
    g = golden["frames_siso"]
    x = np.ascontiguousarray(g["iq"][:g["offs"][12]])
    _, _, po = ol.rx_item(x, max_frames=40)
    want = pkg.blocks.split_messages(bytes(po))
    rx = pkg.Receiver(device=0)
    preac, preconj = rx.presiso(x)
    rx.close()
    out, errs = {}, []

    def worker(k):
        try:
            ch = pkg.blocks.Chain(nant=1, seed=10 + k, max_call=2048 + 1024 * k)
            try:
                out[k] = ch.run(preac, preconj, x)
            finally:
                ch.close()
        except Exception as e:                                   # noqa: BLE001
            errs.append((k, repr(e)))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert out[0] == want and out[1] == want and len(want) >= 11
