"""Multi-GPU plumbing: one process per GPU (torchrun), frames sharded by item index, NO data-path
collective.  The only exchanges are (1) one broadcast of the LUT blob from rank 0 at start-up and
(2) one reduction of the result counters / device times at the end.  Works with any
torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests."""
import numpy as np

from .rx import lut_blob

LUT_MAGIC = 0x4C423843


def shard_range(n_items, rank, world):
    """Contiguous block of items for `rank`: sizes differ by at most one, union = [0, n_items)."""
    base, rem = divmod(int(n_items), int(world))
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


def broadcast_lut(torch, dist, rank, world, device):
    """Rank 0 builds the blob by formula; everybody else receives it.  Returns a uint8 tensor on `device`."""
    import ctypes as C  # noqa: F401
    from . import _cabi
    n = int(_cabi.lib().c8b_lut_size())
    if rank == 0:
        t = torch.from_numpy(lut_blob().copy()).to(device)
    else:
        t = torch.zeros(n, dtype=torch.uint8, device=device)
    if world > 1:
        dist.broadcast(t, src=0)
    return t


def check_lut(blob_u8):
    """Header sanity of a received blob (the library re-checks in c8b_lut_load[_dev])."""
    h = np.frombuffer(np.ascontiguousarray(blob_u8)[:16].tobytes(), "<u4")
    from . import _cabi
    return int(h[0]) == LUT_MAGIC and int(h[2]) == int(_cabi.lib().c8b_lut_size())


def reduce_stats(torch, dist, world, device, sums, maxes):
    """sums: list of ints added over ranks; maxes: list of floats max-ed over ranks."""
    s = torch.tensor([int(x) for x in sums], dtype=torch.int64, device=device)
    m = torch.tensor([float(x) for x in maxes], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return [int(x) for x in s.cpu()], [float(x) for x in m.cpu()]
