"""Synthetic work items for benchmarks and scale tests, built ON THE DEVICE with torch (plumbing only):
unique clean frames made by the reference's generator (tests/golden/*.npz) are replicated, placed at a random
offset inside a zero-padded item, rotated by a per-item CFO and given AWGN -- the channel model of SURVEY 8d
(noise amplitude as tools/performance/perf_siso.py:92)."""
import numpy as np


def make_items(torch, dev, frames, counts, snr_db=30.0, cfo_hz=100e3, pre=(200, 400), post=200, seed=0, rms=0.1875):
    """frames: list of clean frames; each is one complex64 array (1 antenna) or a tuple of two (2 antennas).
    counts[t] items are made from frames[t].  Returns (iq_list [per antenna tensors], off int64[n], len int32[n], kind int32[n])."""
    nant = 2 if isinstance(frames[0], (tuple, list)) else 1
    lens = [int((f[0] if nant == 2 else f).size) + pre[1] + post for f in frames]
    n = int(sum(counts))
    item_len = np.repeat(np.array(lens, np.int64), counts)
    off = np.concatenate([[0], np.cumsum(item_len)[:-1]]).astype(np.int64)
    kind = np.repeat(np.arange(len(frames), dtype=np.int32), counts)
    total = int(item_len.sum())
    out = [torch.zeros(total, dtype=torch.complex64, device=dev) for _ in range(nant)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000003 * seed + 13579)
    sigma = rms / np.sqrt(2.0 * 10 ** (snr_db / 10)) if snr_db is not None else 0.0
    pos = 0
    for t, f in enumerate(frames):
        L, cnt = lens[t], int(counts[t])
        fa = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (f if nant == 2 else (f,))]
        Lf = fa[0].numel()
        step = max(1, (1 << 25) // L)
        for b in range(0, cnt, step):
            e = min(cnt, b + step)
            m = e - b
            p = torch.randint(pre[0], pre[1] + 1, (m, 1), generator=gen, device=dev)
            j = torch.arange(L, device=dev)[None, :] - p
            valid = (j >= 0) & (j < Lf)
            jc = j.clamp(0, Lf - 1)
            cfo = (torch.rand(m, 1, generator=gen, device=dev) * 2 - 1) * cfo_hz
            ph = (2 * np.pi / 20e6) * cfo * torch.arange(L, device=dev, dtype=torch.float32)[None, :]
            rot = torch.polar(torch.ones_like(ph), ph)
            for a in range(nant):
                x = torch.where(valid, fa[a][jc], torch.zeros((), dtype=torch.complex64, device=dev)) * rot
                if sigma:
                    x = x + torch.view_as_complex(torch.randn((m, L, 2), generator=gen, device=dev) * sigma)
                out[a][pos + b * L: pos + e * L] = x.reshape(-1)
        pos += cnt * L
    return out, off, item_len.astype(np.int32), kind
