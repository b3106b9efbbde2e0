#!/usr/bin/env python
"""bench.py -- RX throughput of the 802.11 OFDM receive chain on B200 (BASELINE.json metric:
"RX IQ samples/s & frames/s (20 MHz VHT MCS7)", workload = configs[4]: VHT MCS7 1500-byte frames as
independent work items, 1M-frame batch per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # N>1: launched under torchrun, one rank per GPU
  python bench.py --impl reference ...                         # CPU arm: the reference's own blocks on all host cores

A step = one pass of the whole chain (presiso -> trigger/sync/signal -> demod -> decode) over one batch.
`value`  : samples/s with the batch resident in HBM (results stay on the device), CUDA events on the
           launching stream, max over ranks, whole-job aggregate.
`e2e`    : the same through c8b_rx_batch with HOST (pinned) buffers: H2D of the IQ and D2H of the frame
           records + PDU bytes inside the timed region.
`roofline`: the dominant kernel (Viterbi decode), algorithmic bytes / measured launch time vs the measured
           HBM peak; `stages` gives every kernel (the HBM-streaming ones against the same peak).
`cpu_baseline`: the reference's own receive blocks (compiled unmodified into oracle/_ref) on this host's cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GAP = 200                    # zeros before and after each frame (400-sample gaps)
FRAME_SAMP = 4560            # VHT MCS7, 1500-byte MPDU: 47 symbols (SURVEY 8 sizes)
ITEM = FRAME_SAMP + 2 * GAP
NSYM, NCBPS, TRELLIS, TOTAL_LLR, MPDU_LEN = 47, 312, 12220, 14664, 1500
PDU_STRIDE = 1536
SNR_DB = 30.0
WORKLOAD = "configs[4]: VHT MCS7 1500-byte MPDU (47 sym, 4560 samples + 400 gap), AWGN 30 dB, CFO U(-100,100) kHz"


def cpu_model():
    """host CPU model string (the CPU baseline names the host it ran on)"""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def ncu_profile():
    """the newest committed ncu --set full summary (profiles/ncu_rNN.json)"""
    for name in ("ncu_r02.json", "ncu_r01.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return p
    return None


def ncu_warp_instructions(kernel):
    """warp instructions one launch of `kernel` executed in that capture (a property of the kernel on this workload), or None"""
    try:
        for k in json.load(open(ncu_profile())):
            if k["kernel"] == kernel:
                return float(str(k["warp_instructions"]).split()[0]), float(str(k.get("frames_per_launch", 56832)).split()[0])
    except Exception:
        pass
    return None, None


def ncu_traffic(kernel):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set full summary (profiles/ncu_rNN.json), or None"""
    p = ncu_profile()
    try:
        for k in json.load(open(p)):
            if k["kernel"] == kernel:
                def val(sv):
                    x, u = sv.split()[:2]
                    return float(x) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                return val(k["dram_read"]) + val(k["dram_write"])
    except Exception:
        pass
    return None


def api_arms():
    """The reference-facing API of north_star beside the batched call (rank 0, N = 1, outside the timed region): the live-stream
    session fed 1M-sample pushes (tools/bench_stream.py) and the seven gr::block shells under the thread-per-block mock
    scheduler (tools/bench_blocks.py), both on the reference's dense demo capture; 20 M samples/s = real time of one channel."""
    out = {}
    for key, cmd, pick in (("e2e_stream", [sys.executable, os.path.join(ROOT, "tools", "bench_stream.py"), "1048576"], "1048576"),
                           ("e2e_blocks", [sys.executable, os.path.join(ROOT, "tools", "bench_blocks.py"), "8192", "16"], "thread_per_block")):
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
            d = json.loads(line[7:])[pick]
            out[key] = {"value": d["samples_per_s"], "unit": "samples/s", "x_real_time_20MHz": d["samples_per_s"] / 20e6,
                        "api": ("c8b_stream_push, 1 Mi-sample pushes from pinned memory, %d of %d PDUs" % (d["pdus"], d["sent"])) if key == "e2e_stream"
                        else ("the seven gr::block shells (c8b_blk_work), thread per block, %d messages from %d samples" % (d["messages"], d["samples"]))}
        except Exception as e:                                       # noqa: BLE001 -- an arm that cannot run reports why
            out[key] = {"value": None, "error": repr(e)[:200]}
    return out


def bind_near_gpu(torch, index):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the PCI device) BEFORE any pinned host buffer is
    allocated: first-touch then places the staging memory on the GPU's own NUMA node, so N ranks do not all pull their
    H2D traffic out of one socket's memory.  Returns a short description for the JSON line (None when sysfs says nothing)."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        cpus = open(base + "/local_cpulist").read().strip()
        node = open(base + "/numa_node").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= set(os.sched_getaffinity(0))
        if ids:
            os.sched_setaffinity(0, ids)
            return {"pci": bdf, "numa_node": int(node), "cpus": len(ids)}
    except Exception:
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_batch_device(torch, dev, nframes, seed):
    """Synthetic config-5 batch built ON THE DEVICE from the 16 unique frames the reference's generator made
    (tests/golden/frames_bench.npz): per item gather + CFO + AWGN.  Returns complex64 tensor [nframes*ITEM]."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "frames_bench.npz"))
    base = torch.zeros((16, ITEM), dtype=torch.complex64, device=dev)
    base[:, GAP:GAP + FRAME_SAMP] = torch.from_numpy(g["iq"]).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(13579 + seed)
    sigma = 0.1875 / np.sqrt(2.0 * 10 ** (SNR_DB / 10))           # tools/performance/perf_siso.py:92
    out = torch.empty(nframes * ITEM, dtype=torch.complex64, device=dev)
    nidx = torch.arange(ITEM, device=dev, dtype=torch.float32)
    step = 16384
    for b in range(0, nframes, step):
        e = min(nframes, b + step)
        k = torch.arange(b, e, device=dev) % 16
        cfo = (torch.rand(e - b, generator=gen, device=dev) * 2 - 1) * 100e3
        ph = (2 * np.pi / 20e6) * cfo[:, None] * nidx[None, :]
        x = base[k] * torch.polar(torch.ones_like(ph), ph)
        noise = torch.randn((e - b, ITEM, 2), generator=gen, device=dev) * sigma
        x = x + torch.view_as_complex(noise)
        out[b * ITEM:e * ITEM] = x.reshape(-1)
    return out, g["mpdu"]


def make_batch_tx(torch, dev, pkg, rx, nframes, seed):
    """Synthetic config-5 batch made entirely ON THE DEVICE by the transmit synthesiser: every item carries its OWN random
    1500-byte MPDU (valid FCS, one-MPDU A-MPDU), VHT MCS7, a CFO drawn from U(-100, 100) kHz, AWGN at 30 dB.
    Returns (iq tensor [nframes*ITEM], psdu tensor [nframes, 1504] = what every frame must decode to)."""
    rng = np.random.default_rng(13579 + seed)
    d = np.zeros(nframes, pkg.TXFRAME_DTYPE)
    d["format"], d["mcs"], d["psdu_len"] = 2, 7, MPDU_LEN + 4
    d["psdu_off"] = np.arange(nframes, dtype=np.int64) * (MPDU_LEN + 4)
    d["out_off"] = np.arange(nframes, dtype=np.int64) * ITEM + GAP
    d["cfo_hz"] = rng.uniform(-100e3, 100e3, nframes).astype(np.float32)
    psdu = torch.zeros(nframes * (MPDU_LEN + 4) + 16, dtype=torch.uint8, device=dev)
    out = torch.zeros(nframes * ITEM, dtype=torch.complex64, device=dev)
    torch.cuda.synchronize()
    rx.tx_random_psdu_dev(psdu.data_ptr(), nframes * (MPDU_LEN + 4), d, seed=0x80211 + seed)
    rx.tx_batch_dev(psdu.data_ptr(), nframes * (MPDU_LEN + 4), d, out.data_ptr(), nframes * ITEM, multiplier=12.0, seed=93)
    rx.sync()
    gen = torch.Generator(device=dev)
    gen.manual_seed(13579 + seed)
    sigma = 0.1875 / np.sqrt(2.0 * 10 ** (SNR_DB / 10))           # tools/performance/perf_siso.py:92
    step = 16384 * ITEM
    ov = torch.view_as_real(out)
    for b in range(0, nframes * ITEM, step):
        e = min(nframes * ITEM, b + step)
        ov[b:e] += torch.randn((e - b, 2), generator=gen, device=dev) * sigma
    return out, psdu[:nframes * (MPDU_LEN + 4)].view(nframes, MPDU_LEN + 4)


def cpu_batch(nframes, seed=0):
    """config-5 items for the CPU arm: the generator's 16 frames (tests/golden/frames_bench.npz), per-item CFO + AWGN at 30 dB"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "frames_bench.npz"))
    rng = np.random.default_rng(seed)
    sigma = 0.1875 / np.sqrt(2.0 * 10 ** (SNR_DB / 10))
    iq = np.zeros((nframes, ITEM), np.complex64)
    n = np.arange(ITEM)
    for i in range(nframes):
        iq[i, GAP:GAP + FRAME_SAMP] = g["iq"][i % 16]
        cfo = rng.uniform(-100e3, 100e3)
        iq[i] *= np.exp(2j * np.pi * cfo * n / 20e6).astype(np.complex64)
    iq = (iq + sigma * (rng.standard_normal(iq.shape) + 1j * rng.standard_normal(iq.shape))).astype(np.complex64).reshape(-1)
    return iq, (np.arange(nframes) * ITEM).astype(np.int64), np.full(nframes, ITEM, np.int32)


def cpu_kind():
    """"reference": the reference's own seven blocks (lib/*_impl.cc, unmodified) compiled into oracle/_ref/libgr80211_ref.so
    and run by oracle/ref_chain.cc's scheduler; "port": the restated oracle (only when that library was never built)."""
    import oracle_lib as ol
    return "reference" if ol.have_refchain() else "port"


def cpu_arm(nframes, threads, seed=0, kind=None):
    """The CPU arm on `nframes` config-5 items, frame-parallel over `threads` host threads; returns (seconds, frames ok).
    kind "reference": every thread owns one chain of the reference's blocks and is fed its run of items as one stream,
    presiso included (refchain_bench); kind "port": orx_rx_batch of the restated oracle."""
    import oracle_lib as ol
    kind = kind or cpu_kind()
    iq, off, ln = cpu_batch(nframes, seed)
    if kind == "reference":
        L = ol.refchain_lib()
        ol.oracle()
        cnt = np.zeros(3, np.int64)
        t0 = time.perf_counter()
        rc = L.refchain_bench(ol.c2f(iq), off, ln, nframes, threads, cnt)
        dt = time.perf_counter() - t0
        if rc:
            raise RuntimeError("refchain_bench failed")
        return dt, int(cnt[0])
    fr = np.zeros(nframes, ol.FRAME_DTYPE)
    pdu = np.zeros(nframes * PDU_STRIDE, np.uint8)
    L = ol.oracle()
    t0 = time.perf_counter()
    L.orx_rx_batch(ol.c2f(iq), off, ln, nframes, threads, fr.ctypes.data, pdu, PDU_STRIDE)
    dt = time.perf_counter() - t0
    return dt, int((fr["npdu"] == 1).sum())


CPU_SAMPLE = {"reference": "the reference's own blocks (lib/{trigger,sync,signal,demod,decode}_impl.cc + cloud80211phy.cc, unmodified, "
                           "oracle/_ref/libgr80211_ref.so) under oracle/ref_chain.cc's scheduler, one chain per thread, presiso included",
              "port": "oracle/liboracle_rx.so (orx_rx_batch, the restated chain)"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path -- its seven receive blocks compiled unmodified
    (oracle/_ref/libgr80211_ref.so, built by oracle/Makefile where /root/reference is mounted and shipped prebuilt) -- on all
    host threads, frame-parallel (one chain of blocks per thread, which is kinder to it than GNU Radio's thread-per-block
    scheduler, where the decode block's thread bounds the flowgraph)."""
    if rank != 0:
        return
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    dt, _ = cpu_arm(2 * cores, cores, kind=kind)                            # calibrate
    per_step = int(max(cores, min(4096, (2 * cores / dt) * 3.0)))   # ~3 s of CPU work per step
    for _ in range(args.warmup):
        cpu_arm(max(cores, per_step // 4), cores, kind=kind)
    t = 0.0
    ok = 0
    for s in range(args.steps):
        dt, k = cpu_arm(per_step, cores, seed=s, kind=kind)
        t += dt
        ok += k
    v = per_step * args.steps * ITEM / t
    line = {
        "impl": "reference", "metric": "rx_iq_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "frames_per_s": per_step * args.steps / t,
        "config": {"workload": WORKLOAD, "frames_per_step": per_step, "samples_per_item": ITEM, "frames_ok": ok},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
                         "cpu_model": cpu_model(),
                         "sample": "%d steps x %d config-5 items through %s, %d threads" % (args.steps, per_step, CPU_SAMPLE[kind], cores)},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per GPU per step (config 5: 1M-frame batch)")
    ap.add_argument("--e2e-frames", type=int, default=4 * 56832, help="frames per host-buffer call of the e2e arm (4 pipeline chunks)")
    ap.add_argument("--chunk", type=int, default=56832, help="items per pipeline pass inside the library (3 CTAs x 148 SMs x 128 frames: one full wave of k_viterbi_tp)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --frames per GPU per step (the driver's default contract); strong: ONE batch of --frames frames per step, "
                         "sharded over the ranks by item index (BASELINE configs[4]: the 1M-frame batch split over 1/2/4/8 GPUs)")
    ap.add_argument("--synth", default="tx", choices=["tx", "golden"],
                    help="tx: every frame unique, made on the device by the transmit synthesiser; golden: the generator's 16 frames replicated")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_pkg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    pkg = load_pkg()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(torch, local)                               # pinned staging allocated below lands on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- LUT: built on rank 0, broadcast over NCCL, loaded from device memory (the only exchange) ----
    blob = pkg.parallel.broadcast_lut(torch, dist, rank, world, dev)
    blob_n = blob.numel()
    rx = pkg.Receiver(device=local, chunk_items=args.chunk, blob=(blob.cpu().numpy() if world == 1 else None))
    if world > 1:
        torch.cuda.synchronize()
        rx.load_lut_device(blob.data_ptr(), blob_n)

    nfr = args.frames
    if args.scaling == "strong":                                     # rank r takes the contiguous shard r of the one batch
        lo, hi = pkg.parallel.shard_range(args.frames, rank, world)
        nfr = hi - lo
    psdu_sent = None
    if args.synth == "tx":
        iq, psdu_sent = make_batch_tx(torch, dev, pkg, rx, nfr, seed=rank)
    else:
        iq, mpdus = make_batch_device(torch, dev, nfr, seed=rank)
    off = (np.arange(nfr, dtype=np.int64) * ITEM)
    ln = np.full(nfr, ITEM, np.int32)
    d_frames = torch.zeros(nfr * pkg.FRAME_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_pdu = torch.zeros(nfr * PDU_STRIDE, dtype=torch.uint8, device=dev)
    st = torch.cuda.ExternalStream(rx.stream, device=dev)
    torch.cuda.synchronize()

    def step():
        rx.rx_batch_dev_async(iq.data_ptr(), off, ln, d_frames.data_ptr(), d_pdu.data_ptr(), PDU_STRIDE)

    def barrier():
        if world > 1:
            dist.barrier()
        rx.sync()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    barrier()
    ms = e0.elapsed_time(e1)

    # ---- per-kernel device times: one extra pass with every kernel on ONE stream (the timed steps above overlap the
    #      decode kernel of chunk k with the front end of chunk k+1, which would smear per-kernel event times) ----
    rx_t = pkg.Receiver(device=local, chunk_items=args.chunk, blob=blob.cpu().numpy(), overlap=False)
    rx_t.rx_batch_dev_async(iq.data_ptr(), off, ln, d_frames.data_ptr(), d_pdu.data_ptr(), PDU_STRIDE)
    rx_t.sync()
    rx_t.timing(True)
    rx_t.timing_read(reset=True)
    st_t = torch.cuda.ExternalStream(rx_t.stream, device=dev)
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record(st_t)
    rx_t.rx_batch_dev_async(iq.data_ptr(), off, ln, d_frames.data_ptr(), d_pdu.data_ptr(), PDU_STRIDE)
    t1e.record(st_t)
    rx_t.sync()
    torch.cuda.synchronize()
    ms_serial = t0e.elapsed_time(t1e)
    stage = rx_t.timing_read(reset=True)
    rx_t.close()

    # ---- correctness of what was just timed (outside the timed region) ----
    fr = np.frombuffer(d_frames.cpu().numpy().tobytes(), dtype=pkg.FRAME_DTYPE)
    ok = (fr["status"] == 0) & (fr["npdu"] == 1) & (fr["pdu_bytes"] == MPDU_LEN + 4)
    frames_ok = int(ok.sum())
    if psdu_sent is not None:                                        # EVERY decoded MPDU against the bytes its frame was built from
        okd = torch.from_numpy(ok).to(dev)
        same = torch.zeros(nfr, dtype=torch.bool, device=dev)
        for b in range(0, nfr, 65536):
            e = min(nfr, b + 65536)
            same[b:e] = (d_pdu.view(nfr, PDU_STRIDE)[b:e, 3:3 + MPDU_LEN] == psdu_sent[b:e, 4:4 + MPDU_LEN]).all(dim=1)
        bytes_ok, nchk = int((same & okd).sum().item()), frames_ok
    else:
        chk = np.random.default_rng(1).choice(nfr, size=min(nfr, 512), replace=False)
        pd = d_pdu.view(nfr, PDU_STRIDE)[torch.from_numpy(chk).to(dev)].cpu().numpy()
        bytes_ok = sum(int(bytes(pd[j, 3:3 + MPDU_LEN]) == bytes(mpdus[int(i) % 16])) for j, i in enumerate(chk) if ok[i])
        nchk = int(ok[chk].sum())

    # ---- e2e: host (pinned) buffers through c8b_rx_batch, H2D + D2H inside the timed region ----
    ne = min(args.e2e_frames, nfr)
    calls = max(1, nfr // ne)
    h_iq = torch.empty(ne * ITEM, dtype=torch.complex64, pin_memory=True)
    h_iq.copy_(iq[:ne * ITEM])
    h_np = h_iq.numpy()
    h_frames = torch.zeros(ne * pkg.FRAME_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    h_pdu = torch.zeros(ne * PDU_STRIDE, dtype=torch.uint8, pin_memory=True)
    fr_np = np.frombuffer(h_frames.numpy(), dtype=pkg.FRAME_DTYPE)
    L = pkg._cabi.lib()
    import ctypes as C
    off_e, ln_e = off[:ne].copy(), ln[:ne].copy()

    def e2e_step():
        for _ in range(calls):
            rc = L.c8b_rx_batch(rx.h, C.c_void_p(h_np.ctypes.data), pkg._cabi.ptr(off_e), pkg._cabi.ptr(ln_e), ne, C.c_void_p(fr_np.ctypes.data),
                                C.c_void_p(h_pdu.data_ptr()), PDU_STRIDE)
            if rc:
                raise RuntimeError("c8b_rx_batch: %d" % rc)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_ok = int(((fr_np["status"] == 0) & (fr_np["npdu"] == 1)).sum())
    fc32_records = fr_np[["status", "npdu", "pdu_bytes", "mcs", "len"]].copy()

    # ---- the same through c8b_rx_batch_sc16: the capture in the radio's wire format (interleaved int16), half the H2D bytes ----
    h16 = torch.empty((ne * ITEM, 2), dtype=torch.int16, pin_memory=True)
    for b0 in range(0, ne * ITEM, 1 << 26):
        e0 = min(ne * ITEM, b0 + (1 << 26))
        h16[b0:e0].copy_(torch.view_as_real(iq[b0:e0]).mul(32768.0).round_().clamp_(-32768, 32767).to(torch.int16))
    h16_np = h16.numpy()

    def sc16_step():
        for _ in range(calls):
            rc = L.c8b_rx_batch_sc16(rx.h, C.c_void_p(h16_np.ctypes.data), pkg._cabi.ptr(off_e), pkg._cabi.ptr(ln_e), ne, C.c_void_p(fr_np.ctypes.data),
                                     C.c_void_p(h_pdu.data_ptr()), PDU_STRIDE)
            if rc:
                raise RuntimeError("c8b_rx_batch_sc16: %d" % rc)

    sc16_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sc16_step()
    barrier()
    sc16_s = time.perf_counter() - t0
    sc16_ok = int(((fr_np["status"] == 0) & (fr_np["npdu"] == 1)).sum())
    clocks = sampler.stop() if rank == 0 else None

    # ---- reduce over ranks: max time, sums of work ----
    cnt, tt = pkg.parallel.reduce_stats(torch, dist, world, dev, [frames_ok, nfr, nfr * ITEM, bytes_ok, nchk, e2e_ok, sc16_ok, calls * ne], [ms, e2e_s * 1e3, sc16_s * 1e3])
    ms, e2e_ms, sc16_ms = tt
    frames_ok, frames_total, samples_total, bytes_ok, nchk, e2e_ok, sc16_ok, e2e_frames_total = cnt

    if rank == 0:
        peak, peak_src = peaks()
        k = args.steps
        value = samples_total * k / (ms * 1e-3)
        e2e_v = e2e_frames_total * ITEM * e2e_steps / (e2e_ms * 1e-3)
        nch = (nfr + args.chunk - 1) // args.chunk
        launches = (sum(v[1] for v in stage.values()) + nch) * k * world   # kernels launched inside the timed region, all ranks: 5 stage kernels + k_ndp per chunk
        # algorithmic bytes per launch, per kernel; one launch covers one chunk of items
        per_launch_items = nfr / nch
        alg = {
            "presiso": per_launch_items * ITEM * (8 + 4),                       # 8 B/sample in, 4 B preac out
            "detect": per_launch_items * ITEM * 4,                              # preac scan (+O(1) per frame)
            "header": per_launch_items * 560 * 8,                               # SIG fields + LTF: 8 B/sample once
            "demod": per_launch_items * NSYM * (640 + 4 * NCBPS),               # 1888 B/symbol
            "viterbi": per_launch_items * (4 * TOTAL_LLR + MPDU_LEN + 4),       # LLR read + PDU written
        }
        stages = {}
        for name, (tms, n) in stage.items():
            if n:
                per = tms / n
                gbs = alg[name] / (per * 1e-3) / 1e9
                stages[name] = {"ms_per_launch": per, "launches": n, "share": tms / ms_serial, "alg_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
        vit = stages.get("viterbi", {})
        acs = per_launch_items * TRELLIS * 64 / (vit.get("ms_per_launch", 1) * 1e-3) if vit else None
        # the bound that actually holds for the decode kernel: warp-instruction issue (one per cycle per SM sub-partition)
        winst, wframes = ncu_warp_instructions("k_viterbi_tp")
        issue = None
        if vit and winst and clocks and clocks.get("sm_mhz"):
            nsm = torch.cuda.get_device_properties(local).multi_processor_count
            ach = winst * (per_launch_items / wframes) / (vit["ms_per_launch"] * 1e-3)
            pk = nsm * 4 * clocks["sm_mhz"] * 1e6
            issue = {"bound": "issue", "achieved": ach, "peak": pk, "unit": "warp-inst/s", "frac": ach / pk,
                     "how": "warp instructions per launch from the committed ncu capture (%s) x frames, / the launch time measured here; "
                            "peak = SMs x 4 schedulers x the SM clock sampled during the run" % os.path.basename(ncu_profile() or "none")}
        line = {
            "metric": "rx_iq_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": k, "warmup": args.warmup,
            "ms_per_step": ms / k, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_s": frames_total * k / (ms * 1e-3),
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": nfr, "frames_per_step_all_gpus": frames_total, "samples_per_item": ITEM,
                       "chunk_items": args.chunk,
                       "l2": "inputs larger than L2 (%.1f GB of IQ per GPU per step)" % (nfr * ITEM * 8 / 1e9),
                       "input": ("every frame unique: random 1500-byte MPDUs modulated on the device by the transmit synthesiser (c8b_tx_batch_dev)"
                                 if args.synth == "tx" else "the reference generator's 16 frames replicated on the device"),
                       "frames_ok": frames_ok, "frames_total": frames_total, "mpdu_bytes_checked": "%d/%d identical" % (bytes_ok, nchk)},
            "e2e": {"value": e2e_v, "unit": "samples/s", "h2d_bytes_per_step": calls * ne * ITEM * 8,
                    "d2h_bytes_per_step": calls * ne * (pkg.FRAME_DTYPE.itemsize + PDU_STRIDE), "frames_per_s": e2e_frames_total * e2e_steps / (e2e_ms * 1e-3),
                    "steps": e2e_steps, "calls_per_step": calls, "frames_per_call": ne, "frames_ok_last_call": e2e_ok,
                    "api": "c8b_rx_batch (host pinned buffers; H2D double-buffered per chunk, D2H of frame records + PDU bytes)"},
            "e2e_sc16": {"value": e2e_frames_total * ITEM * e2e_steps / (sc16_ms * 1e-3), "unit": "samples/s",
                         "h2d_bytes_per_step": calls * ne * ITEM * 4, "d2h_bytes_per_step": calls * ne * (pkg.FRAME_DTYPE.itemsize + PDU_STRIDE),
                         "frames_ok_last_call": sc16_ok,
                         "api": "c8b_rx_batch_sc16 (host pinned int16 I/Q = UHD sc16, widened on the device by x / 32768; 16-bit quantisation of the same batch)"},
            "host_binding": numa,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_viterbi_tp", "achieved": vit.get("alg_GBps"), "peak": peak, "unit": "GB/s",
                         "frac": vit.get("frac_of_hbm_peak"), "traffic": ncu_traffic("k_viterbi_tp"), "peak_source": peak_src,
                         "note": "the dominant kernel (soft-Viterbi decode, one thread per frame) is bound by instruction issue / the ALU "
                                 "pipe (64-state add-compare-select per trellis step), not by HBM and not by tensor cores: see acs_per_s and "
                                 "profiles/; its DRAM traffic above the algorithmic bytes is the survivor memory of the full traceback "
                                 "(8 B per trellis step written and read back).  The HBM-streaming kernel is 'demod' in stages.",
                         "acs_per_s": acs, "issue": issue},
            "stages": stages,
            "stages_note": "per-kernel times from one extra single-stream pass (%.1f ms/step); the timed steps overlap k_viterbi of "
                           "chunk k with the front end of chunk k+1 on a second stream" % ms_serial,
            "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            kind = cpu_kind()
            dt, _ = cpu_arm(2 * cores, cores, kind=kind)
            nb = int(max(cores, min(16384, (2 * cores / dt) * 12.0)))       # ~12 s of CPU work
            dt, okc = cpu_arm(nb, cores, seed=1, kind=kind)
            n1 = max(8, nb // (4 * cores))
            dt1, _ = cpu_arm(n1, 1, seed=2, kind=kind)
            line["cpu_baseline"] = {"value": nb * ITEM / dt, "unit": "samples/s", "cores": cores, "kind": kind,
                                    "frames_per_s": nb / dt, "single_thread_samples_per_s": n1 * ITEM / dt1,
                                    "cpu_model": cpu_model(),
                                    "sample": "%d config-5 items through %s, %d threads, %d decoded" % (nb, CPU_SAMPLE[kind], cores, okc)}
            line.update(api_arms())
            if kind == "reference":                                          # the restated oracle beside it, for the record
                dtp, _ = cpu_arm(nb, cores, seed=1, kind="port")
                line["cpu_baseline"]["port_samples_per_s"] = nb * ITEM / dtp
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    rx.close()


if __name__ == "__main__":
    main()
