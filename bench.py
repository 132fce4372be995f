#!/usr/bin/env python
"""bench.py — headline benchmark of the epipolar fusion path (BASELINE.json metric).

A "step" = one forward of the fusion layer over one batch of synthetic (ref, src) feature-map
pairs: BASELINE.json configs[1] = H36M 4-view ResNet-50 256x256 -> N=4 pairs, C=256, 64x64
feature map, K=64 (configs/epipolar/keypoint_h36m_zresidual_fixed.yaml shape, 'z' + ZRESIDUAL,
eval).  metric = views/s (= pairs/s); ms_per_step = forward ms.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]

N>1 (torchrun, one rank per GPU = one camera view per GPU): every rank owns `pairs` frames of
its view, all ranks exchange feature maps (NCCL all-gather over NVLink, inside the timed
region), every rank fuses its view against its nearest-neighbour view.  Weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (pairs per GPU, C, H, W, K, cfg factory name, description)
    "cfg2": dict(N=4, C=256, H=64, W=64, K=64, cfg="cfg_h36m_r50_256",
                 desc="H36M 4-view ResNet-50 256x256 (feature 64x64) C=256 K=64, z+ZRESIDUAL eval"),
    "cfg3": dict(N=4, C=256, H=96, W=96, K=64, cfg="cfg_h36m_r152_384",
                 desc="H36M 4-view ResNet-152 384x384 (feature 96x96) C=256 K=64"),
}
L2_BYTES = 126 * 1024 * 1024


def algorithmic_bytes(N, C, H, W, K, attn=True, corr=True):
    """read feat_ref + feat_src, write out (+KRTs, + emitted attn / corr_pos)."""
    b = 3 * 4 * N * C * H * W + 96 * N
    if attn:
        b += 4 * N * K * H * W
    if corr:
        b += 8 * N * H * W
    return b


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """Polls SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index=0, period=0.005):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.index, self.period = index, period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    _BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU op composition (oracle/torch_port.py)
# ----------------------------------------------------------------------------------------------
def time_cpu_port(wl, steps, warmup, threads=None):
    import torch
    import epipolar_transformers_b200 as epi
    from epipolar_transformers_b200 import synthetic as syn
    from oracle import torch_port
    cfg = getattr(epi, wl["cfg"])()
    cfg.EPIPOLAR.SAMPLESIZE = wl["K"]
    N, C, H, W = wl["N"], wl["C"], wl["H"], wl["W"]
    P1, P2 = syn.pairs_from_ring(N, 4 * H)
    f1 = torch.from_numpy(syn.features(N, C, H, W, "relu_smooth", 11))
    f2 = torch.from_numpy(syn.features(N, C, H, W, "relu_smooth", 12))
    params = syn.z_bn_params(C) if "z" in cfg.EPIPOLAR.PARAMETERIZED else None
    if not threads:
        # "all the host threads it can use": ATen's intra-op pool stops scaling (and then regresses) well
        # before a 100+-core host is full, so pick the fastest of a few pool sizes with one forward each.
        ncpu = os.cpu_count() or 1
        best = None
        for t in sorted({ncpu, max(1, ncpu // 2), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True):
            torch.set_num_threads(t)
            torch_port.forward(cfg, f1, f2, P1, P2, params=params)
            t0 = time.perf_counter()
            torch_port.forward(cfg, f1, f2, P1, P2, params=params)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, t)
        threads = best[1]
    torch.set_num_threads(threads)
    for _ in range(warmup):
        torch_port.forward(cfg, f1, f2, P1, P2, params=params)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        torch_port.forward(cfg, f1, f2, P1, P2, params=params)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(ts))
    return {"ms_per_step": ms, "views_per_s": N / (ms * 1e-3), "cores": threads,
            "sample": "%d full forwards of %s (N=%d pairs) after %d warm-up, oracle/torch_port.py "
                      "(same ATen op sequence as the reference), %d threads" % (steps, wl["desc"], N, warmup, threads)}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    r = time_cpu_port(wl, steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": "epipolar_fusion_forward_views_per_sec", "value": r["views_per_s"], "unit": "views/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "pairs": wl["N"], "C": wl["C"], "feat_hw": [wl["H"], wl["W"]], "K": wl["K"],
                   "device": "host CPU"},
        "cpu_baseline": {"value": r["views_per_s"], "unit": "views/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["views_per_s"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, wl):
    # stdout must carry exactly one JSON line: native libraries (NCCL prints its version with printf when
    # NCCL_DEBUG=VERSION is set on the box) write to fd 1, so fd 1 points at stderr until the line is printed.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import epipolar_transformers_b200 as epi
    from epipolar_transformers_b200 import synthetic as syn, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    N, C, H, W, K = wl["N"], wl["C"], wl["H"], wl["W"], wl["K"]
    cfg = getattr(epi, wl["cfg"])()
    cfg.EPIPOLAR.SAMPLESIZE = K
    has_z = "z" in cfg.EPIPOLAR.PARAMETERIZED
    model = epi.Epipolar(cfg=cfg, variant=args.variant).to(dev).eval()
    if has_z:
        sd = {k: torch.from_numpy(v) for k, v in syn.z_bn_params(C).items()}
        model.load_state_dict(sd, strict=False)

    # cameras: `max(world,4)` views on a ring; single-GPU: pair v = (view v, nearest view), like the
    # H36M test-time batch.  Multi-GPU: rank r owns view r, N frames of it.
    # The same 4-camera H36M-like ring at every N, so per-GPU work does not change with N (weak scaling): at N=1
    # the 4 pairs are (cam v, nearest cam) for v=0..3; at N>1 rank r owns camera r%4 (N frames of it) and fuses
    # against the rank that owns its nearest camera inside the same group of 4 ranks.
    ring = syn.ring_cameras(4, 4 * H)
    ring_src = syn.nearest_source(ring)
    if world == 1:
        src_of = ring_src
        P_ref = torch.from_numpy(ring.astype(np.float32)).to(dev)
        P_src = torch.from_numpy(ring[ring_src].astype(np.float32)).to(dev)
    else:
        cam = np.arange(world) % 4
        src_of = (np.arange(world) // 4) * 4 + ring_src[cam]
        src_of = np.where(src_of < world, src_of, (np.arange(world) // 4) * 4 + (cam ^ 1))      # world not a multiple of 4
        src_of = np.where(src_of < world, src_of, (np.arange(world) + 1) % world)
        P_ref = torch.from_numpy(np.repeat(ring[cam[rank]][None], N, 0).astype(np.float32)).to(dev)
        P_src = torch.from_numpy(np.repeat(ring[cam[src_of[rank]]][None], N, 0).astype(np.float32)).to(dev)

    # rotating input sets so consecutive steps never find their inputs in L2
    set_bytes = 2 * N * C * H * W * 4
    n_sets = max(2, int(np.ceil(3.0 * L2_BYTES / set_bytes)))
    n_sets = min(n_sets, 24)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    symm_hdls = None
    if world > 1 and args.exchange == "peer":
        try:
            import torch.distributed._symmetric_memory as symm
            refs, symm_hdls = [], []
            for _ in range(n_sets):                       # the "backbone output" buffers live in peer-mapped memory
                t = symm.empty(N, C, H, W, dtype=torch.float32, device=dev)
                symm_hdls.append(symm.rendezvous(t, dist.group.WORLD))
                t.copy_(torch.relu(torch.randn(N, C, H, W, device=dev, generator=gen)))
                refs.append(t)
            ok = torch.tensor([1], device=dev)
        except Exception as e:                            # no peer mapping on this box: fall back to the collective
            ok = torch.tensor([0], device=dev)
            sys.stderr.write("bench.py: symmetric memory unavailable (%r), using all-gather\n" % (e,))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            args.exchange = "allgather"
            symm_hdls = None
            refs = [torch.relu(torch.randn(N, C, H, W, device=dev, generator=gen)) for _ in range(n_sets)]
    else:
        refs = [torch.relu(torch.randn(N, C, H, W, device=dev, generator=gen)) for _ in range(n_sets)]
    srcs = [torch.relu(torch.randn(N, C, H, W, device=dev, generator=gen)) for _ in range(n_sets)]
    gathered_flat = torch.empty((world * N, C, H, W), device=dev) if world > 1 else None
    gathered = gathered_flat.view(world, N, C, H, W) if world > 1 else None

    recv_buf = torch.empty((N, C, H, W), device=dev) if world > 1 else None
    consumers = [r for r in range(world) if int(src_of[r]) == rank] if world > 1 else []

    def exchange(f_ref, i=0):
        """the path's only exchange step: this rank needs the feature map of its source view"""
        if args.exchange == "peer" and symm_hdls is not None and f_ref is refs[i % n_sets]:
            h = symm_hdls[i % n_sets]
            h.barrier(channel=0)                                        # every rank's map for this step is in place
            return h.get_buffer(int(src_of[rank]), (N, C, H, W), torch.float32)   # neighbour's HBM, read over NVLink by the staging kernel
        if args.exchange == "allgather":
            dist.all_gather_into_tensor(gathered_flat, f_ref)           # every map to every rank (BASELINE config 4 / MULTITEST)
            return gathered[src_of[rank]]
        ops = [dist.P2POp(dist.irecv, recv_buf, int(src_of[rank]))] + [dist.P2POp(dist.isend, f_ref, r) for r in consumers]
        for req in dist.batch_isend_irecv(ops):                         # NCCL send/recv permutation over NVLink
            req.wait()
        return recv_buf

    def step(i):
        f_ref = refs[i % n_sets]
        if world > 1:
            f_src = exchange(f_ref, i)
        else:
            f_src = srcs[i % n_sets]
        with torch.no_grad():
            return model(f_ref, f_src, P_ref, P_src)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(warmup):
        step(i)
    launches_per_step = lib.epi_last_launch_count()
    barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(warmup + i)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # ---- dominant-kernel duration, CUDA events on the launching stream (fusion call only, no epilogue) ----
    kern_ms = None
    pre = dict(K=K, downsample=cfg.BACKBONE.DOWNSAMPLE, softmax_scale=cfg.EPIPOLAR.SOFTMAXSCALE,
               correct_normalize=cfg.EPIPOLAR.USE_CORRECT_NORMALIZE, variant=args.variant)
    def fusion_call(i):           # the same call as the headline step (same outputs, same epilogue), local source map
        with torch.no_grad():
            return model(refs[i % n_sets], (srcs if world == 1 else refs)[(i + 1) % n_sets], P_ref, P_src)
    for i in range(3):
        fusion_call(i)
    torch.cuda.synchronize()
    # (a) CUDA events recorded by the library on the launching stream around the fused attention kernel alone,
    # (b) events around the whole fusion call (operand staging + pixel ordering + fused kernel) as a cross-check
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_only = []
    lib.epi_kernel_timing_enable(1)
    for i, (a, b) in enumerate(evs):
        a.record()
        fusion_call(i + 3)
        b.record()
        t_k = float(lib.epi_kernel_timing_last_ms())
        if t_k > 0:
            kern_only.append(t_k)
    lib.epi_kernel_timing_enable(0)
    torch.cuda.synchronize()
    call_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    kern_ms = float(np.mean(kern_only)) if kern_only else call_ms

    # ---- end to end through the public module call with HOST buffers (pinned), copies inside the timed region ----
    h_ref = [refs[i].cpu().pin_memory() for i in range(2)]
    h_src = [srcs[i].cpu().pin_memory() for i in range(2)]
    h_P1, h_P2 = P_ref.cpu().pin_memory(), P_src.cpu().pin_memory()
    h_out = torch.empty((N, C, H, W), dtype=torch.float32).pin_memory()
    h_attn = torch.empty((N, K, H, W), dtype=torch.float32).pin_memory()
    h_corr = torch.empty((N, H, W, 2), dtype=torch.float32).pin_memory()

    streamer = epi.HostStreamer(model, dev, depth=2) if world == 1 else None

    def e2e_step(i):
        if streamer is not None:        # H2D of step i+1 overlaps kernels + D2H of step i
            streamer(h_ref[i % 2], h_src[i % 2], h_P1, h_P2, h_out, h_attn, h_corr)
            return
        d_ref = h_ref[i % 2].to(dev, non_blocking=True)
        d_src = exchange(d_ref)
        d_P1 = h_P1.to(dev, non_blocking=True); d_P2 = h_P2.to(dev, non_blocking=True)
        with torch.no_grad():
            o, c, a, _ = model(d_ref, d_src, d_P1, d_P2)
        h_out.copy_(o, non_blocking=True); h_attn.copy_(a, non_blocking=True); h_corr.copy_(c, non_blocking=True)

    def e2e_sync():
        if streamer is not None:
            streamer.synchronize()
        barrier()

    e2e_steps = max(5, min(args.steps, 50))
    for i in range(3):
        e2e_step(i)
    e2e_sync()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    if streamer is not None:
        torch.cuda.current_stream().wait_stream(streamer.s_out)
    e1.record()
    e2e_sync()
    wall_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    e2e_wall_ms = wall_ms / e2e_steps
    clocks = sampler.stop() if sampler else None
    h2d = (2 if world == 1 else 1) * N * C * H * W * 4 + 2 * N * 48
    d2h = N * C * H * W * 4 + N * K * H * W * 4 + N * H * W * 8

    if rank == 0:
        peak, peak_src = measured_peaks()
        balg = algorithmic_bytes(N, C, H, W, K)
        achieved = balg / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(wl["name"], {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "epipolar_fusion_forward_views_per_sec", "value": world * N / (ms_step * 1e-3), "unit": "views/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "pairs_per_gpu": N, "C": C, "feat_hw": [H, W], "K": K,
                       "parallelism": ("1 view per GPU, NCCL %s of per-view feature maps" % ({"allgather": "all-gather", "p2p": "send/recv (each rank receives only its source view)", "peer": "symmetric memory: kernels read the source view from the neighbour GPU over NVLink, 1 barrier/step"}[args.exchange])) if world > 1 else "single GPU",
                       "l2": "rotating %d input sets (%.0f MB > 126 MB L2), no reuse between consecutive steps" % (n_sets, n_sets * set_bytes / 1e6),
                       "outputs": "finalout + attn + corr_pos", "variant": args.variant},
            "clocks": clocks,
            "e2e": {"value": world * N / (e2e_ms * 1e-3), "unit": "views/s", "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall_ms,
                    "how": "pinned host buffers -> HostStreamer(Epipolar) -> pinned host buffers; H2D of step i+1 overlaps kernels + D2H of step i (2 streams)" if world == 1 else "pinned host -> device, exchange, Epipolar, device -> pinned host, one stream",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "fused epipolar attention kernel (geometry+taps+softmax+AV)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes": balg, "kernel_ms": kern_ms, "fusion_call_ms": call_ms,
                         "timing": "CUDA events on the launching stream around the fused attention kernel (recorded inside the C ABI call); fusion_call_ms also covers operand staging and pixel ordering",
                         "peak_source": peak_src},
        }
        if world == 1 and not args.no_cpu_baseline:
            r = time_cpu_port(wl, args.cpu_steps, 1)
            line["cpu_baseline"] = {"value": r["views_per_s"], "unit": "views/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"], "ms_per_step": r["ms_per_step"]}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="auto", choices=["auto", "warp", "tile", "sector"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "p2p", "allgather"], help="multi-GPU exchange of per-view feature maps")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload], name=args.workload)
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
