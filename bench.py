#!/usr/bin/env python
"""bench.py -- headline benchmark of the analytic renderer (warp + alpha-over, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c3]

Metric (BASELINE.json): composited layer-Mpix/s, fwd+bwd = B*L*H*W / 1e6 / t(fwd+bwd).
One "step" = one forward + one backward of the render path over one batch of synthetic layer
stacks.  At N>1 every rank renders its own batch shard (weak scaling, no data-path collective;
SURVEY.md 8e) and the value is sum(units)/max(time).  Prints ONE JSON line on rank 0.

Arms:
  default           the CUDA path through the C ABI; `value` with inputs resident in HBM,
                    `e2e` with pinned HOST buffers and H2D/D2H copies inside the timed region.
  --impl reference  the reference's own CPU implementation of the path (the torch port of the
                    reference chain in oracle/torch_chain.py -- /root/reference does not exist on
                    the GPU box) on all host cores, on a bounded sample of the same workload.
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

WORKLOADS = {
    # name: (B per GPU, L, H, W, storage dtype, description)
    "c1": (8, 7, 256, 256, "float32", "C1: B=8 L=7 256x256 RGBA layers, fp32 storage"),
    "c2": (64, 7, 256, 256, "bfloat16", "C2: B=64 L=7 256x256 RGBA layers, bf16 storage"),
    "c3": (32, 16, 512, 512, "float32", "C3 shard: B=32 (256/8) L=16 512x512 RGBA layers, fp32 storage"),
}
METRIC = "composited layer-Mpix/s fwd+bwd"
UNIT = "layer-Mpix/s"


def algorithmic_bytes(B, L, H, W, s_x, s_o, s_g):
    """SURVEY.md 8d / BASELINE.md 3: every tensor touched once in its storage dtype."""
    fwd = B * L * 4 * H * W * s_x + B * 4 * H * W * s_o + 24 * B * L
    bwd = B * 4 * H * W * s_o + B * L * 4 * H * W * s_x + B * L * 4 * H * W * s_g + 48 * B * L
    return fwd, bwd


def measured_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/roofline_traffic.json, written by tools/update_traffic.py from the .ncu-rep)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh).get(workload)
    except Exception:  # noqa: BLE001
        return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.mask, self.max_mhz, self._stop = [], 0, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:  # noqa: BLE001
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b], "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference chain) -- the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------
def cpu_reference_throughput(L, H, W, sample_B, min_seconds, max_iters, seed=0):
    import torch
    import montage_gan_b200  # noqa: F401
    from montage_gan_b200 import synth
    from oracle import torch_chain as TC
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = synth.make_layers(sample_B, L, H, W, "S", seed=seed)
    th = synth.make_theta(sample_B, L, "I", seed=seed)
    go = synth.make_grad_out(sample_B, H, W, seed=seed)
    TC.fwd_bwd(TC.port_chain, x, th, go, "m11", torch.float32)          # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_iters and (len(times) < 3 or time.perf_counter() - t_start < min_seconds):
        t0 = time.perf_counter()
        TC.fwd_bwd(TC.port_chain, x, th, go, "m11", torch.float32)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": sample_B * L * H * W / 1e6 / med, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"B={sample_B} of the workload (L={L}, {H}x{W}, fp32 as the reference computes), "
                      f"{len(times)} iterations, median {med * 1e3:.1f} ms; torch {torch.__version__} CPU ATen "
                      f"affine_grid+grid_sample + per-sample/per-layer Python over loops + autograd backward"}, times


def run_reference_arm(args, wl):
    B, L, H, W, dtype_name, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = B                              # the whole batch of the workload: same config as the CUDA arm
    t0 = time.perf_counter()
    base, times = cpu_reference_throughput(L, H, W, sample_B, min_seconds=0.0, max_iters=args.steps + args.warmup)
    times = times[-args.steps:] if len(times) > args.steps else times
    ms = 1e3 * sum(times) / len(times)
    value = sample_B * L * H * W / 1e6 / (ms / 1e3)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "storage_dtype": "float32", "data": "synthetic",
            "config": {"workload": desc + f"; the CPU arm runs the whole batch (B={sample_B}) per step, fp32 storage and arithmetic "
                                          "as the reference computes",
                       "theta": "I + 0.25*N(0,1), covering back layer", "layers": "smooth (S) family"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def run_cuda_arm(args, wl):
    import ctypes
    import torch
    import torch.distributed as dist
    import montage_gan_b200  # noqa: F401
    from montage_gan_b200 import _lib, sharding, synth

    B, L, H, W, dtype_name, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    dtype = getattr(torch, dtype_name)
    dt_code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    es = torch.empty((), dtype=dtype).element_size()

    # ---- synthetic inputs: NSETS rotating sets so no step finds its inputs in L2 -----------------
    NSETS = 3
    gen_B = min(B, 16)                      # generate a seeded block on the CPU and tile it
    xs, ths, gos = [], [], []
    for k in range(NSETS):
        seed = 1000 * rank + k
        xb = synth.make_layers(gen_B, L, H, W, "S", seed=seed)
        reps = (B + gen_B - 1) // gen_B
        xs.append(xb.repeat(reps, 1, 1, 1, 1)[:B].to(dev, dtype).contiguous())   # layers repeat every gen_B samples; thetas do not
        ths.append(synth.make_theta(B, L, "I", seed=seed).to(dev))
        gos.append(synth.make_grad_out(B, H, W, "randn", seed=seed).to(dev, dtype))
    out = torch.empty((B, 4, H, W), dtype=dtype, device=dev)
    gx = torch.empty((B, L, 4, H, W), dtype=dtype, device=dev)
    gt = torch.empty((B, L, 2, 3), dtype=torch.float32, device=dev)
    flags = 3
    ws_bytes = lib.mgr_render_backward_workspace_bytes(B, L, H, W, dt_code, 1, flags)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    sav = torch.empty(max(lib.mgr_saved_alpha_bytes(B, L, H, W, dt_code), 1), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    def fwd(k):
        _lib.check(lib.mgr_render_forward(P(xs[k]), None, P(ths[k]), P(out), P(sav), B, L, H, W, dt_code, 0, sp), "forward")

    def bwd(k):
        _lib.check(lib.mgr_render_backward(P(xs[k]), None, P(ths[k]), P(out), P(gos[k]), P(sav), P(gx), P(gt), P(ws),
                                           ws_bytes, B, L, H, W, dt_code, 0, flags, sp), "backward")

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing -----------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        fwd(i % NSETS); bwd(i % NSETS)
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = lib.mgr_kernel_launch_count()
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            k = i % NSETS
            ev[i][0].record(stream); fwd(k)
            ev[i][1].record(stream); bwd(k)
            ev[i][2].record(stream)
        end_ev = torch.cuda.Event(enable_timing=True)
        end_ev.record(stream)
        torch.cuda.synchronize(dev)
        t_wall = time.perf_counter() - t_wall0
    launches = lib.mgr_kernel_launch_count() - launches0
    total_ms = ev[0][0].elapsed_time(end_ev)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    barrier()
    units_per_step = B * L * H * W            # per rank
    _, max_ms, thr = sharding.aggregate_throughput(units_per_step * args.steps, total_ms, dev)
    value = thr / 1e6

    # ---- the same step on translation-only placements (what STNv2c emits), kernels only ---------------
    ths_T = [synth.make_theta(B, L, "T", seed=1000 * rank + k, cover_back=False).to(dev) for k in range(NSETS)]
    ths_I, ths[:] = list(ths), ths_T
    for i in range(3):
        fwd(i % NSETS); bwd(i % NSETS)
    torch.cuda.synchronize(dev)
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record(stream)
    for i in range(args.steps):
        fwd(i % NSETS); bwd(i % NSETS)
    t1e.record(stream)
    torch.cuda.synchronize(dev)
    _, ms_T, thr_T = sharding.aggregate_throughput(units_per_step * args.steps, t0e.elapsed_time(t1e), dev)
    ths[:] = ths_I

    # ---- the same workload with fp32 STORAGE (what the CPU arm computes on), kernels only -------------------------
    thr_32, ms_32 = None, None
    if dtype != torch.float32 and not args.kernels_only:
        x32 = xs[0].float()
        go32 = gos[0].float()
        out32 = torch.empty((B, 4, H, W), dtype=torch.float32, device=dev)
        gx32 = torch.empty((B, L, 4, H, W), dtype=torch.float32, device=dev)
        ws32_bytes = lib.mgr_render_backward_workspace_bytes(B, L, H, W, 0, 1, flags)
        ws32 = torch.empty(max(ws32_bytes, 1), dtype=torch.uint8, device=dev)
        sav32 = torch.empty(max(lib.mgr_saved_alpha_bytes(B, L, H, W, 0), 1), dtype=torch.uint8, device=dev)

        def step32(k):
            _lib.check(lib.mgr_render_forward(P(x32), None, P(ths[k]), P(out32), P(sav32), B, L, H, W, 0, 0, sp), "forward fp32")
            _lib.check(lib.mgr_render_backward(P(x32), None, P(ths[k]), P(out32), P(go32), P(sav32), P(gx32), P(gt), P(ws32),
                                               ws32_bytes, B, L, H, W, 0, 0, flags, sp), "backward fp32")
        n32 = max(3, min(args.steps, 10))
        for i in range(3):
            step32(i % NSETS)
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for i in range(n32):
            step32(i % NSETS)
        a1.record(stream)
        torch.cuda.synchronize(dev)
        _, ms_32, thr_32 = sharding.aggregate_throughput(units_per_step * n32, a0.elapsed_time(a1), dev)
        ms_32 /= n32
        del x32, go32, out32, gx32, ws32, sav32

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + kernels + D2H in the timed region) ----
    e2e_steps = 0 if args.kernels_only else max(3, min(args.steps, 10))
    e2e_value, h2d, d2h = None, 0, 0
    if e2e_steps:
        from montage_gan_b200.host import HostRenderer
        hr = HostRenderer(B, L, H, W, dtype, chunk_B=args.chunk, device=dev)
        hx = [x.cpu().pin_memory() for x in xs[:2]]
        hth = [t_.cpu().pin_memory() for t_ in ths[:2]]
        hgo = [g_.cpu().pin_memory() for g_ in gos[:2]]
        h2d = hx[0].numel() * es + hth[0].numel() * 4 + hgo[0].numel() * es
        d2h = hr.out.numel() * es + hr.grad_x.numel() * es + hr.grad_theta.numel() * 4
        for i in range(2):
            hr.fwd_bwd(hx[i % 2], hth[i % 2], hgo[i % 2])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(e2e_steps):
            hr.fwd_bwd(hx[i % 2], hth[i % 2], hgo[i % 2])       # synchronises: the step's results are on the host
        e1.record(stream)
        torch.cuda.synchronize(dev)
        _, _, thr_e = sharding.aggregate_throughput(units_per_step * e2e_steps, e0.elapsed_time(e1), dev)
        e2e_value = thr_e / 1e6

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        fb, bb = algorithmic_bytes(B, L, H, W, es, es, es)
        # dominant kernel = the backward pass (scatter + theta reduction)
        achieved = bb / 1e9 / (bwd_ms / 1e3)
        tr = measured_traffic(args.workload) or {}
        roofline = {"bound": "hbm", "kernel": "render backward (mgr_render_backward: placements + pass 1 + pass 2)", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": tr.get("bytes"),
                    "traffic_source": tr.get("source"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bb,
                    "forward": {"achieved": fb / 1e9 / (fwd_ms / 1e3), "frac": fb / 1e9 / (fwd_ms / 1e3) / peak,
                                "ms": fwd_ms, "algorithmic_bytes_per_launch": fb, "traffic": tr.get("forward")},
                    "backward_ms": bwd_ms,
                    "fwd_bwd": {"achieved": (fb + bb) / 1e9 / ((fwd_ms + bwd_ms) / 1e3),
                                "frac": (fb + bb) / 1e9 / ((fwd_ms + bwd_ms) / 1e3) / peak,
                                "frac_of_nominal_8TBs": (fb + bb) / 1e9 / ((fwd_ms + bwd_ms) / 1e3) / 8000.0}}
        cpu_base = None
        if not args.kernels_only:
            cpu_base, _ = cpu_reference_throughput(L, H, W, B if B * L * H * W <= 64 * 7 * 256 * 256 else min(B, 8),
                                                   min_seconds=args.cpu_seconds, max_iters=50)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "storage_dtype": dtype_name, "data": "synthetic",
                "config": {"workload": desc + " per GPU, fp32 arithmetic in registers",
                           "dtype_note": f"`dtype` is the arithmetic type (every sample, lerp and composite is fp32); tensors are stored as {dtype_name}",
                           "theta": "I + 0.25*N(0,1), covering back layer", "layers": "smooth (S) family",
                           "l2": f"{NSETS} rotating input sets of {xs[0].numel() * es / 1e6:.0f} MB each (> 126 MB L2)",
                           "sharding": f"batch-sharded, {B} samples per GPU, no data-path collective"},
                "roofline": roofline, "cpu_baseline": cpu_base,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "api": "mgr_render_fwd_bwd_host (C ABI, pinned host buffers, "
                                                   f"chunks of {args.chunk} samples on 3 streams)",
                        "note": "bound by the host link, not the GPU: every step moves h2d + d2h bytes over PCIe (about 55 GB/s "
                                "one way, 40-49 GB/s each way when both directions run, tools/pcie_peak.py); at N > 1 the ranks "
                                "share the host's PCIe root complexes and DRAM, so e2e scales far below the kernels (SCALE_r01: "
                                "1.8x at 8 GPUs against 7.7x device-resident)"},
                "translation_only": {"value": thr_T / 1e6, "unit": UNIT, "ms_per_step": ms_T / args.steps,
                                     "frac": (fb + bb) / 1e9 / (ms_T / args.steps / 1e3) / peak,
                                     "traffic": tr.get("translation_fwd_bwd"),
                                     "note": "same workload with the placements STNv2c emits (pure translations, "
                                             "U(-1,1)); kernels only, inputs resident.  `frac` counts the ALGORITHMIC bytes "
                                             "(every texel once); shifts this large push half of every layer off the canvas and "
                                             "the kernels skip what no pixel samples, so the bytes actually moved (`traffic`, ncu) "
                                             "are fewer"},
                "fp32_storage": None if thr_32 is None else {
                    "value": thr_32 / 1e6, "unit": UNIT, "ms_per_step": ms_32,
                    "frac": sum(algorithmic_bytes(B, L, H, W, 4, 4, 4)) / 1e9 / (ms_32 / 1e3) / peak,
                    "note": "the same batch and placements with fp32 tensors (the storage the CPU arm computes on; one input set "
                            f"of {B * L * 4 * H * W * 4 / 1e6:.0f} MB > L2); kernels only, inputs resident"},
                "gpu_launches": int(launches), "clocks": clocks.summary(),
                "wall_s_timed_region": t_wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline sampling budget")
    ap.add_argument("--chunk", type=int, default=8, help="samples per chunk of the host-buffer pipeline (e2e)")
    ap.add_argument("--kernels-only", action="store_true",
                    help="skip the e2e and cpu_baseline legs (for ncu captures; not a valid bench line)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_cuda_arm(args, wl)


if __name__ == "__main__":
    main()
