#!/usr/bin/env python
"""Developer micro-benchmark: times mgr_render_forward / mgr_render_backward separately through the
C ABI with CUDA events, for several workloads, theta families and kernel paths.  Not the
contract benchmark (that is bench.py); used to iterate on kernels and to fill DESIGN.md tables.

    python tools/kbench.py [--workloads c2,c1] [--thetas I,T,X] [--paths auto,direct] [--iters 20]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import montage_gan_b200  # noqa: E402,F401
from montage_gan_b200 import _lib, synth  # noqa: E402
from bench import WORKLOADS, algorithmic_bytes, measured_peak_gbs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2")
    ap.add_argument("--thetas", default="I")
    ap.add_argument("--paths", default="auto")
    ap.add_argument("--layers", default="S")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--custom", default="", help="B,L,H,W,dtype e.g. 16,16,512,512,float32")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    peak, _ = measured_peak_gbs()
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    wls = [WORKLOADS[w] for w in args.workloads.split(",") if w]
    if args.custom:
        b, l, h, w, d = args.custom.split(",")
        wls.append((int(b), int(l), int(h), int(w), d, "custom"))
    for (B, L, H, W, dtype_name, desc) in wls:
        dtype = getattr(torch, dtype_name)
        dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
        es = 4 if dt == 0 else 2
        gen_B = min(B, 8)
        reps = (B + gen_B - 1) // gen_B
        nsets = 3
        for tf in args.thetas.split(","):
            xs = [synth.make_layers(gen_B, L, H, W, args.layers, seed=k).repeat(reps, 1, 1, 1, 1)[:B].to(dev, dtype).contiguous()
                  for k in range(nsets)]
            # 'P' = pure translations for every layer (what STNv2c emits); 'T' keeps the covering back layer
            ths = [(synth.make_theta(B, L, 'T', seed=k, cover_back=False) if tf == 'P' else synth.make_theta(B, L, tf, seed=k)).to(dev)
                   for k in range(nsets)]
            gos = [synth.make_grad_out(B, H, W, seed=k).to(dev, dtype) for k in range(nsets)]
            out = torch.empty(B, 4, H, W, dtype=dtype, device=dev)
            gx = torch.empty(B, L, 4, H, W, dtype=dtype, device=dev)
            gt = torch.empty(B, L, 2, 3, device=dev)
            wsb = lib.mgr_render_backward_workspace_bytes(B, L, H, W, dt, 1, 3)
            ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
            sav = torch.empty(max(lib.mgr_saved_alpha_bytes(B, L, H, W, dt), 1), dtype=torch.uint8, device=dev)
            sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            for path in args.paths.split(","):
                lib.mgr_set_debug_path({"auto": 0, "direct": 1, "nostencil": 2, "legacy": 3}[path])
                def fwd(k):
                    _lib.check(lib.mgr_render_forward(P(xs[k]), None, P(ths[k]), P(out), P(sav), B, L, H, W, dt, 0, sp), "fwd")
                def bwd(k):
                    _lib.check(lib.mgr_render_backward(P(xs[k]), None, P(ths[k]), P(out), P(gos[k]), P(sav), P(gx), P(gt), P(ws), wsb,
                                                       B, L, H, W, dt, 0, 3, sp), "bwd")
                res = {}
                for name, fn in (("fwd", fwd), ("bwd", bwd)):
                    for i in range(3):
                        fn(i % nsets)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(args.iters):
                        fn(i % nsets)
                    e1.record()
                    torch.cuda.synchronize()
                    res[name] = e0.elapsed_time(e1) / args.iters
                fb, bb = algorithmic_bytes(B, L, H, W, es, es, es)
                tot = res["fwd"] + res["bwd"]
                print(json.dumps({"wl": f"B{B} L{L} {H}x{W} {dtype_name}", "theta": tf, "path": path,
                                  "fwd_us": round(res["fwd"] * 1e3, 1), "bwd_us": round(res["bwd"] * 1e3, 1),
                                  "fwd_frac": round(fb / 1e6 / res["fwd"] / peak, 3), "bwd_frac": round(bb / 1e6 / res["bwd"] / peak, 3),
                                  "tot_frac": round((fb + bb) / 1e6 / tot / peak, 3),
                                  "Mpix_s": round(B * L * H * W / 1e3 / tot)}), flush=True)
            lib.mgr_set_debug_path(0)


if __name__ == "__main__":
    main()
