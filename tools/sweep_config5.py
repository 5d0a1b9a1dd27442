#!/usr/bin/env python
"""BASELINE config 5: sweep layers 4-32, resolution 128-1024, theta families (T pure translation, I random affine,
X extreme scale/rotation) -- fwd+bwd layer-Mpix/s and fraction of the measured HBM peak, fp32 and bf16.
Writes one JSON line per point (tee it into profiles/)."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import montage_gan_b200  # noqa
from montage_gan_b200 import _lib, synth
from bench import algorithmic_bytes, measured_peak_gbs

lib = _lib.load()
dev = torch.device("cuda", 0)
peak, _ = measured_peak_gbs()
P = lambda t: ctypes.c_void_p(t.data_ptr())
budget_px = 64 * 7 * 256 * 256          # layer-pixels per point (config-2 sized), batch adapts
for dtype_name in ("bfloat16", "float32"):
    dtype = getattr(torch, dtype_name); dt = 1 if dtype_name == "bfloat16" else 0; es = 2 if dt else 4
    for res in (128, 256, 512, 1024):
        for L in (4, 8, 16, 32):
            B = max(1, budget_px // (L * res * res))
            for tf in ("P", "I", "X"):
                gB = min(B, 4)
                x = synth.make_layers(gB, L, res, res, "S", seed=0).repeat((B + gB - 1) // gB, 1, 1, 1, 1)[:B].to(dev, dtype).contiguous()
                th = (synth.make_theta(B, L, "T", seed=0, cover_back=False) if tf == "P" else synth.make_theta(B, L, tf, seed=0)).to(dev)
                go = synth.make_grad_out(B, res, res, seed=0).to(dev, dtype)
                out = torch.empty(B, 4, res, res, dtype=dtype, device=dev); gx = torch.empty_like(x); gt = torch.empty(B, L, 2, 3, device=dev)
                sav = torch.empty(max(1, lib.mgr_saved_alpha_bytes(B, L, res, res, dt)), dtype=torch.uint8, device=dev)
                wsb = lib.mgr_render_backward_workspace_bytes(B, L, res, res, dt, 1, 3)
                ws = torch.empty(max(1, wsb), dtype=torch.uint8, device=dev)
                sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                def step():
                    _lib.check(lib.mgr_render_forward(P(x), None, P(th), P(out), P(sav), B, L, res, res, dt, 0, sp), "fwd")
                    _lib.check(lib.mgr_render_backward(P(x), None, P(th), P(out), P(go), P(sav), P(gx), P(gt), P(ws), wsb, B, L, res, res, dt, 0, 3, sp), "bwd")
                for _ in range(2): step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5): step()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                fb, bb = algorithmic_bytes(B, L, res, res, es, es, es)
                print(json.dumps({"dtype": dtype_name, "res": res, "L": L, "B": B, "theta": tf, "ms": round(ms, 3),
                                  "Mpix_s": round(B * L * res * res / 1e3 / ms), "frac": round((fb + bb) / 1e6 / ms / peak, 3)}), flush=True)
                del x, gx, ws, sav
