#!/usr/bin/env python
"""Per-direction bandwidth while both directions are busy, with the H2D side on 1, 2 or 4 streams (developer tool)."""
import json
import torch

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(2 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def run(nh, nd=1, piece=32 << 20):
    sh = [torch.cuda.Stream() for _ in range(nh)]
    sd = [torch.cuda.Stream() for _ in range(nd)]
    torch.cuda.synchronize()
    t0 = E(); t0.record()
    for s in sh + sd:
        s.wait_stream(torch.cuda.current_stream())
    # D2H moves twice as much so that it stays busy for the whole H2D transfer
    for k, off in enumerate(range(0, 2 * n, piece)):
        with torch.cuda.stream(sd[k % nd]):
            h_out[off:off + piece].copy_(d_out[off:off + piece], non_blocking=True)
    sub = piece // nh
    for off in range(0, n, piece):
        for q in range(nh):
            with torch.cuda.stream(sh[q]):
                o = off + q * sub
                d_in[o:o + sub].copy_(h_in[o:o + sub], non_blocking=True)
    eh = []
    for s in sh:
        e = E(); e.record(s); eh.append(e)
    ed = []
    for s in sd:
        e = E(); e.record(s); ed.append(e)
    torch.cuda.synchronize()
    th = max(t0.elapsed_time(e) for e in eh)
    td = max(t0.elapsed_time(e) for e in ed)
    return {"h2d_streams": nh, "d2h_streams": nd, "h2d_GBs_while_d2h_busy": round(n / 1e6 / th, 1), "d2h_GBs_overall": round(2 * n / 1e6 / td, 1)}


run(1)
for nh, nd in ((1, 1), (2, 1), (4, 1), (2, 2), (1, 2)):
    print(json.dumps(run(nh, nd)))
print(json.dumps({"asyncEngineCount": torch.cuda.get_device_properties(0).multi_processor_count and __import__("ctypes").c_int(0).value}))
