#!/usr/bin/env python
"""Run every conv kernel variant on given layer shapes, one launch each, and report time + pipeline timeouts.
usage: variant_probe.py "cin,hw,cout,k,stride,pad[,batch]" ..."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributeddeeplearning_b200.ops import native as nv

def main():
    dev = torch.device("cuda")
    for spec in sys.argv[1:]:
        v = [int(t) for t in spec.split(",")]
        ci, hw, co, k, s, p = v[:6]
        B = v[6] if len(v) > 6 else 256
        x = torch.randn(B, ci, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(co, ci, k, k, device=dev) * 0.05)
        wb = w.permute(0, 2, 3, 1).reshape(co, -1).contiguous().to(torch.bfloat16)
        P = (hw + 2 * p - k) // s + 1
        dy = torch.randn(B, co, P, P, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        kb = k * k * ((ci + 63) // 64)
        for name, cands, run in (("fwd", nv.conv_variants((co + 63) // 64 * 64, B * P * P, kb, False),
                                  lambda: nv.conv_fwd(x, wb, (k, k), s, p, stats=True, cout=co)[0]),
                                 ("dgrad", nv.conv_variants((ci + 63) // 64 * 64, B * hw * hw, k * k * ((co + 63) // 64), True),
                                  lambda: nv.conv_dgrad(dy, wb, x.shape, (k, k), s, p))):
            ref = None
            for c in cands:
                nv.force_variant(c)
                try:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    out = run(); torch.cuda.synchronize()
                    a.record(); out = run(); b.record(); torch.cuda.synchronize()
                    bad = nv.conv_timeouts(raise_error=False)
                    if ref is None and bad is None:
                        ref = out.float()
                    d = float((out.float() - ref).abs().max()) if ref is not None else -1
                    print(f"{spec:24s} {name:5s} {nv.variant_name(c):16s} {a.elapsed_time(b) * 1e3:9.1f} us maxdiff {d:.3g} {'TIMEOUT: ' + bad if bad else ''}", flush=True)
                except RuntimeError as e:
                    print(f"{spec:24s} {name:5s} {nv.variant_name(c):16s} ERROR {str(e).splitlines()[0][:100]}", flush=True)
            nv.force_variant(None)

if __name__ == "__main__":
    main()
