#!/usr/bin/env python
"""Small single-GPU target for Nsight Compute captures (never a bench value).

    ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 4 -c 3 \
        -o gpurun_out/prof_conv python tools/ncu_target.py conv
Targets: conv (3x3 64->64 @56x56 and 1x1 256->64, batch 256: fwd/dgrad/wgrad), bn (BN fwd / bwd), sgd (fused update).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributeddeeplearning_b200.ops import native as nv  # noqa: E402

cl = torch.channels_last
dev = torch.device("cuda")
what = sys.argv[1] if len(sys.argv) > 1 else "conv"
B = int(os.environ.get("NCU_BATCH", 256))
if what == "conv":
    for (ci, hw, co, k, s, p) in [(64, 56, 256, 1, 1, 0), (64, 56, 64, 3, 1, 1), (256, 14, 256, 3, 1, 1), (256, 14, 1024, 1, 1, 0)]:
        x = torch.randn(B, ci, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
        w = torch.randn(co, ci, k, k, device=dev) * 0.05
        wb = w.permute(0, 2, 3, 1).reshape(co, -1).contiguous().to(torch.bfloat16)
        gw = torch.zeros(co, ci, k, k, device=dev).contiguous(memory_format=cl)
        for _ in range(1):
            y, st = nv.conv_fwd(x, wb, (k, k), s, p, stats=True, cout=co)
            dy = torch.randn_like(y)
            nv.conv_dgrad(dy, wb, x.shape, (k, k), s, p)
            nv.conv_wgrad(x, dy, gw, (k, k), s, p)
elif what == "convlong":
    # long-K layers through each kernel variant (forward + dgrad), one launch each: where does the time go?
    for (ci, hw, co, k, s, p) in [(256, 14, 256, 3, 1, 1), (1024, 14, 256, 1, 1, 0)]:
        x = torch.randn(B, ci, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
        w = torch.randn(co, ci, k, k, device=dev) * 0.05
        wb = w.permute(0, 2, 3, 1).reshape(co, -1).contiguous().to(torch.bfloat16)
        P = (hw + 2 * p - k) // s + 1
        dy = torch.randn(B, co, P, P, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
        gw = torch.zeros(co, ci, k, k, device=dev).contiguous(memory_format=cl)
        for v in (1, 2, 3 | (3 << 4), 3 | (3 << 4) | (1 << 8), 3 | (2 << 4)):
            nv.force_variant(v)
            nv.conv_fwd(x, wb, (k, k), s, p, stats=True, cout=co)
            nv.conv_dgrad(dy, wb, x.shape, (k, k), s, p)
        nv.force_variant(None)
        nv.conv_wgrad(x, dy, gw, (k, k), s, p)
elif what == "convwide":
    # store-bound short-K forward layers through the persistent and the one-tile kernel
    for (ci, hw, co, k, s, p) in [(64, 56, 256, 1, 1, 0), (64, 56, 64, 3, 1, 1)]:
        x = torch.randn(B, ci, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
        w = torch.randn(co, ci, k, k, device=dev) * 0.05
        wb = w.permute(0, 2, 3, 1).reshape(co, -1).contiguous().to(torch.bfloat16)
        for v in (2, 1):
            nv.force_variant(v)
            nv.conv_fwd(x, wb, (k, k), s, p, stats=True, cout=co)
        nv.force_variant(None)
elif what == "bn":
    c, hw = 256, 56
    y = torch.randn(B, c, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
    g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    for _ in range(3):
        z, save = nv.bn_act_fwd(y, None, g, b, rm, rv, 1e-5, 0.1, True, None, True)
        dz = torch.randn_like(z)
        nv.bn_act_bwd(dz, z, y, save, g, True, False, torch.zeros(c, device=dev), torch.zeros(c, device=dev), beta=b,
                      had_residual=False)
elif what == "sgd":
    from distributeddeeplearning_b200.parallel import dist
    from distributeddeeplearning_b200.parallel.engine import FusedSGD

    dist.init()
    ps = [torch.nn.Parameter(torch.randn(2048, 512, 3, 3, device=dev).contiguous(memory_format=cl)) for _ in range(3)]
    opt = FusedSGD(ps, lr=0.1, momentum=0.9)
    for _ in range(3):
        for p in ps:
            p._ddl_ready()
        opt.step()
torch.cuda.synchronize()
print("done", what)
