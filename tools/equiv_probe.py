#!/usr/bin/env python
"""How reproducible is ONE training step?  Builds the same ResNet-50 (same seed, same batch) several times on one GPU
with the single-rank engine and compares the weight deltas of the first step pairwise: the noise floor that any
N-rank-vs-1-rank comparison has to be read against.  Also reports per-layer agreement for the worst layers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributeddeeplearning_b200 import models, ops  # noqa: E402
from distributeddeeplearning_b200.data import fixed_synthetic_batch  # noqa: E402
from distributeddeeplearning_b200.parallel import dist  # noqa: E402
from distributeddeeplearning_b200.parallel.engine import FusedSGD  # noqa: E402


def one(model_name, batch, seed, lr=0.01):
    torch.manual_seed(seed)
    m = models.get_model(model_name).cuda().train()
    opt = FusedSGD(m.named_parameters(), lr=lr, local=True)
    data, target = fixed_synthetic_batch(batch, 224, 1000, torch.device("cuda"), seed=seed + 17)
    w0 = opt.W.detach().clone()
    loss = ops.softmax_cross_entropy(m(data), target, 1000)
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    names = [n for n, _ in m.named_parameters()]
    return (opt.W.detach() - w0).double(), float(loss.detach()), opt, names


def main():
    dist.init()
    model = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    runs = [one(model, batch, 7) for _ in range(3)]
    d0 = runs[0][0]
    for i, (d, loss, _, _) in enumerate(runs):
        rel = float((d - d0).norm() / d0.norm())
        cos = float((d * d0).sum() / (d.norm() * d0.norm()))
        print(f"run {i}: loss {loss:.6f}  |delta| {float(d.norm()):.6e}  rel diff to run 0 {rel:.4e}  cosine {cos:.6f}")
    opt, names = runs[0][2], runs[0][3]
    d1 = runs[1][0]
    rows = []
    n = len(opt.params)
    for i in range(n):
        o, cnt = int(opt.plan["param_offset"][i]), opt.params[i].numel()
        a, b = d0[o:o + cnt], d1[o:o + cnt]
        rows.append((float((a - b).norm() / a.norm().clamp_min(1e-30)), names[n - 1 - i], float(a.norm())))
    rows.sort(reverse=True)
    print("worst layers (rel diff between two identical runs, name, |delta|):")
    for r in rows[:12]:
        print(f"  {r[0]:.3e}  {r[1]:40s} {r[2]:.3e}")
    print("best layers:")
    for r in rows[-5:]:
        print(f"  {r[0]:.3e}  {r[1]:40s} {r[2]:.3e}")


if __name__ == "__main__":
    main()
