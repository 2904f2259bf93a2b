"""clock64 timeline of the fused feed-forward cluster kernel (M=1920: 120 CTAs) + mean launch time.
usage: python profiles/mlp_trace.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200.engine import Engine  # noqa: E402

eng = Engine("cuda:0")
g = torch.Generator().manual_seed(1)
M = 1920
x = torch.randn(M, 256, generator=g); w1 = torch.randn(1024, 256, generator=g) / 16; b1 = torch.randn(1024, generator=g) * 0.1
w2 = torch.randn(256, 1024, generator=g) / 32; b2 = torch.randn(256, generator=g) * 0.1; res = torch.randn(M, 256, generator=g)
names = ["entry", "setup done", "mma: first operands landed", "mma: GEMM1 issued", "epi: acc1 ready", "epi: S written", "mma: S k-block 0 ready",
         "mma: GEMM2 issued", "epi: acc2 ready", "epi: partial in smem", "cluster sync 1", "reduction stored", "cluster sync 2"]
for v in [0]:
    tr = torch.zeros(256, 16, dtype=torch.int64, device="cuda")
    eng.mlp(x, w1, b1, w2, b2, res, iters=3)
    eng.mlp(x, w1, b1, w2, b2, res, iters=201, trace=tr)
    print("fused feed-forward, M=%d: %.2f us/launch" % (M, eng.last_ms() * 1e3))
    t = tr.cpu()[:120]
    for i, n in enumerate(names):
        d = (t[:, i] - t[:, 0]).float()
        print("   %-32s median %7.0f   max %7.0f" % (n, d.median().item(), d.max().item()))
