"""Small driver for ncu captures: a few un-graphed p_sample steps of the bench workload (B=64,T=30)
and GEMM microbenchmarks.  usage: python profiles/step_probe.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S  # noqa: E402
from interdiff_b200.engine import Engine  # noqa: E402
from interdiff_b200.diffusion.gaussian_diffusion import get_named_beta_schedule  # noqa: E402
from tests.helpers import mdm_weights  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = Engine("cuda:0")
eng.load_denoiser(mdm_weights("smpl", "auto"), "smpl")
b = S.make_smpl_batch(B=64, T=30)
eng.bind(b["cond"], 30)
eng.init_diffusion(get_named_beta_schedule("cosine", 100))
tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps)).cuda()
gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()
x = tape[0]
for k in range(steps):
    x, _ = eng.p_sample(99 - k, x, tape[k + 1], gt, mask)
torch.cuda.synchronize()
for (M, N, K, sk) in ((1920, 1024, 256, False), (1920, 256, 1024, False), (1920, 256, 1024, True), (1920, 1536, 256, False),
                      (1920, 256, 256, False)):
    r = eng.gemm_microbench(M, N, K, iters=200, gelu=(N == 1024), split_k=sk)
    print("gemm %dx%dx%d%s: %.2f us/launch, %.1f algorithmic TFLOP/s" % (M, N, K, " split-K 2" if sk else "", r["ms"] * 1e3,
                                                                        2.0 * M * N * K / (r["ms"] * 1e-3) / 1e12))
