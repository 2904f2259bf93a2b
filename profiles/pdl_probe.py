"""A/B of programmatic dependent launch on the bench workload (B=64, T=30, 100-step loop, CUDA graph
per step): loop time with the kernels serialised vs overlapped, and a bit-exactness check between the two.
usage: python profiles/pdl_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S  # noqa: E402
from interdiff_b200.engine import Engine  # noqa: E402
from interdiff_b200.diffusion.gaussian_diffusion import get_named_beta_schedule  # noqa: E402
from tests.helpers import mdm_weights  # noqa: E402

eng = Engine("cuda:0")
eng.load_denoiser(mdm_weights("smpl", "auto"), "smpl")
b = S.make_smpl_batch(B=64, T=30)
eng.bind(b["cond"], 30)
eng.init_diffusion(get_named_beta_schedule("cosine", 100))
tape = torch.from_numpy(S.noise_tape(b["gt"].shape, 100)).cuda()
gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()
outs = {}
for on in (0, 1, 0, 1):
    eng.set_dependent_launch(on)
    for _ in range(2):
        out = eng.p_sample_loop(tape, gt, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = eng.p_sample_loop(tape, gt, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    outs[on] = out.clone()
    print("dependent launch %d: %.3f ms / 100 steps -> %.1f steps/s" % (on, ms, 100e3 / ms))
print("max |diff| between the two modes:", (outs[0] - outs[1]).abs().max().item())
