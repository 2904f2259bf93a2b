"""Config 3 (SMPL diffusion + correction, 100 steps, B=64, T=30) throughput and per-kernel timings of
the correction path (SMPL-H LBS, normals, signed NN, projector) with CUDA events."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S  # noqa: E402
from interdiff_b200.diffusion.gaussian_diffusion import get_named_beta_schedule  # noqa: E402
from interdiff_b200.engine import Engine  # noqa: E402
from tests.helpers import mdm_weights, projector_weights  # noqa: E402

B, T, steps = 64, 30, 100
eng = Engine("cuda:0")
eng.load_denoiser(mdm_weights("smpl", "auto"), "smpl")
smplh = S.make_smplh_model(233)
eng.load_body(smplh)
eng.load_projector(projector_weights("auto"), 10, 20)
b = S.make_smpl_batch(B=B, T=T)
eng.bind(b["cond"], T)
eng.bind_correction(b["hand_pose"], b["betas"], b["obj_points"], past_len=10)
eng.init_diffusion(get_named_beta_schedule("cosine", steps))
tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps)).cuda()
gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms2 = timed(lambda: eng.p_sample_loop(tape, gt, mask, correction=False))
ms3 = timed(lambda: eng.p_sample_loop(tape, gt, mask, correction=True))
print("config 2 (no correction): %.2f ms / 100 steps -> %.0f steps/s" % (ms2, 100e3 / ms2))
print("config 3 (correction at t=50,0): %.2f ms / 100 steps -> %.0f steps/s ; one correction step = %.2f ms" % (ms3, 100e3 / ms3, (ms3 - ms2) / 2))
F = T * B
g = torch.Generator().manual_seed(0)
pose = (0.3 * torch.randn(F, 156, generator=g)).cuda()
betas = torch.randn(F, 10, generator=g).cuda()
trans = torch.randn(F, 3, generator=g).cuda()
ms = timed(lambda: eng.lbs(pose, betas, trans, want_jtr=False))
byt = 41.7e6 + F * 84.0e3
print("LBS F=%d: %.3f ms ; algorithmic %.1f MB -> %.0f GB/s ; %.1f GFLOP -> %.1f TFLOP/s" % (F, ms, byt / 1e6, byt / ms / 1e6, F * 33.2e-3, F * 33.2e6 / ms / 1e9))
verts, _ = eng.lbs(pose, betas, trans, want_jtr=False)
ms = timed(lambda: eng.vertex_normals(verts))
print("vertex_normals F=%d: %.3f ms (%.0f GB/s of 2 x 159 MB)" % (F, ms, 2 * F * 6890 * 12 / ms / 1e6))
normals = eng.vertex_normals(verts)
obj = (verts[:, ::4][:, :2048] * 1.1).contiguous()
for pr in (False, True):
    eng.set_nn_pruning(pr)
    ms = timed(lambda: eng.signed_nn(obj, verts, normals), n=3, warm=1)
    print("signed_nn F=%d 2048x6890 (%s): %.3f ms -> %.2f T brute-force-equivalent pair-evals/s" % (
        F, "cluster-pruned" if pr else "brute force", ms, F * 2048 * 6890 / ms / 1e9))
