"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck):
3-step sampling loop (graph off), fused feed-forward, GEMM tile configurations, LBS, normals, pruned NN, correction."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S  # noqa: E402
from interdiff_b200.diffusion.gaussian_diffusion import get_named_beta_schedule  # noqa: E402
from interdiff_b200.engine import Engine  # noqa: E402
from tests.helpers import encoder_weights, metrics_inputs, projector_weights  # noqa: E402

B, T, steps = 3, 30, 3
eng = Engine("cuda:0")
eng.load_denoiser(encoder_weights("random"), "smpl")
smplh = S.make_smplh_model(233)
eng.load_body(smplh)
eng.load_projector(projector_weights("random"), 10, 20)
b = S.make_smpl_batch(B=B, T=T)
eng.bind(b["cond"], T)
eng.bind_correction(b["hand_pose"], b["betas"], b["obj_points"], past_len=10)
eng.init_diffusion(get_named_beta_schedule("cosine", steps))
tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps)).cuda()
gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()
out = eng.p_sample_loop(tape, gt, mask, correction=True, use_graph=False)
out2 = eng.p_sample_loop(tape, gt, mask, correction=False, use_graph=True)
g = torch.Generator().manual_seed(0)
for (M, N, K) in ((200, 1536, 256), (130, 256, 1024), (70, 72, 36)):
    eng.gemm(torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), bias=torch.randn(N, generator=g), gelu=True)
eng.gemm(torch.randn(130, 1024, generator=g), torch.randn(256, 1024, generator=g), split_k=True)
eng.mlp(torch.randn(129, 256, generator=g), torch.randn(1024, 256, generator=g) / 16, torch.randn(1024, generator=g),
        torch.randn(256, 1024, generator=g) / 32, torch.randn(256, generator=g), torch.randn(129, 256, generator=g))
# conditioning path (PointNet++ + encoder), metrics, both fused feed-forward levels
pts = 0.2 * (torch.rand(B, 2048, 3, generator=g) - 0.5)
pc = eng.pointcloud_embed(pts.cuda())
cond = eng.encode_condition(gt[..., :10].contiguous(), pc)
for level in (1, 2, 0):
    eng.set_fused_mlp(level)
    eng.bind(cond, T)
    eng.forward(tape[0], torch.tensor([5, 500, 999]).cuda())
eng.set_fused_mlp(2)
mi = metrics_inputs(smplh, T=3, B=2, P=128)
eng.metrics(**{k: v.cuda() for k, v in mi.items() if k != "faces"})
lw, lb = torch.ones(256), torch.zeros(256)
eng.mlp(torch.randn(129, 256, generator=g), torch.randn(1024, 256, generator=g) / 16, torch.randn(1024, generator=g),
        torch.randn(256, 1024, generator=g) / 32, torch.randn(256, generator=g), torch.randn(129, 256, generator=g), ln_w=lw, ln_b=lb)
torch.cuda.synchronize()
print("sanitizer probe done", float(out.abs().max()), float(out2.abs().max()))
