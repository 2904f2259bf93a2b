"""One correction step (the denoised_fn hook at B=64, T=30, F=1920 frames) a few times - for ncu launch lists / captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S
from interdiff_b200.engine import Engine
from interdiff_b200.weights import bench_weights
eng = Engine("cuda:0")
eng.load_body(S.make_smplh_model(233, sparse_weights=True))
eng.load_projector(bench_weights("correction_smpl"), 10, 20)
b = S.make_smpl_batch(B=64, T=30)
eng.bind_correction(b["hand_pose"], b["betas"], b["obj_points"], past_len=10)
gt = torch.from_numpy(b["gt"]).cuda()
x = (gt + 0.02 * torch.randn(gt.shape, device="cuda")).contiguous()
for _ in range(3):
    eng.correction_apply(x.clone(), gt, 450)
torch.cuda.synchronize()
print("ok")
