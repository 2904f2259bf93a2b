"""LBS F=1920 a few times (for an ncu launch list: python profiles/lbs_only.py [sparse|dense])."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S
from interdiff_b200.engine import Engine
eng = Engine("cuda:0")
eng.load_body(S.make_smplh_model(233, sparse_weights=(len(sys.argv) < 2 or sys.argv[1] != "dense")))
F = 1920
g = torch.Generator().manual_seed(0)
pose, betas, trans = (0.3 * torch.randn(F, 156, generator=g)).cuda(), torch.randn(F, 10, generator=g).cuda(), torch.randn(F, 3, generator=g).cuda()
for _ in range(3):
    v, _ = eng.lbs(pose, betas, trans, want_jtr=False)
torch.cuda.synchronize()
print(float(v.abs().max()))
