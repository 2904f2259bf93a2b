"""Fused decoder-layer kernel: device capacity in clusters and the step time as a function of the batch (wave quantisation)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S
from interdiff_b200.diffusion.gaussian_diffusion import get_named_beta_schedule
from interdiff_b200.engine import Engine
from interdiff_b200.weights import bench_weights
eng = Engine("cuda:0")
print("max active 8-CTA clusters of the fused layer kernel:", eng.lib.idb_debug_max_layer_clusters(eng._h))
eng.load_denoiser(bench_weights("diffusion_smpl"), "smpl")
steps = 50
eng.init_diffusion(get_named_beta_schedule("cosine", steps))
for B in (56, 60, 64, 68, 32, 8):
    b = S.make_smpl_batch(B=B, T=30)
    eng.bind(b["cond"], 30)
    tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps)).cuda()
    gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()
    for lvl in (2, 3):
        eng.set_fused_mlp(lvl)
        for _ in range(3):
            eng.p_sample_loop(tape, gt, mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.p_sample_loop(tape, gt, mask)
        e1.record()
        torch.cuda.synchronize()
        print("B=%3d fused level %d: %.1f us/step" % (B, lvl, e0.elapsed_time(e1) / 5 / steps * 1e3))
