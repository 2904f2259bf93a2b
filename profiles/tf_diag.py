"""Diagnostic: teacher-forced per-step error of configs[1] (B=64, T=30, 100 steps) vs the oracle, both GEMM backends."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200 import synthetic as S
from interdiff_b200.engine import Engine
from interdiff_b200.weights import bench_weights
from oracle import restate as R

B, T, steps = 64, 30, 100
torch.set_num_threads(min(os.cpu_count(), 32))
sd = bench_weights("diffusion_smpl")
b = S.make_smpl_batch(B=B, T=T)
gt, mask, cond = torch.from_numpy(b["gt"]), torch.from_numpy(b["mask"]), torch.from_numpy(b["cond"])
tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps))
tables = R.diffusion_tables(R.named_beta_schedule("cosine", steps))
with torch.no_grad():
    _, traj = R.p_sample_loop(lambda x, t: R.mdm_smpl_forward(sd, x, t, cond, faithful=False), tables, tape, gt, mask, return_trajectory=True)
eng = Engine("cuda:0")
eng.load_denoiser(sd, "smpl")
eng.bind(b["cond"], T)
eng.init_diffusion(R.named_beta_schedule("cosine", steps))
gtd, maskd, taped = gt.cuda(), mask.cuda(), tape.cuda()
res = {}
for backend in ("tcgen05", "simt"):
    eng.set_gemm_backend(backend)
    outs = []
    x_ref = tape[0]
    for k in range(steps):
        got, got0 = eng.p_sample(steps - 1 - k, x_ref.cuda(), taped[k + 1], gtd, maskd)
        outs.append((got.cpu(), got0.cpu()))
        x_ref = traj[k][0]
    res[backend] = outs
def rel(a, b): return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12)).item()
print("step  t    tc:x      tc:x0     simt:x    simt:x0   tc-vs-simt:x0   max|x0|  worst sample (tc x0) [b, c, t]")
for k in range(steps):
    tx, tx0 = rel(res["tcgen05"][k][0], traj[k][0]), rel(res["tcgen05"][k][1], traj[k][1])
    sx, sx0 = rel(res["simt"][k][0], traj[k][0]), rel(res["simt"][k][1], traj[k][1])
    d = (res["tcgen05"][k][1] - traj[k][1]).abs()
    idx = np.unravel_index(int(d.argmax()), d.shape)
    if k < 12 or k % 10 == 0 or max(tx0, sx0) > 1e-4 or k > 90:
        print("%3d %3d  %.2e  %.2e  %.2e  %.2e  %.2e   %8.2f  %s" % (k, steps - 1 - k, tx, tx0, sx, sx0, rel(res["tcgen05"][k][1], res["simt"][k][1]),
                                                                    traj[k][1].abs().max().item(), (idx[0], idx[2], idx[3])))
# per-sample error at the worst step
k = max(range(steps), key=lambda k: rel(res["tcgen05"][k][1], traj[k][1]))
d = (res["tcgen05"][k][1] - traj[k][1]).abs().flatten(1).max(1)[0] / traj[k][1].abs().max()
print("worst step", k, "per-sample rel err (sorted desc, top 8):", sorted([(round(v.item(), 7), i) for i, v in enumerate(d)], reverse=True)[:8])
# same step in a batch of 4 containing the worst sample: does the error follow the sample or the batch size?
w = int(d.argmax())
pick = [w, (w + 1) % B, (w + 2) % B, (w + 3) % B]
eng.set_gemm_backend("tcgen05")
eng.bind(np.ascontiguousarray(b["cond"][:, pick]), T)
x_in = (tape[0] if k == 0 else traj[k - 1][0])[pick].contiguous()
g4, g40 = eng.p_sample(steps - 1 - k, x_in.cuda(), taped[k + 1][pick].contiguous(), gtd[pick].contiguous(), maskd[pick].contiguous())
print("batch of 4 with the worst sample: rel err x0 (sample-wise):", [(p, rel(g40[j].cpu(), traj[k][1][p])) for j, p in enumerate(pick)])
# oracle faithful=True on that sample
with torch.no_grad():
    o2 = R.mdm_smpl_forward(sd, x_in, torch.full((4,), steps - 1 - k), cond[:, pick], faithful=True)
    o1 = R.mdm_smpl_forward(sd, x_in, torch.full((4,), steps - 1 - k), cond[:, pick], faithful=False)
    o64 = R.mdm_smpl_forward({kk: (v.double() if v.is_floating_point() else v) for kk, v in sd.items()}, x_in.double(), torch.full((4,), steps - 1 - k), cond[:, pick].double(), faithful=False)
raw = eng.forward(x_in.cuda(), torch.full((4,), steps - 1 - k).cuda()).cpu()
print("forward only (no inpaint): gpu vs oracle(alg) %.2e, gpu vs oracle(faithful) %.2e, oracle alg vs faithful %.2e" % (rel(raw, o1), rel(raw, o2), rel(o1, o2)))
print("vs float64 oracle: gpu %.2e, oracle f32(alg) %.2e, oracle f32(faithful) %.2e" % (rel(raw, o64), rel(o1, o64), rel(o2, o64)))
