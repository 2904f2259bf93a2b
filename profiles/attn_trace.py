"""clock64 phase timeline of the fused QaN + cross-attention kernel (CTA (0,0)) inside a graph-replayed
sampling loop of the bench workload.  usage: python profiles/attn_trace.py"""
import ctypes as C
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
names = ["start", "const staged (issued)", "pdl wait done", "rows landed + barrier", "pre-LN", "QaN dots", "QaN softmax/3-tap/LN1",
         "xattn dots+reduce", "xattn softmax", "xattn values", "-", "LN2 + stores"]
for on in (0, 1):
    eng.set_dependent_launch(on)
    for _ in range(3):
        eng.p_sample_loop(tape, gt, mask)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    assert eng.lib.idb_debug_attn_trace(buf) == 0
    t = list(buf)
    print("dependent launch %d: total %d cycles" % (on, t[11] - t[0]))
    prev = t[0]
    for i in range(1, 12):
        if i == 10:
            continue
        print("  %-28s %6d" % (names[i], t[i] - prev))
        prev = t[i]
