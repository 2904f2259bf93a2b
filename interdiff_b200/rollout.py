"""Long-horizon autoregressive sampling (reference eval_smpl_long.py:247-285, BASELINE configs[4]) with every step between
two windows on the device: no host round trip, no numpy / scipy per frame.

    window 0 : the caller's inpainting tensor (10 observed frames)
    window k : get_batch (eval_smpl_long.py:26-84) of window k-1's last past_len predicted frames -> conditioning encoder ->
               sampling loop (+ correction hook) -> post-processing (6D -> axis-angle, SMPL-H LBS) -> denormalize (:278)

Upstream calls `denormalize` and `correct` without defining them and its get_batch raises for every batch size
(csrc/rollout.cu header); here denormalize is get_batch's exact inverse (add the accumulated per-sample centroid to
translations, vertices and joints) and correct is the identity.
"""
import torch

from .sampling import sample_postprocess


class RolloutDriver:
    def __init__(self, engine, past_len=10, n_windows=10, correction=True, pc_embedding=None):
        """engine: denoiser loaded (with the conditioning-encoder tensors when conditions are to be re-encoded per window),
        diffusion initialised, body model / projector loaded for the hook.  n_windows: windows AFTER the first one.
        pc_embedding (B,256): the object's point-cloud embedding (constant over the rollout); without it the first window's
        bound condition is kept for every window (benchmarks with synthetic conditions)."""
        self.eng, self.past, self.n_windows, self.correction, self.pc = engine, int(past_len), int(n_windows), correction, pc_embedding

    def run(self, tape, gt, mask, hand_pose, betas, out=None, obj_points=None, keep=("body", "obj", "pelvis")):
        """tape (n+1,B,1,144,T) - reused for every window unless a list of tapes is given; gt / mask (B,1,144,T) of window 0.
        Returns a dict of world-coordinate trajectories over past_len + (1 + n_windows) * future_len frames:
        body (.,B,159), obj (.,B,6), pelvis (.,B,3) (+ verts / jtr when listed in `keep`); `out` receives the last window's
        raw sample."""
        eng, P = self.eng, self.past
        T = gt.shape[-1]
        B = gt.shape[0]
        tapes = tape if isinstance(tape, (list, tuple)) else [tape] * (self.n_windows + 1)
        offset = torch.zeros(B, 3, device=eng.device)
        traj = {k: [] for k in keep}
        cur_gt = gt
        for k in range(self.n_windows + 1):
            if self.pc is not None:
                eng.bind(eng.encode_condition(cur_gt[..., :P].contiguous(), self.pc), T)
            if self.correction and k == 0 and obj_points is not None:
                eng.bind_correction(hand_pose, betas, obj_points, past_len=P)
            sample = eng.p_sample_loop(tapes[k], cur_gt, mask, correction=self.correction, use_graph=True, out=out)
            body, obj, verts, jtr = sample_postprocess(eng, sample, hand_pose, betas)
            nxt = None
            if k < self.n_windows:
                nxt, cen = eng.rollout_next_window(body, obj, jtr, T, P)          # canonical coordinates of THIS window
            # denormalize: back to world coordinates with the centroids accumulated so far
            eng.add_offset_(body, offset, col0=body.shape[2] - 3, K=1)
            eng.add_offset_(obj, offset, col0=3, K=1)
            eng.add_offset_(jtr, offset)
            if "verts" in keep:
                eng.add_offset_(verts, offset)
            pelvis = jtr[:, :, 0].contiguous()
            first = 0 if k == 0 else P                                               # later windows contribute their future frames
            vals = dict(body=body, obj=obj, pelvis=pelvis, verts=verts, jtr=jtr)
            for name in keep:
                traj[name].append(vals[name][first:].clone())
            if nxt is not None:
                offset = offset + cen
                cur_gt = nxt
        return {name: torch.cat(v, dim=0) for name, v in traj.items()}
