"""GPU parity at the sizes and code paths BASELINE.json names (VERDICT r1 "What's weak" 1-3):

  * configs[1] exactly (B=64, T=30, 100-step schedule, shipped weights when present): every one of the 100 steps
    teacher-forced from the oracle's x_t at 2e-4, and the free-running per-step curve against the oracle asserted at
    1e-3 up to the step where the REFERENCE disagrees with itself by 1e-4 (float32 vs float64 / 1 vs N threads,
    profiles/r2_conditioning_probe.json);
  * configs[2]'s product path: idb_p_sample_loop(correction=1) against oracle p_sample_loop + make_denoised_fn on a
    schedule with two active hook steps (decisions exact), plus teacher-forced correction steps of the 1000-step
    schedule at t = 500 and t = 0;
  * the drop-in veneer: sample_once_proj (eval_smpl_short.py:133-215) replayed purely through the mirrored modules,
    with the fused hook object and with a plain-Python denoised_fn built from the mirrored SMPL_Layer /
    vertex_normals / point2point_signed / ObjProjector.sample;
  * the graph cache: rebind / reload / backend switch between two loops.

All product calls go through the C ABI (Engine = ctypes).  Tolerances: max|a-b| / max|b|.
"""
import json
import os
from argparse import Namespace

import numpy as np
import pytest
import torch

from interdiff_b200 import synthetic as S
from oracle import restate as R
from oracle import transforms as tf
from tests.helpers import encoder_weights, mdm_weights, projector_weights, rel, smplh_torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    from interdiff_b200.engine import Engine
    e = Engine("cuda:0")
    yield e
    e.close()


def _cutoff(default=85):
    """first step k at which the REFERENCE's own classes disagree with themselves on configs[1] by more than 1e-4
    (float32 vs float64 run, or 1 vs N threads; oracle/conditioning_probe.py -> profiles/r2_conditioning_probe.json:
    k = 85; the reference's float32 sample is 0.44 away from its float64 sample at the last step)"""
    p = os.path.join(ROOT, "profiles", "r2_conditioning_probe.json")
    if not os.path.exists(p):
        return default, None
    with open(p) as f:
        d = json.load(f)
    return int(d["first_step_self_above_1e4"]), d["reference_f32_vs_f64"]


_ORACLE_CACHE = {}


def _oracle_config2(sd, b, steps):
    key = (id(sd), steps)
    if key not in _ORACLE_CACHE:
        torch.set_num_threads(min(os.cpu_count(), 32))
        gt, mask, cond = torch.from_numpy(b["gt"]), torch.from_numpy(b["mask"]), torch.from_numpy(b["cond"])
        tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps))
        tables = R.diffusion_tables(R.named_beta_schedule("cosine", steps))
        with torch.no_grad():
            _, traj = R.p_sample_loop(lambda x, t: R.mdm_smpl_forward(sd, x, t, cond, faithful=False), tables, tape, gt, mask,
                                      return_trajectory=True)
            # the SAME steps from the same x_t in float64: the rounding-free yardstick for pred_xstart (one forward of
            # this model amplifies fp32 rounding by up to ~1e3 on the early steps of the 100-step schedule, where x_t is
            # noise but t says "almost clean": the float32 oracle itself is up to 1e-3 away from this)
            sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
            truth = []
            x_in = tape[0]
            for k in range(steps):
                _, x0_64 = R.p_sample_step(lambda x, t: R.mdm_smpl_forward(sd64, x, t, cond.double(), faithful=False), tables, x_in.double(),
                                           steps - 1 - k, tape[k + 1].double(), gt.double(), mask)
                truth.append(x0_64)
                x_in = traj[k][0]
        _ORACLE_CACHE[key] = (tape, traj, truth)
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("backend", ["tcgen05", "simt"])
def test_config2_full_loop(eng, backend):
    """BASELINE configs[1] exactly: B=64, T=30, 100 DDPM steps, inpainted past."""
    B, T, steps = 64, 30, 100
    eng.set_gemm_backend(backend)
    sd = mdm_weights("smpl", "auto")
    key_sd = _ORACLE_CACHE.setdefault("sd", sd)       # one weight dict for both backends -> one oracle run
    eng.load_denoiser(key_sd, "smpl")
    b = S.make_smpl_batch(B=B, T=T)
    eng.bind(b["cond"], T)
    eng.init_diffusion(R.named_beta_schedule("cosine", steps))
    tape, traj, truth = _oracle_config2(key_sd, b, steps)
    gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()
    tape_d = tape.cuda()
    # (1) teacher-forced: every one of the 100 steps starts from the ORACLE's x_t.
    #     sample x_{t-1}: 2e-4 against the oracle on every step.
    #     pred_xstart: measured against the float64 evaluation of the same step; this implementation may not be further
    #     from it than a small multiple of the float32 oracle's own distance (or 2e-4), and on (geometric) average over
    #     the 100 steps not further than the float32 oracle.
    worst, worst_k, e_gpu, e_f32 = 0.0, -1, [], []
    x_ref = tape[0]
    for k in range(steps):
        i = steps - 1 - k
        got, got0 = eng.p_sample(i, x_ref.cuda(), tape_d[k + 1], gt, mask)
        e = rel(got, traj[k][0])
        if e > worst:
            worst, worst_k = e, k
        e_gpu.append(rel(got0, truth[k]))
        e_f32.append(rel(traj[k][1], truth[k]))
        x_ref = traj[k][0]
    assert worst < 2e-4, (worst, worst_k)
    bad = [(k, a, c) for k, (a, c) in enumerate(zip(e_gpu, e_f32)) if a > max(2e-4, 5.0 * c)]
    assert not bad, bad[:5]
    assert max(e_gpu) < 2e-3
    gmean = float(np.exp(np.mean(np.log(np.maximum(e_gpu, 1e-9)) - np.log(np.maximum(e_f32, 1e-9)))))
    assert gmean < 2.0, gmean
    # (2) free-running, step by step (the p_sample chain is what the loop replays), curve against the oracle
    x = tape_d[0].clone()
    curve = []
    for k in range(steps):
        x, _ = eng.p_sample(steps - 1 - k, x, tape_d[k + 1], gt, mask)
        curve.append(rel(x, traj[k][0]))
    cut, self_curve = _cutoff()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "config2_curve_%s.json" % backend), "w") as f:
        json.dump(dict(backend=backend, teacher_forced_sample_worst=worst, teacher_forced_sample_worst_step=worst_k,
                       teacher_forced_x0_vs_f64=e_gpu, oracle_f32_x0_vs_f64=e_f32, x0_error_ratio_geomean=gmean, free_running_vs_oracle=curve,
                       reference_f32_vs_f64=self_curve, cutoff_step=cut), f)
    bad = [(k, v) for k, v in enumerate(curve[:cut]) if v > 1e-3]
    assert not bad, ("free-running divergence above 1e-3 before the reference's own self-disagreement exceeds 1e-4", bad[:5], cut)
    # the in-library loop (graph replays) is the same computation: all three graph modes give the same bits
    outs = [eng.p_sample_loop(tape_d, gt, mask, use_graph=m).clone() for m in ("step", "loop", "off")]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel(outs[0], x) < 1e-6          # tokens emitted by the tail kernel == tokens re-derived by p_sample
    assert torch.equal(outs[0][mask], gt[mask])
    eng.set_gemm_backend("tcgen05")


def _correction_ctx(b, smplh_np, psd, past=10, future=20):
    return dict(past_len=past, future_len=future, smpl_dim=132, gt=torch.from_numpy(b["gt"]), hand_pose=torch.from_numpy(b["hand_pose"]),
                betas=torch.from_numpy(b["betas"]), obj_points=torch.from_numpy(b["obj_points"]), smplh=smplh_torch(smplh_np),
                projector=psd)


def _logging_denoised_fn(ctx, log):
    """oracle hook that also records the decisions of each active step"""
    inner = R.make_denoised_fn(ctx)

    def fn(x, t, kw=None):
        if not (t[0] > 500 or t[0] % 50 != 0):
            obs = R.correction_observables(x, ctx)
            log.append((int(t[0]), obs["condition"].clone(), obs["contact"].clone()))
        return inner(x, t, kw)
    return fn


@pytest.mark.parametrize("source", ["random", "ref"])
@pytest.mark.parametrize("mode", ["step", "loop", "off"])
def test_config3_loop_with_correction(eng, smplh_np, mode, source):
    """The product path of configs[2]: idb_p_sample_loop(correction=1) - gate (i <= 500 and i % 50 == 0), predict ->
    hook -> posterior with no re-inpainting, the (t/1000) blend - against the oracle loop with the restated hook, on a
    52-step schedule (hook active at i = 50 and i = 0).
      * decisions (condition / contact) of both hook steps: exact, for both weight sets;
      * the fused loop equals the step-by-step composition of the single-step entry points (each pinned to the oracle
        teacher-forced elsewhere in this file) bit for bit;
      * final sample against the oracle at 1e-3 with the seeded random-init denoiser, whose chain is well conditioned
        (its float32 and float64 oracle runs end 9e-7 apart).  With the shipped checkpoint the 52-step chain is chaotic
        like configs[1]'s (float32 vs float64 ORACLE runs end 0.19 apart, profiles/r2_conditioning_probe.txt), so there
        the final sample is only checked to be finite."""
    B, T, steps = 3, 30, 52
    sd, psd = mdm_weights("smpl", source), projector_weights("auto")
    eng.load_denoiser(sd, "smpl")
    eng.load_body(smplh_np)
    eng.load_projector(psd, 10, 20)
    b = S.make_smpl_batch(B=B, T=T)
    eng.bind(b["cond"], T)
    eng.bind_correction(b["hand_pose"], b["betas"], b["obj_points"], past_len=10)
    betas = R.named_beta_schedule("cosine", steps)
    eng.init_diffusion(betas)
    gt, mask, cond = torch.from_numpy(b["gt"]), torch.from_numpy(b["mask"]), torch.from_numpy(b["cond"])
    tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps))
    key = ("c3", steps, source)
    if key not in _ORACLE_CACHE:
        log = []
        ctx = _correction_ctx(b, smplh_np, psd)
        fwd = lambda x, t: R.mdm_smpl_forward(sd, x, t, cond, faithful=False)
        with torch.no_grad():
            ref = R.p_sample_loop(fwd, R.diffusion_tables(betas), tape, gt, mask, denoised_fn=_logging_denoised_fn(ctx, log))
            plain = R.p_sample_loop(fwd, R.diffusion_tables(betas), tape, gt, mask)
        _ORACLE_CACHE[key] = (ref, plain, log)
    ref, plain, log = _ORACLE_CACHE[key]
    assert [t for t, _, _ in log] == [50, 0]
    gtd, maskd, taped = gt.cuda(), mask.cuda(), tape.cuda()
    cond_log, contact_log = eng.correction_log(4)
    got = eng.p_sample_loop(taped, gtd, maskd, correction=True, use_graph=mode).cpu()
    n_ok = 0
    for k, (_, c_ref, ct_ref) in enumerate(log):
        assert torch.equal(cond_log[k].cpu().bool(), c_ref), k
        assert torch.equal(contact_log[k].cpu().long(), ct_ref), k
        n_ok += 1
    assert n_ok == 2 and int(cond_log[2:].sum()) == 0 and int(contact_log[2:].abs().sum()) == 0      # exactly two hook steps ran
    eng.correction_log(0)
    # the same loop composed from the single-step entry points of the C ABI
    x = taped[0].clone()
    for k, i in enumerate(reversed(range(steps))):
        if i <= 500 and i % 50 == 0:
            x0 = eng.p_sample_predict(i, x, gtd, maskd)
            eng.correction_apply(x0, gtd, i)
            x = eng.p_sample_finish(i, x0, x, taped[k + 1])
        else:
            x, _ = eng.p_sample(i, x, taped[k + 1], gtd, maskd)
    assert torch.equal(got, x.cpu())
    assert torch.isfinite(got).all()
    if source == "random":
        assert rel(got, ref) < 1e-3
        if any(bool(c.any()) for _, c, _ in log):
            assert rel(plain, ref) > 1e-3    # the hook changed the sample: the comparison above is not vacuous
    # at i = 0 the sample IS the corrected pred_xstart (coef1 = 1, coef2 = 0, sigma = 0): where condition holds the object's
    # PAST frames were replaced by the projector blend and NOT re-inpainted (gaussian_diffusion.py:354-376)
    c0 = log[-1][1]
    if c0.any():
        assert not torch.equal(got[c0][:, :, 135:, :10], gt[c0][:, :, 135:, :10])
    assert torch.equal(got[:, :, :135, :10], gt[:, :, :135, :10])      # body channels keep the inpainted past


@pytest.mark.parametrize("i", [500, 0])
def test_correction_step_teacher_forced_1000(eng, smplh_np, i):
    """One p_sample of the 1000-step schedule with the hook active (t = 500: blend 0.5/0.5; t = 0: pure projector output,
    posterior = pred_xstart), from the same x_t: predict -> idb_correction_apply -> finish against the oracle step."""
    B, T, steps = 3, 30, 1000
    sd, psd = mdm_weights("smpl", "auto"), projector_weights("auto")
    eng.load_denoiser(sd, "smpl")
    eng.load_body(smplh_np)
    eng.load_projector(psd, 10, 20)
    b = S.make_smpl_batch(B=B, T=T)
    eng.bind(b["cond"], T)
    eng.bind_correction(b["hand_pose"], b["betas"], b["obj_points"], past_len=10)
    eng.init_diffusion(R.named_beta_schedule("cosine", steps))
    gt, mask, cond = torch.from_numpy(b["gt"]), torch.from_numpy(b["mask"]), torch.from_numpy(b["cond"])
    tables = R.diffusion_tables(R.named_beta_schedule("cosine", steps))
    g = torch.Generator().manual_seed(i + 1)
    a = float(np.sqrt(tables["alphas_cumprod"][i]))
    x_t = a * gt + float(np.sqrt(1 - a * a)) * torch.randn(gt.shape, generator=g)      # q_sample(gt, i)
    noise = torch.randn(gt.shape, generator=g)
    log = []
    ctx = _correction_ctx(b, smplh_np, psd)
    with torch.no_grad():
        ref, ref0 = R.p_sample_step(lambda x, t: R.mdm_smpl_forward(sd, x, t, cond, faithful=False), tables, x_t, i, noise, gt, mask,
                                    denoised_fn=_logging_denoised_fn(ctx, log))
    x0 = eng.p_sample_predict(i, x_t.cuda(), gt.cuda(), mask.cuda())
    x0, dbg = eng.correction_apply(x0, gt.cuda(), i, debug=True)
    got = eng.p_sample_finish(i, x0, x_t.cuda(), noise.cuda())
    assert torch.equal(dbg["condition"].cpu(), log[0][1]) and torch.equal(dbg["contact"].cpu().long(), log[0][2])
    assert rel(x0, ref0) < 2e-4 and rel(got, ref) < 2e-4
    if i == 0:
        assert torch.equal(got, x0)


# ------------------------------------------------------------------------------------------------------------------
# the drop-in veneer
# ------------------------------------------------------------------------------------------------------------------
SMPL_ARGS = dict(embedding_dim=256, smpl_dim=132, use_pointnet2=1, dropout=0.0, num_heads=4, ff_size=1024, activation="gelu",
                 latent_usage="memory", future_len=25, past_len=10, cond_mask_prob=0, diffusion_steps=1000, noise_schedule="cosine",
                 sigma_small=True, weight_v=0.2)


def _mirror_model(steps, source="auto"):
    from interdiff_b200.model.diffusion_smpl import create_model_and_diffusion
    args = Namespace(**{**SMPL_ARGS, "diffusion_steps": steps})
    model, diffusion = create_model_and_diffusion(args)
    sd = encoder_weights(source)
    own = model.state_dict()
    model.load_state_dict({k: (sd[k].reshape(own[k].shape) if k in sd else v) for k, v in own.items()})
    return model.cuda().eval(), diffusion, {k: v for k, v in model.state_dict().items()}


def _mirror_projector():
    from interdiff_b200.model.correction_smpl import ObjProjector
    proj = ObjProjector(Namespace(dct=10, num_verts=67, dropout=0.0, past_len=10, future_len=20))
    psd = projector_weights("auto")
    own = proj.state_dict()
    proj.load_state_dict({k: (psd[k].reshape(own[k].shape) if k in psd else v) for k, v in own.items()})
    return proj.cuda().eval(), psd


class _ObjModel:            # the reference passes the Lightning module; the hook reads `.model` (eval_smpl_short.py:124)
    def __init__(self, m):
        self.model = m


def _python_denoised_fn(past_len):
    """eval_smpl_short.py:84-130 transcribed onto the MIRRORED modules (pytorch3d's pure-torch transforms from
    oracle.transforms, which run on CUDA tensors like any torch code)."""
    from interdiff_b200.data.tools import vertex_normals
    from interdiff_b200.engine import MARKERSET_SSM67_SMPLH
    from interdiff_b200.tools import point2point_signed

    def denoised_fn(x, t, model_kwargs):
        if t[0] > 500 or t[0] % 50 != 0:
            return x
        y = model_kwargs["y"]
        xs = x.squeeze(1).permute(2, 0, 1).contiguous()
        body, obj = xs[..., :135], xs[..., 135:]
        T, B, _ = body[:, :, :-3].shape
        obj_rot = tf.rotation_6d_to_matrix(obj[:, :, :-3].reshape(T, B, 6))
        body_rot = tf.matrix_to_axis_angle(tf.rotation_6d_to_matrix(body[:, :, :-3].reshape(T, B, -1, 6))).reshape(T, B, -1)
        body_pred = torch.cat([body_rot, y["hand_pose"], body[:, :, -3:]], dim=2)
        bb = body_pred.view(T * B, -1)
        verts, jtr, _, _ = y["smpl"](bb[:, :-3], th_betas=y["beta"].view(T * B, -1), th_trans=bb[:, -3:])
        human_verts = verts.view(T, B, -1, 3)[:, :, MARKERSET_SSM67_SMPLH]
        obj_points = y["obj_points"]
        obj_pred = torch.matmul(obj_points.unsqueeze(0), obj_rot.permute(0, 1, 3, 2)) + obj[:, :, -3:].unsqueeze(2)
        faces = y["smpl"].th_faces.unsqueeze(0).repeat(T * B, 1, 1)
        normals = vertex_normals(verts, faces)
        o2h_signed, h2o_signed, o2h_idx, h2o_idx, o2h, h2o = point2point_signed(verts, obj_pred.view(T * B, -1, 3), x_normals=normals, return_vector=True)
        w = torch.zeros(T * B, o2h_signed.size(1), device=x.device)
        w[o2h_signed < 0] = 20
        loss_dist_o = (torch.abs(o2h_signed) * w).view(T, B, -1)
        d = torch.stack([torch.norm(human_verts[t_].unsqueeze(1) - obj_pred[t_].unsqueeze(2), dim=3) for t_ in range(T)])   # (T,B,P,67)
        distance = d.min(dim=3)[0].min(dim=2)[0].mean(dim=0)
        condition = torch.logical_not(torch.logical_and(loss_dist_o[past_len:].mean(dim=2).mean(dim=0) < 0.002, distance < 0.02))
        contact = (d < 0.02).any(dim=2)[past_len:].sum(dim=0)
        gts = y["inpainted_motion"].squeeze(1).permute(2, 0, 1).contiguous()
        obj_gt = gts[..., 135:]
        obj_proj = y["obj_model"].model.sample(obj_gt[:, :, :-3], obj_gt[:, :, -3:], human_verts, contact)
        x_ = torch.cat([body, obj_proj], dim=2).permute(1, 2, 0).unsqueeze(1).contiguous()
        x_ = t[0] / 1000 * x + (1 - t[0] / 1000) * x_
        x[condition] = x_[condition]
        denoised_fn.decisions.append((condition.clone(), contact.clone()))
        return x
    denoised_fn.decisions = []
    return denoised_fn


@pytest.mark.parametrize("steps", [12, 52])
def test_veneer_sample_once_proj(smplh_np, steps):
    """sample_once_proj (eval_smpl_short.py:133-215) through the mirrored API only:
       embedding, gt = model._get_embeddings(batch)  ->  diffusion.p_sample_loop(model, shape, clip_denoised=False, noise,
       model_kwargs, denoised_fn).  (a) denoised_fn = FusedCorrection: whole loop in the library; must equal the Engine path
       on the same tape BIT FOR BIT.  (b) denoised_fn = plain-Python transcription of the reference's hook on the mirrored
       SMPL_Layer / vertex_normals / point2point_signed / ObjProjector.sample (arbitrary-callback path, step by step):
       same decisions, same sample to rounding (its torch-side rotations / object transform round differently from the
       fused kernels).  12-step schedule: hook at i = 0 only; 52 steps: i = 50 and 0."""
    from interdiff_b200.libsmpl.smplpytorch.pytorch.smpl_layer import SMPL_Layer
    from interdiff_b200.sampling import FusedCorrection, draw_tape
    B, T, past = 3, 30, 10
    # 52 steps: seeded random-init weights (a well-conditioned chain, so the end-to-end oracle comparison below means
    # something; with the shipped checkpoint a 52-step chain amplifies rounding to O(0.1), see test_config3_loop_with_correction)
    model, diffusion, msd = _mirror_model(steps, "auto" if steps == 12 else "random")
    proj, psd = _mirror_projector()
    smpl = SMPL_Layer.from_arrays(smplh_np).cuda()
    b = S.make_smpl_batch(B=B, T=T)
    g = torch.Generator().manual_seed(7)
    frames = [dict(smplfit_params=dict(pose=0.3 * torch.randn(B, 156, generator=g), trans=0.3 * torch.randn(B, 3, generator=g)),
                   objfit_params=dict(angle=0.5 * torch.randn(B, 3, generator=g), trans=0.3 * torch.randn(B, 3, generator=g))) for _ in range(T)]
    pts = torch.from_numpy(b["obj_points"]).float()
    batch = dict(frames=frames, obj_points=torch.cat([pts, torch.zeros(B, pts.shape[1], 4)], dim=2))
    with torch.no_grad():
        embedding, gt = model._get_embeddings(batch, device="cuda")
    gt = gt.permute(1, 2, 0).unsqueeze(1).contiguous()                       # (B,1,144,T) as sample_once_proj builds it (:139-141)
    mask = torch.zeros_like(gt, dtype=torch.bool)
    mask[..., :past] = True
    hand_pose = torch.from_numpy(b["hand_pose"]).cuda()
    beta = torch.from_numpy(b["betas"]).cuda()
    kw = lambda: {"y": {"cond": embedding, "inpainted_motion": gt, "inpainting_mask": mask, "hand_pose": hand_pose, "smpl": smpl,
                        "beta": beta, "obj_model": _ObjModel(proj), "obj_points": pts.cuda()}}
    x_T = torch.randn(gt.shape, generator=torch.Generator().manual_seed(3)).cuda()
    # (a) fused hook object
    torch.manual_seed(11)
    fused = diffusion.p_sample_loop(model, gt.shape, clip_denoised=False, noise=x_T, model_kwargs=kw(), denoised_fn=FusedCorrection(past)).clone()
    # Engine path on the same tape
    eng = model.engine_for(embedding.device)
    torch.manual_seed(11)
    tape = draw_tape(eng, x_T, steps)
    direct = eng.p_sample_loop(tape, gt, mask, correction=True).clone()
    assert torch.equal(fused, direct)
    # (b) plain-Python callback on the mirrored modules
    fn = _python_denoised_fn(past)
    torch.manual_seed(11)
    cond_log, contact_log = eng.correction_log(4)
    again = eng.p_sample_loop(tape, gt, mask, correction=True).clone()            # logs the fused decisions
    eng.correction_log(0)
    assert torch.equal(again, direct)
    torch.manual_seed(11)
    py = diffusion.p_sample_loop(model, gt.shape, clip_denoised=False, noise=x_T, model_kwargs=kw(), denoised_fn=fn)
    n_active = 2 if steps > 50 else 1
    assert len(fn.decisions) == n_active
    for k, (c, ct) in enumerate(fn.decisions):
        assert torch.equal(cond_log[k].bool(), c) and torch.equal(contact_log[k].long(), ct.long()), k
    assert rel(py, direct) < (1e-4 if steps == 12 else 1e-3)
    # against the oracle end to end (model weights = the mirrored module's own state dict)
    ctx = dict(past_len=past, future_len=T - past, smpl_dim=132, gt=gt.cpu(), hand_pose=hand_pose.cpu(), betas=beta.cpu(), obj_points=pts,
               smplh=smplh_torch(smplh_np), projector=psd)
    cpu_sd = {k: v.cpu() for k, v in msd.items()}
    with torch.no_grad():
        ref = R.p_sample_loop(lambda x, t: R.mdm_smpl_forward(cpu_sd, x, t, embedding.cpu(), faithful=False),
                              R.diffusion_tables(R.named_beta_schedule("cosine", steps)), tape.cpu(), gt.cpu(), mask.cpu(),
                              denoised_fn=R.make_denoised_fn(ctx))
    assert rel(direct, ref) < 1e-3


def test_veneer_modules(smplh_np):
    """SMPL_Layer.__call__ (incl. the zero-betas branch and the default arguments), tools.point2point_signed,
    data.tools.vertex_normals, ObjProjector.sample and sample_postprocess through the mirrored modules."""
    from interdiff_b200.data.tools import vertex_normals
    from interdiff_b200.libsmpl.smplpytorch.pytorch.smpl_layer import SMPL_Layer
    from interdiff_b200.sampling import sample_postprocess
    from interdiff_b200.tools import point2point_signed
    smplh = smplh_torch(smplh_np)
    arrays = dict(smplh_np)
    arrays["betas"] = np.linspace(-1.0, 1.0, 10, dtype=np.float32)            # template betas != 0: the branch is observable
    smpl = SMPL_Layer.from_arrays(arrays).cuda()
    g = torch.Generator().manual_seed(5)
    Fn = 20
    pose, betas, trans = 0.4 * torch.randn(Fn, 156, generator=g), torch.randn(Fn, 10, generator=g), torch.randn(Fn, 3, generator=g)
    v, j, _, _ = smpl(pose.cuda(), th_betas=betas.cuda(), th_trans=trans.cuda())
    v_ref, j_ref = R.smplh_lbs(smplh, pose, betas, trans)
    assert rel(v, v_ref) < 1e-5 and rel(j, j_ref) < 1e-5
    # all-zero betas of full shape (and the default argument) take the layer's own template betas (smpl_layer.py:96-100)
    tb = torch.from_numpy(arrays["betas"]).view(1, 10).expand(Fn, -1)
    v0_ref, j0_ref = R.smplh_lbs(smplh, pose, tb, trans)
    for given in (torch.zeros(Fn, 10).cuda(), None):
        v0, j0, _, _ = smpl(pose.cuda(), th_trans=trans.cuda()) if given is None else smpl(pose.cuda(), th_betas=given, th_trans=trans.cuda())
        assert rel(v0, v0_ref) < 1e-5 and rel(j0, j0_ref) < 1e-5
    # a second layer sharing the first one's engine reloads it instead of silently reusing the other model
    other = dict(smplh_np)
    other["v_template"] = smplh_np["v_template"] * 1.1
    smpl2 = SMPL_Layer.from_arrays(other).cuda()
    engine = smpl.engine_for(v.device)
    smpl2.load_into(engine)
    smpl.load_into(engine)
    v_again, _ = engine.lbs(pose, betas, trans)
    assert torch.equal(v_again, v)
    # geometry helpers
    faces = smpl.th_faces.unsqueeze(0).repeat(3, 1, 1)
    verts = v[:3].contiguous()
    n = vertex_normals(verts, faces)
    n_ref = R.vertex_normals(verts.cpu(), smplh["faces"])
    assert rel(n, n_ref) < 1e-5
    y = (verts[:, ::17][:, :300] * 1.03).contiguous()
    out = point2point_signed(verts, y, x_normals=n, return_vector=True)
    ref = R.point2point_signed(verts.cpu(), y.cpu(), n_ref)
    assert torch.equal(out[2].cpu().long(), ref[2]) and rel(out[0], ref[0]) < 1e-5 and rel(out[4], ref[4]) < 1e-5
    # ObjProjector.sample
    proj, psd = _mirror_projector()
    T, B = 30, 4
    ang, tr, hv = torch.randn(T, B, 6, generator=g), torch.randn(T, B, 3, generator=g), torch.randn(T, B, 67, 3, generator=g)
    contact = (torch.rand(B, 67, generator=g) < 0.05).long()
    got = proj.sample(ang.cuda(), tr.cuda(), hv.cuda(), contact.cuda())
    with torch.no_grad():
        want = R.obj_projector_sample(psd, ang, tr, hv, contact, 10, 20)
    assert rel(got, want) < 1e-4
    # sample_postprocess
    b = S.make_smpl_batch(B=B, T=T)
    sample = torch.from_numpy(b["gt"])
    ctx = dict(smpl_dim=132, hand_pose=torch.from_numpy(b["hand_pose"]), betas=torch.from_numpy(b["betas"]), smplh=smplh)
    engine.load_body(smplh_np)
    engine._body_owner = None
    body, obj, vv, jj = sample_postprocess(engine, sample.cuda(), ctx["hand_pose"].cuda(), ctx["betas"].cuda())
    rb, ro, rv, rj = R.sample_postprocess(sample, ctx)
    assert rel(body, rb) < 1e-5 and rel(obj, ro) < 1e-5 and rel(vv, rv) < 1e-5 and rel(jj, rj) < 1e-5


def test_unconditioned_sampling_without_inpainting_keys():
    """the inpainting keys are optional upstream (gaussian_diffusion.py:307): p_sample / p_sample_loop must infer T from
    the sample shape"""
    model, diffusion, msd = _mirror_model(6)
    b = S.make_smpl_batch(B=2, T=30)
    cond = torch.from_numpy(b["cond"]).cuda()
    shape = b["gt"].shape
    torch.manual_seed(0)
    x_T = torch.randn(shape, device="cuda")
    out = diffusion.p_sample_loop(model, shape, clip_denoised=False, noise=x_T, model_kwargs={"y": {"cond": cond}})
    assert out.shape == tuple(shape) and torch.isfinite(out).all()
    o = diffusion.p_sample(model, x_T, torch.full((2,), 5, device="cuda"), clip_denoised=False, model_kwargs={"y": {"cond": cond}})
    assert torch.isfinite(o["sample"]).all()


def test_graph_cache_invalidation(eng):
    """the captured graphs bake in workspace / weight pointers, Tm and the GEMM backend.  Loop, then rebind
    with another memory length / reload other weights / switch the backend, loop again with graphs: must equal eager."""
    B, T, steps = 4, 30, 6
    sd = mdm_weights("smpl", "auto")
    eng.load_denoiser(sd, "smpl")
    b = S.make_smpl_batch(B=B, T=T)
    eng.init_diffusion(R.named_beta_schedule("cosine", steps))
    tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps)).cuda()
    gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()

    def both():
        outs = [eng.p_sample_loop(tape, gt, mask, use_graph=m).clone() for m in ("step", "loop", "off")]
        assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[2])
        return outs[2]
    eng.bind(b["cond"], T)
    a0 = both()
    cond7 = np.ascontiguousarray(b["cond"][:7])                 # other Tm: every memory tensor is reallocated
    eng.bind(cond7, T)
    a1 = both()
    assert not torch.equal(a0, a1)
    sd2 = {k: (v * 1.01 if k.endswith("linear1.weight") else v) for k, v in sd.items()}
    eng.load_denoiser(sd2, "smpl")                              # weight reload frees and re-packs everything
    eng.bind(cond7, T)
    a2 = both()
    assert not torch.equal(a1, a2)
    eng.set_gemm_backend("simt")
    a3 = both()
    eng.set_gemm_backend("tcgen05")
    a4 = both()
    assert torch.equal(a4, a2) and rel(a3, a2) < 1e-3
    # fresh gt / mask / tape tensors reuse the captured graphs (stable internal copies + tape slot) and still see the new data
    tape2 = tape.clone()
    tape2[1:] *= 0.5
    l0 = eng.launch_count
    b1 = eng.p_sample_loop(tape2, gt.clone(), mask.clone(), use_graph="loop").clone()
    b2 = eng.p_sample_loop(tape2, gt, mask, use_graph="off").clone()
    assert torch.equal(b1, b2) and not torch.equal(b1, a4)
    assert eng.launch_count > l0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_engine_on_non_current_device():
    """an Engine created for cuda:1 must work while cuda:0 is current (every ABI entry sets its device)."""
    from interdiff_b200.engine import Engine
    torch.cuda.set_device(0)
    e0, e1 = Engine("cuda:0"), Engine("cuda:1")
    sd = mdm_weights("smpl", "auto")
    b = S.make_smpl_batch(B=2, T=30)
    outs = []
    for e in (e0, e1):
        e.load_denoiser(sd, "smpl")
        e.bind(b["cond"], 30)
        x = torch.from_numpy(S.noise_tape(b["gt"].shape, 0)[0]).to(e.device)
        outs.append(e.forward(x, torch.tensor([400, 3]).to(e.device)).cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])
    e0.close()
    e1.close()


def test_batch_slices_bit_identical(eng):
    """multi-GPU runs slice ONE global batch (contiguous B/G samples per rank, noise keyed by the global sample
    index).  A sample's trajectory must not depend on which other samples share its batch: the two halves of a B=64 batch
    (M = 1920 rows: 256-column GEMM tiles) sampled as B=32 batches (M = 960: 128-column tiles) reproduce the full-batch
    result bit for bit, and so do B=8 slices."""
    sd = mdm_weights("smpl", "auto")
    eng.load_denoiser(sd, "smpl")
    B, T, steps = 64, 30, 10
    b = S.make_smpl_batch(B=B, T=T)
    eng.init_diffusion(R.named_beta_schedule("cosine", steps))
    tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps)).cuda()
    gt, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()
    eng.bind(b["cond"], T)
    full = eng.p_sample_loop(tape, gt, mask).clone()
    for G in (2, 8):
        n = B // G
        parts = []
        for r in range(G):
            sl = slice(r * n, (r + 1) * n)
            eng.bind(np.ascontiguousarray(b["cond"][:, sl]), T)
            parts.append(eng.p_sample_loop(tape[:, sl].contiguous(), gt[sl].contiguous(), mask[sl].contiguous()).clone())
        assert torch.equal(torch.cat(parts, dim=0), full), G


def test_rollout_next_window_and_driver(eng, smplh_np):
    """SURVEY 8f rank 2 / BASELINE configs[4]: the device-side step between two windows (get_batch + denormalize) against its
    oracle restatement, and the driver's world-coordinate trajectory: window k+1 starts exactly where window k ended."""
    from interdiff_b200.rollout import RolloutDriver
    from interdiff_b200.sampling import sample_postprocess
    g = torch.Generator().manual_seed(21)
    T, B, P = 30, 3, 10
    body = torch.cat([0.8 * torch.randn(T, B, 66, generator=g), 0.1 * torch.randn(T, B, 90, generator=g), torch.randn(T, B, 3, generator=g)], dim=2)
    body[3, 0, 3:6] = 0.0                                     # small-angle branch of axis_angle_to_quaternion
    obj = torch.cat([0.8 * torch.randn(T, B, 3, generator=g), torch.randn(T, B, 3, generator=g)], dim=2)
    jtr = torch.randn(T, B, 52, 3, generator=g)
    gt, cen = eng.rollout_next_window(body, obj, jtr, T, P)
    gt_ref, cen_ref = R.rollout_next_window(body, obj, jtr, T, P)
    assert torch.equal(cen.cpu(), cen_ref) and rel(gt, gt_ref) < 1e-6
    assert torch.equal(gt[..., P:].cpu(), gt[..., P - 1:P].cpu().expand(-1, -1, -1, T - P))       # future inputs repeat the last past frame
    x = torch.randn(T, B, 5, 3, generator=g).cuda()
    y = eng.add_offset_(x.clone(), cen)
    assert rel(y, x.cpu() + cen_ref.view(1, B, 1, 3)) < 1e-7
    # driver: 1 + 2 windows on a short schedule
    sd, psd = mdm_weights("smpl", "random"), projector_weights("auto")
    eng.load_denoiser(sd, "smpl")
    eng.load_body(smplh_np)
    eng.load_projector(psd, P, T - P)
    b = S.make_smpl_batch(B=B, T=T)
    steps = 6
    eng.bind(b["cond"], T)
    hp, bt = torch.from_numpy(b["hand_pose"]).cuda(), torch.from_numpy(b["betas"]).cuda()
    eng.bind_correction(hp, bt, b["obj_points"], past_len=P)
    eng.init_diffusion(R.named_beta_schedule("cosine", steps))
    tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps)).cuda()
    gt0, mask = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda()
    drv = RolloutDriver(eng, past_len=P, n_windows=2, correction=True)
    traj = drv.run(tape, gt0, mask, hp, bt, keep=("body", "obj", "pelvis", "jtr"))
    F = T - P
    assert traj["body"].shape[0] == P + 3 * F and traj["obj"].shape[0] == P + 3 * F
    # manual composition of the same rollout with the ORACLE's window step (host side), engine loops for the sampling
    offset = torch.zeros(B, 3)
    cur = gt0
    bodies = []
    for k in range(3):
        s = eng.p_sample_loop(tape, cur, mask, correction=True)
        bd, ob, _, jt = sample_postprocess(eng, s, hp, bt)
        bd, ob, jt = bd.cpu(), ob.cpu(), jt.cpu()
        nxt, c = R.rollout_next_window(bd, ob, jt, T, P)
        w = bd.clone()
        w[..., -3:] += offset
        bodies.append(w[0 if k == 0 else P:])
        offset = offset + c
        cur = nxt.cuda()
    # two compositions of the same chain whose window steps differ at rounding level (device vs host get_batch, 1e-6 above);
    # three windows of a free-running sampler amplify that, hence 1e-4 and not 1e-6
    assert rel(traj["body"], torch.cat(bodies)) < 1e-4
    # continuity: the inpainted past of window k+1 (frames [P + k F - P .. ) in world coordinates) IS the end of window k
    assert torch.isfinite(traj["pelvis"]).all()


@pytest.mark.parametrize("source", ["random", "ref"])
def test_skeleton_correction(eng, source):
    """the skeleton correction net on the device (model/correction_skeleton.py:84-135; templated projector
    kernel, n_pre 20, joint stack 9-64-32-64-9), its hook (eval_skeleton.py:80-111) and the in-loop path of the skeleton
    sampler, against the oracle (which is pinned to the reference class with checkpoints/obj_skeleton.ckpt)."""
    from tests.helpers import projector_skeleton_weights
    psd = projector_skeleton_weights(source)
    P, Fu = 10, 10
    T, B = P + Fu, 4
    eng.load_projector_skeleton(psd, P, Fu, n_joints=21)
    g = torch.Generator().manual_seed(31)
    quat = torch.randn(T, B, 4, generator=g)
    quat[0, 0] = torch.tensor([0.0, 0.0, 0.0, 2.0])                       # un-normalised identity
    tr, joints = torch.randn(T, B, 3, generator=g), torch.randn(T, B, 21, 3, generator=g)
    q, t = eng.projector_sample_skeleton(quat, tr, joints)
    with torch.no_grad():
        q_ref, t_ref = R.obj_projector_skeleton_sample(psd, quat, tr, joints, P, Fu)
    assert rel(q, q_ref) < 1e-4 and rel(t, t_ref) < 1e-4
    # the mirrored module (strict checkpoint names) routes to the same kernel
    from interdiff_b200.model.correction_skeleton import ObjProjector
    proj = ObjProjector(Namespace(num_joints=21, dropout=0.0, past_len=P, future_len=Fu, embedding_dim=128))
    own = proj.state_dict()
    proj.load_state_dict({k: (psd[k].reshape(own[k].shape) if k in psd else v) for k, v in own.items()}, strict=True)
    q2, t2 = proj.cuda().eval().sample(quat.cuda(), tr.cuda(), joints.cuda())
    assert torch.equal(q2, q) and torch.equal(t2, t)
    # hook body at an active step
    b = S.make_skeleton_batch(B=B, T=T)
    gt, zp = torch.from_numpy(b["gt"]), torch.from_numpy(b["zero_pose_obj"])
    x = gt + 0.05 * torch.randn(gt.shape, generator=g)
    ctx = dict(gt=gt, zero_pose_obj=zp, projector=psd, past_len=P, future_len=Fu)
    with torch.no_grad():
        want = R.make_denoised_fn_skeleton(ctx)(x.clone(), torch.full((B,), 450), None)
    got = eng.skeleton_correction_apply(x.clone().cuda(), gt, zp, 450)
    assert rel(got, want) < 1e-4
    # in the sampling loop of the skeleton denoiser: hook at i = 50 and i = 0 of a 52-step schedule
    sd = mdm_weights("skeleton", "random")
    eng.load_denoiser(sd, "skeleton")
    eng.bind(b["cond"], T, zero_pose_obj=b["zero_pose_obj"])
    steps = 52
    betas = R.named_beta_schedule("cosine", steps)
    eng.init_diffusion(betas)
    tape = torch.from_numpy(S.noise_tape(b["gt"].shape, steps))
    mask, cond = torch.from_numpy(b["mask"]), torch.from_numpy(b["cond"])
    fwd = lambda xx, tt: R.mdm_skeleton_forward(sd, xx, tt, zp, cond)
    with torch.no_grad():
        ref = R.p_sample_loop(fwd, R.diffusion_tables(betas), tape, gt, mask, denoised_fn=R.make_denoised_fn_skeleton(ctx))
        plain = R.p_sample_loop(fwd, R.diffusion_tables(betas), tape, gt, mask)
    outs = [eng.p_sample_loop(tape.cuda(), gt.cuda(), mask.cuda(), correction=True, use_graph=m).cpu() for m in ("loop", "step", "off")]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel(outs[0], ref) < 1e-3 and rel(plain, ref) > 1e-3


def test_training_losses_forward_values():
    """SURVEY 8f rank 4 (second half): diffusion.training_losses(model, x_start, t, model_kwargs, noise) -> (model_output,
    target) as train_diffusion_smpl.py:61-70 binds it, with per-sample timesteps, against the oracle's forward on the same
    q_sample."""
    model, diffusion, msd = _mirror_model(1000)
    B, T = 4, 30
    b = S.make_smpl_batch(B=B, T=T)
    gt, mask, cond = torch.from_numpy(b["gt"]).cuda(), torch.from_numpy(b["mask"]).cuda(), torch.from_numpy(b["cond"]).cuda()
    t = torch.tensor([999, 500, 37, 0], device="cuda")
    noise = torch.randn(gt.shape, generator=torch.Generator().manual_seed(2)).cuda()
    kw = {"y": {"cond": cond, "inpainted_motion": gt, "inpainting_mask": mask}}
    out, target = diffusion.training_losses(model, gt, t, model_kwargs=kw, noise=noise)
    assert torch.equal(target, gt) and out.shape == gt.shape
    tables = R.diffusion_tables(R.named_beta_schedule("cosine", 1000))
    a = torch.tensor(tables["sqrt_alphas_cumprod"])[t.cpu()].float().view(B, 1, 1, 1)
    s1 = torch.tensor(tables["sqrt_one_minus_alphas_cumprod"])[t.cpu()].float().view(B, 1, 1, 1)
    x_t = a * gt.cpu() + s1 * noise.cpu()
    x_t = (x_t * ~mask.cpu()) + (gt.cpu() * mask.cpu())
    with torch.no_grad():
        ref = R.mdm_smpl_forward({k: v.cpu() for k, v in msd.items()}, x_t, t.cpu(), cond.cpu(), faithful=False)
    assert rel(out, ref) < 3e-4
