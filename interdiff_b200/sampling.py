"""Sampling drivers on top of the Engine (the L4 layer of SURVEY section 1).

sample_smpl_host is the end-to-end call bench.py times: host (pinned) buffers in, host buffer
out, everything in between on the device.  Mirrors reference eval_smpl_short.py:179-192
(sample_once) up to the final sample: model_kwargs = {cond, inpainted_motion, inpainting_mask},
noise = x_T, diffusion.p_sample_loop(model, shape, clip_denoised=False, noise, model_kwargs).
"""
import torch


def draw_tape(engine, x_T, n_steps, seed=None):
    """(n_steps+1, *shape) device tensor: tape[0] = x_T, tape[1:] = per-step eps drawn on the
    device (the reference draws th.randn_like(x) every step, gaussian_diffusion.py:532)."""
    tape = torch.empty((n_steps + 1,) + tuple(x_T.shape), device=engine.device)
    tape[0].copy_(x_T, non_blocking=True)
    g = None
    if seed is not None:
        g = torch.Generator(device=engine.device)
        g.manual_seed(int(seed))
    tape[1:].normal_(generator=g)
    return tape


def sample_smpl_host(engine, h_xT, h_gt, h_mask, h_cond, h_out, seed=None, correction=False, use_graph=True):
    """One sampling call with host tensors (pinned for async copies).  The engine must have its
    denoiser loaded and its diffusion initialised."""
    dev = engine.device
    gt = h_gt.to(dev, non_blocking=True)
    mask = h_mask.to(dev, non_blocking=True)
    cond = h_cond.to(dev, non_blocking=True)
    xT = h_xT.to(dev, non_blocking=True)
    T = gt.shape[-1]
    engine.bind(cond, T)
    # keep the tape buffer across calls (same shape) so the captured graph stays valid
    key = (tuple(gt.shape), engine.n_steps)
    cache = getattr(engine, "_host_loop_cache", None)
    if cache is None or cache[0] != key:
        cache = (key, torch.empty((engine.n_steps + 1,) + tuple(gt.shape), device=dev), torch.empty_like(gt), torch.empty_like(mask),
                 torch.empty_like(gt))
        engine._host_loop_cache = cache
    _, tape, gt_buf, mask_buf, out = cache
    gt_buf.copy_(gt)
    mask_buf.copy_(mask)
    tape[0].copy_(xT)
    g = None
    if seed is not None:
        g = torch.Generator(device=dev)
        g.manual_seed(int(seed))
    tape[1:].normal_(generator=g)
    engine.p_sample_loop(tape, gt_buf, mask_buf, correction=correction, use_graph=use_graph, out=out)
    h_out.copy_(out, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return h_out
