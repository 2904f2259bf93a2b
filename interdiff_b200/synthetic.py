"""Synthetic BEHAVE-shape inputs for tests and bench (there is no network for datasets, and
the SMPL-H model file is licensed and not shipped: reference
libsmpl/smplpytorch/native/models/README.md:1).  numpy only, seeded, device independent.

Shapes follow the reference's batch layout (interdiff/data/dataset_smpl.py:181-203,
interdiff/eval_smpl_short.py:133-150) and SURVEY.md section 8(d).
"""
import zlib

import numpy as np

NUM_VERTS = 6890
NUM_FACES = 13776
NUM_JOINTS = 52
NUM_POSE_BASIS = 459  # 51 joints x 9
NUM_BETAS = 10
NUM_OBJ_POINTS = 2048

# SMPL-H kinematic tree (kintree_table[0]; root stored as uint32(-1) in the pkl)
SMPLH_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
                 20, 22, 23, 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35,
                 21, 37, 38, 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50]


def _rng(seed, tag):
    return np.random.default_rng([int(seed), zlib.crc32(tag.encode())])


def make_sphere_mesh(rings=84, segs=82):
    """Closed genus-0 lat-long mesh with 2 + rings*segs = 6890 vertices and 2*rings*segs =
    13776 outward (CCW) faces -- the same counts as SMPL."""
    v = [(0.0, 1.0, 0.0)]
    for r in range(rings):
        th = np.pi * (r + 1) / (rings + 1)
        for s in range(segs):
            ph = 2 * np.pi * s / segs
            v.append((np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)))
    v.append((0.0, -1.0, 0.0))
    v = np.asarray(v, dtype=np.float64)
    f = []
    idx = lambda r, s: 1 + r * segs + (s % segs)
    for s in range(segs):
        f.append((0, idx(0, s + 1), idx(0, s)))
    for r in range(rings - 1):
        for s in range(segs):
            a, b, c, d = idx(r, s), idx(r, s + 1), idx(r + 1, s), idx(r + 1, s + 1)
            f.append((a, b, c))
            f.append((b, d, c))
    last = len(v) - 1
    for s in range(segs):
        f.append((last, idx(rings - 1, s), idx(rings - 1, s + 1)))
    f = np.asarray(f, dtype=np.int64)
    # enforce outward orientation
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    flip = (n * v[f].mean(1)).sum(1) < 0
    f[flip] = f[flip][:, [0, 2, 1]]
    assert v.shape[0] == NUM_VERTS and f.shape[0] == NUM_FACES
    return v, f


def make_smplh_model(seed=233, sparse_weights=False):
    """Synthetic SMPL-H-shaped body model: dict of float32/int64 arrays with the shapes of the
    buffers SMPL_Layer registers (reference smpl_layer.py:47-64).  sparse_weights: at most 4 NON-ZERO skinning
    weights per vertex (exact zeros elsewhere), which is how the licensed SMPL / SMPL-H models are painted; the
    default keeps a dense 1e-4 tail on every bone (the golden vectors were made with it, and it exercises the
    dense-weights path of the skinning kernel)."""
    rng = _rng(seed, "smplh")
    sph, faces = make_sphere_mesh()
    radii = np.array([0.22, 0.85, 0.14])
    v_template = sph * radii + np.array([0.0, -0.1, 0.0])
    v_template += 0.002 * rng.standard_normal(v_template.shape)
    # rest joints spread along the body; skinning weights by proximity (<=4 dominant bones)
    jpos = np.zeros((NUM_JOINTS, 3))
    jpos[:, 1] = np.linspace(0.7, -0.9, NUM_JOINTS)
    jpos[:, 0] = 0.12 * np.sin(np.arange(NUM_JOINTS) * 1.7)
    jpos[:, 2] = 0.05 * np.cos(np.arange(NUM_JOINTS) * 2.3)
    d = ((v_template[:, None, :] - jpos[None]) ** 2).sum(-1)  # (V, J)
    order = np.argsort(d, axis=1)[:, :4]
    w = np.zeros((NUM_VERTS, NUM_JOINTS))
    ww = np.exp(-np.take_along_axis(d, order, 1) / 0.01) + 1e-3
    np.put_along_axis(w, order, ww, 1)
    tail = 1e-4 * rng.random(w.shape)
    if not sparse_weights:
        w += tail  # dense small tail on every bone
    w /= w.sum(1, keepdims=True)
    # joint regressor: non-negative, row-normalised, localised around each joint
    jr = np.exp(-d.T / 0.005) + 1e-6 * rng.random((NUM_JOINTS, NUM_VERTS))
    jr /= jr.sum(1, keepdims=True)
    shapedirs = 0.01 * rng.standard_normal((NUM_VERTS, 3, NUM_BETAS))
    posedirs = 0.001 * rng.standard_normal((NUM_VERTS, 3, NUM_POSE_BASIS))
    return dict(
        v_template=v_template.astype(np.float32),
        shapedirs=shapedirs.astype(np.float32),
        posedirs=posedirs.astype(np.float32),
        J_regressor=jr.astype(np.float32),
        weights=w.astype(np.float32),
        faces=faces,
        parents=np.asarray(SMPLH_PARENTS, dtype=np.int64),
    )


def _aa_to_rot6d(aa):
    """axis-angle (...,3) -> first two ROWS of the rotation matrix (pytorch3d convention,
    reference model/diffusion_smpl.py:212-213), float64 numpy."""
    th = np.linalg.norm(aa, axis=-1, keepdims=True)
    k = aa / np.maximum(th, 1e-12)
    kx, ky, kz = k[..., 0], k[..., 1], k[..., 2]
    c, s = np.cos(th[..., 0]), np.sin(th[..., 0])
    C = 1 - c
    R = np.stack([
        c + kx * kx * C, kx * ky * C - kz * s, kx * kz * C + ky * s,
        ky * kx * C + kz * s, c + ky * ky * C, ky * kz * C - kx * s,
    ], axis=-1)
    return R


def make_box_points(n, half=0.2, seed=233):
    """n points uniform on the surface of a cube of half-size `half`."""
    rng = _rng(seed, "box")
    p = rng.uniform(-half, half, size=(n, 3))
    face = rng.integers(0, 6, size=n)
    ax, sgn = face // 2, (face % 2) * 2 - 1
    p[np.arange(n), ax] = sgn * half
    return p


def make_smpl_batch(B=64, T=30, past_len=10, seed=233, contact=True):
    """Synthetic sampling batch in the layout the sampler consumes:
       gt (B,1,144,T) = [22 x rot6d | body trans | obj rot6d | obj trans] channels,
       cond (10,B,256), mask (B,1,144,T) bool (True = past frames, keep gt),
       hand_pose (T,B,90), betas (T,B,10), obj_points (B,2048,3)."""
    rng = _rng(seed, "batch%d_%d" % (B, T))
    # smooth pose trajectories: base pose + slow drift
    base = 0.3 * rng.standard_normal((1, B, 22, 3))
    drift = 0.02 * np.cumsum(rng.standard_normal((T, B, 22, 3)), axis=0)
    body_aa = base + drift
    body_trans = 0.3 * rng.standard_normal((1, B, 3)) + 0.01 * np.cumsum(rng.standard_normal((T, B, 3)), 0)
    obj_aa = 0.5 * rng.standard_normal((1, B, 3)) + 0.02 * np.cumsum(rng.standard_normal((T, B, 3)), 0)
    if contact:
        # object close to the body surface so the correction path is exercised
        off = rng.standard_normal((1, B, 3))
        off = off / np.linalg.norm(off, axis=-1, keepdims=True) * np.array([0.45, 0.6, 0.4])
        obj_trans = body_trans + off + 0.01 * np.cumsum(rng.standard_normal((T, B, 3)), 0)
    else:
        obj_trans = 0.4 * rng.standard_normal((1, B, 3)) + 0.01 * np.cumsum(rng.standard_normal((T, B, 3)), 0)
    gt = np.concatenate([
        _aa_to_rot6d(body_aa).reshape(T, B, 132), body_trans,
        _aa_to_rot6d(obj_aa).reshape(T, B, 6), obj_trans], axis=2)  # (T,B,144)
    gt = np.ascontiguousarray(gt.transpose(1, 2, 0)[:, None])  # (B,1,144,T)
    mask = np.zeros(gt.shape, dtype=bool)
    mask[..., :past_len] = True
    cond = rng.standard_normal((10, B, 256))
    hand = 0.1 * rng.standard_normal((1, B, 90)) * np.ones((T, 1, 1))
    betas = rng.standard_normal((1, B, 10)) * np.ones((T, 1, 1))
    pts = np.stack([make_box_points(NUM_OBJ_POINTS, 0.2, seed + b) for b in range(B)])
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(gt=f32(gt), mask=mask, cond=f32(cond), hand_pose=f32(hand), betas=f32(betas),
                obj_points=f32(pts), past_len=past_len)


def make_skeleton_batch(B=2, T=15, past_len=10, seed=233):
    """Config 1: skeleton diffusion inputs. x channels = 63 + 36 + 7 = 106."""
    rng = _rng(seed, "skel%d_%d" % (B, T))
    body = 0.5 * rng.standard_normal((T, B, 21, 3))
    zero_pose_obj = 0.3 * rng.standard_normal((B, 12, 3))
    trans = 0.5 * rng.standard_normal((T, B, 3))
    quat = rng.standard_normal((T, B, 4))
    quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    pose = np.concatenate([trans, quat], axis=-1)  # [trans, quat xyzw]
    obj = 0.5 * rng.standard_normal((T, B, 12, 3))
    gt = np.concatenate([body.reshape(T, B, 63), obj.reshape(T, B, 36), pose], axis=2)
    gt = np.ascontiguousarray(gt.transpose(1, 2, 0)[:, None])
    mask = np.zeros(gt.shape, dtype=bool)
    mask[..., :past_len] = True
    cond = rng.standard_normal((10, B, 256))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(gt=f32(gt), mask=mask, cond=f32(cond), zero_pose_obj=f32(zero_pose_obj), past_len=past_len)


def noise_tape(shape, n_steps, seed=233):
    """Identical noise for the CPU oracle and the GPU path ("identical noise seeds" must mean
    an identical tape: CPU and CUDA torch generators differ).  Entry 0 is x_T, entries 1..n
    are the per-step eps drawn at t = n-1 .. 0 (reference gaussian_diffusion.py:532 draws one
    every step, including t = 0)."""
    rng = _rng(seed, "tape")
    return rng.standard_normal((n_steps + 1,) + tuple(shape)).astype(np.float32)


def fill_state_dict(shapes, seed=233):
    """Deterministic random-init weights for a {name: shape} map using the reference's
    parameter names.  Scales mimic torch's default inits so activations stay O(1)."""
    out = {}
    for name, shape in shapes.items():
        rng = _rng(seed, name)
        shape = tuple(shape)
        leaf = name.split(".")[-1]
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros(shape, dtype=np.int64)
        elif leaf == "running_var":
            out[name] = (0.5 + rng.random(shape)).astype(np.float32)
        elif leaf == "running_mean":
            out[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith("prelu.weight"):
            out[name] = np.full(shape, 0.25, dtype=np.float32)
        elif leaf == "weight" and len(shape) == 1:  # LayerNorm / BatchNorm scale
            out[name] = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif leaf in ("bias", "in_proj_bias"):
            out[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        elif leaf == "queries":
            out[name] = (rng.standard_normal(shape) / np.sqrt(shape[-1])).astype(np.float32)
        elif leaf == "wk":
            out[name] = (rng.standard_normal(shape) / np.sqrt(shape[0])).astype(np.float32)
        elif leaf == "inv_freq":
            d = shape[0] * 2
            out[name] = (1.0 / (10000 ** (np.arange(0, d, 2, dtype=np.float32) / d))).astype(np.float32)
        elif leaf == "pe":
            out[name] = sinusoid_table(shape[0], shape[-1]).reshape(shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            bound = 1.0 / np.sqrt(max(fan_in, 1))
            out[name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    return out


def sinusoid_table(max_len, d_model):
    """pe[p, 2i] = sin(p * w_i), pe[p, 2i+1] = cos(p * w_i), w_i = exp(-2i ln(1e4)/d)
    (reference model/layers.py:14-19), evaluated in float32 like the reference."""
    import torch  # same float32 arithmetic (exp/sin/cos) as the reference buffer
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.numpy()
