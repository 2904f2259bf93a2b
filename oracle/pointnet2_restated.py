"""Restatement of the pointnet2_ops 3.0.0 operators the reference's PointNet2Encoder uses
(erikwijmans/Pointnet2_PyTorch, pointnet2_ops_lib: _ext-src/src/sampling_gpu.cu, ball_query_gpu.cu,
group_points_gpu.cu; pointnet2_utils.py QueryAndGroup; pointnet2_modules.py PointnetSAModuleMSG).
TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED for these operators: the package is a go-to-PyPI dependency of the reference that is not
vendored under /root/reference and cannot be installed here, and the reference ships no golden vectors for
it, so this file follows the published algorithm of that pinned version (requirements.txt: pointnet2_ops 3.0.0):

furthest_point_sampling(xyz (B,N,3), m) -> idx (B,m) int
    one block of block_size = min(2^floor(log2 N), 512) threads per cloud; temp[k] = 1e10; idx[0] = 0;
    for j in 1..m-1: with old = idx[j-1], every thread t walks k = t, t+block_size, ...: points with
    x^2+y^2+z^2 <= 1e-3 are SKIPPED (neither updated nor eligible), d = |p_k - p_old|^2,
    temp[k] = min(temp[k], d), thread-best = the FIRST k of its walk with the largest temp (strict >,
    start best = -1, besti = 0); block reduction keeps the larger value and, on ties, the LOWER thread.
ball_query(radius, nsample, xyz (B,N,3), new_xyz (B,m,3)) -> idx (B,m,nsample) int
    for every centre scan k = 0..N-1 in index order while cnt < nsample; d2 < radius^2 (strict) selects k;
    the first hit also fills all nsample slots (padding); idx stays 0 when nothing is in range.
QueryAndGroup(use_xyz=True): grouped = cat([xyz[idx] - new_xyz, features[idx]], channel)  (xyz FIRST)
PointnetSAModuleMSG: new_xyz = xyz[fps idx]; per scale: group -> Conv2d 1x1 (no bias) + BatchNorm2d + ReLU (x3)
    -> max over the nsample axis; concatenate the scales' channels.
Distances are evaluated in float32 as (dx*dx + dy*dy) + dz*dz without FMA contraction (nvcc may contract
upstream; selections can then differ only for points whose distance ties to the last bit).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _d2(a, b):
    d = (a - b).astype(np.float32)
    return ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32) + \
        (d[..., 2] * d[..., 2]).astype(np.float32)


def furthest_point_sample(xyz, m):
    xyz = np.asarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    bs = max(min(1 << int(np.floor(np.log2(N))), 512), 1)
    out = np.zeros((B, m), dtype=np.int64)
    k_all = np.arange(N)
    owner = k_all % bs                      # thread that walks point k
    for b in range(B):
        p = xyz[b]
        mag = ((p[:, 0] * p[:, 0]).astype(np.float32) + (p[:, 1] * p[:, 1]).astype(np.float32)).astype(np.float32) + \
            (p[:, 2] * p[:, 2]).astype(np.float32)
        live = mag > np.float32(1e-3)
        temp = np.full(N, 1e10, dtype=np.float32)
        old = 0
        for j in range(1, m):
            d = _d2(p, p[old][None])
            temp = np.where(live, np.minimum(temp, d), temp).astype(np.float32)
            # thread-best: first k of the thread's walk with the largest value among live points (start: best=-1, besti=0)
            best_v = np.full(bs, -1.0, dtype=np.float32)
            best_i = np.zeros(bs, dtype=np.int64)
            cand = np.where(live)[0]
            if cand.size:
                v = temp[cand]
                order = np.lexsort((cand, -v.astype(np.float64), owner[cand]))   # by owner, then value desc, then k asc
                oc, vc, kc = owner[cand][order], v[order], cand[order]
                first = np.ones(len(order), dtype=bool)
                first[1:] = oc[1:] != oc[:-1]
                best_v[oc[first]] = vc[first]
                best_i[oc[first]] = kc[first]
            # block reduction: larger value wins, ties keep the lower thread
            t = int(np.lexsort((np.arange(bs), -best_v.astype(np.float64)))[0])
            old = int(best_i[t])
            out[b, j] = old
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = np.asarray(xyz, dtype=np.float32), np.asarray(new_xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    r2 = np.float32(radius) * np.float32(radius)
    idx = np.zeros((B, m, nsample), dtype=np.int64)
    for b in range(B):
        d2 = _d2(new_xyz[b][:, None, :], xyz[b][None, :, :])       # (m, N)
        hit = d2 < r2
        for j in range(m):
            ks = np.nonzero(hit[j])[0][:nsample]
            if ks.size:
                idx[b, j, :] = ks[0]
                idx[b, j, : ks.size] = ks
    return idx


def sa_module_msg(sd, prefix, xyz, features, npoint, radii, nsamples, eps=1e-5):
    """xyz (B,N,3) tensor, features (B,C,N) tensor or None -> new_xyz (B,npoint,3), new_features (B,sum C_out,npoint)."""
    B = xyz.shape[0]
    fps = torch.from_numpy(furthest_point_sample(xyz.numpy(), npoint))
    new_xyz = torch.gather(xyz, 1, fps[..., None].expand(-1, -1, 3))
    outs = []
    for s, (r, ns) in enumerate(zip(radii, nsamples)):
        idx = torch.from_numpy(ball_query(r, ns, xyz.numpy(), new_xyz.numpy()))             # (B,npoint,ns)
        flat = idx.reshape(B, -1)
        g_xyz = torch.gather(xyz, 1, flat[..., None].expand(-1, -1, 3)).view(B, npoint, ns, 3) - new_xyz[:, :, None, :]
        x = g_xyz.permute(0, 3, 1, 2)                                                          # (B,3,npoint,ns)
        if features is not None:
            g_f = torch.gather(features, 2, flat[:, None, :].expand(-1, features.shape[1], -1)).view(B, -1, npoint, ns)
            x = torch.cat([x, g_f], dim=1)
        for li in (0, 3, 6):
            p = "%smlps.%d." % (prefix, s)
            x = F.conv2d(x, sd[p + "%d.weight" % li])
            q = p + "%d." % (li + 1)
            x = F.batch_norm(x, sd[q + "running_mean"], sd[q + "running_var"], sd[q + "weight"], sd[q + "bias"], False, 0.0, eps)
            x = F.relu(x)
        outs.append(x.max(dim=3)[0])
    return new_xyz, torch.cat(outs, dim=1)


def pointnet2_encoder(sd, obj_points, prefix="pcEmbedding."):
    """PointNet2Encoder(c_in=1, c_out=256, num_keypoints=1) as MDM._get_embeddings calls it (reference
    model/layers.py:111-175, model/diffusion_smpl.py:210-211): obj_points (B,P,3) -> pc_embedding (B,256)."""
    pc = torch.cat([obj_points, obj_points.norm(dim=2, keepdim=True)], dim=2)                 # (B,P,4)
    xyz, feat = pc[..., :3].contiguous(), pc[..., 3:].transpose(1, 2).contiguous()
    xyz1, f1 = sa_module_msg(sd, prefix + "SA_modules.0.", xyz, feat, 1024, [0.05, 0.1], [16, 32])
    xyz2, f2 = sa_module_msg(sd, prefix + "SA_modules.1.", xyz1, f1, 1, [0.1, 0.2], [16, 32])
    lin = F.linear(f2.transpose(1, 2), sd[prefix + "Linear.weight"], sd[prefix + "Linear.bias"])   # (B,1,253)
    return torch.cat([xyz2, lin], dim=-1).reshape(obj_points.shape[0], -1)
