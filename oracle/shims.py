"""sys.modules shims that let the reference's OWN Python files import and run on CPU in the
authoring container.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference (/root/reference/interdiff) imports third-party packages that are absent here
and cannot be installed (no network).  Each shim provides exactly the names the hot-path
files touch:

  local_attention.LocalAttention              -> oracle.local_attention_restated
  pytorch3d.transforms.*                      -> oracle.transforms  (0.7.2 semantics)
  pointnet2_ops.pointnet2_modules.PointnetSAModuleMSG -> structural stub (same param names)
  chumpy / chumpy.ch / smplx / human_body_prior / pytorch3d.{loss,ops,structures} -> empty
  chamfer_distance.ChamferDistance            -> brute-force first-minimum squared-L2 argmin

Nothing here is used by the product path.
"""
import sys
import types

import torch
import torch.nn as nn

from . import transforms as _tf
from . import local_attention_restated as _la

_INSTALLED = False


class ChamferDistance(nn.Module):
    """idx1[n,i] = argmin_j |x[n,i]-y[n,j]|^2 (first minimum); idx2 likewise.  The reference
    only consumes the indices (interdiff/tools.py:45-53)."""

    def forward(self, x, y, x_normals=None, y_normals=None):
        idx1, idx2, d1, d2 = [], [], [], []
        for n in range(x.shape[0]):
            d = ((x[n][:, None, :] - y[n][None, :, :]) ** 2).sum(-1)
            m1, i1 = d.min(dim=1)
            m2, i2 = d.min(dim=0)
            idx1.append(i1.int()); idx2.append(i2.int()); d1.append(m1); d2.append(m2)
        return torch.stack(d1), torch.stack(d2), torch.stack(idx1), torch.stack(idx2)


class PointnetSAModuleMSG(nn.Module):
    """Parameter-name-compatible stand-in of pointnet2_ops 3.0.0's module (strict state_dict
    loading of checkpoints/diffusion.ckpt); forward = the restated FPS / ball-query / group ops."""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        self.npoint, self.radii, self.nsamples = npoint, radii, nsamples
        self.mlps = nn.ModuleList()
        for spec in mlps:
            spec = list(spec)
            if use_xyz:
                spec[0] += 3
            layers = []
            for i in range(len(spec) - 1):
                layers += [nn.Conv2d(spec[i], spec[i + 1], 1, bias=False), nn.BatchNorm2d(spec[i + 1]), nn.ReLU(True)]
            self.mlps.append(nn.Sequential(*layers))

    def forward(self, xyz, features):
        """pointnet2_ops semantics via oracle.pointnet2_restated (published algorithm; parity unpinned there)."""
        from . import pointnet2_restated as P2
        sd = {"mlps." + k: v for k, v in self.mlps.state_dict().items()}
        return P2.sa_module_msg(sd, "", xyz, features, self.npoint, self.radii, self.nsamples)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    global _INSTALLED
    if _INSTALLED:
        return
    _mod("local_attention", LocalAttention=_la.LocalAttention)
    p3d = _mod("pytorch3d")
    p3d.transforms = _mod(
        "pytorch3d.transforms",
        axis_angle_to_matrix=_tf.axis_angle_to_matrix,
        matrix_to_rotation_6d=_tf.matrix_to_rotation_6d,
        rotation_6d_to_matrix=_tf.rotation_6d_to_matrix,
        matrix_to_axis_angle=_tf.matrix_to_axis_angle,
        axis_angle_to_quaternion=_tf.axis_angle_to_quaternion,
        quaternion_to_matrix=_tf.quaternion_to_matrix,
        matrix_to_quaternion=_tf.matrix_to_quaternion,
        quaternion_to_axis_angle=_tf.quaternion_to_axis_angle,
    )
    p3d.loss = _mod("pytorch3d.loss")
    p3d.ops = _mod("pytorch3d.ops", cot_laplacian=None)
    p3d.structures = _mod("pytorch3d.structures", Meshes=object)
    pn = _mod("pointnet2_ops")
    pn.pointnet2_modules = _mod("pointnet2_ops.pointnet2_modules", PointnetSAModuleMSG=PointnetSAModuleMSG)
    ch = _mod("chumpy", Ch=type("Ch", (), {}), array=lambda *a, **k: None)
    ch.ch = _mod("chumpy.ch", MatVecMult=None)
    _mod("smplx")
    hbp = _mod("human_body_prior")
    hbp.tools = _mod("human_body_prior.tools", tgm_conversion=None)
    _mod("chamfer_distance", ChamferDistance=ChamferDistance)
    _INSTALLED = True
