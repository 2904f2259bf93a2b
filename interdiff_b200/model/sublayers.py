"""Parameter containers with the reference's attribute / state_dict names for the
"query-and-attend" layers (reference model/sublayers.py:37-203 encoder layer, :206-375 decoder
layer) and the graph-convolution parameters of the correction net (:378-516).

These modules only HOLD parameters (so strict checkpoint loading works unchanged); the arithmetic
runs in libinterdiff_b200.so through interdiff_b200.engine.Engine.
"""
import math

import torch
import torch.nn as nn


class _RelPos(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


class LocalAttentionParams(nn.Module):
    """Stands in for local_attention.LocalAttention(dim=d_model, window_size=1, look_backward=1,
    look_forward=1): its only state is the rotary `rel_pos.inv_freq` buffer."""

    def __init__(self, dim):
        super().__init__()
        self.rel_pos = _RelPos(dim)


def _qan_common(mod, d_model, dim_feedforward, num_queries):
    mod.self_attn = LocalAttentionParams(d_model)
    mod.linear1 = nn.Linear(d_model, dim_feedforward)
    mod.linear2 = nn.Linear(dim_feedforward, d_model)
    mod.queries = nn.Parameter(torch.randn(num_queries, d_model) / math.sqrt(d_model))
    mod.wk = nn.Parameter(torch.randn(num_queries, 1) / math.sqrt(num_queries))
    mod.d_model, mod.num_queries = d_model, num_queries


class TransformerEncoderLayerQaN(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, num_queries=10, window_size=1,
                 activation="relu", layer_norm_eps=1e-5, batch_first=False, norm_first=False, **_):
        super().__init__()
        _qan_common(self, d_model, dim_feedforward, num_queries)
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.nhead = nhead


class TransformerDecoderLayerQaN(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, num_queries=10, window_size=1,
                 activation="relu", layer_norm_eps=1e-5, batch_first=False, norm_first=False, **_):
        super().__init__()
        _qan_common(self, d_model, dim_feedforward, num_queries)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first)
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm3 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.nhead = nhead


class ConvTemporalGraphical(nn.Module):
    def __init__(self, time_dim, joints_dim):
        super().__init__()
        self.T = nn.Parameter(torch.empty(time_dim, time_dim).uniform_(-1, 1) / math.sqrt(time_dim))


class ConvSpatialTemporalGraphical(nn.Module):
    def __init__(self, time_dim, joints_dim):
        super().__init__()
        self.A = nn.Parameter(torch.empty(time_dim, joints_dim, joints_dim).uniform_(-1, 1) / math.sqrt(joints_dim))
        self.T = nn.Parameter(torch.empty(joints_dim, time_dim, time_dim).uniform_(-1, 1) / math.sqrt(time_dim))
