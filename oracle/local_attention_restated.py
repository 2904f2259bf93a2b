"""Restatement of lucidrains `local_attention.LocalAttention` for the one configuration the
reference uses.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The package is a third-party dependency that is NOT under /root/reference and is unpinned
(reference requirements.txt:28).  Constructor kwargs at the reference call sites
(interdiff/model/sublayers.py:79-88, 251-260):
    dim=256, window_size=1, causal=False, look_backward=1, look_forward=1,
    dropout=p, exact_windowsize=False, autopad=True
Call: self_attn(q, k, v, mask=ones(1, T))  with q,k,v of shape (B*N, T, 256)
(interdiff/model/sublayers.py:187, 350).

PARITY UNPINNED at this boundary: the rotary-embedding placement changed between releases.
  rotary='absolute'  (<= 1.5.x): q and k are rotated by their absolute sequence position
                     before bucketing  => relative offsets q_pos - k_pos = {+1, 0, -1}
                     for the (t-1, t, t+1) key slots.
  rotary='bucketed'  (>= 1.6):   rotation after look_around, key slot j in {0,1,2} gets
                     position j, the query gets the LAST position (2)
                     => offsets {2, 1, 0}.
The shipped checkpoint behaves better under 'absolute' (SURVEY.md section 8c), which is the
default everywhere in this repo; the product kernel takes the three offsets as a table.
"""
import os
import torch
import torch.nn as nn
import torch.nn.functional as F

DEFAULT_ROTARY = os.environ.get("INTERDIFF_ORACLE_ROTARY", "absolute")


def rotate_half(x):
    d = x.shape[-1] // 2
    x1, x2 = x[..., :d], x[..., d:]
    return torch.cat((-x2, x1), dim=-1)


def _look_around(x, backward=1, forward=1, pad_value=-1, dim=2):
    """x: (b, windows, window_size, ...) -> concatenates the previous / next windows along
    the in-window axis, padding out-of-range windows with pad_value."""
    t = x.shape[1]
    dims = (len(x.shape) - dim) * (0, 0)
    padded_x = F.pad(x, (*dims, backward, forward), value=pad_value)
    tensors = [padded_x[:, ind:(ind + t), ...] for ind in range(forward + backward + 1)]
    return torch.cat(tensors, dim=dim)


class SinusoidalEmbeddings(nn.Module):
    def __init__(self, dim):
        super().__init__()
        inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)

    def forward(self, n, device):
        t = torch.arange(n, device=device).type_as(self.inv_freq)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq)
        return torch.cat((freqs, freqs), dim=-1)


class LocalAttention(nn.Module):
    def __init__(self, window_size, causal=False, look_backward=1, look_forward=None,
                 dropout=0.0, dim=None, autopad=False, exact_windowsize=False, scale=None,
                 rotary=None, **_unused):
        super().__init__()
        look_forward = look_forward if look_forward is not None else (0 if causal else 1)
        assert not (causal and look_forward > 0)
        assert window_size == 1 and not causal, "restated for the reference configuration only"
        self.window_size = window_size
        self.look_backward = look_backward
        self.look_forward = look_forward
        self.autopad = autopad
        self.scale = scale
        self.dropout = nn.Dropout(dropout)
        self.rotary = rotary or DEFAULT_ROTARY
        self.rel_pos = SinusoidalEmbeddings(dim) if dim is not None else None

    def forward(self, q, k, v, mask=None, input_mask=None):
        b, n, dim_head = q.shape
        scale = self.scale if self.scale is not None else dim_head ** -0.5
        pad_value = -1

        if self.rel_pos is not None and self.rotary == "absolute":
            freqs = self.rel_pos(n, q.device)  # (n, d)
            q = q * freqs.cos() + rotate_half(q) * freqs.sin()
            k = k * freqs.cos() + rotate_half(k) * freqs.sin()

        windows = n // self.window_size
        bq = q.reshape(b, windows, self.window_size, dim_head)
        bk = k.reshape(b, windows, self.window_size, dim_head)
        bv = v.reshape(b, windows, self.window_size, dim_head)
        bq = bq * scale
        bk = _look_around(bk, self.look_backward, self.look_forward, pad_value)
        bv = _look_around(bv, self.look_backward, self.look_forward, pad_value)

        if self.rel_pos is not None and self.rotary == "bucketed":
            freqs = self.rel_pos(bk.shape[-2], q.device)  # (3, d)
            q_freqs = freqs[-bq.shape[-2]:]
            bq = bq * q_freqs.cos() + rotate_half(bq) * q_freqs.sin()
            bk = bk * freqs.cos() + rotate_half(bk) * freqs.sin()

        seq = torch.arange(n, device=q.device)
        b_t = seq.reshape(1, windows, self.window_size)
        bq_k = _look_around(b_t, self.look_backward, self.look_forward, pad_value)
        pad_mask = (bq_k == pad_value)[:, :, None, :]  # (1, w, 1, j)

        sim = torch.einsum("bhie,bhje->bhij", bq, bk)
        mask_value = -torch.finfo(sim.dtype).max
        sim = sim.masked_fill(pad_mask, mask_value)
        # the reference passes an all-True (1, T) key mask: it only re-masks the padded slots
        attn = sim.softmax(dim=-1)
        attn = self.dropout(attn)
        out = torch.einsum("bhij,bhje->bhie", attn, bv)
        return out.reshape(b, n, dim_head)
