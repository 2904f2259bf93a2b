"""Host-side mirrors of reference model/layers.py: PositionalEncoding (:9-26), TimestepEmbedder
(:29-43), PointNet2Encoder (:111-175, parameters only), TransformerEncoder / TransformerDecoder
(:177-264, containers) and ST_GCNN_layer (:271-345, parameters only)."""
import numpy as np
import torch
import torch.nn as nn

from .sublayers import ConvSpatialTemporalGraphical, ConvTemporalGraphical


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
        self.register_buffer("pe", pe.unsqueeze(1))


class TimestepEmbedder(nn.Module):
    def __init__(self, latent_dim, sequence_pos_encoder):
        super().__init__()
        self.latent_dim = latent_dim
        self.sequence_pos_encoder = sequence_pos_encoder
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.SiLU(), nn.Linear(latent_dim, latent_dim))


class _SAModuleParams(nn.Module):
    """pointnet2_ops PointnetSAModuleMSG parameter layout: mlps.{k}.{0,1,3,4,6,7}.* with +3 input
    channels (use_xyz)."""

    def __init__(self, mlps):
        super().__init__()
        self.mlps = nn.ModuleList()
        for spec in mlps:
            spec = [spec[0] + 3] + list(spec[1:])
            layers = []
            for i in range(len(spec) - 1):
                layers += [nn.Conv2d(spec[i], spec[i + 1], 1, bias=False), nn.BatchNorm2d(spec[i + 1]), nn.ReLU(True)]
            self.mlps.append(nn.Sequential(*layers))


class PointNet2Encoder(nn.Module):
    def __init__(self, c_in=6, c_out=128, num_keypoints=256):
        super().__init__()
        self.SA_modules = nn.ModuleList([_SAModuleParams([[c_in, 16, 16, 32], [c_in, 32, 32, 64]]),
                                         _SAModuleParams([[96, 64, 64, 128], [96, 64, 96, 128]])])
        self.num_keypoints, self.c_out = num_keypoints, c_out
        self.Linear = nn.Linear(256, c_out - 3)


class TransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, norm=None):
        super().__init__()
        self.layers, self.norm = encoder_layer, norm


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, norm=None):
        super().__init__()
        self.layers, self.norm = decoder_layer, norm


class ST_GCNN_layer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, time_dim, joints_dim, dropout, version):
        super().__init__()
        assert version in (0, 2) and tuple(kernel_size) == (1, 1) and stride == 1
        self.gcn = ConvTemporalGraphical(time_dim, joints_dim) if version == 0 else ConvSpatialTemporalGraphical(time_dim, joints_dim)
        self.tcn = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), nn.BatchNorm2d(out_channels), nn.Dropout(dropout, inplace=True))
        if in_channels != out_channels:
            self.residual = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        else:
            self.residual = nn.Identity()
        self.prelu = nn.PReLU()
