"""Drop-in for BilinearFusion of /root/reference/models/fusion.py:6-63 (same constructor, forward, state_dict keys;
gates on, use_bilinear=1).  The two BatchNorm1d layers are sml_b200 SyncBatchNorm1d: on one rank they are plain
batch norm, under torch.distributed their (sum, sum^2) statistics are all-reduced.  CUDA only."""
import math

import torch
from torch import nn

from .. import functional as Fn
from ..utils.sync_batchnorm import SyncBatchNorm1d


def _init_max_weights(module):
    for m in module.modules():
        if type(m) == nn.Linear:
            stdv = 1. / math.sqrt(m.weight.size(1))
            m.weight.data.normal_(0, stdv)
            m.bias.data.zero_()


def _bilinear(x1, x2, w, b):
    """nn.Bilinear: y[b,o] = x1[b]^T W[o] x2[b] + bias[o]  as two contractions."""
    B, I = x1.shape
    O, _, J = w.shape
    t = Fn.matmul(x1, w.permute(1, 0, 2).reshape(I, O * J)).view(B, O, J)          # [B,O,J]
    y = Fn.matmul(t, x2.view(B, J, 1)).view(B, O)
    return Fn.add_rowvec(y.view(1, B, O), b.view(1, O)).view(B, O)


class BilinearFusion(nn.Module):
    def __init__(self, skip=1, use_bilinear=1, gate1=1, gate2=1, dim1=32, dim2=32, scale_dim1=1, scale_dim2=1, mmhid=64,
                 dropout_rate=0.25):
        super().__init__()
        if not use_bilinear:
            raise NotImplementedError("sml_b200 BilinearFusion: use_bilinear=0 is not the reference configuration")
        self.skip, self.use_bilinear, self.gate1, self.gate2 = skip, use_bilinear, gate1, gate2
        self.relu = nn.ReLU(inplace=False)
        dim1_og, dim2_og, dim1, dim2 = dim1, dim2, dim1 // scale_dim1, dim2 // scale_dim2
        skip_dim = dim1 + dim2 + 2 if skip else 0
        self.linear_h1 = nn.Sequential(nn.Linear(dim1_og, dim1), nn.ReLU())
        self.linear_z1 = nn.Bilinear(dim1_og, dim2_og, dim1)
        self.linear_o1 = nn.Sequential(nn.Linear(dim1, dim1), nn.ReLU(), nn.Dropout(p=dropout_rate))
        self.linear_h2 = nn.Sequential(nn.Linear(dim2_og, dim2), nn.ReLU())
        self.linear_z2 = nn.Bilinear(dim1_og, dim2_og, dim2)
        self.linear_o2 = nn.Sequential(nn.Linear(dim2, dim2), nn.ReLU(), nn.Dropout(p=dropout_rate))
        self.post_fusion_dropout = nn.Dropout(p=dropout_rate)
        self.encoder1 = nn.Sequential(nn.Linear((dim1 + 1) * (dim2 + 1), mmhid), SyncBatchNorm1d(mmhid), nn.ReLU(),
                                      nn.Dropout(p=dropout_rate))
        self.encoder2 = nn.Sequential(nn.Linear(mmhid + skip_dim, mmhid), SyncBatchNorm1d(mmhid), nn.ReLU(),
                                      nn.Dropout(p=dropout_rate))
        _init_max_weights(self)

    def _drop(self, x, p):
        return Fn.dropout_add(None, x, p, self.training)

    def forward(self, vec1, vec2):
        v1, v2 = Fn.unary(vec1.float(), "relu"), Fn.unary(vec2.float(), "relu")
        p = self.post_fusion_dropout.p
        if self.gate1:
            h1 = Fn.linear(v1, self.linear_h1[0].weight, self.linear_h1[0].bias, relu=True)
            z1 = _bilinear(v1, v2, self.linear_z1.weight, self.linear_z1.bias)
            o1 = self._drop(Fn.linear(Fn.mul(Fn.unary(z1, "sigmoid"), h1), self.linear_o1[0].weight, self.linear_o1[0].bias, relu=True), p)
        else:
            o1 = self._drop(Fn.linear(v1, self.linear_o1[0].weight, self.linear_o1[0].bias, relu=True), p)
        if self.gate2:
            h2 = Fn.linear(v2, self.linear_h2[0].weight, self.linear_h2[0].bias, relu=True)
            z2 = _bilinear(v1, v2, self.linear_z2.weight, self.linear_z2.bias)
            o2 = self._drop(Fn.linear(Fn.mul(Fn.unary(z2, "sigmoid"), h2), self.linear_o2[0].weight, self.linear_o2[0].bias, relu=True), p)
        else:
            o2 = self._drop(Fn.linear(v2, self.linear_o2[0].weight, self.linear_o2[0].bias, relu=True), p)
        B = o1.shape[0]
        one = torch.ones(B, 1, device=o1.device, dtype=o1.dtype)
        o1, o2 = torch.cat((o1, one), 1), torch.cat((o2, one), 1)                   # :56-57
        o12 = Fn.matmul(o1.unsqueeze(2), o2.unsqueeze(1)).flatten(start_dim=1)      # :58
        out = self._drop(o12, p)
        out = Fn.linear(out, self.encoder1[0].weight, self.encoder1[0].bias)
        out = self._drop(Fn.unary(self.encoder1[1](out), "relu"), p)
        if self.skip:
            out = torch.cat((out, o1, o2), 1)
        out = Fn.linear(out, self.encoder2[0].weight, self.encoder2[0].bias)
        return self._drop(Fn.unary(self.encoder2[1](out), "relu"), p)
