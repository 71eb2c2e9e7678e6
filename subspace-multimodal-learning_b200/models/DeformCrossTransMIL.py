"""Drop-in for /root/reference/models/DeformCrossTransMIL.py:28-202 (FusionNet, DeformCrossTransLayer,
DeformCrossTransMIL, Pooler; same attribute names / state_dict keys / forward signature).  CUDA only.

Schedule differences (results match within tests' tolerances): the 2500x-repeated omic vector is folded into
a rank-1 row bias of the fusion GEMM (:104,:110); bags of any square N are accepted (reference: 2500)."""
import torch
from torch import nn

from .. import functional as Fn
from . import _deform_common as dc
from .DeformableAttention1D import DeformCrossAttention1D
from .DeformableAttention2D import DeformCrossAttention2D


class FusionNet(nn.Module):
    def __init__(self, feature_dim=128):
        super().__init__()
        self.fusion_layer = nn.Linear(feature_dim * 2, feature_dim)
        self.feature_dim = feature_dim

    def forward(self, gene_features, image_features):
        """cat(gene_features [B,N,f], image_features) @ W^T + b.  ``image_features`` may be the [B,f] base vector of a
        token-repeated tensor (how DeformCrossTransMIL calls it) — then only a rank-1 row bias is added."""
        f = self.feature_dim
        w, b = self.fusion_layer.weight, self.fusion_layer.bias
        h = Fn.linear(gene_features, w[:, :f].contiguous(), None, mode=dc.gemm_mode(getattr(self, "precision", None)))
        if image_features.dim() == 2:
            return Fn.add_rowvec(h, Fn.linear(image_features, w[:, f:].contiguous(), b))
        return h + Fn.linear(image_features, w[:, f:].contiguous(), b)


class DeformCrossTransLayer(nn.Module):
    def __init__(self, norm_layer=nn.LayerNorm, dim=128):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn2d = DeformCrossAttention2D(dim=128, dim_head=64, heads=8, dropout=0.1, downsample_factor=4, offset_scale=4,
                                             offset_groups=8, offset_kernel_size=6)
        self.attn1d = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)

    def forward(self, x1, x2, attn_dim, return_vgrid):
        n1 = Fn.layer_norm(x1, self.norm.weight, self.norm.bias, self.norm.eps).transpose(1, 2)
        n2 = Fn.layer_norm(x2, self.norm.weight, self.norm.bias, self.norm.eps).transpose(1, 2)
        if attn_dim == 1:
            x = self.attn1d(n1, n2)
            return Fn.dropout_add(x1, x.transpose(1, 2).contiguous(), 0.0, False)
        if not return_vgrid:
            x = self.attn2d(n1, n2, return_vgrid)
            return Fn.dropout_add(x1, x.transpose(1, 2).contiguous(), 0.0, False)
        x, vgrid = self.attn2d(n1, n2, return_vgrid)
        return Fn.dropout_add(x1, x.transpose(1, 2).contiguous(), 0.0, False), vgrid


class Pooler(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        avg = Fn.mean_tokens(hidden_states)
        return Fn.unary(Fn.linear(avg, self.dense.weight, self.dense.bias), "tanh")


class DeformCrossTransMIL(nn.Module):
    def __init__(self, args, n_classes=4):
        super().__init__()
        self.fusion_layer = FusionNet(feature_dim=128)
        self._fc1 = nn.Sequential(nn.Linear(1024, args.path_dim), nn.ReLU())
        self.cls_token = nn.Parameter(torch.randn(1, 1, args.path_dim))
        self.args = args
        self.n_classes = n_classes
        self.layer3 = DeformCrossTransLayer(dim=args.path_dim)
        self.norm = nn.LayerNorm(args.path_dim)
        self._fc2 = nn.Linear(args.path_dim, self.n_classes)
        self.pooler = Pooler(args.path_dim)
        self.multimodal_projection = nn.Linear(args.path_dim, self.args.path_dim)
        self.precision = None

    def set_precision(self, p):
        """arithmetic of the fused deformable attention kernels: "bf16x3" (default, fp32-grade), "bf16" or "fp32" (SIMT)"""
        self.precision = p
        self.fusion_layer.precision = p
        self.layer3.attn1d.precision = p
        self.layer3.attn2d.precision = p

    def forward(self, path, omic):
        mode = dc.gemm_mode(self.precision)                    # engine of the big token-level GEMMs (fc1, fusion)
        path = Fn.linear(path if path.dtype == torch.bfloat16 else path.float(), self._fc1[0].weight, self._fc1[0].bias, relu=True, mode=mode)      # :99-100
        omic = omic.float()
        B, N = path.shape[0], path.shape[1]
        h = self.fusion_layer(path, omic)                                                        # :104-110 (repeat folded)
        vgrid = None
        if self.args.attn_dim == 1:
            cls = self.cls_token.expand(B, -1, -1)
            h = torch.cat((cls, h), dim=1)
            path = torch.cat((cls, path), dim=1)
            h = self.layer3(h, path, 1, self.args.return_vgrid)
            h = Fn.layer_norm(h[:, 0].contiguous(), self.norm.weight, self.norm.bias, self.norm.eps)   # norm(h)[:,0]
        else:
            if self.args.return_vgrid:
                h, vgrid = self.layer3(h, path, 2, True)
            else:
                h = self.layer3(h, path, 2, False)
            h = self.pooler(Fn.layer_norm(h, self.norm.weight, self.norm.bias, self.norm.eps))
        logits = Fn.linear(h, self._fc2.weight, self._fc2.bias)
        encoded = Fn.linear(h, self.multimodal_projection.weight, self.multimodal_projection.bias)
        path_grads = None
        if self.args.return_vgrid:
            # the reference returns the token-repeated omic tensor (:104); keep that contract and tag the [B,128]
            # base so that sml_b200's BatchLoss can use the collapsed Gram
            # A stride-0 view instead of a 64 MB copy: sml_b200's BatchLoss recognises it (stride(1) == 0) and uses the [B,128]
            # base directly.  (A Python attribute would be lost when DistributedDataParallel re-wraps the outputs, and the
            # loss would fall back to gathering / multiplying the full [B, N*128] rows: +17 ms per call at 2 GPUs.)
            omic_rep = omic.unsqueeze(1).expand(B, N, omic.shape[-1])
            omic_rep._sml_base = omic
            return encoded, logits, path_grads, omic_rep, vgrid
        return encoded, logits, path_grads
