"""Drop-ins for the CMTA / MCAT models of /root/reference/models/model.py:559-853 and the encoder blocks of
/root/reference/models/cmta_utils.py:858-948 (TransLayer, PPEG, Transformer_P, Transformer_G): same constructor arguments,
``forward(**kwargs)`` keywords, return tuples and state_dict keys, so reference checkpoints load.  CUDA only.

Hot operators (all on libsml_b200): the 1024 -> 256 patch projection, the Nystrom TransLayers (dim 256, 8 heads x 32,
128 landmarks), PPEG, LayerNorm, the two genomic x WSI cross attentions (MultiheadAttention, raw scores) and the heads.
The MCAT path keeps torch's stock ``nn.TransformerEncoder`` / gated attention head on its <= 6 genomic-group tokens
(mcat_utils is out-of-scope glue on per-slide vectors); everything that touches the N patch tokens is ours.
"""
import math

import torch
from torch import nn

from .. import functional as Fn
from .MultiheadAttention import MultiheadAttention
from .mil import PPEG, TransLayer  # noqa: F401  (cmta_utils.py:858-891 are verbatim copies of mil.py:171-206)
from .fusion import BilinearFusion


def SNN_Block(dim1, dim2, dropout=0.25):
    """models/mcat_utils.py:80-94 / cmta_utils.py:112-121 (parameter container; evaluated by ``_snn``)"""
    return nn.Sequential(nn.Linear(dim1, dim2), nn.ELU(), nn.AlphaDropout(p=dropout, inplace=False))


def _snn(seq, x, training):
    for blk in seq:
        x = Fn.unary(Fn.linear(x, blk[0].weight, blk[0].bias), "elu")
        x = Fn.alpha_dropout(x, blk[2].p, training)
    return x


def _set_precision(module, p):
    for m in module.modules():
        if hasattr(m, "precision"):
            m.precision = p


class Transformer_P(nn.Module):
    """cmta_utils.py:894-926: pad to a square, cls token, TransLayer, PPEG, TransLayer, LayerNorm -> (cls, tokens)."""

    def __init__(self, feature_dim=512):
        super().__init__()
        self.pos_layer = PPEG(dim=feature_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, feature_dim))
        nn.init.normal_(self.cls_token, std=1e-6)
        self.layer1 = TransLayer(dim=feature_dim)
        self.layer2 = TransLayer(dim=feature_dim)
        self.norm = nn.LayerNorm(feature_dim)

    def forward(self, features):
        n = features.shape[1]
        side = int(math.ceil(math.sqrt(n)))
        add = side * side - n
        B = features.shape[0]
        h = torch.cat([self.cls_token.expand(B, -1, -1), features.float(), features[:, :add].float()], dim=1)   # :908-914
        h = self.layer1(h)
        h = self.pos_layer(h, side, side)
        h = self.layer2(h)
        h = Fn.layer_norm(h, self.norm.weight, self.norm.bias, self.norm.eps)
        return h[:, 0], h[:, 1:]


class Transformer_G(nn.Module):
    """cmta_utils.py:929-948: cls token, two TransLayers, LayerNorm on the genomic-group tokens."""

    def __init__(self, feature_dim=512):
        super().__init__()
        self.cls_token = nn.Parameter(torch.randn(1, 1, feature_dim))
        nn.init.normal_(self.cls_token, std=1e-6)
        self.layer1 = TransLayer(dim=feature_dim)
        self.layer2 = TransLayer(dim=feature_dim)
        self.norm = nn.LayerNorm(feature_dim)

    def forward(self, features):
        h = torch.cat((self.cls_token.expand(features.shape[0], -1, -1), features.float()), dim=1)
        h = self.layer1(h)
        h = self.layer2(h)
        h = Fn.layer_norm(h, self.norm.weight, self.norm.bias, self.norm.eps)
        return h[:, 0], h[:, 1:]


class _OmicFrontEnd(nn.Module):
    """shared constructor part of MCAT_Surv / CMTA (model.py:560-585, 715-741)"""

    def _build_front(self, args, omic_sizes, model_size_wsi, model_size_omic):
        self.args = args
        self.omic_sizes = omic_sizes
        self.n_classes = args.label_dim
        self.size_dict_WSI = {"small": [1024, 256, 256], "big": [1024, 512, 384]}
        self.size_dict_omic = {'small': [256, 256], 'big': [1024, 1024, 1024, 256]}
        size = self.size_dict_WSI[model_size_wsi]
        self.wsi_net = nn.Sequential(nn.Linear(size[0], size[1]), nn.ReLU(), nn.Dropout(0.25))
        hidden = self.size_dict_omic[model_size_omic]
        nets = []
        for input_dim in omic_sizes:
            fc = [SNN_Block(dim1=input_dim, dim2=hidden[0])]
            for i, _ in enumerate(hidden[1:]):
                fc.append(SNN_Block(dim1=hidden[i], dim2=hidden[i + 1], dropout=0.25))
            nets.append(nn.Sequential(*fc))
        self.sig_networks = nn.ModuleList(nets)
        self.precision = None
        return size, hidden

    def set_precision(self, p):
        """arithmetic of the token-level contractions: "fp32" (SIMT), "bf16x3" (tcgen05, fp32-grade) or "bf16" """
        _set_precision(self, p)

    def _front(self, kwargs):
        x_path, x_omic_all = kwargs['x_path'], kwargs['x_omic']
        sizes = self.omic_sizes
        x_omic = [x_omic_all[:, sum(sizes[:i]):sum(sizes[:i + 1])].float() for i in range(len(sizes))]
        mode = self.precision or "fp32"
        h_path = Fn.linear(x_path.float(), self.wsi_net[0].weight, self.wsi_net[0].bias, relu=True, mode=mode)
        h_path = Fn.dropout_add(None, h_path, self.wsi_net[2].p, self.training)
        h_omic = [_snn(self.sig_networks[i], x, self.training) for i, x in enumerate(x_omic)]
        return h_path, h_omic

    def _mm(self, a, b):
        if self.fusion == 'concat':
            h = torch.cat([a, b], dim=1)
            h = Fn.linear(h, self.mm[0].weight, self.mm[0].bias, relu=True)
            return Fn.linear(h, self.mm[2].weight, self.mm[2].bias, relu=True)
        if self.fusion == 'bilinear':
            return self.mm(a, b)
        raise NotImplementedError("Fusion [{}] is not implemented".format(self.fusion))

    def _make_mm(self, size):
        if self.fusion == 'concat':
            self.mm = nn.Sequential(nn.Linear(256 * 2, size[2]), nn.ReLU(), nn.Linear(size[2], size[2]), nn.ReLU())
        elif self.fusion == 'bilinear':
            self.mm = BilinearFusion(dim1=256, dim2=256, scale_dim1=8, scale_dim2=8, mmhid=256)
        else:
            self.mm = None

    @staticmethod
    def _survival(logits):
        hazards = Fn.unary(logits, "sigmoid")
        S = torch.cumprod(1 - hazards, dim=1)
        return hazards, S


class CMTA(_OmicFrontEnd):
    """models/model.py:714-853."""

    def __init__(self, args, fusion='concat', omic_sizes=[100, 100, 100, 131], n_classes=4, model_size_wsi: str = 'small',
                 model_size_omic: str = 'small', dropout=0.25):
        super().__init__()
        self.fusion = fusion
        size, hidden = self._build_front(args, omic_sizes, model_size_wsi, model_size_omic)
        self.pathomics_encoder = Transformer_P(feature_dim=hidden[-1])
        self.pathomics_decoder = Transformer_P(feature_dim=hidden[-1])
        self.P_in_G_Att = MultiheadAttention(embed_dim=256, num_heads=1)
        self.G_in_P_Att = MultiheadAttention(embed_dim=256, num_heads=1)
        self.genomics_encoder = Transformer_G(feature_dim=hidden[-1])
        self.genomics_decoder = Transformer_G(feature_dim=hidden[-1])
        self._make_mm(size)
        self.classifier = nn.Linear(size[2], self.n_classes)

    def forward(self, **kwargs):
        h_path, h_omic = self._front(kwargs)
        genomics_features = torch.stack(h_omic).transpose(0, 1)                          # [B, G, 256]  :791
        cls_p_enc, tok_p_enc = self.pathomics_encoder(h_path)                            # :798-802
        cls_g_enc, tok_g_enc = self.genomics_encoder(genomics_features)
        p_in_g, _ = self.P_in_G_Att(tok_p_enc.transpose(1, 0), tok_g_enc.transpose(1, 0), tok_g_enc.transpose(1, 0))   # :805-809
        g_in_p, _ = self.G_in_P_Att(tok_g_enc.transpose(1, 0), tok_p_enc.transpose(1, 0), tok_p_enc.transpose(1, 0))   # :810-814
        cls_p_dec, _ = self.pathomics_decoder(p_in_g.transpose(1, 0))                    # :817-821
        cls_g_dec, _ = self.genomics_decoder(g_in_p.transpose(1, 0))
        fusion = self._mm(Fn.scale(cls_p_enc + cls_p_dec, 0.5), Fn.scale(cls_g_enc + cls_g_dec, 0.5))   # :824-841
        logits = Fn.linear(fusion, self.classifier.weight, self.classifier.bias)
        hazards, S = self._survival(logits)
        return logits, hazards, S, cls_p_enc, cls_p_dec, cls_g_enc, cls_g_dec


class Attn_Net_Gated(nn.Module):
    """models/mcat_utils.py:114-143 (runs on the <= 6 genomic-group tokens; torch glue)"""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super().__init__()
        a = [nn.Linear(L, D), nn.Tanh()]
        b = [nn.Linear(L, D), nn.Sigmoid()]
        if dropout:
            a.append(nn.Dropout(0.25))
            b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*a)
        self.attention_b = nn.Sequential(*b)
        self.attention_c = nn.Linear(D, n_classes)

    def forward(self, x):
        return self.attention_c(self.attention_a(x).mul(self.attention_b(x))), x


class MCAT_Surv(_OmicFrontEnd):
    """models/model.py:559-666: patch projection + genomic SNNs + genomic x WSI co-attention (G queries x N keys, raw scores)
    on libsml_b200; the two 2-layer nn.TransformerEncoder stacks and gated attention heads see G <= 6 tokens per slide."""

    def __init__(self, args, fusion='concat', omic_sizes=[100, 100, 100, 131], n_classes=4, model_size_wsi: str = 'small',
                 model_size_omic: str = 'small', dropout=0.25):
        super().__init__()
        self.fusion = fusion
        size, _ = self._build_front(args, omic_sizes, model_size_wsi, model_size_omic)
        self.coattn = MultiheadAttention(embed_dim=256, num_heads=1)
        enc = nn.TransformerEncoderLayer(d_model=256, nhead=8, dim_feedforward=512, dropout=dropout, activation='relu')
        self.path_transformer = nn.TransformerEncoder(enc, num_layers=2)
        self.path_attention_head = Attn_Net_Gated(L=size[2], D=size[2], dropout=dropout, n_classes=1)
        self.path_rho = nn.Sequential(nn.Linear(size[2], size[2]), nn.ReLU(), nn.Dropout(dropout))
        enc = nn.TransformerEncoderLayer(d_model=256, nhead=8, dim_feedforward=512, dropout=dropout, activation='relu')
        self.omic_transformer = nn.TransformerEncoder(enc, num_layers=2)
        self.omic_attention_head = Attn_Net_Gated(L=size[2], D=size[2], dropout=dropout, n_classes=1)
        self.omic_rho = nn.Sequential(nn.Linear(size[2], size[2]), nn.ReLU(), nn.Dropout(dropout))
        self._make_mm(size)
        self.classifier = nn.Linear(size[2], self.n_classes)

    def _pool(self, trans, head, rho, bag):
        h = trans(bag)                                                                   # [G, B, 256]
        A, h = head(h.squeeze(1))
        A = torch.transpose(A.transpose(0, 2), 1, 0)                                     # [B, 1, G]  :635
        h = torch.bmm(torch.softmax(A, dim=2).float(), h.transpose(0, 1).float())        # :640
        return rho(h).squeeze(), A

    def forward(self, **kwargs):
        h_path, h_omic = self._front(kwargs)
        h_path_bag = h_path.transpose(0, 1)                                              # [N, B, 256]  :620
        h_omic_bag = torch.stack(h_omic)                                                 # [G, B, 256]  :622
        h_path_coattn, A_coattn = self.coattn(h_omic_bag, h_path_bag, h_path_bag)        # :627
        h_p, _ = self._pool(self.path_transformer, self.path_attention_head, self.path_rho, h_path_coattn)
        h_o, _ = self._pool(self.omic_transformer, self.omic_attention_head, self.omic_rho, h_omic_bag)
        h = self._mm(h_p, h_o) if self.fusion == 'concat' else self.mm(h_p.unsqueeze(0), h_o.unsqueeze(0)).squeeze()
        logits = Fn.linear(h, self.classifier.weight, self.classifier.bias)
        hazards, S = self._survival(logits)
        return logits, hazards, S
