"""Oracle (test infrastructure only): the CMTA model and its encoder blocks.

Restates /root/reference/models/cmta_utils.py:894-948 (Transformer_P / Transformer_G; TransLayer and PPEG at :858-891 are
verbatim copies of models/mil.py:171-206, restated in oracle/nystrom.py) and /root/reference/models/model.py:784-853
(CMTA.forward: patch projection, genomic SNNs, two encoders, the two genomic x WSI cross attentions with raw scores, two
decoders, concat fusion, classifier) in plain torch ops, eval mode (all dropouts off).  Parameters are a dict keyed by the
reference ``state_dict`` names.  Pinned by tests/test_oracle_golden.py:test_cmta against tests/golden/cmta.pt, which
tests/golden/make_golden_cmta.py generates from the imported, unmodified reference.
"""
import math

import torch
import torch.nn.functional as F

from .heads import mha_coattn
from .nystrom import layer_norm, ppeg, trans_layer


def transformer_p(x, p, prefix, dim=256):
    """cmta_utils.py:894-926 -> (cls [B,dim], tokens [B,side^2,dim])"""
    n = x.shape[1]
    side = int(math.ceil(math.sqrt(n)))
    add = side * side - n
    h = torch.cat([x, x[:, :add]], dim=1)                                        # :908-910
    h = torch.cat([p[prefix + "cls_token"].expand(h.shape[0], -1, -1), h], dim=1)  # :912-914
    h = trans_layer(h, p, prefix + "layer1.", dim)                               # :916
    h = ppeg(h, p, prefix + "pos_layer.", side, side)                            # :918
    h = trans_layer(h, p, prefix + "layer2.", dim)                               # :920
    h = layer_norm(h, p[prefix + "norm.weight"], p[prefix + "norm.bias"])        # :922
    return h[:, 0], h[:, 1:]


def transformer_g(x, p, prefix, dim=256):
    """cmta_utils.py:929-948"""
    h = torch.cat([p[prefix + "cls_token"].expand(x.shape[0], -1, -1), x], dim=1)
    h = trans_layer(h, p, prefix + "layer1.", dim)
    h = trans_layer(h, p, prefix + "layer2.", dim)
    h = layer_norm(h, p[prefix + "norm.weight"], p[prefix + "norm.bias"])
    return h[:, 0], h[:, 1:]


def snn(x, p, prefix, depth=2):
    """SNN_Block chain: Linear + ELU (+ AlphaDropout, off in eval) (mcat_utils.py:80-94)"""
    for i in range(depth):
        x = F.elu(x @ p["%s%d.0.weight" % (prefix, i)].t() + p["%s%d.0.bias" % (prefix, i)])
    return x


def cmta(x_path, x_omic, p, omic_sizes=(100, 100, 100, 131)):
    """CMTA.forward (model.py:784-853), fusion='concat'.  Returns the reference's 7-tuple."""
    xs, o = [], 0
    for s in omic_sizes:
        xs.append(x_omic[:, o:o + s]); o += s
    h_path = torch.relu(x_path @ p["wsi_net.0.weight"].t() + p["wsi_net.0.bias"])          # :790
    h_omic = torch.stack([snn(x, p, "sig_networks.%d." % i) for i, x in enumerate(xs)]).transpose(0, 1)   # :791-792
    cls_pe, tok_pe = transformer_p(h_path, p, "pathomics_encoder.")                        # :798-802
    cls_ge, tok_ge = transformer_g(h_omic, p, "genomics_encoder.")
    pg, _ = mha_coattn(tok_pe.transpose(1, 0), tok_ge.transpose(1, 0), tok_ge.transpose(1, 0), p, 1, "P_in_G_Att.")   # :805
    gp, _ = mha_coattn(tok_ge.transpose(1, 0), tok_pe.transpose(1, 0), tok_pe.transpose(1, 0), p, 1, "G_in_P_Att.")   # :810
    cls_pd, _ = transformer_p(pg.transpose(1, 0), p, "pathomics_decoder.")                 # :817-821
    cls_gd, _ = transformer_g(gp.transpose(1, 0), p, "genomics_decoder.")
    f = torch.cat([(cls_pe + cls_pd) / 2, (cls_ge + cls_gd) / 2], dim=1)                   # :824-833
    f = torch.relu(f @ p["mm.0.weight"].t() + p["mm.0.bias"])
    f = torch.relu(f @ p["mm.2.weight"].t() + p["mm.2.bias"])
    logits = f @ p["classifier.weight"].t() + p["classifier.bias"]                         # :845
    hazards = torch.sigmoid(logits)
    S = torch.cumprod(1 - hazards, dim=1)
    return logits, hazards, S, cls_pe, cls_pd, cls_ge, cls_gd


def mcat_front(x_path, x_omic, p, omic_sizes):
    """the part of MCAT_Surv.forward that touches the N patch tokens (model.py:612-627): patch projection, genomic SNNs
    and the genomic x WSI co-attention.  Returns (h_path_coattn [G,B,256], raw scores [B,1,G,N])."""
    xs, o = [], 0
    for s in omic_sizes:
        xs.append(x_omic[:, o:o + s]); o += s
    h_path_bag = torch.relu(x_path @ p["wsi_net.0.weight"].t() + p["wsi_net.0.bias"]).transpose(0, 1)
    h_omic_bag = torch.stack([snn(x, p, "sig_networks.%d." % i) for i, x in enumerate(xs)])
    return mha_coattn(h_omic_bag, h_path_bag, h_path_bag, p, 1, "coattn.")
