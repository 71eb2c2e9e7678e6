"""Drop-ins for MaxNet and DeformPathomicNet of /root/reference/models/model.py:142-187, 440-544 (same constructor
arguments, ``forward(**kwargs)`` keywords, 7-tuple return and state_dict keys).  CUDA only.

DeformPathomicNet = two genomic encoders (MaxNet: 4 x Linear + ELU + AlphaDropout, ReLU) + two DeformCrossTransMIL branches
on the SAME bag + concat (or BilinearFusion) + three linear heads.  Everything runs on libsml_b200 kernels; the two
branches share nothing but the input bag (the reference reads x_path twice as well)."""
import math

import torch
from torch import nn
from torch.nn import Parameter

from .. import functional as Fn
from .DeformCrossTransMIL import DeformCrossTransMIL
from .fusion import BilinearFusion


def init_max_weights(module):
    """models/cmta_utils.py:128-139"""
    for m in module.modules():
        if type(m) == nn.Linear:
            stdv = 1.0 / math.sqrt(m.weight.size(1))
            m.weight.data.normal_(0, stdv)
            m.bias.data.zero_()


def define_bifusion(fusion_type, skip=1, use_bilinear=1, gate1=1, gate2=1, dim1=32, dim2=32, scale_dim1=1, scale_dim2=1,
                    mmhid=32, dropout_rate=0.25):
    """models/model.py:131-137"""
    if fusion_type == 'pofusion':
        return BilinearFusion(skip=skip, use_bilinear=use_bilinear, gate1=gate1, gate2=gate2, dim1=dim1, dim2=dim2,
                              scale_dim1=scale_dim1, scale_dim2=scale_dim2, mmhid=mmhid, dropout_rate=dropout_rate)
    raise NotImplementedError('fusion type [%s] is not found' % fusion_type)


class MaxNet(nn.Module):
    def __init__(self, input_dim=59, omic_dim=32, return_grad='False', dropout_rate=0.25, label_dim=1, init_max=True):
        super().__init__()
        hidden = [64, 48, 32, 32]
        self.return_grad = return_grad
        dims = [input_dim, hidden[0], hidden[1], hidden[2], omic_dim]
        self.encoder = nn.Sequential(*[nn.Sequential(nn.Linear(dims[i], dims[i + 1]), nn.ELU(),
                                                     nn.AlphaDropout(p=dropout_rate, inplace=False)) for i in range(4)])
        self.relu = nn.ReLU(inplace=False)
        self.classifier = nn.Sequential(nn.Linear(omic_dim, label_dim))
        if init_max:
            init_max_weights(self)
        self.output_range = Parameter(torch.FloatTensor([6]), requires_grad=False)
        self.output_shift = Parameter(torch.FloatTensor([-3]), requires_grad=False)

    def forward(self, **kwargs):
        x = kwargs['x_omic'].float()
        for blk in self.encoder:
            x = Fn.unary(Fn.linear(x, blk[0].weight, blk[0].bias), "elu")
            x = Fn.alpha_dropout(x, blk[2].p, self.training)
        features = Fn.unary(x, "relu")
        logits = Fn.linear(features, self.classifier[0].weight, self.classifier[0].bias)
        return features, logits, None


class DeformPathomicNet(nn.Module):
    def __init__(self, args):
        super().__init__()
        init_max = True if args.init_type == "max" else False
        self.args = args
        self.omic_net_tumor = MaxNet(input_dim=args.input_size_omic_tumor, omic_dim=args.omic_dim, return_grad=args.return_grad,
                                     dropout_rate=args.dropout_rate, label_dim=args.label_dim, init_max=init_max)
        self.omic_net_immune = MaxNet(input_dim=args.input_size_omic_immune, omic_dim=args.omic_dim, return_grad=args.return_grad,
                                      dropout_rate=args.dropout_rate, label_dim=args.label_dim, init_max=init_max)
        self.pathomic_net_tumor = DeformCrossTransMIL(args)
        self.pathomic_net_immune = DeformCrossTransMIL(args)
        self.bilinear_dim = 20
        if args.fusion_type != "concat":
            self.fusion = define_bifusion(fusion_type=args.fusion_type, skip=args.skip, use_bilinear=args.use_bilinear,
                                          gate1=args.path_gate, gate2=args.omic_gate, dim1=args.path_dim, dim2=args.omic_dim,
                                          scale_dim1=args.path_scale, scale_dim2=args.omic_scale, mmhid=args.mmhid,
                                          dropout_rate=args.dropout_rate)
            self.classifier = nn.Sequential(nn.Linear(args.mmhid, args.label_dim))
        else:
            self.classifier = nn.Linear(args.mmhid * 2, args.label_dim)
        # NB: the reference creates the per-branch heads only on the "concat" path (model.py:466-471) and uses them
        # unconditionally (:521-522); keep the attribute layout of the configuration the reference can actually run
        self.classifier_tumor = nn.Sequential(nn.Linear(args.mmhid, args.label_dim))
        self.classifier_immune = nn.Sequential(nn.Linear(args.mmhid, args.label_dim))
        self.return_grad = args.return_grad
        self.cut_fuse_grad = args.cut_fuse_grad
        self.fusion_type = args.fusion_type
        self.output_range = Parameter(torch.FloatTensor([6]), requires_grad=False)
        self.output_shift = Parameter(torch.FloatTensor([-3]), requires_grad=False)

    def set_precision(self, p):
        self.pathomic_net_tumor.set_precision(p)
        self.pathomic_net_immune.set_precision(p)

    def forward(self, **kwargs):
        x_path = kwargs['x_path']
        omic_vec_tumor, _, _ = self.omic_net_tumor(x_omic=kwargs['x_omic_tumor'])                     # [B, 59] -> [B, 128]
        omic_vec_immune, _, _ = self.omic_net_immune(x_omic=kwargs['x_omic_immune'])                  # [B, 361] -> [B, 128]
        if not self.args.return_vgrid:
            vec_t, hz_pt, grads_t = self.pathomic_net_tumor(path=x_path, omic=omic_vec_tumor)
            vec_i, hz_pi, grads_i = self.pathomic_net_immune(path=x_path, omic=omic_vec_immune)
        else:
            vec_t, hz_pt, grads_t, omic_tumor, vgrid_tumor = self.pathomic_net_tumor(path=x_path, omic=omic_vec_tumor)
            vec_i, hz_pi, grads_i, omic_immune, vgrid_immune = self.pathomic_net_immune(path=x_path, omic=omic_vec_immune)
        a, b = (vec_t.clone().detach(), vec_i.clone().detach()) if self.cut_fuse_grad else (vec_t, vec_i)
        if self.fusion_type == "concat":
            features = torch.cat((a, b), 1)
            cls = self.classifier
        else:
            features = self.fusion(a, b)
            cls = self.classifier[0]
        hazard = Fn.linear(features, cls.weight, cls.bias)
        hazard_tumor = Fn.linear(vec_t, self.classifier_tumor[0].weight, self.classifier_tumor[0].bias)
        hazard_immune = Fn.linear(vec_i, self.classifier_immune[0].weight, self.classifier_immune[0].bias)
        if self.return_grad == "True":
            raise NotImplementedError("return_grad='True' (get_grad_embedding, numpy export) is not on the training path")
        fuse_grads = None
        if self.args.task_type == "survival":
            hazard, hazard_tumor, hazard_immune = (Fn.unary(t, "sigmoid") for t in (hazard, hazard_tumor, hazard_immune))
        if self.args.return_vgrid:
            logits = [hazard_tumor, hazard_immune, hazard, omic_tumor, vgrid_tumor, omic_immune, vgrid_immune]
        else:
            logits = [hazard_tumor, hazard_immune, hazard]
        return features, vec_t, vec_i, logits, fuse_grads, grads_t, grads_i
