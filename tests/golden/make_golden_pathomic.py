"""Golden fixture for DeformPathomicNet (models/model.py:440-544) from the UNMODIFIED reference (authoring container only):

    python tests/golden/make_golden_pathomic.py

Same stubs as make_golden.py; config_mine.yaml values for the model arguments; eval mode (AlphaDropout / attention dropout
off), batch 2, N = 2500 (forced by the reference)."""
import os
import sys
from argparse import Namespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import grads_of, install_stubs, save, shapes_of  # noqa: E402
from synth import randn, synth_state_dict  # noqa: E402


def pathomic_args():
    return Namespace(init_type="max", input_size_omic_tumor=59, input_size_omic_immune=361, omic_dim=128, path_dim=128, mmhid=128,
                     return_grad="False", dropout_rate=0.1, label_dim=4, fusion_type="concat", cut_fuse_grad=False,
                     return_vgrid=True, attn_dim=2, task_type="diag2021")


def main():
    install_stubs()
    from models.model import DeformPathomicNet
    torch.manual_seed(110)
    net = DeformPathomicNet(pathomic_args()).eval()
    shp = shapes_of(net)
    st = synth_state_dict(shp, 111)
    for k in st:
        if k.endswith("to_offsets.2.weight"):
            st[k] = st[k] * 4.0
        if k.endswith("output_range"):
            st[k] = torch.tensor([6.0])
        if k.endswith("output_shift"):
            st[k] = torch.tensor([-3.0])
    net.load_state_dict(st)
    x_path = randn((2, 2500, 1024), 112)
    xt = randn((2, 59), 113).abs().requires_grad_(); xi = randn((2, 361), 114).abs().requires_grad_()
    feats, vt, vi, logits, _, _, _ = net(x_path=x_path, x_omic_tumor=xt, x_omic_immune=xi)
    loss = sum((l * randn(l.shape, 120 + j)).sum() for j, l in enumerate(logits[:3])) + (feats * randn(feats.shape, 125)).sum()
    loss = loss + 0.01 * ((logits[4] * randn(logits[4].shape, 126)).sum() + (logits[6] * randn(logits[6].shape, 127)).sum())
    loss.backward()
    gp = grads_of(net)
    save("pathomic", dict(shapes=shp, seed=111, offw_scale=4.0, feats=feats.detach(), vt=vt.detach(), vi=vi.detach(),
                          logits=[l.detach() for l in logits[:3]], vgrid_tumor=logits[4].detach(), vgrid_immune=logits[6].detach(),
                          omic_tumor_row=logits[3][:, 5].detach().clone(), gxt=xt.grad, gxi=xi.grad,
                          gnorms={k: float(v.norm()) for k, v in gp.items()},
                          gparams={k: v for k, v in gp.items() if v.numel() <= 2048}))


if __name__ == "__main__":
    main()
