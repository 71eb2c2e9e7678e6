"""Golden fixtures for CMTA and MCAT_Surv (models/model.py:559-853, encoder blocks models/cmta_utils.py:858-948) from the
UNMODIFIED reference (authoring container only):

    python tests/golden/make_golden_cmta.py

Same stubs as make_golden.py; eval mode (all dropouts off); batch 2; N = 150 patches (padded to 13 x 13 = 169 by
Transformer_P), 4 genomic groups [100, 100, 100, 131] (the live configuration) and, for the co-attention benchmark shape of
BASELINE.json configs[0], MCAT_Surv with 6 groups [72, 72, 72, 72, 72, 71]."""
import os
import sys
from argparse import Namespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import grads_of, install_stubs, save, shapes_of  # noqa: E402
from synth import randn, synth_state_dict  # noqa: E402


def run(net, x_path, x_omic, n_out, seeds, double=True):
    """The unmodified reference modules are evaluated in fp64 (``net.double()``): torch's CPU fp32 backward of the depthwise
    (33,1) residual conv w.r.t. its weight is off by 6e-2 relative to an explicit tap sum for [2,8,256,32] inputs (mkldnn path;
    fp64 agrees to 1e-15), so fp32 reference gradients would pin the oracle to a library artefact.  Outputs of
    the fp32 and fp64 evaluations agree to 1e-6."""
    dt = torch.float64 if double else torch.float32     # MCAT_Surv casts to .float() internally (model.py:640): fp32 only, no conv in it
    net.to(dt)
    out = net(x_path=x_path.to(dt), x_omic=x_omic.to(dt))
    outs = list(out[:n_out])
    loss = sum((o * randn(o.shape, seeds + j).to(dt)).sum() for j, o in enumerate(outs) if o.dtype.is_floating_point)
    loss.backward()
    return [o.float() for o in outs]


def main():
    install_stubs()
    from models.model import CMTA, MCAT_Surv
    args = Namespace(label_dim=4)
    fx = {}
    # ---- CMTA ------------------------------------------------------------------------------------------------------
    torch.manual_seed(210)
    net = CMTA(args).eval()
    shp = shapes_of(net)
    st = synth_state_dict(shp, 211)
    for k in st:                                   # cls tokens are ~1e-6 at init; keep them O(0.1) so that their gradient path is exercised
        if k.endswith("cls_token"):
            st[k] = st[k] * 0.1
    net.load_state_dict(st)
    x_path = randn((2, 150, 1024), 212)
    x_omic = randn((2, 431), 213).abs()
    outs = run(net, x_path, x_omic, 7, 220)
    gp = grads_of(net)
    fx["cmta"] = dict(shapes=shp, seed=211, outs=[o.detach() for o in outs], gnorms={k: float(v.norm()) for k, v in gp.items()},
                      gparams={k: v.float() for k, v in gp.items() if v.numel() <= 4096})
    # ---- MCAT_Surv, 4 groups (live) and 6 groups (BASELINE configs[0]) --------------------------------------------
    for tag, sizes, seed in (("mcat4", [100, 100, 100, 131], 230), ("mcat6", [72, 72, 72, 72, 72, 71], 240)):
        torch.manual_seed(seed)
        net = MCAT_Surv(args, omic_sizes=sizes).eval()
        shp = shapes_of(net)
        st = synth_state_dict(shp, seed + 1)
        net.load_state_dict(st)
        x_path = randn((2, 150, 1024), seed + 2)
        x_omic = randn((2, 431), seed + 3).abs()
        outs = run(net, x_path, x_omic, 3, seed + 10, double=False)
        gp = grads_of(net)
        fx[tag] = dict(shapes=shp, seed=seed + 1, sizes=sizes, outs=[o.detach() for o in outs],
                       gnorms={k: float(v.norm()) for k, v in gp.items()},
                       gparams={k: v.float() for k, v in gp.items() if v.numel() <= 4096})
    save("cmta", fx)


if __name__ == "__main__":
    main()
