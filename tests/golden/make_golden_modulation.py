"""Golden fixture for the gradient-modulation block: executes the reference's OWN source lines
(/root/reference/train_test.py:20-22 cosine_similarity and :87-184 the `if args.gradient_modulate:` block) verbatim in a
namespace holding synthetic tensors (authoring container only; nothing here runs on the GPU box).

    python tests/golden/make_golden_modulation.py
"""
import os
import sys
import textwrap
import types
from argparse import Namespace

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import randn  # noqa: E402

SRC = "/root/reference/train_test.py"


def reference_block():
    lines = open(SRC).read().split("\n")
    cos = "\n".join(lines[19:22])                                   # :20-22
    block = textwrap.dedent("\n".join(lines[86:184]))               # :87-184
    assert block.startswith("if args.gradient_modulate:") and "i_index = i_index+1" in block
    return cos, block


def run_case(seed, B=6, hs=128, L=4, favour="t"):
    cos, block = reference_block()
    cls = nn.Linear(2 * hs, L)
    with torch.no_grad():
        cls.weight.copy_(randn((L, 2 * hs), seed, 0.1)); cls.bias.copy_(randn((L,), seed + 1, 0.1))
    G = randn((L, 2 * hs), seed + 2)
    # make the halves of rows 0 and 2 conflict strongly, keep row 1 aligned
    G[0, hs:] = -G[0, :hs] + 0.3 * randn((hs,), seed + 3)
    G[2, hs:] = -0.5 * G[2, :hs] + 0.5 * randn((hs,), seed + 4)
    G[1, hs:] = G[1, :hs] + 0.1 * randn((hs,), seed + 5)
    cls.weight.grad = G.clone()
    ft = randn((B, hs), seed + 6); fi = randn((B, hs), seed + 7)
    labels = torch.from_numpy(__import__("numpy").random.default_rng(seed).integers(0, L, size=(B, 12)))
    if favour == "i":          # make the immune branch score higher -> ratio_t < 1 -> tumor half re-projected
        with torch.no_grad():
            for s in range(B):
                fi[s] = 3.0 * cls.weight[labels[s, 5], hs:] / cls.weight[labels[s, 5], hs:].norm() ** 2 * 4
    if favour == "t":          # tumor branch scores higher -> ratio_i < 1 -> immune half re-projected
        with torch.no_grad():
            for s in range(B):
                ft[s] = 3.0 * cls.weight[labels[s, 5], :hs] / cls.weight[labels[s, 5], :hs].norm() ** 2 * 4
    model = types.SimpleNamespace(module=types.SimpleNamespace(classifier=cls))
    ns = dict(torch=torch, F=F, args=Namespace(gradient_modulate=True, mmhid=hs, task_type="diag2021"), model=model,
              pathomic_feat_tumor=ft, pathomic_feat_immune=fi, label=labels, diag2021_loss_func=nn.CrossEntropyLoss(),
              grade_loss_func=None, subtype_loss_func=None, nll_loss_func=None)
    exec(cos, ns)
    with torch.no_grad():
        exec(block, ns)
    return dict(W=cls.weight.detach().clone(), b=cls.bias.detach().clone(), G=G, ft=ft, fi=fi, labels=labels[:, 5].clone(),
                G_out=cls.weight.grad.clone(), ratio_t=float(ns["ratio_t"]))


def main():
    cases = [run_case(300, favour="t"), run_case(310, favour="i"), run_case(320, B=8, favour="i")]
    for c in cases:
        changed = (c["G_out"] != c["G"]).any(dim=1)
        print("ratio_t %.3f rows changed %s" % (c["ratio_t"], changed.tolist()))
    torch.save(cases, os.path.join(HERE, "modulation.pt"))
    print("modulation: %.2f MB" % (os.path.getsize(os.path.join(HERE, "modulation.pt")) / 1e6))


if __name__ == "__main__":
    main()
