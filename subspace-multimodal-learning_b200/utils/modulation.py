"""Gradient modulation of the fusion classifier (/root/reference/train_test.py:87-184) as one device kernel.

    loss.backward()
    modulate_classifier_grads(model.module.classifier, pathomic_feat_tumor, pathomic_feat_immune, label[:, 5])
    optimizer.step()
"""
import torch

from .. import ops
from .._lib import c_i32, check, lib, ptr, stream


@torch.no_grad()
def modulate_classifier_grads(classifier, feat_t, feat_i, labels, return_info=False):
    """classifier: nn.Linear(2 * hs, L) whose ``weight.grad`` is rewritten in place (the reference indexes
    ``classifier.weight.grad[:, :hs]`` / ``[:, hs:]``, train_test.py:160); feat_t / feat_i [B, hs]; labels [B]."""
    W, b, G = classifier.weight, classifier.bias, classifier.weight.grad
    ops.require_cuda(W, b, G, feat_t, feat_i, labels)
    L, two_hs = W.shape
    hs = two_hs // 2
    assert G is not None and G.is_contiguous() and feat_t.shape == feat_i.shape == (labels.shape[0], hs)
    info = torch.empty(4, device=W.device, dtype=torch.float32) if return_info else None
    # every converted tensor is bound to a local that outlives the launch: a temporary freed right after ptr() could be
    # recycled by the caching allocator for the next conversion (two arguments aliasing one block)
    Wc, bc = W.detach().contiguous(), b.detach().contiguous()
    ft, fi = feat_t.detach().float().contiguous(), feat_i.detach().float().contiguous()
    lab = labels.to(torch.int64).contiguous()
    check(lib.sml_grad_modulate(ptr(Wc), ptr(bc), ptr(G), ptr(ft), ptr(fi), ptr(lab), c_i32(labels.shape[0]), c_i32(L), c_i32(hs),
                                ptr(info), stream()), "grad_modulate")
    del Wc, bc, ft, fi, lab
    return info
