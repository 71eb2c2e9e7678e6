"""Oracle (test infrastructure only): the gradient-modulation block of /root/reference/train_test.py:87-184 for the
softmax-score task types (diag2021 / grade / subtype), restated on plain tensors.  Pinned against the reference's own source
lines executed verbatim (tests/golden/make_golden_modulation.py -> tests/golden/modulation.pt)."""
import torch


def modulate(W, b, G, feat_t, feat_i, labels):
    """Returns the modified copy of G [L, 2 hs] (train_test.py:89-92 logits, :118-119 scores, :148-149 ratios, :158-181 rows)."""
    hs = W.shape[1] // 2
    out_t = feat_t @ W[:, :hs].t() + b / 2
    out_i = feat_i @ W[:, hs:].t() + b / 2
    pt, pi = torch.softmax(out_t, 1), torch.softmax(out_i, 1)
    score_t = sum(pt[i, labels[i]] for i in range(pt.shape[0]))
    score_i = sum(pi[i, labels[i]] for i in range(pi.shape[0]))
    ratio_t = score_t / score_i
    ratio_i = 1 / ratio_t
    G = G.clone()
    for r in range(G.shape[0]):
        gt, gi = G[r, :hs].clone(), G[r, hs:].clone()
        sim = torch.dot(gt, gi) / (gt.norm() * gi.norm())
        if sim < 0:
            if ratio_t < 1:
                proj = torch.dot(gt, gi) / gi.norm() ** 2 * gi
                gt = gt - proj
                perpen = gt - proj
                G[r, :hs] = gt.norm() * perpen / perpen.norm()
            elif ratio_i < 1:
                proj = torch.dot(gi, gt) / gt.norm() ** 2 * gt
                gi = gi - proj
                perpen = gi - proj
                G[r, hs:] = gi.norm() * perpen / perpen.norm()
    return G
