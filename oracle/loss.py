"""Oracle (test infrastructure only): subspace losses and their cross-rank exchange.

Restates /root/reference/utils/loss.py:13-40 (BatchLoss), utils/gather.py:5-19
(GatherLayer), models/cmta_utils.py:1212-1228 (OrthogonalLoss) and the
statistics of utils/sync_batchnorm/batchnorm.py:168-195.
"""
import torch


def batch_loss(omic_all, vgrid_all, n_total):
    """BatchLoss.forward after the all_gather+cat (loss.py:22-40).

    omic_all [Ng, ...], vgrid_all [8*Ng, 2, hk, wk] are the rank-major
    concatenations; the ``view(8, Ng, -1)`` is a raw reinterpretation (a15).
    """
    o = omic_all.reshape(n_total, -1)
    v = vgrid_all.reshape(8, n_total, -1)
    s = o @ o.t()
    s = s / s.pow(2).sum(1, keepdim=True).sqrt()
    ts = []
    for i in range(8):
        t = v[i] @ v[i].t()
        ts.append(t / t.pow(2).sum(1, keepdim=True).sqrt())
    t = torch.stack(ts).mean(0)
    return (s - t) ** 2 / n_total


def gather_backward(grads_per_rank, rank):
    """GatherLayer.backward keeps only the local slice, no reduction (gather.py:14-19)."""
    return grads_per_rank[rank].clone()


def _cos(a, b, eps=1e-8):
    na = a.norm(dim=1).clamp_min(eps)
    nb = b.norm(dim=1).clamp_min(eps)
    return (a * b).sum(1) / (na * nb)


def orthogonal_loss(P, P_hat, G, G_hat, gamma=0.5):
    """OrthogonalLoss.forward (cmta_utils.py:1217-1228); barred terms are detached."""
    Pd, Gd = P.detach(), G.detach()
    pos = (1 - _cos(Pd, P_hat).abs()) + (1 - _cos(Gd, G_hat).abs())
    neg = _cos(P, G).abs() + _cos(Pd, G_hat).abs() + _cos(Gd, P_hat).abs()
    return pos + gamma * neg


def sync_bn_stats(x_per_rank, eps=1e-5):
    """Global (sum, sum-of-squares) BN statistics over all ranks' rows
    (sync_batchnorm/batchnorm.py:168-195: biased var clamped at eps, unbiased for running)."""
    x = torch.cat(list(x_per_rank), 0)
    n = x.shape[0]
    s, ss = x.sum(0), (x * x).sum(0)
    mean = s / n
    sumvar = ss - s * mean
    bias_var = sumvar / n
    unbias_var = sumvar / (n - 1)
    inv_std = bias_var.clamp(min=eps) ** -0.5
    return mean, inv_std, unbias_var
