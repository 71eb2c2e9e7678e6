"""Oracle (test infrastructure only): the 1-D continuous-position-bias MLP as an exact piecewise-linear table.

For DeformCrossAttention1D the bias MLP (models/DeformableAttention1D.py:60-102) maps ONE scalar t = sign(dx) log(|dx| + 1)
to `heads / groups` outputs through Linear(1,32) ReLU Linear(32,32) ReLU Linear(32,o): a composition of affine maps and
ReLUs of a scalar, i.e. a continuous piecewise-linear function with at most 32 + 33 * 32 breakpoints.  This module builds that
table (breakpoints, per-segment slope / intercept) and the mapping of per-segment gradient sums back to the MLP parameters.
It is the specification of the table-based 1-D kernel planned in DESIGN.md section 7 ("if the 1-D PWL
table replaces the MLP, count the table ops") and is checked on CPU against `oracle.deform.cpb_mlp`
(tests/test_oracle_golden.py:test_pwl_table_*)."""
import torch


def _mlp_params(p, prefix):
    return (p[prefix + "mlp.0.0.weight"].reshape(-1), p[prefix + "mlp.0.0.bias"], p[prefix + "mlp.1.0.weight"],
            p[prefix + "mlp.1.0.bias"], p[prefix + "mlp.2.weight"], p[prefix + "mlp.2.bias"])


def build_table(p, prefix=""):
    """Returns (breaks [S-1] ascending, slope [S, o], icpt [S, o]): bias(t) = slope[s] * t + icpt[s] on segment s, where
    s = number of breakpoints <= t."""
    w1, b1, W2, b2, W3, b3 = (x.detach().double() for x in _mlp_params(p, prefix))
    # layer-1 kinks: unit k switches at t = -b1 / w1
    k1 = [(-b1[k] / w1[k]).item() for k in range(w1.numel()) if w1[k] != 0]
    edges = sorted(set(k1))
    # inside every interval between layer-1 kinks, layer 2's pre-activations are affine in t: a2_c = A_c t + C_c
    brks = list(edges)
    bounds = [-float("inf")] + edges + [float("inf")]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        tm = _mid(lo, hi)
        act1 = ((w1 * tm + b1) > 0).double()
        A = W2 @ (act1 * w1)
        C = W2 @ (act1 * b1) + b2
        for c in range(A.numel()):
            if A[c] != 0:
                tz = (-C[c] / A[c]).item()
                if lo < tz < hi:
                    brks.append(tz)
    brks = torch.tensor(sorted(set(brks)), dtype=torch.float64)
    bounds = [-float("inf")] + brks.tolist() + [float("inf")]
    slope, icpt = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        tm = _mid(lo, hi)
        act1 = ((w1 * tm + b1) > 0).double()
        a1, c1 = act1 * w1, act1 * b1
        A, C = W2 @ a1, W2 @ c1 + b2
        act2 = ((A * tm + C) > 0).double()
        slope.append(W3 @ (act2 * A))
        icpt.append(W3 @ (act2 * C) + b3)
    return brks, torch.stack(slope), torch.stack(icpt)


def _mid(lo, hi):
    if lo == -float("inf") and hi == float("inf"):
        return 0.0
    if lo == -float("inf"):
        return hi - 1.0
    if hi == float("inf"):
        return lo + 1.0
    return 0.5 * (lo + hi)


def eval_table(t, table):
    """bias [.., o] for scalar positions t [..] (any shape) from the table."""
    brks, slope, icpt = table
    s = torch.searchsorted(brks, t.double().reshape(-1).contiguous(), right=True)
    out = slope[s] * t.double().reshape(-1, 1) + icpt[s]
    return out.reshape(*t.shape, -1)


def segment_sums(t, dbias, table):
    """What the table-based backward kernel accumulates per pair: G0[s, o] = sum dbias, G1[s, o] = sum dbias * t over the
    pairs whose t falls in segment s; plus d loss / d t = dbias . slope[s]."""
    brks, slope, _ = table
    tf = t.double().reshape(-1)
    s = torch.searchsorted(brks, tf.contiguous(), right=True)
    g = dbias.double().reshape(tf.numel(), -1)
    S = slope.shape[0]
    G0 = torch.zeros(S, g.shape[1], dtype=torch.float64).index_add_(0, s, g)
    G1 = torch.zeros(S, g.shape[1], dtype=torch.float64).index_add_(0, s, g * tf.unsqueeze(1))
    dt = (g * slope[s]).sum(-1).reshape(t.shape)
    return G0, G1, dt


def param_grads_from_segments(p, prefix, table, G0, G1):
    """d loss / d (MLP parameters) from the per-segment sums: on segment s, bias = alpha_s t + beta_s with (alpha, beta)
    smooth in the parameters (fixed ReLU pattern), and alpha_s, beta_s are linear in the MLP's values at two interior points
    t_a < t_b of the segment, so the parameter gradient is an ordinary MLP backward on 2 S virtual points."""
    brks, _, _ = table
    bounds = [-float("inf")] + brks.tolist() + [float("inf")]
    ta, tb = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if lo == -float("inf") and hi == float("inf"):
            a, b = -1.0, 1.0
        elif lo == -float("inf"):
            a, b = hi - 2.0, hi - 1.0
        elif hi == float("inf"):
            a, b = lo + 1.0, lo + 2.0
        else:
            a, b = lo + (hi - lo) / 3.0, lo + 2.0 * (hi - lo) / 3.0
        ta.append(a); tb.append(b)
    ta = torch.tensor(ta, dtype=torch.float64); tb = torch.tensor(tb, dtype=torch.float64)
    d = (tb - ta).unsqueeze(1)
    # alpha = (f_b - f_a) / d, beta = f_a - alpha t_a  ->  dL/df_a, dL/df_b
    ga = G0 * (1 + ta.unsqueeze(1) / d) - G1 / d
    gb = -G0 * ta.unsqueeze(1) / d + G1 / d
    names = ["mlp.0.0.weight", "mlp.0.0.bias", "mlp.1.0.weight", "mlp.1.0.bias", "mlp.2.weight", "mlp.2.bias"]
    q = {prefix + n: p[prefix + n].detach().double().requires_grad_() for n in names}
    from .deform import cpb_mlp
    f = cpb_mlp(torch.cat([ta, tb]).unsqueeze(1), q, prefix)
    f.backward(torch.cat([ga, gb]))
    return {k: v.grad for k, v in q.items()}
