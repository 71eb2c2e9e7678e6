"""Deterministic synthetic tensors shared by the golden generator and the tests.

numpy's PCG64 stream is platform independent, so fixtures only need to store
shapes + seed for large inputs / parameters, not the values.
"""
import numpy as np
import torch


def randn(shape, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return torch.from_numpy((rng.standard_normal(tuple(shape)) * scale).astype(np.float32))


def synth_state_dict(shapes, seed):
    """shapes: ordered {name: shape}.  Weights ~ N(0, 1/sqrt(fan_in)), biases ~ N(0, .02),
    norm weights ~ 1 + N(0,.1), running_var > 0."""
    out = {}
    for i, (name, shape) in enumerate(shapes.items()):
        shape = tuple(shape)
        s = seed * 1000 + i
        if len(shape) == 0:                       # num_batches_tracked
            out[name] = torch.zeros((), dtype=torch.int64)
        elif name.endswith("running_var"):
            out[name] = randn(shape, s).abs() + 0.5
        elif name.endswith("running_mean"):
            out[name] = randn(shape, s, 0.1)
        elif len(shape) == 1 and ("norm" in name or ".1.weight" in name) and name.endswith("weight"):
            out[name] = 1.0 + randn(shape, s, 0.1)
        elif len(shape) == 1:
            out[name] = randn(shape, s, 0.02)
        else:
            fan_in = int(np.prod(shape[1:]))
            out[name] = randn(shape, s, 1.0 / np.sqrt(max(fan_in, 1)))
    return out
