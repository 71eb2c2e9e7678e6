"""CUDA-graph capture of a whole train step (the loop body of train_test.py:57-85: forward, losses, backward,
optimizer step) so that a step is ONE graph launch instead of 300-400 kernel launches driven from Python.

    step = GraphedStep(lambda x, y: train_step(x, y), (x_example, y_example))
    loss = step(x, y)            # copies x, y into the graph's static inputs, replays, returns the static output tensor(s)

What makes the step capturable here: no kernel of libsml_b200 synchronises or allocates on the host, every buffer comes from
torch's caching allocator (graph-private pool during capture), tensor maps are kernel parameters, and the dropout masks --
counter-based hashes whose (seed, offset) are passed by value and would be frozen by the capture -- are salted on the device
(``sml_set_seed_salt`` / ``sml_seed_salt_advance``, include/sml_b200.h): the first node of the graph advances the salt, so every
replay draws fresh masks.  Optimizers must be built with ``capturable=True``.
"""
import torch

from ._lib import c_vp, check, lib, stream

_salt = {}


def seed_salt(device):
    """the per-device salt word pair {counter, salt}; installing it makes every dropout kernel of the library salted"""
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _salt:
        t = torch.zeros(2, dtype=torch.int64, device=device)
        _salt[key] = t
        check(lib.sml_set_seed_salt(c_vp(t.data_ptr())), "set_seed_salt")
    return _salt[key]


def advance_salt(device):
    t = seed_salt(device)
    check(lib.sml_seed_salt_advance(c_vp(t.data_ptr()), stream()), "seed_salt_advance")


class GraphedStep:
    """captures ``fn(*inputs)`` (any mix of forward / backward / optimizer step) after ``warmup`` eager runs on a side stream"""

    def __init__(self, fn, example_inputs, warmup=3, device=None):
        self.fn = fn
        self.static_in = [t.clone() if torch.is_tensor(t) else t for t in example_inputs]
        dev = device or next(t.device for t in self.static_in if torch.is_tensor(t))
        self.dev = dev
        seed_salt(dev)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                advance_salt(dev)
                out = fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            advance_salt(dev)
            out = fn(*self.static_in)
        self.static_out = out

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if torch.is_tensor(dst) and src is not dst and src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedDataParallelStep:
    """Data-parallel train step (main.py:117-119 wraps the model in DDP; train_test.py:57-85 is the step) for one process per
    GPU, built around a graph replay: forward + backward are captured ONCE into a CUDA graph that accumulates every parameter
    gradient into one flat buffer, the buffer is averaged over the ranks by ONE NCCL all-reduce (13 MB for TransMIL: tens of
    microseconds over NVLink, so nothing is lost by not overlapping it with the backward the way DDP's bucket hooks do), then
    the optimizer steps.  Parameters are broadcast from rank 0 at construction like DDP does.  Parameters that take no part in
    the loss keep a zero gradient (no unused-parameter search).

        step = GraphedDataParallelStep(model, optimizer, lambda x, y: loss_of(model, x, y), (x_example, y_example))
        loss = step(x, y)                                    # local (this rank's) loss, detached
    """

    def __init__(self, model, optimizer, loss_fn, example_inputs, group=None, warmup=2):
        import torch.distributed as dist
        self.dist, self.group, self.opt = dist, group, optimizer
        params = [p for p in model.parameters() if p.requires_grad]
        with torch.no_grad():
            for p in params:
                dist.broadcast(p, 0, group=group)
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, device=params[0].device, dtype=params[0].dtype)
        o = 0
        for p in params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)   # autograd accumulates in place into these views
            o += p.numel()
        self.params = params

        def fwd_bwd(*inp):
            self.flat.zero_()
            loss = loss_fn(*inp)
            loss.backward()
            return loss.detach()
        self.graphed = GraphedStep(fwd_bwd, example_inputs, warmup=warmup)

    def __call__(self, *inputs):
        loss = self.graphed(*inputs)
        self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.AVG if self.dist.get_backend(self.group) == "nccl" else self.dist.ReduceOp.SUM,
                             group=self.group)
        if self.dist.get_backend(self.group) != "nccl":
            self.flat.div_(self.dist.get_world_size(self.group))
        self.opt.step()
        return loss
