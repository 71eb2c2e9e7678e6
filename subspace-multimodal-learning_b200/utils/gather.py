"""Drop-in for /root/reference/utils/gather.py:5-19: differentiable all_gather whose backward keeps only the local
slice (no reduction).  torch.distributed (NCCL over NVLink on B200, gloo in the CPU tests) is the transport."""
import torch
import torch.distributed as dist


class GatherLayer(torch.autograd.Function):
    """Gather tensors from all process, supporting backward propagation."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        output = [torch.zeros_like(input) for _ in range(dist.get_world_size())]
        dist.all_gather(output, input.contiguous())
        return tuple(output)

    @staticmethod
    def backward(ctx, *grads):
        (input,) = ctx.saved_tensors
        grad_out = torch.zeros_like(input)
        grad_out[:] = grads[dist.get_rank()]
        return grad_out


class PackedGather(torch.autograd.Function):
    """ONE all_gather for several small tensors (omic base + vgrid + ...): each rank packs its tensors into one flat
    buffer; outputs are the rank-major concatenations.  Backward = local slice, like GatherLayer."""

    @staticmethod
    def forward(ctx, *tensors):
        world, rank = dist.get_world_size(), dist.get_rank()
        flat = torch.cat([t.reshape(-1) for t in tensors]).contiguous()
        buf = torch.empty(world * flat.numel(), device=flat.device, dtype=flat.dtype)
        dist.all_gather_into_tensor(buf, flat)
        buf = buf.view(world, flat.numel())
        ctx.shapes = [t.shape for t in tensors]
        ctx.rank = rank
        outs, o = [], 0
        for t in tensors:
            n = t.numel()
            outs.append(buf[:, o:o + n].reshape(world * t.shape[0], *t.shape[1:]).contiguous())
            o += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        r = ctx.rank
        return tuple(g[r * s[0]:(r + 1) * s[0]].contiguous() for g, s in zip(grads, ctx.shapes))
