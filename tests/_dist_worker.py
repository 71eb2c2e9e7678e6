"""Worker for the 2-rank tests (spawned by tests/test_dist_cpu.py and tests/test_gpu_dist.py)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def run(rank, world, port, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from synth import randn
    import sml_b200
    from sml_b200.utils.gather import GatherLayer, PackedGather
    dev = "cuda" if mode == "gpu" else "cpu"
    B = 3
    res = {}
    # ---- gather semantics (gather.py:5-19): forward = rank-major concatenation, backward = own slice only
    a = (randn((B, 5), 10 + rank).to(dev)).requires_grad_()
    v = (randn((8 * B, 2, 3, 3), 20 + rank).to(dev)).requires_grad_()
    ga = torch.cat(GatherLayer.apply(a), dim=0)
    pa, pv = PackedGather.apply(a, v)
    w = randn((world * B, 5), 30).to(dev)
    wv = randn((world * 8 * B, 2, 3, 3), 31).to(dev)
    ((ga * w).sum() + 2.0 * (pa * w).sum() + (pv * wv).sum()).backward()
    res["gather_fwd"] = ga.detach().cpu()
    res["packed_fwd"] = (pa.detach().cpu(), pv.detach().cpu())
    res["ga"] = a.grad.cpu(); res["gv"] = v.grad.cpu()
    res["w"] = w.cpu(); res["wv"] = wv.cpu()
    res["a"] = a.detach().cpu(); res["v"] = v.detach().cpu()
    if mode == "gpu":
        # ---- BatchLoss over 2 ranks == single-process loss on the concatenated batch
        om = randn((B, 128), 40 + rank).abs().to(dev).requires_grad_()
        vg = randn((8 * B, 2, 4, 4), 50 + rank).to(dev).requires_grad_()
        rep = om.unsqueeze(1).repeat(1, 7, 1); rep._sml_base = om
        loss = sml_b200.BatchLoss(B, world)(rep, vg)
        gl = randn(loss.shape, 60).to(dev)
        loss.backward(gl)
        res["bl"] = loss.detach().cpu(); res["bl_gom"] = om.grad.cpu(); res["bl_gvg"] = vg.grad.cpu()
        res["om"] = om.detach().cpu(); res["vg"] = vg.detach().cpu(); res["gl"] = gl.cpu()
        # ---- sync batch norm == batch norm over the concatenated batch
        bn = sml_b200.SyncBatchNorm1d(16).to(dev).train()
        with torch.no_grad():
            bn.weight.copy_(1.0 + 0.1 * randn((16,), 70).to(dev)); bn.bias.copy_(0.1 * randn((16,), 71).to(dev))
        x = randn((4 + rank, 16), 80 + rank).to(dev).requires_grad_()       # ragged per-rank batch
        y = bn(x)
        gy = randn(y.shape, 90 + rank).to(dev)
        y.backward(gy)
        res["bn_x"] = x.detach().cpu(); res["bn_y"] = y.detach().cpu(); res["bn_gy"] = gy.cpu(); res["bn_gx"] = x.grad.cpu()
        res["bn_w"] = bn.weight.detach().cpu(); res["bn_b"] = bn.bias.detach().cpu()
        res["bn_gw"] = bn.weight.grad.cpu(); res["bn_rv"] = bn.running_var.cpu()
    if mode == "gpu":
        # ---- GraphedDataParallelStep == the same steps with eager backward + per-parameter gradient averaging
        from argparse import Namespace
        from sml_b200 import functional as Fn
        import copy
        torch.manual_seed(5)
        net0 = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev).eval()      # eval: no dropout (offsets are host-side)
        net0.set_precision("bf16")
        st0 = copy.deepcopy(net0.state_dict())
        x = randn((2, 600, 1024), 100 + rank).to(dev); y = torch.tensor([rank, 3 - rank], device=dev)
        F = torch.nn.functional

        def make():
            net = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev).eval()
            net.set_precision("bf16"); net.load_state_dict(st0)
            return net, torch.optim.Adam(net.parameters(), lr=1e-3, fused=True, capturable=True)
        net_a, opt_a = make()
        la = []
        for _ in range(3):
            opt_a.zero_grad(set_to_none=True)
            enc, lg, _ = net_a(x)
            loss = F.cross_entropy(lg, y) + 0.0 * enc.sum()
            loss.backward()
            for p_ in net_a.parameters():
                dist.all_reduce(p_.grad); p_.grad.div_(world)
            opt_a.step()
            la.append(float(loss))
        net_b, opt_b = make()
        step = sml_b200.GraphedDataParallelStep(net_b, opt_b, lambda xx, yy: (lambda o: F.cross_entropy(o[1], yy) + 0.0 * o[0].sum())(net_b(xx)),
                                                (x, y))
        with torch.no_grad():                                # the capture's warm-up runs accumulate nothing into the weights
            for k, v in net_b.state_dict().items():
                assert torch.equal(v, st0[k].to(dev)), k
        lb = [float(step(x, y)) for _ in range(3)]
        res["gdp_eager"] = la; res["gdp_graph"] = lb
        res["gdp_w"] = net_b._fc2.weight.detach().cpu(); res["gdp_w_eager"] = net_a._fc2.weight.detach().cpu()
    torch.save(res, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
