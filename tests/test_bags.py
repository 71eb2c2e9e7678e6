"""Bag shard format + loader (reference: data/dataset.py:137-140, main.py:51-65)."""
import os

import numpy as np
import pytest
import torch


def _bags_module():
    import sml_b200
    return sml_b200.data.bags


def test_shard_roundtrip_and_sampler_partition(tmp_path):
    """pack -> open gives torch's round-to-nearest-even bf16 of the fp32 bags and the side arrays unchanged; the per-rank index
    partition equals torch's DistributedSampler (the reference's sampler, main.py:51-54) for every rank / epoch"""
    m = _bags_module()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((6, 37, 24)).astype(np.float32)
    x[0, 0, 0] = np.nan; x[0, 0, 1] = np.inf; x[0, 0, 2] = 1.00390625          # tie: rounds to even
    om, ot, oi = rng.random((6, 431)), rng.random((6, 59)), rng.random((6, 361))
    lab = rng.integers(0, 4, (6, 12))
    path = os.path.join(tmp_path, "t.bag")
    m.pack_bags(path, x, omic=om, omic_tumor=ot, omic_immune=oi, labels=lab)
    sh = m.BagShard.open(path, pin=False)
    want = torch.from_numpy(x).to(torch.bfloat16)
    got = sh.bags
    assert got.shape == (6, 37, 24)
    assert torch.equal(got.view(torch.int16)[~torch.isnan(want)], want.view(torch.int16)[~torch.isnan(want)])
    assert torch.isnan(got[0, 0, 0])
    assert np.array_equal(sh.side["labels"].numpy(), lab)
    assert np.allclose(sh.side["omic_immune"].numpy(), oi.astype(np.float32))
    from torch.utils.data.distributed import DistributedSampler
    for world in (1, 2, 4):
        for rank in range(world):
            for epoch in (0, 3):
                ds = DistributedSampler(range(10), num_replicas=world, rank=rank, shuffle=True, seed=5)
                ds.set_epoch(epoch)
                want_idx = list(ds)
                got_b = m.epoch_indices(10, 2, rank=rank, world=world, shuffle=True, seed=5, epoch=epoch, drop_last=True)
                assert [i for b in got_b for i in b] == want_idx[:len(want_idx) // 2 * 2]
    sh.close()


@pytest.mark.gpu
def test_loader_streams_batches_in_sampler_order(tmp_path):
    m = _bags_module()
    rng = np.random.default_rng(1)
    x = rng.standard_normal((10, 50, 32)).astype(np.float32)
    lab = np.arange(120).reshape(10, 12)
    ot = rng.random((10, 59))
    path = os.path.join(tmp_path, "t.bag")
    m.pack_bags(path, x, omic_tumor=ot, labels=lab)
    sh = m.BagShard.open(path, pin=True)
    assert sh.pinned
    ld = m.BagLoader(sh, 3, "cuda", rank=1, world=2, shuffle=True, seed=7, prefetch=2)
    want_batches = m.epoch_indices(10, 3, rank=1, world=2, shuffle=True, seed=7, epoch=0)
    xs = torch.from_numpy(x).to(torch.bfloat16)
    seen = 0
    for (xp, xo, xt, xi, y), idx in zip(ld, want_batches):
        assert xo is None and xi is None
        assert torch.equal(xp.cpu(), xs[idx]) and torch.equal(y.cpu(), torch.from_numpy(lab)[idx])
        assert torch.allclose(xt.cpu(), torch.from_numpy(ot).float()[idx])
        seen += 1
    assert seen == len(want_batches) == len(ld)
    # streaming across epochs: 7 batches = epochs 0, 1, 2, ... back to back, each consumed while later ones are in flight
    order = []
    for ep in range(8):
        order += m.epoch_indices(10, 3, rank=1, world=2, shuffle=True, seed=7, epoch=ep)
    acc = []
    for k, (xp, _, _, _, y) in enumerate(ld.stream(7)):
        acc.append((xp.float().sum(dim=(1, 2)) * 1.0).clone())          # some work on the consumer stream
        assert torch.equal(y.cpu(), torch.from_numpy(lab)[order[k]])
    torch.cuda.synchronize()
    for k, a in enumerate(acc):
        assert torch.allclose(a.cpu(), xs[order[k]].float().sum(dim=(1, 2)))
    sh.close()


@pytest.mark.gpu
def test_bf16_bags_are_bit_identical_to_fp32_bags_in_bf16_mode():
    """the packed feed changes nothing in "bf16" mode (the GEMM operand is the same bf16 rounding of the bag) and is widened
    exactly in the fp32-grade modes"""
    import sml_b200
    from argparse import Namespace
    torch.manual_seed(0)
    tm = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).cuda().eval()
    x = torch.randn(2, 700, 1024, device="cuda")
    outs = {}
    for mode in ("bf16", "bf16x3"):
        tm.set_precision(mode)
        for name, inp in (("f32", x if mode == "bf16" else x.to(torch.bfloat16).float()), ("packed", x.to(torch.bfloat16))):
            tm.zero_grad(set_to_none=True)
            enc, lg, _ = tm(inp)
            (lg.sum() + enc.sum()).backward()
            outs[(mode, name)] = (lg.detach().clone(), tm._fc1[0].weight.grad.clone())
        assert torch.equal(outs[(mode, "f32")][0], outs[(mode, "packed")][0])
        # same operands in every GEMM; the backward has launch-order-dependent fp32 accumulations (atomics in the column sums
        # and landmark scatter), which the cancellation-heavy _fc1 weight gradient amplifies: compared at 1e-3 of its scale
        ga, gb = outs[(mode, "f32")][1], outs[(mode, "packed")][1]
        assert float((ga - gb).abs().max()) <= 1e-3 * float(ga.abs().max())
    dm = sml_b200.DeformCrossTransMIL(Namespace(path_dim=128, attn_dim=2, return_vgrid=False)).cuda().eval()
    dm.set_precision("bf16")
    xp = torch.randn(1, 2500, 1024, device="cuda"); om = torch.randn(1, 128, device="cuda").abs()
    # the deformable models run their projection GEMMs fp32-grade in every mode: a packed bag is widened exactly
    a = dm(xp.to(torch.bfloat16).float(), om)[1]; b = dm(xp.to(torch.bfloat16), om)[1]
    assert torch.equal(a, b)
