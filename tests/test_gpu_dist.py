"""GPU box: 2 ranks (gloo transport, CUDA tensors, one GPU) — BatchLoss + PackedGather and SyncBatchNorm1d equal
the single-process result on the concatenated batch (SURVEY §8c distributed oracle)."""
import pytest
import torch

import oracle
from _dist_common import check_gather, run_ranks
from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_two_rank_batchloss_and_syncbn():
    res = run_ranks("gpu")
    check_gather(res)
    B = res[0]["om"].shape[0]
    om = torch.cat([r["om"] for r in res]).double().requires_grad_(); vg = torch.cat([r["vg"] for r in res]).double().requires_grad_()
    want = oracle.loss.batch_loss(om.unsqueeze(1).repeat(1, 7, 1), vg, 2 * B)
    for rk, r in enumerate(res):
        assert rel_err(r["bl"], want) < 1e-4
    want.backward(res[0]["gl"].double())
    for rk, r in enumerate(res):
        assert rel_err(r["bl_gom"], om.grad[rk * B:(rk + 1) * B]) < 1e-3
        assert rel_err(r["bl_gvg"], vg.grad[rk * 8 * B:(rk + 1) * 8 * B]) < 1e-3
    # sync BN vs BatchNorm1d on the concatenated (ragged) batch
    x = torch.cat([r["bn_x"] for r in res]).double().requires_grad_()
    bn = torch.nn.BatchNorm1d(16).double().train()
    with torch.no_grad():
        bn.weight.copy_(res[0]["bn_w"]); bn.bias.copy_(res[0]["bn_b"])
    y = bn(x)
    y.backward(torch.cat([r["bn_gy"] for r in res]).double())
    o = 0
    for r in res:
        n = r["bn_x"].shape[0]
        assert rel_err(r["bn_y"], y[o:o + n]) < 1e-5
        assert rel_err(r["bn_gx"], x.grad[o:o + n]) < 1e-4
        assert rel_err(r["bn_rv"], bn.running_var) < 1e-5
        o += n
    assert rel_err(res[0]["bn_gw"] + res[1]["bn_gw"], bn.weight.grad) < 1e-4
    # GraphedDataParallelStep: same losses as the eager data-parallel loop, identical weights on both ranks
    for r in res:
        assert abs(r["gdp_eager"][0] - r["gdp_graph"][0]) <= 1e-6 * abs(r["gdp_eager"][0])
        for a, b in zip(r["gdp_eager"][1:], r["gdp_graph"][1:]):
            assert abs(a - b) <= 5e-3 * abs(a), (r["gdp_eager"], r["gdp_graph"])
    assert torch.equal(res[0]["gdp_w"], res[1]["gdp_w"])
    assert rel_err(res[0]["gdp_w"], res[0]["gdp_w_eager"]) < 5e-3
