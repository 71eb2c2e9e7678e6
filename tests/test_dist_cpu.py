"""CPU, world_size 2, gloo: host-side logic of the cross-rank exchange (GatherLayer / PackedGather)."""
from _dist_common import check_gather, run_ranks


def test_gather_layers_two_ranks_gloo():
    check_gather(run_ranks("cpu"))
