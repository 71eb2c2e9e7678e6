import os
import socket
import subprocess
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def run_ranks(mode, world=2, timeout=600):
    out = tempfile.mkdtemp(prefix="sml_dist_")
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), str(r), str(world), str(port), mode, out])
             for r in range(world)]
    for p in procs:
        assert p.wait(timeout=timeout) == 0, "rank failed"
    return [torch.load(os.path.join(out, "rank%d.pt" % r), weights_only=False) for r in range(world)]


def check_gather(res, world=2):
    B = res[0]["a"].shape[0]
    cat_a = torch.cat([r["a"] for r in res]); cat_v = torch.cat([r["v"] for r in res])
    for rk, r in enumerate(res):
        assert torch.equal(r["gather_fwd"], cat_a)
        assert torch.equal(r["packed_fwd"][0], cat_a) and torch.equal(r["packed_fwd"][1], cat_v)
        # backward keeps only the local slice of the incoming gradient, no reduction (gather.py:14-19)
        assert torch.allclose(r["ga"], 3.0 * r["w"][rk * B:(rk + 1) * B])
        assert torch.allclose(r["gv"], r["wv"][rk * 8 * B:(rk + 1) * 8 * B])
