"""evaluation metrics on the device (utils/metrics.py:9-25) and checkpoint compatibility (train_test.py:263-290)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SML_REFERENCE", "/root/reference")


@pytest.mark.parametrize("seed,n,C", [(0, 236, 4), (1, 57, 3), (2, 40, 4)])
def test_metrics_match_sklearn(seed, n, C):
    """the torch implementation against the sklearn calls the reference makes (imblearn's sensitivity / specificity restated
    from the confusion matrix), including tied scores and a class that is never predicted"""
    from sklearn.metrics import accuracy_score, f1_score, balanced_accuracy_score, confusion_matrix, roc_auc_score, precision_score, recall_score
    import sml_b200
    from sml_b200.utils.evaluate import compute_avg_metrics
    rng = np.random.default_rng(seed)
    y = rng.integers(0, C, n)
    y[:C] = np.arange(C)
    logits = rng.standard_normal((n, C)) + 1.5 * np.eye(C)[y]
    if seed == 2:
        logits[:, C - 1] = -50.0                              # a class no sample is assigned to
        logits = np.round(logits, 1)                          # ties
    act = torch.softmax(torch.from_numpy(logits), -1)
    got = compute_avg_metrics(torch.from_numpy(y), act)
    a = act.numpy(); pred = a.argmax(-1)
    cm = confusion_matrix(y, pred, labels=sorted(set(y) | set(pred)))
    tp = np.diag(cm); fp = cm.sum(0) - tp; fn = cm.sum(1) - tp; tn = cm.sum() - tp - fp - fn
    want = (accuracy_score(y, pred), f1_score(y, pred, average="macro", zero_division=0), roc_auc_score(y, a, multi_class="ovr"),
            balanced_accuracy_score(y, pred), recall_score(y, pred, average="macro", zero_division=0), float(np.mean(tn / (tn + fp))),
            precision_score(y, pred, average="macro", zero_division=0))
    assert np.allclose(got, want, atol=1e-9), (got, want)


def test_auc_is_zero_when_a_class_is_absent():
    from sml_b200.utils.evaluate import compute_avg_metrics
    y = torch.tensor([0, 1, 1, 0, 1])
    act = torch.softmax(torch.randn(5, 3, generator=torch.Generator().manual_seed(0)), -1)
    assert compute_avg_metrics(y, act)[2] == 0.0              # utils/metrics.py:15-19: ValueError -> auc = 0


SCRIPT = r'''
import sys, os, tempfile
ROOT, REF = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests/golden")
from make_golden import install_stubs
install_stubs()
import torch
from make_golden_pathomic import pathomic_args
from models.model import DeformPathomicNet as RefNet                  # the unmodified reference class
import sml_b200
torch.manual_seed(3)
ref = RefNet(pathomic_args())
ours = sml_b200.DeformPathomicNet(pathomic_args())
d = tempfile.mkdtemp()
torch.save(ref.state_dict(), os.path.join(d, "ref.pth"))             # train_test.py:283-288
missing = ours.load_state_dict(torch.load(os.path.join(d, "ref.pth")), strict=True)
torch.save(ours.state_dict(), os.path.join(d, "ours.pth"))
back = torch.load(os.path.join(d, "ours.pth"))
ref2 = RefNet(pathomic_args())
ref2.load_state_dict(back, strict=True)
a, b = ref.state_dict(), ref2.state_dict()
assert list(a.keys()) == list(b.keys()) == list(ours.state_dict().keys())
assert all(torch.equal(a[k], b[k]) for k in a)
print("CKPT_OK", len(a))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_checkpoint_round_trip_with_reference_pth():
    """a .pth written from the unmodified reference DeformPathomicNet loads strict into the drop-in, and the drop-in's .pth loads
    strict back into the reference: same keys, same order, same bits"""
    env = dict(os.environ, PYTHONPATH=REF + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "CKPT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_epoch_val_on_device(tmp_path):
    """epochVal over a BagLoader: one host read at the end, results equal to the reference's formulae on the collected outputs"""
    import sml_b200
    from argparse import Namespace
    from sml_b200.utils.evaluate import epochVal, compute_avg_metrics
    margs = Namespace(init_type="max", input_size_omic_tumor=59, input_size_omic_immune=361, omic_dim=128, path_dim=128, mmhid=128,
                      return_grad="False", dropout_rate=0.1, label_dim=4, fusion_type="concat", cut_fuse_grad=False,
                      return_vgrid=False, attn_dim=2, task_type="diag2021")
    torch.manual_seed(0)
    net = sml_b200.DeformPathomicNet(margs).cuda()
    net.set_precision("bf16")
    rng = np.random.default_rng(0)
    n = 6
    bags = rng.standard_normal((n, 2500, 1024)).astype(np.float32)
    lab = rng.integers(0, 4, (n, 12)); lab[:4, 5] = np.arange(4)
    path = os.path.join(tmp_path, "val.bag")
    sml_b200.data.pack_bags(path, bags, omic=rng.random((n, 431)), omic_tumor=rng.random((n, 59)), omic_immune=rng.random((n, 361)), labels=lab)
    sh = sml_b200.data.BagShard.open(path)
    ld = sml_b200.data.BagLoader(sh, 2, "cuda", shuffle=False, drop_last=False)
    got = epochVal(net, ld, Namespace(task_type="diag2021"))
    assert net.training
    net.eval()
    outs = []
    with torch.no_grad():
        for xp, xo, xt, xi, y in ld:
            outs.append(torch.softmax(net(x_path=xp, x_omic=xo, x_omic_tumor=xt, x_omic_immune=xi)[3][2], 1))
    want = compute_avg_metrics(torch.from_numpy(lab[:, 5]), torch.cat(outs).cpu())
    assert np.allclose(got, want, atol=1e-6), (got, want)
    sh.close()


@pytest.mark.gpu
def test_async_evaluator_runs_beside_training(tmp_path):
    """AsyncEvaluator: the evaluation of a weight snapshot runs on a side stream while the train loop keeps stepping; its
    metrics equal the synchronous epochVal of a model holding the snapshot"""
    import copy
    import sml_b200
    from argparse import Namespace
    from sml_b200.utils.evaluate import AsyncEvaluator, epochVal
    margs = Namespace(init_type="max", input_size_omic_tumor=59, input_size_omic_immune=361, omic_dim=128, path_dim=128, mmhid=128,
                      return_grad="False", dropout_rate=0.1, label_dim=4, fusion_type="concat", cut_fuse_grad=False,
                      return_vgrid=False, attn_dim=2, task_type="diag2021")
    torch.manual_seed(1)
    net = sml_b200.DeformPathomicNet(margs).cuda().train(); net.set_precision("bf16")
    shadow = sml_b200.DeformPathomicNet(margs).cuda(); shadow.set_precision("bf16")
    rng = np.random.default_rng(3)
    n = 4
    lab = rng.integers(0, 4, (n, 12)); lab[:4, 5] = np.arange(4)
    path = os.path.join(tmp_path, "val.bag")
    sml_b200.data.pack_bags(path, rng.standard_normal((n, 2500, 1024)).astype(np.float32), omic=rng.random((n, 431)),
                            omic_tumor=rng.random((n, 59)), omic_immune=rng.random((n, 361)), labels=lab)
    sh = sml_b200.data.BagShard.open(path)
    mk = lambda: sml_b200.data.BagLoader(sh, 2, "cuda", shuffle=False, drop_last=False)
    ev = AsyncEvaluator(shadow, mk, Namespace(task_type="diag2021"))
    snap = copy.deepcopy(net.state_dict())
    assert ev.submit(net, tag=0)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    x = torch.randn(2, 2500, 1024, device="cuda"); xt = torch.rand(2, 59, device="cuda"); xi = torch.rand(2, 361, device="cuda")
    y = torch.tensor([0, 2], device="cuda")
    for _ in range(3):                                       # training goes on (and changes the weights) while the eval runs
        opt.zero_grad(set_to_none=True)
        lg = net(x_path=x, x_omic_tumor=xt, x_omic_immune=xi)[3][2]
        torch.nn.functional.cross_entropy(lg, y).backward()
        opt.step()
    (tag, got), = ev.wait()
    assert tag == 0 and not ev.busy()
    ref = sml_b200.DeformPathomicNet(margs).cuda(); ref.set_precision("bf16"); ref.load_state_dict(snap)
    want = epochVal(ref, mk(), Namespace(task_type="diag2021"))
    assert np.allclose(got, want, atol=1e-6), (got, want)
    assert any(float((a - b).abs().max()) > 0 for a, b in zip(net.state_dict().values(), snap.values()))   # weights did move
    sh.close()
