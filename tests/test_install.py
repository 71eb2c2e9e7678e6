"""install(): the reference's own factory (models/model.py:define_net) must build the sml_b200 drop-ins once
``sml_b200.install()`` has run.  Needs the reference tree (authoring container: /root/reference, or $SML_REFERENCE); on a box
without it only the module registration part is checked."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SML_REFERENCE", "/root/reference")

SCRIPT = r'''
import sys, types, json
ROOT, REF = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests/golden")
have_ref = REF != "-"
if have_ref:
    from make_golden import install_stubs          # matplotlib / lifelines / sksurv / imblearn stubs, .cuda() no-op (SURVEY 8c)
    install_stubs()
import sml_b200
rep = sml_b200.install()
out = {"modules": rep["modules"], "failed": rep["failed"], "patched": rep["patched"]}
import nystrom_attention
assert nystrom_attention.NystromAttention is sml_b200.NystromAttention
if have_ref:
    import yaml
    from argparse import Namespace
    import models.model as ref_model, models.mil as ref_mil, models.cmta_utils as ref_cmta, models.fusion as ref_fusion
    import utils.loss, utils.gather, models.DeformCrossTransMIL as ref_dct
    assert ref_mil.TransMIL is sml_b200.TransMIL and ref_mil.TransLayer is sml_b200.TransLayer and ref_mil.PPEG is sml_b200.PPEG
    assert hasattr(ref_mil, "ABMIL") and ref_mil.ABMIL.__module__ == "models.mil"           # out-of-scope classes stay the reference's
    assert ref_model.TransMIL is sml_b200.TransMIL and ref_model.DeformPathomicNet is sml_b200.DeformPathomicNet
    assert ref_model.CMTA is sml_b200.CMTA and ref_model.MCAT_Surv is sml_b200.MCAT_Surv
    assert ref_cmta.Transformer_P is sml_b200.Transformer_P and ref_cmta.NystromAttention is sml_b200.NystromAttention
    assert ref_fusion.BilinearFusion is sml_b200.BilinearFusion
    assert utils.loss.BatchLoss is sml_b200.BatchLoss and utils.gather.GatherLayer is sml_b200.GatherLayer
    assert ref_dct.DeformCrossTransMIL is sml_b200.DeformCrossTransMIL
    cfg = yaml.safe_load(open(REF + "/config/config_mine.yaml"))
    built = {}
    for mode in ("deformpathomic", "cmta", "mcat"):
        args = Namespace(**dict(cfg, mode=mode, return_grad="False"))
        net = ref_model.define_net(args)                                                    # the reference factory, unchanged
        built[mode] = type(net).__module__ + "." + type(net).__name__
        kinds = {type(m).__module__.split(".")[0] for m in net.modules()
                 if type(m).__name__ in ("NystromAttention", "DeformCrossAttention2D", "MultiheadAttention", "TransLayer", "PPEG")
                 and type(m).__module__ != "torch.nn.modules.activation"}    # MCAT's stock nn.TransformerEncoder on <= 6 tokens
        assert kinds <= {"sml_b200"}, (mode, kinds)
    out["built"] = built
    assert not rep["failed"], rep["failed"]
print("RESULT " + json.dumps(out))
'''


def _run(ref):
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, ref], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    import json
    return json.loads(line[7:])


def test_install_registers_modules():
    out = _run("-")
    assert "models.NystromAttention" in out["modules"] and "utils.loss" in out["modules"]
    # without the reference tree the class swaps are reported as failed, not swallowed
    assert "models.model" in out["failed"] or "models.model" in out["patched"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_install_reference_factory_builds_dropins():
    out = _run(REF)
    assert out["built"] == {"deformpathomic": "sml_b200.models.model.DeformPathomicNet", "cmta": "sml_b200.models.cmta.CMTA",
                            "mcat": "sml_b200.models.cmta.MCAT_Surv"}
