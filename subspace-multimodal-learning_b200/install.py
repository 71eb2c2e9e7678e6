"""Make the reference's own `main.py` / `train_test.py` / `inference.py` run on the B200 kernels without editing
them: register the drop-in modules under the module names the reference imports and swap the classes the reference's
factory (`models/model.py:define_net`) looks up in its own namespace.

    import sys; sys.path.insert(0, "/path/to/Subspace-Multimodal-Learning"); sys.path.insert(0, "/path/to/this/repo")
    import sml_b200; sml_b200.install()          # before or after `import models.model`
    import main                                  # the reference, unchanged

What is replaced:
  * whole modules (imported by name in the reference): nystrom_attention (models/mil.py:24, DeformCrossTransMIL.py:24),
    models.NystromAttention, models.DeformableAttention1D / 2D, models.MultiheadAttention, models.DeformCrossTransMIL,
    utils.loss, utils.gather;
  * classes patched INTO reference modules that also hold out-of-scope code: models.mil.{TransMIL, TransLayer, PPEG}
    (ABMIL / GatedABMIL stay the reference's), models.cmta_utils.{NystromAttention, TransLayer, PPEG, Transformer_P,
    Transformer_G, MultiheadAttention, OrthogonalLoss}, models.fusion.BilinearFusion,
    models.model.{DeformPathomicNet, MaxNet, TransMIL, TransLayer, PPEG, MultiheadAttention, BilinearFusion, CMTA, MCAT_Surv,
    Transformer_P, Transformer_G, DeformCrossTransMIL}, utils.sync_batchnorm.{convert_model, SynchronizedBatchNorm1d}.
The reference's factory, train / eval loops and CLI stay the reference's code.  `install()` returns a report
{"modules": [...], "patched": {module: [names]}, "failed": {what: reason}}; failures are logged, never swallowed.
"""
import importlib
import logging
import sys
import types

log = logging.getLogger("sml_b200.install")


def _patch(report, modname, **names):
    """setattr the given names on an importable reference module; record what happened"""
    try:
        mod = importlib.import_module(modname)
    except Exception as e:                                   # reference not on sys.path, or one of its imports is missing
        report["failed"][modname] = "%s: %s" % (type(e).__name__, e)
        log.warning("sml_b200.install: could not import %s (%s: %s) - its classes were not swapped", modname, type(e).__name__, e)
        return None
    for k, v in names.items():
        setattr(mod, k, v)
    report["patched"].setdefault(modname, []).extend(sorted(names))
    return mod


def install(replace_mil=True):
    from .models import NystromAttention as na, DeformableAttention1D as d1, DeformableAttention2D as d2
    from .models import MultiheadAttention as mha, DeformCrossTransMIL as dct, mil, cmta, fusion, model
    from .utils import loss, gather, sync_batchnorm

    report = {"modules": [], "patched": {}, "failed": {}}
    pkg = types.ModuleType("nystrom_attention")
    pkg.NystromAttention = na.NystromAttention
    sys.modules["nystrom_attention"] = pkg
    table = {
        "models.NystromAttention": na, "models.DeformableAttention1D": d1, "models.DeformableAttention2D": d2,
        "models.MultiheadAttention": mha, "models.DeformCrossTransMIL": dct,
        "utils.loss": loss, "utils.gather": gather,
    }
    for name, mod in table.items():
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        try:                                                 # `import utils.loss; utils.loss.X` needs the attribute on the package
            setattr(importlib.import_module(parent), leaf, mod)
        except Exception:                                    # reference tree not importable: the sys.modules entries still serve
            pass                                             # `from utils.loss import BatchLoss` once it is
    report["modules"] = sorted(table) + ["nystrom_attention"]

    # classes swapped inside reference modules that also hold code outside the hot path
    if replace_mil:
        _patch(report, "models.mil", TransMIL=mil.TransMIL, TransLayer=mil.TransLayer, PPEG=mil.PPEG,
               NystromAttention=na.NystromAttention)
    _patch(report, "models.cmta_utils", NystromAttention=na.NystromAttention, TransLayer=cmta.TransLayer, PPEG=cmta.PPEG,
           Transformer_P=cmta.Transformer_P, Transformer_G=cmta.Transformer_G, MultiheadAttention=mha.MultiheadAttention,
           OrthogonalLoss=loss.OrthogonalLoss)
    _patch(report, "models.fusion", BilinearFusion=fusion.BilinearFusion)
    _patch(report, "utils.sync_batchnorm", convert_model=sync_batchnorm.convert_model,
           SynchronizedBatchNorm1d=sync_batchnorm.SynchronizedBatchNorm1d)
    # the factory define_net() (models/model.py:49-79) resolves class names in its own module namespace (`from .mil import *`
    # etc. copied them there at import time), so the names are swapped there too
    names = dict(DeformPathomicNet=model.DeformPathomicNet, MaxNet=model.MaxNet, DeformCrossTransMIL=dct.DeformCrossTransMIL,
                 MultiheadAttention=mha.MultiheadAttention, BilinearFusion=fusion.BilinearFusion, CMTA=cmta.CMTA,
                 MCAT_Surv=cmta.MCAT_Surv, Transformer_P=cmta.Transformer_P, Transformer_G=cmta.Transformer_G)
    if replace_mil:
        names.update(TransMIL=mil.TransMIL, TransLayer=mil.TransLayer, PPEG=mil.PPEG)
    _patch(report, "models.model", **names)
    return report
