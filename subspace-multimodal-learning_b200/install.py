"""Make the reference's own `main.py` / `train_test.py` / `inference.py` run on the B200 kernels without editing
them: register the drop-in modules under the module names the reference imports.

    import sys; sys.path.insert(0, "/path/to/Subspace-Multimodal-Learning"); sys.path.insert(0, "/path/to/this/repo")
    import sml_b200; sml_b200.install()          # before `import models.model`
    import main                                  # the reference, unchanged

What is replaced:  nystrom_attention.NystromAttention (models/mil.py:24, DeformCrossTransMIL.py:24),
models.NystromAttention, models.DeformableAttention1D / 2D, models.MultiheadAttention, models.DeformCrossTransMIL,
models.fusion.BilinearFusion, models.model.DeformPathomicNet / MaxNet, utils.loss.BatchLoss, utils.gather.GatherLayer, utils.sync_batchnorm.convert_model.
The reference's factory (models/model.py), train/eval loops and CLI stay the reference's code.
"""
import sys
import types


def install(replace_mil=True):
    from . import models as _models
    from .models import NystromAttention as na, DeformableAttention1D as d1, DeformableAttention2D as d2
    from .models import MultiheadAttention as mha, DeformCrossTransMIL as dct, mil
    from .utils import loss, gather, sync_batchnorm

    pkg = types.ModuleType("nystrom_attention")
    pkg.NystromAttention = na.NystromAttention
    sys.modules["nystrom_attention"] = pkg
    table = {
        "models.NystromAttention": na, "models.DeformableAttention1D": d1, "models.DeformableAttention2D": d2,
        "models.MultiheadAttention": mha, "models.DeformCrossTransMIL": dct,
        "utils.loss": loss, "utils.gather": gather,
    }
    if replace_mil:
        table["models.mil_b200"] = mil           # TransMIL / TransLayer / PPEG drop-ins (ABMIL stays the reference's)
    for name, mod in table.items():
        sys.modules[name] = mod
    # BilinearFusion: keep the reference module object (it also holds the unused Trilinear* classes) and swap the class
    try:
        import models.fusion as ref_fusion
        from .models.fusion import BilinearFusion
        ref_fusion.BilinearFusion = BilinearFusion
    except Exception:
        pass
    try:
        import utils.sync_batchnorm as ref_sbn
        ref_sbn.convert_model = sync_batchnorm.convert_model
        ref_sbn.SynchronizedBatchNorm1d = sync_batchnorm.SynchronizedBatchNorm1d
    except Exception:
        pass
    # DeformPathomicNet / MaxNet (models/model.py:142-187, 440-544): the factory define_net() looks the classes up in its own
    # module namespace, so swap them there (needs the reference importable; skipped otherwise)
    try:
        import models.model as ref_model
        from .models.model import DeformPathomicNet, MaxNet
        ref_model.DeformPathomicNet = DeformPathomicNet
        ref_model.MaxNet = MaxNet
    except Exception:
        pass
    return sorted(table)
