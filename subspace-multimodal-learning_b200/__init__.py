"""sml_b200 — B200-native (sm_100a) kernels behind the Subspace-Multimodal-Learning module signatures.

The package directory is ``subspace-multimodal-learning_b200`` (not an importable identifier); import it
through the repo-root shim ``sml_b200`` or ``__graft_entry__.load_package()``.
"""
from . import _lib, ops, functional, nystrom, deform  # noqa: F401
from .models.NystromAttention import NystromAttention, moore_penrose_iter_pinv  # noqa: F401
from .models.mil import TransLayer, PPEG, TransMIL  # noqa: F401
from .models.DeformableAttention1D import DeformCrossAttention1D  # noqa: F401
from .models.DeformableAttention2D import DeformCrossAttention2D  # noqa: F401
from .models.DeformCrossTransMIL import DeformCrossTransMIL, FusionNet, DeformCrossTransLayer, Pooler  # noqa: F401
from .models.MultiheadAttention import MultiheadAttention  # noqa: F401
from .models.fusion import BilinearFusion  # noqa: F401
from .models.model import MaxNet, DeformPathomicNet  # noqa: F401
from .models.cmta import CMTA, MCAT_Surv, Transformer_P, Transformer_G  # noqa: F401
from .utils.loss import BatchLoss, OrthogonalLoss  # noqa: F401
from .utils.gather import GatherLayer, PackedGather  # noqa: F401
from .utils.sync_batchnorm import SyncBatchNorm1d, SynchronizedBatchNorm1d, convert_model  # noqa: F401
from .utils.modulation import modulate_classifier_grads  # noqa: F401
from .install import install  # noqa: F401
from . import data  # noqa: F401
from .graph import GraphedStep, GraphedDataParallelStep  # noqa: F401
