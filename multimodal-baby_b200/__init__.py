"""multimodal-baby_b200: B200-native (sm_100a) implementation of the CVCL contrastive hot path
of wkvong/multimodal-baby behind the reference's MultiModalModel / MultiModalLitModel API.

The directory name carries a hyphen (it mirrors the upstream repository name), so the package
is imported as `multimodal_baby_b200` through the shim module at the repository root.

    from multimodal_baby_b200 import MultiModalModel, MultiModalLitModel, ops

`ops` binds libcvcl_b200.so (hand-written CUDA, C ABI in include/cvcl_b200.h) with ctypes and
registers `cvcl_b200::*` torch.library ops.  No CPU fallback exists: the ops raise on CPU tensors
and `CvclLibraryMissing` if the library was not built (`python multimodal-baby_b200/build.py`).
"""
from . import _cabi, attention_maps, build, ops, sharding, staging  # noqa: F401
from ._cabi import CvclError, CvclLibraryMissing                   # noqa: F401
from .multimodal import (MultiModalModel, PooledTrunk, TextEncoder, VisionEncoder,  # noqa: F401
                         split_trunk_forward)
from .graphed import GraphedContrastiveStep, GraphedLossStep        # noqa: F401
from .optim import FusedAdamW                                       # noqa: F401
from .staging import PinnedBatchStager, batch_trials, multiModalDataset_collate_fn  # noqa: F401
from .multimodal_lit import MultiModalLitModel, WhitespaceTokenizer, load_vocab      # noqa: F401

__version__ = "0.1.0"
