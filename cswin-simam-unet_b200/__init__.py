"""B200-native (sm_100a) CSWin-SimAM-UNet training hot path.

Drop-in for the ``nn.Module`` surface of TrungMasterChef/CSWin-SimAM-UNet:

    from cswin_simam_unet_b200 import CSWinTransformer, UNet, CSWinBlock, LePEAttention, SimAM

SimAM and the cross-shaped stripe attention (with LePE) run as hand-written CUDA kernels behind the
C ABI in ``include/csb200.h`` (``libcsb200.so``); there is no Triton, no multi-backend dispatch and no
CPU fallback.  See DESIGN.md.
"""
from . import capi, functional  # noqa: F401
from .functional import cross_stripe_attention, simam, stripe_attention  # noqa: F401
from .models import CSWinTransformer, DoubleConv, Down, UNet, Up  # noqa: F401
from .modules import (CARAFE, CARAFE4, CSWinBlock, DropPath, LePEAttention, Merge_Block, Mlp,  # noqa: F401
                      SimAM)
from .train import InferStep, TrainStep, bce_from_logits_as_probabilities, synthetic_batch  # noqa: F401
from .data_parallel import GradientAllReducer, shard_of_global_batch  # noqa: F401
from .optim import FusedAdamW, fused_adam  # noqa: F401

__version__ = "0.1.0"
