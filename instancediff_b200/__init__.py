"""instancediff_b200 -- B200-native (sm_100a) implementation of InstanceDiff's reverse-SDE hot path.

Public surface mirrors the reference (zyc-123/InstanceDiff):
  * ``IRSDE`` / ``SDE``      -- drop-in for ``utils/sde_utils.py`` (same constructor, attributes, methods)
  * ``ConditionalUNet``      -- the drift/noise network callable plugged into ``IRSDE.set_model``
  * ``create_sde`` / ``sample_sharded`` -- testUM.py-style driver helpers (``sampler.py``)
  * ``create_model`` / ``RestorationModel`` -- the ``CLIPDriftModel`` protocol testUM.py drives (``model.py``),
    ``checkpoint`` -- ``{iter}_{label}.pth`` naming and ``load_network`` key clean-up

Everything numerical runs in ``libidiff_sm100.so`` (C ABI: ``include/idiff.h``).  There is no CPU
path: constructing ``ConditionalUNet`` or calling a fused op without the built library raises.
"""
from ._lib import IdiffError, LIB_PATH  # noqa: F401
from .sde import IRSDE, SDE  # noqa: F401
from .unet import ConditionalUNet, param_specs  # noqa: F401
from .sampler import create_sde, gather_shards, sample_sharded, shard_bounds  # noqa: F401
from .model import RestorationModel, create_model  # noqa: F401
from . import checkpoint  # noqa: F401
from . import metrics  # noqa: F401
from .parallel import GradientAllReducer  # noqa: F401
from .metrics import calculate_psnr, calculate_ssim, tensor2img  # noqa: F401  (utils/__init__.py:3 exports them next to IRSDE)

__all__ = ["IRSDE", "SDE", "ConditionalUNet", "param_specs", "create_sde", "sample_sharded", "shard_bounds", "gather_shards",
           "RestorationModel", "create_model", "checkpoint", "metrics", "GradientAllReducer", "calculate_psnr", "calculate_ssim", "tensor2img",
           "IdiffError", "LIB_PATH"]
