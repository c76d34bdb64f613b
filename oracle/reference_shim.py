"""Import the UNMODIFIED reference scripts from /root/reference (build container only).

The scripts import ``timm`` and ``matplotlib``, neither of which is installed and neither of which is
on the parity path: ``DropPath`` is ``nn.Identity`` at the constructor-default
drop_path=0 (C:344) and ``trunc_normal_`` only touches initialisation.  Two ``sys.modules`` stand-ins
let the files load; ``main()`` is never called.  The GPU box has no /root/reference: callers must
check ``available()`` first.
"""
import importlib.util
import os
import sys
import types

import torch

REFERENCE_DIR = os.environ.get("CSB200_REFERENCE_DIR", "/root/reference")
_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "train_cswinunet_segmentation.py"))


class _DropPath(torch.nn.Module):
    """Stochastic depth per sample, as timm defines it (only used when drop_path > 0)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def _install_shims():
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})


def load(which: str):
    """which: 'cswin' -> train_cswinunet_segmentation, 'unet' -> train_unet_segmentation."""
    if which in _cache:
        return _cache[which]
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    _install_shims()
    fname = {"cswin": "train_cswinunet_segmentation.py", "unet": "train_unet_segmentation.py"}[which]
    spec = importlib.util.spec_from_file_location(f"_csb200_ref_{which}", os.path.join(REFERENCE_DIR, fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[which] = mod
    return mod
