"""Import the UNMODIFIED reference scripts: from /root/reference in the build container, or from the
install made by ``__graft_entry__.build()`` under ``baseline/_ref/`` (git-ignored, travels to the GPU box
with the snapshot like a built .so — the base contract's ``pip install --target baseline/_ref`` for a
reference that is two plain scripts with no packaging metadata).

The scripts import ``timm`` and ``matplotlib``, neither of which is installed and neither of which is
on the parity path: ``DropPath`` is ``nn.Identity`` at the constructor-default
drop_path=0 (C:344) and ``trunc_normal_`` only touches initialisation.  Two ``sys.modules`` stand-ins
let the files load; ``main()`` is never called.  Callers must check ``available()`` first.
"""
import importlib.util
import os
import sys
import types

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALL_DIR = os.path.join(_ROOT, "baseline", "_ref")
FILES = {"cswin": "train_cswinunet_segmentation.py", "unet": "train_unet_segmentation.py"}
_cache = {}


def _find_dir():
    for d in (os.environ.get("CSB200_REFERENCE_DIR"), "/root/reference", INSTALL_DIR):
        if d and os.path.isfile(os.path.join(d, FILES["cswin"])):
            return d
    return None


REFERENCE_DIR = _find_dir() or "/root/reference"


def available() -> bool:
    return _find_dir() is not None


def install(src: str = "/root/reference") -> bool:
    """Copy the reference scripts, byte for byte, to baseline/_ref/ (never into tracked files)."""
    import shutil
    if not os.path.isfile(os.path.join(src, FILES["cswin"])):
        return False
    os.makedirs(INSTALL_DIR, exist_ok=True)
    for f in FILES.values():
        shutil.copyfile(os.path.join(src, f), os.path.join(INSTALL_DIR, f))
    return True


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
    ref_dir = _find_dir()
    if ref_dir is None:
        raise FileNotFoundError("reference scripts not found (/root/reference or baseline/_ref)")
    _install_shims()
    spec = importlib.util.spec_from_file_location(f"_csb200_ref_{which}", os.path.join(ref_dir, FILES[which]))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[which] = mod
    return mod
