"""Imports the reference's own modules (read-only checkout at /root/reference) for golden-vector generation. Third-party
packages that those modules import at the top but that the functions we call never touch are stubbed (they are absent
from this image); torch's `.cuda()` is made a no-op so that range_projection's GPU round trip
(augmentation_3d.py:262-270) runs on the CPU. Used ONLY by the make_*_golden.py scripts, never by tests or product code."""
import importlib
import sys
import types

REFERENCE = "/root/reference"
_STUBS = ("torchsparse", "torchsparse.utils", "torchsparse.utils.quantize", "torchsparse.utils.collate", "pypatchworkpp",
          "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "open3d", "cv2", "PIL", "PIL.Image", "yacs", "yacs.config",
          "tqdm", "scipy.ndimage")


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


def load(module):
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    for name in _STUBS:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Anything(name)
    import numpy as np
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_  # removed in numpy 2; augmentation_3d.py:248,277 still uses it
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU box: keep the reference's code path, drop the device hop
    return importlib.import_module(module)
