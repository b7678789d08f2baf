"""ORACLE -- test / baseline infrastructure, not product code (see oracle/crf_oracle.py for who may import it).

Imports the UNMODIFIED reference (src/newcrf_layers.py, src/model_mobileV3_large_newCRFs.py) when its sources are
reachable -- /root/reference/src in the build container, baseline/_ref/src if a driver put a copy there -- with the
shims SURVEY.md Appendix A documents:
  * `timm.models.layers` is absent: DropPath / to_2tuple / trunc_normal_ are used at construction time only, with
    rate 0 (newcrf_layers.py:6,107,184,187);
  * `matplotlib` is absent: src/utils.py imports it for colour maps only;
  * `mobilenet_v3_large(pretrained=True)` needs the network: patched to weights=None while the model is built.
On the GPU box neither directory exists (the Python reference cannot travel): every function returns None there and
the callers fall back to the restatement in oracle/model_oracle.py.
"""
from __future__ import annotations

import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference/src", os.path.join(_ROOT, "baseline", "_ref", "src"))


def reference_src():
    for d in _CANDIDATES:
        if os.path.isfile(os.path.join(d, "newcrf_layers.py")):
            return d
    return None


def _install_shims():
    import torch.nn as nn
    if "timm.models.layers" not in sys.modules:
        tl = types.ModuleType("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, p=0.):
                super().__init__()
                self.p = p

            def forward(self, x):
                return x

        tl.DropPath = DropPath
        tl.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
        tl.trunc_normal_ = nn.init.trunc_normal_
        sys.modules.update({"timm": types.ModuleType("timm"), "timm.models": types.ModuleType("timm.models"),
                            "timm.models.layers": tl})
    for m in ("matplotlib", "matplotlib.cm"):
        sys.modules.setdefault(m, types.ModuleType(m))


def reference_layers():
    """The reference's newcrf_layers module, or None when the sources are not present."""
    src = reference_src()
    if src is None:
        return None
    _install_shims()
    if src not in sys.path:
        sys.path.insert(0, src)
    import newcrf_layers as ref
    return ref


def reference_ptmodel():
    """A zero-argument factory for the reference's PTModel (random-init encoder: no network), or None."""
    src = reference_src()
    if src is None:
        return None
    _install_shims()
    if src not in sys.path:
        sys.path.insert(0, src)

    def make():
        import torchvision.models as tvm
        orig = tvm.mobilenet_v3_large
        tvm.mobilenet_v3_large = lambda *a, **k: orig(weights=None)
        try:
            import model_mobileV3_large_newCRFs as ref_model
            return ref_model.PTModel()
        finally:
            tvm.mobilenet_v3_large = orig
    return make
