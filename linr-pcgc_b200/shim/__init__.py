"""Drop-in modules for the two un-vendored natives the reference's hot path imports.

    import linr_pcgc_b200.shim as shim
    shim.install()                      # registers `MinkowskiEngine` and `torchac` in sys.modules
    import MinkowskiEngine as ME        # -> linr_pcgc_b200.shim.MinkowskiEngine (sm_100a kernels behind the C ABI)
    import torchac                      # -> linr_pcgc_b200.shim.torchac       (csrc/rc_host.cpp)

With the shims installed, the reference's own `models/model_core.py`, `models/upsample.py`, `models/resnet.py`,
`models/function_utils.py` and `models/module_utils.py` run unmodified on a B200 (the L2->L1 edge).
This is the slow, layer-by-layer way in; `linr_pcgc_b200.model.LINR_PCGC_Model` is the fused path behind the same
model-level interface.  GPU only: there is no CPU fallback.
"""
import sys


def install(force: bool = False) -> None:
    from . import MinkowskiEngine as _me, torchac as _ta
    for name, mod in (("MinkowskiEngine", _me), ("torchac", _ta)):
        if force or name not in sys.modules:
            sys.modules[name] = mod
