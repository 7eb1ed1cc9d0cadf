"""Drop-in installation: expose this package's implementations under the reference's module names
so that reference code (`from wetectron import _C`, `from wetectron.layers import ROIPool`, ...)
binds to the sm_100a path.  See INTEGRATION.md.

    import odwscl_b200.wetectron_shim as shim; shim.install()        # before importing wetectron.*
"""
import sys
import types


def install(replace_layers: bool = True):
    from . import _C, layers
    pkg = sys.modules.get("wetectron")
    if pkg is None:
        try:
            import wetectron as pkg          # the real reference package, if importable
        except Exception:
            pkg = types.ModuleType("wetectron")
            pkg.__path__ = []
            sys.modules["wetectron"] = pkg
    sys.modules["wetectron._C"] = _C
    pkg._C = _C
    if replace_layers and "wetectron.layers" not in sys.modules:
        mod = types.ModuleType("wetectron.layers")
        for n in ("ROIPool", "roi_pool", "ROIAlign", "roi_align", "nms", "smooth_l1_loss"):
            setattr(mod, n, getattr(layers, n))
        sys.modules["wetectron.layers"] = mod
        pkg.layers = mod
    return pkg
