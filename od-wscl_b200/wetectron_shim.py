"""Drop-in installation: expose this package's implementations under the reference's module names
so that reference code (`from wetectron import _C`, `from wetectron.layers import ROIPool`, ...)
binds to the sm_100a path.  See INTEGRATION.md.

    import odwscl_b200.wetectron_shim as shim; shim.install()        # before importing wetectron.*

What install() does, in order:

1. registers ``sys.modules["wetectron._C"]`` = ``odwscl_b200._C`` (the 14 pybind names of csrc/vision.cpp:9-24);
2. compatibility shims the reference's own imports need on a current toolchain, each ONLY when the real module is
   absent: ``apex.amp`` at opt-level O0 (tools/train_net.py:33-36 raises without it; layers/roi_pool.py:9,55 and
   engine/trainer.py:10 import it) and ``torch._six`` (utils/imports.py:8, removed from torch >= 2.0);
3. if the REAL ``wetectron`` package is importable (the reference checkout is on sys.path), nothing else is replaced:
   the reference's own ``wetectron/layers/*.py`` import ``wetectron._C`` and thereby execute against the sm_100a
   kernels, and ``wetectron.layers`` keeps Conv2d / FrozenBatchNorm2d / the DCN modules the rest of the reference
   imports.  Only when the package is NOT importable a minimal stand-in ``wetectron`` / ``wetectron.layers`` (ROIPool,
   ROIAlign, nms, smooth_l1_loss from odwscl_b200.layers) is created, so code written against those names still runs.
"""
import contextlib
import importlib
import importlib.util
import sys
import types

real_package_found = None          # set by install(): True when the reference's own `wetectron` package was bound


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_apex_o0():
    """`apex.amp` at opt-level O0 = fp32 everywhere: float_function / half_function are the identity, initialize returns
    its arguments, scale_loss yields the loss unscaled (what apex itself does at O0 with loss_scale 1.0).  Installed only
    when apex is not importable."""
    try:
        import apex.amp  # noqa: F401
        return False
    except Exception:
        pass

    def initialize(models, optimizers=None, enabled=True, opt_level="O0", **kw):
        if opt_level not in ("O0", None) and enabled:
            raise RuntimeError("the odwscl_b200 apex stand-in only implements opt_level O0 (fp32); install NVIDIA apex "
                               "for %r" % (opt_level,))
        return models if optimizers is None else (models, optimizers)

    @contextlib.contextmanager
    def scale_loss(loss, optimizers, **kw):
        yield loss

    def _identity_decorator(fn):
        return fn

    amp = _module("apex.amp", initialize=initialize, scale_loss=scale_loss, float_function=_identity_decorator,
                  half_function=_identity_decorator, promote_function=_identity_decorator,
                  init=lambda *a, **k: None, state_dict=lambda: {}, load_state_dict=lambda sd: None,
                  master_params=lambda opt: (p for g in opt.param_groups for p in g["params"]))
    apex = sys.modules.get("apex") or _module("apex")
    apex.amp = amp
    return True


def install_torch_six():
    import torch
    try:
        import torch._six  # noqa: F401
        return False
    except Exception:
        torch._six = _module("torch._six", PY3=True, PY37=True, string_classes=(str,), int_classes=(int,),
                             container_abcs=importlib.import_module("collections.abc"))
        return True


def install(replace_layers: bool = True, compat: bool = True):
    global real_package_found
    from . import _C, layers
    sys.modules["wetectron._C"] = _C                  # before anything imports wetectron.layers
    if compat:
        install_apex_o0()
        install_torch_six()
    pkg = sys.modules.get("wetectron")
    real = pkg is not None and bool(getattr(pkg, "__file__", None))
    if pkg is None:
        try:
            spec = importlib.util.find_spec("wetectron")
        except (ImportError, ValueError):
            spec = None
        if spec is not None and spec.origin:
            pkg = importlib.import_module("wetectron")        # the reference's package __init__ is empty
            real = True
        else:
            pkg = types.ModuleType("wetectron")
            pkg.__path__ = []
            sys.modules["wetectron"] = pkg
    pkg._C = _C
    real_package_found = real
    if real:
        # a layers module imported BEFORE install() captured the old `_C` by value (`from wetectron import _C`): rebind
        for name in ("wetectron.layers.roi_pool", "wetectron.layers.roi_align", "wetectron.layers.nms"):
            m = sys.modules.get(name)
            if m is not None and hasattr(m, "_C"):
                m._C = _C
        return pkg
    if replace_layers and "wetectron.layers" not in sys.modules:
        mod = types.ModuleType("wetectron.layers")
        for n in ("ROIPool", "roi_pool", "ROIAlign", "roi_align", "nms", "smooth_l1_loss"):
            setattr(mod, n, getattr(layers, n))
        sys.modules["wetectron.layers"] = mod
        pkg.layers = mod
    return pkg
