"""Import-time shims that let the UNMODIFIED reference Python (/root/reference/wetectron)
be imported in the build container (torch 2.11, CPU only) so that golden vectors can be
generated from the reference itself.  TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py);
the GPU box has no /root/reference, so nothing at run time depends on this file.

None of the shims touches arithmetic on the hot path, with one documented exception:
`wetectron._C.roi_pool_forward/backward` -- the reference has NO CPU ROIPool
(csrc/ROIPool.h:23,44 raise "Not implemented on the CPU"), so the stand-in routes to
torch.ops.torchvision.roi_pool / _roi_pool_backward (same Caffe2 lineage as
csrc/cuda/ROIPool_cuda.cu:16-108).  `_C.nms` / `_C.roi_align_forward` use the reference's own
CPU extension when oracle/_ref was built (oracle/build_ref.py).
"""
import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("ODWSCL_REFERENCE", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class CfgNode(dict):
    """Minimal yacs.config.CfgNode: attribute dict + merge_from_file/list + freeze."""

    def __init__(self, init=None, *a, **k):
        super().__init__()
        self.__dict__["_frozen"] = False
        for kk, v in (init or {}).items():
            self[kk] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)

    def freeze(self):
        pass

    def defrost(self):
        pass

    def dump(self, **k):
        return str(dict(self))

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and k in self and isinstance(self[k], CfgNode):
                self[k]._merge(v)
            else:
                old = self.get(k)
                if isinstance(v, str) and old is not None and not isinstance(old, str):
                    import ast
                    try:
                        v = ast.literal_eval(v)      # yacs decodes "(0.125,)" style YAML strings
                    except Exception:
                        pass
                if isinstance(old, tuple) and isinstance(v, list):
                    v = tuple(v)
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_list(self, lst):
        import ast
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except Exception:
                    pass
            node[parts[-1]] = v


def install():
    if "wetectron" in sys.modules:
        return
    # (1) torch._six
    torch._six = _mod("torch._six", PY3=True, PY37=True, string_classes=(str,), int_classes=(int,))
    # (2) apex.amp (identity at O0)
    amp = _mod("apex.amp", float_function=lambda f: f, half_function=lambda f: f,
               initialize=lambda model, opt=None, **k: (model, opt), init=lambda *a, **k: None)

    @contextlib.contextmanager
    def scale_loss(loss, optimizer, **k):
        yield loss
    amp.scale_loss = scale_loss
    _mod("apex", amp=amp)
    # (3) yacs
    _mod("yacs"); _mod("yacs.config", CfgNode=CfgNode)
    # (4) fvcore / pycocotools / matplotlib stubs
    _mod("fvcore"); _mod("fvcore.nn")
    _mod("fvcore.nn.weight_init", c2_msra_fill=lambda m: None, c2_xavier_fill=lambda m: None)
    _mod("pycocotools"); _mod("pycocotools.mask"); _mod("pycocotools.coco", COCO=object)
    _mod("pycocotools.cocoeval", COCOeval=object)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _mod("matplotlib"); _mod("matplotlib.pyplot")
    # (5) torch.hub private names (utils/model_zoo.py:9-16)
    import torch.hub as hub
    for n in ("_download_url_to_file", "urlparse", "HASH_REGEX"):
        if not hasattr(hub, n):
            setattr(hub, n, None)
    # (7) Tensor.cuda -> identity on CPU (sim_loss.py:72)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    # (6) wetectron._C stand-in
    sys.path.insert(0, REF_ROOT)
    import torchvision  # noqa: F401

    def roi_pool_forward(inp, rois, scale, ph, pw):
        out, arg = torch.ops.torchvision.roi_pool(inp, rois, scale, ph, pw)
        return out, arg

    def roi_pool_backward(grad, inp, rois, argmax, scale, ph, pw, b, c, h, w):
        return torch.ops.torchvision._roi_pool_backward(grad.contiguous(), rois, argmax, scale, ph, pw, b, c, h, w)

    def _unsupported(*a, **k):
        raise RuntimeError("not available in the oracle shim")

    ref_c = None
    try:
        from . import build_ref
        ref_c = build_ref.load()
    except Exception:
        ref_c = None
    c = types.ModuleType("wetectron._C")
    c.roi_pool_forward = roi_pool_forward
    c.roi_pool_backward = roi_pool_backward
    for n in ("nms", "roi_align_forward", "roi_align_backward", "sigmoid_focalloss_forward",
              "sigmoid_focalloss_backward", "deform_conv_forward", "deform_conv_backward_input",
              "deform_conv_backward_parameters", "modulated_deform_conv_forward",
              "modulated_deform_conv_backward", "deform_psroi_pooling_forward",
              "deform_psroi_pooling_backward"):
        setattr(c, n, getattr(ref_c, n) if ref_c is not None and hasattr(ref_c, n) else _unsupported)
    import importlib
    pkg = importlib.import_module("wetectron")
    sys.modules["wetectron._C"] = c
    pkg._C = c
    return pkg
