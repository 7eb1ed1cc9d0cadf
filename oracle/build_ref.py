"""Recipe: compile the reference's OWN CPU extension sources, unmodified, where they lie under
/root/reference/wetectron/csrc (csrc/vision.cpp + csrc/cpu/*.cpp, i.e. setup.py:21-37 without
WITH_CUDA) into oracle/_ref/ (git-ignored, travels with gpurun).  TEST INFRASTRUCTURE ONLY.

Gives the real `_C.nms` (cpu/nms_cpu.cpp) and `_C.roi_align_forward` (cpu/ROIAlign_cpu.cpp) used to
pin the oracle's legacy-NMS and ROIAlign-forward restatements.  ROIPool has no CPU source in the
reference (csrc/ROIPool.h:23,44).  No reference source is copied into this repository.
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("ODWSCL_REFERENCE", "/root/reference")
_mod = None


def load(build_if_missing: bool = True):
    """Return the compiled module (building it when /root/reference is present), else None."""
    global _mod
    if _mod is not None:
        return _mod
    import torch  # noqa: F401
    so = glob.glob(os.path.join(OUT, "odwscl_refC*.so"))
    if so:
        import importlib.util
        spec = importlib.util.spec_from_file_location("odwscl_refC", so[0])
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
        return _mod
    csrc = os.path.join(REF, "wetectron", "csrc")
    if not build_if_missing or not os.path.isdir(csrc):
        return None
    from torch.utils.cpp_extension import load as jit_load
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(csrc, "*.cpp")) + glob.glob(os.path.join(csrc, "cpu", "*.cpp")))
    _mod = jit_load(name="odwscl_refC", sources=srcs, extra_include_paths=[csrc],
                    extra_cflags=["-O2"], build_directory=OUT, verbose=False)
    return _mod


if __name__ == "__main__":
    m = load()
    print("built" if m is not None else "reference not present; nothing built", file=sys.stderr)
