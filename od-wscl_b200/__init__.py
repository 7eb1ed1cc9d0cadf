"""odwscl_b200 -- B200-native (sm_100a) implementation of OD-WSCL's proposal-feature hot path.

Layout (only what the path needs):
  csrc/        hand-written CUDA kernels + the C ABI (include/odwscl.h) -> lib/libodwscl_sm100.so
  capi.py      ctypes binding (raw device pointers + stream through the C ABI)
  _C.py        mirror of the reference's pybind module `wetectron._C` (csrc/vision.cpp:9-24)
  layers.py, structures.py, config.py, modeling/   host-side mirror of the reference's operator
               surface for this path (same names, argument meaning, state-dict keys)
  wetectron_shim.py   installs the above under the `wetectron.*` names so reference code imports them
"""
__version__ = "0.1.0"
