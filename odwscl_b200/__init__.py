"""Import alias: the package lives in `od-wscl_b200/` (not a valid Python identifier), this module
makes it importable as `odwscl_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "od-wscl_b200")
__path__[:] = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
