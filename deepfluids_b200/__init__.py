"""Import alias: the package lives in `deep-fluids_b200/` (the repo-mandated directory name, which is
not a valid Python identifier); this stub makes it importable as `deepfluids_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "deep-fluids_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
