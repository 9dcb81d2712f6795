"""Import shim: the package directory is named `pbrt-r3_b200/` (not a Python identifier), so `import pbrt_r3_b200`
loads it from there."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pbrt-r3_b200")
_spec = importlib.util.spec_from_file_location("pbrt_r3_b200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_m = importlib.util.module_from_spec(_spec)
sys.modules["pbrt_r3_b200"] = _m
_spec.loader.exec_module(_m)

if __name__ == "__main__":   # `python -m pbrt_r3_b200 scene.pbrt ...`: the reference's command line (pbrt-r3_b200/cli.py)
    from pbrt_r3_b200.cli import main
    sys.exit(main())
