"""Import alias: ``import tskd_b200`` loads the package in ``time-series-kafka-demo_b200/``.

The package directory keeps the repo's name (hyphens and all), which Python cannot import
directly; this shim registers it under the importable name ``tskd_b200``.
"""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "time-series-kafka-demo_b200")
_spec = importlib.util.spec_from_file_location(
    "tskd_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tskd_b200"] = _mod
_spec.loader.exec_module(_mod)
