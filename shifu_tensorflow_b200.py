"""Import alias: the package directory is named `shifu-tensorflow_b200` (a hyphen is not importable), so
`import shifu_tensorflow_b200` loads that directory as a regular package under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shifu-tensorflow_b200")
_spec = importlib.util.spec_from_file_location(
    "shifu_tensorflow_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["shifu_tensorflow_b200"] = _mod
_spec.loader.exec_module(_mod)
