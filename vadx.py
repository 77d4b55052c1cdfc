"""Import shim: `import vadx` loads the product package that lives in the
(non-identifier) directory `voice-activity-detection-vad-onnx_b200/`."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "voice-activity-detection-vad-onnx_b200")
_spec = _ilu.spec_from_file_location("vadx", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["vadx"] = _mod
_spec.loader.exec_module(_mod)
