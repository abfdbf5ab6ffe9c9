"""Import alias: ``import clasfv_b200`` loads the package that lives in
``fully-automated-multi-heartbeat-echocardiography-video-segmentation-and-motion-tracking_b200/``
(the directory name the project layout prescribes is not a valid Python identifier)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                     "fully-automated-multi-heartbeat-echocardiography-video-segmentation-and-motion-tracking_b200")
_spec = _ilu.spec_from_file_location("clasfv_b200", _os.path.join(_DIR, "__init__.py"),
                                     submodule_search_locations=[_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["clasfv_b200"] = _mod
_spec.loader.exec_module(_mod)
