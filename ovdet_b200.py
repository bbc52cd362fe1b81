"""Import alias: ``import ovdet_b200`` loads the package that lives in the
directory ``open-vocabulary-3d-object-detection_b200/`` (a hyphenated name is not
importable directly).  Sub-modules resolve normally afterwards, e.g.
``from ovdet_b200.utils.box_util import generalized_box3d_iou``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "open-vocabulary-3d-object-detection_b200")
_spec = importlib.util.spec_from_file_location(
    "ovdet_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ovdet_b200"] = _mod
_spec.loader.exec_module(_mod)
