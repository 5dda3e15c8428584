"""Import alias: ``import vfr_b200`` loads the package that lives in the directory
``video-fragments-retrieval_b200/`` (a hyphenated directory name is not importable with a
plain ``import`` statement).  After this module runs, ``sys.modules['vfr_b200']`` is the real
package, so ``from vfr_b200 import models, evaluate`` works as usual."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "video-fragments-retrieval_b200")
_spec = importlib.util.spec_from_file_location(
    "vfr_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["vfr_b200"] = _pkg
_spec.loader.exec_module(_pkg)
