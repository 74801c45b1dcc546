"""Registers the package directory `block2-preview_b200/` (not an importable identifier)
as the module `block2_preview_b200`.  Usage: `import b2gpkg; b2g = b2gpkg.load()`."""
import importlib.util
import os
import sys

_NAME = "block2_preview_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "block2-preview_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(root, "__init__.py"),
                                                  submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
