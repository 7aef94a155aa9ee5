"""Imports the hyphen-named package directory under a valid module name."""
import importlib.util
import os
import sys

MODULE_NAME = "rsoderh_raytracing_b200"
PACKAGE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rsoderh-raytracing_b200")
REPO_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if MODULE_NAME in sys.modules:
        return sys.modules[MODULE_NAME]
    spec = importlib.util.spec_from_file_location(
        MODULE_NAME, os.path.join(PACKAGE_DIR, "__init__.py"),
        submodule_search_locations=[PACKAGE_DIR])
    module = importlib.util.module_from_spec(spec)
    sys.modules[MODULE_NAME] = module
    spec.loader.exec_module(module)
    return module
