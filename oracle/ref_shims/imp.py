"""Stand-in for the stdlib ``imp`` module (removed in Python 3.12), which the
reference's factories use (networks/make_network.py:2,8).  Test infrastructure
only: lets the golden generator import the unmodified reference."""
import importlib.util
import sys


def load_source(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module
