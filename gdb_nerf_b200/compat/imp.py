"""``imp`` for Python >= 3.12, where the stdlib module is gone.

The reference's factories call ``imp.load_source(name, path)`` (networks/make_network.py:2,8,
datasets/make_dataset.py:5, evaluators/make_evaluator.py:1, train/trainers/make_trainer.py:2).
Put this directory on ``PYTHONPATH`` to run the unmodified reference scripts:

    PYTHONPATH=/path/to/repo/gdb_nerf_b200/compat python run.py --type evaluate ...
"""
import importlib.util
import sys


def load_source(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module
