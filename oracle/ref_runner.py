"""Runs the UNMODIFIED reference model - TEST / BENCHMARK INFRASTRUCTURE (only tests/, smoke() and bench.py's reference legs
may import this; the product path never does).

The reference is pure Python with no setup.py: "installing" it is staging a copy of its tree under the git-ignored
``baseline/_ref/reference`` (recipe: baseline/README.md), from where it travels to the GPU box with the repository snapshot.
It is imported from there (or from $GDB_REFERENCE, or /root/reference in the build container) with the three stand-ins of
``oracle/ref_shims`` on sys.path for modules that do not exist offline (``imp``, ``nvdiffrast``, ``nerfacc``).
"""
from __future__ import annotations

import os
import sys
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_dir() -> Optional[str]:
    for cand in (os.environ.get("GDB_REFERENCE"), os.path.join(ROOT, "baseline", "_ref", "reference"), "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "networks", "gdb_nerf", "network.py")):
            return cand
    return None


def load_reference_network(cfg, device: str = "cpu", seed: int = 0):
    """``networks.gdb_nerf.network.Network(cfg)`` of the reference, default PyTorch initialisation under ``seed`` (the same
    values gdb_nerf_b200.network.Network gets under the same seed: identical module construction order), eval mode."""
    ref = reference_dir()
    if ref is None:
        raise FileNotFoundError("the reference tree is not staged (baseline/README.md)")
    for path in (ref, os.path.join(HERE, "ref_shims")):
        if path not in sys.path:
            sys.path.insert(0, path)
    import networks.gdb_nerf.network as ref_network      # noqa: E402  (the reference's own module)
    torch.manual_seed(seed)
    return ref_network.Network(cfg).to(device).eval()
