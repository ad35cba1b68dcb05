"""Hot-path configuration recipes.

The reference builds a ``SimpleNamespace`` tree from YAML at import time
(configs/config.py:54-104).  Only three sections are read on the rendering
path: ``fpn``, ``mvs`` and ``nerf``.  The reference yamls
do not travel to the GPU box, so the shipped recipes are restated here as plain
dicts (values from configs/dtu_pretrain.yaml:17-42, dtu_eval.yaml:3-13,
llff_eval.yaml:7-28, nerf_eval.yaml:6-24 and the 4x4 comments at
dtu_pretrain.yaml:23-24,33).
"""
from __future__ import annotations

import copy
from types import SimpleNamespace
from typing import Any, Dict, Iterable, Tuple

_BASE: Dict[str, Any] = {
    "fpn": {"base_channels": 8, "feat_dims": [32, 16, 8], "feat_scales": [0.25, 0.5, 1.0]},
    "mvs": {
        "vol_levels": [0, 1],
        "vol_scales": [0.125, 0.5],
        "ci_scales": [1.0, 1.0],
        "voxel_dim": 8,
        "num_depth": [64, 8],
        "inv_depth": [True, False],
        "num_samples": [8],
        "loss_weight": [0.05],
    },
    "nerf": {
        "bundle_size": 2,
        "global_num_depth": 64,
        "max_num_samples": 6,
        "max_mipmap_level": 3,
        "nerf_hidden_dims": 64,
        "chunk_size": 1000000,
        "is_adaptive": False,
        "viewdir_agg": True,
        "dec_layers": 3,
        "reweighting": False,
    },
}

_RECIPES: Dict[str, Dict[str, Dict[str, Any]]] = {
    # configs/dtu_pretrain.yaml (training: fixed 6 samples / bundle)
    "dtu_pretrain": {},
    # configs/dtu_eval.yaml
    "dtu_eval": {"nerf": {"max_num_samples": 3, "is_adaptive": True, "reweighting": False}},
    # configs/llff_eval.yaml
    "llff_eval": {
        "mvs": {"num_depth": [36, 8]},
        "nerf": {"max_num_samples": 3, "is_adaptive": True, "reweighting": True},
    },
    # configs/nerf_eval.yaml
    "nerf_eval": {"nerf": {"max_num_samples": 6, "is_adaptive": True, "reweighting": True}},
    # nerf_eval + the documented 4x4 bundle recipe
    "nerf_eval_4x4": {
        "mvs": {"vol_levels": [0, 0], "vol_scales": [0.125, 0.25]},
        "nerf": {"max_num_samples": 6, "is_adaptive": True, "reweighting": True, "bundle_size": 4},
    },
}


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def to_namespace(d: Dict[str, Any]) -> SimpleNamespace:
    ns = SimpleNamespace()
    for k, v in d.items():
        setattr(ns, k, to_namespace(v) if isinstance(v, dict) else v)
    return ns


def recipe_dict(name: str, overrides: Iterable[Tuple[str, Any]] = ()) -> Dict[str, Any]:
    if name not in _RECIPES:
        raise KeyError(f"unknown recipe {name!r}; have {sorted(_RECIPES)}")
    d = _merge(copy.deepcopy(_BASE), _RECIPES[name])
    for dotted, value in overrides:
        node = d
        keys = dotted.split(".")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = value
    return d


def make_cfg(name: str = "dtu_eval", overrides: Iterable[Tuple[str, Any]] = ()) -> SimpleNamespace:
    """``cfg`` with the same attribute tree the reference's ``Network(cfg)`` reads."""
    return to_namespace(recipe_dict(name, overrides))
