"""Host plumbing of gdb_nerf_b200/pipeline.py: the mirrors of the reference's to_cuda / load_network."""
import os

import torch

from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.network import Network
from gdb_nerf_b200.pipeline import _checkpoint_path, load_network, to_cuda


def test_to_cuda_walks_containers_and_keeps_meta():
    batch = {"a": torch.ones(2), "nest": {"b": [torch.zeros(1), torch.ones(1)]}, "meta": {"scene": "scan1", "t": torch.ones(1)}, "n": 3}
    out = to_cuda(batch, "cpu")
    assert out["meta"] is batch["meta"] and out["n"] == 3
    assert torch.equal(out["nest"]["b"][1], torch.ones(1)) and isinstance(out["nest"]["b"], list)


def test_load_network_file_and_directory(tmp_path):
    cfg = make_cfg("dtu_eval")
    torch.manual_seed(1)
    src = Network(cfg)
    torch.manual_seed(2)
    dst = Network(cfg)
    d = tmp_path / "trained_model"
    os.makedirs(d)
    torch.save({"net": src.state_dict(), "epoch": 41}, d / "41.pth")
    torch.save({"net": src.state_dict(), "epoch": 7}, d / "7.pth")
    assert _checkpoint_path(str(d)) == str(d / "41.pth")               # highest epoch when there is no latest.pth
    assert load_network(dst, str(d)) == 42
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    torch.save({"net": src.state_dict()}, d / "latest.pth")
    assert _checkpoint_path(str(d)).endswith("latest.pth")
    assert load_network(dst, str(d)) == 0                              # no 'epoch' key -> start from 0 (net_utils.py:108-111)
    assert load_network(dst, str(d / "7.pth")) == 8                    # a file path is taken as is
    assert load_network(dst, str(tmp_path / "missing")) == 0 and load_network(dst, str(d), resume=False) == 0
