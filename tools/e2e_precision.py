#!/usr/bin/env python
"""End-to-end effect of the MLP arithmetic: Network.forward at full size with every precision, cuDNN in true fp32,
differences of ret['rgb'] / depths against the fp32 SIMT arithmetic (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.network import Network
from gdb_nerf_b200.synthetic import WORKLOADS, batch_to, workload_batch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for wl in sys.argv[1:] or ["dtu"]:
    cfg = make_cfg(WORKLOADS[wl]["recipe"])
    torch.manual_seed(0)
    net = Network(cfg).cuda().eval()
    batch = batch_to(workload_batch(wl, B=2), "cuda")
    outs = {}
    with torch.no_grad():
        for prec in (0, 1, 2):
            net.mlp_precision = prec
            ret, _, _ = net(batch)
            outs[prec] = {k: v.double() for k, v in ret.items()}
    near, far = WORKLOADS[wl]["near"], WORKLOADS[wl]["far"]
    for prec in (1, 2):
        d = (outs[prec]["rgb"] - outs[0]["rgb"]).abs()
        dd = (outs[prec]["nerf_depth"] - outs[0]["nerf_depth"]).abs() / (far - near)
        mse = float(((outs[prec]["rgb"] - outs[0]["rgb"]) ** 2).mean())
        print(f"{wl} precision {prec} vs 0: rgb max {d.max():.2e} mean {d.mean():.2e} mse {mse:.2e} | nerf_depth/range max {dd.max():.2e}")
