#!/usr/bin/env python
"""Latency of one forward (B target views per call) eager vs CUDA-graph replay (development tool)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.graphed import GraphedForward
from gdb_nerf_b200.network import Network
from gdb_nerf_b200.synthetic import WORKLOADS, batch_to, workload_batch

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="dtu"); ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
cfg = make_cfg(WORKLOADS[args.workload]["recipe"])
torch.manual_seed(0)
net = Network(cfg).cuda().eval()
for B in (1, 2, 8):
    batch = batch_to(workload_batch(args.workload, B=B), "cuda")
    with torch.no_grad():
        for _ in range(5): net(batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(args.iters): net(batch)
    torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / args.iters * 1e3
    runner = GraphedForward(net, batch)
    for _ in range(3): runner(batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(args.iters): runner(batch)
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / args.iters * 1e3
    print(f"{args.workload} B={B}: eager {eager:.3f} ms/call ({eager / B:.3f} ms/view), graph {graph:.3f} ms/call ({graph / B:.3f} ms/view)")
