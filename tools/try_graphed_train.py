#!/usr/bin/env python
"""Experiment: the DTU pre-training step eager vs captured as one CUDA graph (same initial state, same batches)."""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gdb_nerf_b200.config import make_cfg
from gdb_nerf_b200.graphed import GraphedTrainStep
from gdb_nerf_b200.network import Network
from gdb_nerf_b200.synthetic import batch_to, make_batch

dev = "cuda"
cfg = make_cfg("dtu_pretrain")
torch.manual_seed(0)
net_a = Network(cfg).to(dev).train()
net_b = copy.deepcopy(net_a)
loss_fn = lambda out: out[0]["rgb"].square().mean() + sum(b.square().mean() for b in out[2])
mk = lambda s: batch_to(make_batch(1, 3, 64, 64, 425.0, 905.0, 1446.0 * 64 / 512.0, seed=s, images="smooth", tilt=0.03), dev)
pa = [p for p in net_a.parameters() if p.requires_grad]
pb = [p for p in net_b.parameters() if p.requires_grad]
opt_a = torch.optim.Adam(pa, lr=5e-4)
opt_b = torch.optim.Adam(pb, lr=5e-4, capturable=True)

def eager(batch):
    opt_a.zero_grad(set_to_none=True)
    loss = loss_fn(net_a(batch))
    loss.backward()
    torch.nn.utils.clip_grad_value_(pa, 40)
    opt_a.step()
    return loss

# identical history on both copies: the graphed stepper runs 3 warm-up steps on example batch 0 (capturing executes nothing)
for _ in range(3):
    eager(mk(0))
from gdb_nerf_b200.sharding import allreduce_gradients
import traceback
try:
    stepper = GraphedTrainStep(net_b, opt_b, mk(0), loss_fn, pb, allreduce=allreduce_gradients if len(sys.argv) > 1 else None)
except Exception:
    traceback.print_exc(); raise
for s in (1, 2, 3):
    la, lb = float(eager(mk(s))), float(stepper(mk(s)))
    print(f"step {s}: eager loss {la:.6f} graph loss {lb:.6f}")
d = max(float((x - y).abs().max()) for x, y in zip(pa, pb))
print("max parameter difference after 6 steps", d)
for name, fn in (("eager", lambda: eager(mk(5))), ("graph", lambda: stepper(mk(5)))):
    b5 = mk(5)
    f = (lambda: eager(b5)) if name == "eager" else (lambda: stepper(b5))
    for _ in range(3): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): f()
    torch.cuda.synchronize(); print(name, (time.perf_counter() - t) / 20 * 1e3, "ms/step")
