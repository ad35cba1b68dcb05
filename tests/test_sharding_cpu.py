"""world_size-2 gloo tests of the N>1 host logic (view sharding, timing reduction, image gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gdb_nerf_b200.sharding import allreduce_gradients, gather_images, max_over_ranks, shard_rows, shard_views


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = shard_views(n_views, rank, world)
        # "render": image content encodes the view id
        local = torch.stack([torch.full((3, 2, 2), float(i)) for i in ids]) if ids else torch.zeros(0, 3, 2, 2)
        elapsed = max_over_ranks(10.0 + rank)               # rank 1 is slower
        full = gather_images(local, ids, n_views)
        dist.barrier()
        q.put((rank, ids, elapsed, None if full is None else full[:, 0, 0, 0].tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_view_sharding_gloo():
    world, n_views = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = [r[1] for r in res]
    assert sorted(ids[0] + ids[1]) == list(range(n_views)) and not set(ids[0]) & set(ids[1])      # disjoint cover
    assert all(r[2] == 11.0 for r in res)                                                        # max over ranks
    assert res[0][3] == [0.0, 1.0, 2.0, 3.0, 4.0] and res[1][3] is None                          # gathered in view order


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 2))
        x = torch.full((5, 4), float(rank + 1))
        net[1](net[0](x)).sum().backward()              # the last layer is unused on every rank; the others differ per rank
        if rank == 1:
            net[2].weight.grad = torch.ones(2, 2)       # ... except that rank 1 has a gradient for its weight
        nbytes = allreduce_gradients(net.parameters())
        # plain lists, not tensors: a tensor travels through the queue as a shared file descriptor that dies with this process
        q.put((rank, nbytes, [p.grad.tolist() for p in net.parameters()]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_gloo():
    """One flat-bucket all-reduce averages the gradients; unused parameters behave as zeros (DDP find_unused_parameters)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 2))
    want = None
    for r in range(world):
        net.zero_grad()
        net[1](net[0](torch.full((5, 4), float(r + 1)))).sum().backward()
        gs = [torch.zeros_like(p) if p.grad is None else p.grad.clone() for p in net.parameters()]
        want = gs if want is None else [a + b for a, b in zip(want, gs)]
    want = [g / world for g in want]
    want[4] = torch.full((2, 2), 0.5)                   # net[2].weight: (0 + 1) / 2
    assert res[0][1] == res[1][1] == 4 * sum(p.numel() for p in net.parameters())
    for r in range(world):
        for got, w in zip(res[r][2], want):
            assert torch.allclose(torch.tensor(got), w, atol=1e-6)


def test_shard_helpers_single_process():
    assert shard_views(8, 3, 8) == [3] and shard_views(200, 7, 8) == list(range(7, 200, 8))
    rows = [shard_rows(256, r, 8) for r in range(8)]
    assert rows[0].start == 0 and rows[-1].stop == 256
    assert all(a.stop == b.start for a, b in zip(rows, rows[1:])) and all(r.start % 8 == 0 for r in rows)
    assert max_over_ranks(3.5) == 3.5


def test_reference_loader_mechanism(tmp_path, monkeypatch):
    """networks/make_network.py:5-9 does imp.load_source(cfg.network_module, cfg.network_path).Network(cfg)."""
    import importlib.util
    import sys
    from conftest import ROOT
    from gdb_nerf_b200.config import make_cfg
    spec = importlib.util.spec_from_file_location("imp_shim", os.path.join(ROOT, "oracle", "ref_shims", "imp.py"))
    imp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(imp)
    name = "gdb_nerf_b200.network"
    saved = sys.modules.pop(name, None)
    try:
        mod = imp.load_source(name, os.path.join(ROOT, name.replace(".", "/") + ".py"))
        net = mod.Network(make_cfg("dtu_eval"))
        assert sum(p.numel() for p in net.parameters()) == 962311
    finally:
        if saved is not None:
            sys.modules[name] = saved


def _tile_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gdb_nerf_b200.sharding import gather_row_tiles
        out = []
        for n_rows in (16, 24):                      # equal tiles (one all-gather) and ragged tiles (8 | 16: broadcasts)
            full = torch.arange(n_rows * 5 * 3, dtype=torch.float32).view(1, n_rows, 5, 3)
            plane = torch.arange(n_rows * 5, dtype=torch.float32).view(1, n_rows, 5) * 2.0
            rows = shard_rows(n_rows, rank, world)
            a, b = torch.full_like(full, -1.0), torch.full_like(plane, -1.0)
            a[:, rows.start: rows.stop] = full[:, rows.start: rows.stop]      # what this rank's kernel launch wrote
            b[:, rows.start: rows.stop] = plane[:, rows.start: rows.stop]
            gather_row_tiles([a, b], n_rows, world)
            out.append(bool(torch.equal(a, full) and torch.equal(b, plane)))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_row_tile_gather_gloo():
    """Image-tile split: after gather_row_tiles every rank holds every bundle row of the view."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, [True, True]), (1, [True, True])]
