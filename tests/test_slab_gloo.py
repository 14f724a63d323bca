"""World-size-2 gloo tests (CPU) of the slab plumbing: layer partition, halo
selection, neighbour exchange, ghost-padded ordering.  The expectation is
computed from the global particle set in one process with the reference's
bucket arithmetic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aboria_b200 import slab, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _global_sorted(N, S, periodic):
    pos = synth.uniform_positions(N, 3)
    side = 1.0 / S
    inv = 1.0 / side
    v = np.floor((pos - 0.0) * inv).astype(np.int64)
    key = (v[:, 0] * S + v[:, 1]) * S + v[:, 2]
    order = np.argsort(key, kind="stable")
    return pos, v[:, 0], order


def _worker(rank, world, port, N, S, w, periodic, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, layer, order = _global_sorted(N, S, periodic)
        b = synth.vector(N)
        layers = slab.plan_layers(S, world)
        lo, hi = layers[rank]
        sorted_layer = layer[order]
        mine = order[(sorted_layer >= lo) & (sorted_layer < hi)]  # owned, globally sorted order
        own_layers = sorted_layer[(sorted_layer >= lo) & (sorted_layer < hi)]
        layer_offsets = torch.from_numpy(np.searchsorted(own_layers, np.arange(lo, hi + 1), side="left").astype(np.int64))
        ex = slab.SlabExchange(rank, world, periodic, w, layer_offsets)
        pos_local = ex.assemble(torch.from_numpy(pos[mine]))
        id_local = ex.assemble(torch.from_numpy(mine.astype(np.int64)))
        # expectation: layers lo-w .. hi+w-1 (wrapped when periodic), in sorted order
        want = []
        for u in range(lo - w, hi + w):
            if periodic:
                g = u % S
            elif 0 <= u < S:
                g = u
            else:
                continue
            want.append(order[sorted_layer == g])
        want = np.concatenate(want)
        ok = np.array_equal(id_local.numpy(), want) and np.array_equal(pos_local.numpy(), pos[want])
        # b halo: owned values placed, ghosts filled from the neighbours
        b_local = torch.zeros(ex.n_local, dtype=torch.float64)
        b_local[ex.own_begin:ex.own_end] = torch.from_numpy(b[mine])
        ex.fill_halo(b_local)
        ok = ok and np.array_equal(b_local.numpy(), b[want])
        ok = ok and ex.n_local == len(want) and ex.n_own == len(mine)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("periodic,S,w", [(True, 12, 1), (True, 12, 2), (False, 12, 1), (True, 9, 2)])
def test_slab_exchange_world2(periodic, S, w):
    world, N = 2, 6000
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, N, S, w, periodic, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_plan_layers_and_halo_width():
    assert slab.plan_layers(294, 8) == [(0, 37), (37, 74), (74, 111), (111, 148), (148, 185), (185, 222), (222, 258), (258, 294)]
    assert sum(hi - lo for lo, hi in slab.plan_layers(147, 4)) == 147
    assert slab.halo_width(1.0 / 294, 1.0 / 294) == 1
    assert slab.halo_width(1.09 / 294, 1.0 / 294) == 2
    with pytest.raises(ValueError):
        slab.SlabExchange(0, 1, False, 3, torch.tensor([0, 5, 9]))


def test_plan_layers_balanced():
    # uniform histogram: the even split; a clustered one: near-equal particle counts, contiguous, every rank >= 1 layer
    assert slab.plan_layers_balanced(np.ones(294), 8) == slab.plan_layers(294, 8) or \
        [hi - lo for lo, hi in slab.plan_layers_balanced(np.ones(294), 8)].count(37) == 6
    rng = np.random.default_rng(3)
    h = rng.integers(1, 50, 117).astype(np.float64)
    h[40:48] += 4000.0  # a blob
    for world in (2, 4, 8):
        plan = slab.plan_layers_balanced(h, world)
        assert plan[0][0] == 0 and plan[-1][1] == 117
        assert all(plan[g][1] == plan[g + 1][0] for g in range(world - 1)) and all(hi > lo for lo, hi in plan)
        counts = [h[lo:hi].sum() for lo, hi in plan]
        even = [h[lo:hi].sum() for lo, hi in slab.plan_layers(117, world)]
        assert max(counts) <= max(even)  # never worse than the split by layer count
    with pytest.raises(ValueError):
        slab.plan_layers_balanced(np.ones(3), 4)


def _layout_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank r owns values 100 r + k; it sends its first 2 / last 3 owned entries, receives 3 from below and 2 from above
        n_own = 10
        lower, upper = (rank - 1) % world, (rank + 1) % world
        lay = slab._LocalLayout(rank, world, lower, upper, 3, n_own, 2, 2, 3, None)
        local = torch.zeros(lay.n_local, dtype=torch.float64)
        local[lay.own_begin:lay.own_end] = torch.arange(n_own, dtype=torch.float64) + 100.0 * rank
        lay.fill_halo(local)
        want_lo = torch.arange(n_own - 3, n_own, dtype=torch.float64) + 100.0 * lower   # the lower neighbour's last 3
        want_hi = torch.arange(0, 2, dtype=torch.float64) + 100.0 * upper              # the upper neighbour's first 2
        ok = torch.equal(local[:3], want_lo) and torch.equal(local[lay.own_end:], want_hi)
        out = lay.assemble(torch.arange(n_own, dtype=torch.float64) + 100.0 * rank)
        ok = ok and torch.equal(out, local) and lay.halo_bytes(8) == 40
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_local_layout_halo_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_layout_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
