"""helpers shared by the GPU parity tests"""
import numpy as np
import torch

import aboria_b200 as ab
from oracle import oracle as orc


def build_both(pos, low, high, periodic, n_leaf=10.0, alive=None, variables=None, two_level=None):
    """Runs init_neighbour_search on the oracle (stable sort) and on the GPU
    for the same input; returns (oracle, oracle_out, particles)."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n, D = pos.shape
    o = orc.Oracle(D)
    al = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8).copy()
    out = o.init_neighbour_search(pos, low, high, periodic, n_leaf, alive=al, sort_mode=orc.SORT_STABLE)
    p = ab.Particles(D, n, variables=variables)
    if two_level is not None:
        # build strategy: False = LSD radix sort + gather, True = two-level radix build, "counting" = counting-sort build
        # "records": two-level with the binned copy as one record per particle; "direct": two-level without the
        # bulk-copy staged record move and with the column-by-column final gather (the round-1 kernels)
        p.set_option("two_level_min_n", 0 if two_level in (True, "records", "direct") else 1e18)
        p.set_option("counting_min_n", 0 if two_level == "counting" else 1e18)
        p.set_option("record_aos", 1 if two_level == "records" else 0)
        p.set_option("stage_records", 0 if two_level == "direct" else 1)
        p.set_option("gather_slots", 0 if two_level == "direct" else 1)
        p.set_option("skip_alive_move", 0 if two_level == "direct" else 1)
        p.set_option("bounds_one_sweep", 0 if two_level == "direct" else 1)
    p.set("position", torch.from_numpy(pos.copy()))
    if alive is not None:
        p.set("alive", torch.from_numpy(np.ascontiguousarray(alive, dtype=np.uint8)))
    p.init_neighbour_search(low, high, periodic, n_leaf)
    return o, out, p


def assert_build_equal(o, out, p):
    """bit-exact: order, sorted keys, bucket_begin/end, wrapped positions"""
    na = out["n_alive"]
    assert p.size() == na
    osize, oside = o.grid()
    gsize, gside, nb = p.grid()
    assert list(osize) == list(gsize)
    assert np.array_equal(oside, gside)
    order = p.get_alive_indicies().cpu().numpy()
    assert np.array_equal(order, out["order"])
    q = p.get_query()
    assert np.array_equal(q.bucket_indices.cpu().numpy().view(np.uint32)[:na], out["keys"])
    assert np.array_equal(q.bucket_begin.cpu().numpy().view(np.uint32), out["bucket_begin"])
    assert np.array_equal(q.bucket_end.cpu().numpy().view(np.uint32), out["bucket_end"])
    gp = p.get("position").cpu().numpy()
    assert np.array_equal(gp.view(np.uint64), out["pos"].view(np.uint64))
    assert np.array_equal(p.get("id").cpu().numpy(), out["order"].astype(np.int64))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)
