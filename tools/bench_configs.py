#!/usr/bin/env python
"""Times the BASELINE.json configs c1-c4 (build + one product) on one GPU and
the oracle on a bounded CPU sample of the same workload; prints a JSON list.
Not the driver's bench (that is bench.py, config c5); evidence for profiles/."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aboria_b200 as ab  # noqa: E402
from aboria_b200 import kernels as K  # noqa: E402
from aboria_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

dev = torch.device("cuda:0")


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def run(name, D, N, make_pos, low, high, periodic, n_leaf, radius_of, kern_of, okern, cpu_n):
    pos = make_pos(N)
    p = ab.Particles(D, 0)
    p.resize_from_positions(torch.as_tensor(pos).to(dev))
    p.init_neighbour_search(low, high, periodic, n_leaf)
    size, side, nb = p.grid()
    radius, extra = radius_of(N, side)
    kern = kern_of(extra)
    op = ab.create_sparse_operator(p, p, radius, kern)
    b = torch.from_numpy(synth.vector(p.size() * kern.block_cols)).to(dev)
    cnt, _ = p.pair_stats(radius)
    pairs = int(cnt.long().sum().item())
    pos_dev = torch.as_tensor(pos).to(dev)

    def build():
        p.resize_from_positions(pos_dev.clone())
        p.init_neighbour_search(low, high, periodic, n_leaf)

    ms_build = gpu_time(build)
    ms_mv = gpu_time(lambda: op.matvec(b))
    # CPU sample
    cpos = make_pos(cpu_n)
    o = orc.Oracle(D)
    t0 = time.perf_counter()
    out = o.init_neighbour_search(cpos, low, high, periodic, n_leaf, sort_mode=orc.SORT_STD)
    t1 = time.perf_counter()
    cr, cextra = radius_of(cpu_n, o.grid()[1])
    kid, kparams, BR = okern(cextra)
    _, cpairs = o.sparse_matvec(out["pos"], kid, kparams, cr, synth.vector(cpu_n), BR=BR, BC=1)
    t2 = time.perf_counter()
    return {"config": name, "n": N, "buckets": int(nb), "radius": radius, "pairs": pairs, "pairs_per_row": pairs / N,
            "gpu_ms_build": ms_build, "gpu_ms_matvec": ms_mv, "gpu_pairs_per_s": pairs / (ms_mv * 1e-3),
            "gpu_build_mparticles_per_s": N / (ms_build * 1e-3) / 1e6, "walk_rows": p.last_counters()["walk_rows"],
            "cpu": {"n": cpu_n, "cores": orc.max_threads(), "s_build": t1 - t0, "s_matvec": t2 - t1, "pairs_per_s": cpairs / (t2 - t1),
                    "build_mparticles_per_s": cpu_n / (t1 - t0) / 1e6}}


def main():
    out = []
    out.append(run("c1: 3-D periodic unit cube N=1e5, 1/(r+eps), r=0.05", 3, 100_000, lambda n: synth.uniform_positions(n, 3), 0.0, 1.0, True, 10.0,
                   lambda n, side: (0.05, None), lambda e: K.inv_dist(0.1), lambda e: (orc.K_INV_DIST, [0.1], 1), 100_000))
    out.append(run("c2: 2-D RBF Wendland C2 N=1e6, ~30 nbrs", 2, 1_000_000, lambda n: synth.uniform_positions(n, 2), 0.0, 1.0, False, 10.0,
                   lambda n, side: (np.sqrt(30.0 / (np.pi * n)), 0.5 * np.sqrt(30.0 / (np.pi * n))), lambda h: K.wendland_c2(h),
                   lambda h: (orc.K_WENDLAND_C2, [h], 1), 1_000_000))
    L = (4_000_000 / 0.8442) ** (1.0 / 3.0)
    Lc = (400_000 / 0.8442) ** (1.0 / 3.0)
    out.append(run("c3: 3-D LJ force N=4M periodic, cutoff 2.5 sigma", 3, 4_000_000,
                   lambda n: synth.uniform_positions(n, 3, 0.0, (n / 0.8442) ** (1.0 / 3.0)), 0.0, L, True, 10.0,
                   lambda n, side: (2.5, None), lambda e: K.lj_force(3, 1.0, 1.0), lambda e: (orc.K_LJ_FORCE, [1.0, 1.0], 3), 4_000_000)
               if False else run_c3(L, Lc))
    out.append(run("c4: SPH density N=16M clustered, periodic (1,1,0), r=2h", 3, 16_000_000, lambda n: synth.clustered_positions(n), 0.0, 1.0,
                   [True, True, False], 10.0, lambda n, side: (3.0 * n ** (-1.0 / 3.0), 1.5 * n ** (-1.0 / 3.0)),
                   lambda h: K.sph_density(h, 1.0, 21.0 / (256.0 * np.pi)), lambda h: (orc.K_SPH_DENSITY, [h, 1.0, 21.0 / (256.0 * np.pi)], 1), 1_600_000))
    print(json.dumps(out, indent=1))


def run_c3(L, Lc):
    # the CPU sample uses its own box (same density), N/10
    name = "c3: 3-D LJ force N=4M periodic, cutoff 2.5 sigma"
    N = 4_000_000
    pos = synth.uniform_positions(N, 3, 0.0, L)
    p = ab.Particles(3, 0)
    pos_dev = torch.as_tensor(pos).to(dev)
    p.resize_from_positions(pos_dev.clone())
    p.init_neighbour_search(0.0, L, True, 10.0)
    op = ab.create_sparse_operator(p, p, 2.5, K.lj_force(3, 1.0, 1.0))
    b = torch.ones(N, dtype=torch.float64, device=dev)
    cnt, _ = p.pair_stats(2.5)
    pairs = int(cnt.long().sum().item())

    def build():
        p.resize_from_positions(pos_dev.clone())
        p.init_neighbour_search(0.0, L, True, 10.0)

    ms_build = gpu_time(build)
    ms_mv = gpu_time(lambda: op.matvec(b))
    cn = 400_000
    cpos = synth.uniform_positions(cn, 3, 0.0, Lc)
    o = orc.Oracle(3)
    t0 = time.perf_counter()
    out = o.init_neighbour_search(cpos, 0.0, Lc, True, 10.0, sort_mode=orc.SORT_STD)
    t1 = time.perf_counter()
    _, cpairs = o.sparse_matvec(out["pos"], orc.K_LJ_FORCE, [1.0, 1.0], 2.5, np.ones(cn), BR=3, BC=1)
    t2 = time.perf_counter()
    nb = p.grid()[2]
    return {"config": name, "n": N, "buckets": int(nb), "radius": 2.5, "pairs": pairs, "pairs_per_row": pairs / N, "gpu_ms_build": ms_build,
            "gpu_ms_matvec": ms_mv, "gpu_pairs_per_s": pairs / (ms_mv * 1e-3), "gpu_build_mparticles_per_s": N / (ms_build * 1e-3) / 1e6,
            "walk_rows": p.last_counters()["walk_rows"],
            "cpu": {"n": cn, "cores": orc.max_threads(), "s_build": t1 - t0, "s_matvec": t2 - t1, "pairs_per_s": cpairs / (t2 - t1),
                    "build_mparticles_per_s": cn / (t1 - t0) / 1e6}}


if __name__ == "__main__":
    main()
