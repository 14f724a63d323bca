#!/usr/bin/env python
"""Times the rows next to the hot path on one GPU, with the oracle on
the host beside them: find-by-id (id-map build + find), coeff, assemble (CSR),
bucket-pair traversal + fast cell-list search, scaled distance_search, and
accumulate_within_distance.  Prints a JSON list; evidence for profiles/, not the
driver's bench."""
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


def gpu_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


def cpu_ms(fn):
    t0 = time.perf_counter()
    r = fn()
    return (time.perf_counter() - t0) * 1e3, r


def main():
    N = int(os.environ.get("ABR_NEXT_N", 4_000_000))
    NC = int(os.environ.get("ABR_NEXT_CPU_N", 400_000))
    out = []
    pos = synth.uniform_positions(N, 3)
    p = ab.Particles(3, 0)
    p.resize_from_positions(torch.from_numpy(pos).to(dev))
    rng = np.random.default_rng(1)
    ids = rng.permutation(N).astype(np.int64)
    p.set("id", torch.from_numpy(ids))
    p.init_neighbour_search(0.0, 1.0, True)
    size, side, nb = p.grid()
    r = float(side[0])
    # CPU sample of the same density
    cpos = synth.uniform_positions(NC, 3)
    o = orc.Oracle(3)
    oo = o.init_neighbour_search(cpos, 0.0, 1.0, True)
    _, cside = o.grid()
    cr = float(cside[0])

    # find-by-id
    ms = gpu_ms(lambda: p.init_id_search())
    q = torch.from_numpy(rng.integers(0, 2 * N, size=N).astype(np.int64)).to(dev)
    ms_find = gpu_ms(lambda: p.get_query().find(q))
    cids = rng.permutation(NC).astype(np.uint64)
    t_cpu, (ck, cv) = cpu_ms(lambda: orc.id_map_build(cids))
    cq = rng.integers(0, 2 * NC, size=NC).astype(np.uint64)
    t_cpu_f, _ = cpu_ms(lambda: orc.id_find(ck, cv, cq))
    out.append({"row": "find-by-id: id map build", "n": N, "gpu_ms": ms, "gpu_mitems_per_s": N / ms / 1e3, "cpu_n": NC, "cpu_ms": t_cpu, "cpu_mitems_per_s": NC / t_cpu / 1e3})
    out.append({"row": "find-by-id: find", "n": N, "gpu_ms": ms_find, "gpu_mitems_per_s": N / ms_find / 1e3, "cpu_n": NC, "cpu_ms": t_cpu_f, "cpu_mitems_per_s": NC / t_cpu_f / 1e3})

    # coeff
    op = ab.create_sparse_operator(p, p, r, K.inv_dist(0.1))
    m = 8_000_000
    ii = torch.from_numpy(rng.integers(0, N, size=m)).to(dev)
    jj = torch.from_numpy(rng.integers(0, N, size=m)).to(dev)
    ms = gpu_ms(lambda: op.coeff(ii, jj))
    mc = 800_000
    ci, cj = rng.integers(0, NC, size=mc), rng.integers(0, NC, size=mc)
    t_cpu, _ = cpu_ms(lambda: o.coeff(oo["pos"], oo["pos"], ci, cj, orc.K_INV_DIST, [0.1], cr))
    out.append({"row": "coeff(i, j)", "n": m, "gpu_ms": ms, "gpu_mitems_per_s": m / ms / 1e3, "cpu_n": mc, "cpu_ms": t_cpu, "cpu_mitems_per_s": mc / t_cpu / 1e3})

    # assemble to CSR
    ms = gpu_ms(lambda: op.assemble(), reps=3)
    rp, col, val = op.assemble()
    nnz = int(col.shape[0])
    t_cpu, (crp, ccol, cval) = cpu_ms(lambda: o.assemble(oo["pos"], orc.K_INV_DIST, [0.1], cr))
    out.append({"row": "assemble (CSR, values)", "n": nnz, "gpu_ms": ms, "gpu_mitems_per_s": nnz / ms / 1e3, "cpu_n": int(len(ccol)), "cpu_ms": t_cpu, "cpu_mitems_per_s": len(ccol) / t_cpu / 1e3})
    del rp, col, val

    # bucket pairs + fast cell-list search (bucket side >= r is given: r = side)
    ms = gpu_ms(lambda: p.get_query().neighbouring_buckets(), reps=3)
    npairs = int(p.get_query().neighbouring_buckets()[0].shape[0])
    t_cpu, cb = cpu_ms(lambda: o.bucket_pairs())
    out.append({"row": "get_neighbouring_buckets (pair list)", "n": npairs, "gpu_ms": ms, "gpu_mitems_per_s": npairs / ms / 1e3, "cpu_n": int(len(cb[0])), "cpu_ms": t_cpu, "cpu_mitems_per_s": len(cb[0]) / t_cpu / 1e3})
    ms = gpu_ms(lambda: p.get_query().fast_bucket_search_counts(r), reps=3)
    t_cpu, _ = cpu_ms(lambda: o.fast_bucket_search_counts(cr))
    out.append({"row": "fast cell-list search (neighbour counts)", "n": N, "gpu_ms": ms, "gpu_mitems_per_s": N / ms / 1e3, "cpu_n": NC, "cpu_ms": t_cpu, "cpu_mitems_per_s": NC / t_cpu / 1e3})

    # scaled distance_search (per-query walk)
    ms = gpu_ms(lambda: p.distance_search_stats(r, 2, scale=[1.0, 1.5, 0.75]), reps=3)
    t_cpu, _ = cpu_ms(lambda: o.pair_stats_norm(oo["pos"], cr, 2, scale=[1.0, 1.5, 0.75]))
    out.append({"row": "distance_search<2> with ScaleTransform (count + hash per query)", "n": N, "gpu_ms": ms, "gpu_mitems_per_s": N / ms / 1e3, "cpu_n": NC, "cpu_ms": t_cpu, "cpu_mitems_per_s": NC / t_cpu / 1e3, "cpu_threads": orc.max_threads()})

    # accumulate_within_distance (SPH density sum)
    h = 0.5 * r
    ms = gpu_ms(lambda: ab.accumulate_within_distance(p, p, r, K.sph_density(h, 1.0 / N, 0.0261)))
    t_cpu, _ = cpu_ms(lambda: o.accumulate_within_distance(oo["pos"], orc.K_SPH_DENSITY, [0.5 * cr, 1.0 / NC, 0.0261], cr))
    out.append({"row": "accumulate_within_distance (SPH density)", "n": N, "gpu_ms": ms, "gpu_mitems_per_s": N / ms / 1e3, "cpu_n": NC, "cpu_ms": t_cpu, "cpu_mitems_per_s": NC / t_cpu / 1e3, "cpu_threads": orc.max_threads()})
    for e in out:
        e["speedup_per_item"] = e["gpu_mitems_per_s"] / e["cpu_mitems_per_s"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
