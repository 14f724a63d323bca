"""Seeded differential test: random small configurations of the whole path — dimension, box
(non-cubic, off-origin), per-dimension periodicity, occupancy, cloud shape (uniform, blobs,
lattice on bucket faces, duplicated points), dead / non-finite / out-of-domain particles, build
strategy, cut-off from a fraction of a bucket to several buckets, functor, ordered or symmetric
or bulk-staged product, rows == columns or a separate row set — CUDA path against the oracle.
Bit-exact for the build and the pair sets, rel. L2 <= 1e-12 for the product vectors.

The reference tests the same space case by case (tests/neighbours.h:1252-1327 case lists,
tests/operators.h:810-933, tests/rbf_interpolation.h:326); this walks through it at random with
fixed seeds so that a failure reproduces from its case number."""
import os

import numpy as np
import pytest
import torch

import aboria_b200 as ab
from aboria_b200 import kernels as K
from aboria_b200 import synth
from oracle import oracle as orc
from util import assert_build_equal, build_both, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12
N_CASES = int(os.environ.get("ABR_FUZZ_CASES", "240"))      # a longer hunt: ABR_FUZZ_CASES=1000 ABR_FUZZ_SEED=50000 pytest tests/test_gpu_fuzz.py
SEED0 = int(os.environ.get("ABR_FUZZ_SEED", "9000"))


def make_case(case):
    rng = np.random.default_rng(SEED0 + case)
    D = int(rng.choice([1, 2, 3], p=[0.15, 0.35, 0.5]))
    nmax = {1: 3000, 2: 20000, 3: 30000}[D]
    N = int(np.exp(rng.uniform(0.0, np.log(nmax)))) if rng.random() < 0.5 else int(rng.integers(nmax // 20, nmax))
    low = rng.uniform(-2.0, 1.0, size=D)
    ext = rng.uniform(0.5, 3.0, size=D) if rng.random() < 0.6 else np.full(D, rng.uniform(0.5, 3.0))
    high = low + ext
    periodic = [bool(x) for x in rng.random(D) < 0.5]
    n_leaf = float(rng.choice([1.0, 3.0, 10.0, 37.5]))
    shape = rng.choice(["uniform", "blobs", "lattice", "duplicates"], p=[0.45, 0.3, 0.15, 0.1])
    if shape == "uniform":
        pos = low + rng.random((N, D)) * ext
    elif shape == "blobs":
        nb = int(rng.integers(1, 6))
        centres = low + rng.random((nb, D)) * ext
        which = rng.integers(0, nb, size=N)
        pos = centres[which] + rng.normal(0.0, 0.04 * ext.min(), size=(N, D))
        bg = rng.random(N) < 0.1
        pos[bg] = low + rng.random((int(bg.sum()), D)) * ext
    elif shape == "lattice":
        m = max(1, int(round(N ** (1.0 / D))))
        g = [low[d] + np.arange(m) * (ext[d] / m) for d in range(D)]
        pos = np.stack(np.meshgrid(*g, indexing="ij"), axis=-1).reshape(-1, D)
    else:
        base = low + rng.random((max(1, N // int(rng.choice([7, 90]))), D)) * ext  # 90 copies of a point: buckets beyond 64 rows
        pos = base[rng.integers(0, base.shape[0], size=N)]
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    N = pos.shape[0]
    alive = None
    if rng.random() < 0.3:
        alive = (rng.random(N) < 0.9).astype(np.uint8)
    if rng.random() < 0.3 and N > 4:
        k = max(1, N // 20)
        idx = rng.choice(N, size=k, replace=False)
        pos[idx] = low + (rng.random((k, D)) * 1.6 - 0.3) * ext  # some outside: wrapped or killed per dimension
        pos[idx[0], rng.integers(0, D)] = np.nan
        if k > 1:
            pos[idx[1], rng.integers(0, D)] = np.inf
    strategy = [False, True, "counting"][int(rng.integers(0, 3))]
    if strategy is True:
        strategy = [True, "records", "direct"][case % 3]
    return dict(rng=rng, D=D, N=N, low=low, high=high, periodic=periodic, n_leaf=n_leaf, shape=str(shape), pos=pos, alive=alive,
                strategy=strategy)


def pick_kernel(rng, D, r):
    h = r / 2
    k = int(rng.integers(0, 6))
    if k == 0:
        return K.inv_dist(0.1), orc.K_INV_DIST, [0.1], 1
    if k == 1:
        return K.wendland_c2(h), orc.K_WENDLAND_C2, [h], 1
    if k == 2:
        return K.sph_density(h, 0.5, 1.3), orc.K_SPH_DENSITY, [h, 0.5, 1.3], 1
    if k == 3:
        return K.lj_force(D, 0.4 * r, 1.0), orc.K_LJ_FORCE, [0.4 * r, 1.0], D
    if k == 4:
        return K.linear_spring(D, 2.0, 0.8 * r), orc.K_LINEAR_SPRING, [2.0, 0.8 * r], D
    return K.inv_dist(1e-3 * r), orc.K_INV_DIST, [1e-3 * r], 1


def product_agrees(y, y_o, o, rows, okid, params, r, b, BR, npairs, radius_per_row=None):
    """rel. L2 <= TOL against the oracle.  A sum whose terms cancel (a particle between its own periodic
    images, a perfect lattice under a force kernel: the exact result is 0) leaves only rounding noise in both
    vectors; there the error is measured against the size of the terms, sum_j |K_ij b_j|, taken from the
    oracle's assembled matrix."""
    if rel_l2(y, y_o) <= TOL:
        return True
    if npairs > 5e6:
        return False
    row_ptr, col, vals = o.assemble(rows, okid, params, r, BR=BR, radius_per_row=radius_per_row)
    n_rows = rows.shape[0]
    vals = np.asarray(vals).reshape(len(col), BR)
    terms = np.abs(vals * b[col][:, None])
    mag = np.zeros((n_rows, BR))
    np.add.at(mag, np.repeat(np.arange(n_rows), np.diff(np.append(row_ptr[:n_rows], len(col)).astype(np.int64))), terms)
    return np.linalg.norm(y - y_o) <= TOL * np.linalg.norm(mag)


@pytest.mark.parametrize("case", range(N_CASES))
def test_random_configuration(case):
    c = make_case(case)
    rng, D = c["rng"], c["D"]
    o, out, p = build_both(c["pos"], c["low"], c["high"], c["periodic"], c["n_leaf"], alive=c["alive"], two_level=c["strategy"])
    assert_build_equal(o, out, p)
    n = out["n_alive"]
    if n == 0:
        return
    side = np.asarray(o.grid()[1], dtype=np.float64)
    # cut-off: a fraction of a bucket up to a few buckets, capped so the oracle's pair count stays small
    r = float(side.min() * rng.choice([0.05, 0.4, 0.999, 1.0, 1.3, 1.9, 2.6]))
    dens = n / float(np.prod(c["high"] - c["low"]))
    ball = {1: 2 * r, 2: np.pi * r * r, 3: 4.0 / 3.0 * np.pi * r ** 3}[D]
    if c["shape"] in ("blobs", "duplicates"):
        r = min(r, float(side.min()) * 1.3)
    while n * dens * ball > 2e7 and r > 1e-3:
        r *= 0.7
        ball = {1: 2 * r, 2: np.pi * r * r, 3: 4.0 / 3.0 * np.pi * r ** 3}[D]
    tag = f"case {case}: D={D} N={c['N']} alive={n} {c['shape']} periodic={c['periodic']} n_leaf={c['n_leaf']} r/side={r / side.min():.3f} build={c['strategy']}"

    # neighbour pair sets (count + hash per row), cell-tiled kernel and exact walk
    cnt_o, hs_o = o.pair_stats(out["pos"], r)
    for path in (0, 1):
        try:
            cnt, hs = p.pair_stats(r, path=path)
        except ab.AbrError:
            assert path == 0, tag
            continue
        assert np.array_equal(cnt.cpu().numpy().view(np.uint32), cnt_o), (tag, path)
        assert np.array_equal(hs.cpu().numpy().view(np.uint64), hs_o), (tag, path)

    # the product, in one of its forms
    kern, okid, params, BR = pick_kernel(rng, D, r)
    form = str(rng.choice(["ordered", "symmetric", "staged"]))
    b = synth.vector(n, seed=synth.SEED + case)
    y_o, npairs = o.sparse_matvec(out["pos"], okid, params, r, b, BR=BR, BC=1)
    assert npairs == int(cnt_o.astype(np.int64).sum()), tag
    bt = torch.from_numpy(b).to(p.device)
    op = ab.create_sparse_operator(p, p, r, kern)
    p.set_option("symmetric", 1 if form == "symmetric" else 0)
    p.set_option("matvec_variant", 1 if form == "staged" else 0)
    y = (op * bt).cpu().numpy()
    p.set_option("symmetric", 0)
    p.set_option("matvec_variant", 0)
    assert product_agrees(y, y_o, o, out["pos"], okid, params, r, b, BR, npairs), (tag, form, okid, rel_l2(y, y_o))

    # a separate row set (no search structure of its own, partly outside the domain)
    if rng.random() < 0.5:
        M = int(np.exp(rng.uniform(0.0, np.log(6000))))
        ext = c["high"] - c["low"]
        rows = c["low"] + (rng.random((M, D)) * 1.2 - 0.1) * ext
        p.set_option("xrows_min_n", float(rng.choice([0, 1024])))
        y_or, _ = o.sparse_matvec(rows, okid, params, r, b, BR=BR, BC=1)
        test = ab.Particles(D, M)
        test.set("position", torch.from_numpy(rows.copy()))
        G = ab.create_sparse_operator(test, p, r, kern)
        yr = (G * bt).cpu().numpy()
        assert product_agrees(yr, y_or, o, rows, okid, params, r, b, BR, npairs), (tag, "rows != cols", M, okid, rel_l2(yr, y_or))
        # per-row radii (the FRadius overload, src/Operators.h:478-489), zero and beyond-the-box radii included
        if rng.random() < 0.5:
            rpr = rng.uniform(0.0, 1.5 * r, size=M)
            rpr[rng.integers(0, M)] = 0.0
            y_o2, _ = o.sparse_matvec(rows, okid, params, 0.0, b, BR=BR, BC=1, radius_per_row=rpr)
            G2 = ab.create_sparse_operator(test, p, rpr, kern)
            y2 = (G2 * bt).cpu().numpy()
            assert product_agrees(y2, y_o2, o, rows, okid, params, 0.0, b, BR, npairs, radius_per_row=rpr), (tag, "per-row radius", M, okid, rel_l2(y2, y_o2))
