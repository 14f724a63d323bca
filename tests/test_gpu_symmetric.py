"""The symmetric form of the product (abr_set_option("symmetric", 1)): every unordered pair of
the half stencil evaluated once and added to both rows — the reference's "fast cell-list
search" traversal (src/Search.h:498-764, tests/neighbours.h:281-300) applied to
KernelSparse::evaluate for functors that declare SYMMETRY.  Same pair sets, same vectors
within the 1e-12 budget (summation order only), checked against the oracle on the cases that
stress it: periodic wrap, tiny grids (every neighbour is an image), r > bucket side, odd
(force) kernels, block kernels, heavy buckets, lattice points on bucket faces (every row goes
to the exact walk), dead particles."""
import numpy as np
import pytest
import torch

import aboria_b200 as ab
from aboria_b200 import kernels as K
from aboria_b200 import synth
from oracle import oracle as orc
from util import build_both, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _check(o, out, p, kern, okid, params, r, BR=1, seed=3):
    n = out["n_alive"]
    b = synth.vector(n, seed=synth.SEED + seed)
    y_o, npairs = o.sparse_matvec(out["pos"], okid, params, r, b, BR=BR, BC=1)
    bt = torch.from_numpy(b).to(p.device)
    op = ab.create_sparse_operator(p, p, r, kern)
    p.set_option("symmetric", 0)
    y0 = (op * bt).cpu().numpy()
    p.set_option("symmetric", 1)
    y1 = (op * bt).cpu().numpy()
    launches = p.last_counters()["launches"]
    p.set_option("symmetric", 0)
    assert rel_l2(y0, y_o) <= TOL
    assert rel_l2(y1, y_o) <= TOL, rel_l2(y1, y_o)
    # evaluate accumulates (y += K b) in the symmetric form too
    p.set_option("symmetric", 1)
    y2 = torch.from_numpy(y_o.copy()).to(p.device)
    op.evaluate(y2, bt)
    p.set_option("symmetric", 0)
    assert rel_l2(y2.cpu().numpy(), 2 * y_o) <= TOL
    return launches, npairs


@pytest.mark.parametrize("D,N,periodic,rfac,nn", [
    (3, 20000, True, 1.0, 10), (3, 20000, False, 1.0, 10), (3, 20000, True, 1.7, 10), (3, 20000, [True, False, True], 0.6, 10),
    (2, 20000, True, 1.0, 10), (2, 20000, False, 2.3, 10), (1, 2000, True, 1.0, 10), (1, 2000, False, 1.5, 10),
    (3, 50, True, 1.0, 10), (3, 200, True, 0.9, 100), (3, 3000, True, 1.0, 1), (3, 300000, True, 1.0, 10)])
def test_symmetric_scalar_kernels(D, N, periodic, rfac, nn):
    rng = np.random.default_rng(7 * D + N)
    pos = rng.random((N, D))
    o, out, p = build_both(pos, 0.0, 1.0, periodic, nn)
    side = float(o.grid()[1][0])
    r = rfac * side
    launches, _ = _check(o, out, p, K.inv_dist(0.1), orc.K_INV_DIST, [0.1], r)
    assert launches == 3  # symmetric tiled kernel + combine + exact-walk pass
    h = r / 2
    _check(o, out, p, K.wendland_c2(h), orc.K_WENDLAND_C2, [h], r)
    _check(o, out, p, K.sph_density(h, 0.5, 1.3), orc.K_SPH_DENSITY, [h, 0.5, 1.3], r)


@pytest.mark.parametrize("D,N,periodic,rfac", [(3, 20000, True, 1.0), (3, 20000, False, 1.4), (2, 20000, True, 1.0)])
def test_symmetric_force_kernels(D, N, periodic, rfac):
    # odd kernels (SYMMETRY = -1), D x 1 blocks: block(-dx) = -block(dx)
    rng = np.random.default_rng(5 * D + N)
    pos = rng.random((N, D))
    o, out, p = build_both(pos, 0.0, 1.0, periodic)
    r = rfac * float(o.grid()[1][0])
    _check(o, out, p, K.lj_force(D, 0.4 * r, 1.0), orc.K_LJ_FORCE, [0.4 * r, 1.0], r, BR=D)
    _check(o, out, p, K.linear_spring(D, 2.0, 0.8 * r), orc.K_LINEAR_SPRING, [2.0, 0.8 * r], r, BR=D)


def test_symmetric_lattice_and_dead_particles():
    # lattice points sit exactly on bucket faces: every row is rounding-sensitive and is handed to the
    # exact walk (flagged by its own bucket AND by its partners: the list must not hold duplicates)
    n = 20
    g = (np.arange(n) + 0.0) / n
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    o, out, p = build_both(pos, 0.0, 1.0, True, 1.0)
    _check(o, out, p, K.inv_dist(0.1), orc.K_INV_DIST, [0.1], 1.0001 / n)
    # particles outside a non-periodic domain die; NaNs die
    rng = np.random.default_rng(9)
    pos = rng.uniform(-0.2, 1.2, (30000, 3))
    pos[::101, 0] = np.nan
    o, out, p = build_both(pos, 0.0, 1.0, [False, True, False])
    assert out["n_alive"] < 30000
    _check(o, out, p, K.inv_dist(0.1), orc.K_INV_DIST, [0.1], 0.9 * float(o.grid()[1][0]))


def test_symmetric_clustered_heavy_buckets():
    pos = synth.clustered_positions(400000)
    o, out, p = build_both(pos, 0.0, 1.0, [True, True, False])
    h = 1.5 * 400000 ** (-1.0 / 3.0)
    _check(o, out, p, K.sph_density(h, 1.0 / 400000, 21.0 / (256.0 * np.pi)), orc.K_SPH_DENSITY, [h, 1.0 / 400000, 21.0 / (256.0 * np.pi)], 2 * h)


def test_symmetric_large_properties():
    # 4M particles: symmetric and ordered forms agree; x.(K y) == y.(K x); antisymmetric forces sum to zero
    N = 4_000_000
    L = (N / 0.8442) ** (1.0 / 3.0)
    dev = torch.device("cuda:0")
    pos = synth.torch_uniform_positions(N, 3, 0.0, L, synth.SEED, 0, dev)
    p = ab.Particles(3, 0)
    p.resize_from_positions(pos)
    p.init_neighbour_search(0.0, L, True)
    r = 2.5
    op = ab.create_sparse_operator(p, p, r, K.inv_dist(0.1))
    x = torch.from_numpy(synth.vector(N, seed=5)).to(dev)
    yv = torch.from_numpy(synth.vector(N, seed=6)).to(dev)
    Kx0 = op * x
    p.set_option("symmetric", 1)
    Kx, Ky = op * x, op * yv
    assert float(torch.linalg.norm(Kx - Kx0) / torch.linalg.norm(Kx0)) <= TOL
    lhs, rhs = float(torch.dot(x, Ky)), float(torch.dot(yv, Kx))
    assert abs(lhs - rhs) / abs(lhs) < 1e-12
    opf = ab.create_sparse_operator(p, p, r, K.lj_force(3, 1.0, 1.0))
    f = (opf * torch.ones(N, dtype=torch.float64, device=dev)).view(N, 3)
    assert f.sum(dim=0).abs().max().item() / f.abs().sum().item() < 1e-12
