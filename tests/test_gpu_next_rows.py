"""GPU parity tests for the SURVEY §8f rows next to the hot path: find-by-id,
coeff(i, j), block operators.  Bit-exact for the id map and the found indices;
coefficients equal to the oracle's within 2 ulp-level tolerance (same kernel
functions as the product, TOL 1e-12 relative)."""
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


@pytest.mark.parametrize("N,idmax", [(1, 10), (100, 100), (5000, 1 << 20), (70000, 1 << 31), (70000, (1 << 40) + 12345), (3000, (1 << 63) - 1)])
def test_id_map_parity(N, idmax):
    # ids: unique, arbitrary 64-bit values; compare m_id_map_key / m_id_map_value with the oracle
    rng = np.random.default_rng(N + (idmax & 0xFFFF))
    if idmax <= 4 * N:
        ids = rng.permutation(max(N, idmax))[:N].astype(np.uint64)
    else:
        ids = np.unique(rng.integers(0, idmax, size=2 * N, dtype=np.uint64))
        ids = rng.permutation(ids)[:N]
        ids[0] = np.uint64(idmax - 1) if N > 1 else ids[0]
        ids = np.array(sorted(set(ids.tolist())), dtype=np.uint64)
        ids = rng.permutation(ids)
    n = len(ids)
    p = ab.Particles(3, n)
    p.set("position", torch.from_numpy(synth.uniform_positions(n, 3)))
    p.set("id", torch.from_numpy(ids.view(np.int64).copy()))
    p.init_id_search()
    key, value = p.id_map()
    okey, ovalue = orc.id_map_build(ids)
    assert np.array_equal(key.cpu().numpy().view(np.uint64), okey)
    assert np.array_equal(value.cpu().numpy().view(np.uint64), ovalue)
    q = np.concatenate([ids[: min(n, 1000)], rng.integers(0, idmax, size=1000, dtype=np.uint64), np.array([0, idmax], dtype=np.uint64)])
    got = p.get_query().find(torch.from_numpy(q.view(np.int64).copy())).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, orc.id_find(okey, ovalue, q))


def test_id_search_documentation_example_and_reorder():
    # tests/id_search.h:64-99, then the id map must follow init_neighbour_search's reorder
    # (src/NeighbourSearchBase.h:440-486 runs inside every update_positions)
    N = 100
    rng = np.random.default_rng(3)
    ids = rng.permutation(N).astype(np.int64)
    pos = synth.uniform_positions(N, 3)
    p = ab.Particles(3, N)
    p.set("position", torch.from_numpy(pos.copy()))
    p.set("id", torch.from_numpy(ids.copy()))
    p.init_id_search()
    f = p.get_query().find([2, 2 * N]).cpu().numpy()
    assert p.get("id")[int(f[0])].item() == 2
    assert f[1] == N  # end of the particle vector
    p.init_neighbour_search(0.0, 1.0, True)
    f = p.get_query().find(np.arange(N)).cpu().numpy()
    assert np.array_equal(p.get("id").cpu().numpy()[f], np.arange(N))
    # positions found through the id are the original ones
    assert np.array_equal(p.get("position").cpu().numpy()[f], pos[np.argsort(ids)])
    with pytest.raises(ab.AbrError):
        ab.Particles(3, 4).get_query().find([1])


KERNELS = [
    ("const_sum", lambda D: K.const_sum("a", "a"), orc.K_CONST_SUM, lambda D: [], 1),
    ("const_sum_diff", lambda D: K.const_sum_diff("a", "a"), orc.K_CONST_SUM_DIFF, lambda D: [], 2),
    ("inv_dist", lambda D: K.inv_dist(0.1), orc.K_INV_DIST, lambda D: [0.1], 1),
    ("wendland", lambda D: K.wendland_c2(0.07), orc.K_WENDLAND_C2, lambda D: [0.07], 1),
    ("lj", lambda D: K.lj_force(D, 0.05, 1.0), orc.K_LJ_FORCE, lambda D: [0.05, 1.0], None),
]


@pytest.mark.parametrize("D,periodic", [(1, True), (2, False), (3, True), (3, False)])
@pytest.mark.parametrize("kname,kmake,kid,kparams,br", KERNELS)
def test_coeff_parity(D, periodic, kname, kmake, kid, kparams, br):
    N, r = 600, 0.14
    BR = D if br is None else br
    rng = np.random.default_rng(D * 7 + int(periodic))
    pos = rng.random((N, D))
    a = rng.random(N)
    o, out, p = build_both(pos, 0.0, 1.0, periodic, variables={"a": torch.float64})
    order = out["order"]
    p2 = ab.Particles(D, N, variables={"a": torch.float64})
    p2.set("position", torch.from_numpy(pos.copy()))
    p2.set("a", torch.from_numpy(a.copy()))
    p2.init_neighbour_search(0.0, 1.0, periodic)
    a_sorted = a[order]
    op = ab.create_sparse_operator(p2, p2, r, kmake(D))
    m = 20000
    ii = rng.integers(0, N * BR, size=m)
    jj = rng.integers(0, N, size=m)
    # make sure close pairs are among them: neighbours of the first rows
    cnt, j, im, dx = o.search_point(out["pos"][0], r)
    ii[: len(j)] = 0
    jj[: len(j)] = j
    got = op.coeff(ii, jj).cpu().numpy()
    ref = o.coeff(out["pos"], out["pos"], ii, jj, kid, kparams(D), r, BR=BR, BC=1, row_vars=[a_sorted], col_vars=[a_sorted])
    assert np.array_equal(got == 0.0, ref == 0.0)  # the strict predicate selects the same entries
    assert np.count_nonzero(ref) > 50
    assert rel_l2(got, ref) <= TOL
    with pytest.raises(ValueError):
        op.coeff([N * BR], [0])


def test_coeff_golden_and_strict_predicate():
    # tests/operators.h:873-881 dense matrix [[3,3,0],[3,3,3],[0,3,3]] via coeff
    diameter = 0.1
    pos = np.array([[0, 0, 0], [diameter * 0.9, 0, 0], [diameter * 1.8, 0, 0]], dtype=np.float64)
    p = ab.Particles(3, 3, variables={"s1": torch.float64, "s2": torch.float64})
    p.set("position", torch.from_numpy(pos.copy()))
    p.set("s1", torch.full((3,), 1.0, dtype=torch.float64))
    p.set("s2", torch.full((3,), 2.0, dtype=torch.float64))
    p.init_neighbour_search(-1.0, 1.0, False)
    ids = p.get("id").cpu().numpy()
    C = ab.create_sparse_operator(p, p, diameter, K.const_sum("s1", "s2"))
    ii, jj = np.divmod(np.arange(9), 3)
    c = C.coeff(ii, jj).cpu().numpy().reshape(3, 3)
    expect = np.array([[0.0 if {int(ids[i]), int(ids[j])} == {0, 2} else 3.0 for j in range(3)] for i in range(3)])
    assert np.array_equal(c, expect)
    # a pair at exactly r: in the product (<=), not in coeff (<); minimum image through the boundary
    pos = np.array([[0.125, 0.5, 0.5], [0.375, 0.5, 0.5], [0.9375, 0.5, 0.5]])
    p = ab.Particles(3, 3, variables={"s1": torch.float64, "s2": torch.float64})
    p.set("position", torch.from_numpy(pos.copy()))
    p.set("s1", torch.full((3,), 1.0, dtype=torch.float64))
    p.set("s2", torch.full((3,), 2.0, dtype=torch.float64))
    p.init_neighbour_search(0.0, 1.0, True)
    x = p.get("position").cpu().numpy()[:, 0]
    a, b, cc = [int(np.argmin(np.abs(x - v))) for v in (0.125, 0.375, 0.9375)]
    C = ab.create_sparse_operator(p, p, 0.25, K.const_sum("s1", "s2"))
    c = C.coeff(ii, jj).cpu().numpy().reshape(3, 3)
    assert c[a, b] == 0.0 and c[b, a] == 0.0 and c[a, cc] == 3.0 and c[cc, a] == 3.0
    cnt, _ = p.pair_stats(0.25)
    assert cnt.cpu().numpy()[[a, b, cc]].tolist() == [3, 2, 2]


def test_block_operator():
    # create_block_operator<2,2>(A, B, C, Zero) (src/Operators.h:541-548) over sparse blocks:
    # product and coeff equal those of the dense composition of the blocks
    rng = np.random.default_rng(11)
    n1, n2, r = 500, 300, 0.2
    def make(n, seed):
        p = ab.Particles(3, n)
        p.set("position", torch.from_numpy(synth.uniform_positions(n, 3, seed=seed)))
        p.init_neighbour_search(0.0, 1.0, True)
        return p
    p1, p2 = make(n1, 1), make(n2, 2)
    A = ab.create_sparse_operator(p1, p1, r, K.inv_dist(0.1))
    B = ab.create_sparse_operator(p1, p2, r, K.wendland_c2(0.1))
    Ct = ab.create_sparse_operator(p2, p1, r, K.wendland_c2(0.1))
    Z = ab.create_zero_operator(p2, p2)
    Full = ab.create_block_operator(2, 2, A, B, Ct, Z)
    assert Full.rows() == n1 + n2 and Full.cols() == n1 + n2
    v = torch.from_numpy(rng.random(n1 + n2)).to(p1.device)
    y = (Full * v).cpu().numpy()
    yA = (A * v[:n1]).cpu().numpy() + (B * v[n1:]).cpu().numpy()
    yC = (Ct * v[:n1]).cpu().numpy()
    assert np.array_equal(y[:n1], yA) or rel_l2(y[:n1], yA) <= TOL
    assert np.array_equal(y[n1:], yC)
    # dense composition through coeff: (Full.coeff) x v == Full * v up to the pairs at exactly r (none here)
    ii, jj = np.divmod(np.arange((n1 + n2) * (n1 + n2)), n1 + n2)
    dense = Full.coeff(ii, jj).cpu().numpy().reshape(n1 + n2, n1 + n2)
    assert np.all(dense[n1:, n1:] == 0.0)
    assert rel_l2(dense @ v.cpu().numpy(), y) <= 1e-11
    with pytest.raises(ValueError):
        ab.create_block_operator(2, 2, A, B, Ct)
    with pytest.raises(ValueError):
        ab.create_block_operator(2, 2, A, A, Ct, Z)


def test_accumulate_within_distance_sph_sums():
    # tests/sph.h:295-353: rho[a] = sum(b, norm(dx) < 2h, mass * W(norm(dx), h)) and the vector-valued
    # pressure sum, as AccumulateWithinDistance<std::plus> (src/detail/Contexts.h:247-289)
    N, D = 4000, 3
    h = 1.5 * N ** (-1.0 / 3.0)
    r = 2 * h
    mass, wcon = 1.0 / N, 21.0 / (256.0 * np.pi)
    pos = synth.uniform_positions(N, D, seed=9)
    o, out, p0 = build_both(pos, 0.0, 1.0, [True, True, False])
    p = ab.Particles(D, N, variables={"pdr2": torch.float64})
    p.set("position", torch.from_numpy(pos.copy()))
    p.init_neighbour_search(0.0, 1.0, [True, True, False])
    rho = ab.accumulate_within_distance(p, p, r, K.sph_density(h, mass, wcon)).cpu().numpy()
    rho_ref = o.accumulate_within_distance(out["pos"], orc.K_SPH_DENSITY, [h, mass, wcon], r)
    assert rel_l2(rho, rho_ref) <= TOL
    # non-zero initial value (set_init, src/Symbolic.h:437-439)
    rho7 = ab.accumulate_within_distance(p, p, r, K.sph_density(h, mass, wcon), init=7.0).cpu().numpy()
    assert rel_l2(rho7, o.accumulate_within_distance(out["pos"], orc.K_SPH_DENSITY, [h, mass, wcon], r, init=7.0)) <= TOL
    pdr2 = np.random.default_rng(4).random(N) + 0.5  # P/rho^2 column, in post-reorder order
    p.set("pdr2", torch.from_numpy(pdr2))
    acc = ab.accumulate_within_distance(p, p, r, K.sph_pressure(D, h, mass, wcon, "pdr2")).cpu().numpy()
    acc_ref = o.accumulate_within_distance(out["pos"], orc.K_SPH_PRESSURE, [h, mass, wcon], r, BR=D, row_vars=[pdr2], col_vars=[pdr2])
    assert acc.shape == (N, D)
    assert rel_l2(acc, acc_ref) <= 1e-11  # near-cancelling force sums: looser absolute scale


FAST_CASES = [(1, 14, 0.1, False), (1, 14, 0.1, True), (1, 1000, 0.1, True), (1, 1000, 0.1, False), (2, 1000, 0.1, True),
              (2, 1000, 0.1, False), (2, 1000, 0.5, True), (2, 1000, 0.5, False), (2, 1000, 0.2, True), (2, 1000, 0.2, False),
              (3, 1000, 0.2, True), (3, 1000, 0.2, False), (3, 200000, 0.03, True), (3, 50, 0.9, True), (2, 30, 1.0, True)]


@pytest.mark.parametrize("D,N,r,periodic", FAST_CASES)
def test_bucket_pair_iterator_parity(D, N, r, periodic):
    # get_neighbouring_buckets(query) (src/Search.h:498-764): the device produces the reference
    # iterator's sequence of (i, j, quadrant) bit for bit, and the fast cell-list search over it
    # (tests/neighbours.h:892-951, case list :1330-1366) gives the oracle's / brute-force counts
    rng = np.random.default_rng(17 * D + N)
    pos = rng.uniform(-1.0, 1.0, size=(N, D)).astype(np.float32).astype(np.float64)
    required_bucket_number = N * r ** D / 2.0 ** D
    o, out, p = build_both(pos, -1.0, 1.0, periodic, required_bucket_number)
    bi, bj, qd = o.bucket_pairs()
    gi, gj, gq = p.get_query().neighbouring_buckets()
    assert np.array_equal(gi.cpu().numpy().view(np.uint32), bi)
    assert np.array_equal(gj.cpu().numpy().view(np.uint32), bj)
    assert np.array_equal(gq.cpu().numpy(), qd)
    cnt = p.get_query().fast_bucket_search_counts(r).cpu().numpy().view(np.uint32)
    assert np.array_equal(cnt, o.fast_bucket_search_counts(r))
    if N <= 5000:
        assert np.array_equal(cnt, orc.brute_force_counts(out["pos"], [-1.0] * D, [1.0] * D, periodic, r))


@pytest.mark.parametrize("D,periodic", [(1, True), (2, True), (2, False), (3, True), (3, False)])
@pytest.mark.parametrize("lnorm", [-1, 1, 2])
def test_scale_transform_search_parity(D, periodic, lnorm):
    # distance_search<LN>(query, centre, r, create_scale_transform(scale)) (src/Search.h:794-845,
    # src/Transform.h:140-172): per-query pair sets equal the oracle's, bit for bit
    rng = np.random.default_rng(31 * D + lnorm + int(periodic))
    N = 20000
    pos = rng.uniform(-1.0, 1.0, size=(N, D))
    o, out, p = build_both(pos, -1.0, 1.0, periodic)
    scale = np.array([1.0, 2.5, 0.4])[:D]
    r = 0.12
    queries = np.concatenate([out["pos"][:3000], rng.uniform(-1.2, 1.2, size=(2000, D))])
    cnt, hs = p.distance_search_stats(r, lnorm, queries=queries, scale=scale)
    ocnt, ohs = o.pair_stats_norm(queries, r, lnorm, scale=scale)
    assert np.array_equal(cnt.cpu().numpy().view(np.uint32), ocnt)
    assert np.array_equal(hs.cpu().numpy().view(np.uint64), ohs)
    assert ocnt.sum() > 0
    # tests/neighbours.h:553-561: radius 1 with scale 1/r == radius r (same counts up to pairs at exactly r)
    c_r, _ = p.distance_search_stats(r, 2, queries=queries)
    c_s, _ = p.distance_search_stats(1.0, 2, queries=queries, scale=1.0 / r)
    assert abs(int(c_r.sum()) - int(c_s.sum())) <= 2


@pytest.mark.parametrize("D,N,r,nn,periodic", [(1, 14, 0.1, 1, False), (1, 1000, 0.1, 10, True), (2, 1000, 0.5, 10, True), (2, 1000, 0.5, 10, False),
                                                (2, 1000, 0.2, 10, True), (2, 50000, 0.05, 10, True), (3, 20000, 0.1, 10, True)])
@pytest.mark.parametrize("lnorm", [2, -1])
def test_linear_transform_search_parity(D, N, r, nn, periodic, lnorm):
    # distance_search with create_linear_transform<D>(SkewTransform()) (tests/neighbours.h:1262-1309):
    # per-query pair sets equal the oracle's bit for bit
    rng = np.random.default_rng(D * 100 + N)
    pos = rng.uniform(-1.0, 1.0, size=(N, D)).astype(np.float32).astype(np.float64)
    T = {1: np.array([[0.7]]), 2: np.array([[1.0, 0.3], [0.0, 1.0]]), 3: np.array([[1.0, 0.3, 0.0], [0.0, 1.0, -0.2], [0.1, 0.0, 0.9]])}[D]
    o, out, p = build_both(pos, -1.0, 1.0, periodic, nn)
    queries = np.concatenate([out["pos"][: min(N, 3000)], rng.uniform(-1.2, 1.2, size=(500, D))])
    cnt, hs = p.distance_search_stats(r, lnorm, queries=queries, linear=T)
    ocnt, ohs = o.pair_stats_norm(queries, r, lnorm, linear=T)
    assert np.array_equal(cnt.cpu().numpy().view(np.uint32), ocnt)
    assert np.array_equal(hs.cpu().numpy().view(np.uint64), ohs)
    assert ocnt.sum() > 0
