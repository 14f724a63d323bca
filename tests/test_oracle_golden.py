"""Pins the CPU oracle against the reference's own known-answer tests
(SURVEY.md §4 / §8c).  Each test cites the reference test it ports."""
import numpy as np
import pytest

from oracle import oracle as orc


def test_bucket_indicies():
    # tests/utils.h:52-74
    assert orc.collapse_index_vector([4, 7, 2], [1, 2, 1]) == 19
    assert orc.collapse_index_vector([4, 7, 1, 6], [1, 2, 0, 4]) == 58


def test_point_to_bucket_indicies():
    # tests/utils.h:76-93
    o = orc.Oracle(3)
    o.force_grid(0.0, 1.0, False, 5)
    idx, v = o.point_to_bucket_index([0.5, 0.5, 0.5])
    assert list(v) == [2, 2, 2]
    assert idx == 2 * 5 * 5 + 2 * 5 + 2


def test_lattice_within_distance_1d():
    # tests/iterators.h:49-116 (100 particles, n_leaf 10 -> 10 buckets)
    rng = np.random.default_rng(0)
    o = orc.Oracle(1)
    o.init_neighbour_search(rng.random((100, 1)), 0.0, 1.0, False)
    size, side = o.grid()
    assert list(size) == [10]
    for point, r, expect in [(0.5, 0.05, 2), (0.5, 0.15, 4), (0.1, 0.15, 3), (-0.1, 0.05, 0), (-0.1, 0.15, 1)]:
        assert o.buckets_near_point([point], r)[0] == expect


def test_lattice_within_distance_2d():
    # tests/iterators.h:118-211 (1000 particles -> 10 x 10 buckets)
    rng = np.random.default_rng(1)
    o = orc.Oracle(2)
    o.init_neighbour_search(rng.random((1000, 2)), 0.0, 1.0, False)
    size, side = o.grid()
    assert list(size) == [10, 10]
    cases = [((0.5, 0.5), 0.05, 4), ((0.5, 0.5), 0.1001, 12), ((0.55, 0.55), 0.1001, 9), ((0.55, 0.55), 0.049999, 1),
             ((-0.001, 0.001), 2.0, 100), ((-0.001, 0.001), 0.01, 1), ((1.001, 1.001), 0.01, 1)]
    for point, r, expect in cases:
        assert o.buckets_near_point(point, r)[0] == expect, (point, r)


def test_single_particle():
    # tests/neighbours.h:520-557
    radius = 0.1
    o = orc.Oracle(3)
    o.init_neighbour_search([[0.0, 0.0, 0.0]], -1.0, 1.0, True)
    assert o.search_point([radius / 2, radius / 2, 0], radius)[0] == 1
    assert o.search_point([2 * radius, 0, 0], radius)[0] == 0


def test_two_particles():
    # tests/neighbours.h:559-605
    radius = 0.1
    o = orc.Oracle(3)
    out = o.init_neighbour_search([[0.0, 0.0, 0.0], [radius / 2, 0, 0]], -1.0, 1.0, True)
    c, j, im, dx = o.search_point([1.1 * radius, 0, 0], radius)
    assert c == 1 and out["order"][j[0]] == 1  # found particle has id 1
    assert o.search_point([0.9 * radius, 0, 0], radius)[0] == 2
    assert o.search_point([1.6 * radius, 0, 0], radius)[0] == 0
    assert o.search_point([0.25 * radius, 0.9 * radius, 0], radius)[0] == 2
    assert o.search_point([0.25 * radius, 0.99 * radius, 0], radius)[0] == 0


def _lattice(D, n):
    # tests/neighbours.h:641-661: pos = index*dx + min + dx/2, first index fastest
    idx = np.indices((n,) * D).reshape(D, -1).T[:, ::-1]
    return idx.astype(np.float64) * 1.0 + 0.0 + 0.5


@pytest.mark.parametrize("D,n,r,nn", [(1, 100, 1.5, 10), (2, 50, 1.0001, 10), (2, 50, 1.5, 10), (2, 20, 2.1, 10), (3, 10, 1.9, 10), (3, 10, 1.0001, 10)])
def test_helper_d_regular(D, n, r, nn):
    # tests/neighbours.h:627-686 with the case list :1252-1260 (L2 part: the
    # Gauss circle count in 2-D; in other D compared with brute force)
    pos = _lattice(D, n)
    o = orc.Oracle(D)
    out = o.init_neighbour_search(pos, 0.0, float(n), True, nn)
    cnt, _ = o.pair_stats(out["pos"], r)
    if D == 2:
        n_expect = 0
        for i in range(100):
            n_expect += int(np.floor(r**2 / (4 * i + 1))) - int(np.floor(r**2 / (4 * i + 3)))
        n_expect = 1 + 4 * n_expect
        assert n_expect == {1.0001: 5, 1.5: 9, 2.1: 13}[r]
        assert np.all(cnt == n_expect)
    brute = orc.brute_force_counts(out["pos"], [0.0] * D, [float(n)] * D, True, r)
    assert np.array_equal(cnt, brute)


@pytest.mark.parametrize("D,N,r,nn,periodic", [
    (1, 14, 0.1, 1, False), (1, 14, 0.1, 1, True), (1, 1000, 0.1, 10, True), (1, 1000, 0.1, 10, False),
    (1, 1000, 0.1, 100, True), (1, 1000, 0.1, 100, False), (2, 1000, 0.5, 10, True), (2, 1000, 0.5, 10, False),
    (2, 1000, 0.2, 1, True), (2, 1000, 0.2, 1, False), (3, 1000, 0.2, 100, True), (3, 1000, 0.2, 100, False),
    (3, 1000, 0.2, 10, True), (3, 1000, 0.2, 10, False), (3, 1000, 0.2, 1, True), (3, 1000, 0.2, 1, False)])
@pytest.mark.parametrize("sort_mode", [orc.SORT_STD, orc.SORT_STABLE])
def test_helper_d_random(D, N, r, nn, periodic, sort_mode):
    # tests/neighbours.h:968-1147 with cases :1273-1327 (IdentityTransform);
    # positions are float32 values in [-1,1) as in the reference (:1000-1004)
    rng = np.random.default_rng(1234 + D * 100 + N + int(periodic))
    pos = rng.uniform(-1.0, 1.0, size=(N, D)).astype(np.float32).astype(np.float64)
    o = orc.Oracle(D)
    out = o.init_neighbour_search(pos, -1.0, 1.0, periodic, nn, sort_mode=sort_mode)
    assert out["n_alive"] == N
    cnt, _ = o.pair_stats(out["pos"], r)
    brute = orc.brute_force_counts(out["pos"], [-1.0] * D, [1.0] * D, periodic, r)
    assert np.array_equal(cnt, brute)
    # helper_data_structure (tests/data_structures.h:326-393): every particle lies
    # in its bucket; bucket ranges partition the sorted array
    bb, be, keys = out["bucket_begin"], out["bucket_end"], out["keys"]
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
    assert int((be.astype(np.int64) - bb).sum()) == N
    for c in np.unique(keys):
        assert np.all(keys[bb[c]:be[c]] == c)
    # the two sort modes agree on everything but the within-cell order
    assert sorted(out["order"].tolist()) == list(range(N))


def test_sparse_operator_golden():
    # tests/operators.h:810-933
    diameter = 0.1
    pos = np.array([[0, 0, 0], [diameter * 0.9, 0, 0], [diameter * 1.8, 0, 0]], dtype=np.float64)
    o = orc.Oracle(3)
    out = o.init_neighbour_search(pos, -1.0, 1.0, False)
    order = out["order"]
    s1 = np.full(3, 1.0)
    s2 = np.full(3, 2.0)
    v = np.array([1.0, 2.0, 3.0])
    # Eigen vectors are indexed by post-reorder position
    y, npairs = o.sparse_matvec(out["pos"], orc.K_CONST_SUM, [], diameter, v, row_vars=[s1], col_vars=[s2])
    assert npairs == 7  # C_sparse.nonZeros() == 7
    ids = order
    expect = np.zeros(3)
    for i in range(3):
        for j in range(3):
            if {int(ids[i]), int(ids[j])} == {0, 2}:
                continue
            expect[i] += 3.0 * v[j]
    assert np.array_equal(y, expect)
    if list(ids) == [0, 1, 2]:
        assert list(y) == [9.0, 18.0, 15.0]
    y2, npairs2 = o.sparse_matvec(out["pos"], orc.K_CONST_SUM_DIFF, [], diameter, v, BR=2, BC=1, row_vars=[s1], col_vars=[s2])
    assert npairs2 == 7  # 14 scalar non-zeros
    if list(ids) == [0, 1, 2]:
        # the reference asserts only the first n=3 entries of its `check`
        # vector {9,-3,18,-7,15,-5} (tests/operators.h:928-931); the 4th entry
        # is a typo in the reference (-1-2-3 = -6) that is never compared.
        assert list(y2[:3]) == [9.0, -3.0, 18.0]
        assert list(y2) == [9.0, -3.0, 18.0, -6.0, 15.0, -5.0]


def test_documentation_operator():
    # tests/operators.h:121-311: N=100 uniform in the unit cube, eps=0.1,
    # r=0.1, kernel a_i a_j/(|dx|+eps): K_s*b equals the assembled matrix * b
    N, eps, r = 100, 0.1, 0.1
    rng = np.random.default_rng(7)
    pos = rng.random((N, 3))
    a = rng.random(N)
    o = orc.Oracle(3)
    out = o.init_neighbour_search(pos, 0.0, 1.0, False)
    ps, as_ = out["pos"], a[out["order"]]
    b = np.linspace(0, 1.0, N)
    y, _ = o.sparse_matvec(ps, orc.K_INV_DIST_AA, [eps], r, b, row_vars=[as_], col_vars=[as_])
    dx = ps[None, :, :] - ps[:, None, :]
    d2 = (dx[..., 0] * dx[..., 0] + dx[..., 1] * dx[..., 1]) + dx[..., 2] * dx[..., 2]
    K = np.where(d2 <= r * r, (as_[:, None] * as_[None, :]) / (np.sqrt(d2) + eps), 0.0)
    assert np.allclose(y, K @ b, rtol=1e-13, atol=0)


def test_enforce_domain_and_dead():
    # src/NeighbourSearchBase.h:185-238: periodic wrap, non-periodic kill, non-finite kill
    pos = np.array([[1.25, 0.5], [-0.25, 0.5], [0.5, 1.5], [0.5, np.nan], [0.1, 0.2], [3.75, 0.999]], dtype=np.float64)
    o = orc.Oracle(2)
    o.set_domain([0.0, 0.0], [1.0, 1.0], [True, False], 1.0)
    p = pos.copy()
    out = o.update_positions(p)
    assert list(out["alive"]) == [1, 1, 0, 0, 1, 1]
    assert p[0, 0] == 0.25 and p[1, 0] == 0.75 and p[5, 0] == 0.75
    assert sorted(out["order"].tolist()) == [0, 1, 4, 5]
    assert out["n_alive"] == 4


def test_sparse_operator_assemble_golden():
    # tests/operators.h:852-873: C.assemble(dense / sparse): 7 block entries, all equal to 3;
    # the 2x1 operator has 14 scalar non-zeros (tests/operators.h:934-938)
    diameter = 0.1
    pos = np.array([[0, 0, 0], [diameter * 0.9, 0, 0], [diameter * 1.8, 0, 0]], dtype=np.float64)
    o = orc.Oracle(3)
    out = o.init_neighbour_search(pos, -1.0, 1.0, False)
    s1, s2 = np.full(3, 1.0), np.full(3, 2.0)
    rp, col, val = o.assemble(out["pos"], orc.K_CONST_SUM, [], diameter, row_vars=[s1], col_vars=[s2])
    assert len(col) == 7 and np.all(val == 3.0)
    dense = np.zeros((3, 3))
    for i in range(3):
        for k in range(rp[i], rp[i + 1]):
            dense[i, col[k]] = val[k, 0, 0]
    assert np.array_equal(dense, np.array([[3, 3, 0], [3, 3, 3], [0, 3, 3.0]]))
    rp2, col2, val2 = o.assemble(out["pos"], orc.K_CONST_SUM_DIFF, [], diameter, BR=2, BC=1, row_vars=[s1], col_vars=[s2])
    assert len(col2) * 2 == 14 and np.all(val2[:, 0, 0] == 3.0) and np.all(val2[:, 1, 0] == -1.0)
    # assembled matrix times vector equals the matrix-free product (tests/operators.h:866-869)
    v = np.array([1.0, 2.0, 3.0])
    y, _ = o.sparse_matvec(out["pos"], orc.K_CONST_SUM, [], diameter, v, row_vars=[s1], col_vars=[s2])
    assert np.array_equal(dense @ v, y)


@pytest.mark.parametrize("D,n,r", [(1, 100, 1.5), (2, 50, 1.0001), (2, 50, 1.5), (2, 20, 2.1), (3, 10, 1.9), (3, 10, 1.0001)])
def test_helper_d_regular_linf_box_counts(D, n, r):
    # tests/neighbours.h:675-684: Linf (chebyshev) search on the regular lattice finds
    # (2 floor(r) + 1)^D points
    idx = np.indices((n,) * D).reshape(D, -1).T[:, ::-1]
    pos = idx.astype(np.float64) + 0.5
    o = orc.Oracle(D)
    out = o.init_neighbour_search(pos, 0.0, float(n), True, 10)
    cnt, _ = o.pair_stats_norm(out["pos"], r, -1)
    assert np.all(cnt == (2 * int(np.floor(r)) + 1) ** D)
    # L2 through the generic path equals the euclidean path
    c2, h2 = o.pair_stats_norm(out["pos"], r, 2)
    c2e, h2e = o.pair_stats(out["pos"], r)
    assert np.array_equal(c2, c2e) and np.array_equal(h2, h2e)


def test_manhattan_and_chebyshev_random_vs_brute_force():
    rng = np.random.default_rng(99)
    N, D, r = 800, 3, 0.3
    pos = rng.uniform(-1.0, 1.0, size=(N, D)).astype(np.float32).astype(np.float64)
    for periodic in (True, False):
        o = orc.Oracle(D)
        out = o.init_neighbour_search(pos, -1.0, 1.0, periodic, 10)
        ps = out["pos"]
        shifts = np.array(np.meshgrid(*[[-2.0, 0.0, 2.0]] * D, indexing="ij")).reshape(D, -1).T if periodic else np.zeros((1, D))
        for lnorm, fn in ((-1, lambda d: np.abs(d).max(-1)), (1, lambda d: np.abs(d).sum(-1))):
            cnt, _ = o.pair_stats_norm(ps, r, lnorm)
            brute = np.zeros(N, dtype=np.int64)
            for sft in shifts:
                d = ps[None, :, :] - (ps[:, None, :] + sft)
                brute += (fn(d) <= r).sum(1)
            assert np.array_equal(cnt, brute), (lnorm, periodic)


def test_id_search_golden():
    # tests/id_search.h:64-99: N particles with ids 0..N-1 in shuffled order; find(2) points
    # at the particle with id 2, find(2N) at the end of the particle vector
    N = 100
    ids = np.random.default_rng(0).permutation(N).astype(np.uint64)
    key, value = orc.id_map_build(ids)
    assert np.array_equal(key, np.arange(N, dtype=np.uint64))
    assert np.array_equal(ids[value.astype(np.int64)], key)
    found = orc.id_find(key, value, [2, 2 * N])
    assert ids[int(found[0])] == 2
    assert found[1] == N
    # tests/id_search.h:126-205 (helper_d_random): random ids to look up, brute force as the check
    rng = np.random.default_rng(1)
    ids = rng.choice(10 * N, size=N, replace=False).astype(np.uint64)
    key, value = orc.id_map_build(ids)
    q = rng.integers(0, 10 * N, size=500).astype(np.uint64)
    got = orc.id_find(key, value, q)
    for qi, gi in zip(q, got):
        w = np.where(ids == qi)[0]
        assert gi == (w[0] if len(w) else N)


def test_sparse_operator_coeff_golden():
    # tests/operators.h:873-881: C.coeff(i, j) equals the assembled dense matrix
    # [[3,3,0],[3,3,3],[0,3,3]] (and :939-941 for the 2x1 block operator)
    diameter = 0.1
    pos = np.array([[0, 0, 0], [diameter * 0.9, 0, 0], [diameter * 1.8, 0, 0]], dtype=np.float64)
    o = orc.Oracle(3)
    out = o.init_neighbour_search(pos, -1.0, 1.0, False)
    s1, s2 = np.full(3, 1.0), np.full(3, 2.0)
    ii, jj = np.divmod(np.arange(9), 3)
    c = o.coeff(out["pos"], out["pos"], ii, jj, orc.K_CONST_SUM, [], diameter, row_vars=[s1], col_vars=[s2]).reshape(3, 3)
    ids = out["order"]
    expect = np.array([[0.0 if {int(ids[i]), int(ids[j])} == {0, 2} else 3.0 for j in range(3)] for i in range(3)])
    assert np.array_equal(c, expect)
    ii, jj = np.divmod(np.arange(18), 3)
    c2 = o.coeff(out["pos"], out["pos"], ii, jj, orc.K_CONST_SUM_DIFF, [], diameter, BR=2, BC=1, row_vars=[s1], col_vars=[s2]).reshape(6, 3)
    assert np.array_equal(c2[0::2], expect)
    assert np.array_equal(c2[1::2], -expect / 3.0)
    # the coeff predicate is strict and uses the minimum image (src/detail/Kernels.h:358-367):
    # a pair at exactly r is IN the product (<=) but NOT in coeff (<)
    pos = np.array([[0.125, 0.5, 0.5], [0.375, 0.5, 0.5], [0.9375, 0.5, 0.5]])  # exactly representable
    o = orc.Oracle(3)
    out = o.init_neighbour_search(pos, 0.0, 1.0, True)
    sp = out["pos"]
    r = 0.25
    ii, jj = np.divmod(np.arange(9), 3)
    c = o.coeff(sp, sp, ii, jj, orc.K_CONST_SUM, [], r, row_vars=[s1], col_vars=[s2]).reshape(3, 3)
    cnt, _ = o.pair_stats(sp, r)
    x = sp[:, 0]
    a, b, cc = [int(np.argmin(np.abs(x - v))) for v in (0.125, 0.375, 0.9375)]
    assert c[a, b] == 0.0 and c[b, a] == 0.0          # |dx| == r exactly: excluded by '<'
    assert c[a, cc] == 3.0 and c[cc, a] == 3.0        # 0.125 <-> 0.9375 through the periodic boundary (|dx| = 0.1875)
    assert cnt[a] == 3 and cnt[b] == 2 and cnt[cc] == 2   # the search (<=) does take the pair at exactly r


FAST_CASES = [(1, 14, 0.1, False), (1, 14, 0.1, True), (1, 1000, 0.1, True), (1, 1000, 0.1, False), (2, 1000, 0.1, True),
              (2, 1000, 0.1, False), (2, 1000, 0.5, True), (2, 1000, 0.5, False), (2, 1000, 0.2, True), (2, 1000, 0.2, False),
              (3, 1000, 0.2, True), (3, 1000, 0.2, False)]


@pytest.mark.parametrize("D,N,r,periodic", FAST_CASES)
def test_fast_bucketsearch_vs_brute_force(D, N, r, periodic):
    # tests/neighbours.h:1149-1248 with the case list :1330-1366: neighbour counts found through
    # get_neighbouring_buckets (bucket_pair_iterator, src/Search.h:498-764) equal the brute-force counts
    rng = np.random.default_rng(17 * D + N)
    pos = rng.uniform(-1.0, 1.0, size=(N, D)).astype(np.float32).astype(np.float64)
    o = orc.Oracle(D)
    required_bucket_number = N * r ** D / 2.0 ** D
    out = o.init_neighbour_search(pos, -1.0, 1.0, periodic, required_bucket_number)
    _, side = o.grid()
    assert np.all(side >= r)  # the assumption of the fast search
    bi, bj, qd = o.bucket_pairs()
    assert len(bi) > 0 and np.all(bi < len(out["bucket_begin"])) and np.all(bj < len(out["bucket_begin"]))
    # every unordered pair of touching buckets appears exactly once (no duplicates) when the grid has >= 3 buckets per side
    size, _ = o.grid()
    if np.all(size >= 3):
        keys = set()
        for a, b, q in zip(bi.tolist(), bj.tolist(), map(tuple, qd.tolist())):
            assert (a, b, q) not in keys
            keys.add((a, b, q))
    cnt = o.fast_bucket_search_counts(r)
    bf = orc.brute_force_counts(out["pos"], [-1.0] * D, [1.0] * D, periodic, r)
    assert np.array_equal(cnt, bf)


def test_scale_transform_golden():
    # tests/neighbours.h:553-561: euclidean_search(query, centre, 1.0, create_scale_transform(1/radius))
    # finds what euclidean_search(query, centre, radius) finds
    radius = 0.1
    o = orc.Oracle(3)
    o.init_neighbour_search([[0.0, 0.0, 0.0]], -1.0, 1.0, True)
    cnt, _ = o.pair_stats_norm(np.array([[radius / 2, radius / 2, 0.0], [2 * radius, 0.0, 0.0]]), 1.0, 2, scale=1.0 / radius)
    assert cnt.tolist() == [1, 0]
    # random cloud, anisotropic scale: against a brute force over the periodic images with the same
    # transformed distance (the reference's own check, tests/neighbours.h:739-764, applies the transform
    # to dx before the norm)
    rng = np.random.default_rng(8)
    for D, periodic in [(2, True), (3, False), (3, True)]:
        N = 400
        pos = rng.uniform(-1.0, 1.0, size=(N, D))
        scale = np.array([1.0, 2.0, 0.5])[:D]
        r = 0.3
        o = orc.Oracle(D)
        out = o.init_neighbour_search(pos, -1.0, 1.0, periodic, 5)
        sp = out["pos"]
        cnt, _ = o.pair_stats_norm(sp, r, 2, scale=scale)
        images = np.array(np.meshgrid(*[[-1, 0, 1] if periodic else [0]] * D, indexing="ij")).reshape(D, -1).T
        bf = np.zeros(N, dtype=np.int64)
        for im in images:
            d = (sp[None, :, :] - (sp[:, None, :] + im * 2.0)) * scale
            bf += ((d * d).sum(-1) <= r * r).sum(1)
        assert np.array_equal(cnt, bf)
        # identity scale reproduces the untransformed search bit for bit
        c1, h1 = o.pair_stats_norm(sp, r, 2, scale=1.0)
        c0, h0 = o.pair_stats_norm(sp, r, 2)
        assert np.array_equal(c1, c0) and np.array_equal(h1, h0)


SKEW_CASES = [(1, 14, 0.1, 1, False), (1, 14, 0.1, 1, True), (1, 1000, 0.1, 10, True), (1, 1000, 0.1, 10, False),
              (2, 1000, 0.5, 10, True), (2, 1000, 0.5, 10, False), (2, 1000, 0.2, 10, True), (2, 1000, 0.2, 10, False)]


def _skew(D):
    # tests/neighbours.h:1262-1267 SkewTransform: 1-D 0.7 v; 2-D (v0 + 0.3 v1, v1)
    return np.array([[0.7]]) if D == 1 else np.array([[1.0, 0.3], [0.0, 1.0]])


@pytest.mark.parametrize("D,N,r,nn,periodic", SKEW_CASES)
def test_linear_transform_vs_brute_force(D, N, r, nn, periodic):
    # helper_d_random(..., skew) (tests/neighbours.h:968-1147, cases :1274-1309): the search with a
    # LinearTransform finds what the transform-aware brute force of :739-764 finds
    rng = np.random.default_rng(D * 100 + N)
    pos = rng.uniform(-1.0, 1.0, size=(N, D)).astype(np.float32).astype(np.float64)
    T = _skew(D)
    o = orc.Oracle(D)
    out = o.init_neighbour_search(pos, -1.0, 1.0, periodic, nn)
    sp = out["pos"]
    cnt, _ = o.pair_stats_norm(sp, r, 2, linear=T)
    images = np.array(np.meshgrid(*[[-1, 0, 1] if periodic else [0]] * D, indexing="ij")).reshape(D, -1).T
    bf = np.zeros(N, dtype=np.int64)
    for im in images:
        d = (sp[None, :, :] - sp[:, None, :] - im * 2.0) @ T.T
        bf += ((d * d).sum(-1) <= r * r).sum(1)
    assert np.array_equal(cnt, bf)
