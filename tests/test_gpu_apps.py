"""Application-level GPU tests: the reference's own end-to-end checks for this
path, run through the drop-in interface.

  * RBF interpolation with a compactly supported kernel
    (tests/rbf_interpolation.h:245-417, BASELINE config c2's application): the saddle
    point system W = [[G, P], [P^T, 0]] built with create_block_operator, solved with a
    Krylov method that only needs W * x, error bounds of the reference test.
  * data-structure invariants (tests/data_structures.h:326-393): every particle lies
    inside the bounds of its bucket; bucket ranges tile the particle array.
  * add / delete with the ordered structure (tests/particle_container.h,
    helper_add_delete_particle; tests/neighbours.h helper_d_random deletions).
"""
import numpy as np
import pytest
import torch

import aboria_b200 as ab
from aboria_b200 import kernels as K
from oracle import oracle as orc
from util import build_both

pytestmark = pytest.mark.gpu


class _OnesOperator:
    """create_dense_operator(rows, cols, [](a, b) { return 1.0; }) of the reference test: the
    polynomial (constant) block P of the RBF system.  Dense operators are outside the
    accelerated path; this tiny host-side stand-in only serves the test."""

    def __init__(self, rows, cols):
        self.row_particles, self.col_particles = rows, cols

    def rows(self):
        return self.row_particles.size()

    def cols(self):
        return self.col_particles.size()

    def evaluate(self, y, b):
        y += b.sum()

    def coeff(self, i, j):
        return torch.ones(len(torch.as_tensor(i).reshape(-1)), dtype=torch.float64, device=self.col_particles.device)


def test_rbf_interpolation_compact():
    from scipy.sparse.linalg import LinearOperator, gmres

    funct = lambda x, y: np.exp(-9 * (x - 0.5) ** 2 - 9 * (y - 0.25) ** 2)  # noqa: E731
    N, hfac = 1000, 4.0
    h = hfac * N ** -0.5
    rng = np.random.default_rng(123)
    pts = rng.random((2 * N, 2))
    knots_pos, test_pos = pts[0::2].copy(), pts[1::2].copy()
    knots = ab.Particles(2, N)
    knots.set("position", torch.from_numpy(knots_pos))
    knots.init_neighbour_search(0.0, 1.0, False)
    test = ab.Particles(2, N)
    test.set("position", torch.from_numpy(test_pos))
    augment = ab.Particles(2, 1)
    dev = knots.device

    kernel = K.wendland_c2(h)
    G = ab.create_sparse_operator(knots, knots, 2 * h, kernel)
    P, Pt = _OnesOperator(knots, augment), _OnesOperator(augment, knots)
    Zero = ab.create_zero_operator(augment, augment)
    W = ab.create_block_operator(2, 2, G, P, Pt, Zero)
    G_test = ab.create_sparse_operator(test, knots, 2 * h, kernel)
    W_test = ab.create_block_operator(2, 2, G_test, _OnesOperator(test, augment), Pt, Zero)
    assert W.rows() == N + 1 and W.cols() == N + 1

    kp = knots.get("position").cpu().numpy()  # post-reorder order: vectors are indexed by it
    phi = np.concatenate([funct(kp[:, 0], kp[:, 1]), [0.0]])

    def matvec(x):
        return (W * torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)).cpu().numpy()

    A = LinearOperator((N + 1, N + 1), matvec=matvec, dtype=np.float64)
    gamma, info = gmres(A, phi, restart=101, maxiter=4, rtol=1e-12, atol=0.0)
    assert info >= 0
    gt = torch.from_numpy(gamma).to(dev)
    ev = (W * gt).cpu().numpy()[:N]
    truth = funct(kp[:, 0], kp[:, 1])
    rms_centres = np.sqrt(((ev - truth) ** 2).sum() / (truth ** 2).sum())
    assert rms_centres < 1e-4, rms_centres  # tests/rbf_interpolation.h:395
    ev_t = (W_test * gt).cpu().numpy()[:N]
    truth_t = funct(test_pos[:, 0], test_pos[:, 1])
    rms_away = np.sqrt(((ev_t - truth_t) ** 2).sum() / (truth_t ** 2).sum())
    assert rms_away < 1e-2, rms_away        # tests/rbf_interpolation.h:412


@pytest.mark.parametrize("D,N,nn,periodic", [(2, 20, 5, False), (3, 5000, 10, True), (1, 300, 3, False), (3, 200000, 10, [True, False, True])])
def test_data_structure_invariants(D, N, nn, periodic):
    rng = np.random.default_rng(N)
    pos = rng.random((N, D))
    o, out, p = build_both(pos, 0.0, 1.0, periodic, nn)
    size, side, nb = p.grid()
    q = p.get_query()
    bb = q.bucket_begin.cpu().numpy().astype(np.int64)
    be = q.bucket_end.cpu().numpy().astype(np.int64)
    keys = q.bucket_indices.cpu().numpy().astype(np.int64)[: p.size()]
    pp = p.get("position").cpu().numpy()
    # bucket ranges tile [0, n) in bucket order; a particle's range is its key's
    assert bb[0] == 0 and be[-1] == p.size()
    assert np.array_equal(bb[1:], be[:-1])
    k = np.arange(p.size())
    assert np.all((bb[keys] <= k) & (k < be[keys]))
    # every particle lies inside the bounds of its bucket (tests/data_structures.h:381-387):
    # bounds = bmin + v * side .. bmin + (v + 1) * side, '<=' on both sides
    v = np.stack(np.unravel_index(keys, tuple(int(s) for s in size)), axis=1)
    lo = 0.0 + v * side
    hi = 0.0 + (v + 1) * side
    assert np.all(lo <= pp + 1e-15) and np.all(pp <= hi + 1e-15)


def test_add_delete_with_ordered_structure():
    # helper_add_delete_particle (tests/particle_container.h) / the deletions of helper_d_random
    # (tests/neighbours.h:1099-1147): delete by clearing `alive` + update_positions, add by growing
    # the container + update_positions; ids survive, neighbour counts stay equal to brute force
    rng = np.random.default_rng(42)
    N, r = 3000, 0.12
    pos = rng.uniform(-1.0, 1.0, size=(N, 3))
    p = ab.Particles(3, N)
    p.set("position", torch.from_numpy(pos.copy()))
    p.init_neighbour_search(-1.0, 1.0, True)
    assert p.size() == N

    def check():
        sp = p.get("position").cpu().numpy()
        cnt, _ = p.pair_stats(r)
        bf = orc.brute_force_counts(sp, [-1.0] * 3, [1.0] * 3, True, r)
        assert np.array_equal(cnt.cpu().numpy().view(np.uint32), bf)

    check()
    # delete every 7th particle (by id) and one in the middle
    ids = p.get("id").cpu().numpy()
    kill = (ids % 7 == 0) | (ids == 1234)
    alive = p.get("alive").clone()
    alive[torch.from_numpy(kill).to(alive.device)] = 0
    p.set("alive", alive)
    n_after = p.update_positions()
    assert n_after == N - int(kill.sum()) and p.size() == n_after
    assert not np.any(np.isin(p.get("id").cpu().numpy(), ids[kill]))
    assert np.all(p.get("alive").cpu().numpy() == 1)
    check()
    # add 500 new particles (new ids), some outside the periodic box (they get wrapped)
    M = 500
    new_pos = rng.uniform(-1.5, 1.5, size=(M, 3))
    allpos = torch.cat([p.get("position"), torch.from_numpy(new_pos).to(p.device)])
    allid = torch.cat([p.get("id"), torch.arange(N, N + M, dtype=torch.int64, device=p.device)])
    p.columns = {"position": allpos.contiguous(), "id": allid.contiguous(), "alive": torch.ones(n_after + M, dtype=torch.uint8, device=p.device)}
    p._other = {}
    assert p.update_positions() == n_after + M
    got_ids = np.sort(p.get("id").cpu().numpy())
    assert np.array_equal(got_ids, np.sort(np.concatenate([ids[~kill], np.arange(N, N + M)])))
    sp = p.get("position").cpu().numpy()
    assert np.all((sp >= -1.0) & (sp < 1.0))
    check()


def test_md_linear_spring_steps():
    # tests/md.h:100-190 (helper_md_iterator): N discs in a periodic square, linear spring
    # repulsion within `diameter`, explicit velocity / position update, update_positions every
    # step.  The same loop driven through the drop-in interface on the GPU and through the
    # oracle on the host: trajectories agree (by particle id) after several steps.
    D, N, steps = 2, 2000, 8
    diameter, k_spring, dt, mass = 0.03, 1.0e2, 1e-3, 1.0
    rng = np.random.default_rng(7)
    pos0 = rng.random((N, D))
    vel0 = rng.normal(scale=0.5, size=(N, D))

    # --- GPU, reference-shaped loop ---
    p = ab.Particles(D, N, variables={"velocity": (torch.float64, (D,))})
    p.set("position", torch.from_numpy(pos0.copy()))
    p.set("velocity", torch.from_numpy(vel0.copy()))
    p.init_neighbour_search(0.0, 1.0, True)
    spring = K.linear_spring(D, k_spring, diameter)
    for _ in range(steps):
        force = ab.accumulate_within_distance(p, p, diameter, spring)     # sum over neighbours of -k (d/r - 1) dx
        p.get("velocity").add_(force, alpha=dt / mass)
        p.get("position").add_(p.get("velocity"), alpha=dt)
        p.update_positions()
    assert p.size() == N
    order = np.argsort(p.get("id").cpu().numpy())
    gpos = p.get("position").cpu().numpy()[order]
    gvel = p.get("velocity").cpu().numpy()[order]

    # --- oracle, same loop on the host ---
    o = orc.Oracle(D)
    pos, vel, ids = pos0.copy(), vel0.copy(), np.arange(N)
    out = o.init_neighbour_search(pos, 0.0, 1.0, True)
    pos, vel, ids = out["pos"].copy(), vel[out["order"]], ids[out["order"]]
    for _ in range(steps):
        force = o.accumulate_within_distance(pos, orc.K_LINEAR_SPRING, [k_spring, diameter], diameter, BR=D)
        vel = vel + (dt / mass) * force
        pos = pos + dt * vel
        out = o.init_neighbour_search(pos, 0.0, 1.0, True)
        pos, vel, ids = out["pos"].copy(), vel[out["order"]], ids[out["order"]]
    oo = np.argsort(ids)
    assert np.abs(gpos - pos[oo]).max() < 1e-11
    assert np.abs(gvel - vel[oo]).max() < 1e-9
    # the discs did interact
    assert np.abs(gvel - vel0).max() > 1e-3


def test_host_pipeline_matches_step_by_step():
    # aboria_b200.pipeline.HostPipeline: overlapped host-buffer steps give, for every step, exactly
    # what the plain sequence (upload, init_neighbour_search, K * b, download) gives
    from aboria_b200.pipeline import HostPipeline

    N, steps = 50000, 7
    rng = np.random.default_rng(21)
    side = (10.0 / N) ** (1.0 / 3.0)
    r = 1.3 * side
    kern = K.inv_dist(0.1)
    pos_h = [torch.from_numpy(rng.random((N, 3))).pin_memory() for _ in range(steps)]
    b_h = [torch.from_numpy(rng.random(N)).pin_memory() for _ in range(steps)]
    y_h = [torch.empty(N, dtype=torch.float64).pin_memory() for _ in range(steps)]
    pipe = HostPipeline(3, N, 0.0, 1.0, True, r, kern)
    for k in range(steps):
        pipe.submit(pos_h[k], b_h[k], y_h[k])
    pipe.wait()
    p = ab.Particles(3, N)
    op = ab.create_sparse_operator(p, p, r, kern)
    for k in range(steps):
        p.resize_from_positions(pos_h[k])
        p.init_neighbour_search(0.0, 1.0, True)
        y = op.matvec(b_h[k].to(p.device)).cpu()
        assert torch.equal(y, y_h[k]), k
    # a step in which a particle leaves a non-periodic domain is reported, not multiplied
    pipe2 = HostPipeline(3, N, 0.0, 1.0, False, r, kern)
    bad = pos_h[0].clone()
    bad[17, 1] = 1.5
    pipe2.submit(bad.pin_memory(), b_h[0], y_h[0])
    with pytest.raises(ab.AbrError):
        pipe2.wait()
