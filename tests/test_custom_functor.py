"""The user-functor path (abr_sparse_matvec_custom + abr::sparse_launcher): a device functor
compiled by nvcc in the USER's translation unit (tests/cpp/custom_functor.cu) reads a
per-particle column of the row and of the column particle, as the lambdas of the reference's
create_sparse_operator do (src/Operators.h:478-516; tests/operators.h:251-256).  Checked
against the oracle (same kernel = K_INV_DIST_AA / K_CONST_SUM_DIFF there)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "cpp", "libcustom_functor.so")
TOL = 1e-12


def test_custom_functor_library_is_built():
    if not os.path.exists(SO):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(SO)


def _lib():
    from aboria_b200 import _lib as abl

    abl.lib()  # libabr.so first (RTLD_GLOBAL not needed: the test library carries an rpath)
    L = C.CDLL(SO)
    vp = C.c_void_p
    L.custom_weighted_inv_dist.restype = C.c_int
    L.custom_weighted_inv_dist.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, vp, C.c_double, C.c_double, vp, vp, C.POINTER(C.c_uint64)]
    L.custom_sum_diff.restype = C.c_int
    L.custom_sum_diff.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, vp, C.c_double, vp, vp]
    return L


@pytest.mark.gpu
@pytest.mark.parametrize("N,r,periodic", [(20000, 0.08, True), (5000, 0.11, False), (300000, 0.02, True)])
def test_custom_functor_rows_are_cols(N, r, periodic):
    import aboria_b200 as ab
    from aboria_b200 import synth
    from aboria_b200._lib import check
    from oracle import oracle as orc
    from util import rel_l2

    L = _lib()
    pos = synth.uniform_positions(N, 3)
    rng = np.random.default_rng(7)
    w = rng.uniform(0.5, 1.5, N)
    b = synth.vector(N)
    o = orc.Oracle(3)
    out = o.init_neighbour_search(pos, 0.0, 1.0, periodic)
    order = out["order"]
    w_s, b_s = w[order], b[order]
    y_o, npairs_o = o.sparse_matvec(out["pos"], orc.K_INV_DIST_AA, [0.1], r, b_s, row_vars=[w_s], col_vars=[w_s])

    p = ab.Particles(3, N, variables={"w": torch.float64})
    p.set("position", torch.from_numpy(pos.copy()))
    p.set("w", torch.from_numpy(w.copy()))
    p.init_neighbour_search(0.0, 1.0, periodic)
    assert np.array_equal(p.get("w").cpu().numpy(), w_s)  # the column followed the reorder
    dev = p.device
    bt = torch.from_numpy(b_s).to(dev)
    y = torch.zeros(N, dtype=torch.float64, device=dev)
    npairs = C.c_uint64(0)
    p._sync_stream()
    wt, pt = p.get("w"), p.get("position")
    check(p._h, L.custom_weighted_inv_dist(p._h, pt.data_ptr(), N, 1, wt.data_ptr(), wt.data_ptr(), 0.1, r, bt.data_ptr(), y.data_ptr(), C.byref(npairs)))
    torch.cuda.synchronize()
    assert npairs.value == npairs_o
    assert p.last_counters()["launches"] >= 2  # the cell-tiled kernel + its exact-walk pass (+ a heavy-bucket launch), not the per-row walk (1)
    assert rel_l2(y.cpu().numpy(), y_o) <= TOL

    # 2 x 1 block functor
    s1 = rng.uniform(-1, 1, N)
    y2_o, _ = o.sparse_matvec(out["pos"], orc.K_CONST_SUM_DIFF, [], r, b_s, BR=2, BC=1, row_vars=[s1[order]], col_vars=[s1[order]])
    st = torch.from_numpy(s1[order].copy()).to(dev)
    y2 = torch.zeros(2 * N, dtype=torch.float64, device=dev)
    check(p._h, L.custom_sum_diff(p._h, pt.data_ptr(), N, 1, st.data_ptr(), st.data_ptr(), r, bt.data_ptr(), y2.data_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(y2.cpu().numpy(), y2_o) <= TOL


@pytest.mark.gpu
def test_custom_functor_rows_not_cols():
    # row set without a search structure (tests/rbf_interpolation.h:326): the functor's row
    # column belongs to the ROW set, the column one to the searched set
    import aboria_b200 as ab
    from aboria_b200 import synth
    from aboria_b200._lib import check
    from oracle import oracle as orc
    from util import rel_l2

    L = _lib()
    N, M, r = 30000, 4000, 0.07
    pos = synth.uniform_positions(N, 3)
    rng = np.random.default_rng(11)
    rows = rng.uniform(0.0, 1.0, (M, 3))
    w_col, w_row = rng.uniform(0.5, 1.5, N), rng.uniform(0.5, 1.5, M)
    b = synth.vector(N)
    o = orc.Oracle(3)
    out = o.init_neighbour_search(pos, 0.0, 1.0, True)
    order = out["order"]
    y_o, npairs_o = o.sparse_matvec(rows, orc.K_INV_DIST_AA, [0.1], r, b[order], row_vars=[w_row], col_vars=[w_col[order]])
    p = ab.Particles(3, N, variables={"w": torch.float64})
    p.set("position", torch.from_numpy(pos.copy()))
    p.set("w", torch.from_numpy(w_col.copy()))
    p.init_neighbour_search(0.0, 1.0, True)
    dev = p.device
    rt = torch.from_numpy(rows).to(dev)
    wr = torch.from_numpy(w_row).to(dev)
    bt = torch.from_numpy(b[order]).to(dev)
    y = torch.zeros(M, dtype=torch.float64, device=dev)
    npairs = C.c_uint64(0)
    p._sync_stream()
    check(p._h, L.custom_weighted_inv_dist(p._h, rt.data_ptr(), M, 0, wr.data_ptr(), p.get("w").data_ptr(), 0.1, r, bt.data_ptr(), y.data_ptr(), C.byref(npairs)))
    torch.cuda.synchronize()
    assert npairs.value == npairs_o
    assert rel_l2(y.cpu().numpy(), y_o) <= TOL
