"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``aboria_b200`` never does.  See the header of ``aboria_oracle.cpp``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

K_CONST_SUM, K_CONST_SUM_DIFF, K_INV_DIST, K_INV_DIST_AA = 0, 1, 2, 3
K_WENDLAND_C2, K_LJ_FORCE, K_SPH_DENSITY, K_SPH_PRESSURE = 4, 5, 6, 7
K_LINEAR_SPRING = 8

SORT_STD, SORT_STABLE = 0, 1


def build():
    """Compile liboracle with the committed Makefile (g++ only)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libaboria_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, dp, u8p, ip, up = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_uint)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_collapse_index_vector.restype = C.c_int
        L.orc_collapse_index_vector.argtypes = [C.c_int, up, ip]
        L.orc_set_domain.argtypes = [vp, dp, dp, u8p, C.c_double]
        L.orc_force_grid.argtypes = [vp, dp, dp, u8p, up]
        L.orc_get_grid.argtypes = [vp, up, dp]
        L.orc_point_to_bucket_index.restype = C.c_int
        L.orc_point_to_bucket_index.argtypes = [vp, dp, ip]
        L.orc_update_positions.restype = C.c_long
        L.orc_update_positions.argtypes = [vp, dp, u8p, C.c_size_t, C.c_int]
        L.orc_num_buckets.restype = C.c_size_t
        L.orc_num_buckets.argtypes = [vp]
        L.orc_get_build.argtypes = [vp, ip, up, up, up]
        L.orc_gather.argtypes = [ip, C.c_size_t, vp, vp, C.c_size_t]
        L.orc_update_iterators.argtypes = [vp, dp, C.c_size_t]
        L.orc_buckets_near_point.restype = C.c_long
        L.orc_buckets_near_point.argtypes = [vp, dp, C.c_double, ip, C.c_long]
        L.orc_search_point.restype = C.c_long
        L.orc_search_point.argtypes = [vp, dp, C.c_double, ip, ip, dp, C.c_long]
        L.orc_pair_stats.argtypes = [vp, dp, C.c_size_t, C.c_double, dp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_pair_stats_norm.restype = C.c_int
        L.orc_pair_stats_norm.argtypes = [vp, dp, C.c_size_t, C.c_double, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_pair_stats_norm_scaled.restype = C.c_int
        L.orc_pair_stats_norm_scaled.argtypes = [vp, dp, C.c_size_t, C.c_double, C.c_int, dp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_pair_stats_norm_linear.restype = C.c_int
        L.orc_pair_stats_norm_linear.argtypes = [vp, dp, C.c_size_t, C.c_double, C.c_int, dp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_sparse_matvec.restype = C.c_uint64
        L.orc_sparse_matvec.argtypes = [vp, dp, C.c_size_t, C.c_int, dp, C.POINTER(dp), C.POINTER(dp), C.c_double, dp, C.c_int, C.c_int, dp, dp, C.c_int]
        L.orc_sparse_assemble.restype = C.c_uint64
        L.orc_sparse_assemble.argtypes = [vp, dp, C.c_size_t, C.c_int, dp, C.POINTER(dp), C.POINTER(dp), C.c_double, dp, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), dp]
        L.orc_brute_force_counts.argtypes = [C.c_int, dp, C.c_size_t, dp, dp, C.c_int, C.c_double, C.POINTER(C.c_uint32)]
        u64p = C.POINTER(C.c_uint64)
        L.orc_id_map_build.argtypes = [u64p, C.c_size_t, u64p, u64p]
        L.orc_id_find.argtypes = [u64p, u64p, C.c_size_t, u64p, C.c_size_t, u64p]
        L.orc_sparse_coeff.argtypes = [vp, dp, dp, u64p, u64p, C.c_size_t, C.c_int, dp, C.POINTER(dp), C.POINTER(dp), C.c_double, dp, C.c_int, C.c_int, dp]
        L.orc_bucket_pairs.restype = C.c_uint64
        L.orc_bucket_pairs.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int8), C.c_uint64]
        L.orc_fast_bucket_search_counts.argtypes = [vp, C.c_double, C.POINTER(C.c_uint32)]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def collapse_index_vector(size, v):
    size = np.ascontiguousarray(size, dtype=np.uint32)
    v = np.ascontiguousarray(v, dtype=np.int32)
    return lib().orc_collapse_index_vector(len(size), size.ctypes.data_as(C.POINTER(C.c_uint)), v.ctypes.data_as(C.POINTER(C.c_int)))


def brute_force_counts(pos, bmin, bmax, periodic, r):
    pos = _f64(pos)
    n, D = pos.shape
    out = np.zeros(n, dtype=np.uint32)
    bmin, bmax = _f64(bmin), _f64(bmax)
    lib().orc_brute_force_counts(D, _dp(pos), n, _dp(bmin), _dp(bmax), int(bool(periodic)), float(r) * float(r), out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


class Oracle:
    """Mirror of Particles<...,CellListOrdered> + create_sparse_operator for one
    column particle set (positions only + caller-held variable columns)."""

    def __init__(self, D):
        self.D = D
        self.h = lib().orc_create(D)
        if not self.h:
            raise ValueError("bad dimension")
        self.pos = None  # sorted positions kept alive here
        self.order = None

    def __del__(self):
        if getattr(self, "h", None) and callable(lib):  # (module globals are gone at interpreter shutdown)
            lib().orc_destroy(self.h)
            self.h = None

    def _dom(self, bmin, bmax, periodic):
        bmin = _f64(np.broadcast_to(bmin, (self.D,)))
        bmax = _f64(np.broadcast_to(bmax, (self.D,)))
        per = np.ascontiguousarray(np.broadcast_to(periodic, (self.D,)), dtype=np.uint8)
        return bmin, bmax, per

    def set_domain(self, bmin, bmax, periodic, n_leaf=10.0):
        bmin, bmax, per = self._dom(bmin, bmax, periodic)
        lib().orc_set_domain(self.h, _dp(bmin), _dp(bmax), per.ctypes.data_as(C.POINTER(C.c_uint8)), float(n_leaf))

    def force_grid(self, bmin, bmax, periodic, size):
        bmin, bmax, per = self._dom(bmin, bmax, periodic)
        size = np.ascontiguousarray(np.broadcast_to(size, (self.D,)), dtype=np.uint32)
        lib().orc_force_grid(self.h, _dp(bmin), _dp(bmax), per.ctypes.data_as(C.POINTER(C.c_uint8)), size.ctypes.data_as(C.POINTER(C.c_uint)))

    def grid(self):
        size = np.zeros(self.D, dtype=np.uint32)
        side = np.zeros(self.D, dtype=np.float64)
        lib().orc_get_grid(self.h, size.ctypes.data_as(C.POINTER(C.c_uint)), _dp(side))
        return size, side

    def point_to_bucket_index(self, r):
        r = _f64(r)
        v = np.zeros(self.D, dtype=np.int32)
        idx = lib().orc_point_to_bucket_index(self.h, _dp(r), v.ctypes.data_as(C.POINTER(C.c_int)))
        return idx, v

    def update_positions(self, pos, alive=None, sort_mode=SORT_STABLE):
        """pos (n,D) float64 is wrapped in place, alive (n,) uint8 updated in
        place.  Returns dict(order, keys, bucket_begin, bucket_end, n_alive)."""
        assert pos.dtype == np.float64 and pos.flags.c_contiguous and pos.shape[1] == self.D
        n = pos.shape[0]
        if alive is None:
            alive = np.ones(n, dtype=np.uint8)
        assert alive.dtype == np.uint8 and alive.flags.c_contiguous
        na = lib().orc_update_positions(self.h, _dp(pos), alive.ctypes.data_as(C.POINTER(C.c_uint8)), n, sort_mode)
        nb = lib().orc_num_buckets(self.h)
        order = np.zeros(na, dtype=np.int32)
        keys = np.zeros(na, dtype=np.uint32)
        bb = np.zeros(nb, dtype=np.uint32)
        be = np.zeros(nb, dtype=np.uint32)
        ip, up = C.POINTER(C.c_int), C.POINTER(C.c_uint)
        lib().orc_get_build(self.h, order.ctypes.data_as(ip), keys.ctypes.data_as(up), bb.ctypes.data_as(up), be.ctypes.data_as(up))
        return dict(order=order, keys=keys, bucket_begin=bb, bucket_end=be, n_alive=int(na), alive=alive)

    @staticmethod
    def gather(order, col):
        col = np.ascontiguousarray(col)
        out = np.empty((len(order),) + col.shape[1:], dtype=col.dtype)
        eb = col.dtype.itemsize * int(np.prod(col.shape[1:], dtype=np.int64))
        order = np.ascontiguousarray(order, dtype=np.int32)
        lib().orc_gather(order.ctypes.data_as(C.POINTER(C.c_int)), len(order), col.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), eb)
        return out

    def init_neighbour_search(self, pos, bmin, bmax, periodic, n_leaf=10.0, alive=None, sort_mode=SORT_STABLE):
        """src/Particles.h:445-455: set_domain, update_positions, reorder.
        Returns the build dict plus 'pos' (reordered positions)."""
        pos = np.array(pos, dtype=np.float64, order="C", copy=True)
        self.set_domain(bmin, bmax, periodic, n_leaf)
        out = self.update_positions(pos, alive, sort_mode)
        self.pos = self.gather(out["order"], pos)
        self.order = out["order"]
        lib().orc_update_iterators(self.h, _dp(self.pos), self.pos.shape[0])
        out["pos"] = self.pos
        return out

    def update_iterators(self, pos_sorted):
        self.pos = _f64(pos_sorted)
        lib().orc_update_iterators(self.h, _dp(self.pos), self.pos.shape[0])

    def buckets_near_point(self, point, r, max_out=4096):
        point = _f64(point)
        out = np.zeros((max_out, self.D), dtype=np.int32)
        c = lib().orc_buckets_near_point(self.h, _dp(point), float(r), out.ctypes.data_as(C.POINTER(C.c_int)), max_out)
        return int(c), out[: min(c, max_out)]

    def search_point(self, point, r, max_out=65536):
        point = _f64(point)
        j = np.zeros(max_out, dtype=np.int32)
        im = np.zeros(max_out, dtype=np.int32)
        dx = np.zeros((max_out, self.D), dtype=np.float64)
        ip = C.POINTER(C.c_int)
        c = lib().orc_search_point(self.h, _dp(point), float(r), j.ctypes.data_as(ip), im.ctypes.data_as(ip), _dp(dx), max_out)
        m = min(c, max_out)
        return int(c), j[:m], im[:m], dx[:m]

    def pair_stats(self, row_pos, radius, radius_per_row=None):
        row_pos = _f64(row_pos)
        n = row_pos.shape[0]
        cnt = np.zeros(n, dtype=np.uint32)
        hs = np.zeros(n, dtype=np.uint64)
        rpr = _f64(radius_per_row) if radius_per_row is not None else None
        lib().orc_pair_stats(self.h, _dp(row_pos), n, float(radius), _dp(rpr), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), hs.ctypes.data_as(C.POINTER(C.c_uint64)))
        return cnt, hs

    def pair_stats_norm(self, row_pos, radius, lnorm, scale=None, linear=None):
        """distance_search<lnorm> (src/Search.h:794-831): count and pair-set hash per row;
        scale: ScaleTransform factors (src/Transform.h:140-160); linear: D x D matrix of a
        LinearTransform (:61-137); neither = IdentityTransform"""
        row_pos = _f64(row_pos)
        n = row_pos.shape[0]
        cnt = np.zeros(n, dtype=np.uint32)
        hs = np.zeros(n, dtype=np.uint64)
        if linear is not None:
            mat = _f64(np.asarray(linear, dtype=np.float64).reshape(self.D, self.D))
            rc = lib().orc_pair_stats_norm_linear(self.h, _dp(row_pos), n, float(radius), int(lnorm), _dp(mat), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), hs.ctypes.data_as(C.POINTER(C.c_uint64)))
            if rc:
                raise ValueError("unsupported norm")
            return cnt, hs
        sc = _f64(np.broadcast_to(scale, (self.D,))) if scale is not None else None
        rc = lib().orc_pair_stats_norm_scaled(self.h, _dp(row_pos), n, float(radius), int(lnorm), _dp(sc), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), hs.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc:
            raise ValueError("unsupported norm")
        return cnt, hs

    def sparse_matvec(self, row_pos, kernel_id, params, radius, b, BR=1, BC=1, row_vars=(), col_vars=(), radius_per_row=None, y=None, nthreads=0):
        """y += K b (src/Kernels.h:720-751).  Returns (y, n_pairs)."""
        row_pos = _f64(row_pos)
        n_rows = row_pos.shape[0]
        params = _f64(params if len(params) else [0.0])
        b = _f64(b)
        if y is None:
            y = np.zeros(n_rows * BR, dtype=np.float64)
        rv = [_f64(v) for v in row_vars]
        cv = [_f64(v) for v in col_vars]
        dpp = C.POINTER(C.c_double)
        RV = (dpp * max(1, len(rv)))(*[_dp(v) for v in rv])
        CV = (dpp * max(1, len(cv)))(*[_dp(v) for v in cv])
        rpr = _f64(radius_per_row) if radius_per_row is not None else None
        npairs = lib().orc_sparse_matvec(self.h, _dp(row_pos), n_rows, int(kernel_id), _dp(params), RV, CV, float(radius), _dp(rpr), BR, BC, _dp(b), _dp(y), int(nthreads))
        return y, int(npairs)


def _assemble(self, row_pos, kernel_id, params, radius, BR=1, BC=1, row_vars=(), col_vars=(), radius_per_row=None):
    """K.assemble(triplets) (src/Kernels.h:653-685) as CSR: (row_ptr, col_idx, values[nnz,BR,BC])"""
    row_pos = _f64(row_pos)
    n_rows = row_pos.shape[0]
    params = _f64(params if len(params) else [0.0])
    rv = [_f64(v) for v in row_vars]
    cv = [_f64(v) for v in col_vars]
    dpp = C.POINTER(C.c_double)
    RV = (dpp * max(1, len(rv)))(*[_dp(v) for v in rv])
    CV = (dpp * max(1, len(cv)))(*[_dp(v) for v in cv])
    rpr = _f64(radius_per_row) if radius_per_row is not None else None
    row_ptr = np.zeros(n_rows + 1, dtype=np.uint32)
    u32p, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    args = (self.h, _dp(row_pos), n_rows, int(kernel_id), _dp(params), RV, CV, float(radius), _dp(rpr), BR, BC, row_ptr.ctypes.data_as(u32p))
    nnz = lib().orc_sparse_assemble(*args, None, None)
    col = np.zeros(max(nnz, 1), dtype=np.int32)
    vals = np.zeros((max(nnz, 1), BR, BC), dtype=np.float64)
    lib().orc_sparse_assemble(*args, col.ctypes.data_as(i32p), _dp(vals))
    return row_ptr, col[:nnz], vals[:nnz]


Oracle.assemble = _assemble


def _coeff(self, row_pos, col_pos, ii, jj, kernel_id, params, radius, BR=1, BC=1, row_vars=(), col_vars=(), radius_per_row=None):
    """K.coeff(i, j) (src/Kernels.h:102-112 over detail::sparse_kernel, src/detail/Kernels.h:336-367) for arrays ii, jj"""
    row_pos, col_pos = _f64(row_pos), _f64(col_pos)
    ii = np.ascontiguousarray(ii, dtype=np.uint64)
    jj = np.ascontiguousarray(jj, dtype=np.uint64)
    params = _f64(params if len(params) else [0.0])
    rv = [_f64(v) for v in row_vars]
    cv = [_f64(v) for v in col_vars]
    dpp = C.POINTER(C.c_double)
    RV = (dpp * max(1, len(rv)))(*[_dp(v) for v in rv])
    CV = (dpp * max(1, len(cv)))(*[_dp(v) for v in cv])
    rpr = _f64(radius_per_row) if radius_per_row is not None else None
    out = np.zeros(len(ii), dtype=np.float64)
    u64p = C.POINTER(C.c_uint64)
    lib().orc_sparse_coeff(self.h, _dp(row_pos), _dp(col_pos), ii.ctypes.data_as(u64p), jj.ctypes.data_as(u64p), len(ii), int(kernel_id), _dp(params),
                           RV, CV, float(radius), _dp(rpr), BR, BC, _dp(out))
    return out


Oracle.coeff = _coeff


def _accumulate_within_distance(self, row_pos, kernel_id, params, radius, BR=1, row_vars=(), col_vars=(), init=0.0):
    """AccumulateWithinDistance<std::plus> per row (src/detail/Contexts.h:247-289 sparse_sum_impl):
    sum = init; for b in distance_search<2>(cols, r_a, radius): sum = sum + expr(dx, a, b), in the
    iterator's order.  Restated through orc_sparse_matvec with rhs == 1 (x * 1.0 is exact) and
    lhs preset to init: the same sequential additions in the same order."""
    n = np.asarray(row_pos).shape[0]
    y0 = np.empty((n, BR), dtype=np.float64)
    y0[:] = np.asarray(init, dtype=np.float64)
    ncols = self.pos.shape[0]
    y, _ = self.sparse_matvec(row_pos, kernel_id, params, radius, np.ones(ncols), BR=BR, BC=1, row_vars=row_vars, col_vars=col_vars, y=y0.reshape(-1).copy())
    return y if BR == 1 else y.reshape(n, BR)


Oracle.accumulate_within_distance = _accumulate_within_distance


def _bucket_pairs(self):
    """get_neighbouring_buckets(query) (src/Search.h:498-764): (bucket_i, bucket_j, quadrant[n, D]) in iterator order"""
    n = lib().orc_bucket_pairs(self.h, None, None, None, 0)
    bi = np.zeros(max(n, 1), dtype=np.uint32)
    bj = np.zeros(max(n, 1), dtype=np.uint32)
    qd = np.zeros((max(n, 1), self.D), dtype=np.int8)
    u32p = C.POINTER(C.c_uint32)
    lib().orc_bucket_pairs(self.h, bi.ctypes.data_as(u32p), bj.ctypes.data_as(u32p), qd.ctypes.data_as(C.POINTER(C.c_int8)), n)
    return bi[:n], bj[:n], qd[:n]


def _fast_bucket_search_counts(self, radius):
    """per-particle neighbour counts through the bucket-pair traversal (tests/neighbours.h:892-951)"""
    out = np.zeros(self.pos.shape[0], dtype=np.uint32)
    lib().orc_fast_bucket_search_counts(self.h, float(radius), out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


Oracle.bucket_pairs = _bucket_pairs
Oracle.fast_bucket_search_counts = _fast_bucket_search_counts


def id_map_build(ids):
    """the id map of src/NeighbourSearchBase.h:440-486: (key, value) sorted by id"""
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    key = np.zeros(len(ids), dtype=np.uint64)
    value = np.zeros(len(ids), dtype=np.uint64)
    u64p = C.POINTER(C.c_uint64)
    lib().orc_id_map_build(ids.ctypes.data_as(u64p), len(ids), key.ctypes.data_as(u64p), value.ctypes.data_as(u64p))
    return key, value


def id_find(key, value, query):
    """CellListOrderedQuery::find (src/CellListOrdered.h:379-388): index or n"""
    key = np.ascontiguousarray(key, dtype=np.uint64)
    value = np.ascontiguousarray(value, dtype=np.uint64)
    query = np.ascontiguousarray(query, dtype=np.uint64)
    out = np.zeros(len(query), dtype=np.uint64)
    u64p = C.POINTER(C.c_uint64)
    lib().orc_id_find(key.ctypes.data_as(u64p), value.ctypes.data_as(u64p), len(key), query.ctypes.data_as(u64p), len(query), out.ctypes.data_as(u64p))
    return out


def max_threads():
    return lib().orc_max_threads()


def set_num_threads(n):
    """OpenMP team size of the oracle from here on (overrides OMP_NUM_THREADS)"""
    lib().orc_set_num_threads(int(n))


def host_cores():
    """cores this process may run on (the affinity mask, not OMP_NUM_THREADS)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
