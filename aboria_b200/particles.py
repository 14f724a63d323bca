"""Host-side mirror of the reference interface for the hot path, on top of the
C-ABI (include/abr.h).  Names and argument meaning follow the reference:

  Particles.init_neighbour_search(low, high, periodic, n_particles_in_leaf=10)
      src/Particles.h:445-455
  Particles.update_positions()            src/Particles.h:526-531 (+ reorder :694-724)
  Particles.get_query()                   src/Particles.h (CellListOrderedQuery POD)
  create_sparse_operator(rows, cols, radius, kernel)   src/Operators.h:478-516
  K * b / K.evaluate(y, b)                src/Operators.h:153, src/Kernels.h:720-751

torch is used for device memory and streams only; all compute is in libabr.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import AbrError, KernelDesc, check
from .kernels import Kernel


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Query:
    """Device view handed to kernels (CellListOrderedQuery, src/CellListOrdered.h:285-599)."""

    def __init__(self, particles):
        self.particles = particles

    @property
    def bucket_begin(self):
        return self.particles._bucket_view()[1]

    @property
    def bucket_end(self):
        return self.particles._bucket_view()[2]

    @property
    def bucket_indices(self):
        return self.particles._bucket_view()[0]

    def neighbouring_buckets(self):
        """get_neighbouring_buckets(query) (src/Search.h:857-860): the bucket pairs of the
        fast cell-list search in iterator order, as (bucket_i, bucket_j int32 tensors,
        quadrant int8 tensor [n, D]); position offset of a pair = quadrant * (high - low)."""
        p = self.particles
        p._sync_stream()
        n = C.c_uint64()
        check(p._h, p._lib.abr_bucket_pairs(p._h, None, None, None, 0, C.byref(n)))
        m = max(int(n.value), 1)
        bi = torch.empty(m, dtype=torch.int32, device=p.device)
        bj = torch.empty(m, dtype=torch.int32, device=p.device)
        qd = torch.empty((m, p.D), dtype=torch.int8, device=p.device)
        check(p._h, p._lib.abr_bucket_pairs(p._h, _ptr(bi), _ptr(bj), _ptr(qd), int(n.value), C.byref(n)))
        return bi[: n.value], bj[: n.value], qd[: n.value]

    def fast_bucket_search_counts(self, radius):
        """neighbour counts through the bucket-pair traversal (tests/neighbours.h:892-951)"""
        p = self.particles
        p._sync_stream()
        cnt = torch.zeros(max(p.size(), 1), dtype=torch.int32, device=p.device)
        check(p._h, p._lib.abr_fast_bucket_search_counts(p._h, float(radius), _ptr(cnt)))
        return cnt[: p.size()]

    def find(self, ids):
        """find(id) (src/CellListOrdered.h:379-388) for a batch of ids: the position of
        the particle with that id, or size() (the reference's end pointer) when absent."""
        return self.particles._find_ids(ids)


class Particles:
    """Struct-of-columns particle container on one GPU (Particles<VAR,D,...,
    CellListOrdered>, src/Particles.h:108-112).  Columns: position (n x D f64),
    id (int64, the reference's size_t), alive (uint8), plus user variables."""

    def __init__(self, D=3, n=0, variables=None, device=None):
        if D < 1 or D > _lib.MAX_D:
            raise ValueError("D must be 1, 2 or 3")
        if not torch.cuda.is_available():
            raise AbrError("aboria_b200 needs a CUDA device; there is no CPU fallback")
        self.D = D
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.abr_create(C.byref(self._h), self.device.index or 0, C.c_void_p(stream))
        if rc != 0:
            raise AbrError(f"abr_create failed ({rc}): {self._lib.abr_last_error_string(None).decode()}")
        self.columns = {}
        self._other = {}
        self.columns["position"] = torch.zeros((n, D), dtype=torch.float64, device=self.device)
        self.columns["id"] = torch.arange(n, dtype=torch.int64, device=self.device)
        self.columns["alive"] = torch.ones(n, dtype=torch.uint8, device=self.device)
        for name, spec in (variables or {}).items():
            dtype, shape = spec if isinstance(spec, tuple) else (spec, ())
            self.columns[name] = torch.zeros((n,) + tuple(shape), dtype=dtype, device=self.device)
        self.searchable = False
        self._order = None
        self.n_buckets = 0
        self._id_map = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.abr_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- container ---------------------------------------------------------
    def size(self):
        return self.columns["position"].shape[0]

    __len__ = size

    def get(self, name):
        """get<variable>(particles)"""
        return self.columns[name]

    def set(self, name, value):
        t = torch.as_tensor(value, device=self.device)
        cur = self.columns.get(name)
        if cur is not None:
            t = t.to(cur.dtype)
        if name != "position" and name in self.columns and t.shape[0] != self.size():
            raise ValueError("column length mismatch")
        self.columns[name] = t.contiguous()
        if name == "position":
            # the handle's query points at the tensor the last build reordered into: replacing the column
            # leaves it without a search structure until the next update_positions (the reference would
            # search a stale list; a raw pointer into freed device memory must not be searched at all)
            self.searchable = False

    def resize_from_positions(self, pos):
        """Replace the whole set by n new particles at `pos` (host or device)."""
        pos = torch.as_tensor(pos, dtype=torch.float64)
        n = pos.shape[0]
        cur = self.columns.get("position")
        if cur is not None and cur.shape == pos.shape and cur.is_contiguous():
            # same size as before: refill the existing buffers (no allocation on the hot path)
            cur.copy_(pos, non_blocking=True)
            iota = getattr(self, "_iota", None)
            if iota is None or iota.shape[0] != n:
                iota = self._iota = torch.arange(n, dtype=torch.int64, device=self.device)
            self.columns["id"].copy_(iota)  # ids 0..n-1 (a device copy is 3x faster than regenerating the sequence)
            self.columns["alive"].fill_(1)
            for name in self.columns:
                if name not in ("position", "id", "alive"):
                    self.columns[name].zero_()
            return
        self.searchable = False
        self.columns["position"] = pos.to(self.device, non_blocking=True).contiguous().clone() if pos.device == self.device else pos.to(self.device, non_blocking=True).contiguous()
        self.columns["id"] = torch.arange(n, dtype=torch.int64, device=self.device)
        self.columns["alive"] = torch.ones(n, dtype=torch.uint8, device=self.device)
        for name in list(self.columns):
            if name not in ("position", "id", "alive"):
                old = self.columns[name]
                self.columns[name] = torch.zeros((n,) + tuple(old.shape[1:]), dtype=old.dtype, device=self.device)

    # -- neighbour search ----------------------------------------------------
    def _sync_stream(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(self._h, self._lib.abr_set_stream(self._h, C.c_void_p(stream)))

    def init_neighbour_search(self, low, high, periodic, n_particles_in_leaf=10.0, assume_all_alive=False):
        D = self.D
        low = np.ascontiguousarray(np.broadcast_to(np.asarray(low, dtype=np.float64), (D,)))
        high = np.ascontiguousarray(np.broadcast_to(np.asarray(high, dtype=np.float64), (D,)))
        per = np.ascontiguousarray(np.broadcast_to(np.asarray(periodic), (D,)).astype(np.uint8))
        check(self._h, self._lib.abr_domain_set(self._h, D, low.ctypes.data, high.ctypes.data, per.ctypes.data, float(n_particles_in_leaf)))
        self.low, self.high, self.periodic = low, high, per
        self.update_positions(assume_all_alive=assume_all_alive)
        self.searchable = True

    def force_grid(self, low, high, periodic, size):
        D = self.D
        low = np.ascontiguousarray(np.broadcast_to(np.asarray(low, dtype=np.float64), (D,)))
        high = np.ascontiguousarray(np.broadcast_to(np.asarray(high, dtype=np.float64), (D,)))
        per = np.ascontiguousarray(np.broadcast_to(np.asarray(periodic), (D,)).astype(np.uint8))
        size = np.ascontiguousarray(np.broadcast_to(np.asarray(size), (D,)).astype(np.uint32))
        check(self._h, self._lib.abr_domain_force_grid(self._h, D, low.ctypes.data, high.ctypes.data, per.ctypes.data, size.ctypes.data))
        self.low, self.high, self.periodic = low, high, per
        self.update_positions()
        self.searchable = True

    def grid(self):
        size = np.zeros(self.D, dtype=np.uint32)
        side = np.zeros(self.D, dtype=np.float64)
        nb = C.c_uint64()
        check(self._h, self._lib.abr_domain_get(self._h, size.ctypes.data, side.ctypes.data, C.byref(nb)))
        return size, side, nb.value

    def check_async(self):
        """verify the last asynchronous update (raises if a particle died)"""
        check(self._h, self._lib.abr_check_async(self._h))

    def update_positions(self, assume_all_alive=False):
        """wrap / kill, build the ordered cell list, reorder every column.
        assume_all_alive=True: asynchronous form (no host round trip); the caller
        asserts no particle leaves the domain and calls check_async() later."""
        self._sync_stream()
        n = self.size()
        pos = self.columns["position"]
        alive = self.columns["alive"]
        order = getattr(self, "_order_buf", None)
        if order is None or order.shape[0] != max(n, 1):
            order = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
            self._order_buf = order
        n_alive = C.c_size_t(0)
        # reorder: every column is gathered into the other buffer (room for n), then swapped;
        # the other buffer is kept between calls like the reference's other_data
        names = list(self.columns)
        src = [self.columns[k] for k in names]
        dst = []
        for k, t in zip(names, src):
            o = self._other.get(k)
            if o is not None and o.shape == t.shape and o.dtype == t.dtype and o.is_contiguous() and o.data_ptr() != t.data_ptr():
                dst.append(o)
            else:
                dst.append(torch.empty_like(t))
        nc = len(names)
        SP = (C.c_void_p * nc)(*[t.data_ptr() for t in src])
        DP = (C.c_void_p * nc)(*[t.data_ptr() for t in dst])
        EB = (C.c_size_t * nc)(*[t.element_size() * int(np.prod(t.shape[1:], dtype=np.int64)) for t in src])
        check(self._h, self._lib.abr_update_positions(self._h, _ptr(pos), _ptr(alive), n, nc, SP, DP, EB, _ptr(order),
                                                      None if assume_all_alive else C.byref(n_alive)))
        na = n if assume_all_alive else n_alive.value
        self._order = order[:na]
        self._other = dict(zip(names, src))
        self.columns = {k: t[:na] for k, t in zip(names, dst)}
        self._bound_position = self.columns["position"]  # keeps the tensor behind the handle's query alive until the next build
        self.n_buckets = self.grid()[2]
        if self._id_map:
            self._update_id_map()
        return na

    # -- find by id ------------------------------------------------------------
    def init_id_search(self):
        """Particles::init_id_search (src/Particles.h) -> init_id_map + update
        (src/NeighbourSearchBase.h:294-298, :440-486); kept up to date by every
        later update_positions, as in the reference."""
        self._id_map = True
        self._update_id_map()

    def _update_id_map(self):
        self._sync_stream()
        ids = self.columns["id"]
        check(self._h, self._lib.abr_id_map_build(self._h, _ptr(ids), ids.shape[0]))

    def id_map(self):
        """(m_id_map_key, m_id_map_value) as int64 device tensors"""
        k, v = C.c_void_p(), C.c_void_p()
        n = C.c_size_t()
        check(self._h, self._lib.abr_id_map_get(self._h, C.byref(k), C.byref(v), C.byref(n)))
        check(self._h, self._lib.abr_synchronize(self._h))
        if n.value == 0:
            e = torch.empty(0, dtype=torch.int64, device=self.device)
            return e, e.clone()
        kt = torch.as_tensor(_DevArray(k.value, n.value, "<i8"), device=self.device).clone()
        vt = torch.as_tensor(_DevArray(v.value, n.value, "<i8"), device=self.device).clone()
        return kt, vt

    def _find_ids(self, ids):
        if not self._id_map:
            raise AbrError("init_id_search not called on this particle set")
        self._sync_stream()
        q = torch.as_tensor(ids, dtype=torch.int64, device=self.device).contiguous().reshape(-1)
        out = torch.empty(q.shape[0], dtype=torch.int64, device=self.device)
        check(self._h, self._lib.abr_id_find(self._h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def get_alive_indicies(self):
        """m_alive_indices after the sort: new[k] = old[order[k]]"""
        return self._order

    def get_query(self):
        return Query(self)

    def _bucket_view(self, clone=True, sync=True):
        """(bucket_indices, bucket_begin, bucket_end) as int32 device tensors;
        clone=False returns views borrowed from the handle (valid until the next build);
        sync=False skips the stream synchronisation (stream-ordered consumers only)"""
        ki, bb, be = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nb = C.c_uint64()
        check(self._h, self._lib.abr_celllist_get(self._h, C.byref(ki), C.byref(bb), C.byref(be), C.byref(nb)))
        n = self.size()

        if sync:
            check(self._h, self._lib.abr_synchronize(self._h))

        def view(ptr, count):
            if count == 0 or not ptr.value:
                return torch.empty(0, dtype=torch.int32, device=self.device)
            t = torch.as_tensor(_DevArray(ptr.value, count, "<i4"), device=self.device)
            return t.clone() if clone else t

        return view(ki, n), view(bb, nb.value), view(be, nb.value)

    def last_counters(self):
        c = (C.c_uint64 * 4)()
        check(self._h, self._lib.abr_last_counters(self._h, C.byref(c)))
        return dict(walk_rows=int(c[0]), aliased=int(c[1]), launches=int(c[2]), total_launches=int(c[3]))

    def set_option(self, name, value):
        """tuning knobs of the library (no effect on results), see abr_set_option"""
        check(self._h, self._lib.abr_set_option(self._h, name.encode(), float(value)))

    def distance_search_stats(self, radius, lnorm, queries=None, scale=None, linear=None):
        """distance_search<lnorm> / chebyshev_search (-1) / manhatten_search (1) /
        euclidean_search (2) from `queries` (default: the particles themselves):
        per query the neighbour count and pair-set hash."""
        self._sync_stream()
        qp = self.columns["position"] if queries is None else torch.as_tensor(queries, dtype=torch.float64, device=self.device).contiguous()
        n = qp.shape[0]
        cnt = torch.zeros(n, dtype=torch.int32, device=self.device)
        hs = torch.zeros(n, dtype=torch.int64, device=self.device)
        if linear is not None:  # create_linear_transform<D>(functor) with a linear functor given as its matrix (src/Transform.h:61-137)
            mat = np.ascontiguousarray(np.asarray(linear, dtype=np.float64).reshape(self.D, self.D))
            check(self._h, self._lib.abr_distance_search_stats_linear(self._h, _ptr(qp), n, float(radius), None, int(lnorm), mat.ctypes.data, _ptr(cnt), _ptr(hs)))
            return cnt, hs
        if scale is not None:  # create_scale_transform(scale) (src/Transform.h:140-172)
            sc = np.ascontiguousarray(np.broadcast_to(np.asarray(scale, dtype=np.float64), (self.D,)))
            check(self._h, self._lib.abr_distance_search_stats_scaled(self._h, _ptr(qp), n, float(radius), None, int(lnorm), sc.ctypes.data, _ptr(cnt), _ptr(hs)))
            return cnt, hs
        check(self._h, self._lib.abr_distance_search_stats(self._h, _ptr(qp), n, float(radius), None, int(lnorm), _ptr(cnt), _ptr(hs)))
        return cnt, hs

    def probe_fp64_peak(self):
        """measured DFMA throughput of this device in TFLOP/s"""
        self._sync_stream()
        v = C.c_double()
        check(self._h, self._lib.abr_probe_fp64_peak(self._h, C.byref(v)))
        return v.value

    def pair_stats(self, radius, rows=None, path=-1, radius_per_row=None):
        """per-row neighbour count and pair-set hash of euclidean_search"""
        self._sync_stream()
        rows_are_cols = rows is None
        rp = self.columns["position"] if rows is None else torch.as_tensor(rows, dtype=torch.float64, device=self.device).contiguous()
        n = rp.shape[0]
        cnt = torch.zeros(n, dtype=torch.int32, device=self.device)
        hs = torch.zeros(n, dtype=torch.int64, device=self.device)
        rpr = None
        if radius_per_row is not None:
            rpr = torch.as_tensor(radius_per_row, dtype=torch.float64, device=self.device).contiguous()
        check(self._h, self._lib.abr_pair_stats(self._h, _ptr(rp), n, int(rows_are_cols), float(radius), _ptr(rpr), int(path), _ptr(cnt), _ptr(hs)))
        return cnt, hs


class _DevArray:
    """borrowed device memory exposed through __cuda_array_interface__"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class SparseOperator:
    """MatrixReplacement<1,1,tuple<KernelSparse[Const]>> (src/Operators.h:75-291)."""

    def __init__(self, rows, cols, radius, kernel, radius_per_row=None):
        if not isinstance(kernel, Kernel):
            raise TypeError("kernel must be an aboria_b200.kernels.Kernel")
        self.row_particles = rows
        self.col_particles = cols
        self.radius = float(radius) if radius is not None else 0.0
        self.radius_per_row = radius_per_row
        self.kernel = kernel

    def rows(self):
        return self.row_particles.size() * self.kernel.block_rows

    def cols(self):
        return self.col_particles.size() * self.kernel.block_cols

    def _desc(self):
        k = self.kernel
        d = KernelDesc()
        d.kernel_id, d.block_rows, d.block_cols = k.kernel_id, k.block_rows, k.block_cols
        for i, p in enumerate(k.params):
            d.params[i] = p
        keep = []
        for i, name in enumerate(k.row_vars):
            t = self.row_particles.get(name)
            keep.append(t)
            d.row_vars[i] = t.data_ptr()
        for i, name in enumerate(k.col_vars):
            t = self.col_particles.get(name)
            keep.append(t)
            d.col_vars[i] = t.data_ptr()
        return d, keep

    def evaluate(self, y, b, count_pairs=False):
        """y += K b, device tensors (KernelSparse::evaluate, src/Kernels.h:720-751)."""
        cols, rows = self.col_particles, self.row_particles
        if not cols.searchable:
            raise AbrError("column particle set has no neighbour search (call init_neighbour_search)")
        if b.shape[0] != self.cols() or y.shape[0] != self.rows():
            raise ValueError("vector has incompatible size")
        if b.dtype != torch.float64 or y.dtype != torch.float64 or not b.is_contiguous() or not y.is_contiguous():
            raise ValueError("b and y must be contiguous float64 device tensors")
        cols._sync_stream()
        d, keep = self._desc()
        rpr = None
        if self.radius_per_row is not None:
            rpr = torch.as_tensor(self.radius_per_row, dtype=torch.float64, device=cols.device).contiguous()
        npairs = C.c_uint64(0)
        rp = rows.get("position")
        rc = cols._lib.abr_sparse_matvec(cols._h, _ptr(rp), rows.size(), int(rows is cols), C.byref(d), self.radius, _ptr(rpr), _ptr(b), _ptr(y),
                                         C.byref(npairs) if count_pairs else None)
        check(cols._h, rc)
        del keep
        return npairs.value if count_pairs else None

    def coeff(self, i, j):
        """K.coeff(i, j) (src/Operators.h:149-151 -> src/Kernels.h:102-112 over
        detail::sparse_kernel): matrix entries for index arrays i, j.  Note the strict
        `|dx|^2 < r^2` and the minimum-image dx of this path (src/detail/Kernels.h:358-367)."""
        cols, rows = self.col_particles, self.row_particles
        cols._sync_stream()
        d, keep = self._desc()
        it = torch.as_tensor(i, dtype=torch.int64, device=cols.device).contiguous().reshape(-1)
        jt = torch.as_tensor(j, dtype=torch.int64, device=cols.device).contiguous().reshape(-1)
        if it.shape != jt.shape:
            raise ValueError("i and j must have the same length")
        if it.numel() and (int(it.min()) < 0 or int(it.max()) >= self.rows() or int(jt.min()) < 0 or int(jt.max()) >= self.cols()):
            raise ValueError("coeff: index out of range")  # the reference ASSERTs (src/Kernels.h:103-104)
        rpr = None
        if self.radius_per_row is not None:
            rpr = torch.as_tensor(self.radius_per_row, dtype=torch.float64, device=cols.device).contiguous()
        out = torch.empty(it.shape[0], dtype=torch.float64, device=cols.device)
        rp = rows.get("position")
        cp = cols.get("position")
        check(cols._h, cols._lib.abr_query_set_particles(cols._h, _ptr(cp), cols.size()))
        check(cols._h, cols._lib.abr_sparse_coeff(cols._h, _ptr(rp), rows.size(), C.byref(d), self.radius, _ptr(rpr), _ptr(it), _ptr(jt), it.shape[0], _ptr(out)))
        del keep
        return out

    def assemble(self, values=True):
        """K.assemble(triplets) (src/Kernels.h:653-685) as CSR on the device:
        returns (row_ptr int32[n_rows+1], col_idx int32[nnz], values float64[nnz, BR, BC] | None)."""
        cols, rows = self.col_particles, self.row_particles
        if not cols.searchable:
            raise AbrError("column particle set has no neighbour search (call init_neighbour_search)")
        cols._sync_stream()
        d, keep = self._desc()
        rpr = None
        if self.radius_per_row is not None:
            rpr = torch.as_tensor(self.radius_per_row, dtype=torch.float64, device=cols.device).contiguous()
        n_rows = rows.size()
        rp = rows.get("position")
        row_ptr = torch.zeros(n_rows + 1, dtype=torch.int32, device=cols.device)
        nnz = C.c_uint64(0)
        args = (cols._h, _ptr(rp), n_rows, int(rows is cols), C.byref(d), self.radius, _ptr(rpr), _ptr(row_ptr))
        check(cols._h, cols._lib.abr_sparse_assemble(*args, None, None, 0, C.byref(nnz)))
        n = nnz.value
        col_idx = torch.empty(max(n, 1), dtype=torch.int32, device=cols.device)
        k = self.kernel
        vals = torch.empty((max(n, 1), k.block_rows, k.block_cols), dtype=torch.float64, device=cols.device) if values else None
        check(cols._h, cols._lib.abr_sparse_assemble(*args, _ptr(col_idx), _ptr(vals), n, C.byref(nnz)))
        del keep
        return row_ptr, col_idx[:n], (vals[:n] if values else None)

    def matvec(self, b, out=None):
        """y = K * b  (Eigen zeroes y first, src/detail/Operators.h:219-232).
        out: optional preallocated result vector (zeroed here)."""
        if out is None:
            y = torch.zeros(self.rows(), dtype=torch.float64, device=self.col_particles.device)
        else:
            y = out
            y.zero_()
        self.evaluate(y, b)
        return y

    __mul__ = matvec
    __matmul__ = matvec

    def matvec_host(self, b_host, out_host=None):
        """Host-buffer entry: copies b to the device, applies K, copies y back.
        b_host / out_host are (ideally pinned) CPU tensors or numpy arrays."""
        dev = self.col_particles.device
        bt = torch.as_tensor(b_host, dtype=torch.float64)
        b = bt.to(dev, non_blocking=True)
        y = self.matvec(b)
        if out_host is None:
            return y.cpu()
        out_host.copy_(y, non_blocking=False)
        return out_host


class ZeroOperator:
    """create_zero_operator(rows, cols) (src/Operators.h:531-537, KernelZero): a block of zeros"""

    def __init__(self, rows, cols, block_rows=1, block_cols=1):
        self.row_particles, self.col_particles = rows, cols
        self.block_rows, self.block_cols = block_rows, block_cols

    def rows(self):
        return self.row_particles.size() * self.block_rows

    def cols(self):
        return self.col_particles.size() * self.block_cols

    def evaluate(self, y, b):
        return None

    def coeff(self, i, j):
        it = torch.as_tensor(i).reshape(-1)
        return torch.zeros(it.shape[0], dtype=torch.float64, device=self.col_particles.device)


def create_zero_operator(row_particles, col_particles):
    return ZeroOperator(row_particles, col_particles)


class BlockOperator:
    """create_block_operator<NI,NJ>(blocks...) (src/Operators.h:541-548): an NI x NJ
    arrangement of operators behind one MatrixReplacement.  The product follows
    src/detail/Operators.h:170-198: block (I, J) is evaluated on
    y.segment(start_row(I), rows(I)) and x.segment(start_col(J), cols(J)) and
    accumulates into y; coeff() picks the block that owns (i, j)
    (src/Operators.h:242-251).  Host-side composition only: every block runs its own
    device kernels."""

    def __init__(self, NI, NJ, blocks):
        blocks = list(blocks)
        if len(blocks) != NI * NJ:
            raise ValueError("create_block_operator: need NI*NJ blocks")
        self.NI, self.NJ, self.blocks = NI, NJ, blocks
        for I in range(NI):
            for J in range(NJ):
                blk = self.block(I, J)
                if blk.rows() != self.block(I, 0).rows() or blk.cols() != self.block(0, J).cols():
                    raise ValueError("create_block_operator: block sizes do not line up")

    def block(self, I, J):
        return self.blocks[I * self.NJ + J]

    def _row_starts(self):
        s = [0]
        for I in range(self.NI):
            s.append(s[-1] + self.block(I, 0).rows())
        return s

    def _col_starts(self):
        s = [0]
        for J in range(self.NJ):
            s.append(s[-1] + self.block(0, J).cols())
        return s

    def rows(self):
        return self._row_starts()[-1]

    def cols(self):
        return self._col_starts()[-1]

    def evaluate(self, y, b):
        rs, cs = self._row_starts(), self._col_starts()
        if b.shape[0] != cs[-1] or y.shape[0] != rs[-1]:
            raise ValueError("vector has incompatible size")
        for I in range(self.NI):
            for J in range(self.NJ):
                self.block(I, J).evaluate(y[rs[I]:rs[I + 1]], b[cs[J]:cs[J + 1]])

    def matvec(self, b, out=None):
        dev = b.device
        y = torch.zeros(self.rows(), dtype=torch.float64, device=dev) if out is None else out.zero_()
        self.evaluate(y, b)
        return y

    __mul__ = matvec
    __matmul__ = matvec

    def coeff(self, i, j):
        it = torch.as_tensor(i, dtype=torch.int64).reshape(-1)
        jt = torch.as_tensor(j, dtype=torch.int64).reshape(-1)
        rs, cs = self._row_starts(), self._col_starts()
        dev = None
        out = None
        for I in range(self.NI):
            for J in range(self.NJ):
                m = (it >= rs[I]) & (it < rs[I + 1]) & (jt >= cs[J]) & (jt < cs[J + 1])
                if not bool(m.any()):
                    continue
                v = self.block(I, J).coeff(it[m] - rs[I], jt[m] - cs[J])
                if out is None:
                    dev = v.device
                    out = torch.zeros(it.shape[0], dtype=torch.float64, device=dev)
                out[m.to(dev)] = v
        if out is None:
            raise ValueError("coeff: index out of range")
        return out


def create_block_operator(NI, NJ, *blocks):
    return BlockOperator(NI, NJ, blocks)


def accumulate_within_distance(row_particles, col_particles, radius, kernel, init=0.0):
    """AccumulateWithinDistance<std::plus<T>> evaluated for every row particle
    (src/Symbolic.h:420-444 -> sparse_sum_impl, src/detail/Contexts.h:247-289):
        s_a = init;  for b in euclidean_search(cols, r_a, radius): s_a = s_a + expr(dx, a, b)
    with `expr` given as a kernel function (the restricted expression subset: anything
    written as a device functor F(dx, a, b), scalar or BR x 1 vector valued — e.g. the
    density and pressure sums of tests/sph.h:295-353, kernels.sph_density / sph_pressure).
    It is the sparse product with b == 1 accumulated onto `init`, so it runs on the same
    cell-tiled kernel.  Returns a (n_rows,) or (n_rows, BR) device tensor."""
    if kernel.block_cols != 1:
        raise ValueError("accumulate_within_distance: the summand must be scalar or BR x 1")
    op = SparseOperator(row_particles, col_particles, radius, kernel)
    dev = col_particles.device
    ones = torch.ones(op.cols(), dtype=torch.float64, device=dev)
    y = torch.empty(op.rows(), dtype=torch.float64, device=dev)
    br = kernel.block_rows
    y.view(-1, br)[:] = torch.as_tensor(init, dtype=torch.float64, device=dev)
    op.evaluate(y, ones)
    return y if br == 1 else y.view(-1, br)


def create_sparse_operator(row_particles, col_particles, radius, kernel):
    """create_sparse_operator(rows, cols, radius | radius_function, f)
    (src/Operators.h:478-516).  `radius` may be a float or a per-row array
    (the FRadius overload evaluated on the row particles)."""
    if np.isscalar(radius):
        return SparseOperator(row_particles, col_particles, radius, kernel)
    return SparseOperator(row_particles, col_particles, None, kernel, radius_per_row=radius)
