"""Multi-GPU slab decomposition with halo exchange.

The reference is single process; this is the one parallel strategy the build
adds.  The domain is cut into slabs of bucket LAYERS along dimension 0 —
`collapse_index_vector` makes dimension 0 the slowest index
(src/detail/SpatialUtil.h:49-59), so a slab is a contiguous range of global
bucket numbers AND of the globally sorted particle array: concatenating the
owned ranges of all ranks reproduces the single-GPU cell list exactly.

One process per GPU.  The only data-path communication is with the two slab
neighbours (torch.distributed P2P: NCCL over NVLink on GPUs, gloo in the CPU
tests): halo layers of the sorted columns once per build, halo entries of `b`
once per product.  No global reduction: every row of y is owned by one rank.

  SlabExchange   pure torch + torch.distributed plumbing (device agnostic;
                 covered by the world_size-2 gloo tests)
  SlabParticles  a rank's particles on its GPU: owned build -> halo exchange ->
                 adopt the sorted [ghost_lo | owned | ghost_hi] set (libabr.so)
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist


def plan_layers(n_layers, world):
    """balanced contiguous split of the bucket layers of dimension 0"""
    base, rem = divmod(n_layers, world)
    out, lo = [], 0
    for g in range(world):
        hi = lo + base + (1 if g < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def plan_layers_balanced(layer_counts, world):
    """contiguous split of the bucket layers of dimension 0 by PARTICLE COUNT (prefix over
    the per-layer histogram): rank g ends at the layer where the running count is nearest to
    (g + 1) / world of the total; every rank keeps at least one layer"""
    c = np.asarray(layer_counts, dtype=np.float64)
    n_layers = len(c)
    if n_layers < world:
        raise ValueError("fewer bucket layers than ranks: replicas only (SURVEY §8e)")
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    cuts = [0]
    for g in range(1, world):
        target = total * g / world
        k = int(np.argmin(np.abs(cum - target)))
        k = max(k, cuts[-1] + 1)
        k = min(k, n_layers - (world - g))
        cuts.append(k)
    cuts.append(n_layers)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def halo_width(radius, side0):
    """bucket layers a row can reach: ceil(r / side) with the same rounding
    guard as the tiled kernel (aboria_b200/csrc/abr_matvec.cu)"""
    return max(1, int(math.ceil(radius / side0 - 1e-9)))


class SlabExchange:
    """Halo plumbing for one rank.

    layer_offsets: (own_n + 1,) int64 CPU tensor — offsets of the owned bucket
    layers inside the rank's SORTED owned arrays (layer l = rows
    layer_offsets[l] .. layer_offsets[l+1])."""

    def __init__(self, rank, world, periodic0, w, layer_offsets, group=None):
        self.rank, self.world, self.w, self.group = rank, world, w, group
        lo = layer_offsets.tolist()
        own_n = len(lo) - 1
        if own_n < w:
            raise ValueError(f"rank {rank}: owns {own_n} bucket layers, fewer than the halo width {w}: replicas only (SURVEY §8e)")
        self.n_own = lo[-1]
        self.lower = (rank - 1) % world if (periodic0 or rank > 0) else None
        self.upper = (rank + 1) % world if (periodic0 or rank < world - 1) else None
        # rows of my first / last w owned layers (what the neighbours need)
        self.send_lo = (0, lo[w])
        self.send_hi = (lo[own_n - w], lo[own_n])
        # sizes of the ghost ranges: exchanged once (tiny)
        counts = torch.tensor([self.send_lo[1] - self.send_lo[0], self.send_hi[1] - self.send_hi[0]], dtype=torch.int64)
        self.n_ghost_lo, self.n_ghost_hi = self._exchange_counts(counts)
        self.own_begin = self.n_ghost_lo
        self.own_end = self.n_ghost_lo + self.n_own
        self.n_local = self.own_end + self.n_ghost_hi

    def _dev(self):
        return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")

    def _exchange_counts(self, counts):
        dev = self._dev()
        mine = counts.to(dev)
        from_lower = torch.zeros(1, dtype=torch.int64, device=dev)
        from_upper = torch.zeros(1, dtype=torch.int64, device=dev)
        ops = []
        # order matters when lower == upper (world == 2): sends lo, hi; receives hi-ghost, lo-ghost
        if self.lower is not None:
            ops.append(dist.P2POp(dist.isend, mine[0:1], self.lower, self.group))
        if self.upper is not None:
            ops.append(dist.P2POp(dist.isend, mine[1:2], self.upper, self.group))
        if self.upper is not None:
            ops.append(dist.P2POp(dist.irecv, from_upper, self.upper, self.group))
        if self.lower is not None:
            ops.append(dist.P2POp(dist.irecv, from_lower, self.lower, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return int(from_lower.item()), int(from_upper.item())

    def _halo_ops(self, send_from, recv_into):
        """send_from(which) -> contiguous tensor slice to send; recv_into(which) -> slice to fill"""
        ops = []
        if self.lower is not None and self.send_lo[1] > self.send_lo[0]:
            ops.append(dist.P2POp(dist.isend, send_from("lo"), self.lower, self.group))
        if self.upper is not None and self.send_hi[1] > self.send_hi[0]:
            ops.append(dist.P2POp(dist.isend, send_from("hi"), self.upper, self.group))
        if self.upper is not None and self.n_ghost_hi > 0:
            ops.append(dist.P2POp(dist.irecv, recv_into("hi"), self.upper, self.group))
        if self.lower is not None and self.n_ghost_lo > 0:
            ops.append(dist.P2POp(dist.irecv, recv_into("lo"), self.lower, self.group))
        return ops

    def assemble(self, owned_sorted):
        """[ghost_lo | owned | ghost_hi] for one sorted owned column (any dtype,
        leading dimension = particles).  Ghost rows keep the sender's coordinates:
        the periodic image is applied by the search (cur = r + image*L,
        src/Search.h:188-190), never by moving a particle."""
        out = torch.empty((self.n_local,) + tuple(owned_sorted.shape[1:]), dtype=owned_sorted.dtype, device=owned_sorted.device)
        out[self.own_begin:self.own_end].copy_(owned_sorted)
        self.fill_halo(out)
        return out

    def fill_halo(self, local):
        """refresh the ghost ranges of a local column from the neighbours' owned
        ranges (used for b before every product)"""
        ob = self.own_begin

        def send_from(which):
            a, b = self.send_lo if which == "lo" else self.send_hi
            return local[ob + a: ob + b]

        def recv_into(which):
            return local[: self.n_ghost_lo] if which == "lo" else local[self.own_end:]

        ops = self._halo_ops(send_from, recv_into)
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return local

    def halo_bytes(self, row_bytes):
        return (self.n_ghost_lo + self.n_ghost_hi) * row_bytes


class _LocalLayout:
    """where the owned and ghost ranges of a rank's local (ghost-padded) arrays are, and the halo
    refresh of a per-particle vector (b before every product): contiguous slices, neighbours only"""

    def __init__(self, rank, world, lower, upper, n_lo, n_own, n_hi, n_send_lo, n_send_hi, group):
        self.rank, self.world, self.lower, self.upper, self.group = rank, world, lower, upper, group
        self.n_ghost_lo, self.n_own, self.n_ghost_hi = n_lo, n_own, n_hi
        self.own_begin, self.own_end = n_lo, n_lo + n_own
        self.n_local = n_lo + n_own + n_hi
        self.send_lo = (0, n_send_lo)
        self.send_hi = (n_own - n_send_hi, n_own)

    def fill_halo(self, local):
        ob = self.own_begin
        ops = []
        if self.lower is not None and self.send_lo[1] > self.send_lo[0]:
            ops.append(dist.P2POp(dist.isend, local[ob + self.send_lo[0]: ob + self.send_lo[1]], self.lower, self.group))
        if self.upper is not None and self.send_hi[1] > self.send_hi[0]:
            ops.append(dist.P2POp(dist.isend, local[ob + self.send_hi[0]: ob + self.send_hi[1]], self.upper, self.group))
        if self.upper is not None and self.n_ghost_hi > 0:
            ops.append(dist.P2POp(dist.irecv, local[self.own_end:], self.upper, self.group))
        if self.lower is not None and self.n_ghost_lo > 0:
            ops.append(dist.P2POp(dist.irecv, local[: self.n_ghost_lo], self.lower, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return local

    def assemble(self, owned_sorted):
        out = torch.empty((self.n_local,) + tuple(owned_sorted.shape[1:]), dtype=owned_sorted.dtype, device=owned_sorted.device)
        out[self.own_begin:self.own_end].copy_(owned_sorted)
        return self.fill_halo(out)

    def halo_bytes(self, row_bytes):
        return (self.n_ghost_lo + self.n_ghost_hi) * row_bytes


class SlabParticles:
    """One rank's share of a slab-decomposed particle set on its GPU."""

    def __init__(self, D, low, high, periodic, n_total, n_leaf, radius, rank, world, device, group=None):
        from . import _lib
        from .particles import Particles

        if D < 2:
            raise ValueError("slabs need D >= 2")
        self.D, self.rank, self.world, self.device, self.group = D, rank, world, device, group
        self.low = np.ascontiguousarray(np.broadcast_to(np.asarray(low, dtype=np.float64), (D,)))
        self.high = np.ascontiguousarray(np.broadcast_to(np.asarray(high, dtype=np.float64), (D,)))
        self.periodic = np.ascontiguousarray(np.broadcast_to(np.asarray(periodic), (D,)).astype(np.uint8))
        self.radius = float(radius)
        L = _lib.lib()
        size = np.zeros(D, dtype=np.uint32)
        side = np.zeros(D, dtype=np.float64)
        rc = L.abr_grid_for(D, self.low.ctypes.data, self.high.ctypes.data, float(n_leaf), int(n_total), size.ctypes.data, side.ctypes.data)
        if rc:
            raise RuntimeError("abr_grid_for failed")
        self.size, self.side = size, side
        self.layers = plan_layers(int(size[0]), world)
        self.lo_layer, self.hi_layer = self.layers[rank]
        self.own_n = self.hi_layer - self.lo_layer
        self.w = halo_width(self.radius, float(side[0]))
        if bool(self.periodic[0]) and self.own_n + 2 * self.w > int(size[0]):
            raise ValueError("slab window would wrap onto itself: replicas only (SURVEY §8e)")
        self.p = Particles(D, 0, device=device)
        self.per_layer = int(np.prod(size[1:], dtype=np.int64))
        self.ex = None

    def slab_bounds(self):
        """[x_lo, x_hi) of this rank's slab in dimension 0"""
        s = float(self.side[0])
        return float(self.low[0]) + self.lo_layer * s, float(self.low[0]) + self.hi_layer * s

    def _force(self, win_lo, win_n, own_lo, own_n):
        from ._lib import check

        p = self.p
        check(p._h, p._lib.abr_domain_force_grid(p._h, self.D, self.low.ctypes.data, self.high.ctypes.data, self.periodic.ctypes.data, self.size.ctypes.data))
        check(p._h, p._lib.abr_domain_set_window(p._h, win_lo, win_n, own_lo, own_n))

    def set_layers(self, layers):
        """use another contiguous layer split (plan_layers_balanced) — same on every rank"""
        self.layers = [tuple(map(int, x)) for x in layers]
        self.lo_layer, self.hi_layer = self.layers[self.rank]
        self.own_n = self.hi_layer - self.lo_layer
        if self.own_n < self.w:
            raise ValueError(f"rank {self.rank}: owns {self.own_n} bucket layers, fewer than the halo width {self.w}: replicas only (SURVEY §8e)")
        if bool(self.periodic[0]) and self.own_n + 2 * self.w > int(self.size[0]):
            raise ValueError("slab window would wrap onto itself: replicas only (SURVEY §8e)")
        self._big = None

    def _layers_of(self, pos):
        """bucket layer (dimension 0) of every particle, by the library's own key arithmetic; -1: the build would kill it"""
        from ._lib import check

        p = self.p
        self._force_global()
        p._sync_stream()
        pos = pos.contiguous()
        out = torch.empty(max(pos.shape[0], 1), dtype=torch.int32, device=self.device)
        check(p._h, p._lib.abr_slab_layers(p._h, C.c_void_p(pos.data_ptr()), pos.shape[0], C.c_void_p(out.data_ptr())))
        return out[: pos.shape[0]]

    def layer_histogram(self, pos):
        """global particle count per bucket layer of dimension 0 (one small all-reduce; set-up time only)"""
        S0 = int(self.size[0])
        layer = self._layers_of(pos).long()
        h = torch.bincount(layer[layer >= 0], minlength=S0).to(torch.float64)
        dist.all_reduce(h, group=self.group)
        return h.cpu().numpy()

    def cost_histogram(self, pos, cost):
        """global sum of a per-particle cost (e.g. accepted pairs per row from pair_stats) per bucket layer of
        dimension 0: the histogram to balance by when the work per particle is far from uniform (clustered
        clouds: pairs ~ density^2)"""
        S0 = int(self.size[0])
        layer = self._layers_of(pos).long()
        ok = layer >= 0
        h = torch.zeros(S0, dtype=torch.float64, device=self.device)
        h.index_add_(0, layer[ok], cost.to(torch.float64)[ok])
        dist.all_reduce(h, group=self.group)
        return h.cpu().numpy()

    def distribute(self, pos, columns=None):
        """One-time distribution of a global cloud: every rank hands in ANY share of the particles
        (e.g. a range of ids) and gets back the particles of its own slab (all-to-all by owner; set-up
        time only — the per-step traffic of a moving cloud is the neighbour-only migrate()).  Particles
        the build would kill (outside a non-periodic domain, non-finite) stay where they are."""
        dev = self.device
        columns = dict(columns or {})
        layer = self._layers_of(pos).long()
        his = torch.tensor([hi for _, hi in self.layers], dtype=torch.int64, device=dev)
        owner = torch.bucketize(layer, his, right=True).clamp_(max=self.world - 1)
        owner = torch.where(layer < 0, torch.full_like(owner, self.rank), owner)
        order = torch.argsort(owner, stable=True)
        send = torch.bincount(owner, minlength=self.world)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        send_l, recv_l = send.tolist(), recv.tolist()
        allc = dict(columns)
        allc["position"] = pos
        out = {}
        for k, t in allc.items():
            src = t.index_select(0, order).contiguous()
            dst = torch.empty((sum(recv_l),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            dist.all_to_all_single(dst, src, recv_l, send_l, group=self.group)
            out[k] = dst
        new_pos = out.pop("position")
        return new_pos, out

    def _neighbours(self):
        per0 = bool(self.periodic[0])
        lower = (self.rank - 1) % self.world if (per0 or self.rank > 0) else None
        upper = (self.rank + 1) % self.world if (per0 or self.rank < self.world - 1) else None
        return lower, upper

    def build(self, pos_owned_unsorted, extra_columns=None, assume_all_alive=True):
        """One build of this rank's share, ONE host synchronisation, ONE data exchange:
          1. ordinary build of the owned particles in the ghost-padded window (own layers in the
             middle; the reorder writes straight into the middle of ghost-padded column buffers);
          2. the particle counts of my first / last w layers travel to the neighbours as device
             tensors; the four counts (sent, received) come back to the host in one read;
          3. one batch of P2P operations: the halo slices of every sorted column (contiguous: a
             layer is a contiguous range of the sorted array) and of m_bucket_begin / m_bucket_end;
          4. abr_celllist_patch_ghosts: ghost bucket ranges from the senders' ranges, owned ranges
             shifted — no pass over the particles."""
        from ._lib import check

        p = self.p
        dev = self.device
        lower, upper = self._neighbours()
        has_lo, has_hi = lower is not None, upper is not None
        w, per_layer = self.w, self.per_layer
        own_lo = w if has_lo else 0
        win_lo = self.lo_layer - own_lo
        win_n = self.own_n + own_lo + (w if has_hi else 0)
        # the container's input columns: persistent buffers refilled in place (no allocation, no arange on the
        # hot path); the caller's tensors stay untouched (the build wraps positions in place)
        n_in = pos_owned_unsorted.shape[0]
        extra = extra_columns or {}
        stage = getattr(self, "_stage", None)
        if stage is None or stage["position"].shape[0] != n_in or set(stage) != {"position", "id", "alive", *extra} or any(
                stage[k].dtype != v.dtype or stage[k].shape[1:] != v.shape[1:] for k, v in extra.items()):
            stage = {"position": torch.empty((n_in, self.D), dtype=torch.float64, device=dev), "id": torch.empty(n_in, dtype=torch.int64, device=dev),
                     "alive": torch.empty(n_in, dtype=torch.uint8, device=dev)}
            for k, v in extra.items():
                stage[k] = torch.empty_like(v)
            self._stage = stage
            self._iota = torch.arange(n_in, dtype=torch.int64, device=dev)
        stage["position"].copy_(torch.as_tensor(pos_owned_unsorted), non_blocking=True)
        stage["id"].copy_(self._iota)
        stage["alive"].fill_(1)
        for k, v in extra.items():
            stage[k].copy_(v)
        p.columns = dict(stage)
        p.low, p.high, p.periodic = self.low, self.high, self.periodic
        cap = max(getattr(self, "_cap", 0), n_in // 8 + 4096)
        if hasattr(self, "_test_cap"):  # tests: a rank-dependent reserve, so that only some ranks outgrow it
            cap = self._test_cap(n_in)
        big = getattr(self, "_big", None)
        if big is None or self._big_n != n_in or self._cap != cap or set(big) != set(p.columns) or any(
                big[k].dtype != v.dtype or big[k].shape[1:] != v.shape[1:] for k, v in p.columns.items()):
            big = {k: torch.empty((n_in + 2 * cap,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device) for k, v in p.columns.items()}
            self._big, self._big_n, self._cap = big, n_in, cap
        p._other = {k: big[k][cap: cap + n_in] for k in p.columns}
        # 1. owned build in the padded window
        self._force(win_lo, win_n, own_lo, self.own_n)
        n_own = p.update_positions(assume_all_alive=assume_all_alive)
        self.order_owned = p.get_alive_indicies()
        _, bb, be = p._bucket_view(clone=False, sync=False)
        c0 = own_lo * per_layer                       # first owned bucket
        c1 = c0 + self.own_n * per_layer              # one past the last owned bucket
        # 2. counts: [sent to lower, sent to upper] -> neighbours; everything to the host in ONE read
        first = bb[c0]  # 0 unless a stray particle sits in a lower ghost layer (caught below)
        mine = torch.stack([(bb[c0 + w * per_layer] if self.own_n > w else be[c1 - 1]) - first, be[c1 - 1] - bb[c1 - w * per_layer], be[c1 - 1] - first]).to(torch.int64)
        got = torch.zeros(2, dtype=torch.int64, device=dev)
        ops = []
        if has_lo:
            ops.append(dist.P2POp(dist.isend, mine[0:1], lower, self.group))
        if has_hi:
            ops.append(dist.P2POp(dist.isend, mine[1:2], upper, self.group))
        if has_hi:
            ops.append(dist.P2POp(dist.irecv, got[1:2], upper, self.group))
        if has_lo:
            ops.append(dist.P2POp(dist.irecv, got[0:1], lower, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        n_send_lo, n_send_hi, n_own_dev, n_lo, n_hi = [int(v) for v in torch.cat([mine, got]).tolist()]  # the one host synchronisation
        if n_own_dev != n_own:
            raise RuntimeError(f"slab build: {n_own - n_own_dev} particle(s) died or left this rank's slab during the owned build")
        p.check_async()
        if not has_lo:
            n_lo = 0
        if not has_hi:
            n_hi = 0
        if n_lo > cap or n_hi > cap:
            # halo larger than the reserve on THIS rank (other ranks may be fine: nothing collective may be repeated):
            # move the sorted owned columns into larger ghost-padded buffers — a local copy — and go on
            new_cap = max(n_lo, n_hi) * 5 // 4 + 4096
            big2 = {k: torch.empty((n_in + 2 * new_cap,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device) for k, v in big.items()}
            for k in big:
                big2[k][new_cap: new_cap + n_own].copy_(big[k][cap: cap + n_own])
            big = self._big = big2
            cap = self._cap = new_cap
        # 3. one batch: halo slices of every sorted column + bucket ranges of those layers
        nb_w = w * per_layer
        gh = getattr(self, "_ghost_ranges", None)
        if gh is None or gh.shape[1] != nb_w:
            gh = self._ghost_ranges = torch.empty((4, max(nb_w, 1)), dtype=torch.int32, device=dev)
        ops = []
        cols = {k: big[k][cap - n_lo: cap + n_own + n_hi] for k in p.columns}
        for k in p.columns:
            t = big[k]
            if has_lo and n_send_lo > 0:
                ops.append(dist.P2POp(dist.isend, t[cap: cap + n_send_lo], lower, self.group))
            if has_hi and n_send_hi > 0:
                ops.append(dist.P2POp(dist.isend, t[cap + n_own - n_send_hi: cap + n_own], upper, self.group))
            if has_hi and n_hi > 0:
                ops.append(dist.P2POp(dist.irecv, t[cap + n_own: cap + n_own + n_hi], upper, self.group))
            if has_lo and n_lo > 0:
                ops.append(dist.P2POp(dist.irecv, t[cap - n_lo: cap], lower, self.group))
        for src, (to_lo, to_hi) in ((bb, (0, 2)), (be, (1, 3))):
            if has_lo:
                ops.append(dist.P2POp(dist.isend, src[c0: c0 + nb_w], lower, self.group))
            if has_hi:
                ops.append(dist.P2POp(dist.isend, src[c1 - nb_w: c1], upper, self.group))
            if has_hi:
                ops.append(dist.P2POp(dist.irecv, gh[to_hi][:nb_w], upper, self.group))
            if has_lo:
                ops.append(dist.P2POp(dist.irecv, gh[to_lo][:nb_w], lower, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        p.columns = cols
        p._other = {}  # the reorder buffers of the owned build are now the live columns
        # 4. bucket ranges of the local set
        p._sync_stream()
        pos = p.columns["position"]
        check(p._h, p._lib.abr_celllist_patch_ghosts(p._h, C.c_void_p(pos.data_ptr()), n_lo, n_own, n_hi,
                                                     C.c_void_p(gh[0].data_ptr()) if has_lo else None, C.c_void_p(gh[1].data_ptr()) if has_lo else None,
                                                     C.c_void_p(gh[2].data_ptr()) if has_hi else None, C.c_void_p(gh[3].data_ptr()) if has_hi else None))
        p.searchable = True
        self.ex = _LocalLayout(self.rank, self.world, lower, upper, n_lo, n_own, n_hi, n_send_lo, n_send_hi, self.group)
        return self.ex.n_local

    def migrate(self, pos, columns=None):
        """Neighbour-only particle migration (the moving-particle loops of tests/md.h:318-326
        on more than one GPU): particles whose bucket layer left this rank's slab go to the lower / upper
        neighbour, arrivals are appended.  `pos` (n x D, unsorted, moved) and the other per-particle
        columns are returned as new tensors (unchanged objects when nothing moves anywhere near this rank).
        A particle that moved further than one slab is reported by the next build."""
        from ._lib import check

        p = self.p
        dev = self.device
        columns = dict(columns or {})
        lower, upper = self._neighbours()
        n = pos.shape[0]
        self._force_global()
        cls = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        counts = torch.zeros(3, dtype=torch.int32, device=dev)
        p._sync_stream()
        check(p._h, p._lib.abr_slab_classify(p._h, C.c_void_p(pos.data_ptr()), n, self.lo_layer, self.hi_layer, C.c_void_p(cls.data_ptr()), C.c_void_p(counts.data_ptr())))
        mine = counts[1:3].to(torch.int64)
        got = torch.zeros(2, dtype=torch.int64, device=dev)
        ops = []
        if lower is not None:
            ops.append(dist.P2POp(dist.isend, mine[0:1], lower, self.group))
        if upper is not None:
            ops.append(dist.P2POp(dist.isend, mine[1:2], upper, self.group))
        if upper is not None:
            ops.append(dist.P2POp(dist.irecv, got[1:2], upper, self.group))
        if lower is not None:
            ops.append(dist.P2POp(dist.irecv, got[0:1], lower, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        s_lo, s_hi, r_lo, r_hi = [int(v) for v in torch.cat([mine, got]).tolist()]
        if lower is None and s_lo or upper is None and s_hi:
            raise RuntimeError("slab migrate: a particle left a non-periodic domain towards a missing neighbour")
        if s_lo + s_hi + r_lo + r_hi == 0:
            return pos, columns
        cls = cls[:n]
        idx_lo = torch.nonzero(cls == 1).reshape(-1)
        idx_hi = torch.nonzero(cls == 2).reshape(-1)
        keep = torch.nonzero(cls == 0).reshape(-1)
        allc = dict(columns)
        allc["position"] = pos
        out = {}
        ops = []
        recv = {}
        for k, t in allc.items():
            send_lo = t.index_select(0, idx_lo).contiguous()
            send_hi = t.index_select(0, idx_hi).contiguous()
            new = torch.empty((keep.shape[0] + r_lo + r_hi,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            new[: keep.shape[0]] = t.index_select(0, keep)
            out[k] = new
            recv[k] = (send_lo, send_hi)
            nk = keep.shape[0]
            if lower is not None and s_lo:
                ops.append(dist.P2POp(dist.isend, send_lo, lower, self.group))
            if upper is not None and s_hi:
                ops.append(dist.P2POp(dist.isend, send_hi, upper, self.group))
            if upper is not None and r_hi:
                ops.append(dist.P2POp(dist.irecv, new[nk + r_lo: nk + r_lo + r_hi], upper, self.group))
            if lower is not None and r_lo:
                ops.append(dist.P2POp(dist.irecv, new[nk: nk + r_lo], lower, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        new_pos = out.pop("position")
        return new_pos, out

    def _force_global(self):
        from ._lib import check

        p = self.p
        check(p._h, p._lib.abr_domain_force_grid(p._h, self.D, self.low.ctypes.data, self.high.ctypes.data, self.periodic.ctypes.data, self.size.ctypes.data))

    def owned(self, local):
        return local[self.ex.own_begin: self.ex.own_end]

    def local_vector(self, owned_values):
        """place owned values into a local (ghost-padded) vector"""
        v = torch.zeros((self.ex.n_local,) + tuple(owned_values.shape[1:]), dtype=owned_values.dtype, device=owned_values.device)
        v[self.ex.own_begin: self.ex.own_end] = owned_values
        return v

    def matvec(self, op, b_local):
        """y_owned = (K b)_owned: halo exchange of b, then the local product"""
        self.ex.fill_halo(b_local)
        y = op.matvec(b_local)
        return y


class SlabHostPipeline:
    """Host-buffer steps on a slab rank (the multi-GPU counterpart of
    aboria_b200.pipeline.HostPipeline): every step uploads this rank's positions and b from
    pinned host memory, builds (owned build, halo exchange, adopt), multiplies (halo exchange
    of b, product) and downloads the owned part of y.  The uploads of step k run on a copy
    stream while step k-1 computes on the main stream, the download of step k-1 on a third
    stream; NCCL calls stay on the main stream in the same order on every rank."""

    def __init__(self, sp, op, n_mine, device):
        self.sp, self.op, self.dev = sp, op, device
        self.pos_dev = [torch.empty((n_mine, sp.D), dtype=torch.float64, device=device) for _ in range(2)]
        self.b_dev = [torch.empty(n_mine, dtype=torch.float64, device=device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device)
        self.d2h_stream = torch.cuda.Stream(device)
        self.ready = [None, None]
        self.consumed = [None, None]
        self.pending = None
        self.k = 0
        self.b_local = None

    def submit(self, pos_host, b_host, y_host):
        i = self.k % 2
        self.k += 1
        if self.consumed[i] is not None:
            self.copy_stream.wait_event(self.consumed[i])  # step k-2 has read these buffers
        with torch.cuda.stream(self.copy_stream):
            self.pos_dev[i].copy_(torch.as_tensor(pos_host), non_blocking=True)  # H2D positions
            self.b_dev[i].copy_(torch.as_tensor(b_host), non_blocking=True)      # H2D b
            self.ready[i] = torch.cuda.Event()
            self.ready[i].record(self.copy_stream)
        prev, self.pending = self.pending, (i, y_host)
        if prev is not None:
            self._run(*prev)

    def _run(self, i, y_host):
        sp = self.sp
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(self.ready[i])
        sp.build(self.pos_dev[i])
        if self.b_local is None or self.b_local.shape[0] != sp.ex.n_local:
            self.b_local = torch.zeros(sp.ex.n_local, dtype=torch.float64, device=self.dev)
        self.b_local[sp.ex.own_begin: sp.ex.own_end].copy_(self.b_dev[i])
        self.consumed[i] = torch.cuda.Event()
        self.consumed[i].record(main)
        y = sp.matvec(self.op, self.b_local)
        done = torch.cuda.Event()
        done.record(main)
        self.d2h_stream.wait_event(done)
        with torch.cuda.stream(self.d2h_stream):
            yo = sp.owned(y)
            torch.as_tensor(y_host).copy_(yo, non_blocking=True)                # D2H y
            y.record_stream(self.d2h_stream)

    def wait(self):
        if self.pending is not None:
            prev, self.pending = self.pending, None
            self._run(*prev)
        self.d2h_stream.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()


def bind_to_gpu_numa_node(dev):
    """Run this process on the cores of the NUMA node its GPU hangs off, so that the pinned host buffers it
    allocates next live in that node's memory (first touch).  With eight ranks staging 1 GB per step each from
    ONE node the host side of the box, not PCIe, bounds the end-to-end step (round 1: 0.32 weak-scaling
    efficiency end to end at 8 GPUs).  Returns (node, n_cpus) or None when the topology cannot be read."""
    import os

    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except Exception:
        return None


def run_bench(args, wl, rank, world, dev, metric, unit, emit=None):
    """bench.py body for N > 1: slabs along dimension 0.  `wl` (bench.Workload) names the cloud: weak
    scaling (n_per_gpu particles per GPU), strong scaling (n_total fixed) or the clustered cloud of
    BASELINE config 4.  Every rank generates a RANGE OF IDS of the global cloud — particles anywhere in the
    domain — and the owners are found by a one-time distribution (outside the timed region); the layer
    split is balanced by the global per-layer particle histogram."""
    import json
    import os
    import sys
    import time

    import aboria_b200 as ab
    from aboria_b200 import synth

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import N_LEAF, ClockSampler

    if emit is None:
        emit = lambda line: print(json.dumps(line))  # noqa: E731

    n_total = wl.n_total
    radius = wl.radius
    sp = SlabParticles(3, 0.0, 1.0, wl.periodic, n_total, N_LEAF, radius, rank, world, dev)
    size = int(sp.size[0])
    # this rank's share of the global cloud (a range of ids), then to the owners
    first_id, n_share = wl.share(rank)
    pos_any = wl.positions(rank, dev)
    gid = torch.arange(first_id, first_id + n_share, dtype=torch.int64, device=dev)
    hist = sp.layer_histogram(pos_any)
    sp.set_layers(plan_layers_balanced(hist, world))
    per_rank = [int(hist[lo:hi].sum()) for lo, hi in sp.layers]
    pos_unsorted, cols = sp.distribute(pos_any, {"gid": gid})
    del pos_any, gid
    n_mine = pos_unsorted.shape[0]
    assert n_mine == per_rank[rank], (n_mine, per_rank)
    # b by GLOBAL particle id (compared across GPU counts by id); the product indexes it by post-reorder position
    b_by_id_mine = torch.from_numpy(synth.vector(n_total)).to(dev)[cols["gid"]] if n_total <= 64_000_000 else None
    if b_by_id_mine is None:  # large clouds: generate only this rank's values
        ids_np = cols["gid"].cpu().numpy().astype(np.uint64)
        b_by_id_mine = torch.from_numpy(synth.uniform01(synth.SEED + 1, ids_np)).to(dev)
        del ids_np
    del cols
    op = ab.create_sparse_operator(sp.p, sp.p, radius, wl.kernel)
    state = {}
    balance = "particle count per layer"
    if wl.cloud == "clustered":
        # work per particle is far from uniform on a clustered cloud (pairs ~ density^2): re-split the layers by
        # the accepted-pair count per layer (one build + stats pass, set-up time only) and redistribute
        sp.build(pos_unsorted)
        cnt0, _ = sp.p.pair_stats(radius)
        cost_in = torch.zeros(n_mine, dtype=torch.float64, device=dev)
        cost_in[sp.order_owned.long()] = sp.owned(cnt0).to(torch.float64)
        cost_hist = sp.cost_histogram(pos_unsorted, cost_in)
        del cnt0
        new_layers = plan_layers_balanced(cost_hist, world)
        if [tuple(x) for x in new_layers] != [tuple(x) for x in sp.layers]:
            sp.set_layers(new_layers)
            pos_unsorted, cols2 = sp.distribute(pos_unsorted, {"b": b_by_id_mine})
            b_by_id_mine = cols2["b"]
            n_mine = pos_unsorted.shape[0]
        balance = "accepted pairs per layer (stats pass at set-up)"
        del cost_in

    def step(evs=None):
        sp.build(pos_unsorted)  # copied into the container (the bench input itself stays unsorted)
        if evs is not None:
            evs.append(torch.cuda.Event(enable_timing=True))
            evs[-1].record()
        b_local = state.get("b_local")
        if b_local is None or b_local.shape[0] != sp.ex.n_local:
            b_local = state["b_local"] = torch.zeros(sp.ex.n_local, dtype=torch.float64, device=dev)
        # b is indexed by post-reorder position (src/Kernels.h:745-748); the input is the same every step, so is
        # the order: the owned entries of b are gathered once, the ghosts come from the neighbours every product
        if "b_sorted" not in state:
            state["b_sorted"] = b_by_id_mine[sp.order_owned.long()]
        b_local[sp.ex.own_begin: sp.ex.own_end].copy_(state["b_sorted"])
        y = sp.matvec(op, b_local)
        if evs is not None:
            evs.append(torch.cuda.Event(enable_timing=True))
            evs[-1].record()
        return y

    step()
    cnt, hs = sp.p.pair_stats(radius)
    pairs_t = sp.owned(cnt).long().sum()
    dist.all_reduce(pairs_t)
    pairs = int(pairs_t.item())
    # parity guards before anything is timed: (1) the global pair count is the one a uniform cloud of this
    # density must have (the single-GPU count of the same workload: tests/test_gpu_parity.py::test_c5_32m_sampled_oracle);
    # (2) sampled owned rows: the cell-tiled kernel on the ghost-padded local set finds exactly the pair sets
    # (count + hash) the exact per-row iterator walk finds on this rank
    expect = wl.expected_pairs()
    if expect is not None:
        assert abs(pairs - expect) / expect < 2e-3, f"global pair count {pairs} differs from the expectation {expect:.6g} of a uniform cloud"
    sub = torch.arange(sp.ex.own_begin, sp.ex.own_end, 4001, device=dev)
    cnt_w, hs_w = sp.p.pair_stats(radius, rows=sp.p.get("position")[sub].contiguous(), path=1)
    assert bool((cnt[sub] == cnt_w).all()) and bool((hs[sub] == hs_w).all()), f"rank {rank}: tiled pair sets differ from the exact walk on sampled owned rows"
    pairs_rank = int(sp.owned(cnt).long().sum().item())
    del hs, cnt_w, hs_w, cnt
    for _ in range(max(0, args.warmup - 1)):
        step()
    torch.cuda.synchronize()
    launches0 = sp.p.last_counters()["total_launches"]
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evs = []
    for _ in range(args.steps):
        step(evs)
    e1.record()
    torch.cuda.synchronize()
    ms_mv = float(np.mean([evs[2 * k].elapsed_time(evs[2 * k + 1]) for k in range(args.steps)]))  # b gather + halo exchange of b + product, this rank
    ms_build = float(np.mean([(evs[2 * k - 1] if k else e0).elapsed_time(evs[2 * k]) for k in range(args.steps)]))
    ms_local = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.barrier()
    dist.all_reduce(ms_local, op=dist.ReduceOp.MAX)  # max over ranks
    clocks = sampler.stop() if rank == 0 else None
    launches = sp.p.last_counters()["total_launches"] - launches0
    ms_per_step = float(ms_local.item()) / args.steps
    value = pairs / (ms_per_step * 1e-3)
    per_rank_ms = torch.zeros(world, 3, dtype=torch.float64, device=dev)
    per_rank_ms[rank, 0], per_rank_ms[rank, 1], per_rank_ms[rank, 2] = ms_build, ms_mv, float(n_mine)
    dist.all_reduce(per_rank_ms)

    # end-to-end with host buffers: H2D of positions and b, D2H of y, every step
    numa = bind_to_gpu_numa_node(dev)  # pinned buffers in the memory of the GPU's own NUMA node
    pos_host = torch.empty((n_mine, 3), dtype=torch.float64, pin_memory=True)
    pos_host.copy_(pos_unsorted)
    b_sorted = state["b_sorted"]
    b_host = torch.empty(n_mine, dtype=torch.float64, pin_memory=True)
    b_host.copy_(b_sorted)
    y_host = torch.empty(n_mine, dtype=torch.float64, pin_memory=True)

    pipe = SlabHostPipeline(sp, op, n_mine, dev)
    y_hosts = [y_host, torch.empty(n_mine, dtype=torch.float64, pin_memory=True)]
    for k in range(2):
        pipe.submit(pos_host, b_host, y_hosts[k % 2])
    pipe.wait()
    e2e_steps = max(4, min(args.steps, 10))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        pipe.submit(pos_host, b_host, y_hosts[k % 2])
    pipe.wait()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_sec = float(t_e2e.item()) / e2e_steps
    # the pipelined steps computed what the device-resident step computes
    y_ref = sp.owned(step()).cpu()
    for yh in y_hosts:
        rel = float(torch.linalg.norm(yh - y_ref) / torch.linalg.norm(y_ref))
        assert rel <= 1e-12, f"rank {rank}: pipelined e2e result differs (rel L2 {rel:.3e})"
    # a checksum of the product by particle id: the same number at every GPU count (up to summation order)
    chk = torch.stack([y_ref.sum().to(dev), (y_ref * y_ref).sum().to(dev)])
    dist.all_reduce(chk)
    halo = torch.tensor([float(sp.ex.n_ghost_lo + sp.ex.n_ghost_hi)], dtype=torch.float64, device=dev)
    dist.all_reduce(halo)
    n_all = torch.tensor([float(n_mine)], dtype=torch.float64, device=dev)
    dist.all_reduce(n_all)
    if rank == 0:
        from bench import measured_peaks

        hbm_peak, peak_src = measured_peaks()
        n_loc, n_own = sp.ex.n_local, sp.ex.n_own
        ncells_loc = int(sp.per_layer) * (sp.own_n + 2 * sp.w)
        mv_bytes = n_own * (8 * 3 + 8) + n_loc * (8 * 3 + 8) + 8 * ncells_loc  # SURVEY §8d B_mv for this rank's rows / columns
        roofline = {"bound": "hbm", "kernel": f"abr::tiled_kernel<3, {wl.kernel_name}> on rank 0 (sparse matvec incl. the halo exchange of b)",
                    "achieved": mv_bytes / (ms_mv * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": mv_bytes / (ms_mv * 1e-3) / 1e9 / hbm_peak,
                    "traffic": None, "algorithmic_bytes": mv_bytes, "peak_source": peak_src, "ms_matvec_rank0": ms_mv,
                    "pairs_per_s_matvec_only_rank0": pairs_rank / (ms_mv * 1e-3),
                    "note": "the product is instruction-issue / LSU bound, not HBM bound (DESIGN.md §4.2); same kernel as the 1-GPU line"}
        prm = per_rank_ms.cpu().numpy()
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name + f"; slabs along dim 0 (layer split balanced by {balance}), NCCL halo exchange",
                       "n_particles_per_gpu": [int(v) for v in prm[:, 2]], "n_particles": int(n_all.item()), "buckets": int(np.prod(sp.size)), "radius": radius,
                       "pairs_per_matvec": pairs, "halo_particles_total": int(halo.item()), "parallelism": f"slab{world}", "layers": [list(x) for x in sp.layers],
                       "y_checksum": [float(chk[0]), float(chk[1])], "l2": "inputs exceed the 126 MB L2; no flush needed"},
            "per_rank": {"ms_build": [float(v) for v in prm[:, 0]], "ms_matvec": [float(v) for v in prm[:, 1]]},
            "e2e": {"value": pairs / e2e_sec, "unit": unit, "h2d_bytes_per_step": int(n_all.item()) * 32, "d2h_bytes_per_step": int(n_all.item()) * 8,
                    "ms_per_step": e2e_sec * 1e3, "steps": e2e_steps,
                    "how": "SlabHostPipeline per rank: pinned host buffers (allocated on the GPU's own NUMA node); uploads of step k on a copy stream while step k-1 computes, "
                           "download of y on a third stream; wall clock, max over ranks", "numa_node_rank0": numa[0] if numa else None},
            "gpu_launches": int(launches) * world, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": None,
        }
        emit(line)
    dist.barrier()
    dist.destroy_process_group()
