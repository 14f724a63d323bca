/* abr.h — C-ABI of the B200-native ordered cell list + sparse kernel matvec.
 *
 * This is the drop-in boundary for ONE hot path of Aboria (reference at
 * /root/reference): CellListOrdered build followed by
 * create_sparse_operator(...) * b.  The reference has no FFI; its seam is C++
 * template substitution.  Each entry point below names the
 * reference interface it replaces (file:line under /root/reference).  The C++
 * shim in include/aboria_b200/Aboria.h keeps the reference's own names
 * (Particles, init_neighbour_search, create_sparse_operator, K*b) on top of
 * these calls; see INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; every array argument is a DEVICE pointer
 *    unless its name ends in _host.
 *  - all work is enqueued on the handle's CUDA stream; functions that return a
 *    value through a *_host pointer synchronise that stream first.
 *  - return value 0 = ok, nonzero = error (abr_last_error_string explains).
 *    The reference prints and raises SIGTRAP (src/Log.h:47-62) and never
 *    returns codes; the C++ shim turns a nonzero code into the same CHECK.
 *  - there is NO CPU fallback: every call fails with ABR_ERR_CUDA when no
 *    CUDA device / kernel image is available.
 */
#ifndef ABR_H_
#define ABR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABR_MAX_D 3
#define ABR_MAX_VARS 4
#define ABR_MAX_PARAMS 8

enum abr_status {
  ABR_OK = 0,
  ABR_ERR_INVALID = 1, /* bad argument */
  ABR_ERR_CUDA = 2,    /* CUDA runtime error / no device */
  ABR_ERR_STATE = 3,   /* call order (e.g. matvec before build) */
  ABR_ERR_UNSUPPORTED = 4
};

/* Built-in device functors for F(dx, a_i, b_j) (src/detail/Kernels.h:336-357).
 * Values mirror the reference tests' lambdas; custom functors go through
 * abr_sparse_matvec_custom + include/aboria_b200/device_kernel.cuh. */
enum abr_kernel_id {
  ABR_K_CONST_SUM = 0,      /* s1(a)+s2(b)                 tests/operators.h:842-847 */
  ABR_K_CONST_SUM_DIFF = 1, /* 2x1 (s1(a)+s2(b), s1(a)-s2(b)) tests/operators.h:905-911 */
  ABR_K_INV_DIST = 2,       /* 1/(|dx|+p0)                 SURVEY §8d c1 */
  ABR_K_INV_DIST_AA = 3,    /* a_i a_j/(|dx|+p0)           tests/operators.h:251-256 */
  ABR_K_WENDLAND_C2 = 4,    /* (2-|dx|/p0)^4 (1+2|dx|/p0)  tests/rbf_interpolation.h:310-313 */
  ABR_K_LJ_FORCE = 5,       /* Dx1: 24 p1 (2(p0/r)^12-(p0/r)^6)/r^2 dx  tests/md.h:166-174 pattern */
  ABR_K_SPH_DENSITY = 6,    /* p1 * W(|dx|, p0), p2 = WCON tests/sph.h:154-165 */
  ABR_K_SPH_PRESSURE = 7,   /* Dx1: p1 (rv0_i + cv0_j) F(|dx|,p0) dx    tests/sph.h:140-152, :333-339 */
  ABR_K_LINEAR_SPRING = 8,  /* Dx1: -p0 (p1/|dx| - 1) dx, 0 at |dx| = 0   tests/md.h:166-174 (k = p0, diameter = p1) */
  ABR_K_COUNT_ = 9
};

/* Describes the kernel function of one sparse operator block
 * (KernelSparse<Row,Col,FRadius,FWithDx>, src/Kernels.h:578-607). */
typedef struct abr_kernel_desc {
  int32_t kernel_id;                       /* enum abr_kernel_id */
  int32_t block_rows, block_cols;          /* BR, BC (src/Kernels.h:599-600) */
  int32_t reserved;
  double params[ABR_MAX_PARAMS];           /* captured scalars of the lambda */
  const double *row_vars[ABR_MAX_VARS];    /* per-row-particle variable columns */
  const double *col_vars[ABR_MAX_VARS];    /* per-col-particle variable columns */
} abr_kernel_desc;

typedef struct abr_handle_s *abr_handle;

/* One handle per (GPU, particle set): owns what the search object owns in the
 * reference — m_bucket_begin/end/indices (src/CellListOrdered.h:272-283) and
 * m_alive_indices (src/NeighbourSearchBase.h) — plus scratch.  `stream` is a
 * cudaStream_t (NULL = legacy default stream). */
int abr_create(abr_handle *out, int device, void *stream);
int abr_destroy(abr_handle h);
int abr_set_stream(abr_handle h, void *stream);
int abr_synchronize(abr_handle h);
/* Synchronises and verifies the last asynchronous abr_update_positions: error if
 * a particle died or a bucket index overflowed (then the update must be redone
 * with n_alive_host != NULL). */
int abr_check_async(abr_handle h);
/* Tuning knobs: "two_level_min_n" — particle count from which abr_update_positions uses
 * the two-level (partition + bin-local sort) build; "stage_records" — 1 (default): its
 * record partition stages each tile's column windows in shared memory with
 * cp.async.bulk when they fit, 0: records read directly from L2; "gather_slots" — 1
 * (default): its final reorder keeps the loads of every column in flight at once;
 * "skip_alive_move" — 1 (default): the reordered alive column is written as a run of
 * ones (every particle that survives the reorder is alive) instead of being moved — a
 * caller that stores other non-zero values in `alive` gets them normalised to 1;
 * "bounds_one_sweep" — 1 (default): bucket_begin/end from one sweep over the sorted keys;
 * "record_aos" — 1: binned copy as one record per particle (measured slower, default 0);
 * "phased_gather" — 0/1;
 * "matvec_variant" — 0: cell-tiled kernel gathering candidates from L2, 1: candidates
 * staged in shared memory with cp.async.bulk; "symmetric" — 1: products with rows ==
 * columns whose functor declares SYMMETRY evaluate every unordered pair once (half
 * stencil, src/Search.h:498-764) and add it to both rows with fp64 reductions; results
 * then differ from the ordered kernel by summation order only (<= 1e-12 relative) and
 * are no longer bit-reproducible from run to run.  None of them changes a pair set. */
int abr_set_option(abr_handle h, const char *name, double value);
const char *abr_last_error_string(abr_handle h);
const char *abr_version(void);

/* neighbour_search_base::set_domain + CellListOrdered::set_domain_impl
 * (src/NeighbourSearchBase.h:252-268, src/CellListOrdered.h:132-186).  Host
 * scalar math in the reference's expression order, including the rule that the
 * grid is only recomputed when the alive count leaves [1/2, 2] x the count it
 * was last computed with (:134-135).  Arrays are HOST pointers of length D. */
int abr_domain_set(abr_handle h, int D, const double *bmin_host, const double *bmax_host,
                   const uint8_t *periodic_host, double n_particles_in_leaf);
/* m_size, m_bucket_side_length, number of buckets (host out, length D). */
int abr_domain_get(abr_handle h, uint32_t *size_host, double *side_host, uint64_t *n_buckets_host);
/* Bypass the occupancy rule and fix the grid (tests that build
 * point_to_bucket_index by hand, tests/utils.h:76-93; slab ranks). */
int abr_domain_force_grid(abr_handle h, int D, const double *bmin_host, const double *bmax_host,
                          const uint8_t *periodic_host, const uint32_t *size_host);

/* The grid CellListOrdered::set_domain_impl would choose for n particles
 * (src/CellListOrdered.h:140-157), as a pure host function. */
int abr_grid_for(int D, const double *bmin_host, const double *bmax_host, double n_particles_in_leaf,
                 size_t n, uint32_t *size_host, double *side_host);

/* Multi-GPU slabs (no counterpart in the single-process
 * reference).  After abr_domain_force_grid with the GLOBAL grid, restrict this
 * handle to the bucket layers win_lo .. win_lo+win_n-1 of dimension 0
 * (unwrapped numbering: a window of a periodic dimension may start below 0 or
 * end past size[0]); rows are computed for the local layers own_lo ..
 * own_lo+own_n-1 only (the others are ghost layers received from neighbours).
 * All bucket/key arithmetic stays that of the global grid, so concatenating
 * the owned ranges of all ranks reproduces the single-GPU cell list exactly. */
int abr_domain_set_window(abr_handle h, int win_lo, int win_n, int own_lo, int own_n);

/* Adopt a particle set that is ALREADY sorted by (local) bucket — a rank's
 * [ghost_lo | owned | ghost_hi] concatenation after the halo exchange:
 * computes keys and the bucket ranges, checks sortedness, binds the query.
 * No permutation, no reorder. */
int abr_celllist_adopt_sorted(abr_handle h, double *pos_sorted, uint8_t *alive, size_t n);

/* Slabs, cheaper second step: after abr_update_positions in the ghost-padded window (own
 * layers in the middle, the ghost layers still empty) and the halo exchange, adopt the local
 * array [ghost_lo | owned | ghost_hi] WITHOUT a pass over the particles: owned bucket ranges
 * shift by n_ghost_lo, ghost buckets take the sender's ranges — bb_lo/be_lo: the lower
 * neighbour's m_bucket_begin/end of the layers it sent (own_lo * prod(size[1..]) entries, its own
 * numbering, rebased here), bb_hi/be_hi likewise for the upper neighbour; device pointers, a side
 * without ghost layers passes NULL.  Binds the query to pos_local.  m_bucket_indices (the sorted
 * key array) is not kept on this path. */
int abr_celllist_patch_ghosts(abr_handle h, const double *pos_local, size_t n_ghost_lo, size_t n_own, size_t n_ghost_hi,
                              const uint32_t *bb_lo, const uint32_t *be_lo, const uint32_t *bb_hi, const uint32_t *be_hi);

/* Slabs, particle migration: for every (unsorted, moved) particle of this rank decide from the
 * GLOBAL grid (abr_domain_force_grid) whether its bucket layer in dimension 0 is still inside
 * [lo_layer, hi_layer) — cls 0 — or now belongs to the lower (1) / upper (2) neighbour (the
 * nearer slab face, periodic wrap applied as enforce_domain would); counts3[k] = particles of
 * class k (device, 3 x u32).  Particles outside a non-periodic domain or non-finite stay (the
 * build kills them). */
int abr_slab_classify(abr_handle h, const double *pos, size_t n, int lo_layer, int hi_layer, uint8_t *cls, uint32_t *counts3);
/* The bucket layer of dimension 0 of every particle (the same arithmetic as the build's key; -1 for a
 * particle the build would kill): what an initial distribution of a global cloud over the slab owners,
 * or a per-layer histogram for a balanced split, needs. */
int abr_slab_layers(abr_handle h, const double *pos, size_t n, int32_t *layer_out);

/* neighbour_search_base::update_positions (src/NeighbourSearchBase.h:350-495)
 * for the ordered case + CellListOrdered::update_positions_impl
 * (src/CellListOrdered.h:190-259):
 *   pos   n x D AoS doubles, wrapped IN PLACE (enforce_domain_lambda :185-238)
 *   alive n bytes, cleared IN PLACE for particles outside a non-periodic
 *         domain or with non-finite coordinates
 *   order_out  n int32: on return order_out[0..n_alive) = m_alive_indices after
 *         sort_by_key, i.e. new[k] = old[order[k]]; stable within a bucket.
 *   n_alive_host: number of alive particles (host; stream is synchronised).
 * Bucket arrays stay owned by the handle (abr_celllist_get). */
int abr_celllist_build(abr_handle h, double *pos, uint8_t *alive, size_t n, int32_t *order_out,
                       size_t *n_alive_host);
/* Device views of m_bucket_indices (sorted keys, n_alive), m_bucket_begin,
 * m_bucket_end (n_buckets each) as left by the last build. */
int abr_celllist_get(abr_handle h, const uint32_t **bucket_indices, const uint32_t **bucket_begin,
                     const uint32_t **bucket_end, uint64_t *n_buckets_host);

/* Particles::reorder -> detail::gather of every column
 * (src/Particles.h:694-724, src/detail/Algorithms.h:718-745):
 * dst[c][k] = src[c][order[k]] for k < n_out, element size elem_bytes[c].
 * Out of place; the caller owns both buffers and swaps them (data.swap(other_data)).
 * src/dst/elem_bytes are HOST arrays of ncols entries holding device pointers. */
int abr_gather_columns(abr_handle h, int ncols, const void *const *src_host, void *const *dst_host,
                       const size_t *elem_bytes_host, const int32_t *order, size_t n_out);

/* Particles::update_positions (src/Particles.h:526-531) in one call:
 * abr_celllist_build + abr_gather_columns of every column (the position column,
 * recognised by src_host[c] == pos, must be among them) + update_iterators on
 * the gathered position column.  The reorder is enqueued behind the build on
 * the device (bounded by the device-side alive count), so the whole update
 * costs ONE host synchronisation — the one that returns n_alive_host.
 * dst columns must have room for n elements; their first n_alive are valid.
 * n_alive_host == NULL selects the ASYNCHRONOUS form: nothing is read back, the
 * call returns after enqueueing and the handle assumes every particle stays
 * alive (n_alive == n).  The caller must confirm that with abr_check_async
 * before trusting anything computed from this update. */
int abr_update_positions(abr_handle h, double *pos, uint8_t *alive, size_t n, int ncols,
                         const void *const *src_host, void *const *dst_host,
                         const size_t *elem_bytes_host, int32_t *order_out, size_t *n_alive_host);

/* neighbour_search_base::update_iterators (src/NeighbourSearchBase.h:504-512):
 * point the query at the reordered position column (n_alive x D). */
int abr_query_set_particles(abr_handle h, const double *pos_sorted, size_t n);

/* KernelSparse::evaluate (src/Kernels.h:720-751): y += K b, where rows i with
 * position row_pos[i] search the handle's (column) cell list within
 * radius (or radius_per_row[i], FRadius of src/detail/Kernels.h:336-357) and
 *   y[i*BR + p] += sum_q F(dx, a_i, b_j)(p,q) * b[j*BC + q].
 * rows_are_cols != 0 asserts row_pos is the handle's own sorted position array
 * (the create_sparse_operator(particles, particles, ...) case) and selects the
 * cell-tiled kernel; otherwise rows are arbitrary points (no search structure
 * needed on the row set, tests/rbf_interpolation.h:326).
 * n_pairs_host (optional): accepted (i,j,image) pair count (synchronises). */
int abr_sparse_matvec(abr_handle h, const double *row_pos, size_t n_rows, int rows_are_cols,
                      const abr_kernel_desc *kernel_host, double radius,
                      const double *radius_per_row, const double *b, double *y,
                      uint64_t *n_pairs_host);

/* KernelSparse::assemble to triplets (src/Kernels.h:653-685), as CSR:
 * row i owns entries row_ptr[i] .. row_ptr[i+1]-1, col_idx[k] is the column
 * PARTICLE j and values[k*BR*BC ..] its BR x BC block (row major) — i.e. the
 * triplets (i*BR+ii, j*BC+jj, block(ii,jj)) of the reference, in the reference's
 * own order within a row (the exact iterator walk).  Two-call protocol:
 *   col_idx == NULL : count only — fills row_ptr (n_rows+1, device), returns nnz;
 *   col_idx != NULL : also fills col_idx (capacity >= nnz) and, if non-NULL, values.
 * Block entries, not scalars: nnz * BR * BC scalar non-zeros (7 / 14 in
 * tests/operators.h:871, :938).  nnz must be < 2^32. */
int abr_sparse_assemble(abr_handle h, const double *row_pos, size_t n_rows, int rows_are_cols,
                        const abr_kernel_desc *kernel_host, double radius, const double *radius_per_row,
                        uint32_t *row_ptr, int32_t *col_idx, double *values, size_t capacity,
                        uint64_t *nnz_host);

/* KernelBase::coeff(i, j) over detail::sparse_kernel (src/Kernels.h:102-112,
 * src/detail/Kernels.h:336-367; reached from MatrixReplacement::coeff,
 * src/Operators.h:149-151) for m matrix entries at once: out[q] = K(ii[q], jj[q]).
 * Row particle ii/BR at row_pos, column particle jj/BC of the handle's particle
 * set; dx = correct_dx_for_periodicity(p_col - p_row) (src/Particles.h:480-494)
 * and the entry is non-zero only when dx.squaredNorm() < radius^2 — a STRICT
 * comparison, unlike the `<=` of the search that the product uses.  Needs the
 * domain (abr_domain_set) and abr_query_set_particles, not the cell list. */
int abr_sparse_coeff(abr_handle h, const double *row_pos, size_t n_rows,
                     const abr_kernel_desc *kernel_host, double radius, const double *radius_per_row,
                     const uint64_t *ii, const uint64_t *jj, size_t m, double *out);

/* get_neighbouring_buckets(query) -> bucket_pair_iterator (src/Search.h:498-764,
 * :857-860), the "fast cell-list search" of tests/neighbours.h:281-300: every pair
 * of touching buckets (i, j) once — first inside the domain (j after i in the box
 * around i), then across every periodic boundary with the quadrant q whose position
 * offset is q * (bmax - bmin) — in exactly the order the reference iterator yields
 * them.  bucket_i / bucket_j: collapsed bucket numbers; quadrant: D int8 per pair.
 * Two-call protocol: bucket_i == NULL counts only (n_pairs_host). */
int abr_bucket_pairs(abr_handle h, uint32_t *bucket_i, uint32_t *bucket_j, int8_t *quadrant,
                     size_t capacity, uint64_t *n_pairs_host);
/* The loops the reference documents for that iterator (tests/neighbours.h:892-951):
 * count[k] = number of particles within `radius` of particle k (|dx|^2 < r^2, STRICT,
 * k itself included), every unordered pair tested once through the bucket pairs.
 * Valid when the bucket side is >= radius, as the reference requires. */
int abr_fast_bucket_search_counts(abr_handle h, double radius, uint32_t *count);

/* Find-by-id (neighbour_search_base::init_id_map and the id-map update of
 * update_positions, src/NeighbourSearchBase.h:294-298, :440-486): builds
 * m_id_map_key (ids in ascending order) and m_id_map_value (position of that
 * particle) from the id column in its CURRENT order, i.e. call it after every
 * abr_update_positions, as the reference does inside update_positions.
 * ids are the reference's size_t (u64), assumed unique (src/Particles.h). */
int abr_id_map_build(abr_handle h, const uint64_t *ids, size_t n);
/* device views of m_id_map_key / m_id_map_value (n entries each) */
int abr_id_map_get(abr_handle h, const uint64_t **key, const uint64_t **value, size_t *n_host);
/* CellListOrderedQuery::find (src/CellListOrdered.h:379-388) for m ids at once:
 * index_out[q] = position of the particle with id query_ids[q], or n — the
 * reference's end pointer — when there is none. */
int abr_id_find(abr_handle h, const uint64_t *query_ids, size_t m, uint64_t *index_out);

/* Neighbour-set diagnostics used by the parity tests: per row the number of
 * accepted (j,image) pairs of euclidean_search (src/Search.h:839-845) and an
 * order-independent 64-bit hash of that set.  path 0 = cell-tiled kernel (needs
 * rows_are_cols), path 1 = per-row iterator walk. */
int abr_pair_stats(abr_handle h, const double *row_pos, size_t n_rows, int rows_are_cols,
                   double radius, const double *radius_per_row, int path, uint32_t *count,
                   uint64_t *hash);

/* The p-norm query iterator: distance_search<LNormNumber>, chebyshev_search
 * (lnorm -1), manhatten_search (1), euclidean_search (2) (src/Search.h:794-845)
 * from arbitrary query points: per point the number of (j, image) hits and the
 * order-independent hash of that set.  The device-side iterator itself is
 * abr::search_walk<D, LN> (include/aboria_b200/detail/grid.cuh) for custom kernels.
 * p >= 3 is not offered: the reference evaluates it through std::pow, whose
 * rounding a device pow() does not reproduce. */
int abr_distance_search_stats(abr_handle h, const double *query_pos, size_t n_queries, double radius,
                              const double *radius_per_query, int lnorm, uint32_t *count,
                              uint64_t *hash);

/* The same search in a scaled coordinate system: the Transform argument of
 * distance_search / euclidean_search (src/Search.h:794-845) with a ScaleTransform
 * (create_scale_transform, src/Transform.h:140-172): dx -> dx * scale per dimension
 * in the candidate test (src/Search.h:443) and in the bucket / domain distance tests
 * of the bucket iterator (src/NeighbourSearchBase.h:1790-1799, :1884-1892, :1950-1958).
 * scale_host: D factors (host).  E.g. tests/neighbours.h:553-561 searches radius 1.0
 * with scale 1/radius. */
int abr_distance_search_stats_scaled(abr_handle h, const double *query_pos, size_t n_queries,
                                     double radius, const double *radius_per_query, int lnorm,
                                     const double *scale_host, uint32_t *count, uint64_t *hash);

/* ... and with a LinearTransform (create_linear_transform<D>(functor), src/Transform.h:61-137,
 * :162-166) whose functor is linear, given as its D x D matrix (row major, host): the
 * point transform is the matrix product (zero entries skipped, so the reference tests'
 * SkewTransform `v[0] + 0.3 v[1]`, tests/neighbours.h:1262-1267, is reproduced operation for
 * operation); the box transform uses the constructor's "eigen vertex" (:82-99, :112-136). */
int abr_distance_search_stats_linear(abr_handle h, const double *query_pos, size_t n_queries,
                                     double radius, const double *radius_per_query, int lnorm,
                                     const double *matrix_host, uint32_t *count, uint64_t *hash);

/* Counters of the last abr_sparse_matvec / abr_pair_stats on the tiled path:
 * [0] rows re-done by the exact per-row walk (rounding-sensitive rows),
 * [1] particles whose bucket index overflowed in the last build (forces the
 *     exact walk), [2] kernels launched by the last matvec call,
 * [3] kernels launched by this handle since abr_create. */
int abr_last_counters(abr_handle h, uint64_t counters_host[4]);

/* Custom device functor (user TU compiled by nvcc, see
 * include/aboria_b200/device_kernel.cuh): `launch` is
 * abr::sparse_launcher<D, Functor>::launch and `functor_host` points at the
 * functor object (copied by value into the kernel arguments). */
struct abr_matvec_plan; /* opaque, filled by the library */
typedef int (*abr_launch_fn)(const struct abr_matvec_plan *plan, const void *functor_host);
int abr_sparse_matvec_custom(abr_handle h, const double *row_pos, size_t n_rows, int rows_are_cols,
                             abr_launch_fn launch, const void *functor_host, int block_rows,
                             int block_cols, double radius, const double *radius_per_row,
                             const double *b, double *y, uint64_t *n_pairs_host);

/* Measures the device's fp64 FMA throughput (TFLOP/s, FMA = 2 flop) with a
 * register-resident DFMA loop; the denominator of the matvec's fp64 roofline. */
int abr_probe_fp64_peak(abr_handle h, double *tflops_host);

/* Thin device-memory helpers so that C/C++ hosts need no CUDA headers. */
int abr_malloc(abr_handle h, void **ptr_out, size_t bytes);
int abr_free(abr_handle h, void *ptr);
int abr_memcpy_h2d(abr_handle h, void *dst, const void *src_host, size_t bytes);
int abr_memcpy_d2h(abr_handle h, void *dst_host, const void *src, size_t bytes);
int abr_memset(abr_handle h, void *dst, int value, size_t bytes);
int abr_host_alloc_pinned(void **ptr_out, size_t bytes);
int abr_host_free_pinned(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* ABR_H_ */
