// Internal declarations shared by the translation units of libabr.so.
#ifndef ABR_INTERNAL_H_
#define ABR_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "abr.h"
#include "aboria_b200/detail/grid.cuh"

namespace abr {

// grow-only device scratch buffer
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

// device-resident scalars written by the build / matvec kernels
struct DevScalars {
  uint32_t n_alive;        // first position holding the dead key
  uint32_t n_incell;       // first position with key >= ncells
  uint32_t n_aliased;      // alive particles whose bucket index vector overflowed (v[d] >= size[d])
  uint32_t work_counter;   // dynamic tile scheduler of the tiled matvec
  uint32_t danger_count;   // rows handed to the exact per-row walk
  uint32_t n_outside;      // particles whose bucket layer is not in this rank's window
  uint32_t n_unsorted;     // adopt_sorted: key inversions found
  uint32_t pad[1];
  unsigned long long pair_count;
  unsigned long long heavy_state; // tiled product: heavy-bucket items << 40 | their row batches
  uint32_t heavy_work;            // scheduler of the heavy launch
  uint32_t max_bucket;            // largest bucket occupancy seen by the build (k_boundaries)
};

struct Handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int sm_count = 148;

  // --- neighbour_search_base / CellListOrdered host state -----------------
  bool domain_set = false;
  int D = 3;
  double bmin[MAXD], bmax[MAXD];
  bool periodic[MAXD];
  double n_leaf = 10.0;
  uint32_t size[MAXD] = {1, 1, 1};
  double side[MAXD], inv_side[MAXD];
  size_t size_calculated_with_n = (size_t)-1;
  size_t n_alive_last = 0; // m_alive_indices.size()
  bool grid_forced = false;
  bool windowed = false; // slab window along dim 0 (abr_domain_set_window)
  int win_lo = 0, win_n = 0, own_lo = 0, own_n = 0;

  // --- device state ---------------------------------------------------------
  DevBuf keys[2], idx[2], tile_hist, scan_tmp, gap_list;
  DevBuf idx2, tmp_cols, tile_tab, seg_hist; // two-level build
  DevBuf cs_ko, cs_scratch, cs_scratch2, cs_bins; // counting-sort build (abr_build2.cu)
  size_t counting_min_n = (size_t)-1;             // counting-sort build (abr_build2.cu) from this many particles; off by default: measured slower
                                                  // than the two-level radix build on B200 (3.4 vs 2.7 ms at 32 M, profiles/r2f_counting_build.txt)
  size_t two_level_min_n = (size_t)1 << 20;  // use the two-level build from this many particles (abr_set_option)
  bool gather_slots = true;                  // two-level build: final reorder with the loads of every column in flight at once
  int stage_threads = 512;                   // threads per CTA of the staged record move (512 or 1024)
  bool record_aos = false;                   // staged record move writes one record per particle (RecLayout); measured slower, off
  bool bounds_one_sweep = true;              // bucket ranges in one sweep over the sorted keys (k_boundaries_fill) instead of fills + boundaries + suffix-min scan
  bool skip_alive_move = true;               // two-level build: the reordered alive column is a run of ones, not a gather
  bool stage_records = true;                 // two-level build: bulk-copy staged record move when the tile windows fit in shared memory
  DevBuf bucket_begin, bucket_end;
  DevBuf danger_list;
  DevBuf scan_tmp2, pair_i, pair_j, pair_q; // bucket-pair traversal (abr_pairs.cu)
  DevBuf idm_k[2], idm_i[2], idm_max, id_map_key, id_map_value; // id map (m_id_map_key / m_id_map_value) + sort scratch
  size_t id_map_n = 0;
  Handle *rows_h = nullptr; // internal handle: row points of a rows != columns product bucketed into this grid (abr_matvec.cu)
  size_t xrows_min_n = 1024;  // from this many rows on (abr_set_option("xrows_min_n")); fewer rows: one thread per row
  uint32_t max_bucket = 0; // largest bucket occupancy of the last verified build (hint for the product's heavy-bucket split)
  DevBuf heavy_list;     // tiled product: (bucket, first batch) of the buckets split into row batches
  DevBuf ytmp, row_bits; // symmetric product: accumulation scratch, exact-walk flags
  bool symmetric = false; // abr_set_option("symmetric"): evaluate each unordered pair once for functors that declare SYMMETRY
  DevBuf posb; // packed (x, y, z, b) records of the column particles for the tiled product
  DevScalars *d_scalars = nullptr;
  DevScalars *h_scalars = nullptr; // pinned read-back mirror, written by k_publish_scalars
  const uint32_t *sorted_keys = nullptr;
  uint64_t ncells = 0;
  bool built = false;
  size_t async_pending_n = 0; // particle count of an unverified asynchronous update (0: none)
  uint32_t n_aliased = 0;

  // query binding
  const double *pos_sorted = nullptr;
  size_t n_sorted = 0;

  uint64_t counters[4] = {0, 0, 0, 0};
  uint64_t launches = 0; // kernels launched since abr_create
  int matvec_variant = 0; // tiled product: 0 = gathers from L2 (tiled_kernel), 1 = bulk-copy staging (staged_kernel); abr_set_option("matvec_variant")
  bool phased_gather = false; // L2-windowed reorder for column sets larger than L2 (measured SLOWER on B200: 5.8 vs 3.0 ms build; ABR_PHASED_GATHER=1 enables)
  uint64_t gather_src_n = 0; // source length of the gather in flight (0: unknown, use n_out)

  Grid grid() const;
};

int set_error(Handle *h, int code, const std::string &msg);
int check_cuda(Handle *h, cudaError_t e, const char *what);

constexpr int GP_MAXC = 8; // columns a reorder kernel takes by value

// abr_build.cu
struct ReorderSpec {
  int ncols;
  const void *const *src;
  void *const *dst;
  const size_t *elem_bytes;
};
int build_celllist(Handle *h, double *pos, uint8_t *alive, size_t n, int32_t *order_out,
                   size_t *n_alive_host, const ReorderSpec *reorder, bool presorted = false);
// abr_build2.cu
bool counting_build_applicable(const Handle *h, size_t n, int bits, const ReorderSpec *reorder, const uint8_t *alive);
int build_counting(Handle *h, double *pos, uint8_t *alive, uint32_t n, const Grid &g, int bits, const ReorderSpec *reorder, int32_t *order_out);
int gather_columns(Handle *h, int ncols, const void *const *src, void *const *dst,
                   const size_t *elem_bytes, const int32_t *order, size_t n_out, const uint32_t *n_dev);

int build_id_map(Handle *h, const uint64_t *ids, size_t n);
int find_ids(Handle *h, const uint64_t *query, size_t m, uint64_t *index_out);
void fill_u32(Handle *h, uint32_t *p, uint32_t v, uint64_t n); // kernel fill (no memset node)
void publish_scalars(Handle *h);                               // d_scalars -> pinned h_scalars by a kernel store

// abr_matvec.cu
struct MatvecCall {
  const double *row_pos;
  size_t n_rows;
  int rows_are_cols;
  double radius;
  const double *radius_per_row;
  const double *b;
  double *y;
  // stats mode
  uint32_t *count;
  uint64_t *hash;
  int force_path; // -1 auto, 0 tiled, 1 walk
};
int run_builtin_matvec(Handle *h, const MatvecCall &c, const abr_kernel_desc *k);
int run_pair_stats(Handle *h, const MatvecCall &c);
int run_norm_stats(Handle *h, const MatvecCall &c, int lnorm, int transform_kind = 0, const double *t_host = nullptr);
int run_assemble(Handle *h, const MatvecCall &c, const abr_kernel_desc *k, uint32_t *row_ptr, int32_t *col_idx, double *values,
                 size_t capacity, uint64_t *nnz_host);
int run_coeff(Handle *h, const MatvecCall &c, const abr_kernel_desc *k, const uint64_t *ii, const uint64_t *jj, size_t m, double *out);
// abr_pairs.cu
int run_bucket_pairs(Handle *h, uint32_t *bucket_i, uint32_t *bucket_j, int8_t *quadrant, uint64_t capacity, uint64_t *n_host);
int run_fast_bucket_search_counts(Handle *h, double radius, uint32_t *count);
int run_custom_matvec(Handle *h, const MatvecCall &c, abr_launch_fn launch, const void *functor,
                      int BR, int BC);

} // namespace abr

#define ABR_CUDA(h, expr)                                        \
  do {                                                           \
    cudaError_t e__ = (expr);                                    \
    if (e__ != cudaSuccess) return abr::check_cuda(h, e__, #expr); \
  } while (0)

#endif
