// C-ABI entry points of libabr.so (declared in include/abr.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "abr_internal.h"

namespace abr {

int set_error(Handle *h, int code, const std::string &msg) {
  if (h) h->err = msg;
  return code;
}
int check_cuda(Handle *h, cudaError_t e, const char *what) {
  if (e == cudaSuccess) return ABR_OK;
  if (h) h->err = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return ABR_ERR_CUDA;
}
void host_set_domain_impl(Handle *h, size_t n);
int probe_fp64_peak(Handle *h, double *tflops);

} // namespace abr

namespace abr {
__global__ void __launch_bounds__(256) k_sum_u32(const uint32_t *__restrict__ v, uint64_t n, unsigned long long *__restrict__ out) {
  unsigned long long acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// accepted (i, j, image) pairs of a product: a separate stats pass (diagnostic, not part of
// the product path), reduced on the device
static int count_pairs(Handle *h, const double *row_pos, size_t n_rows, int rows_are_cols, double radius, const double *radius_per_row,
                       uint64_t *n_pairs_host) {
  *n_pairs_host = 0;
  if (n_rows == 0) return ABR_OK;
  uint32_t *cnt = nullptr;
  ABR_CUDA(h, cudaMalloc(&cnt, (n_rows + 3) * sizeof(uint32_t)));
  unsigned long long *total = reinterpret_cast<unsigned long long *>(cnt + ((n_rows + 1) & ~(size_t)1));
  cudaMemsetAsync(total, 0, sizeof(unsigned long long), h->stream);
  MatvecCall s{row_pos, n_rows, rows_are_cols, radius, radius_per_row, nullptr, nullptr, cnt, nullptr, -1};
  int rc = run_pair_stats(h, s);
  if (!rc) {
    const unsigned grid = (unsigned)std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)h->sm_count * 8);
    k_sum_u32<<<grid, 256, 0, h->stream>>>(cnt, n_rows, total);
    h->launches += 1;
    unsigned long long t = 0;
    cudaMemcpyAsync(&t, total, sizeof(t), cudaMemcpyDeviceToHost, h->stream);
    rc = check_cuda(h, cudaStreamSynchronize(h->stream), "pair count");
    *n_pairs_host = t;
  }
  cudaFree(cnt);
  return rc;
}
} // namespace abr

using abr::Handle;

static thread_local std::string g_create_error;

extern "C" {

const char *abr_version(void) { return "aboria_b200 0.1 (sm_100a)"; }

int abr_create(abr_handle *out, int device, void *stream) {
  if (!out) return ABR_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("abr_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    cudaGetLastError();
    return ABR_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "abr_create: bad device index";
    return ABR_ERR_INVALID;
  }
  Handle *h = new (std::nothrow) Handle();
  if (!h) return ABR_ERR_INVALID;
  h->device = device;
  h->stream = static_cast<cudaStream_t>(stream);
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaMalloc(&h->d_scalars, sizeof(abr::DevScalars))) != cudaSuccess ||
      (e = cudaMallocHost(&h->h_scalars, 3 * sizeof(abr::DevScalars))) != cudaSuccess) {
    g_create_error = std::string("abr_create: ") + cudaGetErrorString(e);
    delete h;
    return ABR_ERR_CUDA;
  }
  cudaMemset(h->d_scalars, 0, sizeof(abr::DevScalars));
  if (const char *e = getenv("ABR_PHASED_GATHER")) h->phased_gather = (e[0] != '0');
  if (const char *e = getenv("ABR_MATVEC_VARIANT")) h->matvec_variant = atoi(e);
  if (const char *e = getenv("ABR_SYMMETRIC")) h->symmetric = (e[0] != '0');
  if (const char *e = getenv("ABR_STAGE_RECORDS")) h->stage_records = (e[0] != '0');
  if (const char *e = getenv("ABR_GATHER_SLOTS")) h->gather_slots = (e[0] != '0');
  if (const char *e = getenv("ABR_RECORD_AOS")) h->record_aos = (e[0] != '0');
  if (const char *e = getenv("ABR_SKIP_ALIVE_MOVE")) h->skip_alive_move = (e[0] != '0');
  if (const char *e = getenv("ABR_BOUNDS_ONE_SWEEP")) h->bounds_one_sweep = (e[0] != '0');
  if (const char *e = getenv("ABR_STAGE_THREADS")) h->stage_threads = atoi(e) == 512 ? 512 : 1024;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h->sm_count = sms;
  for (int d = 0; d < abr::MAXD; ++d) {
    h->bmin[d] = 0;
    h->bmax[d] = 1;
    h->periodic[d] = false;
    h->side[d] = 1;
    h->inv_side[d] = 1;
  }
  *out = reinterpret_cast<abr_handle>(h);
  return ABR_OK;
}

int abr_destroy(abr_handle hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->rows_h) { // internal handle of the rows != columns product
    abr_destroy(reinterpret_cast<abr_handle>(h->rows_h));
    h->rows_h = nullptr;
  }
  h->heavy_list.release();
  h->cs_ko.release();
  h->cs_scratch.release();
  h->cs_scratch2.release();
  h->cs_bins.release();
  for (int i = 0; i < 2; ++i) {
    h->keys[i].release();
    h->idx[i].release();
  }
  h->tile_hist.release();
  h->scan_tmp.release();
  h->gap_list.release();
  h->idx2.release();
  h->tmp_cols.release();
  h->tile_tab.release();
  h->seg_hist.release();
  h->bucket_begin.release();
  h->bucket_end.release();
  h->danger_list.release();
  h->ytmp.release();
  h->row_bits.release();
  h->posb.release();
  for (int i = 0; i < 2; ++i) {
    h->idm_k[i].release();
    h->idm_i[i].release();
  }
  h->idm_max.release();
  h->scan_tmp2.release();
  h->pair_i.release();
  h->pair_j.release();
  h->pair_q.release();
  h->id_map_key.release();
  h->id_map_value.release();
  if (h->d_scalars) cudaFree(h->d_scalars);
  if (h->h_scalars) cudaFreeHost(h->h_scalars);
  delete h;
  return ABR_OK;
}

int abr_set_stream(abr_handle hh, void *stream) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  h->stream = static_cast<cudaStream_t>(stream);
  return ABR_OK;
}

int abr_set_option(abr_handle hh, const char *name, double value) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !name) return ABR_ERR_INVALID;
  const std::string k(name);
  if (k == "stage_records") {
    h->stage_records = value != 0;
    return ABR_OK;
  }
  if (k == "bounds_one_sweep") {
    h->bounds_one_sweep = value != 0;
    return ABR_OK;
  }
  if (k == "skip_alive_move") {
    h->skip_alive_move = value != 0;
    return ABR_OK;
  }
  if (k == "record_aos") {
    h->record_aos = value != 0;
    return ABR_OK;
  }
  if (k == "gather_slots") {
    h->gather_slots = value != 0;
    return ABR_OK;
  }
  if (k == "two_level_min_n") {
    h->two_level_min_n = value < 0 ? 0 : (size_t)value;
  } else if (k == "counting_min_n") {
    h->counting_min_n = value < 0 ? 0 : (value > 1e18 ? (size_t)-1 : (size_t)value);
  } else if (k == "phased_gather") {
    h->phased_gather = value != 0;
  } else if (k == "xrows_min_n") {
    h->xrows_min_n = value < 0 ? 0 : (value > 1e18 ? (size_t)-1 : (size_t)value);
  } else if (k == "symmetric") {
    h->symmetric = value != 0;
  } else if (k == "matvec_variant") {
    h->matvec_variant = (int)value;
  } else {
    return abr::set_error(h, ABR_ERR_INVALID, "set_option: unknown option " + k);
  }
  return ABR_OK;
}

int abr_check_async(abr_handle hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->async_pending_n == 0) return ABR_OK;
  const size_t n = h->async_pending_n;
  h->async_pending_n = 0;
  h->max_bucket = h->h_scalars->max_bucket;
  if (h->h_scalars->n_alive != n)
    return abr::set_error(h, ABR_ERR_STATE, "asynchronous update_positions: particles died; results of this update are invalid, redo it synchronously");
  if (h->h_scalars->n_aliased != 0 || h->h_scalars->n_outside != 0)
    return abr::set_error(h, ABR_ERR_STATE, "asynchronous update_positions: a bucket index overflowed; redo the update synchronously");
  return ABR_OK;
}

int abr_synchronize(abr_handle hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaStreamSynchronize(h->stream));
  return ABR_OK;
}

const char *abr_last_error_string(abr_handle hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return g_create_error.c_str();
  return h->err.c_str();
}

static int set_domain_common(Handle *h, int D, const double *bmin, const double *bmax, const uint8_t *periodic) {
  if (D < 1 || D > abr::MAXD) return abr::set_error(h, ABR_ERR_UNSUPPORTED, "domain: D must be 1, 2 or 3");
  if (!bmin || !bmax || !periodic) return abr::set_error(h, ABR_ERR_INVALID, "domain: null pointer");
  for (int d = 0; d < D; ++d)
    if (!(bmax[d] > bmin[d]) || !std::isfinite(bmin[d]) || !std::isfinite(bmax[d]))
      return abr::set_error(h, ABR_ERR_INVALID, "domain: need finite bmin < bmax");
  if (h->domain_set && h->D != D) { // a different particle type: start over
    h->size_calculated_with_n = (size_t)-1;
    h->n_alive_last = 0;
  }
  h->D = D;
  for (int d = 0; d < D; ++d) {
    h->bmin[d] = bmin[d];
    h->bmax[d] = bmax[d];
    h->periodic[d] = periodic[d] != 0;
  }
  h->domain_set = true;
  h->built = false;
  return ABR_OK;
}

int abr_domain_set(abr_handle hh, int D, const double *bmin, const double *bmax, const uint8_t *periodic, double n_leaf) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!(n_leaf > 0)) return abr::set_error(h, ABR_ERR_INVALID, "domain: n_particles_in_leaf must be > 0");
  int rc = set_domain_common(h, D, bmin, bmax, periodic);
  if (rc) return rc;
  h->n_leaf = n_leaf;
  h->grid_forced = false;
  h->windowed = false;
  // set_domain -> set_domain_impl with the current m_alive_indices.size()
  // (src/NeighbourSearchBase.h:252-268, src/CellListOrdered.h:132-139)
  abr::host_set_domain_impl(h, h->n_alive_last);
  return ABR_OK;
}

int abr_domain_force_grid(abr_handle hh, int D, const double *bmin, const double *bmax, const uint8_t *periodic, const uint32_t *size) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!size) return abr::set_error(h, ABR_ERR_INVALID, "force_grid: null size");
  int rc = set_domain_common(h, D, bmin, bmax, periodic);
  if (rc) return rc;
  for (int d = 0; d < D; ++d) {
    if (size[d] == 0) return abr::set_error(h, ABR_ERR_INVALID, "force_grid: zero size");
    h->size[d] = size[d];
    h->side[d] = (h->bmax[d] - h->bmin[d]) / h->size[d];
    h->inv_side[d] = 1.0 / h->side[d];
  }
  h->grid_forced = true;
  h->windowed = false;
  return ABR_OK;
}

int abr_domain_set_window(abr_handle hh, int win_lo, int win_n, int own_lo, int own_n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!h->domain_set || !h->grid_forced) return abr::set_error(h, ABR_ERR_STATE, "set_window: call abr_domain_force_grid (global grid) first");
  if (h->D < 2) return abr::set_error(h, ABR_ERR_UNSUPPORTED, "set_window: slabs need D >= 2");
  const int S0 = (int)h->size[0];
  if (win_n < 1 || win_n > S0 || own_lo < 0 || own_n < 1 || own_lo + own_n > win_n || win_lo < -S0 || win_lo >= S0)
    return abr::set_error(h, ABR_ERR_INVALID, "set_window: bad window");
  if (!h->periodic[0] && (win_lo < 0 || win_lo + win_n > S0))
    return abr::set_error(h, ABR_ERR_INVALID, "set_window: window leaves a non-periodic domain");
  h->windowed = true;
  h->win_lo = win_lo;
  h->win_n = win_n;
  h->own_lo = own_lo;
  h->own_n = own_n;
  h->built = false;
  return ABR_OK;
}

int abr_grid_for(int D, const double *bmin, const double *bmax, double n_leaf, size_t n, uint32_t *size, double *side) {
  // CellListOrdered::set_domain_impl (src/CellListOrdered.h:140-157) as a pure function
  if (D < 1 || D > abr::MAXD || !bmin || !bmax || !size || !side || !(n_leaf > 0)) return ABR_ERR_INVALID;
  if (n_leaf > n) {
    for (int d = 0; d < D; ++d) size[d] = 1;
  } else {
    double total_volume = 1.0;
    for (int d = 0; d < D; ++d) total_volume *= (bmax[d] - bmin[d]);
    const double box_volume = n_leaf / double(n) * total_volume;
    const double box_side_length = std::pow(box_volume, 1.0 / D);
    for (int d = 0; d < D; ++d) {
      size[d] = static_cast<unsigned int>(std::floor((bmax[d] - bmin[d]) / box_side_length));
      if (size[d] == 0) size[d] = 1;
    }
  }
  for (int d = 0; d < D; ++d) side[d] = (bmax[d] - bmin[d]) / size[d];
  return ABR_OK;
}

int abr_celllist_adopt_sorted(abr_handle hh, double *pos_sorted, uint8_t *alive, size_t n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (n > 0 && (!pos_sorted || !alive)) return abr::set_error(h, ABR_ERR_INVALID, "adopt_sorted: null pointer");
  ABR_CUDA(h, cudaSetDevice(h->device));
  size_t n_alive = 0;
  int rc = abr::build_celllist(h, pos_sorted, alive, n, nullptr, &n_alive, nullptr, true);
  if (rc) return rc;
  if (n_alive != n) return abr::set_error(h, ABR_ERR_INVALID, "adopt_sorted: dead or out-of-domain particles in a sorted set");
  h->pos_sorted = pos_sorted;
  h->n_sorted = n;
  return ABR_OK;
}

int abr_domain_get(abr_handle hh, uint32_t *size, double *side, uint64_t *n_buckets) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!h->domain_set) return abr::set_error(h, ABR_ERR_STATE, "domain_get: domain has not been set");
  uint64_t prod = 1;
  for (int d = 0; d < h->D; ++d) {
    if (size) size[d] = h->size[d];
    if (side) side[d] = h->side[d];
    prod *= h->size[d];
  }
  if (n_buckets) *n_buckets = prod;
  return ABR_OK;
}

int abr_celllist_build(abr_handle hh, double *pos, uint8_t *alive, size_t n, int32_t *order_out, size_t *n_alive_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (n > 0 && (!pos || !alive || !order_out)) return abr::set_error(h, ABR_ERR_INVALID, "build: null pointer");
  ABR_CUDA(h, cudaSetDevice(h->device));
  return abr::build_celllist(h, pos, alive, n, order_out, n_alive_host, nullptr);
}

int abr_celllist_get(abr_handle hh, const uint32_t **bucket_indices, const uint32_t **bucket_begin, const uint32_t **bucket_end, uint64_t *n_buckets) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!h->built) return abr::set_error(h, ABR_ERR_STATE, "celllist_get: no build");
  if (bucket_indices) *bucket_indices = h->sorted_keys;
  if (bucket_begin) *bucket_begin = h->bucket_begin.as<uint32_t>();
  if (bucket_end) *bucket_end = h->bucket_end.as<uint32_t>();
  if (n_buckets) *n_buckets = h->ncells;
  return ABR_OK;
}

int abr_gather_columns(abr_handle hh, int ncols, const void *const *src, void *const *dst, const size_t *elem_bytes, const int32_t *order, size_t n_out) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (ncols < 0 || (ncols > 0 && (!src || !dst || !elem_bytes))) return abr::set_error(h, ABR_ERR_INVALID, "gather: null pointer");
  if (n_out > 0 && !order) return abr::set_error(h, ABR_ERR_INVALID, "gather: null order");
  ABR_CUDA(h, cudaSetDevice(h->device));
  return abr::gather_columns(h, ncols, src, dst, elem_bytes, order, n_out, nullptr);
}

int abr_update_positions(abr_handle hh, double *pos, uint8_t *alive, size_t n, int ncols, const void *const *src, void *const *dst,
                         const size_t *elem_bytes, int32_t *order_out, size_t *n_alive_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (n > 0 && (!pos || !alive || !order_out)) return abr::set_error(h, ABR_ERR_INVALID, "update_positions: null pointer");
  if (ncols <= 0 || !src || !dst || !elem_bytes) return abr::set_error(h, ABR_ERR_INVALID, "update_positions: no columns");
  int pos_col = -1;
  for (int c = 0; c < ncols; ++c)
    if (src[c] == (const void *)pos) pos_col = c;
  if (pos_col < 0 && n > 0) return abr::set_error(h, ABR_ERR_INVALID, "update_positions: the position column must be among the columns");
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::ReorderSpec spec{ncols, src, dst, elem_bytes};
  size_t n_alive = n;
  int rc = abr::build_celllist(h, pos, alive, n, order_out, n_alive_host ? &n_alive : nullptr, n > 0 ? &spec : nullptr);
  if (rc) return rc;
  if (n_alive_host) *n_alive_host = n_alive;
  // update_iterators: the query now reads the reordered position column
  h->pos_sorted = n > 0 ? static_cast<const double *>(dst[pos_col]) : nullptr;
  h->n_sorted = n_alive;
  return ABR_OK;
}

int abr_query_set_particles(abr_handle hh, const double *pos_sorted, size_t n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (n > 0 && !pos_sorted) return abr::set_error(h, ABR_ERR_INVALID, "query: null positions");
  if (n != h->n_alive_last) return abr::set_error(h, ABR_ERR_INVALID, "query: n differs from the alive count of the last build");
  h->pos_sorted = pos_sorted;
  h->n_sorted = n;
  return ABR_OK;
}

int abr_sparse_matvec(abr_handle hh, const double *row_pos, size_t n_rows, int rows_are_cols, const abr_kernel_desc *k, double radius,
                      const double *radius_per_row, const double *b, double *y, uint64_t *n_pairs_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, rows_are_cols, radius, radius_per_row, b, y, nullptr, nullptr, -1};
  int rc = abr::run_builtin_matvec(h, c, k);
  if (rc) return rc;
  if (n_pairs_host) return abr::count_pairs(h, row_pos, n_rows, rows_are_cols, radius, radius_per_row, n_pairs_host);
  return ABR_OK;
}

int abr_pair_stats(abr_handle hh, const double *row_pos, size_t n_rows, int rows_are_cols, double radius, const double *radius_per_row,
                   int path, uint32_t *count, uint64_t *hash) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (path != 0 && path != 1 && path != -1) return abr::set_error(h, ABR_ERR_INVALID, "pair_stats: path must be -1, 0 or 1");
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, rows_are_cols, radius, radius_per_row, nullptr, nullptr, count, hash, path};
  return abr::run_pair_stats(h, c);
}

int abr_sparse_assemble(abr_handle hh, const double *row_pos, size_t n_rows, int rows_are_cols, const abr_kernel_desc *k, double radius,
                        const double *radius_per_row, uint32_t *row_ptr, int32_t *col_idx, double *values, size_t capacity,
                        uint64_t *nnz_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (n_rows > 0 && !row_pos) return abr::set_error(h, ABR_ERR_INVALID, "assemble: null row positions");
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, rows_are_cols, radius, radius_per_row, nullptr, nullptr, nullptr, nullptr, -1};
  return abr::run_assemble(h, c, k, row_ptr, col_idx, values, capacity, nnz_host);
}

int abr_sparse_coeff(abr_handle hh, const double *row_pos, size_t n_rows, const abr_kernel_desc *k, double radius, const double *radius_per_row,
                     const uint64_t *ii, const uint64_t *jj, size_t m, double *out) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, 0, radius, radius_per_row, nullptr, nullptr, nullptr, nullptr, 1};
  return abr::run_coeff(h, c, k, ii, jj, m, out);
}

int abr_bucket_pairs(abr_handle hh, uint32_t *bucket_i, uint32_t *bucket_j, int8_t *quadrant, size_t capacity, uint64_t *n_pairs_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  return abr::run_bucket_pairs(h, bucket_i, bucket_j, quadrant, capacity, n_pairs_host);
}

int abr_fast_bucket_search_counts(abr_handle hh, double radius, uint32_t *count) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  return abr::run_fast_bucket_search_counts(h, radius, count);
}

int abr_id_map_build(abr_handle hh, const uint64_t *ids, size_t n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  return abr::build_id_map(h, ids, n);
}

int abr_id_map_get(abr_handle hh, const uint64_t **key, const uint64_t **value, size_t *n_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (key) *key = h->id_map_n ? h->id_map_key.as<uint64_t>() : nullptr;
  if (value) *value = h->id_map_n ? h->id_map_value.as<uint64_t>() : nullptr;
  if (n_host) *n_host = h->id_map_n;
  return ABR_OK;
}

int abr_id_find(abr_handle hh, const uint64_t *query_ids, size_t m, uint64_t *index_out) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  return abr::find_ids(h, query_ids, m, index_out);
}

int abr_distance_search_stats(abr_handle hh, const double *row_pos, size_t n_rows, double radius, const double *radius_per_row, int lnorm,
                              uint32_t *count, uint64_t *hash) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, 0, radius, radius_per_row, nullptr, nullptr, count, hash, 1};
  return abr::run_norm_stats(h, c, lnorm);
}

int abr_distance_search_stats_scaled(abr_handle hh, const double *row_pos, size_t n_rows, double radius, const double *radius_per_row, int lnorm,
                                     const double *scale_host, uint32_t *count, uint64_t *hash) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, 0, radius, radius_per_row, nullptr, nullptr, count, hash, 1};
  return abr::run_norm_stats(h, c, lnorm, scale_host ? 1 : 0, scale_host);
}

int abr_distance_search_stats_linear(abr_handle hh, const double *row_pos, size_t n_rows, double radius, const double *radius_per_row, int lnorm,
                                     const double *matrix_host, uint32_t *count, uint64_t *hash) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, 0, radius, radius_per_row, nullptr, nullptr, count, hash, 1};
  return abr::run_norm_stats(h, c, lnorm, 2, matrix_host);
}

int abr_last_counters(abr_handle hh, uint64_t counters[4]) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !counters) return ABR_ERR_INVALID;
  abr::publish_scalars(h);
  ABR_CUDA(h, cudaStreamSynchronize(h->stream));
  counters[0] = h->h_scalars->danger_count;
  counters[1] = h->n_aliased;
  counters[2] = h->counters[2];
  counters[3] = h->launches;
  return ABR_OK;
}

int abr_sparse_matvec_custom(abr_handle hh, const double *row_pos, size_t n_rows, int rows_are_cols, abr_launch_fn launch,
                             const void *functor, int BR, int BC, double radius, const double *radius_per_row, const double *b,
                             double *y, uint64_t *n_pairs_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::MatvecCall c{row_pos, n_rows, rows_are_cols, radius, radius_per_row, b, y, nullptr, nullptr, -1};
  int rc = abr::run_custom_matvec(h, c, launch, functor, BR, BC);
  if (rc) return rc;
  if (n_pairs_host) return abr::count_pairs(h, row_pos, n_rows, rows_are_cols, radius, radius_per_row, n_pairs_host);
  return ABR_OK;
}

int abr_probe_fp64_peak(abr_handle hh, double *tflops_host) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !tflops_host) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  return abr::probe_fp64_peak(h, tflops_host);
}

int abr_malloc(abr_handle hh, void **ptr, size_t bytes) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !ptr) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaSetDevice(h->device));
  ABR_CUDA(h, cudaMalloc(ptr, bytes ? bytes : 1));
  return ABR_OK;
}
int abr_free(abr_handle hh, void *ptr) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaStreamSynchronize(h->stream));
  ABR_CUDA(h, cudaFree(ptr));
  return ABR_OK;
}
int abr_memcpy_h2d(abr_handle hh, void *dst, const void *src, size_t bytes) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return ABR_OK;
}
int abr_memcpy_d2h(abr_handle hh, void *dst, const void *src, size_t bytes) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  ABR_CUDA(h, cudaStreamSynchronize(h->stream));
  return ABR_OK;
}
int abr_memset(abr_handle hh, void *dst, int value, size_t bytes) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  ABR_CUDA(h, cudaMemsetAsync(dst, value, bytes, h->stream));
  return ABR_OK;
}
int abr_host_alloc_pinned(void **ptr, size_t bytes) {
  if (!ptr) return ABR_ERR_INVALID;
  return cudaMallocHost(ptr, bytes ? bytes : 1) == cudaSuccess ? ABR_OK : ABR_ERR_CUDA;
}
int abr_host_free_pinned(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? ABR_OK : ABR_ERR_CUDA; }

} // extern "C"
