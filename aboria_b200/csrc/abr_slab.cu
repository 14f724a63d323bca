// Multi-GPU slab support (no counterpart in the single-process reference):
// the two device-side steps a rank needs besides its ordinary build.
//
//   abr_celllist_patch_ghosts  after abr_update_positions in the ghost-padded window (own
//       layers in the middle, ghost layers still empty) and the halo exchange: the local array
//       is [ghost_lo | owned | ghost_hi]; owned bucket ranges shift by the lower ghost count,
//       ghost buckets take the SENDER's ranges (contiguous slices of its m_bucket_begin/end for
//       the layers it sent), rebased.  One pass over the window's bucket arrays — the key /
//       boundary passes over the whole local set of abr_celllist_adopt_sorted are gone.
//   abr_slab_classify  migration: which of this rank's (unsorted, moved) particles now belong
//       to the lower / upper neighbour's slab (global grid arithmetic of
//       src/detail/SpatialUtil.h:118-131 in dimension 0, periodic wrap of
//       src/NeighbourSearchBase.h:208-237).
#include <algorithm>

#include "abr_internal.h"

namespace abr {

__global__ void __launch_bounds__(256)
k_patch_ghosts(uint32_t *__restrict__ bb, uint32_t *__restrict__ be, uint32_t n_lo_buckets, uint32_t n_own_buckets, uint32_t n_hi_buckets,
               const uint32_t *__restrict__ bb_lo, const uint32_t *__restrict__ be_lo, const uint32_t *__restrict__ bb_hi,
               const uint32_t *__restrict__ be_hi, uint32_t n_lo, uint32_t n_own) {
  const uint32_t total = n_lo_buckets + n_own_buckets + n_hi_buckets;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
    if (c < n_lo_buckets) {
      const uint32_t base = bb_lo[0];
      bb[c] = bb_lo[c] - base;
      be[c] = be_lo[c] - base;
    } else if (c < n_lo_buckets + n_own_buckets) {
      bb[c] += n_lo;
      be[c] += n_lo;
    } else {
      const uint32_t k = c - n_lo_buckets - n_own_buckets;
      const uint32_t base = bb_hi[0];
      bb[c] = bb_hi[k] - base + n_lo + n_own;
      be[c] = be_hi[k] - base + n_lo + n_own;
    }
  }
}

// cls: 0 stays, 1 goes to the lower neighbour, 2 to the upper one; counts[cls] += 1
__global__ void __launch_bounds__(256)
k_slab_classify(const double *__restrict__ pos, uint32_t n, int D, double bmin0, double bmax0, double inv_side0, int S0, int periodic0, int lo_layer,
                int hi_layer, uint8_t *__restrict__ cls, uint32_t *__restrict__ counts, int32_t *__restrict__ layer_out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = 0;
  if (layer_out) { // layers only (abr_slab_layers)
    if (p < n) {
      double x = pos[(size_t)p * D];
      int layer = -1;
      if (isfinite(x)) {
        if (periodic0) {
          int guard = 0;
          while (x < bmin0 && ++guard < (1 << 20)) x += (bmax0 - bmin0);
          while (x >= bmax0 && ++guard < (1 << 20)) x -= (bmax0 - bmin0);
        }
        layer = (int)floor((x - bmin0) * inv_side0);
        if (layer < 0 || layer >= S0) layer = -1;
      }
      layer_out[p] = layer;
    }
    return;
  }
  if (p < n) {
    double x = pos[(size_t)p * D];
    if (isfinite(x)) {
      if (periodic0) {
        int guard = 0;
        while (x < bmin0 && ++guard < (1 << 20)) x += (bmax0 - bmin0);
        while (x >= bmax0 && ++guard < (1 << 20)) x -= (bmax0 - bmin0);
      }
      const int layer = (int)floor((x - bmin0) * inv_side0);
      if (layer >= 0 && layer < S0 && !(layer >= lo_layer && layer < hi_layer)) {
        // the nearer slab face decides the direction (neighbour-only migration)
        int down = lo_layer - 1 - layer, up = layer - hi_layer;
        if (periodic0) {
          down = ((down % S0) + S0) % S0;
          up = ((up % S0) + S0) % S0;
        } else {
          if (down < 0) down = S0;
          if (up < 0) up = S0;
        }
        c = down <= up ? 1u : 2u;
      }
    }
    cls[p] = (uint8_t)c;
  }
  // one atomic per warp and class
  for (uint32_t k = 1; k <= 2; ++k) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, p < n && c == k);
    if (m && (threadIdx.x & 31) == (uint32_t)(__ffs(m) - 1)) atomicAdd(&counts[k], (uint32_t)__popc(m));
  }
}

} // namespace abr

using abr::Handle;

extern "C" {

int abr_celllist_patch_ghosts(abr_handle hh, const double *pos_local, size_t n_ghost_lo, size_t n_own, size_t n_ghost_hi, const uint32_t *bb_lo,
                              const uint32_t *be_lo, const uint32_t *bb_hi, const uint32_t *be_hi) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!h->windowed || !h->built) return abr::set_error(h, ABR_ERR_STATE, "patch_ghosts: build the owned particles in a slab window first");
  if (n_own != h->n_alive_last) return abr::set_error(h, ABR_ERR_INVALID, "patch_ghosts: n_own differs from the alive count of the owned build");
  ABR_CUDA(h, cudaSetDevice(h->device));
  uint64_t per_layer = 1;
  for (int d = 1; d < h->D; ++d) per_layer *= h->size[d];
  const uint64_t nb_lo = per_layer * (uint64_t)h->own_lo, nb_own = per_layer * (uint64_t)h->own_n,
                 nb_hi = per_layer * (uint64_t)(h->win_n - h->own_lo - h->own_n);
  if ((nb_lo > 0 && (!bb_lo || !be_lo)) || (nb_hi > 0 && (!bb_hi || !be_hi)))
    return abr::set_error(h, ABR_ERR_INVALID, "patch_ghosts: bucket ranges of a ghost side are missing");
  if (n_ghost_lo + n_own + n_ghost_hi >= 0x7FFFFFFFull) return abr::set_error(h, ABR_ERR_UNSUPPORTED, "patch_ghosts: local set too large");
  const uint64_t total = nb_lo + nb_own + nb_hi;
  const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)h->sm_count * 16);
  abr::k_patch_ghosts<<<grid, 256, 0, h->stream>>>(h->bucket_begin.as<uint32_t>(), h->bucket_end.as<uint32_t>(), (uint32_t)nb_lo, (uint32_t)nb_own,
                                                   (uint32_t)nb_hi, bb_lo, be_lo, bb_hi, be_hi, (uint32_t)n_ghost_lo, (uint32_t)n_own);
  h->launches += 1;
  ABR_CUDA(h, cudaGetLastError());
  h->pos_sorted = pos_local;
  h->n_sorted = n_ghost_lo + n_own + n_ghost_hi;
  h->n_alive_last = h->n_sorted;
  h->sorted_keys = nullptr; // m_bucket_indices is not kept for ghost particles
  return ABR_OK;
}

int abr_slab_classify(abr_handle hh, const double *pos, size_t n, int lo_layer, int hi_layer, uint8_t *cls, uint32_t *counts3) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!h->domain_set) return abr::set_error(h, ABR_ERR_STATE, "slab_classify: domain has not been set");
  if (n >= 0xFFFFFFF0ull) return abr::set_error(h, ABR_ERR_UNSUPPORTED, "slab_classify: too many particles");
  if (n > 0 && (!pos || !cls || !counts3)) return abr::set_error(h, ABR_ERR_INVALID, "slab_classify: null pointer");
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::fill_u32(h, counts3, 0u, 3);
  if (n == 0) return ABR_OK;
  abr::k_slab_classify<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(pos, (uint32_t)n, h->D, h->bmin[0], h->bmax[0], h->inv_side[0], (int)h->size[0],
                                                                          h->periodic[0] ? 1 : 0, lo_layer, hi_layer, cls, counts3, nullptr);
  h->launches += 1;
  ABR_CUDA(h, cudaGetLastError());
  return ABR_OK;
}

int abr_slab_layers(abr_handle hh, const double *pos, size_t n, int32_t *layer_out) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return ABR_ERR_INVALID;
  if (!h->domain_set) return abr::set_error(h, ABR_ERR_STATE, "slab_layers: domain has not been set");
  if (n >= 0xFFFFFFF0ull) return abr::set_error(h, ABR_ERR_UNSUPPORTED, "slab_layers: too many particles");
  if (n == 0) return ABR_OK;
  if (!pos || !layer_out) return abr::set_error(h, ABR_ERR_INVALID, "slab_layers: null pointer");
  ABR_CUDA(h, cudaSetDevice(h->device));
  abr::k_slab_classify<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(pos, (uint32_t)n, h->D, h->bmin[0], h->bmax[0], h->inv_side[0], (int)h->size[0],
                                                                          h->periodic[0] ? 1 : 0, 0, 0, nullptr, nullptr, layer_out);
  h->launches += 1;
  ABR_CUDA(h, cudaGetLastError());
  return ABR_OK;
}

} // extern "C"
