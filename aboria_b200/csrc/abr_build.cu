// Ordered cell-list build for sm_100a.
//
// Replaces the Thrust call chain of the reference
//   neighbour_search_base::update_positions   (src/NeighbourSearchBase.h:350-495)
//   CellListOrdered::update_positions_impl    (src/CellListOrdered.h:190-259)
//   Particles::reorder                        (src/Particles.h:694-724)
// with four hand-written stages:
//   k1  enforce_domain + alive + bucket key, fused (one pass over positions)
//   k2  LSD radix sort of (key, index), 8-bit digits, stable
//   k3  bucket_begin / bucket_end from run boundaries of the sorted keys
//       (+ a suffix-min fill for empty buckets, = lower/upper_bound semantics)
//   k4  multi-column gather with coalesced 8-byte-word stores
// Dead particles get the key `key_bound` (> every live key) so the stable sort
// leaves exactly m_alive_indices-after-sort_by_key in front; no scan/scatter.
//
// Compiled with -fmad=false (see grid.cuh).
#include <algorithm>
#include <cmath>

#include "abr_internal.h"
#include "abr_enforce.cuh"

namespace abr {

// ---------------------------------------------------------------------------
// k1: enforce_domain_lambda (src/NeighbourSearchBase.h:208-237) followed by
// point_to_bucket_index (src/detail/SpatialUtil.h:118-131).  One thread per
// particle; the position / alive flag are written back only when they changed
// (same memory image as the reference's unconditional store, fewer bytes).
// ---------------------------------------------------------------------------
constexpr int EK_THREADS = 256;
constexpr int EK_TILE = 4096; // == RS_TILE: one block per radix tile, so the block can hand over the tile's first histogram

// hist != nullptr: also writes the tile's digit histogram of (key >> hist_shift) & 255 in the
// digit-major layout of k_radix_hist (hist[digit * num_tiles + tile]) — the first pass of the
// sort then needs no histogram kernel of its own.
template <int D, bool WINDOWED>
__global__ void __launch_bounds__(EK_THREADS)
k_enforce_key(double *__restrict__ pos, uint8_t *__restrict__ alive, uint32_t n, Grid g,
              uint32_t *__restrict__ keys, DevScalars *sc, uint32_t *__restrict__ hist, int hist_shift, uint32_t num_tiles) {
  __shared__ uint32_t s_hist[256];
  const uint32_t tile = blockIdx.x;
  if (hist) s_hist[threadIdx.x] = 0;
  if (hist) __syncthreads();
#pragma unroll 4
  for (int j = 0; j < EK_TILE / EK_THREADS; ++j) {
    const uint32_t p = tile * EK_TILE + j * EK_THREADS + threadIdx.x;
    if (p >= n) break;
    const uint32_t key = enforce_one<D, WINDOWED>(pos, alive, p, g, sc);
    keys[p] = key;
    if (hist) atomicAdd(&s_hist[(key >> hist_shift) & 255u], 1u);
  }
  if (hist) {
    __syncthreads();
    hist[(size_t)threadIdx.x * num_tiles + tile] = s_hist[threadIdx.x];
  }
}

// ---------------------------------------------------------------------------
// k2: radix sort.  Tile = 256 threads x 16 keys, warp-striped so that the
// order (warp, slot, lane) equals the memory order -> stable ranks.
// ---------------------------------------------------------------------------
constexpr int RS_THREADS = 512;
constexpr int RS_IPT = 8;
constexpr int RS_TILE = RS_THREADS * RS_IPT;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RADIX = 256;

// a set of columns (device pointers + element sizes) passed to kernels by value
struct GatherCols {
  int ncols;
  const uint8_t *src[GP_MAXC];
  uint8_t *dst[GP_MAXC];
  uint32_t eb[GP_MAXC];
};

// Tile table of a SEGMENTED pass (two-level build): tiles never straddle the bins
// of the first-level partition, and the histogram is laid out [bin][digit][tile in
// bin] so that one flat exclusive scan yields bin-local destinations.  Null table
// = dense pass: tile t covers [t*RS_TILE, ...), histogram digit-major [digit][tile].
struct TileTab {
  const uint32_t *start;   // first element of the tile
  const uint32_t *count;   // elements in the tile
  const uint32_t *hbase;   // histogram index of (digit 0, this tile)
  const uint32_t *hstride; // histogram stride between digits
  const uint32_t *total;   // number of tiles in use (device)
};

// One element of one column, src -> dst.  All loads of a batch (up to CB 8-byte words)
// are issued before the first store, so an element costs one memory round trip per
// batch instead of one per word (the reorder kernels are latency bound, profiles/).
template <int CB = 8>
__device__ __forceinline__ void copy_element(const uint8_t *__restrict__ s, uint8_t *__restrict__ t, uint32_t eb) {
  if ((eb & 7u) == 0 && (((uintptr_t)s | (uintptr_t)t) & 7u) == 0) {
    const uint64_t *s8 = reinterpret_cast<const uint64_t *>(s);
    uint64_t *t8 = reinterpret_cast<uint64_t *>(t);
    const uint32_t words = eb / 8;
    for (uint32_t w0 = 0; w0 < words; w0 += CB) {
      uint64_t v[CB];
#pragma unroll
      for (int i = 0; i < CB; ++i)
        if (w0 + i < words) v[i] = __ldg(s8 + w0 + i);
#pragma unroll
      for (int i = 0; i < CB; ++i)
        if (w0 + i < words) t8[w0 + i] = v[i];
    }
  } else if ((eb & 3u) == 0 && (((uintptr_t)s | (uintptr_t)t) & 3u) == 0) {
    const uint32_t *s4 = reinterpret_cast<const uint32_t *>(s);
    uint32_t *t4 = reinterpret_cast<uint32_t *>(t);
    const uint32_t words = eb / 4;
    for (uint32_t w0 = 0; w0 < words; w0 += CB) {
      uint32_t v[CB];
#pragma unroll
      for (int i = 0; i < CB; ++i)
        if (w0 + i < words) v[i] = __ldg(s4 + w0 + i);
#pragma unroll
      for (int i = 0; i < CB; ++i)
        if (w0 + i < words) t4[w0 + i] = v[i];
    }
  } else {
    for (uint32_t w0 = 0; w0 < eb; w0 += CB) {
      uint8_t v[CB];
#pragma unroll
      for (int i = 0; i < CB; ++i)
        if (w0 + i < eb) v[i] = __ldg(s + w0 + i);
#pragma unroll
      for (int i = 0; i < CB; ++i)
        if (w0 + i < eb) t[w0 + i] = v[i];
    }
  }
}

struct TileInfo {
  uint32_t start, count, hbase, hstride;
  bool live;
};
__device__ __forceinline__ TileInfo tile_info(const TileTab &tt, uint32_t tile, uint32_t n, uint32_t num_tiles) {
  TileInfo t;
  if (tt.start) {
    t.live = tile < *tt.total;
    t.start = t.live ? tt.start[tile] : 0;
    t.count = t.live ? tt.count[tile] : 0;
    t.hbase = t.live ? tt.hbase[tile] : 0;
    t.hstride = t.live ? tt.hstride[tile] : 0;
  } else {
    t.live = true;
    t.start = tile * RS_TILE;
    t.count = min((uint32_t)RS_TILE, n - t.start);
    t.hbase = tile;
    t.hstride = num_tiles;
  }
  return t;
}

// per-tile digit histogram
__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t num_tiles,
             uint32_t *__restrict__ hist, const TileTab tt) {
  __shared__ uint32_t s_hist[RADIX];
  const TileInfo ti = tile_info(tt, blockIdx.x, n, num_tiles);
  if (!ti.live) return;
  if (threadIdx.x < RADIX) s_hist[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_IPT; ++j) {
    const uint32_t q = j * RS_THREADS + threadIdx.x;
    if (q < ti.count) atomicAdd(&s_hist[(keys[ti.start + q] >> shift) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < RADIX) hist[(size_t)threadIdx.x * ti.hstride + ti.hbase] = s_hist[threadIdx.x];
}

// bulk-copy engine helpers of the staged record move (k_radix_scatter<2>)
__device__ __forceinline__ void rs_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rs_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void rs_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
// where each column's window of a tile sits in the staging area (bytes from its start, 16-byte aligned)
struct StageLayout {
  uint32_t off[GP_MAXC];
  uint32_t bar_off; // the mbarrier
};
// The binned copy as ONE record per particle (k_radix_scatter<3> writes it, k_gather_records reads it): the
// columns made of 8-byte words one after the other, then a tail word holding the original index (bytes 0..3)
// and the 1/2/4-byte columns (bytes 4..7).  position + id + alive: 5 words, 40 bytes.  The final reorder then
// touches one or two lines per particle instead of one per column.
constexpr int REC_MAXW = 8;
struct RecLayout {
  int nwords;               // 8-byte words per record, tail included
  int nsmall;               // columns packed in the tail word
  uint32_t inv;             // ceil(2^32 / nwords)
  uint32_t wbase[REC_MAXW]; // word w < nwords-1: byte offset of its source inside the staging area (element 0)
  uint32_t wstride[REC_MAXW]; //                    and the element size of its column
  uint8_t wcol[REC_MAXW], wsub[REC_MAXW]; // its column and word index inside the column's element
  uint8_t scol[4], sshift[4], sbytes[4];  // tail word: column, bit shift, size of every small column
  uint32_t soff[4];                       // and the byte offset of its window inside the staging area
};

template <int WARPS> struct RadixScatterSmemT {
  uint32_t warp_hist[WARPS][RADIX];
  uint32_t digit_start[RADIX];
  uint32_t glob_off[RADIX];
  uint32_t s_keys[RS_TILE];
  uint32_t s_idx[RS_TILE];
  uint32_t s_scan[RADIX / 32];
};
using RadixScatterSmem = RadixScatterSmemT<RS_WARPS>;
constexpr int RS2_THREADS = 1024; // the staged record move runs one CTA per SM: twice the warps per tile

// MOVE_RECORDS (first level of the two-level build): besides (key, index) the
// whole particle record — every column in `cols` — moves to its bin.  The tile's
// records are read from a 4096-element window (L1/L2 resident) and written in the
// sorted order of the tile, i.e. in contiguous runs per bin.
//
// MOVE_RECORDS == 2, the staged record move: the column windows are CONTIGUOUS in the input, so one thread
// hands them to the bulk-copy engine (cp.async.bulk -> shared memory, completion on an mbarrier) before
// the ranking starts; the copy runs under the ranking, and the output phase reads the records from
// shared memory (any order, no L1 tag look-ups, no L2 latency) and writes the sorted runs with
// word-parallel coalesced stores.  One CTA per SM (the staging area of a 4096-record tile of
// position + id + alive is 132 KB); chosen on the host when the windows fit and are 16-byte aligned.
template <int MOVE_RECORDS, int THREADS = RS_THREADS>
__global__ void __launch_bounds__(THREADS, MOVE_RECORDS >= 2 ? 1 : 2)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ idx_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ idx_out,
                const uint32_t *__restrict__ tile_offsets, int shift, uint32_t n,
                uint32_t num_tiles, const TileTab tt, const GatherCols cols, const StageLayout stage, const RecLayout rec) {
  constexpr int IPT = RS_TILE / THREADS, WARPS = THREADS / 32;
  using Smem = RadixScatterSmemT<WARPS>;
  extern __shared__ __align__(16) unsigned char rs_raw[];
  Smem &S = *reinterpret_cast<Smem *>(rs_raw);
  unsigned char *const stage_base = rs_raw + ((sizeof(Smem) + 15) & ~(size_t)15);
  auto &warp_hist = S.warp_hist;
  auto &digit_start = S.digit_start;
  auto &glob_off = S.glob_off;
  auto &s_keys = S.s_keys;
  auto &s_idx = S.s_idx;
  auto &s_scan = S.s_scan;

  const TileInfo ti = tile_info(tt, blockIdx.x, n, num_tiles);
  if (!ti.live) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lane_lt = (1u << lane) - 1u;
  // this tile's global offset of digit `tid`: asked for now, needed after the ranking
  const uint32_t my_glob_off = tid < RADIX ? tile_offsets[(size_t)tid * ti.hstride + ti.hbase] : 0u;
  if (MOVE_RECORDS >= 2) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(stage_base + stage.bar_off);
    if (tid == 0) {
      rs_mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      uint32_t total = 0;
      for (int c = 0; c < cols.ncols; ++c) total += (ti.count * cols.eb[c]) & ~15u;
      rs_mbar_arrive_expect_tx(bar, total);
      for (int c = 0; c < cols.ncols; ++c) {
        const uint32_t bytes = (ti.count * cols.eb[c]) & ~15u;
        if (bytes)
          rs_bulk_g2s((uint32_t)__cvta_generic_to_shared(stage_base + stage.off[c]), cols.src[c] + (uint64_t)ti.start * cols.eb[c], bytes, bar);
      }
    }
    // the last bytes of a partial tile's window (the bulk copy moves multiples of 16)
    for (int c = 0; c < cols.ncols; ++c) {
      const uint32_t all = ti.count * cols.eb[c];
      const uint8_t *base = cols.src[c] + (uint64_t)ti.start * cols.eb[c];
      for (uint32_t b = (all & ~15u) + tid; b < all; b += THREADS) stage_base[stage.off[c] + b] = base[b];
    }
  }
  if (MOVE_RECORDS == 1) {
    // the records of this tile are a contiguous window of every column: start pulling it
    // into L2 now, the copy at the end of the kernel then runs at L2 latency
    for (int c = 0; c < cols.ncols; ++c) {
      const uint64_t bytes = (uint64_t)ti.count * cols.eb[c];
      const uint8_t *base = cols.src[c] + (uint64_t)ti.start * cols.eb[c];
      for (uint64_t off = (uint64_t)tid * 128; off < bytes; off += (uint64_t)THREADS * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }
  }
  for (int i = tid; i < WARPS * RADIX; i += THREADS) (&warp_hist[0][0])[i] = 0;
  __syncthreads();

  const uint32_t qbase = warp * (IPT * 32); // position inside the tile
  const uint32_t wbase = ti.start + qbase;
  uint32_t key[IPT], val[IPT];
  uint16_t rank[IPT];
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    const uint32_t p = wbase + j * 32 + lane;
    const bool valid = qbase + j * 32 + lane < ti.count;
    key[j] = valid ? keys_in[p] : 0xFFFFFFFFu;
    val[j] = valid ? (idx_in ? idx_in[p] : p) : 0u;
  }
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    const bool valid = qbase + j * 32 + lane < ti.count;
    const uint32_t d = (key[j] >> shift) & (RADIX - 1);
    // lanes holding the same digit: eight independent ballots (pipelined) instead
    // of MATCH.ANY, whose latency dominated this kernel (profiles/)
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
    if (!valid) peers = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
      const uint32_t vote = __ballot_sync(0xFFFFFFFFu, (d >> bit) & 1u);
      peers &= ((d >> bit) & 1u) ? vote : ~vote;
    }
    const uint32_t prev = warp_hist[warp][d];
    __syncwarp();
    if (valid && (peers & lane_lt) == 0) warp_hist[warp][d] = prev + __popc(peers);
    __syncwarp();
    rank[j] = (uint16_t)(prev + __popc(peers & lane_lt));
  }
  __syncthreads();

  // per digit: exclusive prefix over warps, block total (threads 0..255 own a digit)
  if (tid < RADIX) {
    const int d = tid;
    uint32_t running = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const uint32_t c = warp_hist[w][d];
      warp_hist[w][d] = running;
      running += c;
    }
    const uint32_t total = running;
    glob_off[d] = my_glob_off;
    // exclusive scan of the 256 digit totals
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    digit_start[d] = incl - total; // warp-local exclusive; completed below
  }
  __syncthreads();
  if (tid < RADIX) {
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
      if (w < warp) woff += s_scan[w];
    digit_start[tid] += woff;
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    if (qbase + j * 32 + lane < ti.count) {
      const uint32_t d = (key[j] >> shift) & (RADIX - 1);
      const uint32_t slot = digit_start[d] + warp_hist[warp][d] + rank[j];
      s_keys[slot] = key[j];
      s_idx[slot] = val[j];
    }
  }
  __syncthreads();
  const uint32_t tile_count = ti.count;
  if constexpr (MOVE_RECORDS == 3) {
    // one record per particle: s_keys[s] becomes the output position of sorted slot s, s_idx[s] the record's
    // position inside the staged window (the original index travels in the record's tail word)
    __shared__ uint32_t s_wbase[REC_MAXW], s_wstride[REC_MAXW];
    if (tid == 0) { // (compile-time indices: a thread-dependent index would copy the parameter block to local memory)
#pragma unroll
      for (int w = 0; w < REC_MAXW; ++w) {
        s_wbase[w] = rec.wbase[w];
        s_wstride[w] = rec.wstride[w];
      }
    }
    for (uint32_t s = tid; s < tile_count; s += THREADS) {
      const uint32_t k = s_keys[s];
      const uint32_t d = (k >> shift) & (RADIX - 1);
      const uint32_t out = glob_off[d] + (s - digit_start[d]);
      keys_out[out] = k;
      s_keys[s] = out;
      s_idx[s] -= ti.start;
    }
    rs_mbar_wait((uint32_t)__cvta_generic_to_shared(stage_base + stage.bar_off), 0u);
    __syncthreads();
    // word-parallel over whole records: one contiguous stream of 8-byte stores per bin.  (Column by column —
    // no per-word table, no divergence, 40 % fewer instructions — was slower, 2.69 against 2.56 ms for the
    // build: its stores fill the sectors of a record in three separate sweeps.)
    const uint32_t NW = (uint32_t)rec.nwords, total = tile_count * NW;
    uint64_t *recs = reinterpret_cast<uint64_t *>(cols.dst[0]);
    for (uint32_t g = tid; g < total; g += THREADS) {
      const uint32_t e = __umulhi(g, rec.inv), w = g - e * NW;
      const uint32_t li = s_idx[e];
      uint64_t v;
      if (w + 1 < NW) {
        v = *reinterpret_cast<const uint64_t *>(stage_base + s_wbase[w] + li * s_wstride[w]);
      } else {
        v = (uint64_t)(ti.start + li);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < rec.nsmall) {
            const uint32_t nb = rec.sbytes[q];
            const unsigned char *sp = stage_base + rec.soff[q] + li * nb;
            uint64_t t = sp[0];
            if (nb >= 2) t |= (uint64_t)sp[1] << 8;
            if (nb == 4) t |= ((uint64_t)sp[2] << 16) | ((uint64_t)sp[3] << 24);
            v |= t << rec.sshift[q];
          }
        }
      }
      recs[(uint64_t)s_keys[e] * NW + w] = v;
    }
  } else if constexpr (MOVE_RECORDS == 2) {
    // keys and original indices first; s_keys[s] then becomes the output position of sorted slot s and
    // s_idx[s] the record's position inside the staged window
    for (uint32_t s = tid; s < tile_count; s += THREADS) {
      const uint32_t k = s_keys[s];
      const uint32_t d = (k >> shift) & (RADIX - 1);
      const uint32_t out = glob_off[d] + (s - digit_start[d]);
      keys_out[out] = k;
      const uint32_t src = s_idx[s];
      idx_out[out] = src;
      s_keys[s] = out;
      s_idx[s] = src - ti.start;
    }
    rs_mbar_wait((uint32_t)__cvta_generic_to_shared(stage_base + stage.bar_off), 0u);
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < cols.ncols; ++c) {
      const uint32_t eb = cols.eb[c];
      const unsigned char *win = stage_base + stage.off[c];
      uint8_t *dst = cols.dst[c];
      if ((eb & 7u) == 0 && eb <= 128 && ((uintptr_t)dst & 7u) == 0) {
        // word-parallel: the W words of a record are moved by W consecutive threads
        const uint32_t W = eb >> 3, total = tile_count * W;
        const uint32_t inv = W > 1 ? (uint32_t)((0x100000000ull + W - 1u) / W) : 0u; // g / W == umulhi(g, inv) for g < 2^16, W <= 16
        const uint64_t *w8 = reinterpret_cast<const uint64_t *>(win);
        uint64_t *t8 = reinterpret_cast<uint64_t *>(dst);
        for (uint32_t g = tid; g < total; g += THREADS) {
          const uint32_t e = W > 1 ? __umulhi(g, inv) : g, w = g - e * W;
          t8[(uint64_t)s_keys[e] * W + w] = w8[s_idx[e] * W + w];
        }
      } else if (eb == 4 && ((uintptr_t)dst & 3u) == 0) {
        for (uint32_t e = tid; e < tile_count; e += THREADS)
          reinterpret_cast<uint32_t *>(dst)[s_keys[e]] = reinterpret_cast<const uint32_t *>(win)[s_idx[e]];
      } else if (eb == 1) {
        for (uint32_t e = tid; e < tile_count; e += THREADS) dst[s_keys[e]] = win[s_idx[e]];
      } else {
        for (uint32_t e = tid; e < tile_count; e += THREADS) {
          const unsigned char *sp = win + (size_t)s_idx[e] * eb;
          uint8_t *tp = dst + (uint64_t)s_keys[e] * eb;
          for (uint32_t b = 0; b < eb; ++b) tp[b] = sp[b];
        }
      }
    }
  } else {
    for (uint32_t s = tid; s < tile_count; s += THREADS) {
      const uint32_t k = s_keys[s];
      const uint32_t d = (k >> shift) & (RADIX - 1);
      const uint32_t out = glob_off[d] + (s - digit_start[d]);
      keys_out[out] = k;
      const uint32_t src = s_idx[s];
      idx_out[out] = src;
      if (MOVE_RECORDS == 1) {
        for (int c = 0; c < cols.ncols; ++c) {
          const uint32_t eb = cols.eb[c];
          copy_element<4>(cols.src[c] + (uint64_t)src * eb, cols.dst[c] + (uint64_t)out * eb, eb);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// device-wide scans (three-phase): exclusive prefix sum / inclusive suffix min
// ---------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_IPT = 16;
constexpr int SC_TILE = SC_THREADS * SC_IPT;

struct OpSum {
  __device__ static uint32_t identity() { return 0u; }
  __device__ static uint32_t apply(uint32_t a, uint32_t b) { return a + b; }
};
struct OpMin {
  __device__ static uint32_t identity() { return 0xFFFFFFFFu; }
  __device__ static uint32_t apply(uint32_t a, uint32_t b) { return a < b ? a : b; }
};

template <class Op> __device__ inline uint32_t block_reduce(uint32_t v, uint32_t *s_tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = Op::apply(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  if (lane == 0) s_tmp[warp] = v;
  __syncthreads();
  uint32_t r = Op::identity();
  for (int w = 0; w < SC_THREADS / 32; ++w) r = Op::apply(r, s_tmp[w]);
  __syncthreads();
  return r;
}

// REVERSE = false: element order is index order (prefix); true: reversed (suffix)
template <class Op, bool REVERSE>
__global__ void __launch_bounds__(SC_THREADS)
k_scan_reduce(const uint32_t *__restrict__ in, uint64_t m, uint32_t *__restrict__ partial) {
  __shared__ uint32_t s_tmp[SC_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SC_TILE;
  uint32_t v = Op::identity();
  for (int j = 0; j < SC_IPT; ++j) {
    const uint64_t e = base + (uint64_t)j * SC_THREADS + threadIdx.x;
    if (e < m) v = Op::apply(v, in[REVERSE ? (m - 1 - e) : e]);
  }
  v = block_reduce<Op>(v, s_tmp);
  if (threadIdx.x == 0) partial[blockIdx.x] = v;
}

// single block: exclusive scan of the partials in place
template <class Op>
__global__ void __launch_bounds__(1024) k_scan_partials(uint32_t *partial, uint32_t nb) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = Op::identity();
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t e = base + threadIdx.x;
    const uint32_t v = e < nb ? partial[e] : Op::identity();
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl = Op::apply(incl, t);
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = s_carry;
    for (int w = 0; w < warp; ++w) woff = Op::apply(woff, s_w[w]);
    uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) excl = Op::identity();
    excl = Op::apply(woff, excl);
    if (e < nb) partial[e] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = Op::apply(woff, incl);
    __syncthreads();
  }
}

// phase 3.  MODE 0: out[e] = exclusive prefix sum.
//           MODE 1 (bucket fill): in = bucket_begin candidates (0xFFFFFFFF for
//           empty), out_begin = min(inclusive suffix min, n_incell);
//           out_end[c] = end candidate or out_begin[c] when the bucket is empty.
template <class Op, bool REVERSE, int MODE>
__global__ void __launch_bounds__(SC_THREADS)
k_scan_apply(const uint32_t *in, uint64_t m, const uint32_t *__restrict__ partial, uint32_t *out,
             uint32_t *end_io, const DevScalars *sc) {
  __shared__ uint32_t s_w[SC_THREADS / 32];
  __shared__ uint32_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * SC_TILE;
  if (threadIdx.x == 0) s_carry = partial[blockIdx.x];
  __syncthreads();
  for (int j = 0; j < SC_IPT; ++j) {
    const uint64_t e = base + (uint64_t)j * SC_THREADS + threadIdx.x;
    const uint64_t src = REVERSE ? (m - 1 - e) : e;
    const uint32_t v = e < m ? in[src] : Op::identity();
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl = Op::apply(incl, t);
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = s_carry;
    for (int w = 0; w < warp; ++w) woff = Op::apply(woff, s_w[w]);
    if (MODE == 0) {
      uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
      if (lane == 0) excl = Op::identity();
      if (e < m) out[src] = Op::apply(woff, excl);
    } else {
      if (e < m) {
        uint32_t bb = Op::apply(woff, incl);
        const uint32_t lim = sc->n_incell;
        if (bb > lim) bb = lim;
        out[src] = bb;
        if (end_io[src] == 0xFFFFFFFFu) end_io[src] = bb;
      }
    }
    __syncthreads();
    if (threadIdx.x == SC_THREADS - 1) s_carry = Op::apply(woff, incl);
    __syncthreads();
  }
}

template <class Op, bool REVERSE, int MODE>
static cudaError_t device_scan(Handle *h, const uint32_t *in, uint64_t m, uint32_t *out,
                               uint32_t *end_io) {
  if (m == 0) return cudaSuccess;
  const uint32_t nb = (uint32_t)((m + SC_TILE - 1) / SC_TILE);
  cudaError_t e = h->scan_tmp.reserve((size_t)nb * sizeof(uint32_t));
  if (e != cudaSuccess) return e;
  uint32_t *partial = h->scan_tmp.as<uint32_t>();
  k_scan_reduce<Op, REVERSE><<<nb, SC_THREADS, 0, h->stream>>>(in, m, partial);
  k_scan_partials<Op><<<1, 1024, 0, h->stream>>>(partial, nb);
  k_scan_apply<Op, REVERSE, MODE><<<nb, SC_THREADS, 0, h->stream>>>(in, m, partial, out, end_io,
                                                                    h->d_scalars);
  h->launches += 3;
  return cudaGetLastError();
}

int scan_exclusive_u32(Handle *h, uint32_t *data, uint64_t m) {
  cudaError_t e = device_scan<OpSum, false, 0>(h, data, m, data, nullptr);
  if (e != cudaSuccess) return check_cuda(h, e, "exclusive scan");
  return ABR_OK;
}

// ---------------------------------------------------------------------------
// k3: run boundaries of the sorted keys.  Equivalent to lower_bound /
// upper_bound of every bucket id in the sorted key array
// (src/CellListOrdered.h:229-239) once empty buckets are filled by the
// suffix-min pass above.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_boundaries(const uint32_t *__restrict__ keys, uint32_t n, uint32_t ncells, uint32_t dead_key,
             uint32_t *__restrict__ bb, uint32_t *__restrict__ be, DevScalars *sc) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t k = keys[p];
  const uint32_t kprev = p > 0 ? keys[p - 1] : 0xFFFFFFFFu;
  const uint32_t knext = p + 1 < n ? keys[p + 1] : 0xFFFFFFFFu;
  if (p > 0 && k < kprev) atomicAdd(&sc->n_unsorted, 1u);
  if (p == 0 || k != kprev) {
    if (k < ncells) bb[k] = p;
    if (k >= ncells && (p == 0 || kprev < ncells)) sc->n_incell = p;
    if (k == dead_key) sc->n_alive = p;
  }
  if (k != knext || p + 1 == n) {
    if (k < ncells) be[k] = p + 1;
  }
  // bucket occupancy, sampled where a run of 65 equal keys ends: enough to tell the product that heavy buckets exist
  if (p >= 64 && k < ncells && keys[p - 64] == k && (k != knext || p + 1 == n)) atomicMax(&sc->max_bucket, 65u);
}

// The same bucket ranges in ONE sweep over the sorted keys, every bucket written exactly once — no
// 0xFFFFFFFF fills of both arrays, no suffix-min scan over the buckets afterwards.  The thread at a run
// boundary p (keys[p-1] < keys[p]; p == n closes the array) writes bucket_end of the run that ends,
// bucket_begin of the run that starts, and begin = end = p for the EMPTY buckets between the two, which
// is what lower_bound / upper_bound give them (src/CellListOrdered.h:229-239).  Gaps are filled by the
// whole warp (coalesced); a gap longer than GAP_LONG goes to a list that k_fill_gaps sweeps with the
// whole grid, so a cloud that leaves most of the grid empty costs no more than a fill.
constexpr uint32_t GAP_LONG = 4096;
struct GapList {
  uint32_t *count;  // entries appended
  uint32_t *lo, *hi, *val; // buckets [lo, hi) get begin = end = val
  uint32_t capacity;
};
__global__ void __launch_bounds__(256)
k_boundaries_fill(const uint32_t *__restrict__ keys, uint32_t n, uint32_t ncells, uint32_t dead_key,
                  uint32_t *__restrict__ bb, uint32_t *__restrict__ be, DevScalars *sc, const GapList gl) {
  // four consecutive boundaries per thread: one 16-byte load, the neighbours by shuffle
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p0 = t * 4u; // boundaries p0 .. p0+3 (boundary p sits in front of element p; p == n closes the array)
  const int lane = threadIdx.x & 31;
  uint32_t k[4];
  if (p0 + 3 < n) {
    const uint4 v = *reinterpret_cast<const uint4 *>(keys + p0);
    k[0] = v.x, k[1] = v.y, k[2] = v.z, k[3] = v.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) k[e] = p0 + e < n ? keys[p0 + e] : 0xFFFFFFFFu;
  }
  uint32_t kprev = __shfl_up_sync(0xFFFFFFFFu, k[3], 1);
  if (lane == 0) kprev = (p0 > 0 && p0 <= n) ? keys[p0 - 1] : 0xFFFFFFFFu;
  uint32_t knext4 = __shfl_down_sync(0xFFFFFFFFu, k[0], 1);
  if (lane == 31) knext4 = p0 + 4 < n ? keys[p0 + 4] : 0xFFFFFFFFu;
  uint32_t glo[4], ghi[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint32_t p = p0 + e;
    const uint32_t kk = k[e];
    const uint32_t kp = e == 0 ? kprev : k[e - 1];
    const uint32_t kn = e == 3 ? knext4 : k[e + 1];
    glo[e] = ghi[e] = 0;
    if (p <= n) {
      const uint32_t kc = kk < ncells ? kk : ncells;                   // keys beyond the grid (dead, outside) close it
      const uint32_t kp1 = p > 0 ? (kp < ncells ? kp + 1u : ncells) : 0u; // first bucket after the previous run
      if (p < n) {
        if (p > 0 && kk < kp) atomicAdd(&sc->n_unsorted, 1u);
        if (p == 0 || kk != kp) {
          if (kk >= ncells && (p == 0 || kp < ncells)) sc->n_incell = p;
          if (kk == dead_key) sc->n_alive = p;
        }
        // bucket occupancy, sampled where a run of 65 equal keys ends: tells the product that heavy buckets exist
        if ((kk != kn || p + 1 == n) && p >= 64 && kk < ncells && keys[p - 64] == kk) atomicMax(&sc->max_bucket, 65u);
      }
      if (p == 0 || p == n || kk != kp) {
        if (p > 0 && kp < ncells) be[kp] = p;
        if (kc < ncells) bb[kc] = p;
        if (kp1 < kc) {
          glo[e] = kp1;
          ghi[e] = kc;
        }
      }
    }
  }
  // empty buckets, warp-cooperatively
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint32_t pending = __ballot_sync(0xFFFFFFFFu, ghi[e] > glo[e]);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const uint32_t lo = __shfl_sync(0xFFFFFFFFu, glo[e], src), hi = __shfl_sync(0xFFFFFFFFu, ghi[e], src);
      const uint32_t val = __shfl_sync(0xFFFFFFFFu, p0, src) + e;
      if (hi - lo > GAP_LONG) {
        if (lane == 0) {
          const uint32_t slot = atomicAdd(gl.count, 1u);
          if (slot < gl.capacity) {
            gl.lo[slot] = lo;
            gl.hi[slot] = hi;
            gl.val[slot] = val;
          }
        }
      } else {
        for (uint32_t j = lo + lane; j < hi; j += 32) {
          bb[j] = val;
          be[j] = val;
        }
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_fill_gaps(const GapList gl, uint32_t *__restrict__ bb, uint32_t *__restrict__ be) {
  const uint32_t m = min(*gl.count, gl.capacity);
  for (uint32_t g = 0; g < m; ++g) {
    const uint32_t lo = gl.lo[g], hi = gl.hi[g], val = gl.val[g];
    for (uint32_t j = lo + blockIdx.x * blockDim.x + threadIdx.x; j < hi; j += gridDim.x * blockDim.x) {
      bb[j] = val;
      be[j] = val;
    }
  }
}

// ---------------------------------------------------------------------------
// k4: gather.  8-byte-word columns: one thread per output word, so the stores
// of a warp are one contiguous 256-byte span; the loads of one element are
// contiguous too.  Other element sizes fall back to a byte-granular variant.
// ---------------------------------------------------------------------------
// n_dev (optional): device-side element count (the alive count of the build that
// is still in flight), so the reorder can be enqueued without a host round trip.
__global__ void __launch_bounds__(256)
k_gather_words(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst,
               const int32_t *__restrict__ order, uint64_t n_out, const uint32_t *__restrict__ n_dev,
               uint32_t words) {
  // stage 256 elements in shared memory, then store them as one contiguous,
  // 16-byte-vectorised span (the loads are the random side of the gather)
  extern __shared__ uint64_t s_stage[];
  if (n_dev) n_out = min(n_out, (uint64_t)*n_dev);
  const uint64_t k0 = (uint64_t)blockIdx.x * 256;
  if (k0 >= n_out) return;
  const uint32_t cnt = (uint32_t)min((uint64_t)256, n_out - k0);
  if (threadIdx.x < cnt) {
    const uint64_t *e = src + (uint64_t)order[k0 + threadIdx.x] * words;
    for (uint32_t w = 0; w < words; ++w) s_stage[threadIdx.x * words + w] = __ldg(e + w); // plain nc load: .cs/.no_allocate fetch MORE from HBM (tools/gather_probe.cu)
  }
  __syncthreads();
  const uint32_t total = cnt * words;
  uint64_t *out = dst + k0 * words; // 16-byte aligned: k0 * words * 8 is a multiple of 2048
  if ((total & 1u) == 0 && ((uintptr_t)out & 15u) == 0) {
    const ulonglong2 *sv = reinterpret_cast<const ulonglong2 *>(s_stage);
    ulonglong2 *ov = reinterpret_cast<ulonglong2 *>(out);
    for (uint32_t t = threadIdx.x; t < total / 2; t += 256) ov[t] = sv[t];
  } else {
    for (uint32_t t = threadIdx.x; t < total; t += 256) out[t] = s_stage[t];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_gather_small(const T *__restrict__ src, T *__restrict__ dst, const int32_t *__restrict__ order,
               uint64_t n_out, const uint32_t *__restrict__ n_dev) {
  if (n_dev) n_out = min(n_out, (uint64_t)*n_dev);
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_out) dst[k] = src[order[k]];
}

__global__ void __launch_bounds__(256)
k_gather_bytes(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
               const int32_t *__restrict__ order, uint64_t n_out, const uint32_t *__restrict__ n_dev, uint32_t eb) {
  if (n_dev) n_out = min(n_out, (uint64_t)*n_dev);
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_out * eb) return;
  const uint64_t k = g / eb;
  const uint32_t w = (uint32_t)(g - k * eb);
  dst[g] = src[(uint64_t)order[k] * eb + w];
}

// fp64 FMA throughput probe: the denominator of the matvec's fp64 roofline
// (MEASURED_PEAKS.json has no fp64 entry; SURVEY.md §8d asks for a DFMA microbenchmark)
__global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

int probe_fp64_peak(Handle *h, double *tflops) {
  const int blocks = h->sm_count * 8, iters = 1 << 14;
  double *out = nullptr;
  ABR_CUDA(h, cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, h->stream);
    k_dfma_probe<<<blocks, 256, 0, h->stream>>>(out, iters);
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  ABR_CUDA(h, cudaGetLastError());
  const double flops = 2.0 * 8.0 * (double)iters * blocks * 256.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  return ABR_OK;
}

static inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// Phased gather for column sets larger than L2.  A random gather from an array
// that does not fit in L2 pulls ~91 B from HBM per element whatever its size
// (tools/gather_probe.cu).  Here the SOURCE index range is cut into windows
// whose rows (all columns together) fit in L2; one launch per window streams the
// whole permutation (coalesced, evict-first) and moves only the elements whose
// source lies in the window — every source sector is fetched from HBM once and
// reused from L2 by the other elements of its line.

__global__ void __launch_bounds__(256)
k_gather_phased(const GatherCols cols, const int32_t *__restrict__ order, uint64_t n_out,
                const uint32_t *__restrict__ n_dev, uint32_t lo, uint32_t hi) {
  if (n_dev) n_out = min(n_out, (uint64_t)*n_dev);
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_out) return;
  const uint32_t o = (uint32_t)__ldcs(order + k);
  if (o < lo || o >= hi) return;
  for (int c = 0; c < cols.ncols; ++c) {
    const uint32_t eb = cols.eb[c];
    const uint8_t *s = cols.src[c] + (uint64_t)o * eb;
    uint8_t *d = cols.dst[c] + k * eb;
    if ((eb & 7u) == 0 && (((uintptr_t)s | (uintptr_t)d) & 7u) == 0) {
      for (uint32_t w = 0; w < eb / 8; ++w) __stcs(reinterpret_cast<uint64_t *>(d) + w, __ldg(reinterpret_cast<const uint64_t *>(s) + w));
    } else if ((eb & 3u) == 0 && (((uintptr_t)s | (uintptr_t)d) & 3u) == 0) {
      for (uint32_t w = 0; w < eb / 4; ++w) reinterpret_cast<uint32_t *>(d)[w] = __ldg(reinterpret_cast<const uint32_t *>(s) + w);
    } else {
      for (uint32_t w = 0; w < eb; ++w) d[w] = __ldg(s + w);
    }
  }
}

// Final reorder of the two-level build, all columns in one launch: the permutation
// is read once per particle and every source element sits in the L2-resident bin of
// its destination.
__global__ void __launch_bounds__(256, 5)
k_gather_fused(const GatherCols cols, const uint32_t *__restrict__ perm, uint32_t n_out, const uint32_t *__restrict__ n_dev) {
  // A block moves 256 consecutive output elements of every column.  Columns made of
  // 8-byte words are gathered WORD-parallel: the W words of an element are read by W
  // consecutive lanes, so a warp-wide load touches ~32/W records instead of 32 (the L1 tag
  // stage, one look-up per distinct line per request, is what bounds this kernel:
  // profiles/r1x_ncu_build_kernels.txt) and the stores are one contiguous span.
  __shared__ uint32_t s_perm[256];
  if (n_dev) n_out = min(n_out, *n_dev);
  const uint32_t k0 = blockIdx.x * 256;
  if (k0 >= n_out) return;
  const uint32_t cnt = min(256u, n_out - k0);
  const uint32_t tid = threadIdx.x;
  if (tid < cnt) s_perm[tid] = perm[k0 + tid];
  __syncthreads();
#pragma unroll 1
  for (int c = 0; c < cols.ncols; ++c) {
    const uint32_t eb = cols.eb[c];
    const uint8_t *src = cols.src[c];
    uint8_t *dst = cols.dst[c] + (uint64_t)k0 * eb;
    if ((eb & 7u) == 0 && eb <= 128 && (((uintptr_t)src | (uintptr_t)dst) & 7u) == 0) {
      const uint32_t W = eb >> 3, total = cnt * W;
      const uint32_t inv = (65536u + W - 1u) / W; // g / W == (g * inv) >> 16 for g < 4096, W <= 16
      const uint64_t *s8 = reinterpret_cast<const uint64_t *>(src);
      uint64_t *t8 = reinterpret_cast<uint64_t *>(dst);
      constexpr int U = 4;
      for (uint32_t g0 = tid; g0 < total; g0 += 256 * U) {
        uint64_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t g = g0 + u * 256;
          if (g < total) {
            const uint32_t e = (g * inv) >> 16;
            v[u] = __ldg(s8 + (uint64_t)s_perm[e] * W + (g - e * W));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t g = g0 + u * 256;
          if (g < total) t8[g] = v[u];
        }
      }
    } else if (eb == 4 && (((uintptr_t)src | (uintptr_t)dst) & 3u) == 0) {
      if (tid < cnt) reinterpret_cast<uint32_t *>(dst)[tid] = __ldg(reinterpret_cast<const uint32_t *>(src) + s_perm[tid]);
    } else if (eb == 1) {
      if (tid < cnt) dst[tid] = __ldg(src + s_perm[tid]);
    } else {
      if (tid < cnt) copy_element<4>(src + (uint64_t)s_perm[tid] * eb, dst + (uint64_t)tid * eb, eb);
    }
  }
}

// The same reorder with the loads of EVERY column in flight at once.  k_gather_fused walks the columns one
// after the other, so a block pays one memory latency per column — and the narrow columns (id, alive,
// original index) keep only a few hundred bytes per warp in flight while they wait (ncu: long-scoreboard
// stalls 16 cycles per issue, DRAM 36 %).  Here the host lays the work of a thread out as up to GS_SLOTS
// register slots (slot -> column, word round); all loads are issued, then all stores.
constexpr int GS_SLOTS = 8;
struct GatherSlots {
  int nslots;
  uint8_t col[GS_SLOTS];   // column of the slot
  uint8_t round[GS_SLOTS]; // word-parallel columns: the slot moves word g = tid + 256 * round of the block's words
  uint32_t inv[GP_MAXC];   // ceil(2^32 / W) of the column (W = 8-byte words per element), 0 for W == 1
};
__global__ void __launch_bounds__(256, 6)
k_gather_slots(const GatherCols cols, const GatherSlots slots, const uint32_t *__restrict__ perm, uint32_t n_out,
               const uint32_t *__restrict__ n_dev) {
  __shared__ uint32_t s_perm[256];
  if (n_dev) n_out = min(n_out, *n_dev);
  const uint32_t k0 = blockIdx.x * 256;
  if (k0 >= n_out) return;
  const uint32_t cnt = min(256u, n_out - k0);
  const uint32_t tid = threadIdx.x;
  if (tid < cnt) s_perm[tid] = perm[k0 + tid];
  __syncthreads();
  uint64_t v[GS_SLOTS];
#pragma unroll
  for (int k = 0; k < GS_SLOTS; ++k) {
    if (k < slots.nslots) {
      const int c = slots.col[k];
      const uint32_t eb = cols.eb[c];
      const uint8_t *src = cols.src[c];
      if (eb >= 8) {
        const uint32_t W = eb >> 3, g = tid + 256u * slots.round[k];
        if (g < cnt * W) {
          const uint32_t e = W > 1 ? __umulhi(g, slots.inv[c]) : g;
          v[k] = __ldg(reinterpret_cast<const uint64_t *>(src) + (uint64_t)s_perm[e] * W + (g - e * W));
        }
      } else if (eb == 4) {
        if (tid < cnt) v[k] = __ldg(reinterpret_cast<const uint32_t *>(src) + s_perm[tid]);
      } else {
        if (tid < cnt) v[k] = __ldg(src + s_perm[tid]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < GS_SLOTS; ++k) {
    if (k < slots.nslots) {
      const int c = slots.col[k];
      const uint32_t eb = cols.eb[c];
      uint8_t *dst = cols.dst[c] + (uint64_t)k0 * eb;
      if (eb >= 8) {
        const uint32_t W = eb >> 3, g = tid + 256u * slots.round[k];
        if (g < cnt * W) reinterpret_cast<uint64_t *>(dst)[g] = v[k];
      } else if (eb == 4) {
        if (tid < cnt) reinterpret_cast<uint32_t *>(dst)[tid] = (uint32_t)v[k];
      } else {
        if (tid < cnt) dst[tid] = (uint8_t)v[k];
      }
    }
  }
}
// Final reorder from the one-record-per-particle binned copy (RecLayout): the words of a record are read by
// consecutive lanes (one or two lines per particle), all rounds in flight, then dealt out to the columns.
__global__ void __launch_bounds__(256, 6)
k_gather_records(const GatherCols cols, const RecLayout rec, const uint64_t *__restrict__ recs, uint32_t *__restrict__ order_out,
                 const uint32_t *__restrict__ perm, uint32_t n_out, const uint32_t *__restrict__ n_dev) {
  __shared__ uint32_t s_perm[256];
  __shared__ uint64_t *s_wdst[REC_MAXW]; // word w of a record goes to s_wdst[w][element * s_wW[w]]
  __shared__ uint32_t s_wW[REC_MAXW];
  if (n_dev) n_out = min(n_out, *n_dev);
  const uint32_t k0 = blockIdx.x * 256;
  if (k0 >= n_out) return;
  const uint32_t cnt = min(256u, n_out - k0);
  const uint32_t tid = threadIdx.x;
  if (tid < cnt) s_perm[tid] = perm[k0 + tid];
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < REC_MAXW; ++w) {
      const int c = rec.wcol[w];
      s_wdst[w] = reinterpret_cast<uint64_t *>(cols.dst[c]) + rec.wsub[w];
      s_wW[w] = cols.eb[c] >> 3;
    }
  }
  __syncthreads();
  const uint32_t NW = (uint32_t)rec.nwords, total = cnt * NW;
  uint64_t v[REC_MAXW];
#pragma unroll
  for (int k = 0; k < REC_MAXW; ++k) {
    const uint32_t g = tid + 256u * k;
    if (g < total) {
      const uint32_t e = __umulhi(g, rec.inv);
      v[k] = __ldg(recs + (uint64_t)s_perm[e] * NW + (g - e * NW));
    }
  }
#pragma unroll
  for (int k = 0; k < REC_MAXW; ++k) {
    const uint32_t g = tid + 256u * k;
    if (g < total) {
      const uint32_t e = __umulhi(g, rec.inv), w = g - e * NW;
      if (w + 1 < NW) {
        s_wdst[w][(uint64_t)(k0 + e) * s_wW[w]] = v[k];
      } else {
        order_out[k0 + e] = (uint32_t)v[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < rec.nsmall) {
            const uint32_t c = rec.scol[q], nb = rec.sbytes[q];
            const uint32_t t = (uint32_t)(v[k] >> rec.sshift[q]);
            uint8_t *tp = cols.dst[c] + (uint64_t)(k0 + e) * nb;
            if (nb == 1) tp[0] = (uint8_t)t;
            else if (nb == 2) *reinterpret_cast<uint16_t *>(tp) = (uint16_t)t;
            else *reinterpret_cast<uint32_t *>(tp) = t;
          }
        }
      }
    }
  }
}
// record layout of the columns; false when they do not fit (an element that is neither whole 8-byte words nor
// 1/2/4 bytes, more than REC_MAXW words, more than four bytes of small columns, misaligned buffers)
static bool plan_record_layout(const GatherCols &cols, const StageLayout &stage, RecLayout *r) {
  memset(r, 0, sizeof(*r));
  int nw = 0, small_bytes = 0;
  for (int c = 0; c < cols.ncols; ++c) {
    const uint32_t eb = cols.eb[c];
    if (eb >= 8 && (eb & 7u) == 0) {
      if ((uintptr_t)cols.dst[c] & 7u) return false;
      for (uint32_t sub = 0; sub < eb / 8; ++sub) {
        if (nw == REC_MAXW - 1) return false;
        r->wbase[nw] = stage.off[c] + sub * 8;
        r->wstride[nw] = eb;
        r->wcol[nw] = (uint8_t)c;
        r->wsub[nw] = (uint8_t)sub;
        ++nw;
      }
    } else if (eb == 1 || eb == 2 || eb == 4) {
      if ((uintptr_t)cols.dst[c] & (eb - 1)) return false;
      small_bytes = (small_bytes + (int)eb - 1) / (int)eb * (int)eb; // natural alignment inside the tail word
      if (r->nsmall == 4 || small_bytes + (int)eb > 4) return false;
      r->scol[r->nsmall] = (uint8_t)c;
      r->sshift[r->nsmall] = (uint8_t)(32 + 8 * small_bytes);
      r->sbytes[r->nsmall] = (uint8_t)eb;
      r->soff[r->nsmall] = stage.off[c];
      r->nsmall += 1;
      small_bytes += (int)eb;
    } else {
      return false;
    }
  }
  r->nwords = nw + 1;
  r->inv = (uint32_t)((0x100000000ull + (uint32_t)r->nwords - 1u) / (uint32_t)r->nwords);
  return r->nwords >= 2; // (a record of the tail word alone: umulhi(g, 2^32) is not representable; nothing to gain either)
}

// lays the columns out as register slots; false when they do not fit (wide or odd-sized columns)
static bool plan_gather_slots(const GatherCols &cols, GatherSlots *gs) {
  memset(gs, 0, sizeof(*gs));
  int n = 0;
  for (int c = 0; c < cols.ncols; ++c) {
    const uint32_t eb = cols.eb[c];
    const bool words = eb >= 8 && (eb & 7u) == 0 && eb <= 128 && (((uintptr_t)cols.src[c] | (uintptr_t)cols.dst[c]) & 7u) == 0;
    const bool four = eb == 4 && (((uintptr_t)cols.src[c] | (uintptr_t)cols.dst[c]) & 3u) == 0;
    if (!words && !four && eb != 1) return false;
    const uint32_t W = words ? eb >> 3 : 1;
    gs->inv[c] = W > 1 ? (uint32_t)((0x100000000ull + W - 1u) / W) : 0u;
    for (uint32_t r = 0; r < W; ++r) {
      if (n == GS_SLOTS) return false;
      gs->col[n] = (uint8_t)c;
      gs->round[n] = (uint8_t)r;
      ++n;
    }
  }
  gs->nslots = n;
  return n > 0;
}

// Tile table of the segmented passes from the scanned first-level histogram:
// bin b starts at scanned[b * num_tiles] (digit-major layout, tile 0).
struct TileTabW {
  uint32_t *start, *count, *hbase, *hstride, *total, *bin_tile0;
};
__global__ void __launch_bounds__(RADIX) k_tiletab_bins(const uint32_t *__restrict__ scanned, uint32_t num_tiles, uint32_t n, TileTabW w) {
  __shared__ uint32_t s_nt[RADIX];
  const int b = threadIdx.x;
  const uint32_t lo = scanned[(size_t)b * num_tiles];
  const uint32_t hi = b + 1 < RADIX ? scanned[(size_t)(b + 1) * num_tiles] : n;
  const uint32_t nt = (hi - lo + RS_TILE - 1) / RS_TILE;
  s_nt[b] = nt;
  __syncthreads();
  if (b == 0) {
    uint32_t run = 0;
    for (int i = 0; i < RADIX; ++i) {
      const uint32_t c = s_nt[i];
      s_nt[i] = run;
      run += c;
    }
    *w.total = run;
  }
  __syncthreads();
  w.bin_tile0[b] = s_nt[b];
  const uint32_t t0 = s_nt[b];
  for (uint32_t t = 0; t < nt; ++t) {
    w.start[t0 + t] = lo + t * RS_TILE;
    w.count[t0 + t] = min((uint32_t)RS_TILE, hi - (lo + t * RS_TILE));
    w.hbase[t0 + t] = RADIX * t0 + t;
    w.hstride[t0 + t] = nt;
  }
}

__global__ void __launch_bounds__(256) k_iota(uint32_t *out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

int gather_columns(Handle *h, int ncols, const void *const *src, void *const *dst,
                   const size_t *elem_bytes, const int32_t *order, size_t n_out, const uint32_t *n_dev) {
  if (n_out == 0) return ABR_OK;
  size_t row_bytes = 0;
  for (int c = 0; c < ncols; ++c) {
    if (elem_bytes[c] == 0 || !src[c] || !dst[c]) return set_error(h, ABR_ERR_INVALID, "gather: null column");
    row_bytes += elem_bytes[c];
  }
  // source larger than L2 -> phased gather (window = rows worth ~80 MB of source)
  const size_t kWindowBytes = (size_t)80 << 20;
  if (h->phased_gather && row_bytes * n_out > 2 * kWindowBytes && n_out < 0xFFFFFFFFull) {
    const uint32_t win = (uint32_t)std::max<size_t>(1, kWindowBytes / row_bytes);
    // the source index space is [0, n_src); n_out <= n_src and order[] < n_src.  n_src is
    // not known here, but order holds a permutation prefix of it: windows up to 2^32 cover it
    const uint64_t n_src = h->gather_src_n ? h->gather_src_n : n_out;
    for (int c0 = 0; c0 < ncols; c0 += GP_MAXC) {
      GatherCols gc;
      gc.ncols = std::min(GP_MAXC, ncols - c0);
      for (int c = 0; c < gc.ncols; ++c) {
        gc.src[c] = static_cast<const uint8_t *>(src[c0 + c]);
        gc.dst[c] = static_cast<uint8_t *>(dst[c0 + c]);
        gc.eb[c] = (uint32_t)elem_bytes[c0 + c];
      }
      for (uint64_t lo = 0; lo < n_src; lo += win) {
        // the last window is open-ended: order[] may hold indices >= n_out when dead particles were dropped
        const uint64_t hi = (lo + win >= n_src) ? 0xFFFFFFFFull : lo + win;
        k_gather_phased<<<grid_for(n_out, 256), 256, 0, h->stream>>>(gc, order, n_out, n_dev, (uint32_t)lo, (uint32_t)hi);
        h->launches += 1;
      }
    }
    ABR_CUDA(h, cudaGetLastError());
    return ABR_OK;
  }
  for (int c = 0; c < ncols; ++c) {
    const size_t eb = elem_bytes[c];
    const bool aligned8 = (eb % 8 == 0) && ((uintptr_t)src[c] % 8 == 0) && ((uintptr_t)dst[c] % 8 == 0);
    if (aligned8 && eb <= 128) { // staged path: 256 * eb bytes of shared memory (<= 32 KB)
      const uint32_t words = (uint32_t)(eb / 8);
      k_gather_words<<<grid_for(n_out, 256), 256, 256 * eb, h->stream>>>(
          (const uint64_t *)src[c], (uint64_t *)dst[c], order, n_out, n_dev, words);
    } else if (eb == 4 && (uintptr_t)src[c] % 4 == 0 && (uintptr_t)dst[c] % 4 == 0) {
      k_gather_small<uint32_t><<<grid_for(n_out, 256), 256, 0, h->stream>>>(
          (const uint32_t *)src[c], (uint32_t *)dst[c], order, n_out, n_dev);
    } else if (eb == 1) {
      k_gather_small<uint8_t><<<grid_for(n_out, 256), 256, 0, h->stream>>>(
          (const uint8_t *)src[c], (uint8_t *)dst[c], order, n_out, n_dev);
    } else {
      k_gather_bytes<<<grid_for((uint64_t)n_out * eb, 256), 256, 0, h->stream>>>(
          (const uint8_t *)src[c], (uint8_t *)dst[c], order, n_out, n_dev, (uint32_t)eb);
    }
  }
  h->launches += (uint64_t)ncols;
  ABR_CUDA(h, cudaGetLastError());
  return ABR_OK;
}

// ---------------------------------------------------------------------------
// host driver of the build
// ---------------------------------------------------------------------------
// CellListOrdered::set_domain_impl (src/CellListOrdered.h:132-186), host math.
static void set_domain_impl(Handle *h, size_t n) {
  if (h->grid_forced) return;
  if (n < 0.5 * h->size_calculated_with_n || n > 2 * h->size_calculated_with_n) {
    h->size_calculated_with_n = n;
    const int D = h->D;
    if (h->n_leaf > n) {
      for (int d = 0; d < D; ++d) h->size[d] = 1;
    } else {
      double total_volume = 1.0;
      for (int d = 0; d < D; ++d) total_volume *= (h->bmax[d] - h->bmin[d]);
      const double box_volume = h->n_leaf / double(n) * total_volume;
      const double box_side_length = std::pow(box_volume, 1.0 / D);
      for (int d = 0; d < D; ++d) {
        h->size[d] = static_cast<unsigned int>(std::floor((h->bmax[d] - h->bmin[d]) / box_side_length));
        if (h->size[d] == 0) h->size[d] = 1;
      }
    }
    for (int d = 0; d < D; ++d) {
      h->side[d] = (h->bmax[d] - h->bmin[d]) / h->size[d];
      h->inv_side[d] = 1.0 / h->side[d];
    }
  }
}

void host_set_domain_impl(Handle *h, size_t n) { set_domain_impl(h, n); }

Grid Handle::grid() const {
  Grid g;
  g.D = D;
  uint64_t prod = 1, bound = 0;
  for (int d = 0; d < MAXD; ++d) {
    const bool in = d < D;
    g.size[d] = in ? (int)size[d] : 1;
    g.end[d] = g.size[d] - 1;
    g.periodic[d] = in ? (periodic[d] ? 1 : 0) : 0;
    g.bmin[d] = in ? bmin[d] : 0.0;
    g.bmax[d] = in ? bmax[d] : 1.0;
    g.side[d] = in ? side[d] : 1.0;
    g.inv_side[d] = in ? inv_side[d] : 1.0;
    g.L[d] = g.bmax[d] - g.bmin[d];
    if (in) prod *= size[d];
  }
  // largest key an alive particle can get is collapse(size) (every v[d] == size[d])
  for (int d = 0; d < D; ++d) bound = bound * size[d] + size[d];
  g.ncells = (uint32_t)prod;
  g.key_bound = (uint32_t)(bound + 1);
  g.win_lo = 0;
  g.win_n = g.size[0];
  g.own_lo = 0;
  g.own_n = g.size[0];
  if (windowed) {
    g.win_lo = win_lo;
    g.win_n = win_n;
    g.own_lo = own_lo;
    g.own_n = own_n;
    uint64_t per_layer = 1;
    for (int d = 1; d < D; ++d) per_layer *= size[d];
    g.ncells = (uint32_t)(per_layer * (uint64_t)win_n);
    g.key_bound = g.ncells; // windowed keys are local bucket numbers; overflow keys are rejected
  }
  return g;
}

// Scalars and fills go through tiny kernels, never through cudaMemcpyAsync /
// cudaMemsetAsync: those may be routed to a copy engine, where they queue behind
// whatever bulk host<->device transfer another stream has in flight (measured: a
// 3 ms build stretched to 6 ms behind a 5 ms D2H of the previous step's result).
__global__ void k_set_scalars(DevScalars *d, const DevScalars v) { *d = v; }
// h is pinned host memory (device accessible under UVA): the read-back is a 40-byte store over PCIe
__global__ void k_publish_scalars(DevScalars *h, const DevScalars *d) {
  *h = *d;
  __threadfence_system();
}
__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *p, uint32_t v, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}
// the alive column after a reorder: ones for the n_alive survivors (n_dev: device-side count)
__global__ void __launch_bounds__(256) k_fill_ones_bounded(uint8_t *p, uint32_t n, const uint32_t *__restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const uint32_t i = (blockIdx.x * 256 + threadIdx.x) * 16;
  if (i + 16 <= n && (((uintptr_t)p) & 15u) == 0) {
    *reinterpret_cast<uint4 *>(p + i) = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  } else {
    for (uint32_t k = i; k < min(n, i + 16); ++k) p[k] = 1;
  }
}
void fill_u32(Handle *h, uint32_t *p, uint32_t v, uint64_t n) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  const unsigned grid = (unsigned)(blocks < (uint64_t)h->sm_count * 16 ? blocks : (uint64_t)h->sm_count * 16);
  k_fill_u32<<<grid, 256, 0, h->stream>>>(p, v, n);
  h->launches += 1;
}
void publish_scalars(Handle *h) {
  k_publish_scalars<<<1, 1, 0, h->stream>>>(h->h_scalars, h->d_scalars);
  h->launches += 1;
}

int build_celllist(Handle *h, double *pos, uint8_t *alive, size_t n, int32_t *order_out,
                   size_t *n_alive_host, const ReorderSpec *reorder, bool presorted) {
  if (!h->domain_set) return set_error(h, ABR_ERR_STATE, "build: domain has not been set");
  if (n >= 0x7FFFFFFFull) return set_error(h, ABR_ERR_UNSUPPORTED, "build: n must fit in int32");
  const int D = h->D;
  h->built = false;
  if (n == 0) { // update_n == 0 -> return false (src/NeighbourSearchBase.h:375-376)
    // the reference keeps whatever bucket arrays it had (all ranges empty after
    // set_domain_impl's resize); give the query the same: empty buckets
    const uint64_t prod = h->grid().ncells;
    ABR_CUDA(h, h->bucket_begin.reserve(prod * sizeof(uint32_t)));
    ABR_CUDA(h, h->bucket_end.reserve(prod * sizeof(uint32_t)));
    ABR_CUDA(h, cudaMemsetAsync(h->bucket_begin.p, 0, prod * sizeof(uint32_t), h->stream));
    ABR_CUDA(h, cudaMemsetAsync(h->bucket_end.p, 0, prod * sizeof(uint32_t), h->stream));
    ABR_CUDA(h, h->keys[0].reserve(sizeof(uint32_t)));
    h->sorted_keys = h->keys[0].as<uint32_t>();
    h->ncells = prod;
    h->n_aliased = 0;
    h->n_alive_last = 0;
    h->built = true;
    if (n_alive_host) *n_alive_host = 0;
    return ABR_OK;
  }

  // The reference sizes the grid with the number of ALIVE particles
  // (m_alive_indices.size(), src/CellListOrdered.h:133), known only after
  // enforce_domain.  The grid does not influence which particles die, so k1 is
  // run with the grid computed from n and, if particles died and the recompute
  // rule picks a different grid for n_alive, once more (rare).
  size_t n_for_grid = n;
  const size_t saved_calc = h->size_calculated_with_n;
  uint32_t saved_size[MAXD];
  double saved_side[MAXD], saved_inv[MAXD];
  for (int d = 0; d < MAXD; ++d) {
    saved_size[d] = h->size[d];
    saved_side[d] = h->side[d];
    saved_inv[d] = h->inv_side[d];
  }

  for (int attempt = 0; attempt < 2; ++attempt) {
    set_domain_impl(h, n_for_grid);
    const Grid g = h->grid();
    uint64_t prod = 1;
    for (int d = 0; d < D; ++d) prod *= h->size[d];
    if (prod >= 0x7FFFFFF0ull || (uint64_t)g.key_bound >= 0xFFFFFFF0ull)
      return set_error(h, ABR_ERR_UNSUPPORTED, "build: too many buckets for 32-bit keys");
    prod = g.ncells; // buckets stored locally (the slab window when there is one)
    h->ncells = prod;

    const uint32_t n32 = (uint32_t)n;
    const uint32_t num_tiles = (n32 + RS_TILE - 1) / RS_TILE;
    for (int i = 0; i < 2; ++i) {
      ABR_CUDA(h, h->keys[i].reserve(n * sizeof(uint32_t)));
      ABR_CUDA(h, h->idx[i].reserve(n * sizeof(uint32_t)));
    }
    ABR_CUDA(h, h->tile_hist.reserve((size_t)RADIX * num_tiles * sizeof(uint32_t)));
    ABR_CUDA(h, h->bucket_begin.reserve(prod * sizeof(uint32_t)));
    ABR_CUDA(h, h->bucket_end.reserve(prod * sizeof(uint32_t)));

    // scalars: n_alive = n_incell = n by default (no dead / no overflow keys)
    DevScalars init;
    memset(&init, 0, sizeof(init));
    init.n_alive = n32;
    init.n_incell = n32;
    k_set_scalars<<<1, 1, 0, h->stream>>>(h->d_scalars, init);
    h->launches += 1;

    const unsigned gb = grid_for(n, 256);
    int bits = 1;
    while (bits < 32 && (g.key_bound >> bits) != 0) ++bits;
    // large sets: counting-sort build (abr_build2.cu) — keys, bucket ranges and the reorder of every column in three kernels
    const bool counting = !presorted && counting_build_applicable(h, n, bits, reorder, alive);
    bool two_level = false;
    const uint32_t *perm = nullptr;      // two-level: final position -> position in the binned copy
    const uint32_t *orig_tmp = nullptr;  // two-level: binned position -> original index
    GatherCols tmp_cols;                 // two-level: the binned copy of every column
    tmp_cols.ncols = 0;
    RecLayout rec{};                     // two-level: the binned copy as one record per particle (rec_aos)
    bool rec_aos = false;
    uint8_t *alive_dst = nullptr;        // two-level: the reordered alive column is filled, not moved
    if (counting) {
      const int rc = build_counting(h, pos, alive, n32, g, bits, reorder, order_out);
      if (rc) return rc;
    } else {
    uint32_t *keys0 = h->keys[0].as<uint32_t>();
    // LSD radix sort over the bits of key_bound; the key kernel also produces the tile
    // histograms of the first pass (two-level: of the most significant digit)
    const int passes = presorted ? 0 : (bits + 7) / 8; // adopt_sorted: keys only, no permutation
    two_level = reorder && !presorted && passes >= 2 && n >= h->two_level_min_n && reorder->ncols <= GP_MAXC - 1;
    uint32_t *hist = h->tile_hist.as<uint32_t>();
    uint32_t *first_hist = passes > 0 ? hist : nullptr;
    const int first_shift = two_level ? 8 * (passes - 1) : 0;
    static_assert(EK_TILE == RS_TILE, "k_enforce_key hands its histogram to the radix tiles");
    if (h->windowed) {
      switch (D) {
      case 2: k_enforce_key<2, true><<<num_tiles, EK_THREADS, 0, h->stream>>>(pos, alive, n32, g, keys0, h->d_scalars, first_hist, first_shift, num_tiles); break;
      default: k_enforce_key<3, true><<<num_tiles, EK_THREADS, 0, h->stream>>>(pos, alive, n32, g, keys0, h->d_scalars, first_hist, first_shift, num_tiles); break;
      }
    } else {
      switch (D) {
      case 1: k_enforce_key<1, false><<<num_tiles, EK_THREADS, 0, h->stream>>>(pos, alive, n32, g, keys0, h->d_scalars, first_hist, first_shift, num_tiles); break;
      case 2: k_enforce_key<2, false><<<num_tiles, EK_THREADS, 0, h->stream>>>(pos, alive, n32, g, keys0, h->d_scalars, first_hist, first_shift, num_tiles); break;
      default: k_enforce_key<3, false><<<num_tiles, EK_THREADS, 0, h->stream>>>(pos, alive, n32, g, keys0, h->d_scalars, first_hist, first_shift, num_tiles); break;
      }
    }
    h->launches += 1;

    ABR_CUDA(h, cudaFuncSetAttribute(k_radix_scatter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RadixScatterSmem)));
    int cur = 0;
    const TileTab dense{nullptr, nullptr, nullptr, nullptr, nullptr};
    GatherCols no_cols;
    no_cols.ncols = 0;
    if (two_level) {
      // ---- level 1: stable partition of whole records by the most significant digit ----
      const int top_shift = 8 * (passes - 1);
      size_t tmp_bytes = 0;
      for (int c = 0; c < reorder->ncols; ++c) tmp_bytes += ((reorder->elem_bytes[c] * n + 255) / 256) * 256;
      ABR_CUDA(h, h->tmp_cols.reserve(tmp_bytes));
      ABR_CUDA(h, h->idx2.reserve(n * sizeof(uint32_t)));
      const uint32_t t_bound = num_tiles + RADIX;
      ABR_CUDA(h, h->tile_tab.reserve((size_t)(4 * t_bound + RADIX + 8) * sizeof(uint32_t)));
      ABR_CUDA(h, h->seg_hist.reserve((size_t)RADIX * t_bound * sizeof(uint32_t)));
      GatherCols src_cols;
      {
        uint8_t *base = h->tmp_cols.as<uint8_t>();
        int m = 0;
        for (int c = 0; c < reorder->ncols; ++c) {
          // the alive column does not travel: every particle that survives the reorder is alive (the dead
          // ones sort behind n_alive), so its reordered copy is a run of ones (k_fill_ones_bounded below)
          if (alive && reorder->src[c] == static_cast<const void *>(alive) && reorder->elem_bytes[c] == 1 && h->skip_alive_move) {
            alive_dst = static_cast<uint8_t *>(reorder->dst[c]);
            continue;
          }
          src_cols.src[m] = static_cast<const uint8_t *>(reorder->src[c]);
          src_cols.dst[m] = base;
          src_cols.eb[m] = (uint32_t)reorder->elem_bytes[c];
          tmp_cols.src[m] = base;
          tmp_cols.dst[m] = static_cast<uint8_t *>(reorder->dst[c]);
          tmp_cols.eb[m] = (uint32_t)reorder->elem_bytes[c];
          base += ((reorder->elem_bytes[c] * n + 255) / 256) * 256;
          ++m;
        }
        src_cols.ncols = tmp_cols.ncols = m;
      }
      uint32_t *keys1 = h->keys[1].as<uint32_t>();
      uint32_t *orig = h->idx[1].as<uint32_t>();
      // (the histogram of the top digit came with the keys)
      cudaError_t e = device_scan<OpSum, false, 0>(h, hist, (uint64_t)RADIX * num_tiles, hist, nullptr);
      if (e != cudaSuccess) return check_cuda(h, e, "radix scan");
      // staged record move when every column window of a tile fits in shared memory next to the ranking state
      StageLayout stage{};
      size_t stage_bytes = 0;
      bool staged = h->stage_records;
      for (int c = 0; c < src_cols.ncols; ++c) {
        stage.off[c] = (uint32_t)stage_bytes;
        stage_bytes += ((size_t)RS_TILE * src_cols.eb[c] + 15) & ~(size_t)15;
        staged = staged && (((uintptr_t)src_cols.src[c]) & 15u) == 0;
      }
      stage.bar_off = (uint32_t)stage_bytes;
      const bool wide = h->stage_threads != RS_THREADS;
      const size_t rs_smem = wide ? sizeof(RadixScatterSmemT<RS2_THREADS / 32>) : sizeof(RadixScatterSmem);
      const size_t staged_smem = ((rs_smem + 15) & ~(size_t)15) + stage_bytes + 16;
      staged = staged && staged_smem <= (size_t)227 * 1024;
      rec_aos = staged && h->record_aos && plan_record_layout(tmp_cols, stage, &rec);
      if (rec_aos) {
        ABR_CUDA(h, h->tmp_cols.reserve(std::max(tmp_bytes, (size_t)n * rec.nwords * 8)));
        src_cols.dst[0] = h->tmp_cols.as<uint8_t>(); // the record array
        const size_t smem3 = ((sizeof(RadixScatterSmem) + 15) & ~(size_t)15) + stage_bytes + 16;
        ABR_CUDA(h, cudaFuncSetAttribute(k_radix_scatter<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        k_radix_scatter<3><<<num_tiles, RS_THREADS, smem3, h->stream>>>(keys0, nullptr, keys1, orig, hist, top_shift, n32, num_tiles, dense, src_cols,
                                                                        stage, rec);
      } else if (staged && wide) {
        ABR_CUDA(h, cudaFuncSetAttribute(k_radix_scatter<2, RS2_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem));
        k_radix_scatter<2, RS2_THREADS><<<num_tiles, RS2_THREADS, staged_smem, h->stream>>>(keys0, nullptr, keys1, orig, hist, top_shift, n32, num_tiles,
                                                                                            dense, src_cols, stage, rec);
      } else if (staged) {
        ABR_CUDA(h, cudaFuncSetAttribute(k_radix_scatter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem));
        k_radix_scatter<2><<<num_tiles, RS_THREADS, staged_smem, h->stream>>>(keys0, nullptr, keys1, orig, hist, top_shift, n32, num_tiles, dense,
                                                                              src_cols, stage, rec);
      } else {
        ABR_CUDA(h, cudaFuncSetAttribute(k_radix_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RadixScatterSmem)));
        k_radix_scatter<1><<<num_tiles, RS_THREADS, sizeof(RadixScatterSmem), h->stream>>>(keys0, nullptr, keys1, orig, hist, top_shift, n32,
                                                                                           num_tiles, dense, src_cols, stage, rec);
      }
      // tile table of the segmented passes
      uint32_t *tb = h->tile_tab.as<uint32_t>();
      TileTabW tw{tb, tb + t_bound, tb + 2 * t_bound, tb + 3 * t_bound, tb + 4 * t_bound, tb + 4 * t_bound + 8};
      k_tiletab_bins<<<1, RADIX, 0, h->stream>>>(hist, num_tiles, n32, tw);
      const TileTab seg{tw.start, tw.count, tw.hbase, tw.hstride, tw.total};
      h->launches += 2;
      // ---- level 2: LSD passes over the remaining digits, segmented by bin; they
      //      permute (key, binned position) pairs inside 1/256th of the array ----
      uint32_t *shist = h->seg_hist.as<uint32_t>();
      const uint32_t *kin = keys1;
      const uint32_t *iin = nullptr; // iota
      uint32_t *kbuf[2] = {h->keys[0].as<uint32_t>(), h->keys[1].as<uint32_t>()};
      uint32_t *ibuf[2] = {h->idx[0].as<uint32_t>(), h->idx2.as<uint32_t>()};
      int out = 0;
      for (int pass = 0; pass < passes - 1; ++pass) {
        const int shift = pass * 8;
        fill_u32(h, shist, 0u, (uint64_t)RADIX * t_bound);
        k_radix_hist<<<t_bound, RS_THREADS, 0, h->stream>>>(kin, n32, shift, num_tiles, shist, seg);
        e = device_scan<OpSum, false, 0>(h, shist, (uint64_t)RADIX * t_bound, shist, nullptr);
        if (e != cudaSuccess) return check_cuda(h, e, "segmented radix scan");
        k_radix_scatter<0><<<t_bound, RS_THREADS, sizeof(RadixScatterSmem), h->stream>>>(kin, iin, kbuf[out], ibuf[out], shist, shift, n32,
                                                                                              num_tiles, seg, no_cols, StageLayout{}, RecLayout{});
        h->launches += 2;
        kin = kbuf[out];
        iin = ibuf[out];
        out ^= 1;
      }
      h->sorted_keys = kin;
      perm = iin;
      orig_tmp = orig;
    } else {
      for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * 8;
        const uint32_t *kin = h->keys[cur].as<uint32_t>();
        const uint32_t *iin = pass == 0 ? nullptr : h->idx[cur].as<uint32_t>();
        uint32_t *kout = h->keys[cur ^ 1].as<uint32_t>();
        uint32_t *iout = (pass == passes - 1) ? reinterpret_cast<uint32_t *>(order_out)
                                              : h->idx[cur ^ 1].as<uint32_t>();
        if (pass > 0) k_radix_hist<<<num_tiles, RS_THREADS, 0, h->stream>>>(kin, n32, shift, num_tiles, hist, dense); // pass 0: came with the keys
        cudaError_t e = device_scan<OpSum, false, 0>(h, hist, (uint64_t)RADIX * num_tiles, hist, nullptr);
        if (e != cudaSuccess) return check_cuda(h, e, "radix scan");
        k_radix_scatter<0><<<num_tiles, RS_THREADS, sizeof(RadixScatterSmem), h->stream>>>(kin, iin, kout, iout, hist, shift, n32, num_tiles,
                                                                                                dense, no_cols, StageLayout{}, RecLayout{});
        h->launches += pass > 0 ? 2 : 1;
        cur ^= 1;
      }
      h->sorted_keys = h->keys[cur].as<uint32_t>();
    }

    // bucket ranges
    uint32_t *bb = h->bucket_begin.as<uint32_t>();
    uint32_t *be = h->bucket_end.as<uint32_t>();
    if (h->bounds_one_sweep) {
      // one sweep: every bucket written once (k_boundaries_fill); long runs of empty buckets through a list
      const uint32_t cap = (uint32_t)(prod / GAP_LONG + 2);
      ABR_CUDA(h, h->gap_list.reserve((size_t)(3 * cap + 4) * sizeof(uint32_t)));
      uint32_t *gp = h->gap_list.as<uint32_t>();
      const GapList gl{gp, gp + 4, gp + 4 + cap, gp + 4 + 2 * cap, cap};
      fill_u32(h, gp, 0u, 1);
      k_boundaries_fill<<<grid_for(((uint64_t)n + 4) / 4, 256), 256, 0, h->stream>>>(h->sorted_keys, n32, (uint32_t)prod, g.key_bound, bb, be,
                                                                               h->d_scalars, gl);
      k_fill_gaps<<<h->sm_count * 4, 256, 0, h->stream>>>(gl, bb, be);
      h->launches += 2;
    } else {
      fill_u32(h, bb, 0xFFFFFFFFu, prod);
      fill_u32(h, be, 0xFFFFFFFFu, prod);
      k_boundaries<<<gb, 256, 0, h->stream>>>(h->sorted_keys, n32, (uint32_t)prod, g.key_bound, bb, be, h->d_scalars);
      h->launches += 1;
      cudaError_t e = device_scan<OpMin, true, 1>(h, bb, prod, bb, be);
      if (e != cudaSuccess) return check_cuda(h, e, "bucket fill");
    }

    } // radix builds

    publish_scalars(h);
    if (reorder && !counting) {
      // Particles::reorder enqueued behind the build, bounded by the device-side
      // alive count: the only host round trip of update_positions is the final one
      int rc;
      if (two_level) {
        // the source of every output element lies in the same 1/256th of the binned
        // copy as the element itself: the random side of this gather is served by L2
        GatherCols fc = tmp_cols;
        fc.src[fc.ncols] = reinterpret_cast<const uint8_t *>(orig_tmp); // m_alive_indices: original index of every sorted particle
        fc.dst[fc.ncols] = reinterpret_cast<uint8_t *>(order_out);
        fc.eb[fc.ncols] = sizeof(uint32_t);
        fc.ncols += 1;
        GatherSlots gs;
        if (rec_aos)
          k_gather_records<<<gb, 256, 0, h->stream>>>(tmp_cols, rec, reinterpret_cast<const uint64_t *>(h->tmp_cols.as<uint8_t>()),
                                                      reinterpret_cast<uint32_t *>(order_out), perm, n32, &h->d_scalars->n_alive);
        else if (h->gather_slots && plan_gather_slots(fc, &gs))
          k_gather_slots<<<gb, 256, 0, h->stream>>>(fc, gs, perm, n32, &h->d_scalars->n_alive);
        else
          k_gather_fused<<<gb, 256, 0, h->stream>>>(fc, perm, n32, &h->d_scalars->n_alive);
        if (alive_dst) {
          k_fill_ones_bounded<<<grid_for((n + 15) / 16, 256), 256, 0, h->stream>>>(alive_dst, n32, &h->d_scalars->n_alive);
          h->launches += 1;
        }
        h->launches += 1;
        rc = ABR_OK;
      } else {
        h->gather_src_n = n;
        rc = gather_columns(h, reorder->ncols, reorder->src, reorder->dst, reorder->elem_bytes, order_out, n, &h->d_scalars->n_alive);
        h->gather_src_n = 0;
      }
      if (rc) return rc;
    }
    if (!n_alive_host && reorder && !presorted) {
      // ASYNCHRONOUS update (caller passed no n_alive_host): nothing is read back now.
      // The caller asserts that no particle dies; abr_check_async verifies it later.
      h->n_aliased = 0;
      h->n_alive_last = n;
      h->async_pending_n = n;
      h->built = true;
      return ABR_OK;
    }
    ABR_CUDA(h, cudaStreamSynchronize(h->stream));
    h->async_pending_n = 0;
    const size_t n_alive = h->h_scalars->n_alive;
    h->n_aliased = h->h_scalars->n_aliased;
    h->max_bucket = h->h_scalars->max_bucket;
    if (h->h_scalars->n_outside != 0)
      return set_error(h, ABR_ERR_INVALID, "build: " + std::to_string(h->h_scalars->n_outside) +
                                               " particle(s) lie outside this rank's slab window");
    if (presorted && h->h_scalars->n_unsorted != 0)
      return set_error(h, ABR_ERR_INVALID, "adopt_sorted: positions are not sorted by bucket");

    if (attempt == 0 && n_alive != n && !h->grid_forced) {
      // would the reference (which sees n_alive) have chosen another grid?
      uint32_t got_size[MAXD];
      for (int d = 0; d < MAXD; ++d) got_size[d] = h->size[d];
      h->size_calculated_with_n = saved_calc;
      for (int d = 0; d < MAXD; ++d) {
        h->size[d] = saved_size[d];
        h->side[d] = saved_side[d];
        h->inv_side[d] = saved_inv[d];
      }
      set_domain_impl(h, n_alive);
      bool same = true;
      for (int d = 0; d < D; ++d) same &= (got_size[d] == h->size[d]);
      if (!same) {
        // redo with the grid the reference would use; positions are already
        // wrapped and dead flags set, a second k1 pass is idempotent
        h->size_calculated_with_n = saved_calc;
        for (int d = 0; d < MAXD; ++d) {
          h->size[d] = saved_size[d];
          h->side[d] = saved_side[d];
          h->inv_side[d] = saved_inv[d];
        }
        n_for_grid = n_alive;
        continue;
      }
    }
    h->n_alive_last = n_alive;
    h->built = true;
    if (n_alive_host) *n_alive_host = n_alive;
    return ABR_OK;
  }
  return set_error(h, ABR_ERR_STATE, "build: grid did not converge");
}

// ---------------------------------------------------------------------------
// id map (neighbour_search_base::init_id_map and the id-map update of
// update_positions, src/NeighbourSearchBase.h:294-298, :440-486):
// m_id_map_key = ids of the particles in their (post-reorder) order,
// m_id_map_value = 0..n-1, sort_by_key(key, value).  The 64-bit ids are sorted
// with the same 8-bit LSD passes as the bucket keys, low word first, then (only
// if some id needs more than 32 bits) the high word gathered through the
// permutation; passes above the most significant byte in use are skipped.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_id_split(const uint64_t *__restrict__ ids, uint32_t n, uint32_t *__restrict__ lo,
                                                  unsigned long long *__restrict__ maxid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (i < n) {
    v = ids[i];
    lo[i] = (uint32_t)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, v, o);
    v = t > v ? t : v;
  }
  if ((threadIdx.x & 31) == 0 && v != 0) atomicMax(maxid, v);
}
__global__ void __launch_bounds__(256) k_id_hi_by_perm(const uint64_t *__restrict__ ids, const uint32_t *__restrict__ perm, uint32_t n,
                                                       uint32_t *__restrict__ hi) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) hi[k] = (uint32_t)(ids[perm ? perm[k] : k] >> 32);
}
__global__ void __launch_bounds__(256) k_id_emit(const uint64_t *__restrict__ ids, const uint32_t *__restrict__ perm, uint32_t n,
                                                 uint64_t *__restrict__ key, uint64_t *__restrict__ value) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    const uint32_t src = perm ? perm[k] : k;
    key[k] = ids[src];
    value[k] = src;
  }
}
// src/CellListOrdered.h:379-388 find(id): lower_bound over the sorted keys
__global__ void __launch_bounds__(256) k_id_find(const uint64_t *__restrict__ key, const uint64_t *__restrict__ value, uint64_t n,
                                                 const uint64_t *__restrict__ query, uint64_t m, uint64_t *__restrict__ out) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const uint64_t id = query[q];
  uint64_t lo = 0, len = n;
  while (len > 0) { // std::lower_bound
    const uint64_t half = len >> 1;
    if (key[lo + half] < id) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  out[q] = (lo != n && !(id < key[lo])) ? value[lo] : n;
}

// stable LSD passes over the bytes [0, nbytes) of a 32-bit key array, carrying a
// permutation (null = identity on entry); returns the buffers holding the result
static int lsd_passes_u32(Handle *h, uint32_t *kbuf[2], uint32_t *ibuf[2], int &cur, bool &have_perm, uint32_t n32, int nbytes) {
  const uint32_t num_tiles = (n32 + RS_TILE - 1) / RS_TILE;
  uint32_t *hist = h->tile_hist.as<uint32_t>();
  const TileTab dense{nullptr, nullptr, nullptr, nullptr, nullptr};
  GatherCols no_cols;
  no_cols.ncols = 0;
  ABR_CUDA(h, cudaFuncSetAttribute(k_radix_scatter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RadixScatterSmem)));
  for (int pass = 0; pass < nbytes; ++pass) {
    const int shift = pass * 8;
    k_radix_hist<<<num_tiles, RS_THREADS, 0, h->stream>>>(kbuf[cur], n32, shift, num_tiles, hist, dense);
    cudaError_t e = device_scan<OpSum, false, 0>(h, hist, (uint64_t)RADIX * num_tiles, hist, nullptr);
    if (e != cudaSuccess) return check_cuda(h, e, "id map scan");
    k_radix_scatter<0><<<num_tiles, RS_THREADS, sizeof(RadixScatterSmem), h->stream>>>(kbuf[cur], have_perm ? ibuf[cur] : nullptr, kbuf[cur ^ 1],
                                                                                            ibuf[cur ^ 1], hist, shift, n32, num_tiles, dense, no_cols, StageLayout{}, RecLayout{});
    h->launches += 2;
    cur ^= 1;
    have_perm = true;
  }
  ABR_CUDA(h, cudaGetLastError());
  return ABR_OK;
}

int build_id_map(Handle *h, const uint64_t *ids, size_t n) {
  if (n >= 0xFFFFFFF0ull) return set_error(h, ABR_ERR_UNSUPPORTED, "id map: too many particles");
  h->id_map_n = 0;
  if (n == 0) return ABR_OK;
  if (!ids) return set_error(h, ABR_ERR_INVALID, "id map: null id column");
  const uint32_t n32 = (uint32_t)n;
  const uint32_t num_tiles = (n32 + RS_TILE - 1) / RS_TILE;
  for (int i = 0; i < 2; ++i) {
    ABR_CUDA(h, h->idm_k[i].reserve(n * sizeof(uint32_t)));
    ABR_CUDA(h, h->idm_i[i].reserve(n * sizeof(uint32_t)));
  }
  ABR_CUDA(h, h->tile_hist.reserve((size_t)RADIX * num_tiles * sizeof(uint32_t)));
  ABR_CUDA(h, h->id_map_key.reserve(n * sizeof(uint64_t)));
  ABR_CUDA(h, h->id_map_value.reserve(n * sizeof(uint64_t)));
  ABR_CUDA(h, h->idm_max.reserve(sizeof(unsigned long long)));
  uint32_t *kbuf[2] = {h->idm_k[0].as<uint32_t>(), h->idm_k[1].as<uint32_t>()};
  uint32_t *ibuf[2] = {h->idm_i[0].as<uint32_t>(), h->idm_i[1].as<uint32_t>()};
  unsigned long long *dmax = h->idm_max.as<unsigned long long>();
  fill_u32(h, reinterpret_cast<uint32_t *>(dmax), 0u, 2);
  const unsigned gb = grid_for(n, 256);
  k_id_split<<<gb, 256, 0, h->stream>>>(ids, n32, kbuf[0], dmax);
  h->launches += 1;
  unsigned long long maxid = 0; // not on the hot path: a plain read-back decides how many byte passes are needed
  ABR_CUDA(h, cudaMemcpyAsync(&maxid, dmax, sizeof(maxid), cudaMemcpyDeviceToHost, h->stream));
  ABR_CUDA(h, cudaStreamSynchronize(h->stream));
  auto bytes_of = [](uint32_t v) {
    int b = 0;
    while (v) {
      ++b;
      v >>= 8;
    }
    return b;
  };
  const uint32_t max_hi = (uint32_t)(maxid >> 32);
  const int lo_bytes = max_hi ? 4 : bytes_of((uint32_t)maxid);
  int cur = 0;
  bool have_perm = false;
  int rc = lsd_passes_u32(h, kbuf, ibuf, cur, have_perm, n32, lo_bytes);
  if (rc) return rc;
  if (max_hi) {
    k_id_hi_by_perm<<<gb, 256, 0, h->stream>>>(ids, have_perm ? ibuf[cur] : nullptr, n32, kbuf[cur]);
    h->launches += 1;
    rc = lsd_passes_u32(h, kbuf, ibuf, cur, have_perm, n32, bytes_of(max_hi));
    if (rc) return rc;
  }
  k_id_emit<<<gb, 256, 0, h->stream>>>(ids, have_perm ? ibuf[cur] : nullptr, n32, h->id_map_key.as<uint64_t>(), h->id_map_value.as<uint64_t>());
  h->launches += 1;
  ABR_CUDA(h, cudaGetLastError());
  h->id_map_n = n;
  return ABR_OK;
}

int find_ids(Handle *h, const uint64_t *query, size_t m, uint64_t *index_out) {
  if (m == 0) return ABR_OK;
  if (!query || !index_out) return set_error(h, ABR_ERR_INVALID, "id find: null pointer");
  k_id_find<<<grid_for(m, 256), 256, 0, h->stream>>>(h->id_map_key.as<uint64_t>(), h->id_map_value.as<uint64_t>(), (uint64_t)h->id_map_n, query,
                                                     (uint64_t)m, index_out);
  h->launches += 1;
  ABR_CUDA(h, cudaGetLastError());
  return ABR_OK;
}

} // namespace abr
