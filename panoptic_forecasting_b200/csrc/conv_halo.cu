// Persistent halo-tile tcgen05 kernel for the 3x3 / stride-1 ConvLayers (hardnet.py:16-25; 56 of
// the 70 convs).
//
// A CTA loops over 16-row x 8-column blocks of output pixels.  Per input-channel chunk (16, 32 or
// 64 channels of one channel slice) ONE TMA box of 18 x 10 input pixels is staged per bf16 plane
// and all nine filter taps read it through shifted UMMA descriptors: the 8 pixels of an
// accumulator row group are consecutive swizzled rows of the box, row groups are 10 rows apart
// (SBO = 10 * row pitch), and tap (dy,dx) moves the descriptor start by (dy*10+dx) rows.  (The
// tensor core applies the 32/64/128-byte swizzle to absolute shared-memory address bits, so a
// shifted start needs no re-layout -- verified on B200 against the fp32 kernels.)  Versus one box
// per tap this cuts L2->SMEM activation traffic ~6x; image borders are TMA out-of-bounds zero
// fill == the conv's zero padding.
// Weights (split bf16, K-major) are either RESIDENT in shared memory for the whole kernel (the
// high-resolution layers, where a CTA visits many tiles) or STREAMED per (chunk, tap) through a
// second mbarrier ring.  The fp32 accumulator is double-buffered in TMEM so the epilogue of tile
// i overlaps the MMAs of tile i+1.
// Warp roles: warp 0 TMA producer, warps 1 and 6 MMA issuers (alternate chunks; warp 1 owns TMEM), warps 2..5 the
// epilogue team (TMEM reads, bias / ReLU / pool / interpolation, split, stores).  Instantiations
// conv_halo_kernel<threads, mode, add, lean>:
//   224 threads, mode 0   one epilogue team, every epilogue path (folded / whole row in registers / wide)
//   352 threads, mode 1   a second team (warps 7..10) takes every other 16-channel group of the same accumulator rows
//                         (N tiles >= 32: the 1x1 transition / conv1x1_up layers, base.2, base.3)
//   352 threads, mode 3   two teams on ALTERNATE TILES, team k owns accumulator buffer k (base.1 and the single-chunk
//                         folded layers: a tile is a dozen MMAs, the epilogue is the whole cost)
//   608 threads, mode 1   four teams for N tiles >= 64 (no additive term: 86 registers)
//   352 threads, mode 2   A/B only: two teams on the two groups of a folded N = 32 layer
//   add   = the additive term of the fused conv1x1_up layers (only <352, 1, true, true>)
//   lean  = the common ConvLayer epilogue: ReLU + split-bf16 store, none of the pool / space-to-depth / fp32 switches
#include <atomic>

#include <cuda.h>
#include <cuda_bf16.h>

#include "bgnet.h"
#include "conv_tc.h"
#include "split_bf16.cuh"
#include "tc_common.cuh"

namespace pf {

using namespace tc;

namespace {
// Phase timestamps / wait accounting / ablations (tools/halo_ts.py) are a BUILD option: -DPF_HALO_DBG.  As a run-time flag
// they cost every tile of every epilogue warp an S2R, a 64-bit clock read and their scoreboard stalls (ncu source view of
// conv1x1_up.3: 5 % of the samples) plus a dozen registers in instantiations that spill.
#ifdef PF_HALO_DBG
constexpr bool kHaloDbg = true;
#else
constexpr bool kHaloDbg = false;
#endif
constexpr int kMaxA = 8, kMaxB = 8;
constexpr int kHaloThreads = 224, kHaloThreads8 = 352, kHaloThreads16 = 608;   // 1 / 2 / 4 epilogue teams

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t layout_of(int w) { return w == 64 ? 2u : (w == 32 ? 4u : 6u); }
__device__ __forceinline__ uint32_t align1k(uint32_t x) { return (x + 1023u) & ~1023u; }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// t / d for 0 <= t < 2^24 (tile indices) with a precomputed float reciprocal: one multiply, one conversion, one fix-up
// (a 32-bit integer division by a run-time divisor is ~20 instructions, and every epilogue thread did two per tile)
__device__ __forceinline__ int fast_div(int t, int d, float inv_d) {
  int q = __float2int_rz((float)t * inv_d);
  const int r = t - q * d;
  q += (r >= d) ? 1 : 0;
  q -= (r < 0) ? 1 : 0;
  return q;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
}  // namespace

// All MMAs of one staged activation chunk: KS x KS filter taps x NK K-atoms x 2 MMAs.
// Everything that shapes the instruction stream is a template parameter (chunk width W, atoms NK, taps, weight
// residency), the dy loop stays rolled and the dx / atom loops are unrolled with immediate descriptor offsets.
// The issuing warp is a single instruction stream feeding a tensor pipe that retires one of these small-N MMAs
// every ~40 cycles (shared-memory operand fetch: (4 KB of A + N x 32 B of B) / 128 B per clock, measured), so
// its code has to be branch-free and small enough to stay in the instruction cache: the earlier fully unrolled,
// runtime-predicated version spent 38% of its issue slots in instruction-fetch stalls (ncu source view).
struct ChunkIssue {
  uint32_t d, idesc1, idesc2, el;
  uint32_t sa, a_tile;          // A stage: hi plane at sa, lo plane at sa + a_tile
  uint32_t smem_base, ntile4;             // resident weights: tap tiles [hi rows ; lo rows] are align1k(4 * ntile * W) bytes apart
  uint32_t b_base, b_tile, bar_full_b0, bar_empty_b0;
  int SB;
  uint32_t mc;                  // HaloLayer::cluster: the streamed weight ring is shared by the CTA pair (see the producer)
};

// position in the streamed-weight ring: stage index and the parity its full barrier is waited on.  Kept incrementally:
// `ib % SB` and `ib / SB` with a run-time SB are ~40 dependent instructions, and they sat in the issuing thread between
// the MMAs of consecutive taps (the tensor pipe idles while its issuer computes, see mma_warp_loop)
struct BCursor { int idx; uint32_t ph; };
__device__ __forceinline__ void bcur_advance(BCursor& b, int n, int SB) {
  b.idx += n;
  while (b.idx >= SB) { b.idx -= SB; b.ph ^= 1u; }
}

template <int W, int NK, int KS, bool RES>
__device__ __forceinline__ void issue_chunk(const ChunkIssue& c, uint32_t acc_first, uint32_t& woff, BCursor& ib) {
  constexpr int HX = KS == 1 ? 8 : 10;
  constexpr uint32_t RP = 2u * W;
  constexpr uint32_t LAY = W == 64 ? 2u : (W == 32 ? 4u : 6u);
  constexpr uint32_t A_HI = ((HX * RP) >> 4) | (1u << 14) | (LAY << 29);   // SBO | version | swizzle (upper descriptor word)
  constexpr uint32_t B_HI = ((8u * RP) >> 4) | (1u << 14) | (LAY << 29);
  uint32_t a_hi_lo = ((c.sa & 0x3FFFFu) >> 4) | (1u << 16);
  uint32_t a_lo_lo = (((c.sa + c.a_tile) & 0x3FFFFu) >> 4) | (1u << 16);
  uint32_t accv = acc_first;
#pragma unroll 1
  for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
    for (int dx = 0; dx < KS; ++dx) {
      uint32_t sb;
      int sbi = 0;
      if (RES) {
        sb = c.smem_base + woff;
        woff += align1k(c.ntile4 * (uint32_t)W);
      } else {
        sbi = ib.idx;
        mbar_wait(c.bar_full_b0 + 8u * sbi, ib.ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        sb = c.b_base + sbi * c.b_tile;
        if (++ib.idx == c.SB) { ib.idx = 0; ib.ph ^= 1u; }
      }
      const uint32_t b_lo = ((sb & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
      for (int ka = 0; ka < NK; ++ka) {
        const uint32_t shift = (uint32_t)(dx * (int)(RP >> 4) + 2 * ka);   // immediate
        umma_bf16_w(c.d, a_hi_lo + shift, A_HI, b_lo + 2 * ka, B_HI, c.idesc2, (dx == 0 && ka == 0) ? accv : 1u, c.el);
        umma_bf16_w(c.d, a_lo_lo + shift, A_HI, b_lo + 2 * ka, B_HI, c.idesc1, 1u, c.el);
      }
      if (!RES) {
        // the stage goes back to BOTH producers of a CTA pair: each of them writes its half of the next tile into both CTAs
        if (c.mc) umma_commit_mc_p(c.bar_empty_b0 + 8u * sbi, (uint16_t)3, c.el);
        else umma_commit_p(c.bar_empty_b0 + 8u * sbi, c.el);
      }
    }
    a_hi_lo += (uint32_t)(HX * (int)(RP >> 4));
    a_lo_lo += (uint32_t)(HX * (int)(RP >> 4));
    accv = 1u;
  }
}

// chunk table entry: chunk width | atoms << 8 | slice << 12 | first channel << 16
__host__ __device__ __forceinline__ uint32_t chunk_code(int w, int nk, int seg, int c0) {
  return (uint32_t)w | ((uint32_t)nk << 8) | ((uint32_t)seg << 12) | ((uint32_t)c0 << 16);
}

template <int KS, bool RES>
__device__ __forceinline__ void issue_dispatch(uint32_t code, const ChunkIssue& ci, uint32_t accf, uint32_t& woff, BCursor& ib) {
  switch (code & 0xFFFu) {
    case 64u | (4u << 8): issue_chunk<64, 4, KS, RES>(ci, accf, woff, ib); break;
    case 64u | (3u << 8): issue_chunk<64, 3, KS, RES>(ci, accf, woff, ib); break;
    case 64u | (2u << 8): issue_chunk<64, 2, KS, RES>(ci, accf, woff, ib); break;
    case 64u | (1u << 8): issue_chunk<64, 1, KS, RES>(ci, accf, woff, ib); break;
    case 32u | (2u << 8): issue_chunk<32, 2, KS, RES>(ci, accf, woff, ib); break;
    case 32u | (1u << 8): issue_chunk<32, 1, KS, RES>(ci, accf, woff, ib); break;
    default: issue_chunk<16, 1, KS, RES>(ci, accf, woff, ib); break;
  }
}

// dx-folded 3x3 chunk (HaloLayer::fold): the three taps of a filter row share ONE pair of MMAs.  The box is
// 10 rows x 16 columns, accumulator row m = box pixel (m >> 4, m & 15) shifted down by dy rows (contiguous:
// SBO = 8 rows), and the B tile of (chunk, dy) stacks the dx = 0,1,2 weight rows: [hi0; hi1; hi2; lo0; lo1; lo2].
//   D'[:, 0:6n] += A_hi(dy) * B^T (N = 6n)      D'[:, 0:3n] += A_lo(dy) * [hi0; hi1; hi2]^T (N = 3n)
// D'[(y, x'), dx block] is the contribution of input column x' to output column x' + 1 - dx; the epilogue adds the
// three blocks with a one-lane shift.  6 instead of 18 A fetches per K atom: for n = 16 an MMA pair costs
// 56 + 44 cycles (N = 96 / 48) where three pairs cost 3 x (40 + 39).
template <int W, int NK>
__device__ __forceinline__ void issue_chunk_fold(const ChunkIssue& c, uint32_t acc_first, uint32_t& woff) {
  constexpr uint32_t RP = 2u * W;
  constexpr uint32_t LAY = W == 64 ? 2u : (W == 32 ? 4u : 6u);
  constexpr uint32_t D_HI = ((8u * RP) >> 4) | (1u << 14) | (LAY << 29);
  const uint32_t a_hi_lo = ((c.sa & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t a_lo_lo = (((c.sa + c.a_tile) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const uint32_t sb = c.smem_base + woff;
    woff += align1k(3u * c.ntile4 * (uint32_t)W);
    const uint32_t b_lo = ((sb & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
    for (int ka = 0; ka < NK; ++ka) {
      const uint32_t shift = (uint32_t)(dy * (int)RP + 2 * ka);     // 16 box pixels per filter row, immediate
      umma_bf16_w(c.d, a_hi_lo + shift, D_HI, b_lo + 2 * ka, D_HI, c.idesc2, (dy == 0 && ka == 0) ? acc_first : 1u, c.el);
      umma_bf16_w(c.d, a_lo_lo + shift, D_HI, b_lo + 2 * ka, D_HI, c.idesc1, 1u, c.el);
    }
  }
}

__device__ __forceinline__ void issue_dispatch_fold(uint32_t code, const ChunkIssue& ci, uint32_t accf, uint32_t& woff) {
  switch (code & 0xFFFu) {
    case 64u | (4u << 8): issue_chunk_fold<64, 4>(ci, accf, woff); break;
    case 64u | (3u << 8): issue_chunk_fold<64, 3>(ci, accf, woff); break;
    case 64u | (2u << 8): issue_chunk_fold<64, 2>(ci, accf, woff); break;
    case 64u | (1u << 8): issue_chunk_fold<64, 1>(ci, accf, woff); break;
    case 32u | (2u << 8): issue_chunk_fold<32, 2>(ci, accf, woff); break;
    case 32u | (1u << 8): issue_chunk_fold<32, 1>(ci, accf, woff); break;
    default: issue_chunk_fold<16, 1>(ci, accf, woff); break;
  }
}

struct MmaCtx {
  uint32_t el, tmem_d, smem_base, a_base, a_tile, b_base, b_tile, bar0, idesc1, idesc2;
  int total_tiles, rounds;
  bool dbg;
};

// Two MMA warps take alternate chunks.  tcgen05.mma issue is effectively synchronous with the tensor pipe (a
// ~100-cycle pause of the issuing thread idles the pipe for ~100 cycles: measured), so everything that is not
// an MMA -- mbarrier waits, fences, descriptor set-up, the commits -- has to happen in ANOTHER warp while this
// one issues.  Each warp prepares its next chunk, then waits for its turn on a named barrier; the turn passes
// right after the other warp's last MMA of the previous chunk, which keeps the issue order (and so the
// accumulation order) identical to a single issuing warp.
__device__ __forceinline__ void turn_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void turn_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

template <int KS, bool RES, bool FOLD = false>
__device__ __forceinline__ void mma_warp_loop(const HaloLayer& L, const MmaCtx& m, const int role) {
  const uint32_t el = m.el;
  const int SA = L.stages_a, nchunk = L.nchunk, ntile = L.ntile;
  const uint32_t full_a0 = m.bar0, empty_a0 = m.bar0 + 8u * kMaxA;
  const uint32_t tmem_full0 = m.bar0 + 8u * (2 * kMaxA + 2 * kMaxB + 1), tmem_empty0 = tmem_full0 + 16u;
  ChunkIssue ci;
  ci.idesc1 = m.idesc1; ci.idesc2 = m.idesc2; ci.el = el; ci.a_tile = m.a_tile;
  ci.smem_base = m.smem_base; ci.ntile4 = (uint32_t)ntile * 4u;
  ci.b_base = m.b_base; ci.b_tile = m.b_tile;
  ci.bar_full_b0 = m.bar0 + 8u * (2 * kMaxA); ci.bar_empty_b0 = m.bar0 + 8u * (2 * kMaxA + kMaxB); ci.SB = L.stages_b;
  ci.mc = L.cluster ? 1u : 0u;
  const int my_tiles = m.rounds;
  const int last_ia = my_tiles * nchunk - 1;
  int ia = 0, st = 0, acc = 0;
  BCursor ib = {0, 0u};
  uint32_t ph_a = 0, ph_t = 1;
  long long wait_full = 0, wait_tmem = 0, w0 = 0;
  const bool dbg = kHaloDbg && m.dbg && role == 0;
  for (int tl = 0; tl < my_tiles; ++tl) {
    if ((ia & 1) == role) {                   // owner of the tile's first chunk: the accumulator buffer must be drained
      if (dbg) w0 = clock64();
      mbar_wait(tmem_empty0 + 8u * acc, ph_t);
      if (dbg) wait_tmem += clock64() - w0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    ci.d = m.tmem_d + (uint32_t)(acc * (FOLD ? 6 : 2) * ntile);
    uint32_t woff = 0;
    for (int j = 0; j < nchunk; ++j, ++ia) {
      const uint32_t code = L.chunk[j];
      if ((ia & 1) == role) {
        if (dbg) w0 = clock64();
        mbar_wait(full_a0 + 8u * st, ph_a);
        if (dbg) wait_full += clock64() - w0;
        if (dbg && el && ia == 0) L.dbg_ts[3] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        ci.sa = m.a_base + st * 2 * m.a_tile;
        if (ia > 0) turn_wait(1 + role);
        if (dbg && el && ia < 48) L.dbg_ts[40 + (ia >> 1)] = clock64();
        if (FOLD) issue_dispatch_fold(code, ci, j == 0 ? 0u : 1u, woff);
        else issue_dispatch<KS, RES>(code, ci, j == 0 ? 0u : 1u, woff, ib);
        if (ia < last_ia) turn_pass(2 - role);
        if (dbg && el && ia < 48) L.dbg_ts[64 + (ia >> 1)] = clock64();
        umma_commit_p(empty_a0 + 8u * st, el);
        if (j + 2 >= nchunk) umma_commit_p(tmem_full0 + 8u * acc, el);     // this warp's last chunk of the tile
        if (dbg && el && tl == 0 && j + 1 == nchunk) L.dbg_ts[4] = clock64();
      } else {
        // the other warp's chunk: keep the weight cursor / ring counters in step
        const uint32_t w = code & 0xFFu;
        if (FOLD) woff += 3u * align1k(3u * ci.ntile4 * w);
        else if (RES) woff += (uint32_t)(KS * KS) * align1k(ci.ntile4 * w);
        else bcur_advance(ib, KS * KS, ci.SB);
      }
      if (++st == SA) { st = 0; ph_a ^= 1u; }
    }
    if (acc) ph_t ^= 1u;
    acc ^= 1;
  }
  if (dbg && el) { L.dbg_ts[7] = clock64(); L.dbg_ts[9] = wait_full; L.dbg_ts[10] = wait_tmem; }
}

// MODE 0: every epilogue path (one team); 1: wide path only (2 / 4 teams); 2: folded path only, two teams;
// 3: the N <= 32 paths (folded / whole row in registers) with two teams that take ALTERNATE TILES (HaloLayer::alt): team k
// owns accumulator buffer k, so the epilogue of tile i + 1 runs next to that of tile i instead of behind it
// ADD: the epilogue carries the additive term of the fused conv1x1_up layers (only <352, 1> is instantiated with it: its
// eight source addresses and interpolation weights otherwise sit in the registers of every wide layer)
// LEAN: the epilogue of the common ConvLayer (ReLU, split-bf16 output at the layer's own resolution: no pooling, no
// space-to-depth, no fp32 / argmax output) -- those run-time switches were a dozen uniform branches per 16-channel group,
// and the epilogue warps lost 12-16 % of their issue slots to instruction fetch on them (ncu source view, stall_no_inst)
template <int THREADS, int MODE, bool ADD = false, bool LEAN = false>
__global__ void __launch_bounds__(THREADS, 1) conv_halo_kernel(const HaloLayer L, const CUtensorMap* __restrict__ maps) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxA + 2 * kMaxB + 9];
  __shared__ uint32_t tmem_base_smem;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const bool dbg = kHaloDbg && L.dbg_ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  if (dbg && threadIdx.x == 0) L.dbg_ts[0] = clock64();
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int ntile = L.ntile;
  const uint32_t bar0 = smem_u32(bars);
  auto full_a = [&](int s) { return bar0 + 8u * s; };
  auto empty_a = [&](int s) { return bar0 + 8u * (kMaxA + s); };
  auto full_b = [&](int s) { return bar0 + 8u * (2 * kMaxA + s); };
  auto empty_b = [&](int s) { return bar0 + 8u * (2 * kMaxA + kMaxB + s); };
  const uint32_t wbar = bar0 + 8u * (2 * kMaxA + 2 * kMaxB);
  auto tmem_full = [&](int a) { return bar0 + 8u * (2 * kMaxA + 2 * kMaxB + 1 + a); };
  auto tmem_empty = [&](int a) { return bar0 + 8u * (2 * kMaxA + 2 * kMaxB + 3 + a); };
  auto full_p = [&](int a) { return bar0 + 8u * (2 * kMaxA + 2 * kMaxB + 5 + a); };     // additive-term patch (see HaloLayer::add_pbytes)
  auto empty_p = [&](int a) { return bar0 + 8u * (2 * kMaxA + 2 * kMaxB + 7 + a); };

  const int SA = L.stages_a, SB = L.stages_b;
  const uint32_t a_tile = L.a_tile_bytes;          // one plane of one A stage (1024-aligned)
  const uint32_t b_tile = L.b_tile_bytes;          // one tap of the widest chunk: [hi rows ; lo rows], 1024-aligned
  const uint32_t w_region = L.resident ? L.w_bytes_total : 0u;
  const uint32_t a_base = smem_base + w_region;
  const uint32_t b_base = a_base + SA * 2 * a_tile;
  const uint32_t p_base = b_base + (L.resident ? 0u : (uint32_t)SB * b_tile);   // two patch buffers behind the rings
  const uint32_t p_tx = (uint32_t)(kAddPH * kAddPW) * (uint32_t)(ntile + 4) * 4u;
  const int n0 = blockIdx.y * ntile;
  const int tiles_per_img = L.tiles_x * L.tiles_y;
  const int total_tiles = tiles_per_img * L.batch;
  const float inv_tpi = 1.0f / (float)tiles_per_img, inv_tx = 1.0f / (float)L.tiles_x;
  // CTA pair (HaloLayer::cluster, streamed weights): both CTAs run the SAME number of rounds, because every weight tile
  // is loaded half by each of them into both; the odd CTA's last round may be a ghost tile (computed, not stored)
  const uint32_t crank = L.cluster ? cluster_ctarank() : 0u;
  const int bx0 = L.cluster ? ((int)blockIdx.x & ~1) : (int)blockIdx.x;
  const int rounds = (total_tiles - bx0 + (int)gridDim.x - 1) / (int)gridDim.x;
  const int HX = L.hx, HY = L.hy, NT = L.taps;     // 3x3: 10 x 18 box, 9 taps; 1x1: 8 x 16 box, 1 tap
  const int org = NT == 9 ? 1 : 0;                  // box origin = tile origin - pad

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), L.cluster ? 2 : 1); }   // pair: both CTAs' MMA warps release a stage
    mbar_init(wbar, 1);
    // both MMA warps commit to tmem_full after their last chunk of a tile (a commit only covers the issuing thread's MMAs)
    const uint32_t nepi = MODE == 3 ? 4u : 4u * (uint32_t)(L.epi8 ? L.epi8 : 1);   // epilogue warps that hand a buffer back (epi8 = teams)
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), L.nchunk >= 2 ? 2 : 1); mbar_init(tmem_empty(a), nepi); }
    for (int a = 0; a < 2; ++a) { mbar_init(full_p(a), 1); mbar_init(empty_p(a), nepi); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (L.cluster) cluster_sync_all();       // the peer's barriers exist before anything is multicast to them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_smem;
  if (dbg && threadIdx.x == 0) L.dbg_ts[1] = clock64();
  // Programmatic dependent launch: let the next layer's CTAs start their prologue (barriers, TMEM,
  // weight loads -- nothing that depends on this layer) on idle SMs right away ...
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loops, one elected lane issues) =====
    const uint32_t el = elect_one();
    {
      if (L.resident) {
        mbar_expect_tx_p(wbar, L.w_tx_total, el);
        uint32_t off = 0;
        for (int s = 0; s < L.nseg; ++s) {
          const int cpad = L.seg_cpad[s], w = L.seg_w[s];
          const CUtensorMap* wm = maps + L.w_map[w >> 5];          // 16 -> 0, 32 -> 1, 64 -> 2
          const uint32_t bt = (uint32_t)ntile * 2u * w;            // hi rows, lo rows directly behind them
          for (int c0 = 0; c0 < cpad; c0 += w) {
            if (L.fold) {
              // per filter row one B tile: [hi(dx=0); hi(1); hi(2); lo(0); lo(1); lo(2)], ntile rows each
              for (int dy = 0; dy < 3; ++dy) {
                for (int half = 0; half < 2; ++half)
                  for (int dx = 0; dx < 3; ++dx)
                    tma_load_2d_p(smem_base + off + (uint32_t)(half * 3 + dx) * bt, wm + half, wbar,
                                  L.seg_koff[s] + (dy * 3 + dx) * cpad + c0, n0, el);
                off += align1k(6 * bt);
              }
              continue;
            }
            for (int tap = 0; tap < NT; ++tap) {
              if (!((L.tap_mask >> tap) & 1)) continue;
              const int koff = L.seg_koff[s] + tap * cpad + c0;
              tma_load_2d_p(smem_base + off, wm, wbar, koff, n0, el);
              tma_load_2d_p(smem_base + off + bt, wm + 1, wbar, koff, n0, el);
              off += align1k(2 * bt);
            }
          }
        }
      }
      // ... and do not read the previous layer's activations before it has completed and flushed.
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int ia = 0, sb = 0, st = 0, tp = 0;
      uint32_t ph_b = 1;                      // same bookkeeping for the streamed-weight ring
      uint32_t ph_a = 1;                      // parity to wait on for "stage free"; flips when the ring wraps
      long long wait_acc = 0;
      const int nchunk = L.nchunk;
      for (int rd = 0, t0 = blockIdx.x; rd < rounds; ++rd, t0 += gridDim.x) {
        const int t = t0 < total_tiles ? t0 : total_tiles - 1;      // ghost round of a CTA pair: any valid tile
        const int img = fast_div(t, tiles_per_img, inv_tpi);
        const int r = t - img * tiles_per_img;
        const int ty = fast_div(r, L.tiles_x, inv_tx), tx = r - ty * L.tiles_x;
        const int y0 = L.fold ? ty * 8 : ty * 16, x0 = L.fold ? tx * 14 : tx * 8;
        if (L.add_pbytes) {
          // the low-resolution patch this tile's epilogue interpolates from (origin = source pixel of the tile's first
          // row / column, same float arithmetic as the epilogue); out-of-image rows / columns are zero-filled, never read
          const int pa = tp & 1;
          mbar_wait(empty_p(pa), ((tp >> 1) & 1) ^ 1);
          mbar_expect_tx_p(full_p(pa), p_tx, el);
          tma_load_4d_p(p_base + (uint32_t)pa * L.add_pbytes, maps + L.add_map, full_p(pa), n0,
                        (int)(L.add_sw * (float)x0), (int)(L.add_sh * (float)y0), img, el);
          ++tp;
        }
        for (int j = 0; j < nchunk; ++j, ++ia) {
          const uint32_t code = L.chunk[j];
          const int w = (int)(code & 0xFFu), s = (int)((code >> 12) & 0xFu), c0 = (int)(code >> 16);
          const CUtensorMap* am = maps + L.seg_map[s];
          const uint32_t a_tx = (uint32_t)(HX * HY) * 2u * w;
          long long w0 = 0;
          if (dbg) w0 = clock64();
          mbar_wait(empty_a(st), ph_a);
          if (dbg) wait_acc += clock64() - w0;
          if (dbg && el && ia < 24) L.dbg_ts[16 + ia] = clock64();
          const uint32_t sa = a_base + st * 2 * a_tile;
          if (kHaloDbg && (L.dbg_mode & 1) && ia >= SA) {
            if (el) mbar_arrive(full_a(st));
            __syncwarp();
          } else {
            mbar_expect_tx_p(full_a(st), 2 * a_tx, el);
            tma_load_4d_p(sa, am, full_a(st), c0, x0 - org, y0 - org, img, el);
            tma_load_4d_p(sa + a_tile, am + 1, full_a(st), c0, x0 - org, y0 - org, img, el);
          }
          if (++st == SA) { st = 0; ph_a ^= 1u; }
          if (!L.resident) {
            const int cpad = L.seg_cpad[s];
            const CUtensorMap* wm = maps + L.w_map[w >> 5];
            const uint32_t b_tx = (uint32_t)ntile * 2u * w;
            for (int tap = 0; tap < NT; ++tap) {
              if (!((L.tap_mask >> tap) & 1)) continue;
              mbar_wait(empty_b(sb), ph_b);
              const uint32_t sbp = b_base + sb * b_tile;
              mbar_expect_tx_p(full_b(sb), 2 * b_tx, el);
              const int koff = L.seg_koff[s] + tap * cpad + c0;
              if (L.cluster) {
                // CTA pair: this CTA fetches half of the tile's rows (of the hi and of the lo block) and TMA multicast
                // writes them into both CTAs' rings -- every weight byte crosses L2 -> SM once per pair.  The stage was
                // released by both MMA warps (empty_b counts 2); each CTA's full_b expects the whole tile.
                const int hrows = ntile >> 1;
                const uint32_t hoff = crank * (uint32_t)hrows * 2u * (uint32_t)w;
                const CUtensorMap* wh = maps + L.w_map_half[w >> 5];
                tma_load_2d_mc_p(sbp + hoff, wh, full_b(sb), koff, n0 + (int)crank * hrows, (uint16_t)3, el);
                tma_load_2d_mc_p(sbp + b_tx + hoff, wh + 1, full_b(sb), koff, n0 + (int)crank * hrows, (uint16_t)3, el);
              } else {
                tma_load_2d_p(sbp, wm, full_b(sb), koff, n0, el);
                tma_load_2d_p(sbp + b_tx, wm + 1, full_b(sb), koff, n0, el);
              }
              if (++sb == SB) { sb = 0; ph_b ^= 1u; }
            }
          }
        }
      }
      if (L.cluster) {
        // tail: every stage's last release has arrived (from both CTAs) before this CTA may exit
        for (int k = 0; k < SB; ++k) {
          mbar_wait(empty_b(sb), ph_b);
          if (++sb == SB) { sb = 0; ph_b ^= 1u; }
        }
      }
      if (dbg && el) L.dbg_ts[11] = wait_acc;
    }
  } else if (warp == 1 || warp == 6) {
    // ===== MMA issuers (warp-uniform loops, one elected lane issues; alternate chunks, see mma_warp_loop) =====
    const int role = warp == 1 ? 0 : 1;
    const uint32_t el = elect_one();
    {
      // kind::f16, bf16 x bf16 -> fp32, M = 128.  Two MMAs per 16-channel K atom:
      //   D[:, 0:2n] += A_hi * [W_hi ; W_lo]^T   (N = 2n: the hi and lo weight rows are adjacent in smem)
      //   D[:, 0:n]  += A_lo * W_hi^T            (N = n)
      // the epilogue adds the two column halves.  A (4 KB per MMA) is the shared-memory-bandwidth
      // bound of small-N layers, so reading A_hi once instead of twice is a 1.5x saving.
      const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
      // (folded, n = 24: 3n = 72 is no UMMA N; the A_lo MMA runs with N = 80 and its last 8 columns pick up A_lo * W_lo of
      // the dx = 0 block's first 8 channels -- the lo x lo product the scheme otherwise drops, 2^-16 relative)
      const uint32_t idesc1 = idesc_base | ((uint32_t)((((L.fold ? 3 : 1) * ntile + 15) & ~15) >> 3) << 17);
      const uint32_t idesc2 = idesc_base | ((uint32_t)(((L.fold ? 6 : 2) * ntile) >> 3) << 17);
      if (L.resident) mbar_wait(wbar, 0);
      if (dbg && el && role == 0) L.dbg_ts[2] = clock64();
      MmaCtx mc;
      mc.el = el; mc.tmem_d = tmem_d; mc.smem_base = smem_base; mc.a_base = a_base; mc.a_tile = a_tile;
      mc.b_base = b_base; mc.b_tile = b_tile; mc.bar0 = bar0; mc.total_tiles = total_tiles; mc.rounds = rounds; mc.dbg = dbg;
      mc.idesc1 = idesc1; mc.idesc2 = idesc2;
      const int ks = NT == 1 ? 1 : (L.tap_mask == 0x1FF ? 3 : 2);
      if (L.fold) {
        mma_warp_loop<3, true, true>(L, mc, role);
      } else if (L.resident) {
        if (ks == 3) mma_warp_loop<3, true>(L, mc, role);
        else if (ks == 1) mma_warp_loop<1, true>(L, mc, role);
        else mma_warp_loop<2, true>(L, mc, role);
      } else {
        if (ks == 3) mma_warp_loop<3, false>(L, mc, role);
        else if (ks == 1) mma_warp_loop<1, false>(L, mc, role);
        else mma_warp_loop<2, false>(L, mc, role);
      }
    }
    __syncwarp();
  } else if (warp <= 5 || (THREADS > 224 && (MODE == 3 || L.epi8 > 1))) {
    // ===== epilogue =====
    const int q = warp & 3;                    // TMEM lane quarter (warps 7..10 -> 3, 0, 1, 2)
    const int team = warp >= 7 ? 1 + ((warp - 7) >> 2) : 0;   // extra teams take the other 16-channel groups (round robin)
    const int m = q * 32 + lane;
    int tc_ = 0;
    long long wait_epi = 0, w0 = 0;
    float bias_r[2][16];                 // bias of the CTA's channels when the whole N tile fits two register groups
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int i = 0; i < 16; ++i) bias_r[g][i] = (ntile <= 32 && g * 16 < ntile && n0 + g * 16 + i < L.cout_store) ? __ldg(L.bias + n0 + g * 16 + i) : 0.f;
    // wide path (ntile > 32): the bias of this warp's first two 16-channel groups (all it has when ntile <= 64 with two
    // teams, or <= 32 with one)
    float bias_w[2][16];
    {
      const int gs = (THREADS > 224 && L.epi8) ? L.epi8 : 1;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int ch = n0 + (team + k * gs) * 16 + i;
          // (the four-team instantiation has 96 registers: it keeps only its first group's bias, a second group -- N tiles
          // > 64 -- reads it through L1; holding both spilled 88 bytes per thread and cost base.5 193 -> 265 us)
          bias_w[k][i] = (MODE == 1 && (k == 0 || THREADS <= 352) && (team + k * gs) * 16 < ntile && ch < L.cout_store) ? __ldg(L.bias + ch) : 0.f;
        }
    }
    for (int t0 = blockIdx.x; tc_ < rounds; t0 += gridDim.x, ++tc_) {
      const int acc = tc_ & 1;
      if (MODE == 3 && acc != team) continue;      // alternate-tile teams: the other team's tile (and accumulator buffer)
      const bool ghost = t0 >= total_tiles;        // CTA pair, last round of the odd CTA: drained, nothing stored
      const int t = ghost ? total_tiles - 1 : t0;
      const int img = fast_div(t, tiles_per_img, inv_tpi);
      const int r = t - img * tiles_per_img;
      const int ty = fast_div(r, L.tiles_x, inv_tx), tx = r - ty * L.tiles_x;
      // fold: accumulator row m = input column x' = m & 15 of row m >> 4; output column = x' - 1 + tile origin
      const int oy = L.fold ? ty * 8 + (m >> 4) : ty * 16 + (m >> 3);
      const int ox = L.fold ? tx * 14 + (m & 15) - 1 : tx * 8 + (m & 7);
      const bool inside = !ghost && (oy < L.Hout) && (ox < L.Wout) && (!L.fold || ((m & 15) >= 1 && (m & 15) <= 14));
      if (dbg) w0 = clock64();
      mbar_wait(tmem_full(acc), (tc_ >> 1) & 1);
      if (dbg) wait_epi += clock64() - w0;
      if (dbg && tc_ == 0 && threadIdx.x == 64) L.dbg_ts[5] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // pool: the even/even pixel of each 2x2 block stores the average, into the half-resolution tensor
      const bool pool = !LEAN && L.pool;
      const int s2d_block = LEAN ? 0 : L.s2d_block;
      const bool store_px = pool ? (!ghost && ((m & 9) == 0) && (oy >> 1) < (L.Hout >> 1) && (ox >> 1) < (L.Wout >> 1)) : inside;
      const size_t pix = s2d_block
                             ? (size_t)img * L.out_img_stride + ((size_t)(oy >> 1) * (L.Wout >> 1) + (ox >> 1)) * L.out_cs +
                                   (size_t)((oy & 1) * 2 + (ox & 1)) * s2d_block
                             : pool ? (size_t)img * L.out_img_stride + ((size_t)(oy >> 1) * (L.Wout >> 1) + (ox >> 1)) * L.out_cs
                                      : (size_t)img * L.out_img_stride + ((size_t)oy * L.Wout + ox) * L.out_cs;
      // bilinear source of the optional additive term (conv1x1_up fused with TransitionUp)
      const float *a00 = nullptr, *a01 = nullptr, *a10 = nullptr, *a11 = nullptr;
      uint32_t s00 = 0, s01 = 0, s10 = 0, s11 = 0;       // shared-memory addresses of the four source pixels (patch form)
      float aly = 0.f, alx = 0.f;
      // (the four-team instantiation never carries the additive term: its 96 registers have no room for the eight source
      // addresses, and the fused conv1x1_up layers measured slower with four teams anyway)
      if (ADD && L.add_src && inside) {
        const float fy = L.add_sh * (float)oy, fx = L.add_sw * (float)ox;
        const int ay0 = (int)fy, ax0 = (int)fx;
        const int ay1 = ay0 + (ay0 < L.add_H - 1 ? 1 : 0), ax1 = ax0 + (ax0 < L.add_W - 1 ? 1 : 0);
        aly = fy - (float)ay0; alx = fx - (float)ax0;
        if (L.add_pbytes) {
          const int py0 = (int)(L.add_sh * (float)(ty * 16)), px0 = (int)(L.add_sw * (float)(tx * 8));
          const uint32_t pb = p_base + (uint32_t)acc * L.add_pbytes;
          const uint32_t ps = (uint32_t)(ntile + 4) * 4u;                  // bytes per patch pixel
          s00 = pb + (uint32_t)((ay0 - py0) * kAddPW + (ax0 - px0)) * ps; s01 = pb + (uint32_t)((ay0 - py0) * kAddPW + (ax1 - px0)) * ps;
          s10 = pb + (uint32_t)((ay1 - py0) * kAddPW + (ax0 - px0)) * ps; s11 = pb + (uint32_t)((ay1 - py0) * kAddPW + (ax1 - px0)) * ps;
        } else {
          const float* ab = L.add_src + (size_t)img * L.add_img;
          a00 = ab + ((size_t)ay0 * L.add_W + ax0) * L.add_cs; a01 = ab + ((size_t)ay0 * L.add_W + ax1) * L.add_cs;
          a10 = ab + ((size_t)ay1 * L.add_W + ax0) * L.add_cs; a11 = ab + ((size_t)ay1 * L.add_W + ax1) * L.add_cs;
        }
      }
      if (ADD && L.add_pbytes) mbar_wait(full_p(acc), (tc_ >> 1) & 1);
      const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * (L.fold ? 6 : 2) * ntile);
      // the accumulator buffer goes back to the MMA warps as soon as it has been READ (not after the stores):
      // with only two buffers the MMAs of tile i+2 otherwise wait for the whole epilogue of tile i
      auto release = [&]() {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty(acc));
      };
      // nv = 16, or 8 when only the first eight channels are this CTA's to write (N tile of 24 next to another block)
      auto finish16 = [&](float (&v)[16], const int n, const int nv = 16) {     // v = conv + bias of channels [n, n + 16)
        if (ADD && s00) {
          // four products per channel (the weights are per pixel): the interpolation costs 4 FMAs instead of 3 lerps
          const float w00 = (1.f - aly) * (1.f - alx), w01 = (1.f - aly) * alx, w10 = aly * (1.f - alx), w11 = aly * alx;
          const uint32_t co = (uint32_t)(n - n0) * 4u;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 p = lds_f4(s00 + co + 16u * i), q4 = lds_f4(s01 + co + 16u * i);
            const float4 r4 = lds_f4(s10 + co + 16u * i), s4 = lds_f4(s11 + co + 16u * i);
            v[4 * i + 0] += fmaf(w00, p.x, fmaf(w01, q4.x, fmaf(w10, r4.x, w11 * s4.x)));
            v[4 * i + 1] += fmaf(w00, p.y, fmaf(w01, q4.y, fmaf(w10, r4.y, w11 * s4.y)));
            v[4 * i + 2] += fmaf(w00, p.z, fmaf(w01, q4.z, fmaf(w10, r4.z, w11 * s4.z)));
            v[4 * i + 3] += fmaf(w00, p.w, fmaf(w01, q4.w, fmaf(w10, r4.w, w11 * s4.w)));
          }
        }
        if (ADD && a00) {
          const float ahy = 1.f - aly, ahx = 1.f - alx;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(a00 + n) + i), q4 = __ldg(reinterpret_cast<const float4*>(a01 + n) + i);
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(a10 + n) + i), s4 = __ldg(reinterpret_cast<const float4*>(a11 + n) + i);
            v[4 * i + 0] += ahy * (ahx * p.x + alx * q4.x) + aly * (ahx * r4.x + alx * s4.x);
            v[4 * i + 1] += ahy * (ahx * p.y + alx * q4.y) + aly * (ahx * r4.y + alx * s4.y);
            v[4 * i + 2] += ahy * (ahx * p.z + alx * q4.z) + aly * (ahx * r4.z + alx * s4.z);
            v[4 * i + 3] += ahy * (ahx * p.w + alx * q4.w) + aly * (ahx * r4.w + alx * s4.w);
          }
        }
        if (LEAN || L.relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (pool) {
          // AvgPool2d(2,2) of the ReLU'd outputs: lane bits 0 / 3 are the column / row parity inside the 16 x 8 tile
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t = v[i] + __shfl_xor_sync(0xffffffffu, v[i], 1);
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            v[i] = t * 0.25f;
          }
        }
        if (!store_px || (kHaloDbg && (L.dbg_mode & 2))) return;
        if (!LEAN && L.out_f32) {
          if (L.amax_ncls > 0 && n == 0) {
            // first maximum of the class logits (torch.argmax tie rule), parked in the padding channel for the
            // fused upsample + argmax kernel: a full-resolution pixel whose source pixels agree needs no interpolation
            float best = v[0];
            int arg = 0;
#pragma unroll
            for (int i = 1; i < 15; ++i)
              if (i < L.amax_ncls && v[i] > best) { best = v[i]; arg = i; }
            v[15] = __int_as_float(arg);
          }
          float4* o = reinterpret_cast<float4*>(L.out_f32 + pix + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          return;
        }
        uint4 h[2], l[2];
        uint2 th, tl;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          split_store4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), &th, &tl);
          reinterpret_cast<uint2*>(h)[i] = th;
          reinterpret_cast<uint2*>(l)[i] = tl;
        }
        // one 256-bit store per plane: the pixel's 16 channels are one 32-byte sector (two 128-bit stores wrote each
        // sector in halves)
        if (nv == 16 && (n & 15) == 0) {
          st_global_256(L.out_hi + pix + n, h[0], h[1]);
          st_global_256(L.out_lo + pix + n, l[0], l[1]);
        } else {
          // second block of a 2 x 24 split (channel offset 24: 16-byte aligned only) or a half group
          *reinterpret_cast<uint4*>(L.out_hi + pix + n) = h[0];
          *reinterpret_cast<uint4*>(L.out_lo + pix + n) = l[0];
          if (nv == 16) {
            *reinterpret_cast<uint4*>(L.out_hi + pix + n + 8) = h[1];
            *reinterpret_cast<uint4*>(L.out_lo + pix + n + 8) = l[1];
          }
        }
      };
      if (MODE != 1 && L.fold) {
        // six 16-column pieces per 16 output channels: (hi, lo) parts of the dx = 0, 1, 2 blocks.  Output column x takes
        // block 0 from input column x - 1 (one lane down), block 1 from x, block 2 from x + 1 (one lane up); the 16
        // lanes of an accumulator row are one row of the tile, and lanes 0 / 15 of a row store nothing.
        const int ng = (ntile > 16 && n0 + 16 < L.cout_store) ? 2 : 1;
        if (MODE == 2 && team >= ng) release();           // a team without a group of its own
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g >= ng) break;
          if (MODE == 2 && g != team) continue;           // two teams: one 16-channel group each
          uint32_t h0[16], h1[16], h2[16], l0[16], l1[16], l2[16];
          const uint32_t c = trow + (uint32_t)(g * 16);
          tmem_ld16_nowait(c, h0);
          tmem_ld16_nowait(c + (uint32_t)ntile, h1);
          tmem_ld16_nowait(c + (uint32_t)(2 * ntile), h2);
          tmem_ld16_nowait(c + (uint32_t)(3 * ntile), l0);
          tmem_ld16_nowait(c + (uint32_t)(4 * ntile), l1);
          tmem_ld16_nowait(c + (uint32_t)(5 * ntile), l2);
          tmem_ld_wait16(h0); tmem_ld_wait16(h1); tmem_ld_wait16(h2);
          tmem_ld_wait16(l0); tmem_ld_wait16(l1); tmem_ld_wait16(l2);
          if (MODE == 2 || g == ng - 1) release();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float e0 = __uint_as_float(h0[i]) + __uint_as_float(l0[i]);
            const float e1 = __uint_as_float(h1[i]) + __uint_as_float(l1[i]);
            const float e2 = __uint_as_float(h2[i]) + __uint_as_float(l2[i]);
            const float left = __shfl_up_sync(0xffffffffu, e0, 1), right = __shfl_down_sync(0xffffffffu, e2, 1);
            v[i] = ((left + e1) + right) + bias_r[g][i];
            // n = 24: columns 24..31 of a block are the next block's -- the slot's padding channels are written as zeros
            if (g * 16 + i >= ntile) v[i] = 0.f;
          }
          // channels this CTA may write: its own N tile, and up to the slot's padded end when it is the LAST N block (a
          // single block of 24 couts writes the slot's channels 24..31 as zeros: nobody else does, and stale NaN
          // patterns there would turn into zeros behind the consumers' ReLU -- silently wrong)
          const int lim = (blockIdx.y + 1 < gridDim.y) ? n0 + ntile : L.cout_store;
          finish16(v, n0 + g * 16, lim - (n0 + g * 16) >= 16 ? 16 : 8);
        }
      } else if ((MODE == 0 || MODE == 3) && ntile <= 32) {
        // whole accumulator row in registers (2 or 4 loads in flight), buffer released, then the math and the stores
        uint32_t r0[16], r1[16], r2[16], r3[16];
        const bool two = ntile == 32 && n0 + 16 < L.cout_store;
        tmem_ld16_nowait(trow, r0);
        tmem_ld16_nowait(trow + (uint32_t)ntile, r1);
        if (two) {
          tmem_ld16_nowait(trow + 16u, r2);
          tmem_ld16_nowait(trow + (uint32_t)ntile + 16u, r3);
        }
        tmem_ld_wait16(r0); tmem_ld_wait16(r1);
        if (two) { tmem_ld_wait16(r2); tmem_ld_wait16(r3); }
        release();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + bias_r[0][i];
        finish16(v, n0);
        if (two) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (__uint_as_float(r2[i]) + __uint_as_float(r3[i])) + bias_r[1][i];
          finish16(v, n0 + 16);
        }
      } else if (MODE <= 1) {
        int ngroups = 0;
        for (int c = 0; c < ntile && n0 + c < L.cout_store; c += 16) ++ngroups;
        const int gstep = (THREADS > 224 && L.epi8) ? L.epi8 : 1;
        int last = -1;                                  // this warp's last group
        for (int g = team; g < ngroups; g += gstep) last = g;
        const bool bias_in_regs = MODE == 1 && last >= 0 && last <= team + gstep;   // at most two groups: their bias sits in bias_w
        // two 16-channel groups per round: four TMEM loads in flight, one wait, and the accumulator buffer goes back
        // to the MMA warps right after this warp's last read
        constexpr bool kTwo = THREADS <= 352 && !ADD;            // two groups per round (the four-team and the additive-term
                                                                 // instantiations have no registers for it)
        for (int g = team; g < ngroups; g += (kTwo ? 2 : 1) * gstep) {
          const int g2 = g + gstep;
          const bool has2 = kTwo && g2 < ngroups;
          uint32_t r0[16], r1[16], r2[16], r3[16];
          tmem_ld16_nowait(trow + (uint32_t)(g * 16), r0);
          tmem_ld16_nowait(trow + (uint32_t)(ntile + g * 16), r1);
          if (has2) {
            tmem_ld16_nowait(trow + (uint32_t)(g2 * 16), r2);
            tmem_ld16_nowait(trow + (uint32_t)(ntile + g2 * 16), r3);
          }
          tmem_ld_wait16(r0); tmem_ld_wait16(r1);
          if (has2) { tmem_ld_wait16(r2); tmem_ld_wait16(r3); }
          if (g == last || (has2 && g2 == last)) release();
          float v[16];
          // bias_w[0] / [1] = this warp's first / second group (two-group rounds).  One group per round (four teams): only
          // the first group's bias is in registers
          const bool first_in_regs = kTwo ? bias_in_regs : (MODE == 1 && g == team);
          const bool second_in_regs = !kTwo && THREADS <= 352 && bias_in_regs && g == team + gstep;   // one group per round, second round
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) +
                   (first_in_regs ? bias_w[0][i] : second_in_regs ? bias_w[1][i] : __ldg(L.bias + n0 + g * 16 + i));
          finish16(v, n0 + g * 16);
          if (has2) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v[i] = (__uint_as_float(r2[i]) + __uint_as_float(r3[i])) + (bias_in_regs ? bias_w[1][i] : __ldg(L.bias + n0 + g2 * 16 + i));
            finish16(v, n0 + g2 * 16);
          }
        }
        if (last < 0) release();
      }
      if (ADD && L.add_pbytes) {                // the patch buffer goes back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_p(acc));
      }
      if (dbg && tc_ == 0 && threadIdx.x == 64) L.dbg_ts[6] = clock64();
    }
    if (dbg && threadIdx.x == 64) L.dbg_ts[12] = wait_epi;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (L.cluster) cluster_sync_all();       // neither CTA of a pair leaves while the other may still signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)L.tmem_cols));
  }
  if (dbg && threadIdx.x == 0) L.dbg_ts[8] = clock64();
}

// ------------------------------------------------------------------------------------------
// host side
static CUtensorMapSwizzle swizzle_of(int w) {
  return w == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (w == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode2() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int halo_encode_act_map(CUtensorMap* out, const void* base, int c, int cstride, int W, int H, int N,
                        size_t img_stride_elems, int w, int hx, int hy) {
  EncodeTiledFn enc = get_encode2();
  PF_REQUIRE(enc, PF_ESTATE, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)W * cstride * 2, (cuuint64_t)img_stride_elems * 2};
  cuuint32_t box[4] = {(cuuint32_t)w, (cuuint32_t)hx, (cuuint32_t)hy, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(w), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(halo activation c=%d w=%d) failed: %d", c, w, (int)r);
  return 0;
}

// fp32 NHWC tensor the epilogue interpolates from (fused TransitionUp): box = (ntile + 4) channels x kAddPW x kAddPH
int halo_encode_add_map(CUtensorMap* out, const void* base, int cstride, int W, int H, int N, size_t img_stride_elems,
                        int ntile) {
  EncodeTiledFn enc = get_encode2();
  PF_REQUIRE(enc, PF_ESTATE, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)cstride, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cstride * 4, (cuuint64_t)W * cstride * 4, (cuuint64_t)img_stride_elems * 4};
  cuuint32_t box[4] = {(cuuint32_t)(ntile + 4), (cuuint32_t)kAddPW, (cuuint32_t)kAddPH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(additive-term patch, ntile=%d) failed: %d", ntile, (int)r);
  return 0;
}

int halo_encode_weight_map(CUtensorMap* out, const void* base, int ktot, int nrows, int ntile, int w) {
  EncodeTiledFn enc = get_encode2();
  PF_REQUIRE(enc, PF_ESTATE, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)nrows};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)w, (cuuint32_t)ntile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(w), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(halo weights k=%d w=%d) failed: %d", ktot, w, (int)r);
  return 0;
}

int halo_chunk_width(int cpad) { return cpad >= 64 ? 64 : (cpad >= 32 ? 32 : 16); }

// Fills the shared-memory plan of a layer; returns false when it cannot run on this kernel.
bool halo_plan_smem(HaloLayer* L, size_t* smem_bytes) {
  const size_t patches = 2 * (size_t)L->add_pbytes;             // fused TransitionUp: two staged low-resolution patches
  const size_t budget = 218 * 1024 - patches;
  int wmax = 16;
  size_t w_total = 0, w_tx = 0;
  if (L->taps == 9 && L->tap_mask != 0x1FF && L->tap_mask != 0x1B) return false;   // the issue loops know 3x3 and the 2x2 s2d form
  L->nchunk = 0;
  for (int s = 0; s < L->nseg; ++s) {
    const int w = L->seg_w[s];
    for (int c0 = 0; c0 < L->seg_cpad[s]; c0 += w) {
      if (L->nchunk == kHaloMaxChunks) return false;
      const int nk = (L->seg_cpad[s] - c0 < w ? L->seg_cpad[s] - c0 : w) / 16;
      L->chunk[L->nchunk++] = chunk_code(w, nk, s, c0);
    }
  }
  for (int s = 0; s < L->nseg; ++s) {
    const int w = L->seg_w[s];
    if (w > wmax) wmax = w;
    const int nchunks = (L->seg_cpad[s] + w - 1) / w;
    size_t bt = align_up((size_t)2 * L->ntile * 2 * w, 1024);   // [hi rows ; lo rows] of one tap
    int nact = __builtin_popcount((unsigned)L->tap_mask & ((1u << L->taps) - 1u));
    if (L->fold) { bt = align_up((size_t)6 * L->ntile * 2 * w, 1024); nact = 3; }   // one tile per filter row
    w_total += (size_t)nchunks * nact * bt;
    w_tx += (size_t)nchunks * (L->fold ? 9 : nact) * 2 * (size_t)L->ntile * 2 * w;
  }
  const size_t a_tile = align_up((size_t)L->hx * L->hy * 2 * wmax, 1024);
  const size_t b_tile = align_up((size_t)2 * L->ntile * 2 * wmax, 1024);
  L->a_tile_bytes = (uint32_t)a_tile;
  L->b_tile_bytes = (uint32_t)b_tile;
  if (w_total + 2 * (2 * a_tile) <= budget) {
    L->resident = 1;
    L->w_bytes_total = (uint32_t)w_total;
    L->w_tx_total = (uint32_t)w_tx;
    int sa = (int)((budget - w_total) / (2 * a_tile));
    // ring depth: 4 stages, 8 for the 16-channel chunks (12 KB per stage: with 4 the bytes in flight per SM are too
    // few to cover the DRAM latency -- base.1 went 312 -> 255 us; deeper rings made the wider layers slower)
    const int cap = 2 * a_tile <= 12 * 1024 ? kMaxA : 4;
    L->stages_a = sa > cap ? cap : sa;
    L->stages_b = 1;
    *smem_bytes = w_total + (size_t)L->stages_a * 2 * a_tile + patches + 1024;
    return true;
  }
  if (L->fold) return false;          // the folded form needs resident weights: the caller retries unfolded
  L->resident = 0;
  L->w_bytes_total = 0;
  L->w_tx_total = 0;
  // streamed weights: the deepest activation ring (<= 4) that still leaves >= 6 weight stages; 2 when none does
  int sa = 4, sb = 0;
  for (; sa >= 2; --sa) {
    if ((size_t)sa * 2 * a_tile + 2 * b_tile > budget) continue;
    sb = (int)((budget - (size_t)sa * 2 * a_tile) / b_tile);
    if (sb >= 6 || sa == 2) break;
  }
  if (sa < 2 || sb < 2) return false;
  L->stages_a = sa;
  L->stages_b = sb > kMaxB ? kMaxB : sb;
  *smem_bytes = (size_t)L->stages_a * 2 * a_tile + (size_t)L->stages_b * b_tile + patches + 1024;
  return true;
}

int launch_conv_halo(const HaloLayer& L, const CUtensorMap* maps_dev, int nblocks, size_t smem_bytes, cudaStream_t st) {
  {
    // the attribute is per DEVICE: remember which devices have it (bit per device ordinal; thread-safe)
    static std::atomic<unsigned long long> attr_done{0};
    int dev = 0;
    PF_CHECK_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
      const void* fns[] = {(const void*)conv_halo_kernel<kHaloThreads, 0>, (const void*)conv_halo_kernel<kHaloThreads, 0, false, true>,
                           (const void*)conv_halo_kernel<kHaloThreads8, 2>,
                           (const void*)conv_halo_kernel<kHaloThreads8, 1>, (const void*)conv_halo_kernel<kHaloThreads8, 1, false, true>,
                           (const void*)conv_halo_kernel<kHaloThreads8, 1, true, true>,
                           (const void*)conv_halo_kernel<kHaloThreads16, 1>, (const void*)conv_halo_kernel<kHaloThreads16, 1, false, true>,
                           (const void*)conv_halo_kernel<kHaloThreads8, 3>, (const void*)conv_halo_kernel<kHaloThreads8, 3, false, true>};
      for (const void* f : fns) PF_CHECK_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
      attr_done.fetch_or(bit, std::memory_order_release);
    }
  }
  const int total_tiles = L.tiles_x * L.tiles_y * L.batch;
  int gx = kNumSMs / nblocks;
  if (gx < 1) gx = 1;
  if (gx > total_tiles) gx = total_tiles;
  if (L.cluster) {
    gx &= ~1;                                   // CTA pairs along x
    PF_REQUIRE(gx >= 2 && !L.resident && !L.fold, PF_ESTATE, "halo kernel: bad CTA-pair plan");
  }
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("PF_NO_PDL"); use_pdl = (e && e[0] == '1') ? 0 : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, nblocks);
  const bool add = L.add_src != nullptr;        // additive term: the two-team wide instantiation only
  PF_REQUIRE(!add || (L.epi8 == 2 && !L.fold && !L.alt), PF_ESTATE, "halo kernel: the additive term needs the two-team wide form");
  cfg.blockDim = dim3(L.alt ? kHaloThreads8 : L.epi8 >= 4 ? kHaloThreads16 : (L.epi8 ? kHaloThreads8 : kHaloThreads));
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr_pdl[2];
  int na = 0;
  if (use_pdl) {
    attr_pdl[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr_pdl[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (L.cluster) {
    attr_pdl[na].id = cudaLaunchAttributeClusterDimension;
    attr_pdl[na].val.clusterDim.x = 2; attr_pdl[na].val.clusterDim.y = 1; attr_pdl[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr_pdl;
  cfg.numAttrs = na;
  // the common epilogue (see LEAN); the additive-term layers (ReLU, plain split-bf16 output) always have it
  const bool lean = L.relu && !L.pool && !L.s2d_block && !L.out_f32 && !L.amax_ncls;
  PF_REQUIRE(!add || lean, PF_ESTATE, "halo kernel: the additive-term layers are plain ReLU ConvLayers");
  const void* fn;
  if (L.alt) fn = lean ? (const void*)conv_halo_kernel<kHaloThreads8, 3, false, true> : (const void*)conv_halo_kernel<kHaloThreads8, 3>;
  else if (L.epi8 >= 4) fn = lean ? (const void*)conv_halo_kernel<kHaloThreads16, 1, false, true> : (const void*)conv_halo_kernel<kHaloThreads16, 1>;
  else if (L.epi8 && L.fold) fn = (const void*)conv_halo_kernel<kHaloThreads8, 2>;
  else if (L.epi8 && add) fn = (const void*)conv_halo_kernel<kHaloThreads8, 1, true, true>;
  else if (L.epi8) fn = lean ? (const void*)conv_halo_kernel<kHaloThreads8, 1, false, true> : (const void*)conv_halo_kernel<kHaloThreads8, 1>;
  else fn = lean ? (const void*)conv_halo_kernel<kHaloThreads, 0, false, true> : (const void*)conv_halo_kernel<kHaloThreads, 0>;
  void* args[] = {(void*)&L, (void*)&maps_dev};
  PF_CHECK_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  return 0;
}

}  // namespace pf
