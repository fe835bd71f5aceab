// Tensor-core ConvLayer kernel for sm_100a: implicit GEMM on tcgen05 (UMMA) with TMA-staged
// NHWC tiles and a TMEM accumulator.  Reference op: hardnet.py:16-25 (conv + folded BN + ReLU).
//
// GEMM view per CTA:  D[128 pixels x Ntile couts] += A[128 x K] * W[Ntile x K]^T
//   M = an 8 x 16 block of output pixels of one image (TMA box {64 ch, 16, 8, 1} of the NHWC
//       activation; the 3x3 taps are the same box shifted by (dx-1, dy-1), image borders are
//       zero-filled by TMA out-of-bounds handling == the conv's zero padding),
//   K = taps x sum over input channel slices (the reference's torch.cat is just the K order),
//   operands are K-major, 128-byte swizzled, bf16.
// Precision: activations and weights are split x = hi + lo (two bf16 planes); each 16-channel
// K atom issues three MMAs (hi*Whi, lo*Whi, hi*Wlo) into the same fp32 TMEM accumulator,
// which restores ~fp32 accuracy (3e-5 relative on the logits of the whole net) at 1/3 of the
// bf16 tensor rate -- still ~10x the fp32 SIMT rate.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias/ReLU/split -> global).
#include <atomic>

#include <cuda.h>
#include <cuda_bf16.h>

#include "bgnet.h"
#include "conv_tc.h"
#include "split_bf16.cuh"
#include "tc_common.cuh"

namespace pf {

using namespace tc;

__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const TcLayer L, const CUtensorMap* __restrict__ maps) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kTcMaxStages + 1];
  __shared__ uint32_t tmem_base_smem;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int ntile = L.ntile;
  const uint32_t b_bytes = (uint32_t)ntile * kBlockK * 2;
  const uint32_t stage_bytes = 2 * kATileBytes + 2 * b_bytes;
  const int S = L.stages;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kTcMaxStages + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * kTcMaxStages);

  const int tile = blockIdx.x;
  const int ty = tile / L.tiles_x, tx = tile - ty * L.tiles_x;
  const int y0 = ty * kTileH, x0 = tx * kTileW;
  const int n0 = blockIdx.y * ntile;
  const int img = blockIdx.z;
  const int pad = L.ksize >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_smem;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loops, one elected lane issues) =====
    const uint32_t el = elect_one();
    {
      int it = 0;
      for (int s = 0; s < L.nseg; ++s) {
        const int cpad = L.seg_cpad[s];
        const CUtensorMap* mhi = maps + L.seg_map[s];
        for (int tap = 0; tap < L.taps; ++tap) {
          const int dy = tap / L.ksize, dx = tap - dy * L.ksize;
          for (int c0 = 0; c0 < cpad; c0 += kBlockK, ++it) {
            const int st = it % S;
            mbar_wait(empty_bar(st), ((it / S) & 1) ^ 1);
            const uint32_t sa = smem_base + st * stage_bytes;
            mbar_expect_tx_p(full_bar(st), stage_bytes, el);
            tma_load_4d_p(sa, mhi, full_bar(st), c0, x0 + dx - pad, y0 + dy - pad, img, el);
            tma_load_4d_p(sa + kATileBytes, mhi + 1, full_bar(st), c0, x0 + dx - pad, y0 + dy - pad, img, el);
            const int koff = L.seg_koff[s] + tap * cpad + c0;
            tma_load_2d_p(sa + 2 * kATileBytes, maps + L.w_map, full_bar(st), koff, n0, el);
            tma_load_2d_p(sa + 2 * kATileBytes + b_bytes, maps + L.w_map + 1, full_bar(st), koff, n0, el);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loops, one elected lane issues) =====
    const uint32_t el = elect_one();
    {
      // kind::f16, A/B = bf16 K-major, D = fp32, M = 128, N = ntile
      // two MMAs per K atom (see conv_halo.cu): A_hi x [W_hi;W_lo] (N = 2n) and A_lo x W_hi (N = n)
      const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
      const uint32_t idesc1 = idesc_base | ((uint32_t)(ntile >> 3) << 17);
      const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * ntile) >> 3) << 17);
      int it = 0;
      for (int s = 0; s < L.nseg; ++s) {
        const int cpad = L.seg_cpad[s];
        for (int tap = 0; tap < L.taps; ++tap) {
          for (int c0 = 0; c0 < cpad; c0 += kBlockK, ++it) {
            const int st = it % S;
            mbar_wait(full_bar(st), (it / S) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_base + st * stage_bytes;
            const int nk = min(kBlockK, cpad - c0) >> 4;
            const uint64_t a_hi0 = umma_desc(sa), a_lo0 = umma_desc(sa + kATileBytes);
            const uint64_t b_hi0 = umma_desc(sa + 2 * kATileBytes);   // lo rows follow the hi rows
#pragma unroll 4
            for (int ka = 0; ka < nk; ++ka) {
              const uint64_t ko = (uint64_t)(ka * 2);            // 32 bytes per k-atom, in 16-byte units
              umma_bf16_p(tmem_d, a_hi0 + ko, b_hi0 + ko, idesc2, (it > 0 || ka > 0) ? 1u : 0u, el);
              umma_bf16_p(tmem_d, a_lo0 + ko, b_hi0 + ko, idesc1, 1u, el);
            }
            umma_commit_p(empty_bar(st), el);   // frees the stage once these MMAs have read it
          }
        }
      }
      umma_commit_p(accum_bar, el);             // accumulator complete
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM lane quarter (warp % 4) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;           // accumulator row = pixel index inside the 8x16 tile
    const int oy = y0 + (m >> 4), ox = x0 + (m & 15);
    const bool inside = (oy < L.Hout) && (ox < L.Wout);
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const size_t pix = (size_t)img * L.out_img_stride + ((size_t)oy * L.Wout + ox) * L.out_cs;
    for (int c = 0; c < ntile; c += 16) {
      const int n = n0 + c;
      if (n >= L.cout_store) break;        // warp-uniform
      float v[16], v2[16];
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ntile + c), v2);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = (v[i] + v2[i]) + __ldg(L.bias + n + i);
        if (L.relu) v[i] = fmaxf(v[i], 0.f);
      }
      if (!inside) continue;
      if (L.out_f32) {
        float4* o = reinterpret_cast<float4*>(L.out_f32 + pix + n);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
        uint4 h[2], l[2];
        uint2 th, tl;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          split_store4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), &th, &tl);
          reinterpret_cast<uint2*>(h)[i] = th;
          reinterpret_cast<uint2*>(l)[i] = tl;
        }
        uint4* oh = reinterpret_cast<uint4*>(L.out_hi + pix + n);
        uint4* ol = reinterpret_cast<uint4*>(L.out_lo + pix + n);
        oh[0] = h[0]; oh[1] = h[1];
        ol[0] = l[0]; ol[1] = l[1];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)L.tmem_cols));
  }
}


// ------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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

int tc_encode_act_map(CUtensorMap* out, const void* base, int c, int cstride, int W, int H, int N,
                      size_t img_stride_elems, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  PF_REQUIRE(enc, PF_ESTATE, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)W * cstride * 2, (cuuint64_t)img_stride_elems * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(activation c=%d cs=%d %dx%dx%d) failed: %d", c,
             cstride, W, H, N, (int)r);
  return 0;
}

int tc_encode_weight_map(CUtensorMap* out, const void* base, int ktot, int npad, int ntile) {
  EncodeTiledFn enc = get_encode();
  PF_REQUIRE(enc, PF_ESTATE, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)npad};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)ntile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(weights k=%d n=%d) failed: %d", ktot, npad, (int)r);
  return 0;
}

void tc_pick_tiling(int coutpad, int total_tiles, int* ntile, int* nblocks, int* stages, int* tmem_cols,
                    size_t* smem_bytes) {
  // N <= 128 per CTA.  Layers with few pixel tiles (low resolution) are split along N as well so
  // that about one CTA per SM exists: their cost is per-CTA operand streaming, not L2 traffic.
  int nb = (coutpad + 127) / 128;
  const int want = kNumSMs / total_tiles;      // never more CTAs than SMs
  if (want > nb) nb = want;
  if (nb > coutpad / 16) nb = coutpad / 16;
  int nt = ((coutpad + nb - 1) / nb + 15) / 16 * 16;
  *ntile = nt;
  *nblocks = (coutpad + nt - 1) / nt;
  const size_t stage = 2 * (size_t)kATileBytes + 2 * (size_t)nt * kBlockK * 2;
  int s = (int)((200 * 1024) / stage);
  if (s > kTcMaxStages) s = kTcMaxStages;
  if (s < 2) s = 2;
  *stages = s;
  int cols = 32;
  while (cols < 2 * nt) cols <<= 1;
  *tmem_cols = cols;
  *smem_bytes = stage * s + 1024;
}

int launch_conv_tc(const TcLayer& L, const CUtensorMap* maps_dev, int nblocks, int batch, size_t smem_bytes,
                   cudaStream_t st) {
  {
    // the attribute is per DEVICE: remember which devices have it (bit per device ordinal; thread-safe)
    static std::atomic<unsigned long long> attr_done{0};
    int dev = 0;
    PF_CHECK_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
      PF_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
      attr_done.fetch_or(bit, std::memory_order_release);
    }
  }
  dim3 grid(L.tiles_x * L.tiles_y, nblocks, batch);
  conv_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(L, maps_dev);
  PF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pf
