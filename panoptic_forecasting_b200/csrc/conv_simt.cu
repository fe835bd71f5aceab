// fp32 SIMT ConvLayer kernel (conv + folded BN + ReLU) over NHWC channel-slice segments.
// Reference op: hardnet.py:16-25 (ConvLayer), with the torch.cat of hardnet.py:228/239/256
// eliminated: inputs are read as a list of channel slices, the output is written into a
// channel slice of the consumer's buffer.
//
// This is the bit-faithful (1e-6) path and the in-GPU checker for the tcgen05 path.
// Tiling: CTA = 8x16 output pixels x NT output channels; thread = 8 pixels of one row x 4
// channels (32 fp32 accumulators).  Per 8-channel chunk the CTA stages the input halo tile
// ([8][rows][cols] transposed, row pitch padded to avoid bank conflicts) and the 9-tap weight
// slab in shared memory; each thread reuses one row of halo values across the 3 horizontal taps.
#include <cuda_bf16.h>

#include "bgnet.h"
#include "split_bf16.cuh"

namespace pf {

constexpr int TH = 8, TW = 16, PXT = 8, KC = 8;

template <int KS, int STRIDE>
struct HaloGeom {
  static constexpr int rows = (TH - 1) * STRIDE + KS;
  static constexpr int cols = (TW - 1) * STRIDE + KS;
  static constexpr int pitch = (cols % 2 == 0) ? cols + 1 : cols + 2;  // odd pitch
  static constexpr int na = (PXT - 1) * STRIDE + KS;                  // halo values per thread-row
};

__device__ __forceinline__ float4 load4(const SegView& sv, size_t off, bool split) {
  return load4_any(sv.base, sv.base_lo, off, split);
}

template <int KS, int STRIDE, int NT, bool SPLIT_IN, bool SPLIT_OUT>
__global__ void __launch_bounds__(16 * NT / 4) conv_simt_kernel(const ConvLaunch L) {
  using G = HaloGeom<KS, STRIDE>;
  constexpr int NTHREADS = 16 * NT / 4;
  constexpr int TAPS = KS * KS;
  constexpr int PAD = KS / 2;
  __shared__ float As[KC][G::rows * G::pitch];
  __shared__ __align__(16) float Bs[TAPS][KC][NT];

  const int tid = threadIdx.x;
  const int ng = tid % (NT / 4);
  const int pg = tid / (NT / 4);
  const int py = pg / (TW / PXT);
  const int px0 = (pg % (TW / PXT)) * PXT;

  const int tiles_x = (L.Wout + TW - 1) / TW;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
  const int n0 = blockIdx.y * NT;
  const int img = blockIdx.z;
  const int oy0 = tile_y * TH, ox0 = tile_x * TW;
  const int iy0 = oy0 * STRIDE - PAD, ix0 = ox0 * STRIDE - PAD;

  float acc[PXT][4];
#pragma unroll
  for (int i = 0; i < PXT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int kbase = 0;
  for (int s = 0; s < L.nseg; ++s) {
    const SegView sv = L.segs[s];
    const size_t img_off = (size_t)img * L.in_img_stride[s];
    for (int c0 = 0; c0 < sv.cpad; c0 += KC) {
      // ---- stage A halo tile: each task = one halo pixel x 4 channels (float4)
      for (int task = tid; task < G::rows * G::cols * (KC / 4); task += NTHREADS) {
        const int half = task % (KC / 4);
        const int hp = task / (KC / 4);
        const int hy = hp / G::cols, hx = hp % G::cols;
        const int iy = iy0 + hy, ix = ix0 + hx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < L.Hin && ix >= 0 && ix < L.Win)
          v = load4(sv, img_off + ((size_t)iy * L.Win + ix) * sv.cstride + c0 + half * 4, SPLIT_IN);
        const int o = hy * G::pitch + hx;
        As[half * 4 + 0][o] = v.x;
        As[half * 4 + 1][o] = v.y;
        As[half * 4 + 2][o] = v.z;
        As[half * 4 + 3][o] = v.w;
      }
      // ---- stage B: TAPS x KC x NT weights
      for (int task = tid; task < TAPS * KC * (NT / 4); task += NTHREADS) {
        const int n4 = task % (NT / 4);
        const int kc = (task / (NT / 4)) % KC;
        const int tap = task / (NT / 4 * KC);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + n4 * 4 < L.coutpad)
          v = __ldg(reinterpret_cast<const float4*>(L.w + ((size_t)tap * L.kpad + kbase + c0 + kc) * L.coutpad + n0 + n4 * 4));
        *reinterpret_cast<float4*>(&Bs[tap][kc][n4 * 4]) = v;
      }
      __syncthreads();
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          float a[G::na];
          const float* arow = &As[kc][(py * STRIDE + dy) * G::pitch + px0 * STRIDE];
#pragma unroll
          for (int i = 0; i < G::na; ++i) a[i] = arow[i];
#pragma unroll
          for (int dx = 0; dx < KS; ++dx) {
            const float4 w4 = *reinterpret_cast<const float4*>(&Bs[dy * KS + dx][kc][ng * 4]);
#pragma unroll
            for (int i = 0; i < PXT; ++i) {
              const float av = a[i * STRIDE + dx];
              acc[i][0] = fmaf(av, w4.x, acc[i][0]);
              acc[i][1] = fmaf(av, w4.y, acc[i][1]);
              acc[i][2] = fmaf(av, w4.z, acc[i][2]);
              acc[i][3] = fmaf(av, w4.w, acc[i][3]);
            }
          }
        }
      }
      __syncthreads();
    }
    kbase += sv.cpad;
  }

  // ---- epilogue: + bias, ReLU, write channel slice
  const int n = n0 + ng * 4;
  if (n >= L.cout_store) return;
  const float4 bv = *reinterpret_cast<const float4*>(L.bias + n);
  const int oy = oy0 + py;
  if (oy >= L.Hout) return;
  const size_t orow = L.s2d_block
                          ? (size_t)img * L.out_img_stride + (size_t)(oy >> 1) * (L.Wout >> 1) * L.out_cstride +
                                (size_t)((oy & 1) * 2) * L.s2d_block + n
                          : (size_t)img * L.out_img_stride + (size_t)oy * L.Wout * L.out_cstride + n;
#pragma unroll
  for (int i = 0; i < PXT; ++i) {
    const int ox = ox0 + px0 + i;
    if (ox >= L.Wout) break;
    float4 r = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
    if (L.relu) {
      r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
    }
    const size_t o = L.s2d_block ? orow + (size_t)(ox >> 1) * L.out_cstride + (size_t)(ox & 1) * L.s2d_block
                                 : orow + (size_t)ox * L.out_cstride;
    if (!SPLIT_OUT) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(L.out) + o) = r;
    } else {
      uint2 h, l;
      split_store4(r, &h, &l);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(L.out) + o) = h;
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(L.out_lo) + o) = l;
    }
  }
}

template <int KS, int STRIDE, bool SI, bool SO>
static int launch_nt(const ConvLaunch& L, cudaStream_t st) {
  const int tiles = cdiv(L.Hout, TH) * cdiv(L.Wout, TW);
  const int cs = L.cout_store;
  // NT = 16 for the narrowest layers, else whichever of 32 / 64 pads fewer columns (ties -> 64).
  if (cs <= 16) {
    conv_simt_kernel<KS, STRIDE, 16, SI, SO><<<dim3(tiles, cdiv(cs, 16), L.b), 64, 0, st>>>(L);
  } else if (cdiv(cs, 32) * 32 < cdiv(cs, 64) * 64) {
    conv_simt_kernel<KS, STRIDE, 32, SI, SO><<<dim3(tiles, cdiv(cs, 32), L.b), 128, 0, st>>>(L);
  } else {
    conv_simt_kernel<KS, STRIDE, 64, SI, SO><<<dim3(tiles, cdiv(cs, 64), L.b), 256, 0, st>>>(L);
  }
  PF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <bool SI, bool SO>
static int launch_ks(const ConvLaunch& L, int ksize, int stride, cudaStream_t st) {
  if (ksize == 3 && stride == 1) return launch_nt<3, 1, SI, SO>(L, st);
  if (ksize == 3 && stride == 2) return launch_nt<3, 2, SI, SO>(L, st);
  if (ksize == 1 && stride == 1) return launch_nt<1, 1, SI, SO>(L, st);
  set_error("launch_conv_simt: unsupported ksize=%d stride=%d", ksize, stride);
  return PF_EINVAL;
}

int launch_conv_simt(const ConvLaunch& L, int ksize, int stride, bool split_in, bool split_out, cudaStream_t st) {
  if (!split_in && !split_out) return launch_ks<false, false>(L, ksize, stride, st);
  if (split_in && split_out) return launch_ks<true, true>(L, ksize, stride, st);
  if (split_in && !split_out) return launch_ks<true, false>(L, ksize, stride, st);
  return launch_ks<false, true>(L, ksize, stride, st);
}

}  // namespace pf
