// Split-bf16 activation storage of the tensor-core path: x (fp32) is kept as two bf16 planes,
// hi = bf16_rn(x) and lo = bf16_rn(x - hi), so that hi + lo carries ~16 mantissa bits and the
// three tensor-core products (hi*Whi + lo*Whi + hi*Wlo) reproduce the fp32 convolution to ~3e-5.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace pf {

__device__ __forceinline__ void split1(float x, unsigned* hi, unsigned* lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  *hi = (unsigned)__bfloat16_as_ushort(h);
  *lo = (unsigned)__bfloat16_as_ushort(l);
}

// two values at once: one packed cvt.rn.bf16x2.f32 per plane
__device__ __forceinline__ void split2(float a, float b, unsigned* hi, unsigned* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const unsigned hu = *reinterpret_cast<const unsigned*>(&h);
  const float ra = a - __uint_as_float(hu << 16);
  const float rb = b - __uint_as_float(hu & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  *hi = hu;
  *lo = *reinterpret_cast<const unsigned*>(&l);
}

// four values: packed conversions (F2FP.BF16.F32.PACK_AB, full rate) instead of eight scalar F2F (quarter-rate
// conversion pipe) -- same roundings, bit-identical planes; every conv epilogue goes through this
__device__ __forceinline__ void split_store4(const float4& r, uint2* h, uint2* l) {
  unsigned h01, l01, h23, l23;
  split2(r.x, r.y, &h01, &l01);
  split2(r.z, r.w, &h23, &l23);
  h->x = h01; h->y = h23;
  l->x = l01; l->y = l23;
}

__device__ __forceinline__ float bf16lo(unsigned packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi(unsigned packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// 4 consecutive channels at element offset `off`: fp32 buffer, or (hi, lo) bf16 planes.
__device__ __forceinline__ float4 load4_any(const void* base, const void* base_lo, size_t off, bool split) {
  if (!split) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
  const uint2 l = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base_lo) + off));
  float4 r;
  r.x = bf16lo(h.x) + bf16lo(l.x);
  r.y = bf16hi(h.x) + bf16hi(l.x);
  r.z = bf16lo(h.y) + bf16lo(l.y);
  r.w = bf16hi(h.y) + bf16hi(l.y);
  return r;
}
__device__ __forceinline__ void store4_any(void* base, void* base_lo, size_t off, const float4& r, bool split) {
  if (!split) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = r;
  } else {
    uint2 h, l;
    split_store4(r, &h, &l);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = h;
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base_lo) + off) = l;
  }
}

// One 256-bit global store (sm_100: STG.E.256): a whole 32-byte sector per lane, e.g. the 16 bf16 channels of one pixel
// of one plane.  `p` must be 32-byte aligned.
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// 16 consecutive channels starting at a multiple of 16: fp32 (four 128-bit stores) or one 256-bit store per bf16 plane
__device__ __forceinline__ void store16_any(void* base, void* base_lo, size_t off, const float* v, bool split) {
  if (!split) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    return;
  }
  uint2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_store4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), &h[i], &l[i]);
  st_global_256(reinterpret_cast<__nv_bfloat16*>(base) + off, make_uint4(h[0].x, h[0].y, h[1].x, h[1].y),
                make_uint4(h[2].x, h[2].y, h[3].x, h[3].y));
  st_global_256(reinterpret_cast<__nv_bfloat16*>(base_lo) + off, make_uint4(l[0].x, l[0].y, l[1].x, l[1].y),
                make_uint4(l[2].x, l[2].y, l[3].x, l[3].y));
}

// 8 consecutive channels (128-bit loads/stores on the bf16 planes)
__device__ __forceinline__ void load8_any(const void* base, const void* base_lo, size_t off, bool split, float* v) {
  if (!split) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    return;
  }
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
  const uint4 l = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base_lo) + off));
  v[0] = bf16lo(h.x) + bf16lo(l.x); v[1] = bf16hi(h.x) + bf16hi(l.x);
  v[2] = bf16lo(h.y) + bf16lo(l.y); v[3] = bf16hi(h.y) + bf16hi(l.y);
  v[4] = bf16lo(h.z) + bf16lo(l.z); v[5] = bf16hi(h.z) + bf16hi(l.z);
  v[6] = bf16lo(h.w) + bf16lo(l.w); v[7] = bf16hi(h.w) + bf16hi(l.w);
}
__device__ __forceinline__ void store8_any(void* base, void* base_lo, size_t off, const float* v, bool split) {
  if (!split) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    return;
  }
  uint2 h0, l0, h1, l1;
  split_store4(make_float4(v[0], v[1], v[2], v[3]), &h0, &l0);
  split_store4(make_float4(v[4], v[5], v[6], v[7]), &h1, &l1);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base_lo) + off) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

}  // namespace pf
