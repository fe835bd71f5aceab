// Internal interface of the tcgen05 ConvLayer kernel (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "bgnet.h"

namespace pf {

constexpr int kTcMaxStages = 4;

struct TcLayer {
  int nseg;
  int seg_cpad[kMaxSegs];   // channels of each input slice, padded to 16
  int seg_map[kMaxSegs];    // index of the slice's hi-plane tensor map (lo plane = +1)
  int seg_koff[kMaxSegs];   // first K column of the slice in the packed weight matrix
  int w_map;                // index of the weight hi tensor map (lo = +1)
  int taps, ksize;
  int Hout, Wout, tiles_x, tiles_y;
  int ntile, stages, tmem_cols;
  int cout_store;           // channels written (cout padded to 16)
  int relu;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;           // if set: fp32 output (head), no split
  int out_cs;
  size_t out_img_stride;    // elements
  const float* bias;
};

constexpr int kHaloMaxChunks = 32;
constexpr int kAddPH = 10, kAddPW = 6;   // 16 rows x 8 columns at <= 1/2 scale (align_corners): 8.5 x 4.5 source pixels + 1

// 3x3 / stride-1 layers on the persistent halo kernel (conv_halo.cu)
struct HaloLayer {
  int nseg;
  int seg_cpad[kMaxSegs];   // channels of each input slice, padded to 16
  int seg_w[kMaxSegs];      // chunk width of the slice: 16 / 32 / 64 channels (32 / 64 / 128-byte swizzle)
  int seg_map[kMaxSegs];    // hi-plane halo-box tensor map of the slice (lo = +1)
  int seg_koff[kMaxSegs];   // first K column of the slice in the packed weight matrix
  int w_map[3];             // weight tensor maps (hi; lo = +1) for chunk widths 16 / 32 / 64
  int nchunk;               // staged activation chunks per tile, in K order
  uint32_t chunk[kHaloMaxChunks];   // chunk width | K atoms << 8 | slice << 12 | first channel << 16 (halo_plan_smem fills it)
  int Hout, Wout, tiles_x, tiles_y, batch;
  int taps, hx, hy;         // 9 taps / 10 x 18 box (3x3) or 1 tap / 8 x 16 box (1x1)
  int tap_mask;             // active taps (bit dy*3+dx); 0x1FF normally, 0x1B for the space-to-depth stride-2 conv
  int s2d_block;            // > 0: space-to-depth output (see ConvDesc::s2d_out)
  int ntile, tmem_cols, stages_a, stages_b;
  uint32_t a_tile_bytes, b_tile_bytes;
  int resident;             // weights stay in shared memory for the whole kernel
  int amax_ncls;            // > 0 (fp32 head only): channel 15 of every stored pixel carries argmax over the first amax_ncls channels (int bits)
  int pool;                 // 2x2 average pool fused into the epilogue: out_* address the pooled tensor (Hout/2 x Wout/2)
  int epi8;                 // epilogue teams of 4 warps (0 = one; 2 / 4: extra teams share the TMEM lane quarters and take the other 16-channel groups)
  int cluster;              // streamed weights: CTA pairs (cluster 2 x 1 x 1) share the weight ring, each CTA loads half of every tile and TMA multicast writes it into both
  int w_map_half[3];        // weight tensor maps with a box of ntile / 2 rows (hi; lo = +1) for the pair's half loads
  int alt;                  // ntile <= 32: two epilogue teams on alternate tiles (team k owns accumulator buffer k); excludes epi8
  int fold;                 // 3x3, ntile <= 32, resident: the three taps of a filter row folded into N (10 x 16 box, 8 x 14 output tiles)
  uint32_t w_bytes_total, w_tx_total;
  int cout_store, relu;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;           // if set: fp32 output (head), no split
  int out_cs;
  size_t out_img_stride;
  const float* bias;
  // optional epilogue term: + bilinear(align_corners) interpolation of an fp32 NHWC tensor at the output pixel
  const float* add_src;
  int add_H, add_W, add_cs;
  size_t add_img;
  float add_sh, add_sw;
  // ... staged per tile by TMA into shared memory (add_pbytes > 0): the low-resolution patch a 16 x 8 output tile
  // interpolates from is kAddPH x kAddPW pixels x (ntile + 4) floats (4 floats of padding per pixel keep the
  // epilogue's 128-bit reads of neighbouring patch pixels off the same banks); two buffers, like the accumulators
  int add_map;              // tensor map of add_src: fp32 (C, W, H, N), box (ntile + 4, kAddPW, kAddPH, 1)
  uint32_t add_pbytes;      // bytes of one patch buffer (1024-aligned); 0 = per-pixel global gathers
  long long* dbg_ts;        // optional: CTA 0 writes phase timestamps (clock64) here
  int dbg_mode;             // timing experiments only (results invalid): 1 = no activation TMA after the first ring pass, 2 = no stores
};

int halo_chunk_width(int cpad);
int halo_encode_act_map(CUtensorMap* out, const void* base, int c, int cstride, int W, int H, int N,
                        size_t img_stride_elems, int w, int hx, int hy);
int halo_encode_weight_map(CUtensorMap* out, const void* base, int ktot, int nrows, int ntile, int w);
int halo_encode_add_map(CUtensorMap* out, const void* base, int cstride, int W, int H, int N, size_t img_stride_elems,
                        int ntile);
bool halo_plan_smem(HaloLayer* L, size_t* smem_bytes);
int launch_conv_halo(const HaloLayer& L, const CUtensorMap* maps_dev, int nblocks, size_t smem_bytes, cudaStream_t st);

int tc_encode_act_map(CUtensorMap* out, const void* base, int c, int cstride, int W, int H, int N,
                      size_t img_stride_elems, int box_w = 16, int box_h = 8);
int tc_encode_weight_map(CUtensorMap* out, const void* base, int ktot, int npad, int ntile);
void tc_pick_tiling(int coutpad, int total_tiles, int* ntile, int* nblocks, int* stages, int* tmem_cols,
                    size_t* smem_bytes);
int launch_conv_tc(const TcLayer& L, const CUtensorMap* maps_dev, int nblocks, int batch, size_t smem_bytes,
                   cudaStream_t st);

}  // namespace pf
