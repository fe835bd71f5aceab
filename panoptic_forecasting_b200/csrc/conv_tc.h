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

int tc_encode_act_map(CUtensorMap* out, const void* base, int c, int cstride, int W, int H, int N,
                      size_t img_stride_elems);
int tc_encode_weight_map(CUtensorMap* out, const void* base, int ktot, int npad, int ntile);
void tc_pick_tiling(int coutpad, int* ntile, int* nblocks, int* stages, int* tmem_cols, size_t* smem_bytes);
int launch_conv_tc(const TcLayer& L, const CUtensorMap* maps_dev, int nblocks, int batch, size_t smem_bytes,
                   cudaStream_t st);

}  // namespace pf
