// Internal (non-ABI) description of the HarDNet-70 execution plan shared by the kernel files.
#pragma once
#include <string>
#include <vector>

#include "pf_common.cuh"

namespace pf {

constexpr int kMaxSegs = 10;     // conv1x1_up reads <=5 upsampled block-output slots + <=5 skip slots
constexpr int kChanAlign = 16;   // every activation slot starts at, and is padded to, 16 channels (one bf16 UMMA K atom)

struct SegRef {                  // a channel slice of an NHWC activation buffer
  int buf = -1;
  int coff = 0;                  // first channel (multiple of kChanAlign)
  int c = 0;                     // true channel count
  int cpad() const { return (c + kChanAlign - 1) / kChanAlign * kChanAlign; }
};

struct BufDesc {
  int shift = 0;                 // spatial size = (H >> shift, W >> shift)
  int cstride = 0;               // channels per pixel (sum of padded slots)
  size_t offset_floats = 0;      // per-image offset inside the arena, filled at plan time
  bool always_f32 = false;       // fp32 even with split-bf16 storage (quarter-res logits, low-res conv1x1_up partials)
};

struct ConvDesc {
  std::string name;
  int cin = 0, cout = 0, ksize = 3, stride = 1;
  bool relu = true;
  std::vector<SegRef> in;        // in reference cat order
  SegRef out;
  int kpad = 0;                  // sum of padded input channels
  int coutpad = 0;               // cout padded to 16
  // device weights (fp32 path): w[tap][kpad][coutpad], bias[coutpad]
  float* w_dev = nullptr;
  float* bias_dev = nullptr;
  bool loaded = false;
  // tensor-core storage only: base.1 writes its output space-to-depth (4 phase blocks of 32 channels at
  // half its resolution) so that base.2, the only 3x3 STRIDE-2 ConvLayer besides the first, runs as a
  // stride-1 conv with 4 active taps over that tensor on the tcgen05 halo kernel.
  bool s2d_out = false, s2d_in = false;
  // tensor-core path only: conv1x1_up fused with TransitionUp.  A 1x1 conv commutes with the (per-channel,
  // linear) bilinear upsample, so the first `up_nseg` input slices are convolved at LOW resolution into the fp32
  // buffer `ybuf` (no bias / ReLU) and the high-resolution pass over the skip slices adds its bilinear
  // interpolation in the epilogue: relu(W_skip * skip + up(W_up * x) + b).  No upsampled tensor is materialised.
  int up_nseg = 0;
  int ybuf = -1;
  int pool_step = -1;       // index of the AvgPool2d(2,2) step that consumes this conv's output (hardnet.py:293-294)
  int exec_stride() const { return s2d_in ? 1 : stride; }
  std::vector<float> w_host;     // folded, packed like w_dev (kept for re-packing by the tensor-core path)
  std::vector<float> bias_host;
};

enum StepType { STEP_FIRST = 0, STEP_CONV = 1, STEP_POOL = 2, STEP_UPSAMPLE = 3, STEP_HEAD = 4 };

struct Step {
  StepType type;
  int conv = -1;                 // STEP_FIRST / STEP_CONV / STEP_HEAD(final conv index)
  std::vector<SegRef> in;        // STEP_POOL (1 seg) / STEP_UPSAMPLE (n segs)
  SegRef out;
};

// Device-side views --------------------------------------------------------------------------
struct SegView {
  const void* base;              // pixel (0,0) of image 0, channel coff: fp32, or bf16 hi plane (split storage)
  const void* base_lo;           // bf16 lo plane (split storage) or nullptr
  int cstride;
  int cpad;
};

struct ConvLaunch {
  SegView segs[kMaxSegs];
  int nseg;
  int b, Hin, Win, Hout, Wout;
  size_t in_img_stride[kMaxSegs];  // elements per image for each seg's buffer
  void* out;                     // pixel (0,0) of image 0, channel out.coff (fp32, or bf16 hi plane)
  void* out_lo;                  // bf16 lo plane (split storage) or nullptr
  int out_cstride;
  size_t out_img_stride;
  const float* w;
  const float* bias;
  int kpad, coutpad, cout_store; // cout_store = channels written (cout padded to 16)
  int relu;
  int s2d_block;                 // > 0: write pixel (y,x) to (y/2, x/2), channel block ((y&1)*2+(x&1)) * s2d_block
};

// split_in / split_out: activations stored as bf16 (hi, lo) planes with x = hi + lo (tensor-core path storage)
int launch_conv_simt(const ConvLaunch& L, int ksize, int stride, bool split_in, bool split_out, cudaStream_t st);

}  // namespace pf
