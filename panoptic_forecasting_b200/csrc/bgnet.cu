// Stage B: BGModel / HarDNet-70 forward behind the C ABI.
// Reference: bg_model.py:53-71,91-102; hardnet.py:176-240 (HarDBlock), :243-258 (TransitionUp),
// :262-327 (topology), :353-387 (forward).
//
// Layout: every activation is fp32 NHWC inside one caller-provided arena; every HarDBlock owns a
// single buffer whose channel slots are [block input | layer1 | layer2 | ...], so the reference's
// torch.cat calls disappear: a consumer reads a list of channel slices (SegRef), a producer writes
// its slice.  Slots are aligned/padded to 8 channels; producers write zeros into the pad channels.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "bgnet.h"

namespace pf {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------
// Topology constants (hardnet.py:265-269)
static const int kFirstCh[4] = {16, 24, 32, 48};
static const int kChList[5] = {64, 96, 160, 224, 320};
static const double kGrmul = 1.7;
static const int kGr[5] = {10, 16, 18, 24, 32};
static const int kNLayers[5] = {4, 4, 8, 8, 8};

// hardnet.py:177-194
static void get_link(int layer, int base_ch, int gr, int* out_ch, std::vector<int>* link) {
  if (layer == 0) {
    *out_ch = base_ch;
    if (link) link->clear();
    return;
  }
  double oc = gr;
  std::vector<int> lk;
  for (int i = 0; i < 10; ++i) {
    int dv = 1 << i;
    if (layer % dv == 0) {
      lk.push_back(layer - dv);
      if (i > 0) oc *= kGrmul;
    }
  }
  *out_ch = (int)((int)(oc + 1) / 2) * 2;
  if (link) *link = lk;
}

static int pad8(int c) { return (c + kChanAlign - 1) / kChanAlign * kChanAlign; }

}  // namespace pf

using namespace pf;

struct pf_bgnet {
  int num_classes = 11, num_inputs = 3, use_depth = 1, precision = 0;
  float depth_mean = 0.f, depth_std = 1.f;
  bool depth_norm_set = false;
  std::vector<BufDesc> bufs;
  std::vector<ConvDesc> convs;   // ConvLayers in execution order; last entry = finalConv
  std::vector<Step> steps;
  int first_conv = 0, final_conv = -1;
  int quarter_buf = -1;          // fp32 NHWC [H/4, W/4, 16] logits
  // first conv (labels -> 16 ch) tables: lut[tap][frame][class+1][16], wd[tap][frame][16], bias[16]
  float* first_tab_dev = nullptr;
  size_t first_tab_floats = 0;
  int launches = 0;
  // optional per-step CUDA-event profiling (bench.py roofline): ring of [iters][steps+1] events
  std::vector<cudaEvent_t> prof_ev;
  int prof_cap = 0, prof_iter = 0;

  int new_buf(int shift, int cstride) {
    BufDesc b;
    b.shift = shift;
    b.cstride = cstride;
    bufs.push_back(b);
    return (int)bufs.size() - 1;
  }
  int add_conv(const std::string& name, int cin, int cout, int k, int stride, const std::vector<SegRef>& in,
               SegRef out, bool relu = true) {
    ConvDesc c;
    c.name = name; c.cin = cin; c.cout = cout; c.ksize = k; c.stride = stride; c.in = in; c.out = out; c.relu = relu;
    c.kpad = 0;
    for (auto& s : in) c.kpad += s.cpad();
    c.coutpad = (cout + 15) / 16 * 16;
    convs.push_back(c);
    return (int)convs.size() - 1;
  }
};

namespace pf {

// Builds one HarDBlock: returns the buffer and fills `out_segs` (hardnet.py:234-239: odd layers + last).
// `in_slot` receives the slot the producer of the block input must write to.
static void build_block(pf_bgnet* net, const std::string& prefix, int in_ch, int gr, int n_layers, int shift,
                        SegRef* in_slot, std::vector<SegRef>* out_segs, int* out_ch_total) {
  std::vector<int> ch(n_layers + 1), off(n_layers + 1);
  ch[0] = in_ch;
  for (int l = 1; l <= n_layers; ++l) get_link(l, in_ch, gr, &ch[l], nullptr);
  int total = 0;
  for (int l = 0; l <= n_layers; ++l) { off[l] = total; total += pad8(ch[l]); }
  const int buf = net->new_buf(shift, total);
  in_slot->buf = buf; in_slot->coff = 0; in_slot->c = in_ch;
  for (int l = 1; l <= n_layers; ++l) {
    int oc; std::vector<int> link;
    get_link(l, in_ch, gr, &oc, &link);
    std::vector<SegRef> in;
    int cin = 0;
    for (int k : link) { SegRef s; s.buf = buf; s.coff = off[k]; s.c = ch[k]; in.push_back(s); cin += ch[k]; }
    SegRef out; out.buf = buf; out.coff = off[l]; out.c = oc;
    char nm[96];
    snprintf(nm, sizeof(nm), "%s.layers.%d", prefix.c_str(), l - 1);
    int ci = net->add_conv(nm, cin, oc, 3, 1, in, out);
    Step st; st.type = STEP_CONV; st.conv = ci;
    net->steps.push_back(st);
  }
  out_segs->clear();
  *out_ch_total = 0;
  const int t = n_layers + 1;
  for (int i = 0; i < t; ++i) {
    if (i == t - 1 || i % 2 == 1) {
      SegRef s; s.buf = buf; s.coff = off[i]; s.c = ch[i];
      out_segs->push_back(s);
      *out_ch_total += ch[i];
    }
  }
}

static void build_topology(pf_bgnet* net) {
  const int t = net->num_inputs;
  const int cin0 = (net->num_classes + (net->use_depth ? 1 : 0)) * t;
  // stem (hardnet.py:275-280)
  int s0 = net->new_buf(1, pad8(kFirstCh[0]));
  int s1 = net->new_buf(1, pad8(kFirstCh[1]));
  int s2 = net->new_buf(2, pad8(kFirstCh[2]));
  SegRef r0{s0, 0, kFirstCh[0]}, r1{s1, 0, kFirstCh[1]}, r2{s2, 0, kFirstCh[2]};
  {
    int ci = net->add_conv("model.base.0", cin0, kFirstCh[0], 3, 2, {}, r0);
    net->first_conv = ci;
    Step st; st.type = STEP_FIRST; st.conv = ci; net->steps.push_back(st);
    ci = net->add_conv("model.base.1", kFirstCh[0], kFirstCh[1], 3, 1, {r0}, r1);
    st.type = STEP_CONV; st.conv = ci; net->steps.push_back(st);
    ci = net->add_conv("model.base.2", kFirstCh[1], kFirstCh[2], 3, 2, {r1}, r2);
    st.conv = ci; net->steps.push_back(st);
  }
  // base.3 writes straight into encoder block 0's input slot; patched after the block exists.
  int conv3 = net->add_conv("model.base.3", kFirstCh[2], kFirstCh[3], 3, 1, {r2}, SegRef());
  { Step st; st.type = STEP_CONV; st.conv = conv3; net->steps.push_back(st); }

  int ch = kFirstCh[3];
  int idx = 4;
  std::vector<std::vector<SegRef>> skips;
  std::vector<int> skip_ch;
  int pending_producer = conv3;        // conv whose `out` is the next block's input slot
  int pending_pool_step = -1;          // or a pool step
  std::vector<SegRef> cur_segs;        // "out" of the encoder
  int cur_ch = 0;
  for (int i = 0; i < 5; ++i) {
    const int shift = 2 + i;
    char pfx[64];
    snprintf(pfx, sizeof(pfx), "model.base.%d", idx);
    SegRef in_slot; std::vector<SegRef> outs; int oc;
    build_block(net, pfx, ch, kGr[i], kNLayers[i], shift, &in_slot, &outs, &oc);
    if (pending_producer >= 0) net->convs[pending_producer].out = in_slot;
    if (pending_pool_step >= 0) net->steps[pending_pool_step].out = in_slot;
    pending_producer = -1; pending_pool_step = -1;
    idx++;
    if (i < 4) { skips.push_back(outs); skip_ch.push_back(oc); }
    // 1x1 transition (hardnet.py:292)
    snprintf(pfx, sizeof(pfx), "model.base.%d", idx);
    idx++;
    int pbuf = net->new_buf(shift, pad8(kChList[i]));
    SegRef pout{pbuf, 0, kChList[i]};
    int ci = net->add_conv(pfx, oc, kChList[i], 1, 1, outs, pout);
    { Step st; st.type = STEP_CONV; st.conv = ci; net->steps.push_back(st); }
    ch = kChList[i];
    if (i < 4) {
      Step st; st.type = STEP_POOL; st.in = {pout};
      net->steps.push_back(st);
      pending_pool_step = (int)net->steps.size() - 1;
      idx++;
    } else {
      cur_segs = {pout};
      cur_ch = ch;
    }
  }
  // decoder (hardnet.py:312-322, 365-369)
  for (int j = 0; j < 4; ++j) {
    const int i = 3 - j;
    const int shift = 2 + i;
    // the upsampled tensor keeps the (padded) slot layout of its source slices, so conv1x1_up
    // reads it as the same list of slices followed by the skip's slices (hardnet.py:256 cat order).
    int ctot = 0;
    for (auto& s : cur_segs) ctot += s.cpad();
    int ubuf = net->new_buf(shift, ctot);
    std::vector<SegRef> cat_in;
    {
      int off = 0;
      for (auto& s : cur_segs) { SegRef u{ubuf, off, s.c}; cat_in.push_back(u); off += s.cpad(); }
    }
    SegRef uout{ubuf, 0, ctot};
    { Step st; st.type = STEP_UPSAMPLE; st.in = cur_segs; st.out = uout; net->steps.push_back(st); }
    for (auto& s : skips[i]) cat_in.push_back(s);
    const int ccat = cur_ch + skip_ch[i];
    const int chalf = ccat / 2;
    char nm[64];
    snprintf(nm, sizeof(nm), "model.conv1x1_up.%d", j);
    int c1 = net->add_conv(nm, ccat, chalf, 1, 1, cat_in, SegRef());
    { Step st; st.type = STEP_CONV; st.conv = c1; net->steps.push_back(st); }
    snprintf(nm, sizeof(nm), "model.denseBlocksUp.%d", j);
    SegRef in_slot; std::vector<SegRef> outs; int oc;
    build_block(net, nm, chalf, kGr[i], kNLayers[i], shift, &in_slot, &outs, &oc);
    net->convs[c1].out = in_slot;
    cur_segs = outs;
    cur_ch = oc;
  }
  // finalConv (hardnet.py:325-327,371): 1x1 + bias, no BN / ReLU
  net->quarter_buf = net->new_buf(2, 16);
  SegRef qout{net->quarter_buf, 0, net->num_classes};
  net->final_conv = net->add_conv("model.finalConv", cur_ch, net->num_classes, 1, 1, cur_segs, qout, false);
  { Step st; st.type = STEP_HEAD; st.conv = net->final_conv; net->steps.push_back(st); }
}

// ------------------------------------------------------------------------------------------
// K2: first ConvLayer straight from labels + depth (bg_model.py:53-69 + hardnet.py base.0).
// The 11-way one-hot never exists: each tap contributes a row of a (tap, frame, class) weight
// table; depth planes contribute (d-mean)/std*mask times their 3x3 weights.  3x3, stride 2,
// pad 1, 16 output channels.  CTA = 8x32 output pixels, one pixel per thread.
constexpr int F_TH = 8, F_TW = 32;
constexpr int F_IH = F_TH * 2 + 1, F_IW = F_TW * 2 + 1;

struct FirstParams {
  const uint8_t* labels; const float* depth; const uint8_t* mask;
  const float* tab;    // lut | wd | bias
  float* out;          // NHWC, 16 channels
  int b, t, H, W, Ho, Wo, ncls, use_depth;
  float mean, std;
};

__global__ void __launch_bounds__(F_TH* F_TW) first_conv_kernel(FirstParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lut_floats = 9 * p.t * (p.ncls + 1) * 16;
  const int wd_floats = 9 * p.t * 16;
  float* lut = reinterpret_cast<float*>(smraw);
  float* wd = lut + lut_floats;
  float* bias = wd + wd_floats;
  float* dn = bias + 16;                                  // [t][F_IH][F_IW]
  uint8_t* lab = reinterpret_cast<uint8_t*>(dn + p.t * F_IH * F_IW);   // [t][F_IH][F_IW]
  const int tid = threadIdx.x;
  for (int i = tid; i < lut_floats + wd_floats + 16; i += blockDim.x) lut[i] = p.tab[i];
  const int tiles_x = (p.Wo + F_TW - 1) / F_TW;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
  const int img = blockIdx.y;
  const int oy0 = ty * F_TH, ox0 = tx * F_TW;
  const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 1;
  const size_t N = (size_t)p.H * p.W;
  for (int i = tid; i < p.t * F_IH * F_IW; i += blockDim.x) {
    const int f = i / (F_IH * F_IW);
    const int r = i % (F_IH * F_IW);
    const int hy = r / F_IW, hx = r % F_IW;
    const int iy = iy0 + hy, ix = ix0 + hx;
    uint8_t l = (uint8_t)p.ncls;   // zero row: padding or class id >= num_classes
    float d = 0.f;
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
      const size_t o = ((size_t)img * p.t + f) * N + (size_t)iy * p.W + ix;
      uint8_t lv = p.labels[o];
      if (lv < p.ncls) l = lv;
      if (p.use_depth) {
        // bg_model.py:50-51,67-68: (d - mean) / std, then * mask
        float v = __fdiv_rn(__fadd_rn(p.depth[o], -p.mean), p.std);
        d = p.mask[o] ? v : __fmul_rn(v, 0.0f);
      }
    }
    lab[i] = l;
    dn[i] = d;
  }
  __syncthreads();
  const int py = tid / F_TW, px = tid % F_TW;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = bias[j];
  for (int f = 0; f < p.t; ++f) {
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int tap = dy * 3 + dx;
        const int hi = f * F_IH * F_IW + (py * 2 + dy) * F_IW + px * 2 + dx;
        const int l = lab[hi];
        const float d = dn[hi];
        const float4* row = reinterpret_cast<const float4*>(lut + ((tap * p.t + f) * (p.ncls + 1) + l) * 16);
        const float4* wr = reinterpret_cast<const float4*>(wd + (tap * p.t + f) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = row[q];
          const float4 w = wr[q];
          acc[q * 4 + 0] = fmaf(d, w.x, acc[q * 4 + 0] + a.x);
          acc[q * 4 + 1] = fmaf(d, w.y, acc[q * 4 + 1] + a.y);
          acc[q * 4 + 2] = fmaf(d, w.z, acc[q * 4 + 2] + a.z);
          acc[q * 4 + 3] = fmaf(d, w.w, acc[q * 4 + 3] + a.w);
        }
      }
    }
  }
  const int oy = oy0 + py, ox = ox0 + px;
  if (oy < p.Ho && ox < p.Wo) {
    float4* o = reinterpret_cast<float4*>(p.out + (((size_t)img * p.Ho + oy) * p.Wo + ox) * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      o[q] = make_float4(fmaxf(acc[q * 4 + 0], 0.f), fmaxf(acc[q * 4 + 1], 0.f), fmaxf(acc[q * 4 + 2], 0.f),
                         fmaxf(acc[q * 4 + 3], 0.f));
  }
}

// AvgPool2d(2,2) (hardnet.py:296) NHWC slice -> NHWC slice, 4 channels per thread.
__global__ void avgpool2_kernel(const float* __restrict__ in, int in_cs, size_t in_img, float* __restrict__ out,
                                int out_cs, size_t out_img, int b, int Ho, int Wo, int c4) {
  const size_t total = (size_t)b * Ho * Wo * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    size_t r = i / c4;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho);
    const int img = (int)(r / Ho);
    const int Wi = Wo * 2;
    const float* p0 = in + img * in_img + ((size_t)(2 * y) * Wi + 2 * x) * in_cs + c * 4;
    const float4 a = *reinterpret_cast<const float4*>(p0);
    const float4 bq = *reinterpret_cast<const float4*>(p0 + in_cs);
    const float4 cq = *reinterpret_cast<const float4*>(p0 + (size_t)Wi * in_cs);
    const float4 d = *reinterpret_cast<const float4*>(p0 + (size_t)Wi * in_cs + in_cs);
    float4 o;
    o.x = (a.x + bq.x + cq.x + d.x) * 0.25f;
    o.y = (a.y + bq.y + cq.y + d.y) * 0.25f;
    o.z = (a.z + bq.z + cq.z + d.z) * 0.25f;
    o.w = (a.w + bq.w + cq.w + d.w) * 0.25f;
    *reinterpret_cast<float4*>(out + img * out_img + ((size_t)y * Wo + x) * out_cs + c * 4) = o;
  }
}

// Bilinear align_corners=True upsample (hardnet.py:249-254) of a list of channel slices into one
// contiguous NHWC buffer.  Index/weight arithmetic follows ATen's area_pixel_compute_source_index.
struct UpParams {
  SegView segs[kMaxSegs];
  size_t in_img[kMaxSegs];
  int nseg;
  float* out; int out_cs; size_t out_img;
  int b, Hi, Wi, Ho, Wo, c4_total;
  float sh, sw;
};

__global__ void upsample_bilinear_kernel(UpParams p) {
  const size_t total = (size_t)p.b * p.Ho * p.Wo * p.c4_total;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % p.c4_total) * 4;
    size_t r = i / p.c4_total;
    const int x = (int)(r % p.Wo); r /= p.Wo;
    const int y = (int)(r % p.Ho);
    const int img = (int)(r / p.Ho);
    const int cout = c;
    int s = 0;
    while (c >= p.segs[s].cpad) { c -= p.segs[s].cpad; ++s; }
    const float fy = p.sh * (float)y, fx = p.sw * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < p.Hi - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wi - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* base = p.segs[s].base + img * p.in_img[s] + c;
    const int cs = p.segs[s].cstride;
    const float4 v00 = *reinterpret_cast<const float4*>(base + ((size_t)y0 * p.Wi + x0) * cs);
    const float4 v01 = *reinterpret_cast<const float4*>(base + ((size_t)y0 * p.Wi + x1) * cs);
    const float4 v10 = *reinterpret_cast<const float4*>(base + ((size_t)y1 * p.Wi + x0) * cs);
    const float4 v11 = *reinterpret_cast<const float4*>(base + ((size_t)y1 * p.Wi + x1) * cs);
    float4 o;
    o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    *reinterpret_cast<float4*>(p.out + img * p.out_img + ((size_t)y * p.Wo + x) * p.out_cs + cout) = o;
  }
}

// K5: fused bilinear(align_corners) x4 upsample + argmax (hardnet.py:373-377 + bg_model.py:98).
// Reads the small quarter-resolution logits (L2-resident), writes only the label map unless the
// caller asks for the full-resolution logits.  NHWC16 = internal layout, else NCHW fp32.
template <bool NHWC16>
__global__ void upsample_argmax_kernel(const float* __restrict__ q, int b, int ncls, int h, int w, int fh, int fw,
                                       float sh, float sw, uint8_t* __restrict__ seg8,
                                       long long* __restrict__ seg64, float* __restrict__ full) {
  const size_t total = (size_t)b * fh * fw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % fw);
    size_t r = i / fw;
    const int y = (int)(r % fh);
    const int img = (int)(r / fh);
    const float fy = sh * (float)y, fx = sw * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < ncls; ++c) {
      float v00, v01, v10, v11;
      if (NHWC16) {
        const float* base = q + (size_t)img * h * w * 16 + c;
        v00 = base[((size_t)y0 * w + x0) * 16]; v01 = base[((size_t)y0 * w + x1) * 16];
        v10 = base[((size_t)y1 * w + x0) * 16]; v11 = base[((size_t)y1 * w + x1) * 16];
      } else {
        const float* base = q + ((size_t)img * ncls + c) * h * w;
        v00 = base[(size_t)y0 * w + x0]; v01 = base[(size_t)y0 * w + x1];
        v10 = base[(size_t)y1 * w + x0]; v11 = base[(size_t)y1 * w + x1];
      }
      const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
      if (full) full[(((size_t)img * ncls + c) * fh + y) * fw + x] = v;
      if (v > best) { best = v; arg = c; }
    }
    if (seg8) seg8[i] = (uint8_t)arg;
    if (seg64) seg64[i] = arg;
  }
}

__global__ void nhwc16_to_nchw_kernel(const float* __restrict__ q, float* __restrict__ out, int b, int ncls, int h,
                                      int w) {
  const size_t total = (size_t)b * ncls * h * w;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    size_t r = i / w;
    const int y = (int)(r % h); r /= h;
    const int c = (int)(r % ncls);
    const int img = (int)(r / ncls);
    out[i] = q[(((size_t)img * h + y) * w + x) * 16 + c];
  }
}

// debug helpers: NCHW <-> NHWC(slice)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int b, int c, int h, int w,
                                    int cs) {
  const size_t total = (size_t)b * h * w * cs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cs);
    size_t r = i / cs;
    const int x = (int)(r % w); r /= w;
    const int y = (int)(r % h);
    const int img = (int)(r / h);
    out[i] = ch < c ? in[(((size_t)img * c + ch) * h + y) * w + x] : 0.f;
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int b, int c, int h, int w,
                                    int cs) {
  const size_t total = (size_t)b * c * h * w;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    size_t r = i / w;
    const int y = (int)(r % h); r /= h;
    const int ch = (int)(r % c);
    const int img = (int)(r / c);
    out[i] = in[(((size_t)img * h + y) * w + x) * cs + ch];
  }
}

static int grid_for(size_t total, int threads) {
  size_t g = (total + threads - 1) / threads;
  const size_t cap = (size_t)kNumSMs * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static size_t plan_offsets(const pf_bgnet* net, int H, int W, std::vector<size_t>* img_floats) {
  // returns total floats per image; buffers laid out back to back (256-byte aligned)
  size_t off = 0;
  img_floats->resize(net->bufs.size());
  for (size_t i = 0; i < net->bufs.size(); ++i) {
    const BufDesc& bd = net->bufs[i];
    size_t n = (size_t)(H >> bd.shift) * (W >> bd.shift) * bd.cstride;
    (*img_floats)[i] = n;
    off += align_up(n, 64);
  }
  return off;
}

}  // namespace pf

// ==========================================================================================
extern "C" int pf_version(void) { return 100; }
extern "C" const char* pf_last_error(void) { return pf::g_err; }

extern "C" int pf_bgnet_create(pf_bgnet_t** out, int num_classes, int num_inputs, int use_depth, int precision) {
  PF_REQUIRE(out, PF_EINVAL, "pf_bgnet_create: null out");
  PF_REQUIRE(num_classes > 0 && num_classes <= 16, PF_EINVAL, "pf_bgnet_create: num_classes must be 1..16");
  PF_REQUIRE(num_inputs > 0 && num_inputs <= 8, PF_EINVAL, "pf_bgnet_create: num_inputs must be 1..8");
  PF_REQUIRE(precision == 0 || precision == 1, PF_EINVAL, "pf_bgnet_create: precision must be 0 or 1");
  pf_bgnet* net = new pf_bgnet();
  net->num_classes = num_classes; net->num_inputs = num_inputs; net->use_depth = use_depth ? 1 : 0;
  net->precision = precision;
  build_topology(net);
  *out = net;
  return 0;
}

extern "C" void pf_bgnet_destroy(pf_bgnet_t* net) {
  if (!net) return;
  for (auto& c : net->convs) {
    if (c.w_dev) cudaFree(c.w_dev);
    if (c.bias_dev) cudaFree(c.bias_dev);
  }
  if (net->first_tab_dev) cudaFree(net->first_tab_dev);
  for (auto e : net->prof_ev) cudaEventDestroy(e);
  delete net;
}

extern "C" int pf_bgnet_num_convs(const pf_bgnet_t* net) { return net ? (int)net->convs.size() - 1 : PF_EINVAL; }

extern "C" int pf_bgnet_conv_info(const pf_bgnet_t* net, int i, pf_conv_info_t* info) {
  PF_REQUIRE(net && info && i >= 0 && i < (int)net->convs.size(), PF_EINVAL, "pf_bgnet_conv_info: bad index");
  const ConvDesc& c = net->convs[i];
  info->cin = c.cin; info->cout = c.cout; info->ksize = c.ksize; info->stride = c.stride;
  memset(info->name, 0, sizeof(info->name));
  strncpy(info->name, c.name.c_str(), sizeof(info->name) - 1);
  return 0;
}

static int upload_conv(pf_bgnet* net, int i, const std::vector<double>& wfold /*[cout][cin][k][k]*/,
                       const std::vector<double>& bfold) {
  ConvDesc& c = net->convs[i];
  const int taps = c.ksize * c.ksize;
  if (i == net->first_conv) {
    // tables for K2: lut[tap][f][cls+1][16] | wd[tap][f][16] | bias[16]
    const int t = net->num_inputs, C = net->num_classes;
    const size_t lut_n = (size_t)9 * t * (C + 1) * 16, wd_n = (size_t)9 * t * 16;
    std::vector<float> tab(lut_n + wd_n + 16, 0.f);
    for (int tap = 0; tap < 9; ++tap)
      for (int f = 0; f < t; ++f) {
        for (int cls = 0; cls < C; ++cls)
          for (int o = 0; o < c.cout; ++o)
            tab[((size_t)(tap * t + f) * (C + 1) + cls) * 16 + o] =
                (float)wfold[((size_t)o * c.cin + f * C + cls) * 9 + tap];
        if (net->use_depth)
          for (int o = 0; o < c.cout; ++o)
            tab[lut_n + (size_t)(tap * t + f) * 16 + o] = (float)wfold[((size_t)o * c.cin + t * C + f) * 9 + tap];
      }
    for (int o = 0; o < c.cout; ++o) tab[lut_n + wd_n + o] = (float)bfold[o];
    if (!net->first_tab_dev) PF_CHECK_CUDA(cudaMalloc(&net->first_tab_dev, tab.size() * sizeof(float)));
    net->first_tab_floats = tab.size();
    PF_CHECK_CUDA(cudaMemcpy(net->first_tab_dev, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    c.loaded = true;
    return 0;
  }
  c.w_host.assign((size_t)taps * c.kpad * c.coutpad, 0.f);
  c.bias_host.assign(c.coutpad, 0.f);
  int kp = 0, ci = 0;
  for (auto& s : c.in) {
    for (int ch = 0; ch < s.c; ++ch, ++ci)
      for (int tap = 0; tap < taps; ++tap)
        for (int o = 0; o < c.cout; ++o)
          c.w_host[((size_t)tap * c.kpad + kp + ch) * c.coutpad + o] = (float)wfold[((size_t)o * c.cin + ci) * taps + tap];
    kp += s.cpad();
  }
  for (int o = 0; o < c.cout; ++o) c.bias_host[o] = (float)bfold[o];
  if (!c.w_dev) PF_CHECK_CUDA(cudaMalloc(&c.w_dev, c.w_host.size() * sizeof(float)));
  if (!c.bias_dev) PF_CHECK_CUDA(cudaMalloc(&c.bias_dev, c.bias_host.size() * sizeof(float)));
  PF_CHECK_CUDA(cudaMemcpy(c.w_dev, c.w_host.data(), c.w_host.size() * sizeof(float), cudaMemcpyHostToDevice));
  PF_CHECK_CUDA(cudaMemcpy(c.bias_dev, c.bias_host.data(), c.bias_host.size() * sizeof(float), cudaMemcpyHostToDevice));
  c.loaded = true;
  return 0;
}

extern "C" int pf_bgnet_load_conv(pf_bgnet_t* net, int i, const float* weight, const float* bn_weight,
                                  const float* bn_bias, const float* bn_mean, const float* bn_var, float eps) {
  PF_REQUIRE(net && weight && bn_weight && bn_bias && bn_mean && bn_var, PF_EINVAL, "pf_bgnet_load_conv: null pointer");
  PF_REQUIRE(i >= 0 && i < (int)net->convs.size() - 1, PF_EINVAL, "pf_bgnet_load_conv: bad conv index %d", i);
  const ConvDesc& c = net->convs[i];
  const int taps = c.ksize * c.ksize;
  std::vector<double> wf((size_t)c.cout * c.cin * taps), bf(c.cout);
  for (int o = 0; o < c.cout; ++o) {
    // eval-mode BatchNorm2d folded into the conv (hardnet.py:19-22)
    const double s = (double)bn_weight[o] / sqrt((double)bn_var[o] + (double)eps);
    bf[o] = (double)bn_bias[o] - (double)bn_mean[o] * s;
    for (size_t k = 0; k < (size_t)c.cin * taps; ++k) wf[(size_t)o * c.cin * taps + k] = (double)weight[(size_t)o * c.cin * taps + k] * s;
  }
  return upload_conv(net, i, wf, bf);
}

extern "C" int pf_bgnet_load_final(pf_bgnet_t* net, const float* weight, const float* bias) {
  PF_REQUIRE(net && weight && bias, PF_EINVAL, "pf_bgnet_load_final: null pointer");
  const ConvDesc& c = net->convs[net->final_conv];
  std::vector<double> wf((size_t)c.cout * c.cin), bf(c.cout);
  for (size_t k = 0; k < wf.size(); ++k) wf[k] = weight[k];
  for (int o = 0; o < c.cout; ++o) bf[o] = bias[o];
  return upload_conv(net, net->final_conv, wf, bf);
}

extern "C" int pf_bgnet_set_depth_norm(pf_bgnet_t* net, float mean, float std) {
  PF_REQUIRE(net, PF_EINVAL, "pf_bgnet_set_depth_norm: null handle");
  net->depth_mean = mean; net->depth_std = std; net->depth_norm_set = true;
  return 0;
}

extern "C" size_t pf_bgnet_workspace_bytes(const pf_bgnet_t* net, int b, int H, int W) {
  if (!net || b <= 0 || H <= 0 || W <= 0 || H % 64 || W % 64) return 0;
  std::vector<size_t> img;
  size_t per_img = plan_offsets(net, H, W, &img);
  return per_img * b * sizeof(float) + 256;
}

namespace pf {

struct Arena {
  float* base;
  int b, H, W;
  std::vector<size_t> buf_off;   // floats, start of buffer (all images contiguous per buffer)
  std::vector<size_t> img_floats;
};

static void make_arena(const pf_bgnet* net, void* ws, int b, int H, int W, Arena* a) {
  a->base = reinterpret_cast<float*>(align_up((size_t)ws, 256));
  a->b = b; a->H = H; a->W = W;
  plan_offsets(net, H, W, &a->img_floats);
  a->buf_off.resize(net->bufs.size());
  size_t off = 0;
  for (size_t i = 0; i < net->bufs.size(); ++i) {
    a->buf_off[i] = off;
    off += align_up(a->img_floats[i], 64) * b;
  }
}

static void fill_conv_launch(const pf_bgnet* net, const Arena& a, const ConvDesc& c, ConvLaunch* L) {
  L->nseg = (int)c.in.size();
  for (int s = 0; s < L->nseg; ++s) {
    const SegRef& r = c.in[s];
    L->segs[s].base = a.base + a.buf_off[r.buf] + r.coff;
    L->segs[s].cstride = net->bufs[r.buf].cstride;
    L->segs[s].cpad = r.cpad();
    L->in_img_stride[s] = align_up(a.img_floats[r.buf], 64);
  }
  const BufDesc& ib = net->bufs[c.in[0].buf];
  const BufDesc& ob = net->bufs[c.out.buf];
  L->b = a.b;
  L->Hin = a.H >> ib.shift; L->Win = a.W >> ib.shift;
  L->Hout = a.H >> ob.shift; L->Wout = a.W >> ob.shift;
  L->out = a.base + a.buf_off[c.out.buf] + c.out.coff;
  L->out_cstride = ob.cstride;
  L->out_img_stride = align_up(a.img_floats[c.out.buf], 64);
  L->w = c.w_dev; L->bias = c.bias_dev;
  L->kpad = c.kpad; L->coutpad = c.coutpad;
  L->cout_store = pad8(c.cout);
  L->relu = c.relu ? 1 : 0;
}

}  // namespace pf

extern "C" int pf_bgnet_launches_per_forward(const pf_bgnet_t* net) {
  if (!net) return PF_EINVAL;
  int n = 0;
  for (auto& s : net->steps) n += (s.type == STEP_HEAD) ? 2 : 1;
  return n;
}

extern "C" int pf_bgnet_forward(pf_bgnet_t* net, const uint8_t* labels_dev, const float* depth_dev,
                                const uint8_t* mask_dev, int b, int H, int W, int final_h, int final_w,
                                uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev, float* out_quarter_dev,
                                float* out_full_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  PF_REQUIRE(net && labels_dev && workspace_dev, PF_EINVAL, "pf_bgnet_forward: null pointer");
  PF_REQUIRE(!net->use_depth || (depth_dev && mask_dev), PF_EINVAL, "pf_bgnet_forward: depth inputs required");
  PF_REQUIRE(b > 0 && H > 0 && W > 0 && H % 64 == 0 && W % 64 == 0, PF_EINVAL,
             "pf_bgnet_forward: H and W must be positive multiples of 64 (got %dx%d)", H, W);
  PF_REQUIRE(final_h > 0 && final_w > 0, PF_EINVAL, "pf_bgnet_forward: bad final size");
  PF_REQUIRE(workspace_bytes >= pf_bgnet_workspace_bytes(net, b, H, W), PF_ENOMEM, "pf_bgnet_forward: workspace too small");
  for (auto& c : net->convs) PF_REQUIRE(c.loaded, PF_ESTATE, "pf_bgnet_forward: weights of %s not loaded", c.name.c_str());
  PF_REQUIRE(!net->use_depth || net->depth_norm_set, PF_ESTATE, "pf_bgnet_forward: depth norm not set");
  cudaStream_t st = (cudaStream_t)stream;
  Arena a;
  make_arena(net, workspace_dev, b, H, W, &a);

  const int nsteps = (int)net->steps.size();
  const bool prof = net->prof_iter < net->prof_cap;
  cudaEvent_t* pev = prof ? &net->prof_ev[(size_t)net->prof_iter * (nsteps + 1)] : nullptr;
  int step_i = 0;
  for (const Step& s : net->steps) {
    if (prof) PF_CHECK_CUDA(cudaEventRecord(pev[step_i], st));
    ++step_i;
    switch (s.type) {
      case STEP_FIRST: {
        const ConvDesc& c = net->convs[s.conv];
        FirstParams p;
        p.labels = labels_dev; p.depth = depth_dev; p.mask = mask_dev; p.tab = net->first_tab_dev;
        p.out = a.base + a.buf_off[c.out.buf];
        p.b = b; p.t = net->num_inputs; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2;
        p.ncls = net->num_classes; p.use_depth = net->use_depth; p.mean = net->depth_mean; p.std = net->depth_std;
        const size_t smem = net->first_tab_floats * 4 + (size_t)p.t * F_IH * F_IW * 5 + 16;
        static bool attr_set = false;
        if (!attr_set) {
          PF_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
          attr_set = true;
        }
        PF_REQUIRE(smem <= 160 * 1024, PF_EINVAL, "pf_bgnet_forward: first-conv tables too large");
        dim3 grid(cdiv(p.Ho, F_TH) * cdiv(p.Wo, F_TW), b);
        first_conv_kernel<<<grid, F_TH * F_TW, smem, st>>>(p);
        PF_CHECK_CUDA(cudaGetLastError());
        break;
      }
      case STEP_CONV: {
        const ConvDesc& c = net->convs[s.conv];
        ConvLaunch L;
        fill_conv_launch(net, a, c, &L);
        int rc = launch_conv_simt(L, c.ksize, c.stride, st);
        if (rc) return rc;
        break;
      }
      case STEP_POOL: {
        const SegRef& in = s.in[0];
        const BufDesc& ib = net->bufs[in.buf];
        const BufDesc& ob = net->bufs[s.out.buf];
        const int Ho = H >> ob.shift, Wo = W >> ob.shift;
        const int c4 = in.cpad() / 4;
        const size_t total = (size_t)b * Ho * Wo * c4;
        avgpool2_kernel<<<grid_for(total, 256), 256, 0, st>>>(
            a.base + a.buf_off[in.buf] + in.coff, ib.cstride, align_up(a.img_floats[in.buf], 64),
            a.base + a.buf_off[s.out.buf] + s.out.coff, ob.cstride, align_up(a.img_floats[s.out.buf], 64), b, Ho, Wo, c4);
        PF_CHECK_CUDA(cudaGetLastError());
        break;
      }
      case STEP_UPSAMPLE: {
        UpParams p;
        p.nseg = (int)s.in.size();
        int ctot = 0;
        for (int k = 0; k < p.nseg; ++k) {
          const SegRef& r = s.in[k];
          p.segs[k].base = a.base + a.buf_off[r.buf] + r.coff;
          p.segs[k].cstride = net->bufs[r.buf].cstride;
          p.segs[k].cpad = r.cpad();
          p.in_img[k] = align_up(a.img_floats[r.buf], 64);
          ctot += r.cpad();
        }
        // NOTE: the upsampled buffer is the contiguous concatenation of the PADDED input slices, so
        // its consumer (conv1x1_up) must see the same padded channel positions -> handled below.
        const BufDesc& ib = net->bufs[s.in[0].buf];
        const BufDesc& ob = net->bufs[s.out.buf];
        p.out = a.base + a.buf_off[s.out.buf]; p.out_cs = ob.cstride; p.out_img = align_up(a.img_floats[s.out.buf], 64);
        p.b = b; p.Hi = H >> ib.shift; p.Wi = W >> ib.shift; p.Ho = H >> ob.shift; p.Wo = W >> ob.shift;
        p.c4_total = ctot / 4;
        p.sh = p.Ho > 1 ? (float)(p.Hi - 1) / (float)(p.Ho - 1) : 0.f;
        p.sw = p.Wo > 1 ? (float)(p.Wi - 1) / (float)(p.Wo - 1) : 0.f;
        const size_t total = (size_t)b * p.Ho * p.Wo * p.c4_total;
        upsample_bilinear_kernel<<<grid_for(total, 256), 256, 0, st>>>(p);
        PF_CHECK_CUDA(cudaGetLastError());
        break;
      }
      case STEP_HEAD: {
        const ConvDesc& c = net->convs[s.conv];
        ConvLaunch L;
        fill_conv_launch(net, a, c, &L);
        L.cout_store = 16;
        int rc = launch_conv_simt(L, 1, 1, st);
        if (rc) return rc;
        const int h = H / 4, w = W / 4;
        const float* q = a.base + a.buf_off[net->quarter_buf];
        if (out_quarter_dev) {
          const size_t total = (size_t)b * net->num_classes * h * w;
          nhwc16_to_nchw_kernel<<<grid_for(total, 256), 256, 0, st>>>(q, out_quarter_dev, b, net->num_classes, h, w);
          PF_CHECK_CUDA(cudaGetLastError());
        }
        const float sh = final_h > 1 ? (float)(h - 1) / (float)(final_h - 1) : 0.f;
        const float sw = final_w > 1 ? (float)(w - 1) / (float)(final_w - 1) : 0.f;
        const size_t total = (size_t)b * final_h * final_w;
        upsample_argmax_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(
            q, b, net->num_classes, h, w, final_h, final_w, sh, sw, out_seg_u8_dev, (long long*)out_seg_i64_dev, out_full_dev);
        PF_CHECK_CUDA(cudaGetLastError());
        break;
      }
    }
  }
  if (prof) {
    PF_CHECK_CUDA(cudaEventRecord(pev[nsteps], st));
    net->prof_iter++;
  }
  return 0;
}

extern "C" int pf_bgnet_set_profiling(pf_bgnet_t* net, int max_iters) {
  PF_REQUIRE(net && max_iters >= 0, PF_EINVAL, "pf_bgnet_set_profiling: bad argument");
  for (auto e : net->prof_ev) cudaEventDestroy(e);
  net->prof_ev.clear();
  net->prof_cap = max_iters; net->prof_iter = 0;
  const size_t n = (size_t)max_iters * (net->steps.size() + 1);
  net->prof_ev.resize(n);
  for (size_t i = 0; i < n; ++i) PF_CHECK_CUDA(cudaEventCreate(&net->prof_ev[i]));
  return 0;
}

extern "C" int pf_bgnet_num_steps(const pf_bgnet_t* net) { return net ? (int)net->steps.size() : PF_EINVAL; }

extern "C" int pf_bgnet_step_info(const pf_bgnet_t* net, int k, int* type, int* conv_index) {
  PF_REQUIRE(net && type && conv_index && k >= 0 && k < (int)net->steps.size(), PF_EINVAL, "pf_bgnet_step_info: bad index");
  *type = (int)net->steps[k].type; *conv_index = net->steps[k].conv;
  return 0;
}

extern "C" int pf_bgnet_read_profile(pf_bgnet_t* net, float* ms_per_step, int cap) {
  PF_REQUIRE(net && ms_per_step, PF_EINVAL, "pf_bgnet_read_profile: null pointer");
  const int nsteps = (int)net->steps.size();
  PF_REQUIRE(cap >= nsteps, PF_EINVAL, "pf_bgnet_read_profile: cap < num_steps");
  for (int k = 0; k < nsteps; ++k) ms_per_step[k] = 0.f;
  if (net->prof_iter == 0) return 0;
  PF_CHECK_CUDA(cudaEventSynchronize(net->prof_ev[(size_t)(net->prof_iter - 1) * (nsteps + 1) + nsteps]));
  for (int it = 0; it < net->prof_iter; ++it)
    for (int k = 0; k < nsteps; ++k) {
      float ms = 0.f;
      PF_CHECK_CUDA(cudaEventElapsedTime(&ms, net->prof_ev[(size_t)it * (nsteps + 1) + k], net->prof_ev[(size_t)it * (nsteps + 1) + k + 1]));
      ms_per_step[k] += ms / net->prof_iter;
    }
  return net->prof_iter;
}

extern "C" int pf_upsample_argmax(const float* logits_nchw_dev, int b, int classes, int h, int w, int final_h,
                                  int final_w, uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev,
                                  float* out_full_dev, void* stream) {
  PF_REQUIRE(logits_nchw_dev && (out_seg_u8_dev || out_seg_i64_dev || out_full_dev), PF_EINVAL,
             "pf_upsample_argmax: null pointer");
  PF_REQUIRE(b > 0 && classes > 0 && h > 0 && w > 0 && final_h > 0 && final_w > 0, PF_EINVAL, "pf_upsample_argmax: bad size");
  const float sh = final_h > 1 ? (float)(h - 1) / (float)(final_h - 1) : 0.f;
  const float sw = final_w > 1 ? (float)(w - 1) / (float)(final_w - 1) : 0.f;
  const size_t total = (size_t)b * final_h * final_w;
  upsample_argmax_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      logits_nchw_dev, b, classes, h, w, final_h, final_w, sh, sw, out_seg_u8_dev, (long long*)out_seg_i64_dev, out_full_dev);
  PF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pf_bgnet_debug_conv(pf_bgnet_t* net, int i, const float* x_nchw_dev, int b, int H, int W,
                                   float* y_nchw_dev, void* stream) {
  PF_REQUIRE(net && x_nchw_dev && y_nchw_dev, PF_EINVAL, "pf_bgnet_debug_conv: null pointer");
  PF_REQUIRE(i > 0 && i < (int)net->convs.size(), PF_EINVAL, "pf_bgnet_debug_conv: index must be 1..num_convs");
  const ConvDesc& c = net->convs[i];
  PF_REQUIRE(c.loaded, PF_ESTATE, "pf_bgnet_debug_conv: weights not loaded");
  cudaStream_t st = (cudaStream_t)stream;
  // input: one contiguous NHWC buffer whose channel layout is the padded concatenation of the segs
  const int Ho = (H + c.stride - 1) / c.stride, Wo = (W + c.stride - 1) / c.stride;
  const int cs_out = pad8(c.cout) < 16 && i == net->final_conv ? 16 : pad8(c.cout);
  float *xin = nullptr, *yout = nullptr, *xpk = nullptr;
  PF_CHECK_CUDA(cudaMalloc(&xin, (size_t)b * H * W * c.kpad * 4));
  PF_CHECK_CUDA(cudaMalloc(&xpk, (size_t)b * H * W * c.cin * 4));
  PF_CHECK_CUDA(cudaMalloc(&yout, (size_t)b * Ho * Wo * cs_out * 4));
  PF_CHECK_CUDA(cudaMemsetAsync(xin, 0, (size_t)b * H * W * c.kpad * 4, st));
  // NCHW -> dense NHWC (cin), then scatter each seg into its padded position
  {
    const size_t total = (size_t)b * H * W * c.cin;
    nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, st>>>(x_nchw_dev, xpk, b, c.cin, H, W, c.cin);
    int src = 0, dst = 0;
    for (auto& s : c.in) {
      PF_CHECK_CUDA(cudaMemcpy2DAsync(xin + dst, (size_t)c.kpad * 4, xpk + src, (size_t)c.cin * 4, (size_t)s.c * 4,
                                      (size_t)b * H * W, cudaMemcpyDeviceToDevice, st));
      src += s.c; dst += s.cpad();
    }
  }
  ConvLaunch L;
  L.nseg = (int)c.in.size();
  int dst = 0;
  for (int s = 0; s < L.nseg; ++s) {
    L.segs[s].base = xin + dst; L.segs[s].cstride = c.kpad; L.segs[s].cpad = c.in[s].cpad();
    L.in_img_stride[s] = (size_t)H * W * c.kpad;
    dst += c.in[s].cpad();
  }
  L.b = b; L.Hin = H; L.Win = W; L.Hout = Ho; L.Wout = Wo;
  L.out = yout; L.out_cstride = cs_out; L.out_img_stride = (size_t)Ho * Wo * cs_out;
  L.w = c.w_dev; L.bias = c.bias_dev; L.kpad = c.kpad; L.coutpad = c.coutpad; L.cout_store = cs_out; L.relu = c.relu ? 1 : 0;
  int rc = launch_conv_simt(L, c.ksize, c.stride, st);
  if (rc == 0) {
    const size_t total = (size_t)b * c.cout * Ho * Wo;
    nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, st>>>(yout, y_nchw_dev, b, c.cout, Ho, Wo, cs_out);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("pf_bgnet_debug_conv: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  cudaFree(xin); cudaFree(xpk); cudaFree(yout);
  return rc;
}
