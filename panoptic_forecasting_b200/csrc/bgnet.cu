// Stage B: BGModel / HarDNet-70 forward behind the C ABI.
// Reference: bg_model.py:53-71,91-102; hardnet.py:176-240 (HarDBlock), :243-258 (TransitionUp),
// :262-327 (topology), :353-387 (forward).
//
// Layout: every activation is fp32 NHWC inside one caller-provided arena; every HarDBlock owns a
// single buffer whose channel slots are [block input | layer1 | layer2 | ...], so the reference's
// torch.cat calls disappear: a consumer reads a list of channel slices (SegRef), a producer writes
// its slice.  Slots are aligned/padded to 16 channels; producers write zeros into the pad channels.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "bgnet.h"
#include "conv_tc.h"
#include "split_bf16.cuh"
#include "tc_common.cuh"

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

static int padc(int c) { return (c + kChanAlign - 1) / kChanAlign * kChanAlign; }

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
  std::vector<float> first_wd_host;          // first conv: depth-plane weights + bias, passed to the kernel by value
  int launches = 0;
  // tensor-core path (precision 1): packed split-bf16 weights live in ConvDesc-indexed arrays; the
  // TMA tensor maps depend on the arena address and shape, so they are cached per (ws, b, H, W).
  std::vector<__nv_bfloat16*> wtc_dev;      // per conv: hi [nrows][ktot] then lo [nrows][ktot]
  std::vector<int> wtc_rows;
  struct TcPlan {
    const void* ws = nullptr; int b = 0, H = 0, W = 0;
    CUtensorMap* maps_dev = nullptr;
    std::vector<TcLayer> layers;             // indexed by conv (per-tap kernel: 1x1 convs, fallbacks)
    std::vector<HaloLayer> halos;            // indexed by conv (halo kernel: 3x3 stride-1 and 1x1 convs)
    std::vector<HaloLayer> halos_low;        // indexed by conv: low-resolution half of a fused conv1x1_up
    std::vector<char> pool_fused;            // indexed by step: the pool runs in the producing conv's epilogue
    int head_amax = 0;                       // the head conv parks each quarter-resolution pixel's argmax in channel 15
    std::vector<int> nblocks_low;
    std::vector<size_t> smem_low;
    std::vector<int> nblocks;
    std::vector<size_t> smem;
    std::vector<char> use_tc;                // 0 SIMT, 1 per-tap tcgen05, 2 halo tcgen05
  } plan;
  bool force_simt = false;                   // PF_TC_FORCE_SIMT=1: run every conv on the SIMT kernels (A/B checks)
  bool no_halo = false;                      // PF_TC_NO_HALO=1: 3x3 convs on the per-tap tcgen05 kernel
  bool compact_slots = true;                 // HarDBlock slots are separate compact tensors (PF_INTERLEAVED_SLOTS=1: one wide buffer per block)
  bool fuse_up = false;                      // conv1x1_up fused with TransitionUp (tensor-core path, see ConvDesc::up_nseg)
  float* zero_bias_dev = nullptr;
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
  // Every slot (block input, each layer's output) is its own compact NHWC tensor: a consumer's TMA boxes and a
  // producer's stores then touch whole DRAM pages instead of 32..96-byte pieces of a wide interleaved row
  // (the interleaved block buffer made the 1/4-resolution layers ~1.8x slower on B200).
  std::vector<int> bufs(n_layers + 1);
  if (net->compact_slots) {
    for (int l = 0; l <= n_layers; ++l) { off[l] = 0; bufs[l] = net->new_buf(shift, padc(ch[l])); }
  } else {
    int total = 0;
    for (int l = 0; l <= n_layers; ++l) { off[l] = total; total += padc(ch[l]); }
    const int buf = net->new_buf(shift, total);
    for (int l = 0; l <= n_layers; ++l) bufs[l] = buf;
  }
  in_slot->buf = bufs[0]; in_slot->coff = 0; in_slot->c = in_ch;
  for (int l = 1; l <= n_layers; ++l) {
    int oc; std::vector<int> link;
    get_link(l, in_ch, gr, &oc, &link);
    std::vector<SegRef> in;
    int cin = 0;
    for (int k : link) { SegRef s; s.buf = bufs[k]; s.coff = off[k]; s.c = ch[k]; in.push_back(s); cin += ch[k]; }
    SegRef out; out.buf = bufs[l]; out.coff = off[l]; out.c = oc;
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
      SegRef s; s.buf = bufs[i]; s.coff = off[i]; s.c = ch[i];
      out_segs->push_back(s);
      *out_ch_total += ch[i];
    }
  }
}

static void build_topology(pf_bgnet* net) {
  const int t = net->num_inputs;
  const int cin0 = (net->num_classes + (net->use_depth ? 1 : 0)) * t;
  // stem (hardnet.py:275-280)
  // tensor-core storage: base.1 writes space-to-depth (quarter resolution, 4 x 32 channels), see ConvDesc::s2d_out
  const bool s2d = net->precision == 1;
  int s0 = net->new_buf(1, padc(kFirstCh[0]));
  int s1 = s2d ? net->new_buf(2, 128) : net->new_buf(1, padc(kFirstCh[1]));
  int s2 = net->new_buf(2, padc(kFirstCh[2]));
  SegRef r0{s0, 0, kFirstCh[0]}, r1{s1, 0, kFirstCh[1]}, r2{s2, 0, kFirstCh[2]};
  SegRef r1_in = r1;
  if (s2d) r1_in.c = 128;
  {
    int ci = net->add_conv("model.base.0", cin0, kFirstCh[0], 3, 2, {}, r0);
    net->first_conv = ci;
    Step st; st.type = STEP_FIRST; st.conv = ci; net->steps.push_back(st);
    ci = net->add_conv("model.base.1", kFirstCh[0], kFirstCh[1], 3, 1, {r0}, r1);
    net->convs[ci].s2d_out = s2d;
    st.type = STEP_CONV; st.conv = ci; net->steps.push_back(st);
    ci = net->add_conv("model.base.2", kFirstCh[1], kFirstCh[2], 3, 2, {r1_in}, r2);
    net->convs[ci].s2d_in = s2d;
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
    int pbuf = net->new_buf(shift, padc(kChList[i]));
    SegRef pout{pbuf, 0, kChList[i]};
    int ci = net->add_conv(pfx, oc, kChList[i], 1, 1, outs, pout);
    { Step st; st.type = STEP_CONV; st.conv = ci; net->steps.push_back(st); }
    ch = kChList[i];
    if (i < 4) {
      Step st; st.type = STEP_POOL; st.in = {pout};
      net->steps.push_back(st);
      pending_pool_step = (int)net->steps.size() - 1;
      net->convs[ci].pool_step = pending_pool_step;
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
    std::vector<SegRef> cat_in;
    int ybuf = -1;
    if (net->fuse_up) {
      // conv1x1_up reads the LOW-resolution slices directly (see ConvDesc::up_nseg); no upsampled tensor
      for (auto& s : cur_segs) cat_in.push_back(s);
      ybuf = net->new_buf(shift + 1, (((cur_ch + skip_ch[i]) / 2) + 15) / 16 * 16);
      net->bufs[ybuf].always_f32 = true;
    } else {
      // the upsampled tensor keeps the (padded) slot layout of its source slices, so conv1x1_up
      // reads it as the same list of slices followed by the skip's slices (hardnet.py:256 cat order).
      int ctot = 0;
      for (auto& s : cur_segs) ctot += s.cpad();
      int ubuf = net->new_buf(shift, ctot);
      int off = 0;
      for (auto& s : cur_segs) { SegRef u{ubuf, off, s.c}; cat_in.push_back(u); off += s.cpad(); }
      SegRef uout{ubuf, 0, ctot};
      Step st; st.type = STEP_UPSAMPLE; st.in = cur_segs; st.out = uout; net->steps.push_back(st);
    }
    const int n_up = (int)cur_segs.size();
    for (auto& s : skips[i]) cat_in.push_back(s);
    const int ccat = cur_ch + skip_ch[i];
    const int chalf = ccat / 2;
    char nm[64];
    snprintf(nm, sizeof(nm), "model.conv1x1_up.%d", j);
    int c1 = net->add_conv(nm, ccat, chalf, 1, 1, cat_in, SegRef());
    if (net->fuse_up) { net->convs[c1].up_nseg = n_up; net->convs[c1].ybuf = ybuf; }
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
  net->bufs[net->quarter_buf].always_f32 = true;
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

// depth-plane weights wd[tap][frame][16] and bias[16] of the first conv are KERNEL PARAMETERS (FirstParams::wd): the
// parameter space is a constant bank, so they are uniform operands of FFMA with no shared-memory traffic, and every
// launch carries its own copy -- two nets, or one net on two streams, cannot overwrite each other (a process-wide
// __constant__ array refreshed per call could).
constexpr int kFirstMaxT = 8;

// TMA boxes must start on a 16-byte boundary of the innermost (x) dimension: the window's first column
// ix0 = 64k - 1 is odd, so the depth box starts 3 pixels earlier and the label / mask boxes 15 pixels earlier.
constexpr int F_DOFF = 3, F_LOFF = 15;
constexpr int F_DPITCH = 72;     // staged depth row: 72 floats (288 B), columns [ix0 - 3, ix0 + 69)
constexpr int F_LPITCH = 96;     // staged label / mask row: 96 bytes, columns [ix0 - 15, ix0 + 81)
constexpr int F_DFRAME = 1248;   // floats per staged depth frame (17 x 72 = 1224, padded: TMA destinations are 128-byte aligned)
constexpr int F_LFRAME = 1664;   // bytes per staged label / mask frame (17 x 96 = 1632, padded to 128)

struct FirstMaps {
  alignas(64) CUtensorMap m_depth;   // f32 (W, H, b*t), box 68 x 17 x 1
  alignas(64) CUtensorMap m_label;   // u8  (W, H, b*t), box 80 x 17 x 1
  alignas(64) CUtensorMap m_mask;    // u8  (W, H, b*t), box 80 x 17 x 1
};

struct FirstParams {
  FirstMaps maps;      // tensor maps of the caller's input tensors, by value (kernel parameter space)
  float wd[9 * kFirstMaxT * 16 + 16];   // depth-plane weights [tap][frame][16], then bias[16]
  const float* tab;    // lut | wd | bias | lutsum
  void* out;           // NHWC, 16 channels: fp32, or bf16 hi plane
  void* out_lo;        // bf16 lo plane (split storage)
  int split;
  int b, t, H, W, Ho, Wo, ncls, use_depth;
  float mean, std;
};

// The input window of a tile ((2*8+1) x (2*32+1) pixels x t frames of labels, depth, mask) is staged by
// TMA (3 boxes per frame, image borders zero-filled), then one in-place pass turns it into the two arrays
// the taps read: the table row index of every pixel and its normalised masked depth.
// PERSISTENT: a CTA loops over tiles (grid = a multiple of the SM count) with TWO staging buffers, so the boxes of
// tile k+1 are in flight while tile k is converted and convolved, and the 25 KB of tables are loaded once per CTA
// (the one-tile-per-CTA version issued only 24 % of its slots: every CTA sat through its own TMA latency).
// T > 0: the number of input frames is a compile-time constant, the frame loop unrolls and every depth-plane weight is
// an immediate constant-bank operand of its FFMA (with a runtime frame index each weight first went through a
// uniform-register load, LDCU c[0x3][UR + imm]: 72 of them per frame in front of 144 FFMAs).  T = 0: runtime p.t.
template <int T>
__global__ void __launch_bounds__(F_TH* F_TW) first_conv_kernel(const __grid_constant__ FirstParams p_in) {
  const FirstParams& p = p_in;             // stays in the parameter space (constant bank; TMA needs the maps' addresses there)
  const int pt = T > 0 ? T : p_in.t;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t bars[2];
  const int lut_floats = 9 * pt * (p.ncls + 1) * 16;
  const int wd_floats = 9 * pt * 16;
  const int sum_floats = pt * (p.ncls + 1) * 16;
  const int tab_floats = lut_floats + wd_floats + 16 + sum_floats;
  // dynamic smem: 2 x { [depth/dn: t x 17 x 72 f32][labels: t x 17 x 96 u8][mask: t x 17 x 96 u8] } [tables]
  // TMA destinations must be 128-byte aligned: align the dynamic window by hand
  unsigned char* sm0 = smraw + ((128u - ((uint32_t)__cvta_generic_to_shared(smraw) & 127u)) & 127u);
  const int stage_bytes = pt * (F_DFRAME * 4 + 2 * F_LFRAME);
  float* lut = reinterpret_cast<float*>(sm0 + 2 * stage_bytes);
  const float* lutsum = lut + lut_floats + wd_floats + 16;
  const int tid = threadIdx.x;
  const int tiles_x = (p.Wo + F_TW - 1) / F_TW, tiles_y = (p.Ho + F_TH - 1) / F_TH;
  const int tiles_per_img = tiles_x * tiles_y;
  const int total_tiles = tiles_per_img * p.b;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t stage_tx = (uint32_t)pt * F_IH * (F_DPITCH * 4 * (p.use_depth ? 1 : 0) + F_LPITCH * (p.use_depth ? 2 : 1));
  // thread 0: all boxes of tile `tile` into staging buffer `st`
  auto issue = [&](int tile, int st) {
    const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    const int iy0 = ty * F_TH * 2 - 1, ix0 = tx * F_TW * 2 - 1;
    unsigned char* base = sm0 + st * stage_bytes;
    float* dn = reinterpret_cast<float*>(base);
    uint8_t* lab = base + pt * F_DFRAME * 4;
    uint8_t* msk = lab + pt * F_LFRAME;
    const uint32_t bar_a = bar0 + 8u * st;
    tc::mbar_expect_tx(bar_a, stage_tx);
    for (int f = 0; f < pt; ++f) {
      const int z = img * pt + f;
      // label / mask planes are fetched as 32-bit words starting 15 pixels left of the window
      const int xw = (ix0 - F_LOFF) >> 2;
      const int xd = ix0 - F_DOFF;
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"((uint32_t)__cvta_generic_to_shared(lab + f * F_LFRAME)), "l"(&p.maps.m_label), "r"(bar_a),
                     "r"(xw), "r"(z * p.H + iy0) : "memory");
      if (p.use_depth) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(dn + f * F_DFRAME)), "l"(&p.maps.m_depth), "r"(bar_a),
                       "r"(xd), "r"(z * p.H + iy0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(msk + f * F_LFRAME)), "l"(&p.maps.m_mask), "r"(bar_a),
                       "r"(xw), "r"(z * p.H + iy0) : "memory");
      }
    }
  };
  if (tid == 0) {
    tc::mbar_init(bar0, 1);
    tc::mbar_init(bar0 + 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < total_tiles) issue((int)blockIdx.x, 0);
  }
  // tables while the boxes are in flight
  {
    const float4* src = reinterpret_cast<const float4*>(p.tab);     // every table is a multiple of 16 floats
    float4* dst = reinterpret_cast<float4*>(lut);
    for (int i = tid; i < tab_floats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();                 // barriers initialised / tables visible
  int k = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++k) {
  const int st = k & 1;
  if (tid == 0 && tile + (int)gridDim.x < total_tiles) {
    // the other buffer was last read (generic proxy) before the __syncthreads that ended the previous iteration
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    issue(tile + (int)gridDim.x, st ^ 1);
  }
  const int img = tile / tiles_per_img, r_ = tile - img * tiles_per_img;
  const int ty = r_ / tiles_x, tx = r_ - ty * tiles_x;
  const int oy0 = ty * F_TH, ox0 = tx * F_TW;
  const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 1;
  unsigned char* sbase = sm0 + st * stage_bytes;
  float* dn = reinterpret_cast<float*>(sbase);
  uint8_t* lab = sbase + pt * F_DFRAME * 4;
  uint8_t* msk = lab + pt * F_LFRAME;
  tc::mbar_wait(bar0 + 8u * st, (uint32_t)((k >> 1) & 1));
  // No conversion pass over the staged window (it was a third of the kernel's instructions, ncu source view): the taps
  // clamp the label to the zero row (ids >= num_classes, bg_model.py:54-55) and normalise / mask the depth
  // (bg_model.py:50-51,67-68) as they read.  Tiles on the image border first patch their out-of-image positions (label :=
  // zero row, mask := 0, depth := 0): TMA zero-fills columns outside the image, but label 0 is a real class, and the
  // rows above / below an image are the neighbouring frame's rows in the (W, H * frames) tensor maps.
  if (iy0 < 0 || ix0 < 0 || iy0 + F_IH > p.H || ix0 + F_IW > p.W) {           // tile-uniform
    for (int i = tid; i < pt * F_IH * F_IW; i += blockDim.x) {
      const int f = i / (F_IH * F_IW);
      const int r = i - f * (F_IH * F_IW);
      const int hy = r / F_IW, hx = r - hy * F_IW;
      const int iy = iy0 + hy, ix = ix0 + hx;
      if (!(iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)) {
        lab[f * F_LFRAME + hy * F_LPITCH + hx + F_LOFF] = (uint8_t)p.ncls;
        if (p.use_depth) {         // rows above / below the image are the neighbouring frame's rows in the (W, H * frames) maps
          msk[f * F_LFRAME + hy * F_LPITCH + hx + F_LOFF] = 0;
          dn[f * F_DFRAME + hy * F_DPITCH + hx + F_DOFF] = 0.f;
        }
      }
    }
    __syncthreads();
  }
  const float inv_std = __fdiv_rn(1.0f, p.std), neg_mean = -p.mean;
  const int py = tid / F_TW, px = tid % F_TW;
  // 16 output channels as 8 fp32 pairs: the table rows are added with sm_100's packed FADD2 (two IEEE fp32 adds per
  // issue slot); this kernel is issue-bound (ncu: 472 M warp instructions per 16 frames, 61 % issue-active, no memory stall)
  float2 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(p.wd[9 * pt * 16 + 2 * j], p.wd[9 * pt * 16 + 2 * j + 1]);
#pragma unroll
  for (int f = 0; f < pt; ++f) {
    const int l0 = f * F_LFRAME + (py * 2) * F_LPITCH + px * 2 + F_LOFF;
    const int d0 = f * F_DFRAME + (py * 2) * F_DPITCH + px * 2 + F_DOFF;
    int l[9];
    float d[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        l[dy * 3 + dx] = min((int)lab[l0 + dy * F_LPITCH + dx], p.ncls);
        float dv = 0.f;
        if (p.use_depth) {
          const float v = __fmul_rn(__fadd_rn(dn[d0 + dy * F_DPITCH + dx], neg_mean), inv_std);
          dv = msk[l0 + dy * F_LPITCH + dx] ? v : __fmul_rn(v, 0.0f);
        }
        d[dy * 3 + dx] = dv;
      }
    // depth planes: 9 taps x 16 FFMA with constant-bank weights (uniform-register operands).  Packed FFMA2 with the
    // weights read from shared memory (broadcast LDS.128) was slower: 0.74 vs 0.69 ms per 16 frames.
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* w = p.wd + (tap * pt + f) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j].x = fmaf(d[tap], w[2 * j], acc[j].x);
        acc[j].y = fmaf(d[tap], w[2 * j + 1], acc[j].y);
      }
    }
    // label planes: label maps are piecewise constant, so the 3x3 window of a frame is usually one label:
    // then the nine table rows collapse into one pre-summed row (4 instead of 36 128-bit shared loads).
    bool same = true;
#pragma unroll
    for (int tap = 1; tap < 9; ++tap) same = same && (l[tap] == l[0]);
    if (same) {
      const float4* row = reinterpret_cast<const float4*>(lutsum + (f * (p.ncls + 1) + l[0]) * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = row[q];
        acc[2 * q] = __fadd2_rn(acc[2 * q], make_float2(a.x, a.y));
        acc[2 * q + 1] = __fadd2_rn(acc[2 * q + 1], make_float2(a.z, a.w));
      }
    } else {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4* row = reinterpret_cast<const float4*>(lut + ((tap * pt + f) * (p.ncls + 1) + l[tap]) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = row[q];
          acc[2 * q] = __fadd2_rn(acc[2 * q], make_float2(a.x, a.y));
          acc[2 * q + 1] = __fadd2_rn(acc[2 * q + 1], make_float2(a.z, a.w));
        }
      }
    }
  }
  const int oy = oy0 + py, ox = ox0 + px;
  if (oy < p.Ho && ox < p.Wo) {
    const size_t o = (((size_t)img * p.Ho + oy) * p.Wo + ox) * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[2 * j] = fmaxf(acc[j].x, 0.f); v[2 * j + 1] = fmaxf(acc[j].y, 0.f); }
    store16_any(p.out, p.out_lo, o, v, p.split != 0);
  }
  __syncthreads();                 // every thread is done with this staging buffer before it is refilled
  }
}

// First ConvLayer from DENSE per-class planes: the reference's `convert2onehot: False` input mode (bg_model.py:61-69,
// `inps` is already a float [b, t, C, H, W] tensor of class scores / one-hot planes).  Same weight tables as the label
// form: lut[tap][frame][class] is the folded weight row of input channel frame * C + class, wd the depth planes'.  One
// output pixel per thread, 16 accumulators; the planes are read straight from global memory (stride-2 taps: half of
// every line is used, the other half by the neighbouring tap), the weight rows as broadcast 128-bit shared loads.
// Not on the benched path (the reference's bg configs all use one-hot conversion): written for coverage, not for speed.
struct FirstDenseParams {
  const float* x;        // [b, t, C, H, W]
  const float* depth;    // [b, t, H, W] or null
  const uint8_t* mask;   // [b, t, H, W] or null
  const float* tab;      // lut | wd | bias (| lutsum, unused here)
  void* out;
  void* out_lo;
  int split;
  int b, t, H, W, Ho, Wo, ncls, use_depth;
  float mean, std;
};

__global__ void __launch_bounds__(F_TH* F_TW) first_conv_dense_kernel(const FirstDenseParams p) {
  extern __shared__ __align__(16) float dtab[];
  const int lut_floats = 9 * p.t * (p.ncls + 1) * 16, wd_floats = 9 * p.t * 16;
  for (int i = threadIdx.x; i < lut_floats + wd_floats + 16; i += blockDim.x) dtab[i] = __ldg(p.tab + i);
  __syncthreads();
  const float* lut = dtab;
  const float* wd = dtab + lut_floats;
  const float* bias = wd + wd_floats;
  const int px = threadIdx.x % F_TW, py = threadIdx.x / F_TW;
  const int ox = blockIdx.x * F_TW + px, oy = blockIdx.y * F_TH + py, img = blockIdx.z;
  if (ox >= p.Wo || oy >= p.Ho) return;
  float acc[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = bias[o];
  const size_t plane = (size_t)p.H * p.W;
  int off[9];
  bool in[9];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int iy = 2 * oy + dy - 1, ix = 2 * ox + dx - 1;
      in[dy * 3 + dx] = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      off[dy * 3 + dx] = iy * p.W + ix;
    }
  auto add_plane = [&](const float* src, const float* wrow0, int wstride, bool is_depth, const uint8_t* msk) {
    float v[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      float x = in[tap] ? __ldg(src + off[tap]) : 0.f;
      if (is_depth && in[tap]) {
        // bg_model.py:50-51,67-68: (depth - mean) / std, then * depth_mask
        const float n = __fdiv_rn(__fadd_rn(x, -p.mean), p.std);
        x = __ldg(msk + off[tap]) ? n : __fmul_rn(n, 0.0f);
      }
      v[tap] = x;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float4* row = reinterpret_cast<const float4*>(wrow0 + (size_t)tap * wstride);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w = row[q];
        acc[4 * q + 0] = fmaf(v[tap], w.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(v[tap], w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v[tap], w.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v[tap], w.w, acc[4 * q + 3]);
      }
    }
  };
  for (int f = 0; f < p.t; ++f) {
    for (int cls = 0; cls < p.ncls; ++cls)
      add_plane(p.x + ((size_t)(img * p.t + f) * p.ncls + cls) * plane, lut + ((size_t)f * (p.ncls + 1) + cls) * 16,
                p.t * (p.ncls + 1) * 16, false, nullptr);
    if (p.use_depth)
      add_plane(p.depth + (size_t)(img * p.t + f) * plane, wd + (size_t)f * 16, p.t * 16, true,
                p.mask + (size_t)(img * p.t + f) * plane);
  }
  float v[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) v[o] = fmaxf(acc[o], 0.f);
  store16_any(p.out, p.out_lo, (((size_t)img * p.Ho + oy) * p.Wo + ox) * 16, v, p.split != 0);
}

// tensor maps of the caller's input tensors for first_conv_kernel (encoded per call: pointers are the caller's)
static int first_conv_maps(FirstMaps* p, const uint8_t* labels, const float* depth, const uint8_t* mask, int bt, int H,
                           int W, bool use_depth) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = (EncodeTiledFn)fp;
  }
  PF_REQUIRE(enc, PF_ESTATE, "cuTensorMapEncodeTiled not available from the driver");
  // 2-D maps over (W, H * frames): rows above / below an image belong to the neighbouring frame, the
  // kernel's in-bounds test discards them (only rows outside the whole tensor are zero-filled by TMA)
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H * bt};
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)(W / 4), (cuuint64_t)H * bt};     // bytes viewed as 32-bit words
    cuuint64_t strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {(cuuint32_t)(F_LPITCH / 4), (cuuint32_t)F_IH};
    CUresult r = enc(&p->m_label, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(labels), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(labels) failed: %d", (int)r);
    if (use_depth) {
      r = enc(&p->m_mask, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(mask), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(mask) failed: %d", (int)r);
    }
  }
  if (use_depth) {
    cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    cuuint32_t box[2] = {(cuuint32_t)F_DPITCH, (cuuint32_t)F_IH};
    CUresult r = enc(&p->m_depth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(depth), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PF_REQUIRE(r == CUDA_SUCCESS, PF_EINVAL, "cuTensorMapEncodeTiled(depth) failed: %d", (int)r);
  }
  return 0;
}

// AvgPool2d(2,2) (hardnet.py:296) NHWC slice -> NHWC slice, 4 channels per thread.
struct PoolParams {
  const void* in; const void* in_lo; int in_cs; size_t in_img;
  void* out; void* out_lo; int out_cs; size_t out_img;
  int b, Ho, Wo, Wi, c4, split;      // Wi: input row length (2 * Wo, or 2 * Wo + 1: the pool floors)
};

// grid = (ceil(Wo * groups / 256), Ho, b): one thread per (output pixel, 8-channel group)
__global__ void __launch_bounds__(256) avgpool2_kernel(PoolParams p) {
  const bool sp = p.split != 0;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.Wo * p.c4) return;
  const int x = idx / p.c4, c = idx - x * p.c4;
  const int y = blockIdx.y, img = blockIdx.z;
  const int Wi = p.Wi;
  const size_t p0 = (size_t)img * p.in_img + ((size_t)(2 * y) * Wi + 2 * x) * p.in_cs + c * 8;
  float a[8], bq[8], cq[8], d[8], o[8];
  load8_any(p.in, p.in_lo, p0, sp, a);
  load8_any(p.in, p.in_lo, p0 + p.in_cs, sp, bq);
  load8_any(p.in, p.in_lo, p0 + (size_t)Wi * p.in_cs, sp, cq);
  load8_any(p.in, p.in_lo, p0 + (size_t)Wi * p.in_cs + p.in_cs, sp, d);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = (a[k] + bq[k] + cq[k] + d[k]) * 0.25f;
  store8_any(p.out, p.out_lo, (size_t)img * p.out_img + ((size_t)y * p.Wo + x) * p.out_cs + c * 8, o, sp);
}

// Bilinear align_corners=True upsample (hardnet.py:249-254) of a list of channel slices into one
// contiguous NHWC buffer.  Index/weight arithmetic follows ATen's area_pixel_compute_source_index.
struct UpParams {
  SegView segs[kMaxSegs];
  size_t in_img[kMaxSegs];
  int nseg;
  void* out; void* out_lo; int out_cs; size_t out_img;
  int b, Hi, Wi, Ho, Wo, c4_total, split;
  float sh, sw;
};

// grid = (ceil(ceil(Wo / 8) * groups / 256), Ho, b): one thread per (run of 8 output pixels of a row, 8-channel group).
// Vertical interpolation first, per SOURCE column, and the two columns a pixel needs are carried along the run: at x2
// upsampling a new source column is needed every other output pixel, so a pixel costs about 2 instead of 8 128-bit
// loads and a quarter of the unpack / index arithmetic (the one-pixel-per-thread version executed 373 instructions per
// thread and was issue-bound at 2.5 TB/s: ncu, 245 M warp instructions for the 1/8 -> 1/4 resolution step).
// Weights follow ATen's area_pixel_compute_source_index (align_corners=True); the association
// hx * (hy v00 + ly v10) + lx * (hy v01 + ly v11) differs from ATen's by rounding only.
constexpr int kUpRun = 8;
__global__ void __launch_bounds__(256) upsample_bilinear_kernel(UpParams p) {
  const bool sp = p.split != 0;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int runs = (p.Wo + kUpRun - 1) / kUpRun;
  if (idx >= runs * p.c4_total) return;
  const int run = idx / p.c4_total;
  int c = (idx - run * p.c4_total) * 8;
  const int y = blockIdx.y, img = blockIdx.z;
  const int cout = c;
  int s = 0;
  while (c >= p.segs[s].cpad) { c -= p.segs[s].cpad; ++s; }
  const float fy = p.sh * (float)y;
  const int y0 = (int)fy;
  const int y1 = y0 + (y0 < p.Hi - 1 ? 1 : 0);
  const float ly = fy - (float)y0, hy = 1.f - ly;
  const int cs = p.segs[s].cstride;
  const void* bh = p.segs[s].base;
  const void* bl = p.segs[s].base_lo;
  const size_t row0 = (size_t)img * p.in_img[s] + c + (size_t)y0 * p.Wi * cs;
  const size_t row1 = (size_t)img * p.in_img[s] + c + (size_t)y1 * p.Wi * cs;
  auto column = [&](int cx, float* t) {                     // vertically interpolated source column cx
    float a[8], b_[8];
    load8_any(bh, bl, row0 + (size_t)cx * cs, sp, a);
    load8_any(bh, bl, row1 + (size_t)cx * cs, sp, b_);
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = hy * a[k] + ly * b_[k];
  };
  float tl[8], tr[8], o[8];
  int cl = -1, cr = -1;
  const int xs = run * kUpRun;
  size_t out_off = (size_t)img * p.out_img + ((size_t)y * p.Wo + xs) * p.out_cs + cout;
#pragma unroll
  for (int i = 0; i < kUpRun; ++i, out_off += p.out_cs) {
    const int x = xs + i;
    if (x >= p.Wo) break;
    const float fx = p.sw * (float)x;
    const int x0 = (int)fx;
    const int x1 = x0 + (x0 < p.Wi - 1 ? 1 : 0);
    const float lx = fx - (float)x0, hx = 1.f - lx;
    if (x0 != cl) {
      if (x0 == cr) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tl[k] = tr[k];
      } else {
        column(x0, tl);
      }
      cl = x0;
    }
    if (x1 != cr) {
      if (x1 == cl) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tr[k] = tl[k];
      } else {
        column(x1, tr);
      }
      cr = x1;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = hx * tl[k] + lx * tr[k];
    store8_any(p.out, p.out_lo, out_off, o, sp);
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
    if (NHWC16) {
      // 16 fp32 logits per quarter-res pixel: four 128-bit loads per corner, classes in registers
      const float4* p00 = reinterpret_cast<const float4*>(q + ((size_t)img * h * w + (size_t)y0 * w + x0) * 16);
      const float4* p01 = reinterpret_cast<const float4*>(q + ((size_t)img * h * w + (size_t)y0 * w + x1) * 16);
      const float4* p10 = reinterpret_cast<const float4*>(q + ((size_t)img * h * w + (size_t)y1 * w + x0) * 16);
      const float4* p11 = reinterpret_cast<const float4*>(q + ((size_t)img * h * w + (size_t)y1 * w + x1) * 16);
      const int nq = (ncls + 3) >> 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= nq) break;
        const float4 a = __ldg(p00 + k), b_ = __ldg(p01 + k), c_ = __ldg(p10 + k), d = __ldg(p11 + k);
        float v[4];
        v[0] = hy * (hx * a.x + lx * b_.x) + ly * (hx * c_.x + lx * d.x);
        v[1] = hy * (hx * a.y + lx * b_.y) + ly * (hx * c_.y + lx * d.y);
        v[2] = hy * (hx * a.z + lx * b_.z) + ly * (hx * c_.z + lx * d.z);
        v[3] = hy * (hx * a.w + lx * b_.w) + ly * (hx * c_.w + lx * d.w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = k * 4 + j;
          if (c < ncls) {
            if (full) full[(((size_t)img * ncls + c) * fh + y) * fw + x] = v[j];
            if (v[j] > best) { best = v[j]; arg = c; }
          }
        }
      }
    } else {
      for (int c = 0; c < ncls; ++c) {
        const float* base = q + ((size_t)img * ncls + c) * h * w;
        const float v00 = base[(size_t)y0 * w + x0], v01 = base[(size_t)y0 * w + x1];
        const float v10 = base[(size_t)y1 * w + x0], v11 = base[(size_t)y1 * w + x1];
        const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
        if (full) full[(((size_t)img * ncls + c) * fh + y) * fw + x] = v;
        if (v > best) { best = v; arg = c; }
      }
    }
    if (seg8) seg8[i] = (uint8_t)arg;
    if (seg64) seg64[i] = arg;
  }
}

// Strip variant for the usual x4 case (3*sw < 1, fw % 4 == 0): one thread = 4 consecutive output
// pixels of a row; they touch at most 3 source columns x 2 rows, loaded once (18 instead of 48 LDG.128).
// Vertical interpolation first (once per source column): same weights as upsample_argmax_kernel, sums associated differently.
// AMAX: channel 15 of every source pixel holds its own argmax (written by the head conv's epilogue).  If the six
// source pixels a strip touches agree on class c, then c maximises every convex combination of them (first-maximum
// ties included), so the strip is c without any interpolation; label maps are piecewise constant, so that is the
// common case and the kernel becomes a byte-map expansion.  Strips on a class boundary take the full path.
template <bool FULL, bool AMAX>
__global__ void __launch_bounds__(256) upsample_argmax_strip_kernel(const float* __restrict__ q, int ncls, int h, int w, int fh, int fw,
                                                                    float sh, float sw, uint8_t* __restrict__ seg8,
                                                                    long long* __restrict__ seg64, float* __restrict__ full) {
  // grid = (ceil(fw / 4 / 256), fh, b): one thread per strip of 4 output pixels, 32-bit index math
  const int strip = blockIdx.x * blockDim.x + threadIdx.x;
  if (strip >= (fw >> 2)) return;
  const int xs = strip * 4, y = blockIdx.y, img = blockIdx.z;
  const float fy = sh * (float)y;
  const int y0 = (int)fy;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
  const float ly = fy - (float)y0, hy = 1.f - ly;
  const int c0 = (int)(sw * (float)xs);
  const int cA = c0, cB = min(c0 + 1, w - 1), cC = min(c0 + 2, w - 1);
  const float* base = q + (size_t)img * h * w * 16;
  const float4* p0[3] = {reinterpret_cast<const float4*>(base + (size_t)(y0 * w + cA) * 16),
                         reinterpret_cast<const float4*>(base + (size_t)(y0 * w + cB) * 16),
                         reinterpret_cast<const float4*>(base + (size_t)(y0 * w + cC) * 16)};
  const float4* p1[3] = {reinterpret_cast<const float4*>(base + (size_t)(y1 * w + cA) * 16),
                         reinterpret_cast<const float4*>(base + (size_t)(y1 * w + cB) * 16),
                         reinterpret_cast<const float4*>(base + (size_t)(y1 * w + cC) * 16)};
  if (AMAX && !FULL) {
    const int c00 = __float_as_int(__ldg(reinterpret_cast<const float*>(p0[0]) + 15));
    const int c01 = __float_as_int(__ldg(reinterpret_cast<const float*>(p0[1]) + 15));
    const int c02 = __float_as_int(__ldg(reinterpret_cast<const float*>(p0[2]) + 15));
    const int c10 = __float_as_int(__ldg(reinterpret_cast<const float*>(p1[0]) + 15));
    const int c11 = __float_as_int(__ldg(reinterpret_cast<const float*>(p1[1]) + 15));
    const int c12 = __float_as_int(__ldg(reinterpret_cast<const float*>(p1[2]) + 15));
    if (c00 == c01 && c00 == c02 && c00 == c10 && c00 == c11 && c00 == c12) {
      const size_t o = ((size_t)img * fh + y) * fw + xs;
      if (seg8) *reinterpret_cast<uchar4*>(seg8 + o) = make_uchar4((uint8_t)c00, (uint8_t)c00, (uint8_t)c00, (uint8_t)c00);
      if (seg64) {
        longlong2* o64 = reinterpret_cast<longlong2*>(seg64 + o);
        o64[0] = make_longlong2(c00, c00);
        o64[1] = make_longlong2(c00, c00);
      }
      return;
    }
  }
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int arg[4] = {0, 0, 0, 0};
  float wc[4][3];                                             // weights of the three source columns for each of the 4 pixels
  float lx[4], hx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float fx = sw * (float)(xs + j);
    const int x0 = (int)fx;
    lx[j] = fx - (float)x0; hx[j] = 1.f - lx[j];
    const int ia = x0 - c0;                                   // 0 or 1: left column of pixel j within {cA, cB, cC}
    const int ib = ia + (x0 < w - 1 ? 1 : 0);                 // 0..2: right column
    // one of the three weights is zero: a 3-term weighted sum (FMUL + 2 FFMA per class) instead of two column selects
    // per class and pixel (a fifth of the kernel's instructions); 0 * column adds exactly nothing
    wc[j][0] = (ia == 0 ? hx[j] : 0.f) + (ib == 0 ? lx[j] : 0.f);
    wc[j][1] = (ia == 1 ? hx[j] : 0.f) + (ib == 1 ? lx[j] : 0.f);
    wc[j][2] = ib == 2 ? lx[j] : 0.f;
  }
  const int nq = (ncls + 3) >> 2;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= nq) break;
    float4 t0[3], t1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { t0[c] = __ldg(p0[c] + k); t1[c] = __ldg(p1[c] + k); }
    float4 col[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      col[c] = make_float4(hy * t0[c].x + ly * t1[c].x, hy * t0[c].y + ly * t1[c].y, hy * t0[c].z + ly * t1[c].z,
                           hy * t0[c].w + ly * t1[c].w);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // vertical interpolation once per source column (shared by the strip's pixels), then the horizontal blend
      float v[4];
      v[0] = fmaf(wc[j][2], col[2].x, fmaf(wc[j][1], col[1].x, wc[j][0] * col[0].x));
      v[1] = fmaf(wc[j][2], col[2].y, fmaf(wc[j][1], col[1].y, wc[j][0] * col[0].y));
      v[2] = fmaf(wc[j][2], col[2].z, fmaf(wc[j][1], col[1].z, wc[j][0] * col[0].z));
      v[3] = fmaf(wc[j][2], col[2].w, fmaf(wc[j][1], col[1].w, wc[j][0] * col[0].w));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = k * 4 + e;
        if (c < ncls) {
          if (FULL) full[(((size_t)img * ncls + c) * fh + y) * fw + xs + j] = v[e];
          if (v[e] > best[j]) { best[j] = v[e]; arg[j] = c; }
        }
      }
    }
  }
  const size_t o = ((size_t)img * fh + y) * fw + xs;
  if (seg8) *reinterpret_cast<uchar4*>(seg8 + o) = make_uchar4((uint8_t)arg[0], (uint8_t)arg[1], (uint8_t)arg[2], (uint8_t)arg[3]);
  if (seg64) {
    longlong2* o64 = reinterpret_cast<longlong2*>(seg64 + o);
    o64[0] = make_longlong2(arg[0], arg[1]);
    o64[1] = make_longlong2(arg[2], arg[3]);
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

// debug helpers: NCHW fp32 <-> NHWC (fp32 or split-bf16) with channel stride cs
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, void* out, void* out_lo, int split, int b, int c,
                                    int h, int w, int cs) {
  const size_t total = (size_t)b * h * w * cs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cs);
    size_t r = i / cs;
    const int x = (int)(r % w); r /= w;
    const int y = (int)(r % h);
    const int img = (int)(r / h);
    const float v = ch < c ? in[(((size_t)img * c + ch) * h + y) * w + x] : 0.f;
    if (!split) {
      reinterpret_cast<float*>(out)[i] = v;
    } else {
      unsigned hi, lo;
      split1(v, &hi, &lo);
      reinterpret_cast<unsigned short*>(out)[i] = (unsigned short)hi;
      reinterpret_cast<unsigned short*>(out_lo)[i] = (unsigned short)lo;
    }
  }
}
__global__ void nhwc_to_nchw_kernel(const void* in, const void* in_lo, int split, float* __restrict__ out, int b,
                                    int c, int h, int w, int cs) {
  const size_t total = (size_t)b * c * h * w;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    size_t r = i / w;
    const int y = (int)(r % h); r /= h;
    const int ch = (int)(r % c);
    const int img = (int)(r / c);
    const size_t o = (((size_t)img * h + y) * w + x) * cs + ch;
    if (!split) {
      out[i] = reinterpret_cast<const float*>(in)[o];
    } else {
      const unsigned hi = reinterpret_cast<const unsigned short*>(in)[o];
      const unsigned lo = reinterpret_cast<const unsigned short*>(in_lo)[o];
      out[i] = __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
    }
  }
}

// debug helpers for the space-to-depth layers (split-bf16 storage, 4 phase blocks of 32 channels)
__global__ void nchw_to_s2d_kernel(const float* __restrict__ in, unsigned short* out, unsigned short* out_lo, int b, int c,
                                   int h, int w) {
  const size_t total = (size_t)b * (h / 2) * (w / 2) * 128;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 128);
    size_t r = i / 128;
    const int X = (int)(r % (w / 2)); r /= (w / 2);
    const int Y = (int)(r % (h / 2));
    const int img = (int)(r / (h / 2));
    const int blk = ch / 32, cc = ch % 32;
    const int y = 2 * Y + (blk >> 1), x = 2 * X + (blk & 1);
    const float v = cc < c ? in[(((size_t)img * c + cc) * h + y) * w + x] : 0.f;
    unsigned hi, lo;
    split1(v, &hi, &lo);
    out[i] = (unsigned short)hi;
    out_lo[i] = (unsigned short)lo;
  }
}
__global__ void s2d_to_nchw_kernel(const unsigned short* in, const unsigned short* in_lo, float* __restrict__ out, int b,
                                   int c, int h, int w) {
  const size_t total = (size_t)b * c * h * w;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    size_t r = i / w;
    const int y = (int)(r % h); r /= h;
    const int ch = (int)(r % c);
    const int img = (int)(r / c);
    const size_t o = (((size_t)img * (h / 2) + (y >> 1)) * (w / 2) + (x >> 1)) * 128 + ((y & 1) * 2 + (x & 1)) * 32 + ch;
    out[i] = __uint_as_float((unsigned)in[o] << 16) + __uint_as_float((unsigned)in_lo[o] << 16);
  }
}

static int grid_for(size_t total, int threads) {
  size_t g = (total + threads - 1) / threads;
  const size_t cap = (size_t)kNumSMs * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------
// Arena: all activation buffers of one forward inside the caller's work space.
// fp32 storage: one plane per buffer; split storage: a bf16 hi plane followed by a bf16 lo plane.
// The quarter-resolution logits buffer is always fp32.
struct Arena {
  char* base = nullptr;
  int b = 0, H = 0, W = 0;
  bool split = false;
  std::vector<size_t> off;        // byte offset of the (hi) plane of each buffer
  std::vector<size_t> off_lo;     // byte offset of the lo plane (split storage)
  std::vector<size_t> img_elems;  // elements between consecutive images of a buffer
  size_t total_bytes = 0;

  void* ptr(int buf, int coff) const { return base + off[buf] + (size_t)coff * (split_buf(buf) ? 2 : 4); }
  void* ptr_lo(int buf, int coff) const { return split_buf(buf) ? base + off_lo[buf] + (size_t)coff * 2 : nullptr; }
  bool split_buf(int buf) const { return split && !f32[buf]; }
  std::vector<char> f32;          // buffers that stay fp32 under split storage
};

static void make_arena(const pf_bgnet* net, void* ws, int b, int H, int W, Arena* a) {
  a->base = ws ? reinterpret_cast<char*>(align_up((size_t)ws, 256)) : nullptr;
  a->b = b; a->H = H; a->W = W;
  a->split = net->precision == 1;
  a->f32.resize(net->bufs.size());
  for (size_t i = 0; i < net->bufs.size(); ++i) a->f32[i] = net->bufs[i].always_f32 ? 1 : 0;
  const size_t nb = net->bufs.size();
  a->off.resize(nb); a->off_lo.resize(nb); a->img_elems.resize(nb);
  size_t off = 0;
  for (size_t i = 0; i < nb; ++i) {
    const BufDesc& bd = net->bufs[i];
    const size_t n = align_up((size_t)(H >> bd.shift) * (W >> bd.shift) * bd.cstride, 64);
    a->img_elems[i] = n;
    const bool sp = a->split_buf((int)i);
    a->off[i] = off;
    off += align_up(n * b * (sp ? 2 : 4), 256);
    if (sp) {
      a->off_lo[i] = off;
      off += align_up(n * b * 2, 256);
    }
  }
  a->total_bytes = off + 256;
}

static void fill_conv_launch(const pf_bgnet* net, const Arena& a, const ConvDesc& c, ConvLaunch* L) {
  L->nseg = (int)c.in.size();
  for (int s = 0; s < L->nseg; ++s) {
    const SegRef& r = c.in[s];
    L->segs[s].base = a.ptr(r.buf, r.coff);
    L->segs[s].base_lo = a.ptr_lo(r.buf, r.coff);
    L->segs[s].cstride = net->bufs[r.buf].cstride;
    L->segs[s].cpad = r.cpad();
    L->in_img_stride[s] = a.img_elems[r.buf];
  }
  const BufDesc& ib = net->bufs[c.in[0].buf];
  const BufDesc& ob = net->bufs[c.out.buf];
  L->b = a.b;
  L->Hin = a.H >> ib.shift; L->Win = a.W >> ib.shift;
  L->Hout = a.H >> ob.shift; L->Wout = a.W >> ob.shift;
  L->out = a.ptr(c.out.buf, c.out.coff);
  L->out_lo = a.ptr_lo(c.out.buf, c.out.coff);
  L->out_cstride = ob.cstride;
  L->out_img_stride = a.img_elems[c.out.buf];
  L->w = c.w_dev; L->bias = c.bias_dev;
  L->kpad = c.kpad; L->coutpad = c.coutpad;
  L->cout_store = padc(c.cout);
  L->relu = c.relu ? 1 : 0;
  L->s2d_block = 0;
  if (c.s2d_out) {                     // output buffer is the quarter-res space-to-depth tensor
    L->Hout = L->Hin; L->Wout = L->Win;
    L->s2d_block = 32;
    L->cout_store = 32;
  }
}

// ---- tensor-core path helpers --------------------------------------------------------------
static unsigned short f32_to_bf16_rn(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (unsigned short)(u >> 16);   // inf / nan
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}
static float bf16_to_f32(unsigned short h) {
  unsigned u = (unsigned)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// Packs the folded fp32 weights of conv i as the K-major B operand: W[n][k], k = seg_base +
// tap * cpad(seg) + channel, split into bf16 hi / lo planes.
static int upload_conv_tc(pf_bgnet* net, int i) {
  ConvDesc& c = net->convs[i];
  if (net->wtc_dev.size() < net->convs.size()) {
    net->wtc_dev.resize(net->convs.size(), nullptr);
    net->wtc_rows.resize(net->convs.size(), 0);
  }
  const int taps = c.ksize * c.ksize;
  const int ktot = taps * c.kpad;
  const int nrows = c.coutpad;      // rows beyond it are TMA out-of-bounds zero fill
  std::vector<unsigned short> w((size_t)2 * nrows * ktot, 0);
  int kp = 0, kb = 0;
  for (auto& s : c.in) {
    const int cp = s.cpad();
    for (int tap = 0; tap < taps; ++tap)
      for (int ch = 0; ch < s.c; ++ch)
        for (int o = 0; o < c.cout; ++o) {
          const float v = c.w_host[((size_t)tap * c.kpad + kp + ch) * c.coutpad + o];
          const unsigned short hi = f32_to_bf16_rn(v);
          const unsigned short lo = f32_to_bf16_rn(v - bf16_to_f32(hi));
          const size_t k = (size_t)kb + (size_t)tap * cp + ch;
          w[(size_t)o * ktot + k] = hi;
          w[(size_t)nrows * ktot + (size_t)o * ktot + k] = lo;
        }
    kp += cp;
    kb += taps * cp;
  }
  if (net->wtc_dev[i]) cudaFree(net->wtc_dev[i]);
  PF_CHECK_CUDA(cudaMalloc(&net->wtc_dev[i], w.size() * 2));
  PF_CHECK_CUDA(cudaMemcpy(net->wtc_dev[i], w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  net->wtc_rows[i] = nrows;
  return 0;
}

struct TcIo {                       // where one conv reads and writes (arena or debug buffers)
  const void* in_hi[kMaxSegs]; const void* in_lo[kMaxSegs];
  int in_cs[kMaxSegs]; size_t in_img[kMaxSegs];
  int Hin, Win, Hout, Wout, b;
  void* out_hi; void* out_lo; float* out_f32; int out_cs; size_t out_img;
};

// Appends the tensor maps of conv i to `maps` and fills its TcLayer.
static int build_tc_layer(pf_bgnet* net, int i, const TcIo& io, std::vector<CUtensorMap>* maps, TcLayer* L,
                          int* nblocks, size_t* smem) {
  const ConvDesc& c = net->convs[i];
  PF_REQUIRE(c.stride == 1 && !c.s2d_out, PF_EINVAL, "build_tc_layer: layer needs the halo or SIMT kernel");
  memset(L, 0, sizeof(*L));
  const int taps = c.ksize * c.ksize;
  L->nseg = (int)c.in.size();
  int kb = 0;
  for (int s = 0; s < L->nseg; ++s) {
    const int cp = c.in[s].cpad();
    L->seg_cpad[s] = cp;
    L->seg_koff[s] = kb;
    kb += taps * cp;
    L->seg_map[s] = (int)maps->size();
    CUtensorMap m;
    int rc = tc_encode_act_map(&m, io.in_hi[s], cp, io.in_cs[s], io.Win, io.Hin, io.b, io.in_img[s]);
    if (rc) return rc;
    maps->push_back(m);
    rc = tc_encode_act_map(&m, io.in_lo[s], cp, io.in_cs[s], io.Win, io.Hin, io.b, io.in_img[s]);
    if (rc) return rc;
    maps->push_back(m);
  }
  const int ktot = taps * c.kpad;
  int ntile, nb, stages, cols;
  tc_pick_tiling(c.coutpad, cdiv(io.Wout, 16) * cdiv(io.Hout, 8) * io.b, &ntile, &nb, &stages, &cols, smem);
  *nblocks = nb;
  const int nrows = net->wtc_rows[i];
  L->w_map = (int)maps->size();
  {
    CUtensorMap m;
    int rc = tc_encode_weight_map(&m, net->wtc_dev[i], ktot, nrows, ntile);
    if (rc) return rc;
    maps->push_back(m);
    rc = tc_encode_weight_map(&m, net->wtc_dev[i] + (size_t)nrows * ktot, ktot, nrows, ntile);
    if (rc) return rc;
    maps->push_back(m);
  }
  L->taps = taps; L->ksize = c.ksize;
  L->Hout = io.Hout; L->Wout = io.Wout;
  L->tiles_x = cdiv(io.Wout, 16); L->tiles_y = cdiv(io.Hout, 8);
  L->ntile = ntile; L->stages = stages; L->tmem_cols = cols;
  L->cout_store = (io.out_f32 && i == net->final_conv) ? 16 : (c.s2d_out ? 32 : padc(c.cout));
  L->relu = c.relu ? 1 : 0;
  L->out_hi = reinterpret_cast<__nv_bfloat16*>(io.out_hi);
  L->out_lo = reinterpret_cast<__nv_bfloat16*>(io.out_lo);
  L->out_f32 = io.out_f32;
  L->out_cs = io.out_cs; L->out_img_stride = io.out_img;
  L->bias = c.bias_dev;
  return 0;
}

// Shared-memory plan of the dx-folded form; false when the weights cannot be resident.  A plan left with a 2-deep
// ring of 40 KB stages stages 32-channel chunks instead (twice the boxes, half the stage size) when that gives >= 3.
static bool plan_fold(HaloLayer* T, size_t* smem) {
  T->fold = 1; T->hx = 16; T->hy = 10;
  if (!halo_plan_smem(T, smem)) return false;
  if (T->stages_a < 3) {
    HaloLayer V = *T;
    size_t sm2 = 0;
    bool any = false;
    for (int k = 0; k < V.nseg; ++k)
      if (V.seg_w[k] == 64) { V.seg_w[k] = 32; any = true; }
    if (any && halo_plan_smem(&V, &sm2) && V.stages_a >= 3) { *T = V; *smem = sm2; }
  }
  return true;
}

// Halo-kernel plan of a 3x3 / stride-1 conv.  Returns 1 when the layer does not fit (caller falls
// back to the per-tap kernel), 0 on success, <0 / cudaError on failure.
static int build_halo_layer(pf_bgnet* net, int i, const TcIo& io, std::vector<CUtensorMap>* maps, HaloLayer* L,
                            int* nblocks, size_t* smem, int seg0 = 0, int seg1 = -1, bool add_patch = false) {
  const ConvDesc& c = net->convs[i];
  if (seg1 < 0) seg1 = (int)c.in.size();
  if (c.exec_stride() != 1) return 1;
  memset(L, 0, sizeof(*L));
  const int taps = c.ksize * c.ksize;
  L->taps = taps; L->hx = c.ksize == 3 ? 10 : 8; L->hy = c.ksize == 3 ? 18 : 16;
  L->tap_mask = c.s2d_in ? 0x1B : (1 << taps) - 1;
  L->s2d_block = c.s2d_out ? 32 : 0;
  int ntile, nb, stages, cols;
  size_t dummy;
  tc_pick_tiling(c.coutpad, cdiv(io.Wout, 8) * cdiv(io.Hout, 16) * io.b, &ntile, &nb, &stages, &cols, &dummy);
  L->nseg = seg1 - seg0;
  L->ntile = ntile;
  if (add_patch) L->add_pbytes = (uint32_t)align_up((size_t)kAddPH * kAddPW * (ntile + 4) * 4, 1024);
  int kb = 0;
  bool used[3] = {false, false, false};
  for (int s = 0; s < (int)c.in.size(); ++s) {       // K offsets run over ALL slices; only [seg0, seg1) are read
    const int cp = c.in[s].cpad();
    if (s >= seg0 && s < seg1) {
      L->seg_cpad[s - seg0] = cp;
      L->seg_w[s - seg0] = halo_chunk_width(cp);
      used[L->seg_w[s - seg0] >> 5] = true;
      L->seg_koff[s - seg0] = kb;
    }
    kb += taps * cp;
  }
  // dx-folded form (conv_halo.cu, issue_chunk_fold): full 3x3 layers with at most 32 output channels per CTA
  const char* nf = getenv("PF_HALO_NO_FOLD");
  const bool no_fold = nf && nf[0] == '1';
  // Measured on B200 (batch 8, 1/4 resolution): 58->18 167 -> 115 us, 76->28 230 -> 145 us, 73->18 181 -> 118 us, but
  // 18->10 58 -> 71 us and 16->24 256 -> 378 us: with a single 16/32-channel chunk a tile is a dozen MMAs and the
  // epilogue / stores bound it, where the folded form's 8 x 14 tiles (12.5% idle accumulator rows) only cost.
  int kin = 0;
  for (int s = seg0; s < seg1; ++s) kin += c.in[s].cpad();
  const char* ff = getenv("PF_HALO_FOLD_MIN_K");
  // (round 2: with two epilogue teams on alternate tiles, HaloLayer::alt, the single-chunk layers fold too -- 18->10 at
  // 1/4 resolution 117 -> 100 us, 28->16 at 1/8 40 -> 34 us per 16 frames; folded with ONE team they lose: 117 -> 123)
  const int fold_min_k = ff && ff[0] ? atoi(ff) : 32;
  L->fold = (!no_fold && c.ksize == 3 && L->tap_mask == 0x1FF && ntile <= 32 && kin >= fold_min_k) ? 1 : 0;
  if (L->fold) {
    // The folded form needs resident weights and a ring of >= 3 activation stages.  A 32-cout layer whose weights
    // leave no room for that (K >~ 100) is split along N instead: two CTAs per tile with 16 couts each (half the
    // weights per SM; the activation boxes are fetched twice, from L2).
    const char* fs = getenv("PF_HALO_FOLD_SPLIT");
    const int split_mode = fs && fs[0] ? atoi(fs) : 1;       // 0 never, 1 when n = 32 does not fit, 2 always
    bool planned = false;
    {
      // N tile of 24 for the 18- / 24-output-channel layers (slots are padded to 32): the folded MMAs get N = 144 / 80
      // instead of 192 / 96 (72 + 52 instead of 96 + 56 cycles per K atom and filter row).  A/B PF_HALO_FOLD_N24=0.
      const char* n24 = getenv("PF_HALO_FOLD_N24");
      if (!(n24 && n24[0] == '0') && ntile == 32 && c.coutpad == 32 && c.cout <= 24 && !c.s2d_out) {
        HaloLayer T = *L;
        size_t sm2 = 0;
        T.ntile = 24;
        if (plan_fold(&T, &sm2) && T.stages_a >= 3) { *L = T; *smem = sm2; ntile = 24; planned = true; }
      }
    }
    if (!planned) planned = plan_fold(L, smem);              // false when the weights cannot be resident
    const bool roomy = planned && L->stages_a >= 3;
    // measured (batch 8): 135->28 at 1/8 (weights not resident at n = 32) 115 -> 77 us, 96->18 / 114->30 at 1/16
    // (resident with 2 stages, 4 tiles per SM) 31 -> 22 / 37 -> 27 us, but 91->28 at 1/4 (resident with 2 stages,
    // 63 tiles per SM) 196 -> 248 us: fetching every activation box twice costs more than the shallow ring there
    // (32-channel chunks, plan_fold, give it 4 stages instead: 196 -> 153 us).
    const bool few_tiles = cdiv(io.Wout, 14) * cdiv(io.Hout, 8) * io.b < 16 * kNumSMs;
    if (ntile == 32 && c.coutpad == 32 && split_mode && (!planned || (!roomy && few_tiles) || split_mode == 2)) {
      HaloLayer T = *L;
      size_t sm2 = 0;
      T.ntile = 16;
      if (plan_fold(&T, &sm2) && T.stages_a >= 3) { *L = T; *smem = sm2; ntile = 16; nb = 2; planned = true; }
    }
    if (!planned) {                                          // (resident with 2 stages still beats the unfolded form)
      L->fold = 0; L->hx = 10; L->hy = 18;
      for (int k = 0; k < L->nseg; ++k) L->seg_w[k] = halo_chunk_width(L->seg_cpad[k]);
    }
  }
  if (!L->fold && !halo_plan_smem(L, smem)) {
    if (!L->add_pbytes) return 1;
    L->add_pbytes = 0;                                       // no room for the staged patches: per-pixel gathers
    if (!halo_plan_smem(L, smem)) return 1;
  }
  if (!L->fold && !L->resident) {
    // streamed weights: with 64-channel chunks the activation ring is 2 deep (46 KB per stage) -- the load of chunk
    // j + 2 only starts when chunk j's MMAs have drained -- and an N tile >= 96 leaves 3-5 weight stages of 24-32 KB.
    // 32-channel chunks halve both: 4 activation stages + 8 weight stages.  Measured per 16 frames, all streamed layers
    // on 32-channel chunks: 196->88 119 -> 110 us, 294->118 71 -> 63, 401->118 89 -> 81, 108->46 (then folded in three
    // 16-cout CTAs) 189 -> 173, but 235->52 101 -> 113, 267->24 33 -> 38: the N = 64 layers keep 6-7 weight stages with
    // wide chunks and only pay the extra boxes.  Default: N tile >= 96 or 48.  A/B PF_HALO_STREAM_W=64 / 32: all wide / narrow.
    const char* sw = getenv("PF_HALO_STREAM_W");
    const int wmax = sw && sw[0] ? atoi(sw) : ((ntile >= 96 || ntile == 48) ? 32 : 64);
    if (wmax < 64) {
      HaloLayer V = *L;
      size_t sm2 = 0;
      bool any = false;
      for (int k = 0; k < V.nseg; ++k)
        if (V.seg_w[k] > wmax) { V.seg_w[k] = wmax; any = true; }
      if (any && halo_plan_smem(&V, &sm2) && !V.resident && V.stages_a > L->stages_a) { *L = V; *smem = sm2; }
    }
  }
  if (!L->fold && !no_fold && c.ksize == 3 && L->tap_mask == 0x1FF && ntile == 48 && c.coutpad == 48 && kin >= fold_min_k &&
      !L->resident) {
    // 48 couts whose weights have to be streamed per tile: three folded CTAs of 16 couts with resident weights instead
    // (108->46 at 1/8: 114 -> 102 us)
    // (second half of round 2) two CTAs of 24 couts: MMA N = 144 / 80, the activation boxes fetched twice instead of
    // three times.  PF_HALO_FOLD_48: 0 = unfolded, 3 = the three-way split.
    const char* f48 = getenv("PF_HALO_FOLD_48");
    if (!(f48 && f48[0] == '0')) {
      HaloLayer T = *L;
      size_t sm2 = 0;
      bool done = false;
      if (!(f48 && f48[0] == '3')) {
        T.ntile = 24;
        if (plan_fold(&T, &sm2) && T.stages_a >= 3) { *L = T; *smem = sm2; ntile = 24; nb = 2; done = true; }
      }
      if (!done) {
        T = *L; sm2 = 0;
        T.ntile = 16;
        if (plan_fold(&T, &sm2) && T.stages_a >= 3) { *L = T; *smem = sm2; ntile = 16; nb = 3; }
      }
    }
  }
  used[0] = used[1] = used[2] = false;
  for (int k = 0; k < L->nseg; ++k) used[L->seg_w[k] >> 5] = true;
  for (int s = seg0; s < seg1; ++s) {
    const int k = s - seg0;
    L->seg_map[k] = (int)maps->size();
    CUtensorMap m;
    int rc = halo_encode_act_map(&m, io.in_hi[s], L->seg_cpad[k], io.in_cs[s], io.Win, io.Hin, io.b, io.in_img[s], L->seg_w[k], L->hx, L->hy);
    if (rc) return rc;
    maps->push_back(m);
    rc = halo_encode_act_map(&m, io.in_lo[s], L->seg_cpad[k], io.in_cs[s], io.Win, io.Hin, io.b, io.in_img[s], L->seg_w[k], L->hx, L->hy);
    if (rc) return rc;
    maps->push_back(m);
  }
  const int ktot = taps * c.kpad;
  const int nrows = net->wtc_rows[i];
  for (int k = 0; k < 3; ++k) {
    L->w_map[k] = -1;
    if (!used[k]) continue;
    const int w = 16 << k;
    L->w_map[k] = (int)maps->size();
    CUtensorMap m;
    int rc = halo_encode_weight_map(&m, net->wtc_dev[i], ktot, nrows, ntile, w);
    if (rc) return rc;
    maps->push_back(m);
    rc = halo_encode_weight_map(&m, net->wtc_dev[i] + (size_t)nrows * ktot, ktot, nrows, ntile, w);
    if (rc) return rc;
    maps->push_back(m);
  }
  {
    // CTA pairs for the layers whose weights are streamed per (chunk, tap): at 1/16 resolution and below every tile
    // pulls the layer's whole weight matrix (0.3 - 2.3 MB) from L2, ~40 B per clock and SM when the MMAs are to stay
    // busy, and all 148 SMs do so at once.  The hypothesis was that the L2 -> SM fabric bounds them.  A pair shares one
    // weight stream: each CTA loads half of every tile and TMA multicast delivers it to both.  MEASURED (batch 16): 3-8 %
    // SLOWER on every streamed layer (196->88 119 -> 124 us, 144->52 81 -> 88, 235->52 102 -> 108): the L2 -> SM fabric
    // is not what bounds them, and the pair runs in lock step.  Opt-in (PF_HALO_CLUSTER=1), tested.
    const char* cl = getenv("PF_HALO_CLUSTER");
    const bool want = cl && cl[0] == '1';
    const int tiles = cdiv(io.Wout, 8) * cdiv(io.Hout, 16) * io.b;
    L->cluster = (want && !L->resident && !L->fold && ntile % 16 == 0 && tiles >= 2 && kNumSMs / nb >= 2) ? 1 : 0;
    for (int k = 0; k < 3; ++k) {
      L->w_map_half[k] = -1;
      if (!L->cluster || !used[k]) continue;
      const int w = 16 << k;
      L->w_map_half[k] = (int)maps->size();
      CUtensorMap m;
      int rc = halo_encode_weight_map(&m, net->wtc_dev[i], ktot, nrows, ntile / 2, w);
      if (rc) return rc;
      maps->push_back(m);
      rc = halo_encode_weight_map(&m, net->wtc_dev[i] + (size_t)nrows * ktot, ktot, nrows, ntile / 2, w);
      if (rc) return rc;
      maps->push_back(m);
    }
  }
  *nblocks = nb;
  {
    const char* e8 = getenv("PF_HALO_EPI8");               // A/B: second epilogue team (0 = off, n = minimum N tile)
    const int e8min = e8 && e8[0] ? atoi(e8) : 32;
    L->epi8 = (!L->fold && e8min > 0 && L->ntile >= e8min && L->ntile >= 32) ? 2 : 0;       // epilogue teams (0 = one)
    // four teams (608 threads, one group per round) for N tiles >= 64: base.5 216 -> 185 us, base.8 98 -> 74 us per 16
    // frames; not for the fused conv1x1_up layers, whose interpolating epilogue got slower (365 -> 381 us).
    // A/B PF_HALO_EPI16=0: never.
    const char* ft = getenv("PF_HALO_FOLD_TEAMS");         // A/B: two epilogue teams for folded layers with two 16-channel groups
    if (L->fold && L->ntile == 32 && ft && ft[0] == '1') L->epi8 = 2;
    const char* e16 = getenv("PF_HALO_EPI16");
    const bool never = e16 && e16[0] == '0';               // (the four-team kernel carries no additive term: never for
                                                           // the fused conv1x1_up layers, whatever the switch says)
    if (L->epi8 && L->ntile >= 64 && !never && !L->add_pbytes && !add_patch) L->epi8 = 4;
    // alternate-tile epilogue teams (conv_halo_kernel<352, 3>: team k owns accumulator buffer k) for N tiles <= 32.
    // Measured per 16 frames: base.1 (16->24 at 1/2 resolution, one 16-channel chunk = 18 MMAs per tile, the epilogue
    // is the whole cost) 500 -> 449 us against two column teams; no gain anywhere else -- the one-team K-light layers
    // (18->10: 113 us either way) are bound by the MMAs' shared-memory A fetch, the folded ones by the tensor pipe --
    // and the N = 32 layers with more K (base.2, 30->18) lose 2-8 %.  Default: only the single-16-channel-chunk case.
    // A/B PF_HALO_ALT: bit 0 = unfolded layers with one team, bit 1 = folded layers, bit 2 = every N = 32 unfolded
    // layer that otherwise runs two column teams; 0 = never.
    const char* al = getenv("PF_HALO_ALT");
    const bool alt_auto = !(al && al[0]);
    const int alt = alt_auto ? 0 : atoi(al);
    if (L->ntile <= 32 && !L->add_pbytes && !L->add_src) {
      if (!L->fold && !L->epi8 && (alt & 1)) L->alt = 1;
      // folded: only the K-light layers (one 32-channel chunk = 12 MMAs per tile) are bound by their epilogue; the
      // others are tensor-bound and lose 2-6 us to the extra warps (76->28 274 -> 280, 73->18 210 -> 216)
      if (L->fold && !L->epi8 && ((alt & 2) || (alt_auto && kin <= 32))) L->alt = 1;
      const bool one_small_chunk = L->nchunk == 1 && L->seg_cpad[0] == 16;
      if (!L->fold && L->epi8 == 2 && ((alt & 4) || (alt_auto && one_small_chunk))) { L->alt = 1; L->epi8 = 0; }
    }
  }
  L->Hout = io.Hout; L->Wout = io.Wout; L->batch = io.b;
  L->tiles_x = cdiv(io.Wout, 8); L->tiles_y = cdiv(io.Hout, 16);
  if (L->fold) { L->tiles_x = cdiv(io.Wout, 14); L->tiles_y = cdiv(io.Hout, 8); }
  int tcols = 32;
  while (tcols < (L->fold ? 12 : 4) * ntile) tcols <<= 1;
  L->tmem_cols = tcols;
  L->cout_store = (io.out_f32 && i == net->final_conv) ? 16 : (c.s2d_out ? 32 : padc(c.cout));
  L->relu = c.relu ? 1 : 0;
  L->out_hi = reinterpret_cast<__nv_bfloat16*>(io.out_hi);
  L->out_lo = reinterpret_cast<__nv_bfloat16*>(io.out_lo);
  L->out_f32 = io.out_f32;
  L->out_cs = io.out_cs; L->out_img_stride = io.out_img;
  L->bias = c.bias_dev;
  return 0;
}

static int ensure_tc_plan(pf_bgnet* net, const Arena& a, const void* ws) {
  auto& P = net->plan;
  if (P.ws == ws && P.b == a.b && P.H == a.H && P.W == a.W && P.maps_dev) return 0;
  std::vector<CUtensorMap> maps;
  const size_t nc = net->convs.size();
  P.layers.assign(nc, TcLayer());
  P.halos.assign(nc, HaloLayer());
  P.halos_low.assign(nc, HaloLayer());
  P.nblocks_low.assign(nc, 0);
  P.smem_low.assign(nc, 0);
  P.nblocks.assign(nc, 0);
  if (!net->zero_bias_dev) {
    PF_CHECK_CUDA(cudaMalloc(&net->zero_bias_dev, 512 * sizeof(float)));
    PF_CHECK_CUDA(cudaMemset(net->zero_bias_dev, 0, 512 * sizeof(float)));
  }
  P.smem.assign(nc, 0);
  P.use_tc.assign(nc, 0);
  P.pool_fused.assign(net->steps.size(), 0);
  P.head_amax = 0;
  for (size_t i = 0; i < nc; ++i) {
    const ConvDesc& c = net->convs[i];
    if ((int)i == net->first_conv || c.exec_stride() != 1) continue;
    TcIo io;
    for (size_t s = 0; s < c.in.size(); ++s) {
      const SegRef& r = c.in[s];
      io.in_hi[s] = a.ptr(r.buf, r.coff); io.in_lo[s] = a.ptr_lo(r.buf, r.coff);
      io.in_cs[s] = net->bufs[r.buf].cstride; io.in_img[s] = a.img_elems[r.buf];
    }
    const BufDesc& ib = net->bufs[c.in[0].buf];
    const BufDesc& ob = net->bufs[c.out.buf];
    io.Hin = a.H >> ib.shift; io.Win = a.W >> ib.shift; io.Hout = a.H >> ob.shift; io.Wout = a.W >> ob.shift; io.b = a.b;
    if (c.s2d_out) { io.Hout = io.Hin; io.Wout = io.Win; }
    const bool head = (int)i == net->final_conv;
    io.out_hi = head ? nullptr : a.ptr(c.out.buf, c.out.coff);
    io.out_lo = head ? nullptr : a.ptr_lo(c.out.buf, c.out.coff);
    io.out_f32 = head ? reinterpret_cast<float*>(a.ptr(c.out.buf, c.out.coff)) : nullptr;
    io.out_cs = ob.cstride; io.out_img = a.img_elems[c.out.buf];
    if (c.up_nseg > 0) {
      // fused conv1x1_up: (1) low-resolution 1x1 over the first up_nseg slices -> fp32 ybuf (no bias / ReLU)
      const BufDesc& yb = net->bufs[c.ybuf];
      TcIo lo = io;
      lo.Hin = lo.Hout = a.H >> yb.shift; lo.Win = lo.Wout = a.W >> yb.shift;
      lo.out_hi = lo.out_lo = nullptr;
      lo.out_f32 = reinterpret_cast<float*>(a.ptr(c.ybuf, 0));
      lo.out_cs = yb.cstride; lo.out_img = a.img_elems[c.ybuf];
      int rc = build_halo_layer(net, (int)i, lo, &maps, &P.halos_low[i], &P.nblocks_low[i], &P.smem_low[i], 0, c.up_nseg);
      PF_REQUIRE(rc == 0, rc == 1 ? PF_EINVAL : rc, "fused conv1x1_up: low-resolution plan failed for %s", c.name.c_str());
      P.halos_low[i].relu = 0;
      P.halos_low[i].bias = net->zero_bias_dev;
      // (2) high-resolution 1x1 over the skip slices + bilinear(ybuf) + bias + ReLU
      TcIo hi = io;                                   // the skip slices live at the output resolution
      hi.Hin = io.Hout; hi.Win = io.Wout;
      const float sh = io.Hout > 1 ? (float)(lo.Hout - 1) / (float)(io.Hout - 1) : 0.f;
      const float sw = io.Wout > 1 ? (float)(lo.Wout - 1) / (float)(io.Wout - 1) : 0.f;
      const char* ng = getenv("PF_TC_FUSE_UP_GATHER");       // A/B: per-pixel global gathers instead of the staged patch
      const bool patch = sh <= 0.5f && sw <= 0.5f && !(ng && ng[0] == '1');   // kAddPH x kAddPW covers a tile at <= 1/2 scale
      rc = build_halo_layer(net, (int)i, hi, &maps, &P.halos[i], &P.nblocks[i], &P.smem[i], c.up_nseg, (int)c.in.size(), patch);
      PF_REQUIRE(rc == 0, rc == 1 ? PF_EINVAL : rc, "fused conv1x1_up: high-resolution plan failed for %s", c.name.c_str());
      HaloLayer& hl = P.halos[i];
      hl.add_src = lo.out_f32; hl.add_H = lo.Hout; hl.add_W = lo.Wout; hl.add_cs = lo.out_cs; hl.add_img = lo.out_img;
      hl.add_sh = sh; hl.add_sw = sw;
      hl.epi8 = 2; hl.alt = 0;                        // the additive term lives in conv_halo_kernel<352, 1, true> only
      if (hl.add_pbytes) {
        CUtensorMap m;
        rc = halo_encode_add_map(&m, lo.out_f32, lo.out_cs, lo.Wout, lo.Hout, a.b, lo.out_img, hl.ntile);
        if (rc) return rc;
        hl.add_map = (int)maps.size();
        maps.push_back(m);
      }
      P.use_tc[i] = 3;
      continue;
    }
    const bool need_halo = c.s2d_in || c.s2d_out;
    int rc = (net->no_halo && !need_halo) ? 1 : build_halo_layer(net, (int)i, io, &maps, &P.halos[i], &P.nblocks[i], &P.smem[i]);
    if (rc == 0) {
      P.use_tc[i] = 2;
      if (head) {
        const char* na = getenv("PF_TC_NO_AMAX");
        P.head_amax = (net->num_classes <= 15 && !(na && na[0] == '1')) ? 1 : 0;
        P.halos[i].amax_ncls = P.head_amax ? net->num_classes : 0;
      }
      // AvgPool2d(2,2) behind a 1x1 transition conv: averaged in the conv's epilogue (warp shuffles over the 16 x 8
      // tile), so the full-resolution tensor is never written (base.5 + pool at 1/4 resolution: 162 + 57 us per 8 frames)
      const char* np_ = getenv("PF_TC_NO_FUSE_POOL");
      if (c.pool_step >= 0 && !(np_ && np_[0] == '1') && !P.halos[i].fold && !P.halos[i].s2d_block) {
        const SegRef& po = net->steps[c.pool_step].out;
        HaloLayer& hl = P.halos[i];
        hl.pool = 1;
        hl.out_hi = reinterpret_cast<__nv_bfloat16*>(a.ptr(po.buf, po.coff));
        hl.out_lo = reinterpret_cast<__nv_bfloat16*>(a.ptr_lo(po.buf, po.coff));
        hl.out_cs = net->bufs[po.buf].cstride; hl.out_img_stride = a.img_elems[po.buf];
        P.pool_fused[c.pool_step] = 1;
      }
      continue;
    }
    if (rc != 1) return rc;
    rc = build_tc_layer(net, (int)i, io, &maps, &P.layers[i], &P.nblocks[i], &P.smem[i]);
    if (rc) return rc;
    P.use_tc[i] = 1;
  }
  if (P.maps_dev) cudaFree(P.maps_dev);
  P.maps_dev = nullptr;
  PF_CHECK_CUDA(cudaMalloc(&P.maps_dev, maps.size() * sizeof(CUtensorMap)));
  PF_CHECK_CUDA(cudaMemcpy(P.maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  P.ws = ws; P.b = a.b; P.H = a.H; P.W = a.W;
  return 0;
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
  const char* fs = getenv("PF_TC_FORCE_SIMT");
  net->force_simt = fs && fs[0] == '1';
  const char* nh = getenv("PF_TC_NO_HALO");
  net->no_halo = nh && nh[0] == '1';
  // conv1x1_up commuted with the bilinear TransitionUp (default; PF_TC_FUSE_UP=0 selects the separate upsample kernel).
  // With per-pixel global gathers of the low-resolution partial in the epilogue this was SLOWER than the upsample
  // kernel + wider 1x1 it replaces (conv1x1_up.3: 746 us vs 347 + 197 per 16 frames); with the partial's patch staged
  // per tile by TMA (HaloLayer::add_pbytes) the four levels take 0.92 instead of 1.24 ms and 1.4 GB per step of
  // upsampled tensors are neither written nor read.
  const char* nf = getenv("PF_TC_FUSE_UP");
  net->fuse_up = precision == 1 && !net->force_simt && !net->no_halo && !(nf && nf[0] == '0');
  const char* il = getenv("PF_INTERLEAVED_SLOTS");
  net->compact_slots = !(il && il[0] == '1');
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
  for (auto p : net->wtc_dev) if (p) cudaFree(p);
  if (net->zero_bias_dev) cudaFree(net->zero_bias_dev);
  if (net->plan.maps_dev) cudaFree(net->plan.maps_dev);
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
    // pre-summed rows for windows with a single label: lutsum[f][cls][o] = sum over the 9 taps
    tab.resize(lut_n + wd_n + 16 + (size_t)t * (C + 1) * 16, 0.f);
    for (int f = 0; f < t; ++f)
      for (int cls = 0; cls < C; ++cls)
        for (int o = 0; o < c.cout; ++o) {
          double sacc = 0;
          for (int tap = 0; tap < 9; ++tap) sacc += wfold[((size_t)o * c.cin + f * C + cls) * 9 + tap];
          tab[lut_n + wd_n + 16 + ((size_t)f * (C + 1) + cls) * 16 + o] = (float)sacc;
        }
    if (net->first_tab_dev) { cudaFree(net->first_tab_dev); net->first_tab_dev = nullptr; }
    if (!net->first_tab_dev) PF_CHECK_CUDA(cudaMalloc(&net->first_tab_dev, tab.size() * sizeof(float)));
    net->first_tab_floats = tab.size();
    net->first_wd_host.assign(tab.begin() + lut_n, tab.begin() + lut_n + wd_n + 16);
    PF_CHECK_CUDA(cudaMemcpy(net->first_tab_dev, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    c.loaded = true;
    return 0;
  }
  c.w_host.assign((size_t)taps * c.kpad * c.coutpad, 0.f);
  c.bias_host.assign(c.coutpad, 0.f);
  int kp = 0, ci = 0;
  if (c.s2d_in) {
    // original tap d in {0,1,2} reads input row 2Y+d-1 = quarter-res row Y+t, phase p: d=0 -> (t=-1,p=1),
    // d=1 -> (t=0,p=0), d=2 -> (t=0,p=1).  Virtual 3x3 tap index = (t+1)*3 + ..., channel = phase block * 32 + c.
    static const int tmap[3] = {0, 1, 1}, pmap[3] = {1, 0, 1};
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        const int vt = tmap[dy] * 3 + tmap[dx];
        const int blk = pmap[dy] * 2 + pmap[dx];
        for (int ch = 0; ch < c.cin; ++ch)
          for (int o = 0; o < c.cout; ++o)
            c.w_host[((size_t)vt * c.kpad + blk * 32 + ch) * c.coutpad + o] = (float)wfold[((size_t)o * c.cin + ch) * 9 + dy * 3 + dx];
      }
  } else
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
  if (net->precision == 1) {
    int rc = upload_conv_tc(net, i);
    if (rc) return rc;
    net->plan.ws = nullptr;        // weight addresses may have changed -> re-encode the tensor maps
  }
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

// Input sizes the plan handles: the two stride-2 convs need H/2 and H/4 exact (the reference's ceil == floor then,
// hardnet.py ConvLayer padding k//2), the four AvgPool2d(2,2) floor like the arena's `>> shift`, TransitionUp goes to
// the skip tensor's size whatever the ratio; W % 16: the uint8 label / mask tensor maps of the first conv need 16-byte
// row pitches.  At least one pixel must be left at 1/64 resolution.
static bool size_supported(int H, int W) { return H >= 64 && W >= 64 && H % 4 == 0 && W % 16 == 0; }

extern "C" size_t pf_bgnet_workspace_bytes(const pf_bgnet_t* net, int b, int H, int W) {
  if (!net || b <= 0 || !size_supported(H, W)) return 0;
  Arena a;
  make_arena(net, nullptr, b, H, W, &a);
  return a.total_bytes;
}

extern "C" int pf_bgnet_launches_per_forward(const pf_bgnet_t* net) {
  if (!net) return PF_EINVAL;
  int n = 0;
  for (size_t k = 0; k < net->steps.size(); ++k) {
    const Step& s = net->steps[k];
    if (s.type == STEP_POOL && net->precision == 1 && !net->force_simt && k < net->plan.pool_fused.size() && net->plan.pool_fused[k])
      continue;                                            // averaged in the producing conv's epilogue
    n += (s.type == STEP_HEAD) ? 2 : 1;
    if (s.type == STEP_CONV && net->convs[s.conv].up_nseg > 0) n += 1;
  }
  return n;
}

// Runs conv `ci` of the plan (arena addressing) on the path the handle's precision selects.
static int run_conv(pf_bgnet* net, const Arena& a, int ci, cudaStream_t st) {
  const ConvDesc& c = net->convs[ci];
  const bool head = ci == net->final_conv;
  if (net->precision == 1 && !net->force_simt && net->plan.use_tc[ci] == 3) {
    int rc = launch_conv_halo(net->plan.halos_low[ci], net->plan.maps_dev, net->plan.nblocks_low[ci], net->plan.smem_low[ci], st);
    if (rc) return rc;
    return launch_conv_halo(net->plan.halos[ci], net->plan.maps_dev, net->plan.nblocks[ci], net->plan.smem[ci], st);
  }
  if (net->precision == 1 && !net->force_simt && net->plan.use_tc[ci] == 2)
    return launch_conv_halo(net->plan.halos[ci], net->plan.maps_dev, net->plan.nblocks[ci], net->plan.smem[ci], st);
  if (net->precision == 1 && !net->force_simt && net->plan.use_tc[ci] == 1)
    return launch_conv_tc(net->plan.layers[ci], net->plan.maps_dev, net->plan.nblocks[ci], a.b, net->plan.smem[ci], st);
  ConvLaunch L;
  fill_conv_launch(net, a, c, &L);
  if (head) L.cout_store = 16;
  const bool sp = net->precision == 1;
  return launch_conv_simt(L, c.ksize, c.exec_stride(), sp, sp && !head, st);
}

// labels_dev (uint8 class ids, the one-hot conversion is implicit) or scores_dev (dense per-class planes): exactly one
static int bgnet_forward_impl(pf_bgnet_t* net, const uint8_t* labels_dev, const float* scores_dev, const float* depth_dev,
                              const uint8_t* mask_dev, int b, int H, int W, int final_h, int final_w,
                              uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev, float* out_quarter_dev,
                              float* out_full_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  PF_REQUIRE(net && (labels_dev || scores_dev) && workspace_dev, PF_EINVAL, "pf_bgnet_forward: null pointer");
  PF_REQUIRE(!net->use_depth || (depth_dev && mask_dev), PF_EINVAL, "pf_bgnet_forward: depth inputs required");
  PF_REQUIRE(b > 0 && size_supported(H, W), PF_EINVAL,
             "pf_bgnet_forward: H must be a multiple of 4, W a multiple of 16, both >= 64 (got %dx%d)", H, W);
  PF_REQUIRE(final_h > 0 && final_w > 0, PF_EINVAL, "pf_bgnet_forward: bad final size");
  PF_REQUIRE(workspace_bytes >= pf_bgnet_workspace_bytes(net, b, H, W), PF_ENOMEM, "pf_bgnet_forward: workspace too small");
  for (auto& c : net->convs) PF_REQUIRE(c.loaded, PF_ESTATE, "pf_bgnet_forward: weights of %s not loaded", c.name.c_str());
  PF_REQUIRE(!net->use_depth || net->depth_norm_set, PF_ESTATE, "pf_bgnet_forward: depth norm not set");
  cudaStream_t st = (cudaStream_t)stream;
  Arena a;
  make_arena(net, workspace_dev, b, H, W, &a);
  if (net->precision == 1 && !net->force_simt) {
    int rc = ensure_tc_plan(net, a, workspace_dev);
    if (rc) return rc;
  }
  const int split = a.split ? 1 : 0;

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
        if (scores_dev) {
          FirstDenseParams d;
          d.x = scores_dev; d.depth = depth_dev; d.mask = mask_dev; d.tab = net->first_tab_dev;
          d.out = a.ptr(c.out.buf, 0); d.out_lo = a.ptr_lo(c.out.buf, 0); d.split = split;
          d.b = b; d.t = net->num_inputs; d.H = H; d.W = W; d.Ho = H / 2; d.Wo = W / 2;
          d.ncls = net->num_classes; d.use_depth = net->use_depth; d.mean = net->depth_mean; d.std = net->depth_std;
          const size_t dsm = ((size_t)9 * d.t * (d.ncls + 1) * 16 + (size_t)9 * d.t * 16 + 16) * sizeof(float);
          PF_REQUIRE(dsm <= 160 * 1024, PF_EINVAL, "pf_bgnet_forward_dense: first-conv tables too large");
          PF_CHECK_CUDA(cudaFuncSetAttribute(first_conv_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
          first_conv_dense_kernel<<<dim3(cdiv(d.Wo, F_TW), cdiv(d.Ho, F_TH), b), F_TH * F_TW, dsm, st>>>(d);
          PF_CHECK_CUDA(cudaGetLastError());
          break;
        }
        FirstParams p;
        {
          PF_REQUIRE(((size_t)labels_dev & 15) == 0 && ((size_t)depth_dev & 15) == 0 && ((size_t)mask_dev & 15) == 0,
                     PF_EINVAL, "pf_bgnet_forward: labels / depth / mask must be 16-byte aligned");
          // the maps depend on the caller's pointers: encoded per call (host-side, ~1 us each) and passed by value
          int rc = first_conv_maps(&p.maps, labels_dev, depth_dev, mask_dev, b * net->num_inputs, H, W, net->use_depth != 0);
          if (rc) return rc;
        }
        p.tab = net->first_tab_dev;
        p.out = a.ptr(c.out.buf, 0); p.out_lo = a.ptr_lo(c.out.buf, 0); p.split = split;
        p.b = b; p.t = net->num_inputs; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2;
        p.ncls = net->num_classes; p.use_depth = net->use_depth; p.mean = net->depth_mean; p.std = net->depth_std;
        const size_t smem = net->first_tab_floats * 4 + 2 * (size_t)p.t * (F_DFRAME * 4 + 2 * F_LFRAME) + 256;   // two staging buffers
        {
          const size_t wd_n = (size_t)9 * p.t * 16;
          PF_REQUIRE(net->first_wd_host.size() == wd_n + 16 && p.t <= kFirstMaxT, PF_ESTATE, "pf_bgnet_forward: first conv not loaded");
          memcpy(p.wd, net->first_wd_host.data(), (wd_n + 16) * sizeof(float));
        }
        // the attribute is per device and per function: set it on every launch (a host-side store, no driver round trip)
        PF_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        PF_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        PF_REQUIRE(smem <= 160 * 1024, PF_EINVAL, "pf_bgnet_forward: first-conv tables too large");
        const int total_tiles = cdiv(p.Ho, F_TH) * cdiv(p.Wo, F_TW) * b;
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        const int grid = total_tiles < kNumSMs * per_sm ? total_tiles : kNumSMs * per_sm;   // persistent: whole waves
        if (p.t == 3) first_conv_kernel<3><<<grid, F_TH * F_TW, smem, st>>>(p);
        else first_conv_kernel<0><<<grid, F_TH * F_TW, smem, st>>>(p);
        PF_CHECK_CUDA(cudaGetLastError());
        break;
      }
      case STEP_CONV: {
        int rc = run_conv(net, a, s.conv, st);
        if (rc) return rc;
        break;
      }
      case STEP_POOL: {
        const int k = step_i - 1;
        if (net->precision == 1 && !net->force_simt && k < (int)net->plan.pool_fused.size() && net->plan.pool_fused[k]) break;
        const SegRef& in = s.in[0];
        const BufDesc& ib = net->bufs[in.buf];
        const BufDesc& ob = net->bufs[s.out.buf];
        PoolParams p;
        p.in = a.ptr(in.buf, in.coff); p.in_lo = a.ptr_lo(in.buf, in.coff); p.in_cs = ib.cstride; p.in_img = a.img_elems[in.buf];
        p.out = a.ptr(s.out.buf, s.out.coff); p.out_lo = a.ptr_lo(s.out.buf, s.out.coff); p.out_cs = ob.cstride;
        p.out_img = a.img_elems[s.out.buf];
        p.b = b; p.Ho = H >> ob.shift; p.Wo = W >> ob.shift; p.Wi = W >> ib.shift; p.c4 = in.cpad() / 8; p.split = split;
        avgpool2_kernel<<<dim3(cdiv(p.Wo * p.c4, 256), p.Ho, b), 256, 0, st>>>(p);
        PF_CHECK_CUDA(cudaGetLastError());
        break;
      }
      case STEP_UPSAMPLE: {
        UpParams p;
        p.nseg = (int)s.in.size();
        int ctot = 0;
        for (int k = 0; k < p.nseg; ++k) {
          const SegRef& r = s.in[k];
          p.segs[k].base = a.ptr(r.buf, r.coff);
          p.segs[k].base_lo = a.ptr_lo(r.buf, r.coff);
          p.segs[k].cstride = net->bufs[r.buf].cstride;
          p.segs[k].cpad = r.cpad();
          p.in_img[k] = a.img_elems[r.buf];
          ctot += r.cpad();
        }
        const BufDesc& ib = net->bufs[s.in[0].buf];
        const BufDesc& ob = net->bufs[s.out.buf];
        p.out = a.ptr(s.out.buf, 0); p.out_lo = a.ptr_lo(s.out.buf, 0); p.out_cs = ob.cstride;
        p.out_img = a.img_elems[s.out.buf];
        p.b = b; p.Hi = H >> ib.shift; p.Wi = W >> ib.shift; p.Ho = H >> ob.shift; p.Wo = W >> ob.shift;
        p.c4_total = ctot / 8; p.split = split;
        p.sh = p.Ho > 1 ? (float)(p.Hi - 1) / (float)(p.Ho - 1) : 0.f;
        p.sw = p.Wo > 1 ? (float)(p.Wi - 1) / (float)(p.Wo - 1) : 0.f;
        upsample_bilinear_kernel<<<dim3(cdiv(cdiv(p.Wo, kUpRun) * p.c4_total, 256), p.Ho, b), 256, 0, st>>>(p);
        PF_CHECK_CUDA(cudaGetLastError());
        break;
      }
      case STEP_HEAD: {
        int rc = run_conv(net, a, s.conv, st);
        if (rc) return rc;
        const int h = H / 4, w = W / 4;
        const float* q = reinterpret_cast<const float*>(a.ptr(net->quarter_buf, 0));
        if (out_quarter_dev) {
          const size_t total = (size_t)b * net->num_classes * h * w;
          nhwc16_to_nchw_kernel<<<grid_for(total, 256), 256, 0, st>>>(q, out_quarter_dev, b, net->num_classes, h, w);
          PF_CHECK_CUDA(cudaGetLastError());
        }
        const float sh = final_h > 1 ? (float)(h - 1) / (float)(final_h - 1) : 0.f;
        const float sw = final_w > 1 ? (float)(w - 1) / (float)(final_w - 1) : 0.f;
        const size_t total = (size_t)b * final_h * final_w;
        if (final_w % 4 == 0 && 3.f * sw < 0.999f && (((size_t)out_seg_u8_dev) & 3) == 0 && (((size_t)out_seg_i64_dev) & 15) == 0) {
          const dim3 grid(cdiv(final_w / 4, 256), final_h, b);
          const bool amax = net->precision == 1 && !net->force_simt && net->plan.head_amax && net->plan.use_tc[net->final_conv] == 2;
          if (out_full_dev)
            upsample_argmax_strip_kernel<true, false><<<grid, 256, 0, st>>>(q, net->num_classes, h, w, final_h, final_w, sh, sw, out_seg_u8_dev,
                                                                           (long long*)out_seg_i64_dev, out_full_dev);
          else if (amax)
            upsample_argmax_strip_kernel<false, true><<<grid, 256, 0, st>>>(q, net->num_classes, h, w, final_h, final_w, sh, sw, out_seg_u8_dev,
                                                                           (long long*)out_seg_i64_dev, nullptr);
          else
            upsample_argmax_strip_kernel<false, false><<<grid, 256, 0, st>>>(q, net->num_classes, h, w, final_h, final_w, sh, sw, out_seg_u8_dev,
                                                                            (long long*)out_seg_i64_dev, nullptr);
        }
        else
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

extern "C" int pf_bgnet_forward(pf_bgnet_t* net, const uint8_t* labels_dev, const float* depth_dev,
                                const uint8_t* mask_dev, int b, int H, int W, int final_h, int final_w,
                                uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev, float* out_quarter_dev,
                                float* out_full_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  PF_REQUIRE(labels_dev, PF_EINVAL, "pf_bgnet_forward: null pointer");
  return bgnet_forward_impl(net, labels_dev, nullptr, depth_dev, mask_dev, b, H, W, final_h, final_w, out_seg_u8_dev,
                            out_seg_i64_dev, out_quarter_dev, out_full_dev, workspace_dev, workspace_bytes, stream);
}

extern "C" int pf_bgnet_forward_dense(pf_bgnet_t* net, const float* scores_dev, const float* depth_dev,
                                      const uint8_t* mask_dev, int b, int H, int W, int final_h, int final_w,
                                      uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev, float* out_quarter_dev,
                                      float* out_full_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  PF_REQUIRE(scores_dev, PF_EINVAL, "pf_bgnet_forward_dense: null pointer");
  return bgnet_forward_impl(net, nullptr, scores_dev, depth_dev, mask_dev, b, H, W, final_h, final_w, out_seg_u8_dev,
                            out_seg_i64_dev, out_quarter_dev, out_full_dev, workspace_dev, workspace_bytes, stream);
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

// One ConvLayer on an NCHW fp32 tensor through the SAME kernels the forward uses for it
// (fp32 SIMT, split-bf16 SIMT for stride 2 / PF_TC_FORCE_SIMT, or the tcgen05 kernel).
extern "C" int pf_bgnet_debug_conv(pf_bgnet_t* net, int i, const float* x_nchw_dev, int b, int H, int W,
                                   float* y_nchw_dev, void* stream) {
  PF_REQUIRE(net && x_nchw_dev && y_nchw_dev, PF_EINVAL, "pf_bgnet_debug_conv: null pointer");
  PF_REQUIRE(i > 0 && i < (int)net->convs.size(), PF_EINVAL, "pf_bgnet_debug_conv: index must be 1..num_convs");
  const ConvDesc& c = net->convs[i];
  PF_REQUIRE(c.loaded, PF_ESTATE, "pf_bgnet_debug_conv: weights not loaded");
  cudaStream_t st = (cudaStream_t)stream;
  const bool head = i == net->final_conv;
  const bool split = net->precision == 1;
  const bool split_out = split && !head;
  const bool s2i = c.s2d_in, s2o = c.s2d_out;                 // only set with split storage
  PF_REQUIRE(!(s2i || s2o) || (H % 2 == 0 && W % 2 == 0), PF_EINVAL, "pf_bgnet_debug_conv: even H, W required");
  const int es = c.exec_stride();
  const int He = s2i ? H / 2 : H, We = s2i ? W / 2 : W;       // input extent as the kernel sees it
  const int Ho = (He + es - 1) / es, Wo = (We + es - 1) / es; // conv output extent
  const int cs_out = head ? 16 : (s2o ? 128 : padc(c.cout));
  const size_t in_elems = (size_t)b * He * We * c.kpad;
  const size_t out_elems = s2o ? (size_t)b * (Ho / 2) * (Wo / 2) * 128 : (size_t)b * Ho * Wo * cs_out;
  const size_t esz_in = split ? 2 : 4;
  char *xin = nullptr, *xin_lo = nullptr, *yout = nullptr, *yout_lo = nullptr;
  float* xpk = nullptr;
  PF_CHECK_CUDA(cudaMalloc(&xin, in_elems * esz_in));
  PF_CHECK_CUDA(cudaMalloc(&xin_lo, in_elems * 2));
  PF_CHECK_CUDA(cudaMalloc(&xpk, (size_t)b * H * W * c.kpad * 4));
  PF_CHECK_CUDA(cudaMalloc(&yout, out_elems * 4));
  PF_CHECK_CUDA(cudaMalloc(&yout_lo, out_elems * 2));
  if (s2i) {
    nchw_to_s2d_kernel<<<grid_for(in_elems, 256), 256, 0, st>>>(x_nchw_dev, reinterpret_cast<unsigned short*>(xin),
                                                              reinterpret_cast<unsigned short*>(xin_lo), b, c.cin, H, W);
    PF_CHECK_CUDA(cudaGetLastError());
  } else {
    // NCHW -> NCHW with each input slice moved to its padded channel position -> NHWC (kpad channels)
    PF_CHECK_CUDA(cudaMemsetAsync(xpk, 0, (size_t)b * H * W * c.kpad * 4, st));
    int src = 0, dst = 0;
    for (auto& s : c.in) {
      PF_CHECK_CUDA(cudaMemcpy2DAsync(xpk + (size_t)dst * H * W, (size_t)c.kpad * H * W * 4,
                                      x_nchw_dev + (size_t)src * H * W, (size_t)c.cin * H * W * 4,
                                      (size_t)s.c * H * W * 4, b, cudaMemcpyDeviceToDevice, st));
      src += s.c; dst += s.cpad();
    }
    nchw_to_nhwc_kernel<<<grid_for(in_elems, 256), 256, 0, st>>>(xpk, xin, xin_lo, split ? 1 : 0, b, c.kpad, H, W, c.kpad);
    PF_CHECK_CUDA(cudaGetLastError());
  }
  int rc = 0;
  const bool use_tc = split && !net->force_simt && es == 1;
  CUtensorMap* maps_dev = nullptr;
  if (use_tc) {
    TcIo io;
    int dst = 0;
    for (size_t s = 0; s < c.in.size(); ++s) {
      io.in_hi[s] = xin + (size_t)dst * 2; io.in_lo[s] = xin_lo + (size_t)dst * 2;
      io.in_cs[s] = c.kpad; io.in_img[s] = (size_t)He * We * c.kpad;
      dst += c.in[s].cpad();
    }
    io.Hin = He; io.Win = We; io.Hout = Ho; io.Wout = Wo; io.b = b;
    io.out_hi = head ? nullptr : yout; io.out_lo = head ? nullptr : yout_lo;
    io.out_f32 = head ? reinterpret_cast<float*>(yout) : nullptr;
    io.out_cs = cs_out; io.out_img = out_elems / b;
    std::vector<CUtensorMap> maps;
    TcLayer L;
    HaloLayer HL;
    int nblocks; size_t smem;
    int kind = 2;
    rc = (net->no_halo && !s2i && !s2o) ? 1 : build_halo_layer(net, i, io, &maps, &HL, &nblocks, &smem);
    if (rc == 1) { kind = 1; rc = build_tc_layer(net, i, io, &maps, &L, &nblocks, &smem); }
    if (rc == 0) {
      PF_CHECK_CUDA(cudaMalloc(&maps_dev, maps.size() * sizeof(CUtensorMap)));
      PF_CHECK_CUDA(cudaMemcpyAsync(maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
      PF_CHECK_CUDA(cudaStreamSynchronize(st));
      long long* ts_dev = nullptr;
      if (kind == 2 && getenv("PF_HALO_TS")) {
        cudaMalloc(&ts_dev, 96 * sizeof(long long));
        cudaMemset(ts_dev, 0, 96 * sizeof(long long));
        HL.dbg_ts = ts_dev;
        HL.dbg_mode = atoi(getenv("PF_HALO_TS")) >> 1;   // PF_HALO_TS=1 plain, 3 no-TMA, 5 no-stores, 7 both
        launch_conv_halo(HL, maps_dev, nblocks, smem, st);      // warm (weights / descriptors in L2)
      }
      rc = kind == 2 ? launch_conv_halo(HL, maps_dev, nblocks, smem, st) : launch_conv_tc(L, maps_dev, nblocks, b, smem, st);
      if (ts_dev) {
        long long ts[96];
        cudaStreamSynchronize(st);
        cudaMemcpy(ts, ts_dev, sizeof(ts), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[halo ts] %s tiles=%d nblk=%d res=%d SA=%d SB=%d ntile=%d: setup %lld, w-ready %lld, first-A %lld, "
                "tile0 mma-issued %lld, tile0 acc-ready %lld, tile0 epi-done %lld, all-mma-issued %lld, end %lld; waits: "
                "mma<-A %lld, mma<-tmem %lld, tma<-emptyA %lld, epi<-acc %lld (cycles)\n",
                c.name.c_str(), HL.tiles_x * HL.tiles_y * HL.batch, nblocks, HL.resident, HL.stages_a, HL.stages_b, HL.ntile,
                ts[1] - ts[0], ts[2] - ts[0], ts[3] - ts[0], ts[4] - ts[0], ts[5] - ts[0], ts[6] - ts[0], ts[7] - ts[0],
                ts[8] - ts[0], ts[9], ts[10], ts[11], ts[12]);
        for (int k = 0; k < 24; ++k)
          fprintf(stderr, "   chunk %2d: tma-issued %7lld  a-ready %7lld  mma-issued %7lld\n", k, ts[16 + k] - ts[0], ts[40 + k] - ts[0],
                  ts[64 + k] - ts[0]);
        cudaFree(ts_dev);
      }
    }
  } else {
    ConvLaunch L;
    L.nseg = (int)c.in.size();
    int dst = 0;
    for (int s = 0; s < L.nseg; ++s) {
      L.segs[s].base = xin + (size_t)dst * esz_in; L.segs[s].base_lo = xin_lo + (size_t)dst * 2;
      L.segs[s].cstride = c.kpad; L.segs[s].cpad = c.in[s].cpad();
      L.in_img_stride[s] = (size_t)He * We * c.kpad;
      dst += c.in[s].cpad();
    }
    L.b = b; L.Hin = He; L.Win = We; L.Hout = Ho; L.Wout = Wo;
    L.out = yout; L.out_lo = yout_lo; L.out_cstride = cs_out; L.out_img_stride = out_elems / b;
    L.w = c.w_dev; L.bias = c.bias_dev; L.kpad = c.kpad; L.coutpad = c.coutpad; L.relu = c.relu ? 1 : 0;
    L.cout_store = s2o ? 32 : cs_out;
    L.s2d_block = s2o ? 32 : 0;
    rc = launch_conv_simt(L, c.ksize, es, split, split_out, st);
  }
  if (rc == 0) {
    const size_t total = (size_t)b * c.cout * Ho * Wo;
    if (s2o)
      s2d_to_nchw_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<unsigned short*>(yout),
                                                             reinterpret_cast<unsigned short*>(yout_lo), y_nchw_dev, b,
                                                             c.cout, Ho, Wo);
    else
      nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, st>>>(yout, yout_lo, split_out ? 1 : 0, y_nchw_dev, b, c.cout, Ho, Wo, cs_out);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("pf_bgnet_debug_conv: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  cudaFree(xin); cudaFree(xin_lo); cudaFree(xpk); cudaFree(yout); cudaFree(yout_lo);
  if (maps_dev) cudaFree(maps_dev);
  return rc;
}
