// The generator (reference code/models.py:61-86) as a sequence of tensor-core conv launches, and
// the recurrent clip loop (reference main.py:173-219) kept entirely on the device.
#include <stdlib.h>
#include <vector>

#include "tg_frame.cuh"

namespace tg {

struct GenLayer {
  int kind;        // kConv3x3 / kConvT3x3s2
  int cin, cout;   // real channel counts
  int has_bias;
  size_t w_off;    // element offset of the weight inside the flat state_dict parameter buffer
  size_t b_off;    // element offset of the bias (if any)
  size_t p_off;    // byte offset of the packed blob
};

// state_dict order: conv.0.{weight,bias}, resids.i.0.{weight,bias}, resids.i.2.weight,
// conv_trans.0.{w,b}, conv_trans.2.0.{w,b}, conv_trans.2.2.w, conv_trans.3.0.{w,b},
// conv_trans.3.2.w, conv_trans.4.{w,b}, conv_trans.6.{w,b}, output.{w,b}   
static std::vector<GenLayer> gen_layers(int nres, size_t* n_params, size_t* packed_bytes) {
  std::vector<GenLayer> L;
  size_t po = 0, bo = 0;
  auto add = [&](int kind, int cin, int cout, int has_bias) {
    GenLayer l{kind, cin, cout, has_bias, po, 0, bo};
    po += static_cast<size_t>(cin) * cout * 9;
    if (has_bias) { l.b_off = po; po += cout; }
    bo += tg_packed_conv_bytes(kind, cin, cout);
    L.push_back(l);
  };
  add(kConv3x3, 51, 64, 1);
  for (int i = 0; i < nres; ++i) { add(kConv3x3, 64, 64, 1); add(kConv3x3, 64, 64, 0); }
  add(kConvT3x3s2, 64, 64, 1);
  add(kConv3x3, 64, 64, 1);
  add(kConv3x3, 64, 64, 0);
  add(kConv3x3, 64, 128, 1);
  add(kConv3x3, 128, 128, 0);
  add(kConvT3x3s2, 128, 128, 1);
  add(kConv3x3, 128, 64, 1);
  add(kConv3x3, 64, 3, 1);
  if (n_params) *n_params = po;
  if (packed_bytes) *packed_bytes = bo;
  return L;
}

static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct GenWorkspace {
  size_t x0, a[3], b[2], c[2], d, e, flags, flag_count, rgbx, total;
};
static GenWorkspace gen_ws(int n, int h, int w) {
  GenWorkspace ws;
  const size_t px = static_cast<size_t>(n) * h * w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align256(bytes); return r; };
  ws.x0 = take(px * 64 * 2);
  for (int i = 0; i < 3; ++i) ws.a[i] = take(px * 64 * 2);
  for (int i = 0; i < 2; ++i) ws.b[i] = take(px * 4 * 64 * 2);
  for (int i = 0; i < 2; ++i) ws.c[i] = take(px * 4 * 128 * 2);
  ws.d = take(px * 16 * 128 * 2);
  ws.e = take(px * 16 * 64 * 2);
  // per-item completion counters of the frame kernel: trunk/convT64 items at 1x, 2x items, 4x items
  auto tiles = [&](int s) { return static_cast<size_t>(n) * frame_tiles_max(h * s, w * s); };
  ws.flag_count = 132 * tiles(1) + 16 * tiles(2) + 4 * tiles(4) + 256;   // covers num_resblock <= 64 (+ one pair-padding counter per segment)
  ws.flags = take(ws.flag_count * 4);
  ws.rgbx = take(px * 16 * 16);                    // float4 per HR pixel: interleaved copy of the last output (clip loops)
  ws.total = o;
  return ws;
}

// The 41 layers of one forward as a list (input/output/residual buffers inside the workspace).
static std::vector<FrLayer> gen_plan(const std::vector<GenLayer>& L, int nres, const void* x, void* out,
                                     float* logits, uint8_t* wsp, int n, int h, int w, long long out_nstride,
                                     bool write_rgbx = false, int out_mode = kOutNCHWf32Sigmoid) {
  const GenWorkspace ws = gen_ws(n, h, w);
  std::vector<FrLayer> P;
  int li = 0;
  auto add = [&](const void* in, void* o, const void* resid, int relu, int hh, int ww) {
    FrLayer f{};
    f.kind = L[li].kind;
    f.cin_pad = cin_padded(L[li].cin);
    f.cout_pad = cout_padded(L[li].cout);
    f.out_mode = kOutNHWCbf16;
    f.relu = relu; f.h = hh; f.w = ww;
    f.in = in; f.out = o; f.resid = resid; f.out2 = nullptr;
    f.blob_off = L[li].p_off;
    f.out_nstride = 0;
    P.push_back(f);
    ++li;
  };
  void* a[3] = {wsp + ws.a[0], wsp + ws.a[1], wsp + ws.a[2]};
  add(x, a[0], nullptr, 1, h, w);                                            // conv.0 + ReLU
  int cur = 0;
  for (int i = 0; i < nres; ++i) {                                           // net = block(net) + net
    const int t = (cur + 1) % 3, nx = (cur + 2) % 3;
    add(a[cur], a[t], nullptr, 1, h, w);
    add(a[t], a[nx], a[cur], 0, h, w);
    cur = nx;
  }
  void* b0 = wsp + ws.b[0]; void* b1 = wsp + ws.b[1];
  void* c0 = wsp + ws.c[0]; void* c1 = wsp + ws.c[1];
  void* d = wsp + ws.d; void* e = wsp + ws.e;
  add(a[cur], b0, nullptr, 1, h, w);                                         // conv_trans.0 (x2) + ReLU
  add(b0, b1, nullptr, 1, 2 * h, 2 * w);                                     // conv_trans.2.0 + ReLU
  add(b1, b0, nullptr, 0, 2 * h, 2 * w);                                     // conv_trans.2.2 (no skip)
  add(b0, c0, nullptr, 1, 2 * h, 2 * w);                                     // conv_trans.3.0 + ReLU
  add(c0, c1, nullptr, 0, 2 * h, 2 * w);                                     // conv_trans.3.2
  add(c1, d, nullptr, 1, 2 * h, 2 * w);                                      // conv_trans.4 (x2) + ReLU
  add(d, e, nullptr, 1, 4 * h, 4 * w);                                       // conv_trans.6 + ReLU
  add(e, out, nullptr, 0, 4 * h, 4 * w);                                     // output + sigmoid
  P.back().out_mode = out_mode;
  P.back().out2 = logits;
  P.back().out_nstride = out_nstride;
  P.back().out_rgbx = write_rgbx ? wsp + ws.rgbx : nullptr;
  return P;
}

// flags_zeroed: the per-item completion counters were already cleared by an earlier kernel of the stream
static int gen_forward_impl(const std::vector<GenLayer>& L, const uint8_t* packed, int nres, const void* x,
                            void* out, float* logits, uint8_t* wsp, int n, int h, int w, int amode,
                            long long out_nstride, bool flags_zeroed, cudaStream_t st, bool write_rgbx = false,
                            int out_mode = kOutNCHWf32Sigmoid) {
  TG_CHECK_ARG(out_mode == kOutNCHWf32Sigmoid || amode == TG_AMODE_FRAME, "generator: fp16 / uint8 outputs need TG_AMODE_FRAME");
  const std::vector<FrLayer> P = gen_plan(L, nres, x, out, logits, wsp, n, h, w, out_nstride,
                                          write_rgbx && amode == TG_AMODE_FRAME, out_mode);
  if (amode == TG_AMODE_FRAME) {
    size_t pb = 0;
    for (auto& l : L) pb += tg_packed_conv_bytes(l.kind, l.cin, l.cout);
    return launch_frame(P.data(), static_cast<int>(P.size()), packed, pb, n,
                        reinterpret_cast<uint32_t*>(wsp + gen_ws(n, h, w).flags), gen_ws(n, h, w).flag_count,
                        flags_zeroed, st);
  }
  for (const FrLayer& f : P) {
    const uint8_t* blob = packed + f.blob_off;
    const float* bias = reinterpret_cast<const float*>(blob + packed_weight_bytes(f.cin_pad, f.cout_pad));
    int rc = launch_conv_tc(f.kind, f.out_mode, f.in, blob, bias, f.resid, f.out, f.out2, n, f.h, f.w, f.cin_pad,
                            f.cout_pad, f.relu, amode, f.out_nstride, st);
    if (rc) return rc;
  }
  return TG_OK;
}

}  // namespace tg

using namespace tg;

extern "C" size_t tg_gen_param_count(int num_resblock) {
  size_t np = 0;
  gen_layers(num_resblock, &np, nullptr);
  return np;
}
extern "C" size_t tg_gen_packed_bytes(int num_resblock) {
  size_t pb = 0;
  gen_layers(num_resblock, nullptr, &pb);
  return pb;
}
extern "C" int tg_gen_pack(const float* flat_params, int num_resblock, void* packed, void* stream) {
  TG_CHECK_ARG(flat_params && packed, "gen_pack: null pointer");
  TG_CHECK_ARG(num_resblock >= 0 && num_resblock <= 64, "gen_pack: bad num_resblock %d", num_resblock);
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 255) == 0, "gen_pack: packed must be 256-byte aligned");
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  std::vector<PackJobSpec> jobs;
  for (auto& l : L)
    jobs.push_back(PackJobSpec{l.kind, flat_params + l.w_off, l.has_bias ? flat_params + l.b_off : nullptr, l.cin, l.cout,
                               static_cast<uint8_t*>(packed) + l.p_off});
  return pack_weights_batched(jobs.data(), static_cast<int>(jobs.size()), static_cast<cudaStream_t>(stream));
}
extern "C" size_t tg_gen_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  return gen_ws(n, h, w).total;
}

extern "C" size_t tg_workspace_bytes_gen_forward(int n, int h, int w) { return tg_gen_workspace_bytes(n, h, w); }

extern "C" int tg_gen_forward(const void* packed, int num_resblock, const void* x_nhwc, float* out,
                              float* logits_or_null, void* workspace, size_t workspace_bytes, int n, int h, int w,
                              int amode, void* stream) {
  TG_CHECK_ARG(packed && x_nhwc && out && workspace, "gen_forward: null pointer");
  TG_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "gen_forward: bad shape");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gen_forward: workspace must be 256-byte aligned");
  if (workspace_bytes < gen_ws(n, h, w).total) {
    tg_set_error("gen_forward: workspace too small (%zu < %zu)", workspace_bytes, gen_ws(n, h, w).total);
    return TG_ERR_WORKSPACE;
  }
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  TG_CHECK_ARG(amode == TG_AMODE_HALO || amode == TG_AMODE_DX3 || amode == TG_AMODE_FRAME, "gen_forward: bad amode %d", amode);
  return gen_forward_impl(L, static_cast<const uint8_t*>(packed), num_resblock, x_nhwc, out, logits_or_null,
                          static_cast<uint8_t*>(workspace), n, h, w, amode, 0, false, static_cast<cudaStream_t>(stream));
}

// One recurrent step: frame input (warp of the previous HR estimate + space-to-depth + concat) and
// generator forward.  Batch strides are in elements.
// chained: prev_hr is the unmodified output of the previous step on this workspace, so its interleaved copy in the
// workspace may be gathered instead (frame mode only; every frame-mode step leaves that copy behind).
static int gen_clip_step_impl(const std::vector<GenLayer>& L, const uint8_t* packed, int nres, const float* lr_t,
                              const float* lr_prev, const float* prev_hr, void* out_t, uint8_t* wsp, int n, int h,
                              int w, long long lr_bs, long long prev_bs, long long out_bs, int amode, cudaStream_t st,
                              bool chained = false, int out_mode = kOutNCHWf32Sigmoid) {
  const GenWorkspace ws = gen_ws(n, h, w);
  void* x0 = wsp + ws.x0;
  const bool frame_mode = (amode == TG_AMODE_FRAME);
  static const bool rgbx_on = []() { const char* e = getenv("TG_RGBX"); return !(e && e[0] == '0'); }();   // A/B knob
  if (!rgbx_on) chained = false;
  size_t nflags = 0;
  if (frame_mode) {
    const std::vector<FrLayer> P = gen_plan(L, nres, x0, out_t, nullptr, wsp, n, h, w, out_bs);
    nflags = frame_flag_count(P.data(), static_cast<int>(P.size()), n);
  }
  // the frame-input kernel also clears the frame kernel's completion counters (it runs strictly
  // after the previous frame kernel and strictly before the next one)
  int rc = fused_input_launch(lr_t, lr_prev, prev_hr, x0, n, h, w, lr_bs, prev_bs,
                              frame_mode ? reinterpret_cast<uint32_t*>(wsp + ws.flags) : nullptr, nflags, st,
                              (chained && frame_mode && prev_hr) ? wsp + ws.rgbx : nullptr);
  if (rc) return rc;
  return gen_forward_impl(L, packed, nres, x0, out_t, nullptr, wsp, n, h, w, amode, out_bs, frame_mode, st, rgbx_on, out_mode);
}

static int check_ws(const char* who, const void* workspace, size_t workspace_bytes, int n, int h, int w) {
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  if (workspace_bytes < gen_ws(n, h, w).total) {
    tg_set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, gen_ws(n, h, w).total);
    return TG_ERR_WORKSPACE;
  }
  return TG_OK;
}

static int clip_step_common(const void* packed, int num_resblock, const float* lr_t, const float* lr_prev,
                            const float* prev_hr, float* out_t, void* workspace, size_t workspace_bytes, int n,
                            int h, int w, long long lr_batch_stride, long long prev_batch_stride,
                            long long out_batch_stride, int amode, void* stream, bool chained) {
  TG_CHECK_ARG(packed && lr_t && out_t && workspace, "gen_clip_step: null pointer");
  TG_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "gen_clip_step: bad shape");
  TG_CHECK_ARG(amode == TG_AMODE_HALO || amode == TG_AMODE_DX3 || amode == TG_AMODE_FRAME, "gen_clip_step: bad amode %d", amode);
  TG_CHECK_ARG((lr_prev == nullptr) == (prev_hr == nullptr), "gen_clip_step: lr_prev and prev_hr go together");
  if (int rc = check_ws("gen_clip_step", workspace, workspace_bytes, n, h, w)) return rc;
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  return gen_clip_step_impl(L, static_cast<const uint8_t*>(packed), num_resblock, lr_t, lr_prev, prev_hr, out_t,
                            static_cast<uint8_t*>(workspace), n, h, w, lr_batch_stride, prev_batch_stride,
                            out_batch_stride, amode, static_cast<cudaStream_t>(stream), chained);
}

extern "C" int tg_gen_clip_step(const void* packed, int num_resblock, const float* lr_t, const float* lr_prev,
                                const float* prev_hr, float* out_t, void* workspace, size_t workspace_bytes, int n,
                                int h, int w, long long lr_batch_stride, long long prev_batch_stride,
                                long long out_batch_stride, int amode, void* stream) {
  return clip_step_common(packed, num_resblock, lr_t, lr_prev, prev_hr, out_t, workspace, workspace_bytes, n, h, w,
                          lr_batch_stride, prev_batch_stride, out_batch_stride, amode, stream, false);
}

extern "C" int tg_gen_clip_step_chained(const void* packed, int num_resblock, const float* lr_t, const float* lr_prev,
                                        const float* prev_hr, float* out_t, void* workspace, size_t workspace_bytes,
                                        int n, int h, int w, long long lr_batch_stride, long long prev_batch_stride,
                                        long long out_batch_stride, int amode, void* stream) {
  return clip_step_common(packed, num_resblock, lr_t, lr_prev, prev_hr, out_t, workspace, workspace_bytes, n, h, w,
                          lr_batch_stride, prev_batch_stride, out_batch_stride, amode, stream, true);
}

extern "C" int tg_gen_clip_step_fmt(const void* packed, int num_resblock, const float* lr_t, const float* lr_prev,
                                    int first_frame, void* out_t, int out_format, void* workspace, size_t workspace_bytes,
                                    int n, int h, int w, long long lr_batch_stride, long long out_batch_stride, void* stream) {
  TG_CHECK_ARG(packed && lr_t && out_t && workspace, "gen_clip_step_fmt: null pointer");
  TG_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "gen_clip_step_fmt: bad shape");
  TG_CHECK_ARG(out_format == TG_OUT_F32 || out_format == TG_OUT_F16 || out_format == TG_OUT_U8, "gen_clip_step_fmt: bad out_format %d", out_format);
  TG_CHECK_ARG(first_frame || lr_prev, "gen_clip_step_fmt: lr_prev is needed for every frame but the first");
  TG_CHECK_ARG(out_batch_stride % 3 == 0, "gen_clip_step_fmt: out_batch_stride must be a multiple of 3");
  static const bool rgbx_on = []() { const char* e = getenv("TG_RGBX"); return !(e && e[0] == '0'); }();
  TG_CHECK_ARG(rgbx_on, "gen_clip_step_fmt: needs the workspace's interleaved copy (TG_RGBX=0 disables it)");
  if (int rc = check_ws("gen_clip_step_fmt", workspace, workspace_bytes, n, h, w)) return rc;
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  const int mode = out_format == TG_OUT_F32 ? kOutNCHWf32Sigmoid : (out_format == TG_OUT_F16 ? kOutNCHWf16Sigmoid : kOutNHWCu8Sigmoid);
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  // the previous estimate is gathered from the workspace's interleaved f32 copy (chained step): prev_hr only says "not the first frame"
  const float* prev_marker = first_frame ? nullptr : reinterpret_cast<const float*>(wsp + gen_ws(n, h, w).rgbx);
  return gen_clip_step_impl(L, static_cast<const uint8_t*>(packed), num_resblock, lr_t, first_frame ? nullptr : lr_prev, prev_marker,
                            out_t, wsp, n, h, w, lr_batch_stride, 0, out_batch_stride, TG_AMODE_FRAME,
                            static_cast<cudaStream_t>(stream), /*chained=*/true, mode);
}

extern "C" int tg_gen_clip_forward(const void* packed, int num_resblock, const float* lr, float* out, void* workspace,
                                   size_t workspace_bytes, int n, int t, int h, int w, int amode, void* stream) {
  TG_CHECK_ARG(packed && lr && out && workspace, "gen_clip_forward: null pointer");
  TG_CHECK_ARG(n >= 1 && t >= 1 && h >= 1 && w >= 1, "gen_clip_forward: bad shape");
  TG_CHECK_ARG(amode == TG_AMODE_HALO || amode == TG_AMODE_DX3 || amode == TG_AMODE_FRAME, "gen_clip_forward: bad amode %d", amode);
  if (int rc = check_ws("gen_clip_forward", workspace, workspace_bytes, n, h, w)) return rc;
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  const long long lr_frame = 3LL * h * w, hr_frame = 48LL * h * w;
  const long long lr_bs = lr_frame * t, hr_bs = hr_frame * t;
  for (int f = 0; f < t; ++f) {
    int rc = gen_clip_step_impl(L, static_cast<const uint8_t*>(packed), num_resblock, lr + f * lr_frame,
                                f ? lr + (f - 1) * lr_frame : nullptr, f ? out + (f - 1) * hr_frame : nullptr,
                                out + f * hr_frame, static_cast<uint8_t*>(workspace), n, h, w, lr_bs, hr_bs, hr_bs, amode,
                                static_cast<cudaStream_t>(stream), /*chained=*/true);
    if (rc) return rc;
  }
  return TG_OK;
}

extern "C" int tg_frame_set_pair(int on) {
  tg::frame_set_pair(on);
  return TG_OK;
}

extern "C" int tg_frame_set_trace(void* buf, size_t bytes) {
  tg::frame_set_trace(static_cast<unsigned long long*>(buf), buf ? bytes / 8 : 0);
  return TG_OK;
}

// =====================================================================================================
// Training: forward that keeps every activation, and the backward pass (reference code/train.py:336:
// scaler.scale(gen_loss).backward() through generator.forward; inputs are detached, code/train.py:90,108).
// =====================================================================================================
namespace tg {

struct GenTrainWs {
  size_t x_in;
  std::vector<size_t> net, t;          // net[0..nres] (64ch @1x), t[0..nres-1] = relu(conv1): each family one contiguous block
  size_t b0, b1, b2, c0, c1, d, e;     // upsampling stack activations
  // gradient buffers, one per tensor (the data-gradient chain runs as persistent multi-layer launches and the weight
  // gradients read them afterwards): g_net[0..nres], g_t[0..nres-1] contiguous like net / t
  size_t g_z, g_e, g_d, g_c1, g_c0, g_b[3];
  std::vector<size_t> g_net, g_t;
  size_t flags, flag_count, total;
};
static GenTrainWs gen_train_ws(int n, int h, int w, int nres) {
  GenTrainWs ws;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align256(bytes); return r; };
  const size_t px = static_cast<size_t>(n) * h * w;
  auto family = [&](std::vector<size_t>& v, int count) {    // `count` blocks of px * 128 bytes back to back (no padding between)
    const size_t base = take(px * 128 * static_cast<size_t>(count > 0 ? count : 1));
    for (int i = 0; i < count; ++i) v.push_back(base + i * px * 128);
  };
  ws.x_in = take(px * 128);
  family(ws.net, nres + 1);
  family(ws.t, nres);
  ws.b0 = take(px * 4 * 128); ws.b1 = take(px * 4 * 128); ws.b2 = take(px * 4 * 128);
  ws.c0 = take(px * 4 * 256); ws.c1 = take(px * 4 * 256);
  ws.d = take(px * 16 * 256); ws.e = take(px * 16 * 128);
  ws.g_z = take(px * 16 * 128); ws.g_e = take(px * 16 * 128); ws.g_d = take(px * 16 * 256);
  ws.g_c1 = take(px * 4 * 256); ws.g_c0 = take(px * 4 * 256);
  ws.g_b[0] = take(px * 4 * 128); ws.g_b[1] = take(px * 4 * 128); ws.g_b[2] = take(px * 4 * 128);
  family(ws.g_net, nres + 1);
  family(ws.g_t, nres);
  auto tiles = [&](int s) { return static_cast<size_t>(n) * frame_tiles_max(h * s, w * s); };
  // forward: 2 nres + 2 segments @1x, 8 @2x, 2 @4x; the backward chains need fewer @1x / @2x and 3 @4x (+ pair padding)
  ws.flag_count = static_cast<size_t>(2 * nres + 2) * tiles(1) + 8 * tiles(2) + 3 * tiles(4) + 2 * nres + 16;
  ws.flags = take(ws.flag_count * 4);
  ws.total = o;
  return ws;
}

// forward plan with one buffer per activation (nothing is overwritten).  The workspace is laid out for n_total images;
// the plan covers the n images starting at image n0 (one frame of a clip batch, tg_gen_clip_forward_train).
static std::vector<FrLayer> gen_plan_train(const std::vector<GenLayer>& L, int nres, float* out, uint8_t* wsp, int n_total,
                                           int n0, int h, int w) {
  const GenTrainWs ws = gen_train_ws(n_total, h, w, nres);
  const size_t px = static_cast<size_t>(h) * w;
  std::vector<FrLayer> P;
  int li = 0;
  // img_bytes: bytes of one image in the buffer at `off`
  auto at = [&](size_t off, size_t img_bytes) { return wsp + off + static_cast<size_t>(n0) * img_bytes; };
  auto add = [&](const void* in, void* o, const void* resid, int relu, int hh, int ww) {
    FrLayer f{};
    f.kind = L[li].kind; f.cin_pad = cin_padded(L[li].cin); f.cout_pad = cout_padded(L[li].cout);
    f.out_mode = kOutNHWCbf16; f.relu = relu; f.h = hh; f.w = ww;
    f.in = in; f.out = o; f.resid = resid; f.out2 = nullptr; f.blob_off = L[li].p_off; f.out_nstride = 0;
    P.push_back(f);
    ++li;
  };
  const size_t s1 = px * 128, s2 = px * 4 * 128, s2w = px * 4 * 256, s4w = px * 16 * 256, s4 = px * 16 * 128;
  add(at(ws.x_in, s1), at(ws.net[0], s1), nullptr, 1, h, w);
  for (int k = 0; k < nres; ++k) {
    add(at(ws.net[k], s1), at(ws.t[k], s1), nullptr, 1, h, w);
    add(at(ws.t[k], s1), at(ws.net[k + 1], s1), at(ws.net[k], s1), 0, h, w);
  }
  add(at(ws.net[nres], s1), at(ws.b0, s2), nullptr, 1, h, w);
  add(at(ws.b0, s2), at(ws.b1, s2), nullptr, 1, 2 * h, 2 * w);
  add(at(ws.b1, s2), at(ws.b2, s2), nullptr, 0, 2 * h, 2 * w);
  add(at(ws.b2, s2), at(ws.c0, s2w), nullptr, 1, 2 * h, 2 * w);
  add(at(ws.c0, s2w), at(ws.c1, s2w), nullptr, 0, 2 * h, 2 * w);
  add(at(ws.c1, s2w), at(ws.d, s4w), nullptr, 1, 2 * h, 2 * w);
  add(at(ws.d, s4w), at(ws.e, s4), nullptr, 1, 4 * h, 4 * w);
  add(at(ws.e, s4), out, nullptr, 0, 4 * h, 4 * w);
  P.back().out_mode = kOutNCHWf32Sigmoid;
  return P;
}

// packed data-gradient weights: one blob per layer except conv.0 (its input needs no gradient)
static std::vector<size_t> gen_dgrad_offsets(const std::vector<GenLayer>& L, size_t* total) {
  std::vector<size_t> off(L.size(), 0);
  size_t o = 0;
  for (size_t i = 1; i < L.size(); ++i) {
    off[i] = o;
    o += tg_packed_conv_bytes(L[i].kind == kConv3x3 ? kPackConv3x3Dgrad : kPackConvT3x3s2Dgrad, L[i].cin, L[i].cout);
  }
  if (total) *total = o;
  return off;
}

}  // namespace tg

extern "C" size_t tg_gen_train_workspace_bytes(int n, int h, int w, int num_resblock) {
  if (n <= 0 || h <= 0 || w <= 0 || num_resblock < 0 || num_resblock > 64) return 0;
  return gen_train_ws(n, h, w, num_resblock).total;
}
extern "C" size_t tg_workspace_bytes_gen_train(int n, int h, int w, int num_resblock) {
  return tg_gen_train_workspace_bytes(n, h, w, num_resblock);
}
extern "C" size_t tg_gen_packed_dgrad_bytes(int num_resblock) {
  size_t t = 0;
  gen_dgrad_offsets(gen_layers(num_resblock, nullptr, nullptr), &t);
  return t;
}
extern "C" int tg_gen_pack_dgrad(const float* flat_params, int num_resblock, void* packed_dgrad, void* stream) {
  TG_CHECK_ARG(flat_params && packed_dgrad, "gen_pack_dgrad: null pointer");
  TG_CHECK_ARG(num_resblock >= 0 && num_resblock <= 64, "gen_pack_dgrad: bad num_resblock %d", num_resblock);
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  const std::vector<size_t> off = gen_dgrad_offsets(L, nullptr);
  std::vector<PackJobSpec> jobs;
  for (size_t i = 1; i < L.size(); ++i)
    jobs.push_back(PackJobSpec{L[i].kind == kConv3x3 ? kPackConv3x3Dgrad : kPackConvT3x3s2Dgrad, flat_params + L[i].w_off, nullptr,
                               L[i].cin, L[i].cout, static_cast<uint8_t*>(packed_dgrad) + off[i]});
  return pack_weights_batched(jobs.data(), static_cast<int>(jobs.size()), static_cast<cudaStream_t>(stream));
}

extern "C" int tg_gen_forward_train(const void* packed, int num_resblock, const float* x_nchw, float* out, void* workspace,
                                    size_t workspace_bytes, int n, int h, int w, void* stream) {
  TG_CHECK_ARG(packed && x_nchw && out && workspace, "gen_forward_train: null pointer");
  TG_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "gen_forward_train: bad shape");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gen_forward_train: workspace must be 256-byte aligned");
  const GenTrainWs ws = gen_train_ws(n, h, w, num_resblock);
  if (workspace_bytes < ws.total) {
    tg_set_error("gen_forward_train: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return TG_ERR_WORKSPACE;
  }
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  int rc = tg_pack_nchw_to_nhwc64(x_nchw, wsp + ws.x_in, n, 51, h, w, stream);
  if (rc) return rc;
  const std::vector<FrLayer> P = gen_plan_train(L, num_resblock, out, wsp, n, 0, h, w);
  size_t pb = 0;
  for (auto& l : L) pb += tg_packed_conv_bytes(l.kind, l.cin, l.cout);
  return launch_frame(P.data(), static_cast<int>(P.size()), packed, pb, n, reinterpret_cast<uint32_t*>(wsp + ws.flags),
                      ws.flag_count, false, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_gen_clip_forward_train(const void* packed, int num_resblock, const float* lr, float* out, void* workspace,
                                         size_t workspace_bytes, int b, int t, int h, int w, void* stream) {
  TG_CHECK_ARG(packed && lr && out && workspace, "gen_clip_forward_train: null pointer");
  TG_CHECK_ARG(b >= 1 && t >= 1 && h >= 1 && w >= 1, "gen_clip_forward_train: bad shape");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gen_clip_forward_train: workspace must be 256-byte aligned");
  const GenTrainWs ws = gen_train_ws(b * t, h, w, num_resblock);
  if (workspace_bytes < ws.total) {
    tg_set_error("gen_clip_forward_train: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return TG_ERR_WORKSPACE;
  }
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t pb = 0;
  for (auto& l : L) pb += tg_packed_conv_bytes(l.kind, l.cin, l.cout);
  const long long lr_frame = 3LL * h * w, hr_frame = 48LL * h * w;
  for (int f = 0; f < t; ++f) {
    float* out_f = out + static_cast<long long>(f) * b * hr_frame;
    const std::vector<FrLayer> P = gen_plan_train(L, num_resblock, out_f, wsp, b * t, f * b, h, w);
    const size_t nflags = frame_flag_count(P.data(), static_cast<int>(P.size()), b);
    if (nflags > ws.flag_count) {
      tg_set_error("gen_clip_forward_train: internal flag capacity (%zu > %zu)", nflags, ws.flag_count);
      return TG_ERR_WORKSPACE;
    }
    void* x_f = wsp + ws.x_in + static_cast<size_t>(f) * b * h * w * 128;
    // frame input (flow upscale + warp + s2d + concat, code/train.py:94-107); clears the frame kernel's counters
    int rc = fused_input_launch(lr + f * lr_frame, f ? lr + (f - 1) * lr_frame : nullptr,
                                f ? out + static_cast<long long>(f - 1) * b * hr_frame : nullptr, x_f, b, h, w, lr_frame * t,
                                hr_frame, reinterpret_cast<uint32_t*>(wsp + ws.flags), nflags, st);
    if (rc) return rc;
    rc = launch_frame(P.data(), static_cast<int>(P.size()), packed, pb, b, reinterpret_cast<uint32_t*>(wsp + ws.flags),
                      ws.flag_count, true, st);
    if (rc) return rc;
  }
  return TG_OK;
}

extern "C" int tg_gen_backward(const void* packed_dgrad, int num_resblock, const float* dout, const float* out,
                               float* flat_grad, void* workspace, size_t workspace_bytes, int n, int h, int w,
                               void* stream) {
  TG_CHECK_ARG(packed_dgrad && dout && out && flat_grad && workspace, "gen_backward: null pointer");
  TG_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "gen_backward: bad shape");
  const int nres = num_resblock;
  const GenTrainWs ws = gen_train_ws(n, h, w, nres);
  if (workspace_bytes < ws.total) {
    tg_set_error("gen_backward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return TG_ERR_WORKSPACE;
  }
  auto L = gen_layers(nres, nullptr, nullptr);
  const std::vector<size_t> doff = gen_dgrad_offsets(L, nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  const uint8_t* pd = static_cast<const uint8_t*>(packed_dgrad);
  int rc;
  // weight / bias gradient of layer li given its input x and output gradient dy (sizes of the layer's INPUT)
  auto wgrad = [&](int li, const void* x, const void* dy, int hh, int ww) {
    const GenLayer& l = L[li];
    const int cp = cin_padded(l.cin), op = l.cout <= 64 ? 64 : 128;
    if (l.kind == kConv3x3)      // (the bias gradient rides on the weight-gradient launch)
      return launch_wgrad3x3(x, dy, flat_grad + l.w_off, n, hh, ww, l.cin, l.cout, cp, op, st, l.has_bias ? flat_grad + l.b_off : nullptr);
    int r = launch_wgrad_convT3x3s2(x, dy, flat_grad + l.w_off, n, hh, ww, l.cin, l.cout, cp, op, st);
    if (r || !l.has_bias) return r;
    const long long opx = static_cast<long long>(n) * hh * ww * (l.kind == kConv3x3 ? 1 : 4);
    return launch_bias_grad(dy, opx, op, l.cout, flat_grad + l.b_off, st);
  };
  // dx = dgrad(dy) (+ resid) masked by the saved post-ReLU activation; (hh, ww) = size of the layer's INPUT
  auto dgrad = [&](int li, const void* dy, const void* resid, const void* mask, void* dx, int hh, int ww) {
    const GenLayer& l = L[li];
    const int dci = l.cout <= 64 ? 64 : 128, dco = cin_padded(l.cin);
    const uint8_t* blob = pd + doff[li];
    if (l.kind == kConv3x3) {
      const float* bias = reinterpret_cast<const float*>(blob + packed_weight_bytes(dci, dco));
      return launch_conv_tc(kConv3x3, kOutNHWCbf16, dy, blob, bias, resid, dx, nullptr, n, hh, ww, dci, dco, 0, TG_AMODE_HALO, 0,
                            st, mask);
    }
    const float* bias = reinterpret_cast<const float*>(blob + packed_weight_bytes_k(kConv4x4s2, dci, dco));
    return launch_conv_tc(kConv4x4s2, kOutNHWCbf16, dy, blob, bias, nullptr, dx, nullptr, n, 2 * hh, 2 * ww, dci, dco, 0,
                          TG_AMODE_HALO, 0, st, mask);
  };
  auto B = [&](size_t off) { return static_cast<void*>(wsp + off); };
  const int iOut = 2 * nres + 8, iCt6 = iOut - 1, iCt4 = iOut - 2, iC32 = iOut - 3, iC30 = iOut - 4, iC22 = iOut - 5,
            iC20 = iOut - 6, iCt0 = iOut - 7;
  if (nres == 0) {
    // no trunk: the ConvT data gradient still has to pass conv.0's ReLU; reuse the mask path of a copy-free dgrad is not
    // possible, so this configuration is rejected (the reference default is 16 blocks)
    tg_set_error("gen_backward: num_resblock == 0 is not supported");
    return TG_ERR_BAD_ARG;
  }
  // TG_GEN_BWD_FRAME=0: one conv_tc launch per data gradient and one weight-gradient launch per layer (A/B, fallback)
  static const bool chain_on = []() { const char* e = getenv("TG_GEN_BWD_FRAME"); return !(e && e[0] == '0'); }();
  size_t dgrad_bytes = 0;
  gen_dgrad_offsets(L, &dgrad_bytes);
  // A run of 3x3 data gradients as ONE persistent launch of the frame kernel's data-gradient build (tg_frame.cu, kMask): the
  // layers are chained by tile counters instead of kernel boundaries.  chain[i] = {layer, dy, resid, mask, dx, scale of (h, w)}
  struct Link { int li; size_t dy, resid, mask, dx; int s; bool has_resid, has_mask; };
  auto run_chain = [&](const std::vector<Link>& chain) {
    if (!chain_on) {
      for (const Link& c : chain)
        if ((rc = dgrad(c.li, B(c.dy), c.has_resid ? B(c.resid) : nullptr, c.has_mask ? B(c.mask) : nullptr, B(c.dx), c.s * h, c.s * w))) return rc;
      return TG_OK;
    }
    std::vector<FrLayer> P;
    for (const Link& c : chain) {
      const GenLayer& l = L[c.li];
      FrLayer f{};
      f.kind = kConv3x3; f.cin_pad = l.cout <= 64 ? 64 : 128; f.cout_pad = cin_padded(l.cin);
      f.out_mode = kOutNHWCbf16; f.relu = 0; f.h = c.s * h; f.w = c.s * w;
      f.in = B(c.dy); f.out = B(c.dx); f.resid = c.has_resid ? B(c.resid) : nullptr; f.mask = c.has_mask ? B(c.mask) : nullptr;
      f.blob_off = doff[c.li];
      P.push_back(f);
    }
    const size_t nflags = frame_flag_count(P.data(), static_cast<int>(P.size()), n);
    if (nflags > ws.flag_count) {
      tg_set_error("gen_backward: internal flag capacity (%zu > %zu)", nflags, ws.flag_count);
      return TG_ERR_WORKSPACE;
    }
    return launch_frame(P.data(), static_cast<int>(P.size()), pd, dgrad_bytes, n, reinterpret_cast<uint32_t*>(wsp + ws.flags),
                        ws.flag_count, false, st);
  };

  // final sigmoid
  if ((rc = sigmoid_bwd_pack_launch(dout, out, B(ws.g_z), n, 16LL * h * w, 48LL * h * w, 48LL * h * w, st))) return rc;
  // output conv 64 -> 3 and conv_trans.6 (128 -> 64) @4x
  if ((rc = run_chain({{iOut, ws.g_z, 0, ws.e, ws.g_e, 4, false, true}, {iCt6, ws.g_e, 0, ws.d, ws.g_d, 4, false, true}}))) return rc;
  if ((rc = wgrad(iOut, B(ws.e), B(ws.g_z), 4 * h, 4 * w))) return rc;
  if ((rc = wgrad(iCt6, B(ws.d), B(ws.g_e), 4 * h, 4 * w))) return rc;
  // conv_trans.4: ConvT 128 -> 128, 2x -> 4x
  if ((rc = wgrad(iCt4, B(ws.c1), B(ws.g_d), 2 * h, 2 * w))) return rc;
  if ((rc = dgrad(iCt4, B(ws.g_d), nullptr, nullptr, B(ws.g_c1), 2 * h, 2 * w))) return rc;
  // conv_trans.3.2 (128 -> 128, no bias), .3.0 (64 -> 128 + ReLU; its input b2 has no activation), .2.2 (64 -> 64, no bias),
  // .2.0 (64 -> 64 + ReLU) @2x
  if ((rc = run_chain({{iC32, ws.g_c1, 0, ws.c0, ws.g_c0, 2, false, true}, {iC30, ws.g_c0, 0, 0, ws.g_b[0], 2, false, false},
                       {iC22, ws.g_b[0], 0, ws.b1, ws.g_b[1], 2, false, true}, {iC20, ws.g_b[1], 0, ws.b0, ws.g_b[2], 2, false, true}}))) return rc;
  if ((rc = wgrad(iC32, B(ws.c0), B(ws.g_c1), 2 * h, 2 * w))) return rc;
  if ((rc = wgrad(iC30, B(ws.b2), B(ws.g_c0), 2 * h, 2 * w))) return rc;
  if ((rc = wgrad(iC22, B(ws.b1), B(ws.g_b[0]), 2 * h, 2 * w))) return rc;
  if ((rc = wgrad(iC20, B(ws.b0), B(ws.g_b[1]), 2 * h, 2 * w))) return rc;
  // conv_trans.0: ConvT 64 -> 64, 1x -> 2x
  if ((rc = wgrad(iCt0, B(ws.net[nres]), B(ws.g_b[2]), h, w))) return rc;
  if ((rc = dgrad(iCt0, B(ws.g_b[2]), nullptr, nullptr, B(ws.g_net[nres]), h, w))) return rc;
  // residual trunk, last block first: net[k+1] = conv2(t[k]) + net[k], t[k] = relu(conv1(net[k]) + b):
  //   d t[k] = dgrad(conv2)(d net[k+1]) masked by t[k];  d net[k] = dgrad(conv1)(d t[k]) + d net[k+1]  (net[0] = relu(conv.0): masked)
  {
    std::vector<Link> trunk;
    for (int k = nres - 1; k >= 0; --k) {
      trunk.push_back({2 + 2 * k, ws.g_net[k + 1], 0, ws.t[k], ws.g_t[k], 1, false, true});
      trunk.push_back({1 + 2 * k, ws.g_t[k], ws.g_net[k + 1], ws.net[0], ws.g_net[k], 1, true, k == 0});
    }
    if ((rc = run_chain(trunk))) return rc;
  }
  // the trunk's weight gradients: the 16 conv1 layers (x = net[k], dy = d t[k], with bias) and the 16 conv2 layers (x = t[k],
  // dy = d net[k+1]) as two batched launches
  if (chain_on && nres >= 2) {
    const long long wstride = static_cast<long long>(L[3].w_off) - static_cast<long long>(L[1].w_off);
    if ((rc = launch_wgrad3x3_batched(B(ws.net[0]), B(ws.g_t[0]), flat_grad + L[1].w_off, flat_grad + L[1].b_off, nres, wstride, wstride,
                                      n, h, w, 64, 64, 64, st))) return rc;
    if ((rc = launch_wgrad3x3_batched(B(ws.t[0]), B(ws.g_net[1]), flat_grad + L[2].w_off, nullptr, nres, wstride, 0, n, h, w, 64, 64, 64, st)))
      return rc;
  } else {
    for (int k = nres - 1; k >= 0; --k) {
      if ((rc = wgrad(2 + 2 * k, B(ws.t[k]), B(ws.g_net[k + 1]), h, w))) return rc;
      if ((rc = wgrad(1 + 2 * k, B(ws.net[k]), B(ws.g_t[k]), h, w))) return rc;
    }
  }
  // conv.0: 51 -> 64 + ReLU (its input is detached: no data gradient)
  return wgrad(0, B(ws.x_in), B(ws.g_net[0]), h, w);
}
