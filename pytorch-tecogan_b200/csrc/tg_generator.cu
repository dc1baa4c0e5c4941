// The generator (reference code/models.py:61-86) as a sequence of tensor-core conv launches, and
// the recurrent clip loop (reference main.py:173-219) kept entirely on the device.
#include <vector>

#include "tg_conv_tc.cuh"

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
  size_t x0, a[3], b[2], c[2], d, e, total;
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
  ws.total = o;
  return ws;
}

static int gen_forward_impl(const std::vector<GenLayer>& L, const uint8_t* packed, int nres, const void* x,
                            float* out, float* logits, uint8_t* wsp, int n, int h, int w, int amode,
                            long long out_nstride, cudaStream_t st) {
  const GenWorkspace ws = gen_ws(n, h, w);
  auto blob = [&](int i) { return packed + L[i].p_off; };
  auto bias = [&](int i) {
    return reinterpret_cast<const float*>(blob(i) + packed_weight_bytes(cin_padded(L[i].cin), cout_padded(L[i].cout)));
  };
  auto conv = [&](int i, const void* in, void* o, const void* resid, int relu, int hh, int ww) {
    return launch_conv_tc(L[i].kind, kOutNHWCbf16, in, blob(i), bias(i), resid, o, nullptr, n, hh, ww,
                          cin_padded(L[i].cin), cout_padded(L[i].cout), relu, amode, 0, st);
  };
  int rc, li = 0;
  void* a[3] = {wsp + ws.a[0], wsp + ws.a[1], wsp + ws.a[2]};
  if ((rc = conv(li++, x, a[0], nullptr, 1, h, w))) return rc;               // conv.0 + ReLU
  int cur = 0;
  for (int i = 0; i < nres; ++i) {                                           // net = block(net) + net
    const int t = (cur + 1) % 3, nx = (cur + 2) % 3;
    if ((rc = conv(li++, a[cur], a[t], nullptr, 1, h, w))) return rc;
    if ((rc = conv(li++, a[t], a[nx], a[cur], 0, h, w))) return rc;
    cur = nx;
  }
  void* b0 = wsp + ws.b[0]; void* b1 = wsp + ws.b[1];
  void* c0 = wsp + ws.c[0]; void* c1 = wsp + ws.c[1];
  void* d = wsp + ws.d; void* e = wsp + ws.e;
  if ((rc = conv(li++, a[cur], b0, nullptr, 1, h, w))) return rc;            // conv_trans.0 (x2) + ReLU
  if ((rc = conv(li++, b0, b1, nullptr, 1, 2 * h, 2 * w))) return rc;        // conv_trans.2.0 + ReLU
  if ((rc = conv(li++, b1, b0, nullptr, 0, 2 * h, 2 * w))) return rc;        // conv_trans.2.2 (no skip)
  if ((rc = conv(li++, b0, c0, nullptr, 1, 2 * h, 2 * w))) return rc;        // conv_trans.3.0 + ReLU
  if ((rc = conv(li++, c0, c1, nullptr, 0, 2 * h, 2 * w))) return rc;        // conv_trans.3.2
  if ((rc = conv(li++, c1, d, nullptr, 1, 2 * h, 2 * w))) return rc;         // conv_trans.4 (x2) + ReLU
  if ((rc = conv(li++, d, e, nullptr, 1, 4 * h, 4 * w))) return rc;          // conv_trans.6 + ReLU
  return launch_conv_tc(kConv3x3, kOutNCHWf32Sigmoid, e, blob(li), bias(li), nullptr, out, logits, n, 4 * h, 4 * w,
                        64, 16, 0, amode, out_nstride, st);                  // output + sigmoid
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
  for (auto& l : L) {
    int rc = tg_pack_weights(l.kind, flat_params + l.w_off, l.has_bias ? flat_params + l.b_off : nullptr, l.cin,
                             l.cout, static_cast<uint8_t*>(packed) + l.p_off, stream);
    if (rc) return rc;
  }
  return TG_OK;
}
extern "C" size_t tg_gen_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  return gen_ws(n, h, w).total;
}

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
  return gen_forward_impl(L, static_cast<const uint8_t*>(packed), num_resblock, x_nhwc, out, logits_or_null,
                          static_cast<uint8_t*>(workspace), n, h, w, amode, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_gen_clip_forward(const void* packed, int num_resblock, const float* lr, float* out, void* workspace,
                                   size_t workspace_bytes, int n, int t, int h, int w, int amode, void* stream) {
  TG_CHECK_ARG(packed && lr && out && workspace, "gen_clip_forward: null pointer");
  TG_CHECK_ARG(n >= 1 && t >= 1 && h >= 1 && w >= 1, "gen_clip_forward: bad shape");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gen_clip_forward: workspace must be 256-byte aligned");
  const GenWorkspace ws = gen_ws(n, h, w);
  if (workspace_bytes < ws.total) {
    tg_set_error("gen_clip_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return TG_ERR_WORKSPACE;
  }
  auto L = gen_layers(num_resblock, nullptr, nullptr);
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  void* x0 = wsp + ws.x0;
  const long long lr_frame = 3LL * h * w, hr_frame = 48LL * h * w;
  const long long lr_bs = lr_frame * t, hr_bs = hr_frame * t;
  for (int f = 0; f < t; ++f) {
    const float* lr_t = lr + f * lr_frame;
    const float* lr_prev = f ? lr + (f - 1) * lr_frame : nullptr;
    const float* prev_hr = f ? out + (f - 1) * hr_frame : nullptr;
    int rc = tg_fused_warp_s2d_concat(lr_t, lr_prev, prev_hr, x0, n, h, w, lr_bs, hr_bs, stream);
    if (rc) return rc;
    rc = gen_forward_impl(L, static_cast<const uint8_t*>(packed), num_resblock, x0, out + f * hr_frame, nullptr, wsp, n,
                          h, w, amode, hr_bs, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
  }
  return TG_OK;
}
