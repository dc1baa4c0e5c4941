// The spatio-temporal discriminator forward (reference code/models.py:97-146) as a sequence of tensor-core
// conv launches (3x3 s1 and 4x4 s2), BatchNorm (batch statistics) + LeakyReLU / skip passes and the head.
// All intermediates that a backward pass needs stay in the caller's workspace.
#include <vector>

#include "tg_disc.cuh"

namespace tg {

struct DConv {
  int kind, cin, cout, has_bias;
  size_t w_off, b_off;     // element offsets in the flat parameter buffer
  size_t p_off;            // byte offset in the packed blob
};
struct DBn { int c; size_t g_off, b_off; };   // gamma / beta element offsets in the flat parameter buffer

struct DLayout {
  std::vector<DConv> convs;   // conv.0, block1, [res.0, res.2]*nb, block2, ..., block4, block5
  std::vector<DBn> bns;       // block1, res*nb, block2, res*nb, block3, res*nb, block4, block5
  size_t fc_w, fc_b, n_params, packed_bytes;
};

// state_dict / named_parameters order: conv.0.{w,b}, block1.0.w, block1.1.{w,b},
// resids1.i.0.0.{w,b}, resids1.i.0.2.w, resids1.i.1.{w,b}, block2..., block5.1.{w,b}, fc.{w,b}
static DLayout disc_layout(int nb, int ch, int fc_in) {
  DLayout L;
  size_t po = 0, bo = 0;
  auto conv = [&](int kind, int cin, int cout, int has_bias) {
    DConv c{kind, cin, cout, has_bias, po, 0, bo};
    po += static_cast<size_t>(cin) * cout * (kind == kConv4x4s2 ? 16 : 9);
    if (has_bias) { c.b_off = po; po += cout; }
    bo += tg_packed_conv_bytes(kind, cin, cout);
    L.convs.push_back(c);
  };
  auto bn = [&](int c) {
    DBn b{c, po, po + static_cast<size_t>(c)};
    po += 2 * static_cast<size_t>(c);
    L.bns.push_back(b);
  };
  auto stage = [&](int cin, int cout) {
    conv(kConv4x4s2, cin, cout, 0); bn(cout);
  };
  auto resids = [&](int c) {
    for (int i = 0; i < nb; ++i) { conv(kConv3x3, c, c, 1); conv(kConv3x3, c, c, 0); bn(c); }
  };
  conv(kConv3x3, 27, 64, 1);
  stage(64, 64); resids(64);
  stage(64, ch); resids(ch);
  stage(ch, ch); resids(ch);
  stage(ch, 64);
  stage(64, 3);
  L.fc_w = po; po += fc_in;
  L.fc_b = po; po += 1;
  L.n_params = po;
  L.packed_bytes = bo;
  return L;
}

static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// Workspace: every activation of the forward pass is kept (the backward pass reads them).
constexpr size_t kStatStride = 2 * 128 * 4;     // floats of one BatchNorm layer's statistics: [group (<= 2)][128][4]

struct DWorkspace {
  size_t x_in, a0;                    // packed input, conv.0 output
  // per BN layer: raw conv output (f32), BN(+act/skip) output as f32 (residual stream, features) and as bf16
  // (conv operand); per resblock: relu(conv1) (bf16)
  std::vector<size_t> raw, act, act16, mid;
  size_t r5, y5;                      // block5 raw conv [n,3,hw] f32, head fc input [n,3*hw] f32
  size_t stats, partial, tickets, logit;
  // backward scratch (tg_disc_backward): gradient ping-pong of the residual stream, d(raw conv output), d(relu(conv1)),
  // d(conv.0 pre-activation), d(block5 raw output), d(logit), BatchNorm reduction results
  size_t g[2], d_raw, d_mid, d_a0, d_r5, d_logit, red;
  size_t total;
};
static DWorkspace disc_ws(int n, int h, int w, int nb, int ch) {
  DWorkspace ws;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align256(bytes); return r; };
  const size_t px = static_cast<size_t>(n) * h * w;
  ws.x_in = take(px * 64 * 2);
  ws.a0 = take(px * 64 * 2);
  const int cs[4] = {64, ch, ch, 64};
  for (int s = 0; s < 4; ++s) {
    const size_t p = px >> (2 * (s + 1));
    const int reps = (s < 3) ? 1 + nb : 1;
    for (int i = 0; i < reps; ++i) {
      ws.raw.push_back(take(p * cs[s] * 4));
      ws.act.push_back(take(p * cs[s] * 4));
      ws.act16.push_back(take(p * cs[s] * 2));
      if (i > 0) ws.mid.push_back(take(p * cs[s] * 2));
    }
  }
  const size_t p5 = px >> 10;
  ws.r5 = take(p5 * 3 * 4);
  ws.y5 = take(p5 * 3 * 4);
  ws.stats = take(static_cast<size_t>(5 + 3 * nb) * kStatStride * 4);
  ws.partial = take(bn_partial_floats() * 4);
  ws.tickets = take(256);
  ws.logit = take(static_cast<size_t>(n) * 4);
  const size_t gmax = (px >> 2) * 128 * 2;        // largest gradient tensor below a0: stage 1 (64 ch) / stage 2 (128 ch @ 1/16)
  ws.g[0] = take(gmax); ws.g[1] = take(gmax); ws.d_raw = take(gmax); ws.d_mid = take(gmax);
  ws.d_a0 = take(px * 64 * 2);
  ws.d_r5 = take(p5 * 64 * 2);
  ws.d_logit = take(static_cast<size_t>(n) * 4);
  ws.red = take(2 * 2 * 128 * 4);
  ws.total = o;
  return ws;
}

}  // namespace tg

using namespace tg;

static int check_cfg(int nb, int ch) {
  TG_CHECK_ARG(nb >= 0 && nb <= 16, "discriminator: discrim_resblocks %d out of range", nb);
  TG_CHECK_ARG(ch == 64 || ch == 128, "discriminator: discrim_channels must be 64 or 128 (got %d)", ch);
  return TG_OK;
}

extern "C" size_t tg_disc_param_count(int nb, int ch, int fc_in) {
  if (check_cfg(nb, ch)) return 0;
  return disc_layout(nb, ch, fc_in).n_params;
}
extern "C" size_t tg_disc_packed_bytes(int nb, int ch) {
  if (check_cfg(nb, ch)) return 0;
  return disc_layout(nb, ch, 48).packed_bytes;
}
extern "C" int tg_disc_pack(const float* flat_params, int nb, int ch, void* packed, void* stream) {
  TG_CHECK_ARG(flat_params && packed, "disc_pack: null pointer");
  if (int rc = check_cfg(nb, ch)) return rc;
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 255) == 0, "disc_pack: packed must be 256-byte aligned");
  const DLayout L = disc_layout(nb, ch, 48);
  std::vector<PackJobSpec> jobs;
  for (auto& c : L.convs)
    jobs.push_back(PackJobSpec{c.kind, flat_params + c.w_off, c.has_bias ? flat_params + c.b_off : nullptr, c.cin, c.cout,
                               static_cast<uint8_t*>(packed) + c.p_off});
  return pack_weights_batched(jobs.data(), static_cast<int>(jobs.size()), static_cast<cudaStream_t>(stream));
}
extern "C" size_t tg_disc_workspace_bytes(int n, int h, int w, int nb, int ch) {
  if (n <= 0 || h <= 0 || w <= 0 || check_cfg(nb, ch)) return 0;
  return disc_ws(n, h, w, nb, ch).total;
}

extern "C" size_t tg_workspace_bytes_disc(int n, int h, int w, int nb, int ch) { return tg_disc_workspace_bytes(n, h, w, nb, ch); }

extern "C" int tg_disc_forward(const float* flat_params, const void* packed, int nb, int ch, int fc_in, const float* x,
                               float* prob, float* const* feats, void* const* bn_running, int training,
                               void* workspace, size_t workspace_bytes, int n, int h, int w, void* stream) {
  return tg_disc_forward_groups(flat_params, packed, nb, ch, fc_in, x, prob, feats, bn_running, training, workspace,
                                workspace_bytes, n, 1, h, w, stream);
}

extern "C" int tg_disc_forward_groups(const float* flat_params, const void* packed, int nb, int ch, int fc_in, const float* x,
                                      float* prob, float* const* feats, void* const* bn_running, int training,
                                      void* workspace, size_t workspace_bytes, int n, int groups, int h, int w, void* stream) {
  TG_CHECK_ARG(flat_params && packed && x && prob && workspace, "disc_forward: null pointer");
  TG_CHECK_ARG(groups >= 1 && groups <= 2 && n % groups == 0, "disc_forward: n = %d samples do not split into %d groups", n, groups);
  if (int rc = check_cfg(nb, ch)) return rc;
  TG_CHECK_ARG(n >= 1 && h >= 32 && w >= 32 && (h % 32) == 0 && (w % 32) == 0,
               "disc_forward: input must be [n,27,h,w] with h, w multiples of 32 (got %dx%d)", h, w);
  TG_CHECK_ARG(fc_in == 3 * (h / 32) * (w / 32), "disc_forward: fc in-features %d do not match 3*(h/32)*(w/32) = %d "
               "(code/models.py:123 hard-codes 48 = 128x128 inputs)", fc_in, 3 * (h / 32) * (w / 32));
  TG_CHECK_ARG(training || bn_running, "disc_forward: eval mode needs the running statistics");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "disc_forward: workspace must be 256-byte aligned");
  const DWorkspace ws = disc_ws(n, h, w, nb, ch);
  if (workspace_bytes < ws.total) {
    tg_set_error("disc_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return TG_ERR_WORKSPACE;
  }
  const DLayout L = disc_layout(nb, ch, fc_in);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  float* stats_all = reinterpret_cast<float*>(wsp + ws.stats);
  float* partial = reinterpret_cast<float*>(wsp + ws.partial);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(wsp + ws.tickets);
  TG_CUDA(cudaMemsetAsync(tickets, 0, 256, st));

  int ci = 0, bi = 0;
  // raw = pre-BatchNorm output, kept in f32
  auto conv = [&](const void* in, void* out, int hh, int ww, int act, bool raw) {
    const DConv& c = L.convs[ci++];
    const int cp = cin_padded(c.cin), op = cout_padded(c.cout);
    const uint8_t* blob = pk + c.p_off;
    const float* bias = reinterpret_cast<const float*>(blob + packed_weight_bytes_k(c.kind, cp, op));
    const int mode = c.cout == 3 ? kOutNCHWf32Raw : (raw ? kOutNHWCf32 : kOutNHWCbf16);
    return launch_conv_tc(c.kind, mode, in, blob, bias, nullptr, out, nullptr, n, hh, ww, cp, op, act, TG_AMODE_HALO, 0, st);
  };
  // BatchNorm (+ LeakyReLU or + skip) of the raw conv output of BN layer `bi`
  auto bn = [&](const void* raw, const void* skip, void* out32, void* out16, long long pixels, int act) {
    const DBn& b = L.bns[bi];
    float* stats = stats_all + static_cast<size_t>(bi) * kStatStride;
    float* rm = bn_running ? static_cast<float*>(bn_running[3 * bi + 0]) : nullptr;
    float* rv = bn_running ? static_cast<float*>(bn_running[3 * bi + 1]) : nullptr;
    long long* nbt = bn_running ? static_cast<long long*>(bn_running[3 * bi + 2]) : nullptr;
    int rc = TG_OK;
    // `pixels` is the whole batch; the statistics are per group (= per forward pass of the reference)
    if (training) rc = bn_stats_launch(raw, pixels / groups, b.c, flat_params + b.g_off, flat_params + b.b_off, partial, tickets, stats,
                                       rm, rv, nbt, st, groups);
    else for (int g = 0; g < groups && !rc; ++g)
      rc = bn_fold_running_launch(b.c, flat_params + b.g_off, flat_params + b.b_off, rm, rv, stats + g * 512, st);
    if (rc) return rc;
    ++bi;
    return bn_apply_launch(raw, skip, out32, out16, pixels / groups, b.c, stats, act, st, groups);
  };

  int rc;
  if ((rc = tg_pack_nchw_to_nhwc64(x, wsp + ws.x_in, n, 27, h, w, stream))) return rc;
  if ((rc = conv(wsp + ws.x_in, wsp + ws.a0, h, w, kActLrelu02, false))) return rc;      // conv + lrelu  (:102,127)
  const void* cur16 = wsp + ws.a0;     // bf16 copy: conv operand
  const void* cur32 = nullptr;         // f32 copy: residual stream
  int hh = h, ww = w, li = 0, mi = 0;
  const int cs[4] = {64, ch, ch, 64};
  for (int s = 0; s < 4; ++s) {
    // discriminator_block: conv k4 s2 (no bias) -> BN -> LeakyReLU   (:90-94)
    if ((rc = conv(cur16, wsp + ws.raw[li], hh, ww, kActNone, true))) return rc;
    hh /= 2; ww /= 2;
    const long long pixels = static_cast<long long>(n) * hh * ww;
    if ((rc = bn(wsp + ws.raw[li], nullptr, wsp + ws.act[li], wsp + ws.act16[li], pixels, kActLrelu02))) return rc;
    cur32 = wsp + ws.act[li]; cur16 = wsp + ws.act16[li];
    ++li;
    if (s < 3) {
      for (int i = 0; i < nb; ++i) {                                                    // net = block(net) + net  (:128-130)
        if ((rc = conv(cur16, wsp + ws.mid[mi], hh, ww, kActRelu, false))) return rc;   // conv + bias + ReLU
        if ((rc = conv(wsp + ws.mid[mi], wsp + ws.raw[li], hh, ww, kActNone, true))) return rc;   // conv (no bias)
        if ((rc = bn(wsp + ws.raw[li], cur32, wsp + ws.act[li], wsp + ws.act16[li], pixels, kActNone))) return rc;   // BN, + skip
        cur32 = wsp + ws.act[li]; cur16 = wsp + ws.act16[li];
        ++li; ++mi;
      }
    }
    if (feats && feats[s]) {                                                            // layer_list (:131,135,139,141)
      if ((rc = nhwc_to_nchw_f32_launch(cur32, feats[s], n, cs[s], static_cast<long long>(hh) * ww, st))) return rc;
    }
  }
  // block5 conv (64 -> 3) raw, then BN + LeakyReLU + flatten + fc + sigmoid in the head kernel  (:121,141-145)
  if ((rc = conv(cur16, wsp + ws.r5, hh, ww, kActNone, false))) return rc;
  hh /= 2; ww /= 2;
  {
    const DBn& b = L.bns[bi];
    float* rm = bn_running ? static_cast<float*>(bn_running[3 * bi + 0]) : nullptr;
    float* rv = bn_running ? static_cast<float*>(bn_running[3 * bi + 1]) : nullptr;
    long long* nbt = bn_running ? static_cast<long long*>(bn_running[3 * bi + 2]) : nullptr;
    rc = disc_head_launch(reinterpret_cast<const float*>(wsp + ws.r5), n / groups, hh * ww, flat_params + b.g_off, flat_params + b.b_off,
                          training, rm, rv, nbt, flat_params + L.fc_w, flat_params + L.fc_b,
                          reinterpret_cast<float*>(wsp + ws.y5), stats_all + static_cast<size_t>(bi) * kStatStride,
                          reinterpret_cast<float*>(wsp + ws.logit), prob, st, groups);
  }
  return rc;
}

// =====================================================================================================
// Backward (reference code/train.py:340: scaler.scale(discrim_loss).backward() through discriminator.forward).
// The reference detaches the discriminator's inputs (code/train.py:181,199) and its layer features
// (code/train.py:214), so the gradient enters through `prob` only and stops at conv.0's weights.
// =====================================================================================================
namespace tg {
// packed data-gradient convolutions: one blob per conv except conv.0
static std::vector<size_t> disc_dgrad_offsets(const DLayout& L, size_t* total) {
  std::vector<size_t> off(L.convs.size(), 0);
  size_t o = 0;
  for (size_t i = 1; i < L.convs.size(); ++i) {
    off[i] = o;
    o += tg_packed_conv_bytes(L.convs[i].kind == kConv3x3 ? kPackConv3x3Dgrad : kPackConv4x4s2Dgrad, L.convs[i].cin,
                              L.convs[i].cout);
  }
  if (total) *total = o;
  return off;
}
}  // namespace tg

extern "C" size_t tg_disc_packed_dgrad_bytes(int nb, int ch) {
  if (check_cfg(nb, ch)) return 0;
  size_t t = 0;
  disc_dgrad_offsets(disc_layout(nb, ch, 48), &t);
  return t;
}

extern "C" int tg_disc_pack_dgrad(const float* flat_params, int nb, int ch, void* packed_dgrad, void* stream) {
  TG_CHECK_ARG(flat_params && packed_dgrad, "disc_pack_dgrad: null pointer");
  if (int rc = check_cfg(nb, ch)) return rc;
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(packed_dgrad) & 255) == 0, "disc_pack_dgrad: buffer must be 256-byte aligned");
  const DLayout L = disc_layout(nb, ch, 48);
  const std::vector<size_t> off = disc_dgrad_offsets(L, nullptr);
  std::vector<PackJobSpec> jobs;
  for (size_t i = 1; i < L.convs.size(); ++i) {
    const DConv& c = L.convs[i];
    jobs.push_back(PackJobSpec{c.kind == kConv3x3 ? kPackConv3x3Dgrad : kPackConv4x4s2Dgrad, flat_params + c.w_off, nullptr, c.cin,
                               c.cout, static_cast<uint8_t*>(packed_dgrad) + off[i]});
  }
  return pack_weights_batched(jobs.data(), static_cast<int>(jobs.size()), static_cast<cudaStream_t>(stream));
}

extern "C" int tg_disc_backward(const float* flat_params, const void* packed_dgrad, int nb, int ch, int fc_in,
                                const float* dprob, const float* prob, float* flat_grad, void* workspace,
                                size_t workspace_bytes, int n, int h, int w, void* stream) {
  return tg_disc_backward_groups(flat_params, packed_dgrad, nb, ch, fc_in, dprob, prob, flat_grad, workspace, workspace_bytes, n, 1,
                                 h, w, stream);
}

extern "C" int tg_disc_backward_groups(const float* flat_params, const void* packed_dgrad, int nb, int ch, int fc_in,
                                       const float* dprob, const float* prob, float* flat_grad, void* workspace,
                                       size_t workspace_bytes, int n, int groups, int h, int w, void* stream) {
  TG_CHECK_ARG(flat_params && packed_dgrad && dprob && prob && flat_grad && workspace, "disc_backward: null pointer");
  TG_CHECK_ARG(groups >= 1 && groups <= 2 && n % groups == 0, "disc_backward: n = %d samples do not split into %d groups", n, groups);
  if (int rc = check_cfg(nb, ch)) return rc;
  TG_CHECK_ARG(n >= 1 && h >= 32 && w >= 32 && (h % 32) == 0 && (w % 32) == 0, "disc_backward: bad shape");
  TG_CHECK_ARG(fc_in == 3 * (h / 32) * (w / 32), "disc_backward: fc in-features do not match the input size");
  const DWorkspace ws = disc_ws(n, h, w, nb, ch);
  if (workspace_bytes < ws.total) {
    tg_set_error("disc_backward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return TG_ERR_WORKSPACE;
  }
  const DLayout L = disc_layout(nb, ch, fc_in);
  const std::vector<size_t> doff = disc_dgrad_offsets(L, nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  const uint8_t* pd = static_cast<const uint8_t*>(packed_dgrad);
  float* stats_all = reinterpret_cast<float*>(wsp + ws.stats);
  float* partial = reinterpret_cast<float*>(wsp + ws.partial);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(wsp + ws.tickets);
  float* red = reinterpret_cast<float*>(wsp + ws.red);
  auto B = [&](size_t off) { return static_cast<void*>(wsp + off); };
  int rc;

  // weight (+ bias) gradient of conv `ci`; (hh, ww) = the conv's OUTPUT size
  auto wgrad = [&](int ci, const void* x, const void* dy, int hh, int ww) {
    const DConv& c = L.convs[ci];
    const int cp = cin_padded(c.cin), op = c.cout <= 64 ? 64 : 128;
    if (c.kind == kConv3x3)      // (the bias gradient rides on the weight-gradient launch)
      return launch_wgrad3x3(x, dy, flat_grad + c.w_off, n, hh, ww, c.cin, c.cout, cp, op, st, c.has_bias ? flat_grad + c.b_off : nullptr);
    int r = launch_wgrad_conv4x4s2(x, dy, flat_grad + c.w_off, n, hh, ww, c.cin, c.cout, cp, op, st);
    if (r || !c.has_bias) return r;
    return launch_bias_grad(dy, static_cast<long long>(n) * hh * ww, op, c.cout, flat_grad + c.b_off, st);
  };
  // data gradient of conv `ci`; (hh, ww) = the conv's OUTPUT size; dx has the conv's input size
  auto dgrad = [&](int ci, const void* dy, const void* resid, const void* mask, int mask_mode, void* dx, int hh, int ww) {
    const DConv& c = L.convs[ci];
    if (c.kind == kConv3x3)
      return tg_conv3x3_dgrad(dy, pd + doff[ci], resid, mask, dx, n, hh, ww, c.cin, c.cout, stream);
    return tg_conv4x4s2_dgrad(dy, pd + doff[ci], mask, mask_mode, dx, n, hh, ww, c.cin, c.cout, stream);
  };
  auto bn_bwd = [&](int bi, const void* g_out, const void* raw, const void* act, void* dx, long long pixels) {
    const DBn& b = L.bns[bi];
    return bn_bwd_launch(g_out, raw, act, dx, pixels / groups, b.c, stats_all + static_cast<size_t>(bi) * kStatStride, partial, tickets, red,
                         flat_grad + b.g_off, flat_grad + b.b_off, st, groups);
  };

  const int n_conv = static_cast<int>(L.convs.size()), n_bn = static_cast<int>(L.bns.size());
  // head: dprob -> sigmoid -> fc -> LeakyReLU -> BatchNorm(3) -> d(block5 conv output)
  {
    const DBn& b = L.bns[n_bn - 1];
    const int hw5 = (h / 32) * (w / 32);
    rc = disc_head_bwd_launch(dprob, prob, reinterpret_cast<const float*>(wsp + ws.y5), reinterpret_cast<const float*>(wsp + ws.r5),
                              n / groups, hw5, stats_all + static_cast<size_t>(n_bn - 1) * kStatStride, flat_params + L.fc_w,
                              flat_grad + L.fc_w, flat_grad + L.fc_b, flat_grad + b.g_off, flat_grad + b.b_off,
                              reinterpret_cast<float*>(wsp + ws.d_logit), B(ws.d_r5), st, groups);
    if (rc) return rc;
  }
  // walk the forward structure backwards.  Indices at the END of the forward pass:
  int ci = n_conv - 1;                 // block5 conv
  int bi = n_bn - 2;                   // block4 BN
  int li = static_cast<int>(ws.raw.size()) - 1, mi = static_cast<int>(ws.mid.size()) - 1;
  int cur = 0;
  // block5 conv (64 -> 3, k4 s2): input = block4 output act16[li]
  {
    const int hh = h / 32, ww = w / 32;
    if ((rc = wgrad(ci, B(ws.act16[li]), B(ws.d_r5), hh, ww))) return rc;
    if ((rc = dgrad(ci, B(ws.d_r5), nullptr, nullptr, 0, B(ws.g[cur]), hh, ww))) return rc;
    --ci;
  }
  for (int s = 3; s >= 0; --s) {
    const int hh = h >> (s + 1), ww = w >> (s + 1);
    const long long pixels = static_cast<long long>(n) * hh * ww;
    if (s < 3) {
      for (int i = nb - 1; i >= 0; --i) {
        // act[li] = BN(conv2(mid[mi])) + act[li-1];  mid[mi] = relu(conv1(act16[li-1]) + b)
        if ((rc = bn_bwd(bi, B(ws.g[cur]), B(ws.raw[li]), nullptr, B(ws.d_raw), pixels))) return rc;
        --bi;
        if ((rc = wgrad(ci, B(ws.mid[mi]), B(ws.d_raw), hh, ww))) return rc;
        if ((rc = dgrad(ci, B(ws.d_raw), nullptr, B(ws.mid[mi]), kMaskRelu, B(ws.d_mid), hh, ww))) return rc;
        --ci;
        if ((rc = wgrad(ci, B(ws.act16[li - 1]), B(ws.d_mid), hh, ww))) return rc;
        if ((rc = dgrad(ci, B(ws.d_mid), B(ws.g[cur]), nullptr, 0, B(ws.g[cur ^ 1]), hh, ww))) return rc;
        --ci;
        cur ^= 1; --li; --mi;
      }
    }
    // discriminator_block: act[li] = lrelu(BN(conv4x4s2(input)))
    if ((rc = bn_bwd(bi, B(ws.g[cur]), B(ws.raw[li]), B(ws.act[li]), B(ws.d_raw), pixels))) return rc;
    --bi;
    const void* input = (s == 0) ? B(ws.a0) : B(ws.act16[li - 1]);
    if ((rc = wgrad(ci, input, B(ws.d_raw), hh, ww))) return rc;
    if (s == 0) {
      // input = a0 = lrelu(conv.0(x) + b): the LeakyReLU backward rides on the data gradient's epilogue
      if ((rc = dgrad(ci, B(ws.d_raw), nullptr, B(ws.a0), kMaskLrelu02, B(ws.d_a0), hh, ww))) return rc;
    } else {
      if ((rc = dgrad(ci, B(ws.d_raw), nullptr, nullptr, 0, B(ws.g[cur ^ 1]), hh, ww))) return rc;
      cur ^= 1;
    }
    --ci; --li;
  }
  // conv.0 (27 -> 64, bias): its input is detached in the reference -> weight / bias gradient only
  return wgrad(0, B(ws.x_in), B(ws.d_a0), h, w);
}
