// Error reporting, device checks and the public conv / pack entry points of libtecogan_b200.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "tg_conv_tc.cuh"

static thread_local char g_err[512] = "";

void tg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Per-device read-only configuration ("no global mutable state except a per-device read-only config"):
// the SM count of the CURRENT device, cached per ordinal so that one process can drive several GPUs.
int tg_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TG_MAX_DEVICES) dev = 0;
  return dev;
}
int tg_num_sms() {
  static std::atomic<int> sms[TG_MAX_DEVICES];
  const int dev = tg_current_device();
  int v = sms[dev].load(std::memory_order_relaxed);
  if (!v) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// ------------------------------------------------------------------- launch accounting / profiling
// The launch counter is a relaxed atomic (any thread, any stream).  Per-launch event timing is a MEASUREMENT mode: the
// list is guarded by a mutex, but tg_prof_pre / tg_prof_post of one launch are paired by position, so while it is on
// all launches must come from one host thread (include/tecogan_b200.h says so); off, the hooks cost one atomic load.
struct ProfEntry { cudaEvent_t a, b; int kid; double work; };
static std::atomic<long long> g_launches{0};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof;

void tg_prof_pre(int kernel_id, double work, cudaStream_t stream) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_acquire)) return;
  ProfEntry e{nullptr, nullptr, kernel_id, work};
  cudaEventCreate(&e.a);
  cudaEventCreate(&e.b);
  cudaEventRecord(e.a, stream);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(e);
}
void tg_prof_post(cudaStream_t stream) {
  if (!g_prof_on.load(std::memory_order_acquire)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, stream);
}

extern "C" long long tg_launch_count(void) { return g_launches.load(); }
extern "C" int tg_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& e : g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  g_prof.clear();
  g_prof_on.store(true, std::memory_order_release);
  return TG_OK;
}
extern "C" int tg_profile_end(int max_entries, int* kernel_ids, float* ms, double* work) {
  g_prof_on.store(false, std::memory_order_release);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto& e : g_prof) {
    float t = 0.f;
    cudaEventSynchronize(e.b);
    cudaEventElapsedTime(&t, e.a, e.b);
    if (n < max_entries) { kernel_ids[n] = e.kid; ms[n] = t; work[n] = e.work; ++n; }
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  g_prof.clear();
  return n;
}

extern "C" const char* tg_last_error_string(void) { return g_err; }
extern "C" int tg_version(void) { return 100; }

extern "C" int tg_check_device(void) {
  int dev = 0, major = 0;
  TG_CUDA(cudaGetDevice(&dev));
  TG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    tg_set_error("libtecogan_b200 is built for sm_100a only; device has compute capability %d.x", major);
    return TG_ERR_ARCH;
  }
  return TG_OK;
}

// ------------------------------------------------------------------------------ weight packing
// Packed blob = bf16 blocks [cout_chunk][k_chunk][tap][NT rows (out ch)][64 (in ch)] in MMA issue
// order, followed by the f32 bias padded to cout_pad.  TMA applies the 128B swizzle on load, so
// the global image is plain row-major.
namespace tg {

__device__ __forceinline__ void pack_weights_body(int kind, const float* __restrict__ w, const float* __restrict__ bias,
                                                  int cin, int cout, int cin_pad, int cout_pad, int nt,
                                                  __nv_bfloat16* __restrict__ dst, float* __restrict__ bias_dst) {
  // conv taps in (ky,kx) raster order; transposed conv grouped by INPUT SHIFT, see tg_conv_tc.cuh (kCtKy / kCtKx)
  const int ct_ky[9] = {1, 1, 2, 2, 1, 2, 0, 0, 0};     // == kCtKy / kCtKx (device-side copy)
  const int ct_kx[9] = {1, 2, 2, 1, 0, 0, 2, 1, 0};
  const int kchunks = cin_pad / 64;
  const int ntap = (kind == kConv4x4s2 || kind == kPackConvT3x3s2Dgrad || kind == kPackConv4x4s2Dgrad) ? 16 : 9;
  const long long total = static_cast<long long>(ntap) * cin_pad * cout_pad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci_l = static_cast<int>(i % 64);
    long long r = i / 64;
    const int co_l = static_cast<int>(r % nt); r /= nt;
    int j, kc, chunk;
    if (kind == kPackConv4x4s2Dgrad) {
      // four independent per-phase blobs: [phase][chunk][kc][tap][co][ci]
      const int t = static_cast<int>(r % 4); r /= 4;
      kc = static_cast<int>(r % kchunks); r /= kchunks;
      const int nchunk = cout_pad / nt;
      chunk = static_cast<int>(r % nchunk); r /= nchunk;
      j = static_cast<int>(r) * 4 + t;
    } else {
      j = static_cast<int>(r % ntap); r /= ntap;
      kc = static_cast<int>(r % kchunks); r /= kchunks;
      chunk = static_cast<int>(r);
    }
    const int ci = kc * 64 + ci_l, co = chunk * nt + co_l;
    float v = 0.f;
    if (ci < cin && co < cout) {
      if (kind == kConv3x3) {
        const int ky = j / 3, kx = j % 3;
        v = w[((static_cast<long long>(co) * cin + ci) * 3 + ky) * 3 + kx];
      } else if (kind == kConv4x4s2) {
        // j = phase * 4 + tap: phase = (py,px) input parity, tap = (dy,dx) offset inside the phase box
        const int ph = j >> 2, t = j & 3;
        const int ky = 2 * (t >> 1) + (ph >> 1), kx = 2 * (t & 1) + (ph & 1);
        v = w[((static_cast<long long>(co) * cin + ci) * 4 + ky) * 4 + kx];
      } else if (kind == kPackConv3x3Dgrad) {
        // dX = conv3x3(dY, W'), W'[ci_f][co_f][ky][kx] = W[co_f][ci_f][2-ky][2-kx]; here ci = co_f, co = ci_f, cout = cin_f
        const int ky = j / 3, kx = j % 3;
        v = w[((static_cast<long long>(ci) * cout + co) * 3 + (2 - ky)) * 3 + (2 - kx)];
      } else if (kind == kPackConv4x4s2Dgrad) {
        // adjoint of Conv2d(k4,s2,p1), output phase (py,px) = parity of the dX pixel: dX[2i+py] takes dY rows
        // {i-1 (ky=3), i (ky=1)} for py = 0 and {i (ky=2), i+1 (ky=0)} for py = 1; here ci = co_f, co = ci_f, cout = cin_f
        const int ph = j >> 2, t = j & 3;
        const int ty = t >> 1, tx = t & 1, py = ph >> 1, px = ph & 1;
        const int ky = py ? (ty ? 0 : 2) : (ty ? 1 : 3), kx = px ? (tx ? 0 : 2) : (tx ? 1 : 3);
        v = w[((static_cast<long long>(ci) * cout + co) * 4 + ky) * 4 + kx];
      } else if (kind == kPackConvT3x3s2Dgrad) {
        // dX[ci_f] = conv k3 s2 p1 (dY[co_f], Wt[ci_f][co_f]) == conv k4 s2 p1 with a zero 4th row / column;
        // here ci = co_f, co = ci_f, cin = cout_f.  Tap order as kConv4x4s2.
        const int ph = j >> 2, t = j & 3;
        const int ky = 2 * (t >> 1) + (ph >> 1), kx = 2 * (t & 1) + (ph & 1);
        v = (ky < 3 && kx < 3) ? w[((static_cast<long long>(co) * cin + ci) * 3 + ky) * 3 + kx] : 0.f;
      } else {
        v = w[((static_cast<long long>(ci) * cout + co) * 3 + ct_ky[j]) * 3 + ct_kx[j]];
      }
    }
    dst[i] = __float2bfloat16_rn(v);
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < cout_pad; c += blockDim.x) bias_dst[c] = (bias && c < cout) ? bias[c] : 0.f;
}

__global__ void pack_weights_kernel(int kind, const float* __restrict__ w, const float* __restrict__ bias,
                                    int cin, int cout, int cin_pad, int cout_pad, int nt,
                                    __nv_bfloat16* __restrict__ dst, float* __restrict__ bias_dst) {
  pack_weights_body(kind, w, bias, cin, cout, cin_pad, cout_pad, nt, dst, bias_dst);
}

// All layers of a network in one launch (blockIdx.y = layer): after every optimizer step the training loop re-packs
// ~70 + ~70 small tensors, one launch each was 5 % of a cfg4 step.
__global__ void pack_weights_batched_kernel(const __grid_constant__ PackJobs jobs) {
  const PackJob& j = jobs.j[blockIdx.y];
  pack_weights_body(j.kind, j.w, j.bias, j.cin, j.cout, j.cin_pad, j.cout_pad, j.nt, static_cast<__nv_bfloat16*>(j.dst), j.bias_dst);
}

}  // namespace tg

// kinds 3 / 4 derive the data-gradient convolution of a forward layer: its input channels are the forward
// layer's output channels and vice versa
static void derived_channels(int kind, int cin, int cout, int* dcin, int* dcout, int* launch_kind) {
  const bool dgrad = (kind == tg::kPackConv3x3Dgrad || kind == tg::kPackConvT3x3s2Dgrad || kind == tg::kPackConv4x4s2Dgrad);
  *dcin = dgrad ? cout : cin;
  *dcout = dgrad ? cin : cout;
  *launch_kind = kind == tg::kPackConv3x3Dgrad ? tg::kConv3x3
               : (kind == tg::kPackConvT3x3s2Dgrad || kind == tg::kPackConv4x4s2Dgrad) ? tg::kConv4x4s2 : kind;   // 16-tap size
}

extern "C" size_t tg_packed_conv_bytes(int kind, int cin, int cout) {
  int dci, dco, lk;
  derived_channels(kind, cin, cout, &dci, &dco, &lk);
  const int cp = tg::cin_padded(dci), op = tg::cout_padded(dco);
  size_t b = tg::packed_weight_bytes_k(lk, cp, op) + static_cast<size_t>(op) * 4;
  return (b + 255) & ~static_cast<size_t>(255);
}

// Fills one PackJob (validation + derived channel counts); shared by the single and the batched entry points.
static int make_pack_job(int kind, const float* weight, const float* bias, int cin, int cout, void* packed, tg::PackJob* job) {
  TG_CHECK_ARG(weight && packed, "pack_weights: null pointer");
  TG_CHECK_ARG(kind >= 0 && kind <= 5, "pack_weights: kind must be 0 (conv3x3), 1 (convT3x3s2), 2 (conv4x4s2), "
               "3 (dgrad of conv3x3), 4 (dgrad of convT3x3s2) or 5 (dgrad of conv4x4s2)");
  TG_CHECK_ARG(cin >= 1 && cin <= 128 && cout >= 1 && cout <= 128, "pack_weights: channels out of range");
  TG_CHECK_ARG(!(kind >= 3 && bias), "pack_weights: data-gradient convolutions have no bias");
  int lk;
  { int a, b; derived_channels(kind, cin, cout, &a, &b, &lk); cin = a; cout = b; }   // from here: the derived conv's channels
  const int cp = tg::cin_padded(cin), op = tg::cout_padded(cout);
  job->kind = kind; job->w = weight; job->bias = bias; job->cin = cin; job->cout = cout; job->cin_pad = cp; job->cout_pad = op;
  job->nt = op == 16 ? 16 : 64;
  job->dst = packed;
  job->bias_dst = reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + tg::packed_weight_bytes_k(lk, cp, op));
  return TG_OK;
}

int tg::pack_weights_batched(const PackJobSpec* specs, int n, cudaStream_t stream) {
  static thread_local PackJobs jobs;
  for (int i0 = 0; i0 < n; i0 += kPackBatch) {
    const int m = n - i0 < kPackBatch ? n - i0 : kPackBatch;
    for (int i = 0; i < m; ++i) {
      const PackJobSpec& sp = specs[i0 + i];
      if (int rc = make_pack_job(sp.kind, sp.weight, sp.bias, sp.cin, sp.cout, sp.packed, &jobs.j[i])) return rc;
    }
    tg_prof_pre(TG_K_PACK, 0.0, stream);
    pack_weights_batched_kernel<<<dim3(64, m), 256, 0, stream>>>(jobs);
    tg_prof_post(stream);
    TG_CUDA(cudaGetLastError());
  }
  return TG_OK;
}

extern "C" int tg_pack_weights(int kind, const float* weight, const float* bias, int cin, int cout,
                               void* packed, void* stream) {
  TG_CHECK_ARG(weight && packed, "pack_weights: null pointer");
  TG_CHECK_ARG(kind >= 0 && kind <= 5, "pack_weights: kind must be 0 (conv3x3), 1 (convT3x3s2), 2 (conv4x4s2), "
               "3 (dgrad of conv3x3), 4 (dgrad of convT3x3s2) or 5 (dgrad of conv4x4s2)");
  TG_CHECK_ARG(cin >= 1 && cin <= 128 && cout >= 1 && cout <= 128, "pack_weights: channels out of range");
  TG_CHECK_ARG(!(kind >= 3 && bias), "pack_weights: data-gradient convolutions have no bias");
  int lk;
  { int a, b; derived_channels(kind, cin, cout, &a, &b, &lk); cin = a; cout = b; }   // from here: the derived conv's channels
  const int cp = tg::cin_padded(cin), op = tg::cout_padded(cout);
  const int nt = op == 16 ? 16 : 64;
  auto* dst = static_cast<__nv_bfloat16*>(packed);
  auto* bdst = reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + tg::packed_weight_bytes_k(lk, cp, op));
  tg_prof_pre(TG_K_PACK, 0.0, static_cast<cudaStream_t>(stream));
  tg::pack_weights_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(kind, weight, bias, cin, cout, cp, op, nt, dst, bdst);
  tg_prof_post(static_cast<cudaStream_t>(stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

static const float* packed_bias(const void* packed, int cin_pad, int cout_pad) {
  return reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed) + tg::packed_weight_bytes(cin_pad, cout_pad));
}

extern "C" int tg_conv4x4s2_fwd(const void* x, const void* packed, void* y, int n, int h, int w, int cin_pad,
                                int cout, int act, void* stream) {
  TG_CHECK_ARG(cout == 3 || cout == 64 || cout == 128, "conv4x4s2_fwd: cout must be 3, 64 or 128 (got %d)", cout);
  const int op = tg::cout_padded(cout);
  const float* bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed) +
                                                     tg::packed_weight_bytes_k(tg::kConv4x4s2, cin_pad, op));
  TG_CHECK_ARG(!(cout == 3 && act != 0), "conv4x4s2_fwd: the 3-channel f32 output is raw (act must be 0)");
  return tg::launch_conv_tc(tg::kConv4x4s2, cout == 3 ? tg::kOutNCHWf32Raw : tg::kOutNHWCbf16, x, packed, bias, nullptr, y,
                            nullptr, n, h, w, cin_pad, op, act, TG_AMODE_HALO, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_conv3x3_fwd(const void* x, const void* packed, const void* residual, void* y, int n, int h,
                              int w, int cin_pad, int cout, int relu, int amode, void* stream) {
  TG_CHECK_ARG(cout == 64 || cout == 128, "conv3x3_fwd: cout must be 64 or 128 (got %d)", cout);
  TG_CHECK_ARG(!(relu && residual), "conv3x3_fwd: activation and residual are mutually exclusive");
  return tg::launch_conv_tc(tg::kConv3x3, tg::kOutNHWCbf16, x, packed, packed_bias(packed, cin_pad, cout), residual, y,
                            nullptr, n, h, w, cin_pad, cout, relu, amode, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_convT3x3s2_fwd(const void* x, const void* packed, void* y, int n, int h, int w, int cin,
                                 int cout, int relu, int amode, void* stream) {
  TG_CHECK_ARG(cout == 64 || cout == 128, "convT3x3s2_fwd: cout must be 64 or 128 (got %d)", cout);
  return tg::launch_conv_tc(tg::kConvT3x3s2, tg::kOutNHWCbf16, x, packed, packed_bias(packed, cin, cout), nullptr, y,
                            nullptr, n, h, w, cin, cout, relu, amode, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_conv3x3_out_sigmoid(const void* x, const void* packed, float* out, float* logits, int n, int h,
                                      int w, int amode, void* stream) {
  return tg::launch_conv_tc(tg::kConv3x3, tg::kOutNCHWf32Sigmoid, x, packed, packed_bias(packed, 64, 16), nullptr, out,
                            logits, n, h, w, 64, 16, 0, amode, 0, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------ data gradients
extern "C" int tg_conv3x3_dgrad(const void* dy, const void* packed_dgrad, const void* residual, const void* mask,
                                void* dx, int n, int h, int w, int cin, int cout, void* stream) {
  // dX = conv3x3(dY, rotated/transposed W) (+ residual), zeroed where mask == 0 (ReLU backward)
  const int dci = cout <= 64 ? 64 : 128, dco = cin <= 64 ? 64 : 128;
  return tg::launch_conv_tc(tg::kConv3x3, tg::kOutNHWCbf16, dy, packed_dgrad, packed_bias(packed_dgrad, dci, dco), residual,
                            dx, nullptr, n, h, w, dci, dco, 0, TG_AMODE_HALO, 0, static_cast<cudaStream_t>(stream), mask);
}

extern "C" int tg_convT3x3s2_dgrad(const void* dy, const void* packed_dgrad, const void* mask, void* dx, int n, int h,
                                   int w, int cin, int cout, void* stream) {
  // forward: x [n,h,w,cin] -> y [n,2h,2w,cout]; dX = conv k3 s2 p1 over dY, run as the 4x4 stride-2 kernel
  const int dci = cout <= 64 ? 64 : 128, dco = cin <= 64 ? 64 : 128;
  const float* bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed_dgrad) +
                                                     tg::packed_weight_bytes_k(tg::kConv4x4s2, dci, dco));
  return tg::launch_conv_tc(tg::kConv4x4s2, tg::kOutNHWCbf16, dy, packed_dgrad, bias, nullptr, dx, nullptr, n, 2 * h, 2 * w,
                            dci, dco, 0, TG_AMODE_HALO, 0, static_cast<cudaStream_t>(stream), mask);
}

extern "C" int tg_conv4x4s2_dgrad(const void* dy, const void* packed_dgrad, const void* mask, int mask_mode, void* dx,
                                  int n, int h, int w, int cin, int cout, void* stream) {
  // forward: x [n,2h,2w,cin] -> y [n,h,w,cout]; dX is produced one output-parity phase per launch (2x2 taps each)
  const int dci = cout <= 64 ? 64 : 128, dco = cin <= 64 ? 64 : 128;
  const size_t phase_bytes = static_cast<size_t>(4) * dci * dco * 2;
  const float* zero_bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed_dgrad) + 4 * phase_bytes);
  for (int ph = 0; ph < 4; ++ph) {
    int rc = tg::launch_conv_tc(tg::kConvT4x4s2Phase, tg::kOutNHWCbf16, dy, static_cast<const uint8_t*>(packed_dgrad) + ph * phase_bytes,
                                zero_bias, nullptr, dx, nullptr, n, h, w, dci, dco, 0, TG_AMODE_HALO, 0,
                                static_cast<cudaStream_t>(stream), mask, mask_mode, ph);
    if (rc) return rc;
  }
  return TG_OK;
}
