// BatchNorm2d(eps=1e-3) in training mode (reference code/ops.py:75-77, used by discriminator_block and the
// discriminator's residual blocks, code/models.py:90-94,106,130) on NHWC bf16 activations, plus the
// discriminator head (block5 BatchNorm + LeakyReLU + flatten + Linear + sigmoid, code/models.py:121,141-145).
// HBM/latency-bound elementwise + reduction kernels: 128-bit vector access, fp32 statistics.
#include "tg_disc.cuh"

namespace tg {

constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = 296;   // two per SM: enough loads in flight to stream at HBM rate, short serial tail
constexpr int kBnMaxGroups = 2;     // independent BatchNorm batches inside one launch (the real and the fake pass)

// ---------------------------------------------------------------------------------------------
// Batch statistics of x [P pixels][C] f32 (raw conv outputs are kept in f32: normalisation amplifies rounding).  Every block reduces a strided slice of the pixels to
// per-channel (sum, sum of squares) partials; the last block to finish (atomic ticket) reduces the
// partials in a fixed order (deterministic), writes {mean, rstd, a, b} with y = a*x + b the folded
// normalise+affine, and applies the running-statistics update of nn.BatchNorm2d (momentum 0.1,
// unbiased variance).  The ticket is reset for the next use.
// ---------------------------------------------------------------------------------------------
// Groups (gridDim.y): the discriminator's real and fake forward passes (code/train.py:181,199) run as ONE batch of 2n
// samples through every kernel; BatchNorm statistics stay per pass: group g = samples [g*n, (g+1)*n), its own partials
// and stats block (stats + g*512).  The LAST block of all groups finalizes the groups in order 0, 1, ..., so the running
// statistics receive the two momentum updates in the order two separate forward calls would apply them.
__global__ void __launch_bounds__(kBnThreads)
bn_stats_nhwc_kernel(const float* __restrict__ x, long long pixels, int c, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float momentum, float* __restrict__ partial,
                     unsigned int* __restrict__ ticket, float* __restrict__ stats, float* running_mean,
                     float* running_var, long long* num_batches_tracked) {
  x += static_cast<long long>(blockIdx.y) * pixels * c;
  float* const partial_all = partial;
  partial += static_cast<size_t>(blockIdx.y) * kBnMaxBlocks * 128 * 2;
  __shared__ float s_sum[kBnThreads][9];      // +1 padding
  __shared__ float s_sq[kBnThreads][9];
  __shared__ bool s_last;
  const int groups = c / 8;                                   // 16-byte groups per pixel
  const int pix_per_iter = kBnThreads / groups;
  const int gi = threadIdx.x % groups, pl = threadIdx.x / groups;
  float sum[8], sq[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sum[e] = sq[e] = 0.f;
  for (long long p = static_cast<long long>(blockIdx.x) * pix_per_iter + pl; p < pixels;
       p += static_cast<long long>(gridDim.x) * pix_per_iter) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + p * c) + 2 * gi);
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + p * c) + 2 * gi + 1);
    const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) { sum[e] += f[e]; sq[e] += f[e] * f[e]; }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { s_sum[threadIdx.x][e] = sum[e]; s_sq[threadIdx.x][e] = sq[e]; }
  __syncthreads();
  // one thread per channel sums the pix_per_iter rows that hold it
  if (threadIdx.x < c) {
    const int g = threadIdx.x / 8, e = threadIdx.x % 8;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < pix_per_iter; ++r) { a += s_sum[r * groups + g][e]; b += s_sq[r * groups + g][e]; }
    partial[(static_cast<size_t>(blockIdx.x) * c + threadIdx.x) * 2 + 0] = a;
    partial[(static_cast<size_t>(blockIdx.x) * c + threadIdx.x) * 2 + 1] = b;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < c) {
    for (unsigned int g = 0; g < gridDim.y; ++g) {
      const float* pg = partial_all + static_cast<size_t>(g) * kBnMaxBlocks * 128 * 2;
      float* sg = stats + g * 512;
      double a = 0.0, b = 0.0;
      for (unsigned int k = 0; k < gridDim.x; ++k) {
        a += static_cast<double>(__ldcg(pg + (static_cast<size_t>(k) * c + threadIdx.x) * 2 + 0));
        b += static_cast<double>(__ldcg(pg + (static_cast<size_t>(k) * c + threadIdx.x) * 2 + 1));
      }
      const double mean = a / static_cast<double>(pixels);
      double var = b / static_cast<double>(pixels) - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      const float ga = gamma[threadIdx.x] * rstd;
      sg[threadIdx.x * 4 + 0] = static_cast<float>(mean);
      sg[threadIdx.x * 4 + 1] = rstd;
      sg[threadIdx.x * 4 + 2] = ga;
      sg[threadIdx.x * 4 + 3] = beta[threadIdx.x] - static_cast<float>(mean) * ga;
      if (running_mean) {
        const double unbiased = pixels > 1 ? var * static_cast<double>(pixels) / static_cast<double>(pixels - 1) : var;
        running_mean[threadIdx.x] = (1.f - momentum) * running_mean[threadIdx.x] + momentum * static_cast<float>(mean);
        running_var[threadIdx.x] = (1.f - momentum) * running_var[threadIdx.x] + momentum * static_cast<float>(unbiased);
      }
    }
  }
  if (threadIdx.x == 0) {
    if (num_batches_tracked) *num_batches_tracked += gridDim.y;
    *ticket = 0u;
  }
}

// eval mode: fold the running statistics into {mean, rstd, a, b}
__global__ void bn_fold_running_kernel(int c, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float eps, const float* __restrict__ running_mean,
                                       const float* __restrict__ running_var, float* __restrict__ stats) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float rstd = rsqrtf(running_var[ch] + eps);
  const float ga = gamma[ch] * rstd;
  stats[ch * 4 + 0] = running_mean[ch];
  stats[ch * 4 + 1] = rstd;
  stats[ch * 4 + 2] = ga;
  stats[ch * 4 + 3] = beta[ch] - running_mean[ch] * ga;
}

// y = act(a*x + b) (+ skip); x, skip: [P][C] f32; written twice: y32 (f32: the residual stream / feature maps) and
// y16 (bf16: the next convolution's operand).  act 0 none / 2 LeakyReLU(0.2)
__global__ void __launch_bounds__(kBnThreads)
bn_apply_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ skip, float* __restrict__ y32,
                     __nv_bfloat16* __restrict__ y16, long long pixels, int c, const float* __restrict__ stats, int act) {
  __shared__ float s_a[128], s_b[128];
  {                                                           // group blockIdx.y: its slice of every tensor, its stats block
    const long long goff = static_cast<long long>(blockIdx.y) * pixels * c;
    x += goff;
    if (skip) skip += goff;
    if (y32) y32 += goff;
    if (y16) y16 += goff;
    stats += blockIdx.y * 512;
  }
  for (int i = threadIdx.x; i < c; i += blockDim.x) { s_a[i] = stats[i * 4 + 2]; s_b[i] = stats[i * 4 + 3]; }
  __syncthreads();
  const int groups = c / 4;
  const long long total = pixels * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      f[e] = s_a[g * 4 + e] * f[e] + s_b[g * 4 + e];
      if (act == kActLrelu02) f[e] = f[e] > 0.f ? f[e] : 0.2f * f[e];
    }
    if (skip) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(skip) + i);
      f[0] += s4.x; f[1] += s4.y; f[2] += s4.z; f[3] += s4.w;
    }
    if (y32) reinterpret_cast<float4*>(y32)[i] = make_float4(f[0], f[1], f[2], f[3]);
    if (y16) reinterpret_cast<uint2*>(y16)[i] = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
  }
}

// NHWC f32 [n][hw][c] -> NCHW f32 [n][c][hw] (the feature maps returned in layer_list, code/models.py:131,135,139,141)
__global__ void __launch_bounds__(256)
nhwc_f32_to_nchw_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int c, long long hw) {
  __shared__ float tile[32][129];
  const long long groups = (hw + 31) / 32;
  for (long long gi = blockIdx.x; gi < groups * n; gi += gridDim.x) {
    const long long b = gi / groups, p0 = (gi % groups) * 32;
    for (int i = threadIdx.x; i < 32 * c; i += blockDim.x) {
      const int px = i / c, ch = i % c;
      tile[px][ch] = (p0 + px < hw) ? in[(b * hw + p0 + px) * c + ch] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * c; i += blockDim.x) {
      const int ch = i / 32, px = i % 32;
      if (p0 + px < hw) out[(b * c + ch) * hw + p0 + px] = tile[px][ch];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Discriminator head, one block: raw block5 conv output r [n][3][hw] f32 -> BatchNorm (batch stats, 3 channels)
// -> LeakyReLU(0.2) -> flatten (NCHW order) -> Linear(3*hw, 1) -> sigmoid.   code/models.py:121,141-145.
// Saves y (post-LeakyReLU, the fc input) and {mean, rstd, a, b} for the backward pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
disc_head_kernel(const float* r, int n, int hw, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, float momentum, int training, float* running_mean, float* running_var,
                 long long* num_batches_tracked, const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                 float* y, float* stats, float* logit, float* prob, int ngroups) {
  __shared__ double s_red[2][256];
  __shared__ float s_ab[3][2];
  const int per = n * hw;
  // groups (gridDim.x is 1; the group loop is serial so that the running statistics see the passes in order)
  const float* const r_all = r; float* const y_all = y; float* const stats_all = stats; float* const logit_all = logit; float* const prob_all = prob;
  for (int grp = 0; grp < ngroups; ++grp) {
  r = r_all + static_cast<size_t>(grp) * n * 3 * hw; y = y_all + static_cast<size_t>(grp) * n * 3 * hw;
  stats = stats_all + grp * 512; logit = logit_all ? logit_all + grp * n : nullptr; prob = prob_all + grp * n;
  __syncthreads();
  for (int ch = 0; ch < 3; ++ch) {
    double a = 0.0, b = 0.0;
    if (training) {
      for (int i = threadIdx.x; i < per; i += blockDim.x) {
        const float v = r[(static_cast<long long>(i / hw) * 3 + ch) * hw + i % hw];
        a += v; b += static_cast<double>(v) * v;
      }
    }
    s_red[0][threadIdx.x] = a; s_red[1][threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) { s_red[0][threadIdx.x] += s_red[0][threadIdx.x + s]; s_red[1][threadIdx.x] += s_red[1][threadIdx.x + s]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      double mean, var;
      if (training) {
        mean = s_red[0][0] / per;
        var = s_red[1][0] / per - mean * mean;
        if (var < 0.0) var = 0.0;
        if (running_mean) {
          const double unbiased = per > 1 ? var * per / (per - 1) : var;
          running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * static_cast<float>(mean);
          running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * static_cast<float>(unbiased);
        }
      } else {
        mean = running_mean[ch]; var = running_var[ch];
      }
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      const float ga = gamma[ch] * rstd;
      stats[ch * 4 + 0] = static_cast<float>(mean); stats[ch * 4 + 1] = rstd;
      stats[ch * 4 + 2] = ga; stats[ch * 4 + 3] = beta[ch] - static_cast<float>(mean) * ga;
      s_ab[ch][0] = ga; s_ab[ch][1] = stats[ch * 4 + 3];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
  const int feat = 3 * hw;
  for (int i = threadIdx.x; i < n * feat; i += blockDim.x) {
    const int ch = (i % feat) / hw;
    float v = s_ab[ch][0] * r[i] + s_ab[ch][1];
    y[i] = v > 0.f ? v : 0.2f * v;
  }
  __syncthreads();
  // one warp per sample: dot(y[s], fc_w) + b
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < n; s += blockDim.x / 32) {
    float acc = 0.f;
    for (int k = lane; k < feat; k += 32) acc += y[s * feat + k] * fc_w[k];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float z = acc + fc_b[0];
      if (logit) logit[s] = z;
      prob[s] = 1.f / (1.f + expf(-z));
    }
  }
  }   // groups
}

// ---------------------------------------------------------------------------------------------
// BatchNorm backward (training mode).  g_out = gradient w.r.t. the block output (bf16); if `act` is given the block
// ended in LeakyReLU(0.2) and g = g_out * (act > 0 ? 1 : 0.2), else g = g_out.  With xh = (x - mean) * rstd:
//   dbeta = sum g, dgamma = sum g*xh, dx = gamma*rstd * (g - dbeta/P - xh*dgamma/P).
// Pass 1 reduces (deterministic partials + last-block finalize, like bn_stats), adds dgamma/dbeta into the
// parameter gradient and leaves {dbeta/P, dgamma/P} in `red`; pass 2 writes dx as bf16 (the conv kernels' operand).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ g_out, const float* __restrict__ x, const float* __restrict__ act,
                     long long pixels, int c, const float* __restrict__ stats, float* __restrict__ partial,
                     unsigned int* __restrict__ ticket, float* __restrict__ red, float* dgamma, float* dbeta) {
  __shared__ float s_sum[kBnThreads][9];
  __shared__ float s_sq[kBnThreads][9];
  __shared__ bool s_last;
  {                                                           // group blockIdx.y (see bn_stats_nhwc_kernel)
    const long long goff = static_cast<long long>(blockIdx.y) * pixels * c;
    g_out += goff; x += goff;
    if (act) act += goff;
    stats += blockIdx.y * 512;
  }
  float* const partial_all = partial;
  partial += static_cast<size_t>(blockIdx.y) * kBnMaxBlocks * 128 * 2;
  const int groups = c / 8;
  const int pix_per_iter = kBnThreads / groups;
  const int gi = threadIdx.x % groups, pl = threadIdx.x / groups;
  float mean[8], rstd[8], s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { mean[e] = stats[(gi * 8 + e) * 4 + 0]; rstd[e] = stats[(gi * 8 + e) * 4 + 1]; s1[e] = s2[e] = 0.f; }
  for (long long p = static_cast<long long>(blockIdx.x) * pix_per_iter + pl; p < pixels;
       p += static_cast<long long>(gridDim.x) * pix_per_iter) {
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(g_out + p * c) + gi);
    const uint32_t gu[4] = {gv.x, gv.y, gv.z, gv.w};
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + p * c) + 2 * gi);
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(x + p * c) + 2 * gi + 1);
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    float av[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (act) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(act + p * c) + 2 * gi);
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(act + p * c) + 2 * gi + 1);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float g = (e & 1) ? bf16_hi(gu[e >> 1]) : bf16_lo(gu[e >> 1]);
      if (act && !(av[e] > 0.f)) g *= 0.2f;
      s1[e] += g;
      s2[e] += g * (xv[e] - mean[e]) * rstd[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { s_sum[threadIdx.x][e] = s1[e]; s_sq[threadIdx.x][e] = s2[e]; }
  __syncthreads();
  if (threadIdx.x < c) {
    const int g = threadIdx.x / 8, e = threadIdx.x % 8;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < pix_per_iter; ++r) { a += s_sum[r * groups + g][e]; b += s_sq[r * groups + g][e]; }
    partial[(static_cast<size_t>(blockIdx.x) * c + threadIdx.x) * 2 + 0] = a;
    partial[(static_cast<size_t>(blockIdx.x) * c + threadIdx.x) * 2 + 1] = b;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < c) {
    for (unsigned int g = 0; g < gridDim.y; ++g) {
      const float* pg = partial_all + static_cast<size_t>(g) * kBnMaxBlocks * 128 * 2;
      double a = 0.0, b = 0.0;
      for (unsigned int k = 0; k < gridDim.x; ++k) {
        a += static_cast<double>(__ldcg(pg + (static_cast<size_t>(k) * c + threadIdx.x) * 2 + 0));
        b += static_cast<double>(__ldcg(pg + (static_cast<size_t>(k) * c + threadIdx.x) * 2 + 1));
      }
      dbeta[threadIdx.x] += static_cast<float>(a);
      dgamma[threadIdx.x] += static_cast<float>(b);
      red[g * 256 + threadIdx.x * 2 + 0] = static_cast<float>(a / static_cast<double>(pixels));
      red[g * 256 + threadIdx.x * 2 + 1] = static_cast<float>(b / static_cast<double>(pixels));
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g_out, const float* __restrict__ x, const float* __restrict__ act,
                    __nv_bfloat16* __restrict__ dx, long long pixels, int c, const float* __restrict__ stats,
                    const float* __restrict__ red) {
  __shared__ float s_mean[128], s_rstd[128], s_a[128], s_m1[128], s_m2[128];
  {
    const long long goff = static_cast<long long>(blockIdx.y) * pixels * c;
    g_out += goff; x += goff; dx += goff;
    if (act) act += goff;
    stats += blockIdx.y * 512;
    red += blockIdx.y * 256;
  }
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    s_mean[i] = stats[i * 4 + 0]; s_rstd[i] = stats[i * 4 + 1]; s_a[i] = stats[i * 4 + 2];
    s_m1[i] = red[i * 2 + 0]; s_m2[i] = red[i * 2 + 1];
  }
  __syncthreads();
  const int groups = c / 4;
  const long long total = pixels * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g4 = static_cast<int>(i % groups) * 4;
    const uint2 gv = __ldg(reinterpret_cast<const uint2*>(g_out) + i);
    float g[4] = {bf16_lo(gv.x), bf16_hi(gv.x), bf16_lo(gv.y), bf16_hi(gv.y)};
    const float4 xv4 = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
    if (act) {
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(act) + i);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) if (!(av[e] > 0.f)) g[e] *= 0.2f;
    }
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (xv[e] - s_mean[g4 + e]) * s_rstd[g4 + e];
      o[e] = s_a[g4 + e] * (g[e] - s_m1[g4 + e] - xh * s_m2[g4 + e]);
    }
    reinterpret_cast<uint2*>(dx)[i] = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
  }
}

// Backward of the discriminator head (disc_head_kernel), one block: dprob -> sigmoid -> fc -> LeakyReLU -> BatchNorm(3).
// Writes d(raw block5 conv output) as NHWC bf16 [n][hw][64] (3 real channels) and ADDS the fc / BN parameter gradients.
__global__ void __launch_bounds__(256)
disc_head_bwd_kernel(const float* dprob, const float* prob, const float* y, const float* r, int n, int hw, const float* stats,
                     const float* __restrict__ fc_w, float* d_fc_w, float* d_fc_b, float* dgamma, float* dbeta,
                     float* dlogit, __nv_bfloat16* dr, int ngroups) {
  __shared__ double s_red[2][256];
  __shared__ float s_m[3][2];
  const int feat = 3 * hw, per = n * hw;
  const float* const dprob_all = dprob; const float* const prob_all = prob; const float* const y_all = y; const float* const r_all = r;
  const float* const stats_all = stats; float* const dlogit_all = dlogit; __nv_bfloat16* const dr_all = dr;
  for (int grp = 0; grp < ngroups; ++grp) {
  dprob = dprob_all + grp * n; prob = prob_all + grp * n; y = y_all + static_cast<size_t>(grp) * n * feat; r = r_all + static_cast<size_t>(grp) * n * feat;
  stats = stats_all + grp * 512; dlogit = dlogit_all + grp * n; dr = dr_all + static_cast<size_t>(grp) * per * 64;
  __syncthreads();
  for (int s = threadIdx.x; s < n; s += blockDim.x) dlogit[s] = dprob[s] * prob[s] * (1.f - prob[s]);
  __syncthreads();
  for (int k = threadIdx.x; k < feat; k += blockDim.x) {           // fc weight gradient
    float a = 0.f;
    for (int s = 0; s < n; ++s) a += dlogit[s] * y[s * feat + k];
    d_fc_w[k] += a;
  }
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int s = 0; s < n; ++s) a += dlogit[s];
    d_fc_b[0] += a;
  }
  for (int ch = 0; ch < 3; ++ch) {                                  // BatchNorm reductions per channel
    const float mean = stats[ch * 4 + 0], rstd = stats[ch * 4 + 1];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
      const int s = i / hw, k = ch * hw + i % hw;
      float g = dlogit[s] * fc_w[k];
      if (!(y[s * feat + k] > 0.f)) g *= 0.2f;
      a += g; b += static_cast<double>(g) * (r[s * feat + k] - mean) * rstd;
    }
    s_red[0][threadIdx.x] = a; s_red[1][threadIdx.x] = b;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if (threadIdx.x < st) { s_red[0][threadIdx.x] += s_red[0][threadIdx.x + st]; s_red[1][threadIdx.x] += s_red[1][threadIdx.x + st]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      dbeta[ch] += static_cast<float>(s_red[0][0]);
      dgamma[ch] += static_cast<float>(s_red[1][0]);
      s_m[ch][0] = static_cast<float>(s_red[0][0] / per); s_m[ch][1] = static_cast<float>(s_red[1][0] / per);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < per; i += blockDim.x) {             // dx, NHWC64 bf16
    const int s = i / hw, px = i % hw;
    float o[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int k = ch * hw + px;
      float g = dlogit[s] * fc_w[k];
      if (!(y[s * feat + k] > 0.f)) g *= 0.2f;
      const float xh = (r[s * feat + k] - stats[ch * 4 + 0]) * stats[ch * 4 + 1];
      o[ch] = stats[ch * 4 + 2] * (g - s_m[ch][0] - xh * s_m[ch][1]);
    }
    uint4* dst = reinterpret_cast<uint4*>(dr + static_cast<size_t>(i) * 64);
    dst[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], 0.f), 0u, 0u);
#pragma unroll
    for (int k = 1; k < 8; ++k) dst[k] = make_uint4(0u, 0u, 0u, 0u);
  }
  }   // groups
}

// ------------------------------------------------------------------------------------ launchers
int bn_stats_launch(const void* x, long long pixels, int c, const float* gamma, const float* beta, float* partial,
                    unsigned int* ticket, float* stats, float* running_mean, float* running_var,
                    long long* nbt, cudaStream_t st, int groups) {
  TG_CHECK_ARG(c == 64 || c == 128, "bn_stats: channels must be 64 or 128 (got %d)", c);
  TG_CHECK_ARG(pixels >= 1, "bn_stats: empty batch");
  TG_CHECK_ARG(groups >= 1 && groups <= kBnMaxGroups, "bn: groups must be 1..%d", kBnMaxGroups);
  const int pix_per_iter = kBnThreads / (c / 8);
  long long blocks = (pixels + pix_per_iter * 4 - 1) / (pix_per_iter * 4);
  if (blocks > kBnMaxBlocks) blocks = kBnMaxBlocks;
  if (blocks < 1) blocks = 1;
  tg_prof_pre(TG_K_GLUE, 4.0 * pixels * c * groups, st);
  bn_stats_nhwc_kernel<<<dim3(static_cast<int>(blocks), groups), kBnThreads, 0, st>>>(
      static_cast<const float*>(x), pixels, c, gamma, beta, 1e-3f, 0.1f, partial, ticket, stats, running_mean,
      running_var, nbt);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}
size_t bn_partial_floats() { return static_cast<size_t>(kBnMaxGroups) * kBnMaxBlocks * 128 * 2; }

int bn_fold_running_launch(int c, const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float* stats, cudaStream_t st) {
  bn_fold_running_kernel<<<1, 128, 0, st>>>(c, gamma, beta, 1e-3f, running_mean, running_var, stats);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int bn_apply_launch(const void* x, const void* skip, void* y32, void* y16, long long pixels, int c, const float* stats,
                    int act, cudaStream_t st, int groups) {
  TG_CHECK_ARG(c == 64 || c == 128, "bn_apply: channels must be 64 or 128 (got %d)", c);
  TG_CHECK_ARG(groups >= 1 && groups <= kBnMaxGroups, "bn: groups must be 1..%d", kBnMaxGroups);
  const long long total = pixels * (c / 4);
  long long blocks = (total + kBnThreads - 1) / kBnThreads;
  const long long cap = static_cast<long long>(tg_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  tg_prof_pre(TG_K_GLUE, (skip ? 14.0 : 10.0) * pixels * c * groups, st);
  bn_apply_nhwc_kernel<<<dim3(static_cast<int>(blocks), groups), kBnThreads, 0, st>>>(
      static_cast<const float*>(x), static_cast<const float*>(skip), static_cast<float*>(y32),
      static_cast<__nv_bfloat16*>(y16), pixels, c, stats, act);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int nhwc_to_nchw_f32_launch(const void* in, float* out, int n, int c, long long hw, cudaStream_t st) {
  TG_CHECK_ARG(c <= 128, "nhwc_to_nchw: at most 128 channels");
  const long long groups = (hw + 31) / 32 * n;
  const long long cap = static_cast<long long>(tg_num_sms()) * 8;
  tg_prof_pre(TG_K_GLUE, 8.0 * n * c * hw, st);
  nhwc_f32_to_nchw_f32_kernel<<<static_cast<int>(groups < cap ? groups : cap), 256, 0, st>>>(
      static_cast<const float*>(in), out, n, c, hw);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int disc_head_launch(const float* r, int n, int hw, const float* gamma, const float* beta, int training,
                     float* running_mean, float* running_var, long long* nbt, const float* fc_w, const float* fc_b,
                     float* y, float* stats, float* logit, float* prob, cudaStream_t st, int groups) {
  tg_prof_pre(TG_K_GLUE, 8.0 * n * 3 * hw * groups, st);
  disc_head_kernel<<<1, 256, 0, st>>>(r, n, hw, gamma, beta, 1e-3f, 0.1f, training, running_mean, running_var, nbt, fc_w,
                                      fc_b, y, stats, logit, prob, groups);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int bn_bwd_launch(const void* g_out, const void* x, const void* act, void* dx, long long pixels, int c, const float* stats,
                  float* partial, unsigned int* ticket, float* red, float* dgamma, float* dbeta, cudaStream_t st, int groups) {
  TG_CHECK_ARG(c == 64 || c == 128, "bn_bwd: channels must be 64 or 128 (got %d)", c);
  TG_CHECK_ARG(groups >= 1 && groups <= kBnMaxGroups, "bn: groups must be 1..%d", kBnMaxGroups);
  const int pix_per_iter = kBnThreads / (c / 8);
  long long blocks = (pixels + pix_per_iter * 4 - 1) / (pix_per_iter * 4);
  if (blocks > kBnMaxBlocks) blocks = kBnMaxBlocks;
  if (blocks < 1) blocks = 1;
  tg_prof_pre(TG_K_GLUE, (act ? 10.0 : 6.0) * pixels * c * groups, st);
  bn_bwd_reduce_kernel<<<dim3(static_cast<int>(blocks), groups), kBnThreads, 0, st>>>(
      static_cast<const __nv_bfloat16*>(g_out), static_cast<const float*>(x), static_cast<const float*>(act), pixels, c, stats,
      partial, ticket, red, dgamma, dbeta);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  const long long total = pixels * (c / 4);
  long long ab = (total + kBnThreads - 1) / kBnThreads;
  const long long cap = static_cast<long long>(tg_num_sms()) * 8;
  if (ab > cap) ab = cap;
  tg_prof_pre(TG_K_GLUE, (act ? 12.0 : 8.0) * pixels * c * groups, st);
  bn_bwd_apply_kernel<<<dim3(static_cast<int>(ab), groups), kBnThreads, 0, st>>>(
      static_cast<const __nv_bfloat16*>(g_out), static_cast<const float*>(x), static_cast<const float*>(act),
      static_cast<__nv_bfloat16*>(dx), pixels, c, stats, red);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int disc_head_bwd_launch(const float* dprob, const float* prob, const float* y, const float* r, int n, int hw,
                         const float* stats, const float* fc_w, float* d_fc_w, float* d_fc_b, float* dgamma, float* dbeta,
                         float* dlogit, void* dr, cudaStream_t st, int groups) {
  tg_prof_pre(TG_K_GLUE, 16.0 * n * 3 * hw * groups, st);
  disc_head_bwd_kernel<<<1, 256, 0, st>>>(dprob, prob, y, r, n, hw, stats, fc_w, d_fc_w, d_fc_b, dgamma, dbeta, dlogit,
                                          static_cast<__nv_bfloat16*>(dr), groups);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

}  // namespace tg

// ------------------------------------------------------------------------------------ C ABI (tg_bn_{stats,apply,bwd})
using namespace tg;

namespace {
struct BnWs { size_t partial, ticket, red, total; };
BnWs bn_ws() {
  BnWs w;
  w.partial = 0;
  w.ticket = (bn_partial_floats() * 4 + 255) & ~static_cast<size_t>(255);
  w.red = w.ticket + 256;
  w.total = w.red + kBnMaxGroups * 2 * 128 * 4;
  return w;
}
int bn_check_ws(const char* who, const void* ws, size_t bytes) {
  TG_CHECK_ARG(ws && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, "%s: workspace must be non-null and 256-byte aligned", who);
  if (bytes < bn_ws().total) {
    tg_set_error("%s: workspace too small (%zu < %zu)", who, bytes, bn_ws().total);
    return TG_ERR_WORKSPACE;
  }
  return TG_OK;
}
}  // namespace

extern "C" size_t tg_workspace_bytes_bn(void) { return bn_ws().total; }

extern "C" int tg_bn_stats(const float* x, long long pixels, int c, const float* gamma, const float* beta, float* stats,
                           float* running_mean, float* running_var, long long* num_batches_tracked, void* workspace,
                           size_t workspace_bytes, void* stream) {
  TG_CHECK_ARG(x && gamma && beta && stats, "bn_stats: null pointer");
  if (int rc = bn_check_ws("bn_stats", workspace, workspace_bytes)) return rc;
  const BnWs w = bn_ws();
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TG_CUDA(cudaMemsetAsync(wsp + w.ticket, 0, 256, st));
  return bn_stats_launch(x, pixels, c, gamma, beta, reinterpret_cast<float*>(wsp + w.partial),
                         reinterpret_cast<unsigned int*>(wsp + w.ticket), stats, running_mean, running_var,
                         num_batches_tracked, st);
}

extern "C" int tg_bn_apply(const float* x, const float* skip, float* y_f32, void* y_bf16, long long pixels, int c,
                           const float* stats, int act, void* stream) {
  TG_CHECK_ARG(x && stats && (y_f32 || y_bf16), "bn_apply: null pointer");
  TG_CHECK_ARG(act == kActNone || act == kActLrelu02, "bn_apply: act must be 0 (none) or 2 (LeakyReLU 0.2)");
  TG_CHECK_ARG(pixels >= 1, "bn_apply: empty batch");
  return bn_apply_launch(x, skip, y_f32, y_bf16, pixels, c, stats, act, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_bn_bwd(const void* g_out, const float* x, const float* act_out, void* dx, long long pixels, int c,
                         const float* stats, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                         void* stream) {
  TG_CHECK_ARG(g_out && x && dx && stats && dgamma && dbeta, "bn_bwd: null pointer");
  TG_CHECK_ARG(pixels >= 1, "bn_bwd: empty batch");
  if (int rc = bn_check_ws("bn_bwd", workspace, workspace_bytes)) return rc;
  const BnWs w = bn_ws();
  uint8_t* wsp = static_cast<uint8_t*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TG_CUDA(cudaMemsetAsync(wsp + w.ticket, 0, 256, st));
  return bn_bwd_launch(g_out, x, act_out, dx, pixels, c, stats, reinterpret_cast<float*>(wsp + w.partial),
                       reinterpret_cast<unsigned int*>(wsp + w.ticket), reinterpret_cast<float*>(wsp + w.red), dgamma,
                       dbeta, st);
}
