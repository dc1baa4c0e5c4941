// HBM-bound glue of the TecoGAN frame loop: space-to-depth / depth-to-space, bilinear backward
// warp, bilinear x4 upscale, and the fused producer of the generator input.
// All kernels are pure streaming / gather kernels: coalesced 128-bit accesses on the contiguous
// side, shared-memory staging where the two sides disagree, grid-stride over whole rows.
#include "tg_frame.cuh"

namespace tg {

// ---------------------------------------------------------------------------------------------
// space_to_depth / depth_to_space, r = 4 fast path: one thread moves one 16-byte group
// in[n,c,4y+dy,4x..4x+3]  <->  out[n, c*16+dy*4+{0..3}, y, x].
// 4-byte payload, bit-exact (values are only moved).
// ---------------------------------------------------------------------------------------------
__global__ void s2d4_kernel(const uint4* __restrict__ in, uint32_t* __restrict__ out, int planes, int ho,
                            int wo) {
  // in viewed as [planes][4*ho][wo] uint4 ; out as [planes][16][ho][wo]
  const long long total = static_cast<long long>(planes) * 4 * ho * wo;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % wo);
    long long r = i / wo;
    const int iy = static_cast<int>(r % (4 * ho));
    const long long pl = r / (4 * ho);
    const int y = iy >> 2, dy = iy & 3;
    const uint4 v = __ldg(in + i);
    uint32_t* o = out + ((pl * 16 + dy * 4) * ho + y) * static_cast<long long>(wo) + x;
    const long long ps = static_cast<long long>(ho) * wo;
    o[0] = v.x; o[ps] = v.y; o[2 * ps] = v.z; o[3 * ps] = v.w;
  }
}

__global__ void d2s4_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out, int planes, int hi,
                            int wi) {
  // two output groups per thread iteration: eight 4-byte plane loads in flight feed two 16-byte stores
  const long long total = static_cast<long long>(planes) * 4 * hi * wi;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long ps = static_cast<long long>(hi) * wi;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += 2 * stride) {
    uint4 v[2];
    long long idx[2] = {i, i + stride};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (idx[u] < total) {
        const int x = static_cast<int>(idx[u] % wi);
        long long r = idx[u] / wi;
        const int oy = static_cast<int>(r % (4 * hi));
        const long long pl = r / (4 * hi);
        const int y = oy >> 2, dy = oy & 3;
        const uint32_t* s = in + ((pl * 16 + dy * 4) * hi + y) * static_cast<long long>(wi) + x;
        v[u].x = __ldg(s); v[u].y = __ldg(s + ps); v[u].z = __ldg(s + 2 * ps); v[u].w = __ldg(s + 3 * ps);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (idx[u] < total) out[idx[u]] = v[u];
  }
}

// generic r (any), one thread per element of the depth-side tensor
__global__ void s2d_generic_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int planes,
                                   int ho, int wo, int r, int to_depth) {
  const long long total = static_cast<long long>(planes) * r * r * ho * wo;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % wo);
    long long q = i / wo;
    const int y = static_cast<int>(q % ho); q /= ho;
    const int dx = static_cast<int>(q % r); q /= r;
    const int dy = static_cast<int>(q % r); q /= r;
    const long long pl = q;
    const long long sp = (pl * (static_cast<long long>(ho) * r) + (static_cast<long long>(y) * r + dy)) *
                             (static_cast<long long>(wo) * r) + static_cast<long long>(x) * r + dx;
    if (to_depth) out[i] = __ldg(in + sp);
    else out[sp] = __ldg(in + i);
  }
}

// ---------------------------------------------------------------------------------------------
// Bilinear helpers.  ATen semantics, align_corners=False.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float round_fp16(float v) { return __half2float(__float2half_rn(v)); }

// value of upscale_four(src*pre)[row, col] for one plane (nn.Upsample bilinear x4)
__device__ __forceinline__ float up4_sample(const float* __restrict__ plane, int h, int w, int row, int col,
                                            float pre) {
  float sy = (row + 0.5f) * 0.25f - 0.5f;
  float sx = (col + 0.5f) * 0.25f - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = min(static_cast<int>(sy), h - 1), x0 = min(static_cast<int>(sx), w - 1);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly1 = sy - y0, lx1 = sx - x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float p00 = __ldg(plane + y0 * w + x0) * pre, p01 = __ldg(plane + y0 * w + x1) * pre;
  const float p10 = __ldg(plane + y1 * w + x0) * pre, p11 = __ldg(plane + y1 * w + x1) * pre;
  return ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
}

// One thread = the four outputs 4x..4x+3 of one output row; they read source columns x-1, x, x+1 of two source rows,
// so six loads (not sixteen) feed one 16-byte store.  One CTA walks whole output rows: no per-element 64-bit div/mod.
// Arithmetic is up4_sample's, operand for operand (ATen upsample_bilinear2d, align_corners=False).  The kernel writes
// 16 bytes for every byte it reads: its roofline is the HBM WRITE rate (scripts/glue_bench.py times a fill beside it).
// (A 4x4-outputs-per-thread variant - nine loads, four stores - measured slower: 130 us vs 92 us at cfg3 sizes.)
__global__ void __launch_bounds__(256)
upscale4_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int h, int w, float pre) {
  // one CTA iteration = a PAIR of output rows (2p, 2p+1): both interpolate between the same two source rows, so the
  // six loads of a thread feed two 16-byte stores
  const int wo = 4 * w, ho = 4 * h;
  const long long pairs = static_cast<long long>(planes) * (ho / 2);
  for (long long r = blockIdx.x; r < pairs; r += gridDim.x) {
    const long long pl = r / (ho / 2);
    const int oy = 2 * static_cast<int>(r - pl * (ho / 2));
    float sy0 = (oy + 0.5f) * 0.25f - 0.5f, sy1 = (oy + 1.5f) * 0.25f - 0.5f;
    sy0 = sy0 < 0.f ? 0.f : sy0;
    sy1 = sy1 < 0.f ? 0.f : sy1;
    const int y0 = min(static_cast<int>(sy0), h - 1), y1 = min(y0 + 1, h - 1);   // (== the pair's second row's y0, y1)
    const float la1 = sy0 - y0, la0 = 1.f - la1;               // weights of row oy
    const float lb1 = sy1 - y0, lb0 = 1.f - lb1;               // weights of row oy + 1
    const float* r0 = in + (pl * h + y0) * static_cast<long long>(w);
    const float* r1 = in + (pl * h + y1) * static_cast<long long>(w);
    float4* orow0 = reinterpret_cast<float4*>(out + (pl * ho + oy) * static_cast<long long>(wo));
    float4* orow1 = reinterpret_cast<float4*>(out + (pl * ho + oy + 1) * static_cast<long long>(wo));
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
      const int xm = max(x - 1, 0), xp = min(x + 1, w - 1);
      const float a0 = __ldg(r0 + xm) * pre, a1 = __ldg(r0 + x) * pre, a2 = __ldg(r0 + xp) * pre;
      const float b0 = __ldg(r1 + xm) * pre, b1 = __ldg(r1 + x) * pre, b2 = __ldg(r1 + xp) * pre;
      // sx = (4x + j + 0.5) / 4 - 0.5 is exact in f32, so the ATen weights lx1 = sx - floor(sx) are the constants
      // .625 .875 .125 .375 (j = 0..3); only column 0 differs: sx clamps to 0 -> taps (x, x+1) with weights (1, 0).
      const bool edge = x == 0;
      const float l0 = edge ? a1 : a0, l1 = edge ? a2 : a1;      // taps of outputs j = 0, 1 on row y0
      const float m0 = edge ? b1 : b0, m1 = edge ? b2 : b1;      // ... on row y1
      const float w0 = edge ? 0.f : 0.625f, w1 = edge ? 0.f : 0.875f;
      const float t0 = (1.f - w0) * l0 + w0 * l1, u0 = (1.f - w0) * m0 + w0 * m1;
      const float t1 = (1.f - w1) * l0 + w1 * l1, u1 = (1.f - w1) * m0 + w1 * m1;
      const float t2 = 0.875f * a1 + 0.125f * a2, u2 = 0.875f * b1 + 0.125f * b2;
      const float t3 = 0.625f * a1 + 0.375f * a2, u3 = 0.625f * b1 + 0.375f * b2;
      orow0[x] = make_float4(la0 * t0 + la1 * u0, la0 * t1 + la1 * u1, la0 * t2 + la1 * u2, la0 * t3 + la1 * u3);
      orow1[x] = make_float4(lb0 * t0 + lb1 * u0, lb0 * t1 + lb1 * u1, lb0 * t2 + lb1 * u2, lb0 * t3 + lb1 * u3);
    }
  }
}

// 4-tap bilinear gather with per-tap zero padding (grid_sampler_2d, align_corners=False)
struct Taps {
  int x0, y0;
  float wnw, wne, wsw, wse;
  bool ok_nw, ok_ne, ok_sw, ok_se;
};
__device__ __forceinline__ Taps make_taps(float gx, float gy, int h, int w) {
  Taps t;
  const float ix = ((gx + 1.f) * w - 1.f) / 2.f;
  const float iy = ((gy + 1.f) * h - 1.f) / 2.f;
  const float fx = floorf(ix), fy = floorf(iy);
  const float ax = ix - fx, ay = iy - fy;            // == (ix - ix_nw)
  const float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
  t.wnw = bx * by; t.wne = ax * by; t.wsw = bx * ay; t.wse = ax * ay;
  // clamp before the int conversion: far out-of-range / non-finite coordinates only need to fail
  // the bounds test
  const float cx = fminf(fmaxf(fx, -2.f), static_cast<float>(w) + 1.f);
  const float cy = fminf(fmaxf(fy, -2.f), static_cast<float>(h) + 1.f);
  t.x0 = static_cast<int>(cx); t.y0 = static_cast<int>(cy);
  const bool finite = (fx == fx) && (fy == fy);
  const bool xl = t.x0 >= 0 && t.x0 < w, xr = t.x0 + 1 >= 0 && t.x0 + 1 < w;
  const bool yt = t.y0 >= 0 && t.y0 < h, yb = t.y0 + 1 >= 0 && t.y0 + 1 < h;
  t.ok_nw = finite && xl && yt; t.ok_ne = finite && xr && yt;
  t.ok_sw = finite && xl && yb; t.ok_se = finite && xr && yb;
  return t;
}
__device__ __forceinline__ float gather(const float* __restrict__ plane, int w, const Taps& t) {
  const float* p = plane + static_cast<long long>(t.y0) * w + t.x0;
  float acc = 0.f;
  if (t.ok_nw) acc += __ldg(p) * t.wnw;
  if (t.ok_ne) acc += __ldg(p + 1) * t.wne;
  if (t.ok_sw) acc += __ldg(p + w) * t.wsw;
  if (t.ok_se) acc += __ldg(p + w + 1) * t.wse;
  return acc;
}

__global__ void warp_kernel(const float* __restrict__ img, const float2* __restrict__ grid,
                            float* __restrict__ out, int n, int c, int h, int w, int ho, int wo) {
  const long long total = static_cast<long long>(n) * ho * wo;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / (static_cast<long long>(ho) * wo);
    const long long pix = i - b * static_cast<long long>(ho) * wo;
    const float2 g = __ldg(grid + i);
    const Taps t = make_taps(round_fp16(g.x), round_fp16(g.y), h, w);
    for (int ch = 0; ch < c; ++ch) {
      const float* plane = img + (b * c + ch) * static_cast<long long>(h) * w;
      out[(b * c + ch) * static_cast<long long>(ho) * wo + pix] = gather(plane, w, t);
    }
  }
}

// Four consecutive output pixels per thread: the grid arrives as two 16-byte loads, every plane leaves as one 16-byte
// store, and a thread keeps 16 gathers per plane in flight.  Needs wo % 4 == 0 and 16-byte aligned grid / out rows.
__global__ void __launch_bounds__(256, 4)   // 64 registers -> 4 CTAs per SM: 180 -> 160 us at 4K (5 CTAs: 48 registers, spills, slower)
warp4_kernel(const float* __restrict__ img, const float4* __restrict__ grid, float* __restrict__ out, int n, int c, int h,
             int w, int ho, int wo) {
  const long long plane_o = static_cast<long long>(ho) * wo;
  const long long total = static_cast<long long>(n) * plane_o / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p0 = i * 4;                               // first of the four pixels (never straddles an image: wo % 4 == 0)
    const long long b = p0 / plane_o;
    const long long pix = p0 - b * plane_o;
    const float4 g0 = __ldg(grid + 2 * i), g1 = __ldg(grid + 2 * i + 1);
    Taps t[4];
    t[0] = make_taps(round_fp16(g0.x), round_fp16(g0.y), h, w);
    t[1] = make_taps(round_fp16(g0.z), round_fp16(g0.w), h, w);
    t[2] = make_taps(round_fp16(g1.x), round_fp16(g1.y), h, w);
    t[3] = make_taps(round_fp16(g1.z), round_fp16(g1.w), h, w);
    for (int ch = 0; ch < c; ++ch) {
      const float* plane = img + (b * c + ch) * static_cast<long long>(h) * w;
      float4 v;
      v.x = gather(plane, w, t[0]); v.y = gather(plane, w, t[1]); v.z = gather(plane, w, t[2]); v.w = gather(plane, w, t[3]);
      *reinterpret_cast<float4*>(out + (b * c + ch) * plane_o + pix) = v;
    }
  }
}

// Branch-free form of the same four-pixel kernel for three planes and 32-bit plane indices (every frame this path sees):
// an out-of-range tap keeps a clamped, valid address and gets weight 0 instead of a predicate, the image index is
// blockIdx.y (no 64-bit division per thread) and the plane loop is unrolled.  The kernel is bound by issued instructions,
// not by L1 sectors or DRAM (profiles/r02_summary.md): this halves them.  Same products in the same order as gather().
struct TapsB {
  int o_nw, o_ne, o_sw, o_se;
  float wnw, wne, wsw, wse;
};
__device__ __forceinline__ TapsB make_taps_b(float gx, float gy, int h, int w) {
  TapsB t;
  const float ix = ((gx + 1.f) * w - 1.f) / 2.f;
  const float iy = ((gy + 1.f) * h - 1.f) / 2.f;
  const float fx = floorf(ix), fy = floorf(iy);
  const float ax = ix - fx, ay = iy - fy;
  const float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
  const float cx = fminf(fmaxf(fx, -2.f), static_cast<float>(w) + 1.f);
  const float cy = fminf(fmaxf(fy, -2.f), static_cast<float>(h) + 1.f);
  const int x0 = static_cast<int>(cx), y0 = static_cast<int>(cy);
  const bool finite = (fx == fx) && (fy == fy);
  const bool xl = static_cast<unsigned>(x0) < static_cast<unsigned>(w), xr = static_cast<unsigned>(x0 + 1) < static_cast<unsigned>(w);
  const bool yt = finite && static_cast<unsigned>(y0) < static_cast<unsigned>(h);
  const bool yb = finite && static_cast<unsigned>(y0 + 1) < static_cast<unsigned>(h);
  t.wnw = (xl && yt) ? bx * by : 0.f; t.wne = (xr && yt) ? ax * by : 0.f;
  t.wsw = (xl && yb) ? bx * ay : 0.f; t.wse = (xr && yb) ? ax * ay : 0.f;
  const int xa = min(max(x0, 0), w - 1), xb = min(max(x0 + 1, 0), w - 1);
  const int ra = min(max(y0, 0), h - 1) * w, rb = min(max(y0 + 1, 0), h - 1) * w;
  t.o_nw = ra + xa; t.o_ne = ra + xb; t.o_sw = rb + xa; t.o_se = rb + xb;
  return t;
}
__device__ __forceinline__ float gather_b(const float* __restrict__ plane, const TapsB& t) {
  float acc = __ldg(plane + t.o_nw) * t.wnw;
  acc = fmaf(__ldg(plane + t.o_ne), t.wne, acc);
  acc = fmaf(__ldg(plane + t.o_sw), t.wsw, acc);
  acc = fmaf(__ldg(plane + t.o_se), t.wse, acc);
  return acc;
}
// A warp owns kWarpPx * 32 consecutive pixels; its k-th instruction handles pixels k*32 + lane, so every tap load of a
// smooth field and every store touches one or two 128-byte lines (four CONSECUTIVE pixels per thread made each load a
// 16-byte-strided access: 3.3 L1 wavefronts per request, the L1 data pipe at 85 % - profiles/r02_summary.md).
#ifndef TG_WARP_PX
#define TG_WARP_PX 4
#endif
#ifndef TG_WARP_MINB
#define TG_WARP_MINB 1
#endif
constexpr int kWarpPx = TG_WARP_PX;
__global__ void __launch_bounds__(256, TG_WARP_MINB)
warp4c3_kernel(const float* __restrict__ img, const float2* __restrict__ grid, float* __restrict__ out, int h, int w, int ho,
               int wo) {
  const int plane_i = h * w, plane_o = ho * wo, chunks = (plane_o + 32 * kWarpPx - 1) / (32 * kWarpPx);
  const float* src = img + static_cast<long long>(blockIdx.y) * 3 * plane_i;
  float* dst = out + static_cast<long long>(blockIdx.y) * 3 * plane_o;
  const float2* g = grid + static_cast<long long>(blockIdx.y) * plane_o;
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int ck = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ck < chunks; ck += warps) {
    const int p0 = ck * (32 * kWarpPx) + lane;
    TapsB t[kWarpPx];
#pragma unroll
    for (int k = 0; k < kWarpPx; ++k) {
      const float2 gk = __ldg(g + min(p0 + 32 * k, plane_o - 1));
      t[k] = make_taps_b(round_fp16(gk.x), round_fp16(gk.y), h, w);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* plane = src + ch * plane_i;
      // keep the plane base in a vector register pair: each tap address is then ONE IMAD.WIDE (offset * 4 + base); with the
      // (CTA-uniform) base in uniform registers the compiler emits a LEA / LEA.HI pair per tap instead
      asm volatile("" : "+l"(plane));
      float v[kWarpPx];
#pragma unroll
      for (int k = 0; k < kWarpPx; ++k) v[k] = gather_b(plane, t[k]);
#pragma unroll
      for (int k = 0; k < kWarpPx; ++k)
        if (p0 + 32 * k < plane_o) dst[ch * plane_o + p0 + 32 * k] = v[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused producer of the spatio-temporal discriminator's input (reference code/train.py:139-198, Dt_mergeDs=True):
// out[s, 0:9]   = before9[s]                                  the three target frames of triplet s        (:175)
// out[s, 9:18]  = crop_pad(grid_sample(src frames 3s..3s+2, T_vel))   centre window kept, border zeroed  (:165-174,187-196)
// out[s, 18:27] = bilinear x4 of the three LR frames                                                     (:176-178)
// T_vel of frame m (class m % 3): 0 -> up4(4*gsrc[m]) (the forward 'flow'), 1 -> zeros, 2 -> 2*up4(4*gsrc[m]) - 1
// (preprocess of the backward 'flow', :149; with pingpang the reference takes the flipped forward flow WITHOUT
// preprocess, :155 -> flags bit 1), each [2,Ho,Wo] block re-viewed as [Ho,Wo,2] exactly as the reference's reshape
// does (:156-157).  The velocity field is computed on the fly from the LR planes; nothing but `out` is written.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
disc_input_kernel(const float* __restrict__ before9, const float* __restrict__ src, long long src_stride_b,
                  long long src_stride_t, int ts, const float* __restrict__ gsrc, const float* __restrict__ lr9,
                  float* __restrict__ out, int tb, int h, int w, int crop_off, int flags) {
  const bool grid_fp16 = (flags & 1) != 0, raw_next = (flags & 2) != 0;
  const int ho = 4 * h, wo = 4 * w;
  const long long hw = static_cast<long long>(ho) * wo;
  const long long total = static_cast<long long>(tb) * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(i / hw);
    const long long pix = i - s * hw;
    const int y = static_cast<int>(pix / wo), x = static_cast<int>(pix - static_cast<long long>(y) * wo);
    float* o = out + static_cast<long long>(s) * 27 * hw + pix;
    const float* bf = before9 + static_cast<long long>(s) * 9 * hw + pix;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k * hw] = __ldg(bf + k * hw);
    const bool inside = y >= crop_off && y < ho - crop_off && x >= crop_off && x < wo - crop_off;
    for (int j = 0; j < 3; ++j) {
      const int m = s * 3 + j;
      float v[3] = {0.f, 0.f, 0.f};
      if (inside) {
        float g[2] = {0.f, 0.f};
        const int cls = m % 3;
        if (cls != 1) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const long long flat = pix * 2 + k;                       // index inside the [2,Ho,Wo] block
            const int ch = static_cast<int>(flat / hw);
            const long long rem = flat - ch * hw;
            const int yy = static_cast<int>(rem / wo), xx = static_cast<int>(rem - static_cast<long long>(yy) * wo);
            float u = up4_sample(gsrc + (static_cast<long long>(m) * 2 + ch) * h * w, h, w, yy, xx, 4.f);
            if (cls == 2 && !raw_next) u = __fsub_rn(__fmul_rn(u, 2.f), 1.f);    // preprocess(): image * 2 - 1, no fused rounding
            g[k] = grid_fp16 ? round_fp16(u) : u;
          }
        }
        const Taps t = make_taps(g[0], g[1], ho, wo);
        const float* frame = src + (m / ts) * src_stride_b + (m % ts) * src_stride_t;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = gather(frame + c * hw, wo, t);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) o[(9 + j * 3 + c) * hw] = v[c];
    }
    const float* lp = lr9 + static_cast<long long>(s) * 9 * h * w;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[(18 + k) * hw] = up4_sample(lp + static_cast<long long>(k) * h * w, h, w, y, x, 1.f);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused producer of the generator input (reference main.py:186-213 as one pass):
// one CTA = 8x8 LR pixels = 32x32 HR pixels.  Each thread warps 4 HR pixels (flow computed on the
// fly from LR_{t-1}, rounded to fp16 as the reference does), applies (v+1)/2, and drops the
// results at their space-to-depth channel of an smem tile [64 LR px][64 ch] bf16, which is then
// written out as whole 128-byte NHWC pixel rows.
// ---------------------------------------------------------------------------------------------
constexpr int kFT = 8;   // LR tile edge

// prev_rgbx (optional): the same previous HR estimate as pixel-interleaved float4 {R,G,B,-} [n][4h][4w] (written by
// the frame kernel's output conv next to the planar tensor).  The sample positions of the reference's "flow" are
// scattered, so the gather is bound by L1 sectors touched per warp instruction: one 16-byte tap for three channels
// instead of three 4-byte ones.  Same taps, same weights, same order of additions: bit-identical.
__device__ __forceinline__ void gather3(const float4* __restrict__ img, int w, const Taps& t, float (&v)[3]) {
  const float4* p = img + static_cast<long long>(t.y0) * w + t.x0;
  v[0] = v[1] = v[2] = 0.f;
  if (t.ok_nw) { const float4 a = __ldg(p); v[0] += a.x * t.wnw; v[1] += a.y * t.wnw; v[2] += a.z * t.wnw; }
  if (t.ok_ne) { const float4 a = __ldg(p + 1); v[0] += a.x * t.wne; v[1] += a.y * t.wne; v[2] += a.z * t.wne; }
  if (t.ok_sw) { const float4 a = __ldg(p + w); v[0] += a.x * t.wsw; v[1] += a.y * t.wsw; v[2] += a.z * t.wsw; }
  if (t.ok_se) { const float4 a = __ldg(p + w + 1); v[0] += a.x * t.wse; v[1] += a.y * t.wse; v[2] += a.z * t.wse; }
}

// Branch-free forms for the per-pixel loop below (it is bound by issued instructions): 16-byte taps at clamped addresses
// with zeroed weights (make_taps_b), and the two flow components of one pixel - columns col, col + 1 of one row of the
// x4-upscaled plane - sharing the row part of up4_sample (same arithmetic per value).
__device__ __forceinline__ void gather3_b(const float4* __restrict__ img, const TapsB& t, float (&v)[3]) {
  const float4 a = __ldg(img + t.o_nw), b = __ldg(img + t.o_ne), c = __ldg(img + t.o_sw), d = __ldg(img + t.o_se);
  v[0] = a.x * t.wnw; v[1] = a.y * t.wnw; v[2] = a.z * t.wnw;
  v[0] = fmaf(b.x, t.wne, v[0]); v[1] = fmaf(b.y, t.wne, v[1]); v[2] = fmaf(b.z, t.wne, v[2]);
  v[0] = fmaf(c.x, t.wsw, v[0]); v[1] = fmaf(c.y, t.wsw, v[1]); v[2] = fmaf(c.z, t.wsw, v[2]);
  v[0] = fmaf(d.x, t.wse, v[0]); v[1] = fmaf(d.y, t.wse, v[1]); v[2] = fmaf(d.z, t.wse, v[2]);
}
__device__ __forceinline__ void up4_pair(const float* __restrict__ plane, int h, int w, int row, int col, float pre, float& u0,
                                         float& u1) {
  float sy = (row + 0.5f) * 0.25f - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  const int y0 = min(static_cast<int>(sy), h - 1), y1 = min(y0 + 1, h - 1);
  const float ly1 = sy - y0, ly0 = 1.f - ly1;
  const float* r0 = plane + y0 * w;
  const float* r1 = plane + y1 * w;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float sx = (col + k + 0.5f) * 0.25f - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    const int x0 = min(static_cast<int>(sx), w - 1), x1 = min(x0 + 1, w - 1);
    const float lx1 = sx - x0, lx0 = 1.f - lx1;
    const float p00 = __ldg(r0 + x0) * pre, p01 = __ldg(r0 + x1) * pre;
    const float p10 = __ldg(r1 + x0) * pre, p11 = __ldg(r1 + x1) * pre;
    const float u = ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
    if (k == 0) u0 = u; else u1 = u;
  }
}

__global__ void __launch_bounds__(256)
fused_input_kernel(const float* __restrict__ lr_t, const float* __restrict__ lr_prev,
                   const float* __restrict__ prev_hr, const float4* __restrict__ prev_rgbx,
                   __nv_bfloat16* __restrict__ x, int n, int h, int w,
                   long long lr_bs, long long hr_bs, uint32_t* __restrict__ zero, size_t zero_count, FastDiv fd_wo) {
  __shared__ __align__(16) __nv_bfloat16 tile[kFT * kFT][64];
  // clear the frame kernel's per-item completion counters (this kernel runs between two frame kernels)
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < zero_count;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    zero[i] = 0u;
  const int tiles_x = (w + kFT - 1) / kFT, tiles_y = (h + kFT - 1) / kFT;
  const int ho = 4 * h, wo = 4 * w;
  const uint32_t hw_o = static_cast<uint32_t>(ho) * static_cast<uint32_t>(wo);     // 2 * hw_o < 2^31 (host-checked)
  const int ntiles = n * tiles_x * tiles_y;
  for (int tIdx = blockIdx.x; tIdx < ntiles; tIdx += gridDim.x) {
    const int tx = tIdx % tiles_x;
    const int r = tIdx / tiles_x;
    const int ty = r % tiles_y;
    const int b = r / tiles_y;
    const int lx0 = tx * kFT, ly0 = ty * kFT;
    // channels 0..2 (LR frame t) and 51..63 (zero padding)
    for (int i = threadIdx.x; i < kFT * kFT * 16; i += blockDim.x) {
      const int px = i >> 4, k = i & 15;                    // k: 0..2 LR, 3..15 -> zero pad 51..63
      const int ly = ly0 + px / kFT, lx = lx0 + px % kFT;
      if (k < 3) {
        float v = 0.f;
        if (ly < h && lx < w) v = __ldg(lr_t + b * lr_bs + (static_cast<long long>(k) * h + ly) * w + lx);
        tile[px][k] = __float2bfloat16_rn(v);
      } else {
        tile[px][48 + k] = __float2bfloat16_rn(0.f);
      }
    }
    // channels 3..50: space_to_depth(deprocess(warp(prev_hr)))
    const float* flow = lr_prev != nullptr ? lr_prev + b * lr_bs : nullptr;
    const float4* rgbx = prev_rgbx != nullptr ? prev_rgbx + static_cast<long long>(b) * hw_o : nullptr;
    for (int i = threadIdx.x; i < (4 * kFT) * (4 * kFT); i += blockDim.x) {
      const int hy = i / (4 * kFT), hx = i % (4 * kFT);
      const int oy = 4 * ly0 + hy, ox = 4 * lx0 + hx;
      float v[3] = {0.f, 0.f, 0.f};
      if (prev_hr != nullptr && oy < ho && ox < wo) {
        // grid element (oy,ox,:) = two consecutive floats of the [2,Ho,Wo] planar flow buffer (wo is even: same row)
        const uint32_t f = (static_cast<uint32_t>(oy) * static_cast<uint32_t>(wo) + static_cast<uint32_t>(ox)) * 2u;
        const uint32_t plane = f >= hw_o ? 1u : 0u;
        const uint32_t rem = f - plane * hw_o;
        const uint32_t row = fdiv(rem, fd_wo), col = rem - row * static_cast<uint32_t>(wo);
        float gx, gy;
        up4_pair(flow + plane * static_cast<uint32_t>(h * w), h, w, static_cast<int>(row), static_cast<int>(col), 4.f, gx, gy);
        gx = round_fp16(gx); gy = round_fp16(gy);
        if (rgbx != nullptr) {
          const TapsB t = make_taps_b(gx, gy, ho, wo);
          gather3_b(rgbx, t, v);
        } else {
          const Taps t = make_taps(gx, gy, ho, wo);
          const float* img = prev_hr + b * hr_bs;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) v[ch] = gather(img + ch * static_cast<long long>(hw_o), wo, t);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = (v[ch] + 1.f) / 2.f;
      }
      const int px = (hy >> 2) * kFT + (hx >> 2);
      const int sub = (hy & 3) * 4 + (hx & 3);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) tile[px][3 + ch * 16 + sub] = __float2bfloat16_rn(v[ch]);
    }
    __syncthreads();
    // write whole pixels: 64 px x 128 B, 8 x 16B chunks per pixel
    for (int i = threadIdx.x; i < kFT * kFT * 8; i += blockDim.x) {
      const int px = i >> 3, ck = i & 7;
      const int ly = ly0 + px / kFT, lx = lx0 + px % kFT;
      if (ly < h && lx < w) {
        const uint4 vv = reinterpret_cast<const uint4*>(&tile[px][0])[ck];
        reinterpret_cast<uint4*>(x + ((static_cast<long long>(b) * h + ly) * w + lx) * 64)[ck] = vv;
      }
    }
    __syncthreads();
  }
}

// NCHW f32 [n,c,h,w] -> NHWC bf16 [n,h,w,64] (zero padded).  One CTA = 32 consecutive pixels.
__global__ void __launch_bounds__(256)
pack_nhwc64_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int c,
                   long long hw) {
  // 128 pixels per iteration: c x 128 coalesced plane loads in flight per CTA (32 pixels per iteration left the kernel
  // latency-bound at 2.4 TB/s), transposed through shared memory, written as whole 128-byte pixels
  constexpr int kPx = 128;
  __shared__ __align__(16) __nv_bfloat16 tile[kPx][64 + 8];
  const long long groups = (hw + kPx - 1) / kPx;
  for (long long gi = blockIdx.x; gi < groups * n; gi += gridDim.x) {
    const long long b = gi / groups;
    const long long p0 = (gi % groups) * kPx;
    for (int i = threadIdx.x; i < 64 * kPx; i += blockDim.x) {
      const int ch = i / kPx, px = i % kPx;
      float v = 0.f;
      if (ch < c && p0 + px < hw) v = __ldg(in + (b * c + ch) * hw + p0 + px);
      tile[px][ch] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPx * 8; i += blockDim.x) {
      const int px = i >> 3, ck = i & 7;
      if (p0 + px < hw) {
        const uint4 vv = *reinterpret_cast<const uint4*>(&tile[px][ck * 8]);
        reinterpret_cast<uint4*>(out + (b * hw + p0 + px) * 64)[ck] = vv;
      }
    }
    __syncthreads();
  }
}

static int grid_for(long long work_items, int threads, int per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(tg_num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace tg

using namespace tg;

extern "C" int tg_space_to_depth(const void* in, void* out, int n, int c, int h_out, int w_out, int r,
                                 void* stream) {
  TG_CHECK_ARG(in && out, "space_to_depth: null pointer");
  TG_CHECK_ARG(n >= 0 && c >= 0 && h_out >= 0 && w_out >= 0 && r >= 1, "space_to_depth: bad shape");
  const long long planes = static_cast<long long>(n) * c;
  if (planes == 0 || h_out == 0 || w_out == 0) return TG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tg_prof_pre(TG_K_GLUE, 8.0 * planes * r * r * h_out * w_out, st);   // bytes: 4 in + 4 out per element
  if (r == 4 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    const long long work = planes * 4 * h_out * w_out;
    s2d4_kernel<<<grid_for(work, 256, 16), 256, 0, st>>>(static_cast<const uint4*>(in), static_cast<uint32_t*>(out),
                                                        static_cast<int>(planes), h_out, w_out);
  } else {
    const long long work = planes * r * r * h_out * w_out;
    s2d_generic_kernel<<<grid_for(work, 256, 16), 256, 0, st>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out),
                                                               static_cast<int>(planes), h_out, w_out, r, 1);
  }
  tg_prof_post(static_cast<cudaStream_t>(stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

extern "C" int tg_depth_to_space(const void* in, void* out, int n, int c, int h_in, int w_in, int r, void* stream) {
  TG_CHECK_ARG(in && out, "depth_to_space: null pointer");
  TG_CHECK_ARG(n >= 0 && c >= 0 && h_in >= 0 && w_in >= 0 && r >= 1, "depth_to_space: bad shape");
  const long long planes = static_cast<long long>(n) * c;
  if (planes == 0 || h_in == 0 || w_in == 0) return TG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tg_prof_pre(TG_K_GLUE, 8.0 * planes * r * r * h_in * w_in, st);
  if (r == 4 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const long long work = planes * 4 * h_in * w_in;
    d2s4_kernel<<<grid_for(work, 256, 16), 256, 0, st>>>(static_cast<const uint32_t*>(in), static_cast<uint4*>(out),
                                                        static_cast<int>(planes), h_in, w_in);
  } else {
    const long long work = planes * r * r * h_in * w_in;
    s2d_generic_kernel<<<grid_for(work, 256, 16), 256, 0, st>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out),
                                                               static_cast<int>(planes), h_in, w_in, r, 0);
  }
  tg_prof_post(static_cast<cudaStream_t>(stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

extern "C" int tg_warp_bilinear(const float* img, const float* grid, float* out, int n, int c, int h, int w, int ho,
                                int wo, void* stream) {
  TG_CHECK_ARG(img && grid && out, "warp_bilinear: null pointer");
  TG_CHECK_ARG(n >= 0 && c >= 1 && h >= 1 && w >= 1 && ho >= 0 && wo >= 0, "warp_bilinear: bad shape");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(grid) & 7) == 0, "warp_bilinear: grid must be 8-byte aligned");
  const long long work = static_cast<long long>(n) * ho * wo;
  if (work == 0) return TG_OK;
  tg_prof_pre(TG_K_GLUE, (8.0 * c + 4.0) * n * ho * wo, static_cast<cudaStream_t>(stream));   // f32 img in/out + fp16 grid
  static const bool branch_free = []() { const char* e = getenv("TG_WARP_BRANCHFREE"); return !(e && e[0] == '0'); }();   // A/B knob
  const bool vec4 = (wo & 3) == 0 && ((reinterpret_cast<uintptr_t>(grid) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (branch_free && c == 3 && n <= 65535 && 3LL * h * w < (1LL << 31) && 3LL * ho * wo < (1LL << 31) - 1024) {
    const int chunks = (ho * wo + 32 * kWarpPx - 1) / (32 * kWarpPx);   // one warp per chunk, 8 warps per CTA
    int bx = (chunks + 7) / 8;
    const int cap = (tg_num_sms() * 16 + n - 1) / n;          // ~16 CTAs per SM over all images
    if (bx > cap) bx = cap;
    warp4c3_kernel<<<dim3(bx, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, reinterpret_cast<const float2*>(grid), out, h, w,
                                                                              ho, wo);
  } else if (vec4)
    warp4_kernel<<<grid_for(work / 4, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        img, reinterpret_cast<const float4*>(grid), out, n, c, h, w, ho, wo);
  else
    warp_kernel<<<grid_for(work, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        img, reinterpret_cast<const float2*>(grid), out, n, c, h, w, ho, wo);
  tg_prof_post(static_cast<cudaStream_t>(stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

extern "C" int tg_upscale4_bilinear(const float* in, float* out, int n, int c, int h, int w, float pre_scale,
                                    void* stream) {
  TG_CHECK_ARG(in && out, "upscale4: null pointer");
  TG_CHECK_ARG(n >= 0 && c >= 0 && h >= 1 && w >= 1, "upscale4: bad shape");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "upscale4: out must be 16-byte aligned");
  const long long planes = static_cast<long long>(n) * c;
  if (planes == 0) return TG_OK;
  const long long work = planes * 4 * h * w;                 // one thread per four outputs
  tg_prof_pre(TG_K_GLUE, 4.0 * planes * h * w * 17.0, static_cast<cudaStream_t>(stream));
  upscale4_kernel<<<grid_for(work, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, static_cast<int>(planes),
                                                                                       h, w, pre_scale);
  tg_prof_post(static_cast<cudaStream_t>(stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int tg::fused_input_launch(const float* lr_t, const float* lr_prev, const float* prev_hr, void* x_nhwc, int n, int h,
                           int w, long long lr_batch_stride, long long hr_batch_stride, uint32_t* zero,
                           size_t zero_count, cudaStream_t stream, const void* prev_rgbx) {
  TG_CHECK_ARG(lr_t && x_nhwc, "fused_warp_s2d_concat: null pointer");
  TG_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "fused_warp_s2d_concat: bad shape");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(x_nhwc) & 15) == 0, "fused_warp_s2d_concat: x must be 16-byte aligned");
  TG_CHECK_ARG(32LL * h * w < (1LL << 31), "fused_warp_s2d_concat: frame too large for 32-bit pixel indices (%d x %d)", h, w);
  if (!lr_prev || !prev_hr) { lr_prev = nullptr; prev_hr = nullptr; }
  const int tiles = n * tg_div_up(w, kFT) * tg_div_up(h, kFT);
  int blocks = tiles < tg_num_sms() * 8 ? tiles : tg_num_sms() * 8;
  // algorithmic bytes, SURVEY.md 8(d): 12 B (prev HR f32) + 6.4 B (51 bf16 channels / 16) per HR pixel; the flow
  // is computed on the fly from the LR frame, so the 4 B/px grid read does not exist
  tg_prof_pre(TG_K_FUSED_INPUT, 18.4 * 16.0 * n * h * w, stream);
  if (!prev_hr) prev_rgbx = nullptr;
  fused_input_kernel<<<blocks, 256, 0, stream>>>(lr_t, lr_prev, prev_hr, static_cast<const float4*>(prev_rgbx),
                                                 static_cast<__nv_bfloat16*>(x_nhwc), n, h, w, lr_batch_stride, hr_batch_stride,
                                                 zero, zero_count, make_fastdiv(static_cast<uint32_t>(4 * w)));
  tg_prof_post(stream);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

extern "C" int tg_fused_warp_s2d_concat(const float* lr_t, const float* lr_prev, const float* prev_hr, void* x_nhwc,
                                        int n, int h, int w, long long lr_batch_stride, long long hr_batch_stride,
                                        void* stream) {
  return tg::fused_input_launch(lr_t, lr_prev, prev_hr, x_nhwc, n, h, w, lr_batch_stride, hr_batch_stride, nullptr, 0,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int tg_pack_nchw_to_nhwc64(const float* in, void* out, int n, int c, int h, int w, void* stream) {
  TG_CHECK_ARG(in && out, "pack_nchw_to_nhwc64: null pointer");
  TG_CHECK_ARG(n >= 1 && c >= 1 && c <= 64 && h >= 1 && w >= 1, "pack_nchw_to_nhwc64: bad shape (c must be <= 64)");
  const long long hw = static_cast<long long>(h) * w;
  const long long groups = (hw + 127) / 128 * n;             // 128 pixels per CTA iteration (pack_nhwc64_kernel)
  const long long cap = static_cast<long long>(tg_num_sms()) * 8;
  tg_prof_pre(TG_K_GLUE, (4.0 * c + 128.0) * n * hw, static_cast<cudaStream_t>(stream));
  pack_nhwc64_kernel<<<static_cast<int>(groups < cap ? groups : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, static_cast<__nv_bfloat16*>(out), n, c, hw);
  tg_prof_post(static_cast<cudaStream_t>(stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

// ---------------------------------------------------------------------------------------------
// Backward of the generator's final sigmoid (code/models.py:86), fused with the layout change the data-gradient
// convolution needs: dz = dout * out * (1 - out), NCHW f32 [n,3,hw] -> NHWC bf16 [n,hw,64] (channels 3..63 zero).
// ---------------------------------------------------------------------------------------------
namespace tg {
__global__ void __launch_bounds__(256)
sigmoid_bwd_pack_kernel(const float* __restrict__ dout, const float* __restrict__ out, __nv_bfloat16* __restrict__ dz,
                        int n, long long hw, long long dout_nstride, long long out_nstride) {
  // A warp owns 32 consecutive pixels: every lane computes ITS pixel's three gradients (coalesced plane loads), then the
  // warp writes the 32 x 128-byte rows four pixels per instruction, eight lanes per pixel (whole 128-byte lines; one
  // 16-byte store per lane at a 128-byte stride wrote half sectors and ran at 2 TB/s).
  const long long total = static_cast<long long>(n) * hw;
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long w0 = ((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5) * 32; w0 < total; w0 += warps * 32) {
    const long long i = w0 + lane;
    uint32_t lo = 0u, hi = 0u;
    if (i < total) {
      const long long b = i / hw, px = i - b * hw;
      float g[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float y = __ldg(out + b * out_nstride + c * hw + px);
        g[c] = __ldg(dout + b * dout_nstride + c * hw + px) * y * (1.f - y);
      }
      lo = pack_bf16x2(g[0], g[1]);
      hi = pack_bf16x2(g[2], 0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int src = k * 4 + (lane >> 3);                     // pixel of this lane's 16-byte chunk in round k
      const uint32_t a = __shfl_sync(0xFFFFFFFFu, lo, src), b2 = __shfl_sync(0xFFFFFFFFu, hi, src);
      const long long p = w0 + src;
      if (p < total)
        reinterpret_cast<uint4*>(dz + p * 64)[lane & 7] = (lane & 7) == 0 ? make_uint4(a, b2, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

int sigmoid_bwd_pack_launch(const float* dout, const float* out, void* dz, int n, long long hw, long long dout_nstride,
                            long long out_nstride, cudaStream_t st) {
  const long long total = static_cast<long long>(n) * hw;
  tg_prof_pre(TG_K_GLUE, (24.0 + 128.0) * total, st);
  sigmoid_bwd_pack_kernel<<<grid_for(total, 256, 16), 256, 0, st>>>(dout, out, static_cast<__nv_bfloat16*>(dz), n, hw,
                                                                    dout_nstride, out_nstride);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}
}  // namespace tg

extern "C" int tg_disc_input_assemble(const float* before9, const float* src, long long src_stride_b,
                                      long long src_stride_t, int ts, const float* gsrc, const float* lr9, float* out,
                                      int tb, int h, int w, int crop_off, int grid_fp16, void* stream) {
  TG_CHECK_ARG(before9 && src && gsrc && lr9 && out, "disc_input_assemble: null pointer");
  TG_CHECK_ARG(tb >= 1 && h >= 1 && w >= 1 && ts >= 3 && ts % 3 == 0, "disc_input_assemble: bad shape");
  TG_CHECK_ARG(crop_off >= 0 && 2 * crop_off < 4 * h && 2 * crop_off < 4 * w, "disc_input_assemble: bad crop offset %d", crop_off);
  const long long work = static_cast<long long>(tb) * 16 * h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // bytes per HR pixel of a triplet: 9 targets + 9 gathered + 27 written (f32)
  tg_prof_pre(TG_K_GLUE, 4.0 * 45.0 * work, st);
  tg::disc_input_kernel<<<grid_for(work, 256, 16), 256, 0, st>>>(before9, src, src_stride_b, src_stride_t, ts, gsrc, lr9, out, tb,
                                                                h, w, crop_off, grid_fp16);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}
