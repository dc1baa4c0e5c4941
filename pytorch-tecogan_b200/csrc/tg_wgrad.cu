// Weight gradient of the 3x3 stride-1 convolution on tcgen05 (backward of nn.Conv2d in the generator /
// discriminator, reference code/train.py:336,340 via autograd):
//     dW[co][ci][ky][kx] = sum over (n, y, x) of dY[n, y, x, co] * X[n, y+ky-1, x+kx-1, ci]
// GEMM view: the reduction dimension K is the PIXEL index, M = co, N = ci.  Activations are NHWC, i.e. the
// channel is the contiguous dimension of both operands: both are "MN-major" UMMA operands.  A TMA box
// {64 channels, pixels...} lands in shared memory as one 128-byte row per pixel = exactly the MN-major
// SWIZZLE_128B canonical layout ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units:
//   * 8 consecutive pixels (one tile row) form one 1024-byte K group; the next tile row is SBO bytes further
//     (1024 for the dY tile, 1280 for the X box with its 1-pixel halo in x);
//   * the next 64 channels (second box of a 128-channel tensor) are LBO bytes further.
// The filter-tap shift (ky, kx) is a shift of X in pixels = a shift of the descriptor start address by whole
// 128-byte rows; for 64 input channels the three kx taps are 128 bytes apart and are issued as ONE MMA with
// N = 192 (three "channel blocks" with LBO = 128).
// Work split: grid = (pixel slabs, 3 ky).  A CTA accumulates its slab in TMEM (<= 384 fp32 columns) and adds the
// result into dW with fp32 atomics (split-K over slabs).
#include <stdlib.h>

#include "tg_conv_tc.cuh"

namespace tg {

constexpr int kWgThreads = 192;
constexpr uint32_t kWgSmemLimit = 232448;

struct WgTap { uint32_t b_off; int ky, kx; };   // B view offset inside the staged box; filter tap it belongs to
struct WgGroup {                                // one blockIdx.y: the taps that share one staged B box
  int b_ox, b_oy;                               // box origin relative to (in_scale*x0, in_scale*y0)
  int ntaps;
  WgTap taps[4];
};
struct WgParams {
  int n, h, w, tiles_x, tiles_y, num_items;   // tile space = the resolution of the A operand
  int rows_real, cols_real;           // real channel counts of the A (GEMM M) and B (GEMM N) tensors
  int ks;                             // filter size (3 or 4): dw index = ((m * cols_real + c) * ks + ky) * ks + kx
  int a_boxes, b_boxes;               // padded channels / 64
  int b_in_scale;                     // 1, or 2: B lives at twice the resolution and is staged with element stride 2
  int stack3;                         // conv3x3 with 64 B channels: the three kx taps as one N = 192 MMA
  uint32_t a_bytes, b_bytes, b_pitch, stage_stride;
  int nstages;
  WgGroup groups[4];
  float* dw;                          // f32 weight gradient in PyTorch layout, accumulated into
};

// MN-major SWIZZLE_128B descriptor: [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | version 1 | layout 2
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor, bf16 x bf16 -> f32, both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t s_st = base;
  const uint32_t bar_full = base + p.nstages * p.stage_stride;
  const uint32_t bar_empty = bar_full + 8 * p.nstages;
  const uint32_t bar_done = bar_empty + 8 * p.nstages;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gbase + (bar_done + 8 - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WgGroup& grp = p.groups[blockIdx.y];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < p.nstages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ---------------- TMA producer: dY tile (no halo) + X rows shifted by ky-1 with a 1-pixel halo in x
    int s = 0;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const int tx = it % p.tiles_x;
      const int r = it / p.tiles_x;
      const int ty = r % p.tiles_y;
      const int n = r / p.tiles_y;
      const int x0 = tx * kTileW, y0 = ty * kTileH;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(bar_full + 8 * s, p.a_boxes * p.a_bytes + p.b_boxes * p.b_bytes);
        const uint32_t dst = s_st + s * p.stage_stride;
        for (int c = 0; c < p.a_boxes; ++c)
          tma_load_4d(dst + c * p.a_bytes, &tm_dy, bar_full + 8 * s, c * 64, x0, y0, n);
        for (int c = 0; c < p.b_boxes; ++c)
          tma_load_4d(dst + p.a_boxes * p.a_bytes + c * p.b_bytes, &tm_x, bar_full + 8 * s, c * 64,
                      x0 * p.b_in_scale + grp.b_ox, y0 * p.b_in_scale + grp.b_oy, n);
      }
      __syncwarp();
      if (++s == p.nstages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: D[co][(kx, ci)] += dY^T[co][pixels] * X[pixels][(kx, ci)]
    const bool stack3 = p.stack3 != 0;
    const int ncols = 64 * p.b_boxes;                        // GEMM N of one tap
    const uint32_t idesc = stack3 ? umma_idesc_bf16_mn(128, 192) : umma_idesc_bf16_mn(128, ncols);
    const uint32_t a_lbo = (p.a_boxes == 2) ? p.a_bytes : 0u;     // 64 output channels: rows 64..127 mirror rows 0..63
    int s = 0;
    uint32_t ph = 0, first = 1;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_base = s_st + s * p.stage_stride;
        const uint32_t b_base = a_base + p.a_boxes * p.a_bytes;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {                         // 16 pixels = tile rows 2j, 2j+1
          const uint64_t ad = umma_desc_mn_sw128(a_base + j * 2048, a_lbo, 1024);
          if (stack3) {
            const uint64_t bd = umma_desc_mn_sw128(b_base + 2 * j * p.b_pitch, 128, p.b_pitch);
            umma_bf16(tmem_base, ad, bd, idesc, (first && j == 0) ? 0u : 1u);
          } else {
            for (int t = 0; t < grp.ntaps; ++t) {
              const uint64_t bd = umma_desc_mn_sw128(b_base + 2 * j * p.b_pitch + grp.taps[t].b_off, p.b_bytes, p.b_pitch);
              umma_bf16(tmem_base + t * ncols, ad, bd, idesc, (first && j == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(bar_empty + 8 * s);
      }
      __syncwarp();
      first = 0;
      if (++s == p.nstages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> fp32 atomics into dW[co][ci][ky][kx]
    const int q = warp & 3;
    const int m = q * 32 + lane;                             // accumulator row = channel of the A tensor
    const bool have_work = blockIdx.x < p.num_items;
    if (have_work) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int cstride = 64 * p.b_boxes;                    // columns per tap
      const int ncol = grp.ntaps * cstride;
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (m < p.rows_real) {
          const int t = c0 / cstride;
          const int ky = grp.taps[t].ky, kx = grp.taps[t].kx;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int c = c0 - t * cstride + e;
            if (c < p.cols_real)
              atomicAdd(p.dw + ((static_cast<size_t>(m) * p.cols_real + c) * p.ks + ky) * p.ks + kx, __uint_as_float(v[e]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// 3x3 weight gradient for layers with <= 64 OUTPUT channels (37 of the generator's 41 convs, half of the discriminator's):
// with M = 128 accumulator rows and only 64 dY channels, the kernel above wastes half of every MMA (rows 64..127 mirror rows
// 0..63) and spends three CTAs (one per filter row) on each pixel tile.  Here the filter-row shift is applied to the dY
// operand instead of X - dW[ky][kx] = sum_q dY[q - (ky-1)] (x) X[q + (0, kx-1)] over the X pixels q of the tile - so that
// two filter rows are the two 64-row halves of ONE MMA (the second half = the same dY box one tile row further: LBO = 1024 B):
//   MMA 1: rows [ky = 2 | ky = 1] x N = 192 (three kx taps stacked, as above)      MMA 2: rows [ky = 0 | mirror] x N = 192
// Two MMAs of 138 cycles per 16 pixels for all nine taps instead of three CTAs x 138: 1.5x the tensor-pipe efficiency
// (1.74x for 128 input channels, where the old path needs three N = 128 MMAs per CTA), and one dY box {64, 8, 18} + one X box
// {64, 10, 16} per tile instead of three of each.  grid = (pixel slabs, cin_pad / 64): a CTA owns one 64-channel block of X.
struct WgKyParams {
  int n, h, w, tiles_x, tiles_y, num_items;
  int rows_real, cols_real;            // cout (<= 64), cin
  int nstages;
  uint32_t stage_stride;
  float* dw;
  float* db;                           // optional bias gradient (db[co] += sum of dY), or null
  // batched form (launch_wgrad3x3_ky_batched): blockIdx.z = layer; the layers' X / dY tensors are consecutive blocks of n
  // images each (image index layer * n + i of one tensor map), their gradients dw_stride / db_stride floats apart
  long long dw_stride, db_stride;
  int co_blocks;                       // 64-row blocks of dY channels (cout_pad / 64): blockIdx.z = layer * co_blocks + block
};
constexpr uint32_t kWgKyABytes = 8 * 18 * 128;        // dY box {64 ch, 8, 18}: the tile and one halo row above / below
constexpr uint32_t kWgKyBBytes = 10 * 16 * 128;       // X box {64 ch, 10, 16}: one halo column left / right

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_ky_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                   const __grid_constant__ WgKyParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t s_st = base;
  const uint32_t bar_full = base + p.nstages * p.stage_stride;
  const uint32_t bar_empty = bar_full + 8 * p.nstages;
  const uint32_t bar_done = bar_empty + 8 * p.nstages;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gbase + (bar_done + 8 - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = blockIdx.y;                               // 64-channel block of X (GEMM N)
  const int layer = blockIdx.z / p.co_blocks, coblk = blockIdx.z - layer * p.co_blocks;   // batched launch: this CTA's layer
  const int n_base = layer * p.n;
  const int rows_real = min(64, p.rows_real - coblk * 64);   // output channels of this CTA's 64-row block of dY
  float* const dw = p.dw + layer * p.dw_stride + static_cast<size_t>(coblk) * 64 * p.cols_real * 9;
  float* const db = p.db ? p.db + layer * p.db_stride + coblk * 64 : nullptr;
  // the bias gradient rides on the first channel block's CTAs: warps 2..5 sum the staged dY tiles while the MMAs run
  const bool do_bias = db != nullptr && cblk == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < p.nstages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, do_bias ? 5 : 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const int tx = it % p.tiles_x;
      const int r = it / p.tiles_x;
      const int ty = r % p.tiles_y;
      const int n = r / p.tiles_y;
      const int x0 = tx * kTileW, y0 = ty * kTileH;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(bar_full + 8 * s, kWgKyABytes + kWgKyBBytes);
        const uint32_t dst = s_st + s * p.stage_stride;
        tma_load_4d(dst, &tm_dy, bar_full + 8 * s, coblk * 64, x0, y0 - 1, n_base + n);
        tma_load_4d(dst + kWgKyABytes, &tm_x, bar_full + 8 * s, cblk * 64, x0 - 1, y0, n_base + n);
      }
      __syncwarp();
      if (++s == p.nstages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16_mn(128, 192);
    int s = 0;
    uint32_t ph = 0, first = 1;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_base = s_st + s * p.stage_stride;
        const uint32_t b_base = a_base + kWgKyABytes;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {                           // X tile rows 2j, 2j+1 = 16 pixels of the reduction
          // dY box row of X row rx for filter row ky: rx + 2 - ky (the box starts one row above the tile)
          const uint64_t a21 = umma_desc_mn_sw128(a_base + (2 * j) * 1024, 1024, 1024);       // rows [ky = 2 | ky = 1]
          const uint64_t a0 = umma_desc_mn_sw128(a_base + (2 * j + 2) * 1024, 0, 1024);       // rows [ky = 0 | mirror]
          const uint64_t bd = umma_desc_mn_sw128(b_base + 2 * j * (10 * 128), 128, 10 * 128); // three kx taps stacked along N
          const uint32_t acc = (first && j == 0) ? 0u : 1u;
          umma_bf16(tmem_base, a21, bd, idesc, acc);
          umma_bf16(tmem_base + 192, a0, bd, idesc, acc);
        }
        umma_commit(bar_empty + 8 * s);
      }
      __syncwarp();
      first = 0;
      if (++s == p.nstages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;                               // accumulator row
    const bool have_work = blockIdx.x < p.num_items;
    if (do_bias && have_work) {
      // bias gradient: thread (warp - 2, lane) owns dY channels 2t, 2t+1 (t = 0..31) of a quarter of the tile's pixels;
      // the tile proper is box rows 1..16 (128 pixels).  Reads race with nothing: the stage is released (bar_empty) only
      // after these four warps arrived as well.
      const int tq = warp - 2, ch2 = lane;                    // pixel quarter, channel pair
      float s0 = 0.f, s1 = 0.f;
      int s = 0;
      uint32_t ph = 0;
      for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        mbar_wait(bar_full + 8 * s, ph);
        const uint8_t* tile = gbase + (s_st - base) + s * p.stage_stride + 8 * 128;   // skip the halo row
#pragma unroll 4
        for (int px = tq * 32; px < tq * 32 + 32; ++px) {
          // SWIZZLE_128B: 16-byte chunk c of 128-byte row r sits at chunk c ^ (r & 7) (r counted from the 1024-aligned stage base)
          const int r = px + 8;
          const int chunk = (ch2 >> 2) ^ (r & 7);
          const uint32_t u = *reinterpret_cast<const uint32_t*>(tile + px * 128 + chunk * 16 + (ch2 & 3) * 4);
          s0 += bf16_lo(u);
          s1 += bf16_hi(u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        if (++s == p.nstages) { s = 0; ph ^= 1; }
      }
      if (2 * ch2 < rows_real) atomicAdd(db + 2 * ch2, s0);
      if (2 * ch2 + 1 < rows_real) atomicAdd(db + 2 * ch2 + 1, s1);
    }
    if (have_work) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      // The CTA's result is, for every output channel co, the CONTIGUOUS run dW[co][cblk*64 .. +64][3][3] (576 floats).  An
      // accumulator row holds it strided (column = kx*64 + ci, row block = ky), so a direct atomicAdd per element scatters
      // every warp instruction over 9-10 cache lines; small layers then spend more time in their atomics than in their
      // MMAs.  Instead the rows are transposed through shared memory (the pipeline stages are idle once bar_done fired;
      // row pitch 577 floats: conflict-free for lanes = rows) and added with coalesced atomics: one line per instruction.
      float* stg = reinterpret_cast<float*>(gbase + (s_st - base));
      constexpr int kLd = 577;
      const int co = m & 63;
      for (int blk = 0; blk < 2; ++blk) {                      // accumulator 0: rows [ky 2 | ky 1]; accumulator 1: rows [ky 0 | mirror]
        const int ky = blk == 0 ? (m < 64 ? 2 : 1) : 0;
        const bool rows_ok = blk == 0 || m < 64;
        for (int c0 = 0; c0 < 192; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * 192 + c0, v);
          tmem_ld_wait();
          if (rows_ok) {
            const int kx = c0 / 64;
            float* dst = stg + co * kLd + ((c0 & 63) * 9 + ky * 3 + kx);
#pragma unroll
            for (int e = 0; e < 32; ++e) dst[e * 9] = __uint_as_float(v[e]);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");           // the four epilogue warps (warps 2..5)
      const int ci_valid = min(64, p.cols_real - cblk * 64);
      const int ncols = ci_valid > 0 ? ci_valid * 9 : 0;
      for (int row = warp - 2; row < rows_real; row += 4) {
        float* g = dw + (static_cast<size_t>(row) * p.cols_real + cblk * 64) * 9;
        const float* srow = stg + row * kLd;
        for (int j = lane; j < ncols; j += 32) atomicAdd(g + j, srow[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int launch_wgrad3x3_ky(const void* x, const void* dy, float* dw, float* db, int n, int h, int w, int cin, int cout, int cin_pad,
                              cudaStream_t stream, int nlayers = 1, long long dw_stride = 0, long long db_stride = 0, int cout_pad = 64) {
  WgKyParams p{};
  p.n = n; p.h = h; p.w = w;
  p.tiles_x = tg_div_up(w, kTileW); p.tiles_y = tg_div_up(h, kTileH);
  p.num_items = n * p.tiles_x * p.tiles_y;                   // per layer
  p.rows_real = cout; p.cols_real = cin;
  p.stage_stride = (kWgKyABytes + kWgKyBBytes + 1023u) & ~1023u;
  p.nstages = 5;
  p.dw = dw; p.db = db;
  p.dw_stride = dw_stride; p.db_stride = db_stride;
  p.co_blocks = cout_pad / 64;
  const uint32_t smem_bytes = p.nstages * p.stage_stride + 16 * p.nstages + 64 + 1024;
  TG_CHECK_ARG(smem_bytes <= kWgSmemLimit, "wgrad3x3: stages do not fit in shared memory");
  const cuuint64_t images = static_cast<cuuint64_t>(n) * nlayers;
  CUtensorMap tm_a, tm_b;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(cout_pad), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), images};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(cout_pad) * 2, static_cast<cuuint64_t>(w) * cout_pad * 2,
                             static_cast<cuuint64_t>(h) * w * cout_pad * 2};
    cuuint32_t box[4] = {64, kTileW, kTileH + 2, 1};
    if (int rc = encode_bf16(&tm_a, dy, 4, dims, strides, box)) return rc;
  }
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(cin_pad), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), images};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(cin_pad) * 2, static_cast<cuuint64_t>(w) * cin_pad * 2,
                             static_cast<cuuint64_t>(h) * w * cin_pad * 2};
    cuuint32_t box[4] = {64, kTileW + 2, kTileH, 1};
    if (int rc = encode_bf16(&tm_b, x, 4, dims, strides, box)) return rc;
  }
  static const int tiles_per_slab = []() { const char* e = getenv("TG_WGRAD_TILES_PER_SLAB"); const int v = e ? atoi(e) : 8; return v > 0 ? v : 8; }();
  const int cblocks = cin_pad / 64;
  int slabs = p.num_items / tiles_per_slab;
  // a batch of layers shares the machine: SMs / (layers x channel blocks) slabs per layer (every CTA ends with rows x taps x
  // cols atomics, so few big slabs per layer beat many small ones)
  int max_slabs = tg_num_sms() / (cblocks * nlayers * p.co_blocks);
  if (max_slabs < 1) max_slabs = 1;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  static TgPerDeviceOnce attr_once;
  if (attr_once.need()) TG_CUDA(cudaFuncSetAttribute(wgrad3x3_ky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemLimit));
  tg_prof_pre(TG_K_WGRAD, 2.0 * 9.0 * cin_pad * cout_pad * n * h * w * nlayers, stream);
  wgrad3x3_ky_kernel<<<dim3(slabs, cblocks, nlayers * p.co_blocks), kWgThreads, smem_bytes < 120 * 1024 ? 120 * 1024 : smem_bytes, stream>>>(tm_a, tm_b, p);
  tg_prof_post(stream);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

// nlayers same-shape 3x3 layers (<= 64 output channels) in ONE launch: layer l reads X at x + l * (n*h*w*cin_pad) elements and
// dY at dy + l * (n*h*w*64), and adds into dw + l * dw_stride (and db + l * db_stride).  The residual trunk's 2 x 16 weight
// gradients are two such launches instead of 32: at small crops a per-layer launch is mostly its final atomics and its tail.
int launch_wgrad3x3_batched(const void* x, const void* dy, float* dw, float* db, int nlayers, long long dw_stride, long long db_stride,
                            int n, int h, int w, int cin, int cout, int cin_pad, cudaStream_t stream) {
  TG_CHECK_ARG(x && dy && dw && nlayers >= 1, "wgrad3x3_batched: null pointer / no layers");
  TG_CHECK_ARG(n > 0 && h > 0 && w > 0 && static_cast<long long>(n) * nlayers < (1LL << 31), "wgrad3x3_batched: bad shape");
  TG_CHECK_ARG((cin_pad == 64 || cin_pad == 128) && cin >= 1 && cin <= cin_pad && cout >= 1 && cout <= 64, "wgrad3x3_batched: bad channel counts");
  TG_CHECK_ARG(nlayers <= 65535, "wgrad3x3_batched: too many layers");
  return launch_wgrad3x3_ky(x, dy, dw, db, n, h, w, cin, cout, cin_pad, stream, nlayers, dw_stride, db_stride);
}

static int wgrad_launch_common(WgParams& p, const void* a, int a_pad, const void* b, int b_pad, int bh, int bw,
                               int b_box_w, int b_box_h, int ngroups, double flops, cudaStream_t stream) {
  p.tiles_x = tg_div_up(p.w, kTileW); p.tiles_y = tg_div_up(p.h, kTileH);
  p.num_items = p.n * p.tiles_x * p.tiles_y;
  p.a_boxes = a_pad / 64; p.b_boxes = b_pad / 64;
  p.a_bytes = kTileH * kTileW * 128;                         // 16 KB
  p.b_bytes = static_cast<uint32_t>(b_box_w * b_box_h * 128);
  p.b_pitch = static_cast<uint32_t>(b_box_w * 128);
  p.stage_stride = (p.a_boxes * p.a_bytes + p.b_boxes * p.b_bytes + 1023u) & ~1023u;
  int nstages = 6;
  while (nstages > 1 && nstages * p.stage_stride + 16 * nstages + 64 + 1024 > kWgSmemLimit) --nstages;
  p.nstages = nstages;
  uint32_t smem_bytes = nstages * p.stage_stride + 16 * nstages + 64 + 1024;
  TG_CHECK_ARG(smem_bytes <= kWgSmemLimit, "wgrad: stage does not fit in shared memory");
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;     // one CTA per SM (each allocates all of TMEM)

  CUtensorMap tm_a, tm_b;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(a_pad), static_cast<cuuint64_t>(p.w), static_cast<cuuint64_t>(p.h), static_cast<cuuint64_t>(p.n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(a_pad) * 2, static_cast<cuuint64_t>(p.w) * a_pad * 2,
                             static_cast<cuuint64_t>(p.h) * p.w * a_pad * 2};
    cuuint32_t box[4] = {64, kTileW, kTileH, 1};
    if (int rc = encode_bf16(&tm_a, a, 4, dims, strides, box)) return rc;
  }
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(b_pad), static_cast<cuuint64_t>(bw), static_cast<cuuint64_t>(bh), static_cast<cuuint64_t>(p.n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(b_pad) * 2, static_cast<cuuint64_t>(bw) * b_pad * 2,
                             static_cast<cuuint64_t>(bh) * bw * b_pad * 2};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(b_box_w * p.b_in_scale), static_cast<cuuint32_t>(b_box_h * p.b_in_scale), 1};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(p.b_in_scale), static_cast<cuuint32_t>(p.b_in_scale), 1};
    if (int rc = encode_bf16(&tm_b, b, 4, dims, strides, box, estr)) return rc;
  }
  // split-K over pixel slabs: enough CTAs to fill the machine on big layers, few on tiny ones (every CTA pays
  // rows x taps x cols atomics at the end)
  static const int tiles_per_slab = []() { const char* e = getenv("TG_WGRAD_TILES_PER_SLAB"); const int v = e ? atoi(e) : 8; return v > 0 ? v : 8; }();
  int slabs = p.num_items / tiles_per_slab;
  const int max_slabs = tg_num_sms() / ngroups;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  static TgPerDeviceOnce attr_once;
  if (attr_once.need()) TG_CUDA(cudaFuncSetAttribute(wgrad3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemLimit));
  tg_prof_pre(TG_K_WGRAD, flops, stream);
  wgrad3x3_kernel<<<dim3(slabs, ngroups), kWgThreads, smem_bytes, stream>>>(tm_a, tm_b, p);
  tg_prof_post(stream);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int launch_wgrad3x3(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout, int cin_pad,
                    int cout_pad, cudaStream_t stream, float* db) {
  TG_CHECK_ARG(x && dy && dw, "wgrad3x3: null pointer");
  TG_CHECK_ARG(n > 0 && h > 0 && w > 0, "wgrad3x3: bad shape");
  TG_CHECK_ARG((cin_pad == 64 || cin_pad == 128) && (cout_pad == 64 || cout_pad == 128), "wgrad3x3: padded channels must be 64 or 128");
  TG_CHECK_ARG(cin >= 1 && cin <= cin_pad && cout >= 1 && cout <= cout_pad, "wgrad3x3: bad channel counts");
  static const bool ky_on = []() { const char* e = getenv("TG_WGRAD_KYSTACK"); return !(e && e[0] == '0'); }();   // A/B knob
  // (128 output channels: two 64-row blocks of dY, each its own CTAs - TG_WGRAD_KY128=0 keeps them on the generic kernel)
  static const bool ky128_on = []() { const char* e = getenv("TG_WGRAD_KY128"); return !(e && e[0] == '0'); }();
  if (ky_on && (cout_pad == 64 || ky128_on)) return launch_wgrad3x3_ky(x, dy, dw, db, n, h, w, cin, cout, cin_pad, stream, 1, 0, 0, cout_pad);
  if (db) {                                                  // 128 output channels: the separate HBM-rate reduction
    if (int rc = launch_bias_grad(dy, static_cast<long long>(n) * h * w, cout_pad, cout, db, stream)) return rc;
  }
  WgParams p{};
  p.n = n; p.h = h; p.w = w;
  p.rows_real = cout; p.cols_real = cin; p.ks = 3;           // A = dY (M = co), B = X (N = ci)
  p.b_in_scale = 1;
  p.stack3 = (cin_pad == 64) ? 1 : 0;
  for (int ky = 0; ky < 3; ++ky) {                           // group = filter row: X rows shifted by ky-1, 1-pixel halo in x
    p.groups[ky].b_ox = -1; p.groups[ky].b_oy = ky - 1; p.groups[ky].ntaps = 3;
    for (int kx = 0; kx < 3; ++kx) p.groups[ky].taps[kx] = WgTap{static_cast<uint32_t>(kx * 128), ky, kx};
  }
  p.dw = dw;
  return wgrad_launch_common(p, dy, cout_pad, x, cin_pad, h, w, kTileW + 2, kTileH, 3,
                             2.0 * 9.0 * cin_pad * cout_pad * n * h * w, stream);
}

// ConvTranspose2d(k3, s2, p1, op1): y[2iy+ky-1, 2ix+kx-1, co] += x[iy, ix, ci] * Wt[ci][co][ky][kx]
//   dWt[ci][co][ky][kx] = sum over (n, iy, ix) of x[n, iy, ix, ci] * dY[n, 2iy-1+ky, 2ix-1+kx, co]
// A = x tile (M = ci), B = dY staged per parity phase with element stride 2 (N = co).  Odd rows 2s+1 hold taps
// ky = 0 (s = iy-1) and ky = 2 (s = iy); even rows hold ky = 1.  h, w = size of x.
int launch_wgrad_convT3x3s2(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                            int cin_pad, int cout_pad, cudaStream_t stream) {
  TG_CHECK_ARG(x && dy && dw, "wgrad_convT: null pointer");
  TG_CHECK_ARG(n > 0 && h > 0 && w > 0, "wgrad_convT: bad shape");
  TG_CHECK_ARG((cin_pad == 64 || cin_pad == 128) && (cout_pad == 64 || cout_pad == 128), "wgrad_convT: padded channels must be 64 or 128");
  WgParams p{};
  p.n = n; p.h = h; p.w = w;
  p.rows_real = cin; p.cols_real = cout; p.ks = 3;
  p.b_in_scale = 2;
  p.stack3 = 0;
  const int bw = kTileW + 1;                                 // 9 x 17 samples per phase box
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      WgGroup& g = p.groups[py * 2 + px];
      g.b_ox = px ? -1 : 0; g.b_oy = py ? -1 : 0;            // odd phase: first sample is row/col 2*(t0-1)+1
      g.ntaps = 0;
      for (int ty = 0; ty <= py; ++ty)
        for (int tx = 0; tx <= px; ++tx) {
          const int ky = py ? 2 * ty : 1, kx = px ? 2 * tx : 1;
          g.taps[g.ntaps++] = WgTap{static_cast<uint32_t>((ty * bw + tx) * 128), ky, kx};
        }
    }
  p.dw = dw;
  return wgrad_launch_common(p, x, cin_pad, dy, cout_pad, 2 * h, 2 * w, bw, kTileH + 1, 4,
                             2.0 * 9.0 * cin_pad * cout_pad * n * h * w, stream);
}

// Conv2d(k4, s2, p1): dW[co][ci][ky][kx] = sum over (n, oy, ox) of dY[n, oy, ox, co] * X[n, 2oy-1+ky, 2ox-1+kx, ci]
// A = dY tile (M = co), B = X staged per parity phase (N = ci).  h, w = size of dY (the conv output).
int launch_wgrad_conv4x4s2(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                           int cin_pad, int cout_pad, cudaStream_t stream) {
  TG_CHECK_ARG(x && dy && dw, "wgrad_conv4x4s2: null pointer");
  TG_CHECK_ARG(n > 0 && h > 0 && w > 0, "wgrad_conv4x4s2: bad shape");
  TG_CHECK_ARG((cin_pad == 64 || cin_pad == 128) && (cout_pad == 64 || cout_pad == 128), "wgrad_conv4x4s2: padded channels must be 64 or 128");
  WgParams p{};
  p.n = n; p.h = h; p.w = w;
  p.rows_real = cout; p.cols_real = cin; p.ks = 4;
  p.b_in_scale = 2;
  p.stack3 = 0;
  const int bw = kTileW + 1;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      WgGroup& g = p.groups[py * 2 + px];
      // input row 2oy-1+ky: ky odd -> even rows (samples oy, oy+1), ky even -> odd rows 2s+1 (samples oy-1, oy)
      g.b_ox = px ? 0 : -1; g.b_oy = py ? 0 : -1;
      g.ntaps = 0;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx)
          g.taps[g.ntaps++] = WgTap{static_cast<uint32_t>((ty * bw + tx) * 128), 2 * ty + py, 2 * tx + px};
    }
  p.dw = dw;
  return wgrad_launch_common(p, dy, cout_pad, x, cin_pad, 2 * h, 2 * w, bw, kTileH + 1, 4,
                             2.0 * 16.0 * cin_pad * cout_pad * n * h * w, stream);
}

// db[c] += sum over pixels of dy[p][c]; dy NHWC bf16 [pixels][cpad]
__global__ void __launch_bounds__(256)
bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, long long pixels, int cpad, int c, float* __restrict__ db) {
  __shared__ float s_sum[256][9];
  const int groups = cpad / 8, pix_per_iter = 256 / groups;
  const int gi = threadIdx.x % groups, pl = threadIdx.x / groups;
  float sum[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sum[e] = 0.f;
  for (long long px = static_cast<long long>(blockIdx.x) * pix_per_iter + pl; px < pixels;
       px += static_cast<long long>(gridDim.x) * pix_per_iter) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + px * cpad) + gi);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { sum[2 * e] += bf16_lo(u[e]); sum[2 * e + 1] += bf16_hi(u[e]); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_sum[threadIdx.x][e] = sum[e];
  __syncthreads();
  if (threadIdx.x < c) {
    const int g = threadIdx.x / 8, e = threadIdx.x % 8;
    float a = 0.f;
    for (int r = 0; r < pix_per_iter; ++r) a += s_sum[r * groups + g][e];
    atomicAdd(db + threadIdx.x, a);
  }
}

int launch_bias_grad(const void* dy, long long pixels, int cpad, int c, float* db, cudaStream_t stream) {
  TG_CHECK_ARG(dy && db && (cpad == 64 || cpad == 128) && c >= 1 && c <= cpad, "bias_grad: bad arguments");
  const int pix_per_iter = 256 / (cpad / 8);
  long long blocks = (pixels + pix_per_iter * 8 - 1) / (pix_per_iter * 8);
  const long long cap = static_cast<long long>(tg_num_sms()) * 4;     // enough loads in flight to stream at HBM rate
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  tg_prof_pre(TG_K_GLUE, 2.0 * pixels * cpad, stream);
  bias_grad_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy), pixels, cpad, c, db);
  tg_prof_post(stream);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

}  // namespace tg

static int pad64(int c) { return c <= 64 ? 64 : 128; }

extern "C" int tg_conv3x3_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                                void* stream) {
  return tg::launch_wgrad3x3(x, dy, dw, n, h, w, cin, cout, pad64(cin), pad64(cout), static_cast<cudaStream_t>(stream));
}
extern "C" int tg_conv3x3_wgrad_bias(const void* x, const void* dy, float* dw, float* db, int n, int h, int w, int cin, int cout,
                                     void* stream) {
  return tg::launch_wgrad3x3(x, dy, dw, n, h, w, cin, cout, pad64(cin), pad64(cout), static_cast<cudaStream_t>(stream), db);
}
extern "C" int tg_convT3x3s2_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                                   void* stream) {
  return tg::launch_wgrad_convT3x3s2(x, dy, dw, n, h, w, cin, cout, pad64(cin), pad64(cout), static_cast<cudaStream_t>(stream));
}
extern "C" int tg_conv4x4s2_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                                  void* stream) {
  return tg::launch_wgrad_conv4x4s2(x, dy, dw, n, h, w, cin, cout, pad64(cin), pad64(cout), static_cast<cudaStream_t>(stream));
}
extern "C" int tg_bias_grad(const void* dy, float* db, long long pixels, int c, void* stream) {
  return tg::launch_bias_grad(dy, pixels, pad64(c), c, db, static_cast<cudaStream_t>(stream));
}
