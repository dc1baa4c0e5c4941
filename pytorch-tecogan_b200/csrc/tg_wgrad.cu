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
#include "tg_conv_tc.cuh"

namespace tg {

constexpr int kWgThreads = 192;
constexpr uint32_t kWgSmemLimit = 232448;

struct WgParams {
  int n, h, w, tiles_x, tiles_y, num_items;
  int cin, cout, cin_pad, cout_pad;   // real / padded channel counts
  int a_boxes, b_boxes;               // cout_pad / 64, cin_pad / 64
  uint32_t a_bytes, b_bytes, stage_stride;
  int nstages;
  float* dw;                          // [cout][cin][3][3] f32, accumulated into
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
  const int ky = blockIdx.y;

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
          tma_load_4d(dst + p.a_boxes * p.a_bytes + c * p.b_bytes, &tm_x, bar_full + 8 * s, c * 64, x0 - 1, y0 + ky - 1, n);
      }
      __syncwarp();
      if (++s == p.nstages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: D[co][(kx, ci)] += dY^T[co][pixels] * X[pixels][(kx, ci)]
    const bool ci64 = (p.b_boxes == 1);
    const uint32_t idesc = ci64 ? umma_idesc_bf16_mn(128, 192) : umma_idesc_bf16_mn(128, 128);
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
          if (ci64) {
            const uint64_t bd = umma_desc_mn_sw128(b_base + (2 * j * 10) * 128, 128, 1280);
            umma_bf16(tmem_base, ad, bd, idesc, (first && j == 0) ? 0u : 1u);
          } else {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t bd = umma_desc_mn_sw128(b_base + (2 * j * 10 + kx) * 128, p.b_bytes, 1280);
              umma_bf16(tmem_base + kx * 128, ad, bd, idesc, (first && j == 0) ? 0u : 1u);
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
    const int co = q * 32 + lane;
    const bool have_work = blockIdx.x < p.num_items;
    if (have_work) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int ncol = (p.b_boxes == 1) ? 192 : 384;
      const int cstride = (p.b_boxes == 1) ? 64 : 128;      // columns per kx tap
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (co < p.cout) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = c0 + e;
            const int kx = col / cstride, ci = col - kx * cstride;
            if (ci < p.cin)
              atomicAdd(p.dw + ((static_cast<size_t>(co) * p.cin + ci) * 3 + ky) * 3 + kx, __uint_as_float(v[e]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int launch_wgrad3x3(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout, int cin_pad,
                    int cout_pad, cudaStream_t stream) {
  TG_CHECK_ARG(x && dy && dw, "wgrad3x3: null pointer");
  TG_CHECK_ARG(n > 0 && h > 0 && w > 0, "wgrad3x3: bad shape");
  TG_CHECK_ARG((cin_pad == 64 || cin_pad == 128) && (cout_pad == 64 || cout_pad == 128), "wgrad3x3: padded channels must be 64 or 128");
  TG_CHECK_ARG(cin >= 1 && cin <= cin_pad && cout >= 1 && cout <= cout_pad, "wgrad3x3: bad channel counts");
  WgParams p{};
  p.n = n; p.h = h; p.w = w;
  p.tiles_x = tg_div_up(w, kTileW); p.tiles_y = tg_div_up(h, kTileH);
  p.num_items = n * p.tiles_x * p.tiles_y;
  p.cin = cin; p.cout = cout; p.cin_pad = cin_pad; p.cout_pad = cout_pad;
  p.a_boxes = cout_pad / 64; p.b_boxes = cin_pad / 64;
  p.a_bytes = kTileH * kTileW * 128;            // 16 KB, 1024-aligned
  p.b_bytes = kTileH * (kTileW + 2) * 128;      // 20 KB, 1024-aligned
  p.stage_stride = p.a_boxes * p.a_bytes + p.b_boxes * p.b_bytes;
  int nstages = 6;
  while (nstages > 1 && nstages * p.stage_stride + 16 * nstages + 64 + 1024 > kWgSmemLimit) --nstages;
  p.nstages = nstages;
  p.dw = dw;
  uint32_t smem_bytes = nstages * p.stage_stride + 16 * nstages + 64 + 1024;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one CTA per SM (each allocates all of TMEM)

  CUtensorMap tm_dy, tm_x;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(cout_pad), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(cout_pad) * 2, static_cast<cuuint64_t>(w) * cout_pad * 2,
                             static_cast<cuuint64_t>(h) * w * cout_pad * 2};
    cuuint32_t box[4] = {64, kTileW, kTileH, 1};
    if (int rc = encode_bf16(&tm_dy, dy, 4, dims, strides, box)) return rc;
  }
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(cin_pad), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(cin_pad) * 2, static_cast<cuuint64_t>(w) * cin_pad * 2,
                             static_cast<cuuint64_t>(h) * w * cin_pad * 2};
    cuuint32_t box[4] = {64, kTileW + 2, kTileH, 1};
    if (int rc = encode_bf16(&tm_x, x, 4, dims, strides, box)) return rc;
  }
  // split-K over pixel slabs: enough CTAs to fill the machine on big layers, few on tiny ones (every CTA pays
  // cout x 3 x cin atomics at the end)
  int slabs = p.num_items / 8;
  const int max_slabs = tg_num_sms() / 3;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  static bool attr_done = false;
  if (!attr_done) {
    TG_CUDA(cudaFuncSetAttribute(wgrad3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemLimit));
    attr_done = true;
  }
  tg_prof_pre(TG_K_WGRAD, 2.0 * 9.0 * cin_pad * cout_pad * n * h * w, stream);
  wgrad3x3_kernel<<<dim3(slabs, 3), kWgThreads, smem_bytes, stream>>>(tm_dy, tm_x, p);
  tg_prof_post(stream);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

}  // namespace tg

extern "C" int tg_conv3x3_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                                void* stream) {
  return tg::launch_wgrad3x3(x, dy, dw, n, h, w, cin, cout, tg::cin_padded(cin), cout <= 64 ? 64 : 128,
                             static_cast<cudaStream_t>(stream));
}
