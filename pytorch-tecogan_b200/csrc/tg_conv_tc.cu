// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Replaces the nn.Conv2d / nn.ConvTranspose2d calls of the generator
// (reference code/models.py:54-58,68-76 via code/ops.py:45-63).
//
// GEMM view (per work item = one 16x8 pixel sub-tile of one image):
//   D[128 pixels x NT out-channels] += A_tap[128 pixels x 64 in-channels] * W_tap[NT x 64]^T
// summed over filter taps and 64-channel K chunks.  Activations are NHWC bf16, so one pixel's
// 64 channels are one 128-byte row == one SWIZZLE_128B row of the UMMA K-major canonical layout.
//
//  * A operand: ONE TMA box {64ch, 10, 18} (sub-tile + halo, out-of-bounds zero-filled == the
//    conv's zero padding) per stage.  The nine taps are nine *views* of that box: the UMMA
//    descriptor start address is shifted by (dy*10+dx) rows and the 8-row-group stride (SBO) is
//    the 10-pixel row pitch, so every input byte is fetched from L2 once per item (not 9x).
//    (TG_AMODE_DX3 is the conservative variant: three x-shifted boxes so that every descriptor
//    start stays 1024-byte aligned; 3x L2 traffic.)
//  * B operand: the layer's packed weights for this CTA's 64-wide output-channel chunk stay
//    resident in shared memory for the lifetime of the persistent CTA.
//  * Accumulators live in TMEM (ring of 512/64 columns), so the epilogue of item i overlaps the
//    MMAs of item i+1.  ConvTranspose(k3,s2) runs as 4 output phases = 4 accumulators fed from
//    the same staged A tile (9 (phase,tap) pairs == 9 MMA groups, same FLOPs as a 3x3 conv).
//  * Warp roles: warp0 = TMA producer, warp1 = TMEM owner + single-thread MMA issuer,
//    warps 2..5 = epilogue (TMEM -> registers -> bias/ReLU/residual -> global).
//  * Pair mode (kPair, the default for 64-wide chunks): the CTAs of the two SMs of a TPC form a cluster and run their
//    items in lock-step as ONE tcgen05.mma.cta_group::2 of M = 256; every CTA stages its own item's activations and
//    HALF of the weight rows (32 of the 64 output channels of each tap block), the leader issues for both, the "full"
//    barriers live in the leader and take one arrival + the TMA bytes of each CTA, MMA completion is multicast.  An
//    N = 64 MMA with both operands in shared memory costs 75 cycles alone and 43 as a pair for twice the work
//    (profiles/r01_mma_microbench_v2.txt).  An odd item count is padded with an out-of-range item (zero fill).
#include <stdlib.h>

#include "tg_conv_tc.cuh"

namespace tg {

#ifndef TG_CONV_EPI_WARPS
#define TG_CONV_EPI_WARPS 16
#endif
constexpr int kEpiWarps = TG_CONV_EPI_WARPS;            // (TMEM lane quarter) x (part of the 64 accumulator columns)
// Two SETS of eight warps, set s draining every other item (s, s+2, ...): two items' epilogues in flight.  A warp then
// owns 32 columns of its lane quarter.  (TG_CONV_TWO_SETS=0: all warps on every item, 64 / (warps / 4) columns each.)
#ifndef TG_CONV_TWO_SETS
#define TG_CONV_TWO_SETS 1
#endif
constexpr bool kTwoSets = TG_CONV_TWO_SETS != 0 && kEpiWarps == 16;
constexpr int kSetWarps = kTwoSets ? kEpiWarps / 2 : kEpiWarps;
constexpr int kEpiCols = 64 / (kSetWarps / 4);          // columns per epilogue warp: 32 (8 warps per item) or 16 (16)
static_assert(kEpiWarps == 8 || kEpiWarps == 16, "epilogue warps");
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kSmemLimit = 232448;   // 227 KB opt-in maximum per CTA

struct SmemLayout {
  uint32_t w_off, a_off, bar_off, tmem_ptr_off, bias_off, total;
};

__host__ __device__ inline SmemLayout make_layout(uint32_t w_bytes, uint32_t stage_stride,
                                                  int nstages, int ngroups, int nt) {
  SmemLayout l;
  l.w_off = 0;
  l.a_off = (w_bytes + 1023u) & ~1023u;
  l.bar_off = l.a_off + stage_stride * nstages;
  uint32_t nbar = 1 + 2 * nstages + 2 * ngroups;
  l.tmem_ptr_off = l.bar_off + 8 * nbar;
  l.bias_off = (l.tmem_ptr_off + 4 + 15u) & ~15u;
  l.total = l.bias_off + 4 * nt;
  return l;
}

template <int NT, bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B needs 1024B alignment
  uint8_t* gbase = smem_raw + (base - raw);
  const SmemLayout L = make_layout(p.w_bytes, p.stage_stride, p.nstages, p.ngroups, NT);

  const uint32_t s_w = base + L.w_off;
  const uint32_t s_a = base + L.a_off;
  const uint32_t bar_w = base + L.bar_off;
  const uint32_t bar_afull = bar_w + 8;
  const uint32_t bar_aempty = bar_afull + 8 * p.nstages;
  const uint32_t bar_cfull = bar_aempty + 8 * p.nstages;
  const uint32_t bar_cempty = bar_cfull + 8 * p.ngroups;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gbase + L.tmem_ptr_off);
  float* s_bias = reinterpret_cast<float*>(gbase + L.bias_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.y;                             // 64-wide output channel chunk
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;    // pair mode: rank 0 issues the MMAs
  const uint32_t nctas = kPair ? 2u : 1u;
  const int num_items = kPair ? ((p.num_items + 1) & ~1) : p.num_items;   // (the padding item loads zeros, stores nothing)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    mbar_init(bar_w, nctas);
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(bar_afull + 8 * i, nctas);
      mbar_init(bar_aempty + 8 * i, 1);
    }
    for (int i = 0; i < p.ngroups; ++i) {
      mbar_init(bar_cfull + 8 * i, 1);
      mbar_init(bar_cempty + 8 * i, kSetWarps * nctas);     // one arrive per epilogue warp of the item's set (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NT) {
    const int c = threadIdx.x - 64;
    s_bias[c] = p.bias ? p.bias[chunk * NT + c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();                            // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // leader-side barrier addresses (shared::cluster window; the CTA's own when not paired)
  const uint32_t lbar_w = kPair ? mapa_rank(bar_w, 0) : bar_w;
  const uint32_t lbar_afull = kPair ? mapa_rank(bar_afull, 0) : bar_afull;
  const uint32_t lbar_cempty = kPair ? mapa_rank(bar_cempty, 0) : bar_cempty;
  // Programmatic dependent launch: the next layer's CTAs may start their prologue (barrier init,
  // TMEM alloc, weight TMA) as soon as SMs free up; they block in griddepcontrol.wait until this
  // grid has completed and flushed, before touching any activation.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int blocks_per_chunk = p.stages_per_item * p.ntaps;
  constexpr uint32_t kWBlockBytes = NT * 128 / (kPair ? 2 : 1);   // rows of one tap block held by this CTA x 128 B
  constexpr int kWRows = NT / (kPair ? 2 : 1);

  if (warp == 0) {
    // ================================ TMA producer =========================================
    // Whole warp runs the (uniform) loop; one elected lane issues.  Keeps TMA operands in uniform
    // registers (no per-lane waterfall loops around UTMALDG).
    if (!p.w_stage_bytes && elect_one()) {
      if (kPair) {
        mbar_expect_tx_cluster(lbar_w, p.w_bytes);
        for (int b = 0; b < blocks_per_chunk; ++b)
          tma_load_2d_pair(s_w + b * kWBlockBytes, &tm_w, lbar_w, 0, (chunk * blocks_per_chunk + b) * NT + static_cast<int>(rank) * kWRows);
      } else {
        mbar_expect_tx(bar_w, p.w_bytes);
        for (int b = 0; b < blocks_per_chunk; ++b)
          tma_load_2d(s_w + b * kWBlockBytes, &tm_w, bar_w, 0, (chunk * blocks_per_chunk + b) * NT);
      }
    }
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");      // weights are constants; activations are not
    int s = 0;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int tx = it % p.tiles_x;
      const int r = it / p.tiles_x;
      const int ty = r % p.tiles_y;
      const int n = r / p.tiles_y;                           // == p.n for the padding item: every box out of range -> zeros
      const int x0 = tx * kTileW, y0 = ty * kTileH;
      for (int sidx = 0; sidx < p.stages_per_item; ++sidx) {
        const int kc = sidx / p.nphase, phs = sidx - kc * p.nphase;
        mbar_wait(bar_aempty + 8 * s, ph ^ 1);
        if (elect_one()) {
          const uint32_t dst = s_a + s * p.stage_stride;
          if (kPair) {
            mbar_expect_tx_cluster(lbar_afull + 8 * s, p.stage_bytes + p.w_stage_bytes);
            for (int c = 0; c < p.ncopies; ++c)
              tma_load_4d_pair(dst + c * p.copy_bytes, &tm_a, lbar_afull + 8 * s, kc * 64,
                               x0 * p.in_scale + p.copy_dx[c] + (phs & 1), y0 * p.in_scale + p.box_y0 + (phs >> 1), n);
            if (p.w_stage_bytes)                             // this CTA's half of the weights of this (K chunk, phase)
              for (int j = 0; j < p.ntaps; ++j)
                tma_load_2d_pair(dst + p.a_region + j * kWBlockBytes, &tm_w, lbar_afull + 8 * s, 0,
                                 ((chunk * p.stages_per_item + sidx) * p.ntaps + j) * NT + static_cast<int>(rank) * kWRows);
          } else {
            mbar_expect_tx(bar_afull + 8 * s, p.stage_bytes + p.w_stage_bytes);
            for (int c = 0; c < p.ncopies; ++c)
              tma_load_4d(dst + c * p.copy_bytes, &tm_a, bar_afull + 8 * s, kc * 64,
                          x0 * p.in_scale + p.copy_dx[c] + (phs & 1), y0 * p.in_scale + p.box_y0 + (phs >> 1), n);
            if (p.w_stage_bytes)                             // weights of this (K chunk, phase) ride in the stage
              for (int j = 0; j < p.ntaps; ++j)
                tma_load_2d(dst + p.a_region + j * kWBlockBytes, &tm_w, bar_afull + 8 * s, 0,
                            ((chunk * p.stages_per_item + sidx) * p.ntaps + j) * NT);
          }
        }
        __syncwarp();
        if (++s == p.nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===========================================
    constexpr uint32_t idesc = umma_idesc_bf16(kPair ? 256 : 128, NT);
    if (rank == 0) {                                        // (pair mode: the peer's issuer warp only owns its TMEM)
    if (!p.w_stage_bytes) mbar_wait(bar_w, 0);
    tc_fence_after();
    int s = 0, g = 0;
    uint32_t ph = 0, gph = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      mbar_wait(bar_cempty + 8 * g, gph ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + static_cast<uint32_t>(g * p.n_acc * kAccCols);
      for (int sidx = 0; sidx < p.stages_per_item; ++sidx) {
        mbar_wait(bar_afull + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_base = s_a + s * p.stage_stride;
          const uint32_t w_base = p.w_stage_bytes ? a_base + p.a_region : s_w + sidx * p.ntaps * kWBlockBytes;
          for (int j = 0; j < p.ntaps; ++j) {
            const TcTap tap = p.taps[j];
            const uint64_t ad = umma_desc_sw128(a_base + tap.a_off, p.sbo);
            const uint64_t bd = umma_desc_sw128(w_base + j * kWBlockBytes, 1024);
            const uint32_t d = d_base + tap.acc * kAccCols;
            const uint32_t keep = (sidx > 0 || !tap.first) ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 x (K=16 bf16 = 32 bytes) inside the 128B swizzle row
              if (kPair) umma_bf16_pair(d, ad + 2 * k, bd + 2 * k, idesc, (k > 0) ? 1u : keep);
              else umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (k > 0) ? 1u : keep);
            }
          }
          if (kPair) {
            umma_commit_pair(bar_aempty + 8 * s);            // both CTAs' stages reusable once these MMAs retire
            if (sidx == p.stages_per_item - 1) umma_commit_pair(bar_cfull + 8 * g);
          } else {
            umma_commit(bar_aempty + 8 * s);                 // stage reusable once these MMAs retire
            if (sidx == p.stages_per_item - 1) umma_commit(bar_cfull + 8 * g);   // accumulators of this item complete
          }
        }
        __syncwarp();
        if (++s == p.nstages) { s = 0; ph ^= 1; }
      }
      if (++g == p.ngroups) { g = 0; gph ^= 1; }
    }
    }
  } else {
    // ================================ epilogue (16 warps, two sets) ========================
    // With four warps doing 64 channels per lane the kernel was epilogue-bound.  Sixteen warps in two sets of eight:
    // set s drains the items s, s+2, ... of this CTA (accumulator groups of the same parity; the ring has 8 or 2
    // groups), a warp = one TMEM lane quarter x 32 of the 64 columns.
    const int q = warp & 3;                                // TMEM lane quarter of this warp
    const int set = kTwoSets ? (warp - 2) >> 3 : 0;        // which items this warp works on (two sets)
    const int h2 = ((warp - 2) >> 2) & (kSetWarps / 4 - 1);   // which kEpiCols of the 64 columns
    const int m = q * 32 + lane;                           // GEMM row == pixel within sub-tile
    const int pr = m >> 3, pc = m & 7;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int gstep = kTwoSets ? 2 : 1;
    int g = set % p.ngroups;
    uint32_t gph = 0;
    for (int it = blockIdx.x + set * gridDim.x; it < num_items; it += gstep * gridDim.x) {
      const int tx = it % p.tiles_x;
      const int r = it / p.tiles_x;
      const int ty = r % p.tiles_y;
      const int n = r / p.tiles_y;
      const int iy = ty * kTileH + pr, ix = tx * kTileW + pc;
      const bool valid = (it < p.num_items) && (iy < p.h) && (ix < p.w);
      mbar_wait(bar_cfull + 8 * g, gph);
      tc_fence_after();
      for (int a = 0; a < p.n_acc; ++a) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>((g * p.n_acc + a) * kAccCols);
        const int oy = iy * p.sy + p.acc_oy[a], ox = ix * p.sx + p.acc_ox[a];
        if constexpr (NT == 64) {
          uint32_t v[kEpiCols];
          if constexpr (kEpiCols == 32) tmem_ld_32x32(taddr + h2 * 32, reinterpret_cast<uint32_t(&)[32]>(v));
          else tmem_ld_32x16(taddr + h2 * 16, reinterpret_cast<uint32_t(&)[16]>(v));
          tmem_ld_wait();
          if (a == p.n_acc - 1) {                          // TMEM drained -> hand the slot back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (kPair) mbar_arrive_cluster(lbar_cempty + 8 * g); else mbar_arrive(bar_cempty + 8 * g); }
          }
          if (valid && p.out_mode == kOutNHWCf32) {          // raw f32 (bias, no activation): BatchNorm input
            const size_t pix = (static_cast<size_t>(n) * p.oh + oy) * p.ow + ox;
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + pix * p.oc + chunk * 64);
#pragma unroll
            for (int c4 = 0; c4 < kEpiCols / 4; ++c4)
              dst[h2 * (kEpiCols / 4) + c4] = make_float4(__uint_as_float(v[c4 * 4 + 0]) + s_bias[h2 * kEpiCols + c4 * 4 + 0],
                                                          __uint_as_float(v[c4 * 4 + 1]) + s_bias[h2 * kEpiCols + c4 * 4 + 1],
                                                          __uint_as_float(v[c4 * 4 + 2]) + s_bias[h2 * kEpiCols + c4 * 4 + 2],
                                                          __uint_as_float(v[c4 * 4 + 3]) + s_bias[h2 * kEpiCols + c4 * 4 + 3]);
          } else if (valid) {
            const size_t pix = (static_cast<size_t>(n) * p.oh + oy) * p.ow + ox;
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + pix * p.oc +
                                                  chunk * 64);
            const uint4* res = p.resid ? reinterpret_cast<const uint4*>(
                                             static_cast<const __nv_bfloat16*>(p.resid) +
                                             pix * p.oc + chunk * 64)
                                       : nullptr;
            const uint4* msk = p.mask ? reinterpret_cast<const uint4*>(
                                            static_cast<const __nv_bfloat16*>(p.mask) + pix * p.oc + chunk * 64)
                                      : nullptr;
            {
#pragma unroll
              for (int c8 = 0; c8 < kEpiCols / 8; ++c8) {  // 8 channels = one 16-byte store
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  f[e] = __uint_as_float(v[c8 * 8 + e]) + s_bias[h2 * kEpiCols + c8 * 8 + e];
                  if (p.relu == kActRelu) f[e] = fmaxf(f[e], 0.f);
                  else if (p.relu == kActLrelu02) f[e] = f[e] > 0.f ? f[e] : 0.2f * f[e];
                }
                if (res) {
                  const uint4 rv = __ldg(res + h2 * (kEpiCols / 8) + c8);
                  f[0] += bf16_lo(rv.x); f[1] += bf16_hi(rv.x);
                  f[2] += bf16_lo(rv.y); f[3] += bf16_hi(rv.y);
                  f[4] += bf16_lo(rv.z); f[5] += bf16_hi(rv.z);
                  f[6] += bf16_lo(rv.w); f[7] += bf16_hi(rv.w);
                }
                if (msk) {                                   // ReLU backward: gradient flows where the saved output > 0
                  const uint4 mv = __ldg(msk + h2 * (kEpiCols / 8) + c8);
                  const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    if (p.mask_mode == kMaskLrelu02) {       // saved <= 0 (zero or sign bit): slope 0.2
                      if ((mw[e] & 0x7FFFu) == 0u || (mw[e] & 0x8000u)) f[2 * e] *= 0.2f;
                      if ((mw[e] & 0x7FFF0000u) == 0u || (mw[e] & 0x80000000u)) f[2 * e + 1] *= 0.2f;
                    } else {
                      if ((mw[e] & 0x7FFFu) == 0u) f[2 * e] = 0.f;
                      if ((mw[e] & 0x7FFF0000u) == 0u) f[2 * e + 1] = 0.f;
                    }
                  }
                }
                uint4 o;
                o.x = pack_bf16x2(f[0], f[1]);
                o.y = pack_bf16x2(f[2], f[3]);
                o.z = pack_bf16x2(f[4], f[5]);
                o.w = pack_bf16x2(f[6], f[7]);
                dst[h2 * (kEpiCols / 8) + c8] = o;
              }
            }
          }
        } else {
          uint32_t v[16];
          tmem_ld_32x16(taddr, v);
          tmem_ld_wait();
          if (a == p.n_acc - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (kPair) mbar_arrive_cluster(lbar_cempty + 8 * g); else mbar_arrive(bar_cempty + 8 * g); }
          }
          if (valid && h2 == 0) {                          // (the 3-channel output conv needs one warp per lane quarter)
            const size_t plane = static_cast<size_t>(p.oh) * p.ow;
            const size_t o0 = static_cast<size_t>(n) * p.out_nstride + static_cast<size_t>(oy) * p.ow + ox;
            for (int c = 0; c < p.oc; ++c) {
              const float z = __uint_as_float(v[c]) + s_bias[c];
              if (p.out2) p.out2[o0 + c * plane] = z;
              static_cast<float*>(p.out)[o0 + c * plane] = (p.out_mode == kOutNCHWf32Sigmoid) ? 1.f / (1.f + expf(-z)) : z;
            }
          }
        }
      }
      g += gstep;
      if (g >= p.ngroups) { g -= p.ngroups; gph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();                            // the leader's MMAs read the peer's shared memory
  if (warp == 1) {
    if (kPair) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_bf16(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* elem_strides) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    tg_set_error("cuTensorMapEncodeTiled entry point not available");
    return TG_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides) for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims,
                   strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tg_set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return TG_ERR_CUDA;
  }
  return TG_OK;
}

int cin_padded(int cin) { return cin <= 64 ? 64 : 128; }
int cout_padded(int cout) { return cout <= 16 ? 16 : (cout <= 64 ? 64 : 128); }
size_t packed_weight_bytes(int cin_pad, int cout_pad) {
  return static_cast<size_t>(9) * cin_pad * cout_pad * 2;
}

int launch_conv_tc(int kind, int out_mode, const void* x, const void* packed_w, const float* bias,
                   const void* resid, void* out, float* out2, int n, int h, int w, int cin_pad,
                   int cout_pad, int relu, int amode, long long out_nstride, cudaStream_t stream, const void* mask,
                   int mask_mode, int phase) {
  TG_CHECK_ARG(x && packed_w && out, "conv: null pointer");
  TG_CHECK_ARG(!(mask && out_mode != kOutNHWCbf16), "conv: mask only for the NHWC bf16 output");
  TG_CHECK_ARG(n > 0 && h > 0 && w > 0, "conv: bad shape n=%d h=%d w=%d", n, h, w);
  TG_CHECK_ARG(kind == kConv3x3 || kind == kConvT3x3s2 || kind == kConv4x4s2 || kind == kConvT4x4s2Phase, "conv: bad kind %d", kind);
  const bool tph = (kind == kConvT4x4s2Phase);
  TG_CHECK_ARG(!tph || (phase >= 0 && phase < 4), "conv: bad phase %d", phase);
  TG_CHECK_ARG(cin_pad == 64 || cin_pad == 128, "conv: cin_pad must be 64 or 128 (got %d)", cin_pad);
  TG_CHECK_ARG(cout_pad == 16 || cout_pad == 64 || cout_pad == 128, "conv: cout_pad must be 16/64/128 (got %d)", cout_pad);
  const bool nchw_out = (out_mode == kOutNCHWf32Sigmoid || out_mode == kOutNCHWf32Raw);
  TG_CHECK_ARG(nchw_out == (cout_pad == 16), "conv: cout 16 <=> NCHW f32 output");
  TG_CHECK_ARG(!(out_mode == kOutNHWCf32 && (relu != kActNone || resid)), "conv: the f32 NHWC output is raw (no activation / residual)");
  TG_CHECK_ARG(amode == TG_AMODE_HALO || amode == TG_AMODE_DX3, "conv: bad amode %d", amode);
  TG_CHECK_ARG(relu == kActNone || relu == kActRelu || relu == kActLrelu02, "conv: bad activation %d", relu);
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(packed_w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv: pointers must be 16-byte aligned");
  TG_CHECK_ARG(!(resid && (kind != kConv3x3 || out_mode != kOutNHWCbf16)), "conv: residual only for conv3x3 NHWC");
  const bool s2 = (kind == kConv4x4s2);
  TG_CHECK_ARG(!s2 || ((h % 2) == 0 && (w % 2) == 0), "conv4x4s2: input size %dx%d must be even", h, w);
  if (s2) amode = TG_AMODE_HALO;

  const int nt = cout_pad == 16 ? 16 : 64;
  const int chunks = cout_pad / nt;
  // CTA pairs (cta_group::2) for the 64-wide chunks; TG_CONV_PAIR=0 selects the single-CTA kernel (A/B measurements)
  static const bool pair_on = []() { const char* e = getenv("TG_CONV_PAIR"); return !(e && e[0] == '0'); }();
  const bool pair = pair_on && nt == 64;
  const int wdiv = pair ? 2 : 1;                            // each CTA of a pair holds half of the weight rows
  // tile space: the output resolution for the stride-2 conv, the input resolution otherwise
  const int th = s2 ? h / 2 : h, tw = s2 ? w / 2 : w;

  TcParams p{};
  p.n = n; p.h = th; p.w = tw;
  p.tiles_x = tg_div_up(tw, kTileW);
  p.tiles_y = tg_div_up(th, kTileH);
  p.num_items = n * p.tiles_x * p.tiles_y;
  p.kchunks = cin_pad / 64;
  p.nphase = s2 ? 4 : 1;
  p.stages_per_item = p.kchunks * p.nphase;
  p.in_scale = s2 ? 2 : 1;
  p.ntaps = (s2 || tph) ? 4 : 9;
  // staged box in pixels: tile + halo; the stride-2 conv stages one input-parity phase (every 2nd pixel)
  const int box_w = s2 ? kTileW + 1 : ((amode == TG_AMODE_HALO) ? kTileW + 2 : kTileW);
  const int box_h = s2 ? kTileH + 1 : kTileH + 2;
  p.copy_bytes = static_cast<uint32_t>(box_w * box_h * 128);
  const uint32_t row_pitch = static_cast<uint32_t>(box_w * 128);
  if (amode == TG_AMODE_HALO) {
    p.ncopies = 1;
    p.sbo = row_pitch;
  } else {
    p.ncopies = (kind == kConv3x3) ? 3 : 2;
    p.sbo = 1024;
  }
  // tap tables: (dy,dx) are offsets inside the staged box, in MMA issue order == packed block order
  static const int conv_dy[9] = {0, 0, 0, 1, 1, 1, 2, 2, 2};
  static const int conv_dx[9] = {0, 1, 2, 0, 1, 2, 0, 1, 2};
  const int* ct_dy = kCtDy; const int* ct_dx = kCtDx; const int* ct_acc = kCtAcc; const int* ct_first = kCtFirst;
  for (int j = 0; j < p.ntaps; ++j) {
    int dy, dx;
    if (s2) { dy = j >> 1; dx = j & 1; }                    // kernel tap (2*dy + phase_y, 2*dx + phase_x)
    else if (tph) { dy = (j >> 1) + (phase >> 1); dx = (j & 1) + (phase & 1); }   // dY rows {i-1,i} (even phase) / {i,i+1} (odd)
    else if (kind == kConv3x3) { dy = conv_dy[j]; dx = conv_dx[j]; }
    else { dy = ct_dy[j]; dx = ct_dx[j]; }
    p.taps[j].a_off = (amode == TG_AMODE_HALO) ? static_cast<uint32_t>((dy * box_w + dx) * 128)
                                               : static_cast<uint32_t>(dx * p.copy_bytes + dy * row_pitch);
    p.taps[j].acc = (kind == kConvT3x3s2) ? ct_acc[j] : 0;
    p.taps[j].first = (kind == kConvT3x3s2) ? ct_first[j] : (j == 0);
  }
  const int origin = (kind == kConvT3x3s2) ? 0 : -1;   // conv: box starts at (x0-1,y0-1) (padding 1); convT: (x0,y0)
  for (int c = 0; c < 3; ++c) p.copy_dx[c] = origin + ((amode == TG_AMODE_HALO) ? 0 : c);
  p.box_y0 = origin;
  p.n_acc = (kind == kConvT3x3s2) ? 4 : 1;
  p.stage_bytes = p.ncopies * p.copy_bytes;
  p.a_region = (p.stage_bytes + 1023u) & ~1023u;
  // 4x4 weights (16 taps x Cin x 64) do not fit next to the A ring: stream the block of each (K chunk, phase)
  p.w_stage_bytes = s2 ? static_cast<uint32_t>(p.ntaps * nt * 128 / wdiv) : 0u;
  p.stage_stride = p.a_region + p.w_stage_bytes;
  p.ngroups = 8 / p.n_acc;
  p.w_bytes = s2 ? 0u : static_cast<uint32_t>(p.stages_per_item * p.ntaps * nt * 128 / wdiv);
  // A ring depth from what is left of the 227 KB
  int nstages = 8;
  while (nstages > 0 && make_layout(p.w_bytes, p.stage_stride, nstages, p.ngroups, nt).total + 1024 > kSmemLimit)
    --nstages;
  TG_CHECK_ARG(nstages >= 2 || (nstages >= 1 && p.stages_per_item == 1),
               "conv: shared memory too small for cin=%d cout=%d amode=%d (stages=%d)", cin_pad, cout_pad, amode, nstages);
  p.nstages = nstages;
  const SmemLayout L = make_layout(p.w_bytes, p.stage_stride, nstages, p.ngroups, nt);
  // request > half of the SM so that two CTAs never share an SM (each allocates all 512 TMEM cols)
  uint32_t smem_bytes = L.total + 1024;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;

  p.out_mode = out_mode;
  p.sy = p.sx = (kind == kConvT3x3s2 || tph) ? 2 : 1;
  p.oh = th * p.sy; p.ow = tw * p.sx;
  p.oc = nchw_out ? 3 : cout_pad;
  for (int a = 0; a < kMaxAcc; ++a) { p.acc_oy[a] = (kind == kConvT3x3s2) ? (a >> 1) : 0; p.acc_ox[a] = (kind == kConvT3x3s2) ? (a & 1) : 0; }
  if (tph) { p.acc_oy[0] = phase >> 1; p.acc_ox[0] = phase & 1; }
  p.out_nstride = out_nstride > 0 ? out_nstride : static_cast<long long>(p.oc) * p.oh * p.ow;
  p.relu = relu;
  p.out = out; p.out2 = out2; p.resid = resid; p.bias = bias; p.mask = mask; p.mask_mode = mask_mode;

  CUtensorMap tm_a, tm_w;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(cin_pad), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(cin_pad) * 2, static_cast<cuuint64_t>(w) * cin_pad * 2,
                             static_cast<cuuint64_t>(h) * w * cin_pad * 2};
    // stride-2 conv: boxDim = pixels * elementStride; every 2nd pixel is loaded -> box_w x box_h pixels land in smem
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_w * p.in_scale), static_cast<cuuint32_t>(box_h * p.in_scale), 1};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(p.in_scale), static_cast<cuuint32_t>(p.in_scale), 1};
    int rc = encode_bf16(&tm_a, x, 4, dims, strides, box, estr);
    if (rc) return rc;
  }
  {
    const int rows = chunks * p.stages_per_item * p.ntaps * nt;
    cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(nt / wdiv)};
    int rc = encode_bf16(&tm_w, packed_w, 2, dims, strides, box);
    if (rc) return rc;
  }

  int per_chunk = tg_num_sms() / chunks;
  if (per_chunk < 1) per_chunk = 1;
  int gx = p.num_items < per_chunk ? p.num_items : per_chunk;
  if (pair) gx = (gx + 1) & ~1;                             // whole 2-CTA clusters (an odd item count is padded in-kernel)
  if (pair && gx > per_chunk) gx -= 2;
  if (pair && gx < 2) gx = 2;
  dim3 grid(gx, chunks);
  static TgPerDeviceOnce attr_once[3];
  // algorithmic FLOPs (MAC = 2) on the padded channel counts; bench.py uses SURVEY.md's unpadded figure
  tg_prof_pre(nt == 64 ? TG_K_CONV64 : TG_K_CONV16,
              2.0 * (s2 ? 16.0 : 9.0) * cin_pad * (nt == 64 ? cout_pad : 3) * n * th * tw, stream);
  static const bool use_pdl = []() { const char* e = getenv("TG_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (pair) {
    if (attr_once[2].need()) { TG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)); }
    TG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, true>, tm_a, tm_w, p));
  } else if (nt == 64) {
    if (attr_once[0].need()) { TG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)); }
    TG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, false>, tm_a, tm_w, p));
  } else {
    if (attr_once[1].need()) { TG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)); }
    TG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, false>, tm_a, tm_w, p));
  }
  tg_prof_post(stream);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

size_t packed_weight_bytes_k(int kind, int cin_pad, int cout_pad) {
  return static_cast<size_t>(kind == kConv4x4s2 ? 16 : 9) * cin_pad * cout_pad * 2;
}

}  // namespace tg
