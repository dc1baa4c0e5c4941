// Persistent whole-generator kernel for sm_100a: the 41 convolutions of generator.forward
// (reference code/models.py:78-86) as ONE launch of one CTA per SM.
//
// Why: at 320x180 the 33 trunk layers are 3.6 us of tensor work each; as separate launches they
// cost 13-18 us (launch, prologue, weight fetch, pipeline fill/drain, 3.04-wave quantisation; see
// profiles/r01_summary_v1_perlayer.md).  Here the whole frame is one ordered queue of work items
//     item = (segment = layer x 64-wide Cout chunk, image, pixel tile)
// statically dealt round-robin to the CTAs, and layers are chained by per-tile completion counters
// in global memory instead of kernel boundaries:
//   * warp roles (20 warps): TMA producer, single-thread MMA issuer, 16 epilogue warps in two sets of 8 (set s drains
//     accumulator group s = every other item: two items' epilogues in flight; a warp = one TMEM lane quarter x 32
//     output channels), one publisher warp (lane s publishes set s's tiles), one dependency warp;
//   * an item's TMA producer waits until the <= 3x3 producer tiles of the previous layer that its halo box touches
//     have been published (one red.release.gpu per tile by a publisher lane, after the set's 8 warps arrived on a
//     CTA-local mbarrier; the acquire side - relaxed polls + one fence - runs ahead in the dependency warp and is
//     handed over through an mbarrier ring), then issues the box load; MMAs and epilogues of earlier items never
//     wait on later ones, all CTAs are co-resident (1 per SM) and every dependency points to an earlier item of the
//     queue, so the schedule cannot deadlock;
//   * weights live in two shared-memory slots (72 KB each, 36 KB per CTA of a pair) managed as an LRU pair: a layer with one K chunk
//     prefetches the next layer's block into the idle slot while it computes; a layer with two K
//     chunks keeps both blocks resident.  A block is released (tcgen05.commit -> mbarrier) after the
//     CTA's last item of the segment;
//   * the tensor pipe therefore never drains between layers; TMEM, barriers and tensor maps are
//     set up once per frame.
// The GEMM core: A = activation halo box (TMA, SWIZZLE_128B), B = weights, accumulators in TMEM, K = 16 per MMA.
// 3x3 convs: 4 x 32-pixel "wide" tiles, the three taps of a filter row stacked along N (one MMA of N = 192 per
// filter row and K step, partial sums combined by lane shuffles in the epilogue); transposed convs: 16 x 8 tiles,
// nine row-shifted descriptor views of the box into four output-phase accumulators (N = 64), as in tg_conv_tc.cu.
//
// Pair mode (template kPair, the default): the CTAs of two SMs of a TPC form a cluster and run their items in
// lock-step as ONE tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own item's activation box and HALF of
// the weight rows (N/2), the leader CTA issues for both, every CTA keeps its own 128 accumulator lanes and its own
// epilogue / publisher / dependency warps.  With both operands in shared memory a single-CTA MMA costs 43 + N/2
// cycles (the A fetch is not hidden), the paired one N/2 (profiles/r01_mma_microbench_v2.txt: N=192 138 -> 96 cycles
// for twice the work).  Segments are padded to an even item count; the padding ("phantom") item of the odd CTA loads
// an out-of-range box (zero fill), computes, and is dropped by the epilogue.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "tg_frame.cuh"

namespace tg {

constexpr int kEpiWarps = 16;                          // 4 per TMEM lane quarter: 16 accumulator columns each
// Epilogue organisation: two SETS of 8 warps, set s drains accumulator group s, i.e. every other item, so two items'
// epilogues are in flight at once (a warp = lane quarter x 32 columns, worked off as two 16-column passes with the three
// partial sums loaded one after the other into the same registers).  (All 16 warps on every item measured 8.5 % slower.)
constexpr int kSetWarps = kEpiWarps / 2;              // arrivals per accumulator hand-back / publish
constexpr int kPubWarps = 1;                          // (a second publisher WARP would be the 21st: 80 registers, spills - measured slower)
constexpr int kFrThreads = 32 * (2 + kEpiWarps + kPubWarps + 1);   // producer, MMA, epilogue warps, publisher(s), dependency warp
constexpr int kDepRing = 4;                            // dependency warp runs at most this many items ahead of the producer
constexpr uint32_t kFrSmemLimit = 232448;
constexpr uint32_t kAStride = 24576;                  // stage pitch: {64ch, 32, 6} wide box (24576 B) / {64ch, 10, 18} tall box (23040 B)
constexpr uint32_t kBarBytes = 512;                   // mbarrier area
constexpr int kWBoxRows = 48;                         // weight TMA box: 48 rows x 128 B
template <bool kPair> struct FrCfg {
  static constexpr uint32_t kWSlotBytes = kPair ? 36864u : 73728u;   // 9 taps x 64 x 64 bf16 (half of the rows per CTA of a pair)
  static constexpr int kStages = kPair ? 6 : 3;                       // A ring depth
  static constexpr uint32_t kSmemBytes = 2 * kWSlotBytes + kStages * kAStride + kBarBytes + 16 + kEpiWarps * 64 * 4 +
                                         kFrMaxSegs * sizeof(FrSegS) + 1024;
};
static_assert(FrCfg<true>::kSmemBytes <= kFrSmemLimit && FrCfg<false>::kSmemBytes <= kFrSmemLimit, "smem budget");
constexpr int kGroupCols = 256;                       // TMEM columns per accumulator group (2 groups)
constexpr uint32_t kItemDone = 128;                   // counter value of a published tile

// tap tables for the per-tap paths: [2]: 3x3 conv on the WIDE box, one view per tap (TG_FRAME_TAP=1; row pitch 32 pixels
// = 4096 B); [0]: 3x3 conv on the tall box (TG_FRAME_WIDE=0).  The transposed conv is issued as four N-stacked shift groups.
__constant__ uint32_t c_aoff[3][9] = {
    {0 * 128, 1 * 128, 2 * 128, 10 * 128, 11 * 128, 12 * 128, 20 * 128, 21 * 128, 22 * 128},
    {0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0 * 128, 1 * 128, 2 * 128, 32 * 128, 33 * 128, 34 * 128, 64 * 128, 65 * 128, 66 * 128}};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// L2-only load: the residual buffer is rewritten every third layer, L1 must not serve a stale line
__device__ __forceinline__ void ld_global_cg_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p)
               : "memory");
}

__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}


// Packed f32x2 arithmetic (FADD2): the epilogue's instruction count, not its math, is what it pays for.
__device__ __forceinline__ uint64_t f2_pack(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(uint64_t v) {
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
// 16 f32 accumulator columns of one pixel (8 f32x2 pairs) -> + bias, ReLU, + residual -> 16 bf16 channels (one 32-byte
// store, NHWC).  ReLU commutes with the rounding to bf16, so it is one bf16x2 max per pair; a layer has a ReLU or a
// residual, never both (launch_frame checks).
__device__ __forceinline__ void epi_store_bf16(uint64_t (&a)[8], const float* bias16, uint8_t* dst, const uint8_t* res,
                                               bool relu, const uint8_t* msk = nullptr) {
  const uint64_t* b2 = reinterpret_cast<const uint64_t*>(bias16);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = f2_add(a[e], b2[e]);
  uint32_t o[8];
  if (res) {
    uint32_t rv[8];
    ld_global_cg_v8(res, rv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = f2_to_bf16x2(f2_add(a[e], f2_pack(rv[e] << 16, rv[e] & 0xFFFF0000u)));
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = f2_to_bf16x2(a[e]);
      if (relu) asm("max.bf16x2 %0, %1, %2;" : "=r"(o[e]) : "r"(o[e]), "r"(0u));
    }
  }
  if (msk) {
    // backward of the ReLU that produced the saved activation `msk` (same layout as the output): the gradient passes where
    // the activation is non-zero (tg_conv_tc.cu: kMaskRelu).  The activation belongs to an earlier launch: plain L2 load.
    uint32_t mv[8];
    ld_global_cg_v8(msk, mv);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      o[e] &= ((mv[e] & 0x7FFFu) ? 0xFFFFu : 0u) | ((mv[e] & 0x7FFF0000u) ? 0xFFFF0000u : 0u);
  }
  st_global_v8(dst, o);
}

// Store one channel of one pixel of the network output in the segment's format.  o = n * out_nstride + y * ow + x
// (+ c * plane for the planar formats); the uint8 format is pixel-interleaved: element (o_px * 3 + c).
__device__ __forceinline__ void store_network_output(void* out, int mode, size_t o_planar, size_t o_px, int c, float y) {
  if (mode == kOutNCHWf32Sigmoid) static_cast<float*>(out)[o_planar] = y;
  else if (mode == kOutNCHWf16Sigmoid) static_cast<__half*>(out)[o_planar] = __float2half_rn(y);
  else static_cast<uint8_t*>(out)[o_px * 3 + c] = static_cast<uint8_t>(__float2uint_rz(__fmul_rn(y, 255.f)));
}

// Stall accounting (measurement only, FrProgram::stats != null): cycles a role spends in each of its waits.
template <bool kOn>
struct Tick {
  long long t;
  bool on;        // stats requested
  bool gate;      // current item belongs to the selected segment (FrProgram::stat_seg, -1 = every segment)
  __device__ __forceinline__ void start() { if (kOn && on) t = clock64(); }
  __device__ __forceinline__ void stop(long long& acc) { if (kOn && on && gate) acc += clock64() - t; }
};

// Weight-slot LRU pair.  Producer and MMA warps run the same deterministic state machine over the
// same item sequence, so they agree on slot and load parity without communicating.
struct WSlots {
  int blk0, blk1, mru;
  uint32_t loads0, loads1;
  __device__ __forceinline__ void init() { blk0 = blk1 = -1; mru = 1; loads0 = loads1 = 0; }
  // returns slot; *is_load = block had to be (re)loaded; *nth = number of earlier loads into the slot
  __device__ __forceinline__ int use(int b, bool* is_load, uint32_t* nth) {
    int s;
    if (blk0 == b) { s = 0; *is_load = false; }
    else if (blk1 == b) { s = 1; *is_load = false; }
    else {
      s = 1 - mru;
      if (s == 0) blk0 = b; else blk1 = b;
      *is_load = true;
    }
    *nth = s ? loads1 : loads0;
    if (*is_load) { if (s == 0) ++loads0; else ++loads1; }
    mru = s;
    return s;
  }
};

// Batch decode shared by the producer and the dependency warp: lane j owns item it0 + j*G of this CTA.
//   l0 = segment | image << 8          l1 = box origin x (16 bit) | y << 16
//   l2 = first producer tile x | y << 16     l3 = nx | ny << 4 | (65535/nx + 1) << 8   (0: nothing to wait for)
__device__ __forceinline__ void decode_batch(const FrProgram& P, int it0, int lane, int G, int si_base, uint32_t& l0,
                                             uint32_t& l1, uint32_t& l2, uint32_t& l3, int& l_si) {
  const int my_it = it0 + lane * G;
  l0 = l1 = l2 = l3 = 0;
  l_si = si_base;
  if (my_it < P.total_items) {
    while (my_it >= P.segs[l_si].item_end) ++l_si;
    const FrSeg& S = P.segs[l_si];
    const uint32_t local = static_cast<uint32_t>(my_it - S.item_begin);
    const uint32_t r = fdiv(local, S.fd_tiles_x);
    const int tx = static_cast<int>(local - r * S.tiles_x);
    const uint32_t n = fdiv(r, S.fd_tiles_y);
    const int ty = static_cast<int>(r - n * S.tiles_y);
    const int origin = (S.kind == kConv3x3) ? -1 : 0;
    const int bx0 = tx * S.tile_w + origin, by0 = ty * S.tile_h + origin;
    l0 = static_cast<uint32_t>(l_si) | (n << 8);
    l1 = (static_cast<uint32_t>(bx0) & 0xFFFFu) | (static_cast<uint32_t>(by0) << 16);
    if (S.dep_nseg > 0 && local < static_cast<uint32_t>(S.items_real)) {   // (a pair's padding item waits for nothing)
      // producer tiles of the previous layer touched by the halo box (clipped to the image)
      const int ya = max(by0, 0), yb = min(by0 + S.box_h - 1, S.h - 1);
      const int xa = max(bx0, 0), xb = min(bx0 + S.box_w - 1, S.w - 1);
      const uint32_t tya = fdiv(ya, S.fd_dep_th), tyb = fdiv(yb, S.fd_dep_th);
      const uint32_t txa = fdiv(xa, S.fd_dep_tw), txb = fdiv(xb, S.fd_dep_tw);
      const uint32_t nx = txb - txa + 1, ny = tyb - tya + 1;
      l2 = txa | (tya << 16);
      l3 = nx | (ny << 4) | ((65535u / nx + 1u) << 8);          // q / nx == (q * inv) >> 16 for q < 256
    }
  }
}

// kDbg: the measurement build (TG_FRAME_DBG knobs, segment trace, stall accounting); the production build has none of it.
// kMask: the data-gradient build (tg_gen_backward): bf16 layers may carry a ReLU mask (FrLayer::mask) applied after the residual.
template <bool kPair, bool kDbg, bool kMask = false>
__global__ void __launch_bounds__(kFrThreads, 1) frame_kernel(const __grid_constant__ FrProgram P) {
  const int dbg = kDbg ? P.dbg : 0;
  unsigned long long* const trace = kDbg ? P.trace : nullptr;
  unsigned long long* const stats = kDbg ? P.stats : nullptr;
  const int stat_seg = kDbg ? P.stat_seg : -1;
  constexpr uint32_t kWSlotBytes = FrCfg<kPair>::kWSlotBytes;
  constexpr int kFrStages = FrCfg<kPair>::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;     // the same offset in both CTAs of a pair (same kernel, same layout)
  uint8_t* gbase = smem_raw + (base - raw);

  const uint32_t s_w = base;
  const uint32_t s_a = base + 2 * kWSlotBytes;
  const uint32_t bar0 = s_a + kFrStages * kAStride;
  const uint32_t bar_wfull = bar0, bar_wempty = bar0 + 16;
  const uint32_t bar_afull = bar0 + 32, bar_aempty = bar_afull + 8 * kFrStages;
  const uint32_t bar_cfull = bar_aempty + 8 * kFrStages, bar_cempty = bar_cfull + 16;
  const uint32_t bar_pfull = bar_cempty + 16, bar_pempty = bar_pfull + 16;
  const uint32_t bar_dfull = bar_pempty + 16, bar_dempty = bar_dfull + 8 * kDepRing;
  static_assert(32 + 16 * FrCfg<kPair>::kStages + 64 + 16 * kDepRing <= kBarBytes, "barrier area");
  const uint32_t off_misc = (bar0 + kBarBytes) - base;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gbase + off_misc);
  float* s_bias_all = reinterpret_cast<float*>(gbase + off_misc + 16);   // [kEpiWarps][64]
  FrSegS* segs = reinterpret_cast<FrSegS*>(gbase + off_misc + 16 + kEpiWarps * 64 * 4);   // [nseg]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Pair mode: rank 0 of the cluster (the leader) issues the MMAs.  The "full" barriers (weights, activation stages)
  // and the "accumulator drained" barriers the issuer waits on live in the leader and take one arrival per CTA;
  // everything the MMA completion signals (stage / weight slot free, accumulator ready) is multicast to both CTAs.
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t nctas = kPair ? 2u : 1u;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_wfull + 8 * i, nctas);
      mbar_init(bar_wempty + 8 * i, 1);
      mbar_init(bar_cfull + 8 * i, 1);
      mbar_init(bar_cempty + 8 * i, kSetWarps * nctas);
      mbar_init(bar_pfull + 8 * i, kSetWarps);
      mbar_init(bar_pempty + 8 * i, 1);
    }
    for (int i = 0; i < kFrStages; ++i) {
      mbar_init(bar_afull + 8 * i, nctas);
      mbar_init(bar_aempty + 8 * i, 1);
    }
    for (int i = 0; i < kDepRing; ++i) {
      mbar_init(bar_dfull + 8 * i, 1);
      mbar_init(bar_dempty + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  }
  for (int i = threadIdx.x; i < P.nseg; i += kFrThreads) {
    const FrSeg& S = P.segs[i];
    FrSegS c;
    c.item_begin = S.item_begin; c.item_end = S.item_end; c.items_real = S.items_real;
    c.tiles_x = static_cast<uint16_t>(S.tiles_x); c.tiles_y = static_cast<uint16_t>(S.tiles_y);
    c.h = static_cast<uint16_t>(S.h); c.w = static_cast<uint16_t>(S.w);
    c.oh = static_cast<uint16_t>(S.oh); c.ow = static_cast<uint16_t>(S.ow);
    c.oc = static_cast<uint16_t>(S.oc); c.ch0 = static_cast<uint16_t>(S.ch0);
    c.wide = static_cast<uint8_t>(S.wide); c.out_mode = static_cast<uint8_t>(S.out_mode); c.relu = static_cast<uint8_t>(S.relu);
    c.kind = static_cast<uint8_t>(S.kind); c.nt = static_cast<uint8_t>(S.nt); c.kchunks = static_cast<uint8_t>(S.kchunks);
    c.pad0 = c.pad1 = 0;
    c.fd_tiles_x = S.fd_tiles_x; c.fd_tiles_y = S.fd_tiles_y;
    c.out_nstride = S.out_nstride; c.out = S.out; c.out2 = S.out2; c.resid = S.resid; c.bias = S.bias;
    segs[i] = c;
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();                    // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // leader-side barrier addresses (shared::cluster window; the CTA's own addresses when not paired)
  const uint32_t lbar_wfull = kPair ? mapa_rank(bar_wfull, 0) : bar_wfull;
  const uint32_t lbar_afull = kPair ? mapa_rank(bar_afull, 0) : bar_afull;
  const uint32_t lbar_cempty = kPair ? mapa_rank(bar_cempty, 0) : bar_cempty;

  const int G = gridDim.x;

  if (warp == 0) {
    // ================================ TMA producer =========================================
    // The producer is the one strictly serial role (item -> dependencies -> slot -> TMA), so its per-item latency
    // bounds the item rate of the CTA.  Items are decoded 32 at a time, one per lane (segment search, tile
    // coordinates), and broadcast by shuffles; the dependency wait itself (a gpu-scope acquire of up to 24 completion
    // counters, an L2 round trip plus a fence) is done ahead of time by the dependency warp and handed over through
    // a CTA-local mbarrier ring, which keeps the acquire -> TMA ordering (mbarrier arrive = release.cta, wait =
    // acquire.cta; causality order is transitive).
    if (lane < P.nseg) tma_prefetch_desc(&P.maps[P.segs[lane].map_a]);
    if (lane == 0) tma_prefetch_desc(&P.maps[0]);
    WSlots ws;
    ws.init();
    int st = 0;
    uint32_t ph = 0, dk = 0;
    bool grid_waited = false;
    int si_base = 0;
    int p_si = -1, p_kch = 1;
    uint32_t p_box_bytes = 0;
    const CUtensorMap* p_map = &P.maps[0];
    Tick<kDbg> tk{0, stats != nullptr, true}, tall{0, stats != nullptr, true};
    long long a_wempty = 0, a_dep = 0, a_aempty = 0, a_total = 0, a_issue = 0;
    tall.start();
    for (int it0 = blockIdx.x; it0 < P.total_items; it0 += 32 * G) {
      uint32_t l0, l1, l2, l3;
      int l_si;
      decode_batch(P, it0, lane, G, si_base, l0, l1, l2, l3, l_si);
      const int left = (P.total_items - it0 + G - 1) / G;
      const int nb = left < 32 ? left : 32;
      for (int j = 0; j < nb; ++j) {
        const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, l0, j), b1 = __shfl_sync(0xFFFFFFFFu, l1, j);
        const bool has_dep = __shfl_sync(0xFFFFFFFFu, l3, j) != 0;
        const int si = static_cast<int>(b0 & 0xFFu), n = static_cast<int>(b0 >> 8);
        tk.gate = stat_seg < 0 || si == stat_seg;
        const int bx0 = static_cast<int>(static_cast<int16_t>(b1 & 0xFFFFu)), by0 = static_cast<int>(b1) >> 16;
        const FrSeg& S = P.segs[si];
        if (si != p_si) {                                    // per-segment values in registers: the item path of this strictly
          p_si = si;                                         // serial role reads no kernel-parameter memory
          p_kch = S.kchunks;
          p_box_bytes = static_cast<uint32_t>(S.box_w * S.box_h) * 128u;
          p_map = &P.maps[S.map_a];
        }
        for (int kc = 0; kc < p_kch; ++kc) {
          bool is_load;
          uint32_t nth;
          const int slot = ws.use(si * 2 + kc, &is_load, &nth);
          if (is_load) {
            tk.start();
            mbar_wait(bar_wempty + 8 * slot, (nth & 1) ^ 1);
            tk.stop(a_wempty);
            if (elect_one()) {
              if (kPair) {
                // this CTA's half of every N group: rows [rank * half, (rank + 1) * half) of the group's 2 * half rows
                const uint32_t box = S.w_box_rows;
                uint32_t dst_row = 0, src_row = 0;
                for (int gi = 0; gi < S.w_groups; ++gi) dst_row += static_cast<uint32_t>(S.w_grp_half[gi]);
                mbar_expect_tx_cluster(lbar_wfull + 8 * slot, dst_row * 128u);
                dst_row = 0;
                for (int gi = 0; gi < S.w_groups; ++gi) {
                  const uint32_t half = static_cast<uint32_t>(S.w_grp_half[gi]);
                  for (uint32_t r0 = 0; r0 < half; r0 += box)
                    tma_load_2d_pair(s_w + slot * kWSlotBytes + (dst_row + r0) * 128, &P.maps[S.w_map], lbar_wfull + 8 * slot, 0,
                                     static_cast<int>(S.w_row0[kc] + src_row + rank * half + r0));
                  dst_row += half;
                  src_row += 2 * half;
                }
              } else {
                mbar_expect_tx(bar_wfull + 8 * slot, S.w_rows * 128);
                for (uint32_t r0 = 0; r0 < S.w_rows; r0 += kWBoxRows)
                  tma_load_2d(s_w + slot * kWSlotBytes + r0 * 128, &P.maps[0], bar_wfull + 8 * slot, 0,
                              static_cast<int>(S.w_row0[kc] + r0));
              }
            }
            __syncwarp();
          }
          if (!grid_waited) {   // weights are constants; activations of a previous kernel are not
            asm volatile("griddepcontrol.wait;" ::: "memory");
            grid_waited = true;
          }
          if (kc == 0 && has_dep && !(dbg & 1)) {                  // dependencies acquired by the dependency warp
            const uint32_t ds = dk % kDepRing;
            tk.start();
            mbar_wait(bar_dfull + 8 * ds, (dk / kDepRing) & 1u);
            tk.stop(a_dep);
            __syncwarp();                                            // every lane has seen this phase before it can be reused
            if (lane == 0) mbar_arrive(bar_dempty + 8 * ds);
            ++dk;
          }
          tk.start();
          mbar_wait(bar_aempty + 8 * st, ph ^ 1);
          tk.stop(a_aempty);
          tk.start();
          if (elect_one()) {
            if (!(dbg & 64)) fence_proxy_async_global();   // generic-proxy writes of other CTAs (acquired above) -> async-proxy read
            if (kPair && (dbg & 1024) && tk.gate) {
              mbar_expect_tx_cluster(lbar_afull + 8 * st, 0u);         // measurement: no A load in the selected segment
            } else if (kPair) {
              mbar_expect_tx_cluster(lbar_afull + 8 * st, p_box_bytes);
#ifdef TG_EXP_OUT_SAMEBOX
              // measurement build: every activation load of the segments selected by the mask reads the SAME box (always an L2
              // hit) - is the segment waiting for HBM?  bit 0: output conv (nt 16), bit 1: 128-input-channel layers
              if (((TG_EXP_OUT_SAMEBOX & 1) && S.nt == 16) || ((TG_EXP_OUT_SAMEBOX & 2) && S.kchunks == 2))
                tma_load_4d_pair(s_a + st * kAStride, p_map, lbar_afull + 8 * st, kc * 64, -1, -1, 0);
              else
#endif
              tma_load_4d_pair(s_a + st * kAStride, p_map, lbar_afull + 8 * st, kc * 64, bx0, by0, n);
            } else {
              mbar_expect_tx(bar_afull + 8 * st, p_box_bytes);
              tma_load_4d(s_a + st * kAStride, p_map, bar_afull + 8 * st, kc * 64, bx0, by0, n);
            }
          }
          __syncwarp();
          tk.stop(a_issue);
          if (++st == kFrStages) { st = 0; ph ^= 1; }
        }
      }
      si_base = __shfl_sync(0xFFFFFFFFu, l_si, nb - 1);
    }
    tall.stop(a_total);
    if (stats && lane == 0) {
      unsigned long long* o = stats + blockIdx.x * 32;
      o[9] = a_wempty; o[10] = a_dep; o[11] = a_aempty; o[12] = a_total; o[22] = a_issue;
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA of a pair) ====================
    if (rank == 0) {
    WSlots ws;
    ws.init();
    int si = 0, st = 0, g = 0;
    uint32_t ph = 0, gph = 0;
    int traced_si = -1;
    Tick<kDbg> tk{0, stats != nullptr, true}, tall{0, stats != nullptr, true};
    long long a_cempty = 0, a_wfull = 0, a_afull = 0, a_total = 0, a_issue = 0;
    tall.start();
    // per-segment values live in registers (no shared-memory round trip on the issue path of every item)
    int seg_end = 0, m_wide = 0, m_kchunks = 1, tbl = 0;
    uint32_t idesc = 0, sbo = 0, wtap_bytes = 0;
    for (int it = blockIdx.x; it < P.total_items; it += G) {
      if (it >= seg_end) {
        while (it >= segs[si].item_end) ++si;
        const FrSegS& C = segs[si];
        seg_end = C.item_end; m_wide = C.wide; m_kchunks = C.kchunks;
        idesc = umma_idesc_bf16(kPair ? 256 : 128, C.wide == 1 ? 3 * C.nt : C.nt);
        tbl = C.wide == 2 ? 2 : C.kind;                        // tap table (views inside the staged box)
        sbo = C.wide == 2 ? 1024u : 10u * 128u;                // byte pitch of 8-pixel groups: wide rows are 32 pixels
        wtap_bytes = static_cast<uint32_t>(C.nt) * (kPair ? 64 : 128);   // rows of one tap held by this CTA x 128 B
      }
      if (trace && si != traced_si) {
        if (lane == 0) trace[static_cast<size_t>(si) * G + blockIdx.x] = global_ns();
        traced_si = si;
      }
      tk.gate = stat_seg < 0 || si == stat_seg;
      const bool last_in_seg = (it + G >= seg_end);
      const bool wide = m_wide == 1;
      tk.start();
      if (!(dbg & 8)) {
        if (kPair && !(dbg & 32)) mbar_wait_cluster(bar_cempty + 8 * g, gph ^ 1);
        else mbar_wait(bar_cempty + 8 * g, gph ^ 1);
      }
      tk.stop(a_cempty);
      tc_fence_after();
      const uint32_t d_base = tmem_base + static_cast<uint32_t>(g * kGroupCols);
      for (int kc = 0; kc < m_kchunks; ++kc) {
        bool is_load;
        uint32_t nth;
        const int slot = ws.use(si * 2 + kc, &is_load, &nth);
        tk.start();
        if (is_load) { if (kPair && !(dbg & 32)) mbar_wait_cluster(bar_wfull + 8 * slot, nth & 1); else mbar_wait(bar_wfull + 8 * slot, nth & 1); }
        tk.stop(a_wfull);
        tk.start();
        if (kPair && !(dbg & 32)) mbar_wait_cluster(bar_afull + 8 * st, ph);
        else mbar_wait(bar_afull + 8 * st, ph);
        tk.stop(a_afull);
        tc_fence_after();
        tk.start();
        if (elect_one()) {
          const uint32_t a_base = s_a + st * kAStride;
          const uint32_t w_base = s_w + slot * kWSlotBytes;
          if ((dbg & 512) && tk.gate) {
            // measurement: no MMAs in the selected segment (commits only)
          } else if (wide) {
            // one MMA per (filter row, K=16 step): A = the box shifted by dy rows of 32 pixels (4096 B, so every
            // descriptor keeps the canonical 1024-byte group pitch), B = the row's three taps stacked along N.
            // Fully unrolled, descriptors advanced by immediates: for the N = 48 output conv the issue path, not
            // the MMAs, bounds the item rate.
            const uint64_t ad0 = umma_desc_sw128(a_base, 1024), bd0 = umma_desc_sw128(w_base, 1024);
            const uint64_t wrow = (3 * wtap_bytes) >> 4;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint64_t ad = ad0 + static_cast<uint64_t>((dy * kWideBoxW * 128) >> 4);
              const uint64_t bd = bd0 + dy * wrow;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kPair) umma_bf16_pair(d_base, ad + 2 * k, bd + 2 * k, idesc, (kc > 0 || dy > 0 || k > 0) ? 1u : 0u);
                else umma_bf16(d_base, ad + 2 * k, bd + 2 * k, idesc, (kc > 0 || dy > 0 || k > 0) ? 1u : 0u);
              }
            }
          } else if (m_wide == 2) {
            // one MMA group per tap on the wide box; everything but the two base descriptors is an immediate
            const uint64_t ad0 = umma_desc_sw128(a_base, 1024), bd0 = umma_desc_sw128(w_base, 1024);
            const uint64_t wstep = wtap_bytes >> 4;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const uint64_t ad = ad0 + static_cast<uint64_t>(((j / 3) * kWideBoxW * 128 + (j % 3) * 128) >> 4);
              const uint64_t bd = bd0 + j * wstep;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kPair) umma_bf16_pair(d_base, ad + 2 * k, bd + 2 * k, idesc, (kc > 0 || j > 0 || k > 0) ? 1u : 0u);
                else umma_bf16(d_base, ad + 2 * k, bd + 2 * k, idesc, (kc > 0 || j > 0 || k > 0) ? 1u : 0u);
              }
            }
          } else if (tbl == kConvT3x3s2) {
            // Transposed conv: the nine (phase, tap) products grouped by input shift (tg_conv_tc.cuh: kCt*): four A views
            // (shift (0,0), (0,1), (1,0), (1,1) inside the tall box) times the weight blocks of that shift stacked along N,
            // N = 256 / 128 / 128 / 64, onto accumulator columns 0 / 64 / 128 / 128 ([a0 | a1 | a3 | a2]).  16 MMAs per K
            // chunk instead of 36; per accumulator the products arrive in the per-layer kernel's order.
            constexpr int M = kPair ? 256 : 128;
            constexpr uint32_t kIdesc[4] = {umma_idesc_bf16(M, 256), umma_idesc_bf16(M, 128), umma_idesc_bf16(M, 128), umma_idesc_bf16(M, 64)};
            constexpr uint32_t kAOff[4] = {0u, 1u * 128u, 10u * 128u, 11u * 128u};
            constexpr uint32_t kDCol[4] = {0u, 64u, 128u, 128u};
            constexpr uint32_t kBRow[4] = {0u, 256u, 384u, 512u};             // first weight row of the group (whole block)
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              const uint64_t ad = umma_desc_sw128(a_base + kAOff[gi], sbo);
              const uint64_t bd = umma_desc_sw128(w_base + kBRow[gi] * (kPair ? 64u : 128u), 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (gi > 0 || kc > 0 || k > 0) ? 1u : 0u;
                if (kPair) umma_bf16_pair(d_base + kDCol[gi], ad + 2 * k, bd + 2 * k, kIdesc[gi], acc);
                else umma_bf16(d_base + kDCol[gi], ad + 2 * k, bd + 2 * k, kIdesc[gi], acc);
              }
            }
          } else {
#pragma unroll 1
            for (int j = 0; j < 9; ++j) {                        // 3x3 conv on the tall box (TG_FRAME_WIDE=0), one MMA group per tap
              const uint64_t ad = umma_desc_sw128(a_base + c_aoff[tbl][j], sbo);
              const uint64_t bd = umma_desc_sw128(w_base + j * wtap_bytes, 1024);
              const uint32_t keep = (kc > 0 || j > 0) ? 1u : 0u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kPair) umma_bf16_pair(d_base, ad + 2 * k, bd + 2 * k, idesc, (k > 0) ? 1u : keep);
                else umma_bf16(d_base, ad + 2 * k, bd + 2 * k, idesc, (k > 0) ? 1u : keep);
              }
            }
          }
          if (kPair) {
            umma_commit_pair(bar_aempty + 8 * st);
            if (last_in_seg) umma_commit_pair(bar_wempty + 8 * slot);
            if (kc == m_kchunks - 1) umma_commit_pair(bar_cfull + 8 * g);
          } else {
            umma_commit(bar_aempty + 8 * st);
            if (last_in_seg) umma_commit(bar_wempty + 8 * slot);
            if (kc == m_kchunks - 1) umma_commit(bar_cfull + 8 * g);
          }
        }
        __syncwarp();
        tk.stop(a_issue);
        if (++st == kFrStages) { st = 0; ph ^= 1; }
      }
      g ^= 1;
      if (g == 0) gph ^= 1;
    }
    tall.stop(a_total);
    if (stats && lane == 0) {
      unsigned long long* o = stats + blockIdx.x * 32;
      o[0] = a_cempty; o[1] = a_wfull; o[2] = a_afull; o[3] = a_total; o[20] = a_issue;
    }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ================================ epilogue, two sets of 8 warps =========================
    // Set s (warps 2+8s .. 9+8s) drains accumulator group s = the items k with (k & 1) == s of this CTA's sequence, so
    // the epilogue of item k+1 runs while item k's is still waiting on TMEM, shuffles or stores.  A warp owns one TMEM
    // lane quarter and 32 of the 64 output channels, in two passes of 16; per pass the partial sums P0, P1, P2 are
    // loaded one after the other (16 registers in flight instead of 48) and combined in the order (P0[x-1] + P1[x]) + P2[x+1].
    const int ew = warp - 2;
    const int set = ew >> 3;
    const int q = warp & 3;                                // TMEM lane quarter of this warp
    const int half = (ew >> 2) & 1;                        // which 32 of the 64 accumulator columns
    const int m = q * 32 + lane;
    const int pr = m >> 3, pc = m & 7;
    float* s_bias = s_bias_all + ew * 64;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int si = 0, cur_si = -1;
    uint32_t gph = 0, pk = 0;
    Tick<kDbg> tk{0, stats != nullptr, true}, tall{0, stats != nullptr, true};
    long long a_cfull = 0, a_body = 0, a_pub = 0, a_total = 0, a_tmem = 0, a_hb = 0;
    tall.start();
    float nb0 = 0.f, nb1 = 0.f;                            // bias of segment nb_si, fetched one segment ahead
    int nb_si = -1;
    int seg_end = 0, seg_begin = 0, seg_real = 0, c_tiles_x = 1, c_tiles_y = 1, c_h = 0, c_w = 0;
    FastDiv c_fdx{0, 0}, c_fdy{0, 0};
    uint32_t c_row_bytes = 0, c_px_bytes = 0, c_img_bytes = 0;
    uint8_t* c_out = nullptr;
    const uint8_t* c_res = nullptr;
    const uint8_t* c_msk = nullptr;
    int c_wide = 0, c_mode = 0, c_relu = 0, c_kind = 0;
    const uint32_t bar_cf = bar_cfull + 8 * set, bar_pf = bar_pfull + 8 * set, bar_pe = bar_pempty + 8 * set;
    const uint32_t bar_ce = lbar_cempty + 8 * set, bar_ce_local = bar_cempty + 8 * set;
    auto hand_back = [&]() {                               // TMEM group drained by this warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (kPair) mbar_arrive_cluster(bar_ce); else mbar_arrive(bar_ce_local); }
    };
    for (int it = blockIdx.x + set * G; it < P.total_items; it += 2 * G) {
      if (it >= seg_end) {
        while (it >= segs[si].item_end) ++si;
        const FrSegS& C = segs[si];
        seg_end = C.item_end; seg_begin = C.item_begin; seg_real = C.items_real;
        c_tiles_x = C.tiles_x; c_tiles_y = C.tiles_y; c_fdx = C.fd_tiles_x; c_fdy = C.fd_tiles_y;
        c_h = C.h; c_w = C.w; c_wide = C.wide; c_mode = C.out_mode; c_relu = C.relu; c_kind = C.kind;
        c_px_bytes = C.oc * 2u; c_row_bytes = C.ow * c_px_bytes; c_img_bytes = C.oh * c_row_bytes;   // < 4 GB (launch_frame)
        c_out = static_cast<uint8_t*>(C.out) + C.ch0 * 2u + half * 64u;
        c_res = C.resid ? static_cast<const uint8_t*>(C.resid) + C.ch0 * 2u + half * 64u : nullptr;
        // (bf16 layers have no f32 logits copy: out2 carries the ReLU mask of the data-gradient build)
        c_msk = (kMask && C.out2 && C.out_mode == kOutNHWCbf16) ? reinterpret_cast<const uint8_t*>(C.out2) + C.ch0 * 2u + half * 64u : nullptr;
      }
      const FrSegS& S = segs[si];
      if (si != cur_si) {                                  // warp-private bias copy of this segment
        __syncwarp();
        float b0 = nb0, b1 = nb1;
        if (nb_si != si) {                                 // (first segment, or the set had no item in a segment)
          b0 = lane < S.nt ? S.bias[lane] : 0.f;
          b1 = lane + 32 < S.nt ? S.bias[lane + 32] : 0.f;
        }
        s_bias[lane] = b0;
        s_bias[lane + 32] = b1;
        __syncwarp();
        cur_si = si;
        if (si + 1 < P.nseg) {                             // in flight until the next segment starts
          const FrSegS& N = segs[si + 1];
          nb0 = lane < N.nt ? N.bias[lane] : 0.f;
          nb1 = lane + 32 < N.nt ? N.bias[lane + 32] : 0.f;
          nb_si = si + 1;
        }
      }
      const uint32_t local = static_cast<uint32_t>(it - seg_begin);
      const uint32_t r = fdiv(local, c_fdx);
      const int tx = static_cast<int>(local - r * c_tiles_x);
      const int n = static_cast<int>(fdiv(r, c_fdy));
      const int ty = static_cast<int>(r) - n * c_tiles_y;
      const bool real = local < static_cast<uint32_t>(seg_real);      // false: the padding item of a pair
      tk.gate = stat_seg < 0 || si == stat_seg;
      tk.start();
      mbar_wait(bar_cf, gph);
      tk.stop(a_cfull);
      tk.start();
      tc_fence_after();
      const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(set * kGroupCols);
      if (c_wide == 1) {
        const int wy = ty * kWideH + q, wx = tx * kWideW - 1 + lane;
        const bool wvalid = real && (lane >= 1) && (lane <= kWideW) && (wy < c_h) && (wx < c_w);
        if (c_mode == kOutNHWCbf16) {
          const size_t off = static_cast<size_t>(n) * c_img_bytes + (static_cast<uint32_t>(wy) * c_row_bytes + static_cast<uint32_t>(wx) * c_px_bytes);
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            const uint32_t col = static_cast<uint32_t>(half * 32 + pass * 16);
            uint32_t v[16], vb[16];
            uint64_t a2[8];
            Tick<kDbg> t2{0, stats != nullptr, tk.gate};
            t2.start();
            tmem_ld_32x16(tq + col, v);                    // P0: the left neighbour's value is needed
            tmem_ld_32x16(tq + 64 + col, vb);              // P1 (same round trip)
            tmem_ld_wait();
            t2.stop(a_tmem);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              a2[e] = f2_add(f2_pack(__shfl_up_sync(0xFFFFFFFFu, v[2 * e], 1), __shfl_up_sync(0xFFFFFFFFu, v[2 * e + 1], 1)),
                             f2_pack(vb[2 * e], vb[2 * e + 1]));
            t2.start();
            tmem_ld_32x16(tq + 128 + col, v);              // P2: the right neighbour's value
            tmem_ld_wait();
            t2.stop(a_tmem);
            t2.start();
            if (pass == 1) hand_back();
            t2.stop(a_hb);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              a2[e] = f2_add(a2[e], f2_pack(__shfl_down_sync(0xFFFFFFFFu, v[2 * e], 1), __shfl_down_sync(0xFFFFFFFFu, v[2 * e + 1], 1)));
            if (wvalid && !(dbg & 16))
              epi_store_bf16(a2, s_bias + col, c_out + off + pass * 32u, c_res ? c_res + off + pass * 32u : nullptr, c_relu != 0,
                             (kMask && c_msk) ? c_msk + off + pass * 32u : nullptr);
          }
        } else {
          uint32_t v0[4], v1[4], v2[4];                    // 3 output channels of each of the 3 partial sums
          Tick<kDbg> t2{0, stats != nullptr, tk.gate};
          t2.start();
          tmem_ld_32x4(tq, v0);
          tmem_ld_32x4(tq + 16, v1);
          tmem_ld_32x4(tq + 32, v2);
          tmem_ld_wait();
          t2.stop(a_tmem);
          t2.start();
          hand_back();
          t2.stop(a_hb);
          // half 0: planes 0 and 1, half 1: plane 2.  The two planes of a half-0 warp are independent chains (shuffles,
          // exp, reciprocal) issued side by side; index arithmetic is done once per item.  The epilogue of this 3-channel
          // layer is a latency chain, not a throughput problem: a set spent ~1500 cycles per item in it.
          const int c0 = half == 0 ? 0 : 2;
          const uint32_t p0 = half == 0 ? v0[0] : v0[2], q0 = half == 0 ? v1[0] : v1[2], r0 = half == 0 ? v2[0] : v2[2];
          const float la = __shfl_up_sync(0xFFFFFFFFu, __uint_as_float(p0), 1), lb = __shfl_up_sync(0xFFFFFFFFu, __uint_as_float(v0[1]), 1);
          const float ra = __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(r0), 1), rb = __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(v2[1]), 1);
          const float za = (la + __uint_as_float(q0)) + ra + s_bias[c0];
          const float zb = (lb + __uint_as_float(v1[1])) + rb + s_bias[1];
          // sigmoid: ex2.approx + rcp.approx (relative error ~1e-6, far inside every output tolerance; the f16 / u8 forms
          // are derived from this same value, so the compact formats stay bit-consistent with the f32 output)
          const float ya = __fdividef(1.f, 1.f + __expf(-za));
          const float yb = __fdividef(1.f, 1.f + __expf(-zb));
          if (wvalid) {
            const size_t plane = static_cast<size_t>(S.oh) * S.ow;
            const size_t opx = static_cast<size_t>(static_cast<uint32_t>(wy) * static_cast<uint32_t>(S.ow) + static_cast<uint32_t>(wx));
            const size_t obase = static_cast<size_t>(n) * S.out_nstride + opx;
            const size_t pbase = static_cast<size_t>(n) * (S.out_nstride / 3) + opx;
            float* px = S.resid != nullptr ? static_cast<float*>(const_cast<void*>(S.resid)) + (static_cast<size_t>(n) * plane + opx) * 4 : nullptr;
            if (c0 < S.oc) {
              if (S.out2) S.out2[static_cast<size_t>(n) * 3 * plane + opx + c0 * plane] = za;
              if (S.out && !(dbg & 128)) store_network_output(S.out, c_mode, obase + c0 * plane, pbase, c0, ya);
              // pixel-interleaved second copy (tg_glue.cu: gather3).  The blue warp also writes the unused fourth component:
              // every byte of the copy is written, so no sector is ever partially dirty (a partial sector costs a
              // read-modify-write in ECC DRAM)
              if (px != nullptr && !(dbg & 384)) {
                if (c0 == 2) *reinterpret_cast<float2*>(px + 2) = make_float2(ya, 0.f);
                else px[0] = ya;
              }
            }
            if (half == 0 && 1 < S.oc) {
              if (S.out2) S.out2[static_cast<size_t>(n) * 3 * plane + opx + plane] = zb;
              if (S.out && !(dbg & 128)) store_network_output(S.out, c_mode, obase + plane, pbase, 1, yb);
              if (px != nullptr && !(dbg & 384)) px[1] = yb;
            }
          }
        }
      } else {
        // tall tile: GEMM row m = pixel (m / 8, m % 8); wide box with one view per tap: row = (tile row q, column lane)
        const bool wtap = c_wide == 2;
        const int iy = wtap ? ty * kWideH + q : ty * kTileH + pr, ix = wtap ? tx * kWideW + lane : tx * kTileW + pc;
        const bool valid = real && (iy < c_h) && (ix < c_w) && (!wtap || lane < kWideW);
        const int n_acc = (c_kind == kConv3x3) ? 1 : 4;
        const int sc = (c_kind == kConv3x3) ? 1 : 2;
        for (int a = 0; a < n_acc; ++a) {
          // accumulator a of a transposed conv sits at column kCtAccCol[a] = {0, 64, 192, 128} ([a0 | a1 | a3 | a2])
          const uint32_t taddr = tq + static_cast<uint32_t>(n_acc == 1 ? 0 : (a == 0 ? 0 : (a == 1 ? 64 : (a == 2 ? 192 : 128))));
          const int oy = iy * sc + (a >> 1), ox = ix * sc + (a & 1);
          if (c_mode == kOutNHWCbf16) {
            const size_t off = static_cast<size_t>(n) * c_img_bytes + (static_cast<uint32_t>(oy) * c_row_bytes + static_cast<uint32_t>(ox) * c_px_bytes);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              const uint32_t col = static_cast<uint32_t>(half * 32 + pass * 16);
              uint32_t v[16];
              tmem_ld_32x16(taddr + col, v);
              tmem_ld_wait();
              if (a == n_acc - 1 && pass == 1) hand_back();
              if (valid) {
                uint64_t a2[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) a2[e] = f2_pack(v[2 * e], v[2 * e + 1]);
                epi_store_bf16(a2, s_bias + col, c_out + off + pass * 32u, c_res ? c_res + off + pass * 32u : nullptr, c_relu != 0,
                               (kMask && c_msk) ? c_msk + off + pass * 32u : nullptr);
              }
            }
          } else {                                           // output conv on the tall geometry (TG_FRAME_WIDE=0)
            uint32_t v[16];
            tmem_ld_32x16(taddr, v);
            tmem_ld_wait();
            if (a == n_acc - 1) hand_back();
            if (valid && half == 0) {
              const size_t plane = static_cast<size_t>(S.oh) * S.ow;
              for (int c = 0; c < 3 && c < S.oc; ++c) {
                const size_t o = static_cast<size_t>(n) * S.out_nstride + static_cast<size_t>(oy) * S.ow + ox + c * plane;
                const float z = __uint_as_float(c == 0 ? v[0] : (c == 1 ? v[1] : v[2])) + s_bias[c];
                if (S.out2) S.out2[o] = z;
                static_cast<float*>(S.out)[o] = 1.f / (1.f + expf(-z));
              }
            }
          }
        }
      }
      tk.stop(a_body);
      // the network output has no consumer inside the kernel: nothing to publish
      if (c_mode == kOutNHWCbf16 && real && !(dbg & 2)) {
        tk.start();
        __syncwarp();                                        // orders the 32 lanes' stores before lane 0's release
        if (lane == 0) {
          mbar_wait(bar_pe, (pk & 1u) ^ 1u);                 // the publisher has released this set's previous tile
          mbar_arrive(bar_pf);
        }
        tk.stop(a_pub);
        ++pk;
      }
      gph ^= 1;
    }
    tall.stop(a_total);
    if (stats && warp == 2 && lane == 0) {
      unsigned long long* o = stats + blockIdx.x * 32;
      o[4] = a_cfull; o[5] = a_body; o[6] = a_pub; o[7] = a_total; o[8] = a_tmem; o[17] = a_hb;
    }
  } else if (warp < 2 + kEpiWarps + kPubWarps) {
    // ================================ publisher ============================================
    // A gpu-scope release (MEMBAR.GPU + RED) takes ~1300 cycles, more than an item's MMAs, so with two epilogue sets
    // two releases must be in flight: lanes 0 and 1 of this warp are two independent publishers (independent thread
    // scheduling lets one lane issue while the other sits in its fence), lane pw releasing the tiles of set pw.
    const int pw = lane;
    if (pw < 2) {
      int si = 0;
      uint32_t pk = 0;
      Tick<kDbg> tk{0, stats != nullptr, true};
      long long a_pfull = 0, a_red = 0;
      int seg_end = 0, seg_pub_end = 0;                      // publish items [.., seg_pub_end) of the current segment
      for (int it = blockIdx.x + pw * G; it < P.total_items; it += 2 * G) {
        if (it >= seg_end) {
          while (it >= segs[si].item_end) ++si;
          const FrSegS& C = segs[si];
          seg_end = C.item_end;
          seg_pub_end = C.out_mode == kOutNHWCbf16 ? C.item_begin + C.items_real : 0;
        }
        tk.gate = stat_seg < 0 || si == stat_seg;
        if (it < seg_pub_end && !(dbg & 2)) {
          const uint32_t pg = static_cast<uint32_t>(pw), pph = pk & 1u;   // set pw has its own barrier pair
          tk.start();
          mbar_wait(bar_pfull + 8 * pg, pph);                // all epilogue warps stored (acquire.cta)
          tk.stop(a_pfull);
          tk.start();
          red_release_gpu_add(P.flags + it, kItemDone);      // (flag_off == item_begin) cumulative gpu-scope release
          mbar_arrive(bar_pempty + 8 * pg);
          tk.stop(a_red);
          ++pk;
        }
      }
      if (stats && pw == 0) {
        unsigned long long* o = stats + blockIdx.x * 32;
        o[13] = a_pfull; o[14] = a_red;
      }
    }
  } else if (!(dbg & 1)) {
    // ================================ dependency warp ======================================
    // For every item whose input comes from a layer of this launch: wait until the producer tiles that the item's
    // halo box touches are published (one counter per tile, one lane per counter), two items per round trip:
    // relaxed gpu-scope loads of both items' counters in flight together, one fence.acq_rel.gpu for both in the
    // common already-published case (relaxed load that observed the released value + fence = acquire), then one
    // mbarrier arrive per item for the producer.
    asm volatile("griddepcontrol.wait;" ::: "memory");     // the counters are cleared by the previous kernel of the stream
    uint32_t dk = 0;
    int si_base = 0;
    for (int it0 = blockIdx.x; it0 < P.total_items; it0 += 32 * G) {
      uint32_t l0, l1, l2, l3;
      int l_si;
      decode_batch(P, it0, lane, G, si_base, l0, l1, l2, l3, l_si);
      const int left = (P.total_items - it0 + G - 1) / G;
      const int nb = left < 32 ? left : 32;
      // counter address of this lane for batch item j (nullptr: no counter for this lane); *has = item has dependencies
      auto dep_flag = [&](int j, bool* has) -> const uint32_t* {
        const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, l0, j), b2 = __shfl_sync(0xFFFFFFFFu, l2, j),
                       b3 = __shfl_sync(0xFFFFFFFFu, l3, j);
        *has = b3 != 0;
        if (b3 == 0) return nullptr;
        const FrSeg& S = P.segs[b0 & 0xFFu];
        const uint32_t nx = b3 & 15u, ny = (b3 >> 4) & 15u, inv = b3 >> 8;
        const uint32_t nxy = nx * ny;
        uint32_t q = static_cast<uint32_t>(lane);
        if (q >= nxy * static_cast<uint32_t>(S.dep_nseg)) return nullptr;
        const uint32_t ds = q >= nxy ? 1u : 0u;                     // dep_nseg <= 2 (host-checked)
        q -= ds * nxy;
        const uint32_t qy = (q * inv) >> 16, qx = q - qy * nx;
        return P.flags + P.segs[S.dep_seg0 + ds].flag_off +
               ((b0 >> 8) * S.dep_tiles_y + (b2 >> 16) + qy) * S.dep_tiles_x + (b2 & 0xFFFFu) + qx;
      };
      auto spin = [&](const uint32_t* f, int j) {
        const uint64_t t0 = global_ns();
        uint32_t spins = 0;
        while (ld_relaxed_gpu(f) < kItemDone) {
          __nanosleep(32);
          if ((++spins & 0x3FFu) == 0 && global_ns() - t0 > 4000000000ull) {
            printf("tg: frame dependency timeout item=%d block=%d\n", it0 + j * G, blockIdx.x);
            __trap();
          }
        }
      };
      for (int j = 0; j < nb; j += 2) {
        bool has0, has1 = false;
        const uint32_t* p0 = dep_flag(j, &has0);
        const uint32_t* p1 = (j + 1 < nb) ? dep_flag(j + 1, &has1) : nullptr;
        if (!has0 && !has1) continue;
        const uint32_t v0 = p0 ? ld_relaxed_gpu(p0) : kItemDone;
        const uint32_t v1 = p1 ? ld_relaxed_gpu(p1) : kItemDone;
        // Item j is handed over BEFORE item j+1 is waited for: j+1 may depend on j itself (a layer of exactly G tiles).
        if (v0 < kItemDone) spin(p0, j);
        if (!(dbg & 4)) fence_acq_rel_gpu();
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && v1 < kItemDone) {                  // rare: not yet published at the paired load
            spin(p1, j + 1);
            if (!(dbg & 4)) fence_acq_rel_gpu();
          }
          __syncwarp();
          if (!(h ? has1 : has0)) continue;
          const uint32_t ds = dk % kDepRing;
          mbar_wait(bar_dempty + 8 * ds, ((dk / kDepRing) & 1u) ^ 1u);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_dfull + 8 * ds);
          ++dk;
        }
      }
      si_base = __shfl_sync(0xFFFFFFFFu, l_si, nb - 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();                    // the leader's MMAs read the peer's shared memory and signal its barriers
  if (trace && threadIdx.x == 0) trace[static_cast<size_t>(P.nseg) * G + blockIdx.x] = global_ns();
  if (warp == 1) {
    if (kPair) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host
static unsigned long long* g_trace = nullptr;
static size_t g_trace_words = 0;
void frame_set_trace(unsigned long long* buf, size_t words) { g_trace = buf; g_trace_words = words; }

static int layer_nt(const FrLayer& l) { return l.cout_pad == 16 ? 16 : 64; }

// TG_FRAME_WIDE=0 keeps every layer on the tall 16x8 geometry (A/B measurements)
static bool use_wide(int kind) {
  static const bool on = []() { const char* e = getenv("TG_FRAME_WIDE"); return !(e && e[0] == '0'); }();
  return on && kind == kConv3x3;
}
// Pair mode (cta_group::2) needs every 3x3 conv on the wide geometry and co-resident 2-CTA clusters; TG_FRAME_PAIR=0
// selects the single-CTA kernel (A/B measurements).  Returns the number of clusters that can be resident (0: off).
static std::atomic<int> g_pair_override{-1};                     // tg_frame_set_pair: -1 = TG_FRAME_PAIR / default, 0 = off, 1 = on
void frame_set_pair(int on) { g_pair_override.store(on < 0 ? -1 : (on ? 1 : 0)); }
// Co-residency is what makes the static schedule deadlock-free (every CTA spins on tiles other CTAs produce), so the
// grid is sized from the occupancy the CURRENT device / context reports - an MPS active-thread limit or a green context
// shrinks it - cached per device ordinal.  What the occupancy API cannot see (an unrelated long-running kernel holding
// SMs) cannot hang the GPU either: every wait in the kernel is bounded and traps after 4 s.
struct FrDev { bool probed = false; int pair_clusters = 0; int single_ctas = 0; };
static FrDev g_frdev[TG_MAX_DEVICES];
static std::mutex g_frdev_mu;
static const FrDev& frame_device() {
  const int dev = tg_current_device();
  std::lock_guard<std::mutex> lk(g_frdev_mu);
  FrDev& d = g_frdev[dev];
  if (d.probed) return d;
  d.probed = true;
  bool ok = true;
  ok &= cudaFuncSetAttribute(frame_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrCfg<false>::kSmemBytes) == cudaSuccess;
  ok &= cudaFuncSetAttribute(frame_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrCfg<true>::kSmemBytes) == cudaSuccess;
  ok &= cudaFuncSetAttribute(frame_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrCfg<false>::kSmemBytes) == cudaSuccess;
  ok &= cudaFuncSetAttribute(frame_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrCfg<true>::kSmemBytes) == cudaSuccess;
  ok &= cudaFuncSetAttribute(frame_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrCfg<false>::kSmemBytes) == cudaSuccess;
  ok &= cudaFuncSetAttribute(frame_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FrCfg<true>::kSmemBytes) == cudaSuccess;
  if (!ok) { cudaGetLastError(); return d; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_kernel<false, false>, kFrThreads, FrCfg<false>::kSmemBytes) == cudaSuccess)
    d.single_ctas = (per_sm > 0 ? 1 : 0) * tg_num_sms();             // one CTA per SM (each allocates all of TMEM)
  else cudaGetLastError();
  if (use_wide(kConv3x3)) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tg_num_sms() & ~1);
    cfg.blockDim = dim3(kFrThreads);
    cfg.dynamicSmemBytes = FrCfg<true>::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, frame_kernel<true, false>, &cfg) == cudaSuccess) {
      const int want = tg_num_sms() / 2;
      d.pair_clusters = nc < want ? nc : want;
    } else cudaGetLastError();
  }
  return d;
}
static int pair_clusters_available() { return frame_device().pair_clusters; }
static int pair_clusters() {
  static const bool env_off = []() { const char* e = getenv("TG_FRAME_PAIR"); return e && e[0] == '0'; }();
  const int ov = g_pair_override.load();
  if (ov == 0 || (ov < 0 && env_off)) return 0;
  return pair_clusters_available();
}

size_t frame_tiles(int kind, int h, int w) {
  return use_wide(kind) ? static_cast<size_t>(tg_div_up(w, kWideW)) * tg_div_up(h, kWideH)
                        : static_cast<size_t>(tg_div_up(w, kTileW)) * tg_div_up(h, kTileH);
}
size_t frame_tiles_max(int h, int w) {
  const size_t a = static_cast<size_t>(tg_div_up(w, kWideW)) * tg_div_up(h, kWideH);
  const size_t b = static_cast<size_t>(tg_div_up(w, kTileW)) * tg_div_up(h, kTileH);
  return a > b ? a : b;
}

size_t frame_flag_count(const FrLayer* layers, int nlayers, int n) {
  size_t total = 0;
  for (int i = 0; i < nlayers; ++i)
    total += (static_cast<size_t>(n) * frame_tiles(layers[i].kind, layers[i].h, layers[i].w) + 1) *   // + 1: pair padding
             (layers[i].cout_pad / layer_nt(layers[i]));
  return total;
}

int launch_frame(const FrLayer* layers, int nlayers, const void* packed, size_t packed_bytes, int n,
                 uint32_t* flags, size_t flag_capacity, bool flags_zeroed, cudaStream_t stream) {
  TG_CHECK_ARG(nlayers >= 1 && nlayers + 1 <= kFrMapW32, "frame: too many layers (%d)", nlayers);
  static_assert(kFrMaxSegs <= 256, "segment index is packed into 8 bits");
  TG_CHECK_ARG(packed && flags && (packed_bytes % 128) == 0, "frame: bad packed blob");
  static thread_local FrProgram P;   // 13 KB: keep it off the stack
  memset(&P, 0, sizeof(P));
  const int clusters = pair_clusters();
  const bool pair = clusters > 0;
  P.pair = pair ? 1 : 0;
  {
    cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(packed_bytes / 128)};
    cuuint64_t strides[1] = {128};
    const int rows[3] = {kWBoxRows, 32, 24};
    const int slot[3] = {0, kFrMapW32, kFrMapW24};
    for (int i = 0; i < (pair ? 3 : 1); ++i) {
      cuuint32_t box[2] = {64, static_cast<cuuint32_t>(rows[i])};
      int rc = encode_bf16(&P.maps[slot[i]], packed, 2, dims, strides, box);
      if (rc) return rc;
    }
  }
  int nseg = 0, items = 0;
  bool any_mask = false;
  int first_seg[kFrMaxMaps], nchunks[kFrMaxMaps];
  double flops = 0.0;
  for (int li = 0; li < nlayers; ++li) {
    const FrLayer& l = layers[li];
    TG_CHECK_ARG(l.cin_pad == 64 || l.cin_pad == 128, "frame: cin_pad must be 64 or 128");
    TG_CHECK_ARG(l.cout_pad == 16 || l.cout_pad == 64 || l.cout_pad == 128, "frame: cout_pad must be 16/64/128");
    TG_CHECK_ARG((l.blob_off % 128) == 0, "frame: packed blob offsets must be 128-byte aligned");
    TG_CHECK_ARG(!(l.relu && l.resid), "frame: layer %d has both a ReLU and a residual", li);
    TG_CHECK_ARG(!l.mask || (l.out_mode == kOutNHWCbf16 && !l.relu && !l.out2), "frame: layer %d: a ReLU mask goes with a plain bf16 output", li);
    any_mask |= l.mask != nullptr;
    TG_CHECK_ARG(l.w < 32768 && l.h < 32768 && n < (1 << 23), "frame: layer size out of range");
    TG_CHECK_ARG(4.0 * l.h * l.w * l.cout_pad * 2 < 4294967296.0, "frame: one image of layer %d exceeds 4 GB", li);
    {
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(l.cin_pad), static_cast<cuuint64_t>(l.w),
                            static_cast<cuuint64_t>(l.h), static_cast<cuuint64_t>(n)};
      cuuint64_t strides[3] = {static_cast<cuuint64_t>(l.cin_pad) * 2, static_cast<cuuint64_t>(l.w) * l.cin_pad * 2,
                               static_cast<cuuint64_t>(l.h) * l.w * l.cin_pad * 2};
      cuuint32_t box[4] = {64, kTileW + 2, kTileH + 2, 1};
      if (use_wide(l.kind)) { box[1] = kWideBoxW; box[2] = kWideBoxH; }
      int rc = encode_bf16(&P.maps[1 + li], l.in, 4, dims, strides, box);
      if (rc) return rc;
    }
    const int nt = layer_nt(l);
    const int chunks = l.cout_pad / nt;
    const int kchunks = l.cin_pad / 64;
    const bool wide = use_wide(l.kind);
    const int tile_w = wide ? kWideW : kTileW, tile_h = wide ? kWideH : kTileH;
    const int tiles_x = tg_div_up(l.w, tile_w), tiles_y = tg_div_up(l.h, tile_h);
    const int sc = (l.kind == kConv3x3) ? 1 : 2;
    first_seg[li] = nseg;
    nchunks[li] = chunks;
    flops += 2.0 * 9.0 * l.cin_pad * (nt == 64 ? l.cout_pad : 3) * n * l.h * l.w;
    for (int c = 0; c < chunks; ++c) {
      TG_CHECK_ARG(nseg < kFrMaxSegs, "frame: too many segments");
      FrSeg& S = P.segs[nseg];
      S.item_begin = items;
      S.items_real = n * tiles_x * tiles_y;
      items += pair ? ((S.items_real + 1) & ~1) : S.items_real;
      S.item_end = items;
      S.tiles_x = tiles_x; S.tiles_y = tiles_y; S.h = l.h; S.w = l.w;
      // wide tiles: N-stacked filter rows (1, the default), or - TG_FRAME_TAP=1, pair mode only, where an N=64 MMA costs 43
      // cycles instead of 75 - one MMA group per tap on shifted views of the wide box (2): a third of the accumulator
      // columns and no shuffles, but the A tile is re-read nine times and shared-memory bandwidth makes it slower
      static const bool tap_views = []() { const char* e = getenv("TG_FRAME_TAP"); return e && e[0] == '1'; }();   // off: measured slower
      S.wide = wide ? ((pair && tap_views && nt == 64) ? 2 : 1) : 0; S.tile_w = tile_w; S.tile_h = tile_h;
      S.box_w = wide ? kWideBoxW : kTileW + 2; S.box_h = wide ? kWideBoxH : kTileH + 2;
      S.fd_tiles_x = make_fastdiv(tiles_x); S.fd_tiles_y = make_fastdiv(tiles_y);
      S.map_a = 1 + li;
      S.kchunks = kchunks; S.kind = l.kind; S.nt = nt;
      S.w_rows = 9u * nt;
      if (S.wide == 1) {                                              // one MMA N group per filter row: 3 taps x nt rows
        S.w_groups = 3;
        for (int gi = 0; gi < 3; ++gi) S.w_grp_half[gi] = 3 * nt / 2;
        S.w_box_rows = (3 * nt / 2 == 96) ? kWBoxRows : 3 * nt / 2;   // 96 = 2 x 48-row boxes; the output conv: one 24-row box
      } else if (l.kind == kConvT3x3s2) {                             // four shift groups of N = 256 / 128 / 128 / 64 (nt == 64)
        TG_CHECK_ARG(nt == 64, "frame: the transposed conv needs 64-wide output chunks");
        S.w_groups = 4;
        S.w_grp_half[0] = 128; S.w_grp_half[1] = 64; S.w_grp_half[2] = 64; S.w_grp_half[3] = 32;
        S.w_box_rows = 32;
      } else {                                                        // one group per tap (measurement variants)
        TG_CHECK_ARG(!pair || nt == 64, "frame: per-tap pair mode needs 64-wide chunks");
        S.w_groups = 9;
        for (int gi = 0; gi < 9; ++gi) S.w_grp_half[gi] = nt / 2;
        S.w_box_rows = nt / 2;
      }
      S.w_map = S.w_box_rows == kWBoxRows ? 0 : (S.w_box_rows == 32 ? kFrMapW32 : kFrMapW24);
      TG_CHECK_ARG(!pair || S.w_box_rows == kWBoxRows || S.w_box_rows == 32 || S.w_box_rows == 24,
                   "frame: no pair-mode weight box for layer %d (%d-row boxes)", li, S.w_box_rows);
      for (int kc = 0; kc < kchunks; ++kc)
        S.w_row0[kc] = static_cast<uint32_t>(l.blob_off / 128) + static_cast<uint32_t>(c * kchunks + kc) * S.w_rows;
      S.out_mode = l.out_mode; S.relu = l.relu;
      S.oh = l.h * sc; S.ow = l.w * sc;
      const bool net_out = tc_out_is_network_output(l.out_mode);
      S.oc = net_out ? 3 : l.cout_pad;
      S.ch0 = c * 64;
      S.out_nstride = l.out_nstride > 0 ? l.out_nstride : static_cast<long long>(S.oc) * S.oh * S.ow;
      S.out = l.out; S.out2 = l.mask ? static_cast<float*>(const_cast<void*>(l.mask)) : l.out2;
      S.resid = net_out ? l.out_rgbx : l.resid;
      TG_CHECK_ARG(l.out_rgbx == nullptr || (net_out && wide && l.cout_pad == 16),
                   "frame: the interleaved copy exists for the wide output conv only");
      TG_CHECK_ARG(l.out_mode == kOutNHWCbf16 || net_out, "frame: bad output mode %d", l.out_mode);
      TG_CHECK_ARG(!(net_out && l.out_mode != kOutNCHWf32Sigmoid) || (wide && l.out2 == nullptr && S.out_nstride % 3 == 0),
                   "frame: the fp16 / uint8 network outputs need the wide geometry and take no logits");
      S.bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed) + l.blob_off +
                                              packed_weight_bytes(l.cin_pad, l.cout_pad)) + c * 64;
      if (li == 0) {
        S.dep_seg0 = 0; S.dep_nseg = 0;
      } else {
        const FrLayer& pl = layers[li - 1];
        const int psc = (pl.kind == kConv3x3) ? 1 : 2;
        TG_CHECK_ARG(pl.h * psc == l.h && pl.w * psc == l.w && pl.out == l.in, "frame: layer %d does not consume layer %d", li, li - 1);
        S.dep_seg0 = first_seg[li - 1]; S.dep_nseg = nchunks[li - 1];
        const bool pwide = use_wide(pl.kind);
        const int ptw = pwide ? kWideW : kTileW, pth = pwide ? kWideH : kTileH;
        S.dep_tiles_x = tg_div_up(pl.w, ptw); S.dep_tiles_y = tg_div_up(pl.h, pth);
        S.dep_tw = ptw * psc; S.dep_th = pth * psc;   // producer tile footprint in this layer's input pixels
        S.fd_dep_tw = make_fastdiv(S.dep_tw); S.fd_dep_th = make_fastdiv(S.dep_th);
        // one poll lane per producer tile touched by the halo box
        const int mx = (S.box_w + S.dep_tw - 2) / S.dep_tw + 1, my = (S.box_h + S.dep_th - 2) / S.dep_th + 1;
        TG_CHECK_ARG(S.dep_nseg <= 2 && mx <= 15 && my <= 15 && mx * my * S.dep_nseg <= 32,
                     "frame: layer %d waits on too many producer tiles (%d x %d x %d)", li, mx, my, S.dep_nseg);
      }
      S.flag_off = static_cast<uint32_t>(S.item_begin);
      ++nseg;
    }
  }
  P.nseg = nseg;
  P.total_items = items;
  P.flags = flags;
  {
    static const int dbg = []() { const char* e = getenv("TG_FRAME_DBG"); return e ? atoi(e) : 0; }();
    P.dbg = dbg;
    static const int stat_seg = []() { const char* e = getenv("TG_FRAME_STAT_SEG"); return e ? atoi(e) : -1; }();
    P.stat_seg = stat_seg;
  }
  // one CTA per SM (pair mode: one 2-CTA cluster per TPC); items are dealt round-robin, so any grid size works
  const int max_ctas = pair ? 2 * clusters : frame_device().single_ctas;
  TG_CHECK_ARG(max_ctas >= 1, "frame: the kernel cannot be resident on this device / context (occupancy 0)");
  const int grid = items < max_ctas ? items : max_ctas;           // pair mode: items and max_ctas are even
  P.trace = (g_trace && g_trace_words >= static_cast<size_t>(nseg + 1) * grid) ? g_trace : nullptr;
  // optional stall accounting behind the trace: 16 words per CTA (see Tick)
  P.stats = (P.trace && g_trace_words >= static_cast<size_t>(nseg + 1 + 32) * grid) ? g_trace + static_cast<size_t>(nseg + 1) * grid : nullptr;

  TG_CHECK_ARG(static_cast<size_t>(items) <= flag_capacity, "frame: %d items exceed the flag capacity %zu", items, flag_capacity);
  if (!flags_zeroed) TG_CUDA(cudaMemsetAsync(flags, 0, static_cast<size_t>(items) * sizeof(uint32_t), stream));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kFrThreads);
  cfg.dynamicSmemBytes = pair ? FrCfg<true>::kSmemBytes : FrCfg<false>::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pair ? 2 : 1;
  tg_prof_pre(TG_K_FRAME, flops, stream);
  const bool measure = P.dbg != 0 || P.trace != nullptr;
  if (any_mask) {                                                  // data-gradient chain (no measurement build of it)
    if (pair) TG_CUDA(cudaLaunchKernelEx(&cfg, frame_kernel<true, false, true>, P));
    else TG_CUDA(cudaLaunchKernelEx(&cfg, frame_kernel<false, false, true>, P));
  } else if (pair) {
    if (measure) TG_CUDA(cudaLaunchKernelEx(&cfg, frame_kernel<true, true>, P));
    else TG_CUDA(cudaLaunchKernelEx(&cfg, frame_kernel<true, false>, P));
  } else {
    if (measure) TG_CUDA(cudaLaunchKernelEx(&cfg, frame_kernel<false, true>, P));
    else TG_CUDA(cudaLaunchKernelEx(&cfg, frame_kernel<false, false>, P));
  }
  tg_prof_post(stream);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

}  // namespace tg
