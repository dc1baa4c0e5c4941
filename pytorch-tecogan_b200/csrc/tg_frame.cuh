// Persistent whole-generator kernel ("frame kernel"): all 41 convolutions of one generator forward
// (reference code/models.py:78-86) as ONE launch.  See tg_frame.cu for the design.
#pragma once
#include "tg_conv_tc.cuh"

namespace tg {

constexpr int kFrMaxSegs = 56;     // 41 layers + one extra segment per 128-wide layer
constexpr int kFrMaxMaps = 46;     // [0] packed weights (48-row box), [1 + layer] input activation of the layer,
constexpr int kFrMapW32 = 44;      // [44] / [45] packed weights with 32- / 24-row boxes (pair mode: half of a tap group
constexpr int kFrMapW24 = 45;      //   of the transposed convs / of the output conv per CTA)

// Tile geometries.  "tall": 16 rows x 8 columns, halo box {64ch, 10, 18}, one tcgen05.mma group per filter tap
// (N = nt); used by the transposed convs.  "wide" (3x3 convs): 4 rows x 32 columns of which the inner 30 are
// outputs, box {64ch, 32, 6}; the three dx taps of a filter row share ONE A view and are ONE MMA of N = 3*nt
// (weights of the row stacked along N), and the epilogue adds the three partial sums of neighbouring pixels
// (out[x] = P0[x-1] + P1[x] + P2[x+1], a lane shuffle).  A tcgen05.mma costs ~43 + N/2 cycles with both operands
// in shared memory (profiles/r01_mma_microbench.txt), so 12 MMAs of N=192 per 120 pixels replace 36 of N=64 per 128.
constexpr int kWideW = 30, kWideH = 4, kWideBoxW = 32, kWideBoxH = 6;

// Exact unsigned division by a launch-time constant (Granlund-Montgomery round-up): q = (umulhi(x, m) + x) >> s for
// x < 2^31.  The item decode of every warp role sits on its per-item critical path; a hardware-free integer division
// is ~100 cycles of dependent instructions, this is ~10.
struct FastDiv { uint32_t m, s; };
inline FastDiv make_fastdiv(uint32_t d) {
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  FastDiv f;
  f.m = static_cast<uint32_t>(((1ull << 32) * ((1ull << s) - d)) / d + 1);
  f.s = s;
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fdiv(uint32_t x, FastDiv f) { return (__umulhi(x, f.m) + x) >> f.s; }
#endif

// One segment = one (layer, 64-wide output-channel chunk) pass over all tiles of the layer.
struct FrSeg {
  int item_begin, item_end;   // global item range [begin, end); pair mode pads every segment to an even count
  int items_real;             // items of the segment that exist (local index >= items_real: the pair's padding item)
  int tiles_x, tiles_y;       // tiles per image in the layer's input resolution
  int h, w;                   // input resolution
  int wide;                   // tile geometry (see above)
  int tile_w, tile_h;         // output pixels per tile (input resolution)
  int box_w, box_h;           // staged halo box in pixels
  int map_a;                  // index of the input-activation tensor map
  int kchunks;                // input channels / 64
  int kind;                   // kConv3x3 / kConvT3x3s2
  int nt;                     // UMMA N: 64, or 16 for the 3-channel output conv
  uint32_t w_row0[2];         // first 128-byte row of the weight block of K chunk 0/1 in the packed blob
  uint32_t w_rows;            // rows per weight block (9 taps * nt)
  // pair mode: the block is w_groups MMA-N groups; group g has 2 * w_grp_half[g] rows and each CTA loads its half of it
  // (rows [rank * half, (rank + 1) * half) of the group) in boxes of w_box_rows rows through tensor map w_map
  int w_groups, w_grp_half[9], w_box_rows, w_map;
  // epilogue
  int out_mode, relu, oh, ow, oc, ch0;   // ch0 = first output channel of this chunk
  long long out_nstride;
  void* out;
  float* out2;
  const void* resid;
  const float* bias;          // already offset to ch0
  // tile-level dependencies: the producer layer's segments
  int dep_seg0, dep_nseg;     // first producer segment, count (0 = input comes from a previous kernel)
  int dep_tiles_x, dep_tiles_y, dep_tw, dep_th;   // producer tile = (y / dep_th, x / dep_tw) in this layer's input pixels
  FastDiv fd_tiles_x, fd_tiles_y, fd_dep_tw, fd_dep_th;
  uint32_t flag_off;          // offset of this segment's per-item completion counters
};

// Shared-memory copy of the fields the per-item loops of the MMA, epilogue and publisher warps read.  FrProgram is
// 17 KB of kernel parameters; indexing it dynamically per item costs a chain of constant-cache misses (measured:
// 700-800 cycles of every ~2400-cycle item), so every CTA compacts it once at start.
struct FrSegS {
  int item_begin, item_end, items_real;
  uint16_t tiles_x, tiles_y, h, w, oh, ow, oc, ch0;
  uint8_t wide, out_mode, relu, kind, nt, kchunks, pad0, pad1;
  FastDiv fd_tiles_x, fd_tiles_y;
  long long out_nstride;
  void* out;
  float* out2;                // network-output layers: optional f32 logits; bf16 layers: the optional ReLU mask (FrLayer::mask)
  const void* resid;          // output conv: the optional interleaved float4 copy (FrLayer::out_rgbx) instead
  const float* bias;
};
static_assert(sizeof(FrSegS) == 96, "FrSegS layout");

struct FrProgram {
  CUtensorMap maps[kFrMaxMaps];
  FrSeg segs[kFrMaxSegs];
  int nseg, total_items;
  uint32_t* flags;            // zeroed before the launch; one counter per item (128 = complete)
  unsigned long long* trace;  // optional [nseg+1][grid] globaltimer stamps (tg_frame_set_trace), else null
  unsigned long long* stats;  // optional [grid][16] cycles per wait of every role (behind the trace buffer), else null
  int dbg;                    // measurement-only knobs (TG_FRAME_DBG): 1 no dependency wait, 2 no publish, 4 no acquire fence,
                              //   8 issuer ignores accumulator-drained, 16 no bf16 epilogue stores, 32 cta-scope waits in the pair issuer,
                              //   64 no proxy fence before the A loads, 128 no network-output stores, 256 no interleaved copy,
                              //   512 / 1024: no MMAs / no A loads in segment stat_seg
  int pair;                   // 1: launched as CTA pairs (cta_group::2)
  int stat_seg;               // stall accounting restricted to this segment (TG_FRAME_STAT_SEG, -1 = all)
};

struct FrLayer {              // host-side description of one conv layer of the frame
  int kind, cin_pad, cout_pad, out_mode, relu, h, w;
  const void* in;
  void* out;
  const void* resid;
  float* out2;
  size_t blob_off;            // byte offset of the layer's packed blob
  long long out_nstride;
  void* out_rgbx;             // output conv only (optional): second copy of the result as float4 {R,G,B,0} per pixel
  const void* mask;           // bf16 layers only (optional): saved post-ReLU activation with the output's layout; the result is
                              //   zeroed where it is zero (ReLU backward, applied after the residual) - tg_gen_backward
};

size_t frame_flag_count(const FrLayer* layers, int nlayers, int n);
// tiles of one image of an h x w layer input in the geometry launch_frame picks for `kind`
size_t frame_tiles(int kind, int h, int w);
// upper bound over both geometries (workspace sizing)
size_t frame_tiles_max(int h, int w);
// Builds the program and launches the frame kernel.  `flags` holds `flag_capacity` uint32 counters
// (>= frame_flag_count()); flags_zeroed = an earlier kernel of the stream already cleared them.
int launch_frame(const FrLayer* layers, int nlayers, const void* packed, size_t packed_bytes, int n,
                 uint32_t* flags, size_t flag_capacity, bool flags_zeroed, cudaStream_t stream);

// test / measurement hook: 1 = CTA pairs (default when 2-CTA clusters can be co-resident), 0 = single-CTA kernel,
// -1 = back to the default (TG_FRAME_PAIR environment variable)
void frame_set_pair(int on);
// measurement hook: subsequent frame launches stamp [nseg+1][grid] globaltimer values into buf (null = off)
void frame_set_trace(unsigned long long* buf, size_t words);

// tg_glue.cu: the fused frame-input producer, optionally clearing `zero_count` uint32 at `zero`.
// prev_rgbx: optional pixel-interleaved float4 copy of prev_hr [n][4h][4w] (gathered instead of the planar tensor).
int fused_input_launch(const float* lr_t, const float* lr_prev, const float* prev_hr, void* x_nhwc, int n, int h,
                       int w, long long lr_bs, long long hr_bs, uint32_t* zero, size_t zero_count,
                       cudaStream_t stream, const void* prev_rgbx = nullptr);

// tg_glue.cu: dz = dout * out * (1 - out) as NHWC bf16 [n,hw,64] (backward of the final sigmoid)
int sigmoid_bwd_pack_launch(const float* dout, const float* out, void* dz, int n, long long hw, long long dout_nstride,
                            long long out_nstride, cudaStream_t st);

}  // namespace tg
