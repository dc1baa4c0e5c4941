// Internal interface of the tcgen05 implicit-GEMM convolution kernel (tg_conv_tc.cu).
#pragma once
#include "tg_common.cuh"

namespace tg {

constexpr int kTileH = 16;        // output sub-tile: 16 rows x 8 columns = 128 GEMM rows (UMMA M)
constexpr int kTileW = 8;
constexpr int kAccCols = 64;      // TMEM columns reserved per accumulator
constexpr int kMaxTaps = 16;
constexpr int kMaxAcc = 4;

enum TcKind { kConv3x3 = 0, kConvT3x3s2 = 1, kConv4x4s2 = 2,
              kConvT4x4s2Phase = 6 };   // one output-parity phase of the adjoint of Conv2d(k4,s2,p1): 2x2 taps, stride-2 stores
enum TcOut { kOutNHWCbf16 = 0, kOutNCHWf32Sigmoid = 1, kOutNCHWf32Raw = 2, kOutNHWCf32 = 3,   // 3: pre-BatchNorm conv outputs
             // frame kernel only, compact network outputs for the D2H side of the clip pipeline (TG_OUT_F16 / TG_OUT_U8):
             kOutNCHWf16Sigmoid = 4,      // sigmoid, rounded to fp16, planar [n,3,h,w] (what the reference's autocast path emits)
             kOutNHWCu8Sigmoid = 5 };     // sigmoid, (y * 255) truncated to uint8, interleaved [n,h,w,3] (code/ops.py:234-237)
inline bool tc_out_is_network_output(int m) { return m == kOutNCHWf32Sigmoid || m == kOutNCHWf16Sigmoid || m == kOutNHWCu8Sigmoid; }
// tg_pack_weights kinds beyond TcKind: data-gradient convolutions derived from a forward layer's weights
enum TcPackKind { kPackConv3x3Dgrad = 3, kPackConvT3x3s2Dgrad = 4, kPackConv4x4s2Dgrad = 5 };
enum TcMask { kMaskNone = 0, kMaskRelu = 1, kMaskLrelu02 = 2 };   // backward of the activation that produced `mask`
enum TcAct { kActNone = 0, kActRelu = 1, kActLrelu02 = 2 };   // LeakyReLU(0.2): code/ops.py:71-72

// ConvTranspose2d(k3, s2, p1, op1) as four output phases (oy, ox) = (2y + py, 2x + px), accumulator a = 2*py + px.
// The nine (phase, tap) products are ordered BY INPUT SHIFT (dy, dx) - the A view they read - so that the products of
// one shift are consecutive weight blocks and can be issued as ONE N-stacked MMA into adjacent accumulators:
//   j : 0    1    2    3  |  4    5  |  6    7  |  8
//   k : k11  k12  k22  k21 | k10  k20 | k02  k01 | k00        (ky, kx of the ConvTranspose weight)
//   a : 0    1    3    2  |  1    3  |  3    2  |  3          shift (0,0) | (0,1) | (1,0) | (1,1)
// With the accumulators laid out [a0 | a1 | a3 | a2] the four shifts are MMAs of N = 256 / 128 / 128 / 64 onto column
// offsets 0 / 64 / 128 / 128 (tg_frame.cu); the per-layer kernel issues the nine products one by one.  Per accumulator
// the products are added in the same order in both kernels (bit-identical results).
constexpr int kCtKy[9] = {1, 1, 2, 2, 1, 2, 0, 0, 0};
constexpr int kCtKx[9] = {1, 2, 2, 1, 0, 0, 2, 1, 0};
constexpr int kCtDy[9] = {0, 0, 0, 0, 0, 0, 1, 1, 1};
constexpr int kCtDx[9] = {0, 0, 0, 0, 1, 1, 0, 0, 1};
constexpr int kCtAcc[9] = {0, 1, 3, 2, 1, 3, 3, 2, 3};
constexpr int kCtFirst[9] = {1, 1, 1, 1, 0, 0, 0, 0, 0};
constexpr int kCtAccCol[4] = {0, 64, 192, 128};        // frame kernel: TMEM column of accumulator a ([a0 | a1 | a3 | a2])

// One MMA group = one filter tap on one 64-channel K chunk: 4 x tcgen05.mma (K=16 each).
struct TcTap {
  uint32_t a_off;   // byte offset of the (shifted) A view inside a stage
  uint16_t acc;     // accumulator (output phase) it adds into
  uint16_t first;   // 1 = first tap of this accumulator (clears it on K chunk 0)
};

struct TcParams {
  // tile space (the resolution the 16x8 sub-tiles tile: the conv input resolution)
  int n, h, w, tiles_x, tiles_y, num_items;
  // K loop: an item consumes stages_per_item pipeline stages = kchunks x nphase
  int kchunks;           // input channels / 64
  int nphase;            // 1, or 4 input-parity phases of a stride-2 conv (one strided TMA box each)
  int stages_per_item;
  int in_scale;          // input coordinate = tile coordinate * in_scale (2 for the stride-2 conv)
  int ntaps;             // MMA groups per stage
  TcTap taps[kMaxTaps];
  int n_acc;             // accumulators per item (1 conv, 4 transposed-conv phases)
  // A staging
  int ncopies;           // TMA boxes per stage (1 HALO, 3 DX3)
  int copy_dx[3];        // x origin of each copy relative to tile x0
  int box_y0;            // y origin relative to tile y0
  uint32_t copy_bytes;   // bytes per copy (box bytes)
  uint32_t stage_bytes;  // ncopies * copy_bytes (mbarrier expect_tx)
  uint32_t a_region;     // stage_bytes rounded up to 1024: offset of the streamed weight block in a stage
  uint32_t w_stage_bytes;// 0 = weights resident for the CTA's lifetime; else bytes of weights streamed per stage
  uint32_t stage_stride; // smem distance between stages (1024-aligned)
  uint32_t sbo;          // UMMA stride-byte-offset between 8-row groups of A
  int nstages, ngroups;  // A ring depth, accumulator ring depth
  uint32_t w_bytes;      // resident weight bytes for this CTA's output-channel chunk
  // epilogue
  int out_mode, oh, ow, oc;   // output tensor dims (NHWC: channels oc; NCHW: oc planes)
  int sy, sx;                 // output pixel = input pixel * (sy,sx) + acc offset
  int acc_oy[kMaxAcc], acc_ox[kMaxAcc];
  int relu;                   // TcAct
  long long out_nstride;      // NCHW output: elements between consecutive images
  void* out;
  float* out2;                // optional pre-sigmoid logits (NCHW f32)
  const void* resid;          // optional residual, same layout as out (NHWC bf16)
  const void* mask;           // optional saved activation, same layout as out: ReLU backward zeroes the result where
  int mask_mode;              // mask == 0; LeakyReLU(0.2) backward scales it by 0.2 where mask <= 0  (TcMask)
  const float* bias;          // padded bias for the whole layer (chunk offset added in-kernel)
};

// Launches the kernel for one layer.  x: NHWC bf16 with `cin_pad` channels.
int launch_conv_tc(int kind, int out_mode, const void* x, const void* packed_w, const float* bias,
                   const void* resid, void* out, float* out2, int n, int h, int w, int cin_pad,
                   int cout_pad, int relu, int amode, long long out_nstride, cudaStream_t stream,
                   const void* mask = nullptr, int mask_mode = kMaskRelu, int phase = 0);

// SWIZZLE_128B bf16 tiled tensor map (cuTensorMapEncodeTiled through the runtime's driver entry point)
int encode_bf16(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* elem_strides = nullptr);

// tg_wgrad.cu: dW[cout][cin][3][3] += sum_pixels dY (x) X (f32 atomics); x NHWC bf16 [n,h,w,cin_pad], dy [n,h,w,cout_pad]
// db != null: also the bias gradient db[cout] += sum over pixels of dY (fused into the weight-gradient kernel for <= 64
// output channels, a separate reduction launch otherwise)
int launch_wgrad3x3(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout, int cin_pad,
                    int cout_pad, cudaStream_t stream, float* db = nullptr);
// nlayers consecutive same-shape layers (cout <= 64) in one launch: X / dY blocks of n images each back to back, gradients
// dw_stride / db_stride floats apart (db may be null)
int launch_wgrad3x3_batched(const void* x, const void* dy, float* dw, float* db, int nlayers, long long dw_stride, long long db_stride,
                            int n, int h, int w, int cin, int cout, int cin_pad, cudaStream_t stream);
// ConvTranspose2d(k3,s2,p1,op1): x [n,h,w,cin_pad], dy [n,2h,2w,cout_pad] -> dw [cin][cout][3][3]
int launch_wgrad_convT3x3s2(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                            int cin_pad, int cout_pad, cudaStream_t stream);
// Conv2d(k4,s2,p1): x [n,2h,2w,cin_pad], dy [n,h,w,cout_pad] -> dw [cout][cin][4][4]
int launch_wgrad_conv4x4s2(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                           int cin_pad, int cout_pad, cudaStream_t stream);
int launch_bias_grad(const void* dy, long long pixels, int cpad, int c, float* db, cudaStream_t stream);

// Batched weight packing (tg_api.cu): all layers of a network in one launch.
constexpr int kPackBatch = 64;
struct PackJob {
  const float* w; const float* bias; void* dst; float* bias_dst;
  int kind, cin, cout, cin_pad, cout_pad, nt;
};
struct PackJobs { PackJob j[kPackBatch]; };
struct PackJobSpec { int kind; const float* weight; const float* bias; int cin, cout; void* packed; };   // tg_pack_weights arguments
int pack_weights_batched(const PackJobSpec* specs, int n, cudaStream_t stream);

// packed layout helpers
size_t packed_weight_bytes(int cin_pad, int cout_pad);   // bf16 blocks only, 3x3 kernels
size_t packed_weight_bytes_k(int kind, int cin_pad, int cout_pad);   // ... 16 taps for kConv4x4s2
int cin_padded(int cin);
int cout_padded(int cout);

}  // namespace tg
