/* libtecogan_b200 — C ABI of the B200-native TecoGAN hot path.
 *
 * Plain C, raw device/host pointers and sizes, no torch types.  Every entry point
 *   - returns 0 on success or a negative TG_ERR_* code (never throws, never exits);
 *     tg_last_error_string() then describes the failure;
 *   - is asynchronous with respect to the host: work is enqueued on `stream` (a cudaStream_t
 *     passed as void*), performs no allocation and no synchronisation, and is therefore
 *     CUDA-graph capturable (exception: the *_host entry points, which stage host buffers);
 *   - borrows the pointers for the duration of the enqueued work only; the caller (PyTorch's
 *     caching allocator in the Python mirror) owns all memory.
 *
 * Each function names the reference call site(s) it replaces (paths relative to the
 * reference repo dwight-foster/Pytorch-TecoGAN).  The reference is pure Python over PyTorch
 * and has no FFI of its own; INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Layout conventions
 *   "NCHW f32"  : contiguous float tensors exactly as the reference holds them.
 *   "NHWC bf16" : contiguous __nv_bfloat16 [N][H][W][C], C in {64,128}; the internal activation
 *                 format of the generator (51 input channels are zero-padded to 64).
 */
#ifndef TECOGAN_B200_H
#define TECOGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_OK 0
#define TG_ERR_BAD_ARG (-1)   /* shape / alignment / null pointer */
#define TG_ERR_CUDA (-2)      /* a CUDA runtime or driver call failed */
#define TG_ERR_ARCH (-3)      /* device is not sm_100 */
#define TG_ERR_WORKSPACE (-4) /* workspace too small */

/* A-operand staging of the tensor-core conv kernels (DESIGN.md "A staging modes"). */
#define TG_AMODE_HALO 0 /* one TMA box with halo per stage; taps are row-shifted descriptors */
#define TG_AMODE_DX3 1  /* three x-shifted copies per stage; every descriptor 1024B-aligned */
/* tg_gen_forward / tg_gen_clip_forward only: all 41 layers as ONE persistent kernel chained by
 * per-tile completion counters (HALO staging inside); the default of the Python mirror. */
#define TG_AMODE_FRAME 2

const char* tg_last_error_string(void);
int tg_version(void);
/* 0 if the current device can run this library (compute capability 10.x), else TG_ERR_ARCH. */
int tg_check_device(void);

/* Measurement hooks (bench.py): number of kernels this library launched since load, and optional
 * per-launch CUDA-event timing.  Between tg_profile_begin() and tg_profile_end() every launch is
 * bracketed by events on its stream; tg_profile_end synchronises and returns up to max_entries
 * records (kernel id: 0 conv_tc<64>, 1 output conv, 2 fused frame input, 3 other glue, 4 pack, 5 frame kernel;
 * duration in ms; algorithmic work = FLOPs for convs, bytes for glue).  Not graph-capturable.  The launch counter is
 * thread-safe; the timing mode is a single-threaded measurement mode (while it is on, all launches must come from one
 * host thread) and is the only process-wide mutable state of the library besides the two frame-kernel test hooks below. */
long long tg_launch_count(void);
int tg_profile_begin(void);
int tg_profile_end(int max_entries, int* kernel_ids, float* ms, double* work);

/* Measurement hook for the frame kernel (TG_AMODE_FRAME): while buf != NULL every frame launch writes
 * %globaltimer stamps [segment][cta] (first item of each (layer, Cout chunk) segment per CTA, plus the
 * kernel end in row nseg) into the device buffer; scripts/frame_trace.py turns them into a per-layer
 * timeline.  buf == NULL switches it off. */
int tg_frame_set_trace(void* buf, size_t bytes);
/* Test / measurement hook: the frame kernel runs as CTA pairs (tcgen05.mma.cta_group::2, one 2-CTA cluster per TPC)
 * whenever such clusters can be co-resident, else as the single-CTA kernel.  on = 0 forces the single-CTA kernel,
 * 1 the default choice, -1 returns control to the TG_FRAME_PAIR environment variable.  Both produce identical bits. */
int tg_frame_set_pair(int on);

/* ------------------------------------------------------------------ glue (HBM-bound) ------- */

/* out[n, c*r*r + dy*r + dx, y, x] = in[n, c, r*y+dy, r*x+dx]; bit-exact, any 4-byte element.
 * Replaces the inline view/permute/reshape at main.py:207-212, code/train.py:102-106. */
int tg_space_to_depth(const void* in, void* out, int n, int c, int h_out, int w_out, int r,
                      void* stream);
/* inverse of tg_space_to_depth (== F.pixel_shuffle); in [n, c*r*r, h, w] -> out [n, c, h*r, w*r]. */
int tg_depth_to_space(const void* in, void* out, int n, int c, int h_in, int w_in, int r,
                      void* stream);
/* F.grid_sample(img, grid.half()) bilinear / zeros / align_corners=False.
 * img [n,c,h,w] f32, grid [n,ho,wo,2] f32 (rounded to fp16 inside, as the reference does),
 * out [n,c,ho,wo] f32.  Replaces main.py:203, code/train.py:81,98,165,187.
 * Three-plane inputs take a branch-free path: an out-of-range tap keeps a clamped (valid) address and weight 0, so the
 * result equals grid_sample's zero padding for FINITE images (frames are finite; an Inf / NaN border pixel would leak into
 * samples outside the image, where grid_sample returns 0). */
int tg_warp_bilinear(const float* img, const float* grid, float* out, int n, int c, int h, int w,
                     int ho, int wo, void* stream);
/* nn.Upsample(scale_factor=4, mode="bilinear") (align_corners=False); out = up4(in * pre_scale).
 * in [n,c,h,w] f32 -> out [n,c,4h,4w] f32.  Replaces code/ops.py:98-100 as used at main.py:186. */
int tg_upscale4_bilinear(const float* in, float* out, int n, int c, int h, int w, float pre_scale,
                         void* stream);
/* Fused producer of the generator input for frame t (main.py:186-213 in one pass):
 *   flow  = upscale_four(lr_prev*4)[:,0:2] re-viewed as [n,4h,4w,2]   (computed on the fly)
 *   x     = cat(lr_t, space_to_depth(deprocess(grid_sample(prev_hr, flow.half()))))
 * written as NHWC bf16 [n,h,w,64] (channels 51..63 zero).  lr_prev == NULL or prev_hr == NULL
 * selects the first-frame input cat(lr_t, zeros) (main.py:191-193).
 * lr_t, lr_prev: [n,3,h,w] f32 with batch stride lr_batch_stride elements;
 * prev_hr: [n,3,4h,4w] f32 with batch stride hr_batch_stride elements. */
int tg_fused_warp_s2d_concat(const float* lr_t, const float* lr_prev, const float* prev_hr,
                             void* x_nhwc, int n, int h, int w, long long lr_batch_stride,
                             long long hr_batch_stride, void* stream);
/* NCHW f32 [n,c,h,w] (c <= 64) -> NHWC bf16 [n,h,w,64] zero padded: the boundary conversion of
 * generator.forward(x) (code/models.py:78). */
int tg_pack_nchw_to_nhwc64(const float* in, void* out, int n, int c, int h, int w, void* stream);

/* ------------------------------------------------------ tensor-core convolutions (tcgen05) -- */

/* Packed-weight sizes/packers.  `kind`: 0 = Conv2d 3x3 s1 p1 (weight [cout,cin,3,3], code/ops.py:57-63)
 *                                       1 = ConvTranspose2d 3x3 s2 p1 op1 (weight [cin,cout,3,3], code/ops.py:45-54)
 *                                       2 = Conv2d 4x4 s2 p1 (weight [cout,cin,4,4]; discriminator blocks, code/models.py:90-94).
 * The packed blob holds bf16 weight blocks in MMA issue order followed by the fp32 bias
 * (zero if bias == NULL), padded to the kernel's channel granularity. */
size_t tg_packed_conv_bytes(int kind, int cin, int cout);
int tg_pack_weights(int kind, const float* weight, const float* bias, int cin, int cout,
                    void* packed, void* stream);

/* y = act(conv3x3(x) + bias) (+ residual).  x NHWC bf16 [n,h,w,cin_pad], y NHWC bf16 [n,h,w,cout].
 * cin_pad in {64,128} (51 -> 64), cout in {64,128}.  relu: apply ReLU before the residual add is
 * never needed by the reference; semantics are  y = relu?(conv+bias) + residual?  with at most
 * one of (relu, residual) set.  Replaces nn.Conv2d calls of code/models.py:55-56,68,75. */
int tg_conv3x3_fwd(const void* x, const void* packed, const void* residual, void* y, int n, int h,
                   int w, int cin_pad, int cout, int relu, int amode, void* stream);
/* `relu` of the conv entry points is an activation code: 0 none, 1 ReLU, 2 LeakyReLU(0.2) (code/ops.py:71-72). */
/* y = act(conv4x4s2(x) + bias): x NHWC bf16 [n,h,w,cin_pad] (h, w even) -> y NHWC bf16 [n,h/2,w/2,cout] for
 * cout in {64,128}; for cout == 3 (discriminator block5) y is the raw f32 NCHW [n,3,h/2,w/2] conv output.
 * Replaces the k4 s2 nn.Conv2d of discriminator_block (code/models.py:90-94,104,109,114,119,121). */
int tg_conv4x4s2_fwd(const void* x, const void* packed, void* y, int n, int h, int w, int cin_pad,
                     int cout, int act, void* stream);
/* y = relu?(convT3x3s2(x) + bias); x [n,h,w,c] -> y [n,2h,2w,cout].  code/models.py:72,74. */
int tg_convT3x3s2_fwd(const void* x, const void* packed, void* y, int n, int h, int w, int cin,
                      int cout, int relu, int amode, void* stream);
/* out = sigmoid(conv3x3(x) + bias), cout = 3, x NHWC bf16 [n,h,w,64] -> out NCHW f32 [n,3,h,w]
 * (callers .view the result: main.py:196).  logits != NULL additionally stores the pre-sigmoid
 * values.  code/models.py:76,86. */
int tg_conv3x3_out_sigmoid(const void* x, const void* packed, float* out, float* logits, int n,
                           int h, int w, int amode, void* stream);

/* ------------------------------------------------------ backward kernels (training, code/train.py:336,340) -- */

/* Weight gradient of Conv2d 3x3 s1 p1: dw[cout][cin][3][3] (f32, PyTorch layout) += sum over pixels of
 * dy[n,y,x,co] * x[n,y+ky-1,x+kx-1,ci].  x NHWC bf16 [n,h,w,pad64(cin)], dy NHWC bf16 [n,h,w,pad64(cout)]
 * (channels beyond cin / cout are ignored).  The result is ACCUMULATED with f32 atomics: zero dw first for a
 * plain gradient.  tcgen05 kernel with the pixel index as the GEMM reduction dimension (MN-major operands). */
int tg_conv3x3_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                     void* stream);
/* tg_conv3x3_wgrad plus the bias gradient db[cout] (f32) += sum over pixels of dy, in the same launch when cout <= 64 (the
 * epilogue warps sum the staged dY tiles while the MMAs run), else followed by the tg_bias_grad reduction. */
int tg_conv3x3_wgrad_bias(const void* x, const void* dy, float* dw, float* db, int n, int h, int w, int cin, int cout,
                          void* stream);
/* Same for ConvTranspose2d(k3,s2,p1,op1): x [n,h,w,pad64(cin)], dy [n,2h,2w,pad64(cout)] -> dw [cin][cout][3][3];
 * and for Conv2d(k4,s2,p1): x [n,2h,2w,pad64(cin)], dy [n,h,w,pad64(cout)] -> dw [cout][cin][4][4].  The operand at
 * twice the resolution is staged per parity phase with stride-2 TMA boxes. */
int tg_convT3x3s2_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                        void* stream);
int tg_conv4x4s2_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int cin, int cout,
                       void* stream);
/* db[c] += sum over pixels of dy[p][c] (bias gradient); dy NHWC bf16 [pixels][pad64(c)]. */
int tg_bias_grad(const void* dy, float* db, long long pixels, int c, void* stream);
/* Data gradients.  tg_pack_weights kinds 3 / 4 take the FORWARD layer's weight (and its cin, cout) and pack the
 * weights of the convolution that computes dX from dY (kind 3: conv3x3 with the kernel rotated by 180 degrees and
 * channels transposed; kind 4: the stride-2 3x3 convolution that is the adjoint of ConvTranspose2d(k3,s2,p1,op1)).
 * dx = (conv(dy) + residual?) zeroed where mask == 0 (ReLU backward with the saved post-ReLU activation); all
 * tensors NHWC bf16 padded to 64/128 channels.  tg_conv3x3_dgrad: dy [n,h,w,cout] -> dx [n,h,w,cin].
 * tg_convT3x3s2_dgrad: dy [n,2h,2w,cout] -> dx [n,h,w,cin] (h, w = the forward layer's INPUT size). */
int tg_conv3x3_dgrad(const void* dy, const void* packed_dgrad, const void* residual, const void* mask,
                     void* dx, int n, int h, int w, int cin, int cout, void* stream);
int tg_convT3x3s2_dgrad(const void* dy, const void* packed_dgrad, const void* mask, void* dx, int n, int h,
                        int w, int cin, int cout, void* stream);
/* Adjoint of Conv2d(k4,s2,p1) (pack kind 5): dy [n,h,w,cout] -> dx [n,2h,2w,cin], four output-parity phases of 2x2
 * taps.  mask_mode 1: ReLU backward (zero where mask == 0); 2: LeakyReLU(0.2) backward (x0.2 where mask <= 0). */
int tg_conv4x4s2_dgrad(const void* dy, const void* packed_dgrad, const void* mask, int mask_mode, void* dx,
                       int n, int h, int w, int cin, int cout, void* stream);

/* ----------------------------------------------------------------- generator (41 convs) ---- */

/* Parameters are passed as ONE flat f32 device buffer holding the tensors of
 * generator.state_dict() in its iteration order (conv.0.weight, conv.0.bias, resids.0.0.weight,
 * ... output.bias), so reference checkpoints map 1:1. */
size_t tg_gen_param_count(int num_resblock);
size_t tg_gen_packed_bytes(int num_resblock);
int tg_gen_pack(const float* flat_params, int num_resblock, void* packed, void* stream);
size_t tg_gen_workspace_bytes(int n, int h, int w);
/* generator.forward on an already packed input: x NHWC bf16 [n,h,w,64] -> out NCHW f32
 * [n,3,4h,4w] (sigmoid applied).  code/models.py:78-86. */
int tg_gen_forward(const void* packed, int num_resblock, const void* x_nhwc, float* out,
                   float* logits_or_null, void* workspace, size_t workspace_bytes, int n, int h,
                   int w, int amode, void* stream);
/* Whole recurrent clip on device (main.py:173-219): lr [n,t,3,h,w] f32 -> out [n,t,3,4h,4w] f32.
 * Frames are strictly sequential; the n clips of the batch are independent. */
int tg_gen_clip_forward(const void* packed, int num_resblock, const float* lr, float* out,
                        void* workspace, size_t workspace_bytes, int n, int t, int h, int w,
                        int amode, void* stream);

/* One step of that loop (main.py:199-216): x_t = cat(lr_t, s2d(deprocess(warp(prev_hr, flow(lr_prev))))),
 * out_t = generator(x_t).  lr_prev == prev_hr == NULL selects the first-frame input (main.py:191-195).
 * lr_t/lr_prev [n,3,h,w], prev_hr/out_t [n,3,4h,4w], all f32 with the given batch strides (elements), so a
 * caller can keep frames clip-major ([n,t,...]) or frame-major ([t,n,...]). */
int tg_gen_clip_step(const void* packed, int num_resblock, const float* lr_t, const float* lr_prev,
                     const float* prev_hr, float* out_t, void* workspace, size_t workspace_bytes, int n,
                     int h, int w, long long lr_batch_stride, long long prev_batch_stride,
                     long long out_batch_stride, int amode, void* stream);
/* Same step for a caller that chains frames the way main.py:199-216 does: prev_hr MUST be the out_t of the
 * immediately preceding tg_gen_clip_step[_chained] call on this workspace (same n, h, w), unmodified.  In
 * TG_AMODE_FRAME every step leaves a pixel-interleaved copy of its result in the workspace, and the chained step
 * gathers the warp taps from that copy (one 16-byte load per tap instead of three 4-byte ones); results are
 * bit-identical to tg_gen_clip_step.  tg_gen_clip_forward chains internally. */
int tg_gen_clip_step_chained(const void* packed, int num_resblock, const float* lr_t, const float* lr_prev,
                             const float* prev_hr, float* out_t, void* workspace, size_t workspace_bytes, int n,
                             int h, int w, long long lr_batch_stride, long long prev_batch_stride,
                             long long out_batch_stride, int amode, void* stream);

/* The chained step with a COMPACT result for the host side of the clip pipeline.  The recurrence
 * keeps running on the workspace's f32 interleaved copy of every estimate, so out_t is only what leaves the GPU:
 *   TG_OUT_F32  planar [n,3,4h,4w] float   (the parity default; same bits as tg_gen_clip_step_chained)
 *   TG_OUT_F16  planar [n,3,4h,4w] __half  = the f32 result rounded to nearest fp16: what the reference's GPU path
 *               emits under autocast (main.py:171-172,214)
 *   TG_OUT_U8   interleaved [n,4h,4w,3] uint8 = (uint8)(x * 255): exactly what save_as_gif turns the frames into
 *               before writing them (code/ops.py:234-237: transpose (0,2,3,1), * 255, astype(uint8))
 * first_frame != 0 selects the zero-state input (main.py:191-195) and MUST be used for frame 0 of a clip; every other
 * call must directly follow the previous frame's call on this workspace (same n, h, w).  out_batch_stride is in
 * elements of the output type (a multiple of 3).  TG_AMODE_FRAME only. */
#define TG_OUT_F32 0
#define TG_OUT_F16 1
#define TG_OUT_U8 2
int tg_gen_clip_step_fmt(const void* packed, int num_resblock, const float* lr_t, const float* lr_prev, int first_frame,
                         void* out_t, int out_format, void* workspace, size_t workspace_bytes, int n, int h, int w,
                         long long lr_batch_stride, long long out_batch_stride, void* stream);

/* Generator training (code/train.py:86-111,336): a forward that keeps every activation in the workspace and the
 * backward pass through all 41 layers.  The generator's inputs are detached in the reference (code/train.py:90,108),
 * so no input gradient is produced.
 *   tg_gen_forward_train: x NCHW f32 [n,51,h,w] -> out NCHW f32 [n,3,4h,4w] (one frame-kernel launch).
 *   tg_gen_pack_dgrad   : packs the data-gradient convolutions of every layer from the flat f32 parameters.
 *   tg_gen_backward     : dout = dL/dout [n,3,4h,4w] f32, out = the forward result; ADDS the parameter gradients into
 *                         flat_grad (f32, same flat layout as the parameters; zero it for a plain gradient).
 *                         Uses the workspace of the matching tg_gen_forward_train call. */
size_t tg_gen_train_workspace_bytes(int n, int h, int w, int num_resblock);
size_t tg_gen_packed_dgrad_bytes(int num_resblock);
int tg_gen_pack_dgrad(const float* flat_params, int num_resblock, void* packed_dgrad, void* stream);
int tg_gen_forward_train(const void* packed, int num_resblock, const float* x_nchw, float* out, void* workspace,
                         size_t workspace_bytes, int n, int h, int w, void* stream);
int tg_gen_backward(const void* packed_dgrad, int num_resblock, const float* dout, const float* out,
                    float* flat_grad, void* workspace, size_t workspace_bytes, int n, int h, int w, void* stream);
/* The recurrent generator loop of a training step (code/train.py:86-111) for b clips of t frames, every activation of
 * every frame kept: lr [b,t,3,h,w] f32 (r_inputs as the reference holds it) -> out [t,b,3,4h,4w] f32, FRAME-major.
 * Frame f's input is cat(lr[:,f], s2d(deprocess(warp(out[f-1], flow(lr[:,f-1]))))) (zeros for f = 0), produced by the
 * fused frame-input kernel; the generator inputs are detached in the reference (code/train.py:90,108), so the frames
 * are independent in the backward pass: workspace = tg_gen_train_workspace_bytes(t*b, h, w, ...) holds image f*b + clip,
 * and ONE tg_gen_backward(..., dout [t,b,3,4h,4w], out, ..., n = t*b, ...) back-propagates all frames as one batch. */
int tg_gen_clip_forward_train(const void* packed, int num_resblock, const float* lr, float* out, void* workspace,
                              size_t workspace_bytes, int b, int t, int h, int w, void* stream);

/* -------------------------------------------------- spatio-temporal discriminator ---- */

/* The discriminator's 27-channel input for t_batch = tb frame triplets in one pass (code/train.py:139-198 with
 * Dt_mergeDs=True, pingpang=False): out [tb,27,4h,4w] f32 NCHW =
 *   cat( before9[s]                                          [tb,9,4h,4w]  the triplet's target frames (:175),
 *        crop_pad(grid_sample(src frames 3s..3s+2, T_vel))   warped targets (real, :165) or generator outputs (fake, :187),
 *                                                            centre window kept and a border of crop_off pixels zeroed (:160-174),
 *        bilinear x4 of lr9[s]                               [tb,9,h,w] LR frames (functional.resize, :176-178) ).
 * Frame m = 3s+j is read at src + (m / ts)*src_stride_b + (m % ts)*src_stride_t (elements; [3,4h,4w] planes), so both the
 * clip-major targets and the frame-major generator output of tg_gen_clip_forward_train are addressed in place.
 * T_vel (:147-158) is computed on the fly from gsrc [tb*3,2,h,w], the LR planes the reference up-scales into its
 * velocity field: class m%3 == 0 -> up4(4*gsrc[m]); 1 -> zeros; 2 -> 2*up4(4*gsrc[m]) - 1; each [2,4h,4w] block re-viewed
 * as [4h,4w,2] like the reference's reshape.  grid_fp16 is a bit set: bit 0 rounds the grid to fp16 (the fake path's
 * .half(), :187); bit 1 leaves class 2 un-preprocessed, up4(4*gsrc[m]) (pingpang=True takes the flipped forward flow
 * as it is, :155). */
int tg_disc_input_assemble(const float* before9, const float* src, long long src_stride_b, long long src_stride_t,
                           int ts, const float* gsrc, const float* lr9, float* out, int tb, int h, int w,
                           int crop_off, int grid_fp16, void* stream);


/* discriminator(args) of code/models.py:97-146 with nb = args.discrim_resblocks, ch = args.discrim_channels
 * (64 or 128), fc_in = in-features of fc (48 in the reference = 128x128 inputs; 3*(h/32)*(w/32) in general).
 * Parameters: ONE flat f32 device buffer in named_parameters() order (conv.0.{weight,bias}, block1.0.weight,
 * block1.1.{weight,bias}, resids1.i.0.0.{weight,bias}, resids1.i.0.2.weight, resids1.i.1.{weight,bias}, ...,
 * block5.1.{weight,bias}, fc.{weight,bias}); tg_disc_pack derives the bf16 MMA-ordered conv blocks from it. */
size_t tg_disc_param_count(int nb, int ch, int fc_in);
size_t tg_disc_packed_bytes(int nb, int ch);
int tg_disc_pack(const float* flat_params, int nb, int ch, void* packed, void* stream);
size_t tg_disc_workspace_bytes(int n, int h, int w, int nb, int ch);
/* discriminator.forward (code/models.py:125-146): x NCHW f32 [n,27,h,w] -> prob [n] (sigmoid applied) and, for
 * every non-NULL feats[i] (host array of 4 device pointers), the layer_list feature map i as NCHW f32
 * ([n,64,h/2,w/2], [n,ch,h/4,w/4], [n,ch,h/8,w/8], [n,64,h/16,w/16]).  BatchNorm2d(eps=1e-3) uses batch statistics
 * when training != 0 (the reference never leaves train mode) and then updates the running statistics in place:
 * bn_running is a host array of 3 device pointers per BatchNorm layer {running_mean, running_var,
 * num_batches_tracked (int64)} in module order (block1, resids1.*, block2, resids2.*, block3, resids3.*, block4,
 * block5), or NULL to skip the update.  All activations stay in the workspace for a backward pass. */
int tg_disc_forward(const float* flat_params, const void* packed, int nb, int ch, int fc_in, const float* x,
                    float* prob, float* const* feats, void* const* bn_running, int training,
                    void* workspace, size_t workspace_bytes, int n, int h, int w, void* stream);

/* The same forward for `groups` (1 or 2) independent passes in ONE batch: x holds the groups back to back ([groups * n/groups,
 * 27, h, w]; the real and the fake triplets of code/train.py:181,199), every convolution runs once over all n samples, and
 * BatchNorm uses SEPARATE batch statistics per group and applies the running-statistics updates in group order — bit for
 * bit what two consecutive tg_disc_forward calls produce, in half the launches.  tg_disc_backward_groups is its backward. */
int tg_disc_forward_groups(const float* flat_params, const void* packed, int nb, int ch, int fc_in, const float* x,
                           float* prob, float* const* feats, void* const* bn_running, int training, void* workspace,
                           size_t workspace_bytes, int n, int groups, int h, int w, void* stream);

/* discriminator backward (code/train.py:340, scaler.scale(discrim_loss).backward()).  The reference detaches the
 * discriminator's inputs (code/train.py:181,199) and the layer features (code/train.py:214), so the gradient enters
 * through prob only and no input gradient is produced.
 *   tg_disc_pack_dgrad: packs the data-gradient convolutions of every conv but conv.0 from the flat f32 parameters.
 *   tg_disc_backward  : dprob = dL/dprob [n] f32, prob = the forward result; ADDS the parameter gradients into flat_grad
 *                       (f32, same flat layout as the parameters).  Uses the workspace of the matching tg_disc_forward
 *                       call (training != 0), which holds every activation and the batch statistics. */
size_t tg_disc_packed_dgrad_bytes(int nb, int ch);
int tg_disc_pack_dgrad(const float* flat_params, int nb, int ch, void* packed_dgrad, void* stream);
int tg_disc_backward(const float* flat_params, const void* packed_dgrad, int nb, int ch, int fc_in,
                     const float* dprob, const float* prob, float* flat_grad, void* workspace,
                     size_t workspace_bytes, int n, int h, int w, void* stream);
int tg_disc_backward_groups(const float* flat_params, const void* packed_dgrad, int nb, int ch, int fc_in,
                            const float* dprob, const float* prob, float* flat_grad, void* workspace,
                            size_t workspace_bytes, int n, int groups, int h, int w, void* stream);

/* ------------------------------------------------------------ BatchNorm building blocks ---- */

/* nn.BatchNorm2d(eps=1e-3, momentum=0.1) in training mode (code/ops.py:75-77; code/models.py:90-94,106,130) on NHWC
 * tensors, the three passes tg_disc_forward / tg_disc_backward are made of.  c in {64,128}; `stats` is [c][4] f32 =
 * {mean, rstd, a = gamma*rstd, b = beta - mean*a}; workspace = tg_workspace_bytes_bn() bytes, 256-byte aligned.
 *   tg_bn_stats: batch statistics of x [pixels][c] f32 (deterministic two-level reduction) -> stats; when
 *                running_mean != NULL also the in-place running-statistics update (unbiased variance,
 *                num_batches_tracked += 1) exactly as nn.BatchNorm2d does.
 *   tg_bn_apply: y = act(a*x + b) (+ skip); act 0 none / 2 LeakyReLU(0.2); written as f32 (y_f32) and / or bf16 (y_bf16).
 *   tg_bn_bwd  : g_out [pixels][c] bf16 = dL/dy, x = the forward input, act_out = the forward f32 OUTPUT when the block
 *                ended in LeakyReLU(0.2) (else NULL); dx [pixels][c] bf16 = dL/dx; dgamma / dbeta are ADDED to. */
size_t tg_workspace_bytes_bn(void);
int tg_bn_stats(const float* x, long long pixels, int c, const float* gamma, const float* beta, float* stats,
                float* running_mean, float* running_var, long long* num_batches_tracked, void* workspace,
                size_t workspace_bytes, void* stream);
int tg_bn_apply(const float* x, const float* skip, float* y_f32, void* y_bf16, long long pixels, int c,
                const float* stats, int act, void* stream);
int tg_bn_bwd(const void* g_out, const float* x, const float* act_out, void* dx, long long pixels, int c,
              const float* stats, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------ optimizer step on the flat buckets ---- */

/* Fused multi-tensor Adam.  Replaces scaler.step(optimizer) / scaler.update() of code/train.py:337-338,
 * 341-342 for the optimizers main.py:239-243 builds (torch.optim.Adam, amsgrad=False, weight_decay=0).  `params`,
 * `grads`, `exp_avg`, `exp_avg_sq` are flat f32 buffers of n elements in the networks' flat parameter layout (16-byte
 * aligned); lr, step, inv_scale, found_inf are DEVICE scalars (f32), so the launches are CUDA-graph capturable and a
 * replayed graph follows StepLR / the loss scale.
 *   tg_grad_check_finite: *found_inf = 1 if any gradient is inf / NaN, else 0 (GradScaler.unscale_'s check).
 *   tg_adam_step        : g = grads * (*inv_scale, 1 if NULL), written back to grads when inv_scale != NULL (GradScaler.unscale_
 *                         works in place); skipped entirely when found_inf != NULL and *found_inf != 0;
 *                         t = *step + 1; exp_avg = b1*exp_avg + (1-b1)*g; exp_avg_sq = b2*exp_avg_sq + (1-b2)*g*g;
 *                         params -= (*lr / (1 - b1^t)) * exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - b2^t) + eps).
 *   tg_scaler_update    : *step += 1 unless inf (step may be NULL); GradScaler.update(): on inf scale *= backoff_factor and
 *                         growth_tracker = 0, else growth_tracker += 1 and every growth_interval clean steps
 *                         scale *= growth_factor (scale / growth_tracker may both be NULL: plain optimizer without a scaler). */
int tg_grad_check_finite(const float* grads, long long n, float* found_inf, void* stream);
int tg_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, const float* lr,
                 float beta1, float beta2, float eps, const float* step, const float* inv_scale, const float* found_inf,
                 void* stream);
int tg_scaler_update(float* scale, int* growth_tracker, const float* found_inf, float growth_factor, float backoff_factor,
                     int growth_interval, float* step, void* stream);

/* Workspace queries under the names SURVEY.md 8b lists (tg_workspace_bytes_<op>); the single-layer conv / glue entry
 * points need no workspace.  Same values as tg_gen_workspace_bytes / tg_gen_train_workspace_bytes / tg_disc_workspace_bytes. */
size_t tg_workspace_bytes_gen_forward(int n, int h, int w);
size_t tg_workspace_bytes_gen_train(int n, int h, int w, int num_resblock);
size_t tg_workspace_bytes_disc(int n, int h, int w, int nb, int ch);

#ifdef __cplusplus
}
#endif
#endif /* TECOGAN_B200_H */
