// Internal interface of the discriminator building blocks (tg_bn.cu) used by tg_discriminator.cu.
#pragma once
#include "tg_conv_tc.cuh"

namespace tg {

// stats layout per BN layer and group: [c][4] = {mean, rstd, a = gamma*rstd, b = beta - mean*a}; group g at stats + g*512.
// `pixels` (n, hw for the head) are PER GROUP; tensors hold the groups back to back.
int bn_stats_launch(const void* x, long long pixels, int c, const float* gamma, const float* beta, float* partial,
                    unsigned int* ticket, float* stats, float* running_mean, float* running_var,
                    long long* num_batches_tracked, cudaStream_t st, int groups = 1);
size_t bn_partial_floats();
int bn_fold_running_launch(int c, const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float* stats, cudaStream_t st);
int bn_apply_launch(const void* x, const void* skip, void* y32, void* y16, long long pixels, int c, const float* stats,
                    int act, cudaStream_t st, int groups = 1);
int nhwc_to_nchw_f32_launch(const void* in, float* out, int n, int c, long long hw, cudaStream_t st);
int disc_head_launch(const float* r, int n, int hw, const float* gamma, const float* beta, int training,
                     float* running_mean, float* running_var, long long* nbt, const float* fc_w, const float* fc_b,
                     float* y, float* stats, float* logit, float* prob, cudaStream_t st, int groups = 1);

// BatchNorm backward: g_out (bf16) = gradient of the block output; act != null: the block ended in LeakyReLU(0.2).
// dx (bf16) = gradient of the raw conv output; dgamma/dbeta are ADDED to; red = 2*c floats of scratch.
int bn_bwd_launch(const void* g_out, const void* x, const void* act, void* dx, long long pixels, int c, const float* stats,
                  float* partial, unsigned int* ticket, float* red, float* dgamma, float* dbeta, cudaStream_t st, int groups = 1);
int disc_head_bwd_launch(const float* dprob, const float* prob, const float* y, const float* r, int n, int hw,
                         const float* stats, const float* fc_w, float* d_fc_w, float* d_fc_b, float* dgamma, float* dbeta,
                         float* dlogit, void* dr, cudaStream_t st, int groups = 1);

}  // namespace tg
