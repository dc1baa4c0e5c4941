// Microbenchmark: issue-to-retire cost of tcgen05.mma (M=128, K=16, bf16, SS operands) as a
// function of N and of the A descriptor (aligned SBO=1024 vs row-shifted SBO=1280).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench scripts/mma_bench.cu
#include <cstdio>
#include <cstdlib>

#include "../pytorch-tecogan_b200/csrc/tg_common.cuh"

void tg_set_error(const char*, ...) {}
int tg_num_sms() { return 148; }

using namespace tg;

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: A from smem aligned; 1: A from smem row-shifted (SBO 1280); 2: A from TMEM
// nacc: consecutive MMAs rotate over this many accumulators (1 = one dependent accumulate chain)
template <int NACC>
__global__ void __launch_bounds__(128, 1) k(int n, int mode, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tptr), 512);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t a0 = base, b0 = base + 96 * 1024;
    const uint32_t sbo = mode == 1 ? 1280 : 1024;
    uint64_t ad[36], bd[4];
#pragma unroll
    for (int v = 0; v < 36; ++v) {
      const uint32_t aoff = (mode == 1 ? (uint32_t)((v / 4) * 128 * 11) : (uint32_t)((v / 4) * 16384 % 65536)) + (v & 3) * 32;
      ad[v] = umma_desc_sw128(a0 + aoff, sbo);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) bd[v] = umma_desc_sw128(b0 + v * 32, 1024);
    const uint32_t dstride = (n <= 128) ? 128 : 256;      // accumulator spacing in TMEM columns
    uint32_t ph = 0;
    long long best = 1ll << 60;
    for (int trial = 0; trial < 5; ++trial) {
      const long long t0 = clock64();
      for (int r = 0; r < reps / 36; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int v = 0; v < 36; ++v) {
            const uint32_t d = tm + (uint32_t)(v % NACC) * dstride % 512;
            if (mode == 2) umma_ts(d, tm + 480 + (v & 3) * 8, bd[v & 3], idesc, 1);
            else umma_bf16(d, ad[v], bd[v & 3], idesc, 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph);
      ph ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// mode 3: cta_group::2 (CTA pair, UMMA M = 256): every CTA holds its own 128 A rows and HALF of the B rows (N/2) in
// shared memory; the leader CTA issues for both.  Operand data is irrelevant for timing.
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit_every > 0: after every commit_every MMAs the issuer also issues `ncommit` multicast commits to scratch
// barriers nobody waits on (what the frame kernel does per item: stage free + accumulator ready)
template <int NACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(int n, int rowshift, int reps, long long* out,
                                                                      int commit_every = 0, int ncommit = 0) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t scratch[2];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&scratch[0]), 1); mbar_init(smem_u32(&scratch[1]), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x < 32 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, n);
    const uint32_t a0 = base, b0 = base + 96 * 1024;
    const uint32_t sbo = rowshift ? 1280 : 1024;
    uint64_t ad[36], bd[4];
#pragma unroll
    for (int v = 0; v < 36; ++v) {
      const uint32_t aoff = (rowshift ? (uint32_t)((v / 4) * 128 * 11) : (uint32_t)((v / 4) * 16384 % 65536)) + (v & 3) * 32;
      ad[v] = umma_desc_sw128(a0 + aoff, sbo);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) bd[v] = umma_desc_sw128(b0 + v * 32, 1024);
    const uint32_t dstride = (n <= 128) ? 128 : 256;
    uint32_t ph = 0;
    long long best = 1ll << 60;
    for (int trial = 0; trial < 5; ++trial) {
      const long long t0 = clock64();
      for (int r = 0; r < reps / 36; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int v = 0; v < 36; ++v) {
            umma2_bf16(tm + (uint32_t)(v % NACC) * dstride % 512, ad[v], bd[v & 3], idesc, 1);
            if (commit_every > 0 && (v + 1) % commit_every == 0)
              for (int c = 0; c < ncommit; ++c)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&scratch[c])), "h"((uint16_t)3) : "memory");
          }
        }
        __syncwarp();
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph);
      ph ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x / 2] = best;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int NACC>
void run2(int n, int rowshift, int reps, long long* d, int commit_every = 0, int ncommit = 0) {
  cudaFuncSetAttribute(k2<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k2<NACC><<<148, 128, 200 * 1024>>>(n, rowshift, reps, d, commit_every, ncommit);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("2cta n %d: %s\n", n, cudaGetErrorString(e)); exit(1); }
  long long h[74];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (int i = 0; i < 74; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
  const int r = reps / 36 * 36;
  printf("mode 3 (cta_group::2 M=256, A %s) nacc %d N %3d commits %d per %2d MMAs : %6.1f cycles/MMA (min %6.1f)  -> %3.0f%% of N/2 floor\n",
         rowshift ? "row-shifted" : "aligned", NACC, n, ncommit, commit_every, (double)mx / r, (double)mn / r, 100.0 * (n / 2.0) / ((double)mx / r));
}

template <int NACC>
void run(int n, int mode, int reps, long long* d) {
  cudaFuncSetAttribute(k<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<NACC><<<148, 128, 200 * 1024>>>(n, mode, reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (int i = 0; i < 148; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
  const int r = reps / 36 * 36;
  printf("mode %d (%-18s) nacc %d N %3d : %6.1f cycles/MMA (min %6.1f)  -> %3.0f%% of N/2 floor\n", mode,
         mode == 0 ? "A smem aligned" : mode == 1 ? "A smem row-shifted" : "A tmem", NACC, n, (double)mx / r,
         (double)mn / r, 100.0 * (n / 2.0) / ((double)mx / r));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int reps = 36 * 64;
  const int ns[] = {16, 32, 64, 96, 128, 192, 256};
  for (int mode = 0; mode < 3; ++mode)
    for (int n : ns) {
      run<1>(n, mode, reps, d);
      run<2>(n, mode, reps, d);
      if (n <= 128) run<4>(n, mode, reps, d);
    }
  const int ns2[] = {32, 64, 96, 128, 192, 256};
  for (int rs = 0; rs < 2; ++rs)
    for (int n : ns2) {
      run2<1>(n, rs, reps, d);
      run2<2>(n, rs, reps, d);
    }
  // cost of the per-item commits
  for (int n : {64, 192})
    for (int ce : {36, 12, 4})
      for (int nc : {1, 2}) run2<2>(n, 0, reps, d, ce, nc);
  return 0;
}
