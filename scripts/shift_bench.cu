// Microbenchmark: what tcgen05.shift.down does to Tensor Memory (which rows / columns move) and what it costs, alone and
// interleaved with tcgen05.mma — the question behind DESIGN.md section 7 (1b): could it replace the partial-sum shuffles of
// the N-stacked 3x3 conv epilogue?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/shift_bench scripts/shift_bench.cu
#include <cstdio>
#include <cstdlib>

#include "../pytorch-tecogan_b200/csrc/tg_common.cuh"

void tg_set_error(const char*, ...) {}
int tg_num_sms() { return 148; }

using namespace tg;

__device__ __forceinline__ void tmem_st_32x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return v;
}
__device__ __forceinline__ void tmem_shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}

// phase 0: semantics.  TMEM[lane][col] = lane * 1000 + col for 64 columns; ONE shift at column `scol`; dump the result.
// phase 1: cost of `reps` shifts (round-robin over `ncols8` 8-column blocks) + one commit.
// phase 2: the same shifts interleaved 1:1 with N = 192 MMAs (do they share the tensor pipe?).
__global__ void __launch_bounds__(128, 1) k(int phase, int scol, int reps, int ncols8, uint32_t* dump, long long* cyc) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  const uint32_t tq = tm + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < 64; ++c) tmem_st_32x1(tq + c, static_cast<uint32_t>((warp * 32 + lane) * 1000 + c));
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (phase == 0) {
      if (elect_one()) { tmem_shift_down(tm + scol); umma_commit(smem_u32(&bar)); }
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, 192);
      const uint64_t ad = umma_desc_sw128(base, 1024), bd = umma_desc_sw128(base + 32 * 1024, 1024);
      uint32_t ph = 0;
      long long best = 1ll << 60;
      for (int trial = 0; trial < 5; ++trial) {
        const long long t0 = clock64();
        if (elect_one()) {
          for (int r = 0; r < reps; ++r) {
            if (phase == 2) umma_bf16(tm + 256, ad, bd, idesc, 1);
            if (ncols8 > 0) tmem_shift_down(tm + static_cast<uint32_t>((r % ncols8) * 8));
          }
          umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), ph);
        ph ^= 1;
        const long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
      }
      if (lane == 0) cyc[0] = best;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (phase == 0)
    for (int c = 0; c < 64; ++c) dump[(warp * 32 + lane) * 64 + c] = tmem_ld_32x1(tq + c);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  uint32_t* dump;
  long long* cyc;
  cudaMalloc(&dump, 128 * 64 * 4);
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  static uint32_t h[128 * 64];
  for (int scol : {0, 16}) {
    k<<<1, 128, 100 * 1024>>>(0, scol, 0, 0, dump, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("phase 0 scol %d failed: %s\n", scol, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, dump, sizeof(h), cudaMemcpyDeviceToHost);
    int moved_cols_lo = 64, moved_cols_hi = -1;
    for (int c = 0; c < 64; ++c) {
      bool moved = false;
      for (int l = 0; l < 128; ++l) moved |= h[l * 64 + c] != static_cast<uint32_t>(l * 1000 + c);
      if (moved) { if (c < moved_cols_lo) moved_cols_lo = c; if (c > moved_cols_hi) moved_cols_hi = c; }
    }
    printf("shift at column %d: columns changed [%d, %d]\n", scol, moved_cols_lo, moved_cols_hi);
    const int c = scol;
    for (int l : {0, 1, 2, 15, 16, 17, 30, 31, 32, 33, 34, 63, 64, 65, 95, 96, 97, 126, 127})
      printf("  lane %3d col %2d: now holds the value of lane %u (col %u)\n", l, c, h[l * 64 + c] / 1000, h[l * 64 + c] % 1000);
  }
  for (int phase : {1, 2}) {
    for (int ncols8 : {0, 1, 8, 24}) {
      if (phase == 1 && ncols8 == 0) continue;
      const int reps = 960;
      k<<<1, 128, 100 * 1024>>>(phase, 0, reps, ncols8, dump, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("phase %d failed: %s\n", phase, cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("phase %d (%s), shifts over %2d column blocks: %.1f cycles per iteration\n", phase,
             phase == 1 ? "shifts only" : "one N=192 MMA + one shift", ncols8, static_cast<double>(c) / reps);
    }
  }
  return 0;
}
