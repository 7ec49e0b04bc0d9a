// Microbenchmark (experiment, not product): SM-side throughput of the two ways to push a runlet's keys
// into the L2-resident accumulation ring:
//   mode 0: RED.MAX.U32, lane = channel (2 runlets x 16 channels per warp instruction), cell stride CP words
//   mode 1: cp.reduce.async.bulk max.u32 (TMA reduce), one lane = one runlet, BYTES per runlet from shared memory
//   mode 2: RED.MAX.U32, lane = runlet (32 different cells per instruction)  — the height channel pattern
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/exp/red_bench scripts/ubench/red_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128) k(uint32_t* acc, uint32_t ncell, int cp, int iters, int bytes, int local) {
  extern __shared__ __align__(128) uint32_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = (blockIdx.x * 4 + warp);
  uint32_t* mine = sm + warp * 32 * 32;  // 32 runlets x up to 128 B
  for (int i = lane; i < 32 * 32; i += 32) mine[i] = hash(gw * 1024 + i) | 1u;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  // cells: pseudo-random walk; `local` > 0 keeps a warp's successive cells within a window (coherent scene)
  uint32_t base = hash(gw) % ncell;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      const int s = lane >> 4, c = lane & 15;
      const uint32_t r = hash(gw * 65536u + it * 2 + s);
      const uint32_t cell = local ? (base + (r % local)) % ncell : r % ncell;
      asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(acc + (size_t)cell * cp + c), "r"(r | 1u) : "memory");
    } else if (MODE == 1) {
      const uint32_t r = hash(gw * 65536u + it * 32 + lane);
      const uint32_t cell = local ? (base + (r % local)) % ncell : r % ncell;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.max.u32 [%0], [%1], %2;"
                   ::"l"(acc + (size_t)cell * cp), "r"(smem_u32(mine + lane * 32)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    } else if (MODE == 3) {  // lane = channel, ONE runlet per instruction: lanes 16-31 idle
      const uint32_t r = hash(gw * 65536u + it);
      const uint32_t cell = local ? (base + (r % local)) % ncell : r % ncell;
      if (lane < 16)
        asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(acc + (size_t)cell * cp + lane), "r"(r | 1u) : "memory");
    } else if (MODE == 4) {  // 32 lanes on 32 consecutive words of ONE cell (128 B)
      const uint32_t r = hash(gw * 65536u + it);
      const uint32_t cell = local ? (base + (r % local)) % ncell : r % ncell;
      asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(acc + (size_t)cell * cp + lane), "r"(r | 1u) : "memory");
    } else if (MODE == 5) {  // 4 runlets x 8 channels per instruction
      const int s = lane >> 3, c = lane & 7;
      const uint32_t r = hash(gw * 65536u + it * 4 + s);
      const uint32_t cell = local ? (base + (r % local)) % ncell : r % ncell;
      asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(acc + (size_t)cell * cp + c), "r"(r | 1u) : "memory");
    } else {
      const uint32_t r = hash(gw * 65536u + it * 32 + lane);
      const uint32_t cell = local ? (base + (r % local)) % ncell : r % ncell;
      asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(acc + (size_t)cell * cp), "r"(r | 1u) : "memory");
    }
    if (local && (it & 15) == 15) base = (base + local) % ncell;
  }
  if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const uint32_t ncell = 160000 * 4;  // 4 ring slots
  uint32_t* acc;
  CK(cudaMalloc(&acc, (size_t)ncell * 32 * 4 + 256));
  CK(cudaMemset(acc, 0, (size_t)ncell * 32 * 4 + 256));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int smem = 4 * 32 * 32 * 4;
  auto run = [&](const char* name, int mode, int cp, int ctas_per_sm, int iters, int bytes, int local, double runlets_per_iter) -> int {
    const int grid = 148 * ctas_per_sm;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, 128, smem>>>(acc, ncell, cp, iters, bytes, local);
      if (mode == 1) k<1><<<grid, 128, smem>>>(acc, ncell, cp, iters, bytes, local);
      if (mode == 2) k<2><<<grid, 128, smem>>>(acc, ncell, cp, iters, bytes, local);
      if (mode == 3) k<3><<<grid, 128, smem>>>(acc, ncell, cp, iters, bytes, local);
      if (mode == 4) k<4><<<grid, 128, smem>>>(acc, ncell, cp, iters, bytes, local);
      if (mode == 5) k<5><<<grid, 128, smem>>>(acc, ncell, cp, iters, bytes, local);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_ops = (double)grid * 4 * iters;
    const double runlets = warp_ops * runlets_per_iter;
    printf("%-44s cp=%2d ctas/sm=%d local=%5d: %.3f ms  %.2f cyc/warp-op/SM  %.1f M runlets/ms  (config-2 room needs 7.7M runlets/step)\n",
           name, cp, ctas_per_sm, local, ms, ms * 1e-3 * 1.965e9 / (warp_ops / 148), runlets / ms / 1e6);
    return 0;
  };
  for (int local : {0, 2000}) {
    for (int c : {4, 5}) {
      run("RED lane=channel (2 runlets/instr)", 0, 17, c, 2000, 0, local, 2);
      run("RED lane=channel (2 runlets/instr)", 0, 16, c, 2000, 0, local, 2);
      run("RED lane=runlet (32 cells/instr)", 2, 17, c, 500, 0, local, 32);
      run("RED lane=channel, 1 runlet x 16 ch/instr", 3, 17, c, 2000, 0, local, 1);
      run("RED 1 cell x 32 words/instr (128 B)", 4, 32, c, 2000, 0, local, 1);
      run("RED 4 runlets x 8 ch/instr", 5, 9, c, 2000, 0, local, 4);
      run("bulk reduce 64 B/runlet (32 runlets/instr)", 1, 16, c, 500, 64, local, 32);
      run("bulk reduce 80 B/runlet (32 runlets/instr)", 1, 20, c, 500, 80, local, 32);
      run("bulk reduce 128 B/runlet (32 runlets/instr)", 1, 32, c, 500, 128, local, 32);
    }
  }
  return 0;
}
