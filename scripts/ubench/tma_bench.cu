// Microbenchmark (experiment): how fast can one SM pull [17 rows x 2 KB] tiles with cp.async.bulk (1-D TMA bulk
// copies, one per row) when nothing else happens?  CTAs of 32 threads, `stages` tiles in flight per CTA.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/exp/tma_bench scripts/ubench/tma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) k(const float* __restrict__ src, size_t plane, int rows, int row_bytes, int tiles_per_frame,
                                        int frames, int stages, unsigned* counter, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x;
  const int stage_bytes = rows * (row_bytes + 16);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  if (lane == 0) for (int s = 0; s < stages; ++s)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const unsigned total = (unsigned)tiles_per_frame * frames;
  float acc = 0.f;
  unsigned issued = 0, done = 0;
  unsigned t_of[8];
  auto issue = [&](int s) -> bool {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total) return false;
    const unsigned f = t / tiles_per_frame, i = t % tiles_per_frame;
    if (lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(rows * row_bytes) : "memory");
    }
    __syncwarp();
    for (int r = lane; r < rows; r += 32) {
      const float* p = src + ((size_t)f * rows + r) * plane + (size_t)i * (row_bytes / 4);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + (size_t)s * stage_bytes + r * (row_bytes + 16))), "l"(p), "r"(row_bytes), "r"(smem_u32(bars + s)) : "memory");
    }
    t_of[s & 7] = t;
    return true;
  };
  int live = 0;
  for (int s = 0; s < stages; ++s) if (issue(s)) { ++issued; ++live; }
  while (live > 0) {
    const int s = done % stages;
    const unsigned parity = (done / stages) & 1u;
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bars + s)), "r"(parity) : "memory");
    acc += reinterpret_cast<float*>(smem + (size_t)s * stage_bytes)[lane];
    ++done; --live;
    __syncwarp();
    if (issue(s)) { ++issued; ++live; }
  }
  if (acc == 123.f) *sink = acc;
}

int main() {
  const int rows = 17, row_bytes = 2048, H = 480, W = 640, frames = 64;
  const size_t plane = (size_t)H * W;
  const int tiles = (int)(plane * 4 / row_bytes);
  float* src; unsigned* counter; float* sink;
  CK(cudaMalloc(&src, plane * 4 * rows * frames));
  CK(cudaMemset(src, 0, plane * 4 * rows * frames));
  CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int ctas : {1, 2, 4, 6}) for (int stages : {1, 2, 3}) {
    const int smem = stages * rows * (row_bytes + 16) + 64;
    if ((smem + 1024) * ctas > 227 * 1024) continue;
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(counter, 0, 4));
      cudaEventRecord(e0);
      k<<<148 * ctas, 32, smem>>>(src, plane, rows, row_bytes, tiles, frames, stages, counter, sink);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double bytes = (double)plane * 4 * rows * frames;
    printf("ctas/SM=%d stages/CTA=%d tiles in flight/SM=%2d: %.3f ms  %.0f GB/s  (%.1f B/cycle/SM)\n", ctas, stages, ctas * stages, best,
           bytes / best / 1e6, bytes / best / 1e6 * 1e9 / 148 / 1.965e9 / 1e3 * 1e3 / 1e6 * 1e6 / 1e9 * 1e0);
  }
  return 0;
}
