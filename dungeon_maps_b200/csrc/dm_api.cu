// Library-level entry points: ABI introspection, launch accounting and the HOST-buffer
// variant of the projection (host→device copy, kernels, device→host copy, pipelined over
// two streams so that the PCIe directions and the kernels overlap).
#include <mutex>

#include "dm_common.cuh"

namespace dm {
int64_t g_launches = 0;

namespace {
struct Scratch {
  int device = -1;
  cudaStream_t streams[2] = {nullptr, nullptr};
  void* buf[2] = {nullptr, nullptr};   // per-stream staging: inputs + outputs of one chunk
  size_t buf_bytes[2] = {0, 0};
  void* ws[2] = {nullptr, nullptr};    // per-stream accumulation ring (kept zeroed)
  size_t ws_bytes[2] = {0, 0};
};
Scratch g_scratch;
std::mutex g_mu;

void release_locked() {
  for (int i = 0; i < 2; ++i) {
    if (g_scratch.buf[i]) cudaFree(g_scratch.buf[i]);
    if (g_scratch.ws[i]) cudaFree(g_scratch.ws[i]);
    if (g_scratch.streams[i]) cudaStreamDestroy(g_scratch.streams[i]);
    g_scratch.buf[i] = g_scratch.ws[i] = nullptr;
    g_scratch.buf_bytes[i] = g_scratch.ws_bytes[i] = 0;
    g_scratch.streams[i] = nullptr;
  }
  g_scratch.device = -1;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace
}  // namespace dm

using namespace dm;

extern "C" int dm_abi_version(void) { return DM_ABI_VERSION; }

extern "C" const char* dm_build_info(void) {
  return "dungeon_maps_b200 sm_100a nvcc " __DATE__ " --fmad=false";
}

extern "C" int64_t dm_launch_count(void) { return g_launches; }

extern "C" void dm_release_scratch(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  release_locked();
}

extern "C" int dm_orth_project_host_f32(const float* depth, const float* values, const uint8_t* valid,
                                        const DmProjSample* samples, const DmProjCfg* cfg, int32_t b,
                                        float* topdown, uint8_t* mask, float* height, int32_t device) {
  if (!cfg || b < 0) return DM_EINVAL;
  if (b == 0) return DM_OK;
  if (!depth || !samples || !topdown || !mask) return DM_EINVAL;
  if (cfg->C > 0 && !values) return DM_EINVAL;
  std::lock_guard<std::mutex> lk(g_mu);
  DM_CUDA_OK(cudaSetDevice(device));
  if (g_scratch.device != device) {
    release_locked();
    g_scratch.device = device;
    for (int i = 0; i < 2; ++i) DM_CUDA_OK(cudaStreamCreateWithFlags(&g_scratch.streams[i], cudaStreamNonBlocking));
  }
  const size_t N = (size_t)cfg->H * cfg->W, M = (size_t)cfg->Mh * cfg->Mw;
  const int Cv = cfg->C > 0 ? cfg->C : 1;
  const bool want_h = cfg->C > 0 && cfg->want_height && height;
  // chunk: a few frames per stream so that copies of one chunk overlap kernels of the other
  int chunk = 8;
  if (chunk > b) chunk = b;
  // staging layout of one chunk (every section 256-byte aligned)
  const size_t o_depth = 0;
  const size_t o_values = align_up(o_depth + chunk * N * 4, 256);
  const size_t o_valid = align_up(o_values + (size_t)chunk * cfg->C * N * 4, 256);
  const size_t o_samples = align_up(o_valid + (valid ? chunk * N : 0), 256);
  const size_t o_top = align_up(o_samples + chunk * sizeof(DmProjSample), 256);
  const size_t o_mask = align_up(o_top + (size_t)chunk * Cv * M * 4, 256);
  const size_t o_height = align_up(o_mask + (size_t)chunk * Cv * M, 256);
  const size_t total = align_up(o_height + (want_h ? chunk * M * 4 : 0), 256);
  const size_t ws_need = dm_orth_project_workspace_bytes(cfg, chunk);
  for (int i = 0; i < 2; ++i) {
    if (g_scratch.buf_bytes[i] < total) {
      if (g_scratch.buf[i]) DM_CUDA_OK(cudaFree(g_scratch.buf[i]));
      g_scratch.buf[i] = nullptr; g_scratch.buf_bytes[i] = 0;
      DM_CUDA_OK(cudaMalloc(&g_scratch.buf[i], total));
      g_scratch.buf_bytes[i] = total;
    }
    if (g_scratch.ws_bytes[i] < ws_need) {
      if (g_scratch.ws[i]) DM_CUDA_OK(cudaFree(g_scratch.ws[i]));
      g_scratch.ws[i] = nullptr; g_scratch.ws_bytes[i] = 0;
      DM_CUDA_OK(cudaMalloc(&g_scratch.ws[i], ws_need));
      DM_CUDA_OK(cudaMemset(g_scratch.ws[i], 0, ws_need));
      g_scratch.ws_bytes[i] = ws_need;
    }
  }
  int rc = DM_OK;
  int k = 0;
  for (int f0 = 0; f0 < b && rc == DM_OK; f0 += chunk, k ^= 1) {
    const int nf = (b - f0) < chunk ? (b - f0) : chunk;
    cudaStream_t st = g_scratch.streams[k];
    char* base = static_cast<char*>(g_scratch.buf[k]);
    float* d_depth = reinterpret_cast<float*>(base + o_depth);
    float* d_values = cfg->C > 0 ? reinterpret_cast<float*>(base + o_values) : nullptr;
    uint8_t* d_valid = valid ? reinterpret_cast<uint8_t*>(base + o_valid) : nullptr;
    DmProjSample* d_samples = reinterpret_cast<DmProjSample*>(base + o_samples);
    float* d_top = reinterpret_cast<float*>(base + o_top);
    uint8_t* d_mask = reinterpret_cast<uint8_t*>(base + o_mask);
    float* d_height = want_h ? reinterpret_cast<float*>(base + o_height) : nullptr;
    DM_CUDA_OK(cudaMemcpyAsync(d_depth, depth + (size_t)f0 * N, (size_t)nf * N * 4, cudaMemcpyHostToDevice, st));
    if (d_values)
      DM_CUDA_OK(cudaMemcpyAsync(d_values, values + (size_t)f0 * cfg->C * N, (size_t)nf * cfg->C * N * 4,
                                 cudaMemcpyHostToDevice, st));
    if (d_valid)
      DM_CUDA_OK(cudaMemcpyAsync(d_valid, valid + (size_t)f0 * N, (size_t)nf * N, cudaMemcpyHostToDevice, st));
    DM_CUDA_OK(cudaMemcpyAsync(d_samples, samples + f0, (size_t)nf * sizeof(DmProjSample), cudaMemcpyHostToDevice, st));
    DmProjCfg c = *cfg;
    if (!want_h) c.want_height = 0;
    rc = dm_orth_project_f32(d_depth, d_values, d_valid, d_samples, &c, nf, d_top, d_mask, d_height,
                             g_scratch.ws[k], g_scratch.ws_bytes[k], st);
    if (rc != DM_OK) break;
    DM_CUDA_OK(cudaMemcpyAsync(topdown + (size_t)f0 * Cv * M, d_top, (size_t)nf * Cv * M * 4, cudaMemcpyDeviceToHost, st));
    DM_CUDA_OK(cudaMemcpyAsync(mask + (size_t)f0 * Cv * M, d_mask, (size_t)nf * Cv * M, cudaMemcpyDeviceToHost, st));
    if (d_height)
      DM_CUDA_OK(cudaMemcpyAsync(height + (size_t)f0 * M, d_height, (size_t)nf * M * 4, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < 2; ++i) {
    const cudaError_t e = cudaStreamSynchronize(g_scratch.streams[i]);
    if (e != cudaSuccess && rc == DM_OK) rc = static_cast<int>(e);
  }
  return rc;
}
