// Library-level entry points: ABI introspection, launch accounting and the HOST-buffer
// variant of the projection (host→device copy, kernels, device→host copy as a three-stream
// pipeline over a ring of staging slots, so that both PCIe directions and the kernels overlap).
#include <cstdlib>
#include <mutex>

#include "dm_common.cuh"

namespace dm {
int64_t g_launches = 0;

namespace {
// Pipeline of the HOST-buffer entry: three streams (host→device, kernels, device→host) and a ring of
// kSlots staging slots, chained with events.  The copy engine of each PCIe direction always has the next
// chunk queued (a two-stream A/B scheme left the H2D engine idle while the same stream's D2H ran).
constexpr int kSlots = 4;
struct Scratch {
  int device = -1;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kSlots] = {}, ev_run[kSlots] = {}, ev_out[kSlots] = {};
  void* buf[kSlots] = {};   // staging: inputs + outputs of one chunk
  size_t buf_bytes[kSlots] = {};
  void* ws = nullptr;       // accumulation ring (kept zeroed by the kernel); kernels run on ONE stream
  size_t ws_bytes = 0;
};
Scratch g_scratch;
std::mutex g_mu;

void release_locked() {
  for (int i = 0; i < kSlots; ++i) {
    if (g_scratch.buf[i]) cudaFree(g_scratch.buf[i]);
    if (g_scratch.ev_in[i]) cudaEventDestroy(g_scratch.ev_in[i]);
    if (g_scratch.ev_run[i]) cudaEventDestroy(g_scratch.ev_run[i]);
    if (g_scratch.ev_out[i]) cudaEventDestroy(g_scratch.ev_out[i]);
    g_scratch.buf[i] = nullptr;
    g_scratch.buf_bytes[i] = 0;
    g_scratch.ev_in[i] = g_scratch.ev_run[i] = g_scratch.ev_out[i] = nullptr;
  }
  if (g_scratch.ws) cudaFree(g_scratch.ws);
  g_scratch.ws = nullptr;
  g_scratch.ws_bytes = 0;
  if (g_scratch.s_in) cudaStreamDestroy(g_scratch.s_in);
  if (g_scratch.s_run) cudaStreamDestroy(g_scratch.s_run);
  if (g_scratch.s_out) cudaStreamDestroy(g_scratch.s_out);
  g_scratch.s_in = g_scratch.s_run = g_scratch.s_out = nullptr;
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
    DM_CUDA_OK(cudaStreamCreateWithFlags(&g_scratch.s_in, cudaStreamNonBlocking));
    DM_CUDA_OK(cudaStreamCreateWithFlags(&g_scratch.s_run, cudaStreamNonBlocking));
    DM_CUDA_OK(cudaStreamCreateWithFlags(&g_scratch.s_out, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
      DM_CUDA_OK(cudaEventCreateWithFlags(&g_scratch.ev_in[i], cudaEventDisableTiming));
      DM_CUDA_OK(cudaEventCreateWithFlags(&g_scratch.ev_run[i], cudaEventDisableTiming));
      DM_CUDA_OK(cudaEventCreateWithFlags(&g_scratch.ev_out[i], cudaEventDisableTiming));
    }
  }
  const size_t N = (size_t)cfg->H * cfg->W, M = (size_t)cfg->Mh * cfg->Mw;
  const int Cv = cfg->C > 0 ? cfg->C : 1;
  const bool want_h = cfg->C > 0 && cfg->want_height && height;
  // chunk: ≈ 80 MB of traffic per chunk — large enough for full-rate PCIe copies, small enough that the fill
  // (first chunk in) and drain (last chunk out) of the pipeline stay a few per cent of a 64-frame call
  // (DM_HOST_CHUNK overrides, for experiments)
  const size_t frame_bytes = N * 4 * (1 + (size_t)cfg->C) + (valid ? N : 0) + M * (5 * (size_t)Cv + (want_h ? 4 : 0));
  int chunk = (int)((80u << 20) / (frame_bytes ? frame_bytes : 1));
  if (const char* e = getenv("DM_HOST_CHUNK")) chunk = atoi(e);
  if (chunk < 1) chunk = 1;
  if (chunk > 16) chunk = 16;
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
  for (int i = 0; i < kSlots; ++i) {
    if (g_scratch.buf_bytes[i] < total) {
      if (g_scratch.buf[i]) DM_CUDA_OK(cudaFree(g_scratch.buf[i]));
      g_scratch.buf[i] = nullptr; g_scratch.buf_bytes[i] = 0;
      DM_CUDA_OK(cudaMalloc(&g_scratch.buf[i], total));
      g_scratch.buf_bytes[i] = total;
    }
  }
  if (g_scratch.ws_bytes < ws_need) {
    if (g_scratch.ws) DM_CUDA_OK(cudaFree(g_scratch.ws));
    g_scratch.ws = nullptr; g_scratch.ws_bytes = 0;
    DM_CUDA_OK(cudaMalloc(&g_scratch.ws, ws_need));
    DM_CUDA_OK(cudaMemset(g_scratch.ws, 0, ws_need));
    g_scratch.ws_bytes = ws_need;
  }
  int rc = DM_OK;
  int it = 0;
  for (int f0 = 0; f0 < b && rc == DM_OK; f0 += chunk, ++it) {
    const int nf = (b - f0) < chunk ? (b - f0) : chunk;
    const int k = it % kSlots;
    char* base = static_cast<char*>(g_scratch.buf[k]);
    float* d_depth = reinterpret_cast<float*>(base + o_depth);
    float* d_values = cfg->C > 0 ? reinterpret_cast<float*>(base + o_values) : nullptr;
    uint8_t* d_valid = valid ? reinterpret_cast<uint8_t*>(base + o_valid) : nullptr;
    DmProjSample* d_samples = reinterpret_cast<DmProjSample*>(base + o_samples);
    float* d_top = reinterpret_cast<float*>(base + o_top);
    uint8_t* d_mask = reinterpret_cast<uint8_t*>(base + o_mask);
    float* d_height = want_h ? reinterpret_cast<float*>(base + o_height) : nullptr;
    // host → device (the slot is free once its previous results have left)
    cudaStream_t si = g_scratch.s_in;
    if (it >= kSlots) DM_CUDA_OK(cudaStreamWaitEvent(si, g_scratch.ev_out[k], 0));
    DM_CUDA_OK(cudaMemcpyAsync(d_depth, depth + (size_t)f0 * N, (size_t)nf * N * 4, cudaMemcpyHostToDevice, si));
    if (d_values)
      DM_CUDA_OK(cudaMemcpyAsync(d_values, values + (size_t)f0 * cfg->C * N, (size_t)nf * cfg->C * N * 4,
                                 cudaMemcpyHostToDevice, si));
    if (d_valid)
      DM_CUDA_OK(cudaMemcpyAsync(d_valid, valid + (size_t)f0 * N, (size_t)nf * N, cudaMemcpyHostToDevice, si));
    DM_CUDA_OK(cudaMemcpyAsync(d_samples, samples + f0, (size_t)nf * sizeof(DmProjSample), cudaMemcpyHostToDevice, si));
    DM_CUDA_OK(cudaEventRecord(g_scratch.ev_in[k], si));
    // kernels
    cudaStream_t sr = g_scratch.s_run;
    DM_CUDA_OK(cudaStreamWaitEvent(sr, g_scratch.ev_in[k], 0));
    DmProjCfg c = *cfg;
    if (!want_h) c.want_height = 0;
    rc = dm_orth_project_f32(d_depth, d_values, d_valid, d_samples, &c, nf, d_top, d_mask, d_height,
                             g_scratch.ws, g_scratch.ws_bytes, sr);
    if (rc != DM_OK) break;
    DM_CUDA_OK(cudaEventRecord(g_scratch.ev_run[k], sr));
    // device → host
    cudaStream_t so = g_scratch.s_out;
    DM_CUDA_OK(cudaStreamWaitEvent(so, g_scratch.ev_run[k], 0));
    DM_CUDA_OK(cudaMemcpyAsync(topdown + (size_t)f0 * Cv * M, d_top, (size_t)nf * Cv * M * 4, cudaMemcpyDeviceToHost, so));
    DM_CUDA_OK(cudaMemcpyAsync(mask + (size_t)f0 * Cv * M, d_mask, (size_t)nf * Cv * M, cudaMemcpyDeviceToHost, so));
    if (d_height)
      DM_CUDA_OK(cudaMemcpyAsync(height + (size_t)f0 * M, d_height, (size_t)nf * M * 4, cudaMemcpyDeviceToHost, so));
    DM_CUDA_OK(cudaEventRecord(g_scratch.ev_out[k], so));
  }
  for (cudaStream_t st : {g_scratch.s_in, g_scratch.s_run, g_scratch.s_out}) {
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && rc == DM_OK) rc = static_cast<int>(e);
  }
  return rc;
}
