// Fused orthographic projection: depth (+ value planes) → top-down maps.
// Replaces orth_project, /root/reference/dungeon_maps/maps.py:127-351, including
// scatter_tensor (utils.py:389-492) and the second height scatter (maps.py:335-349).
//
// Design:
//  * No point cloud, no index tensors: one pass over the depth frame computes the cell of
//    every pixel in registers.
//  * The reference reduces every value channel independently by max with ONE shared index
//    (SURVEY.md D2).  So the accumulation buffer is laid out cell-major, channel-minor:
//    acc[cell][CP] of 32-bit order-preserving keys.  A pixel's CU = C(+1 height) updates then
//    hit one or two 128-byte lines instead of CU lines in CU different planes: one RED
//    instruction per pixel-run instead of CU scattered ones.
//  * Channel planes are staged tile by tile in shared memory ([channel][pixel], 128-bit
//    coalesced streaming loads), then re-read transposed: lane = channel, a warp walks its
//    pixels in order and keeps the running max of the current same-cell run in a register;
//    the atomic is issued once per run (neighbouring pixels mostly land in the same cell).
//    Values that cannot change the canvas (v <= fill) are never issued.
//  * acc is a small ring of frame slots that stays resident in the 126 MB L2; the resolve
//    pass decodes keys into the planar (b, C, Mh, Mw) outputs + "changed" masks and zeroes
//    the slot again, so HBM sees only the inputs once and the outputs once.
#include "dm_common.cuh"

namespace dm {

constexpr int kProjThreads = 256;
constexpr int kResolveCells = 256;
constexpr size_t kRingBudgetBytes = 48u << 20;  // accumulation ring kept well inside L2

struct ProjPlan {
  int Cv;     // value channels produced (C, or 1 when the heights are the values)
  int hasH;   // separate height channel accumulated after the values
  int CU;     // Cv + hasH: keys per cell
  int CP;     // cell stride in words (odd → conflict-free transposed smem reads)
  int tile;   // pixels per CTA tile
  int ring;   // frame slots
  int vec;    // 128-bit path usable
  size_t slot_words;
  size_t smem_proj;
  size_t smem_resolve;
};

static ProjPlan make_plan(const DmProjCfg& cfg, int b) {
  ProjPlan p{};
  p.Cv = cfg.C > 0 ? cfg.C : 1;
  p.hasH = (cfg.C > 0 && cfg.want_height) ? 1 : 0;
  p.CU = p.Cv + p.hasH;
  p.CP = (p.CU & 1) ? p.CU : p.CU + 1;
  const size_t M = (size_t)cfg.Mh * cfg.Mw;
  p.slot_words = (M * p.CP + 3) & ~(size_t)3;
  size_t ring = kRingBudgetBytes / (p.slot_words * 4);
  if (ring < 1) ring = 1;
  if (ring > (size_t)b) ring = (size_t)(b > 0 ? b : 1);
  p.ring = (int)ring;
  // staging rows: values + height row; +4 floats keeps rows 16-byte aligned and the
  // transposed LDS.128 conflict-free (row stride ≡ 4 mod 8 words)
  int tile = 1024;
  const int rows = p.CU;
  while (tile > 128 && (size_t)rows * (tile + 4) * 4 + (size_t)tile * 4 > (size_t)100 * 1024) tile >>= 1;
  p.tile = tile;
  p.smem_proj = (size_t)rows * (tile + 4) * 4 + (size_t)tile * 4 + sizeof(DmProjSample);
  p.smem_resolve = (size_t)kResolveCells * p.CP * 4;
  return p;
}

struct ProjDims {
  int Cv, hasH, CU, CP, tile;
  unsigned long long slot_words;
};

// One pixel: validity, cell index (or -1) and the height that goes into the height map.
__device__ __forceinline__ int pixel_cell(const DmProjCfg& cfg, const DmProjSample& sp, int r, int c,
                                          float z, bool ok, float* y_out) {
  if (cfg.has_trunc_depth_max) ok = ok && (z <= cfg.trunc_depth_max);  // maps.py:539-542
  if (cfg.has_trunc_depth_min) ok = ok && (z >= cfg.trunc_depth_min);
  if (cfg.clip_border > 0) {  // maps.py:48-70
    const int k = cfg.clip_border;
    ok = ok && (r >= k) && (r < cfg.H - k) && (c >= k) && (c < cfg.W - k);
  }
  V3 p = unproject(r, c, z, cfg.H, cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.flip_h);
  p = apply_step(sp.to_local, p);                                                  // maps.py:279-284
  if (cfg.has_trunc_height_max) ok = ok && (p.y <= cfg.trunc_height_max);          // maps.py:286-288
  p = apply_step(sp.to_global, p);                                                 // maps.py:290-295
  float xf, zf;
  quantize_f(p.x, p.z, sp.width_offset, sp.height_offset, cfg.map_res, cfg.Mh, cfg.flip_h, &xf, &zf);
  ok = ok && (xf >= 0.0f) && (xf < (float)cfg.Mw) && (zf >= 0.0f) && (zf < (float)cfg.Mh);  // maps.py:1155-1158
  *y_out = p.y;
  return ok ? ((int)zf * cfg.Mw + (int)xf) : -1;  // utils.py:332-370
}

// ---- projection of one tile of one frame ---------------------------------------------------
template <bool VEC>
__device__ __forceinline__ void proj_tile(const float* __restrict__ depth, const float* __restrict__ values,
                                          const uint8_t* __restrict__ valid,
                                          const DmProjSample* __restrict__ samples, const DmProjCfg& cfg,
                                          const ProjDims& d, int frame, int tile_idx,
                                          uint32_t* __restrict__ acc_slot, unsigned char* smem) {
  const int tid = threadIdx.x;
  const int N = cfg.H * cfg.W;
  const int tile = d.tile;
  const int row_stride = tile + 4;
  float* vals = reinterpret_cast<float*>(smem);                      // [CU][tile+4]
  int* cells = reinterpret_cast<int*>(vals + (size_t)d.CU * row_stride);  // [tile]
  DmProjSample* sp_s = reinterpret_cast<DmProjSample*>(cells + tile);

  // per-sample parameters → smem (192 B)
  if (tid < (int)(sizeof(DmProjSample) / 4))
    reinterpret_cast<uint32_t*>(sp_s)[tid] = reinterpret_cast<const uint32_t*>(samples + frame)[tid];
  __syncthreads();
  const DmProjSample& sp = *sp_s;

  const int tile0 = tile_idx * tile;
  const float* dplane = depth + (size_t)frame * N;
  const uint8_t* vplane = valid ? valid + (size_t)frame * N : nullptr;
  const float* vbase = values ? values + (size_t)frame * cfg.C * N : nullptr;
  float* hrow = vals + (size_t)(d.CU - 1) * row_stride;  // C == 0: the only row; hasH: last row

  // ---- phase A: stage value planes, compute cells + heights (thread = 4 consecutive pixels)
  for (int q = tid * 4; q < tile; q += kProjThreads * 4) {
    const int n0 = tile0 + q;
    float z[4];
    bool ok[4];
    if (VEC) {
      if (n0 < N) {
        const float4 z4 = ld_stream_f4(dplane + n0);
        z[0] = z4.x; z[1] = z4.y; z[2] = z4.z; z[3] = z4.w;
        uint32_t vm = 0x01010101u;
        if (vplane) vm = *reinterpret_cast<const uint32_t*>(vplane + n0);
#pragma unroll
        for (int k = 0; k < 4; ++k) ok[k] = ((vm >> (8 * k)) & 0xffu) != 0;
        if (vbase) {
#pragma unroll 8
          for (int ch = 0; ch < cfg.C; ++ch) {
            const float4 v = ld_stream_f4(vbase + (size_t)ch * N + n0);
            *reinterpret_cast<float4*>(vals + (size_t)ch * row_stride + q) = v;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { z[k] = 0.0f; ok[k] = false; }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int n = n0 + k;
        const bool in = n < N;
        z[k] = in ? ld_stream_f1(dplane + n) : 0.0f;
        ok[k] = in && (vplane ? vplane[n] != 0 : true);
      }
      if (vbase) {
        for (int ch = 0; ch < cfg.C; ++ch) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int n = n0 + k;
            vals[(size_t)ch * row_stride + q + k] = n < N ? ld_stream_f1(vbase + (size_t)ch * N + n) : 0.0f;
          }
        }
      }
    }
    int r = n0 / cfg.W;
    int c = n0 - r * cfg.W;
    int cl[4];
    float y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cl[k] = pixel_cell(cfg, sp, r, c, z[k], ok[k], &y[k]);
      if (++c == cfg.W) { c = 0; ++r; }
    }
    *reinterpret_cast<int4*>(cells + q) = make_int4(cl[0], cl[1], cl[2], cl[3]);
    if (d.hasH || cfg.C == 0)
      *reinterpret_cast<float4*>(hrow + q) = make_float4(y[0], y[1], y[2], y[3]);
  }
  __syncthreads();

  // ---- phase B: lane = channel; walk pixels in order, one RED per same-cell run
  const int lane = tid & 31, warp = tid >> 5;
  const int pxw = tile / (kProjThreads / 32);  // pixels per warp, multiple of 4
  const int cu_eff = d.CU < 32 ? d.CU : 32;
  const int streams = d.CU <= 32 ? 32 / d.CU : 1;
  const int passes = d.CU <= 32 ? 1 : (d.CU + 31) / 32;
  const int s = lane / cu_eff;
  const int per = ((pxw / 4 + streams - 1) / streams) * 4;
  int beg = warp * pxw + s * per;
  int end = beg + per;
  if (end > (warp + 1) * pxw) end = (warp + 1) * pxw;
  for (int pass = 0; pass < passes; ++pass) {
    const int c = pass * 32 + (lane - s * cu_eff);
    if (s >= streams || c >= d.CU) continue;
    const bool is_h = d.hasH && (c == d.CU - 1);
    const int is_min = is_h ? 0 : cfg.reduction;
    const float fill = is_h ? -INFINITY : cfg.fill_value;
    const float* row = vals + (size_t)c * row_stride;
    uint32_t* acc_c = acc_slot + c;
    int run_cell = -1;
    float run_v = fill;
    for (int i = beg; i < end; i += 4) {
      const int4 cl4 = *reinterpret_cast<const int4*>(cells + i);
      const float4 v4 = *reinterpret_cast<const float4*>(row + i);
      const int cl[4] = {cl4.x, cl4.y, cl4.z, cl4.w};
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (cl[k] < 0) continue;
        if (cl[k] != run_cell) {
          if (run_cell >= 0 && better(run_v, fill, is_min))
            atomicMax(acc_c + (size_t)run_cell * d.CP, enc_red(run_v, is_min));
          run_cell = cl[k];
          run_v = fill;
        }
        if (better(v[k], run_v, is_min)) run_v = v[k];
      }
    }
    if (run_cell >= 0 && better(run_v, fill, is_min))
      atomicMax(acc_c + (size_t)run_cell * d.CP, enc_red(run_v, is_min));
  }
}

// ---- resolve of kResolveCells cells of one frame -----------------------------------------
__device__ __forceinline__ void resolve_tile(uint32_t* __restrict__ acc_slot, const DmProjCfg& cfg,
                                             const ProjDims& d, int frame, int cell_tile,
                                             float* __restrict__ topdown, uint8_t* __restrict__ mask,
                                             float* __restrict__ height, unsigned char* smem) {
  const int tid = threadIdx.x;
  const int M = cfg.Mh * cfg.Mw;
  const int cell0 = cell_tile * kResolveCells;
  const int ncell = min(kResolveCells, M - cell0);
  const int nw = ncell * d.CP;
  uint32_t* s = reinterpret_cast<uint32_t*>(smem);
  uint32_t* src = acc_slot + (size_t)cell0 * d.CP;  // 16-byte aligned: cell0 % 256 == 0, slot_words % 4 == 0
  for (int i = tid * 4; i < nw; i += kProjThreads * 4) {
    if (i + 3 < nw) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + i));
      if (v.x | v.y | v.z | v.w) __stcg(reinterpret_cast<uint4*>(src + i), make_uint4(0, 0, 0, 0));
      *reinterpret_cast<uint4*>(s + i) = v;
    } else {
      for (int k = i; k < nw; ++k) {
        const uint32_t v = __ldcg(src + k);
        if (v) __stcg(src + k, 0u);
        s[k] = v;
      }
    }
  }
  __syncthreads();
  if (tid < ncell) {
    const int cell = cell0 + tid;
    const uint32_t* mine = s + (size_t)tid * d.CP;
    const size_t obase = (size_t)frame * d.Cv * M + cell;
    for (int c = 0; c < d.Cv; ++c) {
      const uint32_t k = mine[c];
      // utils.py:472-491: canvas starts at fill; a key is only ever stored for a value that
      // beats fill, so "key present" == "cell changed" == mask.
      const float out = k ? dec_red(k, cfg.reduction) : cfg.fill_value;
      st_stream_f1(topdown + obase + (size_t)c * M, out);
      st_stream_u8(mask + obase + (size_t)c * M, k ? 1 : 0);
    }
    if (d.hasH) {
      const uint32_t k = mine[d.Cv];
      st_stream_f1(height + (size_t)frame * M + cell, k ? dec(k) : -INFINITY);  // maps.py:345
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kProjThreads)
proj_kernel(const float* __restrict__ depth, const float* __restrict__ values,
            const uint8_t* __restrict__ valid, const DmProjSample* __restrict__ samples,
            const DmProjCfg cfg, const ProjDims d, int frame0, uint32_t* __restrict__ acc) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int slot = blockIdx.y;
  proj_tile<VEC>(depth, values, valid, samples, cfg, d, frame0 + slot, blockIdx.x,
                 acc + (size_t)slot * d.slot_words, smem);
}

__global__ void __launch_bounds__(kProjThreads)
resolve_kernel(uint32_t* __restrict__ acc, const DmProjCfg cfg, const ProjDims d, int frame0,
               float* __restrict__ topdown, uint8_t* __restrict__ mask, float* __restrict__ height) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int slot = blockIdx.y;
  resolve_tile(acc + (size_t)slot * d.slot_words, cfg, d, frame0 + slot, blockIdx.x, topdown, mask,
               height, smem);
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace dm

using namespace dm;

extern "C" size_t dm_orth_project_workspace_bytes(const DmProjCfg* cfg, int32_t b) {
  if (!cfg || b <= 0 || cfg->Mh <= 0 || cfg->Mw <= 0) return 0;
  const ProjPlan p = make_plan(*cfg, b);
  return p.slot_words * 4 * (size_t)p.ring;
}

extern "C" int dm_orth_project_f32(const float* depth, const float* values, const uint8_t* valid,
                                   const DmProjSample* samples, const DmProjCfg* cfg, int32_t b,
                                   float* topdown, uint8_t* mask, float* height, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  if (!cfg || b < 0) return DM_EINVAL;
  if (b == 0) return DM_OK;
  if (!depth || !samples || !topdown || !mask || !workspace) return DM_EINVAL;
  if (cfg->H <= 0 || cfg->W <= 0 || cfg->Mh <= 0 || cfg->Mw <= 0 || cfg->C < 0) return DM_EINVAL;
  if ((long long)cfg->H * cfg->W >= (1ll << 31) || (long long)cfg->Mh * cfg->Mw >= (1ll << 30)) return DM_EINVAL;
  if (cfg->C > 0 && !values) return DM_EINVAL;
  if (cfg->C > 0 && cfg->want_height && !height) return DM_EINVAL;
  if (cfg->reduction != 0 && cfg->reduction != 1) return DM_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const ProjPlan p = make_plan(*cfg, b);
  if (workspace_bytes < p.slot_words * 4 * (size_t)p.ring) return DM_EWORKSPACE;
  const int N = cfg->H * cfg->W, M = cfg->Mh * cfg->Mw;
  const bool vec = (N % 4 == 0) && aligned(depth, 16) && (!values || aligned(values, 16)) &&
                   (!valid || aligned(valid, 4));
  ProjDims d{p.Cv, p.hasH, p.CU, p.CP, p.tile, (unsigned long long)p.slot_words};
  static bool attr_set = false;
  if (!attr_set) {
    DM_CUDA_OK(cudaFuncSetAttribute(proj_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DM_CUDA_OK(cudaFuncSetAttribute(proj_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DM_CUDA_OK(cudaFuncSetAttribute(resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (p.smem_proj > 200 * 1024 || p.smem_resolve > 200 * 1024) return DM_EINVAL;  // C too large for one pass
  uint32_t* acc = static_cast<uint32_t*>(workspace);
  const int tiles = (N + p.tile - 1) / p.tile;
  const int rtiles = (M + kResolveCells - 1) / kResolveCells;
  for (int f0 = 0; f0 < b; f0 += p.ring) {
    const int nf = (b - f0) < p.ring ? (b - f0) : p.ring;
    dim3 gp(tiles, nf), gr(rtiles, nf);
    if (vec)
      proj_kernel<true><<<gp, kProjThreads, p.smem_proj, stream>>>(depth, values, valid, samples, *cfg, d, f0, acc);
    else
      proj_kernel<false><<<gp, kProjThreads, p.smem_proj, stream>>>(depth, values, valid, samples, *cfg, d, f0, acc);
    DM_LAUNCHED();
    resolve_kernel<<<gr, kProjThreads, p.smem_resolve, stream>>>(acc, *cfg, d, f0, topdown, mask, height);
    DM_LAUNCHED();
  }
  return DM_OK;
}
