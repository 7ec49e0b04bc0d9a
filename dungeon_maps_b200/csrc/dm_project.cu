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
//  * Channel planes are staged tile by tile in shared memory ([channel][pixel]) by TMA bulk
//    copies (cp.async.bulk + mbarrier, L2 evict-first), then re-read transposed: lane = channel,
//    one RED per run of pixels that share a cell (neighbouring pixels mostly do).  Values that
//    cannot change the canvas (v <= fill) are never issued.
//  * ONE persistent launch per call: CTAs pull (frame, tile) work items from a ticket counter.
//    Projection tiles of frame f + lag are interleaved with resolve tiles of frame f, which decode
//    the keys into the planar (b, C, Mh, Mw) outputs + "changed" masks and zero the slot again.
//  * acc is a ring of up to 10 frame slots used SPARSELY: a flag per 64-cell slice says whether any
//    key of the slice was touched; the resolve pass only reads (and re-zeroes) flagged slices, so
//    the lines of a slot that no pixel hit are never brought on chip.  What is resident in the
//    126 MB L2 is the touched part of the ring (a few MB per slot), HBM sees the inputs once and
//    the outputs once, and the lag between projecting and resolving a frame can be several frames
//    — which is what keeps the dependency waits of a deep, many-tiles-in-flight schedule rare.
//  * Shapes whose planes are not 16-byte aligned take the plain-load kernels (proj_kernel /
//    resolve_kernel, chunked over the ring) — same device functions, same results.
#include <cstdio>
#include <cstdlib>

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

#include "dm_project.cuh"

namespace dm {

constexpr int kProjThreads = 256;
constexpr int kResolveCells = 256;
#ifndef DM_MAX_RING
#define DM_MAX_RING 10  // measured at config 2 (room / iid ms): 6: 0.573 / 0.656, 8: 0.496 / 0.586, 10: 0.469 / 0.563, 14: 0.476 / 0.579, 20: 0.480 / 0.614
#endif
constexpr int kMaxRing = DM_MAX_RING;                     // frame slots of the accumulation ring
constexpr size_t kRingBudgetBytes = (size_t)1 << 30;  // ... unless that exceeds 1 GiB of workspace

struct ProjPlan {
  int Cv;     // value channels produced (C, or 1 when the heights are the values)
  int hasH;   // separate height channel accumulated after the values
  int CU;     // Cv + hasH: keys per cell
  int CP;     // cell stride in words (odd → conflict-free transposed smem reads)
  int rows;   // staged rows per tile: C value planes + the depth/height row
  int tile;   // pixels per CTA tile
  int ring;   // frame slots
  int lag;    // resolve(f) is scheduled with proj(f + lag); ring > lag
  int nsl;    // slices (flags) per slot
  size_t slot_words;
  size_t ctrl_bytes;
  size_t flag_bytes;
  size_t stage_bytes;
  size_t smem_tile;      // one stage + sample block (plain-load kernel)
  size_t smem_resolve;
  int ws_warps;          // warp-specialised kernel: consumer warps per CTA (tile = 128 px each)
  int ws_groups;         // ... channel groups a frame's value planes are staged in (each group re-reads the depth row)
  int ws_cg;             // ... value channels per group
  int lean;              // C == 0: hmap_proj_kernel + hmap_resolve_kernel instead of the persistent kernel
  int ws_r2d;            // ... image rows per 2-D tile (tensor-map TMA), 0: tiles of 128 * ws_warps consecutive pixels
  size_t ws_stage_bytes; // warp-specialised kernel: one stage of 128 * ws_warps pixels
  size_t smem_ws;        // stage + barriers/item/sample block
  size_t workspace_bytes() const { return ctrl_bytes + flag_bytes + slot_words * 4 * (size_t)ring; }
};

// Tile rows of the warp-specialised kernel: -1 automatic, 0 forces the 1-D row tiles, 4 / 8 force 2-D tiles of that
// many image rows (test hook dm_debug_set_tile_rows; both layouts give identical results).
static int g_tile_rows = -1;
#ifndef DM_R2D_DEFAULT
#define DM_R2D_DEFAULT 0  // measured (r02m-r02p): row tiles 0.455 ms, 4-row 2-D tiles 0.509 ms, 8-row 0.603 ms per config-2 step
#endif

static ProjPlan make_plan(const DmProjCfg& cfg, int b, int tile_rows = -1) {
  ProjPlan p{};
  p.Cv = cfg.C > 0 ? cfg.C : 1;
  p.hasH = (cfg.C > 0 && cfg.want_height) ? 1 : 0;
  p.CU = p.Cv + p.hasH;
  p.CP = (p.CU & 1) ? p.CU : p.CU + 1;
  p.rows = cfg.C + 1;
  const size_t M = (size_t)cfg.Mh * cfg.Mw;
  p.slot_words = (M * p.CP + 3) & ~(size_t)3;
  size_t ring = kRingBudgetBytes / (p.slot_words * 4);
  // height maps only (C == 0) take the two plain launches of the height-map path: a slot per frame, up to 64
  p.lean = cfg.C == 0 && g_tile_rows != -3;
  if (ring > (size_t)(p.lean ? 64 : kMaxRing)) ring = p.lean ? 64 : kMaxRing;
  if (ring > (size_t)(b > 0 ? b : 1)) ring = (size_t)(b > 0 ? b : 1);
  if (ring < 2) ring = 2;
  p.ring = (int)ring;
  p.lag = p.ring / 2;
  p.nsl = (int)((M + kSliceCells - 1) / kSliceCells);
  p.ctrl_bytes = ((size_t)(kCtrlWords + 2 * (b > 0 ? b : 1)) * 4 + 255) & ~(size_t)255;
  p.flag_bytes = ((size_t)p.ring * p.nsl * kFlagStride * 4 + 255) & ~(size_t)255;
  // staging rows: +4 floats keeps rows 16-byte aligned and the transposed LDS.128
  // conflict-free (row stride ≡ 4 mod 8 words)
  int tile = 1024;
  auto stage = [&](int t) { return (size_t)p.rows * (t + 4) * 4 + (size_t)t * 4; };
  while (tile > 128 && stage(tile) > (size_t)40 * 1024) tile >>= 1;
  p.tile = tile;
  p.stage_bytes = (stage(tile) + 127) & ~(size_t)127;
  p.smem_resolve = (size_t)kResolveCells * p.CP * 4;
  size_t st = p.stage_bytes > p.smem_resolve ? p.stage_bytes : ((p.smem_resolve + 127) & ~(size_t)127);
  p.stage_bytes = st;
  p.smem_tile = st + 256;
  // warp-specialised kernel: 4 consumer warps (512-pixel tiles) while 4 CTAs still fit an SM, else 2 (256-pixel tiles)
  auto ws_stage = [&](int ww, int rows) {
    size_t ws = ((size_t)rows * (128 * ww + 4) * 4 + (size_t)128 * ww * 4 + 127) & ~(size_t)127;
    const size_t ws_res = ((size_t)ww * 64 * p.CP * 4 + 127) & ~(size_t)127;  // ww warps x 64 cells x CP words
    return ws < ws_res ? ws_res : ws;
  };
  auto fits4 = [&](int ww, int rows) { return (ws_stage(ww, rows) + 256 + 1024) * 4 <= (size_t)228 * 1024; };
  // many value channels: stage them in groups of at most 24 planes (each group a tile of its own that re-reads the
  // depth row from L2 and repeats phase A), 2 consumer warps per CTA.  Measured at C = 40, 1280x720, ms per 32 frames
  // (groups x consumer warps): 1x4 2.26, 1x2 1.71, 2x4 1.46, **2x2 1.38**, 3x4 1.54, 3x2 1.59, 4x2 1.89; at C = 16
  // grouping only costs (2x4: 0.62 ms instead of 0.50).
  p.ws_warps = 4;
  p.ws_groups = 1;
  if (!fits4(4, cfg.C + 1)) {
    p.ws_groups = (cfg.C + 23) / 24;
    p.ws_warps = 2;
  }
#ifdef DM_EXPERIMENT_KNOBS  // kernel experiments only (scripts/exp_build.sh builds with -DDM_EXPERIMENT_KNOBS)
  if (const char* e = getenv("DM_WS_WARPS")) {
    const int ww = atoi(e);
    if (ww == 2 || ww == 4) p.ws_warps = ww;
  }
  if (const char* e = getenv("DM_WS_GROUPS")) {
    const int g = atoi(e);
    if (g >= 1 && g <= 8) p.ws_groups = g;
  }
#endif
  if (cfg.C <= 0) p.ws_groups = 1;
  p.ws_cg = cfg.C > 0 ? (cfg.C + p.ws_groups - 1) / p.ws_groups : 0;
  p.ws_groups = cfg.C > 0 ? (cfg.C + p.ws_cg - 1) / p.ws_cg : 1;
  p.ws_r2d = tile_rows >= 0 ? tile_rows : g_tile_rows >= 0 ? g_tile_rows : DM_R2D_DEFAULT;
#ifdef DM_EXPERIMENT_KNOBS
  if (const char* e = getenv("DM_R2D")) if (tile_rows < 0) p.ws_r2d = atoi(e);
#endif
  if (p.ws_r2d != 0 && p.ws_r2d != 4 && p.ws_r2d != 8) p.ws_r2d = 0;
  // 8-row tiles: 8 x 64 pixels (2 consumer warps) keep the stage of a 4 x 128 tile; only without channel groups
  if (p.ws_r2d == 8) {
    if (p.ws_groups == 1 && p.ws_warps == 4) p.ws_warps = 2; else p.ws_r2d = 4;
  }
  if (p.ws_r2d) {
    // (cg + 1) dense planes of R x 32 ww words + the runlet list; a tensor-map box dimension is at most 256
    const size_t pw = (size_t)p.ws_r2d * 32 * p.ws_warps * 4;
    const size_t ws = (size_t)(p.ws_cg + 2) * pw;
    const size_t ws_res = ((size_t)p.ws_warps * 64 * p.CP * 4 + 127) & ~(size_t)127;
    p.ws_stage_bytes = ws < ws_res ? ws_res : ws;
    if (p.ws_cg > 256) p.ws_r2d = 0;
  }
  if (!p.ws_r2d) p.ws_stage_bytes = ws_stage(p.ws_warps, p.ws_cg + 1);
  p.smem_ws = p.ws_stage_bytes + 256;
  return p;
}

struct ProjDims {
  int Cv, hasH, CU, CP, rows, tile, ring, lag, nsl;
  unsigned long long slot_words;
  unsigned long long stage_bytes;
  int groups = 1, cg = 0;  // warp-specialised kernel: channel groups per frame, channels per group
  unsigned long long ws_words = 0;  // 32-bit words of the workspace behind the control block (flags + ring)
};


// ---- phase A: cells + heights of one staged tile ------------------------------------------
// The depth row of the stage is overwritten in place by the heights (it becomes the height
// channel of phase B).  thread = 4 consecutive pixels.
__device__ __forceinline__ void phase_a(const uint8_t* __restrict__ vplane, const DmProjSample& sp,
                                        const DmProjCfg& cfg, int tile, int tile0, float* zrow, int* cells) {
  const int N = cfg.H * cfg.W;
  for (int q = threadIdx.x * 4; q < tile; q += kProjThreads * 4) {
    const int n0 = tile0 + q;
    int cl[4] = {-1, -1, -1, -1};
    float y[4] = {0.f, 0.f, 0.f, 0.f};
    if (n0 < N) {
      const float4 z4 = *reinterpret_cast<const float4*>(zrow + q);
      const float z[4] = {z4.x, z4.y, z4.z, z4.w};
      int r = n0 / cfg.W;
      int c = n0 - r * cfg.W;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = (n0 + k < N) && (vplane ? vplane[n0 + k] != 0 : true);
        cl[k] = pixel_cell(cfg, sp, r, c, z[k], ok, &y[k]);
        if (++c == cfg.W) { c = 0; ++r; }
      }
    }
    *reinterpret_cast<int4*>(cells + q) = make_int4(cl[0], cl[1], cl[2], cl[3]);
    *reinterpret_cast<float4*>(zrow + q) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

// ---- phase B: lane = channel; walk pixels in order, one RED per same-cell run ---------------
template <bool IS_MIN>
__device__ __forceinline__ void emit_runs(const int* __restrict__ cells, const float* __restrict__ row,
                                          int beg, int end, float fill, uint32_t* __restrict__ acc_c, int CP) {
  int run_cell = -1;
  float run_v = fill;
  auto flush = [&]() {
    if (run_cell >= 0 && (IS_MIN ? (run_v < fill) : (run_v > fill)))
      atomicMax(acc_c + (size_t)run_cell * CP, IS_MIN ? ~enc(run_v) : enc(run_v));
  };
  auto visit = [&](int cl, float v) {
    if (cl >= 0) {
      if (cl != run_cell) {
        flush();
        run_cell = cl;
        run_v = fill;
      }
      // fmaxf/fminf drop a NaN operand: torch_scatter's `src > out` never lets NaN win either
      run_v = IS_MIN ? fminf(run_v, v) : fmaxf(run_v, v);
    }
  };
  for (int i = beg; i < end; i += 4) {
    const int4 cl4 = *reinterpret_cast<const int4*>(cells + i);
    const float4 v4 = *reinterpret_cast<const float4*>(row + i);
    visit(cl4.x, v4.x);
    visit(cl4.y, v4.y);
    visit(cl4.z, v4.z);
    visit(cl4.w, v4.w);
  }
  flush();
}

__device__ __forceinline__ void phase_b(const DmProjCfg& cfg, const ProjDims& d, const float* vals,
                                        const int* cells, uint32_t* __restrict__ acc_slot) {
  const int tile = d.tile;
  const int row_stride = tile + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pxw = tile / (kProjThreads / 32);  // pixels per warp, multiple of 4
  const int cu_eff = d.CU < 32 ? d.CU : 32;
  const int streams = d.CU <= 32 ? 32 / d.CU : 1;
  const int passes = d.CU <= 32 ? 1 : (d.CU + 31) / 32;
  const int s = lane / cu_eff;
  const int per = ((pxw / 4 + streams - 1) / streams) * 4;
  const int beg = warp * pxw + s * per;
  int end = beg + per;
  if (end > (warp + 1) * pxw) end = (warp + 1) * pxw;
  for (int pass = 0; pass < passes; ++pass) {
    const int c = pass * 32 + (lane - s * cu_eff);
    if (s >= streams || c >= d.CU) continue;
    // channel c of the cell: value planes first, the height channel last.  Its staged row:
    // value plane c, or the depth/height row (index rows-1) for the height channel and for C == 0.
    const bool is_h = (c == d.Cv);  // only when hasH
    const bool from_hrow = is_h || cfg.C == 0;
    const float* row = vals + (size_t)(from_hrow ? d.rows - 1 : c) * row_stride;
    if (!is_h && cfg.reduction)
      emit_runs<true>(cells, row, beg, end, cfg.fill_value, acc_slot + c, d.CP);
    else
      emit_runs<false>(cells, row, beg, end, is_h ? -INFINITY : cfg.fill_value, acc_slot + c, d.CP);
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

// ================= plain-load kernels (any shape / alignment) ================================

__global__ void __launch_bounds__(kProjThreads)
proj_kernel(const float* __restrict__ depth, const float* __restrict__ values,
            const uint8_t* __restrict__ valid, const DmProjSample* __restrict__ samples,
            const DmProjCfg cfg, const ProjDims d, int frame0, uint32_t* __restrict__ acc) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int slot = blockIdx.y, frame = frame0 + slot;
  const int N = cfg.H * cfg.W;
  const int tile = d.tile, row_stride = tile + 4;
  float* vals = reinterpret_cast<float*>(smem);
  int* cells = reinterpret_cast<int*>(vals + (size_t)d.rows * row_stride);
  DmProjSample* sp_s = reinterpret_cast<DmProjSample*>(smem + d.stage_bytes);
  if (tid < (int)(sizeof(DmProjSample) / 4))
    reinterpret_cast<uint32_t*>(sp_s)[tid] = reinterpret_cast<const uint32_t*>(samples + frame)[tid];
  const int tile0 = blockIdx.x * tile;
  const float* dplane = depth + (size_t)frame * N;
  const float* vbase = values ? values + (size_t)frame * cfg.C * N : nullptr;
  // stage rows with scalar loads (no alignment assumptions)
  for (int q = tid; q < tile; q += kProjThreads) {
    const int n = tile0 + q;
    const bool in = n < N;
    for (int ch = 0; ch < cfg.C; ++ch)
      vals[(size_t)ch * row_stride + q] = in ? ld_stream_f1(vbase + (size_t)ch * N + n) : 0.0f;
    vals[(size_t)cfg.C * row_stride + q] = in ? ld_stream_f1(dplane + n) : 0.0f;
  }
  __syncthreads();
  phase_a(valid ? valid + (size_t)frame * N : nullptr, *sp_s, cfg, tile, tile0, vals + (size_t)cfg.C * row_stride, cells);
  __syncthreads();
  phase_b(cfg, d, vals, cells, acc + (size_t)slot * d.slot_words);
}

__global__ void __launch_bounds__(kProjThreads)
resolve_kernel(uint32_t* __restrict__ acc, const DmProjCfg cfg, const ProjDims d, int frame0,
               float* __restrict__ topdown, uint8_t* __restrict__ mask, float* __restrict__ height) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int slot = blockIdx.y;
  resolve_tile(acc + (size_t)slot * d.slot_words, cfg, d, frame0 + slot, blockIdx.x, topdown, mask,
               height, smem);
}

// ================= persistent TMA kernel =======================================================


// ---- warp-specialised persistent kernel -----------------------------------------------------
// CTA = 1 producer warp + kWsWarps consumer warps, two smem stages.
//   producer : claims tickets, waits for the item's dependency (ring slot free / frame fully
//              projected), fills the stage by TMA bulk copies (one row per lane) and publishes the
//              per-frame completion counters once the consumers have released a stage.
//   consumer : owns 128 pixels (32 lanes x 4) of the 512-pixel tile, or 64 cells of a resolve
//              tile, and does everything warp-locally (no CTA-wide barrier anywhere):
//       A  cells + heights of its 4 pixels, in-thread merge of equal neighbouring cells into
//          "runlets", warp prefix sum → compacted position of every runlet;
//       B1 lane = pixel quad, loop over channels at full lane utilisation: LDS.128, 3 predicated
//          max, store the runlet maxima compacted in place (row c of the stage);
//       B2 lane = channel, loop over runlets only: one RED per runlet covers all channels of the
//          cell (1-2 cache lines); the height channel goes lane = runlet.
#ifdef DM_PROFILE
// build-time instrumentation (scripts/exp.py): cycle counters summed over CTAs in ctrl[16..63]
#define DM_CLK() clock64()
#define DM_ACC(ctrl_, i_, v_) atomicAdd(reinterpret_cast<unsigned long long*>((ctrl_) + 320) + (i_), (unsigned long long)(v_))
#else
#define DM_CLK() 0ll
#define DM_ACC(ctrl_, i_, v_) ((void)0)
#endif
// WW = consumer warps per CTA (template parameter of the kernel): 4 by default; 2, together with channel groups
// (make_plan), when the C + 1 staged rows of a 512-pixel tile would leave fewer than 4 CTAs per SM (many value
// channels) — consumer warps per SM are what sets the throughput.  tile = 128 * WW pixels, resolve tile = 64 * WW cells.
// Measured on the B200 (scripts/exp_ww.sh, ms per 64-frame step, room / iid scene | config 5 shapes, 32 frames):
//   C = 16: WW 8: 0.511 / 0.582   6: 0.511 / 0.579   4: 0.499 / 0.585   3: 0.541 / 0.632   2: 0.551 / 0.682   1: 0.694 / 1.059
//   C = 40: WW 8: 2.32            6: 3.01            4: 2.26            3: 2.04            2: 1.87            1: 2.07
constexpr int kWsMaxWarps = 4;
#ifndef DM_RES_K
#define DM_RES_K 2  // 64-cell slices a consumer warp resolves per resolve ticket (measured: 1: 0.494, 2: 0.470, 3: 0.473, 4: 0.485 ms)
#endif
constexpr int kResK = DM_RES_K;

struct WsItem {
  int kind, frame, idx, tile0, r0, c0, ok, _pad;
};



// FAST: 0 generic steps, 1 local only, 2 local + global.  IS_MIN: reduction of the value channels.
template <int FAST, bool IS_MIN, int WW>
__device__ __forceinline__ void ws_proj_slice(const DmProjCfg& cfg, const ProjDims& d, const WsItem& it,
                                              const DmProjSample& sp, const Rcps& rcp,
                                              const uint8_t* __restrict__ vplane, float* vals, int* lcell,
                                              uint32_t* __restrict__ acc, uint32_t slot_off,
                                              uint32_t* __restrict__ slot_flags, int cw, int lane,
                                              uint64_t* full_vals, uint32_t phase, int ch0, int nch,
                                              long long* tprof) {
  // acc is the kernel parameter (uniform); every RED address is acc + a 32-bit word offset
  [[maybe_unused]] const long long tp0 = DM_CLK();
  constexpr int RS = 128 * WW + 4;
  const int N = cfg.H * cfg.W;
  const int sb = cw * 128;
  const int q = sb + 4 * lane;
  const int n0 = it.tile0 + q;
  float* zrow = vals + nch * RS;  // the depth row follows the group's value rows
  int cl[4] = {-1, -1, -1, -1};
  float y[4] = {0.f, 0.f, 0.f, 0.f};
  // ---- A: cells and heights of my 4 pixels
  if (n0 < N) {
    const float4 z4 = *reinterpret_cast<const float4*>(zrow + q);
    const float z[4] = {z4.x, z4.y, z4.z, z4.w};
    int c = it.c0 + q, r = it.r0;
    while (c >= cfg.W) { c -= cfg.W; ++r; }
    uint32_t vm = 0x01010101u;
    if (vplane) vm = *reinterpret_cast<const uint32_t*>(vplane + n0);
    if (FAST) {
      // maps.py:677-678 column / row factors rn(rn(c - cx) / fx), rn(rn(yy - cy) / fy): exact division
      // through the reciprocals (div_by_rcp), no table in shared memory
      float xn[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xn[k] = div_by_rcp(__fsub_rn((float)(c + k), cfg.cx), cfg.fx, rcp.fx);
      const float yy = cfg.flip_h ? __fsub_rn((float)(cfg.H - 1), (float)r) : (float)r;
      const float yn = div_by_rcp(__fsub_rn(yy, cfg.cy), cfg.fy, rcp.fy);
      bool rowok = true;
      const int kb = cfg.clip_border;
      if (kb > 0) rowok = (r >= kb) && (r < cfg.H - kb);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bool ok = rowok && (((vm >> (8 * k)) & 0xffu) != 0);
        if (kb > 0) ok = ok && (c + k >= kb) && (c + k < cfg.W - kb);
        cl[k] = pixel_cell_fast<FAST == 2, true>(cfg, sp, xn[k], yn, z[k], ok, &y[k], rcp.res);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        cl[k] = pixel_cell(cfg, sp, r, c + k, z[k], ((vm >> (8 * k)) & 0xffu) != 0, &y[k]);
    }
  }
  // in-thread runs: pixel k continues into k+1 when both are valid and share the cell
  const bool p01 = (cl[0] >= 0) && (cl[0] == cl[1]);
  const bool p12 = (cl[1] >= 0) && (cl[1] == cl[2]);
  const bool p23 = (cl[2] >= 0) && (cl[2] == cl[3]);
  const bool t0 = (cl[0] >= 0) && !p01, t1 = (cl[1] >= 0) && !p12, t2 = (cl[2] >= 0) && !p23, t3 = cl[3] >= 0;
  const int cnt = (int)t0 + (int)t1 + (int)t2 + (int)t3;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int padn = (4 - (total & 3)) & 3;  // B2 reads the list in quads; the stale tail of the last one is masked there
  // compacted positions of my runlets (word offsets inside a row); dead pixels park on a scratch
  // word past the slice so that stores need no predicate juggling
  const int o0 = sb + incl - cnt, o1 = o0 + (int)t0, o2 = o1 + (int)t1, o3 = o2 + (int)t2;
  // the list holds word offsets of the cells in the accumulation slot (cell * CP)
  if (t0) lcell[o0] = cl[0] * d.CP;
  if (t1) lcell[o1] = cl[1] * d.CP;
  if (t2) lcell[o2] = cl[2] * d.CP;
  if (t3) lcell[o3] = cl[3] * d.CP;
  {  // sparse ring: flag the 64-cell slices my runlets touch (one store per change of slice, not per runlet)
    const int s0 = cl[0] >> 6, s1 = cl[1] >> 6, s2 = cl[2] >> 6, s3 = cl[3] >> 6;
    const int lastv = t3 ? s3 : t2 ? s2 : t1 ? s1 : t0 ? s0 : -1;
    int prev = __shfl_up_sync(0xffffffffu, lastv, 1);
    if (lane == 0) prev = -1;
    if (t0 && s0 != prev) st_flag(slot_flags + s0 * kFlagStride);
    prev = t0 ? s0 : prev;
    if (t1 && s1 != prev) st_flag(slot_flags + s1 * kFlagStride);
    prev = t1 ? s1 : prev;
    if (t2 && s2 != prev) st_flag(slot_flags + s2 * kFlagStride);
    prev = t2 ? s2 : prev;
    if (t3 && s3 != prev) st_flag(slot_flags + s3 * kFlagStride);
  }
  {  // the depth row becomes the (compacted) height row; C == 0: it is the value channel itself
    const bool hmin = IS_MIN && cfg.C == 0;
    y[1] = p01 ? (hmin ? fminf(y[0], y[1]) : fmaxf(y[0], y[1])) : y[1];
    y[2] = p12 ? (hmin ? fminf(y[1], y[2]) : fmaxf(y[1], y[2])) : y[2];
    y[3] = p23 ? (hmin ? fminf(y[2], y[3]) : fmaxf(y[2], y[3])) : y[3];
    __syncwarp();  // every lane has read its depths
    if (t0) zrow[o0] = y[0];
    if (t1) zrow[o1] = y[1];
    if (t2) zrow[o2] = y[2];
    if (t3) zrow[o3] = y[3];
  }
  // phase A needed the depth row only; the value rows (their own barrier) had this long to arrive
  mbar_wait(full_vals, phase);
#ifdef DM_ABL_NOB  // ablation build (wrong results): load pipeline + phase A + resolve only
  return;
#endif
  [[maybe_unused]] const long long tp1 = DM_CLK();
  // ---- B1: channel loop at full lane utilisation, compaction in place
#define DM_B1_ROW(ROW)                                                     \
  {                                                                        \
    float* row_ = (ROW);                                                   \
    float4 a = *reinterpret_cast<const float4*>(row_ + q);                 \
    a.y = p01 ? red2<IS_MIN>(a.x, a.y) : a.y;                              \
    a.z = p12 ? red2<IS_MIN>(a.y, a.z) : a.z;                              \
    a.w = p23 ? red2<IS_MIN>(a.z, a.w) : a.w;                              \
    __syncwarp();                                                          \
    if (t0) row_[o0] = a.x;                                                \
    if (t1) row_[o1] = a.y;                                                \
    if (t2) row_[o2] = a.z;                                                \
    if (t3) row_[o3] = a.w;                                                \
  }
  {
    int c = 0;
    for (; c + 4 <= nch; c += 4) {
      float* r0 = vals + c * RS;
      float4 a0 = *reinterpret_cast<const float4*>(r0 + q);
      float4 a1 = *reinterpret_cast<const float4*>(r0 + RS + q);
      float4 a2 = *reinterpret_cast<const float4*>(r0 + 2 * RS + q);
      float4 a3 = *reinterpret_cast<const float4*>(r0 + 3 * RS + q);
      a0.y = p01 ? red2<IS_MIN>(a0.x, a0.y) : a0.y; a1.y = p01 ? red2<IS_MIN>(a1.x, a1.y) : a1.y;
      a2.y = p01 ? red2<IS_MIN>(a2.x, a2.y) : a2.y; a3.y = p01 ? red2<IS_MIN>(a3.x, a3.y) : a3.y;
      a0.z = p12 ? red2<IS_MIN>(a0.y, a0.z) : a0.z; a1.z = p12 ? red2<IS_MIN>(a1.y, a1.z) : a1.z;
      a2.z = p12 ? red2<IS_MIN>(a2.y, a2.z) : a2.z; a3.z = p12 ? red2<IS_MIN>(a3.y, a3.z) : a3.z;
      a0.w = p23 ? red2<IS_MIN>(a0.z, a0.w) : a0.w; a1.w = p23 ? red2<IS_MIN>(a1.z, a1.w) : a1.w;
      a2.w = p23 ? red2<IS_MIN>(a2.z, a2.w) : a2.w; a3.w = p23 ? red2<IS_MIN>(a3.z, a3.w) : a3.w;
      __syncwarp();  // loads of these rows are done before any lane compacts into them
      if (t0) { r0[o0] = a0.x; r0[RS + o0] = a1.x; r0[2 * RS + o0] = a2.x; r0[3 * RS + o0] = a3.x; }
      if (t1) { r0[o1] = a0.y; r0[RS + o1] = a1.y; r0[2 * RS + o1] = a2.y; r0[3 * RS + o1] = a3.y; }
      if (t2) { r0[o2] = a0.z; r0[RS + o2] = a1.z; r0[2 * RS + o2] = a2.z; r0[3 * RS + o2] = a3.z; }
      if (t3) { r0[o3] = a0.w; r0[RS + o3] = a1.w; r0[2 * RS + o3] = a2.w; r0[3 * RS + o3] = a3.w; }
    }
    for (; c < nch; ++c) DM_B1_ROW(vals + c * RS)
  }
#undef DM_B1_ROW
  __syncwarp();
  [[maybe_unused]] const long long tp2 = DM_CLK();
  // ---- B2: one RED per (runlet, channel); lane = channel keeps a runlet's keys in 1-2 lines
  const int total4 = total + padn;
  if (cfg.C > 0) {
    const int Cv = nch;  // this group's channels; their keys start at ch0
    const int cu_eff = Cv < 32 ? Cv : 32;
    const int streams = Cv <= 32 ? 32 / Cv : 1;
    const int passes = Cv <= 32 ? 1 : (Cv + 31) / 32;
    const int s = lane / cu_eff;
    const int per = ((total4 / 4 + streams - 1) / streams) * 4;
    const int beg = s * per;
    const int end = min(beg + per, total4);
    for (int pass = 0; pass < passes; ++pass) {
      const int c = pass * 32 + (lane - s * cu_eff);
      if (s >= streams || c >= Cv) continue;
      const float* row = vals + c * RS + sb;
      const int* lc = lcell + sb;
      const uint32_t off_c = slot_off + (uint32_t)(ch0 + c);
      const float fill = cfg.fill_value;
      // runs longer than a pixel quad arrive as neighbouring runlets of one cell: they are folded and the RED goes
      // out when the cell changes; (pc, pv) is the runlet that has not been issued yet, carried across the quads of
      // the list (folding inside a quad only: 70 k REDs per room frame, carried: 56 k; 0.470 -> 0.457 ms)
#ifdef DM_B2_BRANCHFREE
      if (beg >= end) continue;
      uint32_t pc = (uint32_t)lc[beg];
#else
      uint32_t pc = 0xffffffffu;
#endif
      float pv = fill;
      for (int i = beg; i < end; i += 4) {
        uint4 c4 = *reinterpret_cast<const uint4*>(lc + i);
        float4 v4 = *reinterpret_cast<const float4*>(row + i);
        if (i + 4 > total) {  // the last quad of the list is partly stale: its tail continues the last entry with `fill`
          const int n = total - i;
          if (n < 2) { c4.y = c4.x; v4.y = fill; }
          if (n < 3) { c4.z = c4.y; v4.z = fill; }
          c4.w = c4.z; v4.w = fill;
        }
        const bool mp = pc == c4.x, m01 = c4.x == c4.y, m12 = c4.y == c4.z, m23 = c4.z == c4.w;
        red_key_if_run_ends<IS_MIN>(pc, c4.x, pv, fill, acc + (off_c + pc));
        v4.x = mp ? red2<IS_MIN>(pv, v4.x) : v4.x;
        v4.y = m01 ? red2<IS_MIN>(v4.x, v4.y) : v4.y;
        v4.z = m12 ? red2<IS_MIN>(v4.y, v4.z) : v4.z;
        v4.w = m23 ? red2<IS_MIN>(v4.z, v4.w) : v4.w;
        red_key_if_run_ends<IS_MIN>(c4.x, c4.y, v4.x, fill, acc + (off_c + c4.x));
        red_key_if_run_ends<IS_MIN>(c4.y, c4.z, v4.y, fill, acc + (off_c + c4.y));
        red_key_if_run_ends<IS_MIN>(c4.z, c4.w, v4.z, fill, acc + (off_c + c4.z));
        pc = c4.w;
        pv = v4.w;
      }
      red_key_if_run_ends<IS_MIN>(pc, 0xfffffffeu, pv, fill, acc + (off_c + pc));
    }
  }
  // lane = runlet for the heights.  Neighbouring runlets of one cell are folded by a segmented scan over the
  // lanes and only the last one of a run issues its RED.
  if ((d.hasH || cfg.C == 0) && ch0 == 0) {  // once per tile: the first channel group
    const bool hmin = IS_MIN && cfg.C == 0;  // C == 0: the heights are the values (fill / reduction apply)
    const float hfill = cfg.C == 0 ? cfg.fill_value : -INFINITY;  // height channel: max against -inf (maps.py:340-348)
    const uint32_t hoff = slot_off + (cfg.C == 0 ? 0u : (uint32_t)d.Cv);
    for (int base = 0; base < total; base += 32) {
      const int i = base + lane;
      const bool active = i < total;
      const uint32_t cellv = active ? (uint32_t)lcell[sb + i] : 0xffffffffu;
      float v = active ? zrow[sb + i] : hfill;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float pv = __shfl_up_sync(0xffffffffu, v, o);
        const uint32_t pc = __shfl_up_sync(0xffffffffu, cellv, o);
        if (lane >= o && pc == cellv) v = hmin ? fminf(v, pv) : fmaxf(v, pv);
      }
      const uint32_t nc = __shfl_down_sync(0xffffffffu, cellv, 1);
      const bool last = lane == 31 || nc != cellv;
      const bool win = hmin ? (v < hfill) : (v > hfill);
      if (active && last && win) red_max_u32(acc + (hoff + cellv), hmin ? ~enc(v) : enc(v));
    }
  }
#ifdef DM_PROFILE
  [[maybe_unused]] const long long tp3 = DM_CLK();
  tprof[0] += tp1 - tp0; tprof[1] += tp2 - tp1; tprof[2] += tp3 - tp2; tprof[3] += total;
#endif
}

// ---- 2-D tiles (round 2): R image rows x 32 WW columns, staged by ONE tensor-map TMA copy per tile ------------
// A thread owns the R rows of ONE image column.  Vertical structure (walls, furniture) projects onto one map cell,
// so runs fold in-thread down the column first and then across the lanes along each tile row: the runlet list holds
// the runlets that END in tile row 0 (in lane = column order), then those of row 1, ... ("row streams"), neighbours
// in the list that share the cell fold in B2.  Host model of the room frames (same folding rules): 65.5 k RED groups
// per frame with 128 x 1 pixels per warp, 29.8 k with 32 x 4, 22.8 k with 32 x 8.
// Stage layout: plane p (value channel of the group, then the depth plane at index d.cg) = R segments (tile rows) of
// TW = 32 WW words, dense — exactly what cp.async.bulk.tensor.3d writes for a box (TW, R, planes).  Warp cw owns words
// [32 cw, 32 cw + 32) of every segment, for its pixels and — in place — for its compacted runlet list (entry i at
// segment i / 32, word i % 32).  Dense planes are a multiple of 32 words apart, which would make B2's transposed
// reads (lane = channel, same list position) 8-way bank conflicts; the compacted entries of value plane c are
// therefore XOR-swizzled by 4 (c & 3) words inside their 32-word segment (B1 writes, B2 reads them so).
template <int FAST, bool IS_MIN, int WW, int R>
__device__ __forceinline__ void ws_proj_slice2d(const DmProjCfg& cfg, const ProjDims& d, const WsItem& it,
                                                const DmProjSample& sp, const Rcps& rcp,
                                                const uint8_t* __restrict__ vplane, float* vals, int* lcell,
                                                uint32_t* __restrict__ acc, uint32_t slot_off,
                                                uint32_t* __restrict__ slot_flags, int cw, int lane,
                                                uint64_t* full_vals, uint32_t phase, int ch0, int nch,
                                                long long* tprof) {
  [[maybe_unused]] const long long tp0 = DM_CLK();
  constexpr int TW = 32 * WW;
  constexpr int PW = R * TW;  // words per staged plane
  const int colt = (cw << 5) + lane;  // my column inside the tile
  const int wbase = cw << 5;
  // word offset (inside a staged plane) of entry i of this warp's runlet list
  auto pos = [&](int i) { return (i >> 5) * TW + wbase + (i & 31); };
  float* zrow = vals + d.cg * PW;  // the depth plane follows the group's value planes (fixed place: the box is cg planes)
  int cl[R];
  float y[R];
  // ---- A: cells and heights of my R pixels
  {
    const int c = it.c0 + colt;
    const bool colok = c < cfg.W;
    float z[R];
#pragma unroll
    for (int k = 0; k < R; ++k) z[k] = zrow[k * TW + colt];  // rows / columns beyond the image: zero-filled by the TMA
    const int kb = cfg.clip_border;
    bool ok[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int r = it.r0 + k;
      ok[k] = colok && r < cfg.H;
      if (kb > 0) ok[k] = ok[k] && (r >= kb) && (r < cfg.H - kb) && (c >= kb) && (c < cfg.W - kb);
      if (vplane && ok[k]) ok[k] = vplane[(size_t)r * cfg.W + c] != 0;
    }
    if (FAST) {
      const float xn = div_by_rcp(__fsub_rn((float)c, cfg.cx), cfg.fx, rcp.fx);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int r = it.r0 + k;
        const float yy = cfg.flip_h ? __fsub_rn((float)(cfg.H - 1), (float)r) : (float)r;
        const float yn = div_by_rcp(__fsub_rn(yy, cfg.cy), cfg.fy, rcp.fy);
        cl[k] = pixel_cell_fast<FAST == 2, true>(cfg, sp, xn, yn, z[k], ok[k], &y[k], rcp.res);
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        y[k] = 0.f;
        cl[k] = ok[k] ? pixel_cell(cfg, sp, it.r0 + k, c, z[k], true, &y[k]) : -1;
      }
    }
  }
  // in-thread runs down the column: pixel k continues into k + 1 when both are valid and share the cell
  bool p[R], t[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    p[k] = (k + 1 < R) && (cl[k] >= 0) && (cl[k] == cl[k + 1 < R ? k + 1 : k]);
    t[k] = (cl[k] >= 0) && !p[k];
  }
  // compacted positions: row streams
  int o[R];
  int total = 0;
  {
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const unsigned bk = __ballot_sync(0xffffffffu, t[k]);
      o[k] = pos(total + __popc(bk & lt));
      total += __popc(bk);
    }
  }
  const int padn = (4 - (total & 3)) & 3;  // B2 reads the list in quads; the stale tail of the last one is masked there
  // the list holds word offsets of the cells in the accumulation slot (cell * CP)
#pragma unroll
  for (int k = 0; k < R; ++k)
    if (t[k]) lcell[o[k]] = cl[k] * d.CP;
  {  // sparse ring flags: one store per change of slice down my column and against my left neighbour's row
    int prev = -1;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int sk = cl[k] >> 6;
      const int lk = __shfl_up_sync(0xffffffffu, t[k] ? sk : -1, 1);
      if (t[k] && sk != prev && (lane == 0 || sk != lk)) st_flag(slot_flags + sk * kFlagStride);
      prev = t[k] ? sk : prev;
    }
  }
  {  // the depth plane becomes the (compacted) height list; C == 0: it is the value channel itself
    const bool hmin = IS_MIN && cfg.C == 0;
#pragma unroll
    for (int k = 0; k + 1 < R; ++k) y[k + 1] = p[k] ? (hmin ? fminf(y[k], y[k + 1]) : fmaxf(y[k], y[k + 1])) : y[k + 1];
    __syncwarp();  // every lane has read its depths
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (t[k]) zrow[o[k]] = y[k];
  }
  // phase A needed the depth plane only; the value planes (their own barrier) had this long to arrive
  mbar_wait(full_vals, phase);
#ifdef DM_ABL_NOB  // ablation build (wrong results): load pipeline + phase A + resolve only
  return;
#endif
  [[maybe_unused]] const long long tp1 = DM_CLK();
  // ---- B1: channel loop at full lane utilisation, vertical fold, compaction in place.  The entries of plane c are
  // stored XOR-swizzled by 4 (c & 3) words: the planes of one pass (c, c + 4, c + 8, c + 12) share the swizzle, so it
  // costs one LOP per stored row instead of address arithmetic per store, and B2's transposed 128-bit reads are at
  // worst 2-way conflicts
  {
    constexpr int CB = R <= 4 ? 4 : 2;  // planes in flight per pass
#pragma unroll 1
    for (int s4 = 0; s4 < 4; ++s4) {
      int os[R];
#pragma unroll
      for (int k = 0; k < R; ++k) os[k] = o[k] ^ (4 * s4);
#pragma unroll 1
      for (int c = s4; c < nch; c += 4 * CB) {
        float* pl = vals + c * PW;
        float a[CB][R];
#pragma unroll
        for (int u = 0; u < CB; ++u)
          if (c + 4 * u < nch) {
#pragma unroll
            for (int k = 0; k < R; ++k) a[u][k] = pl[4 * u * PW + k * TW + colt];
          }
#pragma unroll
        for (int u = 0; u < CB; ++u)
#pragma unroll
          for (int k = 0; k + 1 < R; ++k) a[u][k + 1] = p[k] ? red2<IS_MIN>(a[u][k], a[u][k + 1]) : a[u][k + 1];
        __syncwarp();  // loads of these planes are done before any lane compacts into them
#pragma unroll
        for (int u = 0; u < CB; ++u)
          if (c + 4 * u < nch) {
#pragma unroll
            for (int k = 0; k < R; ++k)
              if (t[k]) pl[4 * u * PW + os[k]] = a[u][k];
          }
      }
    }
  }
  __syncwarp();
  [[maybe_unused]] const long long tp2 = DM_CLK();
  // ---- B2: one RED per (run, channel); lane = channel keeps a run's keys in 1-2 lines
  const int total4 = total + padn;
  if (cfg.C > 0) {
    const int Cv = nch;  // this group's channels; their keys start at ch0
    const int cu_eff = Cv < 32 ? Cv : 32;
    const int streams = Cv <= 32 ? 32 / Cv : 1;
    const int passes = Cv <= 32 ? 1 : (Cv + 31) / 32;
    const int s = lane / cu_eff;
    const int per = ((total4 / 4 + streams - 1) / streams) * 4;
    const int beg = s * per;
    const int end = min(beg + per, total4);
    for (int pass = 0; pass < passes; ++pass) {
      const int c = pass * 32 + (lane - s * cu_eff);
      if (s >= streams || c >= Cv) continue;
      const float* pl = vals + c * PW;
      const int swz = 4 * (c & 3);
      const uint32_t off_c = slot_off + (uint32_t)(ch0 + c);
      const float fill = cfg.fill_value;
      // neighbouring runlets of one cell (the same cell along a tile row) are folded, the RED goes out when the cell
      // changes: (pc, pv) is the run that has not been issued yet
      uint32_t pc = 0xffffffffu;
      float pv = fill;
      for (int i = beg; i < end; i += 4) {
        const int pi = pos(i);
        uint4 c4 = *reinterpret_cast<const uint4*>(lcell + pi);
        float4 v4 = *reinterpret_cast<const float4*>(pl + (pi ^ swz));
        if (i + 4 > total) {  // the last quad of the list is partly stale: its tail continues the last entry with `fill`
          const int n = total - i;
          if (n < 2) { c4.y = c4.x; v4.y = fill; }
          if (n < 3) { c4.z = c4.y; v4.z = fill; }
          c4.w = c4.z; v4.w = fill;
        }
        const bool mp = pc == c4.x, m01 = c4.x == c4.y, m12 = c4.y == c4.z, m23 = c4.z == c4.w;
        red_key_if_run_ends<IS_MIN>(pc, c4.x, pv, fill, acc + (off_c + pc));
        v4.x = mp ? red2<IS_MIN>(pv, v4.x) : v4.x;
        v4.y = m01 ? red2<IS_MIN>(v4.x, v4.y) : v4.y;
        v4.z = m12 ? red2<IS_MIN>(v4.y, v4.z) : v4.z;
        v4.w = m23 ? red2<IS_MIN>(v4.z, v4.w) : v4.w;
        red_key_if_run_ends<IS_MIN>(c4.x, c4.y, v4.x, fill, acc + (off_c + c4.x));
        red_key_if_run_ends<IS_MIN>(c4.y, c4.z, v4.y, fill, acc + (off_c + c4.y));
        red_key_if_run_ends<IS_MIN>(c4.z, c4.w, v4.z, fill, acc + (off_c + c4.z));
        pc = c4.w;
        pv = v4.w;
      }
      red_key_if_run_ends<IS_MIN>(pc, 0xfffffffeu, pv, fill, acc + (off_c + pc));
    }
  }
  // lane = runlet for the heights: segmented scan over the lanes, the last runlet of a run issues
  if ((d.hasH || cfg.C == 0) && ch0 == 0) {  // once per tile: the first channel group
    const bool hmin = IS_MIN && cfg.C == 0;
    const float hfill = cfg.C == 0 ? cfg.fill_value : -INFINITY;  // maps.py:340-348
    const uint32_t hoff = slot_off + (cfg.C == 0 ? 0u : (uint32_t)d.Cv);
    for (int base = 0; base < total; base += 32) {
      const int i = base + lane;
      const bool active = i < total;
      const int pi = pos(i);
      const uint32_t cellv = active ? (uint32_t)lcell[pi] : 0xffffffffu;
      float v = active ? zrow[pi] : hfill;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const float pv = __shfl_up_sync(0xffffffffu, v, o2);
        const uint32_t pc = __shfl_up_sync(0xffffffffu, cellv, o2);
        if (lane >= o2 && pc == cellv) v = hmin ? fminf(v, pv) : fmaxf(v, pv);
      }
      const uint32_t nc = __shfl_down_sync(0xffffffffu, cellv, 1);
      const bool last = lane == 31 || nc != cellv;
      const bool win = hmin ? (v < hfill) : (v > hfill);
      if (active && last && win) red_max_u32(acc + (hoff + cellv), hmin ? ~enc(v) : enc(v));
    }
  }
#ifdef DM_PROFILE
  [[maybe_unused]] const long long tp3 = DM_CLK();
  tprof[0] += tp1 - tp0; tprof[1] += tp2 - tp1; tprof[2] += tp3 - tp2; tprof[3] += total;
#endif
}

// 64 cells of a resolve tile, warp-local: load the 64 x CP keys (coalesced 128-bit, all loads in
// flight before the first use), zero what was set, and write the planar outputs.  Most of a map
// is empty: a slice without a single key takes a constant-store path.
__device__ __forceinline__ bool ws_resolve_slice(uint32_t* __restrict__ acc_slot, const DmProjCfg& cfg,
                                                 const ProjDims& d, uint32_t* __restrict__ slot_flags, int frame,
                                                 int cell_tile, int tile_cells, int cw, int lane, uint32_t* wres,
                                                 float* __restrict__ topdown,
                                                 uint8_t* __restrict__ mask, float* __restrict__ height) {
  const int M = cfg.Mh * cfg.Mw;
  const int cell0 = cell_tile * tile_cells + cw * 64;
  const int ncell = min(64, M - cell0);
  if (ncell <= 0) return false;
  uint32_t* slice_flag = slot_flags + (size_t)(cell0 >> 6) * kFlagStride;
  const int nw = ncell * d.CP;
  const int nw4 = nw & ~3;
  uint32_t* src = acc_slot + (size_t)cell0 * d.CP;  // 16-byte aligned: cell0 % 64 == 0 → words % 4 == 0
  uint32_t any = 0;
  // sparse ring: a slice nobody flagged holds no key and is not even read
  uint32_t flagged = 0;
  if (lane == 0) {
    flagged = __ldcg(slice_flag);
    if (flagged) __stcg(slice_flag, 0u);
  }
  flagged = __shfl_sync(0xffffffffu, flagged, 0);
  if (!flagged) {
    if (!(ncell == 64 && (M & 3) == 0))
      for (int k = lane; k < nw; k += 32) wres[k] = 0;
  } else {
  for (int base = 0; base < nw4; base += 512) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + (u * 32 + lane) * 4;
      v[u] = i < nw4 ? __ldcg(reinterpret_cast<const uint4*>(src + i)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + (u * 32 + lane) * 4;
      const uint32_t nz = v[u].x | v[u].y | v[u].z | v[u].w;
      any |= nz;
      if (i < nw4) {
        if (nz) __stcg(reinterpret_cast<uint4*>(src + i), make_uint4(0, 0, 0, 0));
        *reinterpret_cast<uint4*>(wres + i) = v[u];
      }
    }
  }
  for (int k = nw4 + lane; k < nw; k += 32) {
    const uint32_t v = __ldcg(src + k);
    if (v) __stcg(src + k, 0u);
    wres[k] = v;
    any |= v;
  }
  }
  const bool occupied = __any_sync(0xffffffffu, any != 0);
  const size_t plane0 = (size_t)frame * d.Cv * M + cell0;
  if (!occupied && ncell == 64 && (M & 3) == 0) {
    // empty slice: 64 x fill per channel, 64 x False per channel.  Every store instruction is a full warp of 16-byte
    // stores — two channels of values (512 B), eight channels of masks when the mask planes are 16-byte aligned —
    // because what a global store costs on the SM is the instruction, not its bytes (r02q ablations): 11 store
    // instructions per slice at C = 16 instead of 33.
    const float f = cfg.fill_value;
    const float4 f4 = make_float4(f, f, f, f);
    for (int c = lane >> 4; c < d.Cv; c += 2) st_stream_f4(topdown + plane0 + (size_t)c * M + (lane & 15) * 4, f4);
    if ((M & 15) == 0) {
      for (int c = lane >> 2; c < d.Cv; c += 8) st_stream_u4(mask + plane0 + (size_t)c * M + (lane & 3) * 16, 0u);
    } else {
      for (int c = lane >> 4; c < d.Cv; c += 2) st_stream_u32(mask + plane0 + (size_t)c * M + (lane & 15) * 4, 0u);
    }
    if (d.hasH && lane < 16)
      st_stream_f4(height + (size_t)frame * M + cell0 + lane * 4, make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY));
    return flagged != 0;
  }
  __syncwarp();
  for (int j = lane; j < ncell; j += 32) {
    const uint32_t* mine = wres + j * d.CP;
    float* tp = topdown + plane0 + j;
    uint8_t* mp = mask + plane0 + j;
    for (int c = 0; c < d.Cv; ++c) {
      const uint32_t k = mine[c];
      const float out = k ? dec_red(k, cfg.reduction) : cfg.fill_value;  // utils.py:472-491
      st_stream_f1(tp, out);
      st_stream_u8(mp, k ? 1 : 0);
      tp += M;
      mp += M;
    }
    if (d.hasH) {
      const uint32_t k = mine[d.Cv];
      st_stream_f1(height + (size_t)frame * M + cell0 + j, k ? dec(k) : -INFINITY);  // maps.py:345
    }
  }
  __syncwarp();
  return flagged != 0;
}

// R2D = 0: tiles of 128 WW consecutive pixels, one bulk copy per plane row (any W % 4 == 0 layout).
// R2D = 4 / 8: tiles of R2D image rows x 32 WW columns, one tensor-map copy (cp.async.bulk.tensor.3d, UTMALDG) for
// the group's value planes and one for the depth plane; tmv / tmd are the tensor maps of `values` (W, H, b C) and
// `depth` (W, H, b), encoded on the host per call.
template <int FAST, bool IS_MIN, int WW, int R2D>
// (the 4-row layout is held to 6 CTAs per SM, the shared-memory limit: left alone ptxas takes 88 registers for it and
// 4 CTAs fit — 0.509 instead of 0.471 ms per config-2 step)
#ifndef DM_WS_MINB
#define DM_WS_MINB 0  // 0: no occupancy target, ptxas's own heuristics (a target of 1 lets it take 80+ registers: 4 CTAs per SM, 0.537 ms)
#endif
__global__ void __launch_bounds__(32 * (WW + 1), (R2D == 4 && WW == 4) ? 6 : DM_WS_MINB)
proj_ws_kernel(const float* __restrict__ depth, const float* __restrict__ values,
               const uint8_t* __restrict__ valid, const DmProjSample* __restrict__ samples,
               const DmProjCfg cfg, const ProjDims d, int b, uint32_t* __restrict__ ctrl,
               uint32_t* __restrict__ flags, uint32_t* __restrict__ acc, float* __restrict__ topdown,
               uint8_t* __restrict__ mask, float* __restrict__ height, const ProjGuard guard,
               const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmd) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kWsWarps = WW, kWsTile = 128 * WW, kWsResolveCells = 64 * WW * kResK;
  constexpr int RS = kWsTile + 4;
  constexpr int TW = 32 * WW;                    // 2-D tiles: columns per tile
  constexpr int PW = (R2D ? R2D : 1) * TW;       // ... words per staged plane
  const Rcps rcp{__frcp_rn(cfg.map_res), __frcp_rn(cfg.fx), __frcp_rn(cfg.fy)};
  // one stage per CTA: latency is hidden by the 5-6 CTAs resident per SM, not by an in-CTA ring
  unsigned char* stage = smem;
  unsigned char* tail = smem + d.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);        // item + depth row of a projection tile
  uint64_t* empty = full + 1;
  uint64_t* full_vals = reinterpret_cast<uint64_t*>(tail + 24);  // value rows of a projection tile
  WsItem* item = reinterpret_cast<WsItem*>(tail + 32);
  DmProjSample* sps = reinterpret_cast<DmProjSample*>(tail + 64);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = cfg.H * cfg.W, M = cfg.Mh * cfg.Mw;
  const int CBk = (cfg.W + TW - 1) / TW;        // 2-D tiles: column blocks per row group
  // projection tickets per frame: tiles x channel groups
  const int P = (R2D ? ((cfg.H + R2D - 1) / (R2D ? R2D : 1)) * CBk : (N + kWsTile - 1) / kWsTile) * d.groups;
  const int R = (M + kWsResolveCells - 1) / kWsResolveCells;
  uint32_t* proj_done = ctrl + kCtrlWords;
  uint32_t* resolve_done = proj_done + b;
  uint32_t* occ_count = reinterpret_cast<uint32_t*>(tail + 16);  // flagged slices seen by this CTA

  // How far the resolve pass trails the projection (in frames) is chosen from how densely the previous call
  // on this workspace filled its maps (ctrl[4], written by that call's last CTA; 0: unknown, assume sparse):
  // the touched part of lag + 2 slots should fit in ~48 MB of L2.  A long lag makes dependency waits rare and
  // lets the producers claim two tickets ahead; densely hit maps fall back to the short schedule.
  int lag = d.lag;
  {
    const uint32_t hint = __ldcg(ctrl + 4);
    if (hint) {
      const float touched = (float)(hint - 1u) * (1.0f / 65536.0f) * (float)d.slot_words * 4.0f;
      const int fit = (int)fminf((float)(48u << 20) / fmaxf(touched, 1.0f), 64.0f) - 2;
      lag = max(min(2, d.lag), min(fit, d.lag));
    }
  }
  const int ahead = lag >= 3 ? 2 : 1;
  // slots in use: a slot that is re-used soon keeps its lines in L2 (a densely hit ring must stay small)
  const int ring = min(d.ring, 2 * lag);
  const unsigned total = (unsigned)(b + lag) * (unsigned)(P + R);

  if (tid == 0) {
    *occ_count = 0;
    mbar_init(full, 1);
    mbar_init(full_vals, 1);
    mbar_init(empty, kWsWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kWsWarps) {
    // ===================== producer =====================
    const uint64_t policy = policy_evict_first();
    int pend_kind = kItemNone, pend_frame = 0;
    uint32_t fills = 0;
    [[maybe_unused]] long long pp[4] = {0, 0, 0, 0};
    // tickets are claimed two ahead so that the round trip of the atomic is never waited for
    unsigned raw0 = 0, raw1 = 0;
    if (lane == 0) {
      raw0 = atomicAdd(ctrl, 1u);
      if (ahead == 2) raw1 = atomicAdd(ctrl, 1u);
    }
    while (true) {
      [[maybe_unused]] const long long tq0 = DM_CLK();
      // ---- while the consumers work on the previous item: decode, read the dependency,
      //      fetch the per-sample parameters (all off the critical path)
      const unsigned t = __shfl_sync(0xffffffffu, raw0, 0);
      if (lane == 0) {
        if (ahead == 2) {
          raw0 = raw1;
          if (t < total) raw1 = atomicAdd(ctrl, 1u);
        } else if (t < total) {
          raw0 = atomicAdd(ctrl, 1u);
        }
      }
      WsItem it{};
      it.ok = 1;
      if (t >= total) {
        it.kind = kItemExit;
      } else {
        decode_ticket(t, b, P, R, lag, &it.kind, &it.frame, &it.idx);
      }
      // dependency: ring slot resolved by its previous tenant / frame fully projected
      const uint32_t* dep = nullptr;
      uint32_t dep_target = 0;
      uint32_t spw0 = 0, spw1 = 0;  // my two words of the sample block
      if (it.kind == kItemProj) {
        const int grp = it.idx % d.groups;  // neighbouring tickets share the tile: its depth row is re-read from L2
        const int ch0 = grp * d.cg;
        it._pad = ch0 | (min(d.cg, cfg.C - ch0) << 16);
        if (R2D) {
          const int ti = it.idx / d.groups, rg = ti / CBk;
          it.r0 = R2D * rg;
          it.c0 = (ti - rg * CBk) * TW;
          it.tile0 = it.r0 * cfg.W + it.c0;
        } else {
          it.tile0 = (it.idx / d.groups) * kWsTile;
          it.r0 = it.tile0 / cfg.W;
          it.c0 = it.tile0 - it.r0 * cfg.W;
        }
        if (it.frame >= ring) { dep = resolve_done + (it.frame - ring); dep_target = (uint32_t)R + guard.dep_bias; }
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(samples + it.frame);
        spw0 = sw[lane];
        if (lane < 16) spw1 = sw[32 + lane];
      } else if (it.kind == kItemResolve) {
        dep = proj_done + it.frame;
        dep_target = (uint32_t)P + guard.dep_bias;
      }
      // acquire load (pairs with the red.release of publish_prev in the CTA that completed the frame), issued here and
      // evaluated after the wait below, so its round trip is never waited for.  The consumers inherit the ordering
      // through the mbarrier hand-off of the item (release.cta by this warp, acquire.cta by theirs).
      uint32_t dep_seen = 0;
      if (dep && lane == 0) dep_seen = ld_acquire(dep);
      // ---- the stage is free once the consumers released the previous item
      [[maybe_unused]] const long long tq1 = DM_CLK();
      mbar_wait(empty, (fills & 1u) ^ 1u);
      [[maybe_unused]] const long long tq2 = DM_CLK();
      int pending = 0;
      if (dep && lane == 0) pending = dep_seen < dep_target;
      pending = __shfl_sync(0xffffffffu, pending, 0);
      const int prev_kind = pend_kind, prev_frame = pend_frame;
      auto publish_prev = [&]() {  // the release of the stage also completes the previous item
        if (lane == 0 && prev_kind != kItemNone)
          red_release_add1((prev_kind == kItemProj ? proj_done : resolve_done) + prev_frame);
      };
      if (pending) {
        // rare: must block.  Publish first — the frame we wait for may need this very tile.
        publish_prev();
        if (lane == 0) it.ok = wait_count(dep, dep_target, ctrl, guard.spin_ns);
        it.ok = __shfl_sync(0xffffffffu, it.ok, 0);
      }
      [[maybe_unused]] const long long tq3 = DM_CLK();
      if (it.kind == kItemProj) {
        reinterpret_cast<uint32_t*>(sps)[lane] = spw0;
        if (lane < 16) reinterpret_cast<uint32_t*>(sps)[32 + lane] = spw1;
      }
      if (lane == 0) *item = it;
      __syncwarp();
      if (R2D && it.kind == kItemProj) {
        // one tensor-map copy for the depth plane (completes `full`: phase A starts) and one for the group's d.cg value
        // planes; rows / columns / planes beyond the tensor are zero-filled and count towards the transaction bytes
        const int ch0 = it._pad & 0xffff, nch = it._pad >> 16;
        if (lane == 0) {
          float* vals = reinterpret_cast<float*>(stage);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(full, (uint32_t)PW * 4u);
          if (nch > 0) mbar_expect_tx(full_vals, (uint32_t)PW * 4u * (uint32_t)d.cg); else mbar_arrive(full_vals);
          tensor_g2s_3d(vals + d.cg * PW, &tmd, it.c0, it.r0, it.frame, full, policy);
          if (nch > 0) tensor_g2s_3d(vals, &tmv, it.c0, it.r0, it.frame * cfg.C + ch0, full_vals, policy);
        }
      } else if (it.kind == kItemProj) {
        const int npx = min(kWsTile, N - it.tile0);
        const uint32_t row_bytes = (uint32_t)npx * 4u;
        const int ch0 = it._pad & 0xffff, nch = it._pad >> 16;
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(full, row_bytes);
          if (nch > 0) mbar_expect_tx(full_vals, row_bytes * (uint32_t)nch); else mbar_arrive(full_vals);
        }
        __syncwarp();
        float* vals = reinterpret_cast<float*>(stage);
        // the depth row goes first and completes `full` on its own: the consumers run phase A (cells, heights,
        // runlets) while the value rows are still in flight
        if (lane == 0) bulk_g2s(vals + nch * RS, depth + (size_t)it.frame * N + it.tile0, row_bytes, full, policy);
        for (int row = lane; row < nch; row += 32)
          bulk_g2s(vals + row * RS, values + ((size_t)it.frame * cfg.C + ch0 + row) * N + it.tile0, row_bytes, full_vals,
                   policy);
      } else {
        if (lane == 0) { mbar_arrive(full); mbar_arrive(full_vals); }
      }
      if (!pending) publish_prev();  // off the critical path: the next item is already on its way
      pend_kind = (it.kind == kItemProj || it.kind == kItemResolve) ? it.kind : kItemNone;
      pend_frame = it.frame;
      ++fills;
#ifdef DM_PROFILE
      pp[0] += tq1 - tq0; pp[1] += tq2 - tq1; pp[2] += tq3 - tq2; pp[3] += DM_CLK() - tq3;
#endif
      if (it.kind == kItemExit) break;
    }
#ifdef DM_PROFILE
    if (lane == 0) { DM_ACC(ctrl, 8, pp[0]); DM_ACC(ctrl, 9, pp[1]); DM_ACC(ctrl, 10, pp[2]); DM_ACC(ctrl, 11, pp[3]); }
#endif
    // the consumers have added their flagged-slice counts; the last CTA out re-arms the control block for
    // the next call and leaves it the density of this call's maps
    mbar_wait(empty, (fills & 1u) ^ 1u);
    uint32_t last = 0;
    if (lane == 0) {
      atomicAdd(ctrl + 5, *reinterpret_cast<volatile uint32_t*>(occ_count));
      __threadfence();
      last = atomicAdd(ctrl + 3, 1u) == gridDim.x - 1 ? 1u : 0u;
      if (last) __threadfence();
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {  // every other CTA has left: a timed-out launch re-zeroes its workspace and tells the host
      scrub_after_timeout(ctrl, flags, d.ws_words, lane, 32, guard.status);
      __syncwarp();
    }
    if (lane == 0) {
      if (last) {
        const unsigned long long flagged = atomicAdd(ctrl + 5, 0u);
        const unsigned long long slices = (unsigned long long)b * (unsigned long long)d.nsl;
        const unsigned long long f16 = slices ? (flagged << 16) / slices : 0ull;
        ctrl[4] = 1u + (uint32_t)(f16 < 65535ull ? f16 : 65535ull);
        ctrl[5] = 0;
        for (int i = 0; i < 2 * b; ++i) proj_done[i] = 0;
        ctrl[0] = 0; ctrl[1] = 0; ctrl[2] = 0; ctrl[3] = 0;
        __threadfence();
      }
    }
  } else {
    // ===================== consumers =====================
    uint32_t uses = 0, my_flagged = 0;
    [[maybe_unused]] long long cp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (true) {
      [[maybe_unused]] const long long tc0 = DM_CLK();
      mbar_wait(full, uses & 1u);
      ++uses;
      const WsItem it = *item;
      [[maybe_unused]] const long long tc1 = DM_CLK();
      if (it.kind == kItemExit) {
        if (lane == 0) {
          atomicAdd(occ_count, my_flagged);
          mbar_arrive(empty);
        }
        break;
      }
      if (it.ok) {
        if (it.kind == kItemProj) {
          float* vals = reinterpret_cast<float*>(stage);
          const int slot = it.frame % ring;
          if constexpr (R2D != 0) {
            int* lcell = reinterpret_cast<int*>(vals + d.rows * PW);
            ws_proj_slice2d<FAST, IS_MIN, WW, (R2D ? R2D : 4)>(
                cfg, d, it, *sps, rcp, valid ? valid + (size_t)it.frame * N : nullptr, vals, lcell, acc,
                (uint32_t)slot * (uint32_t)d.slot_words, flags + (size_t)slot * d.nsl * kFlagStride, warp, lane,
                full_vals, (uses - 1u) & 1u, it._pad & 0xffff, it._pad >> 16, cp);
          } else {
            int* lcell = reinterpret_cast<int*>(vals + d.rows * RS);
            ws_proj_slice<FAST, IS_MIN, WW>(cfg, d, it, *sps, rcp, valid ? valid + (size_t)it.frame * N : nullptr, vals,
                                        lcell, acc, (uint32_t)slot * (uint32_t)d.slot_words,
                                        flags + (size_t)slot * d.nsl * kFlagStride, warp, lane, full_vals,
                                        (uses - 1u) & 1u, it._pad & 0xffff, it._pad >> 16, cp);
          }
#ifdef DM_PROFILE
          cp[4] += tc1 - tc0; cp[7] += 1;
#endif
        } else if (it.kind == kItemResolve) {
          uint32_t* wres = reinterpret_cast<uint32_t*>(stage) + warp * 64 * d.CP;
          const int slot = it.frame % ring;
#pragma unroll 1
          for (int k = 0; k < kResK; ++k)
            my_flagged += ws_resolve_slice(acc + (size_t)slot * d.slot_words, cfg, d,
                                           flags + (size_t)slot * d.nsl * kFlagStride, it.frame, it.idx * kResK + k,
                                           64 * WW, warp, lane, wres, topdown, mask, height) ? 1u : 0u;
#ifdef DM_PROFILE
          cp[5] += tc1 - tc0; cp[6] += DM_CLK() - tc1;
#endif
        }
      } else if (it.kind == kItemProj) {
        // skipped after a dependency timeout: the tile's copies are in flight all the same, and the stage (and the
        // phase of its barrier) may only be handed back once they have landed
        mbar_wait(full_vals, (uses - 1u) & 1u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty);
    }
#ifdef DM_PROFILE
    if (warp == 0 && lane == 0)
      for (int i = 0; i < 8; ++i) DM_ACC(ctrl, i, cp[i]);
#endif
  }
}


// ================= height maps only (C == 0): two plain launches ================================================
// MapBuilder.plot and every caller that asks for a height map project depth alone: 5.2 MB of input and 0.8 MB of output
// per 480 x 640 frame.  The persistent kernel's ticket machinery (913 tickets per frame, cross-CTA dependencies) costs
// more than that work: 100 us per 32 frames where the traffic is worth 10 us (ncu r02h).  Here every frame of the call
// has a key plane of its own (acc: nf x M words, zero between calls), so there is nothing to wait for:
//   hmap_proj_kernel    a thread = 4 consecutive pixels (one 128-bit load): cells + heights with the device functions
//                       of the persistent kernel (same bits), equal neighbours folded in-thread, one RED.MAX per run;
//   hmap_resolve_kernel a thread = 4 cells: key -> value / mask (utils.py:472-491), the plane is zero again after.
template <int FAST, bool IS_MIN>
__global__ void __launch_bounds__(256)
hmap_proj_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ valid,
                 const DmProjSample* __restrict__ samples, const DmProjCfg cfg, int frame0, int vec,
                 uint32_t* __restrict__ acc, unsigned long long slot_words) {
  __shared__ DmProjSample sp;
  const int slot = blockIdx.y, frame = frame0 + slot;
  const int N = cfg.H * cfg.W;
  if (threadIdx.x < (int)(sizeof(DmProjSample) / 4))
    reinterpret_cast<uint32_t*>(&sp)[threadIdx.x] = reinterpret_cast<const uint32_t*>(samples + frame)[threadIdx.x];
  __syncthreads();
  const Rcps rcp{__frcp_rn(cfg.map_res), __frcp_rn(cfg.fx), __frcp_rn(cfg.fy)};
  const float* dplane = depth + (size_t)frame * N;
  const uint8_t* vplane = valid ? valid + (size_t)frame * N : nullptr;
  uint32_t* plane = acc + (size_t)slot * slot_words;
  const float fill = cfg.fill_value;
  for (int n0 = (blockIdx.x * 256 + threadIdx.x) * 4; n0 < N; n0 += gridDim.x * 1024) {
    int cl[4] = {-1, -1, -1, -1};
    float y[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {  // W % 4 == 0, 16-byte aligned planes: the quad lies in one image row
      const float4 z4 = ld_stream_f4(dplane + n0);
      const float z[4] = {z4.x, z4.y, z4.z, z4.w};
      uint32_t vm = 0x01010101u;
      if (vplane) vm = *reinterpret_cast<const uint32_t*>(vplane + n0);
      const int r = n0 / cfg.W, c = n0 - r * cfg.W;
      if (FAST) {
        const float yy = cfg.flip_h ? __fsub_rn((float)(cfg.H - 1), (float)r) : (float)r;
        const float yn = div_by_rcp(__fsub_rn(yy, cfg.cy), cfg.fy, rcp.fy);
        bool rowok = true;
        const int kb = cfg.clip_border;
        if (kb > 0) rowok = (r >= kb) && (r < cfg.H - kb);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float xn = div_by_rcp(__fsub_rn((float)(c + k), cfg.cx), cfg.fx, rcp.fx);
          bool ok = rowok && (((vm >> (8 * k)) & 0xffu) != 0);
          if (kb > 0) ok = ok && (c + k >= kb) && (c + k < cfg.W - kb);
          cl[k] = pixel_cell_fast<FAST == 2, true>(cfg, sp, xn, yn, z[k], ok, &y[k], rcp.res);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          cl[k] = pixel_cell(cfg, sp, r, c + k, z[k], ((vm >> (8 * k)) & 0xffu) != 0, &y[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int n = n0 + k;
        if (n < N) {
          const int r = n / cfg.W, c = n - r * cfg.W;
          cl[k] = pixel_cell(cfg, sp, r, c, ld_stream_f1(dplane + n), vplane ? vplane[n] != 0 : true, &y[k]);
        }
      }
    }
    // runs of equal neighbouring cells fold in-thread; the last pixel of a run issues
    const bool p01 = (cl[0] >= 0) && (cl[0] == cl[1]);
    const bool p12 = (cl[1] >= 0) && (cl[1] == cl[2]);
    const bool p23 = (cl[2] >= 0) && (cl[2] == cl[3]);
    y[1] = p01 ? red2<IS_MIN>(y[0], y[1]) : y[1];
    y[2] = p12 ? red2<IS_MIN>(y[1], y[2]) : y[2];
    y[3] = p23 ? red2<IS_MIN>(y[2], y[3]) : y[3];
    if (cl[0] >= 0 && !p01 && beats<IS_MIN>(y[0], fill)) red_max_u32(plane + cl[0], key_of<IS_MIN>(y[0]));
    if (cl[1] >= 0 && !p12 && beats<IS_MIN>(y[1], fill)) red_max_u32(plane + cl[1], key_of<IS_MIN>(y[1]));
    if (cl[2] >= 0 && !p23 && beats<IS_MIN>(y[2], fill)) red_max_u32(plane + cl[2], key_of<IS_MIN>(y[2]));
    if (cl[3] >= 0 && beats<IS_MIN>(y[3], fill)) red_max_u32(plane + cl[3], key_of<IS_MIN>(y[3]));
  }
}

__global__ void __launch_bounds__(256)
hmap_resolve_kernel(uint32_t* __restrict__ acc, unsigned long long slot_words, const DmProjCfg cfg, int frame0, int vec,
                    float* __restrict__ topdown, uint8_t* __restrict__ mask) {
  const int slot = blockIdx.y, frame = frame0 + slot;
  const int M = cfg.Mh * cfg.Mw;
  uint32_t* plane = acc + (size_t)slot * slot_words;
  float* tp = topdown + (size_t)frame * M;
  uint8_t* mp = mask + (size_t)frame * M;
  const float fill = cfg.fill_value;
  const int is_min = cfg.reduction;
  for (int m0 = (blockIdx.x * 256 + threadIdx.x) * 4; m0 < M; m0 += gridDim.x * 1024) {
    if (vec && m0 + 3 < M) {  // M % 4 == 0 and 16-byte aligned outputs
      const uint4 k = __ldcg(reinterpret_cast<const uint4*>(plane + m0));
      if (k.x | k.y | k.z | k.w) __stcg(reinterpret_cast<uint4*>(plane + m0), make_uint4(0u, 0u, 0u, 0u));
      // utils.py:472-491: a key is only ever stored for a value that beats fill, so "key present" == "changed" == mask
      st_stream_f4(tp + m0, make_float4(k.x ? dec_red(k.x, is_min) : fill, k.y ? dec_red(k.y, is_min) : fill,
                                        k.z ? dec_red(k.z, is_min) : fill, k.w ? dec_red(k.w, is_min) : fill));
      st_stream_u32(mp + m0, (k.x ? 1u : 0u) | (k.y ? 0x100u : 0u) | (k.z ? 0x10000u : 0u) | (k.w ? 0x1000000u : 0u));
    } else {
      for (int m = m0; m < min(m0 + 4, M); ++m) {
        const uint32_t k = __ldcg(plane + m);
        if (k) __stcg(plane + m, 0u);
        st_stream_f1(tp + m, k ? dec_red(k, is_min) : fill);
        st_stream_u8(mp + m, k ? 1 : 0);
      }
    }
  }
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

struct DeviceInfo {
  bool ready = false;
  int sms = 0;
};
static DeviceInfo g_dev[64];

// cuTensorMapEncodeTiled through the runtime's driver entry point query: the library keeps linking cudart only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
#ifndef DM_TM_L2PROMO
#define DM_TM_L2PROMO 3  // CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif
// float32 tensor (W, H, planes) of contiguous H x W planes, box (bw, bh, bp); false: not expressible as a tensor map
static bool plane_tensor_map(CUtensorMap* tm, const float* base, int W, int H, long long planes, int bw, int bh, int bp) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc || planes <= 0 || planes > 0xffffffffll) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};  // bytes, multiples of 16
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bp};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)DM_TM_L2PROMO,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dm

using namespace dm;

extern "C" void dm_debug_set_tile_rows(int32_t rows) { dm::g_tile_rows = rows; }

int dm::hmap_project_keys(const float* depth, const uint8_t* valid, const DmProjSample* samples, const DmProjCfg* cfg,
                          int32_t b, void* workspace, size_t workspace_bytes, cudaStream_t stream, uint32_t** planes_out,
                          unsigned long long* slot_words_out) {
  if (!cfg || !depth || !samples || !workspace || !planes_out || !slot_words_out || b <= 0 || cfg->C != 0) return DM_EINVAL;
  if (cfg->H <= 0 || cfg->W <= 0 || cfg->Mh <= 0 || cfg->Mw <= 0) return DM_EINVAL;
  if ((long long)cfg->H * cfg->W >= (1ll << 31) || (long long)cfg->Mh * cfg->Mw >= (1ll << 30)) return DM_EINVAL;
  if (cfg->reduction != 0 && cfg->reduction != 1) return DM_EINVAL;
  const ProjPlan p = make_plan(*cfg, b);
  if (!p.lean || b > p.ring) return DM_EINVAL;  // the caller falls back to dm_orth_project_f32
  if (workspace_bytes < p.workspace_bytes() || !aligned(workspace, 256)) return DM_EWORKSPACE;
  const int N = cfg->H * cfg->W;
  uint32_t* planes = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + p.ctrl_bytes + p.flag_bytes);
  const int pvec = (cfg->W % 4 == 0) && aligned(depth, 16) && (!valid || aligned(valid, 4));
  void (*kern)(const float*, const uint8_t*, const DmProjSample*, DmProjCfg, int, int, uint32_t*, unsigned long long) = nullptr;
  const bool mn = cfg->reduction != 0;
  switch (cfg->fast_steps) {
    case 1: kern = mn ? hmap_proj_kernel<1, true> : hmap_proj_kernel<1, false>; break;
    case 2: kern = mn ? hmap_proj_kernel<2, true> : hmap_proj_kernel<2, false>; break;
    default: kern = mn ? hmap_proj_kernel<0, true> : hmap_proj_kernel<0, false>; break;
  }
  kern<<<dim3((N + 1023) / 1024, b), 256, 0, stream>>>(depth, valid, samples, *cfg, 0, pvec, planes,
                                                       (unsigned long long)p.slot_words);
  DM_LAUNCHED();
  *planes_out = planes;
  *slot_words_out = (unsigned long long)p.slot_words;
  return DM_OK;
}

extern "C" size_t dm_orth_project_workspace_bytes(const DmProjCfg* cfg, int32_t b) {
  if (!cfg || b <= 0 || cfg->Mh <= 0 || cfg->Mw <= 0) return 0;
  const ProjPlan p = make_plan(*cfg, b);
  return p.workspace_bytes();
}

extern "C" int dm_orth_project_f32(const float* depth, const float* values, const uint8_t* valid,
                                   const DmProjSample* samples, const DmProjCfg* cfg, int32_t b,
                                   float* topdown, uint8_t* mask, float* height, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  DM_TRACE();
  if (!cfg || b < 0) return DM_EINVAL;
  if (b == 0) return DM_OK;
  if (!depth || !samples || !topdown || !mask || !workspace) return DM_EINVAL;
  if (cfg->H <= 0 || cfg->W <= 0 || cfg->Mh <= 0 || cfg->Mw <= 0 || cfg->C < 0) return DM_EINVAL;
  if ((long long)cfg->H * cfg->W >= (1ll << 31) || (long long)cfg->Mh * cfg->Mw >= (1ll << 30)) return DM_EINVAL;
  if (cfg->C > 0 && !values) return DM_EINVAL;
  if (cfg->C > 0 && cfg->want_height && !height) return DM_EINVAL;
  if (cfg->reduction != 0 && cfg->reduction != 1) return DM_EINVAL;
  if (!aligned(workspace, 256)) return DM_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProjPlan p = make_plan(*cfg, b);
  if (workspace_bytes < p.workspace_bytes()) return DM_EWORKSPACE;
  int dev = 0;
  DM_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return DM_EINVAL;
  // an earlier launch on this device ran into the dependency-wait guard: its outputs are garbage (it re-zeroed
  // its workspace itself); report it once, launch nothing
  if (take_timeout(dev)) return DM_ETIMEOUT;
  if (!g_dev[dev].ready) {
    DM_CUDA_OK(cudaFuncSetAttribute(proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    DM_CUDA_OK(cudaFuncSetAttribute(resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
#define DM_WS_ATTR(F, MN) \
    DM_CUDA_OK(cudaFuncSetAttribute(proj_ws_kernel<F, MN, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
    DM_CUDA_OK(cudaFuncSetAttribute(proj_ws_kernel<F, MN, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
    DM_CUDA_OK(cudaFuncSetAttribute(proj_ws_kernel<F, MN, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
    DM_CUDA_OK(cudaFuncSetAttribute(proj_ws_kernel<F, MN, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
    DM_CUDA_OK(cudaFuncSetAttribute(proj_ws_kernel<F, MN, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    DM_WS_ATTR(0, false) DM_WS_ATTR(1, false) DM_WS_ATTR(2, false)
    DM_WS_ATTR(0, true) DM_WS_ATTR(1, true) DM_WS_ATTR(2, true)
#undef DM_WS_ATTR
    DM_CUDA_OK(cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev));
    g_dev[dev].ready = true;
  }
  const int N = cfg->H * cfg->W, M = cfg->Mh * cfg->Mw;
  if (p.lean) {
    uint32_t* planes = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + p.ctrl_bytes + p.flag_bytes);
    const int pvec = (cfg->W % 4 == 0) && aligned(depth, 16) && (!valid || aligned(valid, 4));
    const int rvec = (M % 4 == 0) && aligned(topdown, 16) && aligned(mask, 4);
    void (*kern)(const float*, const uint8_t*, const DmProjSample*, DmProjCfg, int, int, uint32_t*, unsigned long long) = nullptr;
    const bool mn = cfg->reduction != 0;
    switch (cfg->fast_steps) {
      case 1: kern = mn ? hmap_proj_kernel<1, true> : hmap_proj_kernel<1, false>; break;
      case 2: kern = mn ? hmap_proj_kernel<2, true> : hmap_proj_kernel<2, false>; break;
      default: kern = mn ? hmap_proj_kernel<0, true> : hmap_proj_kernel<0, false>; break;
    }
    for (int f0 = 0; f0 < b; f0 += p.ring) {
      const int nf = (b - f0) < p.ring ? (b - f0) : p.ring;
      dim3 gp((N + 1023) / 1024, nf), gr((M + 1023) / 1024, nf);
      kern<<<gp, 256, 0, stream>>>(depth, valid, samples, *cfg, f0, pvec, planes, (unsigned long long)p.slot_words);
      DM_LAUNCHED();
      hmap_resolve_kernel<<<gr, 256, 0, stream>>>(planes, (unsigned long long)p.slot_words, *cfg, f0, rvec, topdown, mask);
      DM_LAUNCHED();
    }
    return DM_OK;
  }
  uint32_t* ctrl = static_cast<uint32_t*>(workspace);
  uint32_t* flags = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + p.ctrl_bytes);
  uint32_t* acc = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + p.ctrl_bytes + p.flag_bytes);
  ProjDims d{p.Cv, p.hasH, p.CU, p.CP, p.rows, p.tile, p.ring, p.lag, p.nsl, (unsigned long long)p.slot_words,
             (unsigned long long)p.stage_bytes};
  const int tiles = (N + p.tile - 1) / p.tile;
  const int rtiles = (M + kResolveCells - 1) / kResolveCells;
  // the warp-specialised TMA kernel needs 16-byte aligned plane starts / row sizes, pixel quads
  // that do not straddle image rows, and two stages that fit in shared memory
  // 2-D tiles need tensor maps of the inputs (16-byte aligned bases and row pitches); anything else takes row tiles
  CUtensorMap tmv{}, tmd{};
  int r2d = p.ws_r2d;
  if (r2d) {
    const int tw = 32 * p.ws_warps;
    bool ok = cfg->W % 4 == 0 && aligned(depth, 16) && (!values || aligned(values, 16)) &&
              plane_tensor_map(&tmd, depth, cfg->W, cfg->H, b, tw, r2d, 1);
    if (ok && cfg->C > 0) ok = plane_tensor_map(&tmv, values, cfg->W, cfg->H, (long long)b * cfg->C, tw, r2d, p.ws_cg);
    if (!ok) {
      r2d = 0;
      p = make_plan(*cfg, b, 0);  // same workspace layout, row tiles
    }
  }
  const int ws_tile = 128 * p.ws_warps, ws_threads = 32 * (p.ws_warps + 1);
  const long long ws_tiles = (r2d ? (long long)((cfg->H + r2d - 1) / r2d) * ((cfg->W + 32 * p.ws_warps - 1) / (32 * p.ws_warps))
                                  : (long long)((N + ws_tile - 1) / ws_tile)) * p.ws_groups;
  const int ws_rtiles = (M + 64 * p.ws_warps * kResK - 1) / (64 * p.ws_warps * kResK);
  const long long ws_total = (long long)(b + p.lag) * (ws_tiles + ws_rtiles);
  const bool ws_ok = g_tile_rows != -2 && (N % 4 == 0) && (cfg->W % 4 == 0) && aligned(depth, 16) && (!values || aligned(values, 16)) &&
                     (!valid || aligned(valid, 4)) && p.smem_ws <= 220 * 1024 && ws_total < (1ll << 31) &&
                     cfg->fast_steps >= 0 && cfg->fast_steps <= 2 && p.ws_cg <= 32 &&
                     (unsigned long long)p.ring * p.slot_words < (1ull << 31);
  if (ws_ok) {
    ProjDims dw = d;
    dw.tile = ws_tile;
    dw.rows = p.ws_cg + 1;
    dw.groups = p.ws_groups;
    dw.cg = p.ws_cg;
    dw.stage_bytes = p.ws_stage_bytes;
    dw.ws_words = (p.workspace_bytes() - p.ctrl_bytes) / 4;
    void (*kern)(const float*, const float*, const uint8_t*, const DmProjSample*, DmProjCfg, ProjDims, int,
                 uint32_t*, uint32_t*, uint32_t*, float*, uint8_t*, float*, ProjGuard, CUtensorMap, CUtensorMap) = nullptr;
    const bool mn = cfg->reduction != 0;
#define DM_WS_PICK(WW, RR)                                                                             \
    switch (cfg->fast_steps) {                                                                         \
      case 1: kern = mn ? proj_ws_kernel<1, true, WW, RR> : proj_ws_kernel<1, false, WW, RR>; break;   \
      case 2: kern = mn ? proj_ws_kernel<2, true, WW, RR> : proj_ws_kernel<2, false, WW, RR>; break;   \
      default: kern = mn ? proj_ws_kernel<0, true, WW, RR> : proj_ws_kernel<0, false, WW, RR>; break;  \
    }
    if (r2d == 8) { DM_WS_PICK(2, 8) }
    else if (r2d == 4) { if (p.ws_warps == 4) { DM_WS_PICK(4, 4) } else { DM_WS_PICK(2, 4) } }
    else { if (p.ws_warps == 4) { DM_WS_PICK(4, 0) } else { DM_WS_PICK(2, 0) } }
#undef DM_WS_PICK
    int per_sm = 0;
    DM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ws_threads, p.smem_ws));
    if (per_sm < 1) return DM_EINVAL;
#ifdef DM_EXPERIMENT_KNOBS
    if (getenv("DM_DEBUG_PLAN")) fprintf(stderr, "proj_ws: per_sm %d sms %d smem %zu threads %d r2d %d\n", per_sm, g_dev[dev].sms, p.smem_ws, ws_threads, r2d);
#endif
    long long grid = (long long)g_dev[dev].sms * per_sm;
    if (grid > ws_total) grid = ws_total;
    kern<<<(unsigned)grid, ws_threads, p.smem_ws, stream>>>(depth, values, valid, samples, *cfg, dw, b, ctrl, flags,
                                                            acc, topdown, mask, height, proj_guard(dev), tmv, tmd);
    DM_LAUNCHED();
    return DM_OK;
  }
  if (p.smem_tile > 220 * 1024) return DM_EINVAL;  // too many channels for one pass
  for (int f0 = 0; f0 < b; f0 += p.ring) {
    const int nf = (b - f0) < p.ring ? (b - f0) : p.ring;
    dim3 gp(tiles, nf), gr(rtiles, nf);
    proj_kernel<<<gp, kProjThreads, p.smem_tile, stream>>>(depth, values, valid, samples, *cfg, d, f0, acc);
    DM_LAUNCHED();
    resolve_kernel<<<gr, kProjThreads, p.smem_resolve, stream>>>(acc, *cfg, d, f0, topdown, mask, height);
    DM_LAUNCHED();
  }
  return DM_OK;
}
