// Fused orthographic projection of depth + CLASS-ID planes: orth_project (maps.py:127-351) for the reference's own
// semantic usage — demos/object_map/run.py:117-124 turns an integer segmentation into one_hot(seg).float() and hands
// the (C, H, W) float planes to orth_project as `value_map`.  This entry takes the (b, 1, H, W) uint8 class ids
// themselves and produces, bit for bit, what the float path produces for their one-hot expansion:
//   * 5 bytes per pixel enter the SM instead of 4·(C+1): at C = 16 the input shrinks 13.6x (host→device copy of the
//     *_host entry, and the device-side algorithmic bytes: 34.3 → 15.0 MB per 480x640 frame);
//   * a cell's C value keys collapse into ONE presence word (bit c = "a pixel of class c landed here"): one
//     RED.OR + one RED.MAX (height key) per run of pixels instead of C + 1 REDs — the SM-side RED path was what
//     separated the float kernel from its load/store pipeline;
//   * the resolve pass rebuilds every channel from (height key, presence word): with reduction max a hit cell holds
//     max(fill, 1) where the class is present and max(fill, 0) where it is not (scatter_max starts from the filled
//     canvas, utils.py:472-477), mask = "differs from fill" (utils.py:489-491); with min a channel keeps 1 only if
//     every pixel of the cell carries that class.  A class id >= C has an all-zero one-hot row (bit C, "other").
//
// Schedule: the persistent ticket scheme of dm_project.cu (projection tiles of frame f + lag interleaved with resolve
// tiles of frame f over a sparse ring of accumulation slots, per-frame completion counters), but with nothing to
// transpose there is nothing to stage: a CTA is one scheduler warp (claims tickets two ahead, acquires the
// dependency, fetches the sample block, publishes completions) and eight worker warps that load their 4 pixels per
// lane straight into registers (one 128-bit depth load + one 32-bit label load, coalesced; the scheduler prefetches
// the lines into L2).  Items travel through two mbarrier-guarded slots, so a worker warp that finishes its part early
// starts on the next item instead of waiting for the slowest warp of the CTA.
#include <type_traits>

#include "dm_project.cuh"

namespace dm {

constexpr int kLblWarps = 8;                        // worker warps per CTA
constexpr int kLblThreads = 32 * (kLblWarps + 1);   // + the scheduler warp
#ifndef DM_LBL_PROJ_J
#define DM_LBL_PROJ_J 1
#endif
constexpr int kLblProjJ = DM_LBL_PROJ_J;            // 128-pixel sub-tiles per worker warp and projection ticket
constexpr int kLblTile = 128 * kLblWarps * kLblProjJ;  // pixels per projection ticket
#ifndef DM_LBL_RES_K
#define DM_LBL_RES_K 2  // measured (room / iid ms per 64 frames): 1: 0.259 / 0.391, 2: 0.252 / 0.352
#endif
constexpr int kLblResK = DM_LBL_RES_K;              // 64-cell slices per worker warp and resolve ticket
constexpr int kLblResCells = 64 * kLblWarps * kLblResK;
#ifndef DM_LBL_RING
#define DM_LBL_RING 64
#endif
// Slots are small (8 or 12 bytes per cell, 1.3 MB per 400x400 frame) and only their touched lines ever travel, so a
// batch of up to kLblMaxRing frames gets a slot per frame: no slot is reused inside the launch, a projection item
// never waits, and — what matters — a resolve item's completion need not be published.  That release is a
// MEMBAR.ALL.GPU behind 40 KB of output stores in flight: 3.6 us each in ncu r02d, a third of the scheduler's time.
constexpr int kLblMaxRing = DM_LBL_RING;
constexpr int kLblList = 128 + 4;                   // runlet list entries per worker warp
constexpr uint32_t kKeyNegInf = 0x007fffffu;        // enc(-inf); smaller non-zero keys only mark "hit" (NaN height)

struct LblPlan {
  int C, W2, CP, ring, lag, nsl;
  size_t slot_words, ctrl_bytes, flag_bytes;
  size_t workspace_bytes() const { return ctrl_bytes + flag_bytes + slot_words * 4 * (size_t)ring; }
};

static LblPlan make_lbl_plan(const DmProjCfg& cfg, int b) {
  LblPlan p{};
  p.C = cfg.C;
  p.W2 = (cfg.C + 1 <= 32) ? 1 : 2;  // presence words per cell: bits 0..C-1 classes, bit C "other"
  p.CP = 1 + p.W2;                   // + the height key
  const size_t M = (size_t)cfg.Mh * cfg.Mw;
  p.slot_words = (M * p.CP + 3) & ~(size_t)3;
  const int bb = b > 0 ? b : 1;
  p.ring = bb < kLblMaxRing ? bb : kLblMaxRing;
  if (p.ring < 2) p.ring = 2;
  p.lag = p.ring / 2;
  p.nsl = (int)((M + kSliceCells - 1) / kSliceCells);
  p.ctrl_bytes = ((size_t)(kCtrlWords + 2 * bb) * 4 + 255) & ~(size_t)255;
  p.flag_bytes = ((size_t)p.ring * p.nsl * kFlagStride * 4 + 255) & ~(size_t)255;
  return p;
}

struct LblDims {
  int C, CP, ring, lag, nsl, hasH;
  int vec_in;   // depth 16-byte / labels and valid 4-byte aligned for every frame: vector loads
  int vec_out;  // 0: scalar stores; 1: float planes 16-byte and mask planes 4-byte aligned for every (frame, channel)
                // (M % 4 == 0); 2: mask planes 16-byte aligned as well (M % 16 == 0)
  unsigned long long slot_words, ws_words;
};

struct LblItem {
  int kind, frame, idx, ok;
};

__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void red_or_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int W2>
using LblBits = std::conditional_t<W2 == 1, uint32_t, unsigned long long>;

// The 4 pixels of a lane as they come from memory: depth, class ids, valid bytes (0x01 each when there is no valid
// map).  Loaded for all sub-tiles of a ticket before the first is worked on, so the latencies overlap.
struct LblQuad {
  float z[4];
  uint32_t lab, vm;
};

__device__ __forceinline__ LblQuad lbl_load_quad(const LblDims& d, const float* __restrict__ dframe,
                                                 const uint8_t* __restrict__ lframe, const uint8_t* __restrict__ vframe,
                                                 int n0, int N) {
  LblQuad q;
  q.z[0] = q.z[1] = q.z[2] = q.z[3] = 0.0f;
  q.lab = 0;
  q.vm = 0;
  if (n0 >= N) return q;
  if (d.vec_in && n0 + 3 < N) {
    const float4 z4 = ld_stream_f4(dframe + n0);
    q.z[0] = z4.x; q.z[1] = z4.y; q.z[2] = z4.z; q.z[3] = z4.w;
    q.lab = ld_stream_u32(lframe + n0);
    q.vm = vframe ? ld_stream_u32(vframe + n0) : 0x01010101u;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool in = n0 + k < N;
      q.z[k] = in ? ld_stream_f1(dframe + n0 + k) : 0.0f;
      q.lab |= (in ? (uint32_t)lframe[n0 + k] : 0u) << (8 * k);
      q.vm |= (in ? (vframe ? (uint32_t)(vframe[n0 + k] != 0) : 1u) : 0u) << (8 * k);
    }
  }
  return q;
}

// ---- projection of 128 pixels by one warp (lane = 4 consecutive pixels) ---------------------------------------
template <int FAST, int W2>
__device__ __forceinline__ void lbl_proj_warp(const DmProjCfg& cfg, const LblDims& d, const DmProjSample& sp,
                                              const Rcps& rcp, const LblQuad& quad, int n0, int lane, uint32_t* list,
                                              uint32_t* __restrict__ acc, uint32_t slot_off,
                                              uint32_t* __restrict__ slot_flags) {
  using Bits = LblBits<W2>;
  const int N = cfg.H * cfg.W;
  int cl[4] = {-1, -1, -1, -1};
  float y[4] = {0.f, 0.f, 0.f, 0.f};
  Bits bt[4] = {0, 0, 0, 0};
  if (n0 < N) {
    const float* z = quad.z;
    const uint32_t lab = quad.lab, vm = quad.vm;
    const int r = n0 / cfg.W, c = n0 - r * cfg.W;
    if (FAST && c + 3 < cfg.W) {
      // maps.py:677-678 column / row factors rn(rn(c - cx) / fx), rn(rn(yy - cy) / fy): exact division through the
      // reciprocals (div_by_rcp), exactly as the float kernel's phase A
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
    } else {  // generic steps, or a quad that wraps into the next image row (W % 4 != 0)
      int rk = r, ck = c;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cl[k] = pixel_cell(cfg, sp, rk, ck, z[k], ((vm >> (8 * k)) & 0xffu) != 0, &y[k]);
        if (++ck == cfg.W) { ck = 0; ++rk; }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int l = (int)((lab >> (8 * k)) & 0xffu);
      bt[k] = (Bits)1 << (l < cfg.C ? l : cfg.C);
    }
  }
  // in-thread runs: pixel k continues into k+1 when both are valid and share the cell
  const bool p01 = (cl[0] >= 0) && (cl[0] == cl[1]);
  const bool p12 = (cl[1] >= 0) && (cl[1] == cl[2]);
  const bool p23 = (cl[2] >= 0) && (cl[2] == cl[3]);
  const bool t0 = (cl[0] >= 0) && !p01, t1 = (cl[1] >= 0) && !p12, t2 = (cl[2] >= 0) && !p23, t3 = cl[3] >= 0;
  y[1] = p01 ? fmaxf(y[0], y[1]) : y[1]; bt[1] = p01 ? (bt[0] | bt[1]) : bt[1];
  y[2] = p12 ? fmaxf(y[1], y[2]) : y[2]; bt[2] = p12 ? (bt[1] | bt[2]) : bt[2];
  y[3] = p23 ? fmaxf(y[2], y[3]) : y[3]; bt[3] = p23 ? (bt[2] | bt[3]) : bt[3];
  const int cnt = (int)t0 + (int)t1 + (int)t2 + (int)t3;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;  // warp-uniform
  uint32_t* lcell = list;
  float* ly = reinterpret_cast<float*>(list + kLblList);
  uint32_t* lb = list + 2 * kLblList;
  const int o0 = incl - cnt, o1 = o0 + (int)t0, o2 = o1 + (int)t1, o3 = o2 + (int)t2;
  auto put = [&](int o, int cell, float yy, Bits b) {
    lcell[o] = (uint32_t)cell;
    ly[o] = yy;
    lb[o] = (uint32_t)b;
    if (W2 == 2) lb[kLblList + o] = (uint32_t)((unsigned long long)b >> 32);
  };
  if (t0) put(o0, cl[0], y[0], bt[0]);
  if (t1) put(o1, cl[1], y[1], bt[1]);
  if (t2) put(o2, cl[2], y[2], bt[2]);
  if (t3) put(o3, cl[3], y[3], bt[3]);
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
  __syncwarp();
  // lane = runlet: neighbouring runlets of one cell (runs longer than a pixel quad) are folded by a segmented scan
  // over the lanes; the last runlet of a run issues the cell's two (three) REDs, which share a 32-byte sector
  for (int base = 0; base < total; base += 32) {
    const int i = base + lane;
    const bool active = i < total;
    const uint32_t cellv = active ? lcell[i] : 0xffffffffu;
    float v = active ? ly[i] : -INFINITY;
    uint32_t blo = active ? lb[i] : 0u, bhi = 0u;
    if (W2 == 2) bhi = active ? lb[kLblList + i] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float pv = __shfl_up_sync(0xffffffffu, v, o);
      const uint32_t pc = __shfl_up_sync(0xffffffffu, cellv, o);
      const uint32_t pl = __shfl_up_sync(0xffffffffu, blo, o);
      uint32_t ph = 0u;
      if (W2 == 2) ph = __shfl_up_sync(0xffffffffu, bhi, o);
      if (lane >= o && pc == cellv) {
        v = fmaxf(v, pv);
        blo |= pl;
        bhi |= ph;
      }
    }
    const uint32_t nc = __shfl_down_sync(0xffffffffu, cellv, 1);
    const bool last = lane == 31 || nc != cellv;
    if (active && last) {
      uint32_t* cellp = acc + (slot_off + cellv * (uint32_t)(1 + W2));
      red_max_u32(cellp, (v != v) ? 1u : enc(v));  // a NaN height never wins (utils.py:475) but the cell was hit
      red_or_u32(cellp + 1, blo);
      if (W2 == 2 && bhi) red_or_u32(cellp + 2, bhi);
    }
  }
  __syncwarp();  // the list is reused by this warp's next tile
}

// One channel of one cell from (height key, presence bits): scatter_max / scatter_min of the one-hot rows into a
// canvas filled with `fill` (utils.py:472-477), and the "changed" mask (utils.py:489-491).
template <int W2>
__device__ __forceinline__ float lbl_channel(uint32_t key, LblBits<W2> bits, int c, float fill, int is_min, uint32_t* m) {
  const bool one = is_min ? (bits == ((LblBits<W2>)1 << c)) : (((bits >> c) & 1) != 0);
  const float v = one ? 1.0f : 0.0f;
  const bool win = key != 0u && (is_min ? (v < fill) : (v > fill));
  *m = win ? 1u : 0u;
  return win ? v : fill;
}

// The keys of the two cells a lane owns in a flagged, fully vectorisable 64-cell slice (cells 2 * lane, 2 * lane + 1):
// loaded with ld.cg, re-zeroed where set.  Issued for all slices of a ticket before the first is decoded.
template <int W2>
struct LblKeys {
  uint32_t k0, k1;
  LblBits<W2> b0, b1;
};

template <int W2>
__device__ __forceinline__ LblKeys<W2> lbl_load_keys(uint32_t* __restrict__ acc_slot, const DmProjCfg& cfg,
                                                     const LblDims& d, uint32_t flagged, int slice, int lane) {
  using Bits = LblBits<W2>;
  constexpr int CP = 1 + W2;
  LblKeys<W2> out{0u, 0u, (Bits)0, (Bits)0};
  const int M = cfg.Mh * cfg.Mw;
  const int cell0 = slice * 64;
  if (!flagged || !d.vec_out || M - cell0 < 64) return out;
  uint32_t* src = acc_slot + (size_t)cell0 * CP;  // 16-byte aligned: cell0 % 64 == 0
  if (W2 == 1) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src) + lane);
    if (v.x | v.y | v.z | v.w) __stcg(reinterpret_cast<uint4*>(src) + lane, make_uint4(0u, 0u, 0u, 0u));
    out.k0 = v.x; out.b0 = (Bits)v.y; out.k1 = v.z; out.b1 = (Bits)v.w;
  } else {
    uint2* s2 = reinterpret_cast<uint2*>(src) + lane * 3;
    const uint2 a = __ldcg(s2), bq = __ldcg(s2 + 1), cq = __ldcg(s2 + 2);
    if (a.x | a.y | bq.x | bq.y | cq.x | cq.y) {
      __stcg(s2, make_uint2(0u, 0u)); __stcg(s2 + 1, make_uint2(0u, 0u)); __stcg(s2 + 2, make_uint2(0u, 0u));
    }
    out.k0 = a.x; out.b0 = (Bits)((unsigned long long)a.y | ((unsigned long long)bq.x << 32));
    out.k1 = bq.y; out.b1 = (Bits)((unsigned long long)cq.x | ((unsigned long long)cq.y << 32));
  }
  return out;
}

// ---- resolve of one 64-cell slice by one warp -----------------------------------------------------------------
template <int W2>
__device__ __forceinline__ void lbl_resolve_slice(uint32_t* __restrict__ acc_slot, const DmProjCfg& cfg,
                                                  const LblDims& d, uint32_t flagged, const LblKeys<W2>& pre,
                                                  int frame, int slice, int lane, float* __restrict__ topdown,
                                                  uint8_t* __restrict__ mask, float* __restrict__ height) {
  using Bits = LblBits<W2>;
  constexpr int CP = 1 + W2;
  const int M = cfg.Mh * cfg.Mw;
  const int cell0 = slice * 64;
  const int ncell = min(64, M - cell0);
  if (ncell <= 0) return;
  const int C = cfg.C;
  const bool vec = d.vec_out && ncell == 64;
  const size_t plane0 = (size_t)frame * C * M + cell0;
  const float fill = cfg.fill_value;
  if (!flagged) {  // nothing landed in these 64 cells: constant stores, the ring is not even read
    if (vec) {
      const float4 f4 = make_float4(fill, fill, fill, fill);
      for (int c = lane >> 4; c < C; c += 2) st_stream_f4(topdown + plane0 + (size_t)c * M + (lane & 15) * 4, f4);
      if (d.vec_out > 1) {  // mask planes 16-byte aligned (M % 16 == 0): 4 lanes cover a channel's 64 bytes
        for (int c = lane >> 2; c < C; c += 8) st_stream_u4(mask + plane0 + (size_t)c * M + (lane & 3) * 16, 0u);
      } else {              // 4-byte aligned (M % 4 == 0): 16 lanes per channel
        for (int c = lane >> 4; c < C; c += 2) st_stream_u32(mask + plane0 + (size_t)c * M + (lane & 15) * 4, 0u);
      }
      if (d.hasH && lane < 16)
        st_stream_f4(height + (size_t)frame * M + cell0 + lane * 4, make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY));
    } else {
      for (int j = lane; j < ncell; j += 32) {
        for (int c = 0; c < C; ++c) {
          st_stream_f1(topdown + plane0 + (size_t)c * M + j, fill);
          st_stream_u8(mask + plane0 + (size_t)c * M + j, 0);
        }
        if (d.hasH) st_stream_f1(height + (size_t)frame * M + cell0 + j, -INFINITY);
      }
    }
    return;
  }
  uint32_t* src = acc_slot + (size_t)cell0 * CP;  // 16-byte aligned: cell0 % 64 == 0
  if (vec) {  // lane owns cells 2 * lane and 2 * lane + 1: their keys were loaded (and re-zeroed) by lbl_load_keys
    const uint32_t k0 = pre.k0, k1 = pre.k1;
    const Bits b0 = pre.b0, b1 = pre.b1;
    float* tp = topdown + plane0 + 2 * lane;
    uint8_t* mp = mask + plane0 + 2 * lane;
    for (int c = 0; c < C; ++c) {
      uint32_t m0, m1;
      const float v0 = lbl_channel<W2>(k0, b0, c, fill, cfg.reduction, &m0);
      const float v1 = lbl_channel<W2>(k1, b1, c, fill, cfg.reduction, &m1);
      st_stream_f2(tp, v0, v1);
      st_stream_u16(mp, m0 | (m1 << 8));
      tp += M;
      mp += M;
    }
    if (d.hasH)  // maps.py:340-348: max against -inf
      st_stream_f2(height + (size_t)frame * M + cell0 + 2 * lane, k0 > kKeyNegInf ? dec(k0) : -INFINITY,
                   k1 > kKeyNegInf ? dec(k1) : -INFINITY);
  } else {
    for (int j = lane; j < ncell; j += 32) {
      uint32_t w[CP];
#pragma unroll
      for (int q = 0; q < CP; ++q) {
        w[q] = __ldcg(src + j * CP + q);
        if (w[q]) __stcg(src + j * CP + q, 0u);
      }
      Bits bits = (Bits)w[1];
      if (W2 == 2) bits = (Bits)((unsigned long long)w[1] | ((unsigned long long)w[CP - 1] << 32));
      for (int c = 0; c < C; ++c) {
        uint32_t m;
        const float v = lbl_channel<W2>(w[0], bits, c, fill, cfg.reduction, &m);
        st_stream_f1(topdown + plane0 + (size_t)c * M + j, v);
        st_stream_u8(mask + plane0 + (size_t)c * M + j, (uint8_t)m);
      }
      if (d.hasH) st_stream_f1(height + (size_t)frame * M + cell0 + j, w[0] > kKeyNegInf ? dec(w[0]) : -INFINITY);
    }
  }
}

// Item slots per CTA and tickets the scheduler prepares at a time (one lane each).  A ticket costs the scheduler an
// acquire load of its dependency, a read of its slice flags and a release of its completion counter — three L2 round
// trips that, taken one ticket at a time, are longer than the ~1.5 us the eight workers need for the item (ncu r02b:
// 37 % of all warp samples were workers polling for the next item).  Taken for kLblBatch tickets by kLblBatch lanes
// they are one round trip each per batch (the slice flags are read by the workers: eight warps at a time).  The other side of the trade: every CTA holds kLblSlots + kLblBatch
// tickets, and all CTAs together must not span more frames than the resolve pass lags behind the projection, or
// the dependency waits stop being rare (r02c, 8 slots + 4 + 4 claimed ahead at lag 8: 1.1 M spins per launch, 0.30 ms
// instead of 0.25) — the host picks the lag from the number of tickets in flight (lbl_schedule).
#ifndef DM_LBL_SLOTS
#define DM_LBL_SLOTS 4
#endif
#ifndef DM_LBL_BATCH
#define DM_LBL_BATCH 2
#endif
constexpr int kLblSlots = DM_LBL_SLOTS;
constexpr int kLblBatch = DM_LBL_BATCH;
constexpr int kLblItemSlices = kLblWarps * kLblResK;  // 64-cell slices of one resolve item

template <int FAST, int W2>
__global__ void __launch_bounds__(kLblThreads, 4)
proj_lbl_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ labels, const uint8_t* __restrict__ valid,
                const DmProjSample* __restrict__ samples, const DmProjCfg cfg, const LblDims d, int b,
                uint32_t* __restrict__ ctrl, uint32_t* __restrict__ flags, uint32_t* __restrict__ acc,
                float* __restrict__ topdown, uint8_t* __restrict__ mask, float* __restrict__ height,
                const ProjGuard guard) {
  // The scheduler posts item k into slot k % kLblSlots (full[slot]) once the workers have released item
  // k - kLblSlots (empty[slot], one arrival per worker warp); a worker warp that finishes its part of an item early
  // starts on the next one instead of waiting for the slowest warp of the CTA (a CTA-wide barrier per item cost
  // 27 % of the warp samples in stall_barrier, ncu r02a).
  constexpr int S = kLblSlots, G = kLblBatch;
  __shared__ __align__(16) LblItem s_item[S];
  __shared__ __align__(16) DmProjSample s_sample[S];
  __shared__ __align__(8) uint64_t s_full[S], s_empty[S];
  __shared__ int2 s_meta[S];  // (kind, frame) of the item in a slot, for the scheduler's own bookkeeping
  __shared__ uint32_t s_list[kLblWarps][(2 + W2) * kLblList];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = cfg.H * cfg.W, M = cfg.Mh * cfg.Mw;
  const int P = (N + kLblTile - 1) / kLblTile;
  const int R = (M + kLblResCells - 1) / kLblResCells;
  const int lag = d.lag, ring = d.ring;
  uint32_t* proj_done = ctrl + kCtrlWords;
  uint32_t* resolve_done = proj_done + b;
  const unsigned total = (unsigned)(b + lag) * (unsigned)(P + R);
  if (tid < S) {
    mbar_init(&s_full[tid], 1);
    mbar_init(&s_empty[tid], kLblWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kLblWarps) {
    // ===================== scheduler =====================
    // items are numbered in posting order; `posted` items have been handed to the workers, the first `published`
    // of them have been waited for (every worker warp released them) and their completion counters incremented
    unsigned posted = 0, published = 0;
    auto ensure_published = [&](unsigned upto) {
      if ((int)(upto - published) <= 0) return;
      for (unsigned j = published; j != upto; ++j) mbar_wait(&s_empty[j % S], (j / S) & 1u);
      if (lane < upto - published) {  // one lane per completed item: one release fence + RED instruction for all
        const int2 m = s_meta[(published + lane) % S];
        if (m.x == kItemProj || (m.x == kItemResolve && b > ring))  // b <= ring: nobody waits for a resolve
          red_release_add1((m.x == kItemProj ? proj_done : resolve_done) + m.y);
      }
      published = upto;
    };
    // tickets are claimed one batch ahead: every L2 round trip the scheduler waits for is a round trip of a memory
    // system that the workers keep saturated (2-3 us each under load), and they add up per batch
    unsigned next_batch = 0;
    if (lane == 0) next_batch = atomicAdd(ctrl, (unsigned)G);
    bool done = false;
    while (!done) {
      const unsigned t0 = __shfl_sync(0xffffffffu, next_batch, 0);
      if (lane == 0 && t0 < total) next_batch = atomicAdd(ctrl, (unsigned)G);
      // ---- lane g prepares ticket t0 + g
      int kind = kItemNone, frame = 0, idx = 0, ok = 1;
      const uint32_t* dep = nullptr;
      uint32_t dep_target = 0;
      if (lane < G) {
        const unsigned t = t0 + (unsigned)lane;
        if (t >= total) kind = kItemExit;
        else decode_ticket(t, b, P, R, lag, &kind, &frame, &idx);
        // dependency: ring slot resolved by its previous tenant / frame fully projected
        if (kind == kItemProj) {
          if (frame >= ring) { dep = resolve_done + (frame - ring); dep_target = (uint32_t)R + guard.dep_bias; }
        } else if (kind == kItemResolve) {
          dep = proj_done + frame;
          dep_target = (uint32_t)P + guard.dep_bias;
        }
      }
      // the sample blocks of the batch (48 words per item) and, for projection items, the input lines on their way
      // into L2 while the workers finish what they have — issued BEFORE the acquire loads, which would hold them back
      uint32_t sw[(G * 48 + 31) / 32];
#pragma unroll
      for (int j = 0; j < (G * 48 + 31) / 32; ++j) {
        const int w = lane + 32 * j, g = w / 48;
        const int gk = __shfl_sync(0xffffffffu, kind, g & (G - 1)), gf = __shfl_sync(0xffffffffu, frame, g & (G - 1));
        sw[j] = 0;
        if (g < G && gk == kItemProj) sw[j] = reinterpret_cast<const uint32_t*>(samples + gf)[w - 48 * g];
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int gk = __shfl_sync(0xffffffffu, kind, g), gf = __shfl_sync(0xffffffffu, frame, g);
        const int gi = __shfl_sync(0xffffffffu, idx, g);
        if (gk != kItemProj) continue;
        const int n0 = gi * kLblTile;
#pragma unroll
        for (int j = 0; j < kLblTile / 1024; ++j) {  // 32 lanes x 128 bytes = 1024 pixels of depth per round
          const int nj = n0 + j * 1024 + lane * 32;
          if (nj < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(depth + (size_t)gf * N + nj));
        }
        if (lane < kLblTile / 128 && n0 + lane * 128 < N)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(labels + (size_t)gf * N + n0 + lane * 128));
      }
      // acquire loads (one instruction for the batch): they pair with the red.release of the CTAs that completed
      // the frames; the other lanes and the workers inherit the ordering through __syncwarp and the mbarrier hand-off
      int pending = 0;
      if (dep) pending = ld_acquire(dep) < dep_target;
      __syncwarp();
      // ---- post the batch in ticket order
      for (int g = 0; g < G; ++g) {
        const int k_ = __shfl_sync(0xffffffffu, kind, g);
        if (k_ == kItemNone) continue;  // a ticket outside the batch (pipeline fill / drain)
        const int fr = __shfl_sync(0xffffffffu, frame, g), ix = __shfl_sync(0xffffffffu, idx, g);
        int okg = 1;
        if (__shfl_sync(0xffffffffu, pending, g)) {
          // rare: must block.  The frame we wait for may need the very items our workers are finishing: everything
          // this CTA was handed so far is completed and published before the wait.
          ensure_published(posted);
          if (lane == g) ok = wait_count(dep, dep_target, ctrl, guard.spin_ns) ? 1 : 0;
          okg = __shfl_sync(0xffffffffu, ok, g);
          __syncwarp();
        }
        // slot posted % S is free once item posted - S has been released by every worker warp
        ensure_published(posted + 1u - (unsigned)S);
        const unsigned s = posted % S;
        if (k_ == kItemProj) {
#pragma unroll
          for (int j = 0; j < (G * 48 + 31) / 32; ++j) {
            const int w = lane + 32 * j;
            if (w / 48 == g) reinterpret_cast<uint32_t*>(&s_sample[s])[w - 48 * g] = sw[j];
          }
        }
        if (lane == 0) {
          s_item[s] = LblItem{k_, fr, ix, okg};
          s_meta[s] = make_int2(k_, fr);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_full[s]);
        ++posted;
        if (k_ == kItemExit) { done = true; break; }
      }
    }
    ensure_published(posted - 1u);  // everything but the exit item, which nobody releases
    // the last CTA out re-arms the control block for the next call (and scrubs the workspace after a timeout)
    uint32_t last = 0;
    if (lane == 0) {
      __threadfence();
      last = atomicAdd(ctrl + 3, 1u) == gridDim.x - 1 ? 1u : 0u;
      if (last) __threadfence();
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      scrub_after_timeout(ctrl, flags, d.ws_words, lane, 32, guard.status);
      __syncwarp();
      if (lane == 0) {
        for (int i = 0; i < 2 * b; ++i) proj_done[i] = 0;
        ctrl[0] = 0; ctrl[1] = 0; ctrl[2] = 0; ctrl[3] = 0;
        __threadfence();
      }
    }
  } else {
    // ===================== workers =====================
    const Rcps rcp{__frcp_rn(cfg.map_res), __frcp_rn(cfg.fx), __frcp_rn(cfg.fy)};
    for (unsigned k = 0;; ++k) {
      const unsigned s = k % S;
      mbar_wait(&s_full[s], (k / S) & 1u);
      const LblItem it = s_item[s];
      if (it.kind == kItemExit) break;
      if (it.ok) {
        const int rslot = it.frame % ring;
        if (it.kind == kItemProj) {
#ifndef DM_LBL_ABL_NOPROJ  // ablation builds (wrong results): what the kernel costs without one of its halves
          const float* dframe = depth + (size_t)it.frame * N;
          const uint8_t* lframe = labels + (size_t)it.frame * N;
          const uint8_t* vframe = valid ? valid + (size_t)it.frame * N : nullptr;
          const int n0 = it.idx * kLblTile + warp * 128 + lane * 4;  // sub-tile j: + j * 128 * kLblWarps
          LblQuad quads[kLblProjJ];
#pragma unroll
          for (int j = 0; j < kLblProjJ; ++j) quads[j] = lbl_load_quad(d, dframe, lframe, vframe, n0 + j * 128 * kLblWarps, N);
#pragma unroll
          for (int j = 0; j < kLblProjJ; ++j)
            lbl_proj_warp<FAST, W2>(cfg, d, s_sample[s], rcp, quads[j], n0 + j * 128 * kLblWarps, lane, s_list[warp], acc,
                                    (uint32_t)rslot * (uint32_t)d.slot_words, flags + (size_t)rslot * d.nsl * kFlagStride);
#endif
        } else if (it.kind == kItemResolve) {
#ifndef DM_LBL_ABL_NORESOLVE
          uint32_t* acc_slot = acc + (size_t)rslot * d.slot_words;
          const int slice0 = (it.idx * kLblWarps + warp) * kLblResK;
          // sparse ring: lane q reads (and clears) the flag of slice q; a slice nobody flagged holds no key
          uint32_t fl = 0;
          if (lane < kLblResK && slice0 + lane < d.nsl) {
            uint32_t* fp = flags + ((size_t)rslot * d.nsl + slice0 + lane) * kFlagStride;
            fl = __ldcg(fp);
            if (fl) __stcg(fp, 0u);
          }
          const uint32_t flagged = __ballot_sync(0xffffffffu, fl != 0u);
          LblKeys<W2> keys[kLblResK];
#pragma unroll
          for (int q = 0; q < kLblResK; ++q)
            keys[q] = lbl_load_keys<W2>(acc_slot, cfg, d, (flagged >> q) & 1u, slice0 + q, lane);
#pragma unroll
          for (int q = 0; q < kLblResK; ++q)
            lbl_resolve_slice<W2>(acc_slot, cfg, d, (flagged >> q) & 1u, keys[q], it.frame, slice0 + q, lane, topdown,
                                  mask, height);
#endif
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);
    }
  }
}

// Lag (frames between the projection of a frame and its resolve) and ring slots in use for a launch of `grid` CTAs:
// the CTAs together hold grid * (slots + batch) tickets; the lag covers 1.5x the frames those span, + 1, within
// what the workspace (sized for `max_ring` slots) and the batch allow; ring = 2 * lag as in dm_project.cu when
// slots have to be reused (b > max_ring).
static void lbl_schedule(long long grid, long long tickets_per_frame, int b, int max_ring, int* lag, int* ring) {
  const double span = (double)grid * (kLblSlots + 2 * kLblBatch) / (double)(tickets_per_frame > 0 ? tickets_per_frame : 1);
  int l = (int)(span * 1.5) + 2;
  if (b <= max_ring) {  // a slot per frame: the lag is free (at most the batch)
    *lag = l < b ? l : b;
    *ring = max_ring;
    return;
  }
  if (l > max_ring / 2) l = max_ring / 2;
  if (l < 1) l = 1;
  *lag = l;
  *ring = 2 * l;
}

static bool lbl_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace dm

using namespace dm;

extern "C" size_t dm_orth_project_labels_workspace_bytes(const DmProjCfg* cfg, int32_t b) {
  if (!cfg || b <= 0 || cfg->Mh <= 0 || cfg->Mw <= 0 || cfg->C <= 0 || cfg->C > 63) return 0;
  return make_lbl_plan(*cfg, b).workspace_bytes();
}

extern "C" int dm_orth_project_labels_f32(const float* depth, const uint8_t* labels, const uint8_t* valid,
                                          const DmProjSample* samples, const DmProjCfg* cfg, int32_t b,
                                          float* topdown, uint8_t* mask, float* height, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
  DM_TRACE();
  if (!cfg || b < 0) return DM_EINVAL;
  if (b == 0) return DM_OK;
  if (!depth || !labels || !samples || !topdown || !mask || !workspace) return DM_EINVAL;
  if (cfg->H <= 0 || cfg->W <= 0 || cfg->Mh <= 0 || cfg->Mw <= 0 || cfg->C <= 0 || cfg->C > 63) return DM_EINVAL;
  if ((long long)cfg->H * cfg->W >= (1ll << 31) - 1024 || (long long)cfg->Mh * cfg->Mw >= (1ll << 29)) return DM_EINVAL;
  if (cfg->want_height && !height) return DM_EINVAL;
  if (cfg->reduction != 0 && cfg->reduction != 1) return DM_EINVAL;
  if (!lbl_aligned(workspace, 256)) return DM_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const LblPlan p = make_lbl_plan(*cfg, b);
  if (workspace_bytes < p.workspace_bytes()) return DM_EWORKSPACE;
  int dev = 0;
  DM_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return DM_EINVAL;
  if (take_timeout(dev)) return DM_ETIMEOUT;  // see dm_orth_project_f32
  static int sms[64] = {};
  if (!sms[dev]) DM_CUDA_OK(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
  const long long N = (long long)cfg->H * cfg->W, M = (long long)cfg->Mh * cfg->Mw;
  LblDims d{};
  d.C = cfg->C; d.CP = p.CP; d.ring = p.ring; d.lag = p.lag; d.nsl = p.nsl;
  d.hasH = cfg->want_height ? 1 : 0;
  d.vec_in = (N % 4 == 0) && lbl_aligned(depth, 16) && lbl_aligned(labels, 4) && (!valid || lbl_aligned(valid, 4));
  d.vec_out = (M % 4 == 0) && lbl_aligned(topdown, 16) && lbl_aligned(mask, 16) && (!d.hasH || lbl_aligned(height, 16));
  if (d.vec_out && M % 16 == 0) d.vec_out = 2;
  d.slot_words = p.slot_words;
  d.ws_words = (p.workspace_bytes() - p.ctrl_bytes) / 4;
  const long long per_frame = (N + kLblTile - 1) / kLblTile + (M + kLblResCells - 1) / kLblResCells;
  const long long tickets = (long long)(b + p.ring) * per_frame;  // the lag of a launch never exceeds the ring
  if (tickets >= (1ll << 31) - (1 << 20) || (unsigned long long)p.ring * p.slot_words >= (1ull << 31)) return DM_EINVAL;
  uint32_t* ctrl = static_cast<uint32_t*>(workspace);
  uint32_t* flags = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + p.ctrl_bytes);
  uint32_t* acc = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + p.ctrl_bytes + p.flag_bytes);
  void (*kern)(const float*, const uint8_t*, const uint8_t*, const DmProjSample*, DmProjCfg, LblDims, int, uint32_t*,
               uint32_t*, uint32_t*, float*, uint8_t*, float*, ProjGuard) = nullptr;
  const int fast = (cfg->fast_steps == 1 || cfg->fast_steps == 2) ? cfg->fast_steps : 0;
#define DM_LBL_PICK(W2)                                           \
  kern = fast == 2 ? proj_lbl_kernel<2, W2> : fast == 1 ? proj_lbl_kernel<1, W2> : proj_lbl_kernel<0, W2>;
  if (p.W2 == 1) { DM_LBL_PICK(1) } else { DM_LBL_PICK(2) }
#undef DM_LBL_PICK
  int per_sm = 0;
  DM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLblThreads, 0));
  if (per_sm < 1) return DM_EINVAL;
  long long grid = (long long)sms[dev] * per_sm;  // persistent: every CTA is resident (the dependency waits rely on it)
  if (grid > tickets) grid = tickets;
  lbl_schedule(grid, per_frame, b, p.ring, &d.lag, &d.ring);
  kern<<<(unsigned)grid, kLblThreads, 0, stream>>>(depth, labels, valid, samples, *cfg, d, b, ctrl, flags, acc, topdown,
                                                   mask, height, proj_guard(dev));
  DM_LAUNCHED();
  return DM_OK;
}
