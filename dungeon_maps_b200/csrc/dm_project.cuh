// Device-side pieces shared by the projection kernels (dm_project.cu: float value planes, dm_labels.cu: class-id
// planes): constants of the sparse accumulation ring, the per-pixel cell arithmetic, ticket decoding, the
// dependency wait and the PTX wrappers.  See dm_project.cu for the design notes.
#pragma once

#include "dm_common.cuh"

namespace dm {

constexpr int kSliceCells = 64;                  // cells per sparse-ring flag (one resolve slice of a warp)
#ifndef DM_FLAG_STRIDE
#define DM_FLAG_STRIDE 8
#endif
constexpr int kFlagStride = DM_FLAG_STRIDE;      // words between two flags: every tile of a frame stores into the
                                                 // same few hundred flags, so each gets its own 32-byte sector
constexpr int kCtrlWords = 512;                  // control block at the head of the workspace
constexpr unsigned long long kSpinLimitNs = 4000000000ull;  // dependency wait guard (bug → no hang)
#ifndef DM_SUSPEND_NS
#define DM_SUSPEND_NS 20000u
#endif

// Guard of the persistent kernels' cross-CTA dependency waits, passed by value with every launch.  A wait that
// exceeds spin_ns raises the sticky word 2 of the workspace's control block and the item is skipped; the LAST CTA of
// the launch then re-zeroes the whole workspace (so later calls on it are sound again), and raises *status — a word of
// mapped pinned host memory, one per device — which the next entry on that device, dm_device_status() and the
// final synchronisation of the *_host entries turn into DM_ETIMEOUT (dm_api.cu).
struct ProjGuard {
  unsigned long long spin_ns;
  uint32_t dep_bias;   // test hook (dm_debug_set_wait_guard): added to every dependency target; 0 in production
  uint32_t* status;    // device pointer of the mapped status word, or nullptr
};
ProjGuard proj_guard(int device);  // dm_api.cu
bool take_timeout(int device);     // dm_api.cu: reads and clears the device's status word

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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The suspend-time hint lets the hardware park the warp until the phase completes (or the hint expires)
// instead of re-issuing the try_wait every few cycles: waiting warps must not eat the issue slots of working ones.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "DM_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DM_DONE;\n\t"
      "bra DM_WAIT;\n\t"
      "DM_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(DM_SUSPEND_NS) : "memory");
}
// TMA bulk copy global → shared, completion on an mbarrier, L2 evict-first (inputs are read once).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// TMA tensor copy global → shared (tile mode, 3-D box at element coordinates (x, y, z) of the tensor map `tm`, a
// __grid_constant__ kernel parameter), completion on an mbarrier.  Out-of-bounds elements arrive as zeros.
__device__ __forceinline__ void tensor_g2s_3d(void* dst, const void* tm, int x, int y, int z, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3, %4}], [%5], %6;"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// thread 0 only.  Returns false on timeout (a scheduling bug, never expected) after raising ctrl[2].
__device__ __forceinline__ bool wait_count(const uint32_t* counter, uint32_t target, uint32_t* ctrl,
                                           unsigned long long spin_ns) {
  if (ld_acquire(counter) >= target) return true;
  const unsigned long long t0 = globaltimer();
  while (ld_acquire(counter) < target) {
    __nanosleep(64);
    if (globaltimer() - t0 > spin_ns) {
      atomicExch(ctrl + 2, 1u);
      return false;
    }
  }
  return true;
}

// Release of a per-frame completion counter: everything this CTA did for the item (its REDs into the ring, its
// key re-zeroing and output stores, ordered before this thread by the CTA-level hand-off) is visible at gpu scope
// before the increment is.  `red.release` compiles to MEMBAR.ALL.GPU + RED; the fence.acq_rel.gpu + relaxed RED it
// replaces added a CCTL.IVALL (L1 invalidate) per ticket (ncu r01l: 53 648 per launch).
__device__ __forceinline__ void red_release_add1(uint32_t* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// Last CTA of a persistent launch, called by `nthreads` threads (tid = 0 .. nthreads - 1) once every other CTA has
// left: when a dependency wait timed out (ctrl[2]) the flags and the ring are re-zeroed — the skipped items left
// keys behind — and the host is told.  `words` = workspace words behind the control block.
__device__ __forceinline__ void scrub_after_timeout(uint32_t* ctrl, uint32_t* flags, unsigned long long words, int tid,
                                                    int nthreads, uint32_t* status) {
  if (__ldcg(ctrl + 2) == 0u) return;
  uint4* w4 = reinterpret_cast<uint4*>(flags);  // 256-byte aligned, a multiple of 16 bytes long
  for (unsigned long long i = tid; i < words / 4; i += nthreads) __stcg(w4 + i, make_uint4(0u, 0u, 0u, 0u));
  __threadfence();
  if (tid == 0 && status) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(status), "r"(1u) : "memory");
    __threadfence_system();
  }
}

enum { kItemProj = 0, kItemResolve = 1, kItemNone = 2, kItemExit = 3 };

// ticket → work item.  Step s holds the P projection tiles of frame s and the R resolve tiles
// of frame s - lag, interleaved evenly so HBM reads and writes mix.
__device__ __forceinline__ void decode_ticket(unsigned t, int b, int P, int R, int lag, int* kind, int* frame,
                                              int* idx) {
  const unsigned per = (unsigned)(P + R);
  const int s = (int)(t / per);
  const int j = (int)(t - (unsigned)s * per);
  // the R resolve tickets are spread evenly among the P projection tickets of the step (HBM reads and writes
  // stay mixed whatever the ratio): ticket j is a resolve ticket when floor((j + 1) R / per) steps up
  const unsigned rb0 = (unsigned)(((unsigned long long)j * (unsigned)R) / per);
  const unsigned rb1 = (unsigned)(((unsigned long long)(j + 1) * (unsigned)R) / per);
  if (rb1 > rb0) {
    *kind = kItemResolve;
    *idx = (int)rb0;
  } else {
    *kind = kItemProj;
    *idx = j - (int)rb0;
  }
  *frame = *kind == kItemProj ? s : s - lag;
  if (*frame < 0 || *frame >= b) *kind = kItemNone;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// sparse-ring slice flag: plain idempotent store, published with the tile's REDs
__device__ __forceinline__ void st_flag(uint32_t* p) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(1u) : "memory");
}
// fire-and-forget reduction (RED, never the returning ATOM form)
__device__ __forceinline__ void red_max_u32(uint32_t* p, uint32_t v) {
#ifdef DM_ABL_NORED  // ablation build (wrong results): what the kernel costs without its REDs
  asm volatile("" ::"l"(p), "r"(v) : "memory");
#else
  asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
// B2's RED sites: "the run that ends here (cell ca, next entry's cell cb differs) holds a value that beats the fill".
// Both tests and the RED are predicated inside one asm block: no branch, no reconvergence bookkeeping, and the two
// half-warp streams of a list no longer diverge at every site (ncu r02h: 9 instructions and a BSSY / BSYNC pair per site).
template <bool IS_MIN>
__device__ __forceinline__ void red_key_if_run_ends(uint32_t ca, uint32_t cb, float v, float fill, uint32_t* p) {
  const uint32_t key = IS_MIN ? ~enc(v) : enc(v);
#ifdef DM_ABL_NORED
  asm volatile("" ::"r"(ca), "r"(cb), "f"(v), "f"(fill), "l"(p), "r"(key) : "memory");
#elif defined(DM_B2_BRANCHFREE)  // every site issues: a RED of key 0 changes nothing (the address must be a real cell)
  {
    const bool on = (ca != cb) && (IS_MIN ? (v < fill) : (v > fill));
    asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(on ? key : 0u) : "memory");
  }
#elif defined(DM_ABL_REDVAR)  // ablation builds (wrong results): what exactly the REDs cost
  {
    const int ln = threadIdx.x & 31;
    bool on = (ca != cb) && (IS_MIN ? (v < fill) : (v > fill));
#if DM_ABL_REDVAR == 2      // half the lanes
    on = on && ((ln & 1) == 0);
#elif DM_ABL_REDVAR == 4    // one lane per run
    on = on && ((ln & 15) == 0);
#endif
#if DM_ABL_REDVAR == 3      // every RED lands in a 1 MB window (L2-resident, no new lines)
    p = reinterpret_cast<uint32_t*>((reinterpret_cast<unsigned long long>(p) & ~0xfffffffull) | (reinterpret_cast<unsigned long long>(p) & 0xffffcull));
#endif
#if DM_ABL_REDVAR == 1      // plain stores instead of reductions
    if (on) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(key) : "memory");
#else
    if (on) asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(key) : "memory");
#endif
  }
#else
  if (IS_MIN)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %0, %1;\n\t"
        "setp.lt.and.f32 p, %2, %3, p;\n\t"
        "@p red.relaxed.gpu.global.max.u32 [%4], %5;\n\t}" ::"r"(ca), "r"(cb), "f"(v), "f"(fill), "l"(p), "r"(key) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %0, %1;\n\t"
        "setp.gt.and.f32 p, %2, %3, p;\n\t"
        "@p red.relaxed.gpu.global.max.u32 [%4], %5;\n\t}" ::"r"(ca), "r"(cb), "f"(v), "f"(fill), "l"(p), "r"(key) : "memory");
#endif
}
// Predicated reduction: no branch, no reconvergence bookkeeping around the RED.
__device__ __forceinline__ void red_max_u32_if(bool pred, uint32_t* p, uint32_t v) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %0, 0;\n\t"
      "@p red.global.max.u32 [%1], %2;\n\t}" ::"r"((int)pred), "l"(p), "r"(v) : "memory");
}

// One pixel on the straight-line path (cfg.fast_steps): xn = rn(rn(c - cx) / fx), yn likewise.
// x / res with one IEEE rounding, through rres = rn(1 / res): product, then two exact-residual corrections
// (Markstein).  Equal to __fdiv_rn(x, res) whenever the quotient is a normal number (checked exhaustively-ish on
// the host, 4e9 random operands over 200 divisors: no mismatch); a non-finite x gives NaN instead of inf, a
// quotient in the denormal range may differ in its last bit — neither can change a bin: the pixel is off the
// map either way, or the difference vanishes in the offset addition that follows (maps.py:1004-1013).
__device__ __forceinline__ float div_by_rcp(float x, float res, float rres) {
  const float q0 = __fmul_rn(x, rres);
  const float q1 = __fmaf_rn(__fmaf_rn(-q0, res, x), rres, q0);
  return __fmaf_rn(__fmaf_rn(-q1, res, x), rres, q1);
}

template <bool GLOBAL, bool RCP_DIV = false>
__device__ __forceinline__ int pixel_cell_fast(const DmProjCfg& cfg, const DmProjSample& sp, float xn, float yn,
                                               float z, bool ok, float* y_out, float rres = 0.0f) {
  if (cfg.has_trunc_depth_max) ok = ok && (z <= cfg.trunc_depth_max);
  if (cfg.has_trunc_depth_min) ok = ok && (z >= cfg.trunc_depth_min);
  const float X = __fmul_rn(xn, z), Y = __fmul_rn(yn, z);
  // pitch about x: R[0] = 1, R[1] = R[2] = R[3] = R[6] = 0 (x passes through), then + (0, h, 0)
  const float* Rl = sp.to_local.R;
  float ly = __fmaf_rn(Rl[7], z, __fmul_rn(Rl[4], Y));
  float lz = __fmaf_rn(Rl[8], z, __fmul_rn(Rl[5], Y));
  ly = __fadd_rn(ly, sp.to_local.t[1]);
  if (cfg.has_trunc_height_max) ok = ok && (ly <= cfg.trunc_height_max);
  float gx = X, gz = lz;
  if (GLOBAL) {  // yaw about y: R[4] = 1, R[1] = R[3] = R[5] = R[7] = 0 (y passes through), + (x, 0, z)
    const float* Rg = sp.to_global.R;
    gx = __fadd_rn(__fmaf_rn(Rg[6], lz, __fmul_rn(Rg[0], X)), sp.to_global.t[0]);
    gz = __fadd_rn(__fmaf_rn(Rg[8], lz, __fmul_rn(Rg[2], X)), sp.to_global.t[2]);
  }
  float xf, zf;
  if (RCP_DIV) {
    const float res = cfg.map_res;
    const float xb = __fadd_rn(div_by_rcp(gx, res, rres), sp.width_offset);
    float zb = __fadd_rn(div_by_rcp(gz, res, rres), sp.height_offset);
    if (cfg.flip_h) zb = __fsub_rn((float)(cfg.Mh - 1), zb);
    xf = floorf(__fadd_rn(xb, 0.5f));
    zf = floorf(__fadd_rn(zb, 0.5f));
  } else {
    quantize_f(gx, gz, sp.width_offset, sp.height_offset, cfg.map_res, cfg.Mh, cfg.flip_h, &xf, &zf);
  }
  ok = ok && (xf >= 0.0f) && (xf < (float)cfg.Mw) && (zf >= 0.0f) && (zf < (float)cfg.Mh);
  *y_out = ly;
  return ok ? ((int)zf * cfg.Mw + (int)xf) : -1;
}

template <bool IS_MIN>
__device__ __forceinline__ float red2(float a, float b) { return IS_MIN ? fminf(a, b) : fmaxf(a, b); }
template <bool IS_MIN>
__device__ __forceinline__ bool beats(float v, float fill) { return IS_MIN ? (v < fill) : (v > fill); }
template <bool IS_MIN>
__device__ __forceinline__ uint32_t key_of(float v) { return IS_MIN ? ~enc(v) : enc(v); }

struct Rcps {
  float res, fx, fy;  // rn(1 / map_res), rn(1 / fx), rn(1 / fy)
};

}  // namespace dm
