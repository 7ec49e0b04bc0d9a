// MapBuilder merge: fuse_topdown_maps, /root/reference/dungeon_maps/maps.py:2181-2287.
//
// The reference turns every cell of every source map back into a 3-D point
// (height_map_to_point_cloud / map_dequantize, maps.py:547-612, 1021-1087), moves it into the
// target frame (maps.py:2059-2060, 2116-2117), takes ONE bounding box over all valid points of
// all samples and channels to size a fresh canvas (maps.py:2146-2179, host sync), re-quantises
// and scatter-maxes (maps.py:2232-2272).  Here that is two fused passes over the source cells,
// neither of which materialises a point: pass 1 reduces the bounding box, pass 2 scatters.
#include "dm_common.cuh"

namespace dm {

constexpr int kFuseThreads = 256;
constexpr int kMaxSources = 8;
constexpr int kGroup = 16;  // mask bytes examined per work item (one 128-bit load)

struct FuseSources {
  DmFuseSource s[kMaxSources];
  long long first_group[kMaxSources + 1];  // prefix of ceil(b*C*h*w / 16) per source
  long long cells[kMaxSources];            // b*C*h*w per source
  int vec_ok[kMaxSources];                 // mask base 16-byte aligned
  int n;
};

// Point of source cell (row r, col c), channel ch, sample smp, in the target frame.
__device__ __forceinline__ V3 source_point(const DmFuseSource& src, int smp, int ch, int r, int c) {
  // maps.py:1081-1086 map_dequantize
  float zb = (float)r;
  if (src.flip_h) zb = __fsub_rn((float)(src.h - 1), zb);
  V3 p;
  p.z = __fmul_rn(__fsub_rn(zb, src.height_offset[smp]), src.map_res);
  p.x = __fmul_rn(__fsub_rn((float)c, src.width_offset[smp]), src.map_res);
  p.y = src.height[(long long)smp * src.height_bstride + (long long)ch * src.height_cstride + (long long)r * src.w + c];
  p = apply_step(src.steps[smp * 2 + 0], p);
  p = apply_step(src.steps[smp * 2 + 1], p);
  return p;
}

// Walks the valid cells of one 16-byte group of a source's flat (b, C, h, w) mask.  Most groups of a
// world map are empty (every environment explored a corner of the batch-wide canvas): those cost one
// 128-bit load and nothing else, so the heights / values of empty regions are never read.
template <typename F>
__device__ __forceinline__ void for_valid_cells(const FuseSources& fs, int C, long long group, F&& visit) {
  int k = 0;
  while (k + 1 < fs.n && group >= fs.first_group[k + 1]) ++k;
  const DmFuseSource& src = fs.s[k];
  const long long i0 = (group - fs.first_group[k]) * kGroup;
  const long long left = fs.cells[k] - i0;
  // one bit per valid cell of the group
  uint32_t bits = 0;
  if (left >= kGroup && fs.vec_ok[k]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src.mask + i0));
    if ((v.x | v.y | v.z | v.w) == 0u) return;
    const uint32_t m[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < kGroup; ++j) bits |= ((m[j >> 2] >> ((j & 3) * 8)) & 0xffu) ? (1u << j) : 0u;
  } else {
    const int cnt = left < kGroup ? (int)left : kGroup;
    for (int j = 0; j < cnt; ++j) bits |= src.mask[i0 + j] ? (1u << j) : 0u;
    if (!bits) return;
  }
  // the loop below has ONE copy of the visitor (an unrolled 16-way version thrashed the instruction cache:
  // 7 stall_no_instruction cycles per issue in ncu)
  const int n = src.h * src.w;
  int smp0, ch0, cell0;
  if (fs.cells[k] < (1ll << 31)) {  // 32-bit divisions whenever the source allows it
    const unsigned sc = (unsigned)i0 / (unsigned)n;
    cell0 = (int)((unsigned)i0 - sc * (unsigned)n);
    smp0 = (int)(sc / (unsigned)C);
    ch0 = (int)(sc - (unsigned)smp0 * (unsigned)C);
  } else {
    const long long sc = i0 / n;
    cell0 = (int)(i0 - sc * n);
    smp0 = (int)(sc / C);
    ch0 = (int)(sc - (long long)smp0 * C);
  }
  const int r0 = cell0 / src.w, c0 = cell0 - r0 * src.w;
  while (bits) {
    const int j = __ffs(bits) - 1;
    bits &= bits - 1;
    int smp = smp0, ch = ch0, cell = cell0 + j, r = r0, c = c0 + j;
    if (cell >= n) {  // the group straddles two planes (n is not a multiple of 16)
      cell -= n;
      if (++ch == C) { ch = 0; ++smp; }
      r = cell / src.w; c = cell - r * src.w;
    } else {
      while (c >= src.w) { c -= src.w; ++r; }
    }
    visit(src, i0 + j, smp, ch, r, c);
  }
}

__global__ void fuse_bbox_init(long long* out) {
  out[0] = 0x7fffffffffffffffLL;          // min_x
  out[1] = (long long)0x8000000000000000ULL;  // max_x
  out[2] = 0x7fffffffffffffffLL;          // min_z
  out[3] = (long long)0x8000000000000000ULL;  // max_z
  out[4] = 0;                              // n_valid
}

__global__ void __launch_bounds__(kFuseThreads)
fuse_bbox_kernel(const __grid_constant__ FuseSources fs, int C, float res, long long total_groups,
                 long long* __restrict__ out) {
  long long mnx = 0x7fffffffffffffffLL, mxx = (long long)0x8000000000000000ULL;
  long long mnz = mnx, mxz = mxx;
  unsigned long long cnt = 0;
  for (long long g = (long long)blockIdx.x * kFuseThreads + threadIdx.x; g < total_groups;
       g += (long long)gridDim.x * kFuseThreads) {
    for_valid_cells(fs, C, g, [&](const DmFuseSource& src, long long, int smp, int ch, int r, int c) {
      const V3 p = source_point(src, smp, ch, r, c);
      // maps.py:2159-2165: map_quantize(width_offset=0., height_offset=0., flip_h=False)
      float xf, zf;
      quantize_f(p.x, p.z, 0.0f, 0.0f, res, 0, 0, &xf, &zf);
      const long long xi = f2i64(xf), zi = f2i64(zf);
      mnx = xi < mnx ? xi : mnx; mxx = xi > mxx ? xi : mxx;
      mnz = zi < mnz ? zi : mnz; mxz = zi > mxz ? zi : mxz;
      ++cnt;
    });
  }
  // warp then block reduction, one atomic per block per quantity
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, mnx, o), b2 = __shfl_xor_sync(0xffffffffu, mxx, o);
    const long long c2 = __shfl_xor_sync(0xffffffffu, mnz, o), d2 = __shfl_xor_sync(0xffffffffu, mxz, o);
    const unsigned long long e2 = __shfl_xor_sync(0xffffffffu, cnt, o);
    mnx = a < mnx ? a : mnx; mxx = b2 > mxx ? b2 : mxx;
    mnz = c2 < mnz ? c2 : mnz; mxz = d2 > mxz ? d2 : mxz;
    cnt += e2;
  }
  __shared__ long long sm[kFuseThreads / 32][4];
  __shared__ unsigned long long sc[kFuseThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sm[warp][0] = mnx; sm[warp][1] = mxx; sm[warp][2] = mnz; sm[warp][3] = mxz; sc[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kFuseThreads / 32; ++w) {
      mnx = sm[w][0] < mnx ? sm[w][0] : mnx; mxx = sm[w][1] > mxx ? sm[w][1] : mxx;
      mnz = sm[w][2] < mnz ? sm[w][2] : mnz; mxz = sm[w][3] > mxz ? sm[w][3] : mxz;
      cnt += sc[w];
    }
    if (cnt) {
      atomicMin(out + 0, mnx); atomicMax(out + 1, mxx);
      atomicMin(out + 2, mnz); atomicMax(out + 3, mxz);
      atomicAdd(reinterpret_cast<unsigned long long*>(out + 4), cnt);
    }
  }
}

// Fresh canvases: topdown = fill (utils.py:472-473), height = -inf (maps.py:2268), mask = false.
// Consecutive lanes store consecutive 16-byte words (a thread owning 64 contiguous bytes made every warp
// store touch 32 different lines: 45 % of DRAM peak in ncu); unaligned bases / the tail take the scalar loop.
__global__ void __launch_bounds__(kFuseThreads)
fuse_fill_kernel(float* __restrict__ topdown, float* __restrict__ height, uint8_t* __restrict__ mask, long long n,
                 float fill, int vec_ok) {
  const long long n16 = vec_ok ? n / 16 : 0;
  const long long tid = (long long)blockIdx.x * kFuseThreads + threadIdx.x, nth = (long long)gridDim.x * kFuseThreads;
  const float4 f4 = make_float4(fill, fill, fill, fill);
  const float4 h4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (long long i = tid; i < n16 * 4; i += nth) st_stream_f4(topdown + i * 4, f4);
  if (height)
    for (long long i = tid; i < n16 * 4; i += nth) st_stream_f4(height + i * 4, h4);
  for (long long i = tid; i < n16; i += nth) *reinterpret_cast<uint4*>(mask + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  for (long long i = n16 * 16 + tid; i < n; i += nth) {
    topdown[i] = fill;
    if (height) height[i] = -INFINITY;
    mask[i] = 0;
  }
}

// mask_inline: the mask (utils.py:489-491: the cell differs from what the canvas was filled with) is
// stored right where a value beats `fill`; with a NaN fill the generic pass below is used instead.
__global__ void __launch_bounds__(kFuseThreads)
fuse_scatter_kernel(const __grid_constant__ FuseSources fs, int C, const DmFuseTarget tgt, long long total_groups,
                    float* __restrict__ topdown, float* __restrict__ height, uint8_t* __restrict__ mask,
                    int mask_inline) {
  const long long M = (long long)tgt.Mh * tgt.Mw;
  for (long long g = (long long)blockIdx.x * kFuseThreads + threadIdx.x; g < total_groups;
       g += (long long)gridDim.x * kFuseThreads) {
    for_valid_cells(fs, C, g, [&](const DmFuseSource& src, long long in_idx, int smp, int ch, int r, int c) {
      const V3 p = source_point(src, smp, ch, r, c);
      float xf, zf;  // maps.py:2232-2238
      quantize_f(p.x, p.z, tgt.width_offset, tgt.height_offset, tgt.map_res, tgt.Mh, tgt.flip_h, &xf, &zf);
      if (!(xf >= 0.0f && xf < (float)tgt.Mw && zf >= 0.0f && zf < (float)tgt.Mh)) return;
      const long long o = ((long long)smp * C + ch) * M + (long long)zf * tgt.Mw + (long long)xf;
      const float v = src.values ? src.values[in_idx] : p.y;  // maps.py:2214-2216
      if (v == v) {
        if (tgt.reduction) atomic_min_f32(topdown + o, v); else atomic_max_f32(topdown + o, v);
        if (mask_inline && better(v, tgt.fill_value, tgt.reduction)) mask[o] = 1;
      }
      if (height && p.y == p.y) atomic_max_f32(height + o, p.y);  // maps.py:2258-2271
    });
  }
}

__global__ void __launch_bounds__(kFuseThreads)
changed_mask_kernel(const float* __restrict__ canvas, long long n, float fill, uint8_t* __restrict__ mask) {
  for (long long i = (long long)blockIdx.x * kFuseThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kFuseThreads) {
    float d = fabsf(__fsub_rn(canvas[i], fill));  // utils.py:489-491
    if (d != d) d = 0.0f;
    mask[i] = d != 0.0f;
  }
}

static int pack_sources(const DmFuseSource* sources, int n, int b, int C, FuseSources* fs, long long* total) {
  if (!sources || n <= 0 || n > kMaxSources || b <= 0 || C <= 0) return DM_EINVAL;
  fs->n = n;
  long long acc = 0;
  for (int i = 0; i < n; ++i) {
    const DmFuseSource& s = sources[i];
    if (!s.height || !s.mask || !s.width_offset || !s.height_offset || !s.steps || s.h <= 0 || s.w <= 0)
      return DM_EINVAL;
    if ((long long)s.h * s.w >= (1ll << 31)) return DM_EINVAL;
    fs->s[i] = s;
    fs->first_group[i] = acc;
    fs->cells[i] = (long long)b * C * s.h * s.w;
    fs->vec_ok[i] = reinterpret_cast<uintptr_t>(s.mask) % 16 == 0;
    acc += (fs->cells[i] + kGroup - 1) / kGroup;
  }
  fs->first_group[n] = acc;
  *total = acc;
  return DM_OK;
}

static unsigned grid_for(long long items) {
  long long blocks = (items + kFuseThreads - 1) / kFuseThreads;
  const long long cap = (long long)kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace dm

using namespace dm;

extern "C" int dm_fuse_bbox_i64(const DmFuseSource* sources, int32_t n_sources, int32_t b, int32_t C,
                                float target_res, int64_t* out, void* stream_) {
  if (!out) return DM_EINVAL;
  FuseSources fs;
  long long total = 0;
  const int rc = pack_sources(sources, n_sources, b, C, &fs, &total);
  if (rc != DM_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  fuse_bbox_init<<<1, 1, 0, stream>>>(reinterpret_cast<long long*>(out));
  DM_LAUNCHED();
  fuse_bbox_kernel<<<grid_for(total), kFuseThreads, 0, stream>>>(fs, C, target_res, total,
                                                                  reinterpret_cast<long long*>(out));
  DM_LAUNCHED();
  return DM_OK;
}

extern "C" int dm_fuse_scatter_f32(const DmFuseSource* sources, int32_t n_sources, int32_t b, int32_t C,
                                   const DmFuseTarget* target, float* topdown, uint8_t* mask, float* height,
                                   void* stream_) {
  if (!target || !topdown || !mask || target->Mh <= 0 || target->Mw <= 0) return DM_EINVAL;
  if (target->reduction != 0 && target->reduction != 1) return DM_EINVAL;
  FuseSources fs;
  long long total = 0;
  const int rc = pack_sources(sources, n_sources, b, C, &fs, &total);
  if (rc != DM_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long n_out = (long long)b * C * target->Mh * target->Mw;
  const int vec_ok = reinterpret_cast<uintptr_t>(topdown) % 16 == 0 && reinterpret_cast<uintptr_t>(mask) % 16 == 0 &&
                     (!height || reinterpret_cast<uintptr_t>(height) % 16 == 0);
  const int mask_inline = target->fill_value == target->fill_value;  // not NaN
  fuse_fill_kernel<<<grid_for((n_out + 3) / 4), kFuseThreads, 0, stream>>>(topdown, height, mask, n_out,
                                                                             target->fill_value, vec_ok);
  DM_LAUNCHED();
  fuse_scatter_kernel<<<grid_for(total), kFuseThreads, 0, stream>>>(fs, C, *target, total, topdown, height, mask,
                                                                    mask_inline);
  DM_LAUNCHED();
  if (!mask_inline) {
    changed_mask_kernel<<<grid_for(n_out), kFuseThreads, 0, stream>>>(topdown, n_out, target->fill_value, mask);
    DM_LAUNCHED();
  }
  return DM_OK;
}

// Opt-in fixed-canvas merge (no reference equivalent of the call; the semantics are those of the
// reference's project(..., canvas=, canvas_masks=), maps.py:1089-1173 / utils.py:462-491): the valid
// cells of the sources are max-merged IN PLACE into canvases that already hold a world map.  No
// bounding box, no host sync, no reallocation: one launch.  Invariant kept: mask == (cell != fill).
extern "C" int dm_fuse_inplace_f32(const DmFuseSource* sources, int32_t n_sources, int32_t b, int32_t C,
                                   const DmFuseTarget* target, float* topdown, uint8_t* mask, float* height,
                                   void* stream_) {
  if (!target || !topdown || !mask || target->Mh <= 0 || target->Mw <= 0) return DM_EINVAL;
  if (target->reduction != 0 && target->reduction != 1) return DM_EINVAL;
  if (!(target->fill_value == target->fill_value)) return DM_EINVAL;  // NaN fill has no in-place mask rule
  FuseSources fs;
  long long total = 0;
  const int rc = pack_sources(sources, n_sources, b, C, &fs, &total);
  if (rc != DM_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  fuse_scatter_kernel<<<grid_for(total), kFuseThreads, 0, stream>>>(fs, C, *target, total, topdown, height, mask, 1);
  DM_LAUNCHED();
  return DM_OK;
}

/* Fills fresh world canvases for dm_fuse_inplace_f32: topdown = fill_value, height = -inf (may be NULL), mask = 0. */
extern "C" int dm_fuse_canvas_init_f32(float* topdown, uint8_t* mask, float* height, int64_t n, float fill_value,
                                       void* stream_) {
  if (!topdown || !mask || n < 0) return DM_EINVAL;
  if (n == 0) return DM_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int vec_ok = reinterpret_cast<uintptr_t>(topdown) % 16 == 0 && reinterpret_cast<uintptr_t>(mask) % 16 == 0 &&
                     (!height || reinterpret_cast<uintptr_t>(height) % 16 == 0);
  fuse_fill_kernel<<<grid_for((n + 3) / 4), kFuseThreads, 0, stream>>>(topdown, height, mask, n, fill_value, vec_ok);
  DM_LAUNCHED();
  return DM_OK;
}
