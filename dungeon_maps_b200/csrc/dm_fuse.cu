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

struct FuseSources {
  DmFuseSource s[kMaxSources];
  long long first_item[kMaxSources + 1];  // prefix of b*C*h*w per source
  int n;
};

// Point of source cell `cell` (row-major in h×w), channel ch, sample smp, in the target frame.
__device__ __forceinline__ V3 source_point(const DmFuseSource& src, int smp, int ch, int cell) {
  const int r = cell / src.w, c = cell - r * src.w;
  // maps.py:1081-1086 map_dequantize
  float zb = (float)r;
  if (src.flip_h) zb = __fsub_rn((float)(src.h - 1), zb);
  V3 p;
  p.z = __fmul_rn(__fsub_rn(zb, src.height_offset[smp]), src.map_res);
  p.x = __fmul_rn(__fsub_rn((float)c, src.width_offset[smp]), src.map_res);
  p.y = src.height[(long long)smp * src.height_bstride + (long long)ch * src.height_cstride + cell];
  p = apply_step(src.steps[smp * 2 + 0], p);
  p = apply_step(src.steps[smp * 2 + 1], p);
  return p;
}

__device__ __forceinline__ bool locate(const FuseSources& fs, long long item, int C, int* si, int* smp,
                                       int* ch, int* cell) {
  int k = 0;
  while (k < fs.n && item >= fs.first_item[k + 1]) ++k;
  if (k >= fs.n) return false;
  const DmFuseSource& src = fs.s[k];
  const long long local = item - fs.first_item[k];
  const int n = src.h * src.w;
  const long long sc = local / n;
  *cell = (int)(local - sc * n);
  *smp = (int)(sc / C);
  *ch = (int)(sc - (long long)(*smp) * C);
  *si = k;
  return true;
}

__global__ void fuse_bbox_init(long long* out) {
  out[0] = 0x7fffffffffffffffLL;          // min_x
  out[1] = (long long)0x8000000000000000ULL;  // max_x
  out[2] = 0x7fffffffffffffffLL;          // min_z
  out[3] = (long long)0x8000000000000000ULL;  // max_z
  out[4] = 0;                              // n_valid
}

__global__ void __launch_bounds__(kFuseThreads)
fuse_bbox_kernel(const FuseSources fs, int C, float res, long long total, long long* __restrict__ out) {
  long long mnx = 0x7fffffffffffffffLL, mxx = (long long)0x8000000000000000ULL;
  long long mnz = mnx, mxz = mxx;
  unsigned long long cnt = 0;
  for (long long item = (long long)blockIdx.x * kFuseThreads + threadIdx.x; item < total;
       item += (long long)gridDim.x * kFuseThreads) {
    int si, smp, ch, cell;
    if (!locate(fs, item, C, &si, &smp, &ch, &cell)) continue;
    const DmFuseSource& src = fs.s[si];
    if (!src.mask[((long long)smp * C + ch) * src.h * src.w + cell]) continue;
    const V3 p = source_point(src, smp, ch, cell);
    // maps.py:2159-2165: map_quantize(width_offset=0., height_offset=0., flip_h=False)
    float xf, zf;
    quantize_f(p.x, p.z, 0.0f, 0.0f, res, 0, 0, &xf, &zf);
    const long long xi = f2i64(xf), zi = f2i64(zf);
    mnx = xi < mnx ? xi : mnx; mxx = xi > mxx ? xi : mxx;
    mnz = zi < mnz ? zi : mnz; mxz = zi > mxz ? zi : mxz;
    ++cnt;
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

__global__ void __launch_bounds__(kFuseThreads)
fuse_fill_kernel(float* __restrict__ topdown, float* __restrict__ height, long long n, float fill) {
  for (long long i = (long long)blockIdx.x * kFuseThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kFuseThreads) {
    topdown[i] = fill;                       // utils.py:472-473
    if (height) height[i] = -INFINITY;       // maps.py:2268
  }
}

__global__ void __launch_bounds__(kFuseThreads)
fuse_scatter_kernel(const FuseSources fs, int C, const DmFuseTarget tgt, long long total,
                    float* __restrict__ topdown, float* __restrict__ height) {
  const long long M = (long long)tgt.Mh * tgt.Mw;
  for (long long item = (long long)blockIdx.x * kFuseThreads + threadIdx.x; item < total;
       item += (long long)gridDim.x * kFuseThreads) {
    int si, smp, ch, cell;
    if (!locate(fs, item, C, &si, &smp, &ch, &cell)) continue;
    const DmFuseSource& src = fs.s[si];
    const long long in_idx = ((long long)smp * C + ch) * src.h * src.w + cell;
    if (!src.mask[in_idx]) continue;
    const V3 p = source_point(src, smp, ch, cell);
    float xf, zf;  // maps.py:2232-2238
    quantize_f(p.x, p.z, tgt.width_offset, tgt.height_offset, tgt.map_res, tgt.Mh, tgt.flip_h, &xf, &zf);
    if (!(xf >= 0.0f && xf < (float)tgt.Mw && zf >= 0.0f && zf < (float)tgt.Mh)) continue;
    const long long o = ((long long)smp * C + ch) * M + (long long)zf * tgt.Mw + (long long)xf;
    const float v = src.values ? src.values[in_idx] : p.y;  // maps.py:2214-2216
    if (v == v) {
      if (tgt.reduction) atomic_min_f32(topdown + o, v); else atomic_max_f32(topdown + o, v);
    }
    if (height && p.y == p.y) atomic_max_f32(height + o, p.y);  // maps.py:2258-2271
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
    fs->s[i] = s;
    fs->first_item[i] = acc;
    acc += (long long)b * C * s.h * s.w;
  }
  fs->first_item[n] = acc;
  *total = acc;
  return DM_OK;
}

static unsigned grid_for(long long items) {
  long long blocks = (items + kFuseThreads - 1) / kFuseThreads;
  const long long cap = (long long)kNumSMs * 8 * 2;
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
  fuse_fill_kernel<<<grid_for(n_out), kFuseThreads, 0, stream>>>(topdown, height, n_out, target->fill_value);
  DM_LAUNCHED();
  fuse_scatter_kernel<<<grid_for(total), kFuseThreads, 0, stream>>>(fs, C, *target, total, topdown, height);
  DM_LAUNCHED();
  changed_mask_kernel<<<grid_for(n_out), kFuseThreads, 0, stream>>>(topdown, n_out, target->fill_value, mask);
  DM_LAUNCHED();
  return DM_OK;
}
