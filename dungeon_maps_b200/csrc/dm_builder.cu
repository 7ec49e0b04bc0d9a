// MapBuilder.step with its host side in C: plot (orth_project with get_height_map, maps.py:2408-2469) + merge
// (fuse_topdown_maps, maps.py:2181-2287) for the case a mapping loop runs thousands of times — height maps of b
// environments, world map in the global frame.  The Python layer made ~1000 interpreter calls per step for a few
// hundred bytes of parameters (0.73 ms per step around 0.25 ms of merge kernels, profiles/r01n); here a step is two
// C calls: every parameter block (projection samples, the two fuse sources) is packed into ONE pinned staging slot,
// uploaded with ONE copy, and the kernels are launched back to back.  Nothing about the arithmetic changes: the
// rotation matrices are formed from sin / cos values the caller computes with the reference's own torch-CPU ops
// (utils.py:303-327; the last ulp matters), by the same float32 operations in the same order.
#include <cmath>
#include <cstring>
#include <new>

#include "dm_common.cuh"

namespace dm {
namespace {

constexpr int kParamSlots = 4;
constexpr int kStepWords = sizeof(DmStep) / 4;           // 16
constexpr int kSampleWords = sizeof(DmProjSample) / 4;   // 48
constexpr int kFuseWords = 2 * kStepWords + 2;           // per sample: two steps, width offset, height offset

}  // namespace
}  // namespace dm

using namespace dm;

struct DmBuilder {
  DmBuilderCfg cfg;
  int device = 0;
  DmPoseCfg pose;      // the constants of the yaw / pitch steps, fused as the projection's rotations are
  int fused_fuse = 0;  // DmStep.fused of the local map's local → global step in the merge
  int local_fast = 0;  // the pitch step has the structure DmProjCfg.fast_steps >= 1 promises
  char* h_params[kParamSlots] = {};
  char* d_params[kParamSlots] = {};
  cudaEvent_t ev[kParamSlots] = {};
  size_t param_bytes = 0;
  unsigned next = 0;
  int64_t* d_bbox = nullptr;
  int64_t* h_bbox = nullptr;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // what dm_builder_plot leaves for dm_builder_merge
  DmFuseSource src[2];
  int n_src = 0;
  cudaEvent_t ev_bbox = nullptr;     // the bounding box has reached the host
  float* prefilled_top = nullptr;    // canvases dm_builder_plot_prefill filled while the host waited for the box
  uint8_t* prefilled_mask = nullptr;
  int64_t prefilled_cells = 0;
};

static void builder_free(DmBuilder* h) {
  if (!h) return;
  for (int i = 0; i < kParamSlots; ++i) {
    if (h->h_params[i]) cudaFreeHost(h->h_params[i]);
    if (h->d_params[i]) cudaFree(h->d_params[i]);
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  }
  if (h->ev_bbox) cudaEventDestroy(h->ev_bbox);
  if (h->d_bbox) cudaFree(h->d_bbox);
  if (h->h_bbox) cudaFreeHost(h->h_bbox);
  if (h->ws) cudaFree(h->ws);
  delete h;
}

extern "C" int dm_builder_create(const DmBuilderCfg* cfg, int32_t device, DmBuilder** out) {
  DM_TRACE();
  if (!cfg || !out || cfg->b <= 0 || cfg->proj.C != 0) return DM_EINVAL;
  if (cfg->proj.H <= 0 || cfg->proj.W <= 0 || cfg->proj.Mh <= 0 || cfg->proj.Mw <= 0) return DM_EINVAL;
  if (cfg->merge_reduction != 0 && cfg->merge_reduction != 1) return DM_EINVAL;
  DmBuilder* h = new (std::nothrow) DmBuilder();
  if (!h) return DM_EINVAL;
  h->cfg = *cfg;
  h->device = device;
  const long long n_proj = (long long)cfg->proj.H * cfg->proj.W;            // points one bmm rotates (utils.py:329)
  const long long n_fuse = (long long)cfg->proj.Mh * cfg->proj.Mw;          // C * h * w points of the local map
  memset(&h->pose, 0, sizeof(h->pose));
  memcpy(h->pose.pitch_R, cfg->pitch_R, sizeof(cfg->pitch_R));
  memcpy(h->pose.yaw_skew, cfg->yaw_skew, sizeof(cfg->yaw_skew));
  memcpy(h->pose.yaw_skew_sq, cfg->yaw_skew_sq, sizeof(cfg->yaw_skew_sq));
  h->pose.cam_height = cfg->cam_height;
  h->pose.fused = 9 * n_proj >= 400;
  h->fused_fuse = 9 * n_fuse >= 400;
  h->local_fast = local_step_is_fast(h->pose);
  h->param_bytes = ((size_t)cfg->b * (kSampleWords + 2 * kFuseWords) * 4 + 255) & ~(size_t)255;
  int rc = DM_OK;
#define DM_TRY(expr)                                                       \
  if (rc == DM_OK) {                                                       \
    const cudaError_t e_ = (expr);                                         \
    if (e_ != cudaSuccess) rc = static_cast<int>(e_);                      \
  }
  for (int i = 0; i < kParamSlots; ++i) {
    DM_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h->h_params[i]), h->param_bytes, cudaHostAllocDefault));
    DM_TRY(cudaMalloc(reinterpret_cast<void**>(&h->d_params[i]), h->param_bytes));
    DM_TRY(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
    if (!h->ev_bbox) DM_TRY(cudaEventCreateWithFlags(&h->ev_bbox, cudaEventDisableTiming));
  }
  DM_TRY(cudaMalloc(reinterpret_cast<void**>(&h->d_bbox), 5 * sizeof(int64_t)));
  DM_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h->h_bbox), 5 * sizeof(int64_t), cudaHostAllocDefault));
  h->ws_bytes = dm_orth_project_workspace_bytes(&cfg->proj, cfg->b);
  if (h->ws_bytes == 0 && rc == DM_OK) rc = DM_EINVAL;
  DM_TRY(cudaMalloc(&h->ws, h->ws_bytes));
  DM_TRY(cudaMemset(h->ws, 0, h->ws_bytes));
#undef DM_TRY
  if (rc != DM_OK) {
    builder_free(h);
    return rc;
  }
  *out = h;
  return DM_OK;
}

extern "C" void dm_builder_destroy(DmBuilder* h) { builder_free(h); }

// Packs one step's parameter blocks into a pinned slot and queues their upload; returns the slot's device base.
//   [DmProjSample x b | local fuse params (steps (b,2), woff (b), hoff (b)) | world fuse params (same layout)]
static int upload_params(DmBuilder* h, const float* pose, const float* sin_yaw, const float* cos_yaw,
                         const DmMapRef* world, cudaStream_t stream, char** d_base, int* fast_steps) {
  const DmBuilderCfg& c = h->cfg;
  const int b = c.b;
  const unsigned slot = h->next++ % kParamSlots;
  DM_CUDA_OK(cudaEventSynchronize(h->ev[slot]));  // the copy that last read this slot has run (it was queued long ago)
  float* w = reinterpret_cast<float*>(h->h_params[slot]);
  float* samples = w;
  float* lsteps = w + (size_t)b * kSampleWords;
  float* lwoff = lsteps + (size_t)b * 2 * kStepWords;
  float* lhoff = lwoff + b;
  float* wsteps = lhoff + b;
  float* wwoff = wsteps + (size_t)b * 2 * kStepWords;
  float* whoff = wwoff + b;
  int fast = h->local_fast ? (c.plot_to_global ? 2 : 1) : 0;
  for (int i = 0; i < b; ++i) {
    float Ry[9];
    yaw_matrix(h->pose, pose[3 * i + 2], sin_yaw[i], cos_yaw[i], Ry);  // utils.py:303-327 for the axis (0, 1, 0)
    const float ty[3] = {pose[3 * i + 0], 0.0f, pose[3 * i + 1]};  // maps.py:889-891
    const float tl[3] = {0.0f, c.cam_height, 0.0f};                // maps.py:795-797
    float* sp = samples + (size_t)i * kSampleWords;
    memset(sp, 0, sizeof(DmProjSample));
    put_step(sp, DM_STEP_ROT_THEN_ADD, c.pitch_R, tl, h->pose.fused);
    put_step(sp + kStepWords, c.plot_to_global ? DM_STEP_ROT_THEN_ADD : DM_STEP_NONE, Ry, ty, h->pose.fused);
    sp[32] = c.width_offset;
    sp[33] = c.height_offset;
    if (fast == 2 && !yaw_step_is_fast(Ry)) fast = 0;
    // the local map as a source of the merge (maps.py:2059-2060): local → global with its own pose, unless it was
    // plotted in the global frame; the target is global (maps.py:2116-2117: no second step)
    put_step(lsteps + (size_t)i * 2 * kStepWords, c.plot_to_global ? DM_STEP_NONE : DM_STEP_ROT_THEN_ADD, Ry, ty,
             h->fused_fuse);
    put_step(lsteps + (size_t)i * 2 * kStepWords + kStepWords, DM_STEP_NONE, nullptr, nullptr, 0);
    lwoff[i] = c.width_offset;
    lhoff[i] = c.height_offset;
    put_step(wsteps + (size_t)i * 2 * kStepWords, DM_STEP_NONE, nullptr, nullptr, 0);
    put_step(wsteps + (size_t)i * 2 * kStepWords + kStepWords, DM_STEP_NONE, nullptr, nullptr, 0);
    wwoff[i] = world ? world->width_offset : 0.0f;
    whoff[i] = world ? world->height_offset : 0.0f;
  }
  DM_CUDA_OK(cudaMemcpyAsync(h->d_params[slot], h->h_params[slot], h->param_bytes, cudaMemcpyHostToDevice, stream));
  DM_CUDA_OK(cudaEventRecord(h->ev[slot], stream));
  *d_base = h->d_params[slot];
  *fast_steps = fast;
  return DM_OK;
}

static void make_sources(DmBuilder* h, char* d_base, float* local_topdown, uint8_t* local_mask, const DmMapRef* world) {
  const DmBuilderCfg& c = h->cfg;
  const int b = c.b;
  float* w = reinterpret_cast<float*>(d_base);
  float* lsteps = w + (size_t)b * kSampleWords;
  float* lwoff = lsteps + (size_t)b * 2 * kStepWords;
  float* lhoff = lwoff + b;
  float* wsteps = lhoff + b;
  float* wwoff = wsteps + (size_t)b * 2 * kStepWords;
  float* whoff = wwoff + b;
  h->n_src = 0;
  if (world) {  // fuse_topdown_maps(world, new): the world map first (maps.py:2471-2508)
    DmFuseSource& s = h->src[h->n_src++];
    s.height = world->topdown; s.values = nullptr; s.mask = world->mask;
    s.height_bstride = (int64_t)world->h * world->w; s.height_cstride = s.height_bstride;
    s.h = world->h; s.w = world->w; s.flip_h = c.proj.flip_h; s.map_res = c.proj.map_res;
    s.width_offset = wwoff; s.height_offset = whoff; s.steps = reinterpret_cast<const DmStep*>(wsteps);
    s.plane_box = world->plane_box;
  }
  DmFuseSource& s = h->src[h->n_src++];
  s.height = local_topdown; s.values = nullptr; s.mask = local_mask;
  s.height_bstride = (int64_t)c.proj.Mh * c.proj.Mw; s.height_cstride = s.height_bstride;
  s.h = c.proj.Mh; s.w = c.proj.Mw; s.flip_h = c.proj.flip_h; s.map_res = c.proj.map_res;
  s.width_offset = lwoff; s.height_offset = lhoff; s.steps = reinterpret_cast<const DmStep*>(lsteps);
  s.plane_box = nullptr;
}

extern "C" int dm_builder_plot(DmBuilder* h, const float* depth, const float* pose, const float* sin_yaw,
                               const float* cos_yaw, float* local_topdown, uint8_t* local_mask, const DmMapRef* world,
                               DmMergeShape* shape, void* stream_) {
  DM_TRACE();
  return dm_builder_plot_prefill(h, depth, pose, sin_yaw, cos_yaw, local_topdown, local_mask, world, shape, nullptr,
                                 nullptr, 0, stream_);
}

// dm_builder_plot + speculation: the GPU is idle from the moment the bounding box is on its way to the host until the
// merge kernels are queued (copy latency, the host computing the canvas shape, five launches) — ~55 us of a 316 us step.
// The canvas of the merge is as a rule in the size class of the old world map, so the caller hands canvases of that
// class in BEFORE the box is known: the fill (the largest merge kernel, 115 us) is queued right behind the box's copy
// and runs while the host waits for the box and prepares the scatter launches.  dm_builder_merge skips its fill when
// `out` is what was prefilled (and large enough); anything else is filled as before.
extern "C" int dm_builder_plot_prefill(DmBuilder* h, const float* depth, const float* pose, const float* sin_yaw,
                                       const float* cos_yaw, float* local_topdown, uint8_t* local_mask,
                                       const DmMapRef* world, DmMergeShape* shape, float* prefill_topdown,
                                       uint8_t* prefill_mask, int64_t prefill_cells, void* stream_) {
  DM_TRACE();
  if (!h || !depth || !pose || !sin_yaw || !cos_yaw || !local_topdown || !local_mask) return DM_EINVAL;
  if (world && (!world->topdown || !world->mask || world->h <= 0 || world->w <= 0)) return DM_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  char* d_base = nullptr;
  int fast = 0;
  int rc = upload_params(h, pose, sin_yaw, cos_yaw, world, stream, &d_base, &fast);
  if (rc != DM_OK) return rc;
  DmProjCfg pc = h->cfg.proj;
  pc.fast_steps = fast;
  pc.want_height = 0;  // C == 0: the topdown map is the height map (maps.py:333-334)
  make_sources(h, d_base, local_topdown, local_mask, world);
  // pass 1 (maps.py:2146-2179): a world map that carries the box its own scatter pass tracked only seeds the reduction
  const bool seeded = world && world->box;
  // the projection's resolve pass and pass 1 over the local map it writes are ONE kernel (dm_fuse.cu:
  // hmap_resolve_bbox_kernel) whenever the depth-only projection can hand over its key planes
  uint32_t* planes = nullptr;
  unsigned long long slot_words = 0;
  rc = hmap_project_keys(depth, nullptr, reinterpret_cast<const DmProjSample*>(d_base), &pc, h->cfg.b, h->ws, h->ws_bytes,
                         stream, &planes, &slot_words);
  if (rc == DM_OK) {
    rc = hmap_resolve_with_bbox(planes, slot_words, &pc, h->cfg.b, local_topdown, local_mask, &h->src[h->n_src - 1],
                                h->cfg.proj.map_res, seeded ? world->box : nullptr, h->d_bbox, stream);
    if (rc != DM_OK) return rc;
    if (world && !seeded) {  // an untracked world map is scanned as before
      rc = fuse_bbox_accumulate(h->src, 1, h->cfg.b, 1, h->cfg.proj.map_res, h->d_bbox, stream);
      if (rc != DM_OK) return rc;
    }
  } else if (rc == DM_EINVAL) {  // more than 64 frames, or the height-map path is switched off: the unfused sequence
    rc = dm_orth_project_f32(depth, nullptr, nullptr, reinterpret_cast<const DmProjSample*>(d_base), &pc, h->cfg.b,
                             local_topdown, local_mask, nullptr, h->ws, h->ws_bytes, stream);
    if (rc != DM_OK) return rc;
    const DmFuseSource* scan = seeded ? &h->src[h->n_src - 1] : h->src;
    rc = dm_fuse_bbox_seeded_i64(scan, seeded ? 1 : h->n_src, h->cfg.b, 1, h->cfg.proj.map_res,
                                 seeded ? world->box : nullptr, h->d_bbox, stream);
    if (rc != DM_OK) return rc;
  } else {
    return rc;
  }
  DM_CUDA_OK(cudaMemcpyAsync(h->h_bbox, h->d_bbox, 5 * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  DM_CUDA_OK(cudaEventRecord(h->ev_bbox, stream));
  h->prefilled_top = nullptr; h->prefilled_mask = nullptr; h->prefilled_cells = 0;
  prefill_cells &= ~(int64_t)15;
  if (prefill_topdown && prefill_mask && prefill_cells > 0) {
    rc = dm_fuse_canvas_init_f32(prefill_topdown, prefill_mask, nullptr, prefill_cells, h->cfg.merge_fill_value, stream);
    if (rc != DM_OK) return rc;
    h->prefilled_top = prefill_topdown; h->prefilled_mask = prefill_mask; h->prefilled_cells = prefill_cells;
  }
  return shape ? dm_builder_plot_wait(h, shape) : DM_OK;  // shape == NULL: the caller waits later (dm_builder_plot_wait)
}

extern "C" int dm_builder_plot_wait(DmBuilder* h, DmMergeShape* shape) {
  DM_TRACE();
  if (!h || !shape) return DM_EINVAL;
  DM_CUDA_OK(cudaEventSynchronize(h->ev_bbox));  // the reference's .item() sync (maps.py:2172-2173)
  const int64_t min_x = h->h_bbox[0], max_x = h->h_bbox[1], min_z = h->h_bbox[2], max_z = h->h_bbox[3];
  shape->n_valid = h->h_bbox[4];
  if (shape->n_valid == 0) return DM_OK;  // maps.py:2217-2225: the caller keeps the last map
  // maps.py:2171-2178: sizes as Python ints, offsets as float32 tensors
  const int64_t map_width = (max_x - min_x) + 2, map_height = (max_z - min_z) + 2;
  if (map_width <= 0 || map_height <= 0 || map_width >= (1ll << 31) || map_height >= (1ll << 31)) return DM_EINVAL;
  shape->map_width = (int32_t)map_width;
  shape->map_height = (int32_t)map_height;
  shape->width_offset = (float)((double)map_width / 2.0) - (float)(max_x + min_x) / 2.0f;
  shape->height_offset = (float)((double)map_height / 2.0) - (float)(max_z + min_z) / 2.0f;
  return DM_OK;
}

extern "C" int dm_builder_merge(DmBuilder* h, const DmMapRef* out, void* stream_) {
  DM_TRACE();
  if (!h || !out || !out->topdown || !out->mask || out->h <= 0 || out->w <= 0 || h->n_src <= 0) return DM_EINVAL;
  const DmBuilderCfg& c = h->cfg;
  DmFuseTarget tgt;
  tgt.Mh = out->h; tgt.Mw = out->w; tgt.flip_h = c.proj.flip_h; tgt.map_res = c.proj.map_res;
  tgt.width_offset = out->width_offset; tgt.height_offset = out->height_offset;
  tgt.fill_value = c.merge_fill_value; tgt.reduction = c.merge_reduction;
  // the prefill covers the first prefilled_cells cells (a multiple of 16: the tail below starts 16-byte aligned) of the
  // canvases the caller guessed; a map that outgrew the guess gets its tail filled here
  const int64_t n_cells = (int64_t)c.b * out->h * out->w;
  const bool prefilled = out->topdown == h->prefilled_top && out->mask == h->prefilled_mask && h->prefilled_cells > 0;
  if (prefilled && n_cells > h->prefilled_cells) {
    const int rf = dm_fuse_canvas_init_f32(out->topdown + h->prefilled_cells, out->mask + h->prefilled_cells, nullptr,
                                           n_cells - h->prefilled_cells, c.merge_fill_value, stream_);
    if (rf != DM_OK) return rf;
  }
  const int rc = fuse_scatter_track(h->src, h->n_src, c.b, 1, &tgt, out->topdown, out->mask, nullptr, out->box,
                                    out->plane_box, prefilled ? 1 : 0, stream_);
  h->prefilled_top = nullptr; h->prefilled_mask = nullptr; h->prefilled_cells = 0;
  h->n_src = 0;
  return rc;
}

extern "C" int dm_builder_step_fixed(DmBuilder* h, const float* depth, const float* pose, const float* sin_yaw,
                                     const float* cos_yaw, float* local_topdown, uint8_t* local_mask,
                                     const DmMapRef* canvas, void* stream_) {
  DM_TRACE();
  if (!h || !depth || !pose || !sin_yaw || !cos_yaw || !local_topdown || !local_mask) return DM_EINVAL;
  if (!canvas || !canvas->topdown || !canvas->mask || canvas->h <= 0 || canvas->w <= 0) return DM_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  char* d_base = nullptr;
  int fast = 0;
  int rc = upload_params(h, pose, sin_yaw, cos_yaw, nullptr, stream, &d_base, &fast);
  if (rc != DM_OK) return rc;
  DmProjCfg pc = h->cfg.proj;
  pc.fast_steps = fast;
  pc.want_height = 0;
  rc = dm_orth_project_f32(depth, nullptr, nullptr, reinterpret_cast<const DmProjSample*>(d_base), &pc, h->cfg.b,
                           local_topdown, local_mask, nullptr, h->ws, h->ws_bytes, stream);
  if (rc != DM_OK) return rc;
  make_sources(h, d_base, local_topdown, local_mask, nullptr);
  const DmBuilderCfg& c = h->cfg;
  DmFuseTarget tgt;
  tgt.Mh = canvas->h; tgt.Mw = canvas->w; tgt.flip_h = c.proj.flip_h; tgt.map_res = c.proj.map_res;
  tgt.width_offset = canvas->width_offset; tgt.height_offset = canvas->height_offset;
  tgt.fill_value = c.merge_fill_value; tgt.reduction = c.merge_reduction;
  rc = dm_fuse_inplace_f32(h->src, 1, c.b, 1, &tgt, canvas->topdown, canvas->mask, nullptr, stream);
  h->n_src = 0;
  return rc;
}
