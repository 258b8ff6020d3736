// sgi_api.cu — the C ABI of include/shadowgi.h: context lifecycle, uploads, pass orchestration, readback.
// The passes replace the reference's per-frame free functions (ShadowMapping/src/main.cpp:350-414,
// SoftShadowMapping/src/main.cpp:756-811,925-1022, ShadowVolumes/src/main.cpp:120-206); see the header for
// the entry-point -> reference mapping.  There is no CPU fallback: without a CUDA device sgi_create fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <cstdint>
#include <utility>
#include <vector>
#include "sgi_internal.cuh"
#include "sgi_moments.cuh"

static void sync_all_streams(sgi_ctx* ctx) {
  cudaStreamSynchronize(ctx->stream);
  cudaStream_t others[] = {ctx->aux_stream, ctx->vis_stream, ctx->copy_stream, ctx->upload_stream, ctx->lane_stream[0], ctx->lane_stream[1], ctx->lane_stream[2], ctx->comm_stream, ctx->comm_stream2};
  for (cudaStream_t s : others) if (s) cudaStreamSynchronize(s);
}

static int ensure_buf(sgi_ctx* ctx, int which, size_t bytes) {
  if (ctx->buf[which] && ctx->buf_bytes[which] == bytes) return SGI_OK;
  if (ctx->buf[which] || ctx->alt[which]) {
    sync_all_streams(ctx);
    if (ctx->buf[which]) cudaFree(ctx->buf[which]);
    if (ctx->alt[which]) cudaFree(ctx->alt[which]);
    ctx->buf[which] = nullptr; ctx->alt[which] = nullptr;
    ctx->sm_reader_cur = ctx->sm_reader_alt = ctx->gb_reader_cur = ctx->gb_reader_alt = -1;   // everything has completed
    ctx->vis_in_flight = false; ctx->gbuf_in_flight = false;
  }
  ctx->buf_bytes[which] = 0;
  if (bytes == 0) return SGI_OK;
  cudaError_t e = cudaMalloc(&ctx->buf[which], bytes);
  if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return SGI_ERR_NOMEM; }
  ctx->buf_bytes[which] = bytes;
  return SGI_OK;
}

// ---- timing ring ---------------------------------------------------------------------------------------------
int sgi_timing_drain(sgi_ctx* ctx) {
  for (int p = 0; p < SGI_PASS_COUNT_; p++) {
    for (int k = 0; k < ctx->ev_n[p]; k++) {
      float ms = 0.f;
      cudaEventSynchronize(ctx->ev[p][k][1]);
      if (cudaEventElapsedTime(&ms, ctx->ev[p][k][0], ctx->ev[p][k][1]) == cudaSuccess) { ctx->pass_ms[p] += ms; ctx->pass_calls[p]++; }
    }
    ctx->ev_n[p] = 0;
  }
  return SGI_OK;
}
int sgi_timing_begin(sgi_ctx* ctx, int pass, cudaStream_t stream) {
  if (!ctx->timing) return -1;
  if (ctx->ev_n[pass] >= SGI_EV_RING) sgi_timing_drain(ctx);
  int slot = ctx->ev_n[pass];
  cudaEventRecord(ctx->ev[pass][slot][0], stream);
  return slot;
}
void sgi_timing_end(sgi_ctx* ctx, int pass, int slot, cudaStream_t stream) {
  if (slot < 0) return;
  cudaEventRecord(ctx->ev[pass][slot][1], stream);
  ctx->ev_n[pass] = slot + 1;
}

// The overflow words are sticky and are looked at on every check, whatever pass raised them (they are host memory: a few loads).
// On the pipelined path the word cannot be charged to one frame - frame k+1 is already queued when frame k's copy is waited for -
// so a raised word condemns every read ticket issued so far (overflow_upto): each of them reports SGI_ERR_OVERFLOW at its
// sgi_read_wait and none of them delivers a truncated frame as good.  `ticket_seq` < 0: a blocking check.
static int check_overflow(sgi_ctx* ctx, long long ticket_seq = -1) {
  for (SgiScratch& sc : ctx->scratch) {
    if (!sc.h_flags) continue;
    sc.overflow_pending = false;
    if (sc.h_flags[0]) {
      sc.h_flags[0] = 0;
      ctx->gbuffer_valid = false; ctx->shadow_map_valid = false;
      ctx->overflow_upto = ctx->read_seq;          // sgi_raster_run grows d_pairs from h_flags[1] on the next call
      ctx->overflow_unreported = true;
    }
  }
  const bool bad = ticket_seq < 0 ? ctx->overflow_unreported : ticket_seq < ctx->overflow_upto;
  if (bad || ticket_seq < 0) ctx->overflow_unreported = false;
  if (bad) { ctx->err = "tile list overflow: the lists were re-sized, run the frame again"; return SGI_ERR_OVERFLOW; }
  return SGI_OK;
}

// called after work that reads the G-buffer / rewrites the mesh has been queued on the main stream: the next G-buffer
// pass (auxiliary stream) must not start before this point, but need not wait for anything queued later
static void mark_gbuffer_use(sgi_ctx* ctx) { if (ctx->ev_fork) cudaEventRecord(ctx->ev_fork, ctx->stream); }

void sgi_wait_reads_of(sgi_ctx* ctx, int which, cudaStream_t writer) {
  int t = ctx->buf_read_ticket[which];
  if (t >= 0 && ctx->read_pending[t]) cudaStreamWaitEvent(writer, ctx->read_done[t], 0);
  ctx->buf_read_ticket[which] = -1;
}

// The G-buffer pass runs on the auxiliary stream so that it overlaps the light-view depth pass; every consumer of
// its outputs (and every host-visible point) first makes the main stream wait for it.
int sgi_join_gbuffer(sgi_ctx* ctx) {
  if (ctx->gbuf_in_flight) {
    SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_gbuf_done, 0));
    ctx->gbuf_in_flight = false;
  }
  return SGI_OK;
}

// The shadow pass runs on the visibility stream; consumers of its result on the main stream wait for it here.
int sgi_join_vis(sgi_ctx* ctx) {
  if (ctx->vis_in_flight && ctx->vis_last >= 0) {
    SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_vis[ctx->vis_last], 0));
    ctx->vis_in_flight = false;
  }
  return SGI_OK;
}

// Before a pass overwrites a render target that a shadow pass still reads (the previous frame's, on the visibility
// stream): switch to the other instance of the target(s) if the whole target is produced and nobody holds its device
// pointer, otherwise make the writer wait.  group 0 = shadow maps, 1 = G-buffer (position, normal, depth, albedo).
static int prepare_target_write(sgi_ctx* ctx, int group, cudaStream_t writer, bool whole) {
  int& cur = group ? ctx->gb_reader_cur : ctx->sm_reader_cur;
  int& alt = group ? ctx->gb_reader_alt : ctx->sm_reader_alt;
  if (cur >= 0 && cudaEventQuery(ctx->ev_vis[cur]) == cudaErrorNotReady) {
    static const int sm_bufs[] = {SGI_BUF_SHADOW_MAP}, gb_bufs[] = {SGI_BUF_GBUF_POS, SGI_BUF_GBUF_NRM, SGI_BUF_CAM_DEPTH, SGI_BUF_GBUF_ALBEDO};
    const int* bufs = group ? gb_bufs : sm_bufs;
    const int nb = group ? 4 : 1;
    bool can_flip = whole && ctx->overlap_passes && !(group ? ctx->gbuf_exposed : ctx->sm_exposed);
    if (can_flip)
      for (int k = 0; k < nb && can_flip; k++)
        if (!ctx->alt[bufs[k]] && ctx->buf[bufs[k]]) {
          if (cudaMalloc(&ctx->alt[bufs[k]], ctx->buf_bytes[bufs[k]]) != cudaSuccess) { cudaGetLastError(); ctx->alt[bufs[k]] = nullptr; can_flip = false; }
        }
    if (can_flip) {
      for (int k = 0; k < nb; k++) {
        std::swap(ctx->buf[bufs[k]], ctx->alt[bufs[k]]);
        // read tickets follow their instance: an asynchronous copy-out of the instance we are switching to must have
        // finished before it is overwritten; the one we leave keeps its pending ticket
        std::swap(ctx->buf_read_ticket[bufs[k]], ctx->alt_read_ticket[bufs[k]]);
        sgi_wait_reads_of(ctx, bufs[k], writer);
      }
      std::swap(cur, alt);
      if (cur >= 0) SGI_CUDA(ctx, cudaStreamWaitEvent(writer, ctx->ev_vis[cur], 0));   // its own last reader: two shadow passes ago
    } else {
      SGI_CUDA(ctx, cudaStreamWaitEvent(writer, ctx->ev_vis[cur], 0));
    }
  }
  cudaGetLastError();
  cur = -1;
  return SGI_OK;
}

extern "C" {

const char* sgi_version(void) { return "shadowgi-b200 0.1 (sm_100a)"; }

void sgi_default_params(sgi_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->technique = SGI_TECH_HARD;
  p->shadow_map_width = p->shadow_map_height = 2048;   // ShadowMapping/src/main.cpp:90-91
  p->shadow_intensity = 0.25f;                         // :877
  p->kernel_order = 7;                                 // Filter order, :859
  p->penumbra_size = 1;                                // :871
  p->blocker_search_size = 7;                          // SoftShadowMapping/src/main.cpp:1611
  p->kernel_size = 15;                                 // :1612
  p->light_source_radius = 8;                          // :1598,1613 (16/2)
  p->max_search = 16;                                  // ShadowMapping/src/main.cpp:869
  p->depth_threshold = 0.0f;
  p->z_near = 1; p->z_far = 1000;
  p->polygon_offset_factor = 4.0f; p->polygon_offset_units = 20.0f;   // :246
  p->sv_depth_func = SGI_DEPTH_LEQUAL;
  p->sv_infinity = 100;                                // ShadowVolumes/src/main.cpp:469
}

int sgi_create(sgi_ctx** out, int device) {
  if (!out) return SGI_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return SGI_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return SGI_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return SGI_ERR_CUDA;
  sgi_ctx* ctx = new (std::nothrow) sgi_ctx();
  if (!ctx) return SGI_ERR_NOMEM;
  ctx->device = device;
  { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ctx->n_sm = v; }
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SGI_ERR_CUDA; }
  ctx->stream = ctx->own_stream;
  if (cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_gbuf_done, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SGI_ERR_CUDA; }
  { const char* e = getenv("SGI_NO_OVERLAP"); ctx->overlap_passes = !(e && e[0] == '1'); }
  { const char* e = getenv("SGI_VIS_STAGED"); ctx->vis_staged = (e && e[0] == '1') ? 1 : 0; }
  { const char* e = getenv("SGI_TILE_SPLIT"); if (e) ctx->tile_split = atoi(e) < 0 ? 0 : atoi(e); }
  { const char* e = getenv("SGI_TILE_ORDER"); ctx->tile_order = (e && e[0] == '0') ? 0 : 1; }
  { const char* e = getenv("SGI_TILE_THREADS"); int v = e ? atoi(e) : 0; ctx->tile_threads = (v == 128 || v == 256 || v == 512 || v == 1024) ? v : 0; }
  { const char* e = getenv("SGI_TILE_BULK"); if (e) ctx->tile_bulk_flush = e[0] != '0'; }
  { const char* e = getenv("SGI_PDL"); if (e) ctx->pdl = e[0] != '0'; }
  { const char* e = getenv("SGI_TILE_BIN_BIG"); if (e) ctx->tile_bin_big = atoi(e); }
  { const char* e = getenv("SGI_TILE_BIN_BIG_WORK"); if (e) ctx->tile_bin_big_work = atoi(e); }
  { const char* e = getenv("SGI_TILE_FEW_WALK"); if (e) ctx->tile_few_walk = e[0] != '0'; }
  { const char* e = getenv("SGI_SV_SPLIT_LISTS"); if (e) ctx->sv_split_lists = e[0] != '0'; }
  { const char* e = getenv("SGI_TILE_DIRECT"); if (e) ctx->tile_direct = atoi(e); }
  { const char* e = getenv("SGI_TILE_STATIC"); if (e) ctx->tile_static_items = e[0] - '0'; }
  { const char* e = getenv("SGI_TILE_REFRESH_FULL"); if (e) ctx->tile_refresh_full = e[0] - '0'; }
  // lowest priority: when CTAs of the next frame's raster passes and of this frame's shadow pass compete for an SM, the raster
  // ones go first (they are latency-bound chains on the critical path); the shadow pass fills what they leave
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  { const char* e = getenv("SGI_VIS_PRIORITY"); if (e && e[0] == '0') prio_lo = 0; }
  if (cudaStreamCreateWithPriority(&ctx->vis_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_s2v, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SGI_ERR_CUDA; }
  for (int k = 0; k < 4; k++) if (cudaEventCreateWithFlags(&ctx->ev_vis[k], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SGI_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_upload_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_geom_main, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SGI_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SGI_ERR_CUDA; }
  for (int k = 0; k < 4; k++) if (cudaEventCreateWithFlags(&ctx->read_done[k], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SGI_ERR_CUDA; }
  for (int b = 0; b < SGI_BUF_COUNT_; b++) { ctx->buf_read_ticket[b] = -1; ctx->alt_read_ticket[b] = -1; }
  sgi_default_params(&ctx->params);
  for (int p = 0; p < SGI_PASS_COUNT_; p++) {
    ctx->ev_n[p] = 0; ctx->pass_ms[p] = 0; ctx->pass_calls[p] = 0;
    for (int k = 0; k < SGI_EV_RING; k++) { ctx->ev[p][k][0] = nullptr; ctx->ev[p][k][1] = nullptr; }
  }
  ctx->light_pos[0] = ctx->light_pos[1] = ctx->light_pos[2] = 0.f;
  *out = ctx;
  return SGI_OK;
}

int sgi_destroy(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  sync_all_streams(ctx);
  for (int b = 0; b < SGI_BUF_COUNT_; b++) { if (ctx->buf[b]) cudaFree(ctx->buf[b]); if (ctx->alt[b]) cudaFree(ctx->alt[b]); }
  sgi_comm_destroy(ctx);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  if (ctx->comm_stream2) cudaStreamDestroy(ctx->comm_stream2);
  if (ctx->d_light_gid) cudaFree(ctx->d_light_gid);
  if (ctx->ev_comm_in) cudaEventDestroy(ctx->ev_comm_in);
  for (int b = 0; b < SGI_BUF_COUNT_; b++) if (ctx->ev_comm_done[b]) cudaEventDestroy(ctx->ev_comm_done[b]);
  for (int b = 0; b < SGI_EDT_NBUF; b++) if (ctx->edt_buf[b]) cudaFree(ctx->edt_buf[b]);
  if (ctx->vis_spare) cudaFree(ctx->vis_spare);
  if (ctx->d_mm) cudaFree(ctx->d_mm);
  for (void* p : {(void*)ctx->d_sv_grp_start, (void*)ctx->d_sv_grp_ent, (void*)ctx->d_sv_cls, (void*)ctx->d_sv_frags}) if (p) cudaFree(p);
  if (ctx->rbssm_buf) cudaFree(ctx->rbssm_buf);
  if (ctx->d_rgb) cudaFree(ctx->d_rgb);
  if (ctx->d_uv) cudaFree(ctx->d_uv);
  for (int k = 0; k < 3; k++) if (ctx->d_tex[k]) cudaFree(ctx->d_tex[k]);
  void* ptrs[] = {ctx->d_xyz_set[0], ctx->d_nrm_set[0], ctx->d_idx_set[0], ctx->d_xyz_set[1], ctx->d_nrm_set[1], ctx->d_idx_set[1], ctx->d_light_trans};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (SgiScratch& sc : ctx->scratch) sgi_raster_free(sc);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_gbuf_done) cudaEventDestroy(ctx->ev_gbuf_done);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  for (int k = 0; k < SGI_LIGHT_LANES; k++) {
    if (ctx->lane_stream[k]) { cudaStreamSynchronize(ctx->lane_stream[k]); cudaStreamDestroy(ctx->lane_stream[k]); }
    if (ctx->ev_lane_done[k]) cudaEventDestroy(ctx->ev_lane_done[k]);
  }
  if (ctx->ev_lane_fork) cudaEventDestroy(ctx->ev_lane_fork);
  for (int k = 0; k < 4; k++) { if (ctx->h_stage[k]) cudaFreeHost(ctx->h_stage[k]); if (ctx->ev_stage[k]) cudaEventDestroy(ctx->ev_stage[k]); }
  if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
  for (int k = 0; k < 4; k++) if (ctx->read_done[k]) cudaEventDestroy(ctx->read_done[k]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
  if (ctx->vis_stream) cudaStreamDestroy(ctx->vis_stream);
  if (ctx->ev_s2v) cudaEventDestroy(ctx->ev_s2v);
  for (int k = 0; k < 4; k++) if (ctx->ev_vis[k]) cudaEventDestroy(ctx->ev_vis[k]);
  if (ctx->ev_upload_done) cudaEventDestroy(ctx->ev_upload_done);
  if (ctx->ev_geom_main) cudaEventDestroy(ctx->ev_geom_main);
  free(ctx->h_light_mvp); free(ctx->h_light_mvp_b);
  for (int p = 0; p < SGI_PASS_COUNT_; p++)
    for (int k = 0; k < SGI_EV_RING; k++) { if (ctx->ev[p][k][0]) cudaEventDestroy(ctx->ev[p][k][0]); if (ctx->ev[p][k][1]) cudaEventDestroy(ctx->ev[p][k][1]); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return SGI_OK;
}

int sgi_set_stream(sgi_ctx* ctx, void* cuda_stream) {
  if (!ctx) return SGI_ERR_INVALID;
  sgi_join_gbuffer(ctx); sgi_join_vis(ctx);
  sync_all_streams(ctx);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  mark_gbuffer_use(ctx);
  return SGI_OK;
}

// Upload stream protocol: begin = wait for every pass queued so far that reads geometry / colours (the depth and
// shadow-volume passes on the main stream, the G-buffer pass on the auxiliary stream); end = the main stream (and through
// ev_fork the auxiliary one) waits for the copies.
static int upload_begin(sgi_ctx* ctx) {
  if (ctx->geom_main_recorded) SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->upload_stream, ctx->ev_geom_main, 0));
  if (ctx->gbuf_done_recorded) SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->upload_stream, ctx->ev_gbuf_done, 0));
  return SGI_OK;
}
static int upload_end(sgi_ctx* ctx) {
  SGI_CUDA(ctx, cudaEventRecord(ctx->ev_upload_done, ctx->upload_stream));
  SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_upload_done, 0));
  return SGI_OK;
}

static bool host_is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// page-locked staging slot of at least `bytes`, free to overwrite (its previous DMA has completed)
static int stage_acquire(sgi_ctx* ctx, int slot, size_t bytes, char** out) {
  if (!ctx->ev_stage[slot]) SGI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_stage[slot], cudaEventDisableTiming));
  else SGI_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage[slot]));
  if (ctx->h_stage_bytes[slot] < bytes) {
    if (ctx->h_stage[slot]) cudaFreeHost(ctx->h_stage[slot]);
    ctx->h_stage[slot] = nullptr; ctx->h_stage_bytes[slot] = 0;
    const size_t cap = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc(&ctx->h_stage[slot], cap, cudaHostAllocDefault) != cudaSuccess) { ctx->err = "cudaHostAlloc (upload staging)"; return SGI_ERR_NOMEM; }
    ctx->h_stage_bytes[slot] = cap;
  }
  *out = (char*)ctx->h_stage[slot];
  return SGI_OK;
}

int sgi_set_mesh(sgi_ctx* ctx, const float* xyz, const float* nrm, int32_t V, const int32_t* idx, int32_t T) {
  if (!ctx || V < 0 || T < 0 || (V > 0 && (!xyz || !nrm)) || (T > 0 && !idx)) { if (ctx) ctx->err = "sgi_set_mesh: bad arguments"; return SGI_ERR_INVALID; }
  {
    int32_t lo = 0, hi = -1;                       // min / max in one vectorisable sweep
    const int64_t n = (int64_t)T * 3;
    if (n > 0) { lo = idx[0]; hi = idx[0]; }
    for (int64_t k = 0; k < n; k++) { const int32_t v = idx[k]; lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
    if (n > 0 && (lo < 0 || hi >= V)) { ctx->err = "sgi_set_mesh: index out of range"; return SGI_ERR_INVALID; }
    if (ctx->sv_track) {        // shadow volumes, silhouette form: the edge groups survive a re-upload of the same indices
      if ((int64_t)ctx->h_idx_copy.size() != n || (n > 0 && memcmp(ctx->h_idx_copy.data(), idx, (size_t)n * 4) != 0)) {
        ctx->h_idx_copy.assign(idx, idx + n);
        ctx->sv_edges_valid = false;
      }
    }
  }
  cudaSetDevice(ctx->device);
  // Upload into the geometry set the frame in flight is NOT using.  Ordering on the main stream is enough: the passes
  // that read this set two uploads ago were queued on (or joined into) the main stream before this copy.
  sgi_join_gbuffer(ctx);
  const int s = ctx->mesh_cur ^ 1;
  if (V != ctx->mesh_V[s] || T != ctx->mesh_T[s]) {
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->upload_stream));
    if (ctx->d_xyz_set[s]) cudaFree(ctx->d_xyz_set[s]);
    if (ctx->d_nrm_set[s]) cudaFree(ctx->d_nrm_set[s]);
    if (ctx->d_idx_set[s]) cudaFree(ctx->d_idx_set[s]);
    ctx->d_xyz_set[s] = ctx->d_nrm_set[s] = nullptr; ctx->d_idx_set[s] = nullptr;
    SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_xyz_set[s], (size_t)(V > 0 ? V : 1) * 12));
    SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_nrm_set[s], (size_t)(V > 0 ? V : 1) * 12));
    SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_idx_set[s], (size_t)(T > 0 ? T : 1) * 12));
    ctx->mesh_V[s] = V; ctx->mesh_T[s] = T;
  }
  {
    const size_t vb = (size_t)V * 12, tb = (size_t)T * 12;
    const void *sx = xyz, *sn = nrm, *si = idx;
    const bool borrow = ctx->borrow_pinned && (V == 0 || (host_is_pinned(xyz) && host_is_pinned(nrm))) && (T == 0 || host_is_pinned(idx));
    int slot = -1;
    if (!borrow) {
      char* stg = nullptr;
      slot = ctx->stage_next_mesh; ctx->stage_next_mesh ^= 1;
      int rc = stage_acquire(ctx, slot, 2 * vb + tb, &stg);
      if (rc) return rc;
      if (V > 0) { memcpy(stg, xyz, vb); memcpy(stg + vb, nrm, vb); }
      if (T > 0) memcpy(stg + 2 * vb, idx, tb);
      sx = stg; sn = stg + vb; si = stg + 2 * vb;
    }
    int rc = upload_begin(ctx);
    if (rc) return rc;
    if (V > 0) {
      SGI_CUDA(ctx, cudaMemcpyAsync(ctx->d_xyz_set[s], sx, vb, cudaMemcpyHostToDevice, ctx->upload_stream));
      SGI_CUDA(ctx, cudaMemcpyAsync(ctx->d_nrm_set[s], sn, vb, cudaMemcpyHostToDevice, ctx->upload_stream));
    }
    if (T > 0) SGI_CUDA(ctx, cudaMemcpyAsync(ctx->d_idx_set[s], si, tb, cudaMemcpyHostToDevice, ctx->upload_stream));
    if (slot >= 0) SGI_CUDA(ctx, cudaEventRecord(ctx->ev_stage[slot], ctx->upload_stream));
    if ((rc = upload_end(ctx))) return rc;
  }
  ctx->mesh_cur = s;
  ctx->d_xyz = ctx->d_xyz_set[s]; ctx->d_nrm = ctx->d_nrm_set[s]; ctx->d_idx = ctx->d_idx_set[s];
  ctx->V = V; ctx->T = T;
  mark_gbuffer_use(ctx);
  ctx->gbuffer_valid = ctx->shadow_map_valid = false; ctx->ids_valid = false; ctx->mm_valid = false;
  ctx->moments_tech = ctx->filtered_tech = -1;        // the moment target / filtered map describe the previous geometry
  return SGI_OK;
}

int sgi_set_mesh_colors(sgi_ctx* ctx, const float* rgb) {
  if (!ctx) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  sgi_join_gbuffer(ctx);
  if (!rgb) { ctx->has_rgb = false; ctx->gbuffer_valid = false; return SGI_OK; }
  if (ctx->V <= 0) { ctx->err = "sgi_set_mesh_colors: set the mesh first"; return SGI_ERR_INVALID; }
  if (ctx->rgb_V != ctx->V) {
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->upload_stream));
    if (ctx->d_rgb) cudaFree(ctx->d_rgb);
    ctx->d_rgb = nullptr;
    SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_rgb, (size_t)ctx->V * 12));
    ctx->rgb_V = ctx->V;
  }
  {
    // stream order is enough for the device buffer: the G-buffer pass that last read it was joined into the main stream
    // above, and the next one forks after this copy (mark_gbuffer_use)
    const void* src = rgb;
    int slot = -1;
    if (!(ctx->borrow_pinned && host_is_pinned(rgb))) {
      char* stg = nullptr;
      slot = 2 + ctx->stage_next_rgb; ctx->stage_next_rgb ^= 1;
      int rc = stage_acquire(ctx, slot, (size_t)ctx->V * 12, &stg);
      if (rc) return rc;
      memcpy(stg, rgb, (size_t)ctx->V * 12);
      src = stg;
    }
    int rc = upload_begin(ctx);
    if (rc) return rc;
    SGI_CUDA(ctx, cudaMemcpyAsync(ctx->d_rgb, src, (size_t)ctx->V * 12, cudaMemcpyHostToDevice, ctx->upload_stream));
    if (slot >= 0) SGI_CUDA(ctx, cudaEventRecord(ctx->ev_stage[slot], ctx->upload_stream));
    if ((rc = upload_end(ctx))) return rc;
  }
  mark_gbuffer_use(ctx);
  ctx->has_rgb = true; ctx->gbuffer_valid = false;
  return SGI_OK;
}

// Mesh::getTextureCoords(): (u, v, texture id) per vertex, the `uv` attribute of GBuffer.vert:15,20 (loadVBOs, VBOs[2])
int sgi_set_mesh_uv(sgi_ctx* ctx, const float* uv) {
  if (!ctx) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  sgi_join_gbuffer(ctx);
  if (!uv) { ctx->has_uv = false; ctx->gbuffer_valid = false; return SGI_OK; }
  if (ctx->V <= 0) { ctx->err = "sgi_set_mesh_uv: set the mesh first"; return SGI_ERR_INVALID; }
  if (ctx->uv_V != ctx->V) {
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->upload_stream));
    if (ctx->d_uv) cudaFree(ctx->d_uv);
    ctx->d_uv = nullptr;
    SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_uv, (size_t)ctx->V * 12));
    ctx->uv_V = ctx->V;
  }
  // (texture coordinates change with the mesh, not per frame: a plain ordered copy on the context's stream, after the passes queued so far)
  sgi_join_vis(ctx);
  SGI_CUDA(ctx, cudaMemcpyAsync(ctx->d_uv, uv, (size_t)ctx->V * 12, cudaMemcpyHostToDevice, ctx->stream));
  SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));        // the caller's array is borrowed for the call only
  mark_gbuffer_use(ctx);
  ctx->has_uv = true; ctx->gbuffer_valid = false;
  return SGI_OK;
}

// MyGLTextureViewer::loadRGBTexture (MyGLTextureViewer.cpp:45-56; Mesh::loadTexture, `m` directive): texture<index> of GBuffer.frag,
// RGB8, row 0 = t 0, GL_LINEAR / GL_REPEAT.  rgb == NULL unbinds.
int sgi_set_texture(sgi_ctx* ctx, int32_t index, const uint8_t* rgb, int32_t width, int32_t height) {
  if (!ctx || index < 0 || index > 2 || (rgb && (width <= 0 || height <= 0 || width > 16384 || height > 16384))) { if (ctx) ctx->err = "sgi_set_texture: bad arguments"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  sgi_join_gbuffer(ctx); sgi_join_vis(ctx);
  SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_tex[index]) cudaFree(ctx->d_tex[index]);
  ctx->d_tex[index] = nullptr; ctx->tex_w[index] = ctx->tex_h[index] = 0;
  ctx->gbuffer_valid = false;
  if (!rgb) return SGI_OK;
  const size_t n = (size_t)width * height;
  std::vector<uchar4> tmp(n);
  for (size_t i = 0; i < n; i++) tmp[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
  SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_tex[index], n * 4));
  SGI_CUDA(ctx, cudaMemcpy(ctx->d_tex[index], tmp.data(), n * 4, cudaMemcpyHostToDevice));
  ctx->tex_w[index] = width; ctx->tex_h[index] = height;
  return SGI_OK;
}

// useTextureForColoring: texture coordinates of the current mesh and at least one bound texture
static bool textures_active(const sgi_ctx* ctx) {
  return ctx->has_uv && ctx->uv_V == ctx->V && (ctx->d_tex[0] || ctx->d_tex[1] || ctx->d_tex[2]);
}

int sgi_set_camera(sgi_ctx* ctx, const float mvp[16], const float mv[16], const float nm[9], int32_t W, int32_t H) {
  if (!ctx || !mvp || !mv || !nm || W <= 0 || H <= 0 || W > 32767 || H > 32767) { if (ctx) ctx->err = "sgi_set_camera: bad arguments"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  sgi_join_gbuffer(ctx);
  memcpy(ctx->cam_mvp, mvp, 64); memcpy(ctx->cam_mv, mv, 64); memcpy(ctx->cam_nm, nm, 36);
  int rc;
  const bool resized = W != ctx->W || H != ctx->H;
  ctx->W = W; ctx->H = H;
  // screen targets are padded to comm_n equal strips (sgi_comm.cu), so that a multi-GPU exchange is one in-place collective
  const size_t px = sgi_padded_pixels(ctx);
  if ((rc = ensure_buf(ctx, SGI_BUF_GBUF_POS, px * 16))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_GBUF_NRM, px * 16))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_CAM_DEPTH, px * 4))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_VISIBILITY, px * 4))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_GBUF_ALBEDO, px * 16))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_SHADED, px * 16))) return rc;
  if (resized) {
    SGI_CUDA(ctx, cudaMemsetAsync(ctx->buf[SGI_BUF_VISIBILITY], 0, px * 4, ctx->stream));
    ctx->scratch[1].sized[SGI_MODE_GBUFFER] = ctx->scratch[0].sized[SGI_MODE_SVCOUNT] = false;
  }
  mark_gbuffer_use(ctx);
  ctx->has_camera = true; ctx->gbuffer_valid = false; ctx->ids_valid = false;
  return SGI_OK;
}

int sgi_set_lights(sgi_ctx* ctx, int32_t N, const float* light_mvp, const float* light_mvp_b, const float lpos[3], int32_t SW, int32_t SH) {
  if (!ctx || N <= 0 || N > SGI_MAX_LIGHTS || !light_mvp || !light_mvp_b || !lpos || SW <= 0 || SH <= 0 || SW > 32767 || SH > 32767) {
    if (ctx) ctx->err = "sgi_set_lights: bad arguments";
    return SGI_ERR_INVALID;
  }
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_SHADOW_MAP, (size_t)N * SW * SH * 4))) return rc;
  if (N != ctx->N) {
    free(ctx->h_light_mvp); free(ctx->h_light_mvp_b);
    ctx->h_light_mvp = (float*)malloc((size_t)N * 64); ctx->h_light_mvp_b = (float*)malloc((size_t)N * 64);
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_light_trans) cudaFree(ctx->d_light_trans);
    ctx->d_light_trans = nullptr;
    SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_light_trans, (size_t)N * 16));
  }
  if (SW != ctx->SW || SH != ctx->SH)
    for (SgiScratch& sc : ctx->scratch) sc.sized[SGI_MODE_DEPTH] = false;
  ctx->N = N; ctx->SW = SW; ctx->SH = SH;
  memcpy(ctx->h_light_mvp, light_mvp, (size_t)N * 64);
  memcpy(ctx->h_light_mvp_b, light_mvp_b, (size_t)N * 64);
  memcpy(ctx->light_pos, lpos, 12);
  ctx->trans_dirty = true;          // lightMVPTrans[] is uploaded lazily by the many-light pass (no sync on the single-light path)
  ctx->shadow_map_valid = false; ctx->mm_valid = false;
  ctx->moments_tech = ctx->filtered_tech = -1;        // a moment target rendered for the previous light (possibly another map size) is stale
  return SGI_OK;
}

int sgi_set_multi_light_common(sgi_ctx* ctx, const float m[16]) {
  if (!ctx) return SGI_ERR_INVALID;
  ctx->has_multi_common = m != nullptr;
  if (m) memcpy(ctx->multi_common, m, 64);
  return SGI_OK;
}

int sgi_set_params(sgi_ctx* ctx, const sgi_params* p) {
  if (!ctx || !p) return SGI_ERR_INVALID;
  if (p->technique < 0 || p->technique > SGI_TECH_PCF_TRICUBIC) { ctx->err = "sgi_set_params: unknown technique"; return SGI_ERR_INVALID; }
  if (p->kernel_order <= 0 || p->blocker_search_size <= 0 || p->kernel_size <= 0 || p->max_search < 0 || p->max_search > 4096 ||
      p->blocker_search_size > SGI_MAX_PCF_TAPS || p->kernel_size > SGI_MAX_PCF_TAPS) {
    ctx->err = "sgi_set_params: kernel sizes must be in 1..64";
    return SGI_ERR_INVALID;
  }
  if (sgi_is_moment_tech(p->technique) && (p->kernel_order < 3 || (p->kernel_order & 1) == 0 || p->kernel_order > (p->technique == SGI_TECH_ESM ? 25 : SGI_MOM_MAX_ORDER))) {
    // the reference only ever builds odd orders (7, then +-2 from the keyboard, main.cpp:545,572-573)
    ctx->err = "sgi_set_params: blur order must be odd and in 3..33 (3..25 for ESM: `uniform float kernel[25]`, LogGaussianFilter.frag:8)";
    return SGI_ERR_INVALID;
  }
  int n1 = sgi_host_pcf_offsets(p->kernel_order, p->penumbra_size, 0, ctx->pcf_off, SGI_MAX_PCF_TAPS);
  int n2 = sgi_host_pcf_offsets(p->kernel_order, p->penumbra_size, 1, ctx->rpcf_off, SGI_MAX_PCF_TAPS);
  if (n1 < 0 || n2 < 0) { ctx->err = "sgi_set_params: more than 64 PCF taps per axis"; return SGI_ERR_INVALID; }
  ctx->pcf_n = n1; ctx->rpcf_n = n2;
  ctx->params = *p;
  ctx->has_params = true;
  return SGI_OK;
}

int sgi_render_shadow_map(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  if (!ctx->d_idx || ctx->N <= 0) { ctx->err = "sgi_render_shadow_map: set mesh and lights first"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  if (!sgi_is_moment_tech(ctx->params.technique)) {
    // (the moment pass writes SGI_BUF_MOMENTS, not the depth map: it must not switch depth-map instances, which would leave
    //  shadow_map_valid pointing at a stale instance)
    int rc0 = prepare_target_write(ctx, 0, ctx->stream, true);      // the depth passes always produce whole maps
    if (rc0) return rc0;
  }
  if (sgi_is_moment_tech(ctx->params.technique)) {
    // VSM / ESM / EVSM / MSM: the light-view pass writes the moment target instead of a depth map (displaySceneFromLightPOV
    // binds Moments / Exponential / ExponentialMoments.frag, main.cpp:227-243).  One light; the shadow pass of the previous
    // frame may still read the filtered map on the visibility stream, but not this target: only the blur reads it.
    if (ctx->N != 1) { ctx->err = "sgi_render_shadow_map: moment shadow maps take one light"; return SGI_ERR_INVALID; }
    int rc = ensure_buf(ctx, SGI_BUF_MOMENTS, (size_t)ctx->SW * ctx->SH * 16);
    if (rc) return rc;
    sgi_wait_reads_of(ctx, SGI_BUF_MOMENTS, ctx->stream);
    int slot = sgi_timing_begin(ctx, SGI_PASS_SHADOW_MAP, ctx->stream);
    SgiRasterJob job;
    memset(&job, 0, sizeof(job));
    job.mode = SGI_MODE_MOMENTS;
    job.xyz = ctx->d_xyz; job.nrm = ctx->d_nrm; job.idx = ctx->d_idx; job.T = ctx->T;
    memcpy(job.mvp, ctx->h_light_mvp, 64);
    job.W = ctx->SW; job.H = ctx->SH;
    job.use_offset = 1; job.factor = ctx->params.polygon_offset_factor; job.units = ctx->params.polygon_offset_units;
    job.mom4 = (float4*)ctx->buf[SGI_BUF_MOMENTS]; job.mom_tech = ctx->params.technique;
    job.z_near = ctx->params.z_near; job.z_far = ctx->params.z_far;
    float minv[16];
    sgi_moments_quantization(job.mq, minv, job.mqt);
    rc = sgi_raster_run(ctx, job, 0, ctx->stream);
    if (rc) return rc;
    sgi_timing_end(ctx, SGI_PASS_SHADOW_MAP, slot, ctx->stream);
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_geom_main, ctx->stream)); ctx->geom_main_recorded = true;
    ctx->moments_tech = ctx->params.technique; ctx->filtered_tech = -1;
    ctx->moments_w = ctx->SW; ctx->moments_h = ctx->SH;
    return SGI_OK;
  }
  sgi_wait_reads_of(ctx, SGI_BUF_SHADOW_MAP, ctx->stream);
  int slot = sgi_timing_begin(ctx, SGI_PASS_SHADOW_MAP, ctx->stream);
  // one light: main stream, scratch set 0.  Several lights: dealt round-robin to SGI_LIGHT_LANES streams with their own
  // scratch sets (forked from / joined into the main stream), so one light's binning overlaps another's tile kernel.
  const bool lanes = ctx->N > 1 && ctx->overlap_passes;
  if (lanes) {
    for (int k = 0; k < SGI_LIGHT_LANES; k++)
      if (!ctx->lane_stream[k]) {
        SGI_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lane_stream[k], cudaStreamNonBlocking));
        SGI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_lane_done[k], cudaEventDisableTiming));
        if (!ctx->ev_lane_fork) SGI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_lane_fork, cudaEventDisableTiming));
      }
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_lane_fork, ctx->stream));
    for (int k = 0; k < SGI_LIGHT_LANES; k++) SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[k], ctx->ev_lane_fork, 0));
  }
  // PCF / PCSS with one light: the tile flush also records the extrema of every 32x32-texel block, dilated afterwards over the
  // reach of the technique's tap window; the shadow pass then decides whole windows with one comparison (sgi_shadow.cu)
  ctx->mm_valid = false;
  const int mm_reach = (ctx->N == 1 && ctx->vis_minmax_cull) ? sgi_minmax_reach(ctx) : 0;
  if (mm_reach > 0) {
    const int tiles_x = (ctx->SW + SGI_TILE - 1) >> SGI_TILE_LOG2, tiles_y = (ctx->SH + SGI_TILE - 1) >> SGI_TILE_LOG2;
    ctx->mm_w = 2 * tiles_x; ctx->mm_h = 2 * tiles_y;
    const size_t need = (size_t)ctx->mm_w * ctx->mm_h * 4 * 6;          // min | max | two dilated (min, max) sets
    if (ctx->mm_bytes < need) {
      sync_all_streams(ctx);
      if (ctx->d_mm) cudaFree(ctx->d_mm);
      ctx->d_mm = nullptr; ctx->mm_bytes = 0;
      SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_mm, need));
      ctx->mm_bytes = need;
    }
  }
  for (int l = 0; l < ctx->N; l++) {
    SgiRasterJob job;
    memset(&job, 0, sizeof(job));
    job.mode = SGI_MODE_DEPTH;
    if (mm_reach > 0) { job.mm_min = ctx->d_mm; job.mm_max = ctx->d_mm + (size_t)ctx->mm_w * ctx->mm_h; job.mm_w = ctx->mm_w; }
    job.xyz = ctx->d_xyz; job.nrm = ctx->d_nrm; job.idx = ctx->d_idx; job.T = ctx->T;
    memcpy(job.mvp, ctx->h_light_mvp + 16 * (size_t)l, 64);
    job.W = ctx->SW; job.H = ctx->SH;
    job.use_offset = 1; job.factor = ctx->params.polygon_offset_factor; job.units = ctx->params.polygon_offset_units;
    job.depth = (float*)ctx->buf[SGI_BUF_SHADOW_MAP] + (size_t)l * ctx->SW * ctx->SH;
    const int lane = l % SGI_LIGHT_LANES;
    int rc = lanes ? sgi_raster_run(ctx, job, 2 + lane, ctx->lane_stream[lane]) : sgi_raster_run(ctx, job, 0, ctx->stream);
    if (rc) return rc;
  }
  if (lanes)
    for (int k = 0; k < SGI_LIGHT_LANES; k++) {
      SGI_CUDA(ctx, cudaEventRecord(ctx->ev_lane_done[k], ctx->lane_stream[k]));
      SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_lane_done[k], 0));
    }
  if (mm_reach > 0) {
    // the dilated set belongs to the shadow-map instance just written (the previous frame's shadow pass may still read the other)
    const int set = (ctx->alt[SGI_BUF_SHADOW_MAP] && (uintptr_t)ctx->buf[SGI_BUF_SHADOW_MAP] > (uintptr_t)ctx->alt[SGI_BUF_SHADOW_MAP]) ? 1 : 0;
    const int R = (mm_reach + 1 + 31) / 32;
    int rc = sgi_minmax_dilate(ctx, set, R, ctx->stream);
    if (rc) return rc;
    ctx->mm_valid = true; ctx->mm_set = set; ctx->mm_radius = R;
  }
  sgi_timing_end(ctx, SGI_PASS_SHADOW_MAP, slot, ctx->stream);
  SGI_CUDA(ctx, cudaEventRecord(ctx->ev_geom_main, ctx->stream)); ctx->geom_main_recorded = true;
  ctx->shadow_map_valid = true;
  return SGI_OK;
}

int sgi_render_gbuffer(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  if (!ctx->d_idx || !ctx->has_camera) { ctx->err = "sgi_render_gbuffer: set mesh and camera first"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  // fork: the auxiliary stream picks up after everything already queued on the main stream (mesh upload, the
  // previous frame's consumers of the G-buffer) and then runs concurrently with what the main stream does next
  sgi_join_gbuffer(ctx);
  cudaStream_t st = ctx->overlap_passes ? ctx->aux_stream : ctx->stream;
  if (ctx->overlap_passes) {
    // ev_fork marks the last point of the main stream that touched the G-buffer or the mesh (mark_gbuffer_use);
    // if a device pointer was lent out, unknown readers may follow it, so fork from "now" instead
    if (ctx->gbuf_exposed) SGI_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    SGI_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_fork, 0));
  }
  {
    const sgi_params& q = ctx->params;
    const bool whole = (q.rect_x1 <= q.rect_x0 || q.rect_y1 <= q.rect_y0) || (q.rect_x0 <= 0 && q.rect_y0 <= 0 && q.rect_x1 >= ctx->W && q.rect_y1 >= ctx->H);
    int rc0 = prepare_target_write(ctx, 1, st, whole);
    if (rc0) return rc0;
  }
  sgi_wait_reads_of(ctx, SGI_BUF_GBUF_POS, st); sgi_wait_reads_of(ctx, SGI_BUF_GBUF_NRM, st); sgi_wait_reads_of(ctx, SGI_BUF_CAM_DEPTH, st);
  sgi_wait_reads_of(ctx, SGI_BUF_GBUF_ALBEDO, st);
  for (int b : {SGI_BUF_GBUF_POS, SGI_BUF_GBUF_NRM, SGI_BUF_CAM_DEPTH, SGI_BUF_GBUF_ALBEDO}) sgi_wait_comm(ctx, b, st);
  // a fused many-light pass still resolving positions from this scratch set's records (previous frame, visibility stream)
  if (ctx->rec_reader >= 0) { SGI_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_vis[ctx->rec_reader], 0)); ctx->rec_reader = -1; }
  ctx->ids_valid = false;
  int slot = sgi_timing_begin(ctx, SGI_PASS_GBUFFER, st);
  SgiRasterJob job;
  memset(&job, 0, sizeof(job));
  job.mode = SGI_MODE_GBUFFER;
  job.xyz = ctx->d_xyz; job.nrm = ctx->d_nrm; job.idx = ctx->d_idx; job.T = ctx->T;
  memcpy(job.mvp, ctx->cam_mvp, 64);
  job.W = ctx->W; job.H = ctx->H;
  job.depth = (float*)ctx->buf[SGI_BUF_CAM_DEPTH];
  job.pos4 = (float4*)ctx->buf[SGI_BUF_GBUF_POS]; job.nrm4 = (float4*)ctx->buf[SGI_BUF_GBUF_NRM];
  const bool rgb_ok = ctx->has_rgb && ctx->rgb_V == ctx->V, tex_ok = textures_active(ctx);
  job.rgb = rgb_ok ? ctx->d_rgb : nullptr; job.albedo4 = (rgb_ok || tex_ok) ? (float4*)ctx->buf[SGI_BUF_GBUF_ALBEDO] : nullptr;
  if (tex_ok) {
    job.uv = ctx->d_uv;
    for (int k = 0; k < 3; k++) { job.tex[k].texels = ctx->d_tex[k]; job.tex[k].w = ctx->tex_w[k]; job.tex[k].h = ctx->tex_h[k]; }
  }
  job.rx0 = ctx->params.rect_x0; job.ry0 = ctx->params.rect_y0; job.rx1 = ctx->params.rect_x1; job.ry1 = ctx->params.rect_y1;
  int rc = sgi_raster_run(ctx, job, 1, st);
  if (rc) return rc;
  sgi_timing_end(ctx, SGI_PASS_GBUFFER, slot, st);
  if (ctx->overlap_passes) {
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_gbuf_done, st));
    ctx->gbuf_in_flight = true; ctx->gbuf_done_recorded = true;
  } else {
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_geom_main, ctx->stream)); ctx->geom_main_recorded = true;
  }
  ctx->gbuffer_valid = true;
  return SGI_OK;
}

int sgi_render_prim_ids(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  if (!ctx->d_idx || !ctx->has_camera) { ctx->err = "sgi_render_prim_ids: set mesh and camera first"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  int rc = ensure_buf(ctx, SGI_BUF_PRIM_ID, sgi_padded_pixels(ctx) * 4);
  if (rc) return rc;
  // same stream protocol as sgi_render_gbuffer: auxiliary stream, forked from the last point of the main stream that touched the mesh
  sgi_join_gbuffer(ctx);
  cudaStream_t st = ctx->overlap_passes ? ctx->aux_stream : ctx->stream;
  if (ctx->overlap_passes) {
    if (ctx->gbuf_exposed) SGI_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    SGI_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_fork, 0));
  }
  // the previous frame's fused many-light pass reads this buffer and this scratch set's records on the visibility stream
  if (ctx->rec_reader >= 0) { SGI_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_vis[ctx->rec_reader], 0)); ctx->rec_reader = -1; }
  sgi_wait_reads_of(ctx, SGI_BUF_PRIM_ID, st);
  sgi_wait_comm(ctx, SGI_BUF_PRIM_ID, st);
  int slot = sgi_timing_begin(ctx, SGI_PASS_GBUFFER, st);
  SgiRasterJob job;
  memset(&job, 0, sizeof(job));
  job.mode = SGI_MODE_IDS;
  job.xyz = ctx->d_xyz; job.nrm = ctx->d_nrm; job.idx = ctx->d_idx; job.T = ctx->T;
  memcpy(job.mvp, ctx->cam_mvp, 64);
  job.W = ctx->W; job.H = ctx->H;
  job.ids = (unsigned int*)ctx->buf[SGI_BUF_PRIM_ID];
  job.rx0 = ctx->params.rect_x0; job.ry0 = ctx->params.rect_y0; job.rx1 = ctx->params.rect_x1; job.ry1 = ctx->params.rect_y1;
  rc = sgi_raster_run(ctx, job, 1, st);
  if (rc) return rc;
  sgi_timing_end(ctx, SGI_PASS_GBUFFER, slot, st);
  if (ctx->overlap_passes) {
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_gbuf_done, st));
    ctx->gbuf_in_flight = true; ctx->gbuf_done_recorded = true;
  } else {
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_geom_main, ctx->stream)); ctx->geom_main_recorded = true;
  }
  ctx->ids_valid = true; ctx->gbuffer_valid = false;      // the scratch set's records now describe this pass
  return SGI_OK;
}

int sgi_filter_shadow_map(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  if (!sgi_is_moment_tech(ctx->params.technique) || ctx->moments_tech != ctx->params.technique || ctx->moments_w != ctx->SW || ctx->moments_h != ctx->SH) {
    ctx->err = "sgi_filter_shadow_map: render the shadow map with a moment technique (VSM / ESM / EVSM / MSM) first";
    return SGI_ERR_INVALID;
  }
  if (!ctx->has_camera) { ctx->err = "sgi_filter_shadow_map: set the camera first (the filtered maps are window-sized, main.cpp:887-888)"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  int rc;
  const size_t bytes = (size_t)ctx->W * ctx->H * 16;
  if ((rc = ensure_buf(ctx, SGI_BUF_MOMENTS_X, bytes))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_MOMENTS_FILTERED, bytes))) return rc;
  // the previous frame's shadow pass (visibility stream) may still be sampling the filtered map
  if ((rc = sgi_join_vis(ctx))) return rc;
  sgi_wait_reads_of(ctx, SGI_BUF_MOMENTS_X, ctx->stream); sgi_wait_reads_of(ctx, SGI_BUF_MOMENTS_FILTERED, ctx->stream);
  const int slot = sgi_timing_begin(ctx, SGI_PASS_MOMENT_FILTER, ctx->stream);
  if ((rc = sgi_moments_filter_run(ctx, ctx->stream))) return rc;
  sgi_timing_end(ctx, SGI_PASS_MOMENT_FILTER, slot, ctx->stream);
  ctx->filtered_tech = ctx->params.technique; ctx->filtered_w = ctx->W; ctx->filtered_h = ctx->H;
  return SGI_OK;
}

void sgi_moment_quantization(float m[16], float m_inverse[16], float t[4]) { sgi_moments_quantization(m, m_inverse, t); }

int sgi_compute_visibility(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  if (sgi_is_moment_tech(ctx->params.technique)) {
    if (!ctx->gbuffer_valid || ctx->filtered_tech != ctx->params.technique || ctx->filtered_w != ctx->W || ctx->filtered_h != ctx->H) {
      ctx->err = "sgi_compute_visibility: render and filter the moment shadow map and render the G-buffer first";
      return SGI_ERR_INVALID;
    }
  } else {
    const bool fused = ctx->params.technique == SGI_TECH_MULTI_HARD && ctx->params.multi_fused;
    if (fused && (!ctx->ids_valid || !ctx->shadow_map_valid)) { ctx->err = "sgi_compute_visibility: multi_fused needs sgi_render_prim_ids and the shadow maps first"; return SGI_ERR_INVALID; }
    if (!fused && (!ctx->gbuffer_valid || !ctx->shadow_map_valid)) { ctx->err = "sgi_compute_visibility: render the shadow map and the G-buffer first"; return SGI_ERR_INVALID; }
    if (ctx->params.technique == SGI_TECH_MULTI_HARD && ctx->params.multi_partial == 2) {
      if (!fused || ctx->mask_total < 1 || (int)ctx->light_gid.size() != ctx->N) {
        ctx->err = "sgi_compute_visibility: multi_partial = 2 needs multi_fused and sgi_set_light_ids for the current lights";
        return SGI_ERR_INVALID;
      }
    }
  }
  cudaSetDevice(ctx->device);
  int rc;
  if (ctx->params.technique == SGI_TECH_MULTI_HARD && ctx->params.multi_partial == 2) {
    if ((rc = ensure_buf(ctx, SGI_BUF_LIGHT_MASK, sgi_padded_pixels(ctx) * (size_t)((ctx->mask_total + 7) / 8)))) return rc;
  }
  // With pass overlap on, the shadow pass goes to the visibility stream: it waits for everything queued on the main stream
  // so far (the depth pass, uploads) and for the G-buffer pass, and the main stream does NOT wait for it - the next frame's
  // depth / G-buffer passes start right away into the other instance of their targets (prepare_target_write).
  cudaStream_t vs = ctx->overlap_passes ? ctx->vis_stream : ctx->stream;
  if (ctx->overlap_passes) {
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_s2v, ctx->stream));
    SGI_CUDA(ctx, cudaStreamWaitEvent(vs, ctx->ev_s2v, 0));
    if (ctx->gbuf_in_flight) SGI_CUDA(ctx, cudaStreamWaitEvent(vs, ctx->ev_gbuf_done, 0));
  } else if ((rc = sgi_join_gbuffer(ctx))) return rc;
  {
    // If the visibility buffer is still being copied out by an asynchronous read (frame pipelining) and this call
    // produces the whole screen, write into a spare buffer instead of making the shadow kernel wait for the copy:
    // otherwise every frame's shadow pass would be chained behind the previous frame's device-to-host transfer.
    const int t = ctx->buf_read_ticket[SGI_BUF_VISIBILITY];
    const sgi_params& q = ctx->params;
    const bool whole = (q.rect_x1 <= q.rect_x0 || q.rect_y1 <= q.rect_y0) || (q.rect_x0 <= 0 && q.rect_y0 <= 0 && q.rect_x1 >= ctx->W && q.rect_y1 >= ctx->H);
    // (not when the caller holds the buffer's device pointer: a swap would make that pointer alternate between frames)
    if (t >= 0 && ctx->read_pending[t] && whole && !ctx->vis_exposed && cudaEventQuery(ctx->read_done[t]) == cudaErrorNotReady) {
      const size_t bytes = ctx->buf_bytes[SGI_BUF_VISIBILITY];
      if (ctx->vis_spare_bytes != bytes) {
        if (ctx->vis_spare) { sync_all_streams(ctx); cudaFree(ctx->vis_spare); }
        ctx->vis_spare = nullptr; ctx->vis_spare_bytes = 0; ctx->vis_spare_ticket = -1;
        SGI_CUDA(ctx, cudaMalloc(&ctx->vis_spare, bytes));
        ctx->vis_spare_bytes = bytes;
      }
      // the spare's own last copy-out (two frames ago) must have finished before it is overwritten
      if (ctx->vis_spare_ticket >= 0 && ctx->read_pending[ctx->vis_spare_ticket])
        cudaStreamWaitEvent(vs, ctx->read_done[ctx->vis_spare_ticket], 0);
      std::swap(ctx->buf[SGI_BUF_VISIBILITY], ctx->vis_spare);
      ctx->vis_spare_ticket = t;
      ctx->buf_read_ticket[SGI_BUF_VISIBILITY] = -1;
    } else {
      cudaGetLastError();
      sgi_wait_reads_of(ctx, SGI_BUF_VISIBILITY, vs);
    }
  }
  if (ctx->params.technique == SGI_TECH_EDTSM_NONCONS || ctx->params.technique == SGI_TECH_EDTSM_CONS) {
    if ((rc = ensure_buf(ctx, SGI_BUF_EDT_NEAREST, (size_t)ctx->W * ctx->H * 4))) return rc;
    sgi_wait_reads_of(ctx, SGI_BUF_EDT_NEAREST, vs);
  }
  // collectives still running on the buffers this pass reads / overwrites (gathered primitive ids; the previous frame's exchange
  // of the visibility buffer)
  sgi_wait_comm(ctx, SGI_BUF_PRIM_ID, vs); sgi_wait_comm(ctx, SGI_BUF_VISIBILITY, vs); sgi_wait_comm(ctx, SGI_BUF_LIGHT_MASK, vs);
  if (ctx->params.technique == SGI_TECH_MULTI_HARD && ctx->params.multi_partial == 2) sgi_wait_reads_of(ctx, SGI_BUF_LIGHT_MASK, vs);
  for (int b : {SGI_BUF_GBUF_POS, SGI_BUF_GBUF_NRM}) sgi_wait_comm(ctx, b, vs);
  int slot = sgi_timing_begin(ctx, SGI_PASS_VISIBILITY, vs);
  rc = sgi_shadow_run(ctx, vs);
  if (rc) return rc;
  sgi_timing_end(ctx, SGI_PASS_VISIBILITY, slot, vs);
  if (ctx->overlap_passes) {
    const int e = ctx->vis_ev_next; ctx->vis_ev_next = (ctx->vis_ev_next + 1) & 3;
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_vis[e], vs));
    ctx->vis_last = e; ctx->vis_in_flight = true;
    ctx->sm_reader_cur = e; ctx->gb_reader_cur = e;        // this pass reads the current instances of both target groups
    if (ctx->params.technique == SGI_TECH_MULTI_HARD && ctx->params.multi_fused) ctx->rec_reader = e;   // ... and the camera pass's records
    // a caller holding the visibility buffer's device pointer queues its own work on the context's stream: keep that
    // stream ordered after the pass for it
    if (ctx->vis_exposed) { if ((rc = sgi_join_vis(ctx))) return rc; }
  } else mark_gbuffer_use(ctx);
  return SGI_OK;
}

int sgi_shade_phong(sgi_ctx* ctx, const float clear_rgba[4]) {
  if (!ctx || !clear_rgba) return SGI_ERR_INVALID;
  if (!ctx->gbuffer_valid) { ctx->err = "sgi_shade_phong: render the G-buffer and compute the visibility first"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  int rc = sgi_join_gbuffer(ctx);
  if (rc) return rc;
  if ((rc = sgi_join_vis(ctx))) return rc;
  for (int b : {SGI_BUF_VISIBILITY, SGI_BUF_GBUF_POS, SGI_BUF_GBUF_NRM, SGI_BUF_GBUF_ALBEDO, SGI_BUF_SHADED}) sgi_wait_comm(ctx, b, ctx->stream);
  sgi_wait_reads_of(ctx, SGI_BUF_SHADED, ctx->stream);
  if (ctx->has_rgb && ctx->rgb_V != ctx->V) { ctx->err = "sgi_shade_phong: colours do not match the current mesh"; return SGI_ERR_INVALID; }
  rc = sgi_shade_run(ctx, clear_rgba);
  mark_gbuffer_use(ctx);
  return rc;
}

int sgi_compute_shadow_volume(sgi_ctx* ctx, const float light_pos[3]) {
  if (!ctx || !light_pos) return SGI_ERR_INVALID;
  if (!ctx->gbuffer_valid) { ctx->err = "sgi_compute_shadow_volume: render the G-buffer (depth pre-pass) first"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  int rc;
  const int zfail = ctx->params.sv_zfail ? 1 : 0, silhouette = ctx->params.sv_silhouette ? 1 : 0;
  const int per = zfail ? 8 : 6;                 // triangles per source triangle: 3 side quads (+ near cap + far cap)
  size_t px = sgi_padded_pixels(ctx);
  if ((rc = ensure_buf(ctx, SGI_BUF_SV_COUNT, px * 4))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_SV_STENCIL, px))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_SV_PRISM_XYZ, (size_t)(ctx->T > 0 ? ctx->T : 1) * 18 * 4))) return rc;
  if ((rc = ensure_buf(ctx, SGI_BUF_SV_PRISM_IDX, (size_t)(ctx->T > 0 ? ctx->T : 1) * per * 3 * 4))) return rc;
  if ((rc = sgi_join_gbuffer(ctx))) return rc;
  for (int b : {SGI_BUF_SV_COUNT, SGI_BUF_SV_STENCIL, SGI_BUF_SV_PRISM_XYZ, SGI_BUF_SV_PRISM_IDX}) { sgi_wait_reads_of(ctx, b, ctx->stream); sgi_wait_comm(ctx, b, ctx->stream); }
  unsigned long long* frags = nullptr;
  if (ctx->sv_count_fragments) {
    if (!ctx->d_sv_frags) SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_sv_frags, 8));
    SGI_CUDA(ctx, cudaMemsetAsync(ctx->d_sv_frags, 0, 8, ctx->stream));
    frags = ctx->d_sv_frags;
  }
  int slot = sgi_timing_begin(ctx, SGI_PASS_SHADOW_VOLUME, ctx->stream);
  if ((rc = sgi_sv_extrude_run(ctx, light_pos, (float*)ctx->buf[SGI_BUF_SV_PRISM_XYZ], (int32_t*)ctx->buf[SGI_BUF_SV_PRISM_IDX], per, silhouette))) return rc;
  SgiRasterJob job;
  memset(&job, 0, sizeof(job));
  job.mode = SGI_MODE_SVCOUNT;
  job.xyz = (const float*)ctx->buf[SGI_BUF_SV_PRISM_XYZ]; job.nrm = nullptr;
  job.idx = (const int32_t*)ctx->buf[SGI_BUF_SV_PRISM_IDX]; job.T = ctx->T * per;
  memcpy(job.mvp, ctx->cam_mvp, 64);
  job.W = ctx->W; job.H = ctx->H;
  job.scene_depth = (const float*)ctx->buf[SGI_BUF_CAM_DEPTH]; job.depth_func = ctx->params.sv_depth_func;
  job.count = (int32_t*)ctx->buf[SGI_BUF_SV_COUNT]; job.stencil = (uint8_t*)ctx->buf[SGI_BUF_SV_STENCIL];
  job.rx0 = ctx->params.rect_x0; job.ry0 = ctx->params.rect_y0; job.rx1 = ctx->params.rect_x1; job.ry1 = ctx->params.rect_y1;
  job.no_far_clip = zfail; job.sv_zfail = zfail; job.sv_caps = zfail; job.frag_counter = frags;
  if ((rc = sgi_raster_run(ctx, job, 0, ctx->stream))) return rc;
  sgi_timing_end(ctx, SGI_PASS_SHADOW_VOLUME, slot, ctx->stream);
  SGI_CUDA(ctx, cudaEventRecord(ctx->ev_geom_main, ctx->stream)); ctx->geom_main_recorded = true;
  mark_gbuffer_use(ctx);
  return SGI_OK;
}

int sgi_divide_selftest(sgi_ctx* ctx, uint64_t n, uint32_t seed, uint64_t* mismatches) {
  if (!ctx || !mismatches) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  unsigned long long m = 0;
  int rc = sgi_divide_selftest_run(ctx, (unsigned long long)n, seed, &m);
  *mismatches = (uint64_t)m;
  return rc;
}

// option "sv_count_fragments": covered prism fragments of the last sgi_compute_shadow_volume (the unit of work SURVEY 8(d) names)
int sgi_sv_fragments(sgi_ctx* ctx, int64_t* fragments) {
  if (!ctx || !fragments) return SGI_ERR_INVALID;
  *fragments = 0;
  if (!ctx->d_sv_frags) return SGI_OK;
  cudaSetDevice(ctx->device);
  unsigned long long v = 0;
  SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  SGI_CUDA(ctx, cudaMemcpy(&v, ctx->d_sv_frags, 8, cudaMemcpyDeviceToHost));
  *fragments = (int64_t)v;
  return SGI_OK;
}

// Orders the context's stream after every pass queued so far on the internal streams (G-buffer, shadow pass), without
// blocking the host: work the caller queues on that stream afterwards sees the finished frame.
int sgi_join(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  int rc = sgi_join_gbuffer(ctx);
  return rc ? rc : sgi_join_vis(ctx);
}

int sgi_synchronize(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  sgi_join_gbuffer(ctx); sgi_join_vis(ctx);
  SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  SGI_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  if (ctx->comm_stream) SGI_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
  if (ctx->comm_stream2) SGI_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream2));
  for (int b = 0; b < SGI_BUF_COUNT_; b++) ctx->comm_pending[b] = false;
  for (int k = 0; k < 4; k++) ctx->read_pending[k] = false;
  if (ctx->timing) sgi_timing_drain(ctx);
  return check_overflow(ctx);
}

int sgi_read(sgi_ctx* ctx, int32_t which, void* dst, size_t bytes) {
  if (!ctx || which < 0 || which >= SGI_BUF_COUNT_ || !dst) return SGI_ERR_INVALID;
  if (!ctx->buf[which] || bytes > ctx->buf_bytes[which]) { ctx->err = "sgi_read: buffer not produced yet or size too large"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  sgi_join_gbuffer(ctx); sgi_join_vis(ctx);
  sgi_wait_comm(ctx, which, ctx->stream);
  SGI_CUDA(ctx, cudaMemcpyAsync(dst, ctx->buf[which], bytes, cudaMemcpyDeviceToHost, ctx->stream));
  mark_gbuffer_use(ctx);
  SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return check_overflow(ctx);
}

int sgi_read_async(sgi_ctx* ctx, int32_t which, void* dst, size_t bytes, int32_t* ticket) {
  if (!ctx || which < 0 || which >= SGI_BUF_COUNT_ || !dst || !ticket) return SGI_ERR_INVALID;
  if (!ctx->buf[which] || bytes > ctx->buf_bytes[which]) { ctx->err = "sgi_read_async: buffer not produced yet or size too large"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  sgi_join_gbuffer(ctx);
  const int t = ctx->read_seq & 3;
  ctx->ticket_seq[t] = ctx->read_seq++;
  if (ctx->read_pending[t]) { SGI_CUDA(ctx, cudaEventSynchronize(ctx->read_done[t])); ctx->read_pending[t] = false; }   // back-pressure
  SGI_CUDA(ctx, cudaEventRecord(ctx->ev_ready, ctx->stream));                 // the buffer's producers are all on / joined into the main stream
  SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready, 0));
  if (ctx->vis_in_flight && ctx->vis_last >= 0)                                 // ... or on the visibility stream (not joined: the main stream keeps going)
    SGI_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_vis[ctx->vis_last], 0));
  sgi_wait_comm(ctx, which, ctx->copy_stream);
  SGI_CUDA(ctx, cudaMemcpyAsync(dst, ctx->buf[which], bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
  SGI_CUDA(ctx, cudaEventRecord(ctx->read_done[t], ctx->copy_stream));
  ctx->read_pending[t] = true;
  ctx->buf_read_ticket[which] = t;
  *ticket = t;
  return SGI_OK;
}

int sgi_read_wait(sgi_ctx* ctx, int32_t ticket) {
  if (!ctx || ticket < 0 || ticket > 3) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->read_pending[ticket]) { SGI_CUDA(ctx, cudaEventSynchronize(ctx->read_done[ticket])); ctx->read_pending[ticket] = false; }
  return check_overflow(ctx, ctx->ticket_seq[ticket]);
}

int sgi_device_ptr(sgi_ctx* ctx, int32_t which, void** dptr, size_t* bytes) {
  if (!ctx || which < 0 || which >= SGI_BUF_COUNT_ || !dptr) return SGI_ERR_INVALID;
  sgi_join_gbuffer(ctx); sgi_join_vis(ctx);   // work queued on the main stream after this call sees finished passes
  sgi_wait_comm(ctx, which, ctx->stream);
  if (which == SGI_BUF_GBUF_POS || which == SGI_BUF_GBUF_NRM || which == SGI_BUF_CAM_DEPTH || which == SGI_BUF_GBUF_ALBEDO) ctx->gbuf_exposed = true;
  if (which == SGI_BUF_SHADOW_MAP) ctx->sm_exposed = true;     // a lent-out pointer pins the instance: no more switching
  if (which == SGI_BUF_VISIBILITY || which == SGI_BUF_EDT_NEAREST) ctx->vis_exposed = true;
  *dptr = ctx->buf[which];
  if (bytes) *bytes = ctx->buf_bytes[which];
  return ctx->buf[which] ? SGI_OK : SGI_ERR_INVALID;
}

int sgi_alloc_host(void** p, size_t bytes) {
  if (!p || bytes == 0) return SGI_ERR_INVALID;
  return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? SGI_OK : SGI_ERR_NOMEM;
}
int sgi_free_host(void* p) { return (p && cudaFreeHost(p) == cudaSuccess) ? SGI_OK : SGI_ERR_INVALID; }
int sgi_register_host(void* p, size_t bytes) {
  if (!p || bytes == 0) return SGI_ERR_INVALID;
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return SGI_ERR_NOMEM; }
  return SGI_OK;
}
int sgi_unregister_host(void* p) {
  if (!p) return SGI_ERR_INVALID;
  if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return SGI_ERR_INVALID; }
  return SGI_OK;
}

int sgi_set_option(sgi_ctx* ctx, const char* name, int32_t value) {
  if (!ctx || !name) return SGI_ERR_INVALID;
  if (!strcmp(name, "vis_staged")) ctx->vis_staged = value != 0;
  else if (!strcmp(name, "overlap_passes")) { sgi_join_gbuffer(ctx); sgi_join_vis(ctx); ctx->overlap_passes = value != 0; }
  else if (!strcmp(name, "tile_order")) ctx->tile_order = value ? 1 : 0;
  else if (!strcmp(name, "tile_split")) ctx->tile_split = value < 0 ? 0 : value;
  else if (!strcmp(name, "borrow_pinned")) ctx->borrow_pinned = value ? 1 : 0;
  else if (!strcmp(name, "sv_tile_cull")) ctx->sv_tile_cull = value ? 1 : 0;
  else if (!strcmp(name, "rbssm_compact")) ctx->rbssm_compact = value ? 1 : 0;
  else if (!strcmp(name, "pcss_early_out")) ctx->pcss_early_out = value ? 1 : 0;
  else if (!strcmp(name, "tile_bulk_flush")) ctx->tile_bulk_flush = value ? 1 : 0;
  else if (!strcmp(name, "pdl")) ctx->pdl = value ? 1 : 0;
  else if (!strcmp(name, "comm_split")) ctx->comm_split = value ? 1 : 0;
  else if (!strcmp(name, "tile_few_walk")) ctx->tile_few_walk = value ? 1 : 0;
  else if (!strcmp(name, "sv_split_lists")) ctx->sv_split_lists = value ? 1 : 0;
  else if (!strcmp(name, "tile_direct")) ctx->tile_direct = value;
  else if (!strcmp(name, "tile_bin_big")) ctx->tile_bin_big = value;
  else if (!strcmp(name, "tile_bin_big_work")) ctx->tile_bin_big_work = value;
  else if (!strcmp(name, "tile_static_items")) ctx->tile_static_items = value < 0 || value > 2 ? 2 : value;
  else if (!strcmp(name, "tile_refresh_full")) ctx->tile_refresh_full = value < 0 || value > 2 ? 2 : value;
  else if (!strcmp(name, "sv_count_fragments")) ctx->sv_count_fragments = value ? 1 : 0;
  else if (!strcmp(name, "vis_minmax_cull")) { ctx->vis_minmax_cull = value ? 1 : 0; ctx->mm_valid = false; }
  else if (!strcmp(name, "tile_threads")) {
    if (value != 0 && value != 128 && value != 256 && value != 512 && value != 1024) { ctx->err = "tile_threads must be 0, 128 (depth pass only), 256, 512 or 1024"; return SGI_ERR_INVALID; }
    ctx->tile_threads = value;
  } else { ctx->err = std::string("unknown option ") + name; return SGI_ERR_INVALID; }
  return SGI_OK;
}

int sgi_enable_timing(sgi_ctx* ctx, int32_t on) {
  if (!ctx) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (on && !ctx->ev[0][0][0]) {
    for (int p = 0; p < SGI_PASS_COUNT_; p++)
      for (int k = 0; k < SGI_EV_RING; k++) {
        SGI_CUDA(ctx, cudaEventCreate(&ctx->ev[p][k][0]));
        SGI_CUDA(ctx, cudaEventCreate(&ctx->ev[p][k][1]));
      }
  }
  if (!on && ctx->timing) { cudaStreamSynchronize(ctx->stream); sgi_timing_drain(ctx); }
  ctx->timing = on != 0;
  return SGI_OK;
}

int sgi_pass_time_ms(sgi_ctx* ctx, int32_t pass, double* total_ms, int64_t* calls) {
  if (!ctx || pass < 0 || pass >= SGI_PASS_COUNT_) return SGI_ERR_INVALID;
  if (total_ms) *total_ms = ctx->pass_ms[pass];
  if (calls) *calls = ctx->pass_calls[pass];
  return SGI_OK;
}

int sgi_reset_timing(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  if (ctx->timing) { cudaStreamSynchronize(ctx->stream); sgi_timing_drain(ctx); }
  for (int p = 0; p < SGI_PASS_COUNT_; p++) { ctx->pass_ms[p] = 0; ctx->pass_calls[p] = 0; }
  return SGI_OK;
}

int sgi_kernel_launches(sgi_ctx* ctx, int64_t* launches) {
  if (!ctx || !launches) return SGI_ERR_INVALID;
  *launches = ctx->launches;
  return SGI_OK;
}

const char* sgi_last_error(sgi_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

}  // extern "C"
