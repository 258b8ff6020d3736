// sgi_comm.cu — multi-GPU exchange of the C ABI (include/shadowgi.h: sgi_comm_*, sgi_gather, sgi_reduce_lights).
// The reference is single-GPU (no collective anywhere); what is partitioned here is its frame:
//   screen tiles  every rank evaluates a strip of the screen (sgi_params.rect_*), one ncclAllGather in place puts the image together
//   lights        renderMonteCarlo (SoftShadowMapping/src/main.cpp:756-811): every rank renders and samples the depth maps of its
//                 own lights; the un-normalised sums are reduce-scattered in place and divided by the light count
//                 (AccurateSoftShadow.frag:127), so each rank ends with the final visibility of its strip
// NCCL is loaded at run time (dlopen libnccl.so.2: in a torch process that is the copy torch already mapped), so the library has no
// link-time dependency on it and loads on machines without NCCL.  The buffers that take part are padded to nranks equal strips
// (ceil(H / nranks) rows each), so every exchange is ONE collective on the buffer itself: no packing, no staging copy.
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <nccl.h>
#include "sgi_internal.cuh"

namespace {

struct NcclApi {
  void* dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;     // NCCL >= 2.18; optional
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return dl && GetUniqueId && CommInitRank && CommDestroy && AllGather && ReduceScatter && GetErrorString; }
};

NcclApi& nccl() {
  static NcclApi api;
  if (!api.dl) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.dl) break; }
    if (api.dl) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.dl, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.dl, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.dl, "ncclCommDestroy");
      api.CommSplit = (decltype(api.CommSplit))dlsym(api.dl, "ncclCommSplit");
      api.AllGather = (decltype(api.AllGather))dlsym(api.dl, "ncclAllGather");
      api.ReduceScatter = (decltype(api.ReduceScatter))dlsym(api.dl, "ncclReduceScatter");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.dl, "ncclGetErrorString");
    }
  }
  return api;
}

#define SGI_NCCL(ctx, expr)                                                                      \
  do {                                                                                           \
    ncclResult_t r__ = (expr);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      (ctx)->err = std::string(#expr) + ": " + nccl().GetErrorString(r__);                       \
      return SGI_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

// AccurateSoftShadow.frag:127: accShadow / count, on the rows of this rank's strip that lie inside the screen
__global__ void __launch_bounds__(256) k_div_rows(float* __restrict__ v, size_t n, float count) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) v[i] = v[i] / count;
}

size_t elem_bytes(int which) {
  switch (which) {
    case SGI_BUF_GBUF_POS: case SGI_BUF_GBUF_NRM: case SGI_BUF_GBUF_ALBEDO: case SGI_BUF_SHADED: return 16;
    case SGI_BUF_CAM_DEPTH: case SGI_BUF_VISIBILITY: case SGI_BUF_SV_COUNT: case SGI_BUF_PRIM_ID: return 4;
    case SGI_BUF_SV_STENCIL: return 1;
    default: return 0;                       // not a screen-sized target
  }
}

}  // namespace

int sgi_strip_rows(const sgi_ctx* ctx) { return ctx->comm_n > 1 ? (ctx->H + ctx->comm_n - 1) / ctx->comm_n : ctx->H; }
size_t sgi_padded_pixels(const sgi_ctx* ctx) { return (size_t)sgi_strip_rows(ctx) * (ctx->comm_n > 1 ? ctx->comm_n : 1) * ctx->W; }

// a consumer (or the next writer) of `which` on `stream` must come after a collective still running on it
void sgi_wait_comm(sgi_ctx* ctx, int which, cudaStream_t stream) {
  if (ctx->comm_pending[which] && ctx->ev_comm_done[which]) cudaStreamWaitEvent(stream, ctx->ev_comm_done[which], 0);
}

extern "C" {

int sgi_comm_unique_id(void* id128, size_t bytes) {
  if (!id128 || bytes < sizeof(ncclUniqueId)) return SGI_ERR_INVALID;
  if (!nccl().ok()) return SGI_ERR_NO_DEVICE;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return SGI_ERR_CUDA;
  memcpy(id128, &id, sizeof(id));
  return SGI_OK;
}

int sgi_comm_init(sgi_ctx* ctx, const void* id128, size_t bytes, int32_t rank, int32_t nranks) {
  if (!ctx || !id128 || bytes < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks) { if (ctx) ctx->err = "sgi_comm_init: bad arguments"; return SGI_ERR_INVALID; }
  if (!nccl().ok()) { ctx->err = "sgi_comm_init: libnccl.so.2 not found"; return SGI_ERR_NO_DEVICE; }
  if (ctx->nccl_comm) { ctx->err = "sgi_comm_init: already initialised"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  sgi_synchronize(ctx);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  SGI_NCCL(ctx, nccl().CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm; ctx->comm_rank = rank; ctx->comm_n = nranks;
  if (!ctx->comm_stream) SGI_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  // Option "comm_split" (set before this call): the reductions get a communicator and a stream of their own (ncclCommSplit, every
  // rank in one colour), so that a frame's id gather does not queue behind the previous frame's reduction on the one stream.
  // Measured at 8 GPUs it did not pay (the collectives' CTAs compete with the raster kernels for the SMs either way): off by default.
  ctx->nccl_comm2 = nullptr;
  if (nranks > 1 && ctx->comm_split && nccl().CommSplit) {
    ncclComm_t c2 = nullptr;
    if (nccl().CommSplit(comm, 0, rank, &c2, nullptr) == ncclSuccess && c2) {
      ctx->nccl_comm2 = c2;
      if (!ctx->comm_stream2) SGI_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream2, cudaStreamNonBlocking));
    }
  }
  if (!ctx->ev_comm_in) SGI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm_in, cudaEventDisableTiming));
  // the strip layout depends on the rank count: targets sized before this call are re-made (padded) by the next sgi_set_camera
  ctx->W = ctx->H = 0; ctx->has_camera = false; ctx->gbuffer_valid = false; ctx->ids_valid = false;
  return SGI_OK;
}

int sgi_comm_destroy(sgi_ctx* ctx) {
  if (!ctx) return SGI_ERR_INVALID;
  if (ctx->nccl_comm) {
    cudaSetDevice(ctx->device);
    sgi_synchronize(ctx);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    if (ctx->comm_stream2) cudaStreamSynchronize(ctx->comm_stream2);
    if (ctx->nccl_comm2) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm2);
    ctx->nccl_comm2 = nullptr;
    nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  ctx->comm_rank = 0; ctx->comm_n = 1;
  for (int b = 0; b < SGI_BUF_COUNT_; b++) ctx->comm_pending[b] = false;
  ctx->W = ctx->H = 0; ctx->has_camera = false; ctx->gbuffer_valid = false; ctx->ids_valid = false;
  return SGI_OK;
}

int sgi_comm_strip(sgi_ctx* ctx, int32_t rank, int32_t* row0, int32_t* row1) {
  if (!ctx || !row0 || !row1 || rank < 0 || rank >= (ctx->comm_n > 1 ? ctx->comm_n : 1) || ctx->H <= 0) { if (ctx) ctx->err = "sgi_comm_strip: set the camera first"; return SGI_ERR_INVALID; }
  const int rows = sgi_strip_rows(ctx);
  *row0 = rank * rows < ctx->H ? rank * rows : ctx->H;
  *row1 = (rank + 1) * rows < ctx->H ? (rank + 1) * rows : ctx->H;
  return SGI_OK;
}

// Orders the communication stream behind everything that may have produced `which`: the main stream as queued so far, the
// G-buffer / id pass on the auxiliary stream, the shadow pass on the visibility stream.
static int comm_begin(sgi_ctx* ctx, int which, cudaStream_t cs) {
  SGI_CUDA(ctx, cudaEventRecord(ctx->ev_comm_in, ctx->stream));
  SGI_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_comm_in, 0));
  if (ctx->gbuf_in_flight) SGI_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_gbuf_done, 0));
  if (ctx->vis_in_flight && ctx->vis_last >= 0) SGI_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_vis[ctx->vis_last], 0));
  sgi_wait_comm(ctx, which, cs);                         // an earlier collective on the same buffer (it may be on the other stream)
  sgi_wait_reads_of(ctx, which, cs);                     // an asynchronous copy-out of the buffer still in flight
  if (!ctx->ev_comm_done[which]) SGI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm_done[which], cudaEventDisableTiming));
  return SGI_OK;
}

int sgi_gather(sgi_ctx* ctx, int32_t which) {
  if (!ctx || which < 0 || which >= SGI_BUF_COUNT_) return SGI_ERR_INVALID;
  if (ctx->comm_n <= 1) return SGI_OK;                   // one rank: the strip is the screen
  const size_t eb = elem_bytes(which);
  if (!eb || !ctx->has_camera) { ctx->err = "sgi_gather: not a screen-sized target (or no camera set)"; return SGI_ERR_INVALID; }
  const size_t strip = (size_t)sgi_strip_rows(ctx) * ctx->W * eb;
  if (!ctx->buf[which] || ctx->buf_bytes[which] < strip * ctx->comm_n) { ctx->err = "sgi_gather: buffer not produced yet (or sized before sgi_comm_init)"; return SGI_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  int rc = comm_begin(ctx, which, ctx->comm_stream);
  if (rc) return rc;
  char* base = (char*)ctx->buf[which];
  SGI_NCCL(ctx, nccl().AllGather(base + strip * ctx->comm_rank, base, strip, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->comm_stream));
  SGI_CUDA(ctx, cudaEventRecord(ctx->ev_comm_done[which], ctx->comm_stream));
  ctx->comm_pending[which] = true;
  return SGI_OK;
}

int sgi_set_light_ids(sgi_ctx* ctx, int32_t n, const int32_t* ids, int32_t total_lights) {
  if (!ctx || n < 0 || n > 32 || total_lights < n || total_lights > 32 || (n > 0 && !ids)) { if (ctx) ctx->err = "sgi_set_light_ids: at most 32 lights in the whole set"; return SGI_ERR_INVALID; }
  unsigned int seen = 0u;
  for (int k = 0; k < n; k++) {
    if (ids[k] < 0 || ids[k] >= total_lights || ((seen >> ids[k]) & 1u)) { ctx->err = "sgi_set_light_ids: indices must be distinct and below the total"; return SGI_ERR_INVALID; }
    seen |= 1u << ids[k];
  }
  ctx->mask_total = total_lights;
  if ((int)ctx->light_gid.size() == n && std::equal(ids, ids + n, ctx->light_gid.begin()) && ctx->d_light_gid) return SGI_OK;     // (called every frame)
  cudaSetDevice(ctx->device);
  if (!ctx->d_light_gid) SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_light_gid, 32 * sizeof(int)));
  sgi_join_vis(ctx);                                             // a pass still reading the previous table
  ctx->light_gid.assign(ids, ids + n);
  int tmp[32] = {0};
  for (int k = 0; k < n; k++) tmp[k] = ids[k];
  SGI_CUDA(ctx, cudaMemcpyAsync(ctx->d_light_gid, tmp, sizeof(tmp), cudaMemcpyHostToDevice, ctx->stream));
  SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));           // tmp is on the stack
  return SGI_OK;
}

static cudaStream_t reduce_stream(sgi_ctx* ctx) { return ctx->nccl_comm2 ? ctx->comm_stream2 : ctx->comm_stream; }
static ncclComm_t reduce_comm(sgi_ctx* ctx) { return (ncclComm_t)(ctx->nccl_comm2 ? ctx->nccl_comm2 : ctx->nccl_comm); }

// light sharding with lit masks: the planes of all ranks are summed (ncclReduceScatter on bytes, disjoint bits), then this rank
// accumulates the visibility of its strip from the union (k_mask_resolve)
static int reduce_light_masks(sgi_ctx* ctx, int32_t total_lights) {
  const int which = SGI_BUF_LIGHT_MASK;
  if (total_lights != ctx->mask_total) { ctx->err = "sgi_reduce_lights: total differs from sgi_set_light_ids"; return SGI_ERR_INVALID; }
  if (!ctx->buf[which] || !ctx->buf[SGI_BUF_VISIBILITY]) { ctx->err = "sgi_reduce_lights: compute the lit masks first"; return SGI_ERR_INVALID; }
  const size_t part = (size_t)sgi_strip_rows(ctx) * ctx->W * ((total_lights + 7) / 8);      // bytes per rank
  unsigned char* base = (unsigned char*)ctx->buf[which];
  cudaStream_t st;
  int rc;
  if (ctx->comm_n > 1) {
    st = reduce_stream(ctx);
    if (ctx->buf_bytes[which] < part * ctx->comm_n) { ctx->err = "sgi_reduce_lights: mask buffer sized before sgi_comm_init"; return SGI_ERR_INVALID; }
    if ((rc = comm_begin(ctx, which, st))) return rc;
    SGI_NCCL(ctx, nccl().ReduceScatter(base, base + part * ctx->comm_rank, part, ncclUint8, ncclSum, reduce_comm(ctx), st));
    sgi_wait_reads_of(ctx, SGI_BUF_VISIBILITY, st);              // the strip's visibility is written on this stream
    sgi_wait_comm(ctx, SGI_BUF_VISIBILITY, st);
  } else {
    if ((rc = sgi_join_vis(ctx))) return rc;
    st = ctx->stream;
  }
  int32_t r0 = 0, r1 = 0;
  sgi_comm_strip(ctx, ctx->comm_rank, &r0, &r1);
  if ((rc = sgi_mask_resolve_run(ctx, r0, r1, st))) return rc;
  if (ctx->comm_n > 1) {
    for (int b : {which, (int)SGI_BUF_VISIBILITY}) {
      if (!ctx->ev_comm_done[b]) SGI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm_done[b], cudaEventDisableTiming));
      SGI_CUDA(ctx, cudaEventRecord(ctx->ev_comm_done[b], st));
      ctx->comm_pending[b] = true;
    }
  }
  return SGI_OK;
}

int sgi_reduce_lights(sgi_ctx* ctx, int32_t total_lights) {
  if (!ctx || total_lights <= 0) return SGI_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->params.multi_partial == 2) return ctx->has_camera ? reduce_light_masks(ctx, total_lights) : SGI_ERR_INVALID;
  if (!ctx->has_camera || !ctx->buf[SGI_BUF_VISIBILITY]) { ctx->err = "sgi_reduce_lights: compute the partial visibility first"; return SGI_ERR_INVALID; }
  const int which = SGI_BUF_VISIBILITY;
  const size_t strip = (size_t)sgi_strip_rows(ctx) * ctx->W;       // floats
  float* base = (float*)ctx->buf[which];
  cudaStream_t st = ctx->comm_n > 1 ? reduce_stream(ctx) : nullptr;
  if (ctx->comm_n > 1) {
    if (ctx->buf_bytes[which] < strip * ctx->comm_n * 4) { ctx->err = "sgi_reduce_lights: visibility buffer sized before sgi_comm_init"; return SGI_ERR_INVALID; }
    int rc = comm_begin(ctx, which, st);
    if (rc) return rc;
    SGI_NCCL(ctx, nccl().ReduceScatter(base, base + strip * ctx->comm_rank, strip, ncclFloat, ncclSum, reduce_comm(ctx), st));
  } else {
    int rc = sgi_join_vis(ctx);                                     // one rank: divide on the context's stream
    if (rc) return rc;
    st = ctx->stream;
  }
  int32_t r0 = 0, r1 = 0;
  sgi_comm_strip(ctx, ctx->comm_rank, &r0, &r1);
  const size_t n = (size_t)(r1 - r0) * ctx->W;
  if (n) {
    k_div_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(base + (size_t)r0 * ctx->W, n, (float)total_lights);
    ctx->launches++;
  }
  SGI_CUDA(ctx, cudaGetLastError());
  if (ctx->comm_n > 1) {
    SGI_CUDA(ctx, cudaEventRecord(ctx->ev_comm_done[which], st));
    ctx->comm_pending[which] = true;
  }
  return SGI_OK;
}

}  // extern "C"
