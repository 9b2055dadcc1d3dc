// C ABI of libray_b200.so: the Futhark-compatible surface of include/ray.h (what futhark/main.c links
// against) plus the extensions of include/ray_b200.h.  No CPU fallback: without a usable CUDA device
// futhark_context_new reports an error and every entry point fails.
#include "../../include/ray_b200.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <vector>

#include "bvh_build.h"
#include "render_params.h"
#include "scene_host.h"

using namespace rayb200;

// ------------------------------------------------------------------------------------------ objects
struct futhark_context_config {
  int device = 0;
  int debugging = 0, profiling = 0, logging = 0;
  int32_t spp = 1;
  int32_t kernel = RAY_B200_KERNEL_AUTO;
  int32_t rank = 0, world = 1;
  int32_t gpus = 1;  // > 1: this ONE process drives that many devices (RAY_GPUS; the drop-in multi-GPU mode of main.c)
  int32_t blocks_per_sm = 4, smem_budget = 48 * 1024, refill_min = 8, tail_from = 8;
  int32_t wq_warps = 0 /* 0 = per scene: 32, or 24 for trees far larger than the caches */, wq_k = 1, wq_spread = 1, wq_packet = -1, wq_refill = 1, wq_ncap = 512, permute = 1, host_build = 0;
  int32_t heavy_first = 0;   // pull long-path tiles to the front of the claim order: 0 off (default: the probe pass costs more than the tail it saves on one GPU, see profiles/), 1/2/4 = probe pixels per tile, -1 = on when spp > 1
  int32_t probe_segments = 8;
  std::string cache_file;
};

struct futhark_context {
  futhark_context_config cfg;
  std::mutex mu;
  char *error = nullptr;
  FILE *log = stderr;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  bool have_timing = false;
  int sm_count = 0, max_smem_optin = 0;
  unsigned long long *counters = nullptr;   // device [4]
  unsigned long long *warp_trace = nullptr; // device [1 + SMs * kWqMaxWarps] when tracing is on (ray_b200_context_trace_warps)
  int trace_warps = 0;                      // warps per CTA of the last traced launch
  float *offsets = nullptr;                 // device sample-offset table of the frame being set up (an entry of offset_tables)
  int32_t offsets_spp = 0;
  struct OffsetTable { int32_t spp; float *dev; };
  std::vector<OffsetTable> offset_tables;   // per-spp cache (ensure_offsets)
  int64_t launches = 0;
  WavefrontBuffers wf;                      // ray queues of the wavefront kernel (grown on demand)
  struct PinnedBlock { unsigned char *ptr; size_t bytes; cudaEvent_t last_use; };
  std::vector<PinnedBlock> pinned_cache;    // page-locked upload buffers of freed prepared scenes, reused by the next prepare_scene
  BvhBuildResult *d_build_result = nullptr, *h_build_result = nullptr;  // device scratch / page-locked host mirror
  int32_t plan_wq_warps = 0, plan_wq_packet = 0;
  // single-process multi-GPU (cfg.gpus > 1): one helper context per extra device; this context is rank 0 and owns them
  std::vector<futhark_context *> peers;
  bool is_peer = false;
  int32_t *peer_tiles = nullptr;        // helper context: this device's compact tile buffer (grow-only)
  size_t peer_tiles_bytes = 0;
  cudaEvent_t peer_done = nullptr;      // helper context: its shard has been rendered
  int32_t *gathered = nullptr;          // rank 0: [gpus][tiles_padded][32] staging for the de-tiling kernel (grow-only)
  size_t gathered_bytes = 0;
  cudaEvent_t ev_gathered = nullptr;    // rank 0: the peer copies of the last frame have read every helper's peer_tiles
  bool gather_pending = false;
  // Per-render scratch.  Lane 0 runs on the context's stream; lane 1 (own stream, created on first use) lets
  // ray_b200_render_batch keep two frames in flight so that one frame's tail is covered by the next frame's start.
  struct Lane {
    cudaStream_t stream = nullptr;             // lane 0: mirrors ctx->stream at each use
    int32_t *work_cursor = nullptr;            // device
    float4 *sample_buf = nullptr;              // warp-queue kernel, spp > 1: per-warp finished-sample colours
    size_t sample_buf_bytes = 0;
    unsigned char *tile_order_block = nullptr; // heavy-first claim order: keys, sorted keys, ids, order, cub temp (one allocation)
    size_t tile_order_bytes = 0;
    TileOrderBuffers tile_order_plan{};
  } lanes[2];
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool profiling_paused = false;
  int64_t renders = 0;
  bool ok = false;
};

struct futhark_opaque_scene {
  HostScene host;
};

struct futhark_opaque_prepared_scene {
  HostScene host;   // kept for store/restore and re-preparation
  int64_t h = 0, w = 0;
  CameraRec cam;
  float root_box[6];
  int32_t max_depth = 0, stale_nodes = 0, refit_sweeps = 0;
  int64_t n = 0;
  std::vector<futhark_opaque_prepared_scene *> peer_prepared;  // single-process multi-GPU: the same scene on every helper device
  DeviceBvh dev;    // everything resident in HBM (one stream-ordered allocation): packed BVH2C + the Karras-order LBVH
  unsigned char *pinned = nullptr;     // page-locked upload buffer (sphere records, or the host-built arrays)
  size_t pinned_bytes = 0;
  cudaEvent_t pinned_event = nullptr;  // completion of the last H2D copy that read `pinned`
  bool host_built = false;
};

struct futhark_i32_2d {
  int32_t *dev = nullptr;
  int64_t shape[2] = {0, 0};
  bool owned = true;
};

// ------------------------------------------------------------------------------------------ helpers
namespace {

void set_error(futhark_context *ctx, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  free(ctx->error);
  ctx->error = strdup(buf);
  if (ctx->cfg.logging || ctx->cfg.debugging) fprintf(ctx->log ? ctx->log : stderr, "[ray_b200] error: %s\n", buf);
}

#define CUDA_TRY(ctx, call)                                                                     \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      set_error(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  return atoi(v);
}

int parse_kernel(const char *v, int dflt) {
  if (!v || !*v) return dflt;
  if (!strcmp(v, "auto")) return RAY_B200_KERNEL_AUTO;
  if (!strcmp(v, "mega")) return RAY_B200_KERNEL_MEGA;
  if (!strcmp(v, "persistent")) return RAY_B200_KERNEL_PERSISTENT;
  if (!strcmp(v, "wavefront")) return RAY_B200_KERNEL_WAVEFRONT;
  if (!strcmp(v, "warpqueue")) return RAY_B200_KERNEL_WARPQUEUE;
  if (!strcmp(v, "streamqueue")) return RAY_B200_KERNEL_STREAMQUEUE;
  return atoi(v);
}

const char *kTuningNames[] = {"kernel", "spp", "blocks_per_sm", "smem_budget", "refill_min", "tail_from", "wq_warps", "wq_k", "wq_spread", "wq_packet", "wq_refill", "wq_ncap", "permute", "heavy_first", "probe_segments", "host_build", "rank", "world", "gpus"};
constexpr int kNumTuning = sizeof(kTuningNames) / sizeof(kTuningNames[0]);

bool bad_ctx(futhark_context *ctx) { return ctx == nullptr || !ctx->ok; }

// Device copy of the sample-offset table for `spp` (offset (0,0) at sample 0).  Tables are cached per spp
// (render_batch mixes sample counts) and never freed while a frame may still read them.  A new table is copied from a
// page-locked staging buffer on the context's stream and the stream is synchronised before it is handed out, so the
// data has landed whichever lane's stream the reading kernel is launched on (a pageable cudaMemcpy on the legacy stream
// only guarantees staging, and nothing would order a non-blocking stream after its DMA).
int ensure_offsets(futhark_context *ctx, int32_t spp) {
  for (auto &e : ctx->offset_tables)
    if (e.spp == spp) { ctx->offsets = e.dev; ctx->offsets_spp = spp; return 0; }
  std::vector<float> table;
  sample_offsets(spp, table);
  const size_t bytes = table.size() * sizeof(float);
  float *stage = nullptr, *dev = nullptr;
  CUDA_TRY(ctx, cudaMallocHost(&stage, bytes));
  memcpy(stage, table.data(), bytes);
  cudaError_t e = cudaMalloc(&dev, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dev, stage, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFreeHost(stage);
  if (e != cudaSuccess) { if (dev) cudaFree(dev); set_error(ctx, "sample-offset table upload failed: %s", cudaGetErrorString(e)); return 1; }
  if (ctx->offset_tables.size() >= 32) {  // evict the oldest; every lane is drained first, so no frame still reads it
    for (auto &L : ctx->lanes) if (L.stream) cudaStreamSynchronize(L.stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->offset_tables.front().dev);
    ctx->offset_tables.erase(ctx->offset_tables.begin());
  }
  ctx->offset_tables.push_back({spp, dev});
  ctx->offsets = dev;
  ctx->offsets_spp = spp;
  return 0;
}

int64_t tiles_total(int64_t h, int64_t w) { return ((h + kTileH - 1) / kTileH) * ((w + kTileW - 1) / kTileW); }
int64_t tiles_of_rank(int64_t h, int64_t w, int32_t rank, int32_t world) {
  const int64_t t = tiles_total(h, w);
  return t / world + ((t % world) > rank ? 1 : 0);
}

int resolve_kernel(const futhark_context *ctx);

// Fills the kernel parameter block for one frame.
int fill_params(futhark_context *ctx, const futhark_opaque_prepared_scene *p, int64_t h, int64_t w, int32_t spp,
                int32_t rank, int32_t world, int32_t *out_pix, float *out_rgb, bool tile_major, RenderParams &P) {
  if (!p || !p->dev.nodes) { set_error(ctx, "render: invalid prepared scene"); return 1; }
  if (h <= 0 || w <= 0 || h > 65536 || w > 65536 || ((h + 3) / 4) * ((w + 7) / 8) > ((int64_t)1 << 25)) {
    // work items are 32-bit: at most 2^30 (padded) pixels per frame
    set_error(ctx, "render: bad image size %lldx%lld", (long long)h, (long long)w);
    return 1;
  }
  if (spp < 1) { set_error(ctx, "render: spp must be >= 1"); return 1; }
  if (world < 1 || rank < 0 || rank >= world) { set_error(ctx, "render: bad shard %d/%d", rank, world); return 1; }
  if (p->max_depth > kStackSize - 1) { set_error(ctx, "render: BVH depth %d exceeds the traversal stack", p->max_depth); return 1; }
  if (ensure_offsets(ctx, spp)) return 1;
  memset(&P, 0, sizeof P);
  P.nodes = p->dev.nodes; P.nodes_soa = p->dev.nodes_soa; P.geom = p->dev.geom; P.colour = p->dev.colour;
  P.n_inner = (int32_t)(p->n - 1); P.n_leaves = (int32_t)p->n;
  P.max_depth = p->max_depth;
  memcpy(P.root_box, p->root_box, sizeof P.root_box);
  // The camera depends on the aspect ratio w/h given to prepare_scene (ray.fut:243-244); render's own
  // h, w only set the pixel grid (ray.fut:246-247) — exactly as in the reference.
  memcpy(P.cam, &p->cam, sizeof P.cam);
  P.W = (int32_t)w; P.H = (int32_t)h; P.spp = spp; P.inv_spp = 1.0f / (float)spp;
  P.offsets = ctx->offsets;
  P.out_pix = out_pix; P.out_rgb = out_rgb; P.tile_major = tile_major ? 1 : 0;
  P.rank = rank; P.world = world;
  P.tiles_x = (int32_t)((w + kTileW - 1) / kTileW); P.tiles_y = (int32_t)((h + kTileH - 1) / kTileH);
  P.n_tiles = tiles_total(h, w);
  P.local_tiles = tiles_of_rank(h, w, rank, world);
  P.n_chunks = (int32_t)((P.local_tiles + 63) / 64);
  {  // stride ~ golden ratio * n_chunks, made coprime to n_chunks
    auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
    int64_t st = std::max<int64_t>(1, (int64_t)(0.6180339887 * (double)P.n_chunks));
    while (gcd(st, P.n_chunks) != 1) st++;
    P.chunk_stride = (int32_t)(P.n_chunks > 1 ? st % P.n_chunks : 0);
    if (P.n_chunks > 1 && P.chunk_stride == 0) P.chunk_stride = 1;
    if (!ctx->cfg.permute) { P.chunk_stride = 1; }
    // modular inverse (extended Euclid): tile chunk q is claimed at position q * inv mod n_chunks (heavy-first order)
    int64_t r0 = P.n_chunks, r1 = P.chunk_stride, t0 = 0, t1 = 1;
    while (r1 != 0) {
      const int64_t qd = r0 / r1, r2 = r0 - qd * r1, t2 = t0 - qd * t1;
      r0 = r1; r1 = r2; t0 = t1; t1 = t2;
    }
    P.chunk_stride_inv = P.n_chunks > 1 ? (int32_t)(((t0 % P.n_chunks) + P.n_chunks) % P.n_chunks) : 0;
  }
  P.probes_per_tile = 1; P.probe_segments = 8;
  P.work_cursor = nullptr;  // the lane's, set by do_render
  P.counters = ctx->counters;
  P.warp_trace = nullptr;
  P.tile_order = nullptr;
  // shared-memory staging plan: BFS prefix of the node array, then the sphere records if they all fit.
  // The warp-queue kernel runs one CTA per SM and gives the staging area whatever its queues leave.
  int64_t budget = std::min<int64_t>(ctx->cfg.smem_budget, ctx->max_smem_optin - 1024) - 128;
  if (resolve_kernel(ctx) == RAY_B200_KERNEL_STREAMQUEUE) {
    const int k = 1;
    const int64_t per_warp = (int64_t)sq_warp_bytes(k, wq_node_capacity(k, p->max_depth));
    int64_t wq_w = ctx->cfg.wq_warps < 1 ? 24 : (ctx->cfg.wq_warps > kWqMaxWarps ? kWqMaxWarps : ctx->cfg.wq_warps);
    while (wq_w > 1 && wq_w * per_warp + 8192 > (int64_t)ctx->max_smem_optin) wq_w--;
    budget = (int64_t)ctx->max_smem_optin - wq_w * per_warp - 512;
    if (budget < 256) { set_error(ctx, "render: stream-queue kernel does not fit shared memory (tree depth %d)", p->max_depth); return 1; }
    ctx->plan_wq_packet = 0;
    ctx->plan_wq_warps = (int32_t)wq_w;
  }
  if (resolve_kernel(ctx) == RAY_B200_KERNEL_WARPQUEUE) {
    // one CTA per SM: as many warps as asked for (<= 32) while their queues leave >= 8 KB for staging;
    // deep trees need bigger node stacks, so they get fewer warps
    const int k = ctx->cfg.wq_k == 1 ? 1 : 2;
    // Packet steps pay off when item-mode node fetches are expensive (part of the tree not staged in shared memory)
    // AND the rays a warp holds are coherent: samples of one pixel (spp > 1) or primary rays of a dense frame.
    // Measured: irreg 64 spp -19 %, irreg 4000^2 1 spp -12 %; rgbbox (fully staged) +1..2 %; 1000^2 1 spp +5 %.
    // The plan is made twice: first assuming packets, to see whether the tree would be fully staged anyway.
    int want_packet = ctx->cfg.wq_packet;
    const bool coherent = spp > 1 || h * w >= ((int64_t)1 << 22);
    int64_t per_warp = 0, wq_w = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
      const bool pk = want_packet != 0;
      per_warp = (int64_t)wq_warp_bytes(k, wq_node_capacity(k, p->max_depth, ctx->cfg.wq_ncap), pk);
      // 32 warps hide latency best while the tree is cache-resident (rgbbox, irreg); the 1 M-sphere tree (64 MB of
      // nodes) runs 8 % faster with 24 warps, i.e. more L1 per warp (profiles/r1_sweep_cta_size.json)
      const int auto_warps = (int64_t)(p->n - 1) * 64 <= ((int64_t)8 << 20) ? 32 : 24;
      wq_w = ctx->cfg.wq_warps < 1 ? auto_warps : (ctx->cfg.wq_warps > kWqMaxWarps ? kWqMaxWarps : ctx->cfg.wq_warps);
      while (wq_w > 1 && wq_w * per_warp + 8192 > (int64_t)ctx->max_smem_optin) wq_w--;
      if (want_packet >= 0) break;
      const int64_t b = (int64_t)ctx->max_smem_optin - wq_w * per_warp - 512 - 128;
      const bool fully_staged = b / 64 >= (int64_t)(p->n - 1);
      want_packet = (!fully_staged && coherent) ? 24 : 0;
      if (want_packet != 0) break;  // the first plan (with packets) stands
    }
    budget = (int64_t)ctx->max_smem_optin - wq_w * per_warp - 512;
    if (budget < 256) { set_error(ctx, "render: warp-queue kernel does not fit shared memory (tree depth %d)", p->max_depth); return 1; }
    ctx->plan_wq_packet = want_packet > 32 ? 32 : want_packet;
    ctx->plan_wq_warps = (int32_t)wq_w;
  }
  int64_t nodes_fit = std::max<int64_t>(0, budget / 64);
  P.smem_nodes = (int32_t)std::min<int64_t>(P.n_inner, nodes_fit);
  const int64_t left = budget - (int64_t)P.smem_nodes * 64;
  P.smem_spheres = (P.smem_nodes == P.n_inner && left >= (int64_t)P.n_leaves * 16) ? P.n_leaves : 0;
  return 0;
}

int resolve_kernel(const futhark_context *ctx) {
  int k = ctx->cfg.kernel;
  if (k == RAY_B200_KERNEL_AUTO) k = RAY_B200_KERNEL_WARPQUEUE;  // fastest measured variant on every BASELINE config
  return k;
}

void free_wavefront(futhark_context *ctx) {
  WavefrontBuffers &b = ctx->wf;
  for (int q = 0; q < 2; q++) {
    if (b.ray_o[q]) cudaFree(b.ray_o[q]);
    if (b.ray_d[q]) cudaFree(b.ray_d[q]);
    if (b.light[q]) cudaFree(b.light[q]);
  }
  if (b.qlen) cudaFree(b.qlen);
  if (b.accum) cudaFree(b.accum);
  memset(&b, 0, sizeof b);
}

// Ray queues: 2 x 48 B per local pixel (+16 B accumulator), resident for the life of the context.
int ensure_wavefront(futhark_context *ctx, int64_t items) {
  WavefrontBuffers &b = ctx->wf;
  b.tail_from = ctx->cfg.tail_from;
  if (b.capacity >= items) return 0;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  free_wavefront(ctx);
  b.tail_from = ctx->cfg.tail_from;
  const size_t bytes = (size_t)items * sizeof(float4);
  for (int q = 0; q < 2; q++) {
    CUDA_TRY(ctx, cudaMalloc(&b.ray_o[q], bytes));
    CUDA_TRY(ctx, cudaMalloc(&b.ray_d[q], bytes));
    CUDA_TRY(ctx, cudaMalloc(&b.light[q], bytes));
  }
  CUDA_TRY(ctx, cudaMalloc(&b.accum, bytes));
  CUDA_TRY(ctx, cudaMalloc(&b.qlen, 2 * (kMaxDepth + 2) * sizeof(int32_t)));
  b.cursor = b.qlen + (kMaxDepth + 2);
  b.capacity = items;
  return 0;
}

// Enqueues one frame on lane `lane_id` (0 = the context's stream).  `timed`: bracket it with the context's timing events.
int do_render(futhark_context *ctx, RenderParams &P, int lane_id = 0, bool timed = true) {
  futhark_context::Lane &L = ctx->lanes[lane_id];
  if (lane_id == 0) L.stream = ctx->stream;
  LaunchConfig lc;
  lc.kernel = resolve_kernel(ctx);
  lc.blocks_per_sm = ctx->cfg.blocks_per_sm;
  lc.sm_count = ctx->sm_count;
  lc.smem_budget = ctx->cfg.smem_budget;
  lc.refill_min = ctx->cfg.refill_min;
  lc.tail_from = ctx->cfg.tail_from;
  lc.wq_warps = ctx->plan_wq_warps > 0 ? ctx->plan_wq_warps : 1;
  lc.wq_k = ctx->cfg.wq_k == 1 ? 1 : 2;
  lc.wq_refill = ctx->cfg.wq_refill < 1 ? 1 : (ctx->cfg.wq_refill > 32 ? 32 : ctx->cfg.wq_refill);
  lc.wq_packet = ctx->plan_wq_packet;
  lc.wq_ncap = ctx->cfg.wq_ncap;
  if (lc.kernel == RAY_B200_KERNEL_WAVEFRONT && ensure_wavefront(ctx, P.local_tiles * kTilePixels)) return 1;
  P.sample_buf = nullptr;
  if ((lc.kernel == RAY_B200_KERNEL_WARPQUEUE || lc.kernel == RAY_B200_KERNEL_STREAMQUEUE) && P.spp > 1 && P.spp <= 65535 && ctx->cfg.wq_spread) {
    // samples of a pixel are spread over a warp's slots; finished colours wait here for the in-order sum
    // (0.6 MB per sample on a B200 with 32 warps).  Above a memory budget, or when the allocation fails, the frame falls
    // back to the pixel-bound variant of the same kernel (kSpread = false), which handles any sample count.
    const size_t need = (size_t)lc.sm_count * lc.wq_warps * kWqRing * (size_t)P.spp * sizeof(float4);
    constexpr size_t kSpreadBudget = (size_t)1 << 30;
    bool have = need <= L.sample_buf_bytes;
    if (!have && need <= kSpreadBudget) {
      CUDA_TRY(ctx, cudaStreamSynchronize(L.stream));
      if (L.sample_buf) CUDA_TRY(ctx, cudaFree(L.sample_buf));
      L.sample_buf = nullptr; L.sample_buf_bytes = 0;
      if (cudaMalloc(&L.sample_buf, need) == cudaSuccess) { L.sample_buf_bytes = need; have = true; }
      else { L.sample_buf = nullptr; cudaGetLastError(); }
    }
    P.sample_buf = have ? L.sample_buf : nullptr;
  }
  // heavy-first claim order (warp-queue kernel): worth its probe pass when a frame is long enough to have a tail to lose
  // — more than one sample per pixel — and pointless when every tile is claimed in the first wave anyway
  const int hf = ctx->cfg.heavy_first;
  const bool heavy_first = lc.kernel == RAY_B200_KERNEL_WARPQUEUE && P.local_tiles > 1 && P.local_tiles < (1ll << 26) &&
                           (hf > 0 || (hf < 0 && P.spp > 1));
  if (heavy_first) {
    P.probes_per_tile = hf >= 4 ? 4 : (hf >= 2 ? 2 : 1);
    P.probe_segments = std::min(std::max(ctx->cfg.probe_segments, 1), kMaxDepth);
    const size_t n = (size_t)P.local_tiles, tmp = tile_order_sort_bytes(P.local_tiles);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t need = 4 * up(4 * n) + up(tmp);
    if (need > L.tile_order_bytes) {
      CUDA_TRY(ctx, cudaStreamSynchronize(L.stream));
      if (L.tile_order_block) CUDA_TRY(ctx, cudaFree(L.tile_order_block));
      L.tile_order_block = nullptr; L.tile_order_bytes = 0;
      CUDA_TRY(ctx, cudaMalloc(&L.tile_order_block, need));
      L.tile_order_bytes = need;
    }
    TileOrderBuffers tb;
    unsigned char *q = L.tile_order_block;
    tb.keys = reinterpret_cast<uint32_t *>(q); q += up(4 * n);
    tb.keys_sorted = reinterpret_cast<uint32_t *>(q); q += up(4 * n);
    tb.ids = reinterpret_cast<int32_t *>(q); q += up(4 * n);
    tb.order = reinterpret_cast<int32_t *>(q); q += up(4 * n);
    tb.sort_tmp = q; tb.sort_tmp_bytes = tmp;
    P.tile_order = tb.order;
    L.tile_order_plan = tb;
  }
  P.work_cursor = L.work_cursor;
  if (ctx->warp_trace && lc.kernel == RAY_B200_KERNEL_WARPQUEUE && lane_id == 0) {
    const size_t n = 1 + (size_t)lc.sm_count * kWqMaxWarps;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->warp_trace, 0, n * sizeof(unsigned long long), L.stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->warp_trace, 0xff, sizeof(unsigned long long), L.stream));
    P.warp_trace = ctx->warp_trace;
    ctx->trace_warps = lc.wq_warps;
  }
  if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_start, L.stream));
  if (lc.kernel == RAY_B200_KERNEL_PERSISTENT || lc.kernel == RAY_B200_KERNEL_WARPQUEUE || lc.kernel == RAY_B200_KERNEL_STREAMQUEUE)
    CUDA_TRY(ctx, cudaMemsetAsync(L.work_cursor, 0, sizeof(int32_t), L.stream));
  if (heavy_first) launch_tile_order(P, L.tile_order_plan, L.stream, &ctx->launches);
  launch_render(P, lc, &ctx->wf, L.stream, &ctx->launches);
  CUDA_TRY(ctx, cudaGetLastError());
  if (timed) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_stop, L.stream));
    ctx->have_timing = true;
  }
  ctx->renders++;
  return 0;
}

// Device memory goes back to the stream-ordered pool (ordered after any render still using it);
// the page-locked upload buffer goes to the context's cache together with the event that guards it.
void free_prepared_device(futhark_context *ctx, futhark_opaque_prepared_scene *p) {
  cudaSetDevice(ctx->cfg.device);
  if (p->dev.block) cudaFreeAsync(p->dev.block, ctx->stream);
  if (p->pinned) {
    if (ctx->pinned_cache.size() < 4) ctx->pinned_cache.push_back({p->pinned, p->pinned_bytes, p->pinned_event});
    else { cudaEventSynchronize(p->pinned_event); cudaEventDestroy(p->pinned_event); cudaFreeHost(p->pinned); }
  }
  p->dev = DeviceBvh();
  p->pinned = nullptr;
  p->pinned_event = nullptr;
}

// A page-locked staging buffer of at least `bytes`, reusing a cached one when possible (main.c frees and
// re-prepares the same scene every run).
int acquire_pinned(futhark_context *ctx, futhark_opaque_prepared_scene *p, size_t bytes) {
  if (p->pinned && p->pinned_bytes >= bytes) {
    CUDA_TRY(ctx, cudaEventSynchronize(p->pinned_event));
    return 0;
  }
  if (p->pinned) {
    CUDA_TRY(ctx, cudaEventSynchronize(p->pinned_event));
    CUDA_TRY(ctx, cudaFreeHost(p->pinned));
    p->pinned = nullptr;
  }
  for (size_t k = 0; k < ctx->pinned_cache.size(); k++) {
    auto &b = ctx->pinned_cache[k];
    if (b.bytes >= bytes && (b.bytes <= 2 * bytes + 4096 || b.bytes <= ((size_t)1 << 20))) {
      CUDA_TRY(ctx, cudaEventSynchronize(b.last_use));  // the copy that last read this block has finished
      if (p->pinned_event) cudaEventDestroy(p->pinned_event);
      p->pinned = b.ptr; p->pinned_bytes = b.bytes; p->pinned_event = b.last_use;
      ctx->pinned_cache.erase(ctx->pinned_cache.begin() + (long)k);
      return 0;
    }
  }
  CUDA_TRY(ctx, cudaMallocHost(&p->pinned, bytes + 64));
  p->pinned_bytes = bytes + 64;
  if (!p->pinned_event) CUDA_TRY(ctx, cudaEventCreateWithFlags(&p->pinned_event, cudaEventDisableTiming));
  return 0;
}

// prepare_scene, device path (default): H2D of the sphere records from page-locked memory, then the whole
// LBVH build + packing as kernels on the context's stream (bvh_build.cu).
double now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int prepare_on_device(futhark_context *ctx, futhark_opaque_prepared_scene *p) {
  const bool timing = ctx->cfg.debugging != 0;
  const double t0 = timing ? now_us() : 0.0;
  const size_t n = p->host.spheres.size();
  const size_t sph_bytes = n * sizeof(SphereRec);
  if (acquire_pinned(ctx, p, sph_bytes)) return 1;
  memcpy(p->pinned, p->host.spheres.data(), sph_bytes);
  float *d_spheres = nullptr;
  CUDA_TRY(ctx, cudaMallocAsync(&d_spheres, sph_bytes, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_spheres, p->pinned, sph_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(p->pinned_event, ctx->stream));
  const double t1 = timing ? now_us() : 0.0;
  p->refit_sweeps = (int32_t)log2f((float)(int64_t)n) + 2;  // bvh.fut:47, host libm as in the reference's C backend
  if (p->dev.block) { CUDA_TRY(ctx, cudaFreeAsync(p->dev.block, ctx->stream)); p->dev = DeviceBvh(); }
  CUDA_TRY(ctx, build_bvh_device(d_spheres, (int64_t)n, p->refit_sweeps, p->dev, ctx->d_build_result, ctx->stream, &ctx->launches));
  CUDA_TRY(ctx, cudaFreeAsync(d_spheres, ctx->stream));
  // the host needs the tree depth (stack sizing) and the root box (kernel parameter) before the first render
  const double t2 = timing ? now_us() : 0.0;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_build_result, ctx->d_build_result, sizeof(BvhBuildResult), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (timing)
    fprintf(ctx->log ? ctx->log : stderr, "[ray_b200] prepare_scene n=%zu: stage+H2D enqueue %.0f us, build enqueue %.0f us, wait %.0f us\n", n,
            t1 - t0, t2 - t1, now_us() - t2);
  memcpy(p->root_box, ctx->h_build_result->root_box, sizeof p->root_box);
  p->max_depth = ctx->h_build_result->max_depth;
  p->stale_nodes = ctx->h_build_result->stale_nodes;
  p->n = (int64_t)n;
  p->host_built = false;
  return 0;
}

// prepare_scene, host path (RAY_HOST_BUILD=1 / tuning "host_build"): scene_host.cpp builds and packs, one H2D copy.
// Kept as an independent implementation the device path is tested against.
int prepare_on_host(futhark_context *ctx, futhark_opaque_prepared_scene *p) {
  Lbvh tree;
  std::string err;
  if (!build_lbvh(p->host, tree, &err)) { set_error(ctx, "%s", err.c_str()); return 1; }
  PackedBvh pk;
  pack_bvh(p->host, tree, pk);
  const int64_t n = tree.n;
  const size_t total = device_bvh_bytes(n);
  if (acquire_pinned(ctx, p, total)) return 1;
  DeviceBvh hostside;  // the same carving, applied to the page-locked buffer
  carve_device_bvh(p->pinned, n, hostside);
  const size_t ni = (size_t)(n - 1);
  memcpy(hostside.nodes, pk.nodes.data(), ni * 64);
  memcpy(hostside.nodes_soa, pk.nodes_soa.data(), ni * 64);
  memcpy(hostside.geom, pk.geom.data(), (size_t)n * 16);
  memcpy(hostside.colour, pk.colour.data(), (size_t)n * 16);
  memcpy(hostside.morton, tree.morton.data(), (size_t)n * 4);
  memcpy(hostside.perm, tree.perm.data(), (size_t)n * 4);
  memcpy(hostside.left, tree.left.data(), ni * 4);
  memcpy(hostside.right, tree.right.data(), ni * 4);
  memcpy(hostside.parent, tree.parent.data(), ni * 4);
  memcpy(hostside.boxes, tree.boxes.data(), ni * 24);
  if (p->dev.block) { CUDA_TRY(ctx, cudaFreeAsync(p->dev.block, ctx->stream)); p->dev = DeviceBvh(); }
  unsigned char *blk = nullptr;
  CUDA_TRY(ctx, cudaMallocAsync(&blk, total, ctx->stream));
  carve_device_bvh(blk, n, p->dev);
  CUDA_TRY(ctx, cudaMemcpyAsync(blk, p->pinned, total, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(p->pinned_event, ctx->stream));
  memcpy(p->root_box, pk.root_box, sizeof p->root_box);
  p->max_depth = tree.max_depth; p->stale_nodes = tree.stale_nodes; p->refit_sweeps = tree.refit_sweeps;
  p->n = n;
  p->host_built = true;
  return 0;
}

int prepare_any(futhark_context *ctx, futhark_opaque_prepared_scene *p) {
  const int64_t n = (int64_t)p->host.spheres.size();
  if (n < 2) { set_error(ctx, "prepare_scene: a scene needs at least 2 spheres (the reference indexes I[0], bvh.fut:65)"); return 1; }
  if (n > (int64_t)1 << 26) { set_error(ctx, "prepare_scene: too many spheres (this build packs leaf indices into 26 bits)"); return 1; }
  return ctx->cfg.host_build ? prepare_on_host(ctx, p) : prepare_on_device(ctx, p);
}

}  // namespace

// ------------------------------------------------------------------------------------------ config / context
extern "C" {

struct futhark_context_config *futhark_context_config_new(void) {
  futhark_context_config *cfg = new (std::nothrow) futhark_context_config;
  return cfg;
}
void futhark_context_config_free(struct futhark_context_config *cfg) { delete cfg; }
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag) { cfg->debugging = flag; }
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag) { cfg->profiling = flag; }
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag) { cfg->logging = flag; }
void futhark_context_config_set_cache_file(struct futhark_context_config *cfg, const char *f) { cfg->cache_file = f ? f : ""; }
void futhark_context_config_set_device(struct futhark_context_config *cfg, const char *s) {
  if (!s) return;
  if (*s == '#') s++;
  cfg->device = atoi(s);
}
int futhark_get_tuning_param_count(void) { return kNumTuning; }
const char *futhark_get_tuning_param_name(int i) { return (i >= 0 && i < kNumTuning) ? kTuningNames[i] : nullptr; }
const char *futhark_get_tuning_param_class(int i) { return (i >= 0 && i < kNumTuning) ? "ray_b200" : nullptr; }
int futhark_context_config_set_tuning_param(struct futhark_context_config *cfg, const char *name, size_t v) {
  if (!strcmp(name, "kernel")) cfg->kernel = (int32_t)v;
  else if (!strcmp(name, "spp")) cfg->spp = (int32_t)v;
  else if (!strcmp(name, "blocks_per_sm")) cfg->blocks_per_sm = (int32_t)v;
  else if (!strcmp(name, "smem_budget")) cfg->smem_budget = (int32_t)v;
  else if (!strcmp(name, "refill_min")) cfg->refill_min = (int32_t)v;
  else if (!strcmp(name, "tail_from")) cfg->tail_from = (int32_t)v;
  else if (!strcmp(name, "wq_warps")) cfg->wq_warps = (int32_t)v;
  else if (!strcmp(name, "wq_k")) cfg->wq_k = (int32_t)v;
  else if (!strcmp(name, "wq_spread")) cfg->wq_spread = (int32_t)v;
  else if (!strcmp(name, "wq_refill")) cfg->wq_refill = (int32_t)v;
  else if (!strcmp(name, "wq_packet")) cfg->wq_packet = (int32_t)v;  // (size_t)-1 = decide per scene
  else if (!strcmp(name, "wq_ncap")) cfg->wq_ncap = (int32_t)v;
  else if (!strcmp(name, "permute")) cfg->permute = (int32_t)v;
  else if (!strcmp(name, "heavy_first")) cfg->heavy_first = (int32_t)v;  // (size_t)-1 = decide per frame
  else if (!strcmp(name, "probe_segments")) cfg->probe_segments = (int32_t)v;
  else if (!strcmp(name, "host_build")) cfg->host_build = (int32_t)v;
  else if (!strcmp(name, "rank")) cfg->rank = (int32_t)v;
  else if (!strcmp(name, "world")) cfg->world = (int32_t)v;
  else if (!strcmp(name, "gpus")) cfg->gpus = (int32_t)v;
  else return 1;
  return 0;
}

struct futhark_context *futhark_context_new(struct futhark_context_config *cfg) {
  futhark_context *ctx = new (std::nothrow) futhark_context;
  if (!ctx) return nullptr;
  if (cfg) ctx->cfg = *cfg;
  // environment overrides: the only extension channel an unmodified futhark/main.c has
  const bool helper = cfg && cfg->gpus < 0;  // helper context of a single-process multi-GPU context: config is final
  if (!helper) ctx->cfg.device = env_int("RAY_DEVICE", ctx->cfg.device);
  ctx->cfg.spp = env_int("RAY_SPP", ctx->cfg.spp);
  ctx->cfg.kernel = parse_kernel(getenv("RAY_KERNEL"), ctx->cfg.kernel);
  if (!helper) ctx->cfg.rank = env_int("RAY_RANK", ctx->cfg.rank);
  if (!helper) ctx->cfg.world = env_int("RAY_WORLD", ctx->cfg.world);
  if (!(cfg && cfg->gpus < 0)) ctx->cfg.gpus = env_int("RAY_GPUS", ctx->cfg.gpus);
  ctx->cfg.blocks_per_sm = env_int("RAY_BLOCKS_PER_SM", ctx->cfg.blocks_per_sm);
  ctx->cfg.smem_budget = env_int("RAY_SMEM_BUDGET", ctx->cfg.smem_budget);
  ctx->cfg.refill_min = env_int("RAY_REFILL_MIN", ctx->cfg.refill_min);
  ctx->cfg.tail_from = env_int("RAY_TAIL_FROM", ctx->cfg.tail_from);
  ctx->cfg.wq_warps = env_int("RAY_WQ_WARPS", ctx->cfg.wq_warps);
  ctx->cfg.wq_k = env_int("RAY_WQ_K", ctx->cfg.wq_k);
  ctx->cfg.wq_spread = env_int("RAY_WQ_SPREAD", ctx->cfg.wq_spread);
  ctx->cfg.wq_packet = env_int("RAY_WQ_PACKET", ctx->cfg.wq_packet);
  ctx->cfg.wq_refill = env_int("RAY_WQ_REFILL", ctx->cfg.wq_refill);
  ctx->cfg.wq_ncap = env_int("RAY_WQ_NCAP", ctx->cfg.wq_ncap);
  ctx->cfg.permute = env_int("RAY_PERMUTE", ctx->cfg.permute);
  ctx->cfg.heavy_first = env_int("RAY_HEAVY_FIRST", ctx->cfg.heavy_first);
  ctx->cfg.probe_segments = env_int("RAY_PROBE_SEGMENTS", ctx->cfg.probe_segments);
  ctx->cfg.host_build = env_int("RAY_HOST_BUILD", ctx->cfg.host_build);
  ctx->cfg.debugging = env_int("RAY_DEBUG", ctx->cfg.debugging);
  memset(&ctx->wf, 0, sizeof ctx->wf);

  auto fail = [&](const char *what, cudaError_t e) {
    set_error(ctx, "futhark_context_new: %s: %s (this library has no CPU fallback)", what, cudaGetErrorString(e));
    return ctx;
  };
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail("no CUDA device", e == cudaSuccess ? cudaErrorNoDevice : e);
  if (ctx->cfg.device < 0 || ctx->cfg.device >= ndev) return fail("device index out of range", cudaErrorInvalidDevice);
  if ((e = cudaSetDevice(ctx->cfg.device)) != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, ctx->cfg.device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  if (prop.major < 10) {
    set_error(ctx, "futhark_context_new: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
              ctx->cfg.device, prop.major, prop.minor);
    return ctx;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaEventCreate(&ctx->ev_start)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&ctx->ev_stop)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaMalloc(&ctx->lanes[0].work_cursor, 64)) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->counters, 4 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->d_build_result, sizeof(BvhBuildResult))) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMallocHost(&ctx->h_build_result, sizeof(BvhBuildResult))) != cudaSuccess) return fail("cudaMallocHost", e);
  if ((e = configure_kernels(ctx->max_smem_optin)) != cudaSuccess) return fail("cudaFuncSetAttribute", e);
  // keep freed frames in the stream-ordered pool: futhark/main.c frees and re-allocates the image every run
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, ctx->cfg.device) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  ctx->ok = true;
  {
    // One-time costs belong to context creation, not to the first timed prepare_scene (main.c:88-100 averages over its
    // runs): build a throw-away two-sphere scene so that the build kernels, CUB's sort kernels, the stream-ordered pool
    // and a page-locked staging block are loaded / allocated now (the render kernels were loaded by configure_kernels).
    futhark_opaque_prepared_scene warm;
    warm.host.spheres = {SphereRec{0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 1.0f, 1.0f}, SphereRec{3.0f, 0.0f, 0.0f, 1.0f, 1.0f, 1.0f, 1.0f}};
    if (prepare_on_device(ctx, &warm) != 0) {
      ctx->ok = false;  // the error message is already set
      return ctx;
    }
    free_prepared_device(ctx, &warm);
    // pre-grow the stream-ordered pool (it keeps what is freed: release threshold = max) and seed the page-locked cache
    void *grow = nullptr;
    if (cudaMallocAsync(&grow, (size_t)64 << 20, ctx->stream) == cudaSuccess) cudaFreeAsync(grow, ctx->stream);
    futhark_context::PinnedBlock seed{nullptr, (size_t)1 << 20, nullptr};
    if (cudaMallocHost(&seed.ptr, seed.bytes) == cudaSuccess && cudaEventCreateWithFlags(&seed.last_use, cudaEventDisableTiming) == cudaSuccess) {
      cudaEventRecord(seed.last_use, ctx->stream);
      ctx->pinned_cache.push_back(seed);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaGetLastError();
    ctx->launches = 0;
  }
  // single-process multi-GPU: helper contexts on devices device+1 .. device+gpus-1 (peer access enabled both ways)
  if (ctx->cfg.gpus > 1 && !(cfg && cfg->gpus < 0)) {
    if (ctx->cfg.device + ctx->cfg.gpus > ndev) {
      ctx->ok = false;
      set_error(ctx, "futhark_context_new: RAY_GPUS=%d needs devices %d..%d but only %d are visible", ctx->cfg.gpus, ctx->cfg.device,
                ctx->cfg.device + ctx->cfg.gpus - 1, ndev);
      return ctx;
    }
    for (int r = 1; r < ctx->cfg.gpus; r++) {
      futhark_context_config pc = ctx->cfg;
      pc.device = ctx->cfg.device + r;
      pc.rank = r; pc.world = ctx->cfg.gpus;
      pc.gpus = -1;  // marks a helper: no recursion, no environment override of the device
      futhark_context *peer = futhark_context_new(&pc);
      if (!peer || !peer->ok) {
        char *pe = peer ? futhark_context_get_error(peer) : nullptr;
        ctx->ok = false;
        set_error(ctx, "futhark_context_new: helper context on device %d failed: %s", pc.device, pe ? pe : "?");
        free(pe);
        if (peer) futhark_context_free(peer);
        return ctx;
      }
      peer->is_peer = true;
      cudaSetDevice(pc.device);
      cudaEventCreateWithFlags(&peer->peer_done, cudaEventDisableTiming);
      cudaDeviceEnablePeerAccess(ctx->cfg.device, 0);
      cudaSetDevice(ctx->cfg.device);
      cudaDeviceEnablePeerAccess(pc.device, 0);
      cudaGetLastError();  // "already enabled" is fine
      ctx->peers.push_back(peer);
    }
    ctx->cfg.rank = 0; ctx->cfg.world = ctx->cfg.gpus;
    cudaSetDevice(ctx->cfg.device);
  }
  return ctx;
}

void futhark_context_free(struct futhark_context *ctx) {
  if (!ctx) return;
  for (futhark_context *peer : ctx->peers) futhark_context_free(peer);
  ctx->peers.clear();
  if (ctx->peer_tiles) { cudaSetDevice(ctx->cfg.device); cudaFree(ctx->peer_tiles); }
  if (ctx->peer_done) cudaEventDestroy(ctx->peer_done);
  if (ctx->gathered) { cudaSetDevice(ctx->cfg.device); cudaFree(ctx->gathered); }
  if (ctx->ev_gathered) cudaEventDestroy(ctx->ev_gathered);
  if (ctx->ok) {
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
  }
  for (auto &b : ctx->pinned_cache) { cudaEventDestroy(b.last_use); cudaFreeHost(b.ptr); }
  if (ctx->ok) free_wavefront(ctx);
  for (auto &L : ctx->lanes) {
    if (L.sample_buf) cudaFree(L.sample_buf);
    if (L.tile_order_block) cudaFree(L.tile_order_block);
    if (L.work_cursor) cudaFree(L.work_cursor);
  }
  if (ctx->lanes[1].stream) cudaStreamDestroy(ctx->lanes[1].stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (auto &e : ctx->offset_tables) cudaFree(e.dev);
  if (ctx->counters) cudaFree(ctx->counters);
  if (ctx->warp_trace) cudaFree(ctx->warp_trace);
  if (ctx->d_build_result) cudaFree(ctx->d_build_result);
  if (ctx->h_build_result) cudaFreeHost(ctx->h_build_result);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  free(ctx->error);
  delete ctx;
}

int futhark_context_sync(struct futhark_context *ctx) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  for (futhark_context *peer : ctx->peers) {
    if (futhark_context_sync(peer)) { char *pe = futhark_context_get_error(peer); set_error(ctx, "helper device: %s", pe ? pe : "?"); free(pe); return 1; }
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

char *futhark_context_get_error(struct futhark_context *ctx) {
  if (!ctx) return nullptr;
  char *e = ctx->error;
  ctx->error = nullptr;
  return e;
}

char *futhark_context_report(struct futhark_context *ctx) {
  if (!ctx) return nullptr;
  std::lock_guard<std::mutex> g(ctx->mu);
  char buf[512];
  float ms = 0.0f;
  if (ctx->ok && ctx->have_timing) {
    cudaEventSynchronize(ctx->ev_stop);
    cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop);
  }
  snprintf(buf, sizeof buf,
           "ray_b200 %s\ndevice: %d (%d SMs)\nrenders: %lld\nkernel launches: %lld\nlast render: %.3f ms (device)\n",
           ray_b200_version(), ctx->cfg.device, ctx->sm_count, (long long)ctx->renders, (long long)ctx->launches, ms);
  return strdup(buf);
}
void futhark_context_set_logging_file(struct futhark_context *ctx, FILE *f) { if (ctx) ctx->log = f; }
void futhark_context_pause_profiling(struct futhark_context *ctx) { if (ctx) ctx->profiling_paused = true; }
void futhark_context_unpause_profiling(struct futhark_context *ctx) { if (ctx) ctx->profiling_paused = false; }
int futhark_context_clear_caches(struct futhark_context *ctx) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  cudaMemPool_t pool;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // lane 1 always joins the context's stream before a call returns
  // grow-only scratch: finished-sample buffers, claim-order tables, wavefront ray queues, page-locked upload buffers
  for (auto &L : ctx->lanes) {
    if (L.sample_buf) cudaFree(L.sample_buf);
    if (L.tile_order_block) cudaFree(L.tile_order_block);
    L.sample_buf = nullptr; L.sample_buf_bytes = 0;
    L.tile_order_block = nullptr; L.tile_order_bytes = 0;
  }
  free_wavefront(ctx);
  for (auto &b : ctx->pinned_cache) { cudaEventDestroy(b.last_use); cudaFreeHost(b.ptr); }
  ctx->pinned_cache.clear();
  if (cudaDeviceGetDefaultMemPool(&pool, ctx->cfg.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  return 0;
}

// ------------------------------------------------------------------------------------------ arrays
struct futhark_i32_2d *futhark_new_i32_2d(struct futhark_context *ctx, const int32_t *data, int64_t d0, int64_t d1) {
  if (bad_ctx(ctx) || d0 < 0 || d1 < 0) return nullptr;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  futhark_i32_2d *a = new futhark_i32_2d;
  a->shape[0] = d0; a->shape[1] = d1;
  const size_t bytes = (size_t)d0 * d1 * sizeof(int32_t);
  if (cudaMallocAsync(&a->dev, bytes ? bytes : 4, ctx->stream) != cudaSuccess) { set_error(ctx, "futhark_new_i32_2d: out of device memory"); delete a; return nullptr; }
  if (bytes && cudaMemcpyAsync(a->dev, data, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { set_error(ctx, "futhark_new_i32_2d: copy failed"); cudaFreeAsync(a->dev, ctx->stream); delete a; return nullptr; }
  cudaStreamSynchronize(ctx->stream);
  return a;
}
struct futhark_i32_2d *futhark_new_raw_i32_2d(struct futhark_context *ctx, void *device_ptr, int64_t d0, int64_t d1) {
  if (bad_ctx(ctx)) return nullptr;
  futhark_i32_2d *a = new futhark_i32_2d;
  a->dev = (int32_t *)device_ptr; a->shape[0] = d0; a->shape[1] = d1; a->owned = false;
  return a;
}
int futhark_free_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr) {
  if (bad_ctx(ctx)) return 1;
  if (!arr) return 0;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  if (arr->owned && arr->dev) CUDA_TRY(ctx, cudaFreeAsync(arr->dev, ctx->stream));
  delete arr;
  return 0;
}
int futhark_values_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr, int32_t *data) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  if (!arr || !data) { set_error(ctx, "futhark_values_i32_2d: null argument"); return 1; }
  const size_t bytes = (size_t)arr->shape[0] * arr->shape[1] * sizeof(int32_t);
  CUDA_TRY(ctx, cudaMemcpyAsync(data, arr->dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // main.c:130-133 reads `data` without syncing
  return 0;
}
void *futhark_values_raw_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr) { (void)ctx; return arr ? arr->dev : nullptr; }
const int64_t *futhark_shape_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr) { (void)ctx; return arr ? arr->shape : nullptr; }

// ------------------------------------------------------------------------------------------ opaque scene
int futhark_free_opaque_scene(struct futhark_context *ctx, struct futhark_opaque_scene *obj) { (void)ctx; delete obj; return 0; }

namespace {
constexpr uint32_t kSceneMagic = 0x53423252u;     // "R2BS"
constexpr uint32_t kPreparedMagic = 0x50423252u;  // "R2BP"
size_t scene_blob_size(const HostScene &s) { return 16 + 7 * sizeof(float) + s.spheres.size() * sizeof(SphereRec); }
void scene_to_blob(const HostScene &s, unsigned char *p, uint32_t magic) {
  const uint64_t n = s.spheres.size();
  memcpy(p, &magic, 4); uint32_t ver = 1; memcpy(p + 4, &ver, 4); memcpy(p + 8, &n, 8);
  float cam[7] = {s.look_from[0], s.look_from[1], s.look_from[2], s.look_at[0], s.look_at[1], s.look_at[2], s.fov};
  memcpy(p + 16, cam, sizeof cam);
  memcpy(p + 16 + sizeof cam, s.spheres.data(), n * sizeof(SphereRec));
}
bool scene_from_blob(HostScene &s, const unsigned char *p, uint32_t magic) {
  uint32_t m, ver; uint64_t n;
  memcpy(&m, p, 4); memcpy(&ver, p + 4, 4); memcpy(&n, p + 8, 8);
  if (m != magic || ver != 1) return false;
  float cam[7];
  memcpy(cam, p + 16, sizeof cam);
  memcpy(s.look_from, cam, 12); memcpy(s.look_at, cam + 3, 12); s.fov = cam[6];
  s.spheres.resize(n);
  memcpy(s.spheres.data(), p + 16 + sizeof cam, n * sizeof(SphereRec));
  return true;
}
}  // namespace

int futhark_store_opaque_scene(struct futhark_context *ctx, const struct futhark_opaque_scene *obj, void **p, size_t *n) {
  if (!ctx || !obj || !n) return 1;
  const size_t sz = scene_blob_size(obj->host);
  *n = sz;
  if (p) {
    if (!*p) *p = malloc(sz);
    if (!*p) { set_error(ctx, "store_opaque_scene: out of memory"); return 1; }
    scene_to_blob(obj->host, (unsigned char *)*p, kSceneMagic);
  }
  return 0;
}
struct futhark_opaque_scene *futhark_restore_opaque_scene(struct futhark_context *ctx, const void *p) {
  if (!ctx || !p) return nullptr;
  futhark_opaque_scene *s = new futhark_opaque_scene;
  if (!scene_from_blob(s->host, (const unsigned char *)p, kSceneMagic)) { set_error(ctx, "restore_opaque_scene: bad blob"); delete s; return nullptr; }
  return s;
}

int futhark_free_opaque_prepared_scene(struct futhark_context *ctx, struct futhark_opaque_prepared_scene *obj) {
  if (!obj) return 0;
  if (ctx && ctx->ok) {
    std::lock_guard<std::mutex> g(ctx->mu);
    for (size_t k = 0; k < obj->peer_prepared.size() && k < ctx->peers.size(); k++)
      futhark_free_opaque_prepared_scene(ctx->peers[k], obj->peer_prepared[k]);
    cudaSetDevice(ctx->cfg.device);
    free_prepared_device(ctx, obj);  // stream-ordered: a render still using it finishes first
  }
  delete obj;
  return 0;
}
// A stored prepared scene is the scene plus the (h, w) it was prepared for; restoring re-runs prepare_scene.
int futhark_store_opaque_prepared_scene(struct futhark_context *ctx, const struct futhark_opaque_prepared_scene *obj, void **p, size_t *n) {
  if (!ctx || !obj || !n) return 1;
  const size_t sz = scene_blob_size(obj->host) + 16;
  *n = sz;
  if (p) {
    if (!*p) *p = malloc(sz);
    if (!*p) { set_error(ctx, "store_opaque_prepared_scene: out of memory"); return 1; }
    scene_to_blob(obj->host, (unsigned char *)*p, kPreparedMagic);
    memcpy((unsigned char *)*p + sz - 16, &obj->h, 8);
    memcpy((unsigned char *)*p + sz - 8, &obj->w, 8);
  }
  return 0;
}
struct futhark_opaque_prepared_scene *futhark_restore_opaque_prepared_scene(struct futhark_context *ctx, const void *p) {
  if (bad_ctx(ctx) || !p) return nullptr;
  futhark_opaque_scene tmp;
  if (!scene_from_blob(tmp.host, (const unsigned char *)p, kPreparedMagic)) { set_error(ctx, "restore_opaque_prepared_scene: bad blob"); return nullptr; }
  const size_t sz = scene_blob_size(tmp.host) + 16;
  int64_t h, w;
  memcpy(&h, (const unsigned char *)p + sz - 16, 8);
  memcpy(&w, (const unsigned char *)p + sz - 8, 8);
  futhark_opaque_prepared_scene *out = nullptr;
  if (futhark_entry_prepare_scene(ctx, &out, h, w, &tmp) != 0) return nullptr;
  return out;
}

// ------------------------------------------------------------------------------------------ entry points
int futhark_entry_rgbbox(struct futhark_context *ctx, struct futhark_opaque_scene **out0) {
  if (bad_ctx(ctx) || !out0) return 1;
  futhark_opaque_scene *s = new futhark_opaque_scene;
  make_rgbbox(s->host);
  *out0 = s;
  return 0;
}
int futhark_entry_irreg(struct futhark_context *ctx, struct futhark_opaque_scene **out0) {
  if (bad_ctx(ctx) || !out0) return 1;
  futhark_opaque_scene *s = new futhark_opaque_scene;
  make_irreg(s->host);
  *out0 = s;
  return 0;
}

int futhark_entry_prepare_scene(struct futhark_context *ctx, struct futhark_opaque_prepared_scene **out0, const int64_t h,
                                const int64_t w, const struct futhark_opaque_scene *scene) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!out0 || !scene) { set_error(ctx, "prepare_scene: null argument"); return 1; }
  if (h <= 0 || w <= 0) { set_error(ctx, "prepare_scene: bad image size"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  futhark_opaque_prepared_scene *p = new futhark_opaque_prepared_scene;
  p->host = scene->host;
  p->h = h; p->w = w;
  p->cam = make_camera(p->host, h, w);
  if (prepare_any(ctx, p)) { free_prepared_device(ctx, p); delete p; return 1; }
  for (futhark_context *peer : ctx->peers) {  // the scene is replicated: every device builds its own LBVH
    futhark_opaque_prepared_scene *pp = nullptr;
    if (futhark_entry_prepare_scene(peer, &pp, h, w, scene)) {
      char *pe = futhark_context_get_error(peer);
      set_error(ctx, "prepare_scene on helper device %d: %s", peer->cfg.device, pe ? pe : "?");
      free(pe);
      for (size_t k = 0; k < p->peer_prepared.size(); k++) futhark_free_opaque_prepared_scene(ctx->peers[k], p->peer_prepared[k]);
      free_prepared_device(ctx, p); delete p;
      return 1;
    }
    p->peer_prepared.push_back(pp);
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  *out0 = p;
  return 0;
}

int ray_b200_entry_render_spp(struct futhark_context *ctx, struct futhark_i32_2d **out0, int64_t h, int64_t w, int32_t spp,
                              const struct futhark_opaque_prepared_scene *p) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!out0) { set_error(ctx, "render: null output"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  futhark_i32_2d *img = new futhark_i32_2d;
  img->shape[0] = h; img->shape[1] = w;
  const size_t bytes = (size_t)(h > 0 ? h : 0) * (size_t)(w > 0 ? w : 0) * sizeof(int32_t);
  cudaError_t e = cudaMallocAsync(&img->dev, bytes ? bytes : 4, ctx->stream);
  if (e != cudaSuccess) { set_error(ctx, "render: cudaMallocAsync: %s", cudaGetErrorString(e)); delete img; return 1; }
  RenderParams P;
  if (!ctx->peers.empty()) {
    // single-process multi-GPU: every device renders its tiles into a compact buffer, device 0 pulls them over
    // NVLink (peer copies ordered by events) and de-tiles — the same data flow as the one-process-per-GPU path,
    // with cudaMemcpyPeerAsync in place of the NCCL gather
    const int world = (int)ctx->peers.size() + 1;
    const int64_t padded = ray_b200_shard_tiles_padded(h, w, world);
    const size_t shard_bytes = (size_t)padded * kTilePixels * sizeof(int32_t);
    auto fail = [&](const char *what) { set_error(ctx, "render (multi-GPU): %s", what); cudaFreeAsync(img->dev, ctx->stream); delete img; return 1; };
    if (!p) return fail("invalid prepared scene");
    if (p->peer_prepared.size() != ctx->peers.size()) return fail("prepared scene was not prepared by this context");
    if (ctx->gathered_bytes < shard_bytes * world) {
      cudaStreamSynchronize(ctx->stream);
      if (ctx->gathered) cudaFree(ctx->gathered);
      ctx->gathered = nullptr; ctx->gathered_bytes = 0;
      if (cudaMalloc(&ctx->gathered, shard_bytes * world) != cudaSuccess) return fail("out of device memory");
      ctx->gathered_bytes = shard_bytes * world;
    }
    for (int r = 1; r < world; r++) {
      futhark_context *peer = ctx->peers[(size_t)r - 1];
      cudaSetDevice(peer->cfg.device);
      if (peer->peer_tiles_bytes < shard_bytes) {
        cudaStreamSynchronize(peer->stream);
        if (peer->peer_tiles) cudaFree(peer->peer_tiles);
        peer->peer_tiles = nullptr; peer->peer_tiles_bytes = 0;
        if (cudaMalloc(&peer->peer_tiles, shard_bytes) != cudaSuccess) { cudaSetDevice(ctx->cfg.device); return fail("out of device memory on a helper device"); }
        peer->peer_tiles_bytes = shard_bytes;
      }
      // the previous frame's peer copy out of peer_tiles (on device 0's stream) must have finished before this helper
      // renders into it again: entries are asynchronous, a caller may issue two renders without a sync in between
      if (ctx->gather_pending) cudaStreamWaitEvent(peer->stream, ctx->ev_gathered, 0);
      if (ray_b200_render_shard_into(peer, peer->peer_tiles, h, w, spp, p->peer_prepared[(size_t)r - 1])) {
        char *pe = futhark_context_get_error(peer);
        cudaSetDevice(ctx->cfg.device);
        set_error(ctx, "render on helper device %d: %s", peer->cfg.device, pe ? pe : "?");
        free(pe);
        cudaFreeAsync(img->dev, ctx->stream); delete img;
        return 1;
      }
      cudaEventRecord(peer->peer_done, peer->stream);
    }
    cudaSetDevice(ctx->cfg.device);
    // rank 0's own shard straight into slot 0 of the gather buffer
    if (fill_params(ctx, p, h, w, spp, 0, world, ctx->gathered, nullptr, true, P)) { cudaFreeAsync(img->dev, ctx->stream); delete img; return 1; }
    if (P.local_tiles < padded)
      cudaMemsetAsync(ctx->gathered + P.local_tiles * kTilePixels, 0, (size_t)(padded - P.local_tiles) * kTilePixels * 4, ctx->stream);
    if (do_render(ctx, P)) { cudaFreeAsync(img->dev, ctx->stream); delete img; return 1; }
    for (int r = 1; r < world; r++) {
      futhark_context *peer = ctx->peers[(size_t)r - 1];
      cudaStreamWaitEvent(ctx->stream, peer->peer_done, 0);
      cudaMemcpyPeerAsync(ctx->gathered + (size_t)r * padded * kTilePixels, ctx->cfg.device, peer->peer_tiles, peer->cfg.device, shard_bytes, ctx->stream);
    }
    if (!ctx->ev_gathered) cudaEventCreateWithFlags(&ctx->ev_gathered, cudaEventDisableTiming);
    cudaEventRecord(ctx->ev_gathered, ctx->stream);  // every helper's tile buffer has been read
    ctx->gather_pending = true;
    launch_detile(ctx->gathered, img->dev, h, w, world, padded, ctx->stream, &ctx->launches);
    if (cudaGetLastError() != cudaSuccess) return fail("de-tiling launch failed");
    *out0 = img;
    return 0;
  }
  // With a shard configured, the row-major frame only receives this rank's tiles; clear the rest.
  if (ctx->cfg.world > 1) cudaMemsetAsync(img->dev, 0, bytes, ctx->stream);
  if (fill_params(ctx, p, h, w, spp, ctx->cfg.rank, ctx->cfg.world, img->dev, nullptr, false, P) || do_render(ctx, P)) {
    cudaFreeAsync(img->dev, ctx->stream);
    delete img;
    return 1;
  }
  *out0 = img;
  return 0;
}

int futhark_entry_render(struct futhark_context *ctx, struct futhark_i32_2d **out0, const int64_t h, const int64_t w,
                         const struct futhark_opaque_prepared_scene *p) {
  if (bad_ctx(ctx)) return 1;
  return ray_b200_entry_render_spp(ctx, out0, h, w, ctx->cfg.spp, p);
}

// ------------------------------------------------------------------------------------------ extensions
int ray_b200_context_set_stream(struct futhark_context *ctx, void *s) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return 0;
}
int ray_b200_context_set_spp(struct futhark_context *ctx, int32_t spp) {
  if (bad_ctx(ctx)) return 1;
  if (spp < 1) { set_error(ctx, "spp must be >= 1"); return 1; }
  ctx->cfg.spp = spp;
  return 0;
}
int ray_b200_context_set_kernel(struct futhark_context *ctx, int32_t k) {
  if (bad_ctx(ctx)) return 1;
  if (k < RAY_B200_KERNEL_AUTO || k > RAY_B200_KERNEL_STREAMQUEUE) { set_error(ctx, "unknown kernel %d", k); return 1; }
  ctx->cfg.kernel = k;
  return 0;
}
int ray_b200_context_set_shard(struct futhark_context *ctx, int32_t rank, int32_t world) {
  if (bad_ctx(ctx)) return 1;
  if (world < 1 || rank < 0 || rank >= world) { set_error(ctx, "bad shard %d/%d", rank, world); return 1; }
  ctx->cfg.rank = rank; ctx->cfg.world = world;
  return 0;
}
int ray_b200_context_device(struct futhark_context *ctx) { return ctx ? ctx->cfg.device : -1; }
int ray_b200_context_last_render_ms(struct futhark_context *ctx, float *ms) {
  if (bad_ctx(ctx) || !ms) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  if (!ctx->have_timing) { set_error(ctx, "no render has been issued yet"); return 1; }
  CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_stop));
  CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
  return 0;
}
int64_t ray_b200_context_launch_count(struct futhark_context *ctx) { return ctx ? ctx->launches : 0; }

int ray_b200_context_trace_warps(struct futhark_context *ctx, int32_t enable) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);
  if (enable && !ctx->warp_trace) {
    CUDA_TRY(ctx, cudaMalloc(&ctx->warp_trace, (1 + (size_t)ctx->sm_count * kWqMaxWarps) * sizeof(unsigned long long)));
  } else if (!enable && ctx->warp_trace) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaFree(ctx->warp_trace));
    ctx->warp_trace = nullptr;
  }
  return 0;
}
int ray_b200_context_warp_trace(struct futhark_context *ctx, float *exit_us, int64_t capacity, int64_t *count) {
  if (bad_ctx(ctx) || !count) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);
  if (!ctx->warp_trace || ctx->trace_warps == 0) { set_error(ctx, "warp_trace: no traced warp-queue render yet"); return 1; }
  const int64_t n = (int64_t)ctx->sm_count * ctx->trace_warps;
  *count = n;
  if (!exit_us) return 0;
  if (capacity < n) { set_error(ctx, "warp_trace: capacity %lld < %lld", (long long)capacity, (long long)n); return 1; }
  std::vector<unsigned long long> h(1 + (size_t)n);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaMemcpy(h.data(), ctx->warp_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; i++) exit_us[i] = h[1 + i] > h[0] ? (float)((double)(h[1 + i] - h[0]) * 1e-3) : 0.0f;
  return 0;
}

int ray_b200_scene_from_arrays(struct futhark_context *ctx, struct futhark_opaque_scene **out0, const float *spheres, int64_t n,
                               const float *cam7) {
  if (bad_ctx(ctx)) return 1;
  if (!out0 || !spheres || !cam7 || n < 0) { set_error(ctx, "scene_from_arrays: bad argument"); return 1; }
  futhark_opaque_scene *s = new futhark_opaque_scene;
  s->host.spheres.resize((size_t)n);
  static_assert(sizeof(SphereRec) == 7 * sizeof(float), "SphereRec must be 7 packed floats");
  memcpy(s->host.spheres.data(), spheres, (size_t)n * sizeof(SphereRec));
  memcpy(s->host.look_from, cam7, 12); memcpy(s->host.look_at, cam7 + 3, 12); s->host.fov = cam7[6];
  *out0 = s;
  return 0;
}
int ray_b200_scene_random(struct futhark_context *ctx, struct futhark_opaque_scene **out0, int64_t n, uint64_t seed) {
  if (bad_ctx(ctx)) return 1;
  if (!out0 || n < 0) { set_error(ctx, "scene_random: bad argument"); return 1; }
  futhark_opaque_scene *s = new futhark_opaque_scene;
  make_random(s->host, n, seed);
  *out0 = s;
  return 0;
}
int64_t ray_b200_scene_num_spheres(struct futhark_context *ctx, const struct futhark_opaque_scene *s) { (void)ctx; return s ? (int64_t)s->host.spheres.size() : -1; }
int ray_b200_scene_get_arrays(struct futhark_context *ctx, const struct futhark_opaque_scene *s, float *spheres, float *cam7) {
  (void)ctx;
  if (!s) return 1;
  if (spheres) memcpy(spheres, s->host.spheres.data(), s->host.spheres.size() * sizeof(SphereRec));
  if (cam7) { memcpy(cam7, s->host.look_from, 12); memcpy(cam7 + 3, s->host.look_at, 12); cam7[6] = s->host.fov; }
  return 0;
}

int ray_b200_prepared_info(struct futhark_context *ctx, const struct futhark_opaque_prepared_scene *p, struct ray_b200_bvh_info *info) {
  if (bad_ctx(ctx) || !p || !info) return 1;
  memset(info, 0, sizeof *info);
  info->n_leaves = p->n; info->n_inner = p->n - 1;
  info->max_depth = p->max_depth; info->refit_sweeps = p->refit_sweeps; info->stale_nodes = p->stale_nodes;
  RenderParams P;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (fill_params(ctx, p, 8, 8, 1, 0, 1, nullptr, nullptr, false, P) == 0) info->smem_nodes = P.smem_nodes;
  memcpy(info->root_box, p->root_box, sizeof info->root_box);
  memcpy(info->camera, &p->cam, sizeof info->camera);
  return 0;
}
int ray_b200_prepared_dump(struct futhark_context *ctx, const struct futhark_opaque_prepared_scene *p, uint32_t *morton, int32_t *perm,
                           int32_t *left, int32_t *right, int32_t *parent, float *boxes) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  if (!p || !p->dev.block) { set_error(ctx, "prepared_dump: invalid prepared scene"); return 1; }
  const size_t n = (size_t)p->n, ni = n - 1;
  const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
  if (morton) CUDA_TRY(ctx, cudaMemcpyAsync(morton, p->dev.morton, n * 4, k, ctx->stream));
  if (perm) CUDA_TRY(ctx, cudaMemcpyAsync(perm, p->dev.perm, n * 4, k, ctx->stream));
  if (left) CUDA_TRY(ctx, cudaMemcpyAsync(left, p->dev.left, ni * 4, k, ctx->stream));
  if (right) CUDA_TRY(ctx, cudaMemcpyAsync(right, p->dev.right, ni * 4, k, ctx->stream));
  if (parent) CUDA_TRY(ctx, cudaMemcpyAsync(parent, p->dev.parent, ni * 4, k, ctx->stream));
  if (boxes) CUDA_TRY(ctx, cudaMemcpyAsync(boxes, p->dev.boxes, ni * 24, k, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
// The packed BVH2C arrays as they sit in HBM: nodes[(n-1)*16 floats], nodes_soa[same], geom[n*4], colour[n*4]. NULLs skipped.
int ray_b200_prepared_packed(struct futhark_context *ctx, const struct futhark_opaque_prepared_scene *p, float *nodes, float *nodes_soa,
                             float *geom, float *colour) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->cfg.device);  // helper contexts of the single-process multi-GPU mode switch devices
  if (!p || !p->dev.block) { set_error(ctx, "prepared_packed: invalid prepared scene"); return 1; }
  const size_t n = (size_t)p->n, ni = n - 1;
  const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
  if (nodes) CUDA_TRY(ctx, cudaMemcpyAsync(nodes, p->dev.nodes, ni * 64, k, ctx->stream));
  if (nodes_soa) CUDA_TRY(ctx, cudaMemcpyAsync(nodes_soa, p->dev.nodes_soa, ni * 64, k, ctx->stream));
  if (geom) CUDA_TRY(ctx, cudaMemcpyAsync(geom, p->dev.geom, n * 16, k, ctx->stream));
  if (colour) CUDA_TRY(ctx, cudaMemcpyAsync(colour, p->dev.colour, n * 16, k, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ray_b200_prepared_reupload(struct futhark_context *ctx, struct futhark_opaque_prepared_scene *p) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!p || !p->dev.block) { set_error(ctx, "prepared_reupload: invalid prepared scene"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  return prepare_any(ctx, p);  // H2D of the sphere records again + the device LBVH build (or the host path)
}
int64_t ray_b200_prepared_device_bytes(struct futhark_context *ctx, const struct futhark_opaque_prepared_scene *p) {
  (void)ctx;
  return p ? (int64_t)p->dev.block_bytes : -1;
}
// Bytes prepare_scene / prepared_reupload copy host -> device (the sphere records on the device-build path).
int64_t ray_b200_prepared_upload_bytes(struct futhark_context *ctx, const struct futhark_opaque_prepared_scene *p) {
  (void)ctx;
  if (!p) return -1;
  return p->host_built ? (int64_t)p->dev.block_bytes : (int64_t)(p->host.spheres.size() * sizeof(SphereRec));
}

int ray_b200_render_into(struct futhark_context *ctx, int32_t *out_pix_dev, float *out_rgb_dev, int64_t h, int64_t w, int32_t spp,
                         const struct futhark_opaque_prepared_scene *p) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!out_pix_dev) { set_error(ctx, "render_into: out_pix_dev is required"); return 1; }
  if (!ctx->peers.empty()) { set_error(ctx, "render_into: not available on a single-process multi-GPU context (RAY_GPUS > 1): only futhark_entry_render / ray_b200_entry_render_spp gather the helper devices' tiles"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  RenderParams P;
  if (fill_params(ctx, p, h, w, spp, ctx->cfg.rank, ctx->cfg.world, out_pix_dev, out_rgb_dev, false, P)) return 1;
  return do_render(ctx, P);
}

int ray_b200_render_host(struct futhark_context *ctx, int32_t *out_pix_host, float *out_rgb_host, int64_t h, int64_t w, int32_t spp,
                         const struct futhark_opaque_prepared_scene *p) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!out_pix_host) { set_error(ctx, "render_host: out_pix_host is required"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const size_t px = (size_t)h * (size_t)w;
  int32_t *d_pix = nullptr;
  float *d_rgb = nullptr;
  CUDA_TRY(ctx, cudaMallocAsync(&d_pix, px * 4, ctx->stream));
  if (out_rgb_host) CUDA_TRY(ctx, cudaMallocAsync(&d_rgb, px * 12, ctx->stream));
  RenderParams P;
  int rc = fill_params(ctx, p, h, w, spp, 0, 1, d_pix, d_rgb, false, P) || do_render(ctx, P);
  if (!rc) {
    if (cudaMemcpyAsync(out_pix_host, d_pix, px * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = 1;
    if (!rc && d_rgb && cudaMemcpyAsync(out_rgb_host, d_rgb, px * 12, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = 1;
    if (rc) set_error(ctx, "render_host: device-to-host copy failed");
  }
  cudaFreeAsync(d_pix, ctx->stream);
  if (d_rgb) cudaFreeAsync(d_rgb, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) { set_error(ctx, "render_host: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; }
  return rc;
}

int64_t ray_b200_shard_tiles(int64_t h, int64_t w, int32_t rank, int32_t world) {
  if (world < 1 || rank < 0 || rank >= world || h <= 0 || w <= 0) return -1;
  return tiles_of_rank(h, w, rank, world);
}
int64_t ray_b200_shard_tiles_padded(int64_t h, int64_t w, int32_t world) {
  if (world < 1 || h <= 0 || w <= 0) return -1;
  return (tiles_total(h, w) + world - 1) / world;
}
int ray_b200_render_shard_into(struct futhark_context *ctx, int32_t *out_tiles_dev, int64_t h, int64_t w, int32_t spp,
                               const struct futhark_opaque_prepared_scene *p) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!out_tiles_dev) { set_error(ctx, "render_shard_into: null output"); return 1; }
  if (!ctx->peers.empty()) { set_error(ctx, "render_shard_into: not available on a single-process multi-GPU context (RAY_GPUS > 1)"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  RenderParams P;
  if (fill_params(ctx, p, h, w, spp, ctx->cfg.rank, ctx->cfg.world, out_tiles_dev, nullptr, true, P)) return 1;
  // ranks that own one tile fewer than the padded count leave a zeroed tail
  const int64_t padded = ray_b200_shard_tiles_padded(h, w, ctx->cfg.world);
  if (P.local_tiles < padded)
    CUDA_TRY(ctx, cudaMemsetAsync(out_tiles_dev + P.local_tiles * kTilePixels, 0, (size_t)(padded - P.local_tiles) * kTilePixels * 4, ctx->stream));
  return do_render(ctx, P);
}
// Several frames in one call, up to two of them in flight: job i runs on lane i % 2, lane 1 forks from the context's
// stream at the start of the call and joins it at the end, so for the caller the batch behaves like one stream-ordered
// operation.  Why: a frame ends on its longest paths (a 50-bounce path is ~50 dependent traversals) while most SMs are
// already idle; the persistent kernel of the NEXT frame cannot start there before the stream order lets it.  With the
// second lane its CTAs take over every SM the moment the previous frame's CTA retires.
int ray_b200_render_batch(struct futhark_context *ctx, const struct ray_b200_render_job *jobs, int32_t n) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (n < 0 || (n > 0 && !jobs)) { set_error(ctx, "render_batch: bad arguments"); return 1; }
  if (n == 0) return 0;
  if (!ctx->peers.empty()) { set_error(ctx, "render_batch: not available on a single-process multi-GPU context (RAY_GPUS > 1)"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  for (int32_t i = 0; i < n; i++) {
    if (!jobs[i].out_dev) { set_error(ctx, "render_batch: job %d has no output buffer", i); return 1; }
    if (jobs[i].shard_layout && jobs[i].out_rgb_dev) { set_error(ctx, "render_batch: job %d: no float output in the shard layout", i); return 1; }
  }
  const int kernel = resolve_kernel(ctx);
  // the wavefront kernel's ray queues and the warp trace exist once per context: those batches run in order on lane 0
  const bool two_lanes = n > 1 && kernel != RAY_B200_KERNEL_WAVEFRONT && !ctx->warp_trace;
  futhark_context::Lane &L1 = ctx->lanes[1];
  if (two_lanes && !L1.stream) {
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&L1.stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaMalloc(&L1.work_cursor, 64));
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
  if (two_lanes) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(L1.stream, ctx->ev_fork, 0));
  }
  int rc = 0;
  for (int32_t i = 0; i < n && !rc; i++) {
    const ray_b200_render_job &j = jobs[i];
    const int lane = two_lanes ? (i & 1) : 0;
    const int32_t spp = j.spp > 0 ? j.spp : ctx->cfg.spp;
    RenderParams P;
    if (j.shard_layout) {
      rc = fill_params(ctx, j.prepared, j.h, j.w, spp, ctx->cfg.rank, ctx->cfg.world, j.out_dev, nullptr, true, P);
      const int64_t padded = rc ? 0 : ray_b200_shard_tiles_padded(j.h, j.w, ctx->cfg.world);
      if (!rc && P.local_tiles < padded &&
          cudaMemsetAsync(j.out_dev + P.local_tiles * kTilePixels, 0, (size_t)(padded - P.local_tiles) * kTilePixels * 4,
                          lane ? L1.stream : ctx->stream) != cudaSuccess) {
        set_error(ctx, "render_batch: memset failed");
        rc = 1;
      }
    } else {
      rc = fill_params(ctx, j.prepared, j.h, j.w, spp, ctx->cfg.rank, ctx->cfg.world, j.out_dev, j.out_rgb_dev, false, P);
    }
    if (!rc) rc = do_render(ctx, P, lane, false);
  }
  // join even after a failure: whatever was enqueued on lane 1 must be ordered before later work on the context's stream
  if (two_lanes) {
    if (cudaEventRecord(ctx->ev_join, L1.stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0) != cudaSuccess) {
      if (!rc) set_error(ctx, "render_batch: join failed");
      rc = 1;
    }
  }
  if (!rc) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    ctx->have_timing = true;
  }
  return rc;
}

int64_t ray_b200_render_job_size(void) { return (int64_t)sizeof(struct ray_b200_render_job); }

int ray_b200_detile(struct futhark_context *ctx, const int32_t *gathered_dev, int32_t *out_pix_dev, int64_t h, int64_t w, int32_t world) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!gathered_dev || !out_pix_dev || world < 1 || h <= 0 || w <= 0) { set_error(ctx, "detile: bad argument"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  launch_detile(gathered_dev, out_pix_dev, h, w, world, ray_b200_shard_tiles_padded(h, w, world), ctx->stream, &ctx->launches);
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

int ray_b200_count_work(struct futhark_context *ctx, int64_t h, int64_t w, int32_t spp, const struct futhark_opaque_prepared_scene *p,
                        struct ray_b200_counters *out) {
  if (bad_ctx(ctx)) return 1;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!out) { set_error(ctx, "count_work: null output"); return 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  int32_t *scratch = nullptr;
  CUDA_TRY(ctx, cudaMallocAsync(&scratch, (size_t)h * w * 4, ctx->stream));
  RenderParams P;
  int rc = fill_params(ctx, p, h, w, spp, 0, 1, scratch, nullptr, false, P);
  if (!rc) {
    cudaMemsetAsync(ctx->counters, 0, 4 * sizeof(unsigned long long), ctx->stream);
    launch_count_work(P, ctx->stream, &ctx->launches);
    unsigned long long host[4] = {0, 0, 0, 0};
    if (cudaMemcpyAsync(host, ctx->counters, sizeof host, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error(ctx, "count_work: %s", cudaGetErrorString(cudaGetLastError()));
      rc = 1;
    }
    out->segments = host[0]; out->node_steps = host[1]; out->box_tests = host[2]; out->leaf_tests = host[3];
  }
  cudaFreeAsync(scratch, ctx->stream);
  return rc;
}

// ---- host-only entry points (no context, no device): the setup-path logic, testable without a GPU ----
int ray_b200_host_scene(const char *name, int64_t n, uint64_t seed, float *spheres, int64_t capacity, float *cam7, int64_t *count) {
  HostScene s;
  if (!name) return 1;
  if (!strcmp(name, "rgbbox")) make_rgbbox(s);
  else if (!strcmp(name, "irreg")) make_irreg(s);
  else if (!strcmp(name, "random")) make_random(s, n, seed);
  else return 1;
  if (count) *count = (int64_t)s.spheres.size();
  if (spheres) {
    if (capacity < (int64_t)s.spheres.size()) return 2;
    memcpy(spheres, s.spheres.data(), s.spheres.size() * sizeof(SphereRec));
  }
  if (cam7) { memcpy(cam7, s.look_from, 12); memcpy(cam7 + 3, s.look_at, 12); cam7[6] = s.fov; }
  return 0;
}
int ray_b200_host_camera(const float *cam7, int64_t h, int64_t w, float *out12) {
  if (!cam7 || !out12 || h <= 0 || w <= 0) return 1;
  HostScene s;
  memcpy(s.look_from, cam7, 12); memcpy(s.look_at, cam7 + 3, 12); s.fov = cam7[6];
  const CameraRec c = make_camera(s, h, w);
  memcpy(out12, &c, sizeof c);
  return 0;
}
int ray_b200_host_lbvh(const float *spheres, int64_t n, uint32_t *morton, int32_t *perm, int32_t *left, int32_t *right,
                       int32_t *parent, float *boxes, int32_t *info4) {
  if (!spheres || n < 0) return 1;
  HostScene s;
  s.spheres.resize((size_t)n);
  memcpy(s.spheres.data(), spheres, (size_t)n * sizeof(SphereRec));
  Lbvh t;
  std::string err;
  if (!build_lbvh(s, t, &err)) return 2;
  if (morton) memcpy(morton, t.morton.data(), t.morton.size() * 4);
  if (perm) memcpy(perm, t.perm.data(), t.perm.size() * 4);
  if (left) memcpy(left, t.left.data(), t.left.size() * 4);
  if (right) memcpy(right, t.right.data(), t.right.size() * 4);
  if (parent) memcpy(parent, t.parent.data(), t.parent.size() * 4);
  if (boxes) memcpy(boxes, t.boxes.data(), t.boxes.size() * 4);
  if (info4) { info4[0] = t.refit_sweeps; info4[1] = t.max_depth; info4[2] = t.stale_nodes; info4[3] = 0; }
  return 0;
}
void ray_b200_host_sample_offsets(int32_t spp, float *table) {
  std::vector<float> t;
  sample_offsets(spp, t);
  memcpy(table, t.data(), t.size() * sizeof(float));
}

const char *ray_b200_version(void) { return "ray_b200 0.1 (sm_100a)"; }

}  // extern "C"
