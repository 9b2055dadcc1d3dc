// sm_100a render kernels of the product build: K0 (thread per pixel: the parity anchor and the counting kernel), the
// tile probe of the heavy-first claim order, K3 (warp work-queue) and the multi-GPU de-tiling kernel, plus the host
// launchers.  K5 (the lane-walk kernel) is in render_lanewalk.cu, the measured-slower alternatives K1 / K2 / K4 in
// render_alt_kernels.cu (compiled with RAYB200_ALL_KERNELS only).  Shared device code: render_common.cuh.
#include <cub/cub.cuh>
#include "render_common.cuh"

namespace rayb200 {

namespace {

// ====================================================================================== K0: megakernel
// One thread per pixel, the whole ray_colour loop inside (what a Futhark GPU backend would emit,
// SURVEY.md §8a A11).  The parity anchor and the "naive GPU" number.
template <bool kCount>
__global__ void __launch_bounds__(128) render_mega_kernel(const __grid_constant__ RenderParams P) {
  const long long k64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  WorkCounters wc;
  int i = 0, j = 0;
  const bool valid = k64 < P.local_tiles * kTilePixels && item_pixel(P, (int)k64, i, j);
  if (valid) {
    const GlobalScene sc{P.nodes, P.geom};
    V3 sum = v3(0.0f, 0.0f, 0.0f);
    for (int s = 0; s < P.spp; s++) {
      Ray r = primary_ray(P, i, j, s);
      V3 light = v3(1.0f, 1.0f, 1.0f), colour;
      int depth = 0;
      while (advance_path<kCount>(sc, P, r, light, depth, colour, wc)) {}
      sum = (s == 0) ? colour : vadd(sum, colour);
    }
    write_pixel(P, (int)k64, i, j, sum);
  } else if (P.tile_major && k64 < P.local_tiles * kTilePixels) {
    P.out_pix[k64] = 0;  // padding pixel of a partial tile
  }
  if (kCount) flush_counters(P, wc);
}

// ====================================================================================== heavy-first claim order
// The persistent kernels end a frame on whatever was claimed last; a 50-bounce path claimed late keeps one warp busy
// for ~50 dependent traversals while the other SMs idle (profiles/r1_trace_tail.json: 44 % of the SM time of an
// irreg 1/8 shard at 64 spp).  Sorting the tiles longest-first cures the tail but un-mixes the frame — first every
// warp waits on L2 for deep nodes, then every warp shades sky — and the bulk gets 20 % slower.  So the claim order
// stays the chunk permutation, and only the tiles whose probe path is still alive after `probe_segments` segments
// ("heavy": possibly 50 bounces) are pulled into its first half: key = position in the permuted sequence, halved for
// heavy tiles, stable-sorted.  What is claimed in the second half ends within probe_segments traversals.  The probe
// traces sample 0 of 1, 2 or 4 pixels per 8x4 tile, so its own latency is bounded by the same cap.  The order only
// changes WHEN a pixel is rendered, never its value.
__global__ void __launch_bounds__(256) tile_probe_kernel(const __grid_constant__ RenderParams P, uint32_t *__restrict__ keys,
                                                         int32_t *__restrict__ ids) {
  const int probes = P.probes_per_tile;  // 1, 2 or 4: the lanes of a tile are neighbours in a warp
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lt = t / probes;
  const int q = (int)(t - lt * probes);
  bool alive = false;
  if (lt < P.local_tiles) {
    const int sub = q == 0 ? 11 : (q == 1 ? 22 : (q == 2 ? 17 : 5));  // pixels (3,1) (6,2) (1,2) (5,0) of the tile
    int i, j;
    if (item_pixel(P, (int)(lt << 5) | sub, i, j)) {
      const GlobalScene sc{P.nodes, P.geom};
      Ray r = primary_ray(P, i, j, 0);
      V3 light = v3(1.0f, 1.0f, 1.0f), colour;
      int depth = 0, segments = 0;
      WorkCounters wc;
      do {
        alive = advance_path<false>(sc, P, r, light, depth, colour, wc);
      } while (alive && ++segments < P.probe_segments);
    }
  }
  if (probes > 1) alive |= __shfl_xor_sync(kFullMask, alive, 1);
  if (probes > 2) alive |= __shfl_xor_sync(kFullMask, alive, 2);
  if (q == 0 && lt < P.local_tiles) {
    const long long chunk = lt >> 6;
    const uint32_t pos = (uint32_t)((chunk * P.chunk_stride_inv) % P.n_chunks) << 6 | (uint32_t)(lt & 63);
    keys[lt] = alive ? pos >> 1 : pos;
    ids[lt] = (int32_t)lt;
  }
}

// Learned claim order: the probe replaced by what the previous frame of the same scene measured (the longest path of every
// tile, recorded by finish_path).  Tiles with a long path keep their place in the permuted sequence but move in front of
// all others, so the 50-bounce paths that end a frame are claimed in its first few percent.
__global__ void __launch_bounds__(256) tile_cost_keys_kernel(const __grid_constant__ RenderParams P, const uint32_t *__restrict__ cost,
                                                             const int long_path, uint32_t *__restrict__ keys, int32_t *__restrict__ ids) {
  const long long lt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lt >= P.local_tiles) return;
  const long long chunk = lt >> 6;
  const uint32_t pos = (uint32_t)((chunk * P.chunk_stride_inv) % P.n_chunks) << 6 | (uint32_t)(lt & 63);
  // long_path > 0: two classes (>= long_path first); long_path < 0: four classes with thresholds 8x / 3x / 1x |long_path|
  const uint32_t c = cost[lt];
  uint32_t cls;
  if (long_path > 0) cls = c >= (uint32_t)long_path ? 0u : 3u;
  else { const uint32_t lp = (uint32_t)(-long_path); cls = c >= 8u * lp ? 0u : (c >= 3u * lp ? 1u : (c >= lp ? 2u : 3u)); }
  keys[lt] = (cls << 29) | pos;
  ids[lt] = (int32_t)lt;
}

// ====================================================================================== K3: warp work-queue
// Dense traversal.  K1/K2 bind a lane to a ray for a whole segment, so a warp's SIMT efficiency is
// mean/max of its lanes' traversal lengths (ncu: 6.5 of 32 lanes active on rgbbox).  The reference's
// fold is order-free and never prunes by the running closest t, so a segment's traversal is just a
// SET of independent (ray, node) box-test items and (ray, leaf) sphere-test items.  K3 therefore
// binds lanes to ITEMS, not rays: every warp keeps R = 32*K rays in flight, their traversal items on
// a warp-private queue in shared memory, and each iteration takes 32 items — whatever rays they belong
// to — so every lane does one node step (two box tests) or one sphere test.  Children are pushed back
// with ballot/popc compaction; a sphere hit is folded into its ray's (t, leaf index) pair with two
// 32-bit shared-memory atomicMin (fold_hit), which is exactly the reference's "smallest t, lowest leaf
// index on ties".  When the queues run dry every ray of the round has its closest hit; the owner lanes
// shade, bounce, and idle slots are refilled, and the next round starts.
//
// Queue order and bound.  The order of the items is free, so it serves the batches (see drain): the
// node queue is a ring, batches take its OLDEST items while it is short (breadth-first: only short
// subtrees are left at the end of a round) and the newest above that (depth-first: bounded growth).
// Taken newest-first only, items pushed in reverse lane order keep the queue sorted by tree depth and
// R + 64*(max_depth+1) entries always suffice; the kernel does not rely on that bound: whenever fewer
// than 96 entries are free it takes items one at a time (a plain DFS adds at most `depth` < 64 entries
// before it shrinks), so any ring of >= 256 entries is safe for any content.  Leaf items are drained
// whenever 32 are available, so that stack never holds more than 31 + 64.
//
// Work distribution.  spp == 1 (kSpread = false): a slot owns a pixel (and, for robustness, all its
// samples), claimed from the global cursor with one warp-aggregated atomicAdd.  spp > 1 (kSpread):
// a warp opens up to kRing pixels at a time and hands their samples, in order, to whichever slots are
// idle — consecutive lanes trace consecutive samples of the same pixel (coherent), the scheduling
// grain is one sample instead of one pixel x spp (no long tails, scales to many GPUs), and every
// finished sample's colour is parked in an L2-resident buffer so that the pixel is summed in SAMPLE
// ORDER when its last sample lands, which keeps the result bit-identical to the sequential definition.

template <int K, bool kSpread, bool kPacket, bool kAllNodes, bool kSpheres>
__global__ void __launch_bounds__(kWqMaxThreads, 1) render_warpqueue_kernel(const __grid_constant__ RenderParams P, const int ncap_arg,
                                                                   const int packet_min, const int refill_min, const int nlow) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (P.warp_trace && threadIdx.x == 0) atomicMin(P.warp_trace, global_timer_ns());
  const float4 *s_nodes, *s_geom;
  stage_scene(P, smem_raw, s_nodes, s_geom);
  const StagedScene<kAllNodes, kSpheres> sc{P.nodes, P.geom, s_nodes, s_geom, P.smem_nodes};

  constexpr int R = 32 * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the node queue is a ring of the largest power of two that fits its allocation (the launcher hands out powers of two
  // unless a caller asked for an odd capacity)
  const int ncap = 1 << (31 - __clz(ncap_arg));
  const unsigned nmask = (unsigned)ncap - 1u;
  // Oldest-first node batches (see drain) need the ring; measured, they pay everywhere except where the packet walk meets
  // spread samples (64 consecutive samples of a pixel in a warp: irreg 64 spp 13.5 ms with the ring against 12.8 without -
  // 8 more instructions per batch buy 1.6 % - while 1-spp frames of the same scenes gain 2-6 %), so that combination
  // keeps the plain stack: no index masks, newest first.
  constexpr bool kDeque = !(kPacket && kSpread);
  auto ring = [&](const int i) { return kDeque ? (int)((unsigned)i & nmask) : i; };
  // Lane masks.  At 64 registers the compiler does not keep them live across the batch bodies; re-reading %lanemask_lt/gt
  // is one instruction where re-deriving them from the thread index is four, but it also changes the register
  // allocation: measured A/B (profiles/r2_sweep_kernel_ab.json) the special registers win on the packet variants
  // (irreg 64 spp 13.88 -> 13.49 ms, 1 M spheres 77.6 -> 76.4) and lose on the fully staged one (rgbbox 38.05 -> 38.64;
  // re-measured after the instruction-cache work: no difference either way, 32.16 vs 32.14 ms).
  unsigned lt_mask, gt_mask;
  if constexpr (kPacket) {
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(gt_mask));
  } else {
    lt_mask = (1u << lane) - 1u;
    gt_mask = lane == 31 ? 0u : ~((2u << lane) - 1u);
  }
  unsigned char *wbase = smem_raw + ((staging_bytes(P) + 127) & ~(size_t)127) + (size_t)warp * wq_warp_bytes(K, ncap_arg, kPacket, kSpread);
  float4 *ray_o = reinterpret_cast<float4 *>(wbase);   // {o.xyz, a = dot d d}
  float4 *ray_i = ray_o + R;                           // {1/d.xyz, 0}
  float4 *ray_d = ray_i + R;                           // {d.xyz, 0}
  float4 *p_light = ray_d + R;                         // {light.rgb, bits(depth)}   owner lane only
  float4 *p_sum = p_light + R;                         // pixel-bound samples: {sum.rgb, bits(sample)}
  int *p_ms = reinterpret_cast<int *>(p_light + R);    // spread samples: ring entry << 16 | sample (4 bytes per slot instead of 16)
  uint32_t *best_t = kSpread ? reinterpret_cast<uint32_t *>(p_ms + R)   // bits(t) of the closest accepted hit (kItemNoHit = none) ...
                             : reinterpret_cast<uint32_t *>(p_sum + R);
  uint32_t *best_l = best_t + R;                       // ... and its leaf, lowest index among equal t (fold_hit)
  int *p_item = reinterpret_cast<int *>(best_l + R);   // work item (pixel) of the slot, -1 = idle
  int *ring_item = p_item + R;                         // spread: pixel item of ring entry m
  int *ring_done = ring_item + kWqRing;                // spread: samples finished, -1 = entry free
  int *pk_node = ring_done + kWqRing;                  // packet walk: deferred (node, lane mask) pairs, warp-uniform
  unsigned *pk_mask = reinterpret_cast<unsigned *>(pk_node + kWqPacketStack);
  uint32_t *lstk = reinterpret_cast<uint32_t *>(kPacket ? pk_mask + kWqPacketStack : reinterpret_cast<unsigned *>(pk_node));
  uint32_t *nstk = lstk + kWqLeafStack;

  const int total = (int)(P.local_tiles * kTilePixels);
  const int total_claims = P.n_chunks << 11;
  const int spp = P.spp;
#pragma unroll
  for (int k = 0; k < K; k++) p_item[lane + 32 * k] = -1;
  if (lane < kWqRing) ring_done[lane] = -1;
  __syncwarp();
  bool exhausted = false;
  int ntop = 0, nhead = 0, ltop = 0;           // warp-uniform: node queue [nhead, ntop) (ring; nhead = 0 outside drain), leaf stack height
  int open_seq = 0, disp_seq = 0, disp_s = 0;  // spread dispenser (warp-uniform): pixels opened / next sample to hand out
  float4 *cbuf = nullptr;                      // spread: [kWqRing][spp] finished-sample colours of this warp
  if (kSpread) cbuf = P.sample_buf + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * kWqRing * (size_t)spp;

  // A path in `slot` has ended with `colour`.
  auto finish_path = [&](const int slot, const V3 colour, const int segs) {
    // learned claim order (recording frame only): the longest path of the tile, sample 0 of its 32 pixels is evidence enough
    if (P.tile_cost && ((kSpread ? p_ms[slot] : __float_as_int(p_sum[slot].w)) & 0xffff) == 0)
      atomicMax(P.tile_cost + (p_item[slot] >> 5), (unsigned)segs);
    if (kSpread) {
      const int ms = p_ms[slot];
      __stcg(cbuf + (size_t)(ms >> 16) * spp + (ms & 0xffff), make_float4(colour.x, colour.y, colour.z, 0.0f));
      atomicAdd(ring_done + (ms >> 16), 1);
      p_item[slot] = -1;
    } else {
      const int item = p_item[slot];
      const float4 ps = p_sum[slot];
      int s = __float_as_int(ps.w);
      const V3 sum = (s == 0) ? colour : vadd(v3(ps.x, ps.y, ps.z), colour);
      s++;
      int pi, pj;
      item_pixel(P, item, pi, pj);
      if (s < spp) {
        const Ray nr = primary_ray(P, pi, pj, s);
        ray_o[slot] = make_float4(nr.o.x, nr.o.y, nr.o.z, 0.0f);
        ray_d[slot] = make_float4(nr.d.x, nr.d.y, nr.d.z, 0.0f);
        p_light[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));
        p_sum[slot] = make_float4(sum.x, sum.y, sum.z, __int_as_float(s));
      } else {
        write_pixel(P, item, pi, pj, sum);
        p_item[slot] = -1;
      }
    }
  };
  // spread: pixels whose last sample has landed are summed IN SAMPLE ORDER and written; frees the ring entry
  auto finalize_pixels = [&]() {
    __syncwarp();
    // four lanes per ring entry, one per colour channel: the in-order sum is a dependent chain of spp - 1 additions per
    // channel, so the three channels run side by side (one lane doing all three cost 3.6 % of irreg's instructions)
    const int m = lane >> 2, ch = lane & 3;
    const bool ready = ring_done[m] == spp;
    if (__any_sync(kFullMask, ready)) {
      float sum = 0.0f;
      if (ready && ch < 3) {
        const float *c = reinterpret_cast<const float *>(cbuf + (size_t)m * spp) + ch;
        sum = __ldcg(c);
#pragma unroll 4
        for (int s = 1; s < spp; s++) sum = sum + __ldcg(c + 4 * s);
      }
      const float g = __shfl_down_sync(kFullMask, sum, 1), b = __shfl_down_sync(kFullMask, sum, 2);
      if (ready && ch == 0) {
        const int item = ring_item[m];
        int pi, pj;
        if (item_pixel(P, item, pi, pj)) write_pixel(P, item, pi, pj, v3(sum, g, b));
        else if (P.tile_major) P.out_pix[item] = 0;
      }
      __syncwarp();  // every lane of the entry has read ring_done before it is reset
      if (ready && ch == 0) ring_done[m] = -1;
    }
    __syncwarp();
  };

  for (;;) {
    unsigned trav = 0;  // bit k: slot lane+32k has a traversal in flight this round
    // warp-uniform: lanes whose slot lane+32k starts a traversal this round.  (The per-k loops below are NOT unrolled:
    // ncu had the K = 2 kernel's instruction working set at 34 KB against a 32 KB instruction cache - 9 % of the fetches
    // missed it and the GPC-level cache behind it ran at 94 % of its peak - and refill / set-up / shading were in it twice.)
    unsigned gomask0 = 0u, gomask1 = 0u;
    for (int pass = 0; pass < 4; pass++) {
      // ---------------------------------------------------------------- hand work to idle slots
      if (kSpread) finalize_pixels();
      int my_idle = 0;
#pragma unroll
      for (int k = 0; k < K; k++) my_idle += p_item[lane + 32 * k] < 0;
      int incl = my_idle;  // inclusive warp scan of the idle counts
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += v;
      }
      const int cnt = __shfl_sync(kFullMask, incl, 31);
      int rank = incl - my_idle;
      if (!kSpread) {
        if (cnt && !exhausted) {
          int base = 0;
          if (lane == 0) base = atomicAdd(P.work_cursor, cnt);
          base = __shfl_sync(kFullMask, base, 0);
#pragma unroll 1
          for (int k = 0; k < K; k++) {
            const int slot = lane + 32 * k;
            if (p_item[slot] < 0) {
              const int c = base + rank++;
              const int item = c < total_claims ? claim_to_item(P, c) : total;
              int pi, pj;
              if (item < total) {
                if (item_pixel(P, item, pi, pj)) {
                  const Ray r = primary_ray(P, pi, pj, 0);
                  p_item[slot] = item;
                  ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, 0.0f);
                  ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
                  p_light[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));
                  p_sum[slot] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
                } else if (P.tile_major) {
                  P.out_pix[item] = 0;
                }
              }
            }
          }
          exhausted = base + cnt >= total_claims;
        }
      } else if (cnt >= refill_min || cnt == R) {
        // (refill_min > 1: wait until several slots are idle, so they get CONSECUTIVE samples and start as a coherent group)
        int avail = (open_seq - disp_seq) * spp - disp_s;
        while (!exhausted && avail < cnt) {  // open more pixels (one cursor claim each) while the ring has room
          const int m = open_seq & (kWqRing - 1);
          if (ring_done[m] != -1) break;
          int c = 0;
          if (lane == 0) c = atomicAdd(P.work_cursor, 1);
          c = __shfl_sync(kFullMask, c, 0);
          if (c >= total_claims) { exhausted = true; break; }
          const int item = claim_to_item(P, c);
          if (item >= total) continue;  // tail of the last (partial) chunk
          __syncwarp();                 // every lane has read ring_done[m] before lane 0 overwrites it
          if (lane == 0) { ring_item[m] = item; ring_done[m] = 0; }
          __syncwarp();
          open_seq++;
          avail += spp;
        }
        const int give = cnt < avail ? cnt : avail;
#pragma unroll 1
        for (int k = 0; k < K; k++) {
          const int slot = lane + 32 * k;
          if (p_item[slot] < 0) {
            const int rk = rank++;
            if (rk < give) {
              int s = disp_s + rk, seq = disp_seq;
              while (s >= spp) { s -= spp; seq++; }
              const int m = seq & (kWqRing - 1);
              const int item = ring_item[m];
              int pi, pj;
              if (item_pixel(P, item, pi, pj)) {
                const Ray r = primary_ray(P, pi, pj, s);
                p_item[slot] = item;
                ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, 0.0f);
                ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
                p_light[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));
                p_ms[slot] = (m << 16) | s;
              } else {
                atomicAdd(ring_done + m, 1);  // padding pixel of a partial tile: nothing to trace
              }
            }
          }
        }
        disp_s += give;
        while (disp_s >= spp) { disp_s -= spp; disp_seq++; }
      }
      // -------------------------------------------------------------- set up this round's segments
      // Root box test by the owner lane; a root miss is shaded (sky) on the spot and the path ends,
      // so sky rays never occupy a traversal round.
#pragma unroll 1
      for (int k = 0; k < K; k++) {
        const int slot = lane + 32 * k;
        bool go = false;
        if (!((trav >> k) & 1u)) {
          while (p_item[slot] >= 0) {
            const float4 ro = ray_o[slot], rd = ray_d[slot];
            Ray r;
            r.o = v3(ro.x, ro.y, ro.z);
            r.d = v3(rd.x, rd.y, rd.z);
            const RayInv q = ray_invariants(r);
            if (box_hit(P.root_box[0], P.root_box[1], P.root_box[2], P.root_box[3], P.root_box[4], P.root_box[5], r, q)) {
              ray_o[slot] = make_float4(ro.x, ro.y, ro.z, q.a);
              ray_i[slot] = make_float4(q.ix, q.iy, q.iz, 0.0f);
              best_t[slot] = kItemNoHit;
              go = true;
              break;
            }
            const float4 pl = p_light[slot];  // miss (ray.fut:141-148)
            V3 light = v3(pl.x, pl.y, pl.z), colour;
            int depth = __float_as_int(pl.w);
            shade_segment(sc, P, r, q.a, -1, 0.0f, light, depth, colour);
            finish_path(slot, colour, depth + 1);
          }
        }
        const unsigned m = __ballot_sync(kFullMask, go);
        if (go) trav |= 1u << k;
        if (k == 0) gomask0 |= m; else gomask1 |= m;  // the traversal of these rays starts at the root after the passes
      }
      // another pass only helps if some slot is idle and there is still work to hand out
      bool idle_left = false;
#pragma unroll
      for (int k = 0; k < K; k++) idle_left |= p_item[lane + 32 * k] < 0;
      const bool more = kSpread ? (!exhausted || (open_seq - disp_seq) * spp - disp_s > 0) : !exhausted;
      if (!more || !__any_sync(kFullMask, idle_left)) break;
    }
    __syncwarp();
    if ((gomask0 | gomask1) == 0u) {
      bool any_active = false;
#pragma unroll
      for (int k = 0; k < K; k++) any_active |= p_item[lane + 32 * k] >= 0;
      const bool undispensed = kSpread && (open_seq - disp_seq) * spp - disp_s > 0;
      if (exhausted && !undispensed && !__any_sync(kFullMask, any_active)) {
        if (kSpread) finalize_pixels();
        if (P.warp_trace && lane == 0) P.warp_trace[1 + blockIdx.x * (blockDim.x >> 5) + warp] = global_timer_ns();
        break;                                                      // frame done for this warp
      }
      continue;                                                     // only sky / padding so far: hand out more
    }

    // ---------------------------------------------------------------- dense traversal of the round
    // Both batch bodies exist twice: a full-warp version (32 items, no lane predicate: the four push flags stay in
    // predicate registers) used while the queues are deep enough, and a partial one for the drain.
    auto leaf_batch = [&](auto full_tag, const int n_part) {
      // closest_hit (ray.fut:78-81) for up to 32 (ray, sphere) pairs
      constexpr bool kFull = decltype(full_tag)::value;
      const int n = kFull ? 32 : n_part;
      bool hit = false;
      uint32_t tb = 0;
      int slot = 0, li = 0;
      if (kFull || lane < n) {
        const uint32_t it = lstk[ltop - 1 - lane];
        slot = (int)(it >> kSlotShift);
        li = (int)(it & kIndexMask);
        const float4 ro = ray_o[slot], rd = ray_d[slot];
        const float4 g = sc.sphere(li);
        Ray r;
        r.o = v3(ro.x, ro.y, ro.z);
        r.d = v3(rd.x, rd.y, rd.z);
        const float t = sphere_t(g.x, g.y, g.z, g.w, r, ro.w, 0.1f, 1000000000.0f);
        hit = t >= 0.0f;
        tb = __float_as_uint(t);
      }
      ltop -= n;
      fold_hit(best_t, best_l, slot, hit, tb, (uint32_t)li);
    };
    auto node_batch = [&](auto full_tag, const bool bottom, const int n_part) {
      // one BVH2C node step (both children's boxes) for up to 32 (ray, node) pairs, taken from the top of the queue
      // (newest = deepest first) or, `bottom` (warp-uniform), from its bottom (oldest = shallowest first)
      constexpr bool kFull = decltype(full_tag)::value;
      const int n = kFull ? 32 : n_part;
      bool pl_node = false, pr_node = false, pl_leaf = false, pr_leaf = false;
      uint32_t tag = 0;
      int lptr = 0, rptr = 0;
      if (kFull || lane < n) {
        const uint32_t it = nstk[ring(kDeque && bottom ? nhead + lane : ntop + ~lane)];
        tag = it & ~kIndexMask;
        const int slot = (int)(it >> kSlotShift), cur = (int)(it & kIndexMask);
        const float4 ro = ray_o[slot], ri = ray_i[slot];
        float4 q0, q1, q2, q3;
        sc.node(cur, q0, q1, q2, q3);
        Ray r;
        r.o = v3(ro.x, ro.y, ro.z);
        r.d = v3(0.0f, 0.0f, 0.0f);
        RayInv q;
        q.ix = ri.x; q.iy = ri.y; q.iz = ri.z; q.a = ro.w;
        lptr = __float_as_int(q0.w);
        rptr = __float_as_int(q1.w);
        const bool hl = box_hit(q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, r, q);
        const bool hr = box_hit(q2.x, q2.y, q2.z, q3.x, q3.y, q3.z, r, q);
        pl_leaf = lptr < 0;          // a leaf child has no box in the reference: always visited
        pr_leaf = rptr < 0;
        pl_node = hl && !pl_leaf;
        pr_node = hr && !pr_leaf;
      }
      __syncwarp();  // top pops: all of them have been read before anything is pushed over them
      if (kDeque && bottom) nhead += n; else ntop -= n;
      const unsigned bl = __ballot_sync(kFullMask, pl_node), br = __ballot_sync(kFullMask, pr_node);
      // reverse lane order: lane 0 popped the top (deepest) item, its children go back on top
      const int nb = ntop + __popc(bl & gt_mask) + __popc(br & gt_mask);
      if (pr_node) nstk[ring(nb)] = tag | (uint32_t)rptr;
      if (pl_node) nstk[ring(nb + (pr_node ? 1 : 0))] = tag | (uint32_t)lptr;
      ntop += __popc(bl) + __popc(br);
      // two batches in three have no leaf child at all (upper tree levels; the scheduling simulation of rgbbox counts
      // 651 of 973): one vote instead of the two ballots, six popc and the stores
      if (__any_sync(kFullMask, pl_leaf || pr_leaf)) {
        const unsigned cl = __ballot_sync(kFullMask, pl_leaf), cr = __ballot_sync(kFullMask, pr_leaf);
        const int lb = ltop + __popc(cl & lt_mask) + __popc(cr & lt_mask);
        if (pl_leaf) lstk[lb] = tag | (uint32_t)(~lptr);
        if (pr_leaf) lstk[lb + (pl_leaf ? 1 : 0)] = tag | (uint32_t)(~rptr);
        ltop += __popc(cl) + __popc(cr);
      }
    };
    using full_t = std::integral_constant<bool, true>;
    using part_t = std::integral_constant<bool, false>;
    // Runs the item queues dry.  Order is free (the fold is a min), so it is chosen for full batches: while the node
    // queue is short (<= nlow items) batches take its OLDEST items, which widens the frontier breadth-first and leaves only
    // deep items (short subtrees) for the end of the round; above nlow they take the newest (depth-first), which bounds
    // the queue.  With top-only popping the last items of a round were the root-level items at the bottom of the stack:
    // ncu had 27.7 of 32 lanes active in the box tests (rgbbox, 64 rays per warp), ~10 partial batches per round; the
    // scheduling simulation of the same rounds gives 0.87 -> 0.98 lane use and 11 % fewer node batches.
    // Overflow guard for ANY content: once fewer than 96 entries are free, items are taken one at a time from the top,
    // a plain DFS that can add at most (tree depth) < 64 entries before it shrinks again.
    auto drain = [&]() {
      __syncwarp();
      const int nroom = ncap - 96;
      for (;;) {
        // the common case first and alone in its loop: full node batches while nothing else is due (ncu: the general
        // dispatch below cost ~12 instructions per batch, 10 % of the kernel's instructions).  One copy of the batch
        // body for both pop directions.
        while ((unsigned)(ntop - nhead - 32) <= (unsigned)(nroom - 32) && ltop < 32) {
          node_batch(full_t{}, kDeque && ntop - nhead <= nlow, 32);
          __syncwarp();
        }
        const int cnt = ntop - nhead;
        if (ltop >= 32) leaf_batch(full_t{}, 32);
        else if (cnt > 0) node_batch(part_t{}, false, cnt > nroom ? 1 : (cnt < 32 ? cnt : 32));
        else if (ltop > 0) leaf_batch(part_t{}, ltop);
        else break;
        __syncwarp();
      }
      ntop = nhead = 0;  // empty: outside the drain the queue is a plain array again
    };
    // Packet walk.  Near the root almost every ray of a warp visits the same nodes, so those node steps are done the
    // cheap way: ONE (node, lane mask) pair for the whole warp, the node fetched once (same address in every lane =
    // a shared-memory broadcast), each lane testing its own slot's ray, leaf children tested inline by the owner lanes
    // (plain read-modify-write of their own `best` word).  As soon as fewer than `packet_min` lanes are left on a node
    // the remaining (ray, node) pairs are handed to the item queue, where lanes are bound to items instead of rays.
    [[maybe_unused]] auto packet_walk = [&](const int k, unsigned mask) {
      const int slot = lane + 32 * k;
      int cur = 0, psp = 0;
      for (;;) {
        const int cnt = __popc(mask);
        bool descended = false;
        if (cnt < packet_min) {
          if (cnt) {
            if (ntop + 32 + 96 > ncap) drain();
            if ((mask >> lane) & 1u) nstk[ntop + __popc(mask & lt_mask)] = ((uint32_t)slot << kSlotShift) | (uint32_t)cur;
            ntop += cnt;
          }
        } else {
          const bool in = (mask >> lane) & 1u;
          float4 q0, q1, q2, q3;
          sc.node(cur, q0, q1, q2, q3);
          const int lptr = __float_as_int(q0.w), rptr = __float_as_int(q1.w);  // same node in every lane
          bool hl = false, hr = false;
          if (in) {
            const float4 ro = ray_o[slot], ri = ray_i[slot];
            Ray r;
            r.o = v3(ro.x, ro.y, ro.z);
            RayInv q;
            q.ix = ri.x; q.iy = ri.y; q.iz = ri.z; q.a = ro.w;
            r.d = v3(0.0f, 0.0f, 0.0f);
            hl = box_hit(q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, r, q);
            hr = box_hit(q2.x, q2.y, q2.z, q3.x, q3.y, q3.z, r, q);
            if (lptr < 0 || rptr < 0) {  // leaf children: visited by every ray that visits this node (bvh.fut:84)
              const float4 rd = ray_d[slot];
              r.d = v3(rd.x, rd.y, rd.z);
              uint32_t bt = best_t[slot], bl = best_l[slot];  // owner lane, no item of this slot is queued yet: plain update
              if (lptr < 0) {
                const float4 g = sc.sphere(~lptr);
                const float t = sphere_t(g.x, g.y, g.z, g.w, r, ro.w, 0.1f, 1000000000.0f);
                const uint32_t tb = __float_as_uint(t), li = (uint32_t)(~lptr);
                if (t >= 0.0f && (tb < bt || (tb == bt && li < bl))) { bt = tb; bl = li; }
              }
              if (rptr < 0) {
                const float4 g = sc.sphere(~rptr);
                const float t = sphere_t(g.x, g.y, g.z, g.w, r, ro.w, 0.1f, 1000000000.0f);
                const uint32_t tb = __float_as_uint(t), li = (uint32_t)(~rptr);
                if (t >= 0.0f && (tb < bt || (tb == bt && li < bl))) { bt = tb; bl = li; }
              }
              best_t[slot] = bt;
              best_l[slot] = bl;
            }
          }
          const unsigned bl = lptr >= 0 ? __ballot_sync(kFullMask, hl) : 0u;
          const unsigned br = rptr >= 0 ? __ballot_sync(kFullMask, hr) : 0u;
          if (bl) {
            if (br) {
              __syncwarp();  // every lane is done reading the entry this may overwrite
              if (lane == 0) { pk_node[psp] = rptr; pk_mask[psp] = br; }
              psp++;
            }
            cur = lptr; mask = bl; descended = true;
          } else if (br) {
            cur = rptr; mask = br; descended = true;
          }
        }
        if (descended) continue;
        if (psp == 0) break;
        psp--;
        __syncwarp();
        cur = pk_node[psp];
        mask = pk_mask[psp];
      }
    };
#pragma unroll
    for (int k = 0; k < K; k++) {
      const unsigned gm = k == 0 ? gomask0 : gomask1;
      if (kPacket && packet_min > 0) {
        if constexpr (kPacket) {
          if (gm) packet_walk(k, gm);
        }
      } else {
        const bool go = (gm >> lane) & 1u;
        if (go) nstk[ntop + __popc(gm & lt_mask)] = (uint32_t)(lane + 32 * k) << kSlotShift;  // (slot, root node 0)
        ntop += __popc(gm);
      }
    }
    drain();

    // ---------------------------------------------------------------- shade: owner lanes finish the segment
#pragma unroll 1
    for (int k = 0; k < K; k++) {
      if (!((trav >> k) & 1u)) continue;
      const int slot = lane + 32 * k;
      const float4 ro = ray_o[slot], rd = ray_d[slot], pl = p_light[slot];
      const uint32_t bt = best_t[slot];
      Ray r;
      r.o = v3(ro.x, ro.y, ro.z);
      r.d = v3(rd.x, rd.y, rd.z);
      V3 light = v3(pl.x, pl.y, pl.z), colour;
      int depth = __float_as_int(pl.w);
      const int j = bt == kItemNoHit ? -1 : (int)best_l[slot];
      const float tb = __uint_as_float(bt);
      if (shade_segment(sc, P, r, ro.w, j, tb, light, depth, colour)) {
        ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, 0.0f);
        ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
        p_light[slot] = make_float4(light.x, light.y, light.z, __int_as_float(depth));
      } else {
        finish_path(slot, colour, depth + 1);
      }
    }
    __syncwarp();
  }
  signal_frame_done(P);
}

// ====================================================================================== peer-frame flags
// (see include/ray_b200.h "peer-memory frames"): the consumer's stream waits for the producers' kernels through a flag in
// its own memory that the producers bump over NVLink; the producers wait for the consumer's "slot is free again".
__global__ void flag_wait_kernel(const uint32_t *flag, uint32_t value, long long timeout_ns, unsigned long long *timeouts) {
  const unsigned long long t0 = global_timer_ns();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int32_t)(v - value) >= 0) return;
    if (timeout_ns > 0 && (long long)(global_timer_ns() - t0) > timeout_ns) { atomicAdd(timeouts, 1ull); return; }
    __nanosleep(200);
  }
}
__global__ void flag_set_kernel(uint32_t *flag, uint32_t value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__global__ void flag_bump_kernel(uint32_t *flag) {
  __threadfence_system();
  atomicAdd_system(flag, 1u);
}

// ====================================================================================== de-tiling (multi-GPU)
// gathered: [world][tiles_padded][32] as an NCCL gather of every rank's compact buffer lays it out.
__global__ void detile_kernel(const int32_t *__restrict__ gathered, int32_t *__restrict__ out, int H, int W, int world,
                              long long tiles_padded, int tiles_x) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)H * W) return;
  const int j = (int)(idx / W), i = (int)(idx - (long long)j * W);
  const long long t = (long long)(j / kTileH) * tiles_x + (i / kTileW);
  const int sub = (j % kTileH) * kTileW + (i % kTileW);
  const long long rank = t % world, lt = t / world;
  out[idx] = gathered[(rank * tiles_padded + lt) * kTilePixels + sub];
}

}  // namespace

// ---------------------------------------------------------------------------------------- host launchers
// Kernels opt in to > 48 KB of dynamic shared memory on first use (opt_in_dynamic_smem), not at context creation:
// futhark_context_new no longer touches ~50 template instantiations it will never launch.
cudaError_t launch_render(const RenderParams &p, const LaunchConfig &lc, const WavefrontBuffers *wf, cudaStream_t stream,
                          int64_t *launches) {
  const long long items = p.local_tiles * kTilePixels;
  if (items <= 0) return cudaSuccess;
  if (lc.kernel == 1) {  // RAY_B200_KERNEL_MEGA
    const int threads = 128;
    const unsigned blocks = (unsigned)((items + threads - 1) / threads);
    render_mega_kernel<false><<<blocks, threads, 0, stream>>>(p);
    (*launches)++;
    return cudaSuccess;
  }
  if (lc.kernel == 6) return launch_lanewalk(p, lc, stream, launches);             // RAY_B200_KERNEL_LANEWALK
  if (lc.kernel != 4) return launch_alt_kernel(p, lc, wf, stream, launches);       // K1 / K2 / K4 (RAYB200_ALL_KERNELS builds)
  // RAY_B200_KERNEL_WARPQUEUE: one CTA per SM, wq_warps warps, 32*wq_k rays in flight per warp
  const bool all_nodes = p.smem_nodes == p.n_inner, sph = p.smem_spheres == p.n_leaves && p.smem_spheres > 0;
  const int k = lc.wq_k == 1 ? 1 : 2;
  const int wthreads = 32 * lc.wq_warps;
  const int ncap = wq_node_capacity(k, p.max_depth, lc.wq_ncap);
  const bool packet = lc.wq_packet > 0;
  // breadth-first below this many queued node items (0 = half the ring, < 0 = never: plain LIFO)
  const int nring = 1 << (31 - __builtin_clz((unsigned)ncap));
  const int nlow = lc.wq_low < 0 ? 0 : (lc.wq_low == 0 ? nring / 2 : (lc.wq_low > nring - 96 ? nring - 96 : lc.wq_low));
  const size_t wsmem = ((staging_bytes(p) + 127) & ~(size_t)127) + (size_t)lc.wq_warps * wq_warp_bytes(k, ncap, packet, p.sample_buf != nullptr);
  long long ctas = lc.sm_count;
  const bool spread = p.sample_buf != nullptr;
  // no more CTAs than there are rays to start at once: a slot takes one SAMPLE when samples are spread, one pixel
  // otherwise (capping by pixels alone left 26 SMs idle on a 125 K-pixel shard at 64 spp with 32 warps per CTA)
  const long long rays = items * (spread ? (long long)p.spp : 1ll);
  const long long useful = (rays + 32 * k * lc.wq_warps - 1) / (32 * k * lc.wq_warps);
  if (ctas > useful) ctas = useful;
  cudaError_t e = cudaSuccess;
#define RAYB200_WQ(KK, SP, PK, A, S)                                                                                        \
  do {                                                                                                                      \
    e = opt_in_dynamic_smem<render_warpqueue_kernel<KK, SP, PK, A, S>>(lc.max_dynamic_smem);                                \
    if (e == cudaSuccess)                                                                                                   \
      render_warpqueue_kernel<KK, SP, PK, A, S><<<(unsigned)ctas, wthreads, wsmem, stream>>>(p, ncap, lc.wq_packet, lc.wq_refill, nlow); \
  } while (0)
#define RAYB200_WQ2(KK, SP, PK)                                                           \
  do {                                                                                    \
    if (all_nodes && sph) RAYB200_WQ(KK, SP, PK, true, true);                             \
    else if (all_nodes) RAYB200_WQ(KK, SP, PK, true, false);                              \
    else if (sph) RAYB200_WQ(KK, SP, PK, false, true);                                    \
    else RAYB200_WQ(KK, SP, PK, false, false);                                            \
  } while (0)
#define RAYB200_WQ3(KK, SP)                                                               \
  do {                                                                                    \
    if (packet) RAYB200_WQ2(KK, SP, true); else RAYB200_WQ2(KK, SP, false);               \
  } while (0)
  if (k == 1) { if (spread) RAYB200_WQ3(1, true); else RAYB200_WQ3(1, false); }
  else { if (spread) RAYB200_WQ3(2, true); else RAYB200_WQ3(2, false); }
#undef RAYB200_WQ3
#undef RAYB200_WQ2
#undef RAYB200_WQ
  if (e == cudaSuccess) (*launches)++;
  return e;
}

// Loads (and opts in to shared memory) the variants the default plan uses - K = 1 / 2, samples spread or not, and the three
// staging cases: whole scene staged, tree partly staged with / without the packet walk - so that the first timed frame of
// main.c does not pay CUDA's lazy function loading (measured: +10 ms on the first frame, +1 ms on main.c's 10-run
// average).  12 of the 32 instantiations, ~70 ms once per process instead of ~300 ms for all kernels of all builds.
cudaError_t preload_default_kernels(int max_dynamic_smem) {
  cudaError_t e = cudaSuccess;
#define RAYB200_PRE(KK, SP, PK, A, S) \
  if (e == cudaSuccess) e = opt_in_dynamic_smem<render_warpqueue_kernel<KK, SP, PK, A, S>>(max_dynamic_smem)
#define RAYB200_PRE3(KK, SP) \
  RAYB200_PRE(KK, SP, false, true, true); RAYB200_PRE(KK, SP, true, false, false); RAYB200_PRE(KK, SP, false, false, false)
  RAYB200_PRE3(1, false); RAYB200_PRE3(1, true); RAYB200_PRE3(2, false); RAYB200_PRE3(2, true);
#undef RAYB200_PRE3
#undef RAYB200_PRE
  return e;
}

size_t tile_order_sort_bytes(int64_t local_tiles) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const int32_t *)nullptr,
                                  (int32_t *)nullptr, (int)local_tiles, 0, 32);
  return bytes;
}

void launch_tile_order(const RenderParams &p, const TileOrderBuffers &b, cudaStream_t stream, int64_t *launches) {
  const long long threads = (long long)p.local_tiles * p.probes_per_tile;
  tile_probe_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(p, b.keys, b.ids);
  size_t bytes = b.sort_tmp_bytes;
  cub::DeviceRadixSort::SortPairs(b.sort_tmp, bytes, b.keys, b.keys_sorted, b.ids, b.order, (int)p.local_tiles, 0,
                                  tile_order_key_bits(p.n_chunks), stream);
  if (launches) *launches += 2;  // the cub sort is counted as one
}

void launch_tile_order_from_cost(const RenderParams &p, const uint32_t *cost, int32_t long_path, const TileOrderBuffers &b,
                                 cudaStream_t stream, int64_t *launches) {
  tile_cost_keys_kernel<<<(unsigned)((p.local_tiles + 255) / 256), 256, 0, stream>>>(p, cost, long_path, b.keys, b.ids);
  size_t bytes = b.sort_tmp_bytes;
  cub::DeviceRadixSort::SortPairs(b.sort_tmp, bytes, b.keys, b.keys_sorted, b.ids, b.order, (int)p.local_tiles, 0, 32, stream);
  if (launches) *launches += 2;
}

void launch_flag_wait(uint32_t *flag, uint32_t value, long long timeout_ns, unsigned long long *timeouts, cudaStream_t stream) {
  flag_wait_kernel<<<1, 1, 0, stream>>>(flag, value, timeout_ns, timeouts);
}
void launch_flag_set(uint32_t *flag, uint32_t value, cudaStream_t stream) { flag_set_kernel<<<1, 1, 0, stream>>>(flag, value); }
void launch_flag_bump(uint32_t *flag, cudaStream_t stream) { flag_bump_kernel<<<1, 1, 0, stream>>>(flag); }

void launch_count_work(const RenderParams &p, cudaStream_t stream, int64_t *launches) {
  const long long items = p.local_tiles * kTilePixels;
  if (items <= 0) return;
  const int threads = 128;
  const unsigned blocks = (unsigned)((items + threads - 1) / threads);
  render_mega_kernel<true><<<blocks, threads, 0, stream>>>(p);
  (*launches)++;
}

void launch_detile(const int32_t *gathered, int32_t *out, int64_t H, int64_t W, int32_t world, int64_t tiles_padded,
                   cudaStream_t stream, int64_t *launches) {
  const long long n = (long long)H * W;
  if (n <= 0) return;
  const int threads = 256;
  const int tiles_x = (int)((W + kTileW - 1) / kTileW);
  detile_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(gathered, out, (int)H, (int)W, world,
                                                                                 tiles_padded, tiles_x);
  (*launches)++;
}

}  // namespace rayb200
