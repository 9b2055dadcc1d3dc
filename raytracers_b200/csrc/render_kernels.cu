// sm_100a render kernels: the reference's per-pixel hot path (ray.fut:126-169 over bvh.fut:61-84)
// re-designed for B200.  Compiled with -fmad=false; see device_math.cuh for the bit-exactness rules.
//
// Traversal.  The reference walks the Karras tree stacklessly (parent pointers, bvh.fut:61-84) and
// tests every box against the ORIGINAL ray interval (0, 1e9) (ray.fut:77), so the set of leaves it
// applies `closest_hit` to is exactly { leaf : every ancestor's box passes aabb_hit } — independent
// of traversal order — and the fold result is the leaf with the smallest accepted t, lowest leaf index
// on ties (strict `<` at ray.fut:40 with a shrinking t_max, leaves folded in ascending order).  We
// visit the same set with a left-first stack DFS over the BVH2C layout (scene_host.h): one node step
// tests both children's boxes, ~half the dependent steps of the reference loop and no re-visits.
// See find_closest for how leaf tests are decoupled from the walk and how ties are broken.
#include "render_params.h"
#include <cub/cub.cuh>
#include "device_math.cuh"

#include <cstdio>
#include <type_traits>

namespace rayb200 {

namespace {

constexpr int kDone = (int)0x80000000;  // traversal sentinel: neither an inner index (>= 0) nor a leaf (~i, i < 2^30)
constexpr unsigned kFullMask = 0xffffffffu;

struct WorkCounters {
  unsigned long long segments = 0, node_steps = 0, box_tests = 0, leaf_tests = 0;
};

// ------------------------------------------------------------------ scene access policies
struct GlobalScene {  // everything through the read-only path (L1/L2)
  const float4 *nodes, *geom;
  __device__ __forceinline__ void node(int cur, float4 &q0, float4 &q1, float4 &q2, float4 &q3) const {
    const float4 *p = nodes + 4 * (size_t)cur;
    q0 = __ldg(p); q1 = __ldg(p + 1); q2 = __ldg(p + 2); q3 = __ldg(p + 3);
  }
  __device__ __forceinline__ float4 sphere(int i) const { return __ldg(geom + i); }
};

template <bool kAllNodes, bool kSpheres>
struct StagedScene {  // top of the tree (BFS prefix) + optionally all spheres in shared memory
  const float4 *nodes, *geom;
  const float4 *s_nodes, *s_geom;
  int smem_nodes;
  __device__ __forceinline__ void node(int cur, float4 &q0, float4 &q1, float4 &q2, float4 &q3) const {
    if (kAllNodes || cur < smem_nodes) {  // component-major in shared memory: 16-B stride -> all 8 bank groups in play
      const float4 *p = s_nodes + cur;
      q0 = p[0]; q1 = p[smem_nodes]; q2 = p[2 * smem_nodes]; q3 = p[3 * smem_nodes];
    } else {
      const float4 *p = nodes + 4 * (size_t)cur;
      q0 = __ldg(p); q1 = __ldg(p + 1); q2 = __ldg(p + 2); q3 = __ldg(p + 3);
    }
  }
  __device__ __forceinline__ float4 sphere(int i) const { return kSpheres ? s_geom[i] : __ldg(geom + i); }
};

// ------------------------------------------------------------------ objs_hit, first half (ray.fut:76-82)
// bvh_fold contains closest_hit (-1, 1e9): returns the winning leaf (or -1) and its t.
//
// Because the reference never prunes by the running closest t, node traversal and sphere tests are
// independent: the walk only *collects* the leaves it reaches (a leaf child is recorded by its
// parent's node step, no extra iteration) and the sphere tests run afterwards in a tight loop.  This
// keeps a warp's lanes in the same loop body instead of ping-ponging between "descend" and "test
// leaf".  Collected leaves are not in ascending order any more, so the reference's tie-break (strict
// `<` while folding leaves in ascending index order = lowest index wins an exact t tie) is applied
// explicitly.  sphere_t is evaluated against the ORIGINAL t_max = 1e9: the value sphere_hit returns
// does not depend on the shrinking t_max, only whether it is accepted does (root2 >= root1, so when
// root1 is rejected for being >= t_max, root2 is too).
constexpr int kLeafBuf = 16;

template <bool kCount, class Scene>
__device__ __forceinline__ void test_leaves(const Scene &sc, const int *leaves, int &nl, const Ray &r, const RayInv &q,
                                            int &best_j, float &best_t, WorkCounters &wc) {
  for (int k = 0; k < nl; k++) {
    const int li = leaves[k];
    const float4 g = sc.sphere(li);
    if (kCount) wc.leaf_tests++;
    const float t = sphere_t(g.x, g.y, g.z, g.w, r, q.a, 0.1f, 1000000000.0f);  // closest_hit, ray.fut:78-81
    if (t >= 0.0f && (t < best_t || (t == best_t && li < best_j))) { best_t = t; best_j = li; }
  }
  nl = 0;
}

template <bool kCount, class Scene>
__device__ __forceinline__ void find_closest(const Scene &sc, const float *root_box, const Ray &r, const RayInv &q,
                                             int &best_j, float &best_t, WorkCounters &wc) {
  best_j = -1;
  best_t = 1000000000.0f;
  if (kCount) { wc.segments++; wc.box_tests++; }
  if (!box_hit(root_box[0], root_box[1], root_box[2], root_box[3], root_box[4], root_box[5], r, q)) return;
  int stack[kStackSize + 1];
  int leaves[kLeafBuf];
  int sp = 1, nl = 0;
  stack[0] = kDone;  // popping the sentinel ends the walk
  int cur = 0;
  while (cur != kDone) {
    if (nl > kLeafBuf - 2) test_leaves<kCount>(sc, leaves, nl, r, q, best_j, best_t, wc);
    float4 q0, q1, q2, q3;
    sc.node(cur, q0, q1, q2, q3);
    const int lptr = __float_as_int(q0.w), rptr = __float_as_int(q1.w);
    const bool hl = box_hit(q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, r, q);
    const bool hr = box_hit(q2.x, q2.y, q2.z, q3.x, q3.y, q3.z, r, q);
    const bool l_leaf = lptr < 0, r_leaf = rptr < 0;
    if (kCount) { wc.node_steps++; wc.box_tests += !l_leaf + !r_leaf; }
    // a leaf child has no box in the reference (bvh.fut:84): it is always visited -> record it
    if (l_leaf) leaves[nl] = ~lptr;
    nl += l_leaf;
    if (r_leaf) leaves[nl] = ~rptr;
    nl += r_leaf;
    // inner children whose box is hit are walked: left first, right deferred on the stack
    const bool tl = hl && !l_leaf, tr = hr && !r_leaf;
    if (tl && tr) stack[sp] = rptr;
    sp += (tl && tr);
    int nxt = tl ? lptr : rptr;
    if (!(tl || tr)) nxt = stack[--sp];
    cur = nxt;
  }
  test_leaves<kCount>(sc, leaves, nl, r, q, best_j, best_t, wc);
}

// ------------------------------------------------------------------ one ray_colour iteration (ray.fut:130-148)
// The part of a ray_colour iteration after the closest-hit search: given the fold result (j, tb),
// re-intersect, scatter or shade the sky.  `a` = dot r.d r.d.  Returns true if the path continues
// (r/light/depth updated), false if it ended with `colour` set.
template <class Scene>
__device__ __forceinline__ bool shade_segment(const Scene &sc, const RenderParams &P, Ray &r, const float a, const int j,
                                              const float tb, V3 &light, int &depth, V3 &colour) {
  if (j >= 0) {
    // objs_hit, second half (ray.fut:83-85): re-intersect the winner with t_min = 0, t_max = t_best + 1
    const float4 g = sc.sphere(j);
    const float t = sphere_t(g.x, g.y, g.z, g.w, r, a, 0.0f, tb + 1.0f);
    if (t >= 0.0f) {
      const V3 c = v3(g.x, g.y, g.z);
      const V3 p = vadd(r.o, vscale(t, r.d));                       // point_at_param, ray.fut:14-15
      const V3 n = vscale(1.0f / g.w, vsub(p, c));                  // ray.fut:42-43
      // scatter (ray.fut:119-124): reflect (normalise r.dir) hit.normal; norm r.dir = sqrt(dot d d) = sqrt(a)
      const V3 unit = vscale(1.0f / sqrtf(a), r.d);
      const V3 refl = vsub(unit, vscale(2.0f * vdot(unit, n), n));  // ray.fut:116-117
      if (vdot(refl, n) > 0.0f) {
        const float4 col = __ldg(P.colour + j);
        r.o = p;
        r.d = refl;
        light = vmul(light, v3(col.x, col.y, col.z));               // ray.fut:135
        depth = depth + 1;
        if (depth < kMaxDepth) return true;
        colour = v3(0.0f, 0.0f, 0.0f);                              // loop exit with colour = light*0 (ray.fut:136)
        return false;
      }
      colour = v3(0.0f, 0.0f, 0.0f);                                // ray.fut:137-140
      return false;
    }
  }
  // miss: sky gradient (ray.fut:141-148)
  const V3 unit = vscale(1.0f / sqrtf(a), r.d);
  const float t = 0.5f * (unit.y + 1.0f);
  const float w1 = 1.0f - t;
  const V3 sky = v3(w1 * 1.0f + t * 0.5f, w1 * 1.0f + t * 0.7f, w1 * 1.0f + t * 1.0f);
  colour = vmul(light, sky);
  return false;
}

// One whole ray_colour iteration for a lane-owned path.  `depth` counts objs_hit calls so far (ray.fut:129).
template <bool kCount, class Scene>
__device__ __forceinline__ bool advance_path(const Scene &sc, const RenderParams &P, Ray &r, V3 &light, int &depth,
                                             V3 &colour, WorkCounters &wc) {
  const RayInv q = ray_invariants(r);
  int j;
  float tb;
  find_closest<kCount>(sc, P.root_box, r, q, j, tb, wc);
  return shade_segment(sc, P, r, q.a, j, tb, light, depth, colour);
}

// get_ray for sample s of pixel (row j, column i): ray.fut:150-154 with pixel j i -> trace_ray (height-j) i
// (ray.fut:167-168) and the spp extension of ray_b200.h (offset (0,0) at s = 0).
__device__ __forceinline__ Ray primary_ray(const RenderParams &P, int i, int j, int s) {
  float u, v;
  if (P.spp == 1) {
    u = (float)i / (float)P.W;
    v = (float)(P.H - j) / (float)P.H;
  } else {
    u = ((float)i + P.offsets[2 * s]) / (float)P.W;
    v = ((float)(P.H - j) + P.offsets[2 * s + 1]) / (float)P.H;
  }
  const float *c = P.cam;
  Ray r;
  r.o = v3(c[0], c[1], c[2]);
  // llc + s*horizontal + t*vertical - origin (ray.fut:111-113), per component, left to right
  r.d = v3(((c[3] + u * c[6]) + v * c[9]) - c[0], ((c[4] + u * c[7]) + v * c[10]) - c[1],
           ((c[5] + u * c[8]) + v * c[11]) - c[2]);
  return r;
}

// item k -> pixel.  Local tile lt = k >> 5 is global tile lt*world + rank; lane position k & 31 inside the 8x4 tile.
__device__ __forceinline__ bool item_pixel(const RenderParams &P, int k, int &i, int &j) {
  const long long t = (long long)(k >> 5) * P.world + P.rank;
  const int sub = k & 31;
  const int ty = (int)(t / P.tiles_x), tx = (int)(t - (long long)ty * P.tiles_x);
  i = tx * kTileW + (sub & (kTileW - 1));
  j = ty * kTileH + (sub >> 3);
  return t < P.n_tiles && i < P.W && j < P.H;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// claim index (value of the global work cursor) -> item index: chunks of 64 tiles in stride-permuted order
__device__ __forceinline__ int claim_to_item(const RenderParams &P, int c) {
  if (P.tile_order) {  // heavy-first order from the probe pass
    const int lt = c >> 5;
    return lt < P.local_tiles ? (__ldg(P.tile_order + lt) << 5) | (c & 31) : (int)(P.local_tiles << 5);
  }
  const int chunk = c >> 11;
  const int perm = (int)(((long long)chunk * P.chunk_stride) % P.n_chunks);
  return (perm << 11) | (c & 2047);
}

__device__ __forceinline__ void write_pixel(const RenderParams &P, int k, int i, int j, V3 sum) {
  const V3 col = (P.spp == 1) ? sum : vscale(P.inv_spp, sum);
  const int pix = pack_pixel(col);
  if (P.tile_major) P.out_pix[k] = pix;
  else P.out_pix[(size_t)j * P.W + i] = pix;
  if (P.out_rgb) {
    float *q = P.out_rgb + 3 * ((size_t)j * P.W + i);
    q[0] = col.x; q[1] = col.y; q[2] = col.z;
  }
}

__device__ __forceinline__ void flush_counters(const RenderParams &P, WorkCounters &wc) {
  for (int o = 16; o > 0; o >>= 1) {
    wc.segments += __shfl_down_sync(kFullMask, wc.segments, o);
    wc.node_steps += __shfl_down_sync(kFullMask, wc.node_steps, o);
    wc.box_tests += __shfl_down_sync(kFullMask, wc.box_tests, o);
    wc.leaf_tests += __shfl_down_sync(kFullMask, wc.leaf_tests, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(P.counters + 0, wc.segments);
    atomicAdd(P.counters + 1, wc.node_steps);
    atomicAdd(P.counters + 2, wc.box_tests);
    atomicAdd(P.counters + 3, wc.leaf_tests);
  }
}

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

// ====================================================================================== TMA staging helpers
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk async copy global -> shared through the TMA unit (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// Stages the BFS prefix of the node array (and optionally all sphere records) into shared memory.
// One elected thread arms the mbarrier with the byte count and issues the bulk copies; everyone waits on it.
__device__ __forceinline__ void stage_scene(const RenderParams &P, unsigned char *smem_raw, const float4 *&s_nodes,
                                            const float4 *&s_geom) {
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  float4 *nodes_dst = reinterpret_cast<float4 *>(smem_raw + 128);
  float4 *geom_dst = nodes_dst + 4 * (size_t)P.smem_nodes;
  const uint32_t node_bytes = (uint32_t)P.smem_nodes * 64u, geom_bytes = (uint32_t)P.smem_spheres * 16u;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, node_bytes + geom_bytes);
    constexpr uint32_t kChunk = 32768;
    // component-major: the first smem_nodes entries of each of the four component arrays
    const uint32_t comp_bytes = (uint32_t)P.smem_nodes * 16u;
    for (int c = 0; c < 4; c++)
      for (uint32_t off = 0; off < comp_bytes; off += kChunk)
        tma_bulk_g2s(reinterpret_cast<unsigned char *>(nodes_dst + (size_t)c * P.smem_nodes) + off,
                     reinterpret_cast<const unsigned char *>(P.nodes_soa + (size_t)c * P.n_inner) + off,
                     min(kChunk, comp_bytes - off), bar);
    for (uint32_t off = 0; off < geom_bytes; off += kChunk)
      tma_bulk_g2s(reinterpret_cast<unsigned char *>(geom_dst) + off, reinterpret_cast<const unsigned char *>(P.geom) + off,
                   min(kChunk, geom_bytes - off), bar);
  }
  mbar_wait(bar, 0);
  s_nodes = nodes_dst;
  s_geom = geom_dst;
}

// ====================================================================================== K1: persistent + refill
// Persistent CTAs (grid = SMs x resident CTAs).  Every lane owns one pixel at a time and runs its
// samples and bounces; whenever enough lanes of a warp are idle the warp claims new pixels from a
// global cursor with one warp-aggregated atomicAdd (ballot + popc), so irreg's empty-sky rows and
// rgbbox's 50-bounce tails never leave a warp mostly empty.  Samples of one pixel are summed in
// sample order in a register, which is what the spp extension requires.
template <bool kAllNodes, bool kSpheres>
__global__ void __launch_bounds__(256, 2) render_persistent_kernel(const __grid_constant__ RenderParams P,
                                                                    const int refill_min) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const float4 *s_nodes, *s_geom;
  stage_scene(P, smem_raw, s_nodes, s_geom);
  const StagedScene<kAllNodes, kSpheres> sc{P.nodes, P.geom, s_nodes, s_geom, P.smem_nodes};

  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int total = (int)(P.local_tiles * kTilePixels);
  const int total_claims = P.n_chunks << 11;
  WorkCounters wc;

  int item = -1, pi = 0, pj = 0, s = 0, depth = 0;
  Ray r;
  V3 light, sum;
  bool exhausted = false;  // warp-uniform: the cursor has run past the end
  r.o = r.d = light = sum = v3(0.0f, 0.0f, 0.0f);

  for (;;) {
    unsigned idle = __ballot_sync(kFullMask, item < 0);
    if (idle && !exhausted && (__popc(idle) >= refill_min || idle == kFullMask)) {
      const int cnt = __popc(idle);
      const int leader = __ffs(idle) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(P.work_cursor, cnt);
      base = __shfl_sync(kFullMask, base, leader);
      if (item < 0) {
        const int c = base + __popc(idle & lt_mask);
        const int k = c < total_claims ? claim_to_item(P, c) : total;
        if (k < total) {
          if (item_pixel(P, k, pi, pj)) {
            item = k;
            s = 0;
            depth = 0;
            r = primary_ray(P, pi, pj, 0);
            light = v3(1.0f, 1.0f, 1.0f);
          } else if (P.tile_major) {
            P.out_pix[k] = 0;
          }
        }
      }
      exhausted = base + cnt >= total_claims;
      idle = __ballot_sync(kFullMask, item < 0);
    }
    if (idle == kFullMask) {
      if (exhausted) break;
      continue;  // every claimed item was a padding pixel: claim again
    }
    if (item >= 0) {
      V3 colour;
      if (!advance_path<false>(sc, P, r, light, depth, colour, wc)) {
        sum = (s == 0) ? colour : vadd(sum, colour);
        s++;
        if (s < P.spp) {
          depth = 0;
          r = primary_ray(P, pi, pj, s);
          light = v3(1.0f, 1.0f, 1.0f);
        } else {
          write_pixel(P, item, pi, pj, sum);
          item = -1;
        }
      }
    }
  }
}

// ====================================================================================== K2: wavefront
// The north-star design: ONE persistent-threads kernel launch per bounce.  Warps claim batches of 32
// rays from the bounce's global queue with an atomic cursor, trace one segment (same traversal as K1,
// BVH staged by TMA), shade, and append the survivors to the next bounce's queue with warp-vote
// compaction (ballot + popc + one atomicAdd per warp), so every bounce runs on densely packed warps
// whatever the image-space distribution of live paths is.  Bounce 0 generates its rays instead of
// reading them; terminated paths write their pixel (spp == 1) or add into an in-order accumulator.
// From bounce `tail_from` on, the handful of surviving rays are run to completion inside one launch
// instead of paying ~40 more near-empty launches.
template <bool kAllNodes, bool kSpheres>
__global__ void __launch_bounds__(256, 2) wavefront_bounce_kernel(const __grid_constant__ RenderParams P,
                                                                   const __grid_constant__ WavefrontBuffers B,
                                                                   const int bounce, const int sample,
                                                                   const int run_to_end) {
  const long long total64 = P.local_tiles * kTilePixels;
  const int n_in = bounce == 0 ? (int)total64 : B.qlen[bounce];
  if ((long long)blockIdx.x * 32 >= n_in) return;  // nothing for this CTA: skip the staging too
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const float4 *s_nodes, *s_geom;
  stage_scene(P, smem_raw, s_nodes, s_geom);
  const StagedScene<kAllNodes, kSpheres> sc{P.nodes, P.geom, s_nodes, s_geom, P.smem_nodes};
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int qi = bounce & 1, qo = qi ^ 1;
  const bool last_sample = sample == P.spp - 1;
  WorkCounters wc;
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(B.cursor + bounce, 32);
    base = __shfl_sync(kFullMask, base, 0);
    if (base >= n_in) break;
    const int idx = base + lane;
    bool active = idx < n_in;
    Ray r;
    V3 light = v3(1.0f, 1.0f, 1.0f);
    int pid = idx, depth = bounce;
    r.o = r.d = v3(0.0f, 0.0f, 0.0f);
    if (active) {
      if (bounce == 0) {
        int i, j;
        active = item_pixel(P, idx, i, j);
        if (active) r = primary_ray(P, i, j, sample);
        else if (P.tile_major && last_sample) P.out_pix[idx] = 0;
      } else {
        const float4 a = B.ray_o[qi][idx], d = B.ray_d[qi][idx], l = B.light[qi][idx];
        r.o = v3(a.x, a.y, a.z);
        r.d = v3(d.x, d.y, d.z);
        light = v3(l.x, l.y, l.z);
        pid = __float_as_int(a.w);
      }
    }
    bool cont = false;
    if (active) {
      V3 colour;
      cont = advance_path<false>(sc, P, r, light, depth, colour, wc);
      if (run_to_end)
        while (cont) cont = advance_path<false>(sc, P, r, light, depth, colour, wc);
      if (!cont) {  // path ended: this sample's colour goes to its pixel, in sample order
        int i, j;
        item_pixel(P, pid, i, j);
        if (P.spp == 1) {
          write_pixel(P, pid, i, j, colour);
        } else {
          V3 sum = colour;
          if (sample > 0) {
            const float4 acc = B.accum[pid];
            sum = vadd(v3(acc.x, acc.y, acc.z), colour);
          }
          if (last_sample) write_pixel(P, pid, i, j, sum);
          else B.accum[pid] = make_float4(sum.x, sum.y, sum.z, 0.0f);
        }
      }
    }
    // warp-vote compaction of the survivors into the next bounce's queue
    const unsigned alive = __ballot_sync(kFullMask, cont);
    if (alive) {
      const int leader = __ffs(alive) - 1;
      int obase = 0;
      if (lane == leader) obase = atomicAdd(B.qlen + bounce + 1, __popc(alive));
      obase = __shfl_sync(kFullMask, obase, leader);
      if (cont) {
        const int o = obase + __popc(alive & lt_mask);
        B.ray_o[qo][o] = make_float4(r.o.x, r.o.y, r.o.z, __int_as_float(pid));
        B.ray_d[qo][o] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
        B.light[qo][o] = make_float4(light.x, light.y, light.z, 0.0f);
      }
    }
  }
}

// ====================================================================================== K3: warp work-queue
// Dense traversal.  K1/K2 bind a lane to a ray for a whole segment, so a warp's SIMT efficiency is
// mean/max of its lanes' traversal lengths (ncu: 6.5 of 32 lanes active on rgbbox).  The reference's
// fold is order-free and never prunes by the running closest t, so a segment's traversal is just a
// SET of independent (ray, node) box-test items and (ray, leaf) sphere-test items.  K3 therefore
// binds lanes to ITEMS, not rays: every warp keeps R = 32*K rays in flight, their traversal items on
// a warp-private LIFO in shared memory, and each iteration pops 32 items — whatever rays they belong
// to — so every lane does one node step (two box tests) or one sphere test.  Children are pushed back
// with ballot/popc compaction; a sphere hit is folded into its ray's 64-bit (t, leaf index) word with
// atomicMin, which is exactly the reference's "smallest t, lowest leaf index on ties".  When the
// stacks run dry every ray of the round has its closest hit; the owner lanes shade, bounce, and idle
// slots are refilled, and the next round starts.
//
// Stack bound: items are pushed in reverse lane order, which keeps the LIFO sorted by tree depth
// (deepest on top); then at most 64 items of any depth are live at once (children of one 32-item
// batch), so R + 64*(max_depth+1) entries always suffice.  Leaf items are drained whenever 32 are
// available, so that stack never holds more than 31 + 64.
//
// Work distribution.  spp == 1 (kSpread = false): a slot owns a pixel (and, for robustness, all its
// samples), claimed from the global cursor with one warp-aggregated atomicAdd.  spp > 1 (kSpread):
// a warp opens up to kRing pixels at a time and hands their samples, in order, to whichever slots are
// idle — consecutive lanes trace consecutive samples of the same pixel (coherent), the scheduling
// grain is one sample instead of one pixel x spp (no long tails, scales to many GPUs), and every
// finished sample's colour is parked in an L2-resident buffer so that the pixel is summed in SAMPLE
// ORDER when its last sample lands, which keeps the result bit-identical to the sequential definition.
constexpr int kSlotShift = 26;                 // item = slot << 26 | index  (index < 2^26; R <= 64 slots)
constexpr uint32_t kIndexMask = (1u << kSlotShift) - 1u;
constexpr unsigned long long kNoHit = ~0ull;

template <int K, bool kSpread, bool kPacket, bool kAllNodes, bool kSpheres>
__global__ void __launch_bounds__(kWqMaxThreads, 1) render_warpqueue_kernel(const __grid_constant__ RenderParams P, const int ncap,
                                                                   const int packet_min, const int refill_min) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (P.warp_trace && threadIdx.x == 0) atomicMin(P.warp_trace, global_timer_ns());
  const float4 *s_nodes, *s_geom;
  stage_scene(P, smem_raw, s_nodes, s_geom);
  const StagedScene<kAllNodes, kSpheres> sc{P.nodes, P.geom, s_nodes, s_geom, P.smem_nodes};

  constexpr int R = 32 * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned gt_mask = lane == 31 ? 0u : ~((2u << lane) - 1u);
  unsigned char *wbase = smem_raw + ((staging_bytes(P) + 127) & ~(size_t)127) + (size_t)warp * wq_warp_bytes(K, ncap, kPacket);
  float4 *ray_o = reinterpret_cast<float4 *>(wbase);   // {o.xyz, a = dot d d}
  float4 *ray_i = ray_o + R;                           // {1/d.xyz, 0}
  float4 *ray_d = ray_i + R;                           // {d.xyz, 0}
  float4 *p_light = ray_d + R;                         // {light.rgb, bits(depth)}   owner lane only
  float4 *p_sum = p_light + R;                         // {sum.rgb, bits(sample)} / spread: w = bits(ring << 16 | sample)
  unsigned long long *best = reinterpret_cast<unsigned long long *>(p_sum + R);  // (bits(t) << 32 | leaf) min-folded
  int *p_item = reinterpret_cast<int *>(best + R);     // work item (pixel) of the slot, -1 = idle
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
  int ntop = 0, ltop = 0;                      // warp-uniform stack heights
  int open_seq = 0, disp_seq = 0, disp_s = 0;  // spread dispenser (warp-uniform): pixels opened / next sample to hand out
  float4 *cbuf = nullptr;                      // spread: [kWqRing][spp] finished-sample colours of this warp
  if (kSpread) cbuf = P.sample_buf + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * kWqRing * (size_t)spp;

  // A path in `slot` has ended with `colour`.
  auto finish_path = [&](const int slot, const V3 colour) {
    if (kSpread) {
      const int ms = __float_as_int(p_sum[slot].w);
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
    if (lane < kWqRing && ring_done[lane] == spp) {
      const int item = ring_item[lane];
      int pi, pj;
      if (item_pixel(P, item, pi, pj)) {
        const float4 *c = cbuf + (size_t)lane * spp;
        const float4 c0 = __ldcg(c);
        V3 sum = v3(c0.x, c0.y, c0.z);
        for (int s = 1; s < spp; s++) {
          const float4 cs = __ldcg(c + s);
          sum = vadd(sum, v3(cs.x, cs.y, cs.z));
        }
        write_pixel(P, item, pi, pj, sum);
      } else if (P.tile_major) {
        P.out_pix[item] = 0;
      }
      ring_done[lane] = -1;
    }
    __syncwarp();
  };

  for (;;) {
    unsigned trav = 0;  // bit k: slot lane+32k has a traversal in flight this round
    unsigned gomask[K];  // warp-uniform: lanes whose slot lane+32k starts a traversal this round
#pragma unroll
    for (int k = 0; k < K; k++) gomask[k] = 0u;
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
#pragma unroll
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
#pragma unroll
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
                p_sum[slot] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float((m << 16) | s));
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
#pragma unroll
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
              best[slot] = kNoHit;
              go = true;
              break;
            }
            const float4 pl = p_light[slot];  // miss (ray.fut:141-148)
            V3 light = v3(pl.x, pl.y, pl.z), colour;
            int depth = __float_as_int(pl.w);
            shade_segment(sc, P, r, q.a, -1, 0.0f, light, depth, colour);
            finish_path(slot, colour);
          }
        }
        const unsigned m = __ballot_sync(kFullMask, go);
        if (go) trav |= 1u << k;
        gomask[k] |= m;  // the traversal of these rays starts at the root after the passes
      }
      // another pass only helps if some slot is idle and there is still work to hand out
      bool idle_left = false;
#pragma unroll
      for (int k = 0; k < K; k++) idle_left |= p_item[lane + 32 * k] < 0;
      const bool more = kSpread ? (!exhausted || (open_seq - disp_seq) * spp - disp_s > 0) : !exhausted;
      if (!more || !__any_sync(kFullMask, idle_left)) break;
    }
    __syncwarp();
    unsigned any_go = 0u;
#pragma unroll
    for (int k = 0; k < K; k++) any_go |= gomask[k];
    if (any_go == 0u) {
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
    // predicate registers) used while the stacks are deep enough, and a partial version for the drain.
    auto leaf_batch = [&](auto full_tag) {
      // closest_hit (ray.fut:78-81) for up to 32 (ray, sphere) pairs
      constexpr bool kFull = decltype(full_tag)::value;
      const int n = kFull ? 32 : ltop;
      if (kFull || lane < n) {
        const uint32_t it = lstk[ltop - 1 - lane];
        const int slot = (int)(it >> kSlotShift), li = (int)(it & kIndexMask);
        const float4 ro = ray_o[slot], rd = ray_d[slot];
        const float4 g = sc.sphere(li);
        Ray r;
        r.o = v3(ro.x, ro.y, ro.z);
        r.d = v3(rd.x, rd.y, rd.z);
        const float t = sphere_t(g.x, g.y, g.z, g.w, r, ro.w, 0.1f, 1000000000.0f);
        if (t >= 0.0f) atomicMin(best + slot, ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)li);
      }
      ltop -= n;
    };
    auto node_batch = [&](auto full_tag, const int n_part) {
      // one BVH2C node step (both children's boxes) for up to 32 (ray, node) pairs
      constexpr bool kFull = decltype(full_tag)::value;
      const int n = kFull ? 32 : n_part;
      bool pl_node = false, pr_node = false, pl_leaf = false, pr_leaf = false;
      uint32_t tag = 0;
      int lptr = 0, rptr = 0;
      if (kFull || lane < n) {
        const uint32_t it = nstk[ntop - 1 - lane];
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
      __syncwarp();  // all pops have been read before anything is pushed over them
      ntop -= n;
      const unsigned bl = __ballot_sync(kFullMask, pl_node), br = __ballot_sync(kFullMask, pr_node);
      const unsigned cl = __ballot_sync(kFullMask, pl_leaf), cr = __ballot_sync(kFullMask, pr_leaf);
      // reverse lane order: lane 0 popped the top (deepest) item, its children go back on top
      const int nb = ntop + __popc(bl & gt_mask) + __popc(br & gt_mask);
      if (pr_node) nstk[nb] = tag | (uint32_t)rptr;
      if (pl_node) nstk[nb + (pr_node ? 1 : 0)] = tag | (uint32_t)lptr;
      ntop += __popc(bl) + __popc(br);
      const int lb = ltop + __popc(cl & lt_mask) + __popc(cr & lt_mask);
      if (pl_leaf) lstk[lb] = tag | (uint32_t)(~lptr);
      if (pr_leaf) lstk[lb + (pl_leaf ? 1 : 0)] = tag | (uint32_t)(~rptr);
      ltop += __popc(cl) + __popc(cr);
    };
    using full_t = std::integral_constant<bool, true>;
    using part_t = std::integral_constant<bool, false>;
    // Runs the item queues dry.  Without packet spills the depth-sorted LIFO never exceeds the proved bound; with
    // spills (arbitrary depths) a guard keeps it safe for ANY content: once fewer than 96 entries are free, items are
    // taken one at a time, a plain DFS that can add at most (tree depth) < 64 entries before it shrinks again.
    auto drain = [&]() {
      __syncwarp();
      while (ntop > 0 || ltop > 0) {
        const bool tight = ntop + 96 > ncap;
        if (ltop >= 32) leaf_batch(full_t{});
        else if (ntop >= 32 && !tight) node_batch(full_t{}, 32);
        else if (ntop > 0) node_batch(part_t{}, tight ? 1 : ntop);
        else leaf_batch(part_t{});
        __syncwarp();
      }
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
              unsigned long long b = best[slot];
              if (lptr < 0) {
                const float4 g = sc.sphere(~lptr);
                const float t = sphere_t(g.x, g.y, g.z, g.w, r, ro.w, 0.1f, 1000000000.0f);
                const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)(~lptr);
                if (t >= 0.0f && key < b) b = key;
              }
              if (rptr < 0) {
                const float4 g = sc.sphere(~rptr);
                const float t = sphere_t(g.x, g.y, g.z, g.w, r, ro.w, 0.1f, 1000000000.0f);
                const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)(~rptr);
                if (t >= 0.0f && key < b) b = key;
              }
              best[slot] = b;
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
      if (kPacket && packet_min > 0) {
        if constexpr (kPacket) {
          if (gomask[k]) packet_walk(k, gomask[k]);
        }
      } else {
        const bool go = (gomask[k] >> lane) & 1u;
        if (go) nstk[ntop + __popc(gomask[k] & lt_mask)] = (uint32_t)(lane + 32 * k) << kSlotShift;  // (slot, root node 0)
        ntop += __popc(gomask[k]);
      }
    }
    drain();

    // ---------------------------------------------------------------- shade: owner lanes finish the segment
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (!((trav >> k) & 1u)) continue;
      const int slot = lane + 32 * k;
      const float4 ro = ray_o[slot], rd = ray_d[slot], pl = p_light[slot];
      const unsigned long long b = best[slot];
      Ray r;
      r.o = v3(ro.x, ro.y, ro.z);
      r.d = v3(rd.x, rd.y, rd.z);
      V3 light = v3(pl.x, pl.y, pl.z), colour;
      int depth = __float_as_int(pl.w);
      const int j = b == kNoHit ? -1 : (int)(unsigned)(b & 0xffffffffu);
      const float tb = __uint_as_float((unsigned)(b >> 32));
      if (shade_segment(sc, P, r, ro.w, j, tb, light, depth, colour)) {
        ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, 0.0f);
        ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
        p_light[slot] = make_float4(light.x, light.y, light.z, __int_as_float(depth));
      } else {
        finish_path(slot, colour);
      }
    }
    __syncwarp();
  }
}

// ====================================================================================== K4: stream queue
// K3 without rounds — a measured NEGATIVE result, kept (like K2) as an alternative with parity tests: on B200 it is
// ~50 % slower than K3 on every config (profiles/r1_sweep_streamqueue_vs_warpqueue.json); the per-item shared-memory
// atomics on the ray counters (siblings of one ray sit next to each other in the LIFO, so they serialise), the done/free
// lists and the sparser refill batches cost more than the round tails they remove.
// In K3 a round of 32*K rays cannot end before its slowest ray has walked its ~15 dependent node
// steps, so every round has a tail of partial batches (ncu: 25 of 32 lanes active).  K4 keeps the queues permanently
// topped up instead: every ray carries a counter of its outstanding items in shared memory (+children -1 per node item,
// -1 per leaf item); the lane that brings a counter to zero puts the ray on a "done" list; done rays are shaded in dense
// 32-wide batches by whichever lanes are free (all per-ray state lives in shared memory, no owner lanes), and their
// slots — or fresh samples for them — go straight back to the root of the tree while the other rays' items keep the
// node and leaf queues full.  The item LIFO is no longer depth-sorted, so its capacity is protected by the same guard
// as K3's packet spills (one item at a time once fewer than 96 entries are free).
template <int K, bool kSpread, bool kAllNodes, bool kSpheres>
__global__ void __launch_bounds__(kWqMaxThreads, 1) render_streamqueue_kernel(const __grid_constant__ RenderParams P, const int ncap,
                                                                               const int refill_min) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const float4 *s_nodes, *s_geom;
  stage_scene(P, smem_raw, s_nodes, s_geom);
  const StagedScene<kAllNodes, kSpheres> sc{P.nodes, P.geom, s_nodes, s_geom, P.smem_nodes};

  constexpr int R = 32 * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned char *wbase = smem_raw + ((staging_bytes(P) + 127) & ~(size_t)127) + (size_t)warp * sq_warp_bytes(K, ncap);
  float4 *ray_o = reinterpret_cast<float4 *>(wbase);   // {o.xyz, a = dot d d}
  float4 *ray_i = ray_o + R;                           // {1/d.xyz, 0}
  float4 *ray_d = ray_i + R;                           // {d.xyz, 0}
  float4 *p_light = ray_d + R;                         // {light.rgb, bits(depth)}
  float4 *p_sum = p_light + R;                         // {sum.rgb, bits(sample)} / spread: w = bits(ring << 16 | sample)
  unsigned long long *best = reinterpret_cast<unsigned long long *>(p_sum + R);  // (bits(t) << 32 | leaf) min-folded
  int *p_item = reinterpret_cast<int *>(best + R);     // pixel item of the slot, -1 = idle
  int *pending = p_item + R;                           // outstanding traversal items of the slot's ray
  int *ring_item = pending + R;
  int *ring_done = ring_item + kWqRing;
  uint32_t *dstk = reinterpret_cast<uint32_t *>(ring_done + kWqRing);  // slots whose ray has finished its traversal
  uint32_t *fstk = dstk + R;                                           // idle slots
  uint32_t *lstk = fstk + R;
  uint32_t *nstk = lstk + kWqLeafStack;

  const int total = (int)(P.local_tiles * kTilePixels);
  const int total_claims = P.n_chunks << 11;
  const int spp = P.spp;
#pragma unroll
  for (int k = 0; k < K; k++) { p_item[lane + 32 * k] = -1; fstk[lane + 32 * k] = (uint32_t)(lane + 32 * k); }
  if (lane < kWqRing) ring_done[lane] = -1;
  __syncwarp();
  bool exhausted = false;
  int ntop = 0, ltop = 0, dtop = 0, ftop = R;   // warp-uniform stack heights
  int open_seq = 0, disp_seq = 0, disp_s = 0;   // spread dispenser (warp-uniform)
  float4 *cbuf = nullptr;
  if (kSpread) cbuf = P.sample_buf + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * kWqRing * (size_t)spp;

  // A path in `slot` has ended with `colour` (any lane may call this for any slot it is handling).
  auto finish_path = [&](const int slot, const V3 colour) {
    if (kSpread) {
      const int ms = __float_as_int(p_sum[slot].w);
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
  auto finalize_pixels = [&]() {
    __syncwarp();
    if (lane < kWqRing && ring_done[lane] == spp) {
      const int item = ring_item[lane];
      int pi, pj;
      if (item_pixel(P, item, pi, pj)) {
        const float4 *c = cbuf + (size_t)lane * spp;
        const float4 c0 = __ldcg(c);
        V3 sum = v3(c0.x, c0.y, c0.z);
        for (int s = 1; s < spp; s++) {
          const float4 cs = __ldcg(c + s);
          sum = vadd(sum, v3(cs.x, cs.y, cs.z));
        }
        write_pixel(P, item, pi, pj, sum);
      } else if (P.tile_major) {
        P.out_pix[item] = 0;
      }
      ring_done[lane] = -1;
    }
    __syncwarp();
  };
  // The slot has a fresh ray (primary or bounced): root box test; sky rays are finished on the spot (which may hand
  // the slot its next sample).  Returns true if the ray enters the tree (its root item is then pushed by the caller).
  auto start_ray = [&](const int slot) -> bool {
    while (p_item[slot] >= 0) {
      const float4 ro = ray_o[slot], rd = ray_d[slot];
      Ray r;
      r.o = v3(ro.x, ro.y, ro.z);
      r.d = v3(rd.x, rd.y, rd.z);
      const RayInv q = ray_invariants(r);
      if (box_hit(P.root_box[0], P.root_box[1], P.root_box[2], P.root_box[3], P.root_box[4], P.root_box[5], r, q)) {
        ray_o[slot] = make_float4(ro.x, ro.y, ro.z, q.a);
        ray_i[slot] = make_float4(q.ix, q.iy, q.iz, 0.0f);
        best[slot] = kNoHit;
        pending[slot] = 1;
        return true;
      }
      const float4 pl = p_light[slot];  // miss (ray.fut:141-148)
      V3 light = v3(pl.x, pl.y, pl.z), colour;
      int depth = __float_as_int(pl.w);
      shade_segment(sc, P, r, q.a, -1, 0.0f, light, depth, colour);
      finish_path(slot, colour);
    }
    return false;
  };
  // After a refill / shade batch: lanes whose ray entered the tree push its root item, lanes whose slot went idle
  // give it back.  `mine`: this lane handled a slot in the batch.
  auto push_roots_and_idle = [&](const bool mine, const int slot, const bool go) {
    const unsigned gm = __ballot_sync(kFullMask, go);
    if (go) nstk[ntop + __popc(gm & lt_mask)] = (uint32_t)slot << kSlotShift;  // (slot, root node 0)
    ntop += __popc(gm);
    const bool idle = mine && !go && p_item[slot] < 0;
    const unsigned im = __ballot_sync(kFullMask, idle);
    if (idle) fstk[ftop + __popc(im & lt_mask)] = (uint32_t)slot;
    ftop += __popc(im);
  };
  auto avail_samples = [&]() { return (open_seq - disp_seq) * spp - disp_s; };

  auto refill = [&]() {
    if (kSpread) finalize_pixels();
    const int f = ftop < 32 ? ftop : 32;
    const bool mine = lane < f;
    const int slot = mine ? (int)fstk[ftop - 1 - lane] : 0;
    __syncwarp();
    ftop -= f;
    if (!kSpread) {
      int base = 0;
      if (lane == 0) base = atomicAdd(P.work_cursor, f);
      base = __shfl_sync(kFullMask, base, 0);
      if (mine) {
        const int c = base + lane;
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
      exhausted = base + f >= total_claims;
    } else {
      int avail = avail_samples();
      while (!exhausted && avail < f) {  // open more pixels while the ring has room
        const int m = open_seq & (kWqRing - 1);
        if (ring_done[m] != -1) break;
        int c = 0;
        if (lane == 0) c = atomicAdd(P.work_cursor, 1);
        c = __shfl_sync(kFullMask, c, 0);
        if (c >= total_claims) { exhausted = true; break; }
        const int item = claim_to_item(P, c);
        if (item >= total) continue;
        __syncwarp();
        if (lane == 0) { ring_item[m] = item; ring_done[m] = 0; }
        __syncwarp();
        open_seq++;
        avail += spp;
      }
      const int give = f < avail ? f : avail;
      if (mine && lane < give) {
        int s = disp_s + lane, seq = disp_seq;
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
          p_sum[slot] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float((m << 16) | s));
        } else {
          atomicAdd(ring_done + m, 1);  // padding pixel of a partial tile: nothing to trace
        }
      }
      disp_s += give;
      while (disp_s >= spp) { disp_s -= spp; disp_seq++; }
    }
    const bool go = mine ? start_ray(slot) : false;
    push_roots_and_idle(mine, slot, go);
  };

  auto shade_batch = [&]() {
    const int d = dtop < 32 ? dtop : 32;
    const bool mine = lane < d;
    const int slot = mine ? (int)dstk[dtop - 1 - lane] : 0;
    __syncwarp();
    dtop -= d;
    bool go = false;
    if (mine) {
      const float4 ro = ray_o[slot], rd = ray_d[slot], pl = p_light[slot];
      const unsigned long long b = best[slot];
      Ray r;
      r.o = v3(ro.x, ro.y, ro.z);
      r.d = v3(rd.x, rd.y, rd.z);
      V3 light = v3(pl.x, pl.y, pl.z), colour;
      int depth = __float_as_int(pl.w);
      const int j = b == kNoHit ? -1 : (int)(unsigned)(b & 0xffffffffu);
      const float tb = __uint_as_float((unsigned)(b >> 32));
      if (shade_segment(sc, P, r, ro.w, j, tb, light, depth, colour)) {
        ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, 0.0f);
        ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
        p_light[slot] = make_float4(light.x, light.y, light.z, __int_as_float(depth));
      } else {
        finish_path(slot, colour);
      }
      go = start_ray(slot);
    }
    push_roots_and_idle(mine, slot, go);
  };

  auto leaf_batch = [&](auto full_tag) {
    constexpr bool kFull = decltype(full_tag)::value;
    const int n = kFull ? 32 : ltop;
    bool done = false;
    int slot = 0;
    if (kFull || lane < n) {
      const uint32_t it = lstk[ltop - 1 - lane];
      slot = (int)(it >> kSlotShift);
      const int li = (int)(it & kIndexMask);
      const float4 ro = ray_o[slot], rd = ray_d[slot];
      const float4 g = sc.sphere(li);
      Ray r;
      r.o = v3(ro.x, ro.y, ro.z);
      r.d = v3(rd.x, rd.y, rd.z);
      const float t = sphere_t(g.x, g.y, g.z, g.w, r, ro.w, 0.1f, 1000000000.0f);
      if (t >= 0.0f) atomicMin(best + slot, ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)li);
      done = atomicSub(pending + slot, 1) == 1;  // this was the ray's last outstanding item
    }
    ltop -= n;
    const unsigned dm = __ballot_sync(kFullMask, done);
    if (done) dstk[dtop + __popc(dm & lt_mask)] = (uint32_t)slot;
    dtop += __popc(dm);
  };
  auto node_batch = [&](auto full_tag, const int n_part) {
    constexpr bool kFull = decltype(full_tag)::value;
    const int n = kFull ? 32 : n_part;
    bool pl_node = false, pr_node = false, pl_leaf = false, pr_leaf = false, done = false;
    uint32_t tag = 0;
    int lptr = 0, rptr = 0, slot = 0;
    if (kFull || lane < n) {
      const uint32_t it = nstk[ntop - 1 - lane];
      tag = it & ~kIndexMask;
      slot = (int)(it >> kSlotShift);
      const int cur = (int)(it & kIndexMask);
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
      pl_leaf = lptr < 0;
      pr_leaf = rptr < 0;
      pl_node = hl && !pl_leaf;
      pr_node = hr && !pr_leaf;
      const int delta = (int)pl_node + (int)pr_node + (int)pl_leaf + (int)pr_leaf - 1;
      if (delta != 0) done = atomicAdd(pending + slot, delta) + delta == 0;
    }
    __syncwarp();  // all pops have been read before anything is pushed over them
    ntop -= n;
    const unsigned bl = __ballot_sync(kFullMask, pl_node), br = __ballot_sync(kFullMask, pr_node);
    const unsigned cl = __ballot_sync(kFullMask, pl_leaf), cr = __ballot_sync(kFullMask, pr_leaf);
    const int nb = ntop + __popc(bl & lt_mask) + __popc(br & lt_mask);
    if (pl_node) nstk[nb] = tag | (uint32_t)lptr;
    if (pr_node) nstk[nb + (pl_node ? 1 : 0)] = tag | (uint32_t)rptr;
    ntop += __popc(bl) + __popc(br);
    const int lb = ltop + __popc(cl & lt_mask) + __popc(cr & lt_mask);
    if (pl_leaf) lstk[lb] = tag | (uint32_t)(~lptr);
    if (pr_leaf) lstk[lb + (pl_leaf ? 1 : 0)] = tag | (uint32_t)(~rptr);
    ltop += __popc(cl) + __popc(cr);
    const unsigned dm = __ballot_sync(kFullMask, done);
    if (done) dstk[dtop + __popc(dm & lt_mask)] = (uint32_t)slot;
    dtop += __popc(dm);
  };
  using full_t = std::integral_constant<bool, true>;
  using part_t = std::integral_constant<bool, false>;

  for (;;) {
    const bool work_left = kSpread ? (!exhausted || avail_samples() > 0) : !exhausted;
    // a refill must be able to hand something out, or the loop would spin on it: samples already opened, or a ring entry
    // that is free / can be finalized right now
    bool can_hand_out = work_left;
    if (kSpread && avail_samples() <= 0) {
      const int rd = ring_done[open_seq & (kWqRing - 1)];
      can_hand_out = !exhausted && (rd == -1 || rd == spp);
    }
    const bool can_refill = ftop > 0 && can_hand_out;
    const bool tight = ntop + 96 > ncap;
    if (dtop >= 32) shade_batch();
    else if (ltop >= 32) leaf_batch(full_t{});
    else if (ntop >= 32 && !tight) node_batch(full_t{}, 32);
    else if (can_refill && ftop >= refill_min && !tight) refill();   // top the queues up before running partial batches
    else if (dtop > 0) shade_batch();
    else if (ntop > 0) node_batch(part_t{}, tight ? 1 : (ntop < 32 ? ntop : 32));
    else if (ltop > 0) leaf_batch(part_t{});
    else if (can_refill) refill();
    else if (kSpread && ftop == R && !work_left) {
      // everything handed out and finished: flush the last pixels; if that frees nothing there is nothing left
      finalize_pixels();
      break;
    } else break;
    __syncwarp();
  }
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
cudaError_t configure_kernels(int max_dynamic_smem) {
  cudaError_t e;
#define RAYB200_SET(k)                                                                          \
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem);   \
  if (e != cudaSuccess) return e;
  RAYB200_SET((render_persistent_kernel<true, true>));
  RAYB200_SET((render_persistent_kernel<true, false>));
  RAYB200_SET((render_persistent_kernel<false, true>));
  RAYB200_SET((render_persistent_kernel<false, false>));
  RAYB200_SET((wavefront_bounce_kernel<true, true>));
  RAYB200_SET((wavefront_bounce_kernel<true, false>));
  RAYB200_SET((wavefront_bounce_kernel<false, true>));
  RAYB200_SET((wavefront_bounce_kernel<false, false>));
#define RAYB200_SET_WQ(KK, SP, PK)                                  \
  RAYB200_SET((render_warpqueue_kernel<KK, SP, PK, true, true>));   \
  RAYB200_SET((render_warpqueue_kernel<KK, SP, PK, true, false>));  \
  RAYB200_SET((render_warpqueue_kernel<KK, SP, PK, false, true>));  \
  RAYB200_SET((render_warpqueue_kernel<KK, SP, PK, false, false>));
  RAYB200_SET_WQ(1, false, false) RAYB200_SET_WQ(1, true, false) RAYB200_SET_WQ(2, false, false) RAYB200_SET_WQ(2, true, false)
  RAYB200_SET_WQ(1, false, true) RAYB200_SET_WQ(1, true, true) RAYB200_SET_WQ(2, false, true) RAYB200_SET_WQ(2, true, true)
#undef RAYB200_SET_WQ
#define RAYB200_SET_SQ(KK, SP)                                    \
  RAYB200_SET((render_streamqueue_kernel<KK, SP, true, true>));   \
  RAYB200_SET((render_streamqueue_kernel<KK, SP, true, false>));  \
  RAYB200_SET((render_streamqueue_kernel<KK, SP, false, true>));  \
  RAYB200_SET((render_streamqueue_kernel<KK, SP, false, false>));
  RAYB200_SET_SQ(1, false) RAYB200_SET_SQ(1, true)
#undef RAYB200_SET_SQ
#undef RAYB200_SET
  return cudaSuccess;
}

void launch_render(const RenderParams &p, const LaunchConfig &lc, const WavefrontBuffers *wf, cudaStream_t stream,
                   int64_t *launches) {
  const long long items = p.local_tiles * kTilePixels;
  if (items <= 0) return;
  if (lc.kernel == 1) {  // RAY_B200_KERNEL_MEGA
    const int threads = 128;
    const unsigned blocks = (unsigned)((items + threads - 1) / threads);
    render_mega_kernel<false><<<blocks, threads, 0, stream>>>(p);
    (*launches)++;
    return;
  }
  const int threads = 256;
  const size_t smem = staging_bytes(p);
  const bool all_nodes = p.smem_nodes == p.n_inner, sph = p.smem_spheres == p.n_leaves && p.smem_spheres > 0;
  long long want = (long long)lc.sm_count * lc.blocks_per_sm;
  if (lc.kernel == 3) {  // RAY_B200_KERNEL_WAVEFRONT: per sample pass, one launch per bounce up to the tail bounce
    const long long max_useful = (items + 31) / 32;
    if (want > max_useful) want = max_useful;
    const int tail = wf->tail_from < 0 ? 0 : (wf->tail_from > kMaxDepth - 1 ? kMaxDepth - 1 : wf->tail_from);
    for (int s = 0; s < p.spp; s++) {
      cudaMemsetAsync(wf->qlen, 0, 2 * (kMaxDepth + 2) * sizeof(int32_t), stream);  // qlen and cursor are contiguous
      for (int b = 0; b <= tail; b++) {
        const int rte = b == tail;
        if (all_nodes && sph) wavefront_bounce_kernel<true, true><<<(unsigned)want, threads, smem, stream>>>(p, *wf, b, s, rte);
        else if (all_nodes) wavefront_bounce_kernel<true, false><<<(unsigned)want, threads, smem, stream>>>(p, *wf, b, s, rte);
        else if (sph) wavefront_bounce_kernel<false, true><<<(unsigned)want, threads, smem, stream>>>(p, *wf, b, s, rte);
        else wavefront_bounce_kernel<false, false><<<(unsigned)want, threads, smem, stream>>>(p, *wf, b, s, rte);
        (*launches)++;
      }
    }
    return;
  }
  if (lc.kernel == 5) {  // RAY_B200_KERNEL_STREAMQUEUE: one CTA per SM, rays refilled continuously (no rounds); 32 rays per warp
    const int k = 1;
    const int wthreads = 32 * lc.wq_warps;
    const int ncap = wq_node_capacity(k, p.max_depth);
    const size_t wsmem = ((staging_bytes(p) + 127) & ~(size_t)127) + (size_t)lc.wq_warps * sq_warp_bytes(k, ncap);
    long long ctas = lc.sm_count;
    const bool spread = p.sample_buf != nullptr;
    // no more CTAs than there are rays to start at once: a slot takes one SAMPLE when samples are spread, one pixel otherwise
    const long long rays = items * (spread ? (long long)p.spp : 1ll);
    const long long useful = (rays + 32 * k * lc.wq_warps - 1) / (32 * k * lc.wq_warps);
    if (ctas > useful) ctas = useful;
#define RAYB200_SQ(KK, SP, A, S) render_streamqueue_kernel<KK, SP, A, S><<<(unsigned)ctas, wthreads, wsmem, stream>>>(p, ncap, lc.wq_refill)
#define RAYB200_SQ2(KK, SP)                                                               \
  do {                                                                                    \
    if (all_nodes && sph) RAYB200_SQ(KK, SP, true, true);                                 \
    else if (all_nodes) RAYB200_SQ(KK, SP, true, false);                                  \
    else if (sph) RAYB200_SQ(KK, SP, false, true);                                        \
    else RAYB200_SQ(KK, SP, false, false);                                                \
  } while (0)
    if (spread) RAYB200_SQ2(1, true); else RAYB200_SQ2(1, false);
#undef RAYB200_SQ2
#undef RAYB200_SQ
    (*launches)++;
    return;
  }
  if (lc.kernel == 4) {  // RAY_B200_KERNEL_WARPQUEUE: one CTA per SM, wq_warps warps, 32*wq_k rays in flight per warp
    const int k = lc.wq_k == 1 ? 1 : 2;
    const int wthreads = 32 * lc.wq_warps;
    const int ncap = wq_node_capacity(k, p.max_depth, lc.wq_ncap);
    const bool packet = lc.wq_packet > 0;
    const size_t wsmem = ((staging_bytes(p) + 127) & ~(size_t)127) + (size_t)lc.wq_warps * wq_warp_bytes(k, ncap, packet);
    long long ctas = lc.sm_count;
    const bool spread = p.sample_buf != nullptr;
    // no more CTAs than there are rays to start at once: a slot takes one SAMPLE when samples are spread, one pixel
    // otherwise (capping by pixels alone left 26 SMs idle on a 125 K-pixel shard at 64 spp with 32 warps per CTA)
    const long long rays = items * (spread ? (long long)p.spp : 1ll);
    const long long useful = (rays + 32 * k * lc.wq_warps - 1) / (32 * k * lc.wq_warps);
    if (ctas > useful) ctas = useful;
#define RAYB200_WQ(KK, SP, PK, A, S) \
  render_warpqueue_kernel<KK, SP, PK, A, S><<<(unsigned)ctas, wthreads, wsmem, stream>>>(p, ncap, lc.wq_packet, lc.wq_refill)
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
    (*launches)++;
    return;
  }
  // RAY_B200_KERNEL_PERSISTENT
  const long long max_useful = (items + threads - 1) / threads;
  if (want > max_useful) want = max_useful;
  const int refill = lc.refill_min < 1 ? 1 : (lc.refill_min > 32 ? 32 : lc.refill_min);
  if (all_nodes && sph) render_persistent_kernel<true, true><<<(unsigned)want, threads, smem, stream>>>(p, refill);
  else if (all_nodes) render_persistent_kernel<true, false><<<(unsigned)want, threads, smem, stream>>>(p, refill);
  else if (sph) render_persistent_kernel<false, true><<<(unsigned)want, threads, smem, stream>>>(p, refill);
  else render_persistent_kernel<false, false><<<(unsigned)want, threads, smem, stream>>>(p, refill);
  (*launches)++;
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
