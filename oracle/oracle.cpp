// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A literal CPU restatement (strict IEEE f32, no FMA contraction) of the reference's Futhark ray
// tracer: futhark/prim.fut, futhark/ray.fut, futhark/bvh.fut, futhark/radixtree.fut and the
// *semantics* (stable sort by 32-bit key) of lib/github.com/diku-dk/sorts/radix_sort.fut.
// Every function cites the reference file:line it restates.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the product library
// (libray_b200.so) never links, loads or calls it.
//
// Parity status: PINNED.  The reference ships no tests for this path, but rgbbox.png / irreg.png
// (500x500, repo root) are outputs of the Futhark program; tests/test_oracle_golden.py checks this
// oracle reproduces both bit-for-bit (fixtures under tests/golden/, made by tools/make_golden.py).
// Extensions with no counterpart in the reference (spp > 1, float framebuffer, custom/random
// scenes) are "parity unpinned": they are defined by this file only.
//
// Third-party arithmetic not under /root/reference: the Futhark compiler's prelude (f32.sqrt/tan/
// log2/max/min, u32.clz, float->int truncation) — restated here with sqrtf/tanf/log2f/fmaxf/fminf/
// __builtin_clz and C casts; the golden images pin those choices.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -pthread; no -march, no -ffast-math).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include <atomic>
#include <mutex>
#include <thread>

namespace {

// ---------------------------------------------------------------- prim.fut
struct V3 { float x, y, z; };                                           // prim.fut:1

inline V3 vec(float x, float y, float z) { return V3{x, y, z}; }        // prim.fut:5
inline V3 vec_add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // prim.fut:12
inline V3 vec_sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // prim.fut:13
inline V3 vec_mul(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // prim.fut:14
inline V3 scale(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }       // prim.fut:17-20
inline float dot(V3 a, V3 b) {                                               // prim.fut:22-24
  V3 v3 = vec_mul(a, b);
  return v3.x + v3.y + v3.z;
}
inline float norm(V3 v) { return sqrtf(dot(v, v)); }                         // prim.fut:26
inline V3 normalise(V3 v) { return scale(1.0f / norm(v), v); }               // prim.fut:28
inline V3 cross(V3 a, V3 b) {                                                // prim.fut:30-33
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct Aabb { V3 min, max; };                                                // prim.fut:36

inline Aabb enclosing(const Aabb &b0, const Aabb &b1) {                      // prim.fut:38-45
  V3 small = vec(fminf(b0.min.x, b1.min.x), fminf(b0.min.y, b1.min.y), fminf(b0.min.z, b1.min.z));
  V3 big = vec(fmaxf(b0.max.x, b1.max.x), fmaxf(b0.max.y, b1.max.y), fmaxf(b0.max.z, b1.max.z));
  return {small, big};
}
inline V3 centre(const Aabb &b) {                                            // prim.fut:47-50
  return {b.min.x + 0.5f * (b.max.x - b.min.x), b.min.y + 0.5f * (b.max.y - b.min.y),
          b.min.z + 0.5f * (b.max.z - b.min.z)};
}

// ---------------------------------------------------------------- ray.fut types
const float scene_epsilon = 0.1f;                                            // ray.fut:3
struct Ray { V3 origin, dir; };                                              // ray.fut:11-12
inline V3 point_at_param(const Ray &r, float t) { return vec_add(r.origin, scale(t, r.dir)); }  // ray.fut:14-15
struct Hit { float t; V3 p, normal, colour; };                               // ray.fut:17-20
struct Sphere { V3 pos, colour; float radius; };                             // ray.fut:22-24

inline Aabb sphere_aabb(const Sphere &s) {                                   // ray.fut:28-30
  V3 rr = {s.radius, s.radius, s.radius};
  return {vec_sub(s.pos, rr), vec_add(s.pos, rr)};
}

// ray.fut:32-51.  Returns true and fills *out on #some.
inline bool sphere_hit(const Sphere &s, const Ray &r, float t_min, float t_max, Hit *out) {
  V3 oc = vec_sub(r.origin, s.pos);
  float a = dot(r.dir, r.dir);
  float b = dot(oc, r.dir);
  float c = dot(oc, oc) - s.radius * s.radius;
  float discriminant = b * b - a * c;
  auto f = [&](float temp) -> bool {
    if (temp < t_max && temp > t_min) {
      out->t = temp;
      out->p = point_at_param(r, temp);
      out->normal = scale(1.0f / s.radius, vec_sub(point_at_param(r, temp), s.pos));
      out->colour = s.colour;
      return true;
    }
    return false;
  };
  if (discriminant <= 0.0f) return false;
  if (f((-b - sqrtf(b * b - a * c)) / a)) return true;
  return f((-b + sqrtf(b * b - a * c)) / a);
}

// ray.fut:53-70
inline bool aabb_hit(const Aabb &box, const Ray &r, float tmin0, float tmax0) {
  auto iter = [](float min_, float max_, float origin_, float dir_, float tmin_, float tmax_, float *tmin_o,
                 float *tmax_o) {
    float invD = 1.0f / dir_;
    float t0 = (min_ - origin_) * invD;
    float t1 = (max_ - origin_) * invD;
    float t0p = t0, t1p = t1;
    if (invD < 0.0f) { t0p = t1; t1p = t0; }
    *tmin_o = fmaxf(t0p, tmin_);
    *tmax_o = fminf(t1p, tmax_);
  };
  float tmin1, tmax1, tmin2, tmax2, tmin3, tmax3;
  iter(box.min.x, box.max.x, r.origin.x, r.dir.x, tmin0, tmax0, &tmin1, &tmax1);
  if (tmax1 <= tmin1) return false;
  iter(box.min.y, box.max.y, r.origin.y, r.dir.y, tmin1, tmax1, &tmin2, &tmax2);
  if (tmax2 <= tmin2) return false;
  iter(box.min.z, box.max.z, r.origin.z, r.dir.z, tmin2, tmax2, &tmin3, &tmax3);
  return !(tmax3 <= tmin3);
}

// ---------------------------------------------------------------- bvh.fut / radixtree.fut
struct Ptr { int32_t tag; int32_t idx; };                 // bvh.fut:24 / radixtree.fut:6  tag 0=#leaf 1=#inner
inline bool ptr_eq(Ptr a, Ptr b) { return a.tag == b.tag && a.idx == b.idx; }
inline Ptr leaf_ptr(int32_t i) { return {0, i}; }
inline Ptr inner_ptr(int32_t i) { return {1, i}; }

struct Inner { Aabb aabb; Ptr left, right; int32_t parent; };  // bvh.fut:26

struct Bvh {                                                    // bvh.fut:28
  std::vector<Sphere> L;
  std::vector<Inner> I;
  std::vector<uint32_t> morton;  // sorted keys (kept for tests)
  std::vector<int32_t> perm;     // L[k] = input[perm[k]] (kept for tests)
  int32_t sweeps = 0;
};

inline uint32_t expand_bits(uint32_t v) {                       // bvh.fut:8-13
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

inline uint32_t morton_3D(V3 p) {                               // bvh.fut:15-22
  float x = fminf(fmaxf(p.x * 1024.0f, 0.0f), 1023.0f);
  float y = fminf(fmaxf(p.y * 1024.0f, 0.0f), 1023.0f);
  float z = fminf(fmaxf(p.z * 1024.0f, 0.0f), 1023.0f);
  uint32_t xx = expand_bits((uint32_t)x);
  uint32_t yy = expand_bits((uint32_t)y);
  uint32_t zz = expand_bits((uint32_t)z);
  return xx * 4u + yy * 2u + zz;
}

inline int32_t clz32(uint32_t x) { return x == 0 ? 32 : __builtin_clz(x); }
inline int32_t sgn32(int32_t x) { return (x > 0) - (x < 0); }

struct RadixNode { Ptr left, right; int32_t parent; };         // radixtree.fut:8

// radixtree.fut:11-72
std::vector<RadixNode> mk_radix_tree(const std::vector<uint32_t> &Lk) {
  const int32_t n = (int32_t)Lk.size();
  auto delta = [&](int32_t i, int32_t j) -> int32_t {          // radixtree.fut:13-21
    if (j >= 0 && j < n) {
      uint32_t Li = Lk[i], Lj = Lk[j];
      if (Li == Lj) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
      return clz32(Li ^ Lj);
    }
    return -1;
  };
  std::vector<RadixNode> out((size_t)(n - 1));
  std::vector<int64_t> pa_idx((size_t)(n - 1)), pb_idx((size_t)(n - 1));
  for (int32_t i = 0; i < n - 1; i++) {                         // radixtree.fut:23-64 (node i)
    int32_t d = sgn32(delta(i, i + 1) - delta(i, i - 1));       // :27
    int32_t delta_min = delta(i, i - d);                        // :30
    int32_t l_max = 2;                                          // :31-33
    while (delta(i, i + l_max * d) > delta_min) l_max *= 2;
    int32_t l = 0;                                              // :36-40
    for (int32_t t = l_max / 2; t > 0; t /= 2)
      if (delta(i, i + (l + t) * d) > delta_min) l += t;
    int32_t j = i + l * d;                                      // :41
    int32_t delta_node = delta(i, j);                           // :44
    int32_t s = 0;                                              // :45-50
    for (int32_t q = 1; q <= l; q *= 2) {
      int32_t t = (l + q * 2 - 1) / (q * 2);
      if (delta(i, i + (s + t) * d) > delta_node) s += t;
    }
    int32_t gamma = i + s * d + std::min(d, 0);                 // :51
    if (std::min(i, j) == gamma) { out[i].left = leaf_ptr(gamma); pa_idx[i] = -1; }   // :54-57
    else { out[i].left = inner_ptr(gamma); pa_idx[i] = gamma; }
    if (std::max(i, j) == gamma + 1) { out[i].right = leaf_ptr(gamma + 1); pb_idx[i] = -1; }  // :59-62
    else { out[i].right = inner_ptr(gamma + 1); pb_idx[i] = gamma + 1; }
    out[i].parent = -1;                                         // :68 replicate (n-1) (-1)
  }
  // radixtree.fut:66-70 scatter (out-of-range index -1 is dropped)
  for (int32_t i = 0; i < n - 1; i++) if (pa_idx[i] >= 0 && pa_idx[i] < n - 1) out[pa_idx[i]].parent = i;
  for (int32_t i = 0; i < n - 1; i++) if (pb_idx[i] >= 0 && pb_idx[i] < n - 1) out[pb_idx[i]].parent = i;
  return out;
}

// bvh.fut:30-59 with bbf = sphere_aabb
Bvh bvh_mk(const std::vector<Sphere> &ts_in) {
  const size_t n = ts_in.size();
  std::vector<V3> centers(n);
  for (size_t k = 0; k < n; k++) centers[k] = centre(sphere_aabb(ts_in[k]));     // :31
  // f32.maximum / f32.minimum = reduce f32.max/min (fmaxf/fminf) with -inf/+inf neutral   :32-37
  float x_max = -INFINITY, y_max = -INFINITY, z_max = -INFINITY;
  float x_min = INFINITY, y_min = INFINITY, z_min = INFINITY;
  for (size_t k = 0; k < n; k++) {
    x_max = fmaxf(x_max, centers[k].x); y_max = fmaxf(y_max, centers[k].y); z_max = fmaxf(z_max, centers[k].z);
    x_min = fminf(x_min, centers[k].x); y_min = fminf(y_min, centers[k].y); z_min = fminf(z_min, centers[k].z);
  }
  auto morton = [&](const Sphere &s) -> uint32_t {                                // :38-41
    V3 c = centre(sphere_aabb(s));
    V3 nrm = {(c.x - x_min) / (x_max - x_min), (c.y - y_min) / (y_max - y_min), (c.z - z_min) / (z_max - z_min)};
    return morton_3D(nrm);
  };
  // :43 radix_sort_by_key morton 32 get_bit — radix_sort.fut:14-32,50-68: LSD passes that each
  // preserve relative order => a stable sort of (key, original index) by key.
  std::vector<uint32_t> keys(n);
  for (size_t k = 0; k < n; k++) keys[k] = morton(ts_in[k]);
  std::vector<int32_t> perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return keys[a] < keys[b]; });
  Bvh out;
  out.L.resize(n);
  out.morton.resize(n);
  out.perm = perm;
  for (size_t k = 0; k < n; k++) { out.L[k] = ts_in[perm[k]]; out.morton[k] = keys[perm[k]]; }
  // :44-46
  std::vector<RadixNode> rt = mk_radix_tree(out.morton);
  std::vector<Inner> inners(rt.size());
  const Aabb empty_aabb = {vec(0, 0, 0), vec(0, 0, 0)};
  for (size_t k = 0; k < rt.size(); k++) inners[k] = {empty_aabb, rt[k].left, rt[k].right, rt[k].parent};
  int32_t depth = (int32_t)log2f((float)(int64_t)n) + 2;                          // :47
  out.sweeps = depth;
  auto get_aabb = [&](const std::vector<Inner> &in, Ptr p) -> Aabb {              // :48-51
    return p.tag == 0 ? sphere_aabb(out.L[p.idx]) : in[p.idx].aabb;
  };
  for (int32_t it = 0; it < depth; it++) {                                        // :57-58 (Jacobi: new from old)
    std::vector<Inner> next(inners.size());
    for (size_t k = 0; k < inners.size(); k++) {
      const Inner &nd = inners[k];
      next[k] = {enclosing(get_aabb(inners, nd.left), get_aabb(inners, nd.right)), nd.left, nd.right, nd.parent};
    }
    inners.swap(next);
  }
  out.I = std::move(inners);
  return out;
}

struct Counters { uint64_t segments = 0, iterations = 0, box_tests = 0, leaf_tests = 0; };

// ray.fut:76-86 objs_hit, with bvh.fut:61-84 bvh_fold inlined literally.
inline bool objs_hit(const Bvh &bvh, const Ray &r, float t_min, float t_max, Hit *out, Counters *cn) {
  int32_t acc_j = -1;
  float acc_t = t_max;
  int32_t cur = 0;
  Ptr prev = inner_ptr(-1);
  cn->segments++;
  while (cur != -1) {                                                  // bvh.fut:64
    cn->iterations++;
    const Inner &node = bvh.I[cur];
    bool from_left = ptr_eq(prev, node.left);
    bool from_right = ptr_eq(prev, node.right);
    bool rec = false;
    Ptr child = {0, 0};
    if (from_left) { rec = true; child = node.right; }                 // bvh.fut:70-71
    else if (!from_right) {                                            // bvh.fut:73-76
      cn->box_tests++;
      if (aabb_hit(node.aabb, r, t_min, t_max)) { rec = true; child = node.left; }   // ray.fut:77 (original t range)
    }
    if (!rec) { prev = inner_ptr(cur); cur = node.parent; }            // bvh.fut:79-80
    else if (child.tag == 1) { prev = inner_ptr(cur); cur = child.idx; }  // bvh.fut:83
    else {                                                             // bvh.fut:84: op acc i L[i]
      cn->leaf_tests++;
      Hit h;
      if (sphere_hit(bvh.L[child.idx], r, scene_epsilon, acc_t, &h)) { acc_j = child.idx; acc_t = h.t; }  // ray.fut:78-81
      prev = child;
    }
  }
  if (acc_j >= 0) return sphere_hit(bvh.L[acc_j], r, t_min, acc_t + 1.0f, out);   // ray.fut:83-85
  return false;
}

struct Camera { V3 origin, llc, horizontal, vertical; };               // ray.fut:88-91

Camera camera(V3 lookfrom, V3 lookat, V3 vup, float vfov, float aspect) {   // ray.fut:93-107
  float theta = vfov * (float)M_PI / 180.0f;
  float half_height = tanf(theta / 2.0f);
  float half_width = aspect * half_height;
  V3 origin = lookfrom;
  V3 w = normalise(vec_sub(lookfrom, lookat));
  V3 u = normalise(cross(vup, w));
  V3 v = cross(w, u);
  Camera c;
  c.origin = lookfrom;
  c.llc = vec_sub(vec_sub(vec_sub(origin, scale(half_width, u)), scale(half_height, v)), w);
  c.horizontal = scale(2.0f * half_width, u);
  c.vertical = scale(2.0f * half_height, v);
  return c;
}

inline Ray get_ray(const Camera &cam, float s, float t) {              // ray.fut:109-114
  return {cam.origin,
          vec_sub(vec_add(vec_add(cam.llc, scale(s, cam.horizontal)), scale(t, cam.vertical)), cam.origin)};
}

inline V3 reflect(V3 v, V3 n) { return vec_sub(v, scale(2.0f * dot(v, n), n)); }   // ray.fut:116-117

inline bool scatter(const Ray &r, const Hit &hit, Ray *scattered, V3 *attenuation) {   // ray.fut:119-124
  V3 reflected = reflect(normalise(r.dir), hit.normal);
  *scattered = {hit.p, reflected};
  if (dot(scattered->dir, hit.normal) > 0.0f) { *attenuation = hit.colour; return true; }
  return false;
}

V3 ray_colour(const Bvh &objs, Ray r, int32_t max_depth, Counters *cn) {    // ray.fut:126-148
  int32_t depth = 0;
  V3 light = vec(1, 1, 1), colour = vec(0, 0, 0);
  while (depth < max_depth) {
    Hit hit;
    if (objs_hit(objs, r, 0.000f, 1000000000.0f, &hit, cn)) {
      Ray scattered; V3 attenuation;
      if (scatter(r, hit, &scattered, &attenuation)) {
        V3 nl = vec_mul(light, attenuation), nc = vec_mul(light, colour);
        r = scattered; depth = depth + 1; light = nl; colour = nc;
      } else {
        colour = vec_mul(light, colour); depth = max_depth;
      }
    } else {
      V3 unit_dir = normalise(r.dir);
      float t = 0.5f * (unit_dir.y + 1.0f);
      V3 bg = {0.5f, 0.7f, 1.0f};
      colour = vec_mul(light, vec_add(scale(1.0f - t, vec(1, 1, 1)), scale(t, bg)));
      depth = max_depth;
    }
  }
  return colour;
}

inline int32_t colour_to_pixel(V3 p) {                                  // ray.fut:158-162
  int32_t ir = (int32_t)(255.99f * p.x);
  int32_t ig = (int32_t)(255.99f * p.y);
  int32_t ib = (int32_t)(255.99f * p.z);
  return (ir << 16) | (ig << 8) | ib;
}

struct Scene { V3 look_from, look_at; float fov; std::vector<Sphere> spheres; };   // ray.fut:171-174
struct Prepared { Bvh objs; Camera cam; };                                        // ray.fut:239

Scene make_rgbbox() {                                                   // ray.fut:176-221
  const int64_t n = 10;
  const float k = 60.0f;
  Scene s;
  auto wall = [&](auto f) { for (int64_t a = 0; a < n; a++) for (int64_t b = 0; b < n; b++) s.spheres.push_back(f(a, b)); };
  const float rad = k / ((float)n * 2.0f);
  wall([&](int64_t y, int64_t z) { return Sphere{{-k / 2.0f, -k / 2.0f + (k / (float)n) * (float)y, -k / 2.0f + (k / (float)n) * (float)z}, {1, 0, 0}, rad}; });
  wall([&](int64_t x, int64_t y) { return Sphere{{-k / 2.0f + (k / (float)n) * (float)x, -k / 2.0f + (k / (float)n) * (float)y, -k / 2.0f}, {1, 1, 0}, rad}; });
  wall([&](int64_t y, int64_t z) { return Sphere{{k / 2.0f, -k / 2.0f + (k / (float)n) * (float)y, -k / 2.0f + (k / (float)n) * (float)z}, {0, 0, 1}, rad}; });
  wall([&](int64_t x, int64_t z) { return Sphere{{-k / 2.0f + (k / (float)n) * (float)x, -k / 2.0f, -k / 2.0f + (k / (float)n) * (float)z}, {1, 1, 1}, rad}; });
  s.look_from = {0.0f, 30.0f, 30.0f};
  s.look_at = {0.0f, -1.0f, -1.0f};
  s.fov = 75.0f;
  return s;
}

Scene make_irreg() {                                                    // ray.fut:223-237
  const int64_t n = 100;
  const float k = 600.0f;
  Scene s;
  for (int64_t x = 0; x < n; x++)
    for (int64_t z = 0; z < n; z++)
      s.spheres.push_back(Sphere{{-k / 2.0f + (k / (float)n) * (float)x, 0.0f, -k / 2.0f + (k / (float)n) * (float)z}, {1, 1, 1}, k / ((float)n * 2.0f)});
  s.look_from = {0.0f, 12.0f, 30.0f};
  s.look_at = {0.0f, 10.0f, -1.0f};
  s.fov = 75.0f;
  return s;
}

// Extension (no reference counterpart): splitmix64 stream, 7 draws/sphere.
Scene make_random(int64_t n, uint64_t seed) {
  Scene s;
  uint64_t st = seed;
  auto next = [&]() -> float {
    uint64_t z = (st += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    return (float)(z >> 40) * 5.9604644775390625e-8f;  // 2^-24
  };
  s.spheres.resize((size_t)n);
  for (int64_t k = 0; k < n; k++) {
    Sphere sp;
    sp.pos.x = -500.0f + 1000.0f * next(); sp.pos.y = -500.0f + 1000.0f * next(); sp.pos.z = -500.0f + 1000.0f * next();
    sp.colour.x = 0.25f + 0.75f * next(); sp.colour.y = 0.25f + 0.75f * next(); sp.colour.z = 0.25f + 0.75f * next();
    sp.radius = 0.5f + 1.5f * next();
    s.spheres[(size_t)k] = sp;
  }
  s.look_from = {0.0f, 0.0f, 1100.0f};
  s.look_at = {0.0f, 0.0f, 0.0f};
  s.fov = 75.0f;
  return s;
}

// Extension: sub-pixel offsets for sample s.  s = 0 -> (0,0): bit-identical to the reference.
inline void sample_offset(int32_t s, float *ox, float *oy) {
  float a = (float)s * 0.7548776662f, b = (float)s * 0.5698402909f;
  *ox = a - floorf(a);
  *oy = b - floorf(b);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI (ctypes-friendly).  All handles are opaque.
extern "C" {

struct oracle_counters { uint64_t segments, iterations, box_tests, leaf_tests; };

void *oracle_scene_rgbbox(void) { return new Scene(make_rgbbox()); }
void *oracle_scene_irreg(void) { return new Scene(make_irreg()); }
void *oracle_scene_random(int64_t n, uint64_t seed) { return new Scene(make_random(n, seed)); }
// spheres: n x 7 floats (pos.xyz, colour.xyz, radius); cam: look_from.xyz, look_at.xyz, fov
void *oracle_scene_custom(const float *spheres, int64_t n, const float *cam7) {
  Scene *s = new Scene;
  s->spheres.resize((size_t)n);
  for (int64_t k = 0; k < n; k++) {
    const float *p = spheres + 7 * k;
    s->spheres[(size_t)k] = Sphere{{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, p[6]};
  }
  s->look_from = {cam7[0], cam7[1], cam7[2]};
  s->look_at = {cam7[3], cam7[4], cam7[5]};
  s->fov = cam7[6];
  return s;
}
void oracle_scene_free(void *s) { delete (Scene *)s; }
int64_t oracle_scene_num_spheres(void *s) { return (int64_t)((Scene *)s)->spheres.size(); }
void oracle_scene_get(void *sv, float *spheres, float *cam7) {
  Scene *s = (Scene *)sv;
  for (size_t k = 0; k < s->spheres.size(); k++) {
    const Sphere &sp = s->spheres[k];
    float *p = spheres + 7 * k;
    p[0] = sp.pos.x; p[1] = sp.pos.y; p[2] = sp.pos.z; p[3] = sp.colour.x; p[4] = sp.colour.y; p[5] = sp.colour.z; p[6] = sp.radius;
  }
  cam7[0] = s->look_from.x; cam7[1] = s->look_from.y; cam7[2] = s->look_from.z;
  cam7[3] = s->look_at.x; cam7[4] = s->look_at.y; cam7[5] = s->look_at.z; cam7[6] = s->fov;
}

// ray.fut:241-244
void *oracle_prepare_scene(int64_t h, int64_t w, void *sv) {
  Scene *s = (Scene *)sv;
  if (s->spheres.size() < 2) return nullptr;  // bvh.fut:65 indexes I[0] (S19)
  Prepared *p = new Prepared;
  p->objs = bvh_mk(s->spheres);
  p->cam = camera(s->look_from, s->look_at, vec(0.0f, 1.0f, 0.0f), s->fov, (float)w / (float)h);
  return p;
}
void oracle_prepared_free(void *p) { delete (Prepared *)p; }
int32_t oracle_prepared_sweeps(void *p) { return ((Prepared *)p)->objs.sweeps; }
// Dumps for structure tests: morton[n], perm[n], left[n-1], right[n-1] (leaf i -> ~i, inner i -> i),
// parent[n-1], boxes[(n-1)*6] (min.xyz,max.xyz), cam[12].  Any pointer may be NULL.
void oracle_prepared_dump(void *pv, uint32_t *morton, int32_t *perm, int32_t *left, int32_t *right, int32_t *parent,
                          float *boxes, float *cam12) {
  Prepared *p = (Prepared *)pv;
  const Bvh &b = p->objs;
  if (morton) std::copy(b.morton.begin(), b.morton.end(), morton);
  if (perm) std::copy(b.perm.begin(), b.perm.end(), perm);
  for (size_t k = 0; k < b.I.size(); k++) {
    const Inner &nd = b.I[k];
    if (left) left[k] = nd.left.tag ? nd.left.idx : ~nd.left.idx;
    if (right) right[k] = nd.right.tag ? nd.right.idx : ~nd.right.idx;
    if (parent) parent[k] = nd.parent;
    if (boxes) { float *q = boxes + 6 * k; q[0] = nd.aabb.min.x; q[1] = nd.aabb.min.y; q[2] = nd.aabb.min.z; q[3] = nd.aabb.max.x; q[4] = nd.aabb.max.y; q[5] = nd.aabb.max.z; }
  }
  if (cam12) {
    const Camera &c = p->cam;
    const float v[12] = {c.origin.x, c.origin.y, c.origin.z, c.llc.x, c.llc.y, c.llc.z, c.horizontal.x, c.horizontal.y, c.horizontal.z, c.vertical.x, c.vertical.y, c.vertical.z};
    std::copy(v, v + 12, cam12);
  }
}

// ray.fut:246-247 render (+ ray.fut:150-169).  Renders output rows j = row_start, row_start+row_step, ...
// (row_start=0,row_step=1: the whole image; other values: a bounded sample for CPU timing).
// spp = 1 is the reference; spp > 1 is the extension of SURVEY.md §8d.  out_pix / out_rgb: [h][w] / [h][w][3],
// either may be NULL; untouched rows are left as they were.  threads <= 0: all cores.
int oracle_render(void *pv, int64_t h, int64_t w, int32_t spp, int32_t *out_pix, float *out_rgb, int32_t row_start,
                  int32_t row_step, int32_t threads, oracle_counters *cnt) {
  Prepared *p = (Prepared *)pv;
  if (!p || h <= 0 || w <= 0 || spp <= 0 || row_step <= 0) return 1;
  const Bvh &objs = p->objs;
  const Camera cam = p->cam;
  std::vector<float> ox((size_t)spp), oy((size_t)spp);
  for (int32_t s = 0; s < spp; s++) sample_offset(s, &ox[(size_t)s], &oy[(size_t)s]);
  const float inv_spp = 1.0f / (float)spp;
  Counters total;
  std::mutex total_mu;
  int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  const int64_t nrows = (h - row_start + row_step - 1) / row_step;
  const int64_t span = 64;                       // work unit: 64 consecutive pixels of one row
  const int64_t spans_per_row = (w + span - 1) / span;
  const int64_t nunits = nrows * spans_per_row;
  if (nthreads > nunits) nthreads = (int)std::max<int64_t>(nunits, 1);
  std::atomic<int64_t> next_unit{0};  // dynamic scheduling (irreg is load-imbalanced by rows)
  auto worker = [&]() {
    Counters local;
    for (;;) {
      int64_t unit = next_unit.fetch_add(1);
      if (unit >= nunits) break;
      const int64_t r = unit / spans_per_row, i0 = (unit % spans_per_row) * span;
      const int64_t j = row_start + r * row_step;                // ray.fut:166-169 tabulate_2d height width
      for (int64_t i = i0; i < std::min(w, i0 + span); i++) {
        V3 sum = vec(0, 0, 0);
        for (int32_t s = 0; s < spp; s++) {
          // ray.fut:150-154 with pixel j i = trace_ray ... (height-j) i (ray.fut:167-168)
          float u = ((float)i + ox[(size_t)s]) / (float)w;
          float v = ((float)(h - j) + oy[(size_t)s]) / (float)h;
          if (spp == 1) { u = (float)i / (float)w; v = (float)(h - j) / (float)h; }
          V3 c = ray_colour(objs, get_ray(cam, u, v), 50, &local);
          sum = (s == 0) ? c : vec_add(sum, c);
        }
        V3 col = (spp == 1) ? sum : scale(inv_spp, sum);
        if (out_pix) out_pix[j * w + i] = colour_to_pixel(col);
        if (out_rgb) { float *q = out_rgb + 3 * (j * w + i); q[0] = col.x; q[1] = col.y; q[2] = col.z; }
      }
    }
    std::lock_guard<std::mutex> g(total_mu);
    total.segments += local.segments; total.iterations += local.iterations;
    total.box_tests += local.box_tests; total.leaf_tests += local.leaf_tests;
  };
  if (nthreads == 1) worker();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
  }
  if (cnt) { cnt->segments = total.segments; cnt->iterations = total.iterations; cnt->box_tests = total.box_tests; cnt->leaf_tests = total.leaf_tests; }
  return 0;
}

// Per-pixel work of the reference traversal (sum over the samples): bvh_fold iterations and objs_hit calls.  Used to
// study how evenly a tile -> GPU assignment spreads a frame (tools/shard_balance.py, tests); same loop as oracle_render.
int oracle_render_cost(void *pv, int64_t h, int64_t w, int32_t spp, int32_t *out_iterations, int32_t *out_segments, int32_t threads) {
  Prepared *p = (Prepared *)pv;
  if (!p || h <= 0 || w <= 0 || spp <= 0 || !out_iterations) return 1;
  std::vector<float> ox((size_t)spp), oy((size_t)spp);
  for (int32_t s = 0; s < spp; s++) sample_offset(s, &ox[(size_t)s], &oy[(size_t)s]);
  int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  std::atomic<int64_t> next_row{0};
  auto worker = [&]() {
    for (;;) {
      const int64_t j = next_row.fetch_add(1);
      if (j >= h) break;
      for (int64_t i = 0; i < w; i++) {
        Counters local;
        for (int32_t s = 0; s < spp; s++) {
          float u = ((float)i + ox[(size_t)s]) / (float)w;
          float v = ((float)(h - j) + oy[(size_t)s]) / (float)h;
          if (spp == 1) { u = (float)i / (float)w; v = (float)(h - j) / (float)h; }
          (void)ray_colour(p->objs, get_ray(p->cam, u, v), 50, &local);
        }
        out_iterations[j * w + i] = (int32_t)local.iterations;
        if (out_segments) out_segments[j * w + i] = (int32_t)local.segments;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
  for (auto &t : pool) t.join();
  return 0;
}

int oracle_num_procs(void) { int n = (int)std::thread::hardware_concurrency(); return n > 0 ? n : 1; }

// unit-test hooks -------------------------------------------------------------------------------
uint32_t oracle_morton_3d(float x, float y, float z) { return morton_3D({x, y, z}); }
int oracle_aabb_hit(const float *box6, const float *ray6) {
  Aabb b = {{box6[0], box6[1], box6[2]}, {box6[3], box6[4], box6[5]}};
  Ray r = {{ray6[0], ray6[1], ray6[2]}, {ray6[3], ray6[4], ray6[5]}};
  return aabb_hit(b, r, 0.0f, 1000000000.0f) ? 1 : 0;
}
// sphere7 = pos, colour, radius.  out10 = t, p.xyz, normal.xyz, colour.xyz
int oracle_sphere_hit(const float *sphere7, const float *ray6, float t_min, float t_max, float *out10) {
  Sphere s = {{sphere7[0], sphere7[1], sphere7[2]}, {sphere7[3], sphere7[4], sphere7[5]}, sphere7[6]};
  Ray r = {{ray6[0], ray6[1], ray6[2]}, {ray6[3], ray6[4], ray6[5]}};
  Hit h;
  if (!sphere_hit(s, r, t_min, t_max, &h)) return 0;
  const float v[10] = {h.t, h.p.x, h.p.y, h.p.z, h.normal.x, h.normal.y, h.normal.z, h.colour.x, h.colour.y, h.colour.z};
  std::copy(v, v + 10, out10);
  return 1;
}
void oracle_sample_offset(int32_t s, float *ox, float *oy) { sample_offset(s, ox, oy); }
// stable sort permutation of u32 keys (radix_sort.fut:65-68 semantics), for the sort KATs
void oracle_sort_perm(const uint32_t *keys, int64_t n, int32_t *perm) {
  std::iota(perm, perm + n, 0);
  std::stable_sort(perm, perm + n, [&](int32_t a, int32_t b) { return keys[a] < keys[b]; });
}
void oracle_radix_tree(const uint32_t *sorted_keys, int64_t n, int32_t *left, int32_t *right, int32_t *parent) {
  std::vector<uint32_t> k(sorted_keys, sorted_keys + n);
  std::vector<RadixNode> t = mk_radix_tree(k);
  for (size_t i = 0; i < t.size(); i++) {
    left[i] = t[i].left.tag ? t[i].left.idx : ~t[i].left.idx;
    right[i] = t[i].right.tag ? t[i].right.idx : ~t[i].right.idx;
    parent[i] = t[i].parent;
  }
}

}  // extern "C"
