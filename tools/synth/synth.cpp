// Deterministic procedural RGB-D frames for tests and bench: a textured
// Manhattan corridor / room seen by a pinhole camera, gray u8 + depth f32.
// Test / bench INPUT GENERATOR, not part of the product: built as tools/synth/libdrfe_synth.so (plain g++),
// declared in tools/synth/drfe_synth.h, loaded by tools/synth/synth.py.  Host only; no GPU, no libm
// transcendental calls (only + - * / sqrt floor) and built with -ffp-contract=off so the
// same seed gives the same bytes on any x86-64 host.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "drfe_synth.h"

namespace {

inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t hash3(uint64_t a, uint64_t b, uint64_t c) { return mix64(mix64(mix64(a) ^ b) ^ c); }
inline float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }
// ~N(0,1): Irwin-Hall with 4 uniforms taken from one 64-bit hash (var 4/12 -> scale sqrt(3))
inline float gauss(uint64_t h) {
  float s = (float)(h & 0xFFFF) + (float)((h >> 16) & 0xFFFF) + (float)((h >> 32) & 0xFFFF) +
            (float)((h >> 48) & 0xFFFF);
  return (s * (1.0f / 65536.0f) - 2.0f) * 1.7320508f;
}

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 norm(V3 a) { float n = std::sqrt(dot(a, a)); return a * (1.0f / n); }

struct Box { V3 lo, hi; int id; };
struct Pillar { float cx, cz, r, ytop; int id; };

// smooth lattice value noise in [0,1]
float vnoise(float a, float b, uint64_t salt) {
  float fa = std::floor(a), fb = std::floor(b);
  float ta = a - fa, tb = b - fb;
  ta = ta * ta * (3.f - 2.f * ta); tb = tb * tb * (3.f - 2.f * tb);
  int64_t ia = (int64_t)fa, ib = (int64_t)fb;
  float v00 = u01(hash3(salt, (uint64_t)ia, (uint64_t)ib)), v10 = u01(hash3(salt, (uint64_t)(ia + 1), (uint64_t)ib));
  float v01 = u01(hash3(salt, (uint64_t)ia, (uint64_t)(ib + 1))), v11 = u01(hash3(salt, (uint64_t)(ia + 1), (uint64_t)(ib + 1)));
  float v0 = v00 + (v10 - v00) * ta, v1 = v01 + (v11 - v01) * ta;
  return v0 + (v1 - v0) * tb;
}

// gray level of surface `id` at in-face coordinates (a,b) metres
float texture(int id, float a, float b) {
  const uint64_t salt = 0xD1F3ull * (uint64_t)(id + 1);
  const int kind = id % 3;
  const float period = 0.14f + 0.05f * (float)(id % 4);
  float base;
  if (kind == 0) {  // checker
    int ia = (int)std::floor(a / period), ib = (int)std::floor(b / period);
    base = ((ia + ib) & 1) ? 178.f : 74.f;
    base += 36.f * (u01(hash3(salt, (uint64_t)(int64_t)ia, (uint64_t)(int64_t)ib)) - 0.5f);
  } else if (kind == 1) {  // bricks: rows of height period/2, every other row shifted
    float bh = period * 0.5f;
    int row = (int)std::floor(b / bh);
    float aa = a + ((row & 1) ? period * 0.5f : 0.f);
    int col = (int)std::floor(aa / period);
    float fa = aa / period - (float)col, fb = b / bh - (float)row;
    bool mortar = fa < 0.07f || fb < 0.14f;
    base = mortar ? 205.f : 92.f + 70.f * u01(hash3(salt, (uint64_t)(int64_t)row, (uint64_t)(int64_t)col));
  } else {  // stripes crossed with sparse tiles
    int ia = (int)std::floor(a / (period * 0.5f));
    int ib = (int)std::floor(b / (period * 2.0f));
    base = (ia & 1) ? 150.f : 96.f;
    if (u01(hash3(salt ^ 77, (uint64_t)(int64_t)(ia >> 1), (uint64_t)(int64_t)ib)) < 0.35f) base = (ia & 1) ? 60.f : 215.f;
  }
  base += 44.f * (vnoise(a * 9.f, b * 9.f, salt ^ 0xA5) - 0.5f);
  base += 20.f * (vnoise(a * 37.f, b * 37.f, salt ^ 0x5A) - 0.5f);
  return base;
}

struct Hit { float t; V3 n; int id; float a, b; };

// slab test; inside != 0 returns the exit face (room shell), else the entry face
bool hit_box(const Box& bx, V3 o, V3 d, bool inside, Hit& h) {
  float tn = -1e30f, tf = 1e30f;
  int an = -1, af = -1;
  const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  const float lo[3] = {bx.lo.x, bx.lo.y, bx.lo.z}, hi[3] = {bx.hi.x, bx.hi.y, bx.hi.z};
  for (int k = 0; k < 3; ++k) {
    if (dd[k] == 0.f) {
      if (oo[k] < lo[k] || oo[k] > hi[k]) return false;
      continue;
    }
    float t0 = (lo[k] - oo[k]) / dd[k], t1 = (hi[k] - oo[k]) / dd[k];
    if (t0 > t1) { float t = t0; t0 = t1; t1 = t; }
    if (t0 > tn) { tn = t0; an = k; }
    if (t1 < tf) { tf = t1; af = k; }
  }
  if (tn > tf || tf <= 0.f) return false;
  float t; int ax;
  if (inside) { t = tf; ax = af; } else { if (tn <= 0.f) return false; t = tn; ax = an; }
  V3 p = o + d * t;
  const float pp[3] = {p.x, p.y, p.z};
  float sgn = (dd[ax] > 0.f) ? 1.f : -1.f;     // side of the slab that was crossed
  int face = ax * 2 + ((sgn > 0.f) == inside ? 1 : 0);
  h.t = t;
  h.n = {ax == 0 ? -sgn : 0.f, ax == 1 ? -sgn : 0.f, ax == 2 ? -sgn : 0.f};
  h.id = bx.id * 6 + face;
  h.a = pp[(ax + 1) % 3]; h.b = pp[(ax + 2) % 3];
  return true;
}

bool hit_pillar(const Pillar& c, V3 o, V3 d, float floor_y, Hit& h) {
  float ox = o.x - c.cx, oz = o.z - c.cz;
  float A = d.x * d.x + d.z * d.z;
  if (A == 0.f) return false;
  float B = ox * d.x + oz * d.z, Cc = ox * ox + oz * oz - c.r * c.r;
  float disc = B * B - A * Cc;
  if (disc <= 0.f) return false;
  float t = (-B - std::sqrt(disc)) / A;
  if (t <= 0.f) return false;
  V3 p = o + d * t;
  if (p.y < c.ytop || p.y > floor_y) return false;
  h.t = t;
  h.n = {(p.x - c.cx) / c.r, 0.f, (p.z - c.cz) / c.r};
  h.id = c.id * 6;
  // arc-length-like coordinate without trig: use x offset scaled (monotone per half)
  h.a = (p.x - c.cx) * 1.6f + (p.z > c.cz ? 3.f : 0.f); h.b = p.y;
  return true;
}

}  // namespace

extern "C" int drfe_synth_frame(int width, int height, int scene, uint32_t seed,
                                float depth_unit_scale, uint8_t* gray, float* depth, float* fx,
                                float* fy, float* cx, float* cy) {
  if (width < 32 || height < 32 || !gray || !depth) return -1;
  const float f = 525.0f * (float)width / 640.0f;
  const float pcx = ((float)width - 1.f) * 0.5f, pcy = ((float)height - 1.f) * 0.5f;
  if (fx) *fx = f;
  if (fy) *fy = f;
  if (cx) *cx = pcx;
  if (cy) *cy = pcy;

  // ---- scene
  std::vector<Box> boxes;
  std::vector<Pillar> pillars;
  Box room;
  const float floor_y = 1.35f;
  if (scene == 0) room = {{-1.15f, -1.25f, -1.0f}, {1.15f, floor_y, 13.0f}, 0};
  else room = {{-2.6f, -1.35f, -1.0f}, {2.6f, floor_y, 6.2f}, 0};
  if (scene >= 1) {
    boxes.push_back({{-2.3f, 0.45f, 3.4f}, {-1.1f, floor_y, 4.9f}, 1});
    boxes.push_back({{0.7f, 0.1f, 4.3f}, {2.1f, floor_y, 5.6f}, 2});
    boxes.push_back({{-0.5f, 0.75f, 2.6f}, {0.35f, floor_y, 3.3f}, 3});
  } else {
    boxes.push_back({{0.75f, 0.55f, 7.0f}, {1.15f, floor_y, 7.8f}, 1});   // cabinet against a wall
  }
  if (scene == 2) {
    pillars.push_back({-1.2f, 2.6f, 0.32f, -1.35f, 5});
    pillars.push_back({1.5f, 3.1f, 0.26f, -1.35f, 6});
  }
  // ---- camera pose from the seed: smooth sweep + small jitter, rotation without trig
  const uint32_t k = seed % 256u;
  const float s = (float)k / 256.0f;
  const uint64_t hs = mix64(seed);
  V3 pos, fwd;
  if (scene == 0) {
    pos = {0.25f * (2.f * s - 1.f) + 0.05f * (u01(hs) - 0.5f), 0.04f * (u01(hs >> 7) - 0.5f), 0.2f + 5.5f * s};
    fwd = {0.22f * (1.f - 2.f * s) + 0.03f * (u01(hs >> 13) - 0.5f), 0.05f + 0.02f * (u01(hs >> 19) - 0.5f), 1.f};
  } else {
    pos = {1.2f * (2.f * s - 1.f), 0.03f * (u01(hs) - 0.5f), 0.1f + 0.9f * s};
    fwd = {-0.45f * (2.f * s - 1.f) + 0.03f * (u01(hs >> 13) - 0.5f), 0.10f + 0.02f * (u01(hs >> 19) - 0.5f), 1.f};
  }
  fwd = norm(fwd);
  V3 right = norm(cross({0.f, 1.f, 0.f}, fwd));
  V3 down = cross(fwd, right);

  std::vector<float> lum((size_t)width * height);
  for (int v = 0; v < height; ++v)
    for (int u = 0; u < width; ++u) {
      const float xn = ((float)u - pcx) / f, yn = ((float)v - pcy) / f;
      V3 d = right * xn + down * yn + fwd;  // camera-frame z component is exactly 1 => depth z = t
      Hit best{}, h{};
      hit_box(room, pos, d, true, best);
      for (const Box& b : boxes)
        if (hit_box(b, pos, d, false, h) && h.t < best.t) best = h;
      for (const Pillar& c : pillars)
        if (hit_pillar(c, pos, d, floor_y, h) && h.t < best.t) best = h;
      const size_t idx = (size_t)v * width + u;
      lum[idx] = texture(best.id + 13 * scene, best.a, best.b);
      // shading: mild dependence on the incidence angle
      const float cosi = -dot(best.n, d) / std::sqrt(dot(d, d));
      lum[idx] *= 0.78f + 0.22f * cosi;
      // depth with Kinect-like noise, dropouts, u16 x5000 quantisation (TUM convention)
      float z = best.t;
      const uint64_t hp = hash3(seed, idx, 0xDEADull);
      z += 1.425e-3f * z * z * gauss(hp);
      bool drop = u01(mix64(hp)) < 0.02f || (cosi < 0.12f && u01(mix64(hp ^ 0x55)) < 0.7f);
      float q = std::floor(z * 5000.0f + 0.5f);
      if (drop || q < 1.f || q > 65535.f) q = 0.f;
      depth[idx] = q * (1.0f / 5000.0f) * depth_unit_scale;
    }
  // mild [1 2 1]^2 blur (clamped borders) + sensor noise
  std::vector<float> tmp((size_t)width * height);
  for (int v = 0; v < height; ++v)
    for (int u = 0; u < width; ++u) {
      int ul = u > 0 ? u - 1 : 0, ur = u < width - 1 ? u + 1 : u;
      const float* r = &lum[(size_t)v * width];
      tmp[(size_t)v * width + u] = 0.25f * r[ul] + 0.5f * r[u] + 0.25f * r[ur];
    }
  for (int v = 0; v < height; ++v)
    for (int u = 0; u < width; ++u) {
      int vu = v > 0 ? v - 1 : 0, vd = v < height - 1 ? v + 1 : v;
      float val = 0.25f * tmp[(size_t)vu * width + u] + 0.5f * tmp[(size_t)v * width + u] +
                  0.25f * tmp[(size_t)vd * width + u];
      val += 2.0f * gauss(hash3(seed, (uint64_t)v * width + u, 0xBEEFull));
      float r = std::floor(val + 0.5f);
      gray[(size_t)v * width + u] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
    }
  return 0;
}
