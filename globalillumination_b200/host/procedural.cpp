// procedural.cpp — see procedural.h.
#include "procedural.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>

namespace sgh {
namespace {

struct Builder {
  std::vector<float> v;
  std::vector<int> f;
  uint32_t rng;
  explicit Builder(uint32_t seed) : rng(seed * 2654435761u + 12345u) {}
  float rnd() { rng = rng * 1664525u + 1013904223u; return (float)((rng >> 8) & 0xFFFF) / 65536.0f; }   // [0,1)
  int vert(float x, float y, float z) { v.push_back(x); v.push_back(y); v.push_back(z); return (int)v.size() / 3 - 1; }
  void tri(int a, int b, int c) { f.push_back(a); f.push_back(b); f.push_back(c); }
  void quad(int a, int b, int c, int d) { tri(a, b, c); tri(a, c, d); }

  // planar grid: origin o, edge vectors u and w, nu x nw cells
  void grid(const float o[3], const float u[3], const float w[3], int nu, int nw) {
    int base = (int)v.size() / 3;
    for (int j = 0; j <= nw; j++)
      for (int i = 0; i <= nu; i++) {
        float a = (float)i / nu, b = (float)j / nw;
        vert(o[0] + a * u[0] + b * w[0], o[1] + a * u[1] + b * w[1], o[2] + a * u[2] + b * w[2]);
      }
    for (int j = 0; j < nw; j++)
      for (int i = 0; i < nu; i++) {
        int p = base + j * (nu + 1) + i;
        quad(p, p + 1, p + nu + 2, p + nu + 1);
      }
  }
  void box(float x0, float y0, float z0, float x1, float y1, float z1, int n) {
    float o[3], u[3], w[3];
    auto G = [&](float ox, float oy, float oz, float ux, float uy, float uz, float wx, float wy, float wz, int a, int b) {
      o[0] = ox; o[1] = oy; o[2] = oz; u[0] = ux; u[1] = uy; u[2] = uz; w[0] = wx; w[1] = wy; w[2] = wz; grid(o, u, w, a, b);
    };
    float dx = x1 - x0, dy = y1 - y0, dz = z1 - z0;
    G(x0, y0, z0, dx, 0, 0, 0, 0, dz, n, n); G(x0, y1, z0, 0, 0, dz, dx, 0, 0, n, n);
    G(x0, y0, z0, 0, dy, 0, dx, 0, 0, n, n); G(x0, y0, z1, dx, 0, 0, 0, dy, 0, n, n);
    G(x0, y0, z0, 0, 0, dz, 0, dy, 0, n, n); G(x1, y0, z0, 0, dy, 0, 0, 0, dz, n, n);
  }
  // vertical column with a slightly bulged profile, closed by caps
  void column(float cx, float cz, float y0, float y1, float r, int seg, int rings) {
    int base = (int)v.size() / 3;
    for (int j = 0; j <= rings; j++) {
      float t = (float)j / rings;
      float rr = r * (1.0f + 0.12f * std::sin(3.14159265f * t)) * ((j == 0 || j == rings) ? 1.35f : 1.0f);
      for (int i = 0; i < seg; i++) {
        float a = 6.2831853f * (float)i / seg;
        vert(cx + rr * std::cos(a), y0 + t * (y1 - y0), cz + rr * std::sin(a));
      }
    }
    for (int j = 0; j < rings; j++)
      for (int i = 0; i < seg; i++) {
        int p = base + j * seg + i, q = base + j * seg + (i + 1) % seg;
        quad(p, q, q + seg, p + seg);
      }
    int cb = vert(cx, y0, cz), ct = vert(cx, y1, cz);
    for (int i = 0; i < seg; i++) {
      tri(cb, base + (i + 1) % seg, base + i);
      tri(ct, base + rings * seg + i, base + rings * seg + (i + 1) % seg);
    }
  }
  // semicircular arch band between two points on a line parallel to x (dir=0) or z (dir=1)
  void arch(float ax, float az, float bx, float bz, float y, float thick, float depth, int seg) {
    float mx = 0.5f * (ax + bx), mz = 0.5f * (az + bz);
    float hx = 0.5f * (bx - ax), hz = 0.5f * (bz - az);
    float R = std::sqrt(hx * hx + hz * hz), ux = hx / R, uz = hz / R;
    float nx = -uz, nz = ux;                 // horizontal normal of the arch plane
    int base = (int)v.size() / 3;
    for (int i = 0; i <= seg; i++) {
      float a = 3.14159265f * (float)i / seg, c = std::cos(a), s = std::sin(a);
      for (int k = 0; k < 2; k++) {
        float rr = (k == 0) ? R : R + thick;
        for (int d = 0; d < 2; d++) {
          float off = (d == 0) ? -0.5f * depth : 0.5f * depth;
          vert(mx - ux * rr * c + nx * off, y + rr * s, mz - uz * rr * c + nz * off);
        }
      }
    }
    for (int i = 0; i < seg; i++) {
      int p = base + 4 * i, q = p + 4;
      quad(p + 0, q + 0, q + 1, p + 1);       // intrados
      quad(p + 2, p + 3, q + 3, q + 2);       // extrados
      quad(p + 0, p + 2, q + 2, q + 0);       // front face
      quad(p + 1, q + 1, q + 3, p + 3);       // back face
    }
  }
  void sphere(float cx, float cy, float cz, float r, int seg, int rings) {
    int base = (int)v.size() / 3;
    for (int j = 0; j <= rings; j++) {
      float th = 3.14159265f * (float)j / rings;
      for (int i = 0; i < seg; i++) {
        float ph = 6.2831853f * (float)i / seg;
        vert(cx + r * std::sin(th) * std::cos(ph), cy + r * std::cos(th), cz + r * std::sin(th) * std::sin(ph));
      }
    }
    for (int j = 0; j < rings; j++)
      for (int i = 0; i < seg; i++) {
        int p = base + j * seg + i, q = base + j * seg + (i + 1) % seg;
        if (j > 0) tri(p, q, q + seg);
        if (j < rings - 1) tri(p, q + seg, p + seg);
      }
  }
};

// A two-storey arcaded atrium with an open roof, in the model-space proportions of the Dabrovic Sponza the
// reference's Configs/Sponza.txt loads (about 35 x 16 x 15 units, long axis x, y up, floor near y = 0), so the
// config's scale 2 / rotate-y 90 / translate -30 and its camera and light land inside it the way they do in the
// original.  ~66 k triangles at detail 1.
void sponza_like(Builder& b, int detail) {
  const float L = 17.5f, Wd = 7.5f, y0 = 0.0f, yM = 6.2f, yT = 13.0f;      // half length, half width, storeys
  const float cl = 12.5f, cw = 3.6f;                                        // courtyard half extents
  float o[3], u[3], w[3];
  auto G = [&](float ox, float oy, float oz, float ux, float uy, float uz, float wx, float wy, float wz, int a, int c) {
    o[0] = ox; o[1] = oy; o[2] = oz; u[0] = ux; u[1] = uy; u[2] = uz; w[0] = wx; w[1] = wy; w[2] = wz; b.grid(o, u, w, a * detail, c * detail);
  };
  G(-L, y0, -Wd, 0, 0, 2 * Wd, 2 * L, 0, 0, 30, 70);                        // ground floor
  G(-L, y0, -Wd, 2 * L, 0, 0, 0, yT + 2, 0, 70, 30);                        // long walls
  G(-L, y0, Wd, 0, yT + 2, 0, 2 * L, 0, 0, 30, 70);
  G(-L, y0, -Wd, 0, yT + 2, 0, 0, 0, 2 * Wd, 30, 30);                       // end walls
  G(L, y0, -Wd, 0, 0, 2 * Wd, 0, yT + 2, 0, 30, 30);
  // upper gallery slab and roof ring around the open courtyard (4 strips each)
  for (int lvl = 0; lvl < 2; lvl++) {
    float y = lvl == 0 ? yM : yT, t = 0.35f;
    b.box(-L, y, -Wd, L, y + t, -cw, 10 * detail);
    b.box(-L, y, cw, L, y + t, Wd, 10 * detail);
    b.box(-L, y, -cw, -cl, y + t, cw, 6 * detail);
    b.box(cl, y, -cw, L, y + t, cw, 6 * detail);
  }
  // colonnades on both storeys along the courtyard edge, with arches between neighbours
  const int nlong = 11, nshort = 3, seg = 24 * detail, rings = 12 * detail, aseg = 18 * detail;
  for (int lvl = 0; lvl < 2; lvl++) {
    float ya = lvl == 0 ? y0 : yM + 0.35f, yb = lvl == 0 ? yM - 1.6f : yT - 1.4f;
    float r = lvl == 0 ? 0.42f : 0.30f;
    for (int side = -1; side <= 1; side += 2) {
      float px = 0, pz = 0;
      for (int i = 0; i < nlong; i++) {
        float x = -cl + 2 * cl * (float)i / (nlong - 1), z = side * cw;
        b.column(x, z, ya, yb, r * (0.95f + 0.1f * b.rnd()), seg, rings);
        if (i > 0) b.arch(px, pz, x, z, yb, 0.45f, 0.7f, aseg);
        px = x; pz = z;
      }
      for (int i = 1; i < nshort - 1 + 1; i++) {
        float z = -cw + 2 * cw * (float)i / nshort, x = side * cl;
        if (i < nshort) b.column(x, z, ya, yb, r * (0.95f + 0.1f * b.rnd()), seg, rings);
      }
      b.arch(side * cl, -cw, side * cl, -cw + 2 * cw / nshort, yb, 0.45f, 0.7f, aseg);
      b.arch(side * cl, -cw + 2 * cw / nshort, side * cl, -cw + 4 * cw / nshort, yb, 0.45f, 0.7f, aseg);
      b.arch(side * cl, -cw + 4 * cw / nshort, side * cl, cw, yb, 0.45f, 0.7f, aseg);
    }
  }
  // a few free-standing props in the courtyard (occluders at different heights), jittered by the seed
  for (int k = 0; k < 6; k++) {
    float x = -9.0f + 3.6f * k + 0.8f * (b.rnd() - 0.5f), z = 1.6f * (b.rnd() - 0.5f);
    b.sphere(x, 0.9f + 0.3f * b.rnd(), z, 0.8f, 28 * detail, 18 * detail);
    b.box(x - 0.5f, 0.0f, z + 1.2f, x + 0.5f, 1.6f + b.rnd(), z + 2.0f, 3 * detail);
  }
  // hanging banners between the upper columns
  for (int k = 0; k < 5; k++) {
    float x = -8.0f + 4.0f * k;
    G(x - 0.7f, yM + 0.6f, -0.4f + 0.8f * b.rnd(), 1.4f, 0, 0, 0, 4.5f, 0.2f, 4, 10);
  }
}

int query_int(const std::string& spec, const char* key, int dflt) {
  size_t p = spec.find(std::string(key) + "=");
  if (p == std::string::npos) return dflt;
  return atoi(spec.c_str() + p + strlen(key) + 1);
}

}  // namespace

bool makeProcedural(const std::string& spec, Mesh* out, std::string* err) {
  std::string name = spec.substr(0, spec.find('?'));
  int seed = query_int(spec, "seed", 1);
  Builder b((uint32_t)seed);
  if (name == "sponza_like") {
    sponza_like(b, query_int(spec, "detail", 1));
  } else if (name == "sphere") {
    int seg = query_int(spec, "seg", 128);
    b.sphere(0.0f, 1.0f, 0.0f, 1.0f, seg, seg / 2);
  } else if (name == "leaves") {
    int n = query_int(spec, "n", 6000);
    for (int k = 0; k < n; k++) {                 // small random quads in a crown-shaped volume
      float th = 6.2831853f * b.rnd(), rr = 2.6f * std::sqrt(b.rnd()), y = 4.0f + 3.5f * b.rnd();
      float x = rr * std::cos(th), z = rr * std::sin(th), s = 0.10f + 0.08f * b.rnd();
      float ax = b.rnd() - 0.5f, ay = b.rnd() - 0.5f, az = b.rnd() - 0.5f;
      int p0 = b.vert(x - s, y - s * ay, z - s * az), p1 = b.vert(x + s, y + s * ax, z - s * az);
      int p2 = b.vert(x + s, y + s * ay, z + s * az), p3 = b.vert(x - s, y - s * ax, z + s * az);
      b.quad(p0, p1, p2, p3);
    }
  } else if (name == "plane") {
    int a = b.vert(-1, 1, -1), c = b.vert(1, 1, -1), d = b.vert(1, 1, 1), e = b.vert(-1, 1, 1);
    b.tri(a, c, d); b.tri(a, d, e);
  } else {
    if (err) *err = "unknown procedural scene \"" + name + "\"";
    return false;
  }
  out->setGeometry(b.v.data(), (int)b.v.size() / 3, b.f.data(), (int)b.f.size() / 3);
  return true;
}

bool substituteMissingAsset(const std::string& path, Mesh* out) {
  auto ends = [&](const char* s) { std::string t(s); return path.size() >= t.size() && path.compare(path.size() - t.size(), t.size(), t) == 0; };
  std::string e;
  if (ends("Sponza/sponza.obj")) return makeProcedural("sponza_like?seed=1", out, &e);
  if (ends("SanDiego/sphere.obj") || ends("Sphere/sphere.obj")) return makeProcedural("sphere?seed=2", out, &e);
  if (ends("TreeWithLeaves/TreeSub1.obj")) return makeProcedural("leaves?seed=3", out, &e);
  return false;
}

}  // namespace sgh
