// sgi_rbssm.cuh — revectorization-based soft shadow mapping (RBSSM), device side.  Included by sgi_shadow.cu after the
// RBSM helpers it builds on (Rb, getdisc4, nc_getdisc_f, nc_getdisc_v: RBSSM.frag:37-186 is the same code as
// NonConservativeSMSR.frag:23-178 with the break flags off).
//
// Replaces SoftShadowMapping/Shaders/SoftShadow/RBSSM.frag (renderSoftShadows with shadowParams.RBSSM):
//   :193-253 computeDiscontinuityLength -> ss_disc_length     :255-280 normalizeDiscontinuitySpace -> ss_normalize
//   :311-1117 smoothONDS -> ss_fill (entering / exiting halves)  :1137-1147 computeDiscontinuity -> ss_classify
//   :1202-1339 RBSSM + :1341-1358 revectorizationBasedShadowMappingSmoothing -> ss_rbssm (blocker search and penumbra
//   width are the PCSS arithmetic, shared with pcss_blockers / pcss_penumbra)
// Discontinuity code: d[0] = x-axis (0 none, .5 left, .25 right, .75 both), d[1] = y-axis (.5 bottom(+y), .25 top(-y),
// .75 both), d[2] = 1 if the centre sample is in shadow ("exiting"), 0 if lit ("entering").  Every expression keeps the
// shader's evaluation order (-fmad=false), so the result equals the CPU oracle bit for bit.
#pragma once

__device__ __forceinline__ float ss_clamp01(float v) { return g_min(g_max(v, 0.0f), 1.0f); }
__device__ __forceinline__ float ss_compress(float n) { return -2.0f - ((0.5f - n) * 2.0f); }                 // :25-29
__device__ __forceinline__ float ss_decompress(float n) { return (0.5f - ((n + 2.0f) * -1.0f) / 2.0f); }      // :31-35
// mix(a, b, step(edge, x)): a when x < edge, else b  (:304-309)
#define SS_PICK(a, b, edge, x) (((x) < (edge)) ? (a) : (b))

// :193-253
__device__ __noinline__ float ss_disc_length(Rb& r, const float d[4], float4 lightCoord, float dx, float dy) {
  const float thr = r.thr;
  float4 c = lightCoord;
  float foundEdgeEnd = 0.0f;
  if (dx == 0.0f && d[0] == 0.0f && d[1] != 0.0f) return -1.0f;
  if (dy == 0.0f && d[0] != 0.0f && d[1] == 0.0f) return -1.0f;
  if (((0.5f - d[0]) * 8.0f - 1.0f) == dx) return 1.0f;
  if (((d[1] - 0.5f) * 8.0f + 1.0f) == dy) return 1.0f;
  float dist = 1.0f;
  const float stx = dx * r.sx, sty = dy * r.sy;
  c.x += stx; c.y += sty;
  for (int it = 0; it < r.max_search; it++) {
    const float dfl = sm_fetch(r.s, c.x, c.y);
    if (d[2] == 0.0f)
      if (fabsf(c.z - dfl) < thr) c.z -= thr;
    const float center = (c.z <= dfl) ? 1.0f : 0.0f;
    if (fabsf(center - d[2]) == 0.0f) {
      foundEdgeEnd = nc_getdisc_f(r, c, 0.0f, 0.0f, d[2]) ? 1.0f : 0.0f;
      break;
    } else {
      if (!nc_getdisc_v(r, c, dx, dy, d[0], d[1], d[2])) break;
    }
    dist += 1.0f;
    c.x += stx; c.y += sty;
    if (d[2] == 1.0f) c.z = r.newDepth;
  }
  return g_mix(-dist, dist, foundEdgeEnd);
}

// :255-280
__device__ __noinline__ float ss_normalize(Rb& r, float ax, float ay, float sub) {
  if (ax < 0.0f && ay < 0.0f) return -1.0f;
  const float edgeLength = g_min(fabsf(ax) + fabsf(ay) - 1.0f, (float)r.max_search);
  const float mx = g_max(ax, ay);
  float n = 1.0f - mx / edgeLength;
  if (ax == ay) n += g_mix(sub / edgeLength, (1.0f - sub) / edgeLength, (sub < 0.5f) ? 0.0f : 1.0f);
  else if (ax == mx) n += (1.0f - sub) / edgeLength;
  else n += sub / edgeLength;
  if (ax > 0.0f && ay > 0.0f) return ss_compress(n);
  return n;
}

// getDisc(coord, dir, vec4(cr, cg, 0.0, type)) with a literal code (e.g. RBSSM.frag:397): only the x- or y-side named by
// the code is examined.  The shader passes the type in .a, but getDisc(vec4,vec2,vec4) tests discType.b, which these
// literals leave at 0: the call answers "is that side neighbour in shadow".
__device__ __forceinline__ int ss_side(Rb& r, float4 c, float dx, float dy, float cr, float cg, float type) {
  (void)type;
  return nc_getdisc_v(r, c, dx, dy, cr, cg, 0.0f);
}

// entering discontinuity (centre lit, d[2] == 0): :314-704
__device__ __noinline__ int ss_fill_entering(Rb& r, float4 lc, float nx, float ny, float d[4], float subx, float suby, float* out) {
  const float sx = r.sx, sy = r.sy;
  if (nx <= -2.0f && ny <= -2.0f && (d[0] == 0.75f || d[1] == 0.75f)) {                         // :317-427
    int left = 1, right = 1, bottom = 1, top = 1;
    if (d[0] == 0.75f) {
      if (d[1] == 0.0f) {
        lc.y += sy;
        top = nc_getdisc_f(r, lc, 0.0f, 1.0f, d[2]);
        lc.y -= 2.0f * sy;
        bottom = nc_getdisc_f(r, lc, 0.0f, 1.0f, d[2]);
        lc.y += sy;
        if (top && bottom) { *out = 0.0f; return 1; }
        lc.y += top ? sy : -sy;
        left = ss_side(r, lc, 0.0f, 1.0f, 0.5f, 0.0f, 0.0f);
        right = !left;
      } else {
        if (d[1] == 0.5f) top = 0;
        else if (d[1] == 0.25f) bottom = 0;
        lc.y += ((0.5f - d[1]) * 8.0f - 1.0f) * sy;
        left = ss_side(r, lc, 0.0f, 1.0f, 0.5f, 0.0f, d[2]);
        right = ss_side(r, lc, 0.0f, 1.0f, 0.25f, 0.0f, d[2]);
        if (left && right) {
          if (ny <= -2.0f) ny = ss_decompress(ny) + 0.5f;
          const float a = ss_clamp01(subx - (ny - 1.0f));
          const float b = ss_clamp01((1.0f - subx) - (ny - 1.0f));
          const float c = SS_PICK(1.0f - suby, suby, d[1], 1.0f);
          *out = g_min(g_min(a, b), c);
          return 1;
        }
      }
    }
    if (d[1] == 0.75f) {
      if (d[0] == 0.0f) {
        lc.x -= sx;
        left = nc_getdisc_f(r, lc, 1.0f, 0.0f, d[2]);
        lc.x += 2.0f * sx;
        right = nc_getdisc_f(r, lc, 1.0f, 0.0f, d[2]);
        lc.x -= sx;
        if (left && right) { *out = 0.0f; return 1; }
        lc.x += left ? -sx : sx;
        bottom = ss_side(r, lc, 0.0f, 0.25f, 0.5f, 0.0f, d[2]);      // :376 (dir.y = 0.25: neither axis is scanned)
        top = !bottom;
      } else {
        if (d[0] == 0.5f) right = 0;
        else if (d[0] == 0.25f) left = 0;
        lc.x -= ((0.5f - d[0]) * 8.0f - 1.0f) * sx;
        bottom = ss_side(r, lc, 1.0f, 0.0f, 0.0f, 0.5f, d[2]);
        top = ss_side(r, lc, 1.0f, 0.0f, 0.0f, 0.25f, d[2]);
        if (bottom && top) {
          if (nx <= -2.0f) nx = ss_decompress(nx) + 0.5f;
          const float a = ss_clamp01(suby - (nx - 1.0f));
          const float b = ss_clamp01((1.0f - suby) - (nx - 1.0f));
          const float c = SS_PICK(1.0f - subx, subx, d[0], 1.0f);
          *out = g_min(g_min(a, b), c);
          return 1;
        }
      }
    }
    if (!left && !bottom) { *out = ss_clamp01((1.0f - subx) - (1.0f - suby)); return 1; }
    else if (!right && !bottom) { *out = ss_clamp01(subx - (1.0f - suby)); return 1; }
    else if (!left && !top) { *out = ss_clamp01((1.0f - suby) - subx); return 1; }
    else if (!right && !top) { *out = ss_clamp01((1.0f - suby) - (1.0f - subx)); return 1; }
  }

  if (d[0] == 0.75f || d[1] == 0.75f) {                                                          // :429-640
    if (d[0] == 0.75f && d[1] != 0.0f) {                                                         // :432-459
      lc.y += ((d[1] - 0.75f) * 4.0f) * sy;
      const int left = ss_side(r, lc, 0.0f, 1.0f, 0.5f, 0.0f, d[2]);
      const int right = ss_side(r, lc, 0.0f, 1.0f, 0.25f, 0.0f, d[2]);
      if (!left && !right) { *out = ss_clamp01(1.0f - ny); return 1; }
      const float sub = SS_PICK(1.0f - suby, suby, 1.0f, d[1]);
      const float a = right ? ss_clamp01((1.0f - subx) - (ny - 1.0f)) : sub;
      const float b = left ? ss_clamp01(subx - (ny - 1.0f)) : sub;
      *out = g_min(a, b);
      return 1;
    }
    if (d[0] == 0.75f && d[1] == 0.0f) {                                                         // :462-546
      int topLeft, topRight, bottomLeft, bottomRight, topCenter, bottomCenter;
      float4 tc = lc, bc = lc;
      tc.y += sy;
      topCenter = !(tc.z <= sm_fetch(r.s, tc.x, tc.y));
      topLeft = ss_side(r, tc, 0.0f, 1.0f, 0.5f, 0.0f, d[2]);
      topRight = ss_side(r, tc, 0.0f, 1.0f, 0.25f, 0.0f, d[2]);
      bc.y -= sy;
      bottomCenter = !(bc.z <= sm_fetch(r.s, bc.x, bc.y));
      bottomLeft = ss_side(r, bc, 0.0f, 1.0f, 0.5f, 0.0f, d[2]);
      bottomRight = ss_side(r, bc, 0.0f, 1.0f, 0.25f, 0.0f, d[2]);
      if (topCenter) { topLeft = 1; topRight = 1; }
      if (bottomCenter) { bottomLeft = 1; bottomRight = 1; }
      if (ny <= -2.0f) {
        ny = ss_decompress(ny) + 0.5f;
        const float a = ss_clamp01(subx - (ny - 1.0f));
        const float b = ss_clamp01((1.0f - subx) - (ny - 1.0f));
        if (!bottomRight || !topRight) *out = a;
        else if (!bottomLeft || !topLeft) *out = b;
        else *out = g_min(a, b);
        return 1;
      }
      if ((!bottomRight && !bottomLeft) || (!topRight && !topLeft)) *out = ss_clamp01(1.0f - ny);
      else if (!bottomRight || !topRight) *out = ss_clamp01(subx - (ny - 1.0f));
      else *out = ss_clamp01((1.0f - subx) - (ny - 1.0f));
      return 1;
    }
    if (d[0] != 0.0f && d[1] == 0.75f) {                                                         // :549-577
      lc.x -= ((0.5f - d[0]) * 8.0f - 1.0f) * sx;
      const int bottom = ss_side(r, lc, 1.0f, 0.0f, 0.0f, 0.5f, d[2]);
      const int top = ss_side(r, lc, 1.0f, 0.0f, 0.0f, 0.25f, d[2]);
      if (!bottom && !top) { *out = ss_clamp01(1.0f - nx); return 1; }
      const float sub = SS_PICK(subx, 1.0f - subx, d[0], 0.25f);
      const float a = top ? ss_clamp01(suby - (nx - 1.0f)) : sub;
      const float b = bottom ? ss_clamp01((1.0f - suby) - (nx - 1.0f)) : sub;
      *out = g_min(a, b);
      return 1;
    }
    if (d[0] == 0.0f && d[1] == 0.75f) {                                                         // :580-638
      int topLeft, topRight, bottomLeft, bottomRight, leftCenter, rightCenter;
      float4 lcl = lc, lcr = lc;
      lcr.x += sx;
      rightCenter = !(lcr.z <= sm_fetch(r.s, lcr.x, lcr.y));
      bottomRight = ss_side(r, lcr, 1.0f, 0.0f, 0.0f, 0.5f, d[2]);
      topRight = ss_side(r, lcr, 1.0f, 0.0f, 0.0f, 0.25f, d[2]);
      lcl.x -= sx;
      leftCenter = !(lcl.z <= sm_fetch(r.s, lcl.x, lcl.y));
      bottomLeft = ss_side(r, lcl, 1.0f, 0.0f, 0.0f, 0.5f, d[2]);
      topLeft = ss_side(r, lcl, 1.0f, 0.0f, 0.0f, 0.25f, d[2]);
      if (rightCenter) { bottomRight = 1; topRight = 1; }
      if (leftCenter) { topLeft = 1; bottomLeft = 1; }
      if (nx <= -2.0f) {
        nx = ss_decompress(nx) + 0.5f;
        const float a = ss_clamp01(suby - (nx - 1.0f));
        const float b = ss_clamp01((1.0f - suby) - (nx - 1.0f));
        if (!bottomRight || !bottomLeft) *out = a;
        else if (!topRight || !topLeft) *out = b;
        else *out = g_min(a, b);
        return 1;
      }
      if ((!bottomRight && !topRight) || (!bottomLeft && !topLeft)) *out = ss_clamp01(1.0f - nx);
      else if (!bottomRight || !bottomLeft) *out = ss_clamp01(suby - (nx - 1.0f));
      else *out = ss_clamp01((1.0f - suby) - (nx - 1.0f));
      return 1;
    }
  }

  if (d[0] > 0.0f && d[1] > 0.0f) {                                                              // corner, :643-702
    lc.x -= ((0.5f - d[0]) * 8.0f - 1.0f) * sx;
    const int horizontal = nc_getdisc_f(r, lc, 1.0f, 0.0f, 0.0f);
    lc.x += ((0.5f - d[0]) * 8.0f - 1.0f) * sx;
    lc.y += ((0.5f - d[1]) * 8.0f - 1.0f) * sy;
    const int vertical = nc_getdisc_f(r, lc, 0.0f, 1.0f, 0.0f);
    if (horizontal && !vertical) d[0] = 0.0f;
    else if (!horizontal && vertical) d[1] = 0.0f;
    else if (!horizontal && !vertical) {
      if (d[1] == 0.5f) { *out = ss_clamp01((1.0f - nx) - (suby - 1.0f)); return 1; }
      else if (d[1] == 0.25f) { *out = ss_clamp01((1.0f - nx) + suby); return 1; }
    } else {
      float a, b;
      if (d[0] == 0.5f && d[1] == 0.5f) {
        a = SS_PICK(1.0f - suby, ss_clamp01(1.0f - (nx - (1.0f - suby))), -2.0f, nx);
        b = SS_PICK(subx, ss_clamp01(1.0f - (ny - subx)), -2.0f, ny);
        *out = g_min(a, b); return 1;
      } else if (d[0] == 0.5f && d[1] == 0.25f) {
        a = SS_PICK(suby, ss_clamp01(1.0f - (nx - suby)), -2.0f, nx);
        b = SS_PICK(subx, ss_clamp01(1.0f - (ny - subx)), -2.0f, ny);
        *out = g_min(a, b); return 1;
      } else if (d[0] == 0.25f && d[1] == 0.5f) {
        a = SS_PICK(1.0f - suby, ss_clamp01(1.0f - (nx - (1.0f - suby))), -2.0f, nx);
        b = SS_PICK(1.0f - subx, ss_clamp01(1.0f - (ny - (1.0f - subx))), -2.0f, ny);
        *out = g_min(a, b); return 1;
      } else if (d[0] == 0.25f && d[1] == 0.25f) {
        a = SS_PICK(suby, ss_clamp01(1.0f - (nx - suby)), -2.0f, nx);
        b = SS_PICK(1.0f - subx, ss_clamp01(1.0f - (ny - (1.0f - subx))), -2.0f, ny);
        *out = g_min(a, b); return 1;
      }
    }
  }
  if (nx <= -2.0f) { *out = SS_PICK(1.0f - suby, suby, d[1], 0.25f); return 1; }                 // :706
  if (ny <= -2.0f) { *out = SS_PICK(subx, 1.0f - subx, d[0], 0.25f); return 1; }                 // :709
  if (d[1] > 0.0f) {                                                                             // :712-719
    if (d[1] == 0.5f) *out = ss_clamp01((1.0f - suby) - (nx - 1.0f));
    else *out = ss_clamp01(suby - (nx - 1.0f));
    return 1;
  }
  if (d[0] > 0.0f) {                                                                             // :722-729
    if (d[0] == 0.5f) *out = ss_clamp01(subx - (ny - 1.0f));
    else *out = ss_clamp01((1.0f - subx) - (ny - 1.0f));
    return 1;
  }
  return 0;
}

// exiting discontinuity (centre in shadow, d[2] == 1): :733-1113
__device__ __noinline__ int ss_fill_exiting(Rb& r, float4 lc, float nx, float ny, float d[4], float subx, float suby, float* out) {
  const float sx = r.sx, sy = r.sy;
  if (d[0] == 0.75f || d[1] == 0.75f) {
    if (d[0] == 0.75f && d[1] == 0.0f) {                                                         // :738-822
      float4 rc = lc;
      rc.x = lc.x - sx;
      int left = ss_side(r, rc, 0.0f, 1.0f, 0.5f, 0.0f, d[2]);
      rc.x = lc.x + sx;
      int right = ss_side(r, rc, 0.0f, 1.0f, 0.25f, 0.0f, d[2]);
      if (left && right) { *out = ss_clamp01(ny); return 1; }
      else if (left || right) {
        if (!left) *out = SS_PICK(1.0f - subx, ss_clamp01(ny - subx), -2.0f, ny);
        else *out = SS_PICK(subx, ss_clamp01(ny - (1.0f - subx)), -2.0f, ny);
        return 1;
      } else {
        rc.y = lc.y + 1.0f * sy;                          // count = 0: mult = 1, even
        rc.x = lc.x - sx;
        left = nc_getdisc_f(r, rc, 1.0f, 0.0f, 0.0f);
        rc.x = lc.x + sx;
        right = nc_getdisc_f(r, rc, 1.0f, 0.0f, 0.0f);
        if (left && right) { *out = ss_clamp01(ny); return 1; }
        if (!left) *out = SS_PICK(1.0f - subx, ss_clamp01(ny - subx), -2.0f, ny);
        else *out = SS_PICK(subx, ss_clamp01(ny - (1.0f - subx)), -2.0f, ny);
        return 1;
      }
    }
    if (d[0] == 0.75f && d[1] != 0.0f) {                                                         // :825-889
      lc.y += ((0.5f - d[1]) * 8.0f - 1.0f) * sy;
      int left = ss_side(r, lc, 0.0f, 1.0f, 0.5f, 0.0f, d[2]);
      int right = ss_side(r, lc, 0.0f, 1.0f, 0.25f, 0.0f, d[2]);
      float a = 0.0f, b = 0.0f;
      if (left && right) { *out = ss_clamp01(ny); return 1; }
      else if (left || right) {
        if (left) a = SS_PICK(suby, 1.0f - suby, d[1], 0.25f);
        else a = SS_PICK(1.0f - subx, ss_clamp01(ny - subx), -2.0f, ny);
        if (right) b = SS_PICK(suby, 1.0f - suby, d[1], 0.25f);
        else b = SS_PICK(subx, ss_clamp01(ny - (1.0f - subx)), -2.0f, ny);
        *out = g_max(a, b); return 1;
      } else {
        if (d[1] == 0.75f) { *out = 0.0f; return 1; }
        float4 rc = lc;
        rc.x = lc.x - sx;
        left = nc_getdisc_f(r, rc, 1.0f, 0.0f, 0.0f);
        rc.x = lc.x + sx;
        right = nc_getdisc_f(r, rc, 1.0f, 0.0f, 0.0f);
        if (left && right) { *out = ss_clamp01(ny); return 1; }
        a = left ? subx : 1.0f - subx;
        b = SS_PICK(suby, 1.0f - suby, d[1], 0.25f);
        *out = g_max(a, b); return 1;
      }
    }
    if (d[0] == 0.0f && d[1] == 0.75f) {                                                         // :892-951
      float4 rc = lc;
      rc.y = lc.y - sy;
      int top = nc_getdisc_f(r, rc, 0.0f, 1.0f, 0.0f);
      rc.y = lc.y + sy;
      int bottom = nc_getdisc_f(r, rc, 0.0f, 1.0f, 0.0f);
      if (bottom && top) { *out = ss_clamp01(nx); return 1; }
      else if (bottom || top) {
        if (!top) *out = SS_PICK(1.0f - suby, ss_clamp01(nx - suby), -2.0f, nx);
        else *out = SS_PICK(suby, ss_clamp01(nx - (1.0f - suby)), -2.0f, nx);
        return 1;
      } else {
        rc.x = lc.x + 1.0f * sx;                          // count = 0: mult = 1, even
        rc.y = lc.y - sy;
        top = nc_getdisc_f(r, rc, 0.0f, 1.0f, 0.0f);
        rc.y = lc.y + sy;
        bottom = nc_getdisc_f(r, rc, 0.0f, 1.0f, 0.0f);
        if (top && bottom) { *out = ss_clamp01(nx); return 1; }
        if (!top) *out = SS_PICK(1.0f - suby, ss_clamp01(nx - suby), -2.0f, nx);
        else *out = SS_PICK(suby, ss_clamp01(nx - (1.0f - suby)), -2.0f, nx);
        return 1;
      }
    }
    if (d[0] != 0.0f && d[1] == 0.75f) {                                                         // :954-1021
      lc.x -= ((0.5f - d[0]) * 8.0f - 1.0f) * sx;
      int bottom = ss_side(r, lc, 1.0f, 0.0f, 0.0f, 0.5f, d[2]);
      int top = ss_side(r, lc, 1.0f, 0.0f, 0.0f, 0.25f, d[2]);
      float a = 0.0f, b = 0.0f;
      if (bottom && top) { *out = ss_clamp01(nx); return 1; }
      else if (bottom || top) {
        if (top) a = SS_PICK(1.0f - subx, subx, d[0], 0.25f);
        else a = SS_PICK(1.0f - suby, ss_clamp01(nx - suby), -2.0f, nx);
        if (bottom) b = SS_PICK(1.0f - subx, subx, d[0], 0.25f);
        else b = SS_PICK(suby, ss_clamp01(nx - (1.0f - suby)), -2.0f, nx);
        *out = g_max(a, b); return 1;
      } else {
        if (d[0] == 0.75f) { *out = 0.0f; return 1; }
        float4 rc = lc;
        rc.y = lc.y - sy;
        top = nc_getdisc_f(r, rc, 0.0f, 1.0f, 0.0f);
        rc.y = lc.y + sy;
        bottom = nc_getdisc_f(r, rc, 0.0f, 1.0f, 0.0f);
        if (top && bottom) { *out = ss_clamp01(nx); return 1; }
        a = top ? suby : 1.0f - suby;
        b = SS_PICK(1.0f - subx, subx, d[0], 0.25f);
        *out = g_max(a, b); return 1;
      }
    }
  }
  if (d[0] > 0.0f && d[1] > 0.0f) {                                                              // corner, :1026-1085
    lc.x += ((0.5f - d[0]) * 8.0f - 1.0f) * sx;
    const int horizontal = nc_getdisc_f(r, lc, 1.0f, 0.0f, d[2]);
    lc.x += ((0.5f - d[0]) * 8.0f - 1.0f) * sx;
    lc.y += ((0.5f - d[1]) * 8.0f - 1.0f) * sy;
    const int vertical = nc_getdisc_f(r, lc, 0.0f, 1.0f, d[2]);
    if (horizontal && !vertical) d[0] = 0.0f;
    else if (!horizontal && vertical) d[1] = 0.0f;
    else if (!horizontal && !vertical) {
      if (d[1] == 0.5f) { *out = ss_clamp01(suby - (1.0f - nx)); return 1; }
      else if (d[1] == 0.25f) { *out = ss_clamp01((1.0f - suby) - (1.0f - nx)); return 1; }
    } else {
      float a, b;
      if (d[0] == 0.5f && d[1] == 0.5f) {
        a = SS_PICK(suby, ss_clamp01(nx - (1.0f - suby)), -2.0f, nx);
        b = SS_PICK(1.0f - subx, ss_clamp01(ny - subx), -2.0f, ny);
        *out = g_max(a, b); return 1;
      } else if (d[0] == 0.5f && d[1] == 0.25f) {
        a = SS_PICK(1.0f - suby, ss_clamp01(nx - suby), -2.0f, nx);
        b = SS_PICK(1.0f - subx, ss_clamp01(ny - subx), -2.0f, ny);
        *out = g_max(a, b); return 1;
      } else if (d[0] == 0.25f && d[1] == 0.5f) {
        a = SS_PICK(suby, ss_clamp01(nx - (1.0f - suby)), -2.0f, nx);
        b = SS_PICK(subx, ss_clamp01(ny - (1.0f - subx)), -2.0f, ny);
        *out = g_max(a, b); return 1;
      } else if (d[0] == 0.25f && d[1] == 0.25f) {
        a = SS_PICK(1.0f - suby, ss_clamp01(nx - suby), -2.0f, nx);
        b = SS_PICK(subx, ss_clamp01(ny - (1.0f - subx)), -2.0f, ny);
        *out = g_max(a, b); return 1;
      }
    }
  }
  if (nx <= -2.0f) { *out = SS_PICK(suby, 1.0f - suby, d[1], 0.25f); return 1; }                 // :1088
  if (ny <= -2.0f) { *out = SS_PICK(1.0f - subx, subx, d[0], 0.25f); return 1; }                 // :1091
  if (d[1] > 0.0f) {                                                                             // :1094-1101
    if (d[1] == 0.5f) *out = ss_clamp01(nx - (1.0f - suby));
    else *out = ss_clamp01(nx - suby);
    return 1;
  }
  if (d[0] > 0.0f) {                                                                             // :1104-1111
    if (d[0] == 0.5f) *out = ss_clamp01(ny - subx);
    else *out = ss_clamp01(ny - (1.0f - subx));
    return 1;
  }
  return 0;
}

// smoothONDS, :311-1117
__device__ __noinline__ float ss_fill(Rb& r, float4 lc, float nx, float ny, const float disc[4], float subx, float suby) {
  float d[4] = {disc[0], disc[1], disc[2], disc[3]};
  float out = 0.0f;
  const int done = (d[2] == 0.0f) ? ss_fill_entering(r, lc, nx, ny, d, subx, suby, &out)
                                  : ss_fill_exiting(r, lc, nx, ny, d, subx, suby, &out);
  if (done) return out;
  return 1.0f - d[2];                                                                            // :1115
}

// :1137-1147
__device__ __forceinline__ void ss_classify(Rb& r, float4 c, float dfl, float d[4]) {
  const float center = (c.z <= dfl) ? 1.0f : 0.0f;
  float dir[4];
  getdisc4(r, c, dir);
  const float ex = fabsf(dir[0] - center), ey = fabsf(dir[1] - center), ez = fabsf(dir[2] - center), ew = fabsf(dir[3] - center);
  d[0] = (2.0f * ex + ey) / 4.0f;
  d[1] = (2.0f * ez + ew) / 4.0f;
  d[2] = 1.0f - center;
  d[3] = 1.0f;
}


// :1202-1339 (live code) + :1341-1358
__device__ __noinline__ float ss_rbssm(const VisArgs& a, Rb& r, float4 c) {
  const sgi_params& p = a.p;
  const TapSrc<false> g = {r.s.d, r.s.w, 0, 0};
  const float averageDepth = pcss_blockers<0, false>(a, r.s, g, c);            // :1149-1176
  const float penumbraWidth = pcss_penumbra(p, averageDepth, c.z);             // :1178-1187
  const float stepSize = 2.0f * penumbraWidth / (float)p.kernel_size;          // :1350
  if (stepSize <= 0.0f || stepSize >= 1.0f) return 1.0f;
  float illuminationCount = 0.0f;
  const float fw = ((float)p.kernel_size - 1.0f) * 0.5f;
  for (int h = (int)(-fw); (float)h <= fw; h++)
    for (int w = (int)(-fw); (float)w <= fw; w++) {
      const float4 lc = make_float4(c.x + ((float)w * penumbraWidth) / fw, c.y + ((float)h * penumbraWidth) / fw, c.z, c.w);
      const float subx = g_fract(lc.x * r.SWf), suby = g_fract(lc.y * r.SHf);
      const float dfl = sm_fetch(r.s, lc.x, lc.y);
      float d[4];
      ss_classify(r, lc, dfl, d);
      if (d[0] > 0.0f || d[1] > 0.0f) {
        const float left = ss_disc_length(r, d, lc, -1.0f, 0.0f), right = ss_disc_length(r, d, lc, 1.0f, 0.0f);
        const float down = ss_disc_length(r, d, lc, 0.0f, -1.0f), up = ss_disc_length(r, d, lc, 0.0f, 1.0f);
        const float nx = ss_normalize(r, left, right, subx), ny = ss_normalize(r, down, up, suby);
        float fill = ss_fill(r, lc, nx, ny, d, subx, suby);
        fill = g_mix(fill, 1.0f, p.shadow_intensity);
        illuminationCount += fill;
      } else {
        illuminationCount += (lc.z <= dfl) ? 1.0f : p.shadow_intensity;
      }
    }
  return illuminationCount / (float)(p.kernel_size * p.kernel_size);
}

// ---- two-kernel form used by sgi_shadow_run ------------------------------------------------------------------------
// One thread per pixel wastes most lanes here: only penumbra pixels (a few % of the screen) run the k x k tap loop, and
// inside it only the taps next to a shadow-map discontinuity walk the map.  So:
//   k_rbssm_prepare  one thread per pixel: pre-evaluation, blocker search, penumbra width; pixels that end there write
//                    their visibility, the others are appended (pixel, light-space coordinate, width) to a work list
//   k_rbssm_taps     one WARP per listed pixel, taps dealt to lanes; each tap's contribution is parked in shared memory
//                    and lane 0 adds them up in the shader's loop order, so the sum is the same fp32 value
struct RbssmItem { float4 c; float pw; int pixel; int pad0, pad1; };
#define SGI_RBSSM_CHUNK 256            // taps summed per round (k = 15: 225 taps, one round)

__device__ __forceinline__ float ss_tap(const VisArgs& a, Rb& r, float4 c, float penumbraWidth, float fw, int w, int h) {
  const float4 lc = make_float4(c.x + ((float)w * penumbraWidth) / fw, c.y + ((float)h * penumbraWidth) / fw, c.z, c.w);
  const float subx = g_fract(lc.x * r.SWf), suby = g_fract(lc.y * r.SHf);
  const float dfl = sm_fetch(r.s, lc.x, lc.y);
  float d[4];
  ss_classify(r, lc, dfl, d);
  if (d[0] > 0.0f || d[1] > 0.0f) {
    const float left = ss_disc_length(r, d, lc, -1.0f, 0.0f), right = ss_disc_length(r, d, lc, 1.0f, 0.0f);
    const float down = ss_disc_length(r, d, lc, 0.0f, -1.0f), up = ss_disc_length(r, d, lc, 0.0f, 1.0f);
    const float nx = ss_normalize(r, left, right, subx), ny = ss_normalize(r, down, up, suby);
    const float fill = ss_fill(r, lc, nx, ny, d, subx, suby);
    return g_mix(fill, 1.0f, a.p.shadow_intensity);
  }
  return (lc.z <= dfl) ? 1.0f : a.p.shadow_intensity;
}

__global__ void __launch_bounds__(256) k_rbssm_prepare(const VisArgs a, RbssmItem* __restrict__ items, int* __restrict__ n_items) {
  const int x = a.rx0 + blockIdx.x * 32 + threadIdx.x, y = a.ry0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= a.rx1 || y >= a.ry1) return;
  const size_t o = (size_t)y * a.W + x;
  const float4 vertex = __ldg(&a.pos4[o]);
  if (vertex.x == 0.0f) { a.vis[o] = 0.0f; return; }
  const float4 normal = __ldg(&a.nrm4[o]);
  const float4 sc = mat4_mul(a.lmvp, vertex);
  float4 c;
  sgi_div3(sc.x, sc.y, sc.z, sc.w, c.x, c.y, c.z);
  c.w = sc.w / sc.w;
  float shadow = pre_evaluation(a, vertex, normal);
  if (sc.w > 0.0f && shadow == 1.0f) {                         // RBSSM.frag:1372
    const Smap s = {a.sm, a.SW, a.SH, a.fw, a.fh};
    const TapSrc<false> g = {s.d, s.w, 0, 0};
    const float averageDepth = pcss_blockers<0, false>(a, s, g, c);
    const float penumbraWidth = pcss_penumbra(a.p, averageDepth, c.z);
    const float stepSize = 2.0f * penumbraWidth / (float)a.p.kernel_size;
    if (stepSize <= 0.0f || stepSize >= 1.0f) shadow = 1.0f;
    else {
      RbssmItem it; it.c = c; it.pw = penumbraWidth; it.pixel = (int)o; it.pad0 = it.pad1 = 0;
      items[atomicAdd(n_items, 1)] = it;
      return;
    }
  }
  a.vis[o] = shadow;
}

__global__ void __launch_bounds__(256) k_rbssm_taps(const VisArgs a, const RbssmItem* __restrict__ items, const int* __restrict__ n_items, int* __restrict__ cursor) {
  __shared__ float vals[8][SGI_RBSSM_CHUNK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = *n_items;
  const Smap s = {a.sm, a.SW, a.SH, a.fw, a.fh};
  const float fw = ((float)a.p.kernel_size - 1.0f) * 0.5f;
  const int w0 = (int)(-fw);
  const int nk = (fw >= 0.0f) ? (int)fw - w0 + 1 : 0;          // taps per axis of `for(int w = -fw; w <= fw; w++)`
  const int ntaps = nk * nk;
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(cursor, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n) break;
    const RbssmItem it = items[i];
    Rb r = {s, a.sx, a.sy, a.p.depth_threshold, a.p.max_search, a.p.shadow_intensity, 0, a.fw, a.fh, 0.0f};
    float illuminationCount = 0.0f;
    for (int base = 0; base < ntaps; base += SGI_RBSSM_CHUNK) {
      for (int t = base + lane; t < min(base + SGI_RBSSM_CHUNK, ntaps); t += 32)
        vals[warp][t - base] = ss_tap(a, r, it.c, it.pw, fw, w0 + t % nk, w0 + t / nk);      // t = (h - w0) * nk + (w - w0)
      __syncwarp();
      if (lane == 0)
        for (int t = base; t < min(base + SGI_RBSSM_CHUNK, ntaps); t++) illuminationCount += vals[warp][t - base];
      __syncwarp();
    }
    if (lane == 0) a.vis[it.pixel] = illuminationCount / (float)(a.p.kernel_size * a.p.kernel_size);
  }
}
