/* b2s_geom.h -- scalar geometric primitives shared by the CUDA kernels and the
 * CPU oracle (see b2s_math.h for why leaf arithmetic is shared).  No loops over
 * hull vertices, contacts, faces or environments live here: GJK, EPA, the
 * manifold cache, PGS and IK are written separately on each side.
 */
#ifndef B2S_GEOM_H_
#define B2S_GEOM_H_

#include "b2s_math.h"

/* ---- closest point of a simplex to the origin (Voronoi-region tests, after
 *      Ericson, "Real-Time Collision Detection" 5.1) --------------------------
 * Input: n in 1..4 points w[0..n).  Output: barycentric weights bary[0..4)
 * (zero for unused vertices), bitmask of used vertices, closest point v.
 * Returns 1 when n == 4 and the origin is inside the tetrahedron (penetration),
 * else 0.  `degenerate` is set when a tetrahedron is flat. */
struct b2s_simplex_result {
  b2s_v3 v;
  float bary[4];
  int used;
  int inside;
  int degenerate;
};

B2S_HD void b2s_closest_segment(b2s_v3 a, b2s_v3 b, int ia, int ib, b2s_simplex_result* r) {
  b2s_v3 ab = b - a;
  float t = -dot(a, ab);
  r->bary[0] = r->bary[1] = r->bary[2] = r->bary[3] = 0.0f;
  if (t <= 0.0f) { r->v = a; r->bary[ia] = 1.0f; r->used = 1 << ia; return; }
  float den = dot(ab, ab);
  if (t >= den) { r->v = b; r->bary[ib] = 1.0f; r->used = 1 << ib; return; }
  t = t / den;
  r->v = a + ab * t;
  r->bary[ia] = 1.0f - t;
  r->bary[ib] = t;
  r->used = (1 << ia) | (1 << ib);
}

B2S_HD void b2s_closest_triangle(b2s_v3 a, b2s_v3 b, b2s_v3 c, int ia, int ib, int ic,
                                 b2s_simplex_result* r) {
  r->bary[0] = r->bary[1] = r->bary[2] = r->bary[3] = 0.0f;
  b2s_v3 ab = b - a, ac = c - a;
  float d1 = -dot(ab, a), d2 = -dot(ac, a);
  if (d1 <= 0.0f && d2 <= 0.0f) { r->v = a; r->bary[ia] = 1.0f; r->used = 1 << ia; return; }
  float d3 = -dot(ab, b), d4 = -dot(ac, b);
  if (d3 >= 0.0f && d4 <= d3) { r->v = b; r->bary[ib] = 1.0f; r->used = 1 << ib; return; }
  float vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
    float t = d1 / (d1 - d3);
    r->v = a + ab * t; r->bary[ia] = 1.0f - t; r->bary[ib] = t; r->used = (1 << ia) | (1 << ib); return;
  }
  float d5 = -dot(ab, c), d6 = -dot(ac, c);
  if (d6 >= 0.0f && d5 <= d6) { r->v = c; r->bary[ic] = 1.0f; r->used = 1 << ic; return; }
  float vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
    float t = d2 / (d2 - d6);
    r->v = a + ac * t; r->bary[ia] = 1.0f - t; r->bary[ic] = t; r->used = (1 << ia) | (1 << ic); return;
  }
  float va = d3 * d6 - d5 * d4;
  if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
    float t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    r->v = b + (c - b) * t; r->bary[ib] = 1.0f - t; r->bary[ic] = t; r->used = (1 << ib) | (1 << ic); return;
  }
  float den = 1.0f / ((va + vb) + vc);
  float tv = vb * den, tw = vc * den;
  r->v = (a + ab * tv) + ac * tw;
  r->bary[ia] = (1.0f - tv) - tw; r->bary[ib] = tv; r->bary[ic] = tw;
  r->used = (1 << ia) | (1 << ib) | (1 << ic);
}

/* is the origin on the outer side of plane (a,b,c) w.r.t. the 4th point d?
 * returns 1 outside, 0 inside, -1 degenerate (d in the plane) */
B2S_HD int b2s_origin_outside_plane(b2s_v3 a, b2s_v3 b, b2s_v3 c, b2s_v3 d) {
  b2s_v3 n = cross(b - a, c - a);
  float sp = -dot(a, n);          /* signed side of the origin */
  float sd = dot(d - a, n);       /* signed side of d */
  if (sd * sd < (1e-12f * len2(n)) * len2(d - a)) return -1;
  return (sp * sd < 0.0f) ? 1 : 0;
}

B2S_HD void b2s_closest_simplex(const b2s_v3* w, int n, b2s_simplex_result* r) {
  r->inside = 0;
  r->degenerate = 0;
  if (n == 1) {
    r->v = w[0];
    r->bary[0] = 1.0f; r->bary[1] = r->bary[2] = r->bary[3] = 0.0f;
    r->used = 1;
    return;
  }
  if (n == 2) { b2s_closest_segment(w[0], w[1], 0, 1, r); return; }
  if (n == 3) { b2s_closest_triangle(w[0], w[1], w[2], 0, 1, 2, r); return; }
  /* tetrahedron */
  int o0 = b2s_origin_outside_plane(w[0], w[1], w[2], w[3]);
  int o1 = b2s_origin_outside_plane(w[0], w[2], w[3], w[1]);
  int o2 = b2s_origin_outside_plane(w[0], w[3], w[1], w[2]);
  int o3 = b2s_origin_outside_plane(w[1], w[3], w[2], w[0]);
  if (o0 < 0 || o1 < 0 || o2 < 0 || o3 < 0) {
    /* flat tetrahedron: fall back to the newest triangle (1,2,3 hold the most recent points) */
    r->degenerate = 1;
    b2s_closest_triangle(w[1], w[2], w[3], 1, 2, 3, r);
    return;
  }
  if (!o0 && !o1 && !o2 && !o3) {
    r->inside = 1;
    r->v = v3(0.0f, 0.0f, 0.0f);
    r->bary[0] = r->bary[1] = r->bary[2] = r->bary[3] = 0.25f;
    r->used = 15;
    return;
  }
  float best = 3.0e38f;
  b2s_simplex_result t;
  t.inside = 0; t.degenerate = 0;
  if (o0) { b2s_closest_triangle(w[0], w[1], w[2], 0, 1, 2, &t); float d = len2(t.v); if (d < best) { best = d; *r = t; } }
  if (o1) { b2s_closest_triangle(w[0], w[2], w[3], 0, 2, 3, &t); float d = len2(t.v); if (d < best) { best = d; *r = t; } }
  if (o2) { b2s_closest_triangle(w[0], w[3], w[1], 0, 3, 1, &t); float d = len2(t.v); if (d < best) { best = d; *r = t; } }
  if (o3) { b2s_closest_triangle(w[1], w[3], w[2], 1, 3, 2, &t); float d = len2(t.v); if (d < best) { best = d; *r = t; } }
  r->inside = 0;
  r->degenerate = 0;
}

/* ---- 6x6 symmetric positive-definite solve (Cholesky), used by the DLS IK --
 * A is row-major 36 floats (only the lower triangle is read), b has 6 entries;
 * the solution overwrites b.  Returns 0 when a pivot is not positive. */
B2S_HD int b2s_chol6_solve(float* A, float* b) {
  for (int j = 0; j < 6; ++j) {
    float s = A[j * 6 + j];
    for (int k = 0; k < j; ++k) s = s - A[j * 6 + k] * A[j * 6 + k];
    if (!(s > 0.0f)) return 0;
    float d = sqrtf(s);
    A[j * 6 + j] = d;
    float inv = 1.0f / d;
    for (int i = j + 1; i < 6; ++i) {
      float t = A[i * 6 + j];
      for (int k = 0; k < j; ++k) t = t - A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = t * inv;
    }
  }
  for (int i = 0; i < 6; ++i) {
    float t = b[i];
    for (int k = 0; k < i; ++k) t = t - A[i * 6 + k] * b[k];
    b[i] = t / A[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    float t = b[i];
    for (int k = i + 1; k < 6; ++k) t = t - A[k * 6 + i] * b[k];
    b[i] = t / A[i * 6 + i];
  }
  return 1;
}

/* ---- persistent-manifold replacement choice ---------------------------------
 * A full manifold (4 points, positions p[0..4) in A's frame, depths) receives a
 * 5th candidate: keep the deepest point, and drop the point whose removal leaves
 * the largest quad area.  Returns the index 0..3 to overwrite. */
B2S_HD int b2s_manifold_replace_index(const b2s_v3* p, const float* depth, b2s_v3 pn, float dn) {
  int deepest = -1;
  float maxpen = dn;
  for (int i = 0; i < 4; ++i) {
    if (depth[i] < maxpen) { deepest = i; maxpen = depth[i]; }
  }
  float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
  if (deepest != 0) r0 = len2(cross(pn - p[1], p[3] - p[2]));
  if (deepest != 1) r1 = len2(cross(pn - p[0], p[3] - p[2]));
  if (deepest != 2) r2 = len2(cross(pn - p[0], p[3] - p[1]));
  if (deepest != 3) r3 = len2(cross(pn - p[0], p[2] - p[1]));
  int best = 0;
  float bv = r0;
  if (r1 > bv) { bv = r1; best = 1; }
  if (r2 > bv) { bv = r2; best = 2; }
  if (r3 > bv) { bv = r3; best = 3; }
  return best;
}

#endif  /* B2S_GEOM_H_ */
