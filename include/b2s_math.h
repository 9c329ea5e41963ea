/* b2s_math.h -- leaf fp32 math shared by the CUDA kernels and the CPU oracle.
 *
 * Why shared: the parity bar for this path is BIT-EXACT integer outputs
 * (contact-pair indices, collision masks, phase states) and fp32 poses that
 * agree to round-off.  That only holds if both sides perform the same IEEE
 * operations in the same order, so the leaf formulas (vector / quaternion /
 * 3x3 algebra, polynomial sin/cos/atan2) live here once.  Everything above the
 * leaves -- GJK, EPA, manifolds, PGS, IK, the controller and phase machines,
 * the raster -- is written twice: warp-cooperative in robovat_b200/csrc (.cu)
 * and scalar/sequential in oracle (.cpp).
 *
 * Both sides MUST be compiled without FMA contraction
 * (nvcc -fmad=false, g++ -ffp-contract=off) and without fast-math.
 *
 * Conventions follow the reference (robovat/math + third_party/transformations):
 *   quaternion order [x, y, z, w]            (transformations.py:1194-1248)
 *   Euler angles 'sxyz' = static roll/pitch/yaw (transformations.py:1034-1081)
 *   Pose.inverse: p' = -p.R, R' = R^T          (robovat/math/pose.py:161-172)
 */
#ifndef B2S_MATH_H_
#define B2S_MATH_H_

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define B2S_HD __host__ __device__ __forceinline__
#else
#define B2S_HD static inline
#endif

#define B2S_PI 3.14159265358979323846f
#define B2S_HALF_PI 1.57079632679489661923f

struct b2s_v3 { float x, y, z; };
struct b2s_q4 { float x, y, z, w; };              /* [x,y,z,w] */
struct b2s_m3 { b2s_v3 r0, r1, r2; };             /* rows */

B2S_HD b2s_v3 v3(float x, float y, float z) { b2s_v3 r; r.x = x; r.y = y; r.z = z; return r; }
B2S_HD b2s_v3 operator+(b2s_v3 a, b2s_v3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
B2S_HD b2s_v3 operator-(b2s_v3 a, b2s_v3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
B2S_HD b2s_v3 operator-(b2s_v3 a) { return v3(-a.x, -a.y, -a.z); }
B2S_HD b2s_v3 operator*(b2s_v3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
B2S_HD b2s_v3 operator*(float s, b2s_v3 a) { return v3(a.x * s, a.y * s, a.z * s); }
/* a + b * s */
B2S_HD b2s_v3 vmad(b2s_v3 a, b2s_v3 b, float s) { return v3(fmaf(b.x, s, a.x), fmaf(b.y, s, a.y), fmaf(b.z, s, a.z)); }
/* explicit fused multiply-adds: one IEEE rounding each, identical on host (-mfma / libm fmaf) and device */
B2S_HD float dot(b2s_v3 a, b2s_v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
B2S_HD b2s_v3 cross(b2s_v3 a, b2s_v3 b) {
  return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
B2S_HD float len2(b2s_v3 a) { return dot(a, a); }
B2S_HD float len(b2s_v3 a) { return sqrtf(dot(a, a)); }
B2S_HD b2s_v3 vmul(b2s_v3 a, b2s_v3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
B2S_HD float vget(b2s_v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* ---- polynomial elementary functions (identical on host and device) ------ */

/* sin and cos of x (|x| < ~1e4), Cody-Waite reduction to [-pi/4, pi/4] and
 * the classic single-precision minimax polynomials. */
B2S_HD void b2s_sincos(float x, float* s, float* c) {
#ifdef B2S_F64
  *s = sin(x); *c = cos(x);       /* double-precision oracle build (oracle/b2o_f64.h): libm, not the fp32 polynomials */
  return;
#endif
  float k = rintf(x * 0.63661977236758134308f);       /* x * 2/pi */
  float r = x - k * 1.5703125f;
  r = r - k * 4.837512969970703125e-4f;
  r = r - k * 7.54978995489188e-8f;
  float z = r * r;
  float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
  float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z
             - 0.5f * z + 1.0f;
  int q = ((int)k) & 3;
  float ss = (q & 1) ? pc : ps;
  float cc = (q & 1) ? ps : pc;
  if (q == 1 || q == 2) cc = -cc;
  if (q == 2 || q == 3) ss = -ss;
  *s = ss;
  *c = cc;
}

/* atan for x >= 0 */
B2S_HD float b2s_atan_pos(float x) {
  float y0;
  if (x > 2.414213562373095f) { y0 = B2S_HALF_PI; x = -(1.0f / x); }
  else if (x > 0.4142135623730950f) { y0 = 0.78539816339744830962f; x = (x - 1.0f) / (x + 1.0f); }
  else { y0 = 0.0f; }
  float z = x * x;
  float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
             - 3.33329491539e-1f) * z * x + x;
  return y0 + p;
}

B2S_HD float b2s_atan2(float y, float x) {
#ifdef B2S_F64
  if (x != 0.0f) return atan2(y, x);
#endif
  if (x == 0.0f) {
    if (y > 0.0f) return B2S_HALF_PI;
    if (y < 0.0f) return -B2S_HALF_PI;
    return 0.0f;
  }
  float a = b2s_atan_pos(fabsf(y / x));
  if (x < 0.0f) a = B2S_PI - a;
  return (y < 0.0f) ? -a : a;
}

/* wrap an angle to [-pi, pi): (a + pi) mod 2pi - pi with python's sign rule
 * (robovat/envs/push/push_env.py:913). */
B2S_HD float b2s_wrap_pi(float a) {
  float t = a + B2S_PI;
  float m = t - floorf(t / (2.0f * B2S_PI)) * (2.0f * B2S_PI);
  return m - B2S_PI;
}

/* ---- quaternions [x,y,z,w] ---------------------------------------------- */

B2S_HD b2s_q4 q4(float x, float y, float z, float w) { b2s_q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

/* Hamilton product a*b (same component formulas as
 * transformations.quaternion_multiply, third_party/transformations.py:1362). */
B2S_HD b2s_q4 qmul(b2s_q4 a, b2s_q4 b) {
  return q4(fmaf(-a.z, b.y, fmaf(a.y, b.z, fmaf(a.x, b.w, a.w * b.x))),
            fmaf(a.z, b.x, fmaf(a.y, b.w, fmaf(-a.x, b.z, a.w * b.y))),
            fmaf(a.z, b.w, fmaf(-a.y, b.x, fmaf(a.x, b.y, a.w * b.z))),
            fmaf(-a.z, b.z, fmaf(-a.y, b.y, fmaf(-a.x, b.x, a.w * b.w))));
}
B2S_HD b2s_q4 qconj(b2s_q4 a) { return q4(-a.x, -a.y, -a.z, a.w); }
B2S_HD b2s_q4 qnormalize(b2s_q4 a) {
  float n = sqrtf(((a.x * a.x + a.y * a.y) + a.z * a.z) + a.w * a.w);
  float inv = 1.0f / n;
  return q4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}

/* rotation matrix of a unit quaternion (transformations.py:1290-1303) */
B2S_HD b2s_m3 q_to_m3(b2s_q4 q) {
  float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
  float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
  b2s_m3 m;
  m.r0 = v3(1.0f - 2.0f * (yy + zz), 2.0f * (xy - wz), 2.0f * (xz + wy));
  m.r1 = v3(2.0f * (xy + wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz - wx));
  m.r2 = v3(2.0f * (xz - wy), 2.0f * (yz + wx), 1.0f - 2.0f * (xx + yy));
  return m;
}
B2S_HD b2s_v3 mmul(b2s_m3 m, b2s_v3 v) { return v3(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); }
/* M^T v */
B2S_HD b2s_v3 mtmul(b2s_m3 m, b2s_v3 v) {
  return v3(fmaf(m.r2.x, v.z, fmaf(m.r1.x, v.y, m.r0.x * v.x)),
            fmaf(m.r2.y, v.z, fmaf(m.r1.y, v.y, m.r0.y * v.x)),
            fmaf(m.r2.z, v.z, fmaf(m.r1.z, v.y, m.r0.z * v.x)));
}
B2S_HD b2s_m3 mmulm(b2s_m3 a, b2s_m3 b) {
  b2s_v3 c0 = v3(b.r0.x, b.r1.x, b.r2.x), c1 = v3(b.r0.y, b.r1.y, b.r2.y), c2 = v3(b.r0.z, b.r1.z, b.r2.z);
  b2s_m3 m;
  m.r0 = v3(dot(a.r0, c0), dot(a.r0, c1), dot(a.r0, c2));
  m.r1 = v3(dot(a.r1, c0), dot(a.r1, c1), dot(a.r1, c2));
  m.r2 = v3(dot(a.r2, c0), dot(a.r2, c1), dot(a.r2, c2));
  return m;
}
B2S_HD b2s_m3 mtranspose(b2s_m3 a) {
  b2s_m3 m;
  m.r0 = v3(a.r0.x, a.r1.x, a.r2.x);
  m.r1 = v3(a.r0.y, a.r1.y, a.r2.y);
  m.r2 = v3(a.r0.z, a.r1.z, a.r2.z);
  return m;
}
B2S_HD b2s_v3 qrot(b2s_q4 q, b2s_v3 v) { return mmul(q_to_m3(q), v); }

/* quaternion about a unit axis (transformations.quaternion_about_axis :1251) */
B2S_HD b2s_q4 q_axis_angle(b2s_v3 axis, float angle) {
  float s, c;
  b2s_sincos(0.5f * angle, &s, &c);
  return q4(axis.x * s, axis.y * s, axis.z * s, c);
}

/* quaternion from Euler 'sxyz' (roll, pitch, yaw); closed form of
 * transformations.quaternion_from_euler(ai, aj, ak, 'sxyz') :1194-1248 */
B2S_HD b2s_q4 q_from_euler(float roll, float pitch, float yaw) {
  float si, ci, sj, cj, sk, ck;
  b2s_sincos(0.5f * roll, &si, &ci);
  b2s_sincos(0.5f * pitch, &sj, &cj);
  b2s_sincos(0.5f * yaw, &sk, &ck);
  float cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
  return q4(cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc, cj * cc + sj * ss);
}

/* Euler 'sxyz' from a rotation matrix (transformations.euler_from_matrix3
 * :1142-1180 with i,j,k = 0,1,2, parity 0, repetition 0, frame 0). */
B2S_HD b2s_v3 euler_from_m3(b2s_m3 m) {
  float cy = sqrtf(m.r0.x * m.r0.x + m.r1.x * m.r1.x);
  float ax, ay, az;
  if (cy > 8.8817841970012523e-16f) {                 /* _EPS = 4 * DBL_EPSILON */
    ax = b2s_atan2(m.r2.y, m.r2.z);
    ay = b2s_atan2(-m.r2.x, cy);
    az = b2s_atan2(m.r1.x, m.r0.x);
  } else {
    ax = b2s_atan2(-m.r1.z, m.r1.y);
    ay = b2s_atan2(-m.r2.x, cy);
    az = 0.0f;
  }
  return v3(ax, ay, az);
}
B2S_HD b2s_v3 euler_from_q(b2s_q4 q) { return euler_from_m3(q_to_m3(q)); }
B2S_HD float yaw_from_q(b2s_q4 q) { return euler_from_q(q).z; }

/* rotation vector (axis * angle, angle in [0, pi]) of a unit quaternion */
B2S_HD b2s_v3 q_to_rotvec(b2s_q4 q) {
  if (q.w < 0.0f) q = q4(-q.x, -q.y, -q.z, -q.w);
  float n = sqrtf((q.x * q.x + q.y * q.y) + q.z * q.z);
  if (n < 1e-12f) return v3(0.0f, 0.0f, 0.0f);
  float ang = 2.0f * b2s_atan2(n, q.w);
  float k = ang / n;
  return v3(q.x * k, q.y * k, q.z * k);
}

/* first-order quaternion integration q <- normalize(q + dt/2 * [w,0] (x) q) */
B2S_HD b2s_q4 q_integrate(b2s_q4 q, b2s_v3 w, float dt) {
  b2s_q4 wq = q4(w.x, w.y, w.z, 0.0f);
  b2s_q4 d = qmul(wq, q);
  float h = 0.5f * dt;
  return qnormalize(q4(q.x + h * d.x, q.y + h * d.y, q.z + h * d.z, q.w + h * d.w));
}

/* two tangents spanning the plane orthogonal to unit n */
B2S_HD void plane_space(b2s_v3 n, b2s_v3* p, b2s_v3* q) {
  if (fabsf(n.z) > 0.70710678118654752440f) {
    float a = n.y * n.y + n.z * n.z;
    float k = 1.0f / sqrtf(a);
    *p = v3(0.0f, -n.z * k, n.y * k);
    *q = v3(a * k, -n.x * p->z, n.x * p->y);
  } else {
    float a = n.x * n.x + n.y * n.y;
    float k = 1.0f / sqrtf(a);
    *p = v3(-n.y * k, n.x * k, 0.0f);
    *q = v3(-n.z * p->y, n.z * p->x, a * k);
  }
}

/* order-preserving map float -> uint32 (for exact integer max/min reductions) */
B2S_HD uint32_t f2ord(float f) {
  union { float f; uint32_t u; } c;
  c.f = f;
  return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}
B2S_HD float ord2f(uint32_t u) {
  union { float f; uint32_t u; } c;
  c.u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return c.f;
}

/* ---- counter-based RNG (Philox-4x32-10), keyed by (seed, global env id) -- */
struct b2s_u4 { uint32_t x, y, z, w; };
B2S_HD void b2s_mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}
B2S_HD b2s_u4 b2s_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0, lo0, hi1, lo1;
    b2s_mulhilo(0xD2511F53u, c0, &hi0, &lo0);
    b2s_mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  b2s_u4 r; r.x = c0; r.y = c1; r.z = c2; r.w = c3; return r;
}
/* uniform in [0,1) with 24 bits */
B2S_HD float b2s_u01(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

#endif  /* B2S_MATH_H_ */
