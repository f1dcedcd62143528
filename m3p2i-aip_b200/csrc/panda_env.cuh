// panda_env.cuh — Franka Panda (7 revolute + 2 prismatic finger DoF) with two free cubes: forward kinematics of the
// URDF chain, velocity-tracked arm, force-limited fingers, box/SDF contact with Coulomb friction, per-step costs.
// Replaces REACTIVE_TAMP.dynamics -> IsaacGymWrapper.step (reactive_tamp.py:63-70, isaacgym_wrapper.py:354-360) and
// Objective.compute_cost for reach / pick / place (cost_functions.py:91-169; skill_utils.py:140-180,224-289) for the
// panda_env scene. Kinematic constants: assets/urdf/franka_description/robots/franka_panda.urdf:27-242.
#pragma once
#include "common.cuh"

namespace m3 {

struct Cube {
  V3 p;
  float qx, qy, qz, qw;
  V3 v, w;
};

struct PandaEnv {
  float q[9], qd[9];
  Cube cube[2];             // cubeA, cubeB
  V3 f_table, f_shelf, f_cubeb;
  // not part of the stored state: which cubes slept through the last sub-step this thread integrated (bit i) and on
  // which fixed box (panda_step's dormant shortcut)
  unsigned slept;
  int support[2];

  DEV void load(const float* p, int stride, int k) {
    const float* s = p + k;
    int f = 0;
    slept = 0u; support[0] = 0; support[1] = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) { q[j] = s[(f++) * stride]; qd[j] = s[(f++) * stride]; }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      Cube& c = cube[i];
      c.p = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
      c.qx = s[f * stride]; c.qy = s[(f + 1) * stride]; c.qz = s[(f + 2) * stride]; c.qw = s[(f + 3) * stride]; f += 4;
      c.v = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
      c.w = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
    }
    f_table = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
    f_shelf = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
    f_cubeb = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]);
  }
  DEV void store(float* p, int stride, int k) const {
    float* s = p + k;
    int f = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) { s[(f++) * stride] = q[j]; s[(f++) * stride] = qd[j]; }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const Cube& c = cube[i];
      s[f * stride] = c.p.x; s[(f + 1) * stride] = c.p.y; s[(f + 2) * stride] = c.p.z; f += 3;
      s[f * stride] = c.qx; s[(f + 1) * stride] = c.qy; s[(f + 2) * stride] = c.qz; s[(f + 3) * stride] = c.qw; f += 4;
      s[f * stride] = c.v.x; s[(f + 1) * stride] = c.v.y; s[(f + 2) * stride] = c.v.z; f += 3;
      s[f * stride] = c.w.x; s[(f + 1) * stride] = c.w.y; s[(f + 2) * stride] = c.w.z; f += 3;
    }
    s[f * stride] = f_table.x; s[(f + 1) * stride] = f_table.y; s[(f + 2) * stride] = f_table.z; f += 3;
    s[f * stride] = f_shelf.x; s[(f + 1) * stride] = f_shelf.y; s[(f + 2) * stride] = f_shelf.z; f += 3;
    s[f * stride] = f_cubeb.x; s[(f + 1) * stride] = f_cubeb.y; s[(f + 2) * stride] = f_cubeb.z;
  }
  DEV float4 state_row() const { return make_float4(q[0], qd[0], q[1], qd[1]); }
};

// ------------------------------------------------------------------ forward kinematics
struct Hand {
  V3 p;
  M33 R;
  V3 v, w;  // twist of the hand-frame origin due to the 7 arm joints
};

constexpr float kHandZ = 0.107f;             // panda_hand_joint origin (urdf:183)
constexpr float kHandYawC = 0.70710678118f;  // cos(-pi/4)
constexpr float kHandYawS = -0.70710678118f; // sin(-pi/4)
constexpr float kFingerZ = 0.0584f;          // panda_finger_joint1/2 origin (urdf:229,237)
constexpr float kGripReachT = 0.12f;         // hand box centre to the farthest point of a finger box, minus the hand box radius

template <int J>
DEV void panda_joint(V3& p, M33& R, V3& v, V3& w, float qj, float qdj, bool want_twist) {
  // joint origin xyz and roll (quarter turns about x) of panda_joint{J+1}
  constexpr float X[7] = {0.0f, 0.0f, 0.0f, 0.0825f, -0.0825f, 0.0f, 0.088f};
  constexpr float Y[7] = {0.0f, 0.0f, -0.316f, 0.0f, 0.384f, 0.0f, 0.0f};
  constexpr float Z[7] = {0.333f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  constexpr int ROLL[7] = {0, -1, 1, 1, -1, 1, 1};
  const V3 d = mul(R, mk(X[J], Y[J], Z[J]));
  if (want_twist) v = v + cross(w, d);  // the frame origin moves with the links before it
  p = p + d;
  if (ROLL[J] == 1) { const V3 c1 = R.cy; R.cy = R.cz; R.cz = -c1; }
  if (ROLL[J] == -1) { const V3 c1 = R.cy; R.cy = -R.cz; R.cz = c1; }
  if (want_twist) w = w + qdj * R.cz;   // joint axis = local +z
  float sn, cs;
  sincosf(qj, &sn, &cs);
  const V3 c0 = R.cx, c1 = R.cy;
  R.cx = cs * c0 + sn * c1;
  R.cy = cs * c1 - sn * c0;
}

// hand pose (+ twist) from the 7 arm joints; the twist is accumulated link by link (v_i = v_{i-1} + w x d)
DEV void panda_hand(const PandaParams& P, const float* q, const float* qd, bool want_twist, Hand& H) {
  V3 p = mk(P.base[0], P.base[1], P.base[2]);
  M33 R = {mk(1, 0, 0), mk(0, 1, 0), mk(0, 0, 1)};
  V3 v = mk(0, 0, 0), w = mk(0, 0, 0);
  panda_joint<0>(p, R, v, w, q[0], qd[0], want_twist);
  panda_joint<1>(p, R, v, w, q[1], qd[1], want_twist);
  panda_joint<2>(p, R, v, w, q[2], qd[2], want_twist);
  panda_joint<3>(p, R, v, w, q[3], qd[3], want_twist);
  panda_joint<4>(p, R, v, w, q[4], qd[4], want_twist);
  panda_joint<5>(p, R, v, w, q[5], qd[5], want_twist);
  panda_joint<6>(p, R, v, w, q[6], qd[6], want_twist);
  const V3 d = kHandZ * R.cz;
  if (want_twist) v = v + cross(w, d);
  p = p + d;
  const V3 c0 = R.cx, c1 = R.cy;
  R.cx = kHandYawC * c0 + kHandYawS * c1;
  R.cy = kHandYawC * c1 - kHandYawS * c0;
  H.p = p; H.R = R; H.v = v; H.w = w;
}

// quaternion (x,y,z,w) of a rotation matrix (Shepperd)
DEV void R_to_quat(const M33& R, float* q) {
  const float r00 = R.cx.x, r11 = R.cy.y, r22 = R.cz.z;
  const float tr = r00 + r11 + r22;
  if (tr > 0.0f) {
    const float s = sqrtf(tr + 1.0f) * 2.0f;
    q[3] = 0.25f * s; q[0] = (R.cy.z - R.cz.y) / s; q[1] = (R.cz.x - R.cx.z) / s; q[2] = (R.cx.y - R.cy.x) / s;
  } else if (r00 > r11 && r00 > r22) {
    const float s = sqrtf(1.0f + r00 - r11 - r22) * 2.0f;
    q[3] = (R.cy.z - R.cz.y) / s; q[0] = 0.25f * s; q[1] = (R.cy.x + R.cx.y) / s; q[2] = (R.cz.x + R.cx.z) / s;
  } else if (r11 > r22) {
    const float s = sqrtf(1.0f + r11 - r00 - r22) * 2.0f;
    q[3] = (R.cz.x - R.cx.z) / s; q[0] = (R.cy.x + R.cx.y) / s; q[1] = 0.25f * s; q[2] = (R.cz.y + R.cy.z) / s;
  } else {
    const float s = sqrtf(1.0f + r22 - r00 - r11) * 2.0f;
    q[3] = (R.cx.y - R.cy.x) / s; q[0] = (R.cz.x + R.cx.z) / s; q[1] = (R.cz.y + R.cy.z) / s; q[2] = 0.25f * s;
  }
}

// rigid-body rows [3][13] (leftfinger, rightfinger, hand) as IsaacGym reports them
DEV void panda_links(const PandaParams& P, const float* q, const float* qd, float* out) {
  Hand H;
  panda_hand(P, q, qd, true, H);
  float quat[4];
  R_to_quat(H.R, quat);
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const float sg = f == 0 ? 1.0f : -1.0f;
    const V3 r = mul(H.R, mk(0.0f, sg * q[7 + f], kFingerZ));
    const V3 pos = H.p + r, vel = H.v + cross(H.w, r) + (sg * qd[7 + f]) * H.R.cy;
    float* o = out + 13 * f;
    o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = quat[0]; o[4] = quat[1]; o[5] = quat[2]; o[6] = quat[3];
    o[7] = vel.x; o[8] = vel.y; o[9] = vel.z; o[10] = H.w.x; o[11] = H.w.y; o[12] = H.w.z;
  }
  float* o = out + 26;
  o[0] = H.p.x; o[1] = H.p.y; o[2] = H.p.z; o[3] = quat[0]; o[4] = quat[1]; o[5] = quat[2]; o[6] = quat[3];
  o[7] = H.v.x; o[8] = H.v.y; o[9] = H.v.z; o[10] = H.w.x; o[11] = H.w.y; o[12] = H.w.z;
}

// ------------------------------------------------------------------ contact solver
// A body as the solver sees it. Free cube: im, ii > 0. Kinematic link: im = ii = 0 with a prescribed twist (v, w)
// about x; a finger additionally has one sliding DoF (speed `slide` along `axis`, inverse mass `ims`).
// Fixed box: everything zero.
struct Dyn3 {
  V3 v, w, x;
  float im, ii;
  V3 axis;
  float slide, ims;
};

DEV V3 point_vel(const Dyn3& B, V3 r) { return B.v + cross(B.w, r) + B.slide * B.axis; }
DEV float eff_mass(const Dyn3& B, V3 r, V3 d) {
  const V3 c = cross(r, d);
  const float a = dot(B.axis, d);
  return B.im + B.ii * dot(c, c) + B.ims * a * a;
}
DEV void apply_impulse(Dyn3& B, V3 r, V3 Pv, float sgn) {
  B.v = B.v + (sgn * B.im) * Pv;
  B.w = B.w + (sgn * B.ii) * cross(r, Pv);
  B.slide += sgn * B.ims * dot(B.axis, Pv);
}

// One contact: unit normal n from B to A, depth > 0 = penetration, at world point c. Returns the impulse on A.
DEV V3 solve_contact3(Dyn3& A, Dyn3& B, V3 n, float depth, V3 c, float mu, float h, const PandaParams& P) {
  const V3 ra = c - A.x, rb = c - B.x;
  V3 rv = point_vel(A, ra) - point_vel(B, rb);
  float vn = dot(rv, n);
  const float kn = eff_mass(A, ra, n) + eff_mass(B, rb, n);
  if (kn <= 0.0f) return mk(0, 0, 0);
  float target;
  if (depth > 0.0f) {
    const float pen = fmaxf(depth - P.slop, 0.0f);
    target = fminf(P.baumgarte * pen / h, P.max_corr_vel);
  } else {
    target = depth / h;
  }
  const float jn = (target - vn) / kn;
  if (jn <= 0.0f) return mk(0, 0, 0);
  V3 Pn = jn * n;
  apply_impulse(A, ra, Pn, 1.0f); apply_impulse(B, rb, Pn, -1.0f);
  rv = point_vel(A, ra) - point_vel(B, rb);
  vn = dot(rv, n);
  V3 t = rv - vn * n;
  const float vt = sqrtf(dot(t, t));
  if (vt < 1e-9f) return Pn;
  t = (1.0f / vt) * t;
  const float kt = eff_mass(A, ra, t) + eff_mass(B, rb, t);
  if (kt <= 0.0f) return Pn;
  const float jt = fminf(vt / kt, mu * jn);
  const V3 Pt = (-jt) * t;
  apply_impulse(A, ra, Pt, 1.0f); apply_impulse(B, rb, Pt, -1.0f);
  return Pn + Pt;
}

// Accumulated-impulse form of solve_contact3. `L` = (normal impulse, tangential impulse vector) the slot has applied
// so far in this sub-step. The normal impulse is clamped on its running total (a later visit may take back what an
// earlier one over-applied), the friction impulse is clamped to the Coulomb cone of the TOTAL normal impulse.
// `first`: the slot's first visit in this sub-step; then `warm` tells whether the slot was in contact in the previous
// sub-step of the same step(): if so P.warm_start times what it held (projected onto the tangent plane of the new
// normal, inside the new cone) is re-applied before the solve, else the slot starts from zero.
// Returns the impulse applied to A by this call.
DEV V3 solve_contact3_acc(Dyn3& A, Dyn3& B, V3 n, float depth, V3 c, float mu, float h, const PandaParams& P, float4& L,
                          bool first, bool warm) {
  const V3 ra = c - A.x, rb = c - B.x;
  V3 done = mk(0, 0, 0);
  if (first) {
    if (warm) {
      const V3 lt0 = mk(L.y, L.z, L.w);
      const float tn = dot(lt0, n);
      const float ln = P.warm_start * L.x;
      V3 lt = P.warm_start * (lt0 - tn * n);
      const float lim = mu * ln, m2 = dot(lt, lt);
      if (m2 > lim * lim) lt = (lim / sqrtf(m2)) * lt;
      L = make_float4(ln, lt.x, lt.y, lt.z);
      const V3 Pw = ln * n + lt;
      apply_impulse(A, ra, Pw, 1.0f); apply_impulse(B, rb, Pw, -1.0f);
      done = Pw;
    } else {
      L = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  }
  V3 rv = point_vel(A, ra) - point_vel(B, rb);
  float vn = dot(rv, n);
  const float kn = eff_mass(A, ra, n) + eff_mass(B, rb, n);
  if (kn <= 0.0f) return done;
  float target;
  if (depth > 0.0f) {
    const float pen = fmaxf(depth - P.slop, 0.0f);
    target = fminf(P.baumgarte * pen / h, P.max_corr_vel);
  } else {
    target = depth / h;
  }
  const float ln = fmaxf(L.x + (target - vn) / kn, 0.0f);
  const V3 Pn = (ln - L.x) * n;
  L.x = ln;
  apply_impulse(A, ra, Pn, 1.0f); apply_impulse(B, rb, Pn, -1.0f);
  rv = point_vel(A, ra) - point_vel(B, rb);
  vn = dot(rv, n);
  V3 t = rv - vn * n;
  const float vt2 = dot(t, t);
  V3 lt = mk(L.y, L.z, L.w);
  if (vt2 >= 1e-18f) {
    const float vt = sqrtf(vt2);
    t = (1.0f / vt) * t;
    const float kt = eff_mass(A, ra, t) + eff_mass(B, rb, t);
    if (kt > 0.0f) lt = lt - (vt / kt) * t;
  }
  const float lim = mu * ln, m2 = dot(lt, lt);
  if (m2 > lim * lim) lt = (lim / sqrtf(m2)) * lt;
  const V3 Pt = lt - mk(L.y, L.z, L.w);
  L.y = lt.x; L.z = lt.y; L.w = lt.z;
  apply_impulse(A, ra, Pt, 1.0f); apply_impulse(B, rb, Pt, -1.0f);
  return done + Pn + Pt;
}

// The same contact equations for the hot pair type, a free cube against a fixed box: every term of the fixed body
// vanishes, r x P is reused from r x n / r x t. Returns the impulse on the cube.
DEV V3 solve_cube_static(V3& v, V3& w, float im, float ii, V3 x, V3 n, float depth, V3 c, float mu, float h,
                         const PandaParams& P) {
  // divisions and the tangent normalisation use the SFU approximations (<= 2 ulp): this function is a third of all
  // instructions of a rollout and the oracle comparison is tolerance based (see tests/test_gpu_parity.py)
  const float inv_h = __frcp_rn(h);
  const V3 ra = c - x;
  V3 va = v + cross(w, ra);
  float vn = dot(va, n);
  const V3 rn = cross(ra, n);
  const float kn = im + ii * dot(rn, rn);
  float target;
  if (depth > 0.0f) {
    const float pen = fmaxf(depth - P.slop, 0.0f);
    target = fminf(P.baumgarte * pen * inv_h, P.max_corr_vel);
  } else {
    target = depth * inv_h;
  }
  const float jn = __fdividef(target - vn, kn);
  if (jn <= 0.0f) return mk(0, 0, 0);
  v = v + (jn * im) * n;
  w = w + (jn * ii) * rn;
  va = v + cross(w, ra);
  vn = dot(va, n);
  V3 t = va - vn * n;
  const float vt2 = dot(t, t);
  if (vt2 < 1e-18f) return jn * n;
  const float ivt = rsqrtf(vt2), vt = vt2 * ivt;
  t = ivt * t;
  const V3 rt = cross(ra, t);
  const float kt = im + ii * dot(rt, rt);
  const float jt = fminf(__fdividef(vt, kt), mu * jn);
  v = v - (jt * im) * t;
  w = w - (jt * ii) * rt;
  return jn * n - jt * t;
}

// Kinematic link (prescribed twist lv, lw about lx; a finger adds one sliding DoF of inverse mass ims along `axis`
// with speed `slide`) against a free cube. Unit normal n points from the cube to the link. Same equations as
// solve_contact3 with the vanishing terms removed. Returns the impulse on the link.
DEV V3 solve_link_cube(V3 lv, V3 lw, V3 lx, V3 axis, float& slide, float ims, V3& v, V3& w, float im, float ii, V3 x,
                       V3 n, float depth, V3 c, float mu, float h, const PandaParams& P) {
  const V3 rl = c - lx, rc = c - x;
  const V3 vl0 = lv + cross(lw, rl);   // part of the link's point velocity that contacts cannot change
  V3 rv = (vl0 + slide * axis) - (v + cross(w, rc));
  float vn = dot(rv, n);
  const V3 rcn = cross(rc, n);
  const float an = dot(axis, n);
  const float kn = ims * an * an + im + ii * dot(rcn, rcn);
  if (kn <= 0.0f) return mk(0, 0, 0);
  float target;
  if (depth > 0.0f) {
    const float pen = fmaxf(depth - P.slop, 0.0f);
    target = fminf(P.baumgarte * pen / h, P.max_corr_vel);
  } else {
    target = depth / h;
  }
  const float jn = (target - vn) / kn;
  if (jn <= 0.0f) return mk(0, 0, 0);
  slide += ims * an * jn;
  v = v - (jn * im) * n;
  w = w - (jn * ii) * rcn;
  rv = (vl0 + slide * axis) - (v + cross(w, rc));
  vn = dot(rv, n);
  V3 t = rv - vn * n;
  const float vt = sqrtf(dot(t, t));
  if (vt < 1e-9f) return jn * n;
  t = (1.0f / vt) * t;
  const V3 rct = cross(rc, t);
  const float at = dot(axis, t);
  const float kt = ims * at * at + im + ii * dot(rct, rct);
  if (kt <= 0.0f) return jn * n;
  const float jt = fminf(vt / kt, mu * jn);
  slide -= ims * at * jt;
  v = v + (jt * im) * t;
  w = w + (jt * ii) * rct;
  return jn * n - jt * t;
}

struct OBox3 {
  V3 c;
  M33 R;
  V3 half;
};

DEV OBox3 obox_of(const Static3& s) {
  OBox3 b;
  b.c = mk(s.c[0], s.c[1], s.c[2]);
  b.R.cx = mk(s.R[0], s.R[3], s.R[6]); b.R.cy = mk(s.R[1], s.R[4], s.R[7]); b.R.cz = mk(s.R[2], s.R[5], s.R[8]);
  b.half = mk(s.half[0], s.half[1], s.half[2]);
  return b;
}
DEV OBox3 obox_of(const Cube& c, const float* half) {
  OBox3 b;
  b.c = c.p; b.R = quat_to_R(c.qx, c.qy, c.qz, c.qw); b.half = mk(half[0], half[1], half[2]);
  return b;
}

// signed-distance test of world point p against box b; normal points out of b along the least-penetration axis
DEV bool point_in_box(V3 p, const OBox3& b, float margin, V3& n, float& depth) {
  const V3 d = mulT(b.R, p - b.c);
  const float qx = fabsf(d.x) - b.half.x, qy = fabsf(d.y) - b.half.y, qz = fabsf(d.z) - b.half.z;
  int ax = 0;
  float qm = qx;
  if (qy > qm) { ax = 1; qm = qy; }
  if (qz > qm) { ax = 2; qm = qz; }
  if (qm >= margin) return false;
  const float sg = comp(d, ax) < 0.0f ? -1.0f : 1.0f;
  n = sg * col(b.R, ax);
  depth = -qm;
  return true;
}

DEV V3 box_corner(const OBox3& b, int i) {
  return b.c + mul(b.R, mk((i & 1) ? b.half.x : -b.half.x, (i & 2) ? b.half.y : -b.half.y, (i & 4) ? b.half.z : -b.half.z));
}

// bounding sphere of a against the box b
DEV bool boxes_near(const OBox3& a, const OBox3& b, float margin) {
  const V3 d = mulT(b.R, a.c - b.c);
  const float ra = sqrtf(dot(a.half, a.half)) + margin;
  const float ex = fmaxf(fabsf(d.x) - b.half.x, 0.0f), ey = fmaxf(fabsf(d.y) - b.half.y, 0.0f),
              ez = fmaxf(fabsf(d.z) - b.half.z, 0.0f);
  return ex * ex + ey * ey + ez * ez <= ra * ra;
}

// Corners of box `ba` (body A) against box `bb` (body B), each hit solved as the pair (A, B) with the normal out of bb
// -- or, FLIP, as the pair (B, A) with the normal reversed (the corners of the second body of a pair in the first).
// `lam` = the 8 accumulator slots of this corner set (nullptr: plain solves), `prev` / `cur` = bit per corner: in
// contact in the previous / this sub-step; `room` = contacts this corner set may still take (counted down).
// Returns the impulse received by the FIRST body of the solved pair.
template <bool FLIP>
DEV V3 corners_vs_box3(Dyn3& A, const OBox3& ba, Dyn3& B, const OBox3& bb, float mu, float h, const PandaParams& P,
                       float4* lam, unsigned prev, unsigned& cur, bool first, int& room) {
  V3 got = mk(0, 0, 0);
  for (int i = 0; i < 8; ++i) {
    const V3 p = box_corner(ba, i);
    V3 n; float depth;
    if (!point_in_box(p, bb, P.contact_margin, n, depth)) continue;
    if (room-- <= 0) continue;   // the list of this cube is full (kLinkCap): the contact is ignored
    cur |= 1u << i;
    if (lam) {
      const bool warm = (prev >> i) & 1u;
      got = got + (FLIP ? solve_contact3_acc(B, A, -n, depth, p, mu, h, P, lam[i], first, warm)
                        : solve_contact3_acc(A, B, n, depth, p, mu, h, P, lam[i], first, warm));
    } else {
      float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // no history: the plain one-shot solve
      got = got + (FLIP ? solve_contact3_acc(B, A, -n, depth, p, mu, h, P, zero, false, false)
                        : solve_contact3_acc(A, B, n, depth, p, mu, h, P, zero, false, false));
    }
  }
  return got;
}

// kinematic link box against a fixed box: penalty force on the fixed body (reported to the collision cost only)
DEV void link_vs_static(const OBox3& lb, const Dyn3& L, const OBox3& sb, float mu, const PandaParams& P, V3& f_acc) {
  if (!boxes_near(lb, sb, 0.0f)) return;
  for (int i = 0; i < 8; ++i) {
    const V3 p = box_corner(lb, i);
    V3 n; float depth;
    if (!point_in_box(p, sb, 0.0f, n, depth)) continue;
    const float fn = P.penalty_stiffness * depth;
    const V3 v = point_vel(L, p - L.x);
    const float vn = dot(v, n);
    const V3 t = v - vn * n;
    const float vt = sqrtf(dot(t, t));
    V3 f = (-fn) * n;
    if (vt > 1e-6f) f = f + (mu * fn / vt) * t;
    f_acc = f_acc + f;
  }
}

// One step() of the integrator, one thread per environment (the sim facade, the producer of the thread-per-sample
// rollout kernel and that kernel itself). Restated lane-cooperatively in panda_team.cuh.
DEV void panda_step(PandaEnv& e, const PandaParams& P, const float* u, float dt, int substeps, int passes) {
  const float h = dt / (float)substeps;
  const float D = P.drive_damping;
  V3 imp_table = mk(0, 0, 0), imp_shelf = mk(0, 0, 0), imp_cubeb = mk(0, 0, 0);
  V3 pen_table = mk(0, 0, 0), pen_shelf = mk(0, 0, 0);
  // accumulated contact impulses of the current sub-step, kept for the warm start of the next one (a step() starts
  // cold). Slots: [0,16) cube i vs its first near fixed box; [16,32) cube/cube (corners of A in B, of B in A);
  // 32 + 32 f + 16 i + 8 ph: link f vs cube i (ph 0: link corners in the cube, 1: cube corners in the link).
  float4 lam[128];
  unsigned prev[4] = {0u, 0u, 0u, 0u};   // bit per slot: in contact in the previous sub-step
  for (int s = 0; s < substeps; ++s) {
    // 1. joint drives: implicit velocity tracking, effort- and speed-limited, no motion into a joint limit
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const float m = j < 7 ? P.joint_inertia[j < 7 ? j : 0] : P.finger_mass;
      const float v = e.qd[j];
      float vs = (m * v + h * D * u[j]) / (m + h * D);
      const float f = D * (u[j] - vs);
      if (f > P.effort[j]) vs = v + h * P.effort[j] / m;
      else if (f < -P.effort[j]) vs = v - h * P.effort[j] / m;
      vs = clampf(vs, -P.qd_limit[j], P.qd_limit[j]);
      if (e.q[j] <= P.q_lower[j] && vs < 0.0f) vs = 0.0f;
      if (e.q[j] >= P.q_upper[j] && vs > 0.0f) vs = 0.0f;
      e.qd[j] = vs;
    }
    // link and cube boxes of this sub-step (positions are fixed until step 5)
    Hand H;
    panda_hand(P, e.q, e.qd, true, H);
    OBox3 lbox[3];
    Dyn3 L[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const float* cen = f < 2 ? P.finger_center : P.hand_center;
      const float* half = f < 2 ? P.finger_half : P.hand_half;
      V3 l = mk(cen[0], cen[1], cen[2]);
      if (f == 0) { l.y += e.q[7]; l.z += kFingerZ; }
      if (f == 1) { l.y = -l.y - e.q[8]; l.z += kFingerZ; }  // mirrored finger geometry (urdf:220)
      lbox[f].c = H.p + mul(H.R, l);
      lbox[f].R = H.R;
      lbox[f].half = mk(half[0], half[1], half[2]);
      L[f].v = H.v; L[f].w = H.w; L[f].x = H.p; L[f].im = 0.0f; L[f].ii = 0.0f;
      if (f < 2) {
        L[f].axis = (f == 0 ? 1.0f : -1.0f) * H.R.cy;
        L[f].slide = e.qd[7 + f];
        L[f].ims = 1.0f / P.finger_mass;
      } else {
        L[f].axis = mk(0, 0, 0); L[f].slide = 0.0f; L[f].ims = 0.0f;
      }
    }
    // dormant cubes: both slept through the previous sub-step and no link box comes within the sum of the bounding radii
    // (+ margin) of either: every test below would repeat its result (nothing of the cubes moved, boxes_near fails for
    // every link), so only the weight the supports carry is booked and the contact stage is skipped
    bool dormant = e.slept == 3u;
    if (dormant) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const V3 hc = mk(P.cube_half[i][0], P.cube_half[i][1], P.cube_half[i][2]);
        const float rc = sqrtf(dot(hc, hc)) + P.contact_margin;
#pragma unroll
        for (int f = 0; f < 3; ++f) {   // the bounding-sphere part of boxes_near(lbox[f], cbox[i]): a superset of lnear
          const V3 dl = lbox[f].c - e.cube[i].p;
          const float rr = sqrtf(dot(lbox[f].half, lbox[f].half)) + rc;
          dormant = dormant && dot(dl, dl) > rr * rr;
        }
      }
    }
    if (dormant) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float wgt = P.cube_mass[i] * P.gravity * h;
        if (e.support[i] == P.idx_table && P.report_cube) imp_table.z -= wgt;
        if (e.support[i] == P.idx_shelf && P.report_cube) imp_shelf.z -= wgt;
        if (i == 1) imp_cubeb.z += wgt;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) prev[q] = 0u;
    } else {
    OBox3 cbox[2];
    Dyn3 C[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      cbox[i] = obox_of(e.cube[i], P.cube_half[i]);
      C[i].v = e.cube[i].v; C[i].w = e.cube[i].w; C[i].x = e.cube[i].p;
      C[i].im = 1.0f / P.cube_mass[i]; C[i].ii = 1.0f / P.cube_inertia[i];
      C[i].axis = mk(0, 0, 0); C[i].slide = 0.0f; C[i].ims = 0.0f;
    }
    // what is close to what, decided once per sub-step: link f / cube i (bit 2 f + i), fixed boxes per cube
    const bool cc_near = boxes_near(cbox[0], cbox[1], P.contact_margin);
    unsigned lnear = 0u, snear[2] = {0u, 0u};
    int first_box[2] = {-1, -1};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int f = 0; f < 3; ++f)
        if (boxes_near(lbox[f], cbox[i], P.contact_margin)) lnear |= 1u << (2 * f + i);
      for (int k = 0; k < P.n_static; ++k)
        if (boxes_near(cbox[i], obox_of(P.st[k]), P.contact_margin)) { snear[i] |= 1u << k; if (first_box[i] < 0) first_box[i] = k; }
    }
    // 2. sleeping: an (almost) motionless cube that rests on its first near fixed box with at least three corners and
    // has no link and no other cube within the contact margin is neither moved nor solved in this sub-step
    bool asleep[2] = {false, false};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!(P.sleep_lin > 0.0f)) continue;
      const Cube& c = e.cube[i];
      if (!(dot(c.v, c.v) < P.sleep_lin * P.sleep_lin && dot(c.w, c.w) < P.sleep_ang * P.sleep_ang)) continue;
      if (cc_near || ((lnear >> i) & 0x15u) || first_box[i] < 0) continue;
      const OBox3 sb = obox_of(P.st[first_box[i]]);
      int cnt = 0;
      for (int ci = 0; ci < 8; ++ci) {
        V3 n; float depth;
        if (point_in_box(box_corner(cbox[i], ci), sb, P.contact_margin, n, depth) && depth > -P.sleep_gap) ++cnt;
      }
      if (cnt < 3) continue;
      asleep[i] = true;
      e.cube[i].v = mk(0, 0, 0); e.cube[i].w = mk(0, 0, 0);
      C[i].v = mk(0, 0, 0); C[i].w = mk(0, 0, 0);
      const float wgt = P.cube_mass[i] * P.gravity * h;   // the support carries the weight
      if (first_box[i] == P.idx_table && P.report_cube) imp_table.z -= wgt;
      if (first_box[i] == P.idx_shelf && P.report_cube) imp_shelf.z -= wgt;
      if (i == 1) imp_cubeb.z += wgt;
    }
    // 3. gravity
#pragma unroll
    for (int i = 0; i < 2; ++i) if (!asleep[i]) C[i].v.z -= P.gravity * h;
    // 4. contacts. Every pass: links (link_sweeps sweeps: the finger - cube - finger chain of a grasp) and cube/cube
    // first, the fixed boxes LAST (what a kinematic link pushes into the table is pushed back out in the same pass).
    unsigned cur[4] = {0u, 0u, 0u, 0u};
    for (int p = 0; p < passes; ++p) {
      for (int sw = 0; sw < P.link_sweeps && lnear; ++sw) {
        const bool first = p == 0 && sw == 0;
        int room[2] = {kLinkCap, kLinkCap};   // link contacts per cube (the same ones in every sweep: positions are fixed)
        for (int f = 0; f < 3; ++f)
          for (int i = 0; i < 2; ++i) {
            if (asleep[i] || !((lnear >> (2 * f + i)) & 1u)) continue;
            const float mu = 0.5f * (P.robot_mu + P.cube_mu[i]);
            const int s0 = 32 + 32 * f + 16 * i;
            unsigned c0 = 0u, c1 = 0u;
            V3 got = corners_vs_box3<false>(L[f], lbox[f], C[i], cbox[i], mu, h, P, lam + s0, (prev[s0 >> 5] >> (s0 & 31)) & 0xffu, c0, first, room[i]);
            got = got + corners_vs_box3<true>(C[i], cbox[i], L[f], lbox[f], mu, h, P, lam + s0 + 8, (prev[s0 >> 5] >> ((s0 & 31) + 8)) & 0xffu, c1, first, room[i]);
            cur[s0 >> 5] |= (c0 | (c1 << 8)) << (s0 & 31);
            if (i == 1) imp_cubeb = imp_cubeb - got;   // `got` = impulse on the link; cubeB received the opposite
          }
      }
      if (cc_near) {
        const float mu = 0.5f * (P.cube_mu[0] + P.cube_mu[1]);
        unsigned c0 = 0u, c1 = 0u;
        int room = 16;
        V3 got = corners_vs_box3<false>(C[0], cbox[0], C[1], cbox[1], mu, h, P, lam + 16, (prev[0] >> 16) & 0xffu, c0, p == 0, room);
        got = got + corners_vs_box3<true>(C[1], cbox[1], C[0], cbox[0], mu, h, P, lam + 24, (prev[0] >> 24) & 0xffu, c1, p == 0, room);
        cur[0] |= (c0 << 16) | (c1 << 24);
        imp_cubeb = imp_cubeb - got;                   // `got` = impulse on cubeA
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (asleep[i]) continue;
        for (unsigned todo = snear[i]; todo; todo &= todo - 1) {
          const int k = __ffs(todo) - 1;
          const OBox3 sb = obox_of(P.st[k]);
          Dyn3 S;
          S.v = mk(0, 0, 0); S.w = mk(0, 0, 0); S.x = sb.c; S.im = 0.0f; S.ii = 0.0f; S.axis = mk(0, 0, 0); S.slide = 0.0f; S.ims = 0.0f;
          const bool acc = k == first_box[i];
          unsigned c0 = 0u;
          int room = 8;
          const V3 got = corners_vs_box3<false>(C[i], cbox[i], S, sb, 0.5f * (P.cube_mu[i] + P.st[k].mu), h, P,
                                                acc ? lam + 8 * i : nullptr, (prev[0] >> (8 * i)) & 0xffu, c0, p == 0, room);
          if (acc) cur[0] |= c0 << (8 * i);
          if (k == P.idx_table && P.report_cube) imp_table = imp_table - got;
          if (k == P.idx_shelf && P.report_cube) imp_shelf = imp_shelf - got;
          if (i == 1) imp_cubeb = imp_cubeb + got;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) prev[q] = cur[q];
#pragma unroll
    for (int i = 0; i < 2; ++i) { e.cube[i].v = C[i].v; e.cube[i].w = C[i].w; }
    e.slept = (asleep[0] ? 1u : 0u) | (asleep[1] ? 2u : 0u);
    e.support[0] = first_box[0]; e.support[1] = first_box[1];
    }   // not dormant
    const bool moved[2] = {!(e.slept & 1u), !(e.slept & 2u)};
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      e.qd[7 + f] = clampf(L[f].slide, -P.qd_limit[7 + f], P.qd_limit[7 + f]);
      L[f].slide = e.qd[7 + f];
    }
    // kinematic links against the fixed bodies named by the collision cost
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      if (P.idx_table >= 0)
        link_vs_static(lbox[f], L[f], obox_of(P.st[P.idx_table]), 0.5f * (P.robot_mu + P.st[P.idx_table].mu), P, pen_table);
      if (P.idx_shelf >= 0)
        link_vs_static(lbox[f], L[f], obox_of(P.st[P.idx_shelf]), 0.5f * (P.robot_mu + P.st[P.idx_shelf].mu), P, pen_shelf);
    }
    // 5. positions
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float qn = e.q[j] + h * e.qd[j];
      if (qn < P.q_lower[j]) { qn = P.q_lower[j]; e.qd[j] = 0.0f; }
      if (qn > P.q_upper[j]) { qn = P.q_upper[j]; e.qd[j] = 0.0f; }
      e.q[j] = qn;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!moved[i]) continue;
      Cube& c = e.cube[i];
      c.p = c.p + h * c.v;
      const float x = c.qx, y = c.qy, z = c.qz, w = c.qw, hh = 0.5f * h;
      const float nx = x + hh * (c.w.x * w + c.w.y * z - c.w.z * y);
      const float ny = y + hh * (c.w.y * w + c.w.z * x - c.w.x * z);
      const float nz = z + hh * (c.w.z * w + c.w.x * y - c.w.y * x);
      const float nw = w - hh * (c.w.x * x + c.w.y * y + c.w.z * z);
      const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
      c.qx = nx * inv; c.qy = ny * inv; c.qz = nz * inv; c.qw = nw * inv;
    }
  }
  const float inv_dt = 1.0f / dt, inv_ns = 1.0f / (float)substeps;
  e.f_table = inv_dt * imp_table + inv_ns * pen_table;
  e.f_shelf = inv_dt * imp_shelf + inv_ns * pen_shelf;
  e.f_cubeb = inv_dt * imp_cubeb;
}

// ------------------------------------------------------------------ costs
// the reference's own quaternion -> matrix formula (skill_utils.py:140-180)
DEV M33 ref_rotmat(float x, float y, float z, float w) {
  const float q0 = w, q1 = x, q2 = y, q3 = z;
  M33 R;
  R.cx = mk(2.0f * (q0 * q0 + q1 * q1) - 1.0f, 2.0f * (q1 * q2 + q0 * q3), 2.0f * (q1 * q3 - q0 * q2));
  R.cy = mk(2.0f * (q1 * q2 - q0 * q3), 2.0f * (q0 * q0 + q2 * q2) - 1.0f, 2.0f * (q2 * q3 + q0 * q1));
  R.cz = mk(2.0f * (q1 * q3 + q0 * q2), 2.0f * (q2 * q3 - q0 * q1), 2.0f * (q0 * q0 + q3 * q3) - 1.0f);
  return R;
}
// min_j (1 - |<a, C[:,j]>|)
DEV float min_axis_cost(V3 a, const M33& C) {
  return fminf(fminf(1.0f - fabsf(dot(a, C.cx)), 1.0f - fabsf(dot(a, C.cy))), 1.0f - fabsf(dot(a, C.cz)));
}
DEV int sel_axis_of(const Cube& c) {
  const M33 C = ref_rotmat(c.qx, c.qy, c.qz, c.qw);
  int b = 0;
  float m = fabsf(C.cx.x);
  if (fabsf(C.cy.x) > m) { b = 1; m = fabsf(C.cy.x); }
  if (fabsf(C.cz.x) > m) { b = 2; }
  return b;
}

DEV float panda_motion_cost(const PandaEnv& e) {
  const float fx = e.f_table.x + 4.0f * e.f_shelf.x + e.f_cubeb.x;
  const float fy = e.f_table.y + 4.0f * e.f_shelf.y + e.f_cubeb.y;
  return (fabsf(fx) + fabsf(fy)) > 0.1f ? 1000.0f : 0.0f;
}

// task cost given the hand pose (cost_functions.py:91-136): cubeA pose, finger openings q7/q8, contact-force cost
DEV float panda_cost_from_hand(const Hand& H, float q7, float q8, const Cube& cubeA, float motion_cost,
                               const RolloutCfg& c, int kg, const PandaRef& ref) {
  const bool second = c.multi_modal && kg >= c.Kg / 2;
  switch (c.task) {
    case M3P2I_TASK_REACH: {
      // ee = mean of the two finger frames (cost_functions.py:92-94)
      const V3 lf = H.p + mul(H.R, mk(0.0f, q7, kFingerZ)), rf = H.p + mul(H.R, mk(0.0f, -q8, kFingerZ));
      V3 g = mk(ref.cube0[0], ref.cube0[1], ref.cube0[2]);
      if (!second) g.z += c.pre_height_diff;
      else {
        g.x -= c.pre_height_diff * c.tilt_cos;
        g.z += c.pre_height_diff * sqrtf(1.0f - c.tilt_cos * c.tilt_cos);
      }
      const V3 d = mk((lf.x + rf.x) / 2.0f - g.x, (lf.y + rf.y) / 2.0f - g.y, (lf.z + rf.z) / 2.0f - g.z);
      const float reach = sqrtf(dot(d, d));
      const M33 C = ref_rotmat(cubeA.qx, cubeA.qy, cubeA.qz, cubeA.qw);
      float cost_z;
      if (!second) cost_z = min_axis_cost(H.R.cz, C);
      else cost_z = fabsf(c.tilt_cos - dot(H.R.cz, col(C, ref.sel_axis)));
      const float tilt = cost_z + min_axis_cost(H.R.cy, C);
      return 10.0f * reach + 3.0f * tilt;
    }
    case M3P2I_TASK_PICK: {
      const V3 d = mk(c.goal[0] - cubeA.p.x, c.goal[1] - cubeA.p.y, c.goal[2] - cubeA.p.z);
      const M33 C = ref_rotmat(cubeA.qx, cubeA.qy, cubeA.qz, cubeA.qw);
      const M33 G = ref_rotmat(c.goal[3], c.goal[4], c.goal[5], c.goal[6]);
      const float ori = min_axis_cost(G.cx, C) + min_axis_cost(G.cy, C);
      return 10.0f * sqrtf(dot(d, d)) + 15.0f * ori + motion_cost;
    }
    case M3P2I_TASK_PLACE: {
      const V3 lf = H.p + mul(H.R, mk(0.0f, q7, kFingerZ)), rf = H.p + mul(H.R, mk(0.0f, -q8, kFingerZ));
      const V3 d = lf - rf;
      return 2.0f * (1.0f - sqrtf(dot(d, d)));
    }
    default: return 0.0f;
  }
}

// The reach cost (cost_functions.py:91-114,138-156) split at the point where rows of OTHER samples enter it.
// reach_parts(): everything a sample knows by itself after a step -- ee position, the three dot products of the hand's
// z axis with the own cube's axes (the one picked by `sel_axis` of another row is chosen later; first mode: slot 0
// holds the finished cost_z), and the y-axis alignment term. reach_combine(): the cost once the batch rows are known.
// Same arithmetic, same order as panda_cost_from_hand's reach branch.
struct ReachParts {
  V3 ee, dz;
  float min_y;
};
DEV ReachParts reach_parts(const Hand& H, float q7, float q8, const Cube& cubeA, const RolloutCfg& c, int kg) {
  const bool second = c.multi_modal && kg >= c.Kg / 2;
  const V3 lf = H.p + mul(H.R, mk(0.0f, q7, kFingerZ)), rf = H.p + mul(H.R, mk(0.0f, -q8, kFingerZ));
  ReachParts r;
  r.ee = mk((lf.x + rf.x) / 2.0f, (lf.y + rf.y) / 2.0f, (lf.z + rf.z) / 2.0f);
  const M33 C = ref_rotmat(cubeA.qx, cubeA.qy, cubeA.qz, cubeA.qw);
  if (!second) r.dz = mk(min_axis_cost(H.R.cz, C), 0.0f, 0.0f);
  else r.dz = mk(dot(H.R.cz, C.cx), dot(H.R.cz, C.cy), dot(H.R.cz, C.cz));
  r.min_y = min_axis_cost(H.R.cy, C);
  return r;
}
DEV float reach_combine(const ReachParts& r, const RolloutCfg& c, int kg, const PandaRef& ref) {
  const bool second = c.multi_modal && kg >= c.Kg / 2;
  V3 g = mk(ref.cube0[0], ref.cube0[1], ref.cube0[2]);
  if (!second) g.z += c.pre_height_diff;
  else {
    g.x -= c.pre_height_diff * c.tilt_cos;
    g.z += c.pre_height_diff * sqrtf(1.0f - c.tilt_cos * c.tilt_cos);
  }
  const V3 d = mk(r.ee.x - g.x, r.ee.y - g.y, r.ee.z - g.z);
  const float reach = sqrtf(dot(d, d));
  const float cost_z = !second ? r.dz.x : fabsf(c.tilt_cos - comp(r.dz, ref.sel_axis));
  const float tilt = cost_z + r.min_y;
  return 10.0f * reach + 3.0f * tilt;
}

DEV float panda_cost(const PandaEnv& e, const PandaParams& P, const RolloutCfg& c, int kg, const PandaRef& ref) {
  Hand H;
  H.p = mk(0, 0, 0); H.R.cx = mk(1, 0, 0); H.R.cy = mk(0, 1, 0); H.R.cz = mk(0, 0, 1); H.v = mk(0, 0, 0); H.w = mk(0, 0, 0);
  if (c.task != M3P2I_TASK_PICK) panda_hand(P, e.q, e.qd, false, H);
  return panda_cost_from_hand(H, e.q[7], e.q[8], e.cube[0], panda_motion_cost(e), c, kg, ref);
}

}  // namespace m3
