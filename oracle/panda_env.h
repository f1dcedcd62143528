/*
 * panda_env.h — CPU restatement of the Panda (7 revolute + 2 prismatic finger DoF) + cubes integrator and
 * of the panda_env task costs. TEST INFRASTRUCTURE ONLY (see m3p2i_oracle.h).
 *
 * Integrator spec (ours; stands in for IsaacGym PhysX, isaacgym_wrapper.py:354-360, PARITY UNPINNED):
 *   arm     : franka_panda.urdf has no <inertial> blocks and the velocity drive has damping 600
 *             (isaacgym_wrapper.py:341-344), gravity off (panda_env/panda.yaml): joints 1-7 are a
 *             velocity-tracked kinematic chain, qd <- implicit drive with reflected inertia `arm_inertia`,
 *             effort / velocity / position limits of the URDF.
 *   fingers : prismatic DoF of mass `finger_mass`, drive force limited to the URDF effort (20 N) and speed
 *             to the URDF velocity limit (0.2 m/s). Contacts act on this one DoF (the other 6 DoF of the
 *             finger box follow the kinematic hand), so a grasp squeezes with finite force.
 *   cubes   : cubeA, cubeB are 3-D rigid boxes (isotropic inertia) under gravity. Contacts: box corners
 *             against the signed distance field of the other box (cube-static, cube-cube, finger/hand-cube),
 *             velocity-level Gauss-Seidel impulses with Coulomb friction, speculative margin and Baumgarte
 *             feedback, `solver_passes` sweeps per substep in a fixed pair order, each with `link_sweeps` (4)
 *             sweeps over the finger/hand-cube contacts (a squeezed cube settles inside the sub-step).
 *             Semi-implicit Euler.
 *   links vs statics: the kinematic hand/finger boxes do not stop at the table / shelf; their penetration is
 *             reported as a penalty contact force (k * depth + Coulomb friction) on the static body, which is
 *             what the collision cost reads (cost_functions.py:158-169). The cubes' own contacts with the table /
 *             shelf stand enter that report only when `report_cube_contacts` is set.
 * FK restates franka_panda.urdf:27-242 (URDF fixed-axis rpy, joint axes +z, fingers +y / -y).
 * Costs restate cost_functions.py:91-169 and skill_utils.py:140-180,224-289.
 */
#ifndef ORACLE_PANDA_ENV_H
#define ORACLE_PANDA_ENV_H

#include <math.h>
#include <string.h>
#include "../include/m3p2i_b200.h"

typedef struct {
  float p[3], q[4], v[3], w[3]; /* position, quaternion xyzw, linear, angular velocity */
} OCube;

/* accumulated impulse of one contact slot: normal magnitude and tangential (world) vector; `stamp` = 1-based index of
 * the last sub-step (of the current step) in which the slot held a contact; `fresh` = a warm-start impulse is pending */
typedef struct { float n, t[3]; int stamp, fresh; } OLam;
typedef struct {
  OLam st[2][8];        /* cube i against its first near fixed box, corner */
  OLam cc[2][8];        /* 0: corners of cubeA in cubeB, 1: corners of cubeB in cubeA */
  OLam lk[3][2][2][8];  /* link f, cube i, 0: link corners in the cube / 1: cube corners in the link */
} OCache;

typedef struct {
  float q[9], qd[9];
  OCube cube[2];      /* 0 = cubeA, 1 = cubeB */
  float f_table[3], f_shelf[3], f_cubeb[3]; /* net contact force during the last step */
} OPandaEnv;

/* values of sample 0 / sample K/2 of the batch that other samples' costs read (cost_functions.py:98,102-103;
 * skill_utils.py:275-279) */
typedef struct {
  float cube0_pos[3];
  int sel_axis;
} OPandaRef;

/* ------------------------------------------------------------------ small vector helpers */
static inline float q_dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void q_cross(const float* a, const float* b, float* o) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline float q_clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* R (row-major 3x3) from quaternion xyzw, standard Hamilton form */
static inline void q_quat_to_R(const float* q, float* R) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1.0f - 2.0f * (y * y + z * z); R[1] = 2.0f * (x * y - w * z); R[2] = 2.0f * (x * z + w * y);
  R[3] = 2.0f * (x * y + w * z); R[4] = 1.0f - 2.0f * (x * x + z * z); R[5] = 2.0f * (y * z - w * x);
  R[6] = 2.0f * (x * z - w * y); R[7] = 2.0f * (y * z + w * x); R[8] = 1.0f - 2.0f * (x * x + y * y);
}
/* quaternion xyzw from a rotation matrix (Shepperd, largest-pivot branch) */
static inline void q_R_to_quat(const float* R, float* q) {
  float tr = R[0] + R[4] + R[8];
  if (tr > 0.0f) {
    float s = sqrtf(tr + 1.0f) * 2.0f;
    q[3] = 0.25f * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    float s = sqrtf(1.0f + R[0] - R[4] - R[8]) * 2.0f;
    q[3] = (R[7] - R[5]) / s; q[0] = 0.25f * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    float s = sqrtf(1.0f + R[4] - R[0] - R[8]) * 2.0f;
    q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25f * s; q[2] = (R[5] + R[7]) / s;
  } else {
    float s = sqrtf(1.0f + R[8] - R[0] - R[4]) * 2.0f;
    q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25f * s;
  }
}

/* ------------------------------------------------------------------ forward kinematics */
/* franka_panda.urdf joint origins (xyz) and roll about x in quarter turns (rpy = (roll,0,0)); Appendix B */
static const float O_PANDA_XYZ[7][3] = {{0.0f, 0.0f, 0.333f},   {0.0f, 0.0f, 0.0f},      {0.0f, -0.316f, 0.0f},
                                        {0.0825f, 0.0f, 0.0f},  {-0.0825f, 0.384f, 0.0f}, {0.0f, 0.0f, 0.0f},
                                        {0.088f, 0.0f, 0.0f}};
static const int O_PANDA_ROLL[7] = {0, -1, 1, 1, -1, 1, 1}; /* multiples of pi/2 */
#define O_HAND_Z 0.107f        /* panda_hand_joint origin, urdf:183 */
#define O_HAND_YAW -0.785398163397f
#define O_FINGER_Z 0.0584f     /* panda_finger_joint1/2 origin, urdf:229,237 */

typedef struct {
  float p[3], R[9];  /* hand frame */
  float v[3], w[3];  /* twist of the hand frame origin (arm joints only) */
} OHand;

/* hand pose + twist from the 7 arm joints */
static inline void o_panda_hand(const M3P2IPandaScene* s, const float* q, const float* qd, OHand* H) {
  float p[3] = {s->base_pos[0], s->base_pos[1], s->base_pos[2]};
  float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float ax[7][3], org[7][3];
  for (int j = 0; j < 7; ++j) {
    /* p += R * xyz */
    const float* t = O_PANDA_XYZ[j];
    for (int r = 0; r < 3; ++r) p[r] += R[3 * r] * t[0] + R[3 * r + 1] * t[1] + R[3 * r + 2] * t[2];
    /* R <- R * Rx(roll): columns (c0, c1, c2) -> roll=+1: (c0, c2, -c1); roll=-1: (c0, -c2, c1) */
    if (O_PANDA_ROLL[j] != 0) {
      float sg = (float)O_PANDA_ROLL[j];
      for (int r = 0; r < 3; ++r) {
        float c1 = R[3 * r + 1], c2 = R[3 * r + 2];
        R[3 * r + 1] = sg * c2; R[3 * r + 2] = -sg * c1;
      }
    }
    for (int r = 0; r < 3; ++r) { ax[j][r] = R[3 * r + 2]; org[j][r] = p[r]; }
    /* R <- R * Rz(q_j) */
    float c = cosf(q[j]), sn = sinf(q[j]);
    for (int r = 0; r < 3; ++r) {
      float c0 = R[3 * r], c1 = R[3 * r + 1];
      R[3 * r] = c * c0 + sn * c1; R[3 * r + 1] = -sn * c0 + c * c1;
    }
  }
  /* fixed hand joint: translate z, yaw -pi/4 */
  for (int r = 0; r < 3; ++r) p[r] += R[3 * r + 2] * O_HAND_Z;
  {
    float c = cosf(O_HAND_YAW), sn = sinf(O_HAND_YAW);
    for (int r = 0; r < 3; ++r) {
      float c0 = R[3 * r], c1 = R[3 * r + 1];
      R[3 * r] = c * c0 + sn * c1; R[3 * r + 1] = -sn * c0 + c * c1;
    }
  }
  memcpy(H->p, p, sizeof(p)); memcpy(H->R, R, sizeof(R));
  H->v[0] = H->v[1] = H->v[2] = 0.0f; H->w[0] = H->w[1] = H->w[2] = 0.0f;
  for (int j = 0; j < 7; ++j) {
    float r[3] = {p[0] - org[j][0], p[1] - org[j][1], p[2] - org[j][2]}, c[3];
    q_cross(ax[j], r, c);
    for (int i = 0; i < 3; ++i) { H->v[i] += qd[j] * c[i]; H->w[i] += qd[j] * ax[j][i]; }
  }
}

/* rigid-body rows [3][13] (leftfinger, rightfinger, hand): pos3, quat4 xyzw, linvel3, angvel3 */
static inline void o_panda_links(const M3P2IPandaScene* s, const float* q, const float* qd, float* out) {
  OHand H;
  o_panda_hand(s, q, qd, &H);
  float quat[4];
  q_R_to_quat(H.R, quat);
  const float sgn[2] = {1.0f, -1.0f};
  for (int f = 0; f < 2; ++f) {
    float l[3] = {0.0f, sgn[f] * q[7 + f], O_FINGER_Z}, r[3], c[3];
    float* o = out + 13 * f;
    for (int i = 0; i < 3; ++i) r[i] = H.R[3 * i] * l[0] + H.R[3 * i + 1] * l[1] + H.R[3 * i + 2] * l[2];
    q_cross(H.w, r, c);
    for (int i = 0; i < 3; ++i) {
      o[i] = H.p[i] + r[i];
      o[7 + i] = H.v[i] + c[i] + sgn[f] * qd[7 + f] * H.R[3 * i + 1];
      o[10 + i] = H.w[i];
    }
    memcpy(o + 3, quat, sizeof(quat));
  }
  float* o = out + 26;
  for (int i = 0; i < 3; ++i) { o[i] = H.p[i]; o[7 + i] = H.v[i]; o[10 + i] = H.w[i]; }
  memcpy(o + 3, quat, sizeof(quat));
}

/* ------------------------------------------------------------------ contact solver */
/* A body as seen by the contact solver. Dynamic cube: v,w point at its velocities, im/ii > 0.
 * Kinematic box: v,w point at a (read-only) twist, im = ii = 0. Finger: kinematic twist + one sliding DoF. */
typedef struct {
  float* v; float* w;
  float im, ii;
  float x[3];          /* point the twist refers to */
  float* slide;        /* finger slide velocity along `axis` (NULL otherwise) */
  float axis[3];
  float ims;           /* inverse mass of the sliding DoF */
  float* acc;          /* optional [3] accumulator of impulses received */
} OSolv3;

static inline void o_point_vel(const OSolv3* B, const float* r, float* out) {
  float c[3] = {0, 0, 0};
  if (B->w) q_cross(B->w, r, c);
  for (int i = 0; i < 3; ++i) {
    out[i] = (B->v ? B->v[i] : 0.0f) + c[i];
    if (B->slide) out[i] += *B->slide * B->axis[i];
  }
}
static inline float o_eff_mass(const OSolv3* B, const float* r, const float* d) {
  float c[3];
  q_cross(r, d, c);
  float k = B->im + B->ii * q_dot3(c, c);
  if (B->slide) { float a = q_dot3(B->axis, d); k += B->ims * a * a; }
  return k;
}
static inline void o_apply_impulse(OSolv3* B, const float* r, const float* P, float sgn) {
  if (B->im > 0.0f) {
    float c[3];
    q_cross(r, P, c);
    for (int i = 0; i < 3; ++i) { B->v[i] += sgn * B->im * P[i]; B->w[i] += sgn * B->ii * c[i]; }
  }
  if (B->slide) *B->slide += sgn * B->ims * q_dot3(B->axis, P);
  if (B->acc) for (int i = 0; i < 3; ++i) B->acc[i] += sgn * P[i];
}

/* one contact: normal n from B to A, depth > 0 = penetration, at world point c */
static inline void o_solve_contact3(OSolv3* A, OSolv3* B, const float* n, float depth, const float* c, float mu,
                                    float h, const M3P2IPandaScene* sc) {
  float ra[3], rb[3], va[3], vb[3], rv[3];
  for (int i = 0; i < 3; ++i) { ra[i] = c[i] - A->x[i]; rb[i] = c[i] - B->x[i]; }
  o_point_vel(A, ra, va); o_point_vel(B, rb, vb);
  for (int i = 0; i < 3; ++i) rv[i] = va[i] - vb[i];
  float vn = q_dot3(rv, n);
  float kn = o_eff_mass(A, ra, n) + o_eff_mass(B, rb, n);
  if (kn <= 0.0f) return;
  float target;
  if (depth > 0.0f) {
    float pen = depth - sc->slop;
    if (pen < 0.0f) pen = 0.0f;
    target = sc->baumgarte * pen / h;
    if (target > sc->max_corr_vel) target = sc->max_corr_vel;
  } else {
    target = depth / h;
  }
  float jn = (target - vn) / kn;
  if (jn <= 0.0f) return;
  float P[3] = {jn * n[0], jn * n[1], jn * n[2]};
  o_apply_impulse(A, ra, P, 1.0f); o_apply_impulse(B, rb, P, -1.0f);
  /* Coulomb friction against the tangential relative velocity after the normal impulse */
  o_point_vel(A, ra, va); o_point_vel(B, rb, vb);
  for (int i = 0; i < 3; ++i) rv[i] = va[i] - vb[i];
  vn = q_dot3(rv, n);
  float t[3] = {rv[0] - vn * n[0], rv[1] - vn * n[1], rv[2] - vn * n[2]};
  float vt = sqrtf(q_dot3(t, t));
  if (vt < 1e-9f) return;
  for (int i = 0; i < 3; ++i) t[i] /= vt;
  float kt = o_eff_mass(A, ra, t) + o_eff_mass(B, rb, t);
  if (kt <= 0.0f) return;
  float jt = vt / kt;
  if (jt > mu * jn) jt = mu * jn;
  float Pt[3] = {-jt * t[0], -jt * t[1], -jt * t[2]};
  o_apply_impulse(A, ra, Pt, 1.0f); o_apply_impulse(B, rb, Pt, -1.0f);
}

/* oriented box for the narrow phase */
typedef struct {
  float c[3], R[9], half[3];
} OBox3;

static inline float o_box_radius(const OBox3* b) { return sqrtf(q_dot3(b->half, b->half)); }

/* signed-distance test of world point p against box b: returns 0 when farther than `margin`;
 * else normal (world, pointing out of b) and depth (= -sdf along the least-penetration axis) */
static inline int o_point_in_box(const float* p, const OBox3* b, float margin, float* n, float* depth) {
  float o[3] = {p[0] - b->c[0], p[1] - b->c[1], p[2] - b->c[2]}, d[3], qv[3];
  for (int i = 0; i < 3; ++i) {
    d[i] = b->R[i] * o[0] + b->R[3 + i] * o[1] + b->R[6 + i] * o[2]; /* R^T o */
    qv[i] = fabsf(d[i]) - b->half[i];
  }
  int ax = 0;
  if (qv[1] > qv[ax]) ax = 1;
  if (qv[2] > qv[ax]) ax = 2;
  if (qv[ax] >= margin) return 0;
  float sg = d[ax] < 0.0f ? -1.0f : 1.0f;
  for (int i = 0; i < 3; ++i) n[i] = sg * b->R[3 * i + ax];
  *depth = -qv[ax];
  return 1;
}

static inline void o_box_corner(const OBox3* b, int i, float* p) {
  float l[3] = {(i & 1) ? b->half[0] : -b->half[0], (i & 2) ? b->half[1] : -b->half[1],
                (i & 4) ? b->half[2] : -b->half[2]};
  for (int r = 0; r < 3; ++r) p[r] = b->c[r] + b->R[3 * r] * l[0] + b->R[3 * r + 1] * l[1] + b->R[3 * r + 2] * l[2];
}

/* corners of box `ba` (body A) against the SDF of `bb` (body B) */
static inline void o_corners_vs_box3(OSolv3* A, const OBox3* ba, OSolv3* B, const OBox3* bb, float mu, float h,
                                     const M3P2IPandaScene* sc, int flip) {
  for (int i = 0; i < 8; ++i) {
    float p[3], n[3], depth;
    o_box_corner(ba, i, p);
    if (!o_point_in_box(p, bb, sc->contact_margin, n, &depth)) continue;
    if (!flip) o_solve_contact3(A, B, n, depth, p, mu, h, sc);
    else { float m[3] = {-n[0], -n[1], -n[2]}; o_solve_contact3(B, A, m, depth, p, mu, h, sc); }
  }
}

static inline int o_boxes_near(const OBox3* a, const OBox3* b, float margin) {
  /* sphere of a against the box b (b may be large and flat) */
  float o[3] = {a->c[0] - b->c[0], a->c[1] - b->c[1], a->c[2] - b->c[2]};
  float ra = o_box_radius(a) + margin, d2 = 0.0f;
  for (int i = 0; i < 3; ++i) {
    float d = fabsf(b->R[i] * o[0] + b->R[3 + i] * o[1] + b->R[6 + i] * o[2]) - b->half[i];
    if (d > 0.0f) d2 += d * d;
  }
  return d2 <= ra * ra;
}

/* two-way corner/SDF contact between boxes of bodies A and B; `both` = 0 tests only A's corners */
static inline void o_box_vs_box3(OSolv3* A, const OBox3* ba, OSolv3* B, const OBox3* bb, float mu, float h,
                                 const M3P2IPandaScene* sc, int both) {
  if (!o_boxes_near(ba, bb, sc->contact_margin)) return;
  o_corners_vs_box3(A, ba, B, bb, mu, h, sc, 0);
  if (both) o_corners_vs_box3(B, bb, A, ba, mu, h, sc, 1);
}

static inline OBox3 o_static_box3(const M3P2IBox* b) {
  OBox3 r;
  memcpy(r.c, b->pos, sizeof(r.c)); memcpy(r.half, b->half, sizeof(r.half));
  q_quat_to_R(b->quat, r.R);
  return r;
}
static inline OBox3 o_cube_box3(const OCube* c, const M3P2IBody* p) {
  OBox3 r;
  memcpy(r.c, c->p, sizeof(r.c)); memcpy(r.half, p->half, sizeof(r.half));
  q_quat_to_R(c->q, r.R);
  return r;
}

/* kinematic link box vs static box: penalty force on the static body (reported only) */
static inline void o_link_vs_static(const OBox3* lb, const OSolv3* L, const OBox3* sb, float mu,
                                    const M3P2IPandaScene* sc, float* f_acc) {
  if (!o_boxes_near(lb, sb, 0.0f)) return;
  for (int i = 0; i < 8; ++i) {
    float p[3], n[3], depth;
    o_box_corner(lb, i, p);
    if (!o_point_in_box(p, sb, 0.0f, n, &depth)) continue;
    float fn = sc->penalty_stiffness * depth;
    float r[3] = {p[0] - L->x[0], p[1] - L->x[1], p[2] - L->x[2]}, v[3];
    o_point_vel(L, r, v);
    float vn = q_dot3(v, n);
    float t[3] = {v[0] - vn * n[0], v[1] - vn * n[1], v[2] - vn * n[2]};
    float vt = sqrtf(q_dot3(t, t));
    /* force on the static body: pushed along -n by the link, dragged along the link's sliding direction */
    for (int k = 0; k < 3; ++k) {
      float f = -fn * n[k];
      if (vt > 1e-6f) f += mu * fn * t[k] / vt;
      f_acc[k] += f;
    }
  }
}

static inline void o_panda_init(OPandaEnv* e, const M3P2IPandaScene* sc, const float* dof, const float* root) {
  memset(e, 0, sizeof(*e));
  for (int j = 0; j < 9; ++j) { e->q[j] = dof[2 * j]; e->qd[j] = dof[2 * j + 1]; }
  const M3P2IBody* bp[2] = {&sc->cube_a, &sc->cube_b};
  for (int i = 0; i < 2; ++i) {
    const float* r = root + 13 * bp[i]->actor;
    memcpy(e->cube[i].p, r, 12); memcpy(e->cube[i].q, r + 3, 16);
    memcpy(e->cube[i].v, r + 7, 12); memcpy(e->cube[i].w, r + 10, 12);
  }
}

/* At most this many link / cube contacts per cube and sub-step are solved (detection order: finger 1, finger 2, hand;
 * link corners in the cube before cube corners in the link). The kernels keep the list in shared memory (kLinkCap). */
#define O_LINK_CAP 32

/* One detected contact of a sub-step. Positions are fixed inside a sub-step, so detection runs once; the passes
 * then visit the records in the fixed solve order. */
typedef struct {
  OSolv3 *A, *B;
  float n[3], depth, p[3], mu;
  OLam* L;       /* accumulator slot, or NULL: plain (non-accumulated) solve */
} OContact;

/* Accumulated-impulse form of o_solve_contact3: the normal impulse is clamped on its running total (a later visit may
 * take back what an earlier one over-applied), the friction impulse is a tangential vector clamped to the Coulomb
 * cone of the TOTAL normal impulse. A slot that was in contact in the previous sub-step of the same step() first
 * re-applies `warm_start` times what it held then (o_detect prepared it; applied here, at the slot's first visit). */
static inline void o_solve_contact3_acc(const OContact* c, float h, const M3P2IPandaScene* sc) {
  OSolv3 *A = c->A, *B = c->B;
  OLam* L = c->L;
  const float* n = c->n;
  float ra[3], rb[3], va[3], vb[3], rv[3];
  for (int i = 0; i < 3; ++i) { ra[i] = c->p[i] - A->x[i]; rb[i] = c->p[i] - B->x[i]; }
  if (L->fresh) {
    float P[3] = {L->n * n[0] + L->t[0], L->n * n[1] + L->t[1], L->n * n[2] + L->t[2]};
    o_apply_impulse(A, ra, P, 1.0f); o_apply_impulse(B, rb, P, -1.0f);
    L->fresh = 0;
  }
  o_point_vel(A, ra, va); o_point_vel(B, rb, vb);
  for (int i = 0; i < 3; ++i) rv[i] = va[i] - vb[i];
  float vn = q_dot3(rv, n);
  float kn = o_eff_mass(A, ra, n) + o_eff_mass(B, rb, n);
  if (kn <= 0.0f) return;
  float target;
  if (c->depth > 0.0f) {
    float pen = c->depth - sc->slop;
    if (pen < 0.0f) pen = 0.0f;
    target = sc->baumgarte * pen / h;
    if (target > sc->max_corr_vel) target = sc->max_corr_vel;
  } else {
    target = c->depth / h;
  }
  float ln = L->n + (target - vn) / kn;
  if (ln < 0.0f) ln = 0.0f;
  float dj = ln - L->n;
  L->n = ln;
  float P[3] = {dj * n[0], dj * n[1], dj * n[2]};
  o_apply_impulse(A, ra, P, 1.0f); o_apply_impulse(B, rb, P, -1.0f);
  o_point_vel(A, ra, va); o_point_vel(B, rb, vb);
  for (int i = 0; i < 3; ++i) rv[i] = va[i] - vb[i];
  vn = q_dot3(rv, n);
  float t[3] = {rv[0] - vn * n[0], rv[1] - vn * n[1], rv[2] - vn * n[2]};
  float vt2 = q_dot3(t, t);
  float lt[3] = {L->t[0], L->t[1], L->t[2]};
  if (vt2 >= 1e-18f) {
    float vt = sqrtf(vt2);
    for (int i = 0; i < 3; ++i) t[i] /= vt;
    float kt = o_eff_mass(A, ra, t) + o_eff_mass(B, rb, t);
    if (kt > 0.0f) { float jt = vt / kt; for (int i = 0; i < 3; ++i) lt[i] -= jt * t[i]; }
  }
  float lim = c->mu * ln, m2 = q_dot3(lt, lt);
  if (m2 > lim * lim) { float k_ = lim / sqrtf(m2); for (int i = 0; i < 3; ++i) lt[i] *= k_; }
  float Pt[3] = {lt[0] - L->t[0], lt[1] - L->t[1], lt[2] - L->t[2]};
  for (int i = 0; i < 3; ++i) L->t[i] = lt[i];
  o_apply_impulse(A, ra, Pt, 1.0f); o_apply_impulse(B, rb, Pt, -1.0f);
}

static inline void o_solve_one(const OContact* c, float h, const M3P2IPandaScene* sc) {
  if (c->L) { o_solve_contact3_acc(c, h, sc); return; }
  /* no slot: every visit starts from zero, which is the plain one-shot solve */
  OLam zero;
  memset(&zero, 0, sizeof(zero));
  OContact tmp = *c;
  tmp.L = &zero;
  o_solve_contact3_acc(&tmp, h, sc);
}

/* corners of box `ba` (body A) against `bb` (body B) -> contact records (normal from B to A; flip: the records are for
 * the pair (B, A) with the normal reversed, as o_corners_vs_box3). `slots` = the 8 accumulators of this corner set or
 * NULL; `sub` = 1-based sub-step index inside the current step(); `room` = records the list may still take (further
 * contacts are ignored). Returns the new count. */
static inline int o_detect(OContact* out, int cnt, OSolv3* A, const OBox3* ba, OSolv3* B, const OBox3* bb, float mu,
                           const M3P2IPandaScene* sc, int flip, OLam* slots, int sub, int room) {
  for (int i = 0; i < 8; ++i) {
    float p[3], n[3], depth;
    o_box_corner(ba, i, p);
    if (!o_point_in_box(p, bb, sc->contact_margin, n, &depth)) continue;
    if (room-- <= 0) continue;   /* the list of this cube is full: the contact is ignored */
    OContact* c = &out[cnt++];
    if (!flip) { c->A = A; c->B = B; memcpy(c->n, n, 12); }
    else { c->A = B; c->B = A; c->n[0] = -n[0]; c->n[1] = -n[1]; c->n[2] = -n[2]; }
    c->depth = depth; memcpy(c->p, p, 12); c->mu = mu; c->L = slots ? &slots[i] : NULL;
    if (!slots) continue;
    OLam* L = &slots[i];
    if (L->stamp == sub - 1 && sub > 1) {
      /* warm start: what the slot held, scaled, in the tangent plane of the new normal, inside the new cone */
      const float* m = c->n;
      float tn = q_dot3(L->t, m);
      float ln = sc->warm_start * L->n;
      float lt[3];
      for (int r = 0; r < 3; ++r) lt[r] = sc->warm_start * (L->t[r] - tn * m[r]);
      float lim = mu * ln, m2 = q_dot3(lt, lt);
      if (m2 > lim * lim) { float k_ = lim / sqrtf(m2); for (int r = 0; r < 3; ++r) lt[r] *= k_; }
      L->n = ln; L->t[0] = lt[0]; L->t[1] = lt[1]; L->t[2] = lt[2];
      L->fresh = 1;
    } else {
      L->n = 0.0f; L->t[0] = L->t[1] = L->t[2] = 0.0f; L->fresh = 0;
    }
    L->stamp = sub;
  }
  return cnt;
}

static inline void o_panda_step(OPandaEnv* e, const M3P2IPandaScene* sc, const M3P2IConfig* cfg, const float* u) {
  const int ns = cfg->substeps;
  const float h = cfg->dt / (float)ns;
  const float D = sc->drive_damping;
  const M3P2IBody* bp[2] = {&sc->cube_a, &sc->cube_b};
  float imp_table[3] = {0, 0, 0}, imp_shelf[3] = {0, 0, 0}, imp_cubeb[3] = {0, 0, 0};
  float pen_table[3] = {0, 0, 0}, pen_shelf[3] = {0, 0, 0};
  /* every step() starts cold: accumulated impulses are carried from one sub-step to the next, not across steps */
  OCache cache;
  memset(&cache, 0, sizeof(cache));
  for (int s = 0; s < ns; ++s) {
    /* 1. joint drives */
    for (int j = 0; j < 9; ++j) {
      float m = j < 7 ? (sc->joint_inertia[j] > 0.0f ? sc->joint_inertia[j] : sc->arm_inertia) : sc->finger_mass;
      float v = e->qd[j];
      float vs = (m * v + h * D * u[j]) / (m + h * D);
      float f = D * (u[j] - vs);
      if (f > sc->effort[j]) vs = v + h * sc->effort[j] / m;
      else if (f < -sc->effort[j]) vs = v - h * sc->effort[j] / m;
      vs = q_clamp(vs, -sc->qd_limit[j], sc->qd_limit[j]);
      /* joint limits: no velocity into a limit */
      if (e->q[j] <= sc->q_lower[j] && vs < 0.0f) vs = 0.0f;
      if (e->q[j] >= sc->q_upper[j] && vs > 0.0f) vs = 0.0f;
      e->qd[j] = vs;
    }
    /* link and cube boxes of this sub-step (positions are fixed until step 5) */
    OHand H;
    o_panda_hand(sc, e->q, e->qd, &H);
    OBox3 lbox[3]; /* left finger, right finger, hand */
    OSolv3 L[3];
    float slide_sign[2] = {1.0f, -1.0f};
    for (int f = 0; f < 3; ++f) {
      const float* cen = f < 2 ? sc->finger_center : sc->hand_center;
      const float* half = f < 2 ? sc->finger_half : sc->hand_half;
      float l[3] = {cen[0], cen[1], cen[2]};
      if (f == 0) { l[1] += e->q[7]; l[2] += O_FINGER_Z; }
      if (f == 1) { l[1] = -l[1] - e->q[8]; l[2] += O_FINGER_Z; } /* mirrored geometry, urdf:220 */
      for (int r = 0; r < 3; ++r)
        lbox[f].c[r] = H.p[r] + H.R[3 * r] * l[0] + H.R[3 * r + 1] * l[1] + H.R[3 * r + 2] * l[2];
      memcpy(lbox[f].R, H.R, sizeof(H.R)); memcpy(lbox[f].half, half, 12);
      memset(&L[f], 0, sizeof(OSolv3));
      L[f].v = H.v; L[f].w = H.w; memcpy(L[f].x, H.p, 12);
      if (f < 2) {
        L[f].slide = &e->qd[7 + f];
        for (int r = 0; r < 3; ++r) L[f].axis[r] = slide_sign[f] * H.R[3 * r + 1];
        L[f].ims = 1.0f / sc->finger_mass;
      }
    }
    OBox3 cbox[2];
    OSolv3 C[2];
    for (int i = 0; i < 2; ++i) {
      cbox[i] = o_cube_box3(&e->cube[i], bp[i]);
      memset(&C[i], 0, sizeof(OSolv3));
      C[i].v = e->cube[i].v; C[i].w = e->cube[i].w; C[i].im = 1.0f / bp[i]->mass; C[i].ii = 1.0f / bp[i]->inertia;
      memcpy(C[i].x, e->cube[i].p, 12);
    }
    C[1].acc = imp_cubeb;
    OSolv3 S[M3P2I_MAX_STATIC];
    OBox3 sbox[M3P2I_MAX_STATIC];
    for (int k = 0; k < sc->n_static; ++k) {
      sbox[k] = o_static_box3(&sc->statics[k]);
      memset(&S[k], 0, sizeof(OSolv3));
      memcpy(S[k].x, sbox[k].c, 12);
      if (k == sc->idx_table && sc->report_cube_contacts) S[k].acc = imp_table;
      if (k == sc->idx_shelf && sc->report_cube_contacts) S[k].acc = imp_shelf;
    }
    const int cc_near = o_boxes_near(&cbox[0], &cbox[1], sc->contact_margin);
    int lnear[3][2], first_box[2] = {-1, -1};
    for (int i = 0; i < 2; ++i) {
      for (int f = 0; f < 3; ++f) lnear[f][i] = o_boxes_near(&lbox[f], &cbox[i], sc->contact_margin);
      for (int k = 0; k < sc->n_static && first_box[i] < 0; ++k)
        if (o_boxes_near(&cbox[i], &sbox[k], sc->contact_margin)) first_box[i] = k;
    }
    /* 2. sleeping: a cube that is (almost) motionless, rests on its first near fixed box with at least three corners
     * and has no link and no other cube within the contact margin is neither moved nor solved in this sub-step */
    int asleep[2] = {0, 0};
    for (int i = 0; i < 2 && sc->sleep_lin > 0.0f; ++i) {
      const OCube* c = &e->cube[i];
      if (!(q_dot3(c->v, c->v) < sc->sleep_lin * sc->sleep_lin && q_dot3(c->w, c->w) < sc->sleep_ang * sc->sleep_ang)) continue;
      if (cc_near || lnear[0][i] || lnear[1][i] || lnear[2][i] || first_box[i] < 0) continue;
      int cnt = 0;
      for (int cI = 0; cI < 8; ++cI) {
        float p[3], n[3], depth;
        o_box_corner(&cbox[i], cI, p);
        if (o_point_in_box(p, &sbox[first_box[i]], sc->contact_margin, n, &depth) && depth > -sc->sleep_gap) ++cnt;
      }
      if (cnt < 3) continue;
      asleep[i] = 1;
      for (int r = 0; r < 3; ++r) { e->cube[i].v[r] = 0.0f; e->cube[i].w[r] = 0.0f; }
      /* the support carries the weight */
      float wgt = bp[i]->mass * sc->gravity * h;
      if (S[first_box[i]].acc) S[first_box[i]].acc[2] -= wgt;
      if (i == 1) imp_cubeb[2] += wgt;
    }
    /* 3. gravity on the cubes */
    for (int i = 0; i < 2; ++i) if (!asleep[i]) e->cube[i].v[2] -= sc->gravity * h;
    /* 4. contacts: detection once per sub-step, in solve order. Accumulators: link/cube and cube/cube contacts, and the
     * contacts of a cube with its FIRST near fixed box (its support); further fixed boxes use the plain solve. */
    OContact lk[3 * 2 * 16], ccl[16], st[2 * M3P2I_MAX_STATIC * 8];
    int n_lk = 0, n_cc = 0, n_st = 0;
    int lk_left[2] = {O_LINK_CAP, O_LINK_CAP};   /* at most O_LINK_CAP link contacts per cube, in detection order */
    for (int f = 0; f < 3; ++f)
      for (int i = 0; i < 2; ++i) {
        if (asleep[i] || !lnear[f][i]) continue;
        float mu = 0.5f * (sc->robot_mu + bp[i]->mu);
        int n0 = n_lk;
        n_lk = o_detect(lk, n_lk, &L[f], &lbox[f], &C[i], &cbox[i], mu, sc, 0, cache.lk[f][i][0], s + 1, lk_left[i]);
        lk_left[i] -= n_lk - n0; n0 = n_lk;
        n_lk = o_detect(lk, n_lk, &C[i], &cbox[i], &L[f], &lbox[f], mu, sc, 1, cache.lk[f][i][1], s + 1, lk_left[i]);
        lk_left[i] -= n_lk - n0;
      }
    if (cc_near) {
      float mu = 0.5f * (bp[0]->mu + bp[1]->mu);
      n_cc = o_detect(ccl, n_cc, &C[0], &cbox[0], &C[1], &cbox[1], mu, sc, 0, cache.cc[0], s + 1, 16);
      n_cc = o_detect(ccl, n_cc, &C[1], &cbox[1], &C[0], &cbox[0], mu, sc, 1, cache.cc[1], s + 1, 16);
    }
    for (int i = 0; i < 2; ++i)
      for (int k = 0; k < sc->n_static; ++k) {
        if (asleep[i] || !o_boxes_near(&cbox[i], &sbox[k], sc->contact_margin)) continue;
        n_st = o_detect(st, n_st, &C[i], &cbox[i], &S[k], &sbox[k], 0.5f * (bp[i]->mu + sc->statics[k].mu), sc, 0,
                        k == first_box[i] ? cache.st[i] : NULL, s + 1, 8);
      }
    const int sweeps = sc->link_sweeps > 0 ? sc->link_sweeps : 2;
    for (int p = 0; p < cfg->solver_passes; ++p) {
      /* links (a few sweeps: the finger - cube - finger chain of a grasp) and cube/cube first, the fixed boxes LAST:
       * what a kinematic link pushes into the table is pushed back out by the table in the same pass */
      for (int sw = 0; sw < sweeps; ++sw)
        for (int j = 0; j < n_lk; ++j) o_solve_one(&lk[j], h, sc);
      for (int j = 0; j < n_cc; ++j) o_solve_one(&ccl[j], h, sc);
      for (int j = 0; j < n_st; ++j) o_solve_one(&st[j], h, sc);
    }
    /* finger speed limit also holds after contacts */
    for (int j = 7; j < 9; ++j) e->qd[j] = q_clamp(e->qd[j], -sc->qd_limit[j], sc->qd_limit[j]);
    /* kinematic links against the static bodies named by the collision cost */
    for (int f = 0; f < 3; ++f) {
      if (sc->idx_table >= 0)
        o_link_vs_static(&lbox[f], &L[f], &sbox[sc->idx_table], 0.5f * (sc->robot_mu + sc->statics[sc->idx_table].mu), sc, pen_table);
      if (sc->idx_shelf >= 0)
        o_link_vs_static(&lbox[f], &L[f], &sbox[sc->idx_shelf], 0.5f * (sc->robot_mu + sc->statics[sc->idx_shelf].mu), sc, pen_shelf);
    }
    /* 5. positions */
    for (int j = 0; j < 9; ++j) {
      float qn = e->q[j] + h * e->qd[j];
      if (qn < sc->q_lower[j]) { qn = sc->q_lower[j]; e->qd[j] = 0.0f; }
      if (qn > sc->q_upper[j]) { qn = sc->q_upper[j]; e->qd[j] = 0.0f; }
      e->q[j] = qn;
    }
    for (int i = 0; i < 2; ++i) {
      if (asleep[i]) continue;
      OCube* c = &e->cube[i];
      for (int r = 0; r < 3; ++r) c->p[r] += h * c->v[r];
      /* q <- normalize(q + h/2 * (w,0) (x) q), xyzw */
      float x = c->q[0], y = c->q[1], z = c->q[2], w = c->q[3];
      float wx = c->w[0], wy = c->w[1], wz = c->w[2], hh = 0.5f * h;
      float nx = x + hh * (wx * w + wy * z - wz * y);
      float ny = y + hh * (wy * w + wz * x - wx * z);
      float nz = z + hh * (wz * w + wx * y - wy * x);
      float nw = w - hh * (wx * x + wy * y + wz * z);
      float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
      c->q[0] = nx * inv; c->q[1] = ny * inv; c->q[2] = nz * inv; c->q[3] = nw * inv;
    }
  }
  const float inv_dt = 1.0f / cfg->dt, inv_ns = 1.0f / (float)ns;
  for (int r = 0; r < 3; ++r) {
    e->f_table[r] = imp_table[r] * inv_dt + pen_table[r] * inv_ns;
    e->f_shelf[r] = imp_shelf[r] * inv_dt + pen_shelf[r] * inv_ns;
    e->f_cubeb[r] = imp_cubeb[r] * inv_dt;
  }
}

/* ------------------------------------------------------------------ costs */
/* skill_utils.py:140-180 (the reference's own quaternion -> matrix formula, row-major) */
static inline void o_ref_rotmat(const float* Q, float* R) {
  float q0 = Q[3], q1 = Q[0], q2 = Q[1], q3 = Q[2];
  R[0] = 2.0f * (q0 * q0 + q1 * q1) - 1.0f; R[1] = 2.0f * (q1 * q2 - q0 * q3); R[2] = 2.0f * (q1 * q3 + q0 * q2);
  R[3] = 2.0f * (q1 * q2 + q0 * q3); R[4] = 2.0f * (q0 * q0 + q2 * q2) - 1.0f; R[5] = 2.0f * (q2 * q3 - q0 * q1);
  R[6] = 2.0f * (q1 * q3 - q0 * q2); R[7] = 2.0f * (q2 * q3 + q0 * q1); R[8] = 2.0f * (q0 * q0 + q3 * q3) - 1.0f;
}
static inline float o_col_dot(const float* A, int ca, const float* B, int cb) {
  return A[ca] * B[cb] + A[3 + ca] * B[3 + cb] + A[6 + ca] * B[6 + cb];
}
/* min_j (1 - |<A[:,ca], B[:,j]>|) */
static inline float o_min_axis_cost(const float* A, int ca, const float* B) {
  float c0 = 1.0f - fabsf(o_col_dot(A, ca, B, 0)), c1 = 1.0f - fabsf(o_col_dot(A, ca, B, 1)),
        c2 = 1.0f - fabsf(o_col_dot(A, ca, B, 2));
  float m = c0 < c1 ? c0 : c1;
  return m < c2 ? m : c2;
}
/* skill_utils.py:224-252 */
static inline float o_ori_cube2goal(const float* cube_q, const float* goal_q) {
  float C[9], G[9];
  o_ref_rotmat(cube_q, C); o_ref_rotmat(goal_q, G);
  return o_min_axis_cost(G, 0, C) + o_min_axis_cost(G, 1, C);
}
/* index of the cube axis (column) whose world-x component is largest in magnitude; first on ties
 * (skill_utils.py:275-277) */
static inline int o_sel_axis(const float* cube_q) {
  float C[9];
  o_ref_rotmat(cube_q, C);
  int b = 0;
  if (fabsf(C[1]) > fabsf(C[b])) b = 1;
  if (fabsf(C[2]) > fabsf(C[b])) b = 2;
  return b;
}
/* skill_utils.py:256-289 */
static inline float o_ori_ee2cube(const float* ee_q, const float* cube_q, float tilt, int sel_axis) {
  float E[9], C[9];
  o_ref_rotmat(ee_q, E); o_ref_rotmat(cube_q, C);
  float cost_z;
  if (tilt == 0.0f) {
    /* the reference stacks (z, x, y): the minimum does not depend on the order */
    cost_z = o_min_axis_cost(E, 2, C);
  } else {
    cost_z = fabsf(tilt - o_col_dot(E, 2, C, sel_axis));
  }
  return cost_z + o_min_axis_cost(E, 1, C);
}

static inline void o_panda_ref(const OPandaEnv* e, OPandaRef* r, int want_axis) {
  if (!want_axis) memcpy(r->cube0_pos, e->cube[0].p, 12);
  else r->sel_axis = o_sel_axis(e->cube[0].q);
}

/* cost_functions.py:158-169 (panda_env branch) */
static inline float o_panda_motion_cost(const OPandaEnv* e) {
  float fx = e->f_table[0] + 4.0f * e->f_shelf[0] + e->f_cubeb[0];
  float fy = e->f_table[1] + 4.0f * e->f_shelf[1] + e->f_cubeb[1];
  return (fabsf(fx) + fabsf(fy)) > 0.1f ? 1000.0f : 0.0f;
}

/* cost_functions.py:19-36 (panda_env tasks). ref = values of sample 0 / sample K/2 at this step. */
static inline float o_panda_cost(const OPandaEnv* e, const M3P2IPandaScene* sc, const M3P2IConfig* cfg, int task,
                                 const float* goal, int kg, const OPandaRef* ref) {
  float links[39];
  o_panda_links(sc, e->q, e->qd, links);
  const float* lf = links; const float* rf = links + 13;
  const int second = cfg->multi_modal && kg >= cfg->num_samples_global / 2;
  switch (task) {
    case M3P2I_TASK_REACH: {
      /* cost_functions.py:91-114 */
      float g[3] = {ref->cube0_pos[0], ref->cube0_pos[1], ref->cube0_pos[2]};
      if (!second) g[2] += cfg->pre_height_diff;
      else {
        g[0] -= cfg->pre_height_diff * cfg->tilt_cos_theta;
        g[2] += cfg->pre_height_diff * sqrtf(1.0f - cfg->tilt_cos_theta * cfg->tilt_cos_theta);
      }
      float d[3];
      for (int i = 0; i < 3; ++i) d[i] = (lf[i] + rf[i]) / 2.0f - g[i];
      float reach = sqrtf(q_dot3(d, d));
      /* cost_functions.py:138-156 */
      float tilt = o_ori_ee2cube(lf + 3, e->cube[0].q, second ? cfg->tilt_cos_theta : 0.0f, ref->sel_axis);
      return 10.0f * reach + 3.0f * tilt;
    }
    case M3P2I_TASK_PICK: {
      /* cost_functions.py:116-125 */
      float d[3] = {goal[0] - e->cube[0].p[0], goal[1] - e->cube[0].p[1], goal[2] - e->cube[0].p[2]};
      float gc = sqrtf(q_dot3(d, d));
      float oc = o_ori_cube2goal(e->cube[0].q, goal + 3);
      return 10.0f * gc + 15.0f * oc + o_panda_motion_cost(e);
    }
    case M3P2I_TASK_PLACE: {
      /* cost_functions.py:127-136 */
      float d[3] = {lf[0] - rf[0], lf[1] - rf[1], lf[2] - rf[2]};
      return 2.0f * (1.0f - sqrtf(q_dot3(d, d)));
    }
    default: return 0.0f;
  }
}

/* IsaacGymWrapper tensor views for one env: dof [18], root [n_actors,13], link [3,13], contact [3,3] */
static inline void o_panda_read(const OPandaEnv* e, const M3P2IPandaScene* sc, const float* root0, float* dof,
                                float* root, float* link, float* contact) {
  if (dof) for (int j = 0; j < 9; ++j) { dof[2 * j] = e->q[j]; dof[2 * j + 1] = e->qd[j]; }
  if (root) {
    memcpy(root, root0, sizeof(float) * 13 * sc->n_actors);
    const M3P2IBody* bp[2] = {&sc->cube_a, &sc->cube_b};
    for (int i = 0; i < 2; ++i) {
      float* r = root + 13 * bp[i]->actor;
      memcpy(r, e->cube[i].p, 12); memcpy(r + 3, e->cube[i].q, 16);
      memcpy(r + 7, e->cube[i].v, 12); memcpy(r + 10, e->cube[i].w, 12);
    }
  }
  if (link) o_panda_links(sc, e->q, e->qd, link);
  if (contact) { memcpy(contact, e->f_table, 12); memcpy(contact + 3, e->f_shelf, 12); memcpy(contact + 6, e->f_cubeb, 12); }
}

#endif
