/*
 * point_env.h — CPU restatement of the planar point-robot integrator and the point_env task costs.
 * TEST INFRASTRUCTURE ONLY (see m3p2i_oracle.h).
 *
 * Integrator spec (ours; stands in for IsaacGym PhysX, isaacgym_wrapper.py:354-360, PARITY UNPINNED):
 *   per env: robot = disc of radius r on two prismatic joints driven by a velocity drive with damping D
 *   (isaacgym_wrapper.py:341-344) integrated implicitly; two movable planar boxes ("box", "dyn-obs") with
 *   Coulomb ground friction; fixed oriented boxes (walls, obs). Each of `substeps` sub-steps of h=dt/substeps:
 *     1. drive + external (suction) forces -> velocities, 2. ground friction, 3. `solver_passes` Gauss-Seidel
 *     sweeps of velocity-level contact impulses in a fixed pair order (robot-statics, robot-box, robot-dynobs,
 *     box-statics, dynobs-statics, box-dynobs) with speculative margin and Baumgarte feedback,
 *     4. explicit position update (semi-implicit Euler).
 *   net contact force on dyn-obs = sum of contact impulses on it over the step / dt.
 * Costs restate cost_functions.py:38-89,158-169 and skill_utils.py:59-94.
 */
#ifndef ORACLE_POINT_ENV_H
#define ORACLE_POINT_ENV_H

#include <math.h>
#include <string.h>
#include "../include/m3p2i_b200.h"

typedef struct {
  float x, y, th, vx, vy, w;
} OBody2;

typedef struct {
  float px, py, vx, vy; /* robot */
  OBody2 b[2];          /* 0 = box, 1 = dyn-obs */
  float f_robot[2];     /* external force applied during the NEXT step (suction), then cleared */
  float f_box[2];
  float f_dyn[2];       /* net xy contact force on dyn-obs during the last step */
} OPointEnv;

typedef struct {
  float cx, cy, hx, hy, c, s, mu;
} OBox2;

static inline float o_sign(float v) { return v < 0.0f ? -1.0f : 1.0f; }
static inline float o_clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline OBox2 o_static_box2(const M3P2IBox* b) {
  OBox2 r;
  float x = b->quat[0], y = b->quat[1], z = b->quat[2], w = b->quat[3];
  float c = 1.0f - 2.0f * (y * y + z * z);
  float s = 2.0f * (w * z + x * y);
  float n = sqrtf(c * c + s * s);
  r.cx = b->pos[0]; r.cy = b->pos[1]; r.hx = b->half[0]; r.hy = b->half[1];
  r.c = c / n; r.s = s / n; r.mu = b->mu;
  return r;
}

static inline OBox2 o_body_box2(const OBody2* b, const M3P2IBody* p) {
  OBox2 r;
  r.cx = b->x; r.cy = b->y; r.hx = p->half[0]; r.hy = p->half[1];
  r.c = cosf(b->th); r.s = sinf(b->th); r.mu = p->mu;
  return r;
}

/* A planar body as seen by the contact solver: pointers to its velocity, inverse mass / inertia, centre. */
typedef struct {
  float* vx; float* vy; float* w;
  float im, ii, x, y;
  float* acc; /* optional [2] accumulator of the impulses applied to this body */
} OSolv2;

/* One contact: normal n points from B to A, depth > 0 is penetration, (cx,cy) the contact point. */
static inline void o_solve_contact2(OSolv2* A, OSolv2* B, float nx, float ny, float depth, float cx, float cy,
                                    float mu, float h, const M3P2IPointScene* sc) {
  float rax = cx - A->x, ray = cy - A->y, rbx = cx - B->x, rby = cy - B->y;
  float wa = A->w ? *A->w : 0.0f, wb = B->w ? *B->w : 0.0f;
  float avx = A->vx ? *A->vx : 0.0f, avy = A->vy ? *A->vy : 0.0f;
  float bvx = B->vx ? *B->vx : 0.0f, bvy = B->vy ? *B->vy : 0.0f;
  float rvx = (avx - wa * ray) - (bvx - wb * rby);
  float rvy = (avy + wa * rax) - (bvy + wb * rbx);
  float vn = rvx * nx + rvy * ny;
  float ran = rax * ny - ray * nx, rbn = rbx * ny - rby * nx;
  float kn = A->im + B->im + A->ii * ran * ran + B->ii * rbn * rbn;
  float vt_target;
  if (depth > 0.0f) {
    float pen = depth - sc->slop;
    if (pen < 0.0f) pen = 0.0f;
    vt_target = sc->baumgarte * pen / h;
    if (vt_target > sc->max_corr_vel) vt_target = sc->max_corr_vel;
  } else {
    vt_target = depth / h;
  }
  if (kn <= 0.0f) return;
  float jn = (vt_target - vn) / kn;
  if (jn <= 0.0f) return;
  /* normal impulse */
  if (A->vx) { *A->vx += jn * A->im * nx; *A->vy += jn * A->im * ny; }
  if (A->w) *A->w += A->ii * ran * jn;
  if (B->vx) { *B->vx -= jn * B->im * nx; *B->vy -= jn * B->im * ny; }
  if (B->w) *B->w -= B->ii * rbn * jn;
  /* friction along t = (-ny, nx) */
  float tx = -ny, ty = nx;
  wa = A->w ? *A->w : 0.0f; wb = B->w ? *B->w : 0.0f;
  avx = A->vx ? *A->vx : 0.0f; avy = A->vy ? *A->vy : 0.0f;
  bvx = B->vx ? *B->vx : 0.0f; bvy = B->vy ? *B->vy : 0.0f;
  rvx = (avx - wa * ray) - (bvx - wb * rby);
  rvy = (avy + wa * rax) - (bvy + wb * rbx);
  float vt = rvx * tx + rvy * ty;
  float rat = rax * ty - ray * tx, rbt = rbx * ty - rby * tx;
  float kt = A->im + B->im + A->ii * rat * rat + B->ii * rbt * rbt;
  float jt = o_clamp(-vt / kt, -mu * jn, mu * jn);
  if (A->vx) { *A->vx += jt * A->im * tx; *A->vy += jt * A->im * ty; }
  if (A->w) *A->w += A->ii * rat * jt;
  if (B->vx) { *B->vx -= jt * B->im * tx; *B->vy -= jt * B->im * ty; }
  if (B->w) *B->w -= B->ii * rbt * jt;
  float ix = jn * nx + jt * tx, iy = jn * ny + jt * ty;
  if (A->acc) { A->acc[0] += ix; A->acc[1] += iy; }
  if (B->acc) { B->acc[0] -= ix; B->acc[1] -= iy; }
}

/* disc (A) against oriented box (B) */
static inline void o_disc_vs_box(OSolv2* A, float r, OSolv2* B, const OBox2* bx, float mu, float h,
                                 const M3P2IPointScene* sc) {
  float ox = A->x - bx->cx, oy = A->y - bx->cy;
  float reach = r + sc->contact_margin;
  float rad2 = bx->hx * bx->hx + bx->hy * bx->hy;
  if (ox * ox + oy * oy > (reach + sqrtf(rad2)) * (reach + sqrtf(rad2))) return; /* broad phase */
  float dx = bx->c * ox + bx->s * oy, dy = -bx->s * ox + bx->c * oy;
  float qx = o_clamp(dx, -bx->hx, bx->hx), qy = o_clamp(dy, -bx->hy, bx->hy);
  float nlx, nly, depth, plx, ply;
  if (qx == dx && qy == dy) { /* centre inside the box */
    float ex = bx->hx - fabsf(dx), ey = bx->hy - fabsf(dy);
    if (ex < ey) { nlx = o_sign(dx); nly = 0.0f; depth = r + ex; plx = nlx * bx->hx; ply = dy; }
    else { nlx = 0.0f; nly = o_sign(dy); depth = r + ey; plx = dx; ply = nly * bx->hy; }
  } else {
    float ddx = dx - qx, ddy = dy - qy;
    float dist = sqrtf(ddx * ddx + ddy * ddy);
    depth = r - dist;
    if (depth <= -sc->contact_margin) return;
    nlx = ddx / dist; nly = ddy / dist; plx = qx; ply = qy;
  }
  float nx = bx->c * nlx - bx->s * nly, ny = bx->s * nlx + bx->c * nly;
  float cx = bx->cx + bx->c * plx - bx->s * ply, cy = bx->cy + bx->s * plx + bx->c * ply;
  o_solve_contact2(A, B, nx, ny, depth, cx, cy, mu, h, sc);
}

/* corners of box A tested against the signed distance field of box B (normal from B to A) */
static inline void o_corners_vs_box(OSolv2* A, const OBox2* ba, OSolv2* B, const OBox2* bb, float mu, float h,
                                    const M3P2IPointScene* sc, int flip) {
  for (int i = 0; i < 4; ++i) {
    float lx = (i & 1) ? ba->hx : -ba->hx, ly = (i & 2) ? ba->hy : -ba->hy;
    float wx = ba->cx + ba->c * lx - ba->s * ly, wy = ba->cy + ba->s * lx + ba->c * ly;
    float ox = wx - bb->cx, oy = wy - bb->cy;
    float dx = bb->c * ox + bb->s * oy, dy = -bb->s * ox + bb->c * oy;
    float qx = fabsf(dx) - bb->hx, qy = fabsf(dy) - bb->hy;
    float qm = qx > qy ? qx : qy;
    if (qm >= sc->contact_margin) continue;
    float nlx, nly, depth;
    if (qx > qy) { nlx = o_sign(dx); nly = 0.0f; depth = -qx; }
    else { nlx = 0.0f; nly = o_sign(dy); depth = -qy; }
    float nx = bb->c * nlx - bb->s * nly, ny = bb->s * nlx + bb->c * nly;
    if (!flip) o_solve_contact2(A, B, nx, ny, depth, wx, wy, mu, h, sc);
    else o_solve_contact2(B, A, -nx, -ny, depth, wx, wy, mu, h, sc);
  }
}

static inline void o_box_vs_box(OSolv2* A, const OBox2* ba, OSolv2* B, const OBox2* bb, float mu, float h,
                                const M3P2IPointScene* sc) {
  float ox = ba->cx - bb->cx, oy = ba->cy - bb->cy;
  float ra = sqrtf(ba->hx * ba->hx + ba->hy * ba->hy), rb = sqrtf(bb->hx * bb->hx + bb->hy * bb->hy);
  float reach = ra + rb + sc->contact_margin;
  /* long thin boxes (walls) have a large bounding radius; the corner tests below do the exact work */
  if (ox * ox + oy * oy > reach * reach) return;
  o_corners_vs_box(A, ba, B, bb, mu, h, sc, 0); /* corners of A inside B: normal B->A */
  o_corners_vs_box(B, bb, A, ba, mu, h, sc, 1); /* corners of B inside A: solved as (A,B) with -n */
}

static inline void o_point_init(OPointEnv* e, const M3P2IPointScene* sc, const float* dof, const float* root) {
  memset(e, 0, sizeof(*e));
  e->px = dof[0]; e->vx = dof[1]; e->py = dof[2]; e->vy = dof[3];
  const M3P2IBody* bp[2] = {&sc->box, &sc->dyn_obs};
  for (int i = 0; i < 2; ++i) {
    const float* r = root + 13 * bp[i]->actor;
    float x = r[3], y = r[4], z = r[5], w = r[6];
    e->b[i].x = r[0]; e->b[i].y = r[1];
    e->b[i].th = atan2f(2.0f * (w * z + x * y), 1.0f - 2.0f * (y * y + z * z));
    e->b[i].vx = r[7]; e->b[i].vy = r[8]; e->b[i].w = r[12];
  }
}

static inline void o_point_step(OPointEnv* e, const M3P2IPointScene* sc, const M3P2IConfig* cfg, const float* u) {
  const int ns = cfg->substeps;
  const float h = cfg->dt / (float)ns;
  const float m = sc->robot_mass, D = sc->drive_damping, E = sc->drive_effort;
  const M3P2IBody* bp[2] = {&sc->box, &sc->dyn_obs};
  float imp_dyn[2] = {0.0f, 0.0f};
  for (int s = 0; s < ns; ++s) {
    /* 1. implicit velocity drive on the two prismatic joints */
    float* rv[2] = {&e->vx, &e->vy};
    for (int a = 0; a < 2; ++a) {
      float v = *rv[a], F = e->f_robot[a];
      float vs = (m * v + h * (D * u[a] + F)) / (m + h * D);
      float f = D * (u[a] - vs);
      if (f > E) vs = v + h * (E + F) / m;
      else if (f < -E) vs = v + h * (-E + F) / m;
      *rv[a] = vs;
    }
    /* 2. external force on the block, Coulomb ground friction on both movable boxes */
    for (int i = 0; i < 2; ++i) {
      OBody2* b = &e->b[i];
      if (i == 0) { b->vx += h * e->f_box[0] / bp[i]->mass; b->vy += h * e->f_box[1] / bp[i]->mass; }
      float mug = 0.5f * (bp[i]->mu + sc->ground_mu);
      float dv = mug * sc->gravity * h;
      float sp = sqrtf(b->vx * b->vx + b->vy * b->vy);
      if (sp <= dv) { b->vx = 0.0f; b->vy = 0.0f; }
      else { float k = 1.0f - dv / sp; b->vx *= k; b->vy *= k; }
      float dw = dv * bp[i]->mass * bp[i]->r_eff / bp[i]->inertia;
      if (fabsf(b->w) <= dw) b->w = 0.0f;
      else b->w -= o_sign(b->w) * dw;
    }
    /* 3. contacts */
    OSolv2 R = {&e->vx, &e->vy, NULL, 1.0f / (m + h * D), 0.0f, e->px, e->py, NULL};
    OSolv2 Bx = {&e->b[0].vx, &e->b[0].vy, &e->b[0].w, 1.0f / sc->box.mass, 1.0f / sc->box.inertia,
                 e->b[0].x, e->b[0].y, NULL};
    OSolv2 Dy = {&e->b[1].vx, &e->b[1].vy, &e->b[1].w, 1.0f / sc->dyn_obs.mass, 1.0f / sc->dyn_obs.inertia,
                 e->b[1].x, e->b[1].y, imp_dyn};
    OBox2 bbox = o_body_box2(&e->b[0], &sc->box), dbox = o_body_box2(&e->b[1], &sc->dyn_obs);
    for (int p = 0; p < cfg->solver_passes; ++p) {
      for (int i = 0; i < sc->n_static; ++i) {
        OBox2 sb = o_static_box2(&sc->statics[i]);
        OSolv2 S = {NULL, NULL, NULL, 0.0f, 0.0f, sb.cx, sb.cy, NULL};
        o_disc_vs_box(&R, sc->robot_radius, &S, &sb, 0.5f * (sc->robot_mu + sb.mu), h, sc);
      }
      o_disc_vs_box(&R, sc->robot_radius, &Bx, &bbox, 0.5f * (sc->robot_mu + bbox.mu), h, sc);
      o_disc_vs_box(&R, sc->robot_radius, &Dy, &dbox, 0.5f * (sc->robot_mu + dbox.mu), h, sc);
      for (int i = 0; i < sc->n_static; ++i) {
        OBox2 sb = o_static_box2(&sc->statics[i]);
        OSolv2 S = {NULL, NULL, NULL, 0.0f, 0.0f, sb.cx, sb.cy, NULL};
        o_box_vs_box(&Bx, &bbox, &S, &sb, 0.5f * (bbox.mu + sb.mu), h, sc);
      }
      for (int i = 0; i < sc->n_static; ++i) {
        OBox2 sb = o_static_box2(&sc->statics[i]);
        OSolv2 S = {NULL, NULL, NULL, 0.0f, 0.0f, sb.cx, sb.cy, NULL};
        o_box_vs_box(&Dy, &dbox, &S, &sb, 0.5f * (dbox.mu + sb.mu), h, sc);
      }
      o_box_vs_box(&Bx, &bbox, &Dy, &dbox, 0.5f * (bbox.mu + dbox.mu), h, sc);
    }
    /* 4. positions */
    e->px += h * e->vx; e->py += h * e->vy;
    for (int i = 0; i < 2; ++i) {
      e->b[i].x += h * e->b[i].vx; e->b[i].y += h * e->b[i].vy; e->b[i].th += h * e->b[i].w;
    }
  }
  e->f_dyn[0] = imp_dyn[0] / cfg->dt; e->f_dyn[1] = imp_dyn[1] / cfg->dt;
  e->f_robot[0] = e->f_robot[1] = e->f_box[0] = e->f_box[1] = 0.0f; /* forces last one step */
}

/* cost_functions.py:158-169 (point_env branch) */
static inline float o_point_motion_cost(const OPointEnv* e) {
  float c = fabsf(e->f_dyn[0]) + fabsf(e->f_dyn[1]);
  return c > 0.1f ? 1000.0f : 0.0f;
}

/* cost_functions.py:41-50 */
static inline void o_point_dist(const OPointEnv* e, const float* goal, float* dist_cost, float* cos_theta) {
  float rbx = e->px - e->b[0].x, rby = e->py - e->b[0].y;
  float bgx = goal[0] - e->b[0].x, bgy = goal[1] - e->b[0].y;
  float d1 = sqrtf(rbx * rbx + rby * rby), d2 = sqrtf(bgx * bgx + bgy * bgy);
  *dist_cost = d1 + d2 * 10.0f;
  *cos_theta = (rbx * bgx + rby * bgy) / (d1 * d2);
}

/* cost_functions.py:52-60 */
static inline float o_point_push_cost(const OPointEnv* e, const float* goal) {
  float dc, ct;
  o_point_dist(e, goal, &dc, &ct);
  float align = ct > 0.0f ? ct : 0.0f;
  return 3.0f * dc + 1.0f * align;
}

/* cost_functions.py:62-89 + skill_utils.py:59-94; writes the suction forces that act during the next step */
static inline float o_point_pull_cost(OPointEnv* e, const float* goal, const M3P2IConfig* cfg, int kg) {
  float pdx = e->b[0].x - e->px, pdy = e->b[0].y - e->py;
  float rbd = sqrtf(pdx * pdx + pdy * pdy);
  int towards = (e->vx * pdx + e->vy * pdy) > 0.0f;
  /* calculate_suction */
  float mag = 1.0f / rbd;
  float ux = pdx * mag, uy = pdy * mag;
  float thr = cfg->num_samples_global == 1 ? 1.5f : 1.8f;
  float fbx = 0.0f, fby = 0.0f, frx = 0.0f, fry = 0.0f;
  if (mag > thr) {
    fbx = o_clamp(-cfg->kp_suction * ux, -500.0f, 500.0f); fby = o_clamp(-cfg->kp_suction * uy, -500.0f, 500.0f);
    frx = o_clamp(cfg->kp_suction * ux, -500.0f, 500.0f); fry = o_clamp(cfg->kp_suction * uy, -500.0f, 500.0f);
  }
  if (towards || (cfg->multi_modal && kg < cfg->num_samples_global / 2)) { fbx = fby = frx = fry = 0.0f; }
  e->f_box[0] = fbx; e->f_box[1] = fby; e->f_robot[0] = frx; e->f_robot[1] = fry;
  float dc, ct;
  o_point_dist(e, goal, &dc, &ct);
  float align = ct < 0.0f ? -ct : 0.0f;
  float vel_cost = (towards && rbd <= 0.5f) ? 0.6f : 0.0f;
  return 3.0f * dc + 3.0f * vel_cost + 7.0f * align;
}

/* cost_functions.py:19-36 (point_env tasks) */
static inline float o_point_cost(OPointEnv* e, const M3P2IConfig* cfg, int task, const float* goal, int kg) {
  switch (task) {
    case M3P2I_TASK_NAVIGATION: {
      float dx = e->px - goal[0], dy = e->py - goal[1];
      return sqrtf(dx * dx + dy * dy) + o_point_motion_cost(e);
    }
    case M3P2I_TASK_PUSH: return o_point_push_cost(e, goal);
    case M3P2I_TASK_PULL: return o_point_pull_cost(e, goal, cfg, kg);
    case M3P2I_TASK_PUSH_PULL: {
      float push = o_point_push_cost(e, goal);
      float pull = o_point_pull_cost(e, goal, cfg, kg);
      return kg < cfg->num_samples_global / 2 ? push : pull;
    }
    default: return 0.0f;
  }
}

#endif
