// point_env.cuh — planar point-robot environment: disc robot on two velocity-driven prismatic joints, two movable
// boxes (pushed block, dynamic obstacle) with ground friction, fixed oriented boxes; per-step task costs.
// Replaces REACTIVE_TAMP.dynamics -> IsaacGymWrapper.step (reactive_tamp.py:63-70, isaacgym_wrapper.py:354-360) and
// Objective.compute_cost for navigation / push / pull / push_pull (cost_functions.py:19-89,158-169; suction:
// skill_utils.py:59-94) for the point_env scene. All state of one env lives in registers of one thread.
#pragma once
#include "common.cuh"

namespace m3 {

struct Body2 {
  float x, y, th, vx, vy, w;
};

struct PointEnv {
  float px, py, vx, vy;     // robot
  Body2 box, dyn;           // pushed block, dynamic obstacle
  float frx, fry, fbx, fby; // suction pair acting during the next step
  float fdx, fdy;           // net xy contact force on the dynamic obstacle during the last step

  // field-major load/store: field f of env k at p[f * stride + k]
  DEV void load(const float* p, int stride, int k) {
    const float* q = p + k;
    px = q[0 * stride]; vx = q[1 * stride]; py = q[2 * stride]; vy = q[3 * stride];
    box.x = q[4 * stride]; box.y = q[5 * stride]; box.th = q[6 * stride];
    box.vx = q[7 * stride]; box.vy = q[8 * stride]; box.w = q[9 * stride];
    dyn.x = q[10 * stride]; dyn.y = q[11 * stride]; dyn.th = q[12 * stride];
    dyn.vx = q[13 * stride]; dyn.vy = q[14 * stride]; dyn.w = q[15 * stride];
    frx = q[16 * stride]; fry = q[17 * stride]; fbx = q[18 * stride]; fby = q[19 * stride];
    fdx = q[20 * stride]; fdy = q[21 * stride];
  }
  DEV void store(float* p, int stride, int k) const {
    float* q = p + k;
    q[0 * stride] = px; q[1 * stride] = vx; q[2 * stride] = py; q[3 * stride] = vy;
    q[4 * stride] = box.x; q[5 * stride] = box.y; q[6 * stride] = box.th;
    q[7 * stride] = box.vx; q[8 * stride] = box.vy; q[9 * stride] = box.w;
    q[10 * stride] = dyn.x; q[11 * stride] = dyn.y; q[12 * stride] = dyn.th;
    q[13 * stride] = dyn.vx; q[14 * stride] = dyn.vy; q[15 * stride] = dyn.w;
    q[16 * stride] = frx; q[17 * stride] = fry; q[18 * stride] = fbx; q[19 * stride] = fby;
    q[20 * stride] = fdx; q[21 * stride] = fdy;
  }
  DEV float4 state_row() const { return make_float4(px, vx, py, vy); }
};

// velocity-level view of a planar body for the contact solver; fixed bodies have im = ii = 0
struct Dyn2 {
  float vx, vy, w, im, ii, x, y;
};

struct OBox2 {
  float cx, cy, hx, hy, c, s;
};

DEV OBox2 obox_of(const Body2& b, float hx, float hy) {
  OBox2 r;
  r.cx = b.x; r.cy = b.y; r.hx = hx; r.hy = hy;
  sincosf(b.th, &r.s, &r.c);
  return r;
}
DEV OBox2 obox_of(const Static2& s) {
  OBox2 r = {s.cx, s.cy, s.hx, s.hy, s.c, s.s};
  return r;
}

// One contact between A and B: unit normal n from B to A, depth > 0 = penetration, contact point c.
// Sequential impulse without accumulation: normal impulse towards the target separating speed, then Coulomb
// friction clamped by mu * jn. Returns the impulse applied to A (B receives the opposite).
DEV void solve_contact2(Dyn2& A, Dyn2& B, float nx, float ny, float depth, float cx, float cy, float mu, float h,
                        const PointParams& P, float& ix, float& iy) {
  ix = 0.0f; iy = 0.0f;
  const float rax = cx - A.x, ray = cy - A.y, rbx = cx - B.x, rby = cy - B.y;
  float rvx = (A.vx - A.w * ray) - (B.vx - B.w * rby);
  float rvy = (A.vy + A.w * rax) - (B.vy + B.w * rbx);
  const float vn = rvx * nx + rvy * ny;
  const float ran = rax * ny - ray * nx, rbn = rbx * ny - rby * nx;
  const float kn = A.im + B.im + A.ii * ran * ran + B.ii * rbn * rbn;
  float target;
  if (depth > 0.0f) {
    float pen = fmaxf(depth - P.slop, 0.0f);
    target = fminf(P.baumgarte * pen / h, P.max_corr_vel);
  } else {
    target = depth / h;
  }
  if (kn <= 0.0f) return;
  const float jn = (target - vn) / kn;
  if (jn <= 0.0f) return;
  A.vx += jn * A.im * nx; A.vy += jn * A.im * ny; A.w += A.ii * ran * jn;
  B.vx -= jn * B.im * nx; B.vy -= jn * B.im * ny; B.w -= B.ii * rbn * jn;
  const float tx = -ny, ty = nx;
  rvx = (A.vx - A.w * ray) - (B.vx - B.w * rby);
  rvy = (A.vy + A.w * rax) - (B.vy + B.w * rbx);
  const float vt = rvx * tx + rvy * ty;
  const float rat = rax * ty - ray * tx, rbt = rbx * ty - rby * tx;
  const float kt = A.im + B.im + A.ii * rat * rat + B.ii * rbt * rbt;
  const float jt = clampf(-vt / kt, -mu * jn, mu * jn);
  A.vx += jt * A.im * tx; A.vy += jt * A.im * ty; A.w += A.ii * rat * jt;
  B.vx -= jt * B.im * tx; B.vy -= jt * B.im * ty; B.w -= B.ii * rbt * jt;
  ix = jn * nx + jt * tx; iy = jn * ny + jt * ty;
}

// the broad-phase tests of disc_vs_box / box_vs_box as predicates (same expressions)
DEV bool disc_near(float px, float py, float reach, const OBox2& bx) {
  const float ox = px - bx.cx, oy = py - bx.cy;
  const float dx = bx.c * ox + bx.s * oy, dy = -bx.s * ox + bx.c * oy;
  return !(fabsf(dx) - bx.hx > reach || fabsf(dy) - bx.hy > reach);
}
DEV bool box_near(const OBox2& ba, const OBox2& bb, float margin) {
  const float ox = ba.cx - bb.cx, oy = ba.cy - bb.cy;
  const float dx = bb.c * ox + bb.s * oy, dy = -bb.s * ox + bb.c * oy;
  const float reach = sqrtf(ba.hx * ba.hx + ba.hy * ba.hy) + margin;
  const float ex = fmaxf(fabsf(dx) - bb.hx, 0.0f), ey = fmaxf(fabsf(dy) - bb.hy, 0.0f);
  return !(ex * ex + ey * ey > reach * reach);
}

// robot disc (A) against an oriented box (B); accB accumulates the impulse received by B
DEV void disc_vs_box(Dyn2& A, float r, Dyn2& B, const OBox2& bx, float mu, float h, const PointParams& P, float& accBx,
                     float& accBy) {
  const float ox = A.x - bx.cx, oy = A.y - bx.cy;
  const float dx = bx.c * ox + bx.s * oy, dy = -bx.s * ox + bx.c * oy;
  // broad phase in the box frame (exact for long thin walls, whose bounding circle would always pass)
  const float reach = r + P.contact_margin;
  if (fabsf(dx) - bx.hx > reach || fabsf(dy) - bx.hy > reach) return;
  const float qx = clampf(dx, -bx.hx, bx.hx), qy = clampf(dy, -bx.hy, bx.hy);
  float nlx, nly, depth, plx, ply;
  if (qx == dx && qy == dy) {  // centre inside the box: push out along the least-penetration face
    const float ex = bx.hx - fabsf(dx), ey = bx.hy - fabsf(dy);
    if (ex < ey) { nlx = signf(dx); nly = 0.0f; depth = r + ex; plx = nlx * bx.hx; ply = dy; }
    else { nlx = 0.0f; nly = signf(dy); depth = r + ey; plx = dx; ply = nly * bx.hy; }
  } else {
    const float ddx = dx - qx, ddy = dy - qy;
    const float dist = sqrtf(ddx * ddx + ddy * ddy);
    depth = r - dist;
    if (depth <= -P.contact_margin) return;
    nlx = ddx / dist; nly = ddy / dist; plx = qx; ply = qy;
  }
  const float nx = bx.c * nlx - bx.s * nly, ny = bx.s * nlx + bx.c * nly;
  const float cx = bx.cx + bx.c * plx - bx.s * ply, cy = bx.cy + bx.s * plx + bx.c * ply;
  float ix, iy;
  solve_contact2(A, B, nx, ny, depth, cx, cy, mu, h, P, ix, iy);
  accBx -= ix; accBy -= iy;
}

// corners of box `ba` (body A) against the signed distance field of `bb` (body B)
template <bool FLIP>
DEV void corners_vs_box(Dyn2& A, const OBox2& ba, Dyn2& B, const OBox2& bb, float mu, float h, const PointParams& P,
                        float& accAx, float& accAy, float& accBx, float& accBy) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lx = (i & 1) ? ba.hx : -ba.hx, ly = (i & 2) ? ba.hy : -ba.hy;
    const float wx = ba.cx + ba.c * lx - ba.s * ly, wy = ba.cy + ba.s * lx + ba.c * ly;
    const float ox = wx - bb.cx, oy = wy - bb.cy;
    const float dx = bb.c * ox + bb.s * oy, dy = -bb.s * ox + bb.c * oy;
    const float qx = fabsf(dx) - bb.hx, qy = fabsf(dy) - bb.hy;
    if (fmaxf(qx, qy) >= P.contact_margin) continue;
    float nlx, nly, depth;
    if (qx > qy) { nlx = signf(dx); nly = 0.0f; depth = -qx; }
    else { nlx = 0.0f; nly = signf(dy); depth = -qy; }
    const float nx = bb.c * nlx - bb.s * nly, ny = bb.s * nlx + bb.c * nly;
    float ix, iy;
    if (!FLIP) {
      solve_contact2(A, B, nx, ny, depth, wx, wy, mu, h, P, ix, iy);
      accAx += ix; accAy += iy; accBx -= ix; accBy -= iy;
    } else {  // the corner belongs to the second body of the pair: solve as (B, A) with the opposite normal
      solve_contact2(B, A, -nx, -ny, depth, wx, wy, mu, h, P, ix, iy);
      accBx += ix; accBy += iy; accAx -= ix; accAy -= iy;
    }
  }
}

// (A, ba) vs (B, bb): corners of A in B, then corners of B in A. acc* accumulate the impulses received.
DEV void box_vs_box(Dyn2& A, const OBox2& ba, Dyn2& B, const OBox2& bb, float mu, float h, const PointParams& P,
                    float& accAx, float& accAy, float& accBx, float& accBy) {
  // broad phase: bounding circle of ba against the box bb in bb's frame. Any contact of either direction (a corner
  // within the margin of the other box) implies dist(centre of ba, bb) <= radius(ba) + margin, so this never drops a
  // pair the corner tests would have found, and unlike a circle-circle test it rejects the 8 m long walls.
  const float ox = ba.cx - bb.cx, oy = ba.cy - bb.cy;
  const float dx = bb.c * ox + bb.s * oy, dy = -bb.s * ox + bb.c * oy;
  const float reach = sqrtf(ba.hx * ba.hx + ba.hy * ba.hy) + P.contact_margin;
  const float ex = fmaxf(fabsf(dx) - bb.hx, 0.0f), ey = fmaxf(fabsf(dy) - bb.hy, 0.0f);
  if (ex * ex + ey * ey > reach * reach) return;
  corners_vs_box<false>(A, ba, B, bb, mu, h, P, accAx, accAy, accBx, accBy);
  corners_vs_box<true>(B, bb, A, ba, mu, h, P, accBx, accBy, accAx, accAy);
}

// one sim step of dt = substeps * h (semi-implicit Euler)
DEV void point_step(PointEnv& e, const PointParams& P, const float* u, float dt, int substeps, int passes) {
  const float h = dt / (float)substeps;
  const float m = P.robot_mass, D = P.drive_damping, E = P.drive_effort;
  float imp_dx = 0.0f, imp_dy = 0.0f, sink_x = 0.0f, sink_y = 0.0f;
  for (int s = 0; s < substeps; ++s) {
    // 1. implicit velocity drive of the two prismatic joints, force-limited
    {
      float vs = (m * e.vx + h * (D * u[0] + e.frx)) / (m + h * D);
      float f = D * (u[0] - vs);
      if (f > E) vs = e.vx + h * (E + e.frx) / m;
      else if (f < -E) vs = e.vx + h * (-E + e.frx) / m;
      e.vx = vs;
      vs = (m * e.vy + h * (D * u[1] + e.fry)) / (m + h * D);
      f = D * (u[1] - vs);
      if (f > E) vs = e.vy + h * (E + e.fry) / m;
      else if (f < -E) vs = e.vy + h * (-E + e.fry) / m;
      e.vy = vs;
    }
    // 2. suction on the block, Coulomb ground friction (linear + torsional) on both movable boxes
    e.box.vx += h * e.fbx / P.box_mass; e.box.vy += h * e.fby / P.box_mass;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      Body2& b = i == 0 ? e.box : e.dyn;
      const float bmu = i == 0 ? P.box_mu : P.dyn_mu, bm = i == 0 ? P.box_mass : P.dyn_mass;
      const float bi = i == 0 ? P.box_inertia : P.dyn_inertia, br = i == 0 ? P.box_reff : P.dyn_reff;
      const float dv = 0.5f * (bmu + P.ground_mu) * P.gravity * h;
      const float sp = sqrtf(b.vx * b.vx + b.vy * b.vy);
      if (sp <= dv) { b.vx = 0.0f; b.vy = 0.0f; }
      else { const float k = 1.0f - dv / sp; b.vx *= k; b.vy *= k; }
      const float dw = dv * bm * br / bi;
      if (fabsf(b.w) <= dw) b.w = 0.0f;
      else b.w -= signf(b.w) * dw;
    }
    // 3. contacts: fixed pair order, `passes` Gauss-Seidel sweeps
    Dyn2 R = {e.vx, e.vy, 0.0f, 1.0f / (m + h * D), 0.0f, e.px, e.py};
    Dyn2 Bx = {e.box.vx, e.box.vy, e.box.w, 1.0f / P.box_mass, 1.0f / P.box_inertia, e.box.x, e.box.y};
    Dyn2 Dy = {e.dyn.vx, e.dyn.vy, e.dyn.w, 1.0f / P.dyn_mass, 1.0f / P.dyn_inertia, e.dyn.x, e.dyn.y};
    const OBox2 bbox = obox_of(e.box, P.box_hx, P.box_hy), dbox = obox_of(e.dyn, P.dyn_hx, P.dyn_hy);
    // Positions are fixed during a sub-step, so which fixed boxes can touch the robot / the block / the obstacle is
    // decided once (the broad-phase tests of disc_vs_box and box_vs_box, one bit per fixed box) and every pass visits
    // only those: same contacts, same order, same arithmetic.
    unsigned near_r = 0u, near_b = 0u, near_d = 0u;
    for (int i = 0; i < P.n_static; ++i) {
      const OBox2 sb = obox_of(P.st[i]);
      if (disc_near(e.px, e.py, P.robot_radius + P.contact_margin, sb)) near_r |= 1u << i;
      if (box_near(bbox, sb, P.contact_margin)) near_b |= 1u << i;
      if (box_near(dbox, sb, P.contact_margin)) near_d |= 1u << i;
    }
    for (int p = 0; p < passes; ++p) {
      for (unsigned m_ = near_r; m_; m_ &= m_ - 1u) {
        const int i = __ffs(m_) - 1;
        Dyn2 S = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, P.st[i].cx, P.st[i].cy};
        disc_vs_box(R, P.robot_radius, S, obox_of(P.st[i]), 0.5f * (P.robot_mu + P.st[i].mu), h, P, sink_x, sink_y);
      }
      disc_vs_box(R, P.robot_radius, Bx, bbox, 0.5f * (P.robot_mu + P.box_mu), h, P, sink_x, sink_y);
      disc_vs_box(R, P.robot_radius, Dy, dbox, 0.5f * (P.robot_mu + P.dyn_mu), h, P, imp_dx, imp_dy);
      for (unsigned m_ = near_b; m_; m_ &= m_ - 1u) {
        const int i = __ffs(m_) - 1;
        Dyn2 S = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, P.st[i].cx, P.st[i].cy};
        box_vs_box(Bx, bbox, S, obox_of(P.st[i]), 0.5f * (P.box_mu + P.st[i].mu), h, P, sink_x, sink_y, sink_x, sink_y);
      }
      for (unsigned m_ = near_d; m_; m_ &= m_ - 1u) {
        const int i = __ffs(m_) - 1;
        Dyn2 S = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, P.st[i].cx, P.st[i].cy};
        box_vs_box(Dy, dbox, S, obox_of(P.st[i]), 0.5f * (P.dyn_mu + P.st[i].mu), h, P, imp_dx, imp_dy, sink_x, sink_y);
      }
      box_vs_box(Bx, bbox, Dy, dbox, 0.5f * (P.box_mu + P.dyn_mu), h, P, sink_x, sink_y, imp_dx, imp_dy);
    }
    e.vx = R.vx; e.vy = R.vy;
    e.box.vx = Bx.vx; e.box.vy = Bx.vy; e.box.w = Bx.w;
    e.dyn.vx = Dy.vx; e.dyn.vy = Dy.vy; e.dyn.w = Dy.w;
    // 4. positions
    e.px += h * e.vx; e.py += h * e.vy;
    e.box.x += h * e.box.vx; e.box.y += h * e.box.vy; e.box.th += h * e.box.w;
    e.dyn.x += h * e.dyn.vx; e.dyn.y += h * e.dyn.vy; e.dyn.th += h * e.dyn.w;
  }
  e.fdx = imp_dx / dt; e.fdy = imp_dy / dt;
  e.frx = e.fry = e.fbx = e.fby = 0.0f;  // applied forces last one step
}

// ------------------------------------------------------------------ costs
DEV void point_dist(const PointEnv& e, const float* goal, float& dist_cost, float& cos_theta) {
  const float rbx = e.px - e.box.x, rby = e.py - e.box.y;
  const float bgx = goal[0] - e.box.x, bgy = goal[1] - e.box.y;
  const float d1 = sqrtf(rbx * rbx + rby * rby), d2 = sqrtf(bgx * bgx + bgy * bgy);
  dist_cost = d1 + d2 * 10.0f;
  cos_theta = (rbx * bgx + rby * bgy) / (d1 * d2);
}

DEV float point_push_cost(const PointEnv& e, const float* goal) {
  float dc, ct;
  point_dist(e, goal, dc, ct);
  return 3.0f * dc + 1.0f * fmaxf(ct, 0.0f);
}

// pull cost; arms the suction force pair for the next step (cost_functions.py:62-89, skill_utils.py:59-94)
DEV float point_pull_cost(PointEnv& e, const float* goal, const RolloutCfg& c, int kg) {
  const float pdx = e.box.x - e.px, pdy = e.box.y - e.py;
  const float rbd = sqrtf(pdx * pdx + pdy * pdy);
  const bool towards = (e.vx * pdx + e.vy * pdy) > 0.0f;
  const float mag = 1.0f / rbd;
  const float ux = pdx * mag, uy = pdy * mag;
  const float thr = c.Kg == 1 ? 1.5f : 1.8f;
  float fbx = 0.0f, fby = 0.0f, frx = 0.0f, fry = 0.0f;
  if (mag > thr) {
    fbx = clampf(-c.kp_suction * ux, -500.0f, 500.0f); fby = clampf(-c.kp_suction * uy, -500.0f, 500.0f);
    frx = clampf(c.kp_suction * ux, -500.0f, 500.0f); fry = clampf(c.kp_suction * uy, -500.0f, 500.0f);
  }
  if (towards || (c.multi_modal && kg < c.Kg / 2)) { fbx = fby = frx = fry = 0.0f; }
  e.fbx = fbx; e.fby = fby; e.frx = frx; e.fry = fry;
  float dc, ct;
  point_dist(e, goal, dc, ct);
  const float align = ct < 0.0f ? -ct : 0.0f;
  const float vel_cost = (towards && rbd <= 0.5f) ? 0.6f : 0.0f;
  return 3.0f * dc + 3.0f * vel_cost + 7.0f * align;
}

DEV float point_cost(PointEnv& e, const RolloutCfg& c, int kg) {
  switch (c.task) {
    case M3P2I_TASK_NAVIGATION: {
      const float dx = e.px - c.goal[0], dy = e.py - c.goal[1];
      const float coll = (fabsf(e.fdx) + fabsf(e.fdy)) > 0.1f ? 1000.0f : 0.0f;
      return sqrtf(dx * dx + dy * dy) + coll;
    }
    case M3P2I_TASK_PUSH: return point_push_cost(e, c.goal);
    case M3P2I_TASK_PULL: return point_pull_cost(e, c.goal, c, kg);
    case M3P2I_TASK_PUSH_PULL: {
      // both halves evaluate the pull cost (it arms the suction state); the first half reports the push cost
      const float push = point_push_cost(e, c.goal);
      const float pull = point_pull_cost(e, c.goal, c, kg);
      return kg < c.Kg / 2 ? push : pull;
    }
    default: return 0.0f;
  }
}

}  // namespace m3
