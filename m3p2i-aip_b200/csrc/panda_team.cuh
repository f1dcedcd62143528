// panda_team.cuh — lane-cooperative Panda step: a TEAM of 16 lanes advances ONE sample.
//
// Why: at the sizes the planner runs at (K = 4096 samples) a thread-per-sample rollout is one warp per SM walking a
// ~23 k-instruction serial chain per step, with 3/4 of the SM sub-partitions idle. The work inside a sample is mostly
// contact detection (independent per box corner) and Gauss-Seidel impulse solves (serial per body, but the two cubes
// are independent of each other). A team splits it like this:
//     lane bit 3  (g) : which cube the lane works for (0 = cubeA, 1 = cubeB); each 8-lane group keeps a replica of
//                       "its" cube's state and applies that cube's impulses in lock-step
//     lane bits 0-2 (c): which box corner the lane tests
// Joint state, forward kinematics and the link boxes are replicated in all 16 lanes. Contact detection runs on all
// corners at once; the solves are broadcast (warp shuffles) and applied in exactly the pair / corner order of the
// thread-per-sample code (panda_env.cuh), so both produce the same trajectory up to fp32 summation order of the
// reported contact forces. All branches that contain shuffles are warp-uniform (decided by __ballot_sync/__any_sync).
#pragma once
#include "panda_env.cuh"

namespace m3 {

constexpr int kTeam = 16;
constexpr unsigned kFull = 0xffffffffu;

DEV V3 shfl3(V3 a, int src) {
  return mk(__shfl_sync(kFull, a.x, src), __shfl_sync(kFull, a.y, src), __shfl_sync(kFull, a.z, src));
}

struct TeamLane {
  int lane, g, c;     // lane in warp, cube group, corner
  int team_base;      // first lane of this team in the warp
  int group_base;     // first lane of this 8-lane group in the warp
};

DEV TeamLane team_lane() {
  TeamLane t;
  t.lane = threadIdx.x & 31;
  t.g = (t.lane >> 3) & 1;
  t.c = t.lane & 7;
  t.team_base = t.lane & 16;
  t.group_base = t.lane & 24;
  return t;
}

// per-lane state of one sample: joints replicated, ONE cube (cube t.g), contact forces replicated after each step
struct TeamEnv {
  float q[9], qd[9];
  Cube cu;
  V3 f_table, f_shelf, f_cubeb;

  DEV void load(const float* p, int stride, int k, int g) {
    const float* s = p + k;
#pragma unroll
    for (int j = 0; j < 9; ++j) { q[j] = s[(2 * j) * stride]; qd[j] = s[(2 * j + 1) * stride]; }
    int f = 18 + 13 * g;
    cu.p = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
    cu.qx = s[f * stride]; cu.qy = s[(f + 1) * stride]; cu.qz = s[(f + 2) * stride]; cu.qw = s[(f + 3) * stride]; f += 4;
    cu.v = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
    cu.w = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]);
    f = 44;
    f_table = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
    f_shelf = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]); f += 3;
    f_cubeb = mk(s[f * stride], s[(f + 1) * stride], s[(f + 2) * stride]);
  }
  // lanes c == 0 of each group write their cube; the team's first lane also writes joints and forces
  DEV void store(float* p, int stride, int k, const TeamLane& t) const {
    if (t.c != 0) return;
    float* s = p + k;
    int f = 18 + 13 * t.g;
    s[f * stride] = cu.p.x; s[(f + 1) * stride] = cu.p.y; s[(f + 2) * stride] = cu.p.z; f += 3;
    s[f * stride] = cu.qx; s[(f + 1) * stride] = cu.qy; s[(f + 2) * stride] = cu.qz; s[(f + 3) * stride] = cu.qw; f += 4;
    s[f * stride] = cu.v.x; s[(f + 1) * stride] = cu.v.y; s[(f + 2) * stride] = cu.v.z; f += 3;
    s[f * stride] = cu.w.x; s[(f + 1) * stride] = cu.w.y; s[(f + 2) * stride] = cu.w.z;
    if (t.g != 0) return;
#pragma unroll
    for (int j = 0; j < 9; ++j) { s[(2 * j) * stride] = q[j]; s[(2 * j + 1) * stride] = qd[j]; }
    f = 44;
    s[f * stride] = f_table.x; s[(f + 1) * stride] = f_table.y; s[(f + 2) * stride] = f_table.z; f += 3;
    s[f * stride] = f_shelf.x; s[(f + 1) * stride] = f_shelf.y; s[(f + 2) * stride] = f_shelf.z; f += 3;
    s[f * stride] = f_cubeb.x; s[(f + 1) * stride] = f_cubeb.y; s[(f + 2) * stride] = f_cubeb.z;
  }
  DEV float4 state_row() const { return make_float4(q[0], qd[0], q[1], qd[1]); }
};

// The generic two-body solve is only reached when a cube touches the other cube or the gripper; keeping one
// out-of-line copy keeps the hot loop small enough for the instruction cache.
__device__ __noinline__ V3 solve_contact3_call(Dyn3& A, Dyn3& B, V3 n, float depth, V3 c, float mu, float h,
                                               const PandaParams& P) {
  return solve_contact3(A, B, n, depth, c, mu, h, P);
}

DEV Dyn3 dyn_cube(V3 v, V3 w, V3 x, float im, float ii) {
  Dyn3 d;
  d.v = v; d.w = w; d.x = x; d.im = im; d.ii = ii; d.axis = mk(0, 0, 0); d.slide = 0.0f; d.ims = 0.0f;
  return d;
}

// Serial application of the contacts found by the lanes [src0, src0+8): for corner j = 0..7 in order, the lane that
// found a hit broadcasts (n, depth, p) and every lane for which `mine` holds applies solve(A, B, sign*n, ...).
// Returns the sum of the impulses applied to A. The loop bounds and the shuffles are warp-uniform.
template <typename Solve>
DEV V3 apply_hits(bool hit, V3 n, float depth, V3 p, int lane, int src0, bool mine, Solve&& solve) {
  const unsigned hb = __ballot_sync(kFull, hit);
  V3 acc = mk(0, 0, 0);
  if (!hb) return acc;
  // corners that hit in ANY 8-lane group of the warp, visited in ascending corner order
  unsigned todo = (hb | (hb >> 8) | (hb >> 16) | (hb >> 24)) & 0xffu;
  while (todo) {
    const int j = __ffs(todo) - 1;
    todo &= todo - 1;
    const int src = src0 + j;
    const bool act = mine && ((hb >> src) & 1u);
    const V3 nj = shfl3(n, src), pj = shfl3(p, src);
    const float dj = __shfl_sync(kFull, depth, src);
    if (act) acc = acc + solve(nj, dj, pj);
  }
  return acc;
}

DEV void team_panda_step(TeamEnv& e, const PandaParams& P, const float* u, float dt, int substeps, int passes,
                         const TeamLane& t) {
  const float h = dt / (float)substeps;
  const float D = P.drive_damping;
  const int g = t.g;
  const float im = 1.0f / P.cube_mass[g], ii = 1.0f / P.cube_inertia[g], mu_c = P.cube_mu[g];
  const float imo = 1.0f / P.cube_mass[g ^ 1], iio = 1.0f / P.cube_inertia[g ^ 1];
  const V3 half_own = mk(P.cube_half[g][0], P.cube_half[g][1], P.cube_half[g][2]);
  const V3 half_oth = mk(P.cube_half[g ^ 1][0], P.cube_half[g ^ 1][1], P.cube_half[g ^ 1][2]);
  const float rad_own = sqrtf(dot(half_own, half_own)), rad_oth = sqrtf(dot(half_oth, half_oth));
  // group-partial impulse sums (identical in the 8 lanes of a group), lane-partial penalty sums
  V3 imp_table = mk(0, 0, 0), imp_shelf = mk(0, 0, 0), imp_cubeb = mk(0, 0, 0), pen = mk(0, 0, 0);
  for (int s = 0; s < substeps; ++s) {
    // 1. joint drives: lane j of the team integrates joint j, the nine results are broadcast
    {
      const int j = min(t.lane & 15, 8);
      float qj = e.q[0], vj = e.qd[0];
#pragma unroll
      for (int i = 1; i < 9; ++i) { if (j == i) { qj = e.q[i]; vj = e.qd[i]; } }
      float uj = u[0];
#pragma unroll
      for (int i = 1; i < 9; ++i) { if (j == i) uj = u[i]; }
      const float m = j < 7 ? P.arm_inertia : P.finger_mass;
      float vs = (m * vj + h * D * uj) / (m + h * D);
      const float f = D * (uj - vs);
      if (f > P.effort[j]) vs = vj + h * P.effort[j] / m;
      else if (f < -P.effort[j]) vs = vj - h * P.effort[j] / m;
      vs = clampf(vs, -P.qd_limit[j], P.qd_limit[j]);
      if (qj <= P.q_lower[j] && vs < 0.0f) vs = 0.0f;
      if (qj >= P.q_upper[j] && vs > 0.0f) vs = 0.0f;
#pragma unroll
      for (int i = 0; i < 9; ++i) e.qd[i] = __shfl_sync(kFull, vs, t.team_base + i);
    }
    // 2. gravity on the group's cube
    e.cu.v.z -= P.gravity * h;
    // 3. geometry of this sub-step (positions are fixed until step 4)
    Hand H;
    panda_hand(P, e.q, e.qd, true, H);
    OBox3 lbox[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const float* cen = f < 2 ? P.finger_center : P.hand_center;
      const float* hf = f < 2 ? P.finger_half : P.hand_half;
      V3 l = mk(cen[0], cen[1], cen[2]);
      if (f == 0) { l.y += e.q[7]; l.z += kFingerZ; }
      if (f == 1) { l.y = -l.y - e.q[8]; l.z += kFingerZ; }
      lbox[f].c = H.p + mul(H.R, l);
      lbox[f].R = H.R;
      lbox[f].half = mk(hf[0], hf[1], hf[2]);
    }
    float slide[2] = {e.qd[7], e.qd[8]};
    OBox3 cb;  // own cube
    cb.c = e.cu.p; cb.R = quat_to_R(e.cu.qx, e.cu.qy, e.cu.qz, e.cu.qw); cb.half = half_own;
    const V3 pc = box_corner(cb, t.c);  // own corner
    V3 v = e.cu.v, w = e.cu.w;
    const V3 x = e.cu.p;
    // the other cube of the sample (for cube-cube contact)
    const V3 xo = shfl3(x, t.lane ^ 8);
    bool cc_near;
    {
      const V3 d = xo - x;
      const float r = rad_own + rad_oth + P.contact_margin;
      cc_near = dot(d, d) <= r * r;  // cheap symmetric pre-test; the exact test follows if it passes
    }
    OBox3 ob;
    ob.c = xo; ob.half = half_oth;
    ob.R = cb.R;
    if (__any_sync(kFull, cc_near)) {
      ob.R.cx = shfl3(cb.R.cx, t.lane ^ 8); ob.R.cy = shfl3(cb.R.cy, t.lane ^ 8); ob.R.cz = shfl3(cb.R.cz, t.lane ^ 8);
      // exact test of the thread-per-sample code: sphere of cubeA against the box of cubeB
      cc_near = cc_near && (g == 0 ? boxes_near(cb, ob, P.contact_margin) : boxes_near(ob, cb, P.contact_margin));
    }
    // link / cube proximity: pair (f, i) is handled by group i
    bool lnear[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) lnear[f] = boxes_near(lbox[f], cb, P.contact_margin);

    // which fixed boxes are close to the own cube (bit k), decided once per sub-step
    unsigned near_mask = 0u;
    for (int k = 0; k < P.n_static; ++k)
      if (boxes_near(cb, obox_of(P.st[k]), P.contact_margin)) near_mask |= 1u << k;
    unsigned near_any = near_mask;
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) near_any |= __shfl_xor_sync(kFull, near_any, o);

    for (int p = 0; p < passes; ++p) {
      // (a) own cube against the fixed boxes
      for (unsigned todo_k = near_any; todo_k; todo_k &= todo_k - 1) {
        const int k = __ffs(todo_k) - 1;
        const OBox3 sb = obox_of(P.st[k]);
        const bool near = (near_mask >> k) & 1u;
        V3 n = mk(0, 0, 0);
        float depth = 0.0f;
        const bool hit = near && point_in_box(pc, sb, P.contact_margin, n, depth);
        const float mu = 0.5f * (mu_c + P.st[k].mu);
        const V3 got = apply_hits(hit, n, depth, pc, t.lane, t.group_base, true, [&](V3 nj, float dj, V3 pj) {
          return solve_cube_static(v, w, im, ii, x, nj, dj, pj, mu, h, P);
        });
        // impulses received by the fixed box = -(impulses on the cube)
        if (k == P.idx_table) imp_table = imp_table - got;
        if (k == P.idx_shelf) imp_shelf = imp_shelf - got;
        if (g == 1) imp_cubeb = imp_cubeb + got;
      }
      // (b) cubeA against cubeB, both ways; every lane of the team applies every impulse to replicas of both cubes
      if (__any_sync(kFull, cc_near)) {
        const V3 vo = shfl3(v, t.lane ^ 8), wo = shfl3(w, t.lane ^ 8);
        Dyn3 A = g == 0 ? dyn_cube(v, w, x, im, ii) : dyn_cube(vo, wo, xo, imo, iio);   // cubeA
        Dyn3 B = g == 0 ? dyn_cube(vo, wo, xo, imo, iio) : dyn_cube(v, w, x, im, ii);   // cubeB
        const float mu = 0.5f * (P.cube_mu[0] + P.cube_mu[1]);
        V3 n = mk(0, 0, 0);
        float depth = 0.0f;
        // corners of cubeA in cubeB (found by group 0), normal out of cubeB
        bool hit = cc_near && g == 0 && point_in_box(pc, ob, P.contact_margin, n, depth);
        V3 got = apply_hits(hit, n, depth, pc, t.lane, t.team_base, cc_near, [&](V3 nj, float dj, V3 pj) {
          return solve_contact3_call(A, B, nj, dj, pj, mu, h, P);
        });
        // corners of cubeB in cubeA (found by group 1), normal out of cubeA -> solve with -n
        hit = cc_near && g == 1 && point_in_box(pc, ob, P.contact_margin, n, depth);
        got = got + apply_hits(hit, n, depth, pc, t.lane, t.team_base + 8, cc_near, [&](V3 nj, float dj, V3 pj) {
          return solve_contact3_call(A, B, -nj, dj, pj, mu, h, P);
        });
        if (cc_near) {
          v = g == 0 ? A.v : B.v;
          w = g == 0 ? A.w : B.w;
          if (g == 1) imp_cubeb = imp_cubeb - got;  // cubeB received -(impulse on cubeA)
        }
      }
      // (c) links against the cubes in the order (f, cubeA), (f, cubeB); pair (f, i) is worked by group i
#pragma unroll
      for (int f = 0; f < 3; ++f) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const bool mine = g == i && lnear[f];
          if (!__any_sync(kFull, mine)) continue;
          Dyn3 L;
          L.v = H.v; L.w = H.w; L.x = H.p; L.im = 0.0f; L.ii = 0.0f;
          if (f < 2) { L.axis = (f == 0 ? 1.0f : -1.0f) * H.R.cy; L.slide = slide[f]; L.ims = 1.0f / P.finger_mass; }
          else { L.axis = mk(0, 0, 0); L.slide = 0.0f; L.ims = 0.0f; }
          Dyn3 C = dyn_cube(v, w, x, im, ii);
          const float mu = 0.5f * (P.robot_mu + mu_c);
          V3 n = mk(0, 0, 0);
          float depth = 0.0f;
          // corners of the link box in the cube, normal out of the cube
          const V3 lp = box_corner(lbox[f], t.c);
          bool hit = mine && point_in_box(lp, cb, P.contact_margin, n, depth);
          V3 got = apply_hits(hit, n, depth, lp, t.lane, t.team_base + 8 * i, mine, [&](V3 nj, float dj, V3 pj) {
            return solve_contact3_call(L, C, nj, dj, pj, mu, h, P);
          });
          // corners of the cube in the link box, normal out of the link -> solve with -n
          hit = mine && point_in_box(pc, lbox[f], P.contact_margin, n, depth);
          got = got + apply_hits(hit, n, depth, pc, t.lane, t.team_base + 8 * i, mine, [&](V3 nj, float dj, V3 pj) {
            return solve_contact3_call(L, C, -nj, dj, pj, mu, h, P);
          });
          if (mine) {
            v = C.v; w = C.w;
            if (i == 1) imp_cubeb = imp_cubeb - got;
          }
          if (f < 2) {
            // the finger's sliding speed is shared by both groups: take it from the group that worked the pair
            const float sl = __shfl_sync(kFull, L.slide, t.team_base + 8 * i);
            const bool worked = __shfl_sync(kFull, (int)mine, t.team_base + 8 * i) != 0;
            if (worked) slide[f] = sl;
          }
        }
      }
    }
    e.cu.v = v; e.cu.w = w;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      slide[f] = clampf(slide[f], -P.qd_limit[7 + f], P.qd_limit[7 + f]);
      e.qd[7 + f] = slide[f];
    }
    // (d) kinematic links against the fixed boxes named by the collision cost: group 0 -> table, group 1 -> shelf
    {
      const int ks = g == 0 ? P.idx_table : P.idx_shelf;
      if (ks >= 0) {
        const OBox3 sb = obox_of(P.st[ks]);
        const float mu = 0.5f * (P.robot_mu + P.st[ks].mu);
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          if (!boxes_near(lbox[f], sb, 0.0f)) continue;
          const V3 lp = box_corner(lbox[f], t.c);
          V3 n; float depth;
          if (!point_in_box(lp, sb, 0.0f, n, depth)) continue;
          const float fn = P.penalty_stiffness * depth;
          V3 vel = H.v + cross(H.w, lp - H.p);
          if (f < 2) vel = vel + slide[f] * ((f == 0 ? 1.0f : -1.0f) * H.R.cy);
          const float vn = dot(vel, n);
          const V3 tv = vel - vn * n;
          const float vt = sqrtf(dot(tv, tv));
          V3 fo = (-fn) * n;
          if (vt > 1e-6f) fo = fo + (mu * fn / vt) * tv;
          pen = pen + fo;
        }
      }
    }
    // 4. positions
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float qn = e.q[j] + h * e.qd[j];
      if (qn < P.q_lower[j]) { qn = P.q_lower[j]; e.qd[j] = 0.0f; }
      if (qn > P.q_upper[j]) { qn = P.q_upper[j]; e.qd[j] = 0.0f; }
      e.q[j] = qn;
    }
    {
      Cube& c = e.cu;
      c.p = c.p + h * c.v;
      const float qx = c.qx, qy = c.qy, qz = c.qz, qw = c.qw, hh = 0.5f * h;
      const float nx = qx + hh * (c.w.x * qw + c.w.y * qz - c.w.z * qy);
      const float ny = qy + hh * (c.w.y * qw + c.w.z * qx - c.w.x * qz);
      const float nz = qz + hh * (c.w.z * qw + c.w.x * qy - c.w.y * qx);
      const float nw = qw - hh * (c.w.x * qx + c.w.y * qy + c.w.z * qz);
      const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
      c.qx = nx * inv; c.qy = ny * inv; c.qz = nz * inv; c.qw = nw * inv;
    }
  }
  // reduce the lane-partial penalties over the 8 lanes of each group, then combine the two groups
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) pen = pen + shfl3(pen, t.lane ^ o);
  const V3 pen_o = shfl3(pen, t.lane ^ 8);
  const V3 pen_table = g == 0 ? pen : pen_o, pen_shelf = g == 0 ? pen_o : pen;
  const V3 it_o = shfl3(imp_table, t.lane ^ 8), is_o = shfl3(imp_shelf, t.lane ^ 8), ib_o = shfl3(imp_cubeb, t.lane ^ 8);
  const V3 it = g == 0 ? imp_table + it_o : it_o + imp_table;   // cubeA's share first, as in the serial order
  const V3 is = g == 0 ? imp_shelf + is_o : is_o + imp_shelf;
  const V3 ib = g == 1 ? imp_cubeb : ib_o;
  const float inv_dt = 1.0f / dt, inv_ns = 1.0f / (float)substeps;
  e.f_table = inv_dt * it + inv_ns * pen_table;
  e.f_shelf = inv_dt * is + inv_ns * pen_shelf;
  e.f_cubeb = inv_dt * ib;
}

// task cost of the sample, identical in all lanes of the team (cubeA's pose is fetched from group 0)
DEV float team_panda_cost(const TeamEnv& e, const PandaParams& P, const RolloutCfg& c, int kg, const PandaRef* ref,
                          const TeamLane& t) {
  PandaEnv full;
#pragma unroll
  for (int j = 0; j < 9; ++j) { full.q[j] = e.q[j]; full.qd[j] = e.qd[j]; }
  const int src = t.team_base;  // a lane of group 0 holds cubeA
  Cube a;
  a.p = shfl3(e.cu.p, src);
  a.qx = __shfl_sync(kFull, e.cu.qx, src); a.qy = __shfl_sync(kFull, e.cu.qy, src);
  a.qz = __shfl_sync(kFull, e.cu.qz, src); a.qw = __shfl_sync(kFull, e.cu.qw, src);
  a.v = mk(0, 0, 0); a.w = mk(0, 0, 0);
  full.cube[0] = a; full.cube[1] = a;
  full.f_table = e.f_table; full.f_shelf = e.f_shelf; full.f_cubeb = e.f_cubeb;
  PandaRef r;
  if (ref) r = *ref;
  else { r.cube0[0] = a.p.x; r.cube0[1] = a.p.y; r.cube0[2] = a.p.z; r.sel_axis = sel_axis_of(a); }
  return panda_cost(full, P, c, kg, r);
}

}  // namespace m3
