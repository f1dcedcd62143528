// panda_far.cuh -- far-field Panda rollouts: the samples whose gripper never comes near a cube, the table or the shelf.
//
// While both cubes sleep (at rest on their support, nothing within the contact margin) the only things that evolve in a
// rollout are the seven velocity-tracked arm joints and the two fingers, nine scalar recurrences nobody else acts on;
// the team kernel's dormant shortcut (panda_team.cuh) already skips all cube work of such a sub-step, exactly. A rollout
// that stays in that regime for its WHOLE horizon needs none of the contact machinery, so it does not have to walk the
// team kernel's serial loop (one 255-register CTA per SM, latency-bound). This file evaluates those rollouts in three
// short lane-parallel phases per sample (16 lanes per sample, two samples per warp, 128-register CTAs of up to 15 warps):
//   1. actions:   lane l draws the perturbed action of steps l, l + 16, ... (mppi.py:392-416) into shared memory and
//                 stores the action planes;
//   2. joints:    lane j < 7 integrates arm joint j, lanes 7 / 8 the fingers, through all T * substeps sub-steps (drive,
//                 limits, position: the arithmetic of the team kernel's run-ahead) and leaves the positions every
//                 iteration sees in shared memory; lane 0 stores the state rows;
//   3. geometry:  lane l takes iterations l, l + 16, ...: forward kinematics from the stored positions, the link box
//                 centres, the dormancy pre-tests of BOTH cubes, the gripper's bounding sphere against table and shelf
//                 (the tests of panda_team.cuh, same expressions), and the cost of the step that ends there.
// A sample is FAR when the cubes sleep in the start state (far_base_asleep: the sleeping rule of the full path applied
// to the broadcast state) and every iteration passes every test: the serial path would have taken its dormant shortcut
// in each of them, so costs, state rows and actions are the ones it would have produced (same functions, same order of
// the cost sums). Every other sample goes onto the NEAR list and is rolled out by the team / thread-per-sample kernel
// launched next, which processes exactly the listed samples (RolloutBufs::near_list / near_count).
// Reach costs read rows 0 and Kg/2 of the batch: one extra warp per CTA replays those two rows the same way; if either
// leaves the far field, all samples of the launch are sent to the near list (the producer CTA of the next kernel then
// publishes the rows as before); if both stay far the rows are the start state's cube pose.
#pragma once
#include "panda_team.cuh"

namespace m3 {

constexpr int kFarLanes = 16;                 // lanes per sample
constexpr int kFarPerWarp = 32 / kFarLanes;   // samples per warp
constexpr int kFarBlockMax = 480;             // 14 sample warps + the warp of the batch rows

// floats of shared memory per sample: actions [T][9], joint positions [n_iter + 1][8] (7 arm joints + pad),
// finger openings [n_iter + 1][2], costs [T]; rounded so that consecutive samples start in different banks
__host__ __device__ inline int far_sample_floats(int T, int ns) {
  const int n = 9 * T + 10 * (T * ns + 1) + T + 16 * (T * ns / 8 + 1);
  return (n + 3) / 4 * 4 + 4;
}

// The sleeping rule of the full path (panda_team.cuh step 3 / panda_env.cuh) for the broadcast start state, minus its
// link test (the far tests of iteration 0 imply it): both cubes (almost) motionless, not touching each other, resting
// on their first near fixed box with at least three corners. Lanes 0-7 / 8-15 (and their copies 16-31) take the corners
// of cubeA / cubeB. Warp-uniform result.
DEV bool far_base_asleep(const RolloutCfg& c, const PandaParams& P, int& k0_out) {
  const int lane = threadIdx.x & 31, g = (lane >> 3) & 1, cn = lane & 7;
  TeamEnv e;
  e.load(c.base_env, 1, 0, g);
  const float h = c.dt / (float)c.substeps;
  const float im = 1.0f / P.cube_mass[g], ii = 1.0f / P.cube_inertia[g];
  const V3 half_own = mk(P.cube_half[g][0], P.cube_half[g][1], P.cube_half[g][2]);
  const V3 half_oth = mk(P.cube_half[g ^ 1][0], P.cube_half[g ^ 1][1], P.cube_half[g ^ 1][2]);
  const float rad_own = sqrtf(dot(half_own, half_own)), rad_oth = sqrtf(dot(half_oth, half_oth));
  OBox3 cb;
  cb.c = e.cu.p; cb.R = quat_to_R(e.cu.qx, e.cu.qy, e.cu.qz, e.cu.qw); cb.half = half_own;
  const V3 x = e.cu.p, v = e.cu.v, w = e.cu.w;
  const V3 ra = corner_arm(cb, cn);
  const int other = lane ^ 8;
  const V3 xo = shfl3(x, other);
  OBox3 ob;
  ob.c = xo; ob.half = half_oth;
  ob.R.cx = shfl3(cb.R.cx, other); ob.R.cy = shfl3(cb.R.cy, other); ob.R.cz = shfl3(cb.R.cz, other);
  bool cc_near;
  {
    const V3 d = xo - x;
    const float r = rad_own + rad_oth + P.contact_margin;
    cc_near = dot(d, d) <= r * r;
    cc_near = cc_near && (g == 0 ? boxes_near(cb, ob, P.contact_margin) : boxes_near(ob, cb, P.contact_margin));
  }
  const bool near_c = cn < P.n_static && boxes_near(cb, obox_of(P.st[min(cn, P.n_static - 1)]), P.contact_margin);
  const unsigned near_mask = (__ballot_sync(kFull, near_c) >> (lane & ~7)) & 0xffu;
  const int k0 = near_mask ? __ffs(near_mask) - 1 : 0;
  const StaticHit sh = static_hit(near_mask != 0u, x + ra, ra, obox_of(P.st[k0]), im, ii, __frcp_rn(h), P);
  const int sup = __popc((__ballot_sync(kFull, sh.sup) >> (lane & ~7)) & 0xffu);
  const bool asleep = P.sleep_lin > 0.0f && dot(v, v) < P.sleep_lin * P.sleep_lin && dot(w, w) < P.sleep_ang * P.sleep_ang &&
                      !cc_near && near_mask != 0u && sup >= 3;
  k0_out = k0;   // support box of the lane's cube (lanes 0-7: cubeA, 8-15: cubeB)
  return __all_sync(kFull, asleep);
}

// One warp = two samples. `k` = row of the lane's sample in this shard (-1: a batch row owned by another shard), `kg`
// its global id, `valid`: the sample exists and its outputs are to be stored. smem = the warp's 2 * far_sample_floats
// floats. Returns (per lane, uniform in a team) whether the sample stayed in the far field; the undiscounted / discounted
// cost sums come back in run / J (every lane of the team).
DEV bool far_team_eval(const RolloutCfg& c, const PandaParams& P, const RolloutBufs& b, float* smem, int k, int kg, bool valid,
                       float& run_out, float& J_out, int& boundary_out) {
  constexpr int NU = 9, TM = kFarLanes;
  const int lane = threadIdx.x & 31, tl = lane & (TM - 1), team_base = lane & ~(TM - 1);
  const int K = c.K, T = c.T, ns = c.substeps, n_iter = T * ns;
  const float h = c.dt / (float)ns, D = P.drive_damping;
  float* const su = smem + (lane / TM) * far_sample_floats(T, ns);   // [T][9]
  float* const sq = su + 9 * T;                                       // [n_iter + 1][8]
  float* const sf = sq + 8 * (n_iter + 1);                            // [n_iter + 1][2]
  float* const sc = sf + 2 * (n_iter + 1);                            // [T]
  float* const sbv = sc + T;                                          // [n_iter / 8 + 1][16] velocities at the boundaries

  // pass 0: the tests of iteration 0 alone (the start state's joint positions, the same for every sample): when the gripper
  // starts next to a cube or the table, nothing of this launch is far and the three phases are not worth running
  // pass 1: the three phases
  bool ok = true;
  int first_bad = n_iter;   // first iteration of this lane that fails a test
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
  if (pass == 0) {
    if (tl < 7) sq[tl] = c.base_env[2 * tl];
    if (tl == 7 || tl == 8) sf[tl - 7] = c.base_env[2 * tl];
    __syncwarp();
  } else {
  // ---- 1. perturbed actions of the whole horizon
  for (int t = tl; t < T; t += TM) {
    float u[NU];
    sample_action<NU>(c, b, kg, k, t, u);
#pragma unroll
    for (int d = 0; d < NU; ++d) {
      su[t * NU + d] = u[d];
      if (valid) b.actions[(size_t)(t * NU + d) * K + k] = u[d];
    }
  }
  __syncwarp();

  // ---- 2. the nine joints through the horizon (lane j: joint j; the arithmetic of the team kernel's run-ahead / finger
  // drives). sq / sf keep the position every iteration sees, entry n_iter the final one. Lanes 9-15 shadow the second
  // finger and write into the pad column of sq / the spare columns of sbv, so that the loop has no lane predicates;
  // all addresses are running pointers.
  {
    const int jo = min(tl, 8);
    const bool arm = jo < 7;
    float qo = c.base_env[2 * jo], vo = c.base_env[2 * jo + 1];
    const float lo_o = P.q_lower[jo], up_o = P.q_upper[jo], vl_o = P.qd_limit[jo], ef_o = P.effort[jo];
    const float m = arm ? P.joint_inertia[min(jo, 6)] : P.finger_mass;
    const float hD = h * D, den = m + h * D, kick = h * ef_o / m;
    float* pq = arm ? sq + jo : (tl < 9 ? sf + (jo - 7) : sq + 7);
    const int stride = (arm || tl >= 9) ? 8 : 2;
    float* pv = sbv + tl;
    const float* pu = su + jo;
    float4* ps = valid && tl == 0 ? b.states + k : nullptr;
    float uj = 0.0f;
#pragma unroll 1
    for (int it = 0, s = 0; it < n_iter; ++it) {
      if (s == 0) { uj = *pu; pu += NU; }
      if ((it & 7) == 0) { *pv = vo; pv += 16; }
      float vs = (m * vo + hD * uj) / den;
      const float f = D * (uj - vs);
      if (f > ef_o) vs = vo + kick;
      else if (f < -ef_o) vs = vo - kick;
      vs = clampf(vs, -vl_o, vl_o);
      if (qo <= lo_o && vs < 0.0f) vs = 0.0f;
      if (qo >= up_o && vs > 0.0f) vs = 0.0f;
      vo = vs;
      *pq = qo; pq += stride;
      float qn = qo + h * vo;
      if (qn < lo_o) { qn = lo_o; vo = 0.0f; }
      if (qn > up_o) { qn = up_o; vo = 0.0f; }
      qo = qn;
      if (++s == ns) {
        // state row of the finished step (q1, qd1, q2, qd2; reactive_tamp.py:66-69)
        const float q2 = __shfl_sync(kFull, qo, team_base + 1), v2 = __shfl_sync(kFull, vo, team_base + 1);
        if (ps) { *ps = make_float4(qo, vo, q2, v2); ps += K; }
        s = 0;
      }
    }
    *pq = qo;
    if ((n_iter & 7) == 0) *pv = vo;
  }
  __syncwarp();
  }   // pass 1

  // ---- 3. geometry tests and costs, lane l: iterations l, l + 16, ...
  TeamEnv e;   // cubeA of the start state (asleep: its pose is the pose of every step)
  e.load(c.base_env, 1, 0, 0);
  Cube a = e.cu;
  a.v = mk(0, 0, 0); a.w = mk(0, 0, 0);
  V3 xb;       // cubeB
  {
    TeamEnv eb;
    eb.load(c.base_env, 1, 0, 1);
    xb = eb.cu.p;
  }
  PandaRef ref;
  ref.cube0[0] = a.p.x; ref.cube0[1] = a.p.y; ref.cube0[2] = a.p.z; ref.sel_axis = sel_axis_of(a);
  const float frad = sqrtf(P.finger_half[0] * P.finger_half[0] + P.finger_half[1] * P.finger_half[1] + P.finger_half[2] * P.finger_half[2]);
  const float hrad = sqrtf(P.hand_half[0] * P.hand_half[0] + P.hand_half[1] * P.hand_half[1] + P.hand_half[2] * P.hand_half[2]);
  float rad_c[2];
#pragma unroll
  for (int g = 0; g < 2; ++g)
    rad_c[g] = sqrtf(P.cube_half[g][0] * P.cube_half[g][0] + P.cube_half[g][1] * P.cube_half[g][1] + P.cube_half[g][2] * P.cube_half[g][2]);
  const int ls = 31 - __clz(ns);   // ns is a power of two (checked by the launcher)
#pragma unroll 1
  for (int i = pass == 0 ? 0 : tl; i <= (pass == 0 ? 0 : n_iter); i += TM) {
    Hand H;
    H.p = mk(0, 0, 0); H.R.cx = mk(1, 0, 0); H.R.cy = mk(0, 1, 0); H.R.cz = mk(0, 0, 1); H.v = mk(0, 0, 0); H.w = mk(0, 0, 0);
    if (i < n_iter || c.task != M3P2I_TASK_PICK) {   // (the pick cost does not read the hand pose: panda_cost)
      float sn[7], cs[7], qd0[7];
      const float* pq = sq + i * 8;
#pragma unroll
      for (int j = 0; j < 7; ++j) { sincosf(pq[j], &sn[j], &cs[j]); qd0[j] = 0.0f; }
      fk_from_sincos(P, sn, cs, qd0, H);
    }
    const float q7 = sf[i * 2], q8 = sf[i * 2 + 1];
    if (i < n_iter) {
      V3 ll[3];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const float* cen = f < 2 ? P.finger_center : P.hand_center;
        V3 l = mk(cen[0], cen[1], cen[2]);
        if (f == 0) { l.y += q7; l.z += kFingerZ; }
        if (f == 1) { l.y = -l.y - q8; l.z += kFingerZ; }
        ll[f] = H.p + mul(H.R, l);
      }
      // the cheap tests of the dormant shortcut first (bounding spheres); where one fails the exact tests of the full
      // path decide: a link box whose sphere does not reach into the cube's box leaves the cube asleep (lnear == 0), and a
      // link box without a corner inside table / shelf adds nothing to the penalty -- the rollout is still contact-free
      bool dorm[2], sph[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const V3 px = g == 0 ? a.p : xb;
        const V3 d2 = ll[2] - px;
        const float rf = frad + rad_c[g] + P.contact_margin, rh = hrad + rad_c[g] + P.contact_margin;
        const float rfar = rh + kGripReach, dd2 = dot(d2, d2);
        dorm[g] = dd2 > rh * rh;
        if (dorm[g] && !(dd2 > rfar * rfar)) {
          const V3 d0 = ll[0] - px, d1 = ll[1] - px;
          dorm[g] = dot(d0, d0) > rf * rf && dot(d1, d1) > rf * rf;
        }
      }
      OBox3 gb;
      gb.c = ll[2]; gb.R = H.R; gb.half = mk(hrad + kGripReach, 0.0f, 0.0f);
      sph[0] = P.idx_table >= 0 && boxes_near(gb, obox_of(P.st[max(P.idx_table, 0)]), 0.0f);
      sph[1] = P.idx_shelf >= 0 && boxes_near(gb, obox_of(P.st[max(P.idx_shelf, 0)]), 0.0f);
      if (!(dorm[0] && dorm[1]) || sph[0] || sph[1]) {
        bool touch = false;
#pragma unroll 1
        for (int f = 0; f < 3; ++f) {
          OBox3 lb;
          lb.c = f == 0 ? ll[0] : (f == 1 ? ll[1] : ll[2]); lb.R = H.R;
          lb.half = f < 2 ? mk(P.finger_half[0], P.finger_half[1], P.finger_half[2]) : mk(P.hand_half[0], P.hand_half[1], P.hand_half[2]);
#pragma unroll 1
          for (int g = 0; g < 2; ++g) {
            if (dorm[g]) continue;
            TeamEnv ec;
            ec.load(c.base_env, 1, 0, g);
            OBox3 cb;
            cb.c = ec.cu.p; cb.R = quat_to_R(ec.cu.qx, ec.cu.qy, ec.cu.qz, ec.cu.qw);
            cb.half = mk(P.cube_half[g][0], P.cube_half[g][1], P.cube_half[g][2]);
            const V3 dl = lb.c - cb.c;
            const float rr = (f < 2 ? frad : hrad) + rad_c[g] + P.contact_margin;
            if (dot(dl, dl) > rr * rr) continue;
            if (boxes_near(lb, cb, P.contact_margin)) touch = true;
          }
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            if (!sph[q]) continue;
            const OBox3 sb = obox_of(P.st[q == 0 ? P.idx_table : P.idx_shelf]);
            if (!boxes_near(lb, sb, 0.0f)) continue;
            for (int cn = 0; cn < 8; ++cn) {
              V3 n; float depth;
              if (point_in_box(box_corner(lb, cn), sb, 0.0f, n, depth)) touch = true;
            }
          }
        }
        ok = ok && !touch;
      }
      if (!ok) first_bad = min(first_bad, i);
    }
    if ((i & (ns - 1)) == 0 && i > 0) {
      // first sub-step of a step (or the end of the horizon): cost of the step that just ended; no contact force is
      // reported in the far field, so the collision cost is zero
      float cost;
      if (c.task == M3P2I_TASK_REACH) cost = reach_combine(reach_parts(H, q7, q8, a, c, kg), c, kg, ref);
      else cost = panda_cost_from_hand(H, q7, q8, a, 0.0f, c, kg, ref);
      sc[(i >> ls) - 1] = cost;
    }
  }
  if (pass == 0 && !__all_sync(kFull, ok)) {
    run_out = 0.0f; J_out = 0.0f; boundary_out = 0;
    return false;
  }
  }   // passes
  __syncwarp();   // the costs in sc are read by every lane of the team below
  const unsigned okb = __ballot_sync(kFull, ok);
  const bool team_ok = ((okb >> team_base) & 0xffffu) == 0xffffu;
  // cost sums in step order (the order of the serial loop)
  float run = 0.0f, J = 0.0f, gam = 1.0f;
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    const float cost = sc[t];
    run += cost;
    J += gam * cost;
    gam *= c.gamma;
  }
  // the costs of the steps before a hand-over are final too (the step that ends at the boundary included)
  if (valid)
    for (int t = tl; t < T; t += TM) b.cost_h[(size_t)t * K + k] = sc[t];
  run_out = run; J_out = J;
#pragma unroll
  for (int o = TM / 2; o > 0; o >>= 1) first_bad = min(first_bad, __shfl_xor_sync(kFull, first_bad, o));
  const int bd = first_bad >> 3;
  boundary_out = bd;
  if (!team_ok && valid && b.far_dump) {
    const int nb = far_boundaries(T, ns);
    for (int idx = tl; idx < (bd + 1) * 9; idx += TM) {
      const int bi = idx / 9, j = idx - bi * 9, it = bi * 8;
      float* d = b.far_dump + ((size_t)k * nb + bi) * 18 + 2 * j;
      d[0] = j < 7 ? sq[it * 8 + j] : sf[it * 2 + (j - 7)];
      d[1] = sbv[bi * 16 + j];
    }
  }
  return team_ok;
}

}  // namespace m3
