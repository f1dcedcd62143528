// params_host.h -- host-side conversion of the interface scene structs (include/m3p2i_b200.h) into the parameter blocks
// the kernels read (common.cuh). Pure functions: used by api.cu and by the host build of the device code (tests/emu).
#pragma once
#include <cmath>
#include <cstring>
#include "common.cuh"

namespace m3 {

inline Static2 make_static2(const M3P2IBox& b) {
  const float x = b.quat[0], y = b.quat[1], z = b.quat[2], w = b.quat[3];
  const float c = 1.0f - 2.0f * (y * y + z * z), s = 2.0f * (w * z + x * y);
  const float n = sqrtf(c * c + s * s);
  Static2 r;
  r.cx = b.pos[0]; r.cy = b.pos[1]; r.hx = b.half[0]; r.hy = b.half[1];
  r.c = c / n; r.s = s / n; r.mu = b.mu; r.rad = sqrtf(r.hx * r.hx + r.hy * r.hy);
  return r;
}

inline Static3 make_static3(const M3P2IBox& b) {
  Static3 r;
  const float x = b.quat[0], y = b.quat[1], z = b.quat[2], w = b.quat[3];
  memcpy(r.c, b.pos, sizeof(r.c));
  memcpy(r.half, b.half, sizeof(r.half));
  r.R[0] = 1.0f - 2.0f * (y * y + z * z); r.R[1] = 2.0f * (x * y - w * z); r.R[2] = 2.0f * (x * z + w * y);
  r.R[3] = 2.0f * (x * y + w * z); r.R[4] = 1.0f - 2.0f * (x * x + z * z); r.R[5] = 2.0f * (y * z - w * x);
  r.R[6] = 2.0f * (x * z - w * y); r.R[7] = 2.0f * (y * z + w * x); r.R[8] = 1.0f - 2.0f * (x * x + y * y);
  r.mu = b.mu;
  return r;
}

inline void build_point_params(const M3P2IPointScene& s, PointParams& p) {
  p.robot_radius = s.robot_radius; p.robot_mass = s.robot_mass; p.robot_mu = s.robot_mu;
  p.drive_damping = s.drive_damping; p.drive_effort = s.drive_effort; p.gravity = s.gravity; p.ground_mu = s.ground_mu;
  p.contact_margin = s.contact_margin; p.baumgarte = s.baumgarte; p.slop = s.slop; p.max_corr_vel = s.max_corr_vel;
  p.box_hx = s.box.half[0]; p.box_hy = s.box.half[1]; p.box_mass = s.box.mass; p.box_inertia = s.box.inertia;
  p.box_mu = s.box.mu; p.box_reff = s.box.r_eff;
  p.dyn_hx = s.dyn_obs.half[0]; p.dyn_hy = s.dyn_obs.half[1]; p.dyn_mass = s.dyn_obs.mass;
  p.dyn_inertia = s.dyn_obs.inertia; p.dyn_mu = s.dyn_obs.mu; p.dyn_reff = s.dyn_obs.r_eff;
  p.n_static = s.n_static;
  for (int i = 0; i < s.n_static; ++i) p.st[i] = make_static2(s.statics[i]);
}

inline void build_panda_params(const M3P2IPandaScene& s, PandaParams& p) {
  memcpy(p.base, s.base_pos, sizeof(p.base));
  p.gravity = s.gravity;
  memcpy(p.q_lower, s.q_lower, sizeof(p.q_lower)); memcpy(p.q_upper, s.q_upper, sizeof(p.q_upper));
  memcpy(p.qd_limit, s.qd_limit, sizeof(p.qd_limit)); memcpy(p.effort, s.effort, sizeof(p.effort));
  p.drive_damping = s.drive_damping; p.finger_mass = s.finger_mass; p.robot_mu = s.robot_mu;
  for (int j = 0; j < 7; ++j) p.joint_inertia[j] = s.joint_inertia[j] > 0.0f ? s.joint_inertia[j] : s.arm_inertia;
  p.warm_start = s.warm_start; p.sleep_lin = s.sleep_lin; p.sleep_ang = s.sleep_ang; p.sleep_gap = s.sleep_gap;
  memcpy(p.finger_half, s.finger_half, 12); memcpy(p.finger_center, s.finger_center, 12);
  memcpy(p.hand_half, s.hand_half, 12); memcpy(p.hand_center, s.hand_center, 12);
  p.contact_margin = s.contact_margin; p.baumgarte = s.baumgarte; p.slop = s.slop; p.max_corr_vel = s.max_corr_vel;
  p.penalty_stiffness = s.penalty_stiffness;
  p.link_sweeps = s.link_sweeps > 0 ? s.link_sweeps : 2;
  p.report_cube = s.report_cube_contacts ? 1 : 0;
  const M3P2IBody* cb[2] = {&s.cube_a, &s.cube_b};
  for (int i = 0; i < 2; ++i) {
    memcpy(p.cube_half[i], cb[i]->half, 12);
    p.cube_mass[i] = cb[i]->mass; p.cube_inertia[i] = cb[i]->inertia; p.cube_mu[i] = cb[i]->mu;
  }
  p.n_static = s.n_static; p.idx_table = s.idx_table; p.idx_shelf = s.idx_shelf;
  for (int i = 0; i < s.n_static; ++i) p.st[i] = make_static3(s.statics[i]);
}

}  // namespace m3
