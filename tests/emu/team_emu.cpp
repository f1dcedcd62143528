// team_emu.cpp -- runs the lane-cooperative Panda rollout (m3p2i-aip_b200/csrc/panda_team.cuh, the body of
// k_rollout_team) on the host, 32 fibers per warp in lock step (cuda_emu.h). TEST INFRASTRUCTURE ONLY: it lets the
// CPU test tier compare the kernel's device code with the oracle without a GPU. Built by tests/emu/build.py.
#include "cuda_emu.h"

#include <vector>

#include "../../m3p2i-aip_b200/csrc/params_host.h"
#include "../../m3p2i-aip_b200/csrc/panda_team.cuh"
#include "../../m3p2i-aip_b200/csrc/panda_far.cuh"

namespace emu { Warp* W = nullptr; }
using namespace m3;

namespace {
struct Launch {
  const RolloutCfg* c;
  const PandaParams* P;
  const RolloutBufs* b;
  int cpl;
};
Launch g_launch;

// far-field evaluation (panda_far.cuh) of the two samples of one warp: what k_rollout_far does per warp
struct FarLaunch {
  int k_first;        // row of the warp's first sample
  bool producer;      // the warp replays rows 0 / Kg/2 of the batch instead
  float* smem;
  int* ok;            // [K] (or [2] for the producer warp)
  float* run;
  float* J;
  int* boundary;      // [K] hand-over boundary of the samples that leave the far field (or nullptr)
};
FarLaunch g_far;

void far_lane_main() {
  const RolloutCfg& c = *g_launch.c;
  const int lane = threadIdx.x & 31, team = lane / kFarLanes;
  const int kraw = g_far.k_first + team;
  bool valid = !g_far.producer && kraw < c.K;
  int k = kraw < c.K ? kraw : c.K - 1, kg = c.offset + k;
  if (g_far.producer) {
    kg = (team == 1 && c.multi_modal) ? c.Kg / 2 : 0;
    k = (kg >= c.offset && kg < c.offset + c.K) ? kg - c.offset : -1;
  }
  float run = 0.0f, J = 0.0f;
  int k0 = 0, bd = 0;
  bool ok = far_base_asleep(c, *g_launch.P, k0);
  if ((lane == 0 || lane == 8) && g_launch.b->far_info) g_launch.b->far_info[lane >> 3] = k0;
  if (ok) ok = far_team_eval(c, *g_launch.P, *g_launch.b, g_far.smem, k, kg, valid, run, J, bd);
  if ((lane & (kFarLanes - 1)) == 0 && (valid || g_far.producer)) {
    const int slot = g_far.producer ? team : kraw;
    g_far.ok[slot] = ok ? 1 : 0;
    if (g_far.run) { g_far.run[slot] = run; g_far.J[slot] = J; }
    if (g_far.boundary && !g_far.producer) g_far.boundary[slot] = bd;
  }
  emu::me().done = true;
  swapcontext(&emu::me().ctx, &emu::W->sched);
}

void lane_main() {
  if (g_launch.cpl == 0) { far_lane_main(); return; }
  if (g_launch.cpl == 1) team_kernel_body<1>(*g_launch.c, *g_launch.P, *g_launch.b);
  else team_kernel_body<2>(*g_launch.c, *g_launch.P, *g_launch.b);
  emu::me().done = true;
  swapcontext(&emu::me().ctx, &emu::W->sched);
}

constexpr size_t kStack = 1 << 20;

long run_warp(emu::Warp& w, unsigned block, unsigned block_dim, unsigned warp_in_block, void* smem) {
  emu::W = &w;
  w.bid = {block, 0, 0}; w.bdim = {block_dim, 1, 1}; w.smem = smem; w.collectives = 0;
  for (int l = 0; l < 32; ++l) {
    emu::Lane& L = w.lane[l];
    L.done = false; L.n_coll = 0; L.tid = {warp_in_block * 32 + (unsigned)l, 0, 0};
    getcontext(&L.ctx);
    L.ctx.uc_stack.ss_sp = L.stack; L.ctx.uc_stack.ss_size = kStack; L.ctx.uc_link = &w.sched;
    makecontext(&L.ctx, lane_main, 0);
  }
  for (;;) {
    int done = 0;
    for (int l = 0; l < 32; ++l) {
      if (w.lane[l].done) { ++done; continue; }
      w.cur = l;
      swapcontext(&w.sched, &w.lane[l].ctx);
      if (w.lane[l].done) ++done;
    }
    if (done == 32) break;
    if (done != 0) {
      // some lanes returned while others wait at a collective: they would hang on the GPU
      bool all = true;
      for (int l = 0; l < 32; ++l) all = all && w.lane[l].done;
      if (!all) {
        int waiting = 0;
        for (int l = 0; l < 32; ++l) waiting += !w.lane[l].done;
        if (waiting != 32 && done != 32) { fprintf(stderr, "emu: %d lanes exited while %d wait at a collective\n", done, waiting); abort(); }
      }
    }
  }
  return w.collectives;
}
}  // namespace

extern "C" {

// Open-loop rollout of caller-supplied actions [K,T,nu] from one broadcast state, exactly what
// m3p2i_rollout_actions launches for panda_env with `lanes` (8 or 16) lanes per sample.
// Outputs (any may be NULL): states [K,T,4], cost_h [K,T], env_end [K,53], collectives (warp collectives executed).
int emu_team_rollout_actions(const M3P2IConfig* cfg, const M3P2IPandaScene* scene, int task, const float* goal,
                             int gripper, const float* dof, const float* root, const float* actions, int lanes,
                             int block_threads, float* out_states, float* out_cost_h, float* out_env_end,
                             long* out_collectives) {
  if (!cfg || !scene || !actions || (lanes != 8 && lanes != 16) || block_threads % 32 || block_threads < 32) return -1;
  const int K = cfg->num_samples, T = cfg->horizon, NU = 9, nf = kPandaEnvFloats;
  PandaParams P;
  memset(&P, 0, sizeof(P));
  M3P2IPandaScene sc = *scene;
  for (int k = 0; k < sc.n_static; ++k) {   // movable statics follow the real state (m3p2i_set_state)
    M3P2IBox& bx = sc.statics[k];
    if (bx.actor >= 0 && bx.actor < sc.n_actors) { memcpy(bx.pos, root + 13 * bx.actor, 12); memcpy(bx.quat, root + 13 * bx.actor + 3, 16); }
  }
  build_panda_params(sc, P);
  RolloutCfg c;
  memset(&c, 0, sizeof(c));
  c.K = K; c.T = T; c.nu = NU; c.Kg = cfg->num_samples_global > 0 ? cfg->num_samples_global : K; c.offset = cfg->sample_offset;
  c.multi_modal = cfg->multi_modal; c.null_action = cfg->sample_null_action; c.noise_mode = M3P2I_NOISE_TABLE;
  c.substeps = cfg->substeps > 0 ? cfg->substeps : 2; c.passes = cfg->solver_passes > 0 ? cfg->solver_passes : 2;
  c.task = task; c.gripper = gripper; c.env_live = 0; c.store_env = 1; c.open_loop = 1; c.align = 0; c.lanes = lanes;
  c.dt = cfg->dt; c.gamma = cfg->gamma; c.u_scale = cfg->u_scale; c.kp_suction = cfg->kp_suction;
  c.pre_height_diff = cfg->pre_height_diff; c.tilt_cos = cfg->tilt_cos_theta;
  memcpy(c.u_min, cfg->u_min, sizeof(c.u_min)); memcpy(c.u_max, cfg->u_max, sizeof(c.u_max)); memcpy(c.sigma, cfg->sigma, sizeof(c.sigma));
  if (goal) memcpy(c.goal, goal, sizeof(float) * 7);
  const bool refs = task == M3P2I_TASK_REACH;
  c.epoch = 64;

  std::vector<float> base(64, 0.0f), env((size_t)nf * K, 0.0f), vel((size_t)NU * K), act_in((size_t)T * NU * K), act((size_t)T * NU * K),
      cost_h((size_t)T * K), J(K), cost_sum(K), seq((size_t)SEQ_COUNT * T * NU, 0.0f);
  std::vector<float4> states((size_t)T * K);
  std::vector<PandaRef> ref_buf(T);
  unsigned flags[16] = {0};
  for (int j = 0; j < 18; ++j) base[j] = dof[j];
  memcpy(base.data() + 18, root + 13 * sc.cube_a.actor, sizeof(float) * 13);
  memcpy(base.data() + 31, root + 13 * sc.cube_b.actor, sizeof(float) * 13);
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < T * NU; ++j) act_in[(size_t)j * K + k] = actions[(size_t)k * T * NU + j];
  RolloutBufs b;
  memset(&b, 0, sizeof(b));
  b.seq = seq.data(); b.actions_in = act_in.data(); memcpy(c.base_env, base.data(), sizeof(c.base_env)); b.env = env.data(); b.vel_target = vel.data();
  b.actions = act.data(); b.states = states.data(); b.cost_h = cost_h.data(); b.J = J.data(); b.cost_sum = cost_sum.data();
  b.refs = refs ? ref_buf.data() : nullptr; b.ref_flags = flags;

  const int cpl = 16 / lanes, teams_per_block = block_threads / lanes, extra = refs ? 1 : 0;
  const int grid = (K + teams_per_block - 1) / teams_per_block + extra;
  g_launch = {&c, &P, &b, cpl};
  emu::Warp* w = new emu::Warp();
  for (int l = 0; l < 32; ++l) w->lane[l].stack = static_cast<char*>(malloc(kStack));
  std::vector<float4> smem((size_t)7 * cpl * block_threads + (size_t)(block_threads / lanes) * 2 * (kRecStride + T));
  long coll = 0;
  for (int blk = 0; blk < grid; ++blk)   // CTA 0 (the producer of the reach rows) first, as on the GPU
    for (int wi = 0; wi < block_threads / 32; ++wi) coll += run_warp(*w, blk, block_threads, wi, smem.data());
  for (int l = 0; l < 32; ++l) free(w->lane[l].stack);
  delete w;
  emu::W = nullptr;
  if (out_states)
    for (int k = 0; k < K; ++k)
      for (int t = 0; t < T; ++t) memcpy(out_states + ((size_t)k * T + t) * 4, &states[(size_t)t * K + k], 16);
  if (out_cost_h)
    for (int k = 0; k < K; ++k)
      for (int t = 0; t < T; ++t) out_cost_h[(size_t)k * T + t] = cost_h[(size_t)t * K + k];
  if (out_env_end)
    for (int k = 0; k < K; ++k)
      for (int f = 0; f < nf; ++f) out_env_end[(size_t)k * nf + f] = env[(size_t)f * K + k];
  if (out_collectives) *out_collectives = coll;
  return 0;
}

// Far-field evaluation of the same open-loop rollout (panda_far.cuh, the body of k_rollout_far): per sample whether it
// stayed in the far field (out_ok [K]); for those, states / cost_h / sums are what the kernel stores. out_prod_ok [2]:
// rows 0 and Kg/2 of the batch (reach).
int emu_far_rollout_actions(const M3P2IConfig* cfg, const M3P2IPandaScene* scene, int task, const float* goal, int gripper,
                            const float* dof, const float* root, const float* actions, int* out_ok, int* out_prod_ok,
                            float* out_states, float* out_cost_h, float* out_cost_sum, float* out_J) {
  if (!cfg || !scene || !actions || !out_ok) return -1;
  const int K = cfg->num_samples, T = cfg->horizon, NU = 9;
  PandaParams P;
  memset(&P, 0, sizeof(P));
  M3P2IPandaScene sc = *scene;
  for (int k = 0; k < sc.n_static; ++k) {
    M3P2IBox& bx = sc.statics[k];
    if (bx.actor >= 0 && bx.actor < sc.n_actors) { memcpy(bx.pos, root + 13 * bx.actor, 12); memcpy(bx.quat, root + 13 * bx.actor + 3, 16); }
  }
  build_panda_params(sc, P);
  RolloutCfg c;
  memset(&c, 0, sizeof(c));
  c.K = K; c.T = T; c.nu = NU; c.Kg = cfg->num_samples_global > 0 ? cfg->num_samples_global : K; c.offset = cfg->sample_offset;
  c.multi_modal = cfg->multi_modal; c.null_action = cfg->sample_null_action; c.noise_mode = M3P2I_NOISE_TABLE;
  c.substeps = cfg->substeps > 0 ? cfg->substeps : 2; c.passes = cfg->solver_passes > 0 ? cfg->solver_passes : 2;
  c.task = task; c.gripper = gripper; c.open_loop = 1; c.lanes = 16;
  c.dt = cfg->dt; c.gamma = cfg->gamma; c.u_scale = cfg->u_scale; c.kp_suction = cfg->kp_suction;
  c.pre_height_diff = cfg->pre_height_diff; c.tilt_cos = cfg->tilt_cos_theta;
  memcpy(c.u_min, cfg->u_min, sizeof(c.u_min)); memcpy(c.u_max, cfg->u_max, sizeof(c.u_max)); memcpy(c.sigma, cfg->sigma, sizeof(c.sigma));
  if (goal) memcpy(c.goal, goal, sizeof(float) * 7);
  if (c.substeps & (c.substeps - 1)) return -2;
  std::vector<float> base(64, 0.0f), act_in((size_t)T * NU * K), act((size_t)T * NU * K), cost_h((size_t)T * K), seq((size_t)SEQ_COUNT * T * NU, 0.0f);
  std::vector<float4> states((size_t)T * K);
  for (int j = 0; j < 18; ++j) base[j] = dof[j];
  memcpy(base.data() + 18, root + 13 * sc.cube_a.actor, sizeof(float) * 13);
  memcpy(base.data() + 31, root + 13 * sc.cube_b.actor, sizeof(float) * 13);
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < T * NU; ++j) act_in[(size_t)j * K + k] = actions[(size_t)k * T * NU + j];
  RolloutBufs b;
  memset(&b, 0, sizeof(b));
  b.seq = seq.data(); b.actions_in = act_in.data(); memcpy(c.base_env, base.data(), sizeof(c.base_env));
  b.actions = act.data(); b.states = states.data(); b.cost_h = cost_h.data();
  g_launch = {&c, &P, &b, 0};
  emu::Warp* w = new emu::Warp();
  for (int l = 0; l < 32; ++l) w->lane[l].stack = static_cast<char*>(malloc(kStack));
  std::vector<float> smem((size_t)kFarPerWarp * far_sample_floats(T, c.substeps));
  for (int k0 = 0; k0 < K; k0 += kFarPerWarp) {
    g_far = {k0, false, smem.data(), out_ok, out_cost_sum, out_J, nullptr};
    run_warp(*w, 0, 32, 0, nullptr);
  }
  if (out_prod_ok) {
    g_far = {0, true, smem.data(), out_prod_ok, nullptr, nullptr, nullptr};
    run_warp(*w, 0, 32, 0, nullptr);
  }
  for (int l = 0; l < 32; ++l) free(w->lane[l].stack);
  delete w;
  emu::W = nullptr;
  if (out_states)
    for (int k = 0; k < K; ++k)
      for (int t = 0; t < T; ++t) memcpy(out_states + ((size_t)k * T + t) * 4, &states[(size_t)t * K + k], 16);
  if (out_cost_h)
    for (int k = 0; k < K; ++k)
      for (int t = 0; t < T; ++t) out_cost_h[(size_t)k * T + t] = cost_h[(size_t)t * K + k];
  return 0;
}

// What a command launches for pick / place: the far-field kernel's device code over all samples, then the team kernel over
// the near list it leaves behind (rows + hand-over boundaries, joint dumps, costs of the steps before the hand-over).
// Outputs as emu_team_rollout_actions (no env_end) + out_far [K] (sample finished by the far-field code) and
// out_boundary [K].
int emu_split_rollout_actions(const M3P2IConfig* cfg, const M3P2IPandaScene* scene, int task, const float* goal, int gripper,
                              const float* dof, const float* root, const float* actions, int lanes, int block_threads,
                              float* out_states, float* out_cost_h, float* out_cost_sum, float* out_J, int* out_far,
                              int* out_boundary) {
  if (!cfg || !scene || !actions || (lanes != 8 && lanes != 16) || block_threads % 32) return -1;
  const int K = cfg->num_samples, T = cfg->horizon, NU = 9;
  PandaParams P;
  memset(&P, 0, sizeof(P));
  M3P2IPandaScene sc = *scene;
  for (int k = 0; k < sc.n_static; ++k) {
    M3P2IBox& bx = sc.statics[k];
    if (bx.actor >= 0 && bx.actor < sc.n_actors) { memcpy(bx.pos, root + 13 * bx.actor, 12); memcpy(bx.quat, root + 13 * bx.actor + 3, 16); }
  }
  build_panda_params(sc, P);
  RolloutCfg c;
  memset(&c, 0, sizeof(c));
  c.K = K; c.T = T; c.nu = NU; c.Kg = cfg->num_samples_global > 0 ? cfg->num_samples_global : K; c.offset = cfg->sample_offset;
  c.multi_modal = cfg->multi_modal; c.null_action = cfg->sample_null_action; c.noise_mode = M3P2I_NOISE_TABLE;
  c.substeps = cfg->substeps > 0 ? cfg->substeps : 2; c.passes = cfg->solver_passes > 0 ? cfg->solver_passes : 2;
  c.task = task; c.gripper = gripper; c.open_loop = 1; c.align = 0; c.lanes = lanes;
  c.dt = cfg->dt; c.gamma = cfg->gamma; c.u_scale = cfg->u_scale; c.kp_suction = cfg->kp_suction;
  c.pre_height_diff = cfg->pre_height_diff; c.tilt_cos = cfg->tilt_cos_theta;
  memcpy(c.u_min, cfg->u_min, sizeof(c.u_min)); memcpy(c.u_max, cfg->u_max, sizeof(c.u_max)); memcpy(c.sigma, cfg->sigma, sizeof(c.sigma));
  if (goal) memcpy(c.goal, goal, sizeof(float) * 7);
  if (c.substeps & (c.substeps - 1)) return -2;
  std::vector<float> base(64, 0.0f), act_in((size_t)T * NU * K), act((size_t)T * NU * K), cost_h((size_t)T * K), J(K), cost_sum(K),
      seq((size_t)SEQ_COUNT * T * NU, 0.0f), dump((size_t)K * far_boundaries(T, c.substeps) * 18, 0.0f);
  std::vector<float4> states((size_t)T * K);
  std::vector<int> ok(K, 0), bd(K, 0), list(K, 0);
  int info[8] = {0}, count = 0, count_next = 0;
  for (int j = 0; j < 18; ++j) base[j] = dof[j];
  memcpy(base.data() + 18, root + 13 * sc.cube_a.actor, sizeof(float) * 13);
  memcpy(base.data() + 31, root + 13 * sc.cube_b.actor, sizeof(float) * 13);
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < T * NU; ++j) act_in[(size_t)j * K + k] = actions[(size_t)k * T * NU + j];
  RolloutBufs b;
  memset(&b, 0, sizeof(b));
  b.seq = seq.data(); b.actions_in = act_in.data(); memcpy(c.base_env, base.data(), sizeof(c.base_env));
  b.actions = act.data(); b.states = states.data(); b.cost_h = cost_h.data(); b.J = J.data(); b.cost_sum = cost_sum.data();
  b.far_info = info; b.far_dump = dump.data();
  const bool refs = task == M3P2I_TASK_REACH;
  std::vector<PandaRef> ref_buf(T);
  unsigned flags[16] = {0};
  c.epoch = 64;
  b.refs = refs ? ref_buf.data() : nullptr; b.ref_flags = flags;
  emu::Warp* w = new emu::Warp();
  for (int l = 0; l < 32; ++l) w->lane[l].stack = static_cast<char*>(malloc(kStack));
  // 1. far-field code, warp by warp (k_rollout_far without its CTA-level bookkeeping)
  g_launch = {&c, &P, &b, 0};
  std::vector<float> fsm((size_t)kFarPerWarp * far_sample_floats(T, c.substeps));
  bool prod_bad = false;
  if (refs) {
    // the warp of the batch rows first; rows that stay far are published as the start state's cube pose
    int pok[2] = {0, 0};
    g_far = {0, true, fsm.data(), pok, nullptr, nullptr, nullptr};
    run_warp(*w, 0, 32, 0, nullptr);
    prod_bad = !(pok[0] && pok[1]);
    info[2] = prod_bad ? 0 : 1;
    if (!prod_bad) {
      Cube a;
      a.p = mk(base[18], base[19], base[20]); a.qx = base[21]; a.qy = base[22]; a.qz = base[23]; a.qw = base[24];
      a.v = mk(0, 0, 0); a.w = mk(0, 0, 0);
      for (int t = 0; t < T; ++t) { ref_buf[t].cube0[0] = a.p.x; ref_buf[t].cube0[1] = a.p.y; ref_buf[t].cube0[2] = a.p.z; ref_buf[t].sel_axis = sel_axis_of(a); }
      flags[0] = flags[1] = c.epoch + (unsigned)T;
    }
  }
  for (int k0 = 0; k0 < K; k0 += kFarPerWarp) {
    g_far = {k0, false, fsm.data(), ok.data(), cost_sum.data(), J.data(), bd.data()};
    run_warp(*w, 0, 32, 0, nullptr);
  }
  for (int k = 0; k < K; ++k) {
    if (prod_bad) { ok[k] = 0; bd[k] = 0; }
    if (!ok[k]) list[count++] = k | (bd[k] << kFarRowBits);
  }
  // 2. the team kernel over the near list
  b.near_list = list.data(); b.near_count = &count; b.near_count_next = &count_next;
  const int cpl = 16 / lanes, teams_per_block = block_threads / lanes;
  const int grid = (K + teams_per_block - 1) / teams_per_block + (refs ? 1 : 0);
  g_launch = {&c, &P, &b, cpl};
  std::vector<float4> smem((size_t)7 * cpl * block_threads + (size_t)(block_threads / lanes) * 2 * (kRecStride + T));
  for (int blk = 0; blk < grid; ++blk)   // CTA 0 (the producer of the reach rows) first, as on the GPU
    for (int wi = 0; wi < block_threads / 32; ++wi) run_warp(*w, blk, block_threads, wi, smem.data());
  for (int l = 0; l < 32; ++l) free(w->lane[l].stack);
  delete w;
  emu::W = nullptr;
  for (int k = 0; k < K; ++k) {
    for (int t = 0; t < T; ++t) {
      if (out_states) memcpy(out_states + ((size_t)k * T + t) * 4, &states[(size_t)t * K + k], 16);
      if (out_cost_h) out_cost_h[(size_t)k * T + t] = cost_h[(size_t)t * K + k];
    }
    if (out_cost_sum) out_cost_sum[k] = cost_sum[k];
    if (out_J) out_J[k] = J[k];
    if (out_far) out_far[k] = ok[k];
    if (out_boundary) out_boundary[k] = bd[k];
  }
  return 0;
}

// The same open-loop rollout with the thread-per-sample device code (panda_step / panda_cost of panda_env.cuh, the
// body of k_rollout<panda_env>, k_sim_step and the producer rows). pick / place only (no batch rows are read).
int emu_thread_rollout_actions(const M3P2IConfig* cfg, const M3P2IPandaScene* scene, int task, const float* goal,
                               int gripper, const float* dof, const float* root, const float* actions,
                               float* out_states, float* out_cost_h, float* out_env_end) {
  if (!cfg || !scene || !actions || task == M3P2I_TASK_REACH) return -1;
  const int K = cfg->num_samples, T = cfg->horizon, NU = 9, nf = kPandaEnvFloats;
  PandaParams P;
  memset(&P, 0, sizeof(P));
  M3P2IPandaScene sc = *scene;
  for (int k = 0; k < sc.n_static; ++k) {
    M3P2IBox& bx = sc.statics[k];
    if (bx.actor >= 0 && bx.actor < sc.n_actors) { memcpy(bx.pos, root + 13 * bx.actor, 12); memcpy(bx.quat, root + 13 * bx.actor + 3, 16); }
  }
  build_panda_params(sc, P);
  RolloutCfg c;
  memset(&c, 0, sizeof(c));
  c.K = K; c.T = T; c.nu = NU; c.Kg = cfg->num_samples_global > 0 ? cfg->num_samples_global : K; c.offset = cfg->sample_offset;
  c.multi_modal = cfg->multi_modal; c.task = task; c.gripper = gripper;
  c.substeps = cfg->substeps > 0 ? cfg->substeps : 2; c.passes = cfg->solver_passes > 0 ? cfg->solver_passes : 2;
  c.dt = cfg->dt; c.gamma = cfg->gamma; c.u_scale = cfg->u_scale; c.pre_height_diff = cfg->pre_height_diff; c.tilt_cos = cfg->tilt_cos_theta;
  if (goal) memcpy(c.goal, goal, sizeof(float) * 7);
  std::vector<float> base(64, 0.0f);
  for (int j = 0; j < 18; ++j) base[j] = dof[j];
  memcpy(base.data() + 18, root + 13 * sc.cube_a.actor, sizeof(float) * 13);
  memcpy(base.data() + 31, root + 13 * sc.cube_b.actor, sizeof(float) * 13);
  std::vector<float> env_out(nf);
  for (int k = 0; k < K; ++k) {
    PandaEnv e;
    e.load(base.data(), 1, 0);
    for (int t = 0; t < T; ++t) {
      float u[NU];
      for (int d = 0; d < NU; ++d) u[d] = c.u_scale * actions[((size_t)k * T + t) * NU + d];
      if (cfg->sample_null_action && c.offset + k == c.Kg - 1)
        for (int d = 0; d < NU; ++d) u[d] = 0.0f;
      panda_step(e, P, u, c.dt, c.substeps, c.passes);
      PandaRef r;
      r.cube0[0] = e.cube[0].p.x; r.cube0[1] = e.cube[0].p.y; r.cube0[2] = e.cube[0].p.z; r.sel_axis = sel_axis_of(e.cube[0]);
      const float cost = panda_cost(e, P, c, c.offset + k, r);
      if (out_cost_h) out_cost_h[(size_t)k * T + t] = cost;
      if (out_states) { const float4 row = e.state_row(); memcpy(out_states + ((size_t)k * T + t) * 4, &row, 16); }
    }
    if (out_env_end) { e.store(env_out.data(), 1, 0); memcpy(out_env_end + (size_t)k * nf, env_out.data(), sizeof(float) * nf); }
  }
  return 0;
}

}  // extern "C"
