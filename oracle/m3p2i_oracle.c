/*
 * m3p2i_oracle.c — CPU oracle of the M3P2I hot path (sampling, rollout, costs, softmin update).
 * TEST INFRASTRUCTURE ONLY — see m3p2i_oracle.h for what is pinned against the reference and what is not.
 * Plain C, fp32 arithmetic, one sample at a time (OpenMP over samples when timed as the CPU baseline).
 */
#include "m3p2i_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "point_env.h"
#include "panda_env.h"
#include "halton_spline.h"

static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_get_threads(void) { return g_threads; }

struct Oracle {
  M3P2IConfig cfg;
  M3P2IPointScene ps;
  M3P2IPandaScene qs;
  int have_scene, have_state;
  int env_live;        /* 1: rollouts start from the persistent per-env state (a sim_step happened since set_state);
                          0: every env starts from the state given to set_state */
  int task, goal_len, gripper;
  float goal[8];
  float dof0[2 * M3P2I_MAX_NU];
  float root0[32 * 13];
  float* delta;        /* [K,T,nu] or NULL */
  float* delta_row0;   /* [T,nu] noise row of GLOBAL sample 0 (table mode, shards that do not own it) or NULL */
  float* filt;         /* [T,T] or NULL */
  M3P2IPlannerState st;
  /* results of the last command, reference layouts */
  float* actions;      /* [K,T,nu] (scaled by u_scale, as fed to the update, mppi.py:317) */
  float* states;       /* [K,T,4] */
  float* cost_h;       /* [K,T] */
  float* cost_sum;     /* [K] */
  float* J_local;      /* [K] */
  float* J_global;     /* [Kg] */
  float* weights;      /* [3,Kg] */
  M3P2ICommandInfo info;
  float scale[3], inv_eta[3];
  /* persistent sim facade */
  OPointEnv* penv;
  OPandaEnv* qenv;
  float* vel_target;   /* [K,nu] */
};

/* ------------------------------------------------------------------ Philox4x32-10 + Box-Muller */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                 uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* four N(0,1) draws for (global sample kg, time t, dimension group g) */
static inline void o_normal4(uint64_t seed, uint32_t kg, uint32_t t, uint32_t g, float z[4]) {
  uint32_t r[4];
  philox4x32_10(kg, t, g, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  for (int i = 0; i < 2; ++i) {
    float u1 = ((float)(r[2 * i] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(r[2 * i + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.0f * logf(u1));
    float ang = 6.283185307179586f * u2;
    z[2 * i] = rad * cosf(ang);
    z[2 * i + 1] = rad * sinf(ang);
  }
}

static inline float o_noise(const Oracle* o, int k, int t, int d) {
  const M3P2IConfig* c = &o->cfg;
  if (c->noise_mode == M3P2I_NOISE_TABLE) {
    return o->delta ? o->delta[((size_t)k * c->horizon + t) * c->nu + d] : 0.0f;
  }
  float z[4];
  if (c->noise_mode != M3P2I_NOISE_PHILOX_SPLINE) {
    o_normal4(c->seed, (uint32_t)(c->sample_offset + k), (uint32_t)t, (uint32_t)(d >> 2), z);
    return z[d & 3];
  }
  /* M3P2I_NOISE_PHILOX_SPLINE (include/m3p2i_b200.h): uniform quadratic B-spline over nseg + 2 control points drawn
   * at counters 0x40000000 + i, rescaled to unit variance */
  const int T = c->horizon, nseg = T / 4 > 2 ? T / 4 : 2;
  const float s = T > 1 ? (float)t * (float)nseg / (float)(T - 1) : 0.0f;
  int i0 = (int)s < nseg - 1 ? (int)s : nseg - 1;
  const float f = s - (float)i0;
  float w[3] = {0.5f * (1.0f - f) * (1.0f - f), 0.5f + f - f * f, 0.5f * f * f};
  const float inv = 1.0f / sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  float acc = 0.0f;
  for (int j = 0; j < 3; ++j) {
    o_normal4(c->seed, (uint32_t)(c->sample_offset + k), 0x40000000u + (uint32_t)(i0 + j), (uint32_t)(d >> 2), z);
    acc += (w[j] * inv) * z[d & 3];
  }
  return acc;
}

/* ------------------------------------------------------------------ lifecycle */
Oracle* orc_create(const M3P2IConfig* cfg) {
  if (!cfg || cfg->num_samples < 1 || cfg->horizon < 1 || cfg->horizon > M3P2I_MAX_HORIZON || cfg->nu < 1 ||
      cfg->nu > M3P2I_MAX_NU)
    return NULL;
  Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
  o->cfg = *cfg;
  if (o->cfg.num_samples_global <= 0) o->cfg.num_samples_global = o->cfg.num_samples;
  if (o->cfg.solver_passes <= 0) o->cfg.solver_passes = 2;
  if (o->cfg.substeps <= 0) o->cfg.substeps = 2;
  size_t K = cfg->num_samples, T = cfg->horizon, nu = cfg->nu, Kg = o->cfg.num_samples_global;
  o->actions = (float*)calloc(K * T * nu, 4);
  o->states = (float*)calloc(K * T * 4, 4);
  o->cost_h = (float*)calloc(K * T, 4);
  o->cost_sum = (float*)calloc(K, 4);
  o->J_local = (float*)calloc(K, 4);
  o->J_global = (float*)calloc(Kg, 4);
  o->weights = (float*)calloc(3 * Kg, 4);
  o->vel_target = (float*)calloc(K * nu, 4);
  o->st.beta = 1.0;
  for (int d = 0; d < M3P2I_MAX_NU; ++d) o->st.cov_action[d] = cfg->sigma[d] * cfg->sigma[d]; /* mppi.py:175 */
  o->task = cfg->env_type == M3P2I_ENV_POINT ? M3P2I_TASK_NAVIGATION : M3P2I_TASK_REACH;
  return o;
}

void orc_destroy(Oracle* o) {
  if (!o) return;
  free(o->actions); free(o->states); free(o->cost_h); free(o->cost_sum); free(o->J_local); free(o->J_global);
  free(o->weights); free(o->vel_target); free(o->delta); free(o->delta_row0); free(o->filt); free(o->penv); free(o->qenv);
  free(o);
}

int orc_set_scene_point(Oracle* o, const M3P2IPointScene* s) {
  if (!o || !s || o->cfg.env_type != M3P2I_ENV_POINT || s->n_static > M3P2I_MAX_STATIC || s->n_actors > 32) return -1;
  o->ps = *s; o->have_scene = 1; return 0;
}
int orc_set_scene_panda(Oracle* o, const M3P2IPandaScene* s) {
  if (!o || !s || o->cfg.env_type != M3P2I_ENV_PANDA || s->n_static > M3P2I_MAX_STATIC || s->n_actors > 32) return -1;
  o->qs = *s; o->have_scene = 1; return 0;
}

static int o_ndof(const Oracle* o) { return o->cfg.env_type == M3P2I_ENV_POINT ? 2 : 9; }
static int o_nactors(const Oracle* o) { return o->cfg.env_type == M3P2I_ENV_POINT ? o->ps.n_actors : o->qs.n_actors; }

int orc_set_state(Oracle* o, const float* dof, const float* root) {
  if (!o || !dof || !root || !o->have_scene) return -1;
  memcpy(o->dof0, dof, sizeof(float) * 2 * o_ndof(o));
  memcpy(o->root0, root, sizeof(float) * 13 * o_nactors(o));
  /* movable bodies that the integrator keeps fixed (the floating dyn-obs plate) follow the real state */
  {
    const int ns = o->cfg.env_type == M3P2I_ENV_POINT ? o->ps.n_static : o->qs.n_static;
    M3P2IBox* st = o->cfg.env_type == M3P2I_ENV_POINT ? o->ps.statics : o->qs.statics;
    for (int k = 0; k < ns; ++k)
      if (st[k].actor >= 0 && st[k].actor < o_nactors(o)) {
        memcpy(st[k].pos, root + 13 * st[k].actor, 12);
        memcpy(st[k].quat, root + 13 * st[k].actor + 3, 16);
      }
  }
  o->have_state = 1;
  o->env_live = 0;
  return 0;
}

int orc_set_objective(Oracle* o, int task, const float* goal, int goal_len, int gripper) {
  if (!o || goal_len < 0 || goal_len > 7 || (goal_len && !goal)) return -1;
  o->task = task; o->goal_len = goal_len; o->gripper = gripper;
  memset(o->goal, 0, sizeof(o->goal));
  for (int i = 0; i < goal_len; ++i) o->goal[i] = goal[i];
  return 0;
}

int orc_set_noise_table(Oracle* o, const float* delta) {
  if (!o) return -1;
  size_t n = (size_t)o->cfg.num_samples * o->cfg.horizon * o->cfg.nu;
  if (!delta) { free(o->delta); o->delta = NULL; return 0; }
  if (!o->delta) o->delta = (float*)malloc(n * 4);
  memcpy(o->delta, delta, n * 4);
  return 0;
}

int orc_set_noise_row0(Oracle* o, const float* row0) {
  if (!o) return -1;
  size_t n = (size_t)o->cfg.horizon * o->cfg.nu;
  if (!row0) { free(o->delta_row0); o->delta_row0 = NULL; return 0; }
  if (!o->delta_row0) o->delta_row0 = (float*)malloc(n * 4);
  memcpy(o->delta_row0, row0, n * 4);
  return 0;
}

int orc_set_noise_halton_spline(Oracle* o, int knot_scale, int degree, float smoothing, const unsigned short* perms,
                                int perm_stride) {
  if (!o || knot_scale < 1) return -1;
  const int K = o->cfg.num_samples, T = o->cfg.horizon, nu = o->cfg.nu;
  float* d = (float*)malloc((size_t)K * T * nu * 4);
  omp_set_num_threads(g_threads);
  int rc = hs_table(K, o->cfg.sample_offset, T, nu, knot_scale, degree, (double)smoothing, perms, perm_stride, d);
  if (!rc) rc = orc_set_noise_table(o, d);
  if (!rc && o->cfg.sample_offset != 0) {
    rc = hs_table(1, 0, T, nu, knot_scale, degree, (double)smoothing, perms, perm_stride, d);
    if (!rc) rc = orc_set_noise_row0(o, d);
  }
  free(d);
  return rc;
}

int orc_get_noise(Oracle* o, float* out) {
  if (!o || !out) return -1;
  const M3P2IConfig* c = &o->cfg;
  for (int k = 0; k < c->num_samples; ++k)
    for (int t = 0; t < c->horizon; ++t)
      for (int d = 0; d < c->nu; ++d) out[((size_t)k * c->horizon + t) * c->nu + d] = o_noise(o, k, t, d);
  return 0;
}

int orc_get_planner_state(Oracle* o, M3P2IPlannerState* out) { if (!o || !out) return -1; *out = o->st; return 0; }
int orc_set_planner_state(Oracle* o, const M3P2IPlannerState* in) { if (!o || !in) return -1; o->st = *in; return 0; }

int orc_set_filter_matrix(Oracle* o, const float* S) {
  if (!o) return -1;
  size_t n = (size_t)o->cfg.horizon * o->cfg.horizon;
  if (!S) { free(o->filt); o->filt = NULL; return 0; }
  if (!o->filt) o->filt = (float*)malloc(n * 4);
  memcpy(o->filt, S, n * 4);
  return 0;
}

/* ------------------------------------------------------------------ rollout */
typedef union { OPointEnv p; OPandaEnv q; } OEnv;

static void o_env_init(const Oracle* o, OEnv* e) {
  if (o->cfg.env_type == M3P2I_ENV_POINT) o_point_init(&e->p, &o->ps, o->dof0, o->root0);
  else o_panda_init(&e->q, &o->qs, o->dof0, o->root0);
}
/* start state of local env k: the persistent env after a sim_step, else the broadcast state (reactive_tamp.py:45-48) */
static void o_env_start(const Oracle* o, int k, OEnv* e) {
  if (o->env_live && k >= 0 && k < o->cfg.num_samples) {
    if (o->cfg.env_type == M3P2I_ENV_POINT) e->p = o->penv[k]; else e->q = o->qenv[k];
  } else o_env_init(o, e);
}
/* the rollout leaves the K envs (and their velocity targets) where it ended, as the reference's sim does */
static void o_env_store(Oracle* o, int k, const OEnv* e, const float* u_last) {
  if (o->cfg.env_type == M3P2I_ENV_POINT) { if (o->penv) o->penv[k] = e->p; }
  else if (o->qenv) o->qenv[k] = e->q;
  if (o->penv || o->qenv) memcpy(o->vel_target + (size_t)k * o->cfg.nu, u_last, sizeof(float) * o->cfg.nu);
}
static void o_env_step(const Oracle* o, OEnv* e, const float* u) {
  if (o->cfg.env_type == M3P2I_ENV_POINT) o_point_step(&e->p, &o->ps, &o->cfg, u);
  else o_panda_step(&e->q, &o->qs, &o->cfg, u);
}
static float o_env_cost(const Oracle* o, OEnv* e, int kg, const OPandaRef* ref) {
  if (o->cfg.env_type == M3P2I_ENV_POINT) return o_point_cost(&e->p, &o->cfg, o->task, o->goal, kg);
  return o_panda_cost(&e->q, &o->qs, &o->cfg, o->task, o->goal, kg, ref);
}
static void o_env_state_row(const Oracle* o, const OEnv* e, float* row) {
  if (o->cfg.env_type == M3P2I_ENV_POINT) { row[0] = e->p.px; row[1] = e->p.vx; row[2] = e->p.py; row[3] = e->p.vy; }
  else { row[0] = e->q.q[0]; row[1] = e->q.qd[0]; row[2] = e->q.q[1]; row[3] = e->q.qd[1]; }
}

/* mppi.py:275-332 for one sample; a_in = the perturbed action row-block [T,nu] of this sample.
 * `refs[t]` holds what the panda reach cost reads from OTHER rows of the batch after the same step: the cube
 * position of sample 0 (cost_functions.py:98,102-103) and, multi-modal, the cube axis chosen from the first row
 * of the second half (skill_utils.py:275-279 called on [half_samples:], cost_functions.py:151-152) */
static void o_rollout_sample(Oracle* o, int k, const float* a_in, const OPandaRef* refs) {
  const M3P2IConfig* c = &o->cfg;
  const int T = c->horizon, nu = c->nu, kg = c->sample_offset + k;
  OEnv e;
  o_env_start(o, k, &e);
  float run = 0.0f;
  for (int t = 0; t < T; ++t) {
    float u[M3P2I_MAX_NU];
    for (int d = 0; d < nu; ++d) u[d] = c->u_scale * a_in[t * nu + d];
    if (c->sample_null_action && kg == c->num_samples_global - 1)
      for (int d = 0; d < nu; ++d) u[d] = 0.0f;
    o_env_step(o, &e, u);
    OPandaRef self_ref;
    if (!refs && c->env_type == M3P2I_ENV_PANDA) { o_panda_ref(&e.q, &self_ref, 0); o_panda_ref(&e.q, &self_ref, 1); }
    float cost = o_env_cost(o, &e, kg, refs ? &refs[t] : &self_ref);
    run += cost; /* cost_samples += c, mppi.py:308 */
    o->cost_h[(size_t)k * T + t] = cost;
    o_env_state_row(o, &e, &o->states[((size_t)k * T + t) * 4]);
    for (int d = 0; d < nu; ++d) o->actions[((size_t)k * T + t) * nu + d] = u[d];
  }
  o->cost_sum[k] = run;
  o_env_store(o, k, &e, o->actions + ((size_t)k * T + (T - 1)) * nu);
}

/* one row of the GLOBAL batch (kg = 0 or K/2), replayed so every sample can see its state */
static void o_rollout_ref(Oracle* o, int kg, const float* a, OPandaRef* refs, int want_axis) {
  const M3P2IConfig* c = &o->cfg;
  OEnv e;
  o_env_start(o, kg - c->sample_offset, &e);
  for (int t = 0; t < c->horizon; ++t) {
    float u[M3P2I_MAX_NU];
    for (int d = 0; d < c->nu; ++d) u[d] = c->u_scale * a[t * c->nu + d];
    if (c->sample_null_action && kg == c->num_samples_global - 1)
      for (int d = 0; d < c->nu; ++d) u[d] = 0.0f;
    o_env_step(o, &e, u);
    o_panda_ref(&e.q, &refs[t], want_axis);
  }
}

/* mppi.py:386-416 for one sample: perturbed action sequence [T,nu] */
static void o_perturbed_action(const Oracle* o, int k, float* a) {
  const M3P2IConfig* c = &o->cfg;
  const int T = c->horizon, nu = c->nu, Kg = c->num_samples_global, kg = c->sample_offset + k;
  const int half = Kg / 2;
  for (int t = 0; t < T; ++t)
    for (int d = 0; d < nu; ++d) {
      float delta = o_noise(o, k, t, d);
      if (kg == Kg - 1) delta = 0.0f; /* delta[-1] = Z_seq, mppi.py:392 */
      /* scale_tril = sqrt(cov_action) (mppi.py:176,394,516): the configured sigma unless update_cov adapts it */
      float scaled = delta * (c->update_cov ? sqrtf(o->st.cov_action[d]) : c->sigma[d]);
      const float* mean = c->multi_modal ? (kg < half ? o->st.mean_action_1 : o->st.mean_action_2) : o->st.mean_action;
      float v = mean[t * nu + d] + scaled;
      v = fmaxf(fminf(v, c->u_max[d]), c->u_min[d]); /* scale_ctrl clamp, mppi_utils.py:36 */
      if (c->multi_modal) {
        if (kg == 0) v = o->st.best_traj_1[t * nu + d];
        if (kg == half) v = o->st.best_traj_2[t * nu + d];
      }
      if (c->env_type == M3P2I_ENV_PANDA && d >= 7) {
        if (o->gripper == M3P2I_GRIPPER_OPEN) v = 1.5f;
        else if (o->gripper == M3P2I_GRIPPER_CLOSE) v = -1.5f;
      }
      a[t * nu + d] = v;
    }
}

static int o_needs_refs(const Oracle* o) {
  return o->cfg.env_type == M3P2I_ENV_PANDA && o->task == M3P2I_TASK_REACH;
}

/* perturbed action of GLOBAL row kg (0 or K/2) as every shard can reconstruct it */
static int o_global_row_action(const Oracle* o, int kg, const float* actions_in, float* a) {
  const M3P2IConfig* c = &o->cfg;
  const int T = c->horizon, nu = c->nu, kl = kg - c->sample_offset;
  if (actions_in) {
    if (kl < 0 || kl >= c->num_samples) return -1; /* open-loop rollouts need the row in this shard */
    memcpy(a, actions_in + (size_t)kl * T * nu, sizeof(float) * T * nu);
    return 0;
  }
  if (kl >= 0 && kl < c->num_samples) { o_perturbed_action(o, kl, a); return 0; }
  Oracle tmp = *o;
  tmp.cfg.sample_offset = kg; /* local row 0 of tmp == global row kg */
  tmp.delta = NULL;
  if (c->noise_mode == M3P2I_NOISE_TABLE && !c->multi_modal) {
    if (kg != 0 || !o->delta_row0) return -1;
    tmp.delta = o->delta_row0;
  }
  o_perturbed_action(&tmp, 0, a);
  return 0;
}

static int o_rollout_all(Oracle* o, const float* actions_in /* [K,T,nu] or NULL = sample */) {
  const M3P2IConfig* c = &o->cfg;
  const int K = c->num_samples, T = c->horizon, nu = c->nu;
  OPandaRef* refs = NULL;
  if (o_needs_refs(o)) {
    refs = (OPandaRef*)calloc(T, sizeof(OPandaRef));
    float a0[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
    if (o_global_row_action(o, 0, actions_in, a0)) { free(refs); return -1; }
    o_rollout_ref(o, 0, a0, refs, 0);
    if (c->multi_modal) {
      if (o_global_row_action(o, c->num_samples_global / 2, actions_in, a0)) { free(refs); return -1; }
      o_rollout_ref(o, c->num_samples_global / 2, a0, refs, 1);
    }
  }
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int k = 0; k < K; ++k) {
    float a[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
    if (actions_in) memcpy(a, actions_in + (size_t)k * T * nu, sizeof(float) * T * nu);
    else o_perturbed_action(o, k, a);
    o_rollout_sample(o, k, a, refs);
  }
  free(refs);
  /* discounted cost J_k = cost_to_go(costs, gamma_seq)[:,0]: reversed cumulative sum, mppi_utils.py:106-113 */
  float gseq[M3P2I_MAX_HORIZON];
  gseq[0] = 1.0f;
  for (int t = 1; t < T; ++t) gseq[t] = gseq[t - 1] * c->gamma; /* torch.cumprod, mppi.py:182 */
  for (int k = 0; k < K; ++k) {
    float acc = 0.0f;
    for (int t = T - 1; t >= 0; --t) acc += gseq[t] * o->cost_h[(size_t)k * T + t];
    o->J_local[k] = acc / gseq[0];
  }
  if (o->penv || o->qenv) o->env_live = 1;
  return 0;
}

/* ------------------------------------------------------------------ update */
/* m3p2i.py:24-44; costs already min-shifted; beta starts at beta0; returns eta, writes scale = (float)(-1/beta) */
static float o_beta_search(const float* J, int n, float jmin, double beta0, float ub, float lb, float* scale_out,
                           double* beta_out, int* iters) {
  double beta = beta0;
  for (;;) {
    float scale = (float)(-1.0 / beta);
    float eta = 0.0f;
    for (int i = 0; i < n; ++i) eta += expf(scale * (J[i] - jmin));
    ++*iters;
    if (eta > ub) beta = beta * 0.9;
    else if (eta < lb) beta = beta * 1.2;
    else { *scale_out = scale; *beta_out = beta; return eta; } /* also taken when eta is NaN */
    *scale_out = scale; *beta_out = beta;
  }
}

static int o_argmin_first(const float* J, int n) {
  int b = 0;
  for (int i = 1; i < n; ++i) if (J[i] < J[b]) b = i;
  return b;
}

/* Phase 2: weights of all Kg samples from the gathered discounted costs (mppi.py:430-456, m3p2i.py:46-64) */
static void o_compute_stats(Oracle* o) {
  const M3P2IConfig* c = &o->cfg;
  const int Kg = c->num_samples_global, half = Kg / 2;
  const float* J = o->J_global;
  M3P2ICommandInfo* in = &o->info;
  memset(in, 0, sizeof(*in));
  in->near_samples = -1;   /* the oracle has no far-field split: every sample takes the full rollout */
  int lo[3] = {0, 0, half}, n[3] = {Kg, half, Kg - half};
  int nsets = c->multi_modal ? 3 : 1;
  for (int s = 0; s < nsets; ++s) {
    const float* Js = J + lo[s];
    int bi = o_argmin_first(Js, n[s]);
    float jmin = Js[bi];
    float scale, eta;
    double beta_used;
    if (c->multi_modal) {
      /* update_infinite_beta is called with self.beta / beta_1 / beta_2, which are never written back: 1 */
      eta = o_beta_search(Js, n[s], jmin, 1.0, 10.0f, 3.0f, &scale, &beta_used, &in->beta_iters);
    } else {
      beta_used = o->st.beta;
      scale = (float)(-1.0 / beta_used);
      eta = 0.0f;
      for (int i = 0; i < n[s]; ++i) eta += expf(scale * (Js[i] - jmin));
    }
    float inv = 1.0f / eta;
    for (int i = 0; i < n[s]; ++i) o->weights[(size_t)s * Kg + lo[s] + i] = inv * expf(scale * (Js[i] - jmin));
    in->eta[s] = eta; in->beta[s] = (float)beta_used; in->min_cost[s] = jmin; in->best_idx[s] = lo[s] + bi;
    o->scale[s] = scale; o->inv_eta[s] = inv;
  }
  if (c->multi_modal) {
    float wp = 0.0f, wq = 0.0f;
    for (int i = 0; i < half; ++i) wp += o->weights[i];
    for (int i = half; i < Kg; ++i) wq += o->weights[i];
    in->weight_push = wp; in->weight_pull = wq;
  }
}

int orc_partials_len(Oracle* o) { return 7 * o->cfg.horizon * o->cfg.nu + 1; }

/* Phase 3: this shard's weighted action sums, its best-trajectory rows, its sum of undiscounted costs */
static void o_compute_partials(Oracle* o, float* part) {
  const M3P2IConfig* c = &o->cfg;
  const int K = c->num_samples, T = c->horizon, nu = c->nu, Kg = c->num_samples_global, half = Kg / 2;
  const int TN = T * nu;
  memset(part, 0, sizeof(float) * (7 * TN + 1));
  int nsets = c->multi_modal ? 3 : 1;
  for (int k = 0; k < K; ++k) {
    int kg = c->sample_offset + k;
    const float* a = o->actions + (size_t)k * TN;
    float w0 = o->weights[kg];
    for (int i = 0; i < TN; ++i) { part[i] += w0 * a[i]; part[6 * TN + 1 + i] += w0 * a[i] * a[i]; }
    if (c->multi_modal) {
      int s = kg < half ? 1 : 2;
      float ws = o->weights[(size_t)s * Kg + kg];
      for (int i = 0; i < TN; ++i) part[s * TN + i] += ws * a[i];
    }
    for (int s = 0; s < nsets; ++s)
      if (o->info.best_idx[s] == kg)
        for (int i = 0; i < TN; ++i) part[(3 + s) * TN + i] = a[i];
    part[6 * TN] += o->cost_sum[k];
  }
}

/* Phase 4: mppi.py:494-503 / m3p2i.py:75-87, beta adaptation mppi.py:446-454, filter mppi.py:257-263 */
static void o_finish(Oracle* o, const float* part, float* out_action, float* out_cost_total) {
  const M3P2IConfig* c = &o->cfg;
  const int K = c->num_samples, T = c->horizon, nu = c->nu, Kg = c->num_samples_global;
  const int TN = T * nu;
  const float a1 = (float)(1.0 - (double)c->step_size_mean), a2 = c->step_size_mean;
  for (int i = 0; i < TN; ++i) o->st.mean_action[i] = a1 * o->st.mean_action[i] + a2 * part[i];
  if (c->multi_modal) {
    for (int i = 0; i < TN; ++i) {
      o->st.mean_action_1[i] = part[TN + i];
      o->st.mean_action_2[i] = part[2 * TN + i];
      o->st.best_traj_1[i] = part[4 * TN + i];
      o->st.best_traj_2[i] = part[5 * TN + i];
    }
  } else {
    for (int i = 0; i < TN; ++i) o->st.best_traj[i] = part[3 * TN + i];
    if (c->update_cov) {
      /* mppi.py:505-516: delta = actions - NEW mean; cov_update_d = mean_t sum_k w_k delta_ktd^2, from the moments
       * sum w a^2, sum w a and sum w = 1; step_size_cov = 0.7, kappa = 0.005 (mppi.py:202-203) */
      for (int d = 0; d < nu; ++d) {
        double acc = 0.0;
        for (int t = 0; t < T; ++t) {
          const double m = o->st.mean_action[t * nu + d];
          acc += (double)part[6 * TN + 1 + t * nu + d] - 2.0 * m * (double)part[t * nu + d] + m * m;
        }
        const float cov_update = (float)(acc / (double)T);
        o->st.cov_action[d] = (1.0f - 0.7f) * o->st.cov_action[d] + 0.7f * cov_update;
        o->st.cov_action[d] += 0.005f;
      }
    }
    if (c->env_type == M3P2I_ENV_PANDA) {
      if (o->info.eta[0] > 20.0f) o->st.beta = o->st.beta * 0.9;
      else if (o->info.eta[0] < 10.0f) o->st.beta = o->st.beta * 1.2;
    }
  }
  o->info.mean_cost_sum = part[6 * TN] / (float)Kg;
  if (out_cost_total)
    for (int k = 0; k < K; ++k) out_cost_total[k] = o->cost_sum[k] + o->info.mean_cost_sum; /* mppi.py:325 */
  if (out_action) {
    if (c->filter_u && o->filt) {
      for (int t = 0; t < T; ++t)
        for (int d = 0; d < nu; ++d) {
          float acc = 0.0f;
          for (int j = 0; j < T; ++j) acc += o->filt[t * T + j] * o->st.mean_action[j * nu + d];
          out_action[t * nu + d] = acc;
        }
    } else memcpy(out_action, o->st.mean_action, sizeof(float) * TN);
  }
}

/* mppi.py:266-273 */
static void o_shift(float* seq, int T, int nu) {
  for (int t = 0; t + 1 < T; ++t)
    for (int d = 0; d < nu; ++d) seq[t * nu + d] = seq[(t + 1) * nu + d];
  /* action_seq[-1] = saved_action: a view of the pre-roll tensor -> the last row keeps its old value */
}

static void o_shift_all(Oracle* o) {
  const int T = o->cfg.horizon, nu = o->cfg.nu;
  o_shift(o->st.mean_action, T, nu);
  if (o->cfg.multi_modal) {
    o_shift(o->st.mean_action_1, T, nu); o_shift(o->st.mean_action_2, T, nu);
    o_shift(o->st.best_traj_1, T, nu); o_shift(o->st.best_traj_2, T, nu);
  }
}

int orc_phase_rollout(Oracle* o, float* out_J_local) {
  if (!o || !o->have_scene || !o->have_state) return -1;
  o_shift_all(o);
  if (o_rollout_all(o, NULL)) return -1;
  if (out_J_local) memcpy(out_J_local, o->J_local, sizeof(float) * o->cfg.num_samples);
  return 0;
}

int orc_phase_partials(Oracle* o, const float* J_global, float* out_partials) {
  if (!o || !J_global || !out_partials) return -1;
  memcpy(o->J_global, J_global, sizeof(float) * o->cfg.num_samples_global);
  o_compute_stats(o);
  o_compute_partials(o, out_partials);
  return 0;
}

int orc_phase_finish(Oracle* o, const float* partials_sum, float* out_action, float* out_cost_total,
                     M3P2ICommandInfo* info) {
  if (!o || !partials_sum) return -1;
  o_finish(o, partials_sum, out_action, out_cost_total);
  if (info) *info = o->info;
  return 0;
}

int orc_command(Oracle* o, float* out_action, float* out_cost_total, M3P2ICommandInfo* info) {
  if (!o || o->cfg.num_samples != o->cfg.num_samples_global) return -1;
  int rc = orc_phase_rollout(o, NULL);
  if (rc) return rc;
  float* part = (float*)malloc(sizeof(float) * orc_partials_len(o));
  rc = orc_phase_partials(o, o->J_local, part);
  if (!rc) rc = orc_phase_finish(o, part, out_action, out_cost_total, info);
  free(part);
  return rc;
}

int orc_rollout_actions(Oracle* o, const float* actions, float* out_states, float* out_cost_h) {
  if (!o || !actions || !o->have_scene || !o->have_state) return -1;
  const M3P2IConfig* c = &o->cfg;
  if (o_rollout_all(o, actions)) return -1;
  if (out_states) memcpy(out_states, o->states, sizeof(float) * (size_t)c->num_samples * c->horizon * 4);
  if (out_cost_h) memcpy(out_cost_h, o->cost_h, sizeof(float) * (size_t)c->num_samples * c->horizon);
  return 0;
}

int orc_update_only(Oracle* o, const float* cost_h, const float* actions, float* out_mean, M3P2ICommandInfo* info) {
  if (!o || !cost_h || !actions || o->cfg.num_samples != o->cfg.num_samples_global) return -1;
  const M3P2IConfig* c = &o->cfg;
  const int K = c->num_samples, T = c->horizon, nu = c->nu;
  memcpy(o->cost_h, cost_h, sizeof(float) * (size_t)K * T);
  memcpy(o->actions, actions, sizeof(float) * (size_t)K * T * nu);
  float gseq[M3P2I_MAX_HORIZON];
  gseq[0] = 1.0f;
  for (int t = 1; t < T; ++t) gseq[t] = gseq[t - 1] * c->gamma;
  for (int k = 0; k < K; ++k) {
    float acc = 0.0f, run = 0.0f;
    for (int t = T - 1; t >= 0; --t) acc += gseq[t] * cost_h[(size_t)k * T + t];
    for (int t = 0; t < T; ++t) run += cost_h[(size_t)k * T + t];
    o->J_local[k] = acc; o->cost_sum[k] = run;
  }
  float* part = (float*)malloc(sizeof(float) * orc_partials_len(o));
  orc_phase_partials(o, o->J_local, part);
  int filt = o->cfg.filter_u;
  o->cfg.filter_u = 0;
  o_finish(o, part, out_mean, NULL);
  o->cfg.filter_u = filt;
  if (info) *info = o->info;
  free(part);
  return 0;
}

/* mppi.py:248-254: the n largest weights (ties: lower index first), positions (0,2) of the state rows */
int orc_top_trajs(Oracle* o, int n, int32_t* out_idx, float* out_w, float* out_trajs) {
  if (!o || n < 1) return -1;
  const M3P2IConfig* c = &o->cfg;
  const int Kg = c->num_samples_global, T = c->horizon;
  if (n > Kg) return -1;
  char* used = (char*)calloc(Kg, 1);
  for (int j = 0; j < n; ++j) {
    int b = -1;
    for (int i = 0; i < Kg; ++i) {
      if (used[i]) continue;
      float w = o->weights[i];
      if (b < 0 || w > o->weights[b] || (isnan(o->weights[b]) && !isnan(w))) b = i;
    }
    used[b] = 1;
    if (out_idx) out_idx[j] = b;
    if (out_w) out_w[j] = o->weights[b];
    if (out_trajs) {
      int k = b - c->sample_offset;
      for (int t = 0; t < T; ++t) {
        float x = 0.0f, y = 0.0f;
        if (k >= 0 && k < c->num_samples) { x = o->states[((size_t)k * T + t) * 4 + 0]; y = o->states[((size_t)k * T + t) * 4 + 2]; }
        out_trajs[((size_t)j * T + t) * 2 + 0] = x; out_trajs[((size_t)j * T + t) * 2 + 1] = y;
      }
    }
  }
  free(used);
  return 0;
}

int orc_read_buffer(Oracle* o, int which, float* out, size_t count) {
  if (!o || !out) return -1;
  const M3P2IConfig* c = &o->cfg;
  size_t K = c->num_samples, T = c->horizon, nu = c->nu, Kg = c->num_samples_global;
  const float* src; size_t n;
  switch (which) {
    case M3P2I_BUF_ACTIONS: src = o->actions; n = K * T * nu; break;
    case M3P2I_BUF_STATES: src = o->states; n = K * T * 4; break;
    case M3P2I_BUF_COST_HORIZON: src = o->cost_h; n = K * T; break;
    case M3P2I_BUF_COST_DISC: src = o->J_global; n = Kg; break;
    case M3P2I_BUF_COST_SUM: src = o->cost_sum; n = K; break;
    case M3P2I_BUF_WEIGHTS: src = o->weights; n = 3 * Kg; break;
    default: return -1;
  }
  if (count < n) return -1;
  if (which == M3P2I_BUF_ACTIONS && c->u_scale != 1.0f) { /* self.actions /= u_scale, mppi.py:420 */
    for (size_t i = 0; i < n; ++i) out[i] = src[i] / c->u_scale;
    return 0;
  }
  memcpy(out, src, n * 4);
  return 0;
}

/* ------------------------------------------------------------------ persistent sim facade */
int orc_sim_reset(Oracle* o) {
  if (!o || !o->have_scene || !o->have_state) return -1;
  const int K = o->cfg.num_samples;
  if (o->cfg.env_type == M3P2I_ENV_POINT) {
    if (!o->penv) o->penv = (OPointEnv*)malloc(sizeof(OPointEnv) * K);
    for (int k = 0; k < K; ++k) o_point_init(&o->penv[k], &o->ps, o->dof0, o->root0);
  } else {
    if (!o->qenv) o->qenv = (OPandaEnv*)malloc(sizeof(OPandaEnv) * K);
    for (int k = 0; k < K; ++k) o_panda_init(&o->qenv[k], &o->qs, o->dof0, o->root0);
  }
  o->env_live = 1;
  return 0;
}

int orc_sim_set_velocity_target(Oracle* o, const float* u) {
  if (!o || !u) return -1;
  memcpy(o->vel_target, u, sizeof(float) * (size_t)o->cfg.num_samples * o->cfg.nu);
  return 0;
}

int orc_sim_apply_forces(Oracle* o, const float* f_robot, const float* f_box) {
  if (!o || o->cfg.env_type != M3P2I_ENV_POINT) return -1;
  if (!o->env_live && orc_sim_reset(o)) return -1;
  for (int k = 0; k < o->cfg.num_samples; ++k) {
    if (f_robot) { o->penv[k].f_robot[0] = f_robot[2 * k]; o->penv[k].f_robot[1] = f_robot[2 * k + 1]; }
    if (f_box) { o->penv[k].f_box[0] = f_box[2 * k]; o->penv[k].f_box[1] = f_box[2 * k + 1]; }
  }
  return 0;
}

int orc_sim_step(Oracle* o) {
  if (!o) return -1;
  if (!o->env_live && orc_sim_reset(o)) return -1; /* set_state happened: broadcast it first */
  const int K = o->cfg.num_samples, nu = o->cfg.nu;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int k = 0; k < K; ++k) {
    if (o->cfg.env_type == M3P2I_ENV_POINT) o_point_step(&o->penv[k], &o->ps, &o->cfg, o->vel_target + (size_t)k * nu);
    else o_panda_step(&o->qenv[k], &o->qs, &o->cfg, o->vel_target + (size_t)k * nu);
  }
  return 0;
}

int orc_sim_write(Oracle* o, const float* dof, const float* root) {
  if (!o || !o->have_scene) return -1;
  if (!o->have_state) return -1; /* static actor poses come from set_state */
  if (!o->env_live && orc_sim_reset(o)) return -1;
  const int K = o->cfg.num_samples, na = o_nactors(o), nd = 2 * o_ndof(o);
  float* rtmp = (float*)malloc(sizeof(float) * 13 * na);
  float dtmp[2 * M3P2I_MAX_NU];
  for (int k = 0; k < K; ++k) {
    /* current values of env k, overridden by what the caller supplied */
    if (o->cfg.env_type == M3P2I_ENV_POINT) {
      const OPointEnv* e = &o->penv[k];
      dtmp[0] = e->px; dtmp[1] = e->vx; dtmp[2] = e->py; dtmp[3] = e->vy;
    } else {
      for (int j = 0; j < 9; ++j) { dtmp[2 * j] = o->qenv[k].q[j]; dtmp[2 * j + 1] = o->qenv[k].qd[j]; }
    }
    const float* d = dof ? dof + (size_t)k * nd : dtmp;
    const float* r = root ? root + (size_t)k * 13 * na : NULL;
    if (!r) { /* rebuild the root rows of env k from its bodies */
      float* keep = rtmp;
      Oracle* oo = o;
      if (oo->cfg.env_type == M3P2I_ENV_POINT) {
        memcpy(keep, o->root0, sizeof(float) * 13 * na);
        const M3P2IBody* bp[2] = {&o->ps.box, &o->ps.dyn_obs};
        for (int i = 0; i < 2; ++i) {
          float* row = keep + 13 * bp[i]->actor;
          const OBody2* b = &o->penv[k].b[i];
          row[0] = b->x; row[1] = b->y; row[3] = 0; row[4] = 0; row[5] = sinf(0.5f * b->th); row[6] = cosf(0.5f * b->th);
          row[7] = b->vx; row[8] = b->vy; row[12] = b->w;
        }
      } else o_panda_read(&o->qenv[k], &o->qs, o->root0, NULL, keep, NULL, NULL);
      r = keep;
    }
    if (o->cfg.env_type == M3P2I_ENV_POINT) {
      float f_robot[2] = {o->penv[k].f_robot[0], o->penv[k].f_robot[1]}, f_box[2] = {o->penv[k].f_box[0], o->penv[k].f_box[1]};
      o_point_init(&o->penv[k], &o->ps, d, r);
      o->penv[k].f_robot[0] = f_robot[0]; o->penv[k].f_robot[1] = f_robot[1];
      o->penv[k].f_box[0] = f_box[0]; o->penv[k].f_box[1] = f_box[1];
    } else o_panda_init(&o->qenv[k], &o->qs, d, r);
  }
  free(rtmp);
  return 0;
}

/* Objective.compute_cost on the persistent envs (cost_functions.py:19-36) */
int orc_sim_cost(Oracle* o, float* out) {
  if (!o || !out) return -1;
  if (!o->env_live && orc_sim_reset(o)) return -1;
  const M3P2IConfig* c = &o->cfg;
  OPandaRef ref;
  memset(&ref, 0, sizeof(ref));
  if (o_needs_refs(o)) {
    if (c->sample_offset != 0 || c->num_samples != c->num_samples_global) return -1;
    o_panda_ref(&o->qenv[0], &ref, 0);
    o_panda_ref(&o->qenv[c->multi_modal ? c->num_samples_global / 2 : 0], &ref, 1);
  }
  for (int k = 0; k < c->num_samples; ++k) {
    OEnv* e = c->env_type == M3P2I_ENV_POINT ? (OEnv*)&o->penv[k] : (OEnv*)&o->qenv[k];
    out[k] = o_env_cost(o, e, c->sample_offset + k, &ref);
  }
  return 0;
}

static void o_yaw_quat(float th, float* q) { q[0] = 0.0f; q[1] = 0.0f; q[2] = sinf(0.5f * th); q[3] = cosf(0.5f * th); }

int orc_sim_read(Oracle* o, float* dof, float* root, float* link, float* contact) {
  if (!o) return -1;
  if (!o->env_live && orc_sim_reset(o)) return -1;
  const int K = o->cfg.num_samples;
  if (o->cfg.env_type == M3P2I_ENV_POINT) {
    const int na = o->ps.n_actors;
    for (int k = 0; k < K; ++k) {
      const OPointEnv* e = &o->penv[k];
      if (dof) { float* d = dof + 4 * k; d[0] = e->px; d[1] = e->vx; d[2] = e->py; d[3] = e->vy; }
      if (root) {
        float* r = root + (size_t)k * na * 13;
        memcpy(r, o->root0, sizeof(float) * na * 13);
        const M3P2IBody* bp[2] = {&o->ps.box, &o->ps.dyn_obs};
        for (int i = 0; i < 2; ++i) {
          float* row = r + 13 * bp[i]->actor;
          row[0] = e->b[i].x; row[1] = e->b[i].y;
          o_yaw_quat(e->b[i].th, row + 3);
          row[7] = e->b[i].vx; row[8] = e->b[i].vy; row[9] = 0.0f;
          row[10] = 0.0f; row[11] = 0.0f; row[12] = e->b[i].w;
        }
      }
      if (link) {
        float* l = link + (size_t)k * 13;
        memset(l, 0, sizeof(float) * 13);
        l[0] = e->px; l[1] = e->py; l[2] = 0.05f; l[6] = 1.0f; l[7] = e->vx; l[8] = e->vy;
      }
      if (contact) { float* c = contact + (size_t)k * 3; c[0] = e->f_dyn[0]; c[1] = e->f_dyn[1]; c[2] = 0.0f; }
    }
  } else {
    for (int k = 0; k < K; ++k)
      o_panda_read(&o->qenv[k], &o->qs, o->root0, dof ? dof + (size_t)k * 18 : NULL,
                   root ? root + (size_t)k * o->qs.n_actors * 13 : NULL, link ? link + (size_t)k * 39 : NULL,
                   contact ? contact + (size_t)k * 9 : NULL);
  }
  return 0;
}

int orc_panda_fk(const M3P2IPandaScene* s, const float* q, const float* qd, float* link_state) {
  if (!s || !q || !link_state) return -1;
  float zero[9] = {0};
  o_panda_links(s, q, qd ? qd : zero, link_state);
  return 0;
}

/* ------------------------------------------------------------------ halton-spline noise table (halton_spline.h) */
int orc_halton_spline_table(int K, int offset, int T, int nu, int knot_scale, int degree, double smoothing,
                            const unsigned short* perms, int perm_stride, float* out) {
  if (!out || K <= 0 || T <= 0 || T > 256 || nu <= 0 || nu > 16 || knot_scale <= 0) return -1;
  omp_set_num_threads(g_threads);
  return hs_table(K, offset, T, nu, knot_scale, degree, smoothing, perms, perm_stride, out);
}
/* skill_utils.bspline on one knot vector */
int orc_bspline_samples(const float* cv, int m, int T, int degree, double smoothing, float* out) {
  if (!cv || !out || m <= degree || m > HS_MAXM || T <= 0 || T > 256) return -1;
  hs_bspline_samples(cv, m, T, degree, smoothing, out);
  return 0;
}
