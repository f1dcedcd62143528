// api.cu — C ABI of libm3p2i_b200.so (include/m3p2i_b200.h): handle life cycle, host<->device staging, launch
// sequencing of one planner tick, the IsaacGymWrapper-style sim facade, and the K-sharded multi-GPU path (NCCL).
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "params_host.h"

using namespace m3;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(M3P2I_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                                      ":" + std::to_string(__LINE__) + ")");                             \
  } while (0)

// ---- NCCL, bound at run time so that single-GPU users need no NCCL at load time
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& why) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define SYM(field, name)                                                     \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name));               \
  if (!field) { why = std::string("dlsym ") + name + " failed"; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
  }
};
Nccl g_nccl;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (p && n >= count) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc(&p, sizeof(T) * std::max<size_t>(count, 1));
    if (e == cudaSuccess) { n = count; e = cudaMemset(p, 0, sizeof(T) * std::max<size_t>(count, 1)); }
    // The memset runs on the legacy default stream and may still be pending when this returns; the handles work on
    // non-blocking streams, which do not wait for it. Without this sync a kernel or copy enqueued right after an
    // allocation could be overwritten by the late memset (seen as a rare run-to-run difference of closed loops).
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct M3P2IHandle_ {
  M3P2IConfig cfg;
  int device = 0;
  int sm_count = 0;   // multiprocessors of `device` (the rollout kernel shape is chosen against it)
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evr = nullptr;
  bool have_scene = false, have_state = false, env_live = false, env_alloc = false, base_dirty = false;
  M3P2IPointScene ps_in;
  M3P2IPandaScene qs_in;
  PointParams pp;
  PandaParams qp;
  int task = 0, gripper = 0;
  float goal[8] = {0};
  int n_actors = 0, ndof = 0, nf = 0;
  std::vector<float> root0;  // [n_actors*13] last root state given to set_state
  // device buffers
  DevBuf<float> noise, noise_row0, seq, actions_in, base, env, vel_target, actions, cost_h, J, cost_sum, J_global,
      weights, partials, filt, cost_total, result, links, scratch;
  DevBuf<float4> states;
  DevBuf<PandaRef> refs;
  DevBuf<unsigned> ref_flags;
  unsigned ref_epoch = 0;
  DevBuf<int> near_list, near_count, far_info;   // far-field split (panda_far.cuh): rows left for the full rollout, two counters
  DevBuf<float> far_dump;                        // joint states of those rows at their hand-over boundaries
  unsigned far_epoch = 0;
  const int* near_count_used = nullptr;   // counter of the last rollout's far-field split (nullptr: it did not run)
  DevBuf<Stats> stats;
  DevBuf<M3P2ICommandInfo> info;
  bool have_noise = false, have_row0 = false, have_filt = false, have_evr = false;
  bool timed = false;   // the last command recorded its timing events (a caller that does not ask for the info scalars
                        // gets none: three event records less on the stream, k_stats overlaps the rollout's tail)
  // pinned host staging: results (grows on demand) and, separately, the packed base env written by set_state and
  // uploaded lazily (it must survive a reallocation of the result buffer)
  float* pin = nullptr;
  size_t pin_n = 0;
  float* pin_base = nullptr;   // [64] packed start state (host staging)
  float* mirror = nullptr;     // mapped pinned host memory: [2 T nu] result rows + M3P2ICommandInfo, written by the kernels
  float* mirror_dev = nullptr; // its device address
  M3P2ICommandInfo last_info;
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // exchange over peer memory (m3p2i_peer_export / m3p2i_peer_attach): layout in mailbox_layout()
  float* mailbox = nullptr;                 // this rank's mailbox (cudaMalloc, exported with cudaIpc)
  float* peer_box[kMaxPeers] = {nullptr};   // every rank's mailbox as mapped here ([rank] == mailbox)
  bool peer_ipc[kMaxPeers] = {false};       // mapping came from cudaIpcOpenMemHandle (close it on destroy)
  bool peer_on = false;
  unsigned peer_epoch = 0;
  unsigned peer_timeout_ms = 30000;   // M3P2I_PEER_TIMEOUT_MS at attach time
};

namespace {

typedef M3P2IHandle_ H;

int ndof_of(const H* h) { return h->cfg.env_type == M3P2I_ENV_POINT ? 2 : 9; }
int nf_of(const H* h) { return h->cfg.env_type == M3P2I_ENV_POINT ? kPointEnvFloats : kPandaEnvFloats; }

// one env in field order from the reference tensors (dof_state, root_state)
void pack_env(const H* h, const float* dof, const float* root, float* f) {
  memset(f, 0, sizeof(float) * h->nf);
  if (h->cfg.env_type == M3P2I_ENV_POINT) {
    f[0] = dof[0]; f[1] = dof[1]; f[2] = dof[2]; f[3] = dof[3];
    const M3P2IBody* bp[2] = {&h->ps_in.box, &h->ps_in.dyn_obs};
    for (int i = 0; i < 2; ++i) {
      const float* r = root + 13 * bp[i]->actor;
      const float x = r[3], y = r[4], z = r[5], w = r[6];
      float* o = f + 4 + 6 * i;
      o[0] = r[0]; o[1] = r[1];
      o[2] = atan2f(2.0f * (w * z + x * y), 1.0f - 2.0f * (y * y + z * z));
      o[3] = r[7]; o[4] = r[8]; o[5] = r[12];
    }
  } else {
    for (int j = 0; j < 18; ++j) f[j] = dof[j];
    const M3P2IBody* bp[2] = {&h->qs_in.cube_a, &h->qs_in.cube_b};
    for (int i = 0; i < 2; ++i) memcpy(f + 18 + 13 * i, root + 13 * bp[i]->actor, sizeof(float) * 13);
  }
}

// root rows + dof row of one env from its fields (inverse of pack_env; fixed actors keep the rows of root0)
void unpack_env(const H* h, const float* f, int stride, float* dof, float* root) {
  auto F = [&](int i) { return f[(size_t)i * stride]; };
  if (h->cfg.env_type == M3P2I_ENV_POINT) {
    if (dof) { dof[0] = F(0); dof[1] = F(1); dof[2] = F(2); dof[3] = F(3); }
    if (root) {
      memcpy(root, h->root0.data(), sizeof(float) * 13 * h->n_actors);
      const M3P2IBody* bp[2] = {&h->ps_in.box, &h->ps_in.dyn_obs};
      for (int i = 0; i < 2; ++i) {
        float* r = root + 13 * bp[i]->actor;
        const int o = 4 + 6 * i;
        const float th = F(o + 2);
        r[0] = F(o); r[1] = F(o + 1);
        r[3] = 0.0f; r[4] = 0.0f; r[5] = sinf(0.5f * th); r[6] = cosf(0.5f * th);
        r[7] = F(o + 3); r[8] = F(o + 4); r[9] = 0.0f; r[10] = 0.0f; r[11] = 0.0f; r[12] = F(o + 5);
      }
    }
  } else {
    if (dof) for (int j = 0; j < 18; ++j) dof[j] = F(j);
    if (root) {
      memcpy(root, h->root0.data(), sizeof(float) * 13 * h->n_actors);
      const M3P2IBody* bp[2] = {&h->qs_in.cube_a, &h->qs_in.cube_b};
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 13; ++j) root[13 * bp[i]->actor + j] = F(18 + 13 * i + j);
    }
  }
}

int ensure_pin(H* h, size_t n) {
  if (h->pin_n >= n) return 0;
  if (h->pin) cudaFreeHost(h->pin);
  h->pin = nullptr; h->pin_n = 0;
  CK(cudaMallocHost(&h->pin, sizeof(float) * n));
  h->pin_n = n;
  return 0;
}

int ensure_env(H* h) {
  if (h->env_alloc) return 0;
  const size_t K = h->cfg.num_samples;
  CK(h->env.alloc((size_t)h->nf * K));
  CK(h->vel_target.alloc((size_t)h->cfg.nu * K));
  h->env_alloc = true;
  return 0;
}

int upload_base(H* h);

// broadcast the set_state state into the K persistent envs
int materialize(H* h) {
  if (h->env_live) return 0;
  if (!h->have_state) return fail(M3P2I_ERR_STATE, "set_state has not been called");
  int rc = ensure_env(h);
  if (rc) return rc;
  if ((rc = upload_base(h))) return rc;
  launch_sim_reset(h->cfg.env_type, h->base.p, h->env.p, h->cfg.num_samples, h->stream);
  CK(cudaGetLastError());
  h->env_live = true;
  return 0;
}

// lanes per sample of the rollout kernel: the lane-cooperative team kernels shorten the per-sample serial chain and
// win while the GPU is not full; beyond that the redundant work of a team costs more than it hides.
// cfg.lanes_per_sample: 0 = choose, 1 = thread per sample, 8 / 16 = team of that size.
// Measured on B200 (profiles/r02_ksweep_*.csv), 148 SMs, CTAs of 7 warps, one CTA per SM (full register file):
//   K <= 14 * SMs (2072):  16 lanes, one wave
//   K <= 84 * SMs (12432):  8 lanes, one to three waves of 4144 samples (0.13 / 0.25 / 0.37-0.40 ms rollout at rest)
//   larger K: one thread per sample (0.44 ms flat up to K = 16384, then throughput-bound: 0.54 ms at K = 32768)
int rollout_lanes(const H* h) {
  if (h->cfg.env_type != M3P2I_ENV_PANDA) return 1;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("M3P2I_LANES");
    forced = e ? atoi(e) : 0;
  }
  int want = forced ? forced : h->cfg.lanes_per_sample;
  if (want == 1 || want == 8 || want == 16) return want;
  const int sms = h->sm_count > 0 ? h->sm_count : 148;
  if (h->cfg.num_samples <= 14 * sms) return 16;
  if (h->cfg.num_samples <= 84 * sms) return 8;
  return 1;
}

RolloutCfg make_rcfg(const H* h) {
  const M3P2IConfig& c = h->cfg;
  RolloutCfg r;
  memset(&r, 0, sizeof(r));
  r.K = c.num_samples; r.T = c.horizon; r.nu = c.nu; r.Kg = c.num_samples_global; r.offset = c.sample_offset;
  r.multi_modal = c.multi_modal; r.null_action = c.sample_null_action; r.noise_mode = c.noise_mode;
  r.substeps = c.substeps; r.passes = c.solver_passes; r.task = h->task; r.gripper = h->gripper;
  r.env_live = h->env_live ? 1 : 0; r.store_env = h->env_alloc ? 1 : 0; r.open_loop = 0;
  r.lanes = rollout_lanes(h);
  { static int a = -1; if (a < 0) { const char* e = getenv("M3P2I_TEAM_ALIGN"); a = e ? atoi(e) : 1; } r.align = a; }
  r.dt = c.dt; r.gamma = c.gamma; r.u_scale = c.u_scale; r.kp_suction = c.kp_suction;
  r.pre_height_diff = c.pre_height_diff; r.tilt_cos = c.tilt_cos_theta;
  memcpy(r.u_min, c.u_min, sizeof(r.u_min)); memcpy(r.u_max, c.u_max, sizeof(r.u_max));
  memcpy(r.sigma, c.sigma, sizeof(r.sigma));
  memcpy(r.goal, h->goal, sizeof(r.goal));
  r.seed_lo = (uint32_t)c.seed; r.seed_hi = (uint32_t)(c.seed >> 32);
  if (h->pin_base) memcpy(r.base_env, h->pin_base, sizeof(r.base_env));
  return r;
}

RolloutBufs make_rbufs(const H* h) {
  RolloutBufs b;
  b.noise = h->have_noise ? h->noise.p : nullptr;
  b.noise_row0 = h->have_row0 ? h->noise_row0.p : nullptr;
  b.seq = h->seq.p; b.actions_in = nullptr;
  b.sigma_dev = h->cfg.update_cov ? h->stats.p->sigma : nullptr; b.env = h->env.p; b.vel_target = h->vel_target.p;
  b.actions = h->actions.p; b.states = h->states.p; b.cost_h = h->cost_h.p; b.J = h->J.p; b.cost_sum = h->cost_sum.p;
  b.refs = nullptr;
  b.ref_flags = h->ref_flags.p;
  b.near_list = nullptr; b.near_count = nullptr; b.near_count_next = nullptr; b.far_info = nullptr; b.far_dump = nullptr;
  memset(&b.peer, 0, sizeof(b.peer));
  return b;
}

// words of one parity of a mailbox: Jg[Kg] | part[kMaxPeers][np] | jflag[kMaxPeers] | pflag[kMaxPeers]
struct MailboxLayout {
  size_t np, off_part, off_jflag, off_pflag, stride;
};
MailboxLayout mailbox_layout(const H* h) {
  MailboxLayout m;
  const size_t Kg = ((size_t)h->cfg.num_samples_global + 31) / 32 * 32;
  m.np = (7 * (size_t)h->cfg.horizon * h->cfg.nu + 1 + 31) / 32 * 32;
  m.off_part = Kg;
  m.off_jflag = m.off_part + kMaxPeers * m.np;
  m.off_pflag = m.off_jflag + 32;
  m.stride = m.off_pflag + 32;
  return m;
}
// pointers of this command's parity; call once per command after ++peer_epoch
void fill_peer(const H* h, PeerPush* push, PeerReduce* red) {
  const MailboxLayout m = mailbox_layout(h);
  const size_t par = (h->peer_epoch & 1u) * m.stride;
  if (push) {
    push->n = h->nranks; push->rank = h->rank; push->epoch = h->peer_epoch;
    for (int r = 0; r < h->nranks; ++r) {
      push->Jg[r] = h->peer_box[r] + par;
      push->jflag[r] = reinterpret_cast<unsigned*>(h->peer_box[r] + par + m.off_jflag);
    }
    push->ticket = h->ref_flags.p + 3;
  }
  if (red) {
    red->n = h->nranks; red->rank = h->rank; red->np = (int)m.np; red->epoch = h->peer_epoch;
    red->timeout_ms = h->peer_timeout_ms;
    red->jflag_local = reinterpret_cast<const unsigned*>(h->mailbox + par + m.off_jflag);
    red->part_local = h->mailbox + par + m.off_part;
    red->pflag_local = reinterpret_cast<const unsigned*>(h->mailbox + par + m.off_pflag);
    for (int r = 0; r < h->nranks; ++r) {
      red->part[r] = h->peer_box[r] + par + m.off_part;
      red->pflag[r] = reinterpret_cast<unsigned*>(h->peer_box[r] + par + m.off_pflag);
    }
    red->error = h->ref_flags.p + 4;
  }
}

UpdateCfg make_ucfg(const H* h, int shift) {
  const M3P2IConfig& c = h->cfg;
  UpdateCfg u;
  u.K = c.num_samples; u.T = c.horizon; u.nu = c.nu; u.Kg = c.num_samples_global; u.offset = c.sample_offset;
  u.multi_modal = c.multi_modal; u.env_type = c.env_type; u.filter_u = c.filter_u && h->have_filt; u.shift = shift;
  u.gamma = c.gamma; u.step_size_mean = c.step_size_mean;
  u.fuse_finish = 0;
  u.update_cov = c.update_cov && !c.multi_modal;
  u.stage_J = 0;
  return u;
}

// one rank owns every sample: the shard's discounted costs ARE the global ones (no gather, no copy)
bool local_is_global(const H* h) {
  return h->cfg.num_samples == h->cfg.num_samples_global && (h->nranks == 1 || !h->comm);
}

UpdateBufs make_ubufs(const H* h) {
  UpdateBufs b;
  b.J_global = local_is_global(h) ? h->J.p : h->J_global.p; b.weights = h->weights.p; b.stats = h->stats.p; b.actions = h->actions.p;
  b.cost_sum = h->cost_sum.p; b.partials = h->partials.p; b.seq = h->seq.p; b.filt = h->have_filt ? h->filt.p : nullptr;
  b.cost_total = h->cost_total.p; b.result = h->result.p; b.info = h->info.p;
  b.host_result = h->mirror_dev;
  b.host_info = h->mirror_dev ? reinterpret_cast<M3P2ICommandInfo*>(h->mirror_dev + 2 * (size_t)h->cfg.horizon * h->cfg.nu) : nullptr;
  b.near_count = h->near_count_used;
  b.done_counter = h->ref_flags.p + 2;
  b.stats_scratch = h->ref_flags.p + 8;
  memset(&b.peer, 0, sizeof(b.peer));
  return b;
}

// where the gathered discounted costs of the last command live
const float* j_global_ptr(const H* h) {
  if (h->peer_on && h->nranks > 1 && h->peer_epoch) return h->mailbox + (h->peer_epoch & 1u) * mailbox_layout(h).stride;
  return local_is_global(h) ? h->J.p : h->J_global.p;
}

bool needs_refs(const H* h) { return h->cfg.env_type == M3P2I_ENV_PANDA && h->task == M3P2I_TASK_REACH; }

int check_ready(const H* h) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  if (!h->have_scene) return fail(M3P2I_ERR_STATE, "set_scene_* has not been called");
  if (!h->have_state) return fail(M3P2I_ERR_STATE, "set_state has not been called");
  CK(cudaSetDevice(h->device));
  return 0;
}

int upload_base(H* h) {
  if (!h->base_dirty) return 0;
  // h->pin_base[0..nf) holds the packed base env (written by set_state)
  CK(cudaMemcpyAsync(h->base.p, h->pin_base, sizeof(float) * h->nf, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));   // set_state may rewrite the staging slot right away (sim facade path only)
  h->base_dirty = false;
  return 0;
}

// phase 1: sample + rollout of the local shard; J of the local shard lands in h->J (and, single rank, J_global)
int run_rollout(H* h, int* launches, const float* actions_in_dev, bool push_peers = false) {
  int rc = 0;
  RolloutCfg c = make_rcfg(h);   // carries the start state (base_env): no upload
  RolloutBufs b = make_rbufs(h);
  const bool refs = needs_refs(h);
  if (actions_in_dev) { c.open_loop = 1; b.actions_in = actions_in_dev; }
  if (refs) {
    const int half = c.Kg / 2;
    const bool own0 = c.offset == 0, own_half = half >= c.offset && half < c.offset + c.K;
    if (c.open_loop && (!own0 || (c.multi_modal && !own_half)))
      return fail(M3P2I_ERR_STATE, "open-loop reach rollouts need rows 0 and K/2 of the batch in this shard");
    if (!c.open_loop && !c.multi_modal && !own0 && c.noise_mode == M3P2I_NOISE_TABLE && !h->have_row0)
      return fail(M3P2I_ERR_STATE, "table noise: shards that do not own sample 0 need m3p2i_set_noise_row0");
    b.refs = h->refs.p;
    h->ref_epoch += 64;   // > max horizon: flags of this launch run from epoch+1 to epoch+T
    c.epoch = h->ref_epoch;
  }
  if (h->peer_on && push_peers) fill_peer(h, &b.peer, nullptr);
  if (h->near_list.p) {
    // counters alternate: this command's was cleared by the previous far-field launch (or is still zero)
    b.near_list = h->near_list.p;
    b.near_count = h->near_count.p + (h->far_epoch & 1u);
    b.near_count_next = h->near_count.p + ((h->far_epoch + 1u) & 1u);
    b.far_info = h->far_info.p; b.far_dump = h->far_dump.p;
    if (far_rollout_applies(h->cfg.env_type, c, b, refs)) { ++h->far_epoch; h->near_count_used = b.near_count; }
    else h->near_count_used = nullptr;
  }
  launch_rollout(h->cfg.env_type, c, &h->pp, &h->qp, b, refs, h->stream, launches);
  CK(cudaGetLastError());
  if (h->env_alloc) h->env_live = true;
  return 0;
}

int run_update(H* h, int shift, int* launches, bool fuse_finish = false) {
  UpdateCfg u = make_ucfg(h, shift);
  u.fuse_finish = fuse_finish ? 1 : 0;
  UpdateBufs b = make_ubufs(h);
  if (h->peer_on && fuse_finish && h->nranks > 1) {
    fill_peer(h, nullptr, &b.peer);
    b.J_global = h->mailbox + (h->peer_epoch & 1u) * mailbox_layout(h).stride;
  }
  launch_stats(u, b, h->stream, launches);
  launch_wsum(u, b, h->stream, launches);
  CK(cudaGetLastError());
  return 0;
}

int run_finish(H* h, int shift, int* launches) {
  UpdateCfg u = make_ucfg(h, shift);
  UpdateBufs b = make_ubufs(h);
  launch_finish(u, b, h->stream, launches);
  CK(cudaGetLastError());
  return 0;
}

int gather_J(H* h) {
  const size_t K = h->cfg.num_samples;
  if (h->nranks == 1 || !h->comm) {
    if (h->cfg.num_samples != h->cfg.num_samples_global)
      return fail(M3P2I_ERR_STATE, "sharded handle without a communicator: use the m3p2i_phase_* calls or m3p2i_comm_init");
    return 0;   // make_ubufs reads the costs where the rollout left them
  }
  ncclResult_t r = g_nccl.AllGather(h->J.p, h->J_global.p, K, ncclFloat, h->comm, h->stream);
  if (r != ncclSuccess) return fail(M3P2I_ERR_NCCL, std::string("ncclAllGather: ") + g_nccl.GetErrorString(r));
  return 0;
}

int reduce_partials(H* h) {
  if (h->nranks == 1 || !h->comm) return 0;
  const size_t n = 7 * (size_t)h->cfg.horizon * h->cfg.nu + 1;
  ncclResult_t r = g_nccl.AllReduce(h->partials.p, h->partials.p, n, ncclFloat, ncclSum, h->comm, h->stream);
  if (r != ncclSuccess) return fail(M3P2I_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
  return 0;
}

int fetch(H* h, float* out_action, float* out_cost_total, M3P2ICommandInfo* info, bool unfiltered) {
  const size_t TN = (size_t)h->cfg.horizon * h->cfg.nu, K = h->cfg.num_samples;
  const size_t need = 2 * TN + sizeof(M3P2ICommandInfo) / sizeof(float) + 1 + (out_cost_total ? K : 0);
  int rc = ensure_pin(h, std::max<size_t>(need, 256));
  if (rc) return rc;
  float* p = h->pin;
  M3P2ICommandInfo* pi = reinterpret_cast<M3P2ICommandInfo*>(p + 2 * TN);
  const bool mirrored = h->mirror != nullptr;   // the update kernels wrote both straight into mapped host memory
  if (!mirrored) {
    CK(cudaMemcpyAsync(p, h->result.p, sizeof(float) * 2 * TN, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(pi, h->info.p, sizeof(M3P2ICommandInfo), cudaMemcpyDeviceToHost, h->stream));
  }
  float* pc = p + 2 * TN + sizeof(M3P2ICommandInfo) / sizeof(float) + 1;
  if (out_cost_total) CK(cudaMemcpyAsync(pc, h->cost_total.p, sizeof(float) * K, cudaMemcpyDeviceToHost, h->stream));
  unsigned* perr = reinterpret_cast<unsigned*>(p + 2 * TN + sizeof(M3P2ICommandInfo) / sizeof(float));
  *perr = 0u;
  if (h->peer_on) CK(cudaMemcpyAsync(perr, h->ref_flags.p + 4, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (*perr) {
    // report once: clear the flag so that later commands are judged on their own exchange
    CK(cudaMemsetAsync(h->ref_flags.p + 4, 0, sizeof(unsigned), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return fail(M3P2I_ERR_STATE, "peer exchange timed out: a rank did not deliver its costs / partial sums");
  }
  if (mirrored) {
    p = h->mirror;
    pi = reinterpret_cast<M3P2ICommandInfo*>(h->mirror + 2 * TN);
  }
  if (out_action) memcpy(out_action, unfiltered ? p + TN : p, sizeof(float) * TN);
  if (out_cost_total) memcpy(out_cost_total, pc, sizeof(float) * K);
  const float kms = h->last_info.kernel_ms, rms = h->last_info.rollout_ms;
  const int nl = h->last_info.launches;
  h->last_info = *pi;
  h->last_info.kernel_ms = kms;
  h->last_info.rollout_ms = rms;
  h->last_info.launches = nl;
  h->last_info.rollout_lanes = rollout_lanes(h);
  if (info) *info = h->last_info;
  return 0;
}

int command_device(H* h, bool timed) {
  int rc = check_ready(h);
  if (rc) return rc;
  int launches = 0;
  h->timed = timed;
  h->have_evr = false;
  if (timed) CK(cudaEventRecord(h->ev0, h->stream));
  const bool peers = h->peer_on && h->nranks > 1;   // exchange fused into the kernels over peer memory
  if (peers) ++h->peer_epoch;
  if ((rc = run_rollout(h, &launches, nullptr, peers))) return rc;
  if (timed) {
    CK(cudaEventRecord(h->evr, h->stream));
    h->have_evr = true;
  }
  if (!peers && (rc = gather_J(h))) return rc;
  const bool fused = peers || h->nranks == 1 || !h->comm;   // no host-visible exchange between sums and mean update
  if ((rc = run_update(h, 1, &launches, fused))) return rc;
  if (!fused) {
    if ((rc = reduce_partials(h))) return rc;
    if ((rc = run_finish(h, 1, &launches))) return rc;
  }
  if (timed) CK(cudaEventRecord(h->ev1, h->stream));
  h->last_info.launches = launches;
  return 0;
}

int finish_timing(H* h) {
  float ms = 0.0f;
  h->last_info.rollout_lanes = rollout_lanes(h);
  if (!h->timed) {
    h->last_info.kernel_ms = 0.0f; h->last_info.rollout_ms = 0.0f;
    return 0;
  }
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_info.kernel_ms = ms;
  h->last_info.rollout_lanes = rollout_lanes(h);
  h->last_info.rollout_ms = 0.0f;
  if (h->have_evr) {
    CK(cudaEventElapsedTime(&ms, h->ev0, h->evr));
    h->last_info.rollout_ms = ms;
    h->have_evr = false;
  }
  return 0;
}

}  // namespace

// ====================================================================================================== C ABI
extern "C" {

const char* m3p2i_last_error(void) { return g_err.c_str(); }
int m3p2i_version(void) { return 100; }

int m3p2i_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int m3p2i_abi_sizeof(const char* name) {
  if (!name) return -1;
#define SZ(T) if (!strcmp(name, #T)) return (int)sizeof(T)
  SZ(M3P2IConfig); SZ(M3P2IBox); SZ(M3P2IBody); SZ(M3P2IPointScene); SZ(M3P2IPandaScene); SZ(M3P2IPlannerState);
  SZ(M3P2ICommandInfo); SZ(M3P2IPeerHandle);
#undef SZ
  return -1;
}

int m3p2i_create(const M3P2IConfig* cfg, int device, m3p2i_handle* out) {
  if (!cfg || !out) return fail(M3P2I_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->num_samples < 1 || cfg->horizon < 1 || cfg->horizon > M3P2I_MAX_HORIZON)
    return fail(M3P2I_ERR_ARG, "num_samples >= 1 and 1 <= horizon <= 64 required");
  if ((cfg->env_type == M3P2I_ENV_POINT && cfg->nu != 2) || (cfg->env_type == M3P2I_ENV_PANDA && cfg->nu != 9) ||
      (cfg->env_type != M3P2I_ENV_POINT && cfg->env_type != M3P2I_ENV_PANDA))
    return fail(M3P2I_ERR_ARG, "env_type/nu mismatch: point_env has nu=2, panda_env has nu=9");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(M3P2I_ERR_NO_DEVICE, "no CUDA device: libm3p2i_b200 has no CPU path");
  }
  if (device < 0 || device >= ndev) return fail(M3P2I_ERR_ARG, "bad device index");
  CK(cudaSetDevice(device));
  H* h = new H();
  h->cfg = *cfg;
  M3P2IConfig& c = h->cfg;
  if (c.num_samples_global <= 0) c.num_samples_global = c.num_samples;
  if (c.solver_passes <= 0) c.solver_passes = 2;
  if (c.substeps <= 0) c.substeps = 2;
  if (c.sample_offset < 0 || c.sample_offset + c.num_samples > c.num_samples_global) {
    delete h;
    return fail(M3P2I_ERR_ARG, "sample_offset + num_samples exceeds num_samples_global");
  }
  h->device = device;
  if (cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) h->sm_count = 0;
  h->ndof = ndof_of(h); h->nf = nf_of(h);
  h->task = c.env_type == M3P2I_ENV_POINT ? M3P2I_TASK_NAVIGATION : M3P2I_TASK_REACH;
  const size_t K = c.num_samples, T = c.horizon, nu = c.nu, Kg = c.num_samples_global, TN = T * nu;
  cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  h->stream = h->own_stream;
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e == cudaSuccess) e = cudaEventCreate(&h->evr);
  if (e == cudaSuccess) e = h->seq.alloc(SEQ_COUNT * TN);
  if (e == cudaSuccess) e = h->base.alloc(64);
  if (e == cudaSuccess) e = h->actions.alloc(K * TN);
  if (e == cudaSuccess) e = h->states.alloc(K * T);
  if (e == cudaSuccess) e = h->cost_h.alloc(K * T);
  if (e == cudaSuccess) e = h->J.alloc(K);
  if (e == cudaSuccess) e = h->cost_sum.alloc(K);
  if (e == cudaSuccess) e = h->J_global.alloc(Kg);
  if (e == cudaSuccess) e = h->weights.alloc(3 * Kg);
  if (e == cudaSuccess) e = h->partials.alloc(7 * TN + 1);
  if (e == cudaSuccess) e = h->cost_total.alloc(K);
  if (e == cudaSuccess) e = h->result.alloc(2 * TN);
  if (e == cudaSuccess) e = h->refs.alloc(T);
  if (e == cudaSuccess) e = h->ref_flags.alloc(16);  // [0,1] producer progress, [2] fused-update CTA counter,
                                                      // [3] peer-push ticket, [4] peer-wait error,
                                                      // [8] k_stats CTA counter, [9..11] beta iterations per set
  if (e == cudaSuccess) e = h->stats.alloc(1);
  if (e == cudaSuccess) e = h->info.alloc(1);
  if (e == cudaSuccess && c.env_type == M3P2I_ENV_PANDA) e = h->near_list.alloc(K);
  if (e == cudaSuccess && c.env_type == M3P2I_ENV_PANDA) e = h->near_count.alloc(2);
  if (e == cudaSuccess && c.env_type == M3P2I_ENV_PANDA) e = h->far_info.alloc(8);
  if (e == cudaSuccess && c.env_type == M3P2I_ENV_PANDA && c.substeps > 0)
    e = h->far_dump.alloc(K * (size_t)far_boundaries((int)T, c.substeps) * 18);
  if (e == cudaSuccess) {
    Stats s;
    memset(&s, 0, sizeof(s));
    s.beta = 1.0;
    for (int d = 0; d < kMaxNu; ++d) { s.sigma[d] = c.sigma[d]; s.cov[d] = c.sigma[d] * c.sigma[d]; }   // mppi.py:175-176
    e = cudaMemcpy(h->stats.p, &s, sizeof(s), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);   // pageable H2D may still be in flight
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("m3p2i_create: ") + cudaGetErrorString(e);
    m3p2i_destroy(h);
    return fail(M3P2I_ERR_CUDA, msg);
  }
  memset(&h->last_info, 0, sizeof(h->last_info));
  int rc = ensure_pin(h, 4096);
  if (rc) { m3p2i_destroy(h); return rc; }
  if (cudaMallocHost(&h->pin_base, sizeof(float) * 64) != cudaSuccess) {
    m3p2i_destroy(h);
    return fail(M3P2I_ERR_CUDA, "m3p2i_create: cudaMallocHost failed");
  }
  memset(h->pin_base, 0, sizeof(float) * 64);
  {
    // result rows + info as the kernels write them, readable by the host right after the stream sync
    const size_t bytes = sizeof(float) * 2 * (size_t)cfg->horizon * cfg->nu + sizeof(M3P2ICommandInfo);
    void* dev = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->mirror), bytes, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&dev, h->mirror, 0) != cudaSuccess) {
      cudaGetLastError();
      if (h->mirror) cudaFreeHost(h->mirror);
      h->mirror = nullptr;   // fall back to D2H copies
    } else {
      memset(h->mirror, 0, bytes);
      h->mirror_dev = static_cast<float*>(dev);
    }
  }
  *out = h;
  return 0;
}

void m3p2i_destroy(m3p2i_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (int r = 0; r < kMaxPeers; ++r)
    if (h->peer_ipc[r] && h->peer_box[r]) cudaIpcCloseMemHandle(h->peer_box[r]);
  if (h->mailbox) cudaFree(h->mailbox);
  h->noise.release(); h->noise_row0.release(); h->seq.release(); h->actions_in.release(); h->base.release();
  h->env.release(); h->vel_target.release(); h->actions.release(); h->cost_h.release(); h->J.release();
  h->cost_sum.release(); h->J_global.release(); h->weights.release(); h->partials.release(); h->filt.release();
  h->cost_total.release(); h->result.release(); h->links.release(); h->scratch.release(); h->states.release();
  h->refs.release(); h->ref_flags.release(); h->stats.release(); h->info.release();
  h->near_list.release(); h->near_count.release(); h->far_info.release(); h->far_dump.release();
  if (h->pin) cudaFreeHost(h->pin);
  if (h->pin_base) cudaFreeHost(h->pin_base);
  if (h->mirror) cudaFreeHost(h->mirror);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->evr) cudaEventDestroy(h->evr);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int m3p2i_set_stream(m3p2i_handle h, void* cuda_stream) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  CK(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return 0;
}

int m3p2i_set_scene_point(m3p2i_handle h, const M3P2IPointScene* s) {
  if (!h || !s) return fail(M3P2I_ERR_ARG, "null argument");
  if (h->cfg.env_type != M3P2I_ENV_POINT) return fail(M3P2I_ERR_ARG, "handle was created for panda_env");
  if (s->n_static < 0 || s->n_static > M3P2I_MAX_STATIC || s->n_actors < 1 || s->n_actors > 32)
    return fail(M3P2I_ERR_ARG, "n_static <= 8 and 1 <= n_actors <= 32 required");
  if (s->box.actor < 0 || s->box.actor >= s->n_actors || s->dyn_obs.actor < 0 || s->dyn_obs.actor >= s->n_actors)
    return fail(M3P2I_ERR_ARG, "box / dyn_obs actor rows out of range");
  h->ps_in = *s;
  h->n_actors = s->n_actors;
  build_point_params(h->ps_in, h->pp);
  h->have_scene = true;
  return 0;
}

int m3p2i_set_scene_panda(m3p2i_handle h, const M3P2IPandaScene* s) {
  if (!h || !s) return fail(M3P2I_ERR_ARG, "null argument");
  if (h->cfg.env_type != M3P2I_ENV_PANDA) return fail(M3P2I_ERR_ARG, "handle was created for point_env");
  if (s->n_static < 0 || s->n_static > M3P2I_MAX_STATIC || s->n_actors < 1 || s->n_actors > 32)
    return fail(M3P2I_ERR_ARG, "n_static <= 8 and 1 <= n_actors <= 32 required");
  if (s->cube_a.actor < 0 || s->cube_a.actor >= s->n_actors || s->cube_b.actor < 0 || s->cube_b.actor >= s->n_actors)
    return fail(M3P2I_ERR_ARG, "cube actor rows out of range");
  h->qs_in = *s;
  h->n_actors = s->n_actors;
  build_panda_params(h->qs_in, h->qp);
  h->have_scene = true;
  return 0;
}

int m3p2i_set_state(m3p2i_handle h, const float* dof, const float* root) {
  if (!h || !dof || !root) return fail(M3P2I_ERR_ARG, "null argument");
  if (!h->have_scene) return fail(M3P2I_ERR_STATE, "set_scene_* has not been called");
  h->root0.assign(root, root + 13 * (size_t)h->n_actors);
  // movable bodies that the integrator keeps fixed (e.g. the floating dyn-obs plate) follow the real state
  if (h->cfg.env_type == M3P2I_ENV_PANDA) {
    for (int k = 0; k < h->qs_in.n_static; ++k) {
      M3P2IBox& b = h->qs_in.statics[k];
      if (b.actor >= 0 && b.actor < h->n_actors) {
        memcpy(b.pos, root + 13 * b.actor, 12); memcpy(b.quat, root + 13 * b.actor + 3, 16);
        h->qp.st[k] = make_static3(b);
      }
    }
  } else {
    for (int k = 0; k < h->ps_in.n_static; ++k) {
      M3P2IBox& b = h->ps_in.statics[k];
      if (b.actor >= 0 && b.actor < h->n_actors) {
        memcpy(b.pos, root + 13 * b.actor, 12); memcpy(b.quat, root + 13 * b.actor + 3, 16);
        h->pp.st[k] = make_static2(b);
      }
    }
  }
  // (the staging slot is read on the host when a launch is prepared, or uploaded synchronously by the sim facade)
  pack_env(h, dof, root, h->pin_base);
  h->base_dirty = true;
  h->have_state = true;
  h->env_live = false;
  return 0;
}

int m3p2i_set_objective(m3p2i_handle h, int task, const float* goal, int goal_len, int gripper) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  if (goal_len < 0 || goal_len > 7 || (goal_len && !goal)) return fail(M3P2I_ERR_ARG, "goal_len in [0,7] required");
  if (task < 0 || task > M3P2I_TASK_PLACE) return fail(M3P2I_ERR_ARG, "unknown task id");
  const bool point_task = task <= M3P2I_TASK_PUSH_PULL;
  if (point_task != (h->cfg.env_type == M3P2I_ENV_POINT)) return fail(M3P2I_ERR_ARG, "task does not belong to this env_type");
  if (gripper < 0 || gripper > M3P2I_GRIPPER_CLOSE) return fail(M3P2I_ERR_ARG, "unknown gripper command");
  h->task = task; h->gripper = gripper;
  memset(h->goal, 0, sizeof(h->goal));
  for (int i = 0; i < goal_len; ++i) h->goal[i] = goal[i];
  return 0;
}

int m3p2i_set_noise_table(m3p2i_handle h, const float* delta) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  if (!delta) { h->have_noise = false; return 0; }
  const size_t K = h->cfg.num_samples, TN = (size_t)h->cfg.horizon * h->cfg.nu;
  CK(cudaSetDevice(h->device));
  CK(h->noise.alloc(K * TN));
  CK(h->scratch.alloc(K * TN));
  CK(cudaMemcpyAsync(h->scratch.p, delta, sizeof(float) * K * TN, cudaMemcpyHostToDevice, h->stream));
  launch_transpose(h->scratch.p, h->noise.p, (int)K, (int)TN, h->stream);  // [K][T*nu] -> [T*nu][K]
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  h->have_noise = true;
  return 0;
}

int m3p2i_set_noise_halton_spline(m3p2i_handle h, int knot_scale, int degree, float smoothing, const uint16_t* perms,
                                  int perm_stride) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  const int K = h->cfg.num_samples, T = h->cfg.horizon, nu = h->cfg.nu;
  if (h->cfg.noise_mode != M3P2I_NOISE_TABLE)
    return fail(M3P2I_ERR_STATE, "the halton-spline table needs noise_mode = M3P2I_NOISE_TABLE (this planner samples in the kernel)");
  if (knot_scale < 1 || degree < 1 || degree > 3) return fail(M3P2I_ERR_ARG, "knot_scale >= 1 and degree 1..3 expected");
  const int m = T / knot_scale, ndims = m * nu;
  if (m <= degree) return fail(M3P2I_ERR_ARG, "horizon / knot_scale must exceed the spline degree (the reference YAMLs: horizon >= 12)");
  if (m > 16) return fail(M3P2I_ERR_ARG, "at most 16 knot points per spline (horizon / knot_scale)");
  if (!(smoothing > 0.0f)) return fail(M3P2I_ERR_ARG, "smoothing factor must be positive");
  std::vector<int> bases(ndims);
  {
    int found = 0;
    for (int c = 2; found < ndims; ++c) {
      bool prime = true;
      for (int j = 2; j * j <= c; ++j) if (c % j == 0) { prime = false; break; }
      if (prime) bases[found++] = c;
    }
  }
  if (perms) {
    if (perm_stride < bases[ndims - 1]) return fail(M3P2I_ERR_ARG, "perm_stride is smaller than the largest base");
    for (int d = 0; d < ndims; ++d) {   // every row must be a permutation of its base's digits
      std::vector<char> seen(bases[d], 0);
      for (int q = 0; q < bases[d]; ++q) {
        const int v = perms[(size_t)d * perm_stride + q];
        if (v >= bases[d] || seen[v]) return fail(M3P2I_ERR_ARG, "perms rows must be permutations of 0 .. base-1");
        seen[v] = 1;
      }
    }
  }
  CK(cudaSetDevice(h->device));
  DevBuf<int> dbases;
  DevBuf<unsigned short> dperms;
  CK(dbases.alloc(ndims));
  CK(cudaMemcpyAsync(dbases.p, bases.data(), sizeof(int) * ndims, cudaMemcpyHostToDevice, h->stream));
  if (perms) {
    CK(dperms.alloc((size_t)ndims * perm_stride));
    CK(cudaMemcpyAsync(dperms.p, perms, sizeof(unsigned short) * (size_t)ndims * perm_stride, cudaMemcpyHostToDevice, h->stream));
  }
  CK(h->noise.alloc((size_t)K * T * nu));
  launch_halton_spline(h->noise.p, K, h->cfg.sample_offset, T, nu, m, degree, (double)smoothing, dbases.p,
                       perms ? dperms.p : nullptr, perm_stride, h->stream);
  CK(cudaGetLastError());
  h->have_noise = true;
  if (h->cfg.sample_offset != 0) {
    CK(h->noise_row0.alloc((size_t)T * nu));
    launch_halton_spline(h->noise_row0.p, 1, 0, T, nu, m, degree, (double)smoothing, dbases.p, perms ? dperms.p : nullptr,
                         perm_stride, h->stream);
    CK(cudaGetLastError());
    h->have_row0 = true;
  }
  const cudaError_t e = cudaStreamSynchronize(h->stream);
  dbases.release(); dperms.release();
  CK(e);
  return 0;
}

int m3p2i_set_noise_row0(m3p2i_handle h, const float* row0) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  if (!row0) { h->have_row0 = false; return 0; }
  const size_t TN = (size_t)h->cfg.horizon * h->cfg.nu;
  CK(h->noise_row0.alloc(TN));
  CK(cudaMemcpyAsync(h->noise_row0.p, row0, sizeof(float) * TN, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_row0 = true;
  return 0;
}

int m3p2i_get_noise(m3p2i_handle h, float* out) {
  if (!h || !out) return fail(M3P2I_ERR_ARG, "null argument");
  const size_t K = h->cfg.num_samples, TN = (size_t)h->cfg.horizon * h->cfg.nu;
  CK(h->scratch.alloc(2 * K * TN));
  const float* src;
  if (h->cfg.noise_mode != M3P2I_NOISE_TABLE) {
    RolloutCfg c = make_rcfg(h);
    launch_noise_dump(c, h->scratch.p, h->stream);
    src = h->scratch.p;
  } else {
    if (!h->have_noise) { memset(out, 0, sizeof(float) * K * TN); return 0; }
    src = h->noise.p;
  }
  launch_transpose(src, h->scratch.p + K * TN, (int)TN, (int)K, h->stream);  // [T*nu][K] -> [K][T*nu]
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->scratch.p + K * TN, sizeof(float) * K * TN, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_get_planner_state(m3p2i_handle h, M3P2IPlannerState* out) {
  if (!h || !out) return fail(M3P2I_ERR_ARG, "null argument");
  const size_t TN = (size_t)h->cfg.horizon * h->cfg.nu;
  std::vector<float> tmp(SEQ_COUNT * TN);
  Stats s;
  CK(cudaMemcpyAsync(tmp.data(), h->seq.p, sizeof(float) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&s, h->stats.p, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memset(out, 0, sizeof(*out));
  float* dst[SEQ_COUNT] = {out->mean_action, out->mean_action_1, out->mean_action_2, out->best_traj, out->best_traj_1,
                           out->best_traj_2};
  for (int i = 0; i < SEQ_COUNT; ++i) memcpy(dst[i], tmp.data() + i * TN, sizeof(float) * TN);
  out->beta = s.beta;
  for (int d = 0; d < kMaxNu; ++d) out->cov_action[d] = s.cov[d];
  return 0;
}

int m3p2i_set_planner_state(m3p2i_handle h, const M3P2IPlannerState* in) {
  if (!h || !in) return fail(M3P2I_ERR_ARG, "null argument");
  const size_t TN = (size_t)h->cfg.horizon * h->cfg.nu;
  std::vector<float> tmp(SEQ_COUNT * TN);
  const float* src[SEQ_COUNT] = {in->mean_action, in->mean_action_1, in->mean_action_2, in->best_traj, in->best_traj_1,
                                 in->best_traj_2};
  for (int i = 0; i < SEQ_COUNT; ++i) memcpy(tmp.data() + i * TN, src[i], sizeof(float) * TN);
  CK(cudaMemcpyAsync(h->seq.p, tmp.data(), sizeof(float) * tmp.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(&h->stats.p->beta, &in->beta, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  bool have_cov = false;
  for (int d = 0; d < h->cfg.nu; ++d) have_cov = have_cov || in->cov_action[d] > 0.0f;
  float cs[2 * kMaxNu];
  if (have_cov) {   // a state captured before cov_action existed (all zero) keeps the current variance
    for (int d = 0; d < kMaxNu; ++d) { cs[d] = in->cov_action[d]; cs[kMaxNu + d] = sqrtf(fmaxf(in->cov_action[d], 0.0f)); }
    CK(cudaMemcpyAsync(h->stats.p->cov, cs, sizeof(float) * kMaxNu, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->stats.p->sigma, cs + kMaxNu, sizeof(float) * kMaxNu, cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_set_filter_matrix(m3p2i_handle h, const float* S) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  if (!S) { h->have_filt = false; return 0; }
  const size_t T = h->cfg.horizon;
  CK(h->filt.alloc(T * T));
  CK(cudaMemcpyAsync(h->filt.p, S, sizeof(float) * T * T, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_filt = true;
  return 0;
}

int m3p2i_command(m3p2i_handle h, float* out_action, float* out_cost_total, M3P2ICommandInfo* info) {
  int rc = command_device(h, info != nullptr);
  if (rc) return rc;
  if ((rc = fetch(h, out_action, out_cost_total, nullptr, false))) return rc;
  if ((rc = finish_timing(h))) return rc;
  if (info) *info = h->last_info;
  return 0;
}

int m3p2i_command_resident(m3p2i_handle h, M3P2ICommandInfo* info) {
  int rc = command_device(h, info != nullptr);
  if (rc) return rc;
  if (info) {  // asking for the scalars forces a sync; pass NULL to keep the stream running
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = fetch(h, nullptr, nullptr, nullptr, false))) return rc;
    if ((rc = finish_timing(h))) return rc;
    *info = h->last_info;
  }
  return 0;
}

int m3p2i_fetch_result(m3p2i_handle h, float* out_action, float* out_cost_total) {
  if (!h) return fail(M3P2I_ERR_ARG, "null handle");
  int rc = fetch(h, out_action, out_cost_total, nullptr, false);
  if (rc) return rc;
  return finish_timing(h);
}

int m3p2i_rollout_actions(m3p2i_handle h, const float* actions, float* out_states, float* out_cost_h) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!actions) return fail(M3P2I_ERR_ARG, "null actions");
  const size_t K = h->cfg.num_samples, T = h->cfg.horizon, TN = T * h->cfg.nu;
  CK(h->actions_in.alloc(K * TN));
  CK(h->scratch.alloc(std::max(K * TN, K * T * 4)));
  CK(cudaMemcpyAsync(h->scratch.p, actions, sizeof(float) * K * TN, cudaMemcpyHostToDevice, h->stream));
  launch_transpose(h->scratch.p, h->actions_in.p, (int)K, (int)TN, h->stream);
  int launches = 0;
  if ((rc = run_rollout(h, &launches, h->actions_in.p))) return rc;
  CK(cudaStreamSynchronize(h->stream));
  if (out_states && (rc = m3p2i_read_buffer(h, M3P2I_BUF_STATES, out_states, K * T * 4))) return rc;
  if (out_cost_h && (rc = m3p2i_read_buffer(h, M3P2I_BUF_COST_HORIZON, out_cost_h, K * T))) return rc;
  return 0;
}

int m3p2i_sample_actions(m3p2i_handle h, float* out_actions) {
  if (!h || !out_actions) return fail(M3P2I_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  const size_t K = h->cfg.num_samples, TN = (size_t)h->cfg.horizon * h->cfg.nu;
  if (h->cfg.noise_mode == M3P2I_NOISE_TABLE && !h->have_noise)
    return fail(M3P2I_ERR_STATE, "table noise mode: m3p2i_set_noise_table has not been called");
  CK(h->scratch.alloc(2 * K * TN));
  RolloutCfg c = make_rcfg(h);
  RolloutBufs b = make_rbufs(h);
  c.preshifted = 1;
  const float us = c.u_scale;
  c.u_scale = 1.0f;       // perturbed_action is stored before u_scale is applied (mppi.py:297,411)
  c.null_action = 0;      // the null action is applied inside the T-loop of the caller (mppi.py:300-302)
  launch_sample_actions(h->cfg.env_type, c, b, h->seq.p, h->scratch.p, h->stream);
  launch_transpose(h->scratch.p, h->scratch.p + K * TN, (int)TN, (int)K, h->stream);   // -> [K][T*nu]
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_actions, h->scratch.p + K * TN, sizeof(float) * K * TN, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  (void)us;
  return 0;
}

int m3p2i_update_only(m3p2i_handle h, const float* cost_h, const float* actions, float* out_mean, M3P2ICommandInfo* info) {
  if (!h || !cost_h || !actions) return fail(M3P2I_ERR_ARG, "null argument");
  if (h->cfg.num_samples != h->cfg.num_samples_global) return fail(M3P2I_ERR_STATE, "update_only needs an unsharded handle");
  const size_t K = h->cfg.num_samples, T = h->cfg.horizon, TN = T * h->cfg.nu;
  CK(h->scratch.alloc(K * TN));
  CK(cudaMemcpyAsync(h->scratch.p, actions, sizeof(float) * K * TN, cudaMemcpyHostToDevice, h->stream));
  launch_transpose(h->scratch.p, h->actions.p, (int)K, (int)TN, h->stream);
  CK(cudaMemcpyAsync(h->scratch.p, cost_h, sizeof(float) * K * T, cudaMemcpyHostToDevice, h->stream));
  launch_transpose(h->scratch.p, h->cost_h.p, (int)K, (int)T, h->stream);
  int launches = 0, rc;
  h->timed = true; h->have_evr = false;
  CK(cudaEventRecord(h->ev0, h->stream));
  launch_discount(h->cost_h.p, h->J.p, h->cost_sum.p, (int)K, (int)T, h->cfg.gamma, h->stream, &launches);
  if ((rc = gather_J(h))) return rc;
  if ((rc = run_update(h, 0, &launches))) return rc;
  if ((rc = run_finish(h, 0, &launches))) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  h->last_info.launches = launches;
  if ((rc = fetch(h, out_mean, nullptr, nullptr, true))) return rc;
  if ((rc = finish_timing(h))) return rc;
  if (info) *info = h->last_info;
  return 0;
}

int m3p2i_get_buffer(m3p2i_handle h, int which, void** dev_ptr, size_t* bytes) {
  if (!h || !dev_ptr || !bytes) return fail(M3P2I_ERR_ARG, "null argument");
  const size_t K = h->cfg.num_samples, T = h->cfg.horizon, nu = h->cfg.nu, Kg = h->cfg.num_samples_global;
  switch (which) {
    case M3P2I_BUF_ACTIONS: *dev_ptr = h->actions.p; *bytes = 4 * K * T * nu; break;
    case M3P2I_BUF_STATES: *dev_ptr = h->states.p; *bytes = 16 * K * T; break;
    case M3P2I_BUF_COST_HORIZON: *dev_ptr = h->cost_h.p; *bytes = 4 * K * T; break;
    case M3P2I_BUF_COST_DISC: *dev_ptr = const_cast<float*>(j_global_ptr(h)); *bytes = 4 * Kg; break;
    case M3P2I_BUF_COST_SUM: *dev_ptr = h->cost_sum.p; *bytes = 4 * K; break;
    case M3P2I_BUF_WEIGHTS: *dev_ptr = h->weights.p; *bytes = 12 * Kg; break;
    case M3P2I_BUF_NOISE: *dev_ptr = h->have_noise ? h->noise.p : nullptr; *bytes = h->have_noise ? 4 * K * T * nu : 0; break;
    default: return fail(M3P2I_ERR_ARG, "unknown buffer id");
  }
  return 0;
}

int m3p2i_read_buffer(m3p2i_handle h, int which, float* out, size_t count) {
  if (!h || !out) return fail(M3P2I_ERR_ARG, "null argument");
  const size_t K = h->cfg.num_samples, T = h->cfg.horizon, nu = h->cfg.nu, Kg = h->cfg.num_samples_global;
  const float* src = nullptr;
  size_t n = 0;
  int rows = 0, cols = 0;  // device layout [rows][cols] -> host [cols][rows]
  switch (which) {
    case M3P2I_BUF_ACTIONS: src = h->actions.p; n = K * T * nu; rows = (int)(T * nu); cols = (int)K; break;
    case M3P2I_BUF_COST_HORIZON: src = h->cost_h.p; n = K * T; rows = (int)T; cols = (int)K; break;
    case M3P2I_BUF_STATES: src = reinterpret_cast<const float*>(h->states.p); n = K * T * 4; break;
    case M3P2I_BUF_COST_DISC: src = j_global_ptr(h); n = Kg; break;
    case M3P2I_BUF_COST_SUM: src = h->cost_sum.p; n = K; break;
    case M3P2I_BUF_WEIGHTS: src = h->weights.p; n = 3 * Kg; break;
    case M3P2I_BUF_NOISE: return m3p2i_get_noise(h, out);
    default: return fail(M3P2I_ERR_ARG, "unknown buffer id");
  }
  if (count < n) return fail(M3P2I_ERR_ARG, "output array too small");
  if (rows) {
    CK(h->scratch.alloc(n));
    launch_transpose(src, h->scratch.p, rows, cols, h->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->scratch.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (which == M3P2I_BUF_ACTIONS && h->cfg.u_scale != 1.0f)  // self.actions /= u_scale (mppi.py:420)
      for (size_t i = 0; i < n; ++i) out[i] /= h->cfg.u_scale;
    return 0;
  }
  if (which == M3P2I_BUF_STATES) {  // [T][K][4] -> [K][T][4]
    std::vector<float> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), src, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t t = 0; t < T; ++t)
      for (size_t k = 0; k < K; ++k) memcpy(out + (k * T + t) * 4, tmp.data() + (t * K + k) * 4, 16);
    return 0;
  }
  CK(cudaMemcpyAsync(out, src, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_top_trajs(m3p2i_handle h, int n, int32_t* out_idx, float* out_w, float* out_trajs) {
  if (!h || n < 1) return fail(M3P2I_ERR_ARG, "bad argument");
  const int Kg = h->cfg.num_samples_global, T = h->cfg.horizon, K = h->cfg.num_samples, off = h->cfg.sample_offset;
  if (n > Kg) return fail(M3P2I_ERR_ARG, "n exceeds the number of samples (torch.topk would raise)");
  std::vector<float> w(Kg);
  CK(cudaMemcpyAsync(w.data(), h->weights.p, sizeof(float) * Kg, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  // n largest, NaN last, ties broken by the lower index (stable partial sort)
  std::vector<int> idx(Kg);
  for (int i = 0; i < Kg; ++i) idx[i] = i;
  auto better = [&](int a, int b) {
    const bool na = std::isnan(w[a]), nb = std::isnan(w[b]);
    if (na != nb) return nb;
    if (w[a] != w[b]) return w[a] > w[b];
    return a < b;
  };
  std::partial_sort(idx.begin(), idx.begin() + n, idx.end(), better);
  std::vector<float> row((size_t)T * 4);
  for (int j = 0; j < n; ++j) {
    const int g = idx[j], k = g - off;
    if (out_idx) out_idx[j] = g;
    if (out_w) out_w[j] = w[g];
    if (out_trajs) {
      float* o = out_trajs + (size_t)j * T * 2;
      if (k >= 0 && k < K) {
        CK(cudaMemcpy2DAsync(row.data(), 16, reinterpret_cast<const float*>(h->states.p) + (size_t)k * 4,
                             sizeof(float4) * K, 16, T, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int t = 0; t < T; ++t) { o[2 * t] = row[4 * t]; o[2 * t + 1] = row[4 * t + 2]; }
      } else {
        memset(o, 0, sizeof(float) * T * 2);
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- sim facade
int m3p2i_sim_reset(m3p2i_handle h) {
  int rc = check_ready(h);
  if (rc) return rc;
  if ((rc = upload_base(h))) return rc;
  h->env_live = false;
  return materialize(h);
}

int m3p2i_sim_set_velocity_target(m3p2i_handle h, const float* u) {
  if (!h || !u) return fail(M3P2I_ERR_ARG, "null argument");
  int rc = ensure_env(h);
  if (rc) return rc;
  const size_t K = h->cfg.num_samples, nu = h->cfg.nu;
  std::vector<float> t(K * nu);
  for (size_t k = 0; k < K; ++k)
    for (size_t d = 0; d < nu; ++d) t[d * K + k] = u[k * nu + d];
  CK(cudaMemcpyAsync(h->vel_target.p, t.data(), sizeof(float) * K * nu, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_sim_apply_forces(m3p2i_handle h, const float* f_robot, const float* f_box) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (h->cfg.env_type != M3P2I_ENV_POINT) return fail(M3P2I_ERR_ARG, "external forces are only modelled in point_env");
  if ((rc = upload_base(h)) || (rc = materialize(h))) return rc;
  const size_t K = h->cfg.num_samples;
  std::vector<float> t(2 * K);
  if (f_robot) {  // env fields 16,17 = f_robot xy
    for (size_t k = 0; k < K; ++k) { t[k] = f_robot[2 * k]; t[K + k] = f_robot[2 * k + 1]; }
    CK(cudaMemcpyAsync(h->env.p + 16 * K, t.data(), sizeof(float) * 2 * K, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (f_box) {    // env fields 18,19 = f_box xy
    for (size_t k = 0; k < K; ++k) { t[k] = f_box[2 * k]; t[K + k] = f_box[2 * k + 1]; }
    CK(cudaMemcpyAsync(h->env.p + 18 * K, t.data(), sizeof(float) * 2 * K, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int m3p2i_sim_step(m3p2i_handle h) {
  int rc = check_ready(h);
  if (rc) return rc;
  if ((rc = upload_base(h)) || (rc = materialize(h))) return rc;
  RolloutCfg c = make_rcfg(h);
  launch_sim_step(h->cfg.env_type, c, &h->pp, &h->qp, h->env.p, h->vel_target.p, h->stream);
  CK(cudaGetLastError());
  return 0;
}

int m3p2i_sim_cost(m3p2i_handle h, float* out_cost) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!out_cost) return fail(M3P2I_ERR_ARG, "null argument");
  if (needs_refs(h) && h->cfg.num_samples != h->cfg.num_samples_global)
    return fail(M3P2I_ERR_STATE, "reach cost on a sharded sim needs rows 0 and K/2 of the batch");
  if ((rc = upload_base(h)) || (rc = materialize(h))) return rc;
  RolloutCfg c = make_rcfg(h);
  launch_sim_cost(h->cfg.env_type, c, &h->pp, &h->qp, h->env.p, h->cost_total.p, h->stream);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_cost, h->cost_total.p, sizeof(float) * h->cfg.num_samples, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_sim_write(m3p2i_handle h, const float* dof, const float* root) {
  int rc = check_ready(h);
  if (rc) return rc;
  if ((rc = upload_base(h)) || (rc = materialize(h))) return rc;
  const size_t K = h->cfg.num_samples;
  const int nf = h->nf, na = h->n_actors, nd = 2 * h->ndof;
  std::vector<float> env((size_t)nf * K), f(nf), d(nd), r((size_t)13 * na);
  CK(cudaMemcpyAsync(env.data(), h->env.p, sizeof(float) * env.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (size_t k = 0; k < K; ++k) {
    unpack_env(h, env.data() + k, (int)K, d.data(), r.data());
    pack_env(h, dof ? dof + k * nd : d.data(), root ? root + k * 13 * na : r.data(), f.data());
    const int nstate = h->cfg.env_type == M3P2I_ENV_POINT ? 16 : 44;  // forces / contact forces are kept
    for (int i = 0; i < nstate; ++i) env[(size_t)i * K + k] = f[i];
  }
  CK(cudaMemcpyAsync(h->env.p, env.data(), sizeof(float) * env.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_sim_read(m3p2i_handle h, float* dof, float* root, float* link, float* contact) {
  int rc = check_ready(h);
  if (rc) return rc;
  if ((rc = upload_base(h)) || (rc = materialize(h))) return rc;
  const size_t K = h->cfg.num_samples;
  const int nf = h->nf, na = h->n_actors, nd = 2 * h->ndof;
  std::vector<float> env((size_t)nf * K);
  CK(cudaMemcpyAsync(env.data(), h->env.p, sizeof(float) * env.size(), cudaMemcpyDeviceToHost, h->stream));
  std::vector<float> links;
  if (link && h->cfg.env_type == M3P2I_ENV_PANDA) {
    CK(h->links.alloc(K * 39));
    launch_sim_links(&h->qp, h->env.p, h->links.p, (int)K, h->stream);
    CK(cudaGetLastError());
    links.resize(K * 39);
    CK(cudaMemcpyAsync(links.data(), h->links.p, sizeof(float) * K * 39, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  for (size_t k = 0; k < K; ++k) {
    unpack_env(h, env.data() + k, (int)K, dof ? dof + k * nd : nullptr, root ? root + k * 13 * na : nullptr);
    auto F = [&](int i) { return env[(size_t)i * K + k]; };
    if (h->cfg.env_type == M3P2I_ENV_POINT) {
      if (link) {
        float* l = link + k * 13;
        memset(l, 0, sizeof(float) * 13);
        l[0] = F(0); l[1] = F(2); l[2] = 0.05f; l[6] = 1.0f; l[7] = F(1); l[8] = F(3);
      }
      if (contact) { contact[3 * k] = F(20); contact[3 * k + 1] = F(21); contact[3 * k + 2] = 0.0f; }
    } else {
      if (link) memcpy(link + k * 39, links.data() + k * 39, sizeof(float) * 39);
      if (contact) for (int i = 0; i < 9; ++i) contact[9 * k + i] = F(44 + i);
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- multi-GPU
int m3p2i_partials_len(m3p2i_handle h) { return h ? 7 * h->cfg.horizon * h->cfg.nu + 1 : -1; }

int m3p2i_phase_rollout(m3p2i_handle h, float* out_J_local) {
  int rc = check_ready(h);
  if (rc) return rc;
  int launches = 0;
  h->timed = true; h->have_evr = false;
  CK(cudaEventRecord(h->ev0, h->stream));
  if ((rc = run_rollout(h, &launches, nullptr))) return rc;
  h->last_info.launches = launches;
  if (out_J_local)
    CK(cudaMemcpyAsync(out_J_local, h->J.p, sizeof(float) * h->cfg.num_samples, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_phase_partials(m3p2i_handle h, const float* J_global, float* out_partials) {
  if (!h || !J_global || !out_partials) return fail(M3P2I_ERR_ARG, "null argument");
  const size_t n = 7 * (size_t)h->cfg.horizon * h->cfg.nu + 1;
  CK(cudaMemcpyAsync(local_is_global(h) ? h->J.p : h->J_global.p, J_global, sizeof(float) * h->cfg.num_samples_global,
                     cudaMemcpyHostToDevice, h->stream));
  int launches = 0, rc;
  if ((rc = run_update(h, 1, &launches))) return rc;
  h->last_info.launches += launches;
  CK(cudaMemcpyAsync(out_partials, h->partials.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int m3p2i_phase_finish(m3p2i_handle h, const float* partials_sum, float* out_action, float* out_cost_total,
                       M3P2ICommandInfo* info) {
  if (!h || !partials_sum) return fail(M3P2I_ERR_ARG, "null argument");
  const size_t n = 7 * (size_t)h->cfg.horizon * h->cfg.nu + 1;
  CK(cudaMemcpyAsync(h->partials.p, partials_sum, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  int launches = 0, rc;
  if ((rc = run_finish(h, 1, &launches))) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  h->last_info.launches += launches;
  if ((rc = fetch(h, out_action, out_cost_total, nullptr, false))) return rc;
  if ((rc = finish_timing(h))) return rc;
  if (info) *info = h->last_info;
  return 0;
}

int m3p2i_comm_unique_id(void* out_id128) {
  if (!out_id128) return fail(M3P2I_ERR_ARG, "null argument");
  std::string why;
  if (!g_nccl.load(why)) return fail(M3P2I_ERR_NCCL, why);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(M3P2I_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
  memcpy(out_id128, &id, 128);
  return 0;
}

int m3p2i_comm_init(m3p2i_handle h, int rank, int nranks, const void* id128) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(M3P2I_ERR_ARG, "bad argument");
  if ((long long)h->cfg.num_samples * nranks != h->cfg.num_samples_global || h->cfg.sample_offset != rank * h->cfg.num_samples)
    return fail(M3P2I_ERR_ARG, "K must be split evenly: num_samples * nranks == num_samples_global, offset = rank * num_samples");
  std::string why;
  if (!g_nccl.load(why)) return fail(M3P2I_ERR_NCCL, why);
  CK(cudaSetDevice(h->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (r != ncclSuccess) return fail(M3P2I_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
  h->rank = rank; h->nranks = nranks;
  h->peer_on = false;   // the NCCL collectives are the exchange from now on
  return 0;
}

// ---------------------------------------------------------------- exchange over NVLink peer memory
// Replaces the two NCCL collectives of a sharded command by stores into peer HBM issued from inside the rollout and
// weighted-sum kernels (common.cuh, "exchange over NVLink peer memory"). Protocol: every rank calls
// m3p2i_peer_export, the 96-byte descriptors are exchanged by the host (any transport), every rank calls
// m3p2i_peer_attach with all of them, and a host barrier follows before the first command.
int m3p2i_peer_export(m3p2i_handle h, M3P2IPeerHandle* out) {
  if (!h || !out) return fail(M3P2I_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  const size_t bytes = 2 * mailbox_layout(h).stride * sizeof(float);
  if (!h->mailbox) {
    CK(cudaMalloc(&h->mailbox, bytes));
    CK(cudaMemset(h->mailbox, 0, bytes));
    CK(cudaDeviceSynchronize());
  }
  memset(out, 0, sizeof(*out));
  cudaIpcMemHandle_t ipc;
  static_assert(sizeof(ipc) == sizeof(out->ipc), "cudaIpcMemHandle_t is 64 bytes");
  // same-process peers (several handles in one process) use the pointer; the IPC handle may be unavailable then
  if (cudaIpcGetMemHandle(&ipc, h->mailbox) == cudaSuccess) memcpy(out->ipc, &ipc, sizeof(ipc));
  else cudaGetLastError();
  out->pid = (int64_t)getpid();
  out->ptr = (uint64_t)reinterpret_cast<uintptr_t>(h->mailbox);
  out->device = h->device;
  out->bytes = (uint64_t)bytes;
  return 0;
}

int m3p2i_peer_attach(m3p2i_handle h, int rank, int nranks, const M3P2IPeerHandle* all) {
  if (!h || !all || nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks)
    return fail(M3P2I_ERR_ARG, "bad argument (at most 8 ranks)");
  if (!h->mailbox) return fail(M3P2I_ERR_STATE, "m3p2i_peer_export has not been called on this handle");
  if ((long long)h->cfg.num_samples * nranks != h->cfg.num_samples_global || h->cfg.sample_offset != rank * h->cfg.num_samples)
    return fail(M3P2I_ERR_ARG, "K must be split evenly: num_samples * nranks == num_samples_global, offset = rank * num_samples");
  CK(cudaSetDevice(h->device));
  const uint64_t bytes = 2 * mailbox_layout(h).stride * sizeof(float);
  for (int r = 0; r < nranks; ++r) {
    if (all[r].bytes != bytes) return fail(M3P2I_ERR_ARG, "peer mailbox size differs: K_global, horizon and nu must match on all ranks");
    if (r == rank) { h->peer_box[r] = h->mailbox; continue; }
    if (all[r].pid == (int64_t)getpid()) {
      if (all[r].device != h->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, h->device, all[r].device));
        if (!can) return fail(M3P2I_ERR_STATE, "no peer access between the devices of two handles");
        cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
      }
      h->peer_box[r] = reinterpret_cast<float*>(static_cast<uintptr_t>(all[r].ptr));
    } else {
      cudaIpcMemHandle_t ipc;
      memcpy(&ipc, all[r].ipc, sizeof(ipc));
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
      h->peer_box[r] = static_cast<float*>(p);
      h->peer_ipc[r] = true;
    }
  }
  if (const char* e = getenv("M3P2I_PEER_TIMEOUT_MS")) h->peer_timeout_ms = (unsigned)std::max(1, atoi(e));
  h->rank = rank; h->nranks = nranks; h->peer_on = true; h->peer_epoch = 0;
  return 0;
}

}  // extern "C"
