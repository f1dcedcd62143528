// kernels.cu — the MPPI hot path on sm_100a: fused sample -> rollout -> cost kernel, softmin statistics,
// weighted action sums, mean update. One thread owns one sample trajectory; all HBM traffic is sample-fastest SoA.
#include <cstdlib>

#include "kernels.cuh"
#include "panda_env.cuh"
#include "point_env.cuh"
#include "rollout_common.cuh"

namespace m3 {

constexpr int kRolloutBlock = 32;
constexpr int kTeamBlockMax = 224;  // team kernel: up to 7 warps (14 samples) per CTA, size chosen per launch (team_block)
constexpr int kStatsBlock = 1024;
constexpr int kSumBlock = 256;

template <int ENV> struct EnvOf;
template <> struct EnvOf<M3P2I_ENV_POINT> { using Env = PointEnv; using Params = PointParams; static constexpr int NU = 2; };
template <> struct EnvOf<M3P2I_ENV_PANDA> { using Env = PandaEnv; using Params = PandaParams; static constexpr int NU = 9; };

DEV void env_step(PointEnv& e, const PointParams& P, const float* u, const RolloutCfg& c) {
  point_step(e, P, u, c.dt, c.substeps, c.passes);
}
DEV void env_step(PandaEnv& e, const PandaParams& P, const float* u, const RolloutCfg& c) {
  panda_step(e, P, u, c.dt, c.substeps, c.passes);
}
DEV float env_cost(PointEnv& e, const PointParams&, const RolloutCfg& c, int kg, const PandaRef*) {
  return point_cost(e, c, kg);
}
DEV float env_cost(PandaEnv& e, const PandaParams& P, const RolloutCfg& c, int kg, const PandaRef* ref) {
  PandaRef r;
  if (ref) r = *ref;
  else { r.cube0[0] = e.cube[0].p.x; r.cube0[1] = e.cube[0].p.y; r.cube0[2] = e.cube[0].p.z; r.sel_axis = sel_axis_of(e.cube[0]); }
  return panda_cost(e, P, c, kg, r);
}

// ------------------------------------------------------------------ fused rollout
// noise -> perturbation -> clamp -> T x (dynamics step, task cost, discounted accumulate) with per-step stores of
// the action planes, the float4 state row and the cost (mppi.py:275-332 with reactive_tamp.py:63-73 inlined).
// producer for the thread-per-sample kernel: thread 0 replays global row 0, thread 1 global row Kg/2
template <typename Params>
DEV void produce_refs(const RolloutCfg&, const Params&, const RolloutBufs&, int) {}
template <>
DEV void produce_refs<PandaParams>(const RolloutCfg& c, const PandaParams& P, const RolloutBufs& b, int which) {
  if (which > (c.multi_modal ? 1 : 0)) return;
  const int kg = which == 0 ? 0 : c.Kg / 2;
  const int kl = (kg >= c.offset && kg < c.offset + c.K) ? kg - c.offset : -1;
  PandaEnv e;
  if (c.env_live && kl >= 0) e.load(b.env, c.K, kl);
  else e.load(c.base_env, 1, 0);
  float u[9];
  for (int t = 0; t < c.T; ++t) {
    sample_action<9>(c, b, kg, kl, t, u);
    panda_step(e, P, u, c.dt, c.substeps, c.passes);
    ref_publish(b, which, t, c.epoch, e.cube[0], !c.multi_modal);
  }
}

template <int ENV>
__global__ void __launch_bounds__(kRolloutBlock)
k_rollout(const __grid_constant__ RolloutCfg c, const __grid_constant__ typename EnvOf<ENV>::Params P, const RolloutBufs b) {
  using Env = typename EnvOf<ENV>::Env;
  constexpr int NU = EnvOf<ENV>::NU;
  pdl_launch_dependents();   // k_stats may be set up while this grid runs; it waits for its completion (pdl_wait)
  const bool use_refs = ENV == M3P2I_ENV_PANDA && b.refs != nullptr;
  // after k_rollout_far (panda_far.cuh): only the listed samples
  const bool listed = ENV == M3P2I_ENV_PANDA && b.near_list != nullptr;
  int count = c.K;
  if (listed) count = __ldcg(b.near_count);
  if (listed && count < c.near_thread_min) return;   // few samples left: the team kernel launched before this one took them
  if (use_refs && blockIdx.x == 0) {
    // (rows that stayed in the far field were published by k_rollout_far)
    if (ENV == M3P2I_ENV_PANDA && count > 0 && !(listed && __ldcg(b.far_info + 2))) produce_refs(c, P, b, threadIdx.x);
    return;
  }
  const int kraw = (blockIdx.x - (use_refs ? 1 : 0)) * blockDim.x + threadIdx.x;
  if (kraw >= count) return;
  const int k = listed ? (b.near_list[kraw] & ((1 << kFarRowBits) - 1)) : kraw;   // (always from iteration 0 here)
  const int K = c.K, kg = c.offset + k;
  Env e;
  if (c.env_live) e.load(b.env, K, k);
  else e.load(c.base_env, 1, 0);
  float run = 0.0f, J = 0.0f, g = 1.0f;
  float u[NU];
  for (int t = 0; t < c.T; ++t) {
    sample_action<NU>(c, b, kg, k, t, u);
    env_step(e, P, u, c);
    PandaRef ref;
    if (use_refs) ref = ref_wait(b, t, c.epoch, c.multi_modal != 0);
    const float cost = env_cost(e, P, c, kg, use_refs ? &ref : nullptr);
    run += cost;
    J += g * cost;
    g *= c.gamma;
#pragma unroll
    for (int d = 0; d < NU; ++d) b.actions[(size_t)(t * NU + d) * K + k] = u[d];
    b.states[(size_t)t * K + k] = e.state_row();
    b.cost_h[(size_t)t * K + k] = cost;
  }
  b.cost_sum[k] = run;
  b.J[k] = J;
  if (b.peer.n) {
#ifndef M3_EMU
    push_J_store(b.peer, c.offset, k, J);
    const unsigned m = __activemask();   // the lanes of this warp that own a sample
    __syncwarp(m);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) push_J_commit(b.peer, K, (unsigned)__popc(m));
#endif
  }
  if (c.store_env) {
    e.store(b.env, K, k);
#pragma unroll
    for (int d = 0; d < NU; ++d) b.vel_target[(size_t)d * K + k] = u[d];
  }
}

}  // namespace m3

#include "panda_team.cuh"
#include "panda_far.cuh"
#include "halton_spline.cuh"

namespace m3 {

// Lane-cooperative variant for panda_env: 16 / CPL lanes per sample (panda_team.cuh), 2 * CPL samples per warp.
// CTA 0 is the producer of the batch rows read by the reach cost when b.refs is set.
// One CTA per SM (the contact records and accumulators take most of the shared memory): the kernel may use the whole
// register file (no spills); larger grids run in waves.
template <int CPL>
__global__ void __launch_bounds__(kTeamBlockMax, 1)
k_rollout_team(const __grid_constant__ RolloutCfg c, const __grid_constant__ PandaParams P, const RolloutBufs b) {
  pdl_launch_dependents();   // k_stats may be set up while this grid runs; it waits for its completion (pdl_wait)
  if (b.near_list) pdl_wait();   // behind k_rollout_far: its near list, dumps and costs must be complete and visible
  team_kernel_body<CPL>(c, P, b);
}

// Far-field rollouts (panda_far.cuh): two samples per warp; with b.refs one more warp replays rows 0 and Kg/2 of the
// batch. Samples that stay in the far field for the whole horizon are finished here; the others are listed for the
// rollout kernel launched next (deterministic order inside a CTA, CTAs in the order of their atomic reservation).
__global__ void __launch_bounds__(kFarBlockMax, 1)
k_rollout_far(const __grid_constant__ RolloutCfg c, const __grid_constant__ PandaParams P, const RolloutBufs b) {
  pdl_launch_dependents();   // the rollout kernel over the near list is set up while this grid runs (it waits in pdl_wait)
  M3_DYNAMIC_SMEM(float4, far_smem);
  __shared__ int s_prod_bad, s_near[kFarBlockMax / 32 + 1], s_base;
  const bool use_refs = b.refs != nullptr;
  const int warps = blockDim.x / 32, sample_warps = warps - (use_refs ? 1 : 0);
  const int w = threadIdx.x / 32, lane = threadIdx.x & 31, team = lane / kFarLanes;
  const bool prod = use_refs && w == sample_warps;
  const int cta_first = blockIdx.x * sample_warps * kFarPerWarp;
  const int kraw = cta_first + w * kFarPerWarp + team;
  const bool valid = !prod && kraw < c.K;
  int k = kraw < c.K ? kraw : c.K - 1, kg = c.offset + k;
  if (prod) {
    kg = (team == 1 && c.multi_modal) ? c.Kg / 2 : 0;
    k = (kg >= c.offset && kg < c.offset + c.K) ? kg - c.offset : -1;
  }
  if (threadIdx.x == 0) {
    s_prod_bad = 0;
    if (blockIdx.x == 0) *b.near_count_next = 0;
  }
  __syncthreads();
  float run = 0.0f, J = 0.0f;
  int k0 = 0, bd = 0;   // support box of the lane's cube; hand-over boundary of a sample that leaves the far field
  bool ok = far_base_asleep(c, P, k0);   // warp-uniform
  if (blockIdx.x == 0 && w == 0 && (lane == 0 || lane == 8)) b.far_info[lane >> 3] = k0;
  if (ok) ok = far_team_eval(c, P, b, reinterpret_cast<float*>(far_smem) + (size_t)w * kFarPerWarp * far_sample_floats(c.T, c.substeps),
                             k, kg, valid, run, J, bd);
  if (prod && !ok) s_prod_bad = 1;
  __syncthreads();
  if (s_prod_bad) { ok = false; bd = 0; }
  if (use_refs && blockIdx.x == 0) {
    // rows 0 / Kg/2 of the batch stayed in the far field: their cubeA is the start state's at every step. Publish them
    // for the samples of the near list (the rollout kernel's producer then has nothing to replay: far_info[2]).
    if (threadIdx.x == 0) b.far_info[2] = s_prod_bad ? 0 : 1;
    if (!s_prod_bad) {
      TeamEnv e0;
      e0.load(c.base_env, 1, 0, 0);
      const int axis = sel_axis_of(e0.cu);
      for (int t = threadIdx.x; t < c.T; t += blockDim.x) {
        PandaRef* r = b.refs + t;
        r->cube0[0] = e0.cu.p.x; r->cube0[1] = e0.cu.p.y; r->cube0[2] = e0.cu.p.z; r->sel_axis = axis;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        *(volatile unsigned*)(b.ref_flags + 0) = c.epoch + (unsigned)c.T;
        *(volatile unsigned*)(b.ref_flags + 1) = c.epoch + (unsigned)c.T;
      }
    }
  }
  // near list: rank inside the CTA from warp ballots, one reservation per CTA
  const bool writer = valid && (lane & (kFarLanes - 1)) == 0;
  const unsigned nb = __ballot_sync(0xffffffffu, writer && !ok);
  if (lane == 0) s_near[w] = __popc(nb);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < warps; ++i) { const int n = s_near[i]; s_near[i] = tot; tot += n; }
    s_near[warps] = tot;
    s_base = tot ? atomicAdd(b.near_count, tot) : 0;
  }
  __syncthreads();
  if (writer && !ok) b.near_list[s_base + s_near[w] + __popc(nb & ((1u << lane) - 1u))] = k | (bd << kFarRowBits);
  if (writer && ok) {
    b.cost_sum[k] = run;
    b.J[k] = J;
    if (b.peer.n) push_J_store(b.peer, c.offset, k, J);
  }
  if (b.peer.n) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int here = min(sample_warps * kFarPerWarp, c.K - cta_first) - s_near[warps];
      if (here > 0) push_J_commit(b.peer, c.K, (unsigned)here);
    }
  }
}

// Threads per CTA of the rollout kernel. The kernel is latency-bound (one serial chain per sample), so small CTAs
// that spread the samples over all 148 SMs win; M3P2I_ROLLOUT_BLOCK overrides the choice (8..32) for experiments.
static int rollout_block(int K) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("M3P2I_ROLLOUT_BLOCK");
    forced = e ? atoi(e) : 0;
    if (forced != 0 && (forced < 1 || forced > kRolloutBlock)) forced = 0;
  }
  if (forced) return forced;
  return K <= 148 * 16 * 4 ? 16 : kRolloutBlock;
}

// CTA size of the team kernel (one CTA per SM, up to 7 warps). The rollout runs in waves of equally long CTAs, so its
// duration is set by the SM that received the most warps: pick the warps-per-CTA w in 2..7 that minimises
// ceil(#CTAs / #SMs) * w, preferring large CTAs (their warps are re-aligned every sub-step and share
// instruction-cache lines). K = 4096 with 8-lane teams: w = 7 -> 147 CTAs, one per SM.
// Function attributes (dynamic shared memory above 48 KB) are per device: true the first time `what` is asked for on
// the current device of this process (handles of one process may live on different GPUs).
static bool first_use_on_device(int what) {
  static unsigned char done[3][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (done[what][dev]) return false;
  done[what][dev] = 1;
  return true;
}

static int team_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

static int team_block(int K, int extra, int per_warp) {
  const int sms = team_sms();
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("M3P2I_TEAM_BLOCK");
    forced = e ? atoi(e) : 0;
    if (forced % 32 || forced < 32 || forced > kTeamBlockMax) forced = 0;
  }
  if (forced) return forced;
  int best_w = 2, best_load = 1 << 30;
  for (int w = 2; w <= kTeamBlockMax / 32; ++w) {
    const int ctas = (K + per_warp * w - 1) / (per_warp * w) + extra;
    const int load = ((ctas + sms - 1) / sms) * w;
    if (load <= best_load) { best_load = load; best_w = w; }
  }
  return 32 * best_w;
}

// ------------------------------------------------------------------ programmatic dependent launch (host side)
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// shape of the far-field launch: sample warps per CTA (<= 14) that keeps the shared memory within `smem_cap` and spreads
// the CTAs evenly over the SMs; 0 = the horizon does not fit
static int far_sample_warps(int K, int T, int ns, int extra_warp, size_t smem_cap) {
  const size_t per_warp = (size_t)kFarPerWarp * far_sample_floats(T, ns) * sizeof(float);
  int wmax = kFarBlockMax / 32 - 1;
  while (wmax > 0 && per_warp * (wmax + extra_warp) > smem_cap) --wmax;
  if (wmax <= 0) return 0;
  const int sms = team_sms();
  int best_w = 1, best_load = 1 << 30;
  for (int w = 1; w <= wmax; ++w) {
    const int ctas = (K + kFarPerWarp * w - 1) / (kFarPerWarp * w);
    const int load = ((ctas + sms - 1) / sms) * w;
    if (load <= best_load) { best_load = load; best_w = w; }
  }
  return best_w;
}

bool far_rollout_applies(int env_type, const RolloutCfg& c, const RolloutBufs& b, bool need_refs) {
  const char* env = getenv("M3P2I_FAR");   // M3P2I_FAR=0: every sample through the full rollout kernel (A/B tests)
  const bool enabled = env ? atoi(env) != 0 : true;
  if (!enabled || env_type != M3P2I_ENV_PANDA || c.env_live || c.store_env || !b.near_list) return false;
  if (c.K >= (1 << kFarRowBits)) return false;
  // hand-over boundaries are multiples of 8 iterations and must be step boundaries: 1, 2, 4 or 8 sub-steps
  if (c.substeps <= 0 || c.substeps > 8 || (c.substeps & (c.substeps - 1))) return false;
  return far_sample_warps(c.K, c.T, c.substeps, need_refs ? 1 : 0, 200 * 1024) > 0;
}

void launch_rollout(int env_type, const RolloutCfg& c, const PointParams* pp, const PandaParams* qp,
                    const RolloutBufs& b_in, bool need_refs, cudaStream_t st, int* launches) {
  RolloutBufs b = b_in;
  if (far_rollout_applies(env_type, c, b, need_refs)) {
    const int extra_warp = need_refs ? 1 : 0;
    const int w = far_sample_warps(c.K, c.T, c.substeps, extra_warp, 200 * 1024);
    const int fgrid = (c.K + kFarPerWarp * w - 1) / (kFarPerWarp * w);
    const size_t fsmem = (size_t)(w + extra_warp) * kFarPerWarp * far_sample_floats(c.T, c.substeps) * sizeof(float);
    if (first_use_on_device(2)) cudaFuncSetAttribute(k_rollout_far, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_rollout_far<<<fgrid, 32 * (w + extra_warp), fsmem, st>>>(c, *qp, b);
    ++*launches;
  } else {
    b.near_list = nullptr; b.near_count = nullptr; b.near_count_next = nullptr; b.far_info = nullptr; b.far_dump = nullptr;
  }
  const int block = rollout_block(c.K);
  const int extra = need_refs ? 1 : 0;   // CTA 0 = producer of the batch rows every reach cost reads
  const int grid = (c.K + block - 1) / block + extra;
  if (env_type == M3P2I_ENV_PANDA && c.lanes == 1 && b.near_list && c.K > 84 * team_sms()) {
    // K beyond the team kernel's range, but the far-field kernel has usually left only a few samples: launch the team
    // kernel for up to kTeamMax listed samples and the thread-per-sample kernel for more; the count decides on the device
    const int kTeamMax = 84 * team_sms();
    RolloutCfg ct = c;
    ct.lanes = 8; ct.near_team_max = kTeamMax;
    RolloutBufs bt = b;
    // (a K-sample launch shape with K = kTeamMax: CTAs beyond the count leave at once)
    ct.K = c.K;
    const int tb = team_block(kTeamMax, extra, 4);
    const int tgrid = (kTeamMax + tb / 8 - 1) / (tb / 8) + extra;
    const size_t smem = (size_t)7 * 2 * sizeof(float4) * tb + (size_t)(tb / 8) * 2 * kRecStride * sizeof(float4) +
                        (need_refs ? (size_t)(tb / 8) * 2 * c.T * sizeof(float4) : 0);
    if (first_use_on_device(0)) {
      const int max_smem = 7 * 2 * (int)sizeof(float4) * kTeamBlockMax +
                           (kTeamBlockMax / 8) * 2 * (kRecStride + kMaxT) * (int)sizeof(float4);
      cudaFuncSetAttribute(k_rollout_team<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      cudaFuncSetAttribute(k_rollout_team<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    }
    launch_pdl(k_rollout_team<2>, dim3(tgrid), dim3(tb), smem, st, ct, *qp, bt);
    ++*launches;
    RolloutCfg c1 = c;
    c1.near_thread_min = kTeamMax + 1;
    k_rollout<M3P2I_ENV_PANDA><<<grid, block, 0, st>>>(c1, *qp, b);
    ++*launches;
    return;
  }
  if (env_type == M3P2I_ENV_POINT) {
    k_rollout<M3P2I_ENV_POINT><<<grid, block, 0, st>>>(c, *pp, b);
  } else if (c.lanes == 16 || c.lanes == 8) {
    const int per_warp = 32 / c.lanes;
    const int tb = team_block(c.K, extra, per_warp);
    const int teams_per_block = tb / c.lanes;
    const int tgrid = (c.K + teams_per_block - 1) / teams_per_block + extra;
    // dynamic shared memory: the contact accumulators of panda_team.cuh, 7 * CPL float4 per thread
    // + (reach) the deferred cost ingredients, 2 float4 per step and sample
    // + the link / cube contact records of a sub-step, one list of kRecStride float4 per (sample, cube)
    const size_t smem = (size_t)7 * (16 / c.lanes) * sizeof(float4) * tb +
                        (size_t)(tb / c.lanes) * 2 * kRecStride * sizeof(float4) +
                        (need_refs ? (size_t)(tb / c.lanes) * 2 * c.T * sizeof(float4) : 0);
    if (first_use_on_device(0)) {
      // CPL = 2, largest CTA, longest horizon: above the 48 KB default
      const int max_smem = 7 * 2 * (int)sizeof(float4) * kTeamBlockMax +
                           (kTeamBlockMax / 8) * 2 * (kRecStride + kMaxT) * (int)sizeof(float4);
      cudaFuncSetAttribute(k_rollout_team<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      cudaFuncSetAttribute(k_rollout_team<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    }
    if (b.near_list) {
      // programmatic dependent launch behind k_rollout_far: the launch latency hides behind that kernel
      if (c.lanes == 16) launch_pdl(k_rollout_team<1>, dim3(tgrid), dim3(tb), smem, st, c, *qp, b);
      else launch_pdl(k_rollout_team<2>, dim3(tgrid), dim3(tb), smem, st, c, *qp, b);
    } else if (c.lanes == 16) k_rollout_team<1><<<tgrid, tb, smem, st>>>(c, *qp, b);
    else k_rollout_team<2><<<tgrid, tb, smem, st>>>(c, *qp, b);
  } else {
    k_rollout<M3P2I_ENV_PANDA><<<grid, block, 0, st>>>(c, *qp, b);
  }
  ++*launches;
}

// ------------------------------------------------------------------ block reductions (fixed order => reproducible)
DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int BLOCK>
DEV float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < BLOCK / 32) ? sh[threadIdx.x] : 0.0f;
  if (wid == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  r = sh[0];
  return r;
}
// minimum and the FIRST index attaining it
template <int BLOCK>
DEV void block_argmin(float v, int i, float* shv, int* shi, float& vmin, int& imin) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  __syncthreads();
  if (lane == 0) { shv[wid] = v; shi[wid] = i; }
  __syncthreads();
  if (wid == 0) {
    v = (threadIdx.x < BLOCK / 32) ? shv[threadIdx.x] : INFINITY;
    i = (threadIdx.x < BLOCK / 32) ? shi[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (threadIdx.x == 0) { shv[0] = v; shi[0] = i; }
  }
  __syncthreads();
  vmin = shv[0]; imin = shi[0];
}

// ------------------------------------------------------------------ softmin statistics
// +1: eta above the band of mppi.py:446-454 (beta shrinks), -1: below it (beta grows), 0: inside
DEV int eta_adapt(float eta) { return eta > 20.0f ? 1 : (eta < 10.0f ? -1 : 0); }
// _exp_util (mppi.py:430-456) / _multi_modal_exp_util + update_infinite_beta (m3p2i.py:24-64): min-shift, exp,
// normaliser eta, on-the-fly beta search, weights. One CTA; the beta search loops on the device (no host sync).
__global__ void __launch_bounds__(kStatsBlock) k_stats(const UpdateCfg u, const UpdateBufs b) {
  __shared__ float shv[32];
  __shared__ int shi[32];
  __shared__ int is_last;
  const int Kg = u.Kg, half = Kg / 2;
  const float* J = b.J_global;
  pdl_launch_dependents();   // k_wsum may be set up now; it waits for this grid before it reads anything
  pdl_wait();                // the rollout (or whatever precedes on the stream) has completed
  // sharded over peer memory: J_global is the local mailbox; every rank's slice has landed once its flag is up
  Stats* S = b.stats;
  if (b.peer.n) wait_flags(b.peer.jflag_local, b.peer.n, b.peer.epoch, b.peer.timeout_ms, b.peer.error,
                           blockIdx.x == 0 ? &S->peer_wait_ms[0] : nullptr);
  else if (blockIdx.x == 0 && threadIdx.x == 0) { S->peer_wait_ms[0] = 0.0f; S->peer_wait_ms[1] = 0.0f; }
  // multi-modal: the three weight sets (all / first half / second half) are independent searches: one CTA each
  const int s = blockIdx.x;
  int iters = 0;
  // (the weight rows are zero outside the ranges written below: cleared once at allocation, the ranges never change)
  {
    const int lo = s == 2 ? half : 0, n = s == 0 ? Kg : (s == 1 ? half : Kg - half);
    // the set's costs are read three or more times (minimum, one exp-sum per beta of the search, weights): stage them
    // in shared memory on the first pass when they fit (u.stage_J: the launch reserved 4 Kg bytes)
    extern __shared__ float sJ[];
    const bool staged = u.stage_J != 0;
    float v = INFINITY;
    int vi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += kStatsBlock) {
      const float x = J[lo + i];
      if (staged) sJ[i] = x;
      if (x < v) { v = x; vi = i; }
    }
    if (staged) J = sJ - lo;   // J[lo + i] below reads the staged copy (made visible by block_argmin's barriers)
    float jmin; int imin;
    block_argmin<kStatsBlock>(v, vi, shv, shi, jmin, imin);
    if (imin == 0x7fffffff) { imin = 0; jmin = J[lo]; }  // all-NaN / all-inf costs
    double beta = u.multi_modal ? 1.0 : S->beta;  // multi-modal: the search restarts from 1 every call (m3p2i.py:58-60)
    float scale, eta;
    for (;;) {
      scale = (float)(-1.0 / beta);
      float acc = 0.0f;
      for (int i = threadIdx.x; i < n; i += kStatsBlock) acc += expf(scale * (J[lo + i] - jmin));
      eta = block_sum<kStatsBlock>(acc, shv);
      if (!u.multi_modal) break;
      ++iters;
      if (eta > 10.0f) beta = beta * 0.9;
      else if (eta < 3.0f) beta = beta * 1.2;
      else break;  // also taken when eta is NaN (more than 10 samples tie at the minimum), as in the reference
      if (iters > 100000) break;
    }
    const float inv = 1.0f / eta;
    for (int i = threadIdx.x; i < n; i += kStatsBlock)
      b.weights[(size_t)s * Kg + lo + i] = inv * expf(scale * (J[lo + i] - jmin));
    if (threadIdx.x == 0) {
      S->scale[s] = scale; S->inv_eta[s] = inv; S->eta[s] = eta; S->beta_used[s] = (float)beta; S->jmin[s] = jmin;
      S->best_idx[s] = lo + imin;
    }
  }
  if (!u.multi_modal) {
    if (threadIdx.x == 0) {
      S->beta_iters = 0;
      S->weight_push = 0.0f; S->weight_pull = 0.0f;
      if (u.env_type == M3P2I_ENV_PANDA) {  // beta adapts across calls (mppi.py:446-454)
        if (eta_adapt(S->eta[0]) > 0) S->beta = S->beta * 0.9;
        else if (eta_adapt(S->eta[0]) < 0) S->beta = S->beta * 1.2;
      }
    }
    return;
  }
  // the CTA that finishes last adds up the iterations and evaluates get_pull_preference (m3p2i.py:16-22) on the
  // weights of the full set (written by CTA 0)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    b.stats_scratch[1 + s] = (unsigned)iters;
    __threadfence();
    const unsigned done = atomicAdd(b.stats_scratch, 1u);
    is_last = done == gridDim.x - 1;
    if (is_last) *b.stats_scratch = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float a = 0.0f, c = 0.0f;
  for (int i = threadIdx.x; i < half; i += kStatsBlock) a += __ldcg(b.weights + i);
  for (int i = half + threadIdx.x; i < Kg; i += kStatsBlock) c += __ldcg(b.weights + i);
  a = block_sum<kStatsBlock>(a, shv);
  c = block_sum<kStatsBlock>(c, shv);
  if (threadIdx.x == 0) {
    S->weight_push = a; S->weight_pull = c;
    const volatile unsigned* it = b.stats_scratch + 1;
    S->beta_iters = (int)(it[0] + it[1] + it[2]);
  }
}

void launch_stats(const UpdateCfg& u, const UpdateBufs& b, cudaStream_t st, int* launches) {
  // dynamic shared memory: the costs of the largest set, when they fit beside the reduction scratch
  constexpr size_t kStageMax = 200 * 1024;
  if (first_use_on_device(1)) cudaFuncSetAttribute(k_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStageMax);
  UpdateCfg us = u;
  const size_t bytes = sizeof(float) * (size_t)u.Kg;
  us.stage_J = bytes <= kStageMax ? 1 : 0;
  launch_pdl(k_stats, dim3(u.multi_modal ? 3 : 1), dim3(kStatsBlock), us.stage_J ? bytes : 0, st, us, b);
  ++*launches;
}

// ------------------------------------------------------------------ weighted action sums
// CTA j < T*nu reduces plane j of the action buffer over this shard's K samples with the three weight sets
// (mppi.py:498-499, m3p2i.py:80-86) and gathers the best rows (mppi.py:494-496, m3p2i.py:75-78);
// CTA T*nu sums the undiscounted costs (the mean term of mppi.py:325).
DEV void finish_body(const UpdateCfg& u, const UpdateBufs& b, float* smean);

__global__ void __launch_bounds__(kSumBlock) k_wsum(const UpdateCfg u, const UpdateBufs b) {
  extern __shared__ float smean[];  // [T*nu] + [T*T], used by the last CTA when the finish step is fused in
  __shared__ float sh[32];
  __shared__ int is_last;
  const int K = u.K, Kg = u.Kg, half = Kg / 2, TN = u.T * u.nu, j = blockIdx.x;
  pdl_wait();   // k_stats has completed (weights, best indices)
  if (j == TN) {
    float a = 0.0f;
    for (int k = threadIdx.x; k < K; k += kSumBlock) a += b.cost_sum[k];
    a = block_sum<kSumBlock>(a, sh);
    if (threadIdx.x == 0) b.partials[6 * TN] = a;
    if (u.fuse_finish && !b.peer.n) {
      // single rank: this CTA knows the mean already, so cost_total (mppi.py:282-284,325) is written here, in parallel
      // with the weighted sums, instead of by the last CTA
      const float mean_cost = a / (float)Kg;
      for (int k = threadIdx.x; k < K; k += kSumBlock) b.cost_total[k] = b.cost_sum[k] + mean_cost;
    }
  } else {
  const float* plane = b.actions + (size_t)j * K;
  const float* w0 = b.weights + u.offset;
  const float* w1 = b.weights + (size_t)Kg + u.offset;
  const float* w2 = b.weights + 2 * (size_t)Kg + u.offset;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, sq = 0.0f;
  if (((K | u.offset | Kg) & 3) == 0) {
    // 128-bit loads, all of a thread's requests in flight at once (the plane and the weights are L2-resident: the loop
    // is bound by load latency, not bandwidth)
    const float4* p4 = reinterpret_cast<const float4*>(plane);
    const float4* w04 = reinterpret_cast<const float4*>(w0);
    const float4* w14 = reinterpret_cast<const float4*>(w1);
    const float4* w24 = reinterpret_cast<const float4*>(w2);
#pragma unroll 4
    for (int q = threadIdx.x; q < K / 4; q += kSumBlock) {
      const float4 a = p4[q], w = w04[q];
      s0 += w.x * a.x; s0 += w.y * a.y; s0 += w.z * a.z; s0 += w.w * a.w;
      sq += w.x * a.x * a.x; sq += w.y * a.y * a.y; sq += w.z * a.z * a.z; sq += w.w * a.w * a.w;
      if (u.multi_modal) {
        // (half is a multiple of 4 here: a float4 never straddles the mode boundary)
        if (u.offset + 4 * q < half) { const float4 v = w14[q]; s1 += v.x * a.x; s1 += v.y * a.y; s1 += v.z * a.z; s1 += v.w * a.w; }
        else { const float4 v = w24[q]; s2 += v.x * a.x; s2 += v.y * a.y; s2 += v.z * a.z; s2 += v.w * a.w; }
      }
    }
  } else {
  for (int k = threadIdx.x; k < K; k += kSumBlock) {
    const float a = plane[k];
    s0 += w0[k] * a;
    sq += w0[k] * a * a;   // second moment for the covariance update (mppi.py:508-516)
    if (u.multi_modal) {
      if (u.offset + k < half) s1 += w1[k] * a;
      else s2 += w2[k] * a;
    }
  }
  }
  s0 = block_sum<kSumBlock>(s0, sh);
  if (u.multi_modal) { s1 = block_sum<kSumBlock>(s1, sh); s2 = block_sum<kSumBlock>(s2, sh); }
  if (u.update_cov) sq = block_sum<kSumBlock>(sq, sh);
  if (threadIdx.x == 0) {
    b.partials[j] = s0; b.partials[TN + j] = s1; b.partials[2 * TN + j] = s2; b.partials[6 * TN + 1 + j] = sq;
    const int nsets = u.multi_modal ? 3 : 1;
    for (int s = 0; s < 3; ++s) {
      float row = 0.0f;
      if (s < nsets) {
        const int bl = b.stats->best_idx[s] - u.offset;
        if (bl >= 0 && bl < K) row = plane[bl];
      }
      b.partials[(3 + s) * TN + j] = row;
    }
  }
  }
  if (!u.fuse_finish) return;
  // single-GPU path: the CTA that finishes last applies the mean update (k_finish's work) in the same launch
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(b.done_counter, 1u);
    is_last = done == gridDim.x - 1;
    if (is_last) *b.done_counter = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (b.peer.n) {
    // all-reduce over peer memory: this rank's packed partial sums go into box [rank] of every mailbox ...
    const PeerReduce& p = b.peer;
    const int NP = 7 * TN + 1;
    const volatile float* mine = b.partials;
    for (int i = threadIdx.x; i < NP; i += kSumBlock) {
      const float v = mine[i];
      for (int r = 0; r < p.n; ++r) p.part[r][(size_t)p.rank * p.np + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      for (int r = 0; r < p.n; ++r) *(volatile unsigned*)(p.pflag[r] + p.rank) = p.epoch;
    // ... and the n boxes of the own mailbox are added in rank order (the same bits on every rank)
    wait_flags(p.pflag_local, p.n, p.epoch, p.timeout_ms, p.error, &b.stats->peer_wait_ms[1]);
    for (int i = threadIdx.x; i < NP; i += kSumBlock) {
      float a = 0.0f;
      for (int r = 0; r < p.n; ++r) a += __ldcg(p.part_local + (size_t)r * p.np + i);
      b.partials[i] = a;
    }
    __threadfence();
    __syncthreads();
  }
  finish_body(u, b, smean);
}

void launch_wsum(const UpdateCfg& u, const UpdateBufs& b, cudaStream_t st, int* launches) {
  launch_pdl(k_wsum, dim3(u.T * u.nu + 1), dim3(kSumBlock), sizeof(float) * (u.T * u.nu + u.T * u.T), st, u, b);
  ++*launches;
}

// ------------------------------------------------------------------ mean update, filter, cost_total
// mppi.py:494-503 / m3p2i.py:75-87 on the (all-reduced) partial sums; the one-step shift of the stored mean
// (mppi.py:237,266-273) is folded in; Savitzky-Golay as a [T,T] matrix (mppi.py:257-263); cost_total aliasing
// quirk cost_total = sum_t c + mean_k(sum_t c) (mppi.py:282-284,325).
DEV void finish_body(const UpdateCfg& u, const UpdateBufs& b, float* smean /* shared, [T*nu] */) {
  const int T = u.T, nu = u.nu, TN = T * nu;
  const float a2 = u.step_size_mean, a1 = (float)(1.0 - (double)u.step_size_mean);
  const volatile float* part = b.partials;   // written by other CTAs in the fused launch
  float* seq = b.seq;
  float* sfilt = smean + TN;                 // the Savitzky-Golay matrix [T][T], staged once by all threads
  const bool filter = u.filter_u && b.filt;
  if (filter)
    for (int i = threadIdx.x; i < T * T; i += kSumBlock) sfilt[i] = b.filt[i];
  for (int i = threadIdx.x; i < TN; i += kSumBlock) {
    const int t = i / nu, d = i - t * nu;
    const int ts = u.shift ? min(t + 1, T - 1) : t;
    smean[i] = a1 * seq[SEQ_MEAN * TN + ts * nu + d] + a2 * part[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TN; i += kSumBlock) {
    seq[SEQ_MEAN * TN + i] = smean[i];
    if (u.multi_modal) {
      seq[SEQ_MEAN1 * TN + i] = part[TN + i];
      seq[SEQ_MEAN2 * TN + i] = part[2 * TN + i];
      seq[SEQ_BEST1 * TN + i] = part[4 * TN + i];
      seq[SEQ_BEST2 * TN + i] = part[5 * TN + i];
    } else {
      seq[SEQ_BEST * TN + i] = part[3 * TN + i];
    }
    float out = smean[i];
    if (filter) {
      const int t = i / nu, d = i - t * nu;
      float acc = 0.0f;
      for (int jj = 0; jj < T; ++jj) acc += sfilt[t * T + jj] * smean[jj * nu + d];
      out = acc;
    }
    b.result[i] = out;
    b.result[TN + i] = smean[i];
    if (b.host_result) { b.host_result[i] = out; b.host_result[TN + i] = smean[i]; }
  }
  if (u.update_cov && !u.multi_modal) {
    // mppi.py:505-516: delta = actions - NEW mean; cov_update_d = mean_t sum_k w_k delta_ktd^2 from the moments
    // sum w a^2, sum w a (sum w = 1); step_size_cov = 0.7, kappa = 0.005 (mppi.py:202-203)
    __syncthreads();
    if ((int)threadIdx.x < nu) {
      const int d = threadIdx.x;
      double acc = 0.0;
      for (int t = 0; t < T; ++t) {
        const double m = smean[t * nu + d];
        acc += (double)part[6 * TN + 1 + t * nu + d] - 2.0 * m * (double)part[t * nu + d] + m * m;
      }
      Stats* S = b.stats;
      float cov = (1.0f - 0.7f) * S->cov[d] + 0.7f * (float)(acc / (double)T);
      cov += 0.005f;
      S->cov[d] = cov;
      S->sigma[d] = sqrtf(cov);
    }
  }
  const float mean_cost = part[6 * TN] / (float)u.Kg;
  if (!(u.fuse_finish && !b.peer.n))   // (single-rank commands: written by k_wsum's cost-sum CTA)
    for (int k = threadIdx.x; k < u.K; k += kSumBlock) b.cost_total[k] = b.cost_sum[k] + mean_cost;
  if (threadIdx.x == 0) {
    const Stats* S = b.stats;
    M3P2ICommandInfo* in = b.info;
    for (int s = 0; s < 3; ++s) {
      const bool on = s == 0 || u.multi_modal;
      in->eta[s] = on ? S->eta[s] : 0.0f; in->beta[s] = on ? S->beta_used[s] : 0.0f;
      in->min_cost[s] = on ? S->jmin[s] : 0.0f; in->best_idx[s] = on ? S->best_idx[s] : 0;
    }
    in->weight_push = S->weight_push; in->weight_pull = S->weight_pull;
    in->mean_cost_sum = mean_cost;
    in->beta_iters = S->beta_iters;
    in->peer_wait_ms[0] = S->peer_wait_ms[0]; in->peer_wait_ms[1] = S->peer_wait_ms[1];
    in->near_samples = b.near_count ? __ldcg(b.near_count) : -1;
    if (b.host_info) *b.host_info = *in;
  }
}

__global__ void __launch_bounds__(kSumBlock) k_finish(const UpdateCfg u, const UpdateBufs b) {
  extern __shared__ float smean[];  // [T*nu] new mean + [T*T] filter matrix
  finish_body(u, b, smean);
}

void launch_finish(const UpdateCfg& u, const UpdateBufs& b, cudaStream_t st, int* launches) {
  k_finish<<<1, kSumBlock, sizeof(float) * (u.T * u.nu + u.T * u.T), st>>>(u, b);
  ++*launches;
}

// ------------------------------------------------------------------ generic callback path: shift + sample only
// MPPI._shift_action applied in place to the stored sequences (mppi.py:237-242,266-273)
__global__ void k_shift_seq(float* seq, int T, int nu, int multi_modal) {
  extern __shared__ float tmp[];
  const int TN = T * nu;
  for (int q = 0; q < SEQ_COUNT; ++q) {
    const bool on = q == SEQ_MEAN || (multi_modal && (q == SEQ_MEAN1 || q == SEQ_MEAN2 || q == SEQ_BEST1 || q == SEQ_BEST2));
    if (!on) continue;
    for (int i = threadIdx.x; i < TN; i += blockDim.x) {
      const int t = i / nu, d = i - t * nu;
      tmp[i] = seq[q * TN + min(t + 1, T - 1) * nu + d];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TN; i += blockDim.x) seq[q * TN + i] = tmp[i];
    __syncthreads();
  }
}

template <int NU>
__global__ void k_sample_actions(const __grid_constant__ RolloutCfg c, const RolloutBufs b, float* out /*[T][NU][K]*/) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.K) return;
  float u[NU];
  for (int t = 0; t < c.T; ++t) {
    sample_action<NU>(c, b, c.offset + k, k, t, u);
#pragma unroll
    for (int d = 0; d < NU; ++d) out[(size_t)(t * NU + d) * c.K + k] = u[d];
  }
}

void launch_sample_actions(int env_type, const RolloutCfg& c, const RolloutBufs& b, float* seq, float* out, cudaStream_t st) {
  k_shift_seq<<<1, 256, sizeof(float) * c.T * c.nu, st>>>(seq, c.T, c.nu, c.multi_modal);
  const int grid = (c.K + 127) / 128;
  if (env_type == M3P2I_ENV_POINT) k_sample_actions<2><<<grid, 128, 0, st>>>(c, b, out);
  else k_sample_actions<9><<<grid, 128, 0, st>>>(c, b, out);
}

__global__ void k_discount(const float* cost_h, float* J, float* cost_sum, int K, int T, float gamma) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float run = 0.0f, acc = 0.0f, g = 1.0f;
  for (int t = 0; t < T; ++t) {
    const float c = cost_h[(size_t)t * K + k];
    run += c; acc += g * c; g *= gamma;
  }
  J[k] = acc; cost_sum[k] = run;
}

void launch_discount(const float* cost_h, float* J, float* cost_sum, int K, int T, float gamma, cudaStream_t st,
                     int* launches) {
  k_discount<<<(K + 127) / 128, 128, 0, st>>>(cost_h, J, cost_sum, K, T, gamma);
  ++*launches;
}

// ------------------------------------------------------------------ persistent K-env sim facade
__global__ void k_sim_reset(const float* base, float* env, int K, int nf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  for (int f = 0; f < nf; ++f) env[(size_t)f * K + k] = base[f];
}

void launch_sim_reset(int env_type, const float* base, float* env, int K, cudaStream_t st) {
  const int nf = env_type == M3P2I_ENV_POINT ? kPointEnvFloats : kPandaEnvFloats;
  k_sim_reset<<<(K + 127) / 128, 128, 0, st>>>(base, env, K, nf);
}

template <int ENV>
__global__ void __launch_bounds__(kRolloutBlock)
k_sim_step(const __grid_constant__ RolloutCfg c, const __grid_constant__ typename EnvOf<ENV>::Params P, float* env,
           const float* vel_target) {
  constexpr int NU = EnvOf<ENV>::NU;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.K) return;
  typename EnvOf<ENV>::Env e;
  e.load(env, c.K, k);
  float u[NU];
#pragma unroll
  for (int d = 0; d < NU; ++d) u[d] = vel_target[(size_t)d * c.K + k];
  env_step(e, P, u, c);
  e.store(env, c.K, k);
}

void launch_sim_step(int env_type, const RolloutCfg& c, const PointParams* pp, const PandaParams* qp, float* env,
                     const float* vel_target, cudaStream_t st) {
  const int grid = (c.K + kRolloutBlock - 1) / kRolloutBlock;
  if (env_type == M3P2I_ENV_POINT) k_sim_step<M3P2I_ENV_POINT><<<grid, kRolloutBlock, 0, st>>>(c, *pp, env, vel_target);
  else k_sim_step<M3P2I_ENV_PANDA><<<grid, kRolloutBlock, 0, st>>>(c, *qp, env, vel_target);
}

// Objective.compute_cost on the persistent envs; reach costs read rows 0 and K/2 of the batch directly
template <int ENV>
__global__ void __launch_bounds__(kRolloutBlock)
k_sim_cost(const __grid_constant__ RolloutCfg c, const __grid_constant__ typename EnvOf<ENV>::Params P, float* env,
           float* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.K) return;
  typename EnvOf<ENV>::Env e;
  e.load(env, c.K, k);
  PandaRef ref;
  const PandaRef* rp = nullptr;
  if (ENV == M3P2I_ENV_PANDA && c.task == M3P2I_TASK_REACH) {
    PandaEnv r0;
    r0.load(env, c.K, 0);
    ref.cube0[0] = r0.cube[0].p.x; ref.cube0[1] = r0.cube[0].p.y; ref.cube0[2] = r0.cube[0].p.z;
    if (c.multi_modal) r0.load(env, c.K, c.Kg / 2);
    ref.sel_axis = sel_axis_of(r0.cube[0]);
    rp = &ref;
  }
  out[k] = env_cost(e, P, c, c.offset + k, rp);
  e.store(env, c.K, k);  // the pull cost arms the suction forces
}

void launch_sim_cost(int env_type, const RolloutCfg& c, const PointParams* pp, const PandaParams* qp, float* env,
                     float* out_cost, cudaStream_t st) {
  const int grid = (c.K + kRolloutBlock - 1) / kRolloutBlock;
  if (env_type == M3P2I_ENV_POINT) k_sim_cost<M3P2I_ENV_POINT><<<grid, kRolloutBlock, 0, st>>>(c, *pp, env, out_cost);
  else k_sim_cost<M3P2I_ENV_PANDA><<<grid, kRolloutBlock, 0, st>>>(c, *qp, env, out_cost);
}

__global__ void k_sim_links(const __grid_constant__ PandaParams P, const float* env, float* links, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  PandaEnv e;
  e.load(env, K, k);
  float out[39];
  panda_links(P, e.q, e.qd, out);
  for (int i = 0; i < 39; ++i) links[(size_t)k * 39 + i] = out[i];
}

void launch_sim_links(const PandaParams* qp, const float* env, float* links, int K, cudaStream_t st) {
  k_sim_links<<<(K + 63) / 64, 64, 0, st>>>(*qp, env, links, K);
}

__global__ void k_noise_dump(const __grid_constant__ RolloutCfg c, float* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.K) return;
  const int kg = c.offset + k;
  for (int t = 0; t < c.T; ++t) {
    float z[4];
    for (int d = 0; d < c.nu; ++d) {
      if ((d & 3) == 0) noise4(c.noise_mode, c.seed_lo, c.seed_hi, (uint32_t)kg, t, c.T, (uint32_t)(d >> 2), z);
      out[(size_t)(t * c.nu + d) * c.K + k] = z[d & 3];
    }
  }
}

void launch_noise_dump(const RolloutCfg& c, float* out, cudaStream_t st) {
  k_noise_dump<<<(c.K + 127) / 128, 128, 0, st>>>(c, out);
}

__global__ void k_transpose(const float* in, float* out, int rows, int cols) {
  __shared__ float tile[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8)
    if (x < cols && y + j < rows) tile[threadIdx.y + j][threadIdx.x] = in[(size_t)(y + j) * cols + x];
  __syncthreads();
  x = blockIdx.y * 32 + threadIdx.x; y = blockIdx.x * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8)
    if (x < rows && y + j < cols) out[(size_t)(y + j) * rows + x] = tile[threadIdx.x][threadIdx.y + j];
}

void launch_halton_spline(float* out, int K, int offset, int T, int nu, int m, int degree, double smoothing,
                          const int* bases, const unsigned short* perms, int perm_stride, cudaStream_t st) {
  const int n = K * nu, block = 64;   // a few KB of fp64 scratch per thread: small CTAs spread it over all SMs
  k_halton_spline<<<(n + block - 1) / block, block, 0, st>>>(out, K, offset, T, nu, m, degree, smoothing, bases, perms, perm_stride);
}

void launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  k_transpose<<<grid, block, 0, st>>>(in, out, rows, cols);
}

}  // namespace m3
