// kernels.cuh — launch interface between the C ABI (api.cu) and the device code (kernels.cu).
#pragma once
#include "common.cuh"

namespace m3 {

struct UpdateCfg {
  int K, T, nu, Kg, offset, multi_modal, env_type, filter_u, shift;
  int fuse_finish;   // k_wsum's last CTA also runs the finish step (no exchange between them: single rank)
  int update_cov;    // adapt the per-dimension noise variance (mppi.py:508-516; single-mode only)
  int stage_J;       // k_stats keeps the costs of its weight set in shared memory (set by launch_stats)
  float gamma, step_size_mean;
};

// device-resident planner scalars and the statistics of the last softmin (written by k_stats)
struct Stats {
  double beta;          // single-mode inverse temperature carried across commands (mppi.py:184,446-454)
  float scale[3];       // (float)(-1/beta) used for each weight set (all, first half, second half)
  float inv_eta[3];
  float eta[3];
  float beta_used[3];
  float jmin[3];
  int best_idx[3];      // global index of the best sample of each set
  float weight_push, weight_pull;
  int beta_iters;
  float peer_wait_ms[2];   // device time spent in the two waits of the peer-memory exchange (0 when unsharded)
  float cov[kMaxNu];       // per-dimension noise variance cov_action (mppi.py:175,514-515) and its square root, the
  float sigma[kMaxNu];     //   noise scale scale_tril of the next command (mppi.py:176,516); used when update_cov
};

struct UpdateBufs {
  const float* J_global;   // [Kg]
  float* weights;          // [3][Kg]
  Stats* stats;
  const float* actions;    // [T][nu][K]
  const float* cost_sum;   // [K]
  float* partials;         // [7*T*nu + 1]: sums / best rows (6 T nu), sum of costs, sum w a^2 (T nu)
  float* seq;              // [SEQ_COUNT][T*nu]
  const float* filt;       // [T][T] or nullptr
  float* cost_total;       // [K]
  float* result;           // [T*nu] filtered action, followed by [T*nu] unfiltered mean
  M3P2ICommandInfo* info;  // device copy
  float* host_result;      // the same two rows in mapped pinned host memory (written by the kernel: no D2H copy)
  M3P2ICommandInfo* host_info;
  const int* near_count;   // this command's counter of the far-field split (panda_far.cuh) or nullptr: reported in the info
  unsigned* done_counter;  // CTA completion counter of the fused wsum + finish launch
  unsigned* stats_scratch; // [0] CTA completion counter of the multi-modal k_stats, [1..3] beta iterations per set
  PeerReduce peer;
};

// whether launch_rollout will run the far-field kernel (panda_far.cuh) in front of the rollout kernel for this command
bool far_rollout_applies(int env_type, const RolloutCfg& c, const RolloutBufs& b, bool need_refs);
void launch_rollout(int env_type, const RolloutCfg& c, const PointParams* pp, const PandaParams* qp,
                    const RolloutBufs& b, bool need_refs, cudaStream_t st, int* launches);
void launch_stats(const UpdateCfg& u, const UpdateBufs& b, cudaStream_t st, int* launches);
void launch_wsum(const UpdateCfg& u, const UpdateBufs& b, cudaStream_t st, int* launches);
void launch_finish(const UpdateCfg& u, const UpdateBufs& b, cudaStream_t st, int* launches);
// J[k] = sum_t gamma^t cost_h[t][k], cost_sum[k] = sum_t cost_h[t][k] for caller-supplied cost_h (update_only)
void launch_discount(const float* cost_h, float* J, float* cost_sum, int K, int T, float gamma, cudaStream_t st,
                     int* launches);

// shift the stored sequences in place, then write the perturbed actions of this tick (pre-u_scale) to out [T][nu][K]
void launch_sample_actions(int env_type, const RolloutCfg& c, const RolloutBufs& b, float* seq, float* out, cudaStream_t st);

void launch_sim_reset(int env_type, const float* base, float* env, int K, cudaStream_t st);
void launch_sim_step(int env_type, const RolloutCfg& c, const PointParams* pp, const PandaParams* qp, float* env,
                     const float* vel_target, cudaStream_t st);
void launch_sim_cost(int env_type, const RolloutCfg& c, const PointParams* pp, const PandaParams* qp, float* env,
                     float* out_cost, cudaStream_t st);
void launch_sim_links(const PandaParams* qp, const float* env, float* links /*[K][39]*/, int K, cudaStream_t st);
void launch_noise_dump(const RolloutCfg& c, float* out /*[T][nu][K]*/, cudaStream_t st);
// halton-spline noise table (halton_spline.cuh): out [T][nu][K] for global samples offset .. offset + K
void launch_halton_spline(float* out, int K, int offset, int T, int nu, int m, int degree, double smoothing,
                          const int* bases, const unsigned short* perms, int perm_stride, cudaStream_t st);
// [rows][cols] -> [cols][rows]
void launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st);

}  // namespace m3
