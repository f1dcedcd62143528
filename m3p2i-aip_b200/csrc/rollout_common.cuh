// rollout_common.cuh -- device code shared by the rollout kernels (thread-per-sample in kernels.cu, lane-cooperative in
// panda_team.cuh): action sampling / perturbation, the producer / consumer hand-over of the batch rows every panda
// reach cost reads, and the all-gather of the discounted costs fused into the rollout (peer memory).
#pragma once
#include "panda_env.cuh"

namespace m3 {

// Perturbed action of GLOBAL sample kg at step t (mppi.py:392-416). `kl` is its row in this shard's tables, or -1.
// The planner sequences are read time-shifted by one step (MPPI._shift_action, mppi.py:266-273): the shift itself
// is applied to the stored mean in k_finish.
// mean + sigma * delta, clamp, mode / gripper / null-action overrides for given unit noise `delta` (mppi.py:392-416)
template <int NU>
DEV void perturb_action(const RolloutCfg& c, const RolloutBufs& b, int kg, int t, const float* delta_in, float* u) {
  const int TN = c.T * NU, half = c.Kg / 2;
  const int ts = c.preshifted ? t : min(t + 1, c.T - 1);
  const float* mean = b.seq + (c.multi_modal ? (kg < half ? SEQ_MEAN1 : SEQ_MEAN2) : SEQ_MEAN) * TN + ts * NU;
#pragma unroll
  for (int d = 0; d < NU; ++d) {
    const float delta = kg == c.Kg - 1 ? 0.0f : delta_in[d];
    float v = mean[d] + delta * (b.sigma_dev ? b.sigma_dev[d] : c.sigma[d]);
    v = fmaxf(fminf(v, c.u_max[d]), c.u_min[d]);
    if (c.multi_modal) {
      if (kg == 0) v = b.seq[SEQ_BEST1 * TN + ts * NU + d];
      if (kg == half) v = b.seq[SEQ_BEST2 * TN + ts * NU + d];
    }
    if (NU == 9 && d >= 7) {
      if (c.gripper == M3P2I_GRIPPER_OPEN) v = 1.5f;
      else if (c.gripper == M3P2I_GRIPPER_CLOSE) v = -1.5f;
    }
    u[d] = c.u_scale * v;
  }
  if (c.null_action && kg == c.Kg - 1) {
#pragma unroll
    for (int d = 0; d < NU; ++d) u[d] = 0.0f;
  }
}

template <int NU>
DEV void sample_action(const RolloutCfg& c, const RolloutBufs& b, int kg, int kl, int t, float* u) {
  const int K = c.K, TN = c.T * NU, half = c.Kg / 2;
  if (c.open_loop) {
#pragma unroll
    for (int d = 0; d < NU; ++d) u[d] = c.u_scale * b.actions_in[(size_t)(t * NU + d) * K + kl];
  } else {
    const int ts = c.preshifted ? t : min(t + 1, c.T - 1);
    const float* mean = b.seq + (c.multi_modal ? (kg < half ? SEQ_MEAN1 : SEQ_MEAN2) : SEQ_MEAN) * TN + ts * NU;
    float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int d = 0; d < NU; ++d) {
      float delta;
      if (c.noise_mode != M3P2I_NOISE_TABLE) {
        // the block that only feeds dim 8 is not drawn when the gripper command overrides the finger targets below
        const bool unused = NU == 9 && d == 8 && (c.gripper == M3P2I_GRIPPER_OPEN || c.gripper == M3P2I_GRIPPER_CLOSE);
        if ((d & 3) == 0 && !unused) noise4(c.noise_mode, c.seed_lo, c.seed_hi, (uint32_t)kg, t, c.T, (uint32_t)(d >> 2), z);
        delta = z[d & 3];
      } else if (kl >= 0) {
        delta = b.noise ? b.noise[(size_t)(t * NU + d) * K + kl] : 0.0f;
      } else {
        delta = b.noise_row0 ? b.noise_row0[t * NU + d] : 0.0f;
      }
      if (kg == c.Kg - 1) delta = 0.0f;  // delta[-1] = 0: the mean itself is always a sample (mppi.py:392)
      float v = mean[d] + delta * (b.sigma_dev ? b.sigma_dev[d] : c.sigma[d]);
      v = fmaxf(fminf(v, c.u_max[d]), c.u_min[d]);
      if (c.multi_modal) {
        if (kg == 0) v = b.seq[SEQ_BEST1 * TN + ts * NU + d];
        if (kg == half) v = b.seq[SEQ_BEST2 * TN + ts * NU + d];
      }
      if (NU == 9 && d >= 7) {
        if (c.gripper == M3P2I_GRIPPER_OPEN) v = 1.5f;
        else if (c.gripper == M3P2I_GRIPPER_CLOSE) v = -1.5f;
      }
      u[d] = c.u_scale * v;
    }
  }
  if (c.null_action && kg == c.Kg - 1) {
#pragma unroll
    for (int d = 0; d < NU; ++d) u[d] = 0.0f;
  }
}

// ------------------------------------------------------------------ rows 0 and Kg/2 of the batch (panda reach)
// Every sample's reach cost reads, after each step, the cube position of sample 0 (cost_functions.py:98,102-103)
// and the cube axis picked from the first row of the second half (skill_utils.py:275-279 on [half_samples:],
// cost_functions.py:151-152). CTA 0 of the rollout grid is a PRODUCER: it replays those two rows (a shard that does
// not own them reconstructs them from the planner state) and publishes refs[t] step by step; all other CTAs are
// consumers that wait for step t before evaluating their cost. CTA 0 is dispatched first, so the producer is
// resident before any consumer can spin, whatever the grid size.
DEV void ref_publish(const RolloutBufs& b, int which, int t, unsigned epoch, const Cube& cubeA, bool with_axis) {
  volatile PandaRef* r = b.refs + t;
  if (which == 0) { r->cube0[0] = cubeA.p.x; r->cube0[1] = cubeA.p.y; r->cube0[2] = cubeA.p.z; }
  if (which == 1 || with_axis) r->sel_axis = sel_axis_of(cubeA);
  __threadfence();
  *(volatile unsigned*)(b.ref_flags + which) = epoch + (unsigned)t + 1u;
}
DEV PandaRef ref_wait(const RolloutBufs& b, int t, unsigned epoch, bool two) {
  const unsigned target = epoch + (unsigned)t + 1u;
  while ((int)(*(volatile unsigned*)(b.ref_flags + 0) - target) < 0) {}
  if (two) while ((int)(*(volatile unsigned*)(b.ref_flags + 1) - target) < 0) {}
  __threadfence();
  const volatile PandaRef* r = b.refs + t;
  PandaRef out;
  out.cube0[0] = r->cube0[0]; out.cube0[1] = r->cube0[1]; out.cube0[2] = r->cube0[2]; out.sel_axis = r->sel_axis;
  return out;
}

// ------------------------------------------------------------------ all-gather of J fused into the rollout
// The thread that owns sample k stores its J into Jg[offset + k] of EVERY mailbox once its rollout is complete
// (push_J_store: plain remote stores). ONE thread per warp (thread-per-sample kernel) or per CTA (team kernel) then
// commits them after a barrier that orders its companions' stores before it: a system-scope fence, the arrival count on
// the rank's ticket, and -- by whoever completes the K arrivals -- the flags in every mailbox. A peer that sees
// jflag == epoch therefore sees every J. (One fence and one atomic per warp / CTA instead of one per sample: at
// K = 4096 that took 8 - 17 us out of a sharded rollout.)
DEV void push_J_store(const PeerPush& p, int offset, int k, float J) {
  for (int r = 0; r < p.n; ++r) p.Jg[r][offset + k] = J;
}
DEV void push_J_commit(const PeerPush& p, int K, unsigned count) {
  __threadfence_system();
  if (atomicAdd(p.ticket, count) == (unsigned)K - count) {
    *p.ticket = 0u;   // the next launch on this stream starts from zero
    __threadfence_system();
    for (int r = 0; r < p.n; ++r) *(volatile unsigned*)(p.jflag[r] + p.rank) = p.epoch;
  }
}

// Threads 0..n-1 of the CTA poll one flag each until it reaches `epoch`; bounded, so that a peer that died cannot
// hang this GPU (the command then reports M3P2I_ERR_STATE through *error).
DEV unsigned long long global_ns() {
#ifndef M3_EMU
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#else
  return 0ull;
#endif
}
// Returns (in every thread) nothing; thread 0 gets the time the CTA spent here in *waited_ms when that is not null.
DEV void wait_flags(const unsigned* flags, int n, unsigned epoch, unsigned timeout_ms, unsigned* error,
                    float* waited_ms = nullptr) {
  const unsigned long long t_in = threadIdx.x == 0 ? global_ns() : 0ull;
  if ((int)threadIdx.x < n) {
    const volatile unsigned* f = flags + threadIdx.x;
    const unsigned long long t0 = global_ns(), limit = (unsigned long long)timeout_ms * 1000000ull;
    unsigned polls = 0;
    while ((int)(*f - epoch) < 0) {
      if ((++polls & 1023u) == 0u && global_ns() - t0 > limit) { *error = 1u; break; }
      __nanosleep(64);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && waited_ms) *waited_ms = (float)(global_ns() - t_in) * 1e-6f;
}

}  // namespace m3
