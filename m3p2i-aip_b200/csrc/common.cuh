// common.cuh — device-side parameter blocks and small math shared by the rollout / update / sim kernels.
// Layout of everything in HBM is structure-of-arrays with the sample index fastest, so that a warp's
// loads and stores of one field are one contiguous 128-byte line.
#pragma once
#ifndef M3_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "../../include/m3p2i_b200.h"

#define DEV __device__ __forceinline__

// Two constructs that the host build of the device code (tests/emu: the lock-step warp emulator) replaces:
// an optimisation barrier that pins nine floats as values, and the CTA's dynamic shared memory.
#ifndef M3_EMU
#define M3_PIN_VALUES9(a, b, c, d, e, f, g, h, i) \
  asm volatile("" : "+f"(a), "+f"(b), "+f"(c), "+f"(d), "+f"(e), "+f"(f), "+f"(g), "+f"(h), "+f"(i))
#define M3_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif

namespace m3 {

constexpr int kMaxNu = M3P2I_MAX_NU;
constexpr int kMaxStatic = M3P2I_MAX_STATIC;
constexpr int kMaxT = M3P2I_MAX_HORIZON;

// number of floats of one environment in the persistent env buffer (field-major, [field][K])
constexpr int kPointEnvFloats = 22;
constexpr int kPandaEnvFloats = 53;
// panda_env: at most this many link / cube contacts per cube and sub-step are solved (detection order: finger 1,
// finger 2, hand; link corners in the cube before cube corners in the link); the oracle has the same limit
constexpr int kLinkCap = 32;

// planner sequences kept on the device, each [T*nu]
enum Seq { SEQ_MEAN = 0, SEQ_MEAN1, SEQ_MEAN2, SEQ_BEST, SEQ_BEST1, SEQ_BEST2, SEQ_COUNT };

struct Static2 {       // planar oriented box (point_env walls / obstacle)
  float cx, cy, hx, hy, c, s, mu, rad;
};
struct Static3 {       // 3-D oriented box (panda_env table / stands / plate)
  float c[3], R[9], half[3], mu;
};

struct PointParams {
  float robot_radius, robot_mass, robot_mu, drive_damping, drive_effort, gravity, ground_mu;
  float contact_margin, baumgarte, slop, max_corr_vel;
  float box_hx, box_hy, box_mass, box_inertia, box_mu, box_reff;
  float dyn_hx, dyn_hy, dyn_mass, dyn_inertia, dyn_mu, dyn_reff;
  int n_static;
  Static2 st[kMaxStatic];
};

struct PandaParams {
  float base[3], gravity;
  float q_lower[9], q_upper[9], qd_limit[9], effort[9];
  float drive_damping, finger_mass, robot_mu;
  float joint_inertia[7];   // inertia the velocity drive of arm joint j works against
  float warm_start, sleep_lin, sleep_ang, sleep_gap;
  float finger_half[3], finger_center[3], hand_half[3], hand_center[3];
  float contact_margin, baumgarte, slop, max_corr_vel, penalty_stiffness;
  float cube_half[2][3], cube_mass[2], cube_inertia[2], cube_mu[2];
  int n_static, idx_table, idx_shelf, link_sweeps, report_cube;
  Static3 st[kMaxStatic];
};

// per-command scalars (everything a rollout thread needs that is not per-sample)
struct RolloutCfg {
  int K, T, nu, Kg, offset;
  int multi_modal, null_action, noise_mode, substeps, passes;
  int task, gripper;
  int env_live;     // 1: start from the persistent env buffer, 0: from the broadcast base state
  int store_env;    // write the end state (and last velocity target) back to the env buffer
  int preshifted;   // the stored sequences were already shifted (m3p2i_sample_actions): read them at t, not t+1
  int open_loop;    // actions are supplied (m3p2i_rollout_actions) instead of sampled
  unsigned epoch;   // ref_flags reach epoch + t + 1 when step t of this launch has been published
  int align;        // team kernel: 1 = re-align the CTA's warps with a barrier every sub-step (I-cache sharing)
  int lanes;        // lanes per sample: 1 = one thread per sample, 16 = lane-cooperative team (panda_env)
  // large K behind the far-field kernel: both rollout kernels are launched over the near list and the count decides on the
  // device which one works -- the team kernel up to near_team_max samples (0: no limit), the thread-per-sample kernel
  // above near_thread_min - 1 (0: always)
  int near_team_max, near_thread_min;
  float dt, gamma, u_scale, kp_suction, pre_height_diff, tilt_cos;
  float u_min[kMaxNu], u_max[kMaxNu], sigma[kMaxNu];
  float goal[8];
  uint32_t seed_lo, seed_hi;
  // the packed start state (one env, field-major: m3p2i_set_state) travels with the launch: no H2D copy per command,
  // and every thread reads it from the constant bank
  float base_env[56];
};

// values of sample 0 / sample K/2 of the global batch read by every sample's reach cost
struct PandaRef {
  float cube0[3];
  int sel_axis;
};

// ------------------------------------------------------------------ exchange over NVLink peer memory (sharded K)
// Every rank owns a MAILBOX in its own HBM that all peers map (cudaIpc): two parities of
//   Jg[Kg] | part[n][NP] | jflag[n] | pflag[n]
// The rollout kernel of rank r stores its discounted costs straight into Jg[offset_r ...] of EVERY mailbox (its own
// included) and the last sample to finish raises jflag[r] there: rollout + all-gather in one kernel. The weighted-sum
// kernel's last CTA does the same with the packed partial sums (part[r], pflag[r]) and then adds the n boxes of its
// own mailbox in rank order: sums + all-reduce + mean update in one kernel, bitwise identical on every rank.
// Flags carry the command's epoch; parity = epoch & 1 (a rank can be at most one command ahead of a peer).
constexpr int kMaxPeers = 8;

struct PeerPush {          // rollout side
  int n, rank;             // n == 0: no peers (single rank, or NCCL / host exchange)
  unsigned epoch;
  float* Jg[kMaxPeers];        // Jg of this parity in the mailbox of rank i
  unsigned* jflag[kMaxPeers];  // jflag[0..n) of this parity in the mailbox of rank i
  unsigned* ticket;            // local: number of samples of this launch whose J has been pushed
};

struct PeerReduce {        // update side
  int n, rank, np;         // np = padded length of one partials box
  unsigned epoch;
  unsigned timeout_ms;     // give up (set *error) after waiting this long for a flag: a dead peer must not hang the GPU
  const unsigned* jflag_local;   // [n] this parity, local mailbox (k_stats waits for them)
  const float* part_local;       // [n][np] this parity, local mailbox
  const unsigned* pflag_local;   // [n]
  float* part[kMaxPeers];        // part[0..n) of this parity in the mailbox of rank i
  unsigned* pflag[kMaxPeers];
  unsigned* error;               // local: set to 1 when a wait gave up
};

struct RolloutBufs {
  const float* noise;      // [T][nu][K] (table mode) or nullptr
  const float* noise_row0; // [T][nu] noise of GLOBAL sample 0 for shards that do not own it, or nullptr
  const float* seq;        // [SEQ_COUNT][T*nu] planner sequences (un-shifted; the kernel reads them shifted)
  const float* actions_in; // [T][nu][K] open-loop actions or nullptr
  const float* sigma_dev;  // [nu] adapted noise scale (update_cov) or nullptr: use RolloutCfg::sigma
  float* env;              // [fields][K] persistent envs
  float* vel_target;       // [nu][K]
  float* actions;          // [T][nu][K]
  float4* states;          // [T][K] (x, vx, y, vy) / (q1, qd1, q2, qd2)
  float* cost_h;           // [T][K]
  float* J;                // [K]  discounted cost
  float* cost_sum;         // [K]  undiscounted sum
  PandaRef* refs;          // [T] rows 0 / Kg/2 of the batch, published step by step by the producer CTA
  unsigned* ref_flags;     // [2] progress counters of the two producers (cube position, cube axis)
  // far-field split (panda_far.cuh): k_rollout_far finishes the samples that stay in the far field and lists the others;
  // the rollout kernel launched after it processes exactly near_list[0 .. *near_count). nullptr: all K samples.
  int* near_list;          // [K] rows of the samples that need the full rollout
  int* near_count;         // their number (this command's counter)
  int* near_count_next;    // the other parity's counter, cleared for the next command
  int* far_info;           // [0], [1]: support box of cubeA / cubeB in the start state (the sleeping cubes' weight goes there)
  float* far_dump;         // [K][far_boundaries][18] joint states of the listed samples at their hand-over boundaries
  PeerPush peer;
};

// Hand-over of a sample that leaves the far field at iteration i: the rollout kernel does not have to start it from
// iteration 0 -- up to the last multiple of 8 iterations before i nothing but the nine joints moved. far_dump keeps, per
// listed sample, (position, velocity) of the nine joints at the iterations 0, 8, 16, ... up to that boundary:
// [K][far_boundaries][18]; the near list entry carries the boundary (bits 26..) next to the row (bits 0..25).
__host__ __device__ inline int far_boundaries(int T, int ns) { return T * ns / 8 + 1; }
constexpr int kFarRowBits = 26;

// ------------------------------------------------------------------ small math
// ------------------------------------------------------------------ programmatic dependent launch
// The three kernels of a command depend on each other grid-to-grid. Launched with programmatic stream serialization the
// next kernel's CTAs are scheduled as soon as the previous grid has issued launch_dependents and SM resources are free;
// they then wait in pdl_wait() until the previous grid has completed and its writes are visible. That takes the
// launch latency of k_stats / k_wsum out of the command (the rollout's last CTAs finish while they are being set up).
DEV void pdl_wait() {
#ifndef M3_EMU
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
DEV void pdl_launch_dependents() {
#ifndef M3_EMU
  asm volatile("griddepcontrol.launch_dependents;");
#endif
}
DEV float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
DEV float signf(float v) { return v < 0.0f ? -1.0f : 1.0f; }

struct V3 {
  float x, y, z;
};
DEV V3 mk(float x, float y, float z) { V3 r = {x, y, z}; return r; }
DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
DEV V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
DEV V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DEV float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct M33 {  // columns
  V3 cx, cy, cz;
};
DEV V3 mul(const M33& R, V3 l) { return l.x * R.cx + l.y * R.cy + l.z * R.cz; }          // R l
DEV V3 mulT(const M33& R, V3 o) { return mk(dot(R.cx, o), dot(R.cy, o), dot(R.cz, o)); }  // R^T o
DEV V3 col(const M33& R, int i) { return i == 0 ? R.cx : (i == 1 ? R.cy : R.cz); }

// rotation matrix of a unit quaternion (x,y,z,w)
DEV M33 quat_to_R(float x, float y, float z, float w) {
  M33 R;
  R.cx = mk(1.0f - 2.0f * (y * y + z * z), 2.0f * (x * y + w * z), 2.0f * (x * z - w * y));
  R.cy = mk(2.0f * (x * y - w * z), 1.0f - 2.0f * (x * x + z * z), 2.0f * (y * z + w * x));
  R.cz = mk(2.0f * (x * z + w * y), 2.0f * (y * z - w * x), 1.0f - 2.0f * (x * x + y * y));
  return R;
}

// ------------------------------------------------------------------ Philox4x32-10 counter RNG
DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// four N(0,1) draws for counter (global sample, time step, dimension group); Box-Muller on 24-bit uniforms
DEV void normal4(uint32_t seed_lo, uint32_t seed_hi, uint32_t kg, uint32_t t, uint32_t g, float z[4]) {
  uint32_t r[4];
  philox4x32_10(kg, t, g, 0u, seed_lo, seed_hi, r);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float u1 = ((float)(r[2 * i] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(r[2 * i + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    z[2 * i] = rad * cs;
    z[2 * i + 1] = rad * sn;
  }
}


// M3P2I_NOISE_PHILOX_SPLINE: control points c_0 .. c_{nseg+1} ~ N(0,1) per (sample, dimension) at counters
// t = 0x40000000 + i; the value at step t is the uniform quadratic B-spline at s = t * nseg / (T - 1), divided by the
// norm of the three blending weights so that every step keeps unit variance.
DEV void spline_weights(int t, int T, int& i0, float w[3]) {
  const int nseg = max(T / 4, 2);
  const float s = T > 1 ? (float)t * (float)nseg / (float)(T - 1) : 0.0f;
  i0 = min((int)s, nseg - 1);
  const float f = s - (float)i0;
  w[0] = 0.5f * (1.0f - f) * (1.0f - f);
  w[1] = 0.5f + f - f * f;
  w[2] = 0.5f * f * f;
  const float inv = 1.0f / sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  w[0] *= inv; w[1] *= inv; w[2] *= inv;
}

// four unit normals of (global sample kg, step t, dimension group g) for the counter-based noise modes
DEV void noise4(int noise_mode, uint32_t seed_lo, uint32_t seed_hi, uint32_t kg, int t, int T, uint32_t g, float z[4]) {
  if (noise_mode != M3P2I_NOISE_PHILOX_SPLINE) {
    normal4(seed_lo, seed_hi, kg, (uint32_t)t, g, z);
    return;
  }
  int i0;
  float w[3], c[4];
  spline_weights(t, T, i0, w);
  z[0] = z[1] = z[2] = z[3] = 0.0f;
#pragma unroll 1
  for (int j = 0; j < 3; ++j) {
    normal4(seed_lo, seed_hi, kg, 0x40000000u + (uint32_t)(i0 + j), g, c);
    const float wj = j == 0 ? w[0] : (j == 1 ? w[1] : w[2]);
    z[0] += wj * c[0]; z[1] += wj * c[1]; z[2] += wj * c[2]; z[3] += wj * c[3];
  }
}

}  // namespace m3
