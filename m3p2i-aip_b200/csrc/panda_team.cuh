// panda_team.cuh — lane-cooperative Panda rollout: a TEAM of lanes advances ONE sample.
//
// Why: at the sizes the planner runs at (K = 4096 samples) a thread-per-sample rollout is one warp per SM walking a
// ~20 k-instruction serial chain per step, with 3/4 of the SM sub-partitions idle. The work inside a sample is mostly
// contact detection (independent per box corner) and Gauss-Seidel impulse solves (serial per body, but the two cubes
// are independent of each other). A team splits it like this (CPL = corners per lane, 1 or 2):
//     group bit (g)   : which cube the lane works for (0 = cubeA, 1 = cubeB); each group of 8 / CPL lanes keeps a
//                       replica of "its" cube's state and applies that cube's impulses in lock-step
//     lane in group c : the lane tests box corners c, c + 8 / CPL, ... (CPL of them)
// so a team is 16 / CPL lanes and a warp carries 2 * CPL samples. CPL = 1 has the shortest per-sample chain (best
// while the GPU is far from full), CPL = 2 halves the number of warp-instructions per sample because the serial
// solves, the forward kinematics and the cost are replicated over 8 instead of 16 lanes.
// The arm runs ahead of the contacts: no contact acts on the seven velocity-tracked arm joints, so once per block of
// 16 / CPL iterations lane j advances arm joint j through the whole block and lane l runs the forward kinematics of
// iteration l of the block; an iteration then fetches its hand pose with shuffles (team_rollout). Finger joints, link
// boxes and the cost are replicated in all lanes of a team. Contact detection and every geometry-only contact term
// run on all corners at once; the velocity part of the solves is broadcast (warp shuffles) and applied in exactly the
// pair / corner order of the thread-per-sample code (panda_env.cuh), so both produce the same trajectory up to fp32
// rounding of reordered sums.
// All branches that contain shuffles are warp-uniform (decided by __ballot_sync / __any_sync).
//
// Code size is a first-order concern here: with ~14 resident warps per SM at different program counters the kernel
// is instruction-fetch bound as soon as its loop body outgrows the instruction cache (ncu: 75 % `no_inst` stalls with
// the unrolled version in contact-rich states). Hence ONE copy of everything: one forward-kinematics site (the cost
// of step t is evaluated from the hand pose of the first sub-step of step t+1 — same joint positions), link / cube
// pairs, the two detection directions and the corner slots as rolled loops, one inlined copy of each contact solve.
#pragma once
#include "rollout_common.cuh"

namespace m3 {

constexpr unsigned kFull = 0xffffffffu;
constexpr float kGripReach = kGripReachT;

template <int CPL>
struct TeamShape {
  static_assert(CPL == 1 || CPL == 2, "corners per lane: 1 (16-lane teams) or 2 (8-lane teams)");
  static constexpr int kGroup = 8 / CPL;    // lanes per cube
  static constexpr int kTeam = 16 / CPL;    // lanes per sample
};

DEV V3 shfl3(V3 a, int src) {
  return mk(__shfl_sync(kFull, a.x, src), __shfl_sync(kFull, a.y, src), __shfl_sync(kFull, a.z, src));
}
DEV V3 sel3(bool c, V3 a, V3 b) { return mk(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z); }

// corner i of box b relative to its centre (box_corner(b, i) == b.c + corner_arm(b, i))
DEV V3 corner_arm(const OBox3& b, int i) {
  return mul(b.R, mk((i & 1) ? b.half.x : -b.half.x, (i & 2) ? b.half.y : -b.half.y, (i & 4) ? b.half.z : -b.half.z));
}

// fold a warp ballot onto the lanes of one group: bit j = lane j of ANY group of the warp voted
template <int G>
DEV unsigned fold(unsigned hb) {
  hb |= hb >> 16;
  hb |= hb >> 8;
  if (G <= 4) hb |= hb >> 4;
  return hb & ((1u << G) - 1u);
}

struct TeamLane {
  int lane, g, c;     // lane in warp, cube group, lane in group (first corner)
  int tl;             // lane in team
  int team_base;      // first lane of this team in the warp
  int group_base;     // first lane of this group in the warp
  int other;          // the lane with the same corners in the other group of the team
};

template <int CPL>
DEV TeamLane team_lane() {
  constexpr int G = TeamShape<CPL>::kGroup, TM = TeamShape<CPL>::kTeam;
  TeamLane t;
  t.lane = threadIdx.x & 31;
  t.g = (t.lane / G) & 1;
  t.c = t.lane & (G - 1);
  t.tl = t.lane & (TM - 1);
  t.team_base = t.lane & ~(TM - 1);
  t.group_base = t.lane & ~(G - 1);
  t.other = t.lane ^ G;
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
};

// The generic two-body solve is only reached when the two cubes touch each other; one out-of-line copy.
__device__ __noinline__ V3 solve_contact3_call(Dyn3& A, Dyn3& B, V3 n, float depth, V3 c, float mu, float h,
                                               const PandaParams& P, float4& L, bool first) {
  // the slot's warm start was prepared by its owner (warm_prepare): `first` only has to apply it
  V3 done = mk(0, 0, 0);
  if (first) {
    const V3 Pw = L.x * n + mk(L.y, L.z, L.w);
    apply_impulse(A, c - A.x, Pw, 1.0f); apply_impulse(B, c - B.x, Pw, -1.0f);
    done = Pw;
  }
  return done + solve_contact3_acc(A, B, n, depth, c, mu, h, P, L, false, false);
}

DEV Dyn3 dyn_cube(V3 v, V3 w, V3 x, float im, float ii) {
  Dyn3 d;
  d.v = v; d.w = w; d.x = x; d.im = im; d.ii = ii; d.axis = mk(0, 0, 0); d.slide = 0.0f; d.ims = 0.0f;
  return d;
}

// Serial application of the contacts found by the G lanes [src0, src0+G): for lane j = 0..G-1 in order, the lane
// that found a hit broadcasts (n, depth, p) together with its accumulator slot L and every lane for which `mine` holds
// applies solve(n, depth, p, L); the lane that owns the slot keeps the updated L. Returns the sum of the impulses the
// solves report. The loop bounds and the shuffles are warp-uniform.
DEV float4 shfl4(float4 a, int src) {
  return make_float4(__shfl_sync(kFull, a.x, src), __shfl_sync(kFull, a.y, src), __shfl_sync(kFull, a.z, src),
                     __shfl_sync(kFull, a.w, src));
}
template <int G, typename Solve>
DEV V3 apply_hits(bool hit, V3 n, float depth, V3 p, float4& L, int src0, bool mine, Solve&& solve) {
  const unsigned hb = __ballot_sync(kFull, hit);
  const int lane = threadIdx.x & 31;
  V3 acc = mk(0, 0, 0);
  // lanes that hit in ANY group of the warp, visited in ascending order
  unsigned todo = fold<G>(hb);
  while (todo) {
    const int j = __ffs(todo) - 1;
    todo &= todo - 1;
    const int src = src0 + j;
    const bool act = mine && ((hb >> src) & 1u);
    const V3 nj = shfl3(n, src), pj = shfl3(p, src);
    const float dj = __shfl_sync(kFull, depth, src);
    float4 Lj = shfl4(L, src);
    if (act) {
      acc = acc + solve(nj, dj, pj, Lj);
      if (lane == src) L = Lj;
    }
  }
  return acc;
}

// One own-corner / fixed-box contact as the lane that found it describes it. Positions do not move inside a
// sub-step, so everything that depends on geometry only (lever arm ra, ra x n, 1 / normal effective mass, target
// normal speed) is worked out ONCE, by all corner lanes in parallel, and only the velocity-dependent part of
// solve_cube_static is left in the serial Gauss-Seidel chain.
struct StaticHit {
  bool hit, sup;   // sup: the corner is within sleep_gap of the box (it supports the cube)
  V3 n, rn;
  float ikn, target;
};

DEV StaticHit static_hit(bool near, V3 pc, V3 ra, const OBox3& sb, float im, float ii, float inv_h, const PandaParams& P) {
  StaticHit s;
  s.n = mk(0, 0, 0);
  float depth = 0.0f;
  s.hit = near && point_in_box(pc, sb, P.contact_margin, s.n, depth);
  s.sup = s.hit && depth > -P.sleep_gap;
  s.rn = cross(ra, s.n);
  s.ikn = __fdividef(1.0f, im + ii * dot(s.rn, s.rn));
  s.target = depth > 0.0f ? fminf(P.baumgarte * fmaxf(depth - P.slop, 0.0f) * inv_h, P.max_corr_vel) : depth * inv_h;
  return s;
}

// The warm start of an accumulator slot, prepared by the lane that owns it before the serial chain starts: what the
// slot held in the previous sub-step (if it was in contact then), scaled, projected onto the tangent plane of the new
// normal and clamped to the new cone -- or zero. The serial chain then just applies it at the slot's first visit.
DEV float4 warm_prepare(float4 L, bool warm, V3 n, float mu, float k) {
  if (!warm) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  const V3 lt0 = mk(L.y, L.z, L.w);
  const float tn = dot(lt0, n);
  const float ln = k * L.x;
  V3 lt = k * (lt0 - tn * n);
  const float lim = mu * ln, m2 = dot(lt, lt);
  if (m2 > lim * lim) lt = (lim * rsqrtf(m2)) * lt;
  return make_float4(ln, lt.x, lt.y, lt.z);
}

// Serial application (ascending corner order, as apply_hits) of the fixed-box contacts of the own cube: the equations
// of solve_contact3_acc for a free cube against a fixed body, with the geometry terms taken from the StaticHit of the
// lane that found the contact. `L` = this lane's accumulator for the slot (prepared by warm_prepare when `first`);
// `keep` = the slot accumulates over visits (the cube's first near fixed box); otherwise every visit starts from zero,
// which is the plain one-shot solve.
template <int G>
DEV V3 apply_static_hits(const StaticHit& s, V3 ra, int group_base, V3& v, V3& w, float im, float ii, float mu, float4& L,
                         bool keep, bool first) {
  const unsigned hb = __ballot_sync(kFull, s.hit);
  const int lane = threadIdx.x & 31;
  V3 acc = mk(0, 0, 0);
  unsigned todo = fold<G>(hb);
  while (todo) {
    const int j = __ffs(todo) - 1;
    todo &= todo - 1;
    const int src = group_base + j;
    const V3 n = shfl3(s.n, src), rn = shfl3(s.rn, src), r = shfl3(ra, src);
    const float ikn = __shfl_sync(kFull, s.ikn, src), target = __shfl_sync(kFull, s.target, src);
    float4 Lj = shfl4(L, src);
    if (!((hb >> src) & 1u)) continue;
    if (!keep) Lj = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (first && keep) {
      const V3 Pw = Lj.x * n + mk(Lj.y, Lj.z, Lj.w);
      v = v + im * Pw;
      w = w + ii * cross(r, Pw);
      acc = acc + Pw;
    }
    const float ln = fmaxf(Lj.x + (target - (dot(v, n) + dot(w, rn))) * ikn, 0.0f);
    const float dj = ln - Lj.x;
    v = v + (dj * im) * n;
    w = w + (dj * ii) * rn;
    const V3 va = v + cross(w, r);
    const float vn = dot(va, n);
    V3 t = va - vn * n;
    const float vt2 = dot(t, t);
    V3 lt = mk(Lj.y, Lj.z, Lj.w);
    if (vt2 >= 1e-18f) {
      const float ivt = rsqrtf(vt2), vt = vt2 * ivt;
      t = ivt * t;
      const V3 rt = cross(r, t);
      const float kt = im + ii * dot(rt, rt);
      lt = lt - __fdividef(vt, kt) * t;
    }
    const float lim = mu * ln, m2 = dot(lt, lt);
    if (m2 > lim * lim) lt = (lim * rsqrtf(m2)) * lt;
    const V3 Pt = lt - mk(Lj.y, Lj.z, Lj.w);
    v = v + im * Pt;
    w = w + ii * cross(r, Pt);
    acc = acc + dj * n + Pt;
    if (keep && lane == src) L = make_float4(ln, lt.x, lt.y, lt.z);
  }
  return acc;
}

// The same split for a kinematic link (prescribed twist; a finger adds one sliding DoF) against the own cube:
// geometry and the link's prescribed point velocity per contact in parallel, velocity part serial. Normal `n` points
// from the cube to the link (solve_link_cube's convention).
struct LinkHit {
  bool hit;
  V3 n, rcn, rc, vl0;
  float an, ikn, target;
};

struct Hand;
DEV LinkHit link_hit(bool hit, V3 n, float depth, V3 pt, V3 hv, V3 hw, V3 hp, V3 axis, float ims, V3 x, float im,
                     float ii, float inv_h, const PandaParams& P) {
  LinkHit s;
  s.hit = hit; s.n = n;
  s.rc = pt - x;
  s.vl0 = hv + cross(hw, pt - hp);
  s.rcn = cross(s.rc, n);
  s.an = dot(axis, n);
  s.ikn = __fdividef(1.0f, ims * s.an * s.an + im + ii * dot(s.rcn, s.rcn));
  s.target = depth > 0.0f ? fminf(P.baumgarte * fmaxf(depth - P.slop, 0.0f) * inv_h, P.max_corr_vel) : depth * inv_h;
  return s;
}

// The link / cube contacts of a sub-step are detected ONCE (positions do not move inside a sub-step) and kept as
// RECORDS in shared memory, one list per (sample, cube) in the solve order (link 0, 1, 2; corners of the link in the
// cube, then corners of the cube in the link; ascending corner). A record is four float4:
//   [0] n, target   [1] rc, an   [2] vl0, ikn   [3] rc x n, index of the slot's accumulator in the CTA's accumulators
// At most kLinkCap records per cube (further contacts, in detection order, are ignored -- oracle and thread-per-sample
// kernel do the same). kRecStride = float4 per list; the odd 4-word pad puts the lists of a warp in different banks.
constexpr int kRecStride = 4 * kLinkCap + 1;

// The serial Gauss-Seidel chain over the record list of one cube, walked by the cube's first lane (`n_own` = length of
// its list; 0 in every other lane, which only takes part in the votes). A visit is the equations of solve_contact3_acc
// for a kinematic link (one sliding DoF for a finger) against a free cube, geometry terms from the record, accumulator
// of the slot read and written in shared memory. `first`: the slot's first visit of the sub-step applies its warm start
// (prepared by warm_prepare at detection). The link of the record (bits 24.. of its last word) selects the sliding DoF:
// slide[0] / slide[1] for the fingers, none for the hand.
// Most contacts of a grasp are OPEN (speculative, gap > 0): 3.6 of 5.5 carry nothing and receive nothing -- their
// friction step would clamp to the empty cone and change no velocity. A lane therefore runs ahead over the idle records
// of ITS list with the cheap part of the visit (is the accumulator empty -- bit j of `nz` -- and does the normal impulse
// stay zero?) and the lanes of the warp only meet for the records that need the full visit: the number of full visits
// per sweep is the largest number of ACTIVE contacts among the warp's samples, not the longest list.
// `nz`: bit j = the accumulator of record j holds something (set at detection from the warm start, kept up to date here).
// Returns the impulse on the links when WANT_SUM (cubeB's contacts are reported), else zero.
template <bool WANT_SUM>
DEV V3 walk_link_records(const float4* list, int n_own, unsigned& nz, float4* lam_base, V3 ycol, float* slide, float ims_f,
                         V3& v, V3& w, float im, float ii, float mu, bool first) {
  V3 acc = mk(0, 0, 0);
  int j = 0;
  for (;;) {
    float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
    bool found = false;
    while (j < n_own) {
      const float4* r = list + 4 * j;
      q0 = r[0]; q1 = r[1]; q2 = r[2]; q3 = r[3];
      if ((nz >> j) & 1u) { found = true; break; }
      // empty accumulator: the visit does something only if the normal impulse becomes positive now
      const int fq = __float_as_int(q3.w) >> 24;
      const float slq = fq == 1 ? slide[1] : slide[0];
      const V3 nq = mk(q0.x, q0.y, q0.z);
      const float vnq = dot(mk(q2.x, q2.y, q2.z), nq) + slq * q1.w - (dot(v, nq) + dot(w, mk(q3.x, q3.y, q3.z)));
      if (fmaxf((q0.w - vnq) * q2.w, 0.0f) != 0.0f) { found = true; break; }
      ++j;
    }
    if (!__any_sync(kFull, found)) break;
    if (found) {
      const V3 n = mk(q0.x, q0.y, q0.z), rc = mk(q1.x, q1.y, q1.z), vl0 = mk(q2.x, q2.y, q2.z), rcn = mk(q3.x, q3.y, q3.z);
      const float target = q0.w, an = q1.w, ikn = q2.w;
      const int meta = __float_as_int(q3.w), f = meta >> 24;
      // link f: finger 1 slides along +y of the hand, finger 2 along -y, the hand box (f = 2) has no sliding DoF
      const V3 axis = f == 0 ? ycol : (f == 1 ? -ycol : mk(0, 0, 0));
      const float ims = f < 2 ? ims_f : 0.0f;
      float sl = f == 1 ? slide[1] : slide[0];
      float4* const Lp = lam_base + (meta & 0xffffff);
      const float4 Lj = *Lp;
      if (first) {
        const V3 Pw = Lj.x * n + mk(Lj.y, Lj.z, Lj.w);   // on the link; the cube receives -Pw
        sl += ims * dot(axis, Pw);
        v = v - im * Pw;
        w = w - ii * cross(rc, Pw);
        if (WANT_SUM) acc = acc + Pw;
      }
      const float vn0 = dot(vl0, n) + sl * an - (dot(v, n) + dot(w, rcn));
      const float ln = fmaxf(Lj.x + (target - vn0) * ikn, 0.0f);
      const float dj = ln - Lj.x;
      sl += ims * an * dj;
      v = v - (dj * im) * n;
      w = w - (dj * ii) * rcn;
      const V3 rv = (vl0 + sl * axis) - (v + cross(w, rc));
      const float vn = dot(rv, n);
      V3 t = rv - vn * n;
      const float vt2 = dot(t, t);
      V3 lt = mk(Lj.y, Lj.z, Lj.w);
      if (vt2 >= 1e-18f) {
        const float ivt = rsqrtf(vt2), vt = vt2 * ivt;
        t = ivt * t;
        const V3 rct = cross(rc, t);
        const float at = dot(axis, t);
        const float kt = ims * at * at + im + ii * dot(rct, rct);
        lt = lt - __fdividef(vt, kt) * t;
      }
      const float lim = mu * ln, m2 = dot(lt, lt);
      if (m2 > lim * lim) lt = (lim * rsqrtf(m2)) * lt;
      const V3 Pt = lt - mk(Lj.y, Lj.z, Lj.w);
      sl += ims * dot(axis, Pt);
      v = v - im * Pt;
      w = w - ii * cross(rc, Pt);
      *Lp = make_float4(ln, lt.x, lt.y, lt.z);
      // (ln == 0 empties the cone: lt == 0 too)
      nz = (ln != 0.0f || lt.x != 0.0f || lt.y != 0.0f || lt.z != 0.0f) ? (nz | (1u << j)) : (nz & ~(1u << j));
      if (f == 0) slide[0] = sl;
      if (f == 1) slide[1] = sl;
      if (WANT_SUM) acc = acc + dj * n + Pt;
      ++j;
    }
  }
  return acc;
}

// hand pose + twist from the sines / cosines and speeds of the seven arm joints (the arithmetic of panda_hand)
DEV void fk_from_sincos(const PandaParams& P, const float* sn, const float* cs, const float* qd, Hand& H) {
  V3 p = mk(P.base[0], P.base[1], P.base[2]);
  M33 R = {mk(1, 0, 0), mk(0, 1, 0), mk(0, 0, 1)};
  V3 v = mk(0, 0, 0), w = mk(0, 0, 0);
  constexpr float X[7] = {0.0f, 0.0f, 0.0f, 0.0825f, -0.0825f, 0.0f, 0.088f};
  constexpr float Y[7] = {0.0f, 0.0f, -0.316f, 0.0f, 0.384f, 0.0f, 0.0f};
  constexpr float Z[7] = {0.333f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  constexpr int ROLL[7] = {0, -1, 1, 1, -1, 1, 1};
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const V3 d = mul(R, mk(X[i], Y[i], Z[i]));
    v = v + cross(w, d);
    p = p + d;
    if (ROLL[i] == 1) { const V3 c1 = R.cy; R.cy = R.cz; R.cz = -c1; }
    if (ROLL[i] == -1) { const V3 c1 = R.cy; R.cy = -R.cz; R.cz = c1; }
    w = w + qd[i] * R.cz;
    const V3 c0 = R.cx, c1 = R.cy;
    R.cx = cs[i] * c0 + sn[i] * c1;
    R.cy = cs[i] * c1 - sn[i] * c0;
  }
  const V3 d = kHandZ * R.cz;
  v = v + cross(w, d);
  p = p + d;
  const V3 c0 = R.cx, c1 = R.cy;
  R.cx = kHandYawC * c0 + kHandYawS * c1;
  R.cy = kHandYawC * c1 - kHandYawS * c0;
  H.p = p; H.R = R; H.v = v; H.w = w;
}

// The whole rollout of one sample by one team. `k` = row of the sample in this shard's buffers (or -1 for a producer
// team replaying a row of another shard), `kg` its global id. Producer teams publish refs instead of costs.
template <int CPL>
DEV void team_rollout(const RolloutCfg& c, const PandaParams& P, const RolloutBufs& b, const TeamLane& t, int k, int kg,
                      bool valid, bool producer, int which, int it0 = 0) {
  constexpr int NU = 9;
  constexpr int G = TeamShape<CPL>::kGroup, TM = TeamShape<CPL>::kTeam;
  const int K = c.K, T = c.T, ns = c.substeps;
  const bool writer = valid && t.tl == 0;
  const bool use_refs = b.refs != nullptr;
  const int g = t.g;
  const float h = c.dt / (float)ns, D = P.drive_damping;
  const float im = 1.0f / P.cube_mass[g], ii = 1.0f / P.cube_inertia[g], mu_c = P.cube_mu[g];
  const float imo = 1.0f / P.cube_mass[g ^ 1], iio = 1.0f / P.cube_inertia[g ^ 1];
  const V3 half_own = mk(P.cube_half[g][0], P.cube_half[g][1], P.cube_half[g][2]);
  const V3 half_oth = mk(P.cube_half[g ^ 1][0], P.cube_half[g ^ 1][1], P.cube_half[g ^ 1][2]);
  const float rad_own = sqrtf(dot(half_own, half_own)), rad_oth = sqrtf(dot(half_oth, half_oth));

  TeamEnv e;
  if (c.env_live && k >= 0) e.load(b.env, K, k, g);
  else e.load(c.base_env, 1, 0, g);
  float run = 0.0f, J = 0.0f, gam = 1.0f;
  if (it0 > 0) {
    // hand-over from the far-field kernel (panda_far.cuh) at iteration it0 (warp-uniform; a multiple of the run-ahead
    // block and of the sub-steps): until then only the joints moved -- their state at the boundary comes from the
    // dump --, both cubes slept where they started, and the costs of the finished steps are in cost_h
    const float* d = b.far_dump + ((size_t)k * far_boundaries(T, ns) + (it0 >> 3)) * 18;
#pragma unroll
    for (int j = 0; j < 9; ++j) { e.q[j] = d[2 * j]; e.qd[j] = d[2 * j + 1]; }
    e.cu.v = mk(0, 0, 0); e.cu.w = mk(0, 0, 0);
    const int done = it0 / ns;
    for (int ts = 0; ts < done; ++ts) {
      const float cost = b.cost_h[(size_t)ts * K + k];
      run += cost;
      J += gam * cost;
      gam *= c.gamma;
    }
  }
  // group-partial impulse sums (identical in the lanes of a group), lane-partial penalty sums, per step
  V3 imp_table = mk(0, 0, 0), imp_shelf = mk(0, 0, 0), imp_cubeb = mk(0, 0, 0), pen = mk(0, 0, 0);

  // ---- the arm runs ahead of the contacts. The seven arm joints are velocity-tracked and no contact acts on them,
  // so their trajectory (and the hand pose) depends on the sampled actions only. Every TM iterations the team
  // advances the arm TM sub-steps at once: lane j owns joint j (drive, limits, integration, one sincosf site), its
  // (sin, cos, speed) are broadcast, and lane l keeps the snapshot of iteration blk0 + l and runs ONE forward
  // kinematics for it. Inside the block an iteration then fetches its hand pose with 18 shuffles instead of
  // recomputing the kinematic chain in every lane, and the action of a step is drawn once per lane per block.
  const int jo = min(t.tl, 6);   // own arm joint (lanes 7.. shadow joint 6, their results are never read)
  float qo = e.q[0], vo = e.qd[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) { if (jo == i) { qo = e.q[i]; vo = e.qd[i]; } }
  const float lo_o = P.q_lower[jo], up_o = P.q_upper[jo], vl_o = P.qd_limit[jo], ef_o = P.effort[jo], m_o = P.joint_inertia[jo];
  Hand Hl;          // hand pose + twist of iteration blk0 + t.tl
  float ul[NU];     // action of the step of iteration blk0 + t.tl
#pragma unroll
  for (int d = 0; d < NU; ++d) ul[d] = 0.0f;
  float uf[2] = {0.0f, 0.0f};   // finger velocity targets of the current step
  // accumulated contact impulses, kept from one sub-step to the next inside a step() for the warm start
  // (solve_contact3_acc): own corners against the support box in registers; the link / cube (slot f * 2 CPL + phs)
  // and cube / cube (slot 6 CPL + sl) contacts of this lane in shared memory, [slot][thread]
  float4 lam_st[CPL];
#pragma unroll
  for (int sl = 0; sl < CPL; ++sl) lam_st[sl] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  unsigned prev_st = 0u, prev_lk = 0u, prev_cc = 0u;   // bit per slot: in contact in the previous sub-step
  M3_DYNAMIC_SMEM(float4, slam_base);
  float4* const slam = slam_base + threadIdx.x;
  // reach with batch rows from the producer: per sample [T][2] float4 of cost ingredients behind the accumulators
  const bool defer_reach = use_refs && !producer && c.task == M3P2I_TASK_REACH;
  // link / cube contact records of the current sub-step: the list of the own cube (kRecStride float4 per list)
  float4* const srec = slam_base + 7 * CPL * blockDim.x + ((threadIdx.x / TM) * 2 + g) * kRecStride;
  float4* const sreach = slam_base + 7 * CPL * blockDim.x + (blockDim.x / TM) * 2 * kRecStride + (threadIdx.x / TM) * 2 * T;
#pragma unroll
  for (int q = 0; q < 7 * CPL; ++q) slam[q * blockDim.x] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);

  // bounding radii of the link boxes; hand + fingers lie inside a sphere of radius hrad + kGripReach about the hand
  // box centre (the fingers reach 0.12 m from it: finger joint offset + finger box + stroke)
  const float frad = sqrtf(P.finger_half[0] * P.finger_half[0] + P.finger_half[1] * P.finger_half[1] + P.finger_half[2] * P.finger_half[2]);
  const float hrad = sqrtf(P.hand_half[0] * P.hand_half[0] + P.hand_half[1] * P.hand_half[1] + P.hand_half[2] * P.hand_half[2]);
  bool was_asleep = it0 > 0; // own cube slept through the previous sub-step of this rollout
  int k0_keep = it0 > 0 ? b.far_info[g] : 0;   // its support box then
  const int n_iter = T * ns;
  // a sleeping cube's sub-step: the support carries its weight (reported forces)
  auto book_weight = [&](int k0s) {
    const float wgt = P.cube_mass[g] * P.gravity * h;
    if (k0s == P.idx_table && P.report_cube) imp_table.z -= wgt;
    if (k0s == P.idx_shelf && P.report_cube) imp_shelf.z -= wgt;
    if (g == 1) imp_cubeb.z += wgt;
  };
  // end of a step: the contact forces the cost reads, from the impulse / penalty sums of its sub-steps
  auto finish_forces = [&]() {
    V3 pr = pen;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) pr = pr + shfl3(pr, t.lane ^ o);
    const V3 pen_o = shfl3(pr, t.other);
    const V3 pen_table = g == 0 ? pr : pen_o, pen_shelf = g == 0 ? pen_o : pr;
    const V3 it_o = shfl3(imp_table, t.other), is_o = shfl3(imp_shelf, t.other), ib_o = shfl3(imp_cubeb, t.other);
    const V3 itab = g == 0 ? imp_table + it_o : it_o + imp_table;   // cubeA's share first, as in the serial order
    const V3 ishf = g == 0 ? imp_shelf + is_o : is_o + imp_shelf;
    const V3 icb = g == 1 ? imp_cubeb : ib_o;
    const float inv_dt = 1.0f / c.dt, inv_ns = 1.0f / (float)ns;
    e.f_table = inv_dt * itab + inv_ns * pen_table;
    e.f_shelf = inv_dt * ishf + inv_ns * pen_shelf;
    e.f_cubeb = inv_dt * icb;
  };
  // ---- far field (see the block in the loop): needs sub-steps that tile the run-ahead blocks
  const bool far_cfg = (ns & (ns - 1)) == 0 && ns <= TM && !(use_refs && !producer && c.task != M3P2I_TASK_REACH);
  int far_failed = -1;       // first iteration of the block in which the far-field attempt failed
#pragma unroll 1
  for (int it = it0, step = it0 / ns, s = 0; it <= n_iter; ++it, s = (s + 1 == ns ? 0 : s + 1), step += (s == 0)) {
    const bool last = it == n_iter;
    const bool handed = it0 > 0 && it == it0;   // the cost of the step that ended at the hand-over is already counted
    // Re-align the warps of the CTA once per sub-step: they then walk the same stretch of this (large) loop body at
    // about the same time and share its instruction-cache lines instead of evicting each other's.
    if (c.align) __syncthreads();
    const int lb_ = it & (TM - 1);   // position of this iteration in its block
    if (lb_ == 0) {
      // ---- perturbed action (mppi.py:392-416) of the step of the own iteration
      const int itl = it + t.tl;
      const int stepl = min(itl / ns, T - 1);
      const bool mine_live = itl < n_iter;
      if (mine_live) {
        sample_action<NU>(c, b, kg, k, stepl, ul);
        if (valid && itl == stepl * ns) {
#pragma unroll
          for (int d = 0; d < NU; ++d) b.actions[(size_t)(stepl * NU + d) * K + k] = ul[d];
        }
      }
      float ssn[7], scs[7], sqd[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) { ssn[i] = 0.0f; scs[i] = 1.0f; sqd[i] = 0.0f; }
      int stepi = step, si = s;   // step / sub-step of iteration it + l
#pragma unroll 1
      for (int l = 0; l < TM; ++l, si = (si + 1 == ns ? 0 : si + 1), stepi += (si == 0)) {
        const int iti = it + l;
        if (iti > n_iter) break;
        const bool lasti = iti == n_iter;
        const int src = t.team_base + l;
        // own joint's velocity target from the lane that drew this iteration's action
        float uj = 0.0f;
#pragma unroll
        for (int i = 0; i < 7; ++i) { const float v = __shfl_sync(kFull, ul[i], src); if (jo == i) uj = v; }
        if (!lasti) {
          // 1. joint drive of the own joint
          const float m = m_o;
          float vs = (m * vo + h * D * uj) / (m + h * D);
          const float f = D * (uj - vs);
          if (f > ef_o) vs = vo + h * ef_o / m;
          else if (f < -ef_o) vs = vo - h * ef_o / m;
          vs = clampf(vs, -vl_o, vl_o);
          if (qo <= lo_o && vs < 0.0f) vs = 0.0f;
          if (qo >= up_o && vs > 0.0f) vs = 0.0f;
          vo = vs;
        }
        float sno, cso;
        sincosf(qo, &sno, &cso);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const float a_ = __shfl_sync(kFull, sno, t.team_base + i), b_ = __shfl_sync(kFull, cso, t.team_base + i);
          const float c_ = __shfl_sync(kFull, vo, t.team_base + i);
          if (t.tl == l) { ssn[i] = a_; scs[i] = b_; sqd[i] = c_; }
        }
        if (!lasti) {
          // 4. position of the own joint
          float qn = qo + h * vo;
          if (qn < lo_o) { qn = lo_o; vo = 0.0f; }
          if (qn > up_o) { qn = up_o; vo = 0.0f; }
          qo = qn;
          // state row of a finished step (q1, qd1, q2, qd2; reactive_tamp.py:66-69): lane 0 owns joint 1 and fetches
          // joint 2 from lane 1, so every sample stores ONE float4 and the samples of a warp one contiguous segment
          if (si == ns - 1) {
            const float q2 = __shfl_sync(kFull, qo, t.team_base + 1), v2 = __shfl_sync(kFull, vo, t.team_base + 1);
            if (writer) b.states[(size_t)stepi * K + k] = make_float4(qo, vo, q2, v2);
          }
        }
      }
      fk_from_sincos(P, ssn, scs, sqd, Hl);
    }
    // ---- far field: the rest of this run-ahead block in ONE lane-parallel pass. When both cubes of every sample of the
    // warp slept through the previous sub-step, the only things that evolve are the arm (already advanced through the
    // block above) and the two fingers (a cheap recurrence nobody else acts on). Lane l then evaluates iteration blk0 + l
    // on its own -- link box centres from its own forward kinematics, the dormancy pre-tests of BOTH cubes, the
    // bounding-sphere test of the gripper against table and shelf, and the cost of the step that ends there. If every
    // iteration of the block passes every test, the serial path would have taken its dormant shortcut and skipped the
    // link penalties in each of them: the block is committed (costs added in step order) and the loop jumps to the next
    // block. Otherwise nothing is kept and the iterations run one by one as before (exact either way).
    if (far_cfg && s == 0 && !last && far_failed != it - lb_ && __all_sync(kFull, was_asleep)) {
      const int lend = min(TM, n_iter - (it - lb_));   // iterations [lb_, lend) of the block are left
      // fingers through the block (every lane the same): drive, limits, position; lane l keeps the openings iteration l sees
      float fq[2] = {e.q[7], e.q[8]}, fv[2] = {e.qd[7], e.qd[8]}, fu[2] = {uf[0], uf[1]};
      float q7s = fq[0], q8s = fq[1];
      {
        int si = 0;
#pragma unroll 1
        for (int l = lb_; l < lend; ++l, si = (si + 1 == ns ? 0 : si + 1)) {
          if (si == 0) { fu[0] = __shfl_sync(kFull, ul[7], t.team_base + l); fu[1] = __shfl_sync(kFull, ul[8], t.team_base + l); }
          if (t.tl == l) { q7s = fq[0]; q8s = fq[1]; }
#pragma unroll
          for (int jf = 0; jf < 2; ++jf) {
            const float qj = fq[jf], vj = fv[jf], uj = fu[jf];
            const float m = P.finger_mass;
            float vs = (m * vj + h * D * uj) / (m + h * D);
            const float f = D * (uj - vs);
            if (f > P.effort[7 + jf]) vs = vj + h * P.effort[7 + jf] / m;
            else if (f < -P.effort[7 + jf]) vs = vj - h * P.effort[7 + jf] / m;
            vs = clampf(vs, -P.qd_limit[7 + jf], P.qd_limit[7 + jf]);
            if (qj <= P.q_lower[7 + jf] && vs < 0.0f) vs = 0.0f;
            if (qj >= P.q_upper[7 + jf] && vs > 0.0f) vs = 0.0f;
            vs = clampf(vs, -P.qd_limit[7 + jf], P.qd_limit[7 + jf]);
            float qn = qj + h * vs;
            if (qn < P.q_lower[7 + jf]) { qn = P.q_lower[7 + jf]; vs = 0.0f; }
            if (qn > P.q_upper[7 + jf]) { qn = P.q_upper[7 + jf]; vs = 0.0f; }
            fq[jf] = qn; fv[jf] = vs;
          }
        }
      }
      const bool act = t.tl >= lb_ && t.tl < lend;          // this lane holds an iteration of the rest of the block
      const bool ends_step = act && (t.tl & (ns - 1)) == 0;  // ... the first sub-step of a step: the cost of the step before
      const int step_l = step + ((t.tl - lb_) >> (31 - __clz(ns)));
      V3 ll[3];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const float* cen = f < 2 ? P.finger_center : P.hand_center;
        V3 l = mk(cen[0], cen[1], cen[2]);
        if (f == 0) { l.y += q7s; l.z += kFingerZ; }
        if (f == 1) { l.y = -l.y - q8s; l.z += kFingerZ; }
        ll[f] = Hl.p + mul(Hl.R, l);
      }
      bool ok = true;
      {
        const V3 xo = shfl3(e.cu.p, t.other);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const V3 px = q == 0 ? e.cu.p : xo;
          const float rad = q == 0 ? rad_own : rad_oth;
          const V3 d2 = ll[2] - px;
          const float rf = frad + rad + P.contact_margin, rh = hrad + rad + P.contact_margin;
          const float rfar = rh + kGripReach, dd2 = dot(d2, d2);
          bool dorm = dd2 > rh * rh;
          if (dorm && !(dd2 > rfar * rfar)) {
            const V3 d0 = ll[0] - px, d1 = ll[1] - px;
            dorm = dot(d0, d0) > rf * rf && dot(d1, d1) > rf * rf;
          }
          ok = ok && dorm;
        }
        OBox3 gb;
        gb.c = ll[2]; gb.R = Hl.R; gb.half = mk(hrad + kGripReach, 0.0f, 0.0f);
        if (P.idx_table >= 0) ok = ok && !boxes_near(gb, obox_of(P.st[P.idx_table]), 0.0f);
        if (P.idx_shelf >= 0) ok = ok && !boxes_near(gb, obox_of(P.st[P.idx_shelf]), 0.0f);
      }
      if (__all_sync(kFull, ok || !act)) {
        // forces the previous step reported (read by the cost at lane lb_), then those of a step slept through
        const V3 ft0 = e.f_table, fs0 = e.f_shelf, fb0 = e.f_cubeb;
        imp_table = mk(0, 0, 0); imp_shelf = mk(0, 0, 0); imp_cubeb = mk(0, 0, 0); pen = mk(0, 0, 0);
        for (int q = 0; q < ns; ++q) book_weight(k0_keep);
        finish_forces();
        const int src = t.team_base;  // a lane of group 0 holds cubeA
        Cube a;
        a.p = shfl3(e.cu.p, src);
        a.qx = __shfl_sync(kFull, e.cu.qx, src); a.qy = __shfl_sync(kFull, e.cu.qy, src);
        a.qz = __shfl_sync(kFull, e.cu.qz, src); a.qw = __shfl_sync(kFull, e.cu.qw, src);
        a.v = mk(0, 0, 0); a.w = mk(0, 0, 0);
        const bool costs = ends_step && step_l > 0 && !(handed && t.tl == lb_);
        const int ps = step_l - 1;
        float cost_l = 0.0f;
        if (producer) {
          if (costs && (which == 0 || (which == 1 && c.multi_modal))) {
            PandaRef* r = b.refs + ps;
            if (which == 0) { r->cube0[0] = a.p.x; r->cube0[1] = a.p.y; r->cube0[2] = a.p.z; }
            if (which == 1 || !c.multi_modal) r->sel_axis = sel_axis_of(a);
          }
        } else if (defer_reach) {
          if (costs) {
            const ReachParts rp = reach_parts(Hl, q7s, q8s, a, c, kg);
            float4* slot = sreach + 2 * ps;
            slot[0] = make_float4(rp.ee.x, rp.ee.y, rp.ee.z, rp.min_y);
            slot[1] = make_float4(rp.dz.x, rp.dz.y, rp.dz.z, 0.0f);
          }
        } else {
          if (costs) {
            PandaRef ref;
            ref.cube0[0] = a.p.x; ref.cube0[1] = a.p.y; ref.cube0[2] = a.p.z; ref.sel_axis = sel_axis_of(a);
            const bool fst = t.tl == lb_;
            const V3 ft = fst ? ft0 : e.f_table, fs = fst ? fs0 : e.f_shelf, fb = fst ? fb0 : e.f_cubeb;
            const float fx = ft.x + 4.0f * fs.x + fb.x, fy = ft.y + 4.0f * fs.y + fb.y;
            const float motion = (fabsf(fx) + fabsf(fy)) > 0.1f ? 1000.0f : 0.0f;
            cost_l = panda_cost_from_hand(Hl, q7s, q8s, a, motion, c, kg, ref);
            if (valid) b.cost_h[(size_t)ps * K + k] = cost_l;
          }
          int st = step;
#pragma unroll 1
          for (int l = lb_; l < lend; l += ns, ++st) {
            const float cl = __shfl_sync(kFull, cost_l, t.team_base + l);
            if (st > 0 && !(handed && l == lb_)) { run += cl; J += gam * cl; gam *= c.gamma; }
          }
        }
        e.q[7] = fq[0]; e.q[8] = fq[1]; e.qd[7] = fv[0]; e.qd[8] = fv[1];
        uf[0] = fu[0]; uf[1] = fu[1];
#pragma unroll
        for (int sl = 0; sl < CPL; ++sl) lam_st[sl] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        prev_st = 0u; prev_lk = 0u; prev_cc = 0u;
        // the barriers of the iterations that are skipped (every warp of the CTA passes the same number of them)
        if (c.align) for (int q = lb_ + 1; q < lend; ++q) __syncthreads();
        const int adv = lend - lb_;
        it += adv - 1; step += adv / ns - 1; s = ns - 1;   // the loop header moves on to the first iteration after the block
        continue;
      }
      far_failed = it - lb_;
    }
    // ---- hand pose of this iteration, from the lane that ran its forward kinematics
    Hand H;
    {
      const int src = t.team_base + lb_;
      H.p = shfl3(Hl.p, src); H.v = shfl3(Hl.v, src); H.w = shfl3(Hl.w, src);
      H.R.cx = shfl3(Hl.R.cx, src); H.R.cy = shfl3(Hl.R.cy, src); H.R.cz = shfl3(Hl.R.cz, src);
      if (s == 0 && !last) {
        uf[0] = __shfl_sync(kFull, ul[7], src); uf[1] = __shfl_sync(kFull, ul[8], src);
        imp_table = mk(0, 0, 0); imp_shelf = mk(0, 0, 0); imp_cubeb = mk(0, 0, 0); pen = mk(0, 0, 0);
      }
    }
    if (!last) {
      // ---- 1. finger drives: lanes with an even / odd index integrate the left / right finger
      const int jf = t.tl & 1;
      const float qj = jf ? e.q[8] : e.q[7], vj = jf ? e.qd[8] : e.qd[7], uj = jf ? uf[1] : uf[0];
      const float m = P.finger_mass;
      float vs = (m * vj + h * D * uj) / (m + h * D);
      const float f = D * (uj - vs);
      if (f > P.effort[7 + jf]) vs = vj + h * P.effort[7 + jf] / m;
      else if (f < -P.effort[7 + jf]) vs = vj - h * P.effort[7 + jf] / m;
      vs = clampf(vs, -P.qd_limit[7 + jf], P.qd_limit[7 + jf]);
      if (qj <= P.q_lower[7 + jf] && vs < 0.0f) vs = 0.0f;
      if (qj >= P.q_upper[7 + jf] && vs > 0.0f) vs = 0.0f;
      e.qd[7] = __shfl_sync(kFull, vs, t.team_base);
      e.qd[8] = __shfl_sync(kFull, vs, t.team_base + 1);
    }

    if (s == 0 && step > 0 && !handed) {
      // ---- cost of the step that just ended (same joint positions as this FK; the drives only changed velocities)
      const int ps = step - 1;
      if (producer) {
        // rows 0 / Kg/2 of the batch for every sample's reach cost: plain stores per step, ONE fence + flag at the end
        // (the consumers finish their costs after their own loop, see defer_reach)
        if (t.tl == 0 && (which == 0 || (which == 1 && c.multi_modal))) {
          PandaRef* r = b.refs + ps;
          if (which == 0) { r->cube0[0] = e.cu.p.x; r->cube0[1] = e.cu.p.y; r->cube0[2] = e.cu.p.z; }
          if (which == 1 || !c.multi_modal) r->sel_axis = sel_axis_of(e.cu);
        }
      } else {
        const int src = t.team_base;  // a lane of group 0 holds cubeA
        Cube a;
        a.p = shfl3(e.cu.p, src);
        a.qx = __shfl_sync(kFull, e.cu.qx, src); a.qy = __shfl_sync(kFull, e.cu.qy, src);
        a.qz = __shfl_sync(kFull, e.cu.qz, src); a.qw = __shfl_sync(kFull, e.cu.qw, src);
        a.v = mk(0, 0, 0); a.w = mk(0, 0, 0);
        if (defer_reach) {
          // reach: the cost needs rows of OTHER samples (the producer's refs[ps]); keep what this sample knows now and
          // finish all T costs after the loop, when the producer is long done -- no hand-over per step
          const ReachParts rp = reach_parts(H, e.q[7], e.q[8], a, c, kg);
          if (t.tl == 0) {
            float4* slot = sreach + 2 * ps;
            slot[0] = make_float4(rp.ee.x, rp.ee.y, rp.ee.z, rp.min_y);
            slot[1] = make_float4(rp.dz.x, rp.dz.y, rp.dz.z, 0.0f);
          }
        } else {
          PandaRef ref;
          if (use_refs) ref = ref_wait(b, ps, c.epoch, c.multi_modal != 0);
          else { ref.cube0[0] = a.p.x; ref.cube0[1] = a.p.y; ref.cube0[2] = a.p.z; ref.sel_axis = sel_axis_of(a); }
          const float fx = e.f_table.x + 4.0f * e.f_shelf.x + e.f_cubeb.x, fy = e.f_table.y + 4.0f * e.f_shelf.y + e.f_cubeb.y;
          const float motion = (fabsf(fx) + fabsf(fy)) > 0.1f ? 1000.0f : 0.0f;
          const float cost = panda_cost_from_hand(H, e.q[7], e.q[8], a, motion, c, kg, ref);
          run += cost;
          J += gam * cost;
          gam *= c.gamma;
          if (writer) b.cost_h[(size_t)ps * K + k] = cost;
        }
      }
    }
    if (last) break;

    // ---- 2. geometry of this sub-step (positions are fixed until step 5)
    if (s == 0) { prev_st = 0u; prev_lk = 0u; prev_cc = 0u; }   // a step() starts cold (no warm start across steps)
    V3 lc[3];  // centres of the left-finger, right-finger and hand boxes
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const float* cen = f < 2 ? P.finger_center : P.hand_center;
      V3 l = mk(cen[0], cen[1], cen[2]);
      if (f == 0) { l.y += e.q[7]; l.z += kFingerZ; }
      if (f == 1) { l.y = -l.y - e.q[8]; l.z += kFingerZ; }
      lc[f] = H.p + mul(H.R, l);
    }
    const V3 fhalf = mk(P.finger_half[0], P.finger_half[1], P.finger_half[2]);
    const V3 hhalf = mk(P.hand_half[0], P.hand_half[1], P.hand_half[2]);
    float slide[2] = {e.qd[7], e.qd[8]};
    // ---- dormant cubes: a cube that slept through the previous sub-step (velocity zero, supported, nothing near) and
    // that none of the three link boxes can reach (their centre-distance pre-tests of the full path fail) sleeps through
    // this one too -- every test of the full path below would come out as before (nothing of the cube moved). When that
    // holds for every cube of the warp, all the cube work of the sub-step is skipped; only the weight the support
    // carries is booked.
    bool asleep = true;
    bool dormant;
    {
      // the centre-distance pre-tests of the full path (see lnear below), all three failing
      const V3 d2 = lc[2] - e.cu.p;
      const float rf = frad + rad_own + P.contact_margin, rh = hrad + rad_own + P.contact_margin;
      const float rfar = rh + kGripReach, dd2 = dot(d2, d2);
      dormant = was_asleep && dd2 > rh * rh;
      if (dormant && !(dd2 > rfar * rfar)) {   // inside the gripper's bounding sphere: the two finger boxes decide
        const V3 d0 = lc[0] - e.cu.p, d1 = lc[1] - e.cu.p;
        dormant = dot(d0, d0) > rf * rf && dot(d1, d1) > rf * rf;
      }
    }
    if (__all_sync(kFull, dormant)) {
      book_weight(k0_keep);
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl) lam_st[sl] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      prev_st = 0u; prev_lk = 0u; prev_cc = 0u;
    } else {
    OBox3 cb;  // own cube
    cb.c = e.cu.p; cb.R = quat_to_R(e.cu.qx, e.cu.qy, e.cu.qz, e.cu.qw); cb.half = half_own;
    // keep the nine entries as values: under register pressure ptxas otherwise re-derives them from the quaternion
    // inside the contact loops (26 instructions per use, 12 % of all instructions in a grasp state)
    M3_PIN_VALUES9(cb.R.cx.x, cb.R.cx.y, cb.R.cx.z, cb.R.cy.x, cb.R.cy.y, cb.R.cy.z, cb.R.cz.x, cb.R.cz.y, cb.R.cz.z);
    V3 v = e.cu.v, w = e.cu.w;
    const V3 x = e.cu.p;
    // own corners, as lever arms about the cube centre (corner = x + ra); slot s is corner t.c + s * G
    V3 ra[CPL];
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) ra[sl] = corner_arm(cb, t.c + sl * G);
    // the other cube of the sample (for cube-cube contact)
    const V3 xo = shfl3(x, t.other);
    bool cc_near;
    {
      const V3 d = xo - x;
      const float r = rad_own + rad_oth + P.contact_margin;
      cc_near = dot(d, d) <= r * r;  // cheap symmetric pre-test; the exact test follows if it passes
    }
    OBox3 ob;
    ob.c = xo; ob.half = half_oth; ob.R = cb.R;
    if (__any_sync(kFull, cc_near)) {
      ob.R.cx = shfl3(cb.R.cx, t.other); ob.R.cy = shfl3(cb.R.cy, t.other); ob.R.cz = shfl3(cb.R.cz, t.other);
      // exact test of the thread-per-sample code: sphere of cubeA against the box of cubeB
      cc_near = cc_near && (g == 0 ? boxes_near(cb, ob, P.contact_margin) : boxes_near(ob, cb, P.contact_margin));
    }
    // which fixed boxes / link boxes are close to the own cube, decided once per sub-step:
    // lane c of a group tests fixed boxes c, c + G, ... (n_static <= 8); one ballot per slot collects the verdicts
    unsigned near_mask = 0u;
    bool near_c[CPL];
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) {
      const int kb = t.c + sl * G;
      near_c[sl] = kb < P.n_static && boxes_near(cb, obox_of(P.st[min(kb, P.n_static - 1)]), P.contact_margin);
      const unsigned bal = __ballot_sync(kFull, near_c[sl]);
      near_mask |= ((bal >> t.group_base) & ((1u << G) - 1u)) << (sl * G);
    }
    unsigned lnear = 0u;
#pragma unroll 1
    for (int f = 0; f < 3; ++f) {
      OBox3 lb;
      lb.c = f == 0 ? lc[0] : (f == 1 ? lc[1] : lc[2]); lb.R = H.R; lb.half = f < 2 ? fhalf : hhalf;
      // centre-distance pre-test (a superset of the exact sphere-vs-box test below)
      const V3 dl = lb.c - cb.c;
      const float rr = (f < 2 ? frad : hrad) + rad_own + P.contact_margin;
      if (dot(dl, dl) > rr * rr) continue;
      if (boxes_near(lb, cb, P.contact_margin)) lnear |= 1u << f;
    }
    // geometry-only part of the own corners' contacts with the cube's FIRST near fixed box, its support (StaticHit):
    // these are the contacts that keep accumulators; the passes below reuse them
    const float inv_h = __frcp_rn(h);
    const int k0 = near_mask ? __ffs(near_mask) - 1 : 0;
    StaticHit sh0[CPL];
    {
      const OBox3 sb = obox_of(P.st[k0]);
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl) sh0[sl] = static_hit(near_mask != 0u, x + ra[sl], ra[sl], sb, im, ii, inv_h, P);
    }
    // ---- 3. sleeping (decided per cube, i.e. per group): an (almost) motionless cube that rests on its first near
    // fixed box with at least three corners and has no link and no other cube within the contact margin is neither
    // moved nor solved in this sub-step
    asleep = false;
    {
      int sup = 0;
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl)
        sup += __popc((__ballot_sync(kFull, sh0[sl].sup) >> t.group_base) & ((1u << G) - 1u));
      asleep = P.sleep_lin > 0.0f && dot(v, v) < P.sleep_lin * P.sleep_lin && dot(w, w) < P.sleep_ang * P.sleep_ang &&
               !cc_near && lnear == 0u && near_mask != 0u && sup >= 3;
    }
    if (asleep) {
      v = mk(0, 0, 0); w = mk(0, 0, 0);
      book_weight(k0);
#pragma unroll
      for (int sl = 0; sl < CPL; ++sl) sh0[sl].hit = false;
    } else {
      v.z -= P.gravity * h;   // gravity
    }
    // fixed boxes any awake cube of the warp is near (the loops below must be warp-uniform)
    unsigned near_any = 0u;
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) near_any |= fold<G>(__ballot_sync(kFull, near_c[sl] && !asleep)) << (sl * G);
    if (asleep) near_mask = 0u;
    // accumulators of the own corners against box k0: warm start from the previous sub-step of this step
    unsigned cur_st = 0u, cur_lk = 0u, cur_cc = 0u;
#pragma unroll
    for (int sl = 0; sl < CPL; ++sl) {
      lam_st[sl] = warm_prepare(lam_st[sl], sh0[sl].hit && ((prev_st >> sl) & 1u), sh0[sl].n, 0.5f * (mu_c + P.st[k0].mu), P.warm_start);
      if (sh0[sl].hit) cur_st |= 1u << sl;
    }
    // ---- link / cube contacts of this sub-step: detection ONCE, by all corner lanes, into the record list of the own
    // cube (shared memory). Both groups detect at the same time (group g: link f against cube g); the serial order
    // only matters for the solves below.
    int nrec = 0;                  // records in the list of the own cube
    unsigned nz = 0u;              // bit j: the accumulator of record j of that list holds something
    unsigned nmax[2] = {0u, 0u};   // warp-uniform: longest list of cubeA / cubeB in the warp
    const bool any_link = __any_sync(kFull, lnear != 0u);
    if (any_link) {
      __syncwarp();   // the accumulators below were last written by the group's solves of the previous sub-step
      int cnt = 0;
#pragma unroll 1
      for (int f = 0; f < 3; ++f) {
        const bool mine = (lnear >> f) & 1u;
        if (__any_sync(kFull, mine)) {
          OBox3 lb;
          lb.c = f == 0 ? lc[0] : (f == 1 ? lc[1] : lc[2]); lb.R = H.R; lb.half = f < 2 ? fhalf : hhalf;
          const V3 axis = f == 0 ? H.R.cy : (f == 1 ? -H.R.cy : mk(0, 0, 0));
          const float ims = f < 2 ? 1.0f / P.finger_mass : 0.0f;
          const float mu = 0.5f * (P.robot_mu + mu_c);
#pragma unroll 1
          for (int phs = 0; phs < 2 * CPL; ++phs) {
            // ph 0: corners of the link box in the cube, normal out of the cube;
            // ph 1: corners of the cube in the link box, normal out of the link -> solve with -n
            const int ph = phs / CPL, sl = phs - ph * CPL;
            const V3 pt = ph == 0 ? box_corner(lb, t.c + sl * G) : x + ((CPL > 1 && sl > 0) ? ra[CPL - 1] : ra[0]);
            OBox3 bx;
            bx.c = sel3(ph == 0, cb.c, lb.c); bx.half = sel3(ph == 0, cb.half, lb.half);
            bx.R.cx = sel3(ph == 0, cb.R.cx, lb.R.cx); bx.R.cy = sel3(ph == 0, cb.R.cy, lb.R.cy);
            bx.R.cz = sel3(ph == 0, cb.R.cz, lb.R.cz);
            V3 n = mk(0, 0, 0);
            float depth = 0.0f;
            const bool hit = mine && point_in_box(pt, bx, P.contact_margin, n, depth);
            const unsigned hb = __ballot_sync(kFull, hit);
            if (!hb) continue;
            const unsigned gb = (hb >> t.group_base) & ((1u << G) - 1u);   // hits of the own group, bit = lane in group
            const int pos = cnt + __popc(gb & ((1u << t.c) - 1u));
            cnt += __popc(gb);
            if (hit && pos < kLinkCap) {
              const float sg = ph == 0 ? 1.0f : -1.0f;
              const LinkHit lh = link_hit(hit, sg * n, depth, pt, H.v, H.w, H.p, axis, ims, x, im, ii, inv_h, P);
              // accumulator of this lane's slot (f, direction, corner slot) of the own cube: warm start from the
              // previous sub-step of this step, applied by the slot's first visit below
              const int slot = f * (2 * CPL) + phs;
              float4* Ls = slam + slot * blockDim.x;
              const float4 Lw = warm_prepare(*Ls, (prev_lk >> slot) & 1u, lh.n, mu, P.warm_start);
              *Ls = Lw;
              if (Lw.x != 0.0f || Lw.y != 0.0f || Lw.z != 0.0f || Lw.w != 0.0f) nz |= 1u << pos;
              cur_lk |= 1u << slot;
              float4* r = srec + 4 * pos;
              r[0] = make_float4(lh.n.x, lh.n.y, lh.n.z, lh.target);
              r[1] = make_float4(lh.rc.x, lh.rc.y, lh.rc.z, lh.an);
              r[2] = make_float4(lh.vl0.x, lh.vl0.y, lh.vl0.z, lh.ikn);
              r[3] = make_float4(lh.rcn.x, lh.rcn.y, lh.rcn.z, __int_as_float((f << 24) | (slot * (int)blockDim.x + (int)threadIdx.x)));
            }
          }
          cnt = min(cnt, kLinkCap);
        }
      }
      // longest list of cubeA / cubeB in the warp: the bounds of the (warp-uniform) loops of the sweeps
      nrec = cnt;
#pragma unroll
      for (int o = 1; o < G; o <<= 1) nz |= __shfl_xor_sync(kFull, nz, o);   // the group's records, known to its first lane
      nmax[0] = __reduce_max_sync(kFull, g == 0 ? (unsigned)cnt : 0u);
      nmax[1] = __reduce_max_sync(kFull, g == 1 ? (unsigned)cnt : 0u);
      __syncwarp();   // records and prepared accumulators are read by the other lanes of the group
    }

#pragma unroll 1
    for (int p = 0; p < c.passes; ++p) {
      // (a) links against the cubes. The serial order of the model is (f, cubeA), (f, cubeB) for f = 0, 1, 2 in every
      // sweep; cubeA's whole list followed by cubeB's gives the same result (contacts of different cubes only meet in
      // a finger's sliding speed, and those of the same finger keep their order). The first lane of the cube's group
      // walks the list (it alone reads and writes the accumulators of the records); the cube's new velocity and the
      // fingers' sliding speeds go back to the replicas afterwards. P.link_sweeps sweeps: the finger - cube - finger
      // chain of a grasp only settles after a few sweeps over its own contacts.
      if (nmax[0] | nmax[1]) {
        const float ims_f = 1.0f / P.finger_mass, mu = 0.5f * (P.robot_mu + mu_c);
#pragma unroll 1
        for (int sw = 0; sw < P.link_sweeps; ++sw) {
          const bool first = p == 0 && sw == 0;
          if (nmax[0]) {
            const int lead = t.team_base;
            (void)walk_link_records<false>(srec, t.lane == lead ? nrec : 0, nz, slam_base, H.R.cy, slide, ims_f, v, w, im, ii, mu, first);
            const V3 vl = shfl3(v, lead), wl = shfl3(w, lead);
            if (g == 0) { v = vl; w = wl; }
            slide[0] = __shfl_sync(kFull, slide[0], lead); slide[1] = __shfl_sync(kFull, slide[1], lead);
          }
          if (nmax[1]) {
            const int lead = t.team_base + G;
            V3 got = walk_link_records<true>(srec, t.lane == lead ? nrec : 0, nz, slam_base, H.R.cy, slide, ims_f, v, w, im, ii, mu, first);
            const V3 vl = shfl3(v, lead), wl = shfl3(w, lead);
            got = shfl3(got, lead);
            if (g == 1) { v = vl; w = wl; imp_cubeb = imp_cubeb - got; }
            slide[0] = __shfl_sync(kFull, slide[0], lead); slide[1] = __shfl_sync(kFull, slide[1], lead);
          }
        }
      }
      // (b) cubeA against cubeB, both ways; every lane of the team applies every impulse to replicas of both cubes
      if (__any_sync(kFull, cc_near)) {
        const V3 vo = shfl3(v, t.other), wo = shfl3(w, t.other);
        Dyn3 A = g == 0 ? dyn_cube(v, w, x, im, ii) : dyn_cube(vo, wo, xo, imo, iio);   // cubeA
        Dyn3 B = g == 0 ? dyn_cube(vo, wo, xo, imo, iio) : dyn_cube(v, w, x, im, ii);   // cubeB
        const float mu = 0.5f * (P.cube_mu[0] + P.cube_mu[1]);
        V3 got = mk(0, 0, 0);
#pragma unroll 1
        for (int phs = 0; phs < 2 * CPL; ++phs) {
          // ph 0: corners of cubeA in cubeB (found by group 0), normal out of cubeB;
          // ph 1: corners of cubeB in cubeA (found by group 1), normal out of cubeA -> solve with -n
          const int ph = phs / CPL, sl = phs - ph * CPL;
          const V3 pc = x + ((CPL > 1 && sl > 0) ? ra[CPL - 1] : ra[0]);
          V3 n = mk(0, 0, 0);
          float depth = 0.0f;
          const bool hit = cc_near && g == ph && point_in_box(pc, ob, P.contact_margin, n, depth);
          const float sg = ph == 0 ? 1.0f : -1.0f;
          // accumulator of the own corner slot against the other cube (the group that finds the contact owns it)
          float4* Ls = slam + (6 * CPL + sl) * blockDim.x;
          float4 Lv = *Ls;
          if (p == 0) {
            Lv = warm_prepare(Lv, hit && ((prev_cc >> sl) & 1u), sg * n, mu, P.warm_start);
            if (hit) cur_cc |= 1u << sl;
          }
          got = got + apply_hits<G>(hit, n, depth, pc, Lv, t.team_base + G * ph, cc_near, [&](V3 nj, float dj, V3 pj, float4& Lj) {
            return solve_contact3_call(A, B, sg * nj, dj, pj, mu, h, P, Lj, p == 0);
          });
          if (hit) *Ls = Lv;
        }
        if (cc_near) {
          v = g == 0 ? A.v : B.v;
          w = g == 0 ? A.w : B.w;
          if (g == 1) imp_cubeb = imp_cubeb - got;  // cubeB received -(impulse on cubeA)
        }
      }
      // (c) own cube against the fixed boxes, LAST in the pass: what a kinematic link pushes into the table is pushed
      // back out by the table in the same pass
#pragma unroll 1
      for (unsigned todo_k = near_any; todo_k; todo_k &= todo_k - 1) {
        const int ks = __ffs(todo_k) - 1;
        const OBox3 sb = obox_of(P.st[ks]);
        const bool near = (near_mask >> ks) & 1u;
        const bool keep = near && ks == k0;
        const float mu = 0.5f * (mu_c + P.st[ks].mu);
        V3 got = mk(0, 0, 0);
#pragma unroll 1
        for (int sl = 0; sl < CPL; ++sl) {
          const V3 r = (CPL > 1 && sl > 0) ? ra[CPL - 1] : ra[0];
          StaticHit sh;
          if (keep) sh = (CPL > 1 && sl > 0) ? sh0[CPL - 1] : sh0[0];
          else sh = static_hit(near, x + r, r, sb, im, ii, inv_h, P);
          float4 Lr = (CPL > 1 && sl > 0) ? lam_st[CPL - 1] : lam_st[0];
          got = got + apply_static_hits<G>(sh, r, t.group_base, v, w, im, ii, mu, Lr, keep, p == 0);
          if (CPL > 1 && sl > 0) lam_st[CPL - 1] = Lr; else lam_st[0] = Lr;
        }
        // impulses received by the fixed box = -(impulses on the cube)
        if (ks == P.idx_table && P.report_cube) imp_table = imp_table - got;
        if (ks == P.idx_shelf && P.report_cube) imp_shelf = imp_shelf - got;
        if (g == 1) imp_cubeb = imp_cubeb + got;
      }
    }
    prev_st = cur_st; prev_lk = cur_lk; prev_cc = cur_cc;
    e.cu.v = v; e.cu.w = w;
    k0_keep = k0;
    }   // not dormant
    was_asleep = asleep;
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
        // one bounding sphere for hand + fingers first (centre = hand box centre, radius covers the three boxes)
        OBox3 gb;
        gb.c = lc[2]; gb.R = H.R; gb.half = mk(hrad + kGripReach, 0.0f, 0.0f);
        const bool grip_near = boxes_near(gb, sb, 0.0f);
#pragma unroll 1
        for (int f = 0; f < 3 && grip_near; ++f) {
          OBox3 lb;
          lb.c = f == 0 ? lc[0] : (f == 1 ? lc[1] : lc[2]); lb.R = H.R; lb.half = f < 2 ? fhalf : hhalf;
          if (!boxes_near(lb, sb, 0.0f)) continue;
#pragma unroll 1
          for (int sl = 0; sl < CPL; ++sl) {
            const V3 lp = box_corner(lb, t.c + sl * G);
            V3 n; float depth;
            if (!point_in_box(lp, sb, 0.0f, n, depth)) continue;
            const float fn = P.penalty_stiffness * depth;
            V3 vel = H.v + cross(H.w, lp - H.p);
            if (f == 0) vel = vel + slide[0] * H.R.cy;
            if (f == 1) vel = vel - slide[1] * H.R.cy;
            const float vn = dot(vel, n);
            const V3 tv = vel - vn * n;
            const float vt = sqrtf(dot(tv, tv));
            V3 fo = (-fn) * n;
            if (vt > 1e-6f) fo = fo + (mu * fn / vt) * tv;
            pen = pen + fo;
          }
        }
      }
    }
    // ---- 5. positions
#pragma unroll
    for (int j = 7; j < 9; ++j) {   // the arm joints were integrated by the run-ahead
      float qn = e.q[j] + h * e.qd[j];
      if (qn < P.q_lower[j]) { qn = P.q_lower[j]; e.qd[j] = 0.0f; }
      if (qn > P.q_upper[j]) { qn = P.q_upper[j]; e.qd[j] = 0.0f; }
      e.q[j] = qn;
    }
    if (!asleep) {
      Cube& cu = e.cu;
      cu.p = cu.p + h * cu.v;
      const float qx = cu.qx, qy = cu.qy, qz = cu.qz, qw = cu.qw, hh = 0.5f * h;
      const float nx = qx + hh * (cu.w.x * qw + cu.w.y * qz - cu.w.z * qy);
      const float ny = qy + hh * (cu.w.y * qw + cu.w.z * qx - cu.w.x * qz);
      const float nz = qz + hh * (cu.w.z * qw + cu.w.x * qy - cu.w.y * qx);
      const float nw = qw - hh * (cu.w.x * qx + cu.w.y * qy + cu.w.z * qz);
      const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
      cu.qx = nx * inv; cu.qy = ny * inv; cu.qz = nz * inv; cu.qw = nw * inv;
    }
    if (s == ns - 1) {
      // ---- end of the step: reported contact forces
      finish_forces();
    }
  }
  if (producer) {
    if (t.tl == 0 && (which == 0 || (which == 1 && c.multi_modal))) {
      __threadfence();
      *(volatile unsigned*)(b.ref_flags + which) = c.epoch + (unsigned)T;
    }
    return;
  }
  if (defer_reach) {
    // the T reach costs, now that the rows of the batch they read are published: ONE wait for the producer's last
    // step (flags count steps), then refs[0..T) are plain loads. The lanes of a team share the steps (lane l takes
    // steps l, l + TM, ...: cost ingredients from shared memory, the row of the batch from L2) and add up their parts.
    if (t.tl == 0) (void)ref_wait(b, T - 1, c.epoch, c.multi_modal != 0);
    __syncwarp();
    // (after a hand-over the steps before it0 are finished: their costs were added above, gam = gamma ^ (it0 / ns))
    float g_l = gam, g_stride = 1.0f, run_l = 0.0f, J_l = 0.0f;
    for (int q = 0; q < TM; ++q) { if (q < t.tl) g_l *= c.gamma; g_stride *= c.gamma; }
#pragma unroll 2
    for (int ps = it0 / ns + t.tl; ps < T; ps += TM) {
      PandaRef ref;
      ref.cube0[0] = __ldcg(&b.refs[ps].cube0[0]); ref.cube0[1] = __ldcg(&b.refs[ps].cube0[1]);
      ref.cube0[2] = __ldcg(&b.refs[ps].cube0[2]); ref.sel_axis = __ldcg(&b.refs[ps].sel_axis);
      const float4 p0 = sreach[2 * ps], p1 = sreach[2 * ps + 1];
      ReachParts rp;
      rp.ee = mk(p0.x, p0.y, p0.z); rp.min_y = p0.w; rp.dz = mk(p1.x, p1.y, p1.z);
      const float cost = reach_combine(rp, c, kg, ref);
      run_l += cost;
      J_l += g_l * cost;
      g_l *= g_stride;
      if (valid) b.cost_h[(size_t)ps * K + k] = cost;
    }
#pragma unroll
    for (int o = TM / 2; o > 0; o >>= 1) {
      run_l += __shfl_xor_sync(kFull, run_l, o);
      J_l += __shfl_xor_sync(kFull, J_l, o);
    }
    run += run_l;
    J += J_l;
  }
  if (writer) {
    b.cost_sum[k] = run;
    b.J[k] = J;
    if (b.peer.n) push_J_store(b.peer, c.offset, k, J);   // committed per CTA in team_kernel_body
  }
  if (c.store_env) {
    // arm joints back from their owner lanes (warp-uniform branch: every team of the launch stores or none does)
#pragma unroll
    for (int i = 0; i < 7; ++i) { e.q[i] = __shfl_sync(kFull, qo, t.team_base + i); e.qd[i] = __shfl_sync(kFull, vo, t.team_base + i); }
    if (valid) {
      e.store(b.env, K, k, t);
      if (t.tl == ((n_iter - 1) & (TM - 1))) {   // the lane that drew the action of the last step
#pragma unroll
        for (int d = 0; d < NU; ++d) b.vel_target[(size_t)d * K + k] = ul[d];
      }
    }
  }
}

// Body of k_rollout_team: which sample (or producer row) this thread's team works on.
template <int CPL>
DEV void team_kernel_body(const RolloutCfg& c, const PandaParams& P, const RolloutBufs& b) {
  constexpr int TM = TeamShape<CPL>::kTeam;
  const TeamLane t = team_lane<CPL>();
  const bool use_refs = b.refs != nullptr;
  const bool producer = use_refs && blockIdx.x == 0;
  const int which = t.lane / TM;   // producer CTA: team 0 replays global row 0, team 1 global row Kg/2
  // after k_rollout_far: only the listed samples (the far-field ones are finished); CTAs without work leave at once
  int count = c.K;
  if (b.near_list) count = __ldcg(b.near_count);
  const int cta_first = ((blockIdx.x - (use_refs ? 1 : 0)) * blockDim.x) / TM;
  if (count <= 0 || (!producer && cta_first >= count)) return;
  if (c.near_team_max > 0 && count > c.near_team_max) return;   // the thread-per-sample kernel takes this one
  if (producer && b.near_list && __ldcg(b.far_info + 2)) return;   // k_rollout_far published the rows (they stayed far)
  const int kraw = ((blockIdx.x - (use_refs ? 1 : 0)) * blockDim.x + threadIdx.x) / TM;
  const bool valid = !producer && kraw < count;
  int k = kraw < count ? kraw : count - 1;   // idle teams shadow the last sample so that every shuffle has 32 lanes
  int it_start = 0;
  if (b.near_list && !producer) {
    // listed sample: row and hand-over boundary (panda_far.cuh). The warp starts at the earliest boundary of its
    // samples, rounded down to a run-ahead block; a sample handed over later just repeats a few far-field iterations.
    const int entry = b.near_list[k];
    k = entry & ((1 << kFarRowBits) - 1);
    const unsigned bd = __reduce_min_sync(kFull, (unsigned)(entry >> kFarRowBits));
    it_start = ((int)bd * 8) & ~(TM - 1);
  }
  int kg = c.offset + k;
  if (producer) {
    kg = (which == 1 && c.multi_modal) ? c.Kg / 2 : 0;
    k = (kg >= c.offset && kg < c.offset + c.K) ? kg - c.offset : -1;
  }
  team_rollout<CPL>(c, P, b, t, k, kg, valid, producer, which, it_start);
  if (b.peer.n && !producer) {
    // sharded over peer memory: the CTA's J values are in every mailbox; one thread orders them and counts them in
    __syncthreads();
    if (threadIdx.x == 0) {
      const int first = cta_first;
      const int cnt = min((int)blockDim.x / TM, count - first);
      if (cnt > 0) push_J_commit(b.peer, c.K, (unsigned)cnt);
    }
  }
}

}  // namespace m3
