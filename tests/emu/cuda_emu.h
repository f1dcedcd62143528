// cuda_emu.h -- just enough of the CUDA device environment to compile the team rollout's device code
// (m3p2i-aip_b200/csrc/*.cuh) as HOST code and run it warp by warp in lock step. TEST INFRASTRUCTURE ONLY.
//
// Each of the 32 lanes of a warp is a fiber (ucontext). A warp collective (__shfl_sync, __ballot_sync, __any_sync)
// publishes the lane's operand, yields, and reads the other lanes' operands once every lane has published: the
// scheduler resumes the lanes round-robin, so when a lane is resumed all 32 have reached the same collective (two
// operand buffers, alternating, keep a fast lane from overwriting what a slow one still has to read). Every collective
// records its source line; lanes that meet at DIFFERENT collectives (a non-uniform branch around a shuffle -- undefined
// behaviour on the GPU) abort the run with both line numbers. Warps run one after the other; __syncthreads is
// therefore not available (the emulator runs the kernel with RolloutCfg::align = 0).
#pragma once
#define M3_EMU 1
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
typedef void* cudaStream_t;

namespace emu {
struct Idx { unsigned x, y, z; };
struct Lane {
  ucontext_t ctx;
  char* stack;
  bool done;
  unsigned n_coll;
  Idx tid;
};
struct Warp {
  Lane lane[32];
  ucontext_t sched;
  int cur;
  uint32_t val[2][32];
  int site[2][32];
  Idx bid, bdim;
  void* smem;
  long collectives;
};
extern Warp* W;
inline Lane& me() { return W->lane[W->cur]; }
inline void publish(uint32_t v, int site) {
  Lane& L = me();
  const int par = L.n_coll & 1;
  W->val[par][W->cur] = v;
  W->site[par][W->cur] = site;
  swapcontext(&L.ctx, &W->sched);
  for (int i = 0; i < 32; ++i)
    if (W->site[par][i] != site) {
      fprintf(stderr, "emu: divergent warp collective: lane %d at line %d, lane %d at line %d\n", W->cur, site, i, W->site[par][i]);
      abort();
    }
}
inline uint32_t fetch(int src) {
  Lane& L = me();
  return W->val[L.n_coll & 1][src & 31];
}
inline void finish() { ++me().n_coll; ++W->collectives; }
template <typename T>
inline T shfl(T v, int src, int site) {
  static_assert(sizeof(T) == 4, "32-bit operands only");
  uint32_t u;
  memcpy(&u, &v, 4);
  publish(u, site);
  u = fetch(src);
  finish();
  T r;
  memcpy(&r, &u, 4);
  return r;
}
inline unsigned ballot(bool p, int site) {
  publish(p ? 1u : 0u, site);
  unsigned m = 0u;
  for (int i = 0; i < 32; ++i) m |= (fetch(i) & 1u) << i;
  finish();
  return m;
}
inline unsigned reduce_max(unsigned v, int site) {
  publish(v, site);
  unsigned m = 0u;
  for (int i = 0; i < 32; ++i) m = std::max(m, (unsigned)fetch(i));
  finish();
  return m;
}
}  // namespace emu

#define threadIdx (emu::me().tid)
#define blockIdx (emu::W->bid)
#define blockDim (emu::W->bdim)
#define __shfl_sync(mask, v, src) emu::shfl((v), (src), __LINE__)
#define __shfl_xor_sync(mask, v, x) emu::shfl((v), (int)(emu::W->cur ^ (x)), __LINE__)
#define __ballot_sync(mask, p) emu::ballot((p), __LINE__)
#define __any_sync(mask, p) (emu::ballot((p), __LINE__) != 0u)
#define __all_sync(mask, p) (emu::ballot((p), __LINE__) == 0xffffffffu)
#define __syncwarp() ((void)emu::ballot(true, __LINE__))
#define __reduce_max_sync(mask, v) emu::reduce_max((v), __LINE__)
#define __reduce_min_sync(mask, v) (~emu::reduce_max(~(unsigned)(v), __LINE__))
#define __syncthreads() do { fprintf(stderr, "emu: __syncthreads is not emulated (run with align = 0)\n"); abort(); } while (0)
#define M3_PIN_VALUES9(a, b, c, d, e, f, g, h, i) do { } while (0)
#define M3_DYNAMIC_SMEM(type, name) type* name = static_cast<type*>(emu::W->smem)

static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __float_as_int(float a) { int r; memcpy(&r, &a, 4); return r; }
static inline float __int_as_float(int a) { float r; memcpy(&r, &a, 4); return r; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
extern "C" void sincosf(float, float*, float*);
