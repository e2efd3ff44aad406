// tests/warp_emu.hpp — a 32-lane SIMT warp emulated on the CPU, so that the engine's warp-group traversal (csrc/bvh.cuh,
// compiled by g++ with -DB2R_WARP_EMU) can be executed and checked without a GPU.  Test infrastructure only.
//
// Each lane is a ucontext fiber; a warp-wide primitive (__ballot_sync, __shfl_sync, __reduce_*_sync, ...) stores the lane's
// operand and yields to a round-robin scheduler, so by the time the lane is resumed every lane has stored its operand: the lanes
// advance in lockstep from collective to collective, which is exactly the guarantee the device code relies on (all its
// collectives sit in warp-uniform control flow with a full mask).  Operands are double-buffered by collective parity, so a
// fast lane's NEXT collective never overwrites a slot a slower lane still has to read.
#pragma once
#include <ucontext.h>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <functional>
#include <vector>

namespace wemu {
constexpr int W = 32;
struct Warp {
  ucontext_t sched, lane[W];
  std::vector<char> stack[W];
  bool done[W];
  int cur = 0;
  unsigned long long slot[2][W];
  unsigned long long ncoll[W];
  std::function<void(int)> body;
  unsigned base_thread = 0;
};
inline Warp*& g() { static Warp* w = nullptr; return w; }
}  // namespace wemu

// ---- the CUDA built-ins the traversal uses, as host functions (global namespace, like the real ones)
struct WemuIdx { unsigned x = 0, y = 0, z = 0; };
static WemuIdx threadIdx, blockIdx;
static WemuIdx blockDim{128, 1, 1}, gridDim{1, 1, 1};

namespace wemu {
inline void trampoline() {
  Warp* w = g();
  const int l = w->cur;
  w->body(l);
  w->done[l] = true;
  swapcontext(&w->lane[l], &w->sched);
}
// run body(lane) on 32 lockstep lanes; thread index of lane l = base_thread + l
inline void run_warp(unsigned base_thread, std::function<void(int)> body) {
  Warp w;
  g() = &w;
  w.body = std::move(body);
  w.base_thread = base_thread;
  for (int l = 0; l < W; l++) {
    w.done[l] = false;
    w.ncoll[l] = 0;
    w.stack[l].resize(512 * 1024);
    getcontext(&w.lane[l]);
    w.lane[l].uc_stack.ss_sp = w.stack[l].data();
    w.lane[l].uc_stack.ss_size = w.stack[l].size();
    w.lane[l].uc_link = &w.sched;
    makecontext(&w.lane[l], (void (*)())trampoline, 0);
  }
  bool any = true;
  while (any) {
    any = false;
    for (int l = 0; l < W; l++) {
      if (w.done[l]) continue;
      any = true;
      w.cur = l;
      threadIdx.x = base_thread + l;
      swapcontext(&w.sched, &w.lane[l]);
    }
  }
  // SIMT sanity: every lane must have gone through the same number of collectives
  for (int l = 1; l < W; l++)
    if (w.ncoll[l] != w.ncoll[0]) { fprintf(stderr, "warp_emu: lanes diverged around a collective (%llu vs %llu)\n", w.ncoll[l], w.ncoll[0]); abort(); }
  g() = nullptr;
}
// the one primitive: every lane contributes 64 bits and sees everybody's
inline void exchange(unsigned long long v, unsigned long long out[W]) {
  Warp* w = g();
  const int l = w->cur;
  const int b = (int)(w->ncoll[l]++ & 1ull);
  w->slot[b][l] = v;
  swapcontext(&w->lane[l], &w->sched);
  threadIdx.x = w->base_thread + l;
  for (int i = 0; i < W; i++) out[i] = w->slot[b][i];
}
inline int lane_id() { return g()->cur; }
template <class T> inline unsigned long long to_bits(T v) { unsigned long long u = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(unsigned long long u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }
}  // namespace wemu

inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned long long o[wemu::W];
  wemu::exchange(pred ? 1ull : 0ull, o);
  unsigned m = 0;
  for (int i = 0; i < wemu::W; i++) if (o[i]) m |= 1u << i;
  return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
  unsigned long long o[wemu::W];
  wemu::exchange(wemu::to_bits(v), o);
  return wemu::from_bits<T>(o[src & 31]);
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  unsigned long long o[wemu::W];
  const int me = wemu::lane_id();
  wemu::exchange(wemu::to_bits(v), o);
  return wemu::from_bits<T>(o[(me ^ lane_mask) & 31]);
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  unsigned long long o[wemu::W];
  wemu::exchange(v, o);
  unsigned r = 0;
  for (int i = 0; i < wemu::W; i++) r |= (unsigned)o[i];
  return r;
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  unsigned long long o[wemu::W];
  wemu::exchange(v, o);
  unsigned r = 0xffffffffu;
  for (int i = 0; i < wemu::W; i++) r = (unsigned)o[i] < r ? (unsigned)o[i] : r;
  return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  unsigned long long o[wemu::W];
  wemu::exchange(v, o);
  unsigned r = 0;
  for (int i = 0; i < wemu::W; i++) r = (unsigned)o[i] > r ? (unsigned)o[i] : r;
  return r;
}
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline unsigned __float_as_uint(float f) { return (unsigned)wemu::to_bits(f); }
inline float __uint_as_float(unsigned u) { return wemu::from_bits<float>(u); }
// position of the offset-th set bit of mask at or above bit `base` (offset >= 1); 0xffffffff if there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  for (unsigned b = base; b < 32; b++)
    if ((mask >> b) & 1u) { if (--offset <= 0) return b; }
  return 0xffffffffu;
}
inline void __syncthreads() {}
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
