// cuda_emu.h — a small SIMT emulator for the CPU test suite (test infrastructure only).
//
// Runs a __global__ function of this repo on the host, one CTA at a time, every CUDA thread of
// the CTA as a cooperative fibre (ucontext) on ONE OS thread: __syncthreads / __syncwarp /
// shuffles / ballots are rendezvous points at which a fibre yields until its CTA / warp has
// arrived; between them a fibre runs uninterrupted, so shared-memory atomics are plain
// read-modify-writes.  __shared__ variables become function statics (one CTA runs at a time),
// dynamic shared memory is a host buffer.  It models the semantics the kernels rely on
// (barriers, full-mask warp collectives, atomics), not timing, memory spaces or divergence
// rules: full-mask collectives must be reached by every live lane of the warp.
//
// Use: #include "cuda_emu.h", then #include the .cu file (with KML_HOST_EMULATION defined the
// kernels' sources skip their launchers and CUDA-runtime helpers), then
//   kml_emu::launch(grid, block, dyn_smem_bytes, [&] { my_kernel(args); });
#pragma once
#include <cuda_runtime.h>  // host-side types only: uint2, dim3, cudaStream_t
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

// every system header the product sources use, BEFORE the CUDA keywords become macros
// (libstdc++ spells some attributes __noinline__ / __forceinline__ itself)
#include <dlfcn.h>
#include <float.h>
#include <limits.h>
#include <time.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#if __has_include(<nccl.h>)
#include <nccl.h>
#endif

#define KML_HOST_EMULATION 1
#undef __device__
#undef __global__
#undef __shared__
#undef __host__
#undef __forceinline__
#undef __noinline__
#undef __launch_bounds__
#define __device__
#define __global__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)

namespace kml_emu {

struct Idx3 { unsigned x = 0, y = 0, z = 0; };
struct State {
  Idx3 tid, bid, bdim, gdim;       // of the fibre that is running
  int n = 0, cur = 0;              // threads per CTA, running fibre
  std::vector<ucontext_t> ctx;
  std::vector<char> done;
  std::vector<char*> stacks;
  ucontext_t sched;
  int alive = 0, bar_arrived = 0;
  unsigned bar_gen = 0;
  int w_alive[32] = {0}, w_arrived[32] = {0};
  unsigned w_gen[32] = {0};
  uint64_t xchg[32][32];           // [warp][lane] exchange slots of the warp collectives
  std::vector<unsigned char> dyn;  // dynamic shared memory of the running CTA
  size_t dyn_bytes = 0;
  const std::function<void()>* body = nullptr;
  int schedule = -1;               // 0 ascending, 1 descending, 2 random; -1 = read KML_EMU_SCHEDULE
  uint64_t rng = 1;
  std::vector<int> order;
};
inline State& S() { static State s; return s; }

inline void yield_() { State& s = S(); swapcontext(&s.ctx[s.cur], &s.sched); }
inline void block_barrier() {
  State& s = S();
  const unsigned gen = s.bar_gen;
  if (++s.bar_arrived >= s.alive) { s.bar_arrived = 0; ++s.bar_gen; return; }
  while (s.bar_gen == gen) yield_();
}
inline void warp_barrier() {
  State& s = S();
  const int w = s.cur >> 5;
  const unsigned gen = s.w_gen[w];
  if (++s.w_arrived[w] >= s.w_alive[w]) { s.w_arrived[w] = 0; ++s.w_gen[w]; return; }
  while (s.w_gen[w] == gen) yield_();
}
template <class T>
inline T exchange(T v, int src_lane, bool valid) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  State& s = S();
  const int w = s.cur >> 5, lane = s.cur & 31;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  s.xchg[w][lane] = bits;
  warp_barrier();
  T r = v;
  if (valid) memcpy(&r, &s.xchg[w][src_lane & 31], sizeof(T));
  warp_barrier();
  return r;
}
inline void fibre_entry() {
  State& s = S();
  (*s.body)();
  const int t = s.cur;
  s.done[t] = 1;
  --s.alive;
  --s.w_alive[t >> 5];
  // a thread that leaves releases barriers the rest of its CTA / warp is already waiting at
  if (s.alive > 0 && s.bar_arrived >= s.alive) { s.bar_arrived = 0; ++s.bar_gen; }
  const int w = t >> 5;
  if (s.w_alive[w] > 0 && s.w_arrived[w] >= s.w_alive[w]) { s.w_arrived[w] = 0; ++s.w_gen[w]; }
  swapcontext(&s.ctx[t], &s.sched);
}

// Runs body() once per CUDA thread of a grid x block launch (1-D block, 1- or 2-D grid).
inline void launch(Idx3 grid, int block_threads, size_t dyn_smem, const std::function<void()>& body) {
  // one launch at a time: the emulator state and the kernels' __shared__ statics are global, and
  // the library's query lanes call in from several host threads
  static std::mutex launch_mu;
  std::lock_guard<std::mutex> launch_lock(launch_mu);
  State& s = S();
  if (s.schedule < 0) {
    const char* e = getenv("KML_EMU_SCHEDULE");
    s.schedule = (e && !strncmp(e, "reverse", 7)) ? 1 : (e && !strncmp(e, "random", 6)) ? 2 : 0;
    if (s.schedule == 2 && e[6] == ':') s.rng = strtoull(e + 7, nullptr, 10) * 2 + 1;
  }
  constexpr size_t kStack = 512 * 1024;
  if (block_threads < 1 || block_threads > 1024) throw std::runtime_error("kml_emu: bad block size");
  // a GPU refuses these with cudaErrorInvalidConfiguration; the emulator must not run them as no-ops
  if (grid.x == 0 || grid.x > 2147483647u || grid.y > 65535u) {
    fprintf(stderr, "kml_emu: launch with an invalid grid (%u x %u)\n", grid.x, grid.y);
    throw std::runtime_error("kml_emu: invalid grid");
  }
  if (dyn_smem > 227 * 1024) throw std::runtime_error("kml_emu: more dynamic shared memory than an SM has");
  s.n = block_threads;
  s.ctx.resize(block_threads);
  s.done.assign(block_threads, 0);
  while ((int)s.stacks.size() < block_threads) s.stacks.push_back((char*)malloc(kStack));
  // uninitialised on a GPU: poison it; 64 canary bytes behind it catch a CTA that writes past its
  // dynamic shared memory (compute-sanitizer is closed on the GPU pool, so the bounds are checked here)
  s.dyn.assign(dyn_smem + 16 + 64, 0xCD);
  s.dyn_bytes = dyn_smem;
  s.body = &body;
  s.bdim.x = block_threads; s.bdim.y = s.bdim.z = 1;
  s.gdim = grid;
  if (s.gdim.y == 0) s.gdim.y = 1;
  s.gdim.z = 1;
  for (unsigned by = 0; by < s.gdim.y; ++by)
    for (unsigned bx = 0; bx < s.gdim.x; ++bx) {
      s.bid.x = bx; s.bid.y = by; s.bid.z = 0;
      s.alive = block_threads; s.bar_arrived = 0;
      for (int w = 0; w < 32; ++w) { s.w_arrived[w] = 0; s.w_alive[w] = std::max(0, std::min(32, block_threads - 32 * w)); }
      std::fill(s.done.begin(), s.done.end(), 0);
      for (int t = 0; t < block_threads; ++t) {
        getcontext(&s.ctx[t]);
        s.ctx[t].uc_stack.ss_sp = s.stacks[t];
        s.ctx[t].uc_stack.ss_size = kStack;
        s.ctx[t].uc_link = &s.sched;
        makecontext(&s.ctx[t], (void (*)())fibre_entry, 0);
      }
      int guard = 0;
      while (s.alive > 0) {
        const int before = s.alive;
        const unsigned g0 = s.bar_gen;
        for (int k = 0; k < block_threads; ++k) {
          // which fibre runs next: ascending thread id, descending, or a fresh random order per
          // pass (KML_EMU_SCHEDULE=reverse | random[:seed]) — results must not depend on it
          int t = k;
          if (s.schedule == 1) t = block_threads - 1 - k;
          else if (s.schedule == 2) {
            if (k == 0) {
              s.order.resize(block_threads);
              for (int i = 0; i < block_threads; ++i) s.order[i] = i;
              for (int i = block_threads - 1; i > 0; --i) {
                s.rng = s.rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(s.order[i], s.order[(int)((s.rng >> 33) % (uint64_t)(i + 1))]);
              }
            }
            t = s.order[k];
          }
          if (s.done[t]) continue;
          s.cur = t;
          s.tid.x = t; s.tid.y = s.tid.z = 0;
          swapcontext(&s.sched, &s.ctx[t]);
        }
        // every pass must retire a thread or pass a barrier eventually; 1e6 idle passes = deadlock
        guard = (s.alive == before && s.bar_gen == g0) ? guard + 1 : 0;
        if (guard > 1000000) throw std::runtime_error("kml_emu: CTA deadlocked at a barrier");
      }
      {  // canary behind the CTA's dynamic shared memory
        const unsigned char* base = (const unsigned char*)(((uintptr_t)s.dyn.data() + 15) & ~(uintptr_t)15);
        for (size_t i = s.dyn_bytes; i < s.dyn_bytes + 48; ++i)
          if (base[i] != 0xCD) {
            fprintf(stderr, "kml_emu: CTA (%u,%u) wrote %zu bytes past its %zu bytes of dynamic shared memory\n", bx, by,
                    i - s.dyn_bytes + 1, s.dyn_bytes);
            throw std::runtime_error("kml_emu: dynamic shared memory overrun");
          }
      }
    }
}
inline unsigned char* dynamic_smem() {
  State& s = S();
  return (unsigned char*)(((uintptr_t)s.dyn.data() + 15) & ~(uintptr_t)15);
}

}  // namespace kml_emu

// launch + dynamic shared memory macros of csrc/common.cuh, emulated
namespace kml_emu {
inline Idx3 idx3(int x) { Idx3 r; r.x = (unsigned)x; r.y = 1; return r; }
inline Idx3 idx3(unsigned x) { Idx3 r; r.x = x; r.y = 1; return r; }
inline Idx3 idx3(const dim3& d) { Idx3 r; r.x = d.x; r.y = d.y; return r; }
}  // namespace kml_emu
template <class K> inline cudaError_t cudaFuncSetAttribute(K*, cudaFuncAttribute, int) { return cudaSuccess; }
#define KML_UNPAREN(...) __VA_ARGS__
#define KML_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kml_emu::launch(kml_emu::idx3(grid), (int)(block), (size_t)(smem), [&] { KML_UNPAREN kernel(__VA_ARGS__); })
#define KML_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(kml_emu::dynamic_smem())

#define threadIdx (kml_emu::S().tid)
#define blockIdx (kml_emu::S().bid)
#define blockDim (kml_emu::S().bdim)
#define gridDim (kml_emu::S().gdim)

static inline void __syncthreads() { kml_emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { kml_emu::warp_barrier(); }
static inline void __threadfence() {}  // one OS thread: program order is memory order
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return kml_emu::exchange(v, src, true); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return kml_emu::exchange(v, (kml_emu::S().cur & 31) ^ m, true);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = kml_emu::S().cur & 31;
  return kml_emu::exchange(v, lane - d, lane - d >= 0);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = kml_emu::S().cur & 31;
  return kml_emu::exchange(v, lane + d, lane + d < 32);
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
  kml_emu::State& s = kml_emu::S();
  const int w = s.cur >> 5, lane = s.cur & 31;
  s.xchg[w][lane] = pred ? 1u : 0u;
  kml_emu::warp_barrier();
  unsigned r = 0;
  for (int l = 0; l < s.w_alive[w] && l < 32; ++l) r |= (unsigned)(s.xchg[w][l] & 1u) << l;
  kml_emu::warp_barrier();
  return r;
}
static inline int __all_sync(unsigned m, bool pred) {
  const unsigned b = __ballot_sync(m, pred);
  const int n = kml_emu::S().w_alive[kml_emu::S().cur >> 5];
  return b == (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u));
}
static inline int __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
template <class T, class V> static inline T atomicAdd(T* p, V v) { const T old = *p; *p = (T)(old + (T)v); return old; }
template <class T, class V> static inline T atomicMin(T* p, V v) { const T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <class T, class V> static inline T atomicMax(T* p, V v) { const T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline unsigned __vsadu4(unsigned a, unsigned b) {  // sum of absolute differences of the four bytes
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const int x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255;
    r += (unsigned)(x > y ? x - y : y - x);
  }
  return r;
}
using std::max;
using std::min;

// ---- mbarrier + bulk copy (common.cuh), modelled in the 64-bit barrier word: pending
// transaction bytes, pending arrivals, the arrival count it is re-armed with, phase parity.
namespace kml {
struct EmuMbar { uint32_t tx; uint16_t pending; uint8_t count, phase; };
static_assert(sizeof(EmuMbar) == 8, "mbarrier word");
inline void emu_mbar_try_complete(EmuMbar* b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->count; }
}
inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  b->tx = 0; b->pending = (uint16_t)count; b->count = (uint8_t)count; b->phase = 0;
}
inline void fence_mbar_init() {}
// arrive + expect: the phase cannot complete before `bytes` have landed
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  if (b->pending == 0) throw std::runtime_error("kml_emu: mbarrier arrival beyond its count");
  b->tx += bytes;
  b->pending--;
}
// the copy is performed at once (an early read is not caught); a wait on the wrong parity, a
// byte count that does not match the expectation or a misaligned copy is
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  if ((bytes & 15u) || ((uintptr_t)dst_smem & 15u) || ((uintptr_t)src_gmem & 15u))
    throw std::runtime_error("kml_emu: cp.async.bulk needs 16-byte aligned addresses and size");
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  if (b->tx < bytes) throw std::runtime_error("kml_emu: bulk copy larger than the expected transaction bytes");
  memcpy(dst_smem, src_gmem, bytes);
  b->tx -= bytes;
  emu_mbar_try_complete(b);
}
// returns once the phase with parity `parity` has completed
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  while ((uint32_t)(b->phase & 1) == (parity & 1u)) kml_emu::yield_();
}
}  // namespace kml
