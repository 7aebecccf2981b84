// common.cuh — shared device/host helpers for libkml (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libkml is written for sm_100a (B200) only"
#endif

namespace kml {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define KML_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      throw ::kml::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
  } while (0)

#ifndef KML_HOST_EMULATION
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
#else
constexpr int kNumSMs = 2;    // tests/emu/: grid-stride kernels sized by the SM count stay cheap to emulate
#endif

// Kernel launch and dynamic shared memory as macros, so that the CPU test suite can compile the
// same .cu text for the host and run the kernels under tests/emu/cuda_emu.h (which defines its own
// versions and KML_HOST_EMULATION).  A kernel name is passed in parentheses (template arguments
// may contain commas):  KML_LAUNCH((k<8, 64>), grid, block, smem, stream, args...);
#ifndef KML_HOST_EMULATION
#define KML_UNPAREN(...) __VA_ARGS__
#define KML_LAUNCH(kernel, grid, block, smem, stream, ...) \
  KML_UNPAREN kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define KML_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

// growable device buffer (doubling); contents preserved on growth
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // ensure capacity >= n elements; keep the first `keep` elements
  void reserve(size_t n, size_t keep, cudaStream_t s) {
    if (n <= cap) return;
    size_t nc = cap ? cap : 1024;
    while (nc < n) nc *= 2;
    T* np = nullptr;
    KML_CUDA(cudaMalloc(&np, nc * sizeof(T)));
    if (p && keep) KML_CUDA(cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) {
      KML_CUDA(cudaStreamSynchronize(s));
      cudaFree(p);
    }
    p = np;
    cap = nc;
  }
  // exact-size scratch (no preservation)
  void scratch(size_t n) {
    if (n <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    KML_CUDA(cudaMalloc(&p, n * sizeof(T)));
    cap = n;
  }
};

// pinned host staging buffer
template <class T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  void scratch(size_t n) {
    if (n <= cap) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    KML_CUDA(cudaMallocHost(&p, n * sizeof(T)));
    cap = n;
  }
};

#ifdef __CUDACC__
// ---- mbarrier + bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) --------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
#endif

}  // namespace kml
