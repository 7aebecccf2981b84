// fake_cudart.cpp — the CUDA runtime entry points libkml's host code calls, on host memory, for
// the emulated build of the whole library (tests/emu/build_libkml_emu.py): "device" memory is
// malloc'ed, copies and memsets happen at once, streams and events are tokens (kernels run to
// completion inside their launch, so everything is always in order), one "device" that reports
// compute capability 10.0.  Test infrastructure only.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

extern "C" {

cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated runtime error"; }
// the versioned name cuda_runtime_api.h maps cudaGetDeviceProperties to
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof *p);
  strcpy(p->name, "kml_emu (host)");
  p->major = 10; p->minor = 0; p->multiProcessorCount = 2;
  return cudaSuccess;
}
#ifdef cudaGetDeviceProperties
#undef cudaGetDeviceProperties
#endif
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d) { return cudaGetDeviceProperties_v2(p, d); }

cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  memset(a, 0, sizeof *a);
  a->type = cudaMemoryTypeUnregistered;  // every host array takes the staging path of batch_upload
  a->hostPointer = const_cast<void*>(p);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { if (n) memset(p, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free((void*)s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free((void*)e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }

}  // extern "C"
