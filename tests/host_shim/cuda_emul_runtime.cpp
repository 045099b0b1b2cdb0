// cuda_emul_runtime.cpp -- TEST INFRASTRUCTURE: the CUDA runtime calls the launchers and csrc/capi.cu make, with "device" memory being
// host memory, streams and events being no-ops (every launch is synchronous under tests/host_shim/cuda_emul.h) and exactly one device.
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

extern "C" {
extern "C" int cuda_emul_take_launch_error(void);
cudaError_t cudaGetLastError(void) { return cuda_emul_take_launch_error() ? cudaErrorInvalidConfiguration : cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
// CUDA_EMUL_DEVICES=n pretends to have n devices (all of them the host): enough to run the one-process multi-GPU host logic (csrc/multi.cu)
static int emul_devices() { const char* e = std::getenv("CUDA_EMUL_DEVICES"); const int n = e ? std::atoi(e) : 1; return n > 0 ? n : 1; }
static thread_local int emul_current_device = 0;
cudaError_t cudaGetDeviceCount(int* n) { *n = emul_devices(); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emul_devices()) return cudaErrorInvalidDevice; emul_current_device = d; return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t bytes) { return posix_memalign(p, 256, bytes ? bytes : 256) ? cudaErrorMemoryAllocation : (std::memset(*p, 0xFF, bytes), cudaSuccess); }
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeAsync(void* p, cudaStream_t) { std::free(p); return cudaSuccess; }
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t*, int) { return cudaErrorNotSupported; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = emul_current_device; return cudaSuccess; }
cudaError_t cudaDeviceGetPCIBusId(char*, int, int) { return cudaErrorInvalidDevice; }     // no PCI device: csdrb_host_alloc keeps the default placement
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return posix_memalign(p, 256, bytes ? bytes : 256) ? cudaErrorMemoryAllocation : cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t bytes) { std::memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { std::memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) { std::memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < height; r++) std::memmove(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "cuda_emul"; }
}
