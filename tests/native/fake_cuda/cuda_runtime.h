// Host-only stand-in for the CUDA runtime, used ONLY by the ThreadSanitizer harness of mailbox.cpp: device memory is host
// memory, an IPC handle carries the pointer, a "stream" executes in the calling thread.  It lets TSAN see the complete
// host-side protocol (table lookups, shm doorbells, payload / flag ordering) with real threads.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef struct CUstream_st* cudaStream_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
constexpr unsigned cudaIpcMemLazyEnablePeerAccess = 1, cudaHostRegisterMapped = 2, cudaHostRegisterPortable = 1;

inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n); return *p ? cudaSuccess : 1; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
