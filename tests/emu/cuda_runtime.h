// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE, not product code.
//
// A stand-in for <cuda_runtime.h> that lets the kernel sources under peleanalysis_b200/csrc be compiled by g++ and run on
// the CPU under a small CUDA execution-model emulator (cuemu.cpp): every CUDA thread of a block is a fiber, blocks run one
// after the other, __syncthreads / warp shuffles / mbarriers / bulk async copies keep their semantics (asynchronous copies
// complete at pseudo-random later scheduler ticks), and a round in which no fiber makes progress is reported as a deadlock.
// It checks what can be checked without a GPU -- descriptor use, indexing, the producer/consumer protocol of the TMA ring,
// host-side sequencing -- against the same oracle and golden vectors as the GPU tests (tests/test_emu_parity.py).  It says
// nothing about PTX semantics, memory-model races inside a warp instruction, or performance.
//
// The emulated library (tests/emu/_build/libpelestencil_emu.so) is built and loaded ONLY by tests; the product binding
// (peleanalysis_b200/capi.py) opens lib/libpelestencil_b200.so and nothing else, and build() never builds this.
#ifndef PA_TESTS_EMU_CUDA_RUNTIME_H
#define PA_TESTS_EMU_CUDA_RUNTIME_H

#ifndef PA_HOST_EMULATION
#error "tests/emu/cuda_runtime.h is only for the PA_HOST_EMULATION test build"
#endif

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <utility>

// ---- declaration specifiers ----------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static          // blocks run one after the other, so a function-local static IS the block's shared memory

// ---- basic types -----------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(16))) double2 { double x, y; };
inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }

enum cudaError_t {
    cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidConfiguration = 9,
    cudaErrorPeerAccessAlreadyEnabled = 704, cudaErrorNotSupported = 801
};
typedef struct CUstream_st* cudaStream_t;
typedef struct CUevent_st* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major, minor; char name[64]; };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaPitchedPtr { void* ptr; size_t pitch, xsize, ysize; };
struct cudaPos { size_t x, y, z; };
struct cudaExtent { size_t width, height, depth; };
struct cudaMemcpy3DParms { cudaPitchedPtr srcPtr; cudaPos srcPos; cudaPitchedPtr dstPtr; cudaPos dstPos; cudaExtent extent; cudaMemcpyKind kind; };
inline cudaPitchedPtr make_cudaPitchedPtr(void* p, size_t pitch, size_t xs, size_t ys) { return cudaPitchedPtr{p, pitch, xs, ys}; }
inline cudaPos make_cudaPos(size_t x, size_t y, size_t z) { return cudaPos{x, y, z}; }
inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return cudaExtent{w, h, d}; }

// ---- runtime API (host memory stands in for device memory) -------------------------------------------------------
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int d);
cudaError_t cudaMallocBytes(void** p, size_t n);
template <class T> cudaError_t cudaMalloc(T** p, size_t n) { return cudaMallocBytes(reinterpret_cast<void**>(p), n); }
cudaError_t cudaFree(void* p);
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaHostRegister(void* p, size_t n, unsigned flags);
cudaError_t cudaHostUnregister(void* p);
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st);
cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms* p, cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
// streams and events: the emulator executes everything in issue order, which is one of the orders the events allow
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* ev, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t ev);
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st);
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned flags);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
template <class F> cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- execution model ---------------------------------------------------------------------------------------------
namespace cuemu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
unsigned char* dyn_smem();
void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
// run `work` now, or -- CUEMU_DEFER_SIDE=1 and `st` was made by cudaStreamCreate* -- when something waits for that stream
// (cudaStreamWaitEvent on an event recorded in it, cudaStreamSynchronize): the LATEST order the events allow, where the
// default is the earliest.  A pass that is correct under both has its cross-stream dependencies covered by events.
void submit(cudaStream_t st, std::function<void()> work);
void syncthreads();
void syncwarp(unsigned mask);
void named_bar(int id, int count);                          // bar.sync id, count
// every participating lane publishes `bytes` bytes; returns a pointer to the 32 published slots (8 bytes each)
const unsigned long long* warp_exchange(unsigned mask, const void* v, size_t bytes);
// mbarrier / asynchronous copies (addresses are plain host pointers)
void mbar_init(uint64_t* bar, uint32_t count);
void mbar_expect_tx(uint64_t* bar, uint32_t bytes);        // arrive + expect-tx
void mbar_arrive(uint64_t* bar);
void mbar_wait(uint64_t* bar, uint32_t parity);
void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar);
void cp_async_8(void* dst, const void* src);
void cp_async_16(void* dst, const void* src);                    // cp.async.cg.shared.global 16 (no mbarrier: commit / wait groups)
void cp_async_n(void* dst, const void* src, int bytes);           // cp.async 4 / 8 / 16 bytes in the same groups
void cp_async_commit();                                           // cp.async.commit_group
void cp_async_wait(int n);                                        // cp.async.wait_group n
void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes);   // cp.async.bulk.global.shared::cta.bulk_group
void bulk_commit();                                               // cp.async.bulk.commit_group
void bulk_wait_read(int n);                                       // cp.async.bulk.wait_group.read n
void cp_async_arrive_noinc(uint64_t* bar);

template <class... KA>
struct Launcher {
    dim3 g, b;
    size_t smem;
    cudaStream_t st;
    void (*k)(KA...);
    template <class... A>
    void operator()(A&&... a) const {
        std::tuple<std::decay_t<KA>...> args(std::forward<A>(a)...);
        void (*fn)(KA...) = k;
        const dim3 gg = g, bb = b;
        const size_t sm = smem;
        submit(st, [fn, args, gg, bb, sm]() { launch_impl(gg, bb, sm, [&]() { std::apply(fn, args); }); });
    }
};
template <class... KA>
Launcher<KA...> launcher(dim3 g, dim3 b, size_t smem, cudaStream_t st, void (*k)(KA...)) { return Launcher<KA...>{g, b, smem, st, k}; }
}  // namespace cuemu

#define threadIdx (::cuemu::g_threadIdx)
#define blockIdx (::cuemu::g_blockIdx)
#define blockDim (::cuemu::g_blockDim)
#define gridDim (::cuemu::g_gridDim)
#define PA_LAUNCH(grid, block, smem, stream, ...) ::cuemu::launcher((grid), (block), (smem), (stream), __VA_ARGS__)
#define PA_DYN_SMEM(name) unsigned char* name = ::cuemu::dyn_smem()

inline void __syncthreads() { cuemu::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { cuemu::syncwarp(mask); }
template <class T> T __shfl_sync(unsigned mask, T v, int src, int = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const unsigned long long* s = cuemu::warp_exchange(mask, &v, sizeof(T));
    T r; std::memcpy(&r, &s[src & 31], sizeof(T)); return r;
}
template <class T> T __shfl_up_sync(unsigned mask, T v, unsigned delta, int = 32) {
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned long long* s = cuemu::warp_exchange(mask, &v, sizeof(T));
    const int src = lane - (int)delta;
    T r = v; if (src >= 0) std::memcpy(&r, &s[src], sizeof(T)); return r;
}
template <class T> T __shfl_down_sync(unsigned mask, T v, unsigned delta, int = 32) {
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned long long* s = cuemu::warp_exchange(mask, &v, sizeof(T));
    const int src = lane + (int)delta;
    T r = v; if (src <= 31) std::memcpy(&r, &s[src], sizeof(T)); return r;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }

// ---- device intrinsics -----------------------------------------------------------------------------------------
inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)(u >> 32); }
inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)u; }
inline double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &u, 8); return x; }
inline double __longlong_as_double(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
inline long long __double_as_longlong(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
inline float __int_as_float(int v) { float x; std::memcpy(&x, &v, 4); return x; }
inline double __drcp_rn(double x) { return 1.0 / x; }

#endif
