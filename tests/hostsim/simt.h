// simt.h -- a tiny deterministic SIMT emulator (TEST INFRASTRUCTURE ONLY).
//
// The build container has nvcc but no GPU.  To check the *arithmetic and indexing* of the
// sm_100a kernels against the oracle before spending GPU minutes, the kernel sources under
// level-s2fm_official_b200/csrc are also compiled by g++ with -DLS_HOSTSIM against this header.
// Every CUDA thread of a block becomes a ucontext fiber; __syncthreads / __syncwarp / warp
// shuffles / votes yield to a round-robin scheduler with the real barrier semantics (a warp
// barrier releases when all live lanes of that warp wait on it, a block barrier when all
// live threads do).  Atomics are plain read-modify-writes (one OS thread).
//
// Nothing here is product code: the product library is built by nvcc only, has no CPU path
// and the python package refuses to load anything but the sm_100a library.
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

namespace simt {

enum Wait { RUN = 0, WAIT_WARP = 1, WAIT_BLOCK = 2, DONE = 3 };

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    int state = RUN;
    unsigned tid = 0;
};

struct State {
    dim3 threadIdx, blockIdx, blockDim, gridDim;
    char* smem = nullptr;
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int cur = -1;
    uint64_t xchg[1024];   // shuffle / vote exchange slots, one per thread
    std::function<void()> body;
};

inline State& S() { static State s; return s; }

inline void yield(int why) {
    State& s = S();
    Fiber& f = s.fibers[s.cur];
    f.state = why;
    swapcontext(&f.ctx, &s.sched);
}

inline void trampoline() {
    State& s = S();
    s.body();
    s.fibers[s.cur].state = DONE;
    swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

// Runs one block: all fibers round-robin until everybody is DONE.
inline void run_block(unsigned nthreads, size_t smem_bytes, dim3 bidx) {
    State& s = S();
    const size_t STACK = 256 * 1024;
    s.fibers.resize(nthreads);
    std::vector<char> smem(smem_bytes + 64, 0);
    for (size_t i = 0; i < smem.size(); ++i) smem[i] = (char)0xCD;   // garbage, like the real thing
    s.smem = smem.data() + (64 - ((uintptr_t)smem.data() & 63)) % 64;
    s.blockIdx = bidx;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = s.fibers[t];
        if (!f.stack) f.stack = (char*)malloc(STACK);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK;
        f.ctx.uc_link = &s.sched;
        f.state = RUN;
        f.tid = t;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    for (;;) {
        bool progressed = false;
        unsigned done = 0;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber& f = s.fibers[t];
            if (f.state == RUN) {
                s.cur = (int)t;
                s.threadIdx = dim3(t, 0, 0);
                swapcontext(&s.sched, &f.ctx);
                progressed = true;
            }
            if (f.state == DONE) ++done;
        }
        if (done == nthreads) break;
        // release warp barriers
        for (unsigned w = 0; w * 32 < nthreads; ++w) {
            unsigned lo = w * 32, hi = lo + 32 < nthreads ? lo + 32 : nthreads;
            bool all = true, any = false;
            for (unsigned t = lo; t < hi; ++t) {
                if (s.fibers[t].state == WAIT_WARP) any = true;
                else if (s.fibers[t].state != DONE) all = false;
            }
            if (any && all) {
                for (unsigned t = lo; t < hi; ++t) if (s.fibers[t].state == WAIT_WARP) s.fibers[t].state = RUN;
                progressed = true;
            }
        }
        // release the block barrier
        {
            bool all = true, any = false;
            for (unsigned t = 0; t < nthreads; ++t) {
                if (s.fibers[t].state == WAIT_BLOCK) any = true;
                else if (s.fibers[t].state != DONE) all = false;
            }
            if (any && all) {
                for (unsigned t = 0; t < nthreads; ++t) if (s.fibers[t].state == WAIT_BLOCK) s.fibers[t].state = RUN;
                progressed = true;
            }
        }
        if (!progressed) {
            fprintf(stderr, "simt: deadlock (divergent barrier) in block %u\n", bidx.x);
            abort();
        }
    }
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& f) {
    State& s = S();
    s.gridDim = grid;
    s.blockDim = block;
    s.body = std::function<void()>(f);
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) run_block(block.x * block.y * block.z, smem_bytes, dim3(bx, by, 0));
}

template <class T>
inline T xchg_read(unsigned src_tid) {
    T v;
    memcpy(&v, &S().xchg[src_tid], sizeof(T));
    return v;
}
template <class T>
inline void xchg_write(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    memcpy(&S().xchg[S().threadIdx.x], &v, sizeof(T));
}

}  // namespace simt

#define threadIdx (simt::S().threadIdx)
#define blockIdx (simt::S().blockIdx)
#define blockDim (simt::S().blockDim)
#define gridDim (simt::S().gridDim)

static inline void __syncthreads() { simt::yield(simt::WAIT_BLOCK); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::yield(simt::WAIT_WARP); }
static inline void __threadfence() {}

template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    simt::xchg_write(v);
    simt::yield(simt::WAIT_WARP);
    unsigned tid = threadIdx.x, base = tid & ~31u, lane = tid & 31u;
    unsigned seg = lane & ~(unsigned)(width - 1);
    T r = simt::xchg_read<T>(base + seg + ((unsigned)src & (unsigned)(width - 1)));
    simt::yield(simt::WAIT_WARP);
    return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    simt::xchg_write(v);
    simt::yield(simt::WAIT_WARP);
    unsigned tid = threadIdx.x;
    T r = simt::xchg_read<T>((tid & ~31u) + ((tid & 31u) ^ (unsigned)m));
    simt::yield(simt::WAIT_WARP);
    return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    simt::xchg_write(v);
    simt::yield(simt::WAIT_WARP);
    unsigned tid = threadIdx.x, lane = tid & 31u;
    T r = lane >= d ? simt::xchg_read<T>(tid - d) : v;
    simt::yield(simt::WAIT_WARP);
    return r;
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    simt::xchg_write(v);
    simt::yield(simt::WAIT_WARP);
    unsigned tid = threadIdx.x, lane = tid & 31u;
    T r = lane + d < 32 ? simt::xchg_read<T>(tid + d) : v;
    simt::yield(simt::WAIT_WARP);
    return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    simt::xchg_write<unsigned>(pred ? 1u : 0u);
    simt::yield(simt::WAIT_WARP);
    unsigned base = threadIdx.x & ~31u, r = 0;
    unsigned n = blockDim.x * blockDim.y * blockDim.z;
    for (unsigned l = 0; l < 32 && base + l < n; ++l) r |= simt::xchg_read<unsigned>(base + l) << l;
    simt::yield(simt::WAIT_WARP);
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }

static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline float2 atomicAdd(float2* p, float2 v) { float2 o = *p; p->x += v.x; p->y += v.y; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdividef(float a, float b) { return a / b; }

typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "hostsim"; }
