// hd.h -- host/device portability shims for the step phases.
//
// The step pipeline is written once as templated phase functions (step_phases.h).  On the product
// path they are compiled by nvcc for sm_100a and driven by CUDA "team" contexts (one CTA per arena,
// or the whole cooperative grid for a huge arena).  The same functions also compile as plain C++
// for the test-only host emulation under tests/emu/ (used to debug the parallel formulations
// against the reference on GPU-less development containers; never shipped, never a fallback).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#define MG_D __device__ __forceinline__
#else
#define MG_HD inline
#define MG_D inline
#endif

namespace mg {

// ---- atomics (device: hardware atomics; host emulation: single-threaded plain ops) -------------
MG_HD int atomic_exch(int *p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicExch(p, v);
#else
    int o = *p; *p = v; return o;
#endif
}
MG_HD int atomic_min(int *p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicMin(p, v);
#else
    int o = *p; if (v < o) *p = v; return o;
#endif
}
MG_HD int atomic_max(int *p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicMax(p, v);
#else
    int o = *p; if (v > o) *p = v; return o;
#endif
}
MG_HD int atomic_add(int *p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    int o = *p; *p = o + v; return o;
#endif
}
MG_HD int atomic_or(int *p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicOr(p, v);
#else
    int o = *p; *p = o | v; return o;
#endif
}
MG_HD float atomic_addf(float *p, float v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    float o = *p; *p = o + v; return o;
#endif
}
MG_HD void atomic_add64(long long *p, long long v) {
#if defined(__CUDA_ARCH__)
    atomicAdd((unsigned long long *)p, (unsigned long long)v);
#else
    *p += v;
#endif
}

// Per-thread "this item needs no further sweeps" marks of a relaxation (phase_move_relax).  Device: one register,
// the first 32 items of the thread (a 512-thread team visits ~4 movers per thread).  Host emulation (one thread
// visits every mover): unbounded, so the CPU test-suite exercises the same skipping logic on every mover.
#if defined(__CUDACC__)
struct SettledMask {
    unsigned bits = 0;
    MG_HD bool get(int j) const { return j < 32 && ((bits >> j) & 1u); }
    MG_HD void set(int j) { if (j < 32) bits |= 1u << j; }
};
#else
}  // namespace mg
#include <vector>
namespace mg {
struct SettledMask {
    std::vector<bool> v;
    bool get(int j) const { return j < (int)v.size() && v[j]; }
    void set(int j) { if (j >= (int)v.size()) v.resize(j + 1, false); v[j] = true; }
};
#endif

// small read-only tables (range offsets): through the non-coherent L1 path on the device
MG_HD int ld_ro(const int *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// loads that must observe in-place updates made by other threads during a relaxation sweep
MG_HD int ld_volatile(const int *p) { return *(const volatile int *)p; }
MG_HD unsigned char ld_volatile(const unsigned char *p) { return *(const volatile unsigned char *)p; }
MG_HD void st_volatile(int *p, int v) { *(volatile int *)p = v; }
MG_HD void st_volatile(unsigned char *p, unsigned char v) { *(volatile unsigned char *)p = v; }

// ---- minstd_rand0 (std::default_random_engine in libstdc++): x <- 16807 x mod (2^31 - 1) -------
// reference: src/gridworld/GridWorld.h:109 (engine), GridWorld.cc:29 (seed 0), :465-468 (shuffle)
static const uint32_t MINSTD_M = 2147483647u;
static const uint32_t MINSTD_A = 16807u;

MG_HD uint32_t mulmod31(uint32_t a, uint32_t b) {
    uint64_t p = (uint64_t)a * (uint64_t)b;           // < 2^62
    p = (p & MINSTD_M) + (p >> 31);                   // < 2^32
    p = (p & MINSTD_M) + (p >> 31);
    if (p >= MINSTD_M) p -= MINSTD_M;
    return (uint32_t)p;
}

// a^k mod M using the table pow2[b] = a^(2^b) mod M
MG_HD uint32_t minstd_pow(const uint32_t *pow2, uint32_t k) {
    uint32_t r = 1;
    for (int b = 0; k; ++b, k >>= 1)
        if (k & 1u) r = mulmod31(r, pow2[b]);
    return r;
}

MG_HD uint32_t minstd_seed(long long seed_as_int) {
    // std::linear_congruential_engine::seed(s): x = s mod m, 0 -> 1; the reference passes
    // (unsigned long)(int)value (GridWorld.cc:145)
    unsigned long long s = (unsigned long long)seed_as_int;
    uint32_t x = (uint32_t)(s % MINSTD_M);
    return x == 0 ? 1u : x;
}

}  // namespace mg
