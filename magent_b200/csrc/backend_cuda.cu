// backend_cuda.cu -- the sm_100a backend: CUDA runtime plumbing + every kernel of the engine.
//
// Kernels:
//   step_kernel_cta   one CTA per arena runs the whole step pipeline (run_step) with __syncthreads
//                     between phases; arenas are looped grid-stride.          [many small arenas]
//   step_kernel_grid  cooperative launch; the whole grid is one team per arena with grid.sync()
//                     between phases.                                          [one huge arena]
//   cull_kernel_*     clear_dead: stable compaction into the ping-pong SoA buffers.
//   offsets_kernel    prefix of per-arena counts -> ABI concatenation offsets.
//   minimap_*         per (arena, group) histogram of coarse cells + normalisation.
//   obs_render_kernel the observation gather: the HBM-write-bound kernel the roofline is quoted on.
//   info_kernel       id/pos/alive/reward gathers and the action scatter.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

#include "backend.h"
#include "obs_phases.h"

namespace cg = cooperative_groups;

namespace mg {
[[noreturn]] void fatal(const char *fmt, ...);
namespace be {

#define CUDA_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) mg::fatal("CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e),   \
                                         __FILE__, __LINE__, #expr);                              \
    } while (0)

static int g_device = -1, g_sms = 0;
static long long g_launches = 0;
static bool g_profile = false;
static cudaEvent_t g_ev0, g_ev1;
static double g_obs_ms = 0.0;
static long long g_obs_launches = 0;

const char *name() { return "cuda-sm_100a"; }

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

bool init(int device, std::string *err) {
    int n = device_count();
    if (n <= 0) { if (err) *err = "cudaGetDeviceCount found no device"; return false; }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= n) { if (err) *err = "device_id out of range"; return false; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { if (err) *err = cudaGetErrorString(e); return false; }
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    g_device = device;
    g_sms = prop.multiProcessorCount;
    return true;
}
int sm_count() { return g_sms; }

void *dmalloc(size_t bytes) { void *p = nullptr; CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 16)); return p; }
void dfree(void *p) { if (p) cudaFree(p); }
void dmemset(void *p, int byte, size_t bytes) { CUDA_CHECK(cudaMemsetAsync(p, byte, bytes, 0)); }
void h2d(void *dst, const void *src, size_t bytes) { if (bytes) CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); }
void d2h(void *dst, const void *src, size_t bytes) { if (bytes) CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); }
void d2d(void *dst, const void *src, size_t bytes) { if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, 0)); }
void *host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void host_free(void *p) { if (p) cudaFreeHost(p); }
bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
void sync() { CUDA_CHECK(cudaDeviceSynchronize()); }
long long launch_count() { return g_launches; }
void profile_enable(bool on) {
    if (on && !g_profile) { CUDA_CHECK(cudaEventCreate(&g_ev0)); CUDA_CHECK(cudaEventCreate(&g_ev1)); }
    g_profile = on;
    g_obs_ms = 0.0; g_obs_launches = 0;
}
void profile_read(double *ms, long long *n) { *ms = g_obs_ms; *n = g_obs_launches; }

static void post_launch(const char *what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) mg::fatal("kernel launch failed (%s): %s", what, cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------
// team contexts
__device__ __forceinline__ int block_excl_scan(int v, int &total) {
    __shared__ int warp_sums[32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < nw) warp_sums[lane] = wi - w;          // exclusive prefix of warp totals
        if (lane == 31) block_total = wi;
    }
    __syncthreads();
    int excl = incl - v + warp_sums[wid];
    total = block_total;
    __syncthreads();
    return excl;
}

struct CtaCtx {
    MG_HD int tid() const {
#if defined(__CUDA_ARCH__)
        return threadIdx.x;
#else
        return 0;
#endif
    }
    MG_HD int nth() const {
#if defined(__CUDA_ARCH__)
        return blockDim.x;
#else
        return 1;
#endif
    }
    MG_HD void sync() {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
    template <class P, class Em>
    MG_HD int scan(int n, P pred, Em emit) {
#if defined(__CUDA_ARCH__)
        int running = 0;
        for (int tile = 0; tile < n; tile += blockDim.x) {
            int i = tile + threadIdx.x;
            int p = i < n ? pred(i) : 0;
            int tot;
            int ex = block_excl_scan(p, tot);
            if (p) emit(i, running + ex);
            running += tot;
        }
        return running;
#else
        return 0;
#endif
    }
};

struct GridCtx {
    int *scratch;       // [2][4096]
    int parity;
    MG_HD int tid() const {
#if defined(__CUDA_ARCH__)
        return blockIdx.x * blockDim.x + threadIdx.x;
#else
        return 0;
#endif
    }
    MG_HD int nth() const {
#if defined(__CUDA_ARCH__)
        return gridDim.x * blockDim.x;
#else
        return 1;
#endif
    }
    MG_HD void sync() {
#if defined(__CUDA_ARCH__)
        cg::this_grid().sync();
#endif
    }
    template <class P, class Em>
    MG_HD int scan(int n, P pred, Em emit) {
#if defined(__CUDA_ARCH__)
        // contiguous chunk per CTA keeps the emitted order == index order
        const int nb = gridDim.x, bid = blockIdx.x;
        const int L = (n + nb - 1) / nb;
        const int lo = min(n, bid * L), hi = min(n, lo + L);
        int cnt = 0;
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) cnt += pred(i);
        int cta_total;
        block_excl_scan(cnt, cta_total);
        int *sc = scratch + (parity & 1) * 4096;
        parity ^= 1;
        if (threadIdx.x == 0) sc[bid] = cta_total;
        cg::this_grid().sync();
        int below = 0, all = 0;
        for (int b = threadIdx.x; b < nb; b += blockDim.x) { int v = sc[b]; all += v; if (b < bid) below += v; }
        int base, total;
        block_excl_scan(below, base);
        block_excl_scan(all, total);
        int running = base;
        for (int tile = lo; tile < hi; tile += blockDim.x) {
            int i = tile + threadIdx.x;
            int p = i < hi ? pred(i) : 0;
            int tot;
            int ex = block_excl_scan(p, tot);
            if (p) emit(i, running + ex);
            running += tot;
        }
        return total;
#else
        return 0;
#endif
    }
};

__device__ __forceinline__ void load_engine(EngineDev *sE, const EngineDev *gE) {
    const int *src = (const int *)gE;
    int *dst = (int *)sE;
    for (int i = threadIdx.x; i < (int)(sizeof(EngineDev) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

constexpr int STEP_THREADS = 512;

__global__ void __launch_bounds__(STEP_THREADS) step_kernel_cta(const EngineDev *gE, StepArgs S) {
    __shared__ EngineDev sE;
    load_engine(&sE, gE);
    CtaCtx c;
    for (int a = blockIdx.x; a < sE.A; a += gridDim.x) run_step(c, sE, S, a);
}

__global__ void __launch_bounds__(STEP_THREADS) step_kernel_grid(const EngineDev *gE, StepArgs S) {
    __shared__ EngineDev sE;
    load_engine(&sE, gE);
    GridCtx c;
    c.scratch = sE.team_scratch;
    c.parity = 0;
    for (int a = 0; a < sE.A; ++a) run_step(c, sE, S, a);
}

__global__ void __launch_bounds__(STEP_THREADS) cull_kernel_cta(const EngineDev *gE, unsigned curmask) {
    __shared__ EngineDev sE;
    load_engine(&sE, gE);
    CtaCtx c;
    for (int a = blockIdx.x; a < sE.A; a += gridDim.x) run_cull(c, sE, curmask, a);
}

__global__ void __launch_bounds__(STEP_THREADS) cull_kernel_grid(const EngineDev *gE, unsigned curmask) {
    __shared__ EngineDev sE;
    load_engine(&sE, gE);
    GridCtx c;
    c.scratch = sE.team_scratch;
    c.parity = 0;
    for (int a = 0; a < sE.A; ++a) run_cull(c, sE, curmask, a);
}

static const int GRID_MODE_THRESHOLD = 32768;    // agents per arena above which the whole grid teams up

static int coop_grid(const void *kernel) {
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, STEP_THREADS, 0));
    if (per_sm < 1) mg::fatal("cooperative kernel does not fit on an SM");
    int g = per_sm * g_sms;
    return g > 4096 ? 4096 : g;
}

void launch_step(const EngineDev *dE, const EngineDev &hE, const StepArgs &S, int max_agents) {
    if (max_agents > GRID_MODE_THRESHOLD) {
        int grid = coop_grid((const void *)step_kernel_grid);
        StepArgs s = S;
        void *args[] = {(void *)&dE, (void *)&s};
        CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)step_kernel_grid, dim3(grid), dim3(STEP_THREADS), args, 0, 0));
        post_launch("step_kernel_grid");
    } else {
        int grid = hE.A < 8 * g_sms ? hE.A : 8 * g_sms;
        step_kernel_cta<<<grid, STEP_THREADS>>>(dE, S);
        post_launch("step_kernel_cta");
    }
}

void launch_cull(const EngineDev *dE, const EngineDev &hE, unsigned curmask, int max_agents) {
    if (max_agents > GRID_MODE_THRESHOLD) {
        int grid = coop_grid((const void *)cull_kernel_grid);
        void *args[] = {(void *)&dE, (void *)&curmask};
        CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)cull_kernel_grid, dim3(grid), dim3(STEP_THREADS), args, 0, 0));
        post_launch("cull_kernel_grid");
    } else {
        int grid = hE.A < 8 * g_sms ? hE.A : 8 * g_sms;
        cull_kernel_cta<<<grid, STEP_THREADS>>>(dE, curmask);
        post_launch("cull_kernel_cta");
    }
}

// off[g][0..A] = exclusive prefix of n[g][0..A-1]; one CTA per group
__global__ void __launch_bounds__(1024) offsets_kernel(const EngineDev *gE) {
    const int g = blockIdx.x;
    const int A = gE->A;
    const int *n = gE->n + (size_t)g * A;
    int *off = gE->off + (size_t)g * (A + 1);
    int running = 0;
    for (int tile = 0; tile < A; tile += blockDim.x) {
        int a = tile + threadIdx.x;
        int v = a < A ? n[a] : 0;
        int tot;
        int ex = block_excl_scan(v, tot);
        if (a < A) off[a] = running + ex;
        running += tot;
    }
    if (threadIdx.x == 0) off[A] = running;
}

void launch_offsets(const EngineDev *dE, const EngineDev &hE) {
    offsets_kernel<<<hE.G, 1024>>>(dE);
    post_launch("offsets_kernel");
}

// arena of the idx-th agent in the concatenation of group g
__device__ __forceinline__ int locate_arena(const int *off, int A, int idx) {
    int lo = 0, hi = A;                 // invariant: off[lo] <= idx < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) info_kernel(const EngineDev *gE, unsigned curmask, int kind, int g,
                                                   void *buf, int n_total) {
    const EngineDev &E = *gE;
    const int *off = E.off + (size_t)g * (E.A + 1);
    const AgentSoA &s = E.grp[g].soa[(curmask >> g) & 1u];
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_total; o += gridDim.x * blockDim.x) {
        int a = E.A == 1 ? 0 : locate_arena(off, E.A, o);
        long gi = (long)a * E.grp[g].cap + (o - off[a]);
        switch (kind) {
            case INFO_ID: ((int *)buf)[o] = s.id[gi]; break;
            case INFO_POS: ((int2 *)buf)[o] = make_int2(s.x[gi], s.y[gi]); break;
            case INFO_ALIVE: ((unsigned char *)buf)[o] = (s.flags[gi] & FLAG_DEAD) ? 0 : 1; break;
            case INFO_REWARD: ((float *)buf)[o] = s.next_reward[gi] + E.hdr[a].grp_reward[g]; break;
            case INFO_HP: ((float *)buf)[o] = s.hp[gi]; break;
            case INFO_ACTION_SCATTER: s.act[gi] = ((const int *)buf)[o]; break;
        }
    }
}

void launch_info(const EngineDev *dE, const EngineDev &, unsigned curmask, int kind, int group, void *buf, int n_total) {
    int grid = (n_total + 255) / 256;
    if (grid > 8 * g_sms) grid = 8 * g_sms;
    info_kernel<<<grid, 256>>>(dE, curmask, kind, group, buf, n_total);
    post_launch("info_kernel");
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) random_actions_kernel(const EngineDev *gE, unsigned curmask, int g,
                                                             unsigned long long seed, int n_total) {
    const EngineDev &E = *gE;
    const int *off = E.off + (size_t)g * (E.A + 1);
    const AgentSoA &s = E.grp[g].soa[(curmask >> g) & 1u];
    const unsigned na = (unsigned)E.grp[g].n_action;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_total; o += gridDim.x * blockDim.x) {
        int a = E.A == 1 ? 0 : locate_arena(off, E.A, o);
        long gi = (long)a * E.grp[g].cap + (o - off[a]);
        s.act[gi] = (int)((splitmix64(seed ^ ((unsigned long long)o * 0xD1342543DE82EF95ull)) >> 33) % na);
    }
}

void launch_random_actions(const EngineDev *dE, const EngineDev &, unsigned curmask, int group,
                           unsigned long long seed, int n_total) {
    int grid = (n_total + 255) / 256;
    if (grid > 8 * g_sms) grid = 8 * g_sms;
    random_actions_kernel<<<grid, 256>>>(dE, curmask, group, seed, n_total);
    post_launch("random_actions_kernel");
}

// ------------------------------------------------------------------------------------------------
// minimap: counts per (arena, group, coarse cell) then value = (float)count / (float)group size
__global__ void __launch_bounds__(256) minimap_hist_kernel(const EngineDev *gE, unsigned curmask, int og, int chunk) {
    extern __shared__ int hist[];
    const EngineDev &E = *gE;
    const int ag = blockIdx.y;               // a * G + j
    const int a = ag / E.G, j = ag - a * E.G;
    const int vw = E.grp[og].view_w, vh = E.grp[og].view_h, cells = vw * vh;
    const int n = E.n[j * E.A + a];
    const int lo = blockIdx.x * chunk;
    if (lo >= n) return;
    const int hi = min(n, lo + chunk);
    for (int k = threadIdx.x; k < cells; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const AgentSoA &s = E.grp[j].soa[(curmask >> j) & 1u];
    const int scale_h = (E.H + vh - 1) / vh, scale_w = (E.W + vw - 1) / vw;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        long gi = (long)a * E.grp[j].cap + i;
        atomicAdd(&hist[(s.y[gi] / scale_h) * vw + s.x[gi] / scale_w], 1);
    }
    __syncthreads();
    int *out = E.mm_count + (size_t)ag * cells;
    for (int k = threadIdx.x; k < cells; k += blockDim.x)
        if (hist[k]) atomicAdd(&out[k], hist[k]);
}

__global__ void __launch_bounds__(256) minimap_norm_kernel(const EngineDev *gE, int og, float *mm_val, int total) {
    const EngineDev &E = *gE;
    const int cells = E.grp[og].view_w * E.grp[og].view_h;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        int ag = k / cells;
        int a = ag / E.G, j = ag - a * E.G;
        mm_val[k] = (float)E.mm_count[k] / (float)E.n[j * E.A + a];       // GridWorld.cc:350-357
    }
}

void launch_minimap(const EngineDev *dE, const EngineDev &hE, unsigned curmask, int og, float *mm_val) {
    const int cells = hE.grp[og].view_w * hE.grp[og].view_h;
    const int total = hE.A * hE.G * cells;
    CUDA_CHECK(cudaMemsetAsync(hE.mm_count, 0, (size_t)total * 4, 0));
    int cap_max = 0;
    for (int g = 0; g < hE.G; ++g) cap_max = cap_max > hE.grp[g].cap ? cap_max : hE.grp[g].cap;
    const int chunk = 4096;
    dim3 grid((cap_max + chunk - 1) / chunk, hE.A * hE.G);
    minimap_hist_kernel<<<grid, 256, cells * sizeof(int)>>>(dE, curmask, og, chunk);
    post_launch("minimap_hist_kernel");
    int g2 = (total + 255) / 256;
    if (g2 > 8 * g_sms) g2 = 8 * g_sms;
    minimap_norm_kernel<<<g2, 256>>>(dE, og, mm_val, total);
    post_launch("minimap_norm_kernel");
}

// ------------------------------------------------------------------------------------------------
// obs_render_kernel: the observation gather (GridWorld.cc:292-401 + Map::extract_view Map.cc:129-207).
//
// One CTA composes a tile of OBS_TA consecutive agents of the ABI concatenation in shared memory
// (each warp one agent at a time, each lane one view cell: n_channel floats, stride n_channel words =>
// conflict-free for odd channel counts) and streams the tile out with 16-byte coalesced stores; the
// tile's byte range in the output is contiguous and 16-byte aligned because OBS_TA % 4 == 0.
// Algorithmic traffic per agent: 4*(view_h*view_w*n_channel + feature) bytes written.
constexpr int OBS_THREADS = 256;
constexpr int OBS_TA = 8;

struct ObsHdr { int a, x, y, cx, cy, i; };

__global__ void __launch_bounds__(OBS_THREADS) obs_render_kernel(const EngineDev *gE, ObsArgs O, const float *mm_val, int n_total) {
    extern __shared__ __align__(16) float tile[];
    __shared__ EngineDev sE;
    __shared__ ObsHdr hdr[OBS_TA];
    load_engine(&sE, gE);
    const EngineDev &E = sE;
    const int g = O.group;
    const GroupDev &G = E.grp[g];
    const int cells = G.view_w * G.view_h, C = E.n_channel, rec = cells * C, F = G.feature_size;
    const int *off = E.off + (size_t)g * (E.A + 1);
    const AgentSoA &s = E.grp[g].soa[(O.curmask >> g) & 1u];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = OBS_THREADS / 32;

    for (int t0 = blockIdx.x * OBS_TA; t0 < n_total; t0 += gridDim.x * OBS_TA) {
        const int cnt = min(OBS_TA, n_total - t0);
        if (threadIdx.x < cnt) {
            int o = t0 + threadIdx.x;
            int a = E.A == 1 ? 0 : locate_arena(off, E.A, o);
            int i = o - off[a];
            long gi = (long)a * G.cap + i;
            ObsHdr h;
            h.a = a; h.i = i; h.x = s.x[gi]; h.y = s.y[gi];
            h.cx = -1; h.cy = -1;
            if (mm_val) minimap_cell(E, G.view_w, G.view_h, h.x, h.y, h.cx, h.cy);
            hdr[threadIdx.x] = h;
        }
        __syncthreads();
        for (int ag = warp; ag < cnt; ag += nwarp) {
            const ObsHdr h = hdr[ag];
            const float *mm = mm_val ? mm_val + (size_t)h.a * E.G * cells : nullptr;
            float *dst = tile + (size_t)ag * rec;
            for (int cell = lane; cell < cells; cell += 32) {
                int vy = cell / G.view_w, vx = cell - vy * G.view_w;
                obs_compose_cell(E, O.curmask, h.a, g, h.x, h.y, h.cx, h.cy, vy, vx, mm, dst + cell * C);
            }
        }
        // features go straight to global memory, element-wise (coalesced)
        for (int q = threadIdx.x; q < cnt * F; q += OBS_THREADS) {
            int ag = q / F, k = q - ag * F;
            O.feature[(size_t)(t0 + ag) * F + k] = obs_feature_elem(E, O.curmask, hdr[ag].a, g, hdr[ag].i, k);
        }
        __syncthreads();
        // stream the tile out
        float *gout = O.view + (size_t)t0 * rec;
        const int nflt = cnt * rec;
        if ((((size_t)gout) & 15) == 0) {
            const int nv = nflt >> 2;
            const float4 *src4 = (const float4 *)tile;
            float4 *dst4 = (float4 *)gout;
            for (int q = threadIdx.x; q < nv; q += OBS_THREADS) __stcs(dst4 + q, src4[q]);
            for (int q = (nv << 2) + threadIdx.x; q < nflt; q += OBS_THREADS) __stcs(gout + q, tile[q]);
        } else {
            for (int q = threadIdx.x; q < nflt; q += OBS_THREADS) __stcs(gout + q, tile[q]);
        }
        __syncthreads();
    }
}

void launch_obs(const EngineDev *dE, const EngineDev &hE, const ObsArgs &O, const float *mm_val, int n_total) {
    const GroupDev &G = hE.grp[O.group];
    const size_t smem = (size_t)OBS_TA * G.view_w * G.view_h * hE.n_channel * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(obs_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int tiles = (n_total + OBS_TA - 1) / OBS_TA;
    int grid = tiles < 16 * g_sms ? tiles : 16 * g_sms;
    if (g_profile) CUDA_CHECK(cudaEventRecord(g_ev0, 0));
    obs_render_kernel<<<grid, OBS_THREADS, smem>>>(dE, O, mm_val, n_total);
    post_launch("obs_render_kernel");
    if (g_profile) {
        CUDA_CHECK(cudaEventRecord(g_ev1, 0));
        CUDA_CHECK(cudaEventSynchronize(g_ev1));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
        g_obs_ms += ms; ++g_obs_launches;
    }
}

}  // namespace be
}  // namespace mg
