// probe_host_expand.cu -- how fast can the HOST side of env_get_observation be?  (profiles/README.md, round 2)
//
// The reference ABI hands the engine a host buffer that must end up holding one dense float32 record per observer
// (4732 B for battle).  Two ways to fill it: (a) DMA the dense records over PCIe, (b) ship a compact wire record and
// let host threads write the dense bytes (non-temporal stores out of an L1/L2-resident tile).  This probe measures,
// on the box it runs on:  the D2H DMA rate into pinned memory, the aggregate non-temporal store rate of T host threads
// into pinned and into pageable memory, and both at the same time.
//
//   nvcc -O3 -std=c++17 -Xcompiler -mavx2,-pthread probe_host_expand.cu -o probe_host_expand && ./probe_host_expand
#include <cuda_runtime.h>
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__attribute__((target("avx512f"))) static void stream512(char *dst, const char *src, size_t n) {
    for (size_t i = 0; i < n; i += 64) _mm512_stream_si512((__m512i *)(dst + i), _mm512_load_si512((const __m512i *)(src + i)));
}
__attribute__((target("avx2"))) static void stream256(char *dst, const char *src, size_t n) {
    for (size_t i = 0; i < n; i += 32) _mm256_stream_si256((__m256i *)(dst + i), _mm256_load_si256((const __m256i *)(src + i)));
}

// T threads fill [dst, dst + bytes) by repeating a `tile` byte source (cache-resident), chunks dealt by an atomic counter
static double fill(char *dst, size_t bytes, int T, size_t tile, bool avx512, int mode /*0 nt, 1 memcpy, 2 memset*/) {
    const size_t chunk = 4u << 20;
    const size_t n_chunks = bytes / chunk;
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    double t0 = now();
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            char *src = (char *)aligned_alloc(64, tile);
            memset(src, t + 1, tile);
            for (;;) {
                size_t c = next.fetch_add(1);
                if (c >= n_chunks) break;
                char *d = dst + c * chunk;
                for (size_t o = 0; o < chunk; o += tile) {
                    size_t n = chunk - o < tile ? chunk - o : tile;   // tile and chunk are multiples of 64
                    if (mode == 2) memset(d + o, 0, n);
                    else if (mode == 1) memcpy(d + o, src, n);
                    else if (avx512) stream512(d + o, src, n);
                    else stream256(d + o, src, n);
                }
            }
            _mm_sfence();
            free(src);
        });
    for (auto &x : th) x.join();
    return (double)(n_chunks * chunk) / (now() - t0) / 1e9;
}

int main(int argc, char **argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const size_t GB = 1ull << 30;
    const size_t bytes = 4 * GB;
    cpu_set_t set;
    sched_getaffinity(0, sizeof set, &set);
    printf("affinity cpus: %d\n", CPU_COUNT(&set));
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) { char b[128]; if (fgets(b, 128, f)) printf("cgroup cpu.max: %s", b); fclose(f); }
    system("lscpu | egrep 'Model name|Socket|NUMA|Core|Thread' ; nvidia-smi topo -m 2>/dev/null | head -12; cat /proc/meminfo | head -3");
    const bool avx512 = __builtin_cpu_supports("avx512f");
    printf("avx512f: %d\n", (int)avx512);
    char *pinned, *dev;
    double t0 = now();
    CK(cudaHostAlloc((void **)&pinned, bytes, cudaHostAllocDefault));
    printf("cudaHostAlloc 4 GiB: %.2f s\n", now() - t0);
    CK(cudaMalloc((void **)&dev, bytes));
    CK(cudaMemset(dev, 1, bytes));
    char *pageable = (char *)aligned_alloc(4096, bytes);
    t0 = now(); memset(pageable, 0, bytes); printf("first touch of 4 GiB pageable: %.2f s\n", now() - t0);
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));

    // (a) DMA alone
    for (int rep = 0; rep < 2; ++rep) {
        t0 = now();
        CK(cudaMemcpyAsync(pinned, dev, bytes, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        printf("D2H DMA alone (4 GiB, pinned): %.1f GB/s\n", bytes / (now() - t0) / 1e9);
    }
    t0 = now();
    CK(cudaMemcpy(pageable, dev, GB, cudaMemcpyDeviceToHost));
    printf("D2H into PAGEABLE memory (1 GiB): %.1f GB/s\n", GB / (now() - t0) / 1e9);
    // chunked DMA (16 MiB pieces, as the engine would issue them)
    t0 = now();
    for (size_t o = 0; o < bytes; o += 16u << 20) CK(cudaMemcpyAsync(pinned + o, dev + o, 16u << 20, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    printf("D2H DMA in 16 MiB pieces: %.1f GB/s\n", bytes / (now() - t0) / 1e9);

    // (b) host threads alone
    const int Ts[] = {1, 4, 8, 12, 16, 24, 32, 64};
    for (int T : Ts) {
        double nt = fill(pinned, bytes, T, 18944, avx512, 0);
        double nt64 = fill(pinned, bytes, T, 75712, avx512, 0);
        double mc = fill(pinned, bytes, T, 18944, avx512, 1);
        double ms = fill(pinned, bytes, T, 1 << 20, avx512, 2);
        double pg = fill(pageable, bytes, T, 18944, avx512, 0);
        printf("T=%2d  NT(19K tile) %.1f  NT(76K tile) %.1f  memcpy %.1f  memset %.1f  NT->pageable %.1f  GB/s\n", T, nt, nt64, mc, ms, pg);
        fflush(stdout);
    }
    // (c) both: DMA into the upper half while T threads fill the lower half
    for (int T : {8, 12, 14, 16, 24, 32}) {
        std::atomic<int> stop(0);
        double dma_bytes = 0, dma_t = 0;
        std::thread dma([&]() {
            double a = now();
            while (!stop.load()) {
                for (size_t o = 0; o < 2 * GB; o += 16u << 20) cudaMemcpyAsync(pinned + 2 * GB + o, dev + o, 16u << 20, cudaMemcpyDeviceToHost, s);
                cudaStreamSynchronize(s);
                dma_bytes += 2.0 * GB;
            }
            dma_t = now() - a;
        });
        double cpu = 0;
        for (int r = 0; r < 3; ++r) cpu = fill(pinned, 2 * GB, T, 18944, avx512, 0);
        stop = 1;
        dma.join();
        printf("concurrent: T=%2d NT %.1f GB/s + DMA %.1f GB/s = %.1f GB/s\n", T, cpu, dma_bytes / dma_t / 1e9, cpu + dma_bytes / dma_t / 1e9);
        fflush(stdout);
    }
    return 0;
}
