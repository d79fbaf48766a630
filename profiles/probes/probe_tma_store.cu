// probe_tma_store.cu -- what is the ceiling of the render kernel's output pattern?
// Persistent CTAs each own one shared-memory buffer and bulk-store it (cp.async.bulk.global.shared::cta) to
// consecutive `chunk`-byte slots of a large output, round-robin over CTAs, waiting for the previous store's
// smem read-out before re-issuing (exactly the render kernel's store discipline, nothing else).
// usage: probe_tma_store  -> prints GB/s for a sweep of chunk sizes / alignments / CTAs per SM / mode
//   mode 0: TMA bulk store; mode 1: st.global.v4 by all threads (coalesced) from smem; mode 2: mode 0 + wait_group 0
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void store_kernel(unsigned char *out, size_t n_chunks, unsigned chunk, int mode) {
    extern __shared__ __align__(128) unsigned char buf[];
    for (unsigned q = threadIdx.x; q < chunk / 4; q += blockDim.x) ((unsigned *)buf)[q] = q * 2654435761u + blockIdx.x;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        unsigned char *dst = out + c * (size_t)chunk;
        if (mode == 1) {
            for (unsigned q = threadIdx.x; q < chunk / 16; q += blockDim.x) ((uint4 *)dst)[q] = ((uint4 *)buf)[q];
        } else {
            if (threadIdx.x == 0) {
                if (mode == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
            // (the render kernel composes the tile here)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(buf)), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t total = (size_t)2574336000ull;            // bytes of one render launch in bench.py
    unsigned char *out; CK(cudaMalloc(&out, total + (1 << 20)));
    unsigned char *flush; const size_t flush_bytes = 512u << 20; CK(cudaMalloc(&flush, flush_bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const unsigned chunks[] = {18928, 18944, 9472, 37856, 16384, 32768, 65536};
    const int ctas[] = {4, 8, 11, 16};
    CK(cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int mode = 0; mode < 3; ++mode)
        for (unsigned chunk : chunks)
            for (int c : ctas) {
                if ((size_t)c * (chunk + 1024) > 227 * 1024) continue;
                const size_t n_chunks = total / chunk;
                float best = 1e9f, sum = 0;
                const int reps = 5;
                for (int r = 0; r < reps + 1; ++r) {
                    CK(cudaMemsetAsync(flush, r, flush_bytes));
                    CK(cudaEventRecord(e0));
                    store_kernel<<<c * sms, 128, chunk>>>(out, n_chunks, chunk, mode);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (r == 0) continue;
                    if (ms < best) best = ms;
                    sum += ms;
                }
                printf("PROBE mode %d chunk %6u ctas/sm %2d  mean %.3f ms  %.0f GB/s   best %.0f GB/s\n", mode, chunk, c,
                       sum / reps, n_chunks * (double)chunk / (sum / reps) / 1e6, n_chunks * (double)chunk / best / 1e6);
                fflush(stdout);
            }
    // reference point: cudaMemset of the same range
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(out, 1, total)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("PROBE memset %.3f ms %.0f GB/s\n", ms, total / ms / 1e6);
    }
    return 0;
}
