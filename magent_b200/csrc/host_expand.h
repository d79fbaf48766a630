// host_expand.h -- the host half of env_get_observation for HOST buffers.
//
// The reference writes its observations straight into the caller's host memory: one memset of the whole buffer, then a
// handful of sparse writes per observer (src/gridworld/GridWorld.cc:310-311, 363-397).  The B200 engine does the gather on
// the GPU; for a host buffer it ships the result as compact wire records (backend.h: WireHdr / WireMark, ~60 B per observer
// instead of 4732 B) and the threads of this file write the dense float32 records the ABI promises: the arena's minimap
// channels, the self marker, the marked cells, zeros everywhere else -- every byte of the caller's buffer, exactly the
// bytes the reference leaves there.  Plain C++ (no CUDA): non-temporal stores out of a cache-resident record.
#pragma once
#include <stddef.h>
#include <functional>
#include "backend.h"

namespace mg {

struct ExpandGeom {
    int rec;                             // floats per record = view_h * view_w * n_channel
    int C, cells, G;
    int minimap;
    int mm_ch[MG_MAX_GROUPS];            // observation channel of group j's minimap (GridWorld.cc:897-913)
};

// number of threads the host side uses (callers included): MAGENT_B200_HOST_THREADS, else the usable cores (affinity
// mask capped by the cgroup CPU quota) divided by LOCAL_WORLD_SIZE, at most 16 (24 on hosts with 48+ usable cores)
int host_threads();
void set_host_threads(int n);           // 0 = back to the default

// Run fn(tid) for tid in [0, n_threads) on the process-wide worker pool; the calling thread is tid 0.  Returns when all
// have returned.  Jobs of different engines are serialised.
void host_parallel(int n_threads, const std::function<void(int)> &fn);

// Write the dense records of W into out[n_total][rec].  wait_wave(w) must return once wave w of the wire staging is in
// host memory; it is called from the calling thread only, for w = 0 .. n_waves-1 in order.
void expand_views(const ExpandGeom &geom, const be::WireDesc &W, float *out, const std::function<void(int)> &wait_wave);

// ---- NUMA: one socket's memory controllers take ~200 GB/s of streamed writes; a two-socket host takes twice that if each
// half of a buffer lives on its own node and is written by threads of that node.
int numa_nodes();                        // nodes with CPUs this process may use (1 when the host is not NUMA)
// `bytes` of zeroed, page-aligned host memory in 32 MB stripes that alternate over the nodes: every stripe was first touched
// -- and is therefore resident -- on its node, and any prefix of the block is balanced (populations shrink).  expand_views()
// recognises pointers into such a block and deals the chunks that start in a node's stripes to the threads pinned to that
// node.  nullptr when the mapping fails.
void *numa_split_alloc(size_t bytes);
bool numa_split_free(void *p);           // false when p is not a block of numa_split_alloc
size_t numa_split_size(const void *p);   // 0 when p is not the base of such a block

// dst[0, bytes) = src[0, bytes) with all pool threads (non-temporal stores)
void parallel_copy(void *dst, const void *src, size_t bytes);

}  // namespace mg
