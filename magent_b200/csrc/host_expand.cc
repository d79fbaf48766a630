// host_expand.cc -- see host_expand.h.  Plain C++: a small persistent thread pool and the wire -> dense expansion.
#include "host_expand.h"

#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace mg {

// ---------------------------------------------------------------------------------------------
// how many host threads
static int usable_cores() {
    int n = 1;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {                 // cgroup v2: "<quota> <period>" or "max <period>"
        char q[64]; long period = 0;
        if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
            long quota = atol(q);
            if (quota > 0) { int c = (int)(quota / period); if (c < 1) c = 1; if (c < n) n = c; }
        }
        fclose(f);
    } else if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {     // cgroup v1
        long quota = -1, period = 0;
        if (fscanf(g, "%ld", &quota) != 1) quota = -1;
        fclose(g);
        if (FILE *h = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(h, "%ld", &period) != 1) period = 0; fclose(h); }
        if (quota > 0 && period > 0) { int c = (int)(quota / period); if (c < 1) c = 1; if (c < n) n = c; }
    }
    return n < 1 ? 1 : n;
}

static std::atomic<int> g_threads_override{0};
void set_host_threads(int n) { g_threads_override.store(n < 0 ? 0 : (n > 256 ? 256 : n)); }

int host_threads() {
    const int o = g_threads_override.load(std::memory_order_relaxed);
    if (o > 0) return o;
    static const int n = []() {
        if (const char *e = getenv("MAGENT_B200_HOST_THREADS")) { int v = atoi(e); if (v >= 1) return v > 256 ? 256 : v; }
        int c = usable_cores();
        if (const char *w = getenv("LOCAL_WORLD_SIZE")) { int v = atoi(w); if (v > 1) c = c / v; }     // torchrun: ranks share the box
        // the memory system saturates well before the core count of a big host (profiles/README.md, round 2: one socket
        // takes its ~200 GB/s from 8 streaming threads, the expansion's bookkeeping costs ~20 % on top): 16 threads, 24
        // when there are cores to spare
        const int cap = c >= 48 ? 24 : 16;
        if (c > cap) c = cap;
        return c < 1 ? 1 : c;
    }();
    return n;
}

// ---------------------------------------------------------------------------------------------
// NUMA topology from sysfs (no libnuma in the image): the nodes that own CPUs, and their CPU sets
namespace {
struct Topo {
    std::vector<cpu_set_t> node_cpus;    // one entry per node that has CPUs
    std::vector<int> cpu_node;           // cpu -> index into node_cpus, -1 unknown
};
const Topo &topo() {
    static const Topo t = []() {
        Topo T;
        if (const char *e = getenv("MAGENT_B200_NUMA")) { if (!strcmp(e, "off") || !strcmp(e, "0")) return T; }
        if (const char *e = getenv("MAGENT_B200_NUMA_FAKE")) {       // tests on one-node hosts: n pretend nodes over the same CPUs
            const int n = atoi(e);
            if (n >= 2 && n <= 8) {
                cpu_set_t all; CPU_ZERO(&all);
                if (sched_getaffinity(0, sizeof all, &all) != 0) return T;
                for (int k = 0; k < n; ++k) T.node_cpus.push_back(all);
                for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &all)) { if ((int)T.cpu_node.size() <= c) T.cpu_node.resize(c + 1, -1); T.cpu_node[c] = c % n; }
                return T;
            }
        }
        for (int node = 0; node < 64; ++node) {
            char path[96];
            snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
            FILE *f = fopen(path, "r");
            if (!f) continue;
            char buf[4096];
            cpu_set_t set; CPU_ZERO(&set);
            int any = 0;
            if (fgets(buf, sizeof buf, f)) {
                for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
                    int lo = 0, hi = 0;
                    const int k = sscanf(tok, "%d-%d", &lo, &hi);
                    if (k < 1) continue;
                    if (k == 1) hi = lo;
                    for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); any = 1; }
                }
            }
            fclose(f);
            if (!any) continue;
            const int idx = (int)T.node_cpus.size();
            T.node_cpus.push_back(set);
            for (int c = 0; c < CPU_SETSIZE; ++c)
                if (CPU_ISSET(c, &set)) { if ((int)T.cpu_node.size() <= c) T.cpu_node.resize(c + 1, -1); T.cpu_node[c] = idx; }
        }
        if (T.node_cpus.size() < 2) { T.node_cpus.clear(); T.cpu_node.clear(); }
        return T;
    }();
    return t;
}
// node a pool thread belongs to: workers alternate over the nodes (and pin themselves there); the calling thread (tid 0)
// stays wherever the caller put it
int node_of_tid(int tid) {
    const Topo &T = topo();
    const int n = (int)T.node_cpus.size();
    if (n < 2) return 0;
    if (tid == 0) {
        const int cpu = sched_getcpu();
        return (cpu >= 0 && cpu < (int)T.cpu_node.size() && T.cpu_node[cpu] >= 0) ? T.cpu_node[cpu] : 0;
    }
    // ranks of one box (torchrun) start their rotation at different nodes: with few threads per rank the workers of all
    // ranks would otherwise crowd one node
    static const int rank_shift = []() { const char *e = getenv("LOCAL_RANK"); return e ? atoi(e) : 0; }();
    return (tid + rank_shift) % n;
}
}  // namespace

int numa_nodes() { const int n = (int)topo().node_cpus.size(); return n < 2 ? 1 : n; }

// ---------------------------------------------------------------------------------------------
// worker pool: threads are created on first use and live for the life of the process (never joined: the
// pool object is intentionally leaked so that no destructor races with a worker at exit)
namespace {
struct Pool {
    std::mutex job_mu;                   // one job at a time
    std::mutex mu;
    std::condition_variable cv;
    const std::function<void(int)> *fn = nullptr;
    int n = 0;                           // threads taking part in the current job (tid < n)
    unsigned long gen = 0;
    std::atomic<int> remaining{0};
    std::vector<std::thread> workers;    // worker k has tid k + 1

    void worker(int tid) {
        if (numa_nodes() > 1) {             // stay on one node: the chunks dealt to this thread live in that node's memory
            const cpu_set_t &set = topo().node_cpus[node_of_tid(tid)];
            pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), &set);      // best effort (a cpuset may forbid it)
        }
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int)> *f = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return gen != seen; });
                seen = gen;
                if (tid < n) f = fn;
            }
            if (f) {
                (*f)(tid);
                remaining.fetch_sub(1, std::memory_order_acq_rel);
            }
        }
    }
    void run(int want, const std::function<void(int)> &f) {
        std::lock_guard<std::mutex> job(job_mu);
        if (want < 1) want = 1;
        while ((int)workers.size() < want - 1) {
            const int tid = (int)workers.size() + 1;
            workers.emplace_back([this, tid] { worker(tid); });
            workers.back().detach();
        }
        if (want > 1) {
            std::lock_guard<std::mutex> lk(mu);
            fn = &f; n = want;
            remaining.store(want - 1, std::memory_order_release);
            ++gen;
        }
        if (want > 1) cv.notify_all();
        f(0);
        while (remaining.load(std::memory_order_acquire) > 0) {
            for (int i = 0; i < 64; ++i) _mm_pause();
            sched_yield();
        }
    }
};
std::atomic<Pool *> g_pool{nullptr};
std::mutex g_pool_mu;
void forget_pool_in_child() { g_pool.store(nullptr); }      // worker threads do not survive fork()
Pool *pool() {
    Pool *p = g_pool.load(std::memory_order_acquire);
    if (p) return p;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    p = g_pool.load();
    if (!p) {
        static bool hooked = false;
        if (!hooked) { pthread_atfork(nullptr, nullptr, forget_pool_in_child); hooked = true; }
        p = new Pool();
        g_pool.store(p, std::memory_order_release);
    }
    return p;
}
}  // namespace

void host_parallel(int n_threads, const std::function<void(int)> &fn) { pool()->run(n_threads, fn); }

// ---------------------------------------------------------------------------------------------
// streaming copy: dst gets n bytes of src; full 64-byte lines of dst are written with non-temporal stores (no read for
// ownership, no cache pollution), the partial lines at either end with ordinary stores
namespace {
__attribute__((target("avx512f"))) void stream_body_512(char *d, const char *s, size_t n) {       // d 64-aligned, n % 64 == 0
    size_t i = 0;
    for (; i + 256 <= n; i += 256) {
        const __m512i a = _mm512_loadu_si512((const void *)(s + i)), b = _mm512_loadu_si512((const void *)(s + i + 64));
        const __m512i c = _mm512_loadu_si512((const void *)(s + i + 128)), e = _mm512_loadu_si512((const void *)(s + i + 192));
        _mm512_stream_si512((__m512i *)(d + i), a); _mm512_stream_si512((__m512i *)(d + i + 64), b);
        _mm512_stream_si512((__m512i *)(d + i + 128), c); _mm512_stream_si512((__m512i *)(d + i + 192), e);
    }
    for (; i < n; i += 64) _mm512_stream_si512((__m512i *)(d + i), _mm512_loadu_si512((const void *)(s + i)));
}
__attribute__((target("avx2"))) void stream_body_256(char *d, const char *s, size_t n) {
    for (size_t i = 0; i < n; i += 64) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(s + i)), b = _mm256_loadu_si256((const __m256i *)(s + i + 32));
        _mm256_stream_si256((__m256i *)(d + i), a); _mm256_stream_si256((__m256i *)(d + i + 32), b);
    }
}
void stream_body_sse2(char *d, const char *s, size_t n) {
    for (size_t i = 0; i < n; i += 16) _mm_stream_si128((__m128i *)(d + i), _mm_loadu_si128((const __m128i *)(s + i)));
}
typedef void (*StreamFn)(char *, const char *, size_t);
StreamFn pick_stream() {
    __builtin_cpu_init();
    if (getenv("MAGENT_B200_HOST_ISA")) {
        const char *e = getenv("MAGENT_B200_HOST_ISA");
        if (!strcmp(e, "sse2")) return stream_body_sse2;
        if (!strcmp(e, "avx2") && __builtin_cpu_supports("avx2")) return stream_body_256;
    }
    if (__builtin_cpu_supports("avx512f")) return stream_body_512;
    if (__builtin_cpu_supports("avx2")) return stream_body_256;
    return stream_body_sse2;
}
const StreamFn g_stream = pick_stream();

// copies of less than a cache line (record sizes are multiples of 4 bytes): no libc call on the per-record path
__attribute__((target("avx512f"))) void small_copy_512(char *d, const char *s, size_t n) {     // n % 4 == 0, n < 64: one masked move
    const __mmask16 m = (__mmask16)((1u << (n >> 2)) - 1u);
    _mm512_mask_storeu_epi32(d, m, _mm512_maskz_loadu_epi32(m, s));
}
void small_copy_words(char *d, const char *s, size_t n) {
    for (size_t i = 0; i < n; i += 4) { uint32_t v; memcpy(&v, s + i, 4); memcpy(d + i, &v, 4); }
}
typedef void (*SmallFn)(char *, const char *, size_t);
const SmallFn g_small = (g_stream == stream_body_512) ? small_copy_512 : small_copy_words;
inline void small_copy(char *d, const char *s, size_t n) {
    if ((n & 3) == 0) g_small(d, s, n);
    else memcpy(d, s, n);
}

// A sequential byte stream into the caller's buffer.  Records are not multiples of 64 bytes, so the last bytes of one
// record and the first bytes of the next share a cache line: they are collected in `carry` and leave as ONE full-line
// non-temporal store -- no partial-line store (read for ownership) and no libc call per record.  Only the first and the
// last line of a chunk can be partial.
struct LineWriter {
    alignas(64) char carry[64];
    size_t carry_n = 0;
    char *d = nullptr;                       // next destination byte; 64-byte aligned whenever bytes are carried
    void begin(char *dst) { d = dst; carry_n = 0; }
    void put(const char *s, size_t n) {
        if (carry_n) {
            size_t need = 64 - carry_n;
            if (need > n) need = n;
            small_copy(carry + carry_n, s, need);
            carry_n += need; s += need; n -= need;
            if (carry_n == 64) { g_stream(d, carry, 64); d += 64; carry_n = 0; }
            if (!n) return;
        } else {
            size_t head = (size_t)(-(uintptr_t)d) & 63;        // only the first record of a chunk can start off a line
            if (head) {
                if (head > n) head = n;
                memcpy(d, s, head);
                d += head; s += head; n -= head;
                if (!n) return;
            }
        }
        const size_t body = n & ~(size_t)63;
        if (body) { g_stream(d, s, body); d += body; s += body; n -= body; }
        if (n) { small_copy(carry, s, n); carry_n = n; }
    }
    void end() { if (carry_n) { memcpy(d, carry, carry_n); d += carry_n; carry_n = 0; } }
};

inline void stream_out(char *d, const char *s, size_t n) {
    size_t head = (size_t)(-(uintptr_t)d) & 63;
    if (head > n) head = n;
    if (head) memcpy(d, s, head);
    const size_t body = (n - head) & ~(size_t)63;
    if (body) g_stream(d + head, s + head, body);
    const size_t tail = n - head - body;
    if (tail) memcpy(d + head + body, s + head + body, tail);
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// NUMA-split blocks
namespace {
// stripes of `stripe` bytes alternate over the nodes, so that any prefix of the block (populations shrink) is balanced
struct SplitBlock { char *base; size_t bytes; int parts; size_t stripe; };
std::mutex g_split_mu;
std::vector<SplitBlock> g_split;
bool split_lookup(const void *p, SplitBlock *out) {
    std::lock_guard<std::mutex> lk(g_split_mu);
    for (const SplitBlock &b : g_split)
        if ((const char *)p >= b.base && (const char *)p < b.base + b.bytes) { *out = b; return true; }
    return false;
}
}  // namespace

void *numa_split_alloc(size_t bytes) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t len = (bytes + page - 1) / page * page;
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    const int parts = numa_nodes();
    SplitBlock b;
    b.base = (char *)p; b.bytes = len; b.parts = parts;
    b.stripe = (size_t)32 << 20;
    // first touch decides where a page lives: every stripe is zeroed by a pool thread of its node
    int T = host_threads();
    if (T < parts) T = parts;
    const size_t n_stripes = (len + b.stripe - 1) / b.stripe;
    std::vector<std::atomic<size_t>> next(parts);
    for (auto &x : next) x.store(0);
    host_parallel(T, [&](int tid) {
        const int node = parts > 1 ? node_of_tid(tid) % parts : 0;
        for (;;) {
            const size_t i = next[node].fetch_add(1);                 // i-th stripe of this node
            const size_t st = i * parts + node;
            if (st >= n_stripes) break;
            const size_t o = st * b.stripe;
            memset(b.base + o, 0, std::min(b.stripe, len - o));
        }
    });
    // a node without a thread (very few threads) leaves its part untouched: it is zero anyway (anonymous mapping)
    {
        std::lock_guard<std::mutex> lk(g_split_mu);
        g_split.push_back(b);
    }
    return p;
}

bool numa_split_free(void *p) {
    SplitBlock b;
    {
        std::lock_guard<std::mutex> lk(g_split_mu);
        auto it = std::find_if(g_split.begin(), g_split.end(), [&](const SplitBlock &x) { return x.base == (char *)p; });
        if (it == g_split.end()) return false;
        b = *it;
        g_split.erase(it);
    }
    munmap(b.base, b.bytes);
    return true;
}

size_t numa_split_size(const void *p) {
    std::lock_guard<std::mutex> lk(g_split_mu);
    for (const SplitBlock &b : g_split) if (b.base == (const char *)p) return b.bytes;
    return 0;
}

void parallel_copy(void *dst, const void *src, size_t bytes) {
    const size_t piece = (size_t)1 << 20;
    const size_t n_pieces = (bytes + piece - 1) / piece;
    int T = host_threads();
    if ((size_t)T > n_pieces) T = (int)(n_pieces ? n_pieces : 1);
    std::atomic<size_t> next{0};
    host_parallel(T, [&](int) {
        for (;;) {
            const size_t p = next.fetch_add(1);
            if (p >= n_pieces) break;
            const size_t o = p * piece, n = bytes - o < piece ? bytes - o : piece;
            stream_out((char *)dst + o, (const char *)src + o, n);
        }
        _mm_sfence();
    });
}

// ---------------------------------------------------------------------------------------------
// wire -> dense.  Each thread keeps ONE record in a cache-resident scratch buffer: the zeros and the minimap channels of
// the arena it is working in.  Per observer it undoes the previous observer's marks and self marker, applies the new
// ones and streams the record to the caller's buffer; the record is rebuilt only when the arena changes.
namespace {
struct Scratch {
    float *rec = nullptr;
    int arena = -1;
    const be::WireMark *prev = nullptr;
    int prev_n = 0;
    int prev_self = 0xffff;
};

inline void rebuild(const ExpandGeom &g, const be::WireDesc &W, Scratch &s, int arena) {
    memset(s.rec, 0, (size_t)g.rec * sizeof(float));
    if (g.minimap) {
        const float *rows = W.mm + (size_t)arena * W.mm_stride;
        for (int j = 0; j < g.G; ++j) {
            float *d = s.rec + g.mm_ch[j];
            const float *r = rows + (size_t)j * g.cells;
            for (int c = 0; c < g.cells; ++c) d[(size_t)c * g.C] = r[c];           // GridWorld.cc:374-383
        }
    }
    s.arena = arena; s.prev = nullptr; s.prev_n = 0; s.prev_self = 0xffff;
}

// bring scratch record s to show observer o (whose marks start at m): undo what it showed before, apply the new marks
inline void prepare(const ExpandGeom &g, const be::WireDesc &W, Scratch &s, const be::WireHdr &h, const be::WireMark *m) {
    float *rec = s.rec;
    if (h.arena != s.arena) rebuild(g, W, s, h.arena);
    else {
        for (int k = 0; k < s.prev_n; ++k) {                                      // undo the previous observer
            const unsigned off = s.prev[k].off;
            rec[off & ~be::WIRE_HAS_HP] = 0.0f;
            if (off & be::WIRE_HAS_HP) rec[(off & ~be::WIRE_HAS_HP) + 1] = 0.0f;
        }
        if (g.minimap && s.prev_self != 0xffff) {
            const float *rows = W.mm + (size_t)s.arena * W.mm_stride;
            for (int j = 0; j < g.G; ++j) rec[(size_t)s.prev_self * g.C + g.mm_ch[j]] = rows[(size_t)j * g.cells + s.prev_self];
        }
    }
    if (g.minimap && h.self_cell != 0xffff) {                                     // self marker, GridWorld.cc:382
        const float *rows = W.mm + (size_t)h.arena * W.mm_stride;
        for (int j = 0; j < g.G; ++j) {
            const float v = rows[(size_t)j * g.cells + h.self_cell];
            if (v == v) rec[(size_t)h.self_cell * g.C + g.mm_ch[j]] = v + 1.0f;   // 0/0 of an empty group stays the NaN it is
        }
    }
    for (int k = 0; k < (int)h.count; ++k) {                                      // Map::extract_view, Map.cc:183-199
        const unsigned off = m[k].off;
        rec[off & ~be::WIRE_HAS_HP] = 1.0f;
        if (off & be::WIRE_HAS_HP) rec[(off & ~be::WIRE_HAS_HP) + 1] = m[k].val;
    }
    s.prev = m; s.prev_n = h.count; s.prev_self = h.self_cell;
}

// Two scratch records take turns: while record i streams out of one, record i + 1 is prepared in the other -- the 64-byte
// loads of the copy never hit a 4-byte store that is still in flight (no failed store-to-load forwarding).
void expand_chunk(const ExpandGeom &g, const be::WireDesc &W, float *out, int chunk, Scratch (&s)[2]) {
    const size_t o0 = (size_t)chunk * be::WIRE_CHUNK;
    const size_t o1 = o0 + be::WIRE_CHUNK < (size_t)W.n_total ? o0 + be::WIRE_CHUNK : (size_t)W.n_total;
    if (o0 >= o1) return;
    const be::WireMark *m = W.marks + W.chunk_base[chunk];
    const size_t rec_bytes = (size_t)g.rec * sizeof(float);
    LineWriter lw;
    lw.begin((char *)(out + o0 * (size_t)g.rec));
    be::WireHdr h = W.hdr[o0];
    prepare(g, W, s[0], h, m);
    m += h.count;
    int cur = 0;
    for (size_t o = o0; o < o1; ++o, cur ^= 1) {
        if (o + 1 < o1) {
            h = W.hdr[o + 1];
            prepare(g, W, s[cur ^ 1], h, m);
            m += h.count;
        }
        lw.put((const char *)s[cur].rec, rec_bytes);
    }
    lw.end();
}
}  // namespace

void expand_views(const ExpandGeom &geom, const be::WireDesc &W, float *out, const std::function<void(int)> &wait_wave) {
    int T = host_threads();
    if (T > W.n_chunks) T = W.n_chunks > 0 ? W.n_chunks : 1;
    // chunk ranges per NUMA part of the caller's buffer (one range when the buffer is not a numa_split_alloc block):
    // part k holds the chunks whose first byte lies in it
    const size_t rec_bytes = (size_t)geom.rec * sizeof(float), chunk_bytes = rec_bytes * be::WIRE_CHUNK;
    int n_parts = 1;
    std::vector<int> part_chunks[8];          // chunk ids of every part, ascending
    SplitBlock blk;
    if (numa_nodes() > 1 && T > 1 && split_lookup(out, &blk) && blk.parts > 1 && blk.parts <= 8) {
        n_parts = blk.parts;
        const size_t off0 = (size_t)((const char *)out - blk.base);
        for (int c = 0; c < W.n_chunks; ++c) part_chunks[((off0 + (size_t)c * chunk_bytes) / blk.stripe) % n_parts].push_back(c);
    } else if (numa_nodes() > 1 && numa_nodes() <= 8 && T > 1) {
        // somebody else's buffer (plain numpy memory of the reference wrapper, ...): chunk c always goes to the threads of
        // node c mod n.  Pages nobody has touched yet become resident where they are first written -- on that node -- and
        // every later call writes them from the same node again
        n_parts = numa_nodes();
        for (int c = 0; c < W.n_chunks; ++c) part_chunks[c % n_parts].push_back(c);
    } else {
        part_chunks[0].resize(W.n_chunks);
        for (int c = 0; c < W.n_chunks; ++c) part_chunks[0][c] = c;
    }
    std::atomic<int> next[8];
    for (int k = 0; k < 8; ++k) next[k].store(0);
    std::atomic<int> wave_ready[16];
    for (int w = 0; w < 16; ++w) wave_ready[w].store(0);
    int fetched = 0;                         // single-thread mode: waves fetched so far (in queue order)
    auto work = [&](int tid) {
        Scratch s[2];
        void *mem = nullptr;
        const size_t rec_alloc = ((size_t)geom.rec * sizeof(float) + 127) & ~(size_t)63;
        if (posix_memalign(&mem, 64, 2 * rec_alloc) != 0) abort();
        s[0].rec = (float *)mem;
        s[1].rec = (float *)((char *)mem + rec_alloc);
        if (tid == 0 && T == 1) {            // a single thread has to fetch the waves itself, in step with its work
            for (int c = 0; c < W.n_chunks; ++c) {
                const int w = c / W.chunks_per_wave;
                while (!wave_ready[w].load(std::memory_order_relaxed)) { const int q = W.wave_order[fetched++]; wait_wave(q); wave_ready[q].store(1); }
                expand_chunk(geom, W, out, c, s);
            }
        } else {
            if (tid == 0)                     // the calling thread owns the CUDA side: it announces the waves as they land
                for (int q = 0; q < W.n_waves; ++q) { const int w = W.wave_order[q]; wait_wave(w); wave_ready[w].store(1, std::memory_order_release); }
            const int home = n_parts > 1 ? node_of_tid(tid) % n_parts : 0;
            for (int turn = 0; turn < n_parts; ++turn) {       // own part first, then help the others (remote writes)
                const int k = (home + turn) % n_parts;
                for (;;) {
                    const int i = next[k].fetch_add(1);
                    if (i >= (int)part_chunks[k].size()) break;
                    const int c = part_chunks[k][i];
                    const int w = c / W.chunks_per_wave;
                    while (!wave_ready[w].load(std::memory_order_acquire)) { for (int i = 0; i < 32; ++i) _mm_pause(); }
                    expand_chunk(geom, W, out, c, s);
                }
            }
        }
        _mm_sfence();
        free(mem);
    };
    host_parallel(T, work);
}

}  // namespace mg
