// backend_emu.cc -- TEST INFRASTRUCTURE ONLY.
//
// A single-threaded host implementation of magent_b200/csrc/backend.h that runs the very same
// phase functions (step_phases.h / obs_phases.h) the CUDA kernels run, with a team of one thread.
// Purpose: debug the *parallel formulations* (shuffle replay, rank-ordered attack relaxation, claimant
// list move relaxation, reward programs, compaction) against the compiled reference on development
// containers that have no GPU.  It is built by tests/emu/build.sh into tests/_emu/libmagent_emu.so,
// is never part of magent_b200/lib/libmagent.so, and nothing in the product path can load it.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../magent_b200/csrc/backend.h"
#include "../../magent_b200/csrc/obs_phases.h"

namespace mg {
namespace be {

struct HostCtx {
    int tid() const { return 0; }
    int nth() const { return 1; }
    void sync() {}
    void sync_cta() {}
    bool is_cta_leader() const { return true; }
    GroupEnum en_[2];
    GroupEnum *enums() { return en_; }
    int *flags(ArenaHdr *hdr) { return hdr->changed; }
    void add_count(const EngineDev &E, int kind, long long v) { E.counters[kind] += v; }
    void flush_counts(const EngineDev &) {}
    template <class P, class Em> int scan(int n, P pred, Em emit) {
        int run = 0;
        for (int i = 0; i < n; ++i) if (pred(i)) emit(i, run++);
        return run;
    }
};

// the emulation's "device context": the wire staging of the last obs_wire_begin
struct Ctx {
    std::vector<WireHdr> hdr;
    std::vector<WireMark> marks;
    std::vector<long long> base;
    std::vector<float> mm;
    std::vector<int> counts;
};

const char *name() { return "emu"; }
Ctx *create(int, std::string *) { return new Ctx(); }
void destroy(Ctx *c) { delete c; }
int device_of(const Ctx *) { return 0; }
int device_count() { return 1; }
int sm_count(const Ctx *) { return 1; }
void *stream_handle(const Ctx *) { return nullptr; }
void *dmalloc(Ctx *, size_t b) { return calloc(1, b ? b : 1); }
void dfree(Ctx *, void *p) { free(p); }
void dmemset(Ctx *, void *p, int byte, size_t n) { memset(p, byte, n); }
void h2d(Ctx *, void *d, const void *s, size_t n) { memcpy(d, s, n); }
void d2h(Ctx *, void *d, const void *s, size_t n) { memcpy(d, s, n); }
void d2d(Ctx *, void *d, const void *s, size_t n) { memcpy(d, s, n); }
void *host_alloc(size_t b) { return malloc(b ? b : 1); }
void host_free(void *p) { free(p); }
bool host_register(void *, size_t) { return false; }
void host_unregister(void *) {}
bool is_device_ptr(const void *) { return false; }
bool is_pinned_host_ptr(const void *) { return false; }
void sync(Ctx *) {}

void launch_step(Ctx *, const EngineDev *dE, const EngineDev &, const StepArgs &S, int) {
    HostCtx c;
    for (int a = 0; a < dE->A; ++a) run_step(c, *dE, S, a);
}
void launch_cull(Ctx *, const EngineDev *dE, const EngineDev &, unsigned curmask, int) {
    HostCtx c;
    for (int a = 0; a < dE->A; ++a) run_cull(c, *dE, curmask, a);
}
void launch_offsets(Ctx *, const EngineDev *dE, const EngineDev &) {
    const EngineDev &E = *dE;
    for (int g = 0; g < E.G; ++g) {
        int *off = E.off + (size_t)g * (E.A + 1);
        off[0] = 0;
        for (int a = 0; a < E.A; ++a) off[a + 1] = off[a] + E.n[g * E.A + a];
    }
}
void launch_obs_prepare(Ctx *, const EngineDev *dE, const EngineDev &, unsigned curmask, int og, float *mm_val) {
    if (!mm_val) return;
    const EngineDev &E = *dE;
    const GroupDev &OG = E.grp[og];
    int cells = OG.view_w * OG.view_h;
    for (int a = 0; a < E.A; ++a)
        for (int j = 0; j < E.G; ++j) {
            std::vector<int> cnt(cells, 0);
            int n = E.n[j * E.A + a];
            const AgentSoA &s = cur_soa(E, curmask, j);
            int total = 0;
            for (int i = 0; i < n; ++i) {
                if (OG.can_absorb && (s.flags[gidx(E, a, j, i)] & FLAG_ABSORBED)) continue;   // GridWorld.cc:343-347
                int cx, cy;
                minimap_cell(E, OG.view_w, OG.view_h, s.x[gidx(E, a, j, i)], s.y[gidx(E, a, j, i)], cx, cy);
                cnt[cy * OG.view_w + cx]++;
                total++;
            }
            float *out = mm_val + ((size_t)a * E.G + j) * cells;
            for (int k = 0; k < cells; ++k) out[k] = (float)cnt[k] / (float)total;
        }
}
// software float32 -> binary16, round to nearest even (the emulation's own statement of the f16 hand-off format)
static uint16_t f32_to_f16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
    const uint32_t ex = (u >> 23) & 0xffu, man = u & 0x7fffffu;
    if (ex == 0xffu) {
        if (!man) return sign | 0x7c00u;
        uint16_t r = (uint16_t)(0x7c00u + (man >> 13));
        if (r == 0x7c00u) ++r;
        return sign | r;
    }
    const int e = (int)ex - 127 + 15;
    if (e >= 31) return sign | 0x7c00u;
    if (e <= 0) {
        if (e < -10) return sign;
        const uint32_t m = man | 0x800000u;
        const int shift = 14 - e;
        uint32_t r = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (r & 1u))) ++r;
        return sign | (uint16_t)r;
    }
    uint32_t r = ((uint32_t)e << 10) | (man >> 13);
    const uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
    return sign | (uint16_t)r;
}
void launch_obs(Ctx *, const EngineDev *dE, const EngineDev &, const ObsArgs &O, const float *mm_val, int) {
    const EngineDev &E = *dE;
    int g = O.group;
    const GroupDev &G = E.grp[g];
    int cells = G.view_w * G.view_h, C = E.n_channel;
    const AgentSoA &s = cur_soa(E, O.curmask, g);
    std::vector<float> rec((size_t)cells * C), feat((size_t)G.feature_size);
    for (int a = 0; a < E.A; ++a) {
        int n = E.n[g * E.A + a], base = E.off[(size_t)g * (E.A + 1) + a];
        const float *mm = mm_val ? mm_val + (size_t)a * E.G * cells : nullptr;
        for (int i = 0; i < n; ++i) {
            long gi = gidx(E, a, g, i);
            int cx = -1, cy = -1;
            if (mm) minimap_cell(E, G.view_w, G.view_h, s.x[gi], s.y[gi], cx, cy);
            float *out = O.half ? rec.data() : (float *)O.view + (size_t)(base + i) * cells * C;
            for (int vy = 0; vy < G.view_h; ++vy)
                for (int vx = 0; vx < G.view_w; ++vx)
                    obs_compose_cell(E, O.curmask, a, g, s.x[gi], s.y[gi], s.dir[gi], cx, cy, vy, vx, mm,
                                     out + (size_t)(vy * G.view_w + vx) * C);
            float *fo = O.half ? feat.data() : (float *)O.feature + (size_t)(base + i) * G.feature_size;
            obs_feature(E, O.curmask, a, g, i, fo);
            if (O.half) {
                uint16_t *hv = (uint16_t *)O.view + (size_t)(base + i) * cells * C;
                for (int q = 0; q < cells * C; ++q) hv[q] = f32_to_f16(rec[q]);
                uint16_t *hf = (uint16_t *)O.feature + (size_t)(base + i) * G.feature_size;
                for (int q = 0; q < G.feature_size; ++q) hf[q] = f32_to_f16(feat[q]);
            }
        }
    }
}
void launch_info(Ctx *, const EngineDev *dE, const EngineDev &, unsigned curmask, int kind, int g, void *buf, int) {
    const EngineDev &E = *dE;
    const AgentSoA &s = cur_soa(E, curmask, g);
    for (int a = 0; a < E.A; ++a) {
        int n = E.n[g * E.A + a], base = E.off[(size_t)g * (E.A + 1) + a];
        for (int i = 0; i < n; ++i) {
            long gi = gidx(E, a, g, i);
            int o = base + i;
            switch (kind) {
                case INFO_ID: ((int *)buf)[o] = s.id[gi]; break;
                case INFO_POS: ((int *)buf)[2 * o] = s.x[gi]; ((int *)buf)[2 * o + 1] = s.y[gi]; break;
                case INFO_ALIVE: ((unsigned char *)buf)[o] = (s.flags[gi] & FLAG_DEAD) ? 0 : 1; break;
                case INFO_REWARD: ((float *)buf)[o] = s.next_reward[gi] + E.hdr[a].grp_reward[g]; break;
                case INFO_HP: ((float *)buf)[o] = s.hp[gi]; break;
                case INFO_ACTION_SCATTER: s.act[gi] = ((const int *)buf)[o]; break;
            }
        }
    }
}
void launch_random_actions(Ctx *, const EngineDev *dE, const EngineDev &, unsigned curmask, int g,
                           unsigned long long seed, int) {
    const EngineDev &E = *dE;
    const AgentSoA &s = cur_soa(E, curmask, g);
    unsigned long long x = seed | 1;
    for (int a = 0; a < E.A; ++a)
        for (int i = 0; i < E.n[g * E.A + a]; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            s.act[gidx(E, a, g, i)] = (int)(x % (unsigned)E.grp[g].n_action);
        }
}

void counts_fetch_begin(Ctx *c, const int *dev_off, size_t n) { c->counts.assign(dev_off, dev_off + n); }
const int *counts_fetch_wait(Ctx *c) { return c->counts.data(); }
void read_done(Ctx *, const EngineDev &hE, int *done_words) { memcpy(done_words, hE.done, sizeof(int) * hE.A); }
void launch_done_to_device(Ctx *, const EngineDev *dE, const EngineDev &, int *dev_done) {
    int all = 1;
    for (int a = 0; a < dE->A; ++a) all &= dE->done[a] & 1;
    *dev_done = all;
}

// wire records of one observation (backend.h): derived from the same per-cell composition the dense emulation uses, so
// the engine's host expansion (host_expand.cc) can be checked against the reference without a GPU
void obs_wire_begin(Ctx *c, const EngineDev *dE, const EngineDev &, const ObsArgs &O, const float *mm_val, int n_total,
                    WireDesc *out) {
    const EngineDev &E = *dE;
    const int g = O.group;
    const GroupDev &G = E.grp[g];
    const int cells = G.view_w * G.view_h, C = E.n_channel;
    const AgentSoA &s = cur_soa(E, O.curmask, g);
    const int n_chunks = (n_total + WIRE_CHUNK - 1) / WIRE_CHUNK;
    c->hdr.assign(n_total, WireHdr());
    c->marks.clear();
    c->base.assign(n_chunks + 1, 0);
    const int mm_stride = (E.G * cells + 3) & ~3;
    if (mm_val) {
        c->mm.assign((size_t)E.A * mm_stride, 0.0f);
        for (int a = 0; a < E.A; ++a) memcpy(&c->mm[(size_t)a * mm_stride], mm_val + (size_t)a * E.G * cells, sizeof(float) * E.G * cells);
    }
    std::vector<float> cell(C);
    for (int a = 0; a < E.A; ++a) {
        const int n = E.n[g * E.A + a], base = E.off[(size_t)g * (E.A + 1) + a];
        for (int i = 0; i < n; ++i) {
            const long gi = gidx(E, a, g, i);
            const int o = base + i;
            if (o % WIRE_CHUNK == 0) c->base[o / WIRE_CHUNK] = (long long)c->marks.size();
            int cx = -1, cy = -1;
            if (mm_val) minimap_cell(E, G.view_w, G.view_h, s.x[gi], s.y[gi], cx, cy);
            WireHdr h;
            h.arena = a;
            h.self_cell = mm_val ? (unsigned short)(cy * G.view_w + cx) : (unsigned short)0xffff;
            int count = 0;
            for (int vy = 0; vy < G.view_h; ++vy)
                for (int vx = 0; vx < G.view_w; ++vx) {
                    obs_compose_cell(E, O.curmask, a, g, s.x[gi], s.y[gi], s.dir[gi], -1, -1, vy, vx, nullptr, cell.data());
                    const int w = (vy * G.view_w + vx) * C;
                    if (cell[0] != 0.0f) { c->marks.push_back({(unsigned)w, 0.0f}); ++count; }
                    if (E.food_mode && cell[1] != 0.0f) { c->marks.push_back({(unsigned)(w + 1), 0.0f}); ++count; }
                    for (int j = 0; j < E.G; ++j) {
                        const int ch = obs_channel(E, g, j);
                        if (cell[ch] != 0.0f) { c->marks.push_back({(unsigned)(w + ch) | WIRE_HAS_HP, cell[ch + 1]}); ++count; }
                    }
                }
            h.count = (unsigned short)count;
            c->hdr[o] = h;
            obs_feature(E, O.curmask, a, g, i, (float *)O.feature + (size_t)o * G.feature_size);
        }
    }
    c->base[n_chunks] = (long long)c->marks.size();
    c->marks.push_back({0u, 0.0f});
    out->hdr = c->hdr.data(); out->marks = c->marks.data(); out->chunk_base = c->base.data();
    out->mm = mm_val ? c->mm.data() : nullptr; out->mm_stride = mm_val ? mm_stride : 0;
    out->n_total = n_total; out->n_chunks = n_chunks;
    out->chunks_per_wave = (n_chunks + 7) / 8 < 2 ? 2 : (n_chunks + 7) / 8;
    out->n_waves = (n_chunks + out->chunks_per_wave - 1) / out->chunks_per_wave;
    for (int q = 0; q < out->n_waves; ++q) { const int half = (out->n_waves + 1) / 2; out->wave_order[q] = (q & 1) ? half + (q >> 1) : (q >> 1); }
}
void obs_wire_wait(Ctx *, int) {}
void dma_d2h_async(Ctx *, void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void dma_wait(Ctx *, int) {}

bool capture_begin(Ctx *) { return false; }       // no graphs in the emulation
int capture_end(Ctx *) { return -1; }
bool capturing(const Ctx *) { return false; }
void graph_launch(Ctx *, int) {}
void graph_destroy_all(Ctx *) {}

long long launch_count() { return 0; }
void profile_enable(Ctx *, bool) {}
void profile_read(Ctx *, double *ms, long long *n) { *ms = 0; *n = 0; }

}  // namespace be
}  // namespace mg
