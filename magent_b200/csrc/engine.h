// engine.h -- host side of the B200 grid-world engine (plain C++, no CUDA headers).
//
// Mirrors the reference GridWorld class surface (src/gridworld/GridWorld.h:38-128) behind the C ABI in
// shim.cc.  Setup-time work (agent types and range tables, reward-rule compilation, episode placement
// with the engine RNG) runs on the host against a host image of the arenas; the first hot-path call
// uploads the image to HBM and from then on every step-loop call is a kernel launch (backend.h).
#pragma once
#include <stdint.h>
#include <map>
#include <string>
#include <vector>
#include "dev_types.h"

namespace mg {
namespace be { struct Ctx; }

[[noreturn]] void fatal(const char *fmt, ...);

// Range tables (reference src/gridworld/Range.h:104-190)
struct RangeTab {
    int width = 0, height = 0, count = 0;
    int x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    std::vector<unsigned char> mask;    // is_in_range, row-major [height][width]
    std::vector<int> dx, dy;            // num2delta
};
RangeTab make_circle_range(float radius, float inner_radius, int parity);
RangeTab make_sector_range(float angle, float radius, int parity);

// reference src/gridworld/AgentType.h:17-48
struct AgentTypeDef {
    std::string name;
    int width = 1, length = 1;
    float speed = 1.0f, hp = 1.0f;
    float view_radius = 1, view_angle = 360, attack_radius = 0, attack_angle = 0;
    float hear_radius = 0, speak_radius = 0;
    int speak_ability = 0;
    float damage = 0, trace = 0, eat_ability = 0, step_recover = 0, kill_supply = 0, food_supply = 0;
    bool attack_in_group = false, can_absorb = false;
    float step_reward = 0, kill_reward = 0, dead_penalty = 0, attack_penalty = 0;
    int view_x_offset = 0, view_y_offset = 0, att_x_offset = 0, att_y_offset = 0;
    RangeTab view, attack, move;
    int move_base = 0, turn_base = 0, attack_base = 0, n_action = 0;
};

struct HostGroup {                      // one group of one arena, in vector order
    std::vector<int> x, y, id, act, op_obj;
    std::vector<float> hp, next_reward, last_reward;
    std::vector<unsigned char> last_op, flags, dir;
    int dead_ct = 0;
    int n_cull = 0;                     // size after the last clear_dead (reference Agent::index is refreshed only there)
    float grp_reward = 0.0f;
    int size() const { return (int)x.size(); }
    void clear();
    void resize(int n);
};

struct HostArena {
    std::vector<int> occ;
    std::vector<float> food;            // food_mode: amount per OCC_FOOD cell
    std::vector<HostGroup> groups;
    uint32_t rng = 1;
    int id_counter = 0;
    int done = 0;
};

struct SymbolDef { int group = 0, index = 0; };
struct NodeDef { int op = OP_NULL; std::vector<int> raw; };
struct RuleDef { int on = 0; std::vector<int> recv; std::vector<float> values; bool is_terminal = false, auto_value = false; };

class Engine {
public:
    Engine();
    ~Engine();

    // ---- reference Environment / GridWorld surface
    void set_config(const char *key, void *p_value);
    void register_agent_type(const char *name, int n, const char **keys, float *values);
    void new_group(const char *type_name, int *handle);
    void define_agent_symbol(int no, int group, int index);
    void define_event_node(int no, int op, int *inputs, int n_inputs);
    void add_reward_rule(int on, int *receivers, float *values, int n_receiver, bool is_terminal, bool auto_value);

    void reset();
    void add_agents(int group, int n, const char *method, const int *pos_x, const int *pos_y, const int *dir);
    void get_observation(int group, void **bufs, int half = 0);     // half: compact f16 hand-off (extension)
    void set_action(int group, const int *actions);
    void step(int *done);
    void get_reward(int group, float *buf);
    void clear_dead();
    void get_info(int group, const char *name, void *buf);
    void set_goal(int group, const char *method, const int *buf);
    void render();

    void render_next_file();

    // ---- extensions
    void select_arena(int a) { sel_arena_ = a; }
    void random_actions(int group, unsigned long long seed);
    int get_counters(long long *out, int cap);
    void sync();
    void get_io_stats(long long *out, int cap);      // bytes so far: device->host over PCIe, host->device, written by host threads
    // CUDA graphs for launch-bound workloads: the device-pointer calls between begin and end become one graph
    void graph_begin();
    int graph_end();
    void graph_launch(int id, int times);
    void *stream();                      // cudaStream_t of this engine's kernels (nullptr before the first device call)
    void set_profiling(bool on);
    void get_profile(double *ms, long long *n);

private:
    // configuration
    int W_ = 0, H_ = 0;
    bool food_mode_ = false, turn_mode_ = false, minimap_mode_ = false, goal_mode_ = false;
    int embedding_size_ = 0;
    int A_ = 1;
    int device_id_ = -1;
    be::Ctx *bx_ = nullptr;             // this engine's device context (device, streams, scratch): backend.h
    std::map<std::string, AgentTypeDef> types_;
    std::vector<const AgentTypeDef *> group_type_;
    std::vector<SymbolDef> symbols_;
    std::vector<NodeDef> nodes_;
    std::vector<RuleDef> rules_;
    std::vector<RuleDev> compiled_rules_;
    int n_allq_ = 0;
    bool may_have_dead_ = true;         // some arena reported dead agents since the last clear_dead (bit 1 of EngineDev::done)
    bool rules_compiled_ = false;
    bool was_reset_ = false;
    int nsep_ = 1;
    bool large_map_ = false;
    int sel_arena_ = -1;

    // replay dump (reference RenderGenerator, src/gridworld/RenderGenerator.cc)
    std::string render_dir_;
    int file_ct_ = 0, frame_ct_ = 0, frame_per_file_ = 10000;
    bool first_render_ = true;
    struct AttackEvent { int id, x, y; };
    std::vector<AttackEvent> attack_events_;
    void collect_attack_events();

    // host image
    std::vector<HostArena> arenas_;
    enum Where { HOST, DEVICE } where_ = HOST;
    std::vector<int> h_off_;            // [G][A+1] prefix counts (valid in both states)

    // device image
    EngineDev hE_;                      // host copy of the device block (device pointers inside)
    EngineDev *dE_ = nullptr;
    std::vector<void *> dev_allocs_;
    std::vector<int> cap_;              // per group capacity on the device
    unsigned curmask_ = 0;
    std::vector<int> order_;            // groups with set_action this step, in call order
    void *d_view_stage_ = nullptr, *d_feat_stage_ = nullptr;
    float *d_mm_val_ = nullptr;
    size_t view_stage_bytes_ = 0, feat_stage_bytes_ = 0;
    void *d_io_stage_ = nullptr; size_t io_stage_bytes_ = 0;
    void *h_feat_stage_ = nullptr; size_t h_feat_stage_bytes_ = 0;      // page-locked staging of the feature rows (host path)
    enum { IO_D2H = 0, IO_H2D = 1, IO_HOST_WRITTEN = 2, IO_US_WIRE = 3, IO_US_EXPAND = 4, IO_US_FEATURE = 5, IO_N = 6 };
    long long io_[IO_N] = {0, 0, 0, 0, 0, 0};       // step-loop traffic (hot calls only) and host-path phase times
    bool done_stale_ = false;           // arenas_[a].done not refreshed by the last step (device-pointer done)
    int host_path_ = -1;                // env_get_observation into host memory: 1 wire records + host expansion, 0 dense DMA
    unsigned long long rand_calls_ = 0;
    // get_observation pre-pass products (hp_norm plane, minimap) stay valid until the state changes
    unsigned long long state_version_ = 1, prep_version_ = 0;
    int prep_vw_ = 0, prep_vh_ = 0;
    bool prep_skip_absorbed_ = false;

    int G() const { return (int)group_type_.size(); }
    int group2channel(int g) const;
    int n_channel() const { return group2channel(G()); }
    int feature_size(int g) const;
    int total(int g) const { return h_off_[(size_t)g * (A_ + 1) + A_]; }
    int count(int g, int a) const { return h_off_[(size_t)g * (A_ + 1) + a + 1] - h_off_[(size_t)g * (A_ + 1) + a]; }
    void refresh_host_counts();
    bool counts_unknown_ = false;       // steps ran inside replayed graphs: the device's offset table is ahead of h_off_
    unsigned capture_mask_ = 0; int capture_culls_ = 0;
    bool counts_pending_ = false;       // clear_dead queued a fetch of the new counts; h_off_ still holds the old (upper-bound) ones
    void settle_counts();
    void check_group(int g, const char *where) const;

    void ensure_backend();
    void compile_rules();
    void to_device();
    void to_host(bool keep_device_authoritative);
    void free_device();
    void *dalloc(size_t bytes);
    void *io_stage(size_t bytes);
    void stage_reserve(void *&p, size_t &have, size_t need);
    void get_observation_wire(int group, void **bufs);
    int max_agents_per_arena() const;

    // host-side placement (reference Map.cc:49-115, GridWorld.cc:180-290)
    bool host_is_blank(const HostArena &ar, int x, int y, int w, int h) const;
    int host_add_agent(HostArena &ar, int g, int x, int y, int dir);
    int host_add_wall(HostArena &ar, int x, int y);
    void host_random_blank(HostArena &ar, int w, int h, int &x, int &y);
};

}  // namespace mg
