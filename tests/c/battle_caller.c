/*
 * battle_caller.c -- a plain-C caller of the drop-in boundary (include/magent_runtime_api.h), test infrastructure.
 *
 *     battle_caller <library.so> [steps] [map_size] [agents_per_group]
 *
 * dlopen()s the given engine library -- the compiled reference (oracle/_ref/libmagent.so), the C restatement or
 * this repository's CUDA library: all export the reference ABI (src/runtime_api.h:20-55) -- builds the battle game
 * the way python/magent/builtin/config/battle.py + gridworld.py:19-115 do, plays `steps` steps of a fixed
 * pseudo-random action stream with HOST buffers and prints one line of FNV-1a checksums per step.  The test
 * (tests/test_c_caller_*.py) runs the same binary against two libraries and requires identical output, so it
 * checks the boundary exactly as INTEGRATION.md section 3 describes it: plain pointers and sizes, no Python.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "magent_runtime_api.h"

#define LOAD(name) do { *(void **)(&p_##name) = dlsym(lib, #name); \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; } } while (0)

static int (*p_env_new_game)(EnvHandle *, const char *);
static int (*p_env_delete_game)(EnvHandle);
static int (*p_env_config_game)(EnvHandle, const char *, void *);
static int (*p_env_reset)(EnvHandle);
static int (*p_env_get_observation)(EnvHandle, GroupHandle, float **);
static int (*p_env_set_action)(EnvHandle, GroupHandle, const int *);
static int (*p_env_step)(EnvHandle, int *);
static int (*p_env_get_reward)(EnvHandle, GroupHandle, float *);
static int (*p_env_get_info)(EnvHandle, GroupHandle, const char *, void *);
static int (*p_gridworld_register_agent_type)(EnvHandle, const char *, int, const char **, float *);
static int (*p_gridworld_new_group)(EnvHandle, const char *, GroupHandle *);
static int (*p_gridworld_add_agents)(EnvHandle, GroupHandle, int, const char *, const int *, const int *, const int *);
static int (*p_gridworld_clear_dead)(EnvHandle);
static int (*p_gridworld_define_agent_symbol)(EnvHandle, int, int, int);
static int (*p_gridworld_define_event_node)(EnvHandle, int, int, int *, int);
static int (*p_gridworld_add_reward_rule)(EnvHandle, int, int *, float *, int, bool, bool);

static uint64_t fnv(uint64_t h, const void *data, size_t n) {
    const unsigned char *p = (const unsigned char *)data;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <library.so> [steps] [map_size] [agents_per_group]\n", argv[0]); return 2; }
    int steps = argc > 2 ? atoi(argv[2]) : 30, size = argc > 3 ? atoi(argv[3]) : 40, n_add = argc > 4 ? atoi(argv[4]) : 150;
    void *lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    LOAD(env_new_game); LOAD(env_delete_game); LOAD(env_config_game); LOAD(env_reset); LOAD(env_get_observation);
    LOAD(env_set_action); LOAD(env_step); LOAD(env_get_reward); LOAD(env_get_info);
    LOAD(gridworld_register_agent_type); LOAD(gridworld_new_group); LOAD(gridworld_add_agents);
    LOAD(gridworld_clear_dead); LOAD(gridworld_define_agent_symbol); LOAD(gridworld_define_event_node);
    LOAD(gridworld_add_reward_rule);

    EnvHandle g = NULL;
    if (p_env_new_game(&g, "GridWorld") != 0 || !g) { fprintf(stderr, "env_new_game failed\n"); return 3; }
    bool yes = true; int emb = 10, seed = 5;
    p_env_config_game(g, "map_width", &size);
    p_env_config_game(g, "map_height", &size);
    p_env_config_game(g, "minimap_mode", &yes);
    p_env_config_game(g, "embedding_size", &emb);
    p_env_config_game(g, "seed", &seed);

    /* builtin/config/battle.py:14-20, flattened as gridworld.py:69-86 does */
    const char *keys[] = {"width", "length", "hp", "speed", "view_radius", "view_angle", "attack_radius", "attack_angle",
                          "damage", "step_recover", "step_reward", "kill_reward", "dead_penalty", "attack_penalty"};
    float vals[] = {1, 1, 10, 2, 6, 360, 1.5f, 360, 2, 0.1f, -0.005f, 5, -0.1f, -0.1f};
    p_gridworld_register_agent_type(g, "small", 14, keys, vals);

    /* battle.py:25-30: a = any agent of group 0, b = any agent of group 1; attack(a,b) -> a +0.2, attack(b,a) -> b +0.2 */
    p_gridworld_define_agent_symbol(g, 0, 0, -1);
    p_gridworld_define_agent_symbol(g, 1, 1, -1);
    int ab[2] = {0, 1}, ba[2] = {1, 0};
    p_gridworld_define_event_node(g, 0, 7 /* OP_ATTACK, grid_def.h:20 */, ab, 2);
    p_gridworld_define_event_node(g, 1, 7, ba, 2);
    int recv_a[1] = {0}, recv_b[1] = {1}; float val[1] = {0.2f};
    p_gridworld_add_reward_rule(g, 0, recv_a, val, 1, false, false);
    p_gridworld_add_reward_rule(g, 1, recv_b, val, 1, false, false);

    GroupHandle grp[2];
    p_gridworld_new_group(g, "small", &grp[0]);
    p_gridworld_new_group(g, "small", &grp[1]);

    int vs[3], fs[3], as[3];
    p_env_get_info(g, grp[0], "view_space", vs);
    p_env_get_info(g, grp[0], "feature_space", fs);
    p_env_get_info(g, grp[0], "action_space", as);
    printf("spaces view %dx%dx%d feature %d actions %d\n", vs[0], vs[1], vs[2], fs[0], as[0]);

    p_env_reset(g);
    int wall_x[4] = {size / 2, size / 2, size / 2 + 1, size / 2 - 1}, wall_y[4] = {size / 2, size / 2 + 1, size / 2, size / 2};
    p_gridworld_add_agents(g, -1, 4, "custom", wall_x, wall_y, NULL);          /* walls: group -1 (GridWorld.cc:180-206) */
    p_gridworld_add_agents(g, grp[0], n_add, "random", NULL, NULL, NULL);
    p_gridworld_add_agents(g, grp[1], n_add, "random", NULL, NULL, NULL);

    size_t view_sz = (size_t)vs[0] * vs[1] * vs[2], cap = (size_t)n_add;
    float *view = (float *)malloc(cap * view_sz * sizeof(float));
    float *feat = (float *)malloc(cap * fs[0] * sizeof(float));
    float *rew = (float *)malloc(cap * sizeof(float));
    int *act = (int *)malloc(cap * sizeof(int)), *pos = (int *)malloc(cap * 2 * sizeof(int)), *ids = (int *)malloc(cap * sizeof(int));
    unsigned char *alive = (unsigned char *)malloc(cap);
    uint32_t lcg = 12345u;
    for (int t = 0; t < steps; ++t) {
        uint64_t h_view = 1469598103934665603ull, h_feat = h_view, h_pos = h_view, h_id = h_view, h_alive = h_view;
        double rsum = 0.0;
        int num[2], done = 0;
        for (int k = 0; k < 2; ++k) {
            p_env_get_info(g, grp[k], "num", &num[k]);
            float *bufs[2] = {view, feat};
            if (num[k] > 0) p_env_get_observation(g, grp[k], bufs);
            h_view = fnv(h_view, view, (size_t)num[k] * view_sz * sizeof(float));
            h_feat = fnv(h_feat, feat, (size_t)num[k] * fs[0] * sizeof(float));
        }
        for (int k = 0; k < 2; ++k) {
            for (int i = 0; i < num[k]; ++i) { lcg = lcg * 1664525u + 1013904223u; act[i] = (int)((lcg >> 8) % (uint32_t)as[0]); }
            p_env_set_action(g, grp[k], act);
        }
        p_env_step(g, &done);
        for (int k = 0; k < 2; ++k) {
            p_env_get_reward(g, grp[k], rew);
            p_env_get_info(g, grp[k], "pos", pos);
            p_env_get_info(g, grp[k], "id", ids);
            p_env_get_info(g, grp[k], "alive", alive);
            for (int i = 0; i < num[k]; ++i) rsum += rew[i];
            h_pos = fnv(h_pos, pos, (size_t)num[k] * 2 * sizeof(int));
            h_id = fnv(h_id, ids, (size_t)num[k] * sizeof(int));
            h_alive = fnv(h_alive, alive, (size_t)num[k]);
        }
        p_gridworld_clear_dead(g);
        printf("t %3d num %d %d done %d view %016llx feat %016llx pos %016llx id %016llx alive %016llx reward %.4f\n",
               t, num[0], num[1], done, (unsigned long long)h_view, (unsigned long long)h_feat,
               (unsigned long long)h_pos, (unsigned long long)h_id, (unsigned long long)h_alive, rsum);
    }
    p_env_delete_game(g);
    free(view); free(feat); free(rew); free(act); free(pos); free(ids); free(alive);
    return 0;
}
