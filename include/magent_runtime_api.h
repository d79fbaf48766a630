/*
 * magent_runtime_api.h -- C ABI of the B200-native grid-world engine (libmagent.so).
 *
 * This is the drop-in boundary: the symbols, argument order and argument meaning are exactly those of the
 * reference runtime library (reference: src/runtime_api.h:20-61, implemented in src/runtime_api.cc:15-163),
 * so the reference's own Python binding (python/magent/gridworld.py, via ctypes) and this repository's
 * mirror of it (magent_b200/gridworld.py) can load either library.  All functions return 0 (the reference
 * always returns 0); env_new_game additionally returns -1 when no CUDA device is usable.  Fatal conditions
 * print a message and abort the process, as the reference's LOG(FATAL) does (src/utility/utility.h:77-103).
 *
 * Buffers are owned by the caller.  Every buffer argument of the step loop (observation, action, reward,
 * id/pos/alive info) may be a HOST pointer (copied across PCIe inside the call) or a CUDA DEVICE pointer
 * (read / written in place, no copy); the engine tells them apart with cudaPointerGetAttributes.
 *
 * Plain C types only: no C++, CUDA or torch types cross this boundary.
 */
#ifndef MAGENT_B200_RUNTIME_API_H
#define MAGENT_B200_RUNTIME_API_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *EnvHandle;   /* reference: src/Environment.h:36  (Environment*)  */
typedef int GroupHandle;   /* reference: src/Environment.h:12                  */

/* ---- game (reference: src/runtime_api.h:20-22) ------------------------------------------------------- */
/* name must be "GridWorld"; "DiscreteSnake" (a different game, out of scope) is refused with -1.         */
int env_new_game(EnvHandle *game, const char *name);
int env_delete_game(EnvHandle game);
/* p_value is read as int / bool / char* depending on the key (reference: src/gridworld/GridWorld.cc:120-149):
 *   map_width, map_height, embedding_size, seed : int      food_mode, turn_mode, minimap_mode, goal_mode : bool
 *   render_dir : char*
 * extension keys (FATAL in the reference, so they cannot collide): num_arenas : int, device_id : int      */
int env_config_game(EnvHandle game, const char *name, void *p_value);

/* ---- run step (reference: src/runtime_api.h:25-29) --------------------------------------------------- */
int env_reset(EnvHandle game);
/* buffer[0] = view float32 [n][view_h][view_w][n_channel] (NHWC), buffer[1] = feature float32 [n][feature]
 * (reference: src/gridworld/GridWorld.cc:292-401).  Every byte of both buffers is written.               */
int env_get_observation(EnvHandle game, GroupHandle group, float **buffer);
/* actions: int32 [n], n = current size of the group incl. dead-but-unculled agents
 * (reference: src/gridworld/GridWorld.cc:403-454).  The order of calls across groups is the move order.  */
int env_set_action(EnvHandle game, GroupHandle group, const int *actions);
int env_step(EnvHandle game, int *done);                               /* GridWorld.cc:456-631 */
int env_get_reward(EnvHandle game, GroupHandle group, float *buffer);  /* GridWorld.cc:694-704 */

/* ---- info getter (reference: src/runtime_api.h:32, src/gridworld/GridWorld.cc:709-894) --------------- */
/* names: num, id, pos, alive, action_space, view_space, feature_space, view2attack, attack_base,
 *        global_minimap, mean_info, walls_info, groups_info, render_window_info, attack_event, both_attack
 * extension names: arena_num (int[num_arenas]), arena_done (int[num_arenas]), hp (float[n], test aid)     */
int env_get_info(EnvHandle game, GroupHandle group, const char *name, void *buffer);

/* ---- render (reference: src/runtime_api.h:35-36): replay dump of arena 0, byte-identical to the files the
 * reference's RenderGenerator writes (src/gridworld/RenderGenerator.cc:107-185); cold path, host snapshot ---- */
int env_render(EnvHandle game);
int env_render_next_file(EnvHandle game);

/* ---- gridworld specials (reference: src/runtime_api.h:42-55) ------------------------------------------ */
int gridworld_register_agent_type(EnvHandle game, const char *name, int n, const char **keys, float *values);
int gridworld_new_group(EnvHandle game, const char *agent_type_name, GroupHandle *group);
/* group == -1 adds walls; method in {"random","custom","fill"}; for "fill" pos_x = {x, y, width, height, dir}
 * (reference: src/gridworld/GridWorld.cc:180-290) */
int gridworld_add_agents(EnvHandle game, GroupHandle group, int n, const char *method,
                         const int *pos_x, const int *pos_y, const int *dir);
int gridworld_clear_dead(EnvHandle game);                              /* GridWorld.cc:633-665 */
int gridworld_set_goal(EnvHandle game, GroupHandle group, const char *method, const int *linear_buffer);
int gridworld_define_agent_symbol(EnvHandle game, int no, int group, int index);      /* RewardEngine.cc:28-35 */
int gridworld_define_event_node(EnvHandle game, int no, int op, int *inputs, int n_inputs); /* :37-49 */
int gridworld_add_reward_rule(EnvHandle game, int on, int *receiver, float *value, int n_receiver,
                              bool is_terminal, bool auto_value);                      /* :51-69 */

/* ---- the second game of the reference (src/runtime_api.h:60-61): out of scope, exported for link parity */
int discrete_snake_clear_dead(EnvHandle game);
int discrete_snake_add_object(EnvHandle game, int obj_id, int n, const char *method, const int *linear_buffer);

#ifdef __cplusplus
}
#endif
#endif /* MAGENT_B200_RUNTIME_API_H */
