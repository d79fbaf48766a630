/*
 * gridworld_oracle.c -- TEST INFRASTRUCTURE ONLY: a plain-C, single-threaded restatement of the
 * reference GridWorld step path, behind the same C ABI (src/runtime_api.h:20-61).
 *
 * What it is for: an independent CPU checker for the CUDA engine that travels to machines where
 * /root/reference does not exist.  It is PINNED against the unmodified reference: tests/test_oracle_cpu.py
 * replays every committed golden vector (tests/golden/ *.npz, recorded from oracle/_ref = the reference
 * compiled from its own sources by oracle/Makefile) through this file and requires bit-identical results.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; nothing under
 * magent_b200/ references it.
 *
 * It restates the SEQUENTIAL semantics (the reference is only deterministic with OMP_NUM_THREADS=1,
 * SURVEY.md §0): each function cites the reference lines it follows.  Not restated (no shipped config on
 * the hot path uses them; the functions abort with a message): OP_ALIGN, DiscreteSnake.
 */
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))
#define MAXG 16
#define MAXT 16

enum { OP_AND, OP_OR, OP_NOT, OP_KILL, OP_AT, OP_IN, OP_COLLIDE, OP_ATTACK, OP_DIE, OP_IN_A_LINE, OP_ALIGN, OP_NULL };
enum { CELL_EMPTY = -1, CELL_WALL = -2, CELL_FOOD = -3 };     /* food: what a killed agent leaves behind in food_mode (Map.cc:276-283) */

static void die(const char *msg, const char *arg) {
    fprintf(stderr, "[gridworld_oracle FATAL] %s%s\n", msg, arg ? arg : "");
    abort();
}

/* ---- ranges: Range.h:149-190 (CircleRange), same double arithmetic ---- */
typedef struct { int width, height, count, x1, y1, x2, y2; unsigned char *in; int *dx, *dy; } Range;

static Range circle_range(float radius, float inner, int parity) {
    const double eps = 1e-8;
    Range r;
    r.width = 2 * (int)(radius + eps) + parity;
    int center = (int)radius;
    if (r.width % 2 != parity) r.width++;
    r.height = r.width;
    r.in = calloc((size_t)r.width * r.width + 1, 1);
    r.dx = calloc((size_t)r.width * r.width + 1, sizeof(int));
    r.dy = calloc((size_t)r.width * r.width + 1, sizeof(int));
    r.count = 0;
    double delta = parity == 0 ? 0.5 : 0.0;
    for (int i = 0; i < r.width; i++)
        for (int j = 0; j < r.width; j++) {
            double ax = fabs(j - center + delta), ay = fabs(i - center + delta);
            double d = sqrt(ax * ax + ay * ay);
            if (d < radius + eps && d > inner - eps) {
                r.in[i * r.width + j] = 1;
                r.dx[r.count] = j - center; r.dy[r.count] = i - center; r.count++;
            }
        }
    r.x1 = r.y1 = -center;
    r.x2 = r.y2 = r.width - center - 1;
    return r;
}

/* Range.h:104-144 (SectorRange): a sector of `angle` degrees opening to the north, in a rectangle above the anchor */
static Range sector_range(float angle, float radius, int parity) {
    static const double PI_REF = 3.1415926536;    /* Range.h:16 */
    const double eps = 0.00001;
    Range r;
    r.height = (int)(radius + 0.5);
    r.width = (int)(2 * radius * sin(angle / 2 * (PI_REF / 180)) + 0.5);
    if (r.width % 2 != parity) r.width--;
    if (r.width < 0) r.width = 0;
    size_t n = (size_t)r.width * r.height;
    r.in = calloc(n + 1, 1);
    r.dx = calloc(n + 1, sizeof(int));
    r.dy = calloc(n + 1, sizeof(int));
    r.count = 0;
    for (int i = 0; i < r.height; i++)
        for (int j = 0; j < r.width; j++) {
            double ax = fabs(j - (r.width - 1) / 2.0), ay = fabs((double)(r.height - i));
            double d = sqrt(ax * ax + ay * ay);
            if (d < radius + 0.2 + eps && ax / ay < tan(angle / 2 * PI_REF / 180) + eps) {
                r.in[i * r.width + j] = 1;
                r.dx[r.count] = j - r.width / 2; r.dy[r.count] = i - r.height; r.count++;
            }
        }
    r.x1 = -r.width / 2; r.y1 = -r.height;
    r.x2 = (r.width - 1) / 2; r.y2 = -1;
    return r;
}

/* ---- agent types: AgentType.cc:30-123 ---- */
typedef struct {
    char name[64];
    int width, length;
    float speed, hp, view_radius, view_angle, attack_radius, attack_angle;
    float damage, step_recover, kill_supply, eat_ability, food_supply;
    int attack_in_group, can_absorb;
    float step_reward, kill_reward, dead_penalty, attack_penalty;
    Range view, attack, move;
    int turn_base, attack_base, n_action;
} Type;

/* ---- agents live in a pool; cells and groups refer to pool slots ---- */
typedef struct {
    int id, group, index, x, y, dir, action, last_op, op_obj, involved;   /* dir: EAST 0, SOUTH 1, WEST 2, NORTH 3 (grid_def.h:15) */
    bool dead, absorbed;
    float hp, next_reward, last_reward;
} Agent;

typedef struct { int type, n, cap, dead_ct; int *slot; float reward; } Group;
typedef struct { int agent, action; } Act;
enum { EAST = 0, SOUTH = 1, WEST = 2, NORTH = 3 };
typedef struct { Act *v; int n, cap; } ActBuf;

typedef struct { int group, index, entity; } Symbol;               /* RewardEngine.h:17-32 */
typedef struct { int op, nraw, raw[8]; int related[16], nrel; int isub[16], iobj[16], ninf; } Node;
typedef struct { int on, nrecv, recv[8]; float val[8]; bool terminal, trigger; int nin, in[16], inf[16]; } Rule;

typedef struct {
    int w, h, minimap_mode, goal_mode, turn_mode, food_mode, embedding, reset_done, rules_ready;
    uint32_t rng;                                                   /* minstd_rand0 state */
    int ntype; Type type[MAXT];
    int ngroup; Group grp[MAXG];
    Agent *pool; int npool, cappool;
    int *cell;
    float *food;                                                    /* amount per CELL_FOOD cell */
    int id_counter, nsep, large;
    ActBuf attack, move[17], turn[17];
    Symbol sym[32]; int nsym;
    Node node[32]; int nnode;
    Rule rule[16]; int nrule;
    /* replay dump: RenderGenerator.h, GridWorld.cc:18,97,484-509,797-842,939-948 */
    char render_dir[1024];
    int first_render_done, file_ct, frame_ct;                       /* first_render starts true (GridWorld.cc:18) */
    int *ev; int nev, capev;                                        /* attack events of the last step: id, x, y */
} Env;

static uint32_t rng_draw(Env *e) {                                  /* x <- 16807 x mod (2^31-1) */
    e->rng = (uint32_t)(((uint64_t)e->rng * 16807u) % 2147483647u);
    return e->rng;
}
static void push(ActBuf *b, int agent, int action) {
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 256; b->v = realloc(b->v, sizeof(Act) * b->cap); }
    b->v[b->n].agent = agent; b->v[b->n].action = action; b->n++;
}
static int g2c(const Env *e, int g) { return 1 + (e->food_mode ? 1 : 0) + g * (2 + (e->minimap_mode ? 1 : 0)); }      /* GridWorld.cc:915-924 */
static int feature_size(const Env *e, int g) {                                                /* :926-934 */
    return e->embedding + e->type[e->grp[g].type].n_action + 1 + (e->goal_mode ? 2 : 0) + (e->minimap_mode ? 2 : 0);
}

/* ---- map helpers: Map.cc:454-513 ---- */
static bool blank_area(const Env *e, int x, int y, int w, int h, int self) {
    if (x < 0 || y < 0 || x + w >= e->w || y + h >= e->h) return false;
    for (int i = 0; i < w; i++)
        for (int j = 0; j < h; j++) {
            int c = e->cell[(y + j) * e->w + x + i];
            if (c == CELL_WALL || c == CELL_FOOD || (c >= 0 && c != self)) return false;   /* food is an occupier too */
        }
    return true;
}
static void paint(Env *e, int x, int y, int w, int h, int v) {
    for (int i = 0; i < w; i++) for (int j = 0; j < h; j++) e->cell[(y + j) * e->w + x + i] = v;
}
static int first_other(const Env *e, int x, int y, int w, int h, int self) {                  /* get_collide */
    if (x < 0 || y < 0 || x + w >= e->w || y + h >= e->h) return -1;
    for (int i = 0; i < w; i++)
        for (int j = 0; j < h; j++) {
            int c = e->cell[(y + j) * e->w + x + i];
            if (c >= 0 && c != self) return c;
        }
    return -1;
}

/* ================================================================ ABI ============================ */
API int env_new_game(void **game, const char *name) {
    if (strcmp(name, "GridWorld") != 0) die("unsupported game ", name);
    Env *e = calloc(1, sizeof(Env));
    e->rng = 1;                                                     /* seed(0) -> state 1 (GridWorld.cc:29) */
    *game = e;
    return 0;
}
API int env_delete_game(void *game) { free(game); return 0; }      /* (leaks the arrays: test infrastructure) */

API int env_config_game(void *game, const char *key, void *p) {    /* GridWorld.cc:120-149 */
    Env *e = game;
    int iv = *(int *)p; bool bv = *(bool *)p;
    if (!strcmp(key, "map_width")) e->w = iv;
    else if (!strcmp(key, "map_height")) e->h = iv;
    else if (!strcmp(key, "minimap_mode")) e->minimap_mode = bv;
    else if (!strcmp(key, "goal_mode")) e->goal_mode = bv;
    else if (!strcmp(key, "turn_mode")) e->turn_mode = bv;
    else if (!strcmp(key, "embedding_size")) e->embedding = iv;
    else if (!strcmp(key, "render_dir")) { strncpy(e->render_dir, (const char *)p, sizeof e->render_dir - 1); }
    else if (!strcmp(key, "seed")) {
        uint32_t s = (uint32_t)(((unsigned long long)(long long)iv) % 2147483647ull);
        e->rng = s ? s : 1;
    } else if (!strcmp(key, "food_mode")) e->food_mode = bv;
    else die("invalid argument in set_config : ", key);
    return 0;
}

API int gridworld_register_agent_type(void *game, const char *name, int n, const char **keys, float *values) {
    Env *e = game;
    Type *t = &e->type[e->ntype++];
    memset(t, 0, sizeof *t);
    strncpy(t->name, name, 63);
    t->width = t->length = 1; t->speed = 1; t->hp = 1; t->view_radius = 1; t->view_angle = 360;
    for (int i = 0; i < n; i++) {                                   /* AgentType.cc:52-83 */
        const char *k = keys[i]; float v = values[i];
        if (!strcmp(k, "width")) t->width = (int)(v + 0.5);
        else if (!strcmp(k, "length")) t->length = (int)(v + 0.5);
        else if (!strcmp(k, "speed")) t->speed = v;
        else if (!strcmp(k, "hp")) t->hp = v;
        else if (!strcmp(k, "view_radius")) t->view_radius = v;
        else if (!strcmp(k, "view_angle")) t->view_angle = v;
        else if (!strcmp(k, "attack_radius")) t->attack_radius = v;
        else if (!strcmp(k, "attack_angle")) t->attack_angle = v;
        else if (!strcmp(k, "damage")) t->damage = v;
        else if (!strcmp(k, "step_recover")) t->step_recover = v;
        else if (!strcmp(k, "kill_supply")) t->kill_supply = v;
        else if (!strcmp(k, "eat_ability")) t->eat_ability = v;
        else if (!strcmp(k, "food_supply")) t->food_supply = v;
        else if (!strcmp(k, "attack_in_group")) t->attack_in_group = (int)(v + 0.5) != 0;
        else if (!strcmp(k, "step_reward")) t->step_reward = v;
        else if (!strcmp(k, "kill_reward")) t->kill_reward = v;
        else if (!strcmp(k, "dead_penalty")) t->dead_penalty = v;
        else if (!strcmp(k, "attack_penalty")) t->attack_penalty = v;
        else if (!strcmp(k, "can_absorb")) t->can_absorb = (int)(v + 0.5) != 0;
        else if (!strcmp(k, "hear_radius") || !strcmp(k, "speak_radius") || !strcmp(k, "speak_ability") ||
                 !strcmp(k, "trace") ||
                 !strcmp(k, "view_x_offset") || !strcmp(k, "view_y_offset") || !strcmp(k, "att_x_offset") ||
                 !strcmp(k, "att_y_offset") || !strcmp(k, "turn_x_offset") || !strcmp(k, "turn_y_offset")) {}
        else die("invalid agent config : ", k);
    }
    int parity = t->width % 2;                                      /* AgentType.cc:86-118 */
    if (t->view_angle < 180) t->view = sector_range(t->view_angle, t->view_radius, parity);
    else t->view = circle_range(t->view_radius, 0, parity);
    if (t->attack_angle >= 180) t->attack = circle_range(t->attack_radius, t->width / 2.0f, parity);
    else t->attack = sector_range(t->attack_angle, t->attack_radius, parity);
    t->move = circle_range(t->speed, 0, 1);
    t->turn_base = t->move.count;                                   /* AgentType.cc:110-118: [moves][turn L, R][attacks] */
    t->attack_base = t->turn_base + (e->turn_mode ? 2 : 0);
    t->n_action = t->attack_base + t->attack.count;
    return 0;
}

API int gridworld_new_group(void *game, const char *type_name, int *handle) {    /* GridWorld.cc:160-169 */
    Env *e = game;
    for (int i = 0; i < e->ntype; i++)
        if (!strcmp(e->type[i].name, type_name)) {
            *handle = e->ngroup;
            memset(&e->grp[e->ngroup], 0, sizeof(Group));
            e->grp[e->ngroup++].type = i;
            return 0;
        }
    die("invalid name of agent type in new_group : ", type_name);
    return 0;
}

API int gridworld_define_agent_symbol(void *game, int no, int group, int index) {            /* RewardEngine.cc:28-35 */
    Env *e = game;
    if (no >= e->nsym) e->nsym = no + 1;
    e->sym[no].group = group; e->sym[no].index = index; e->sym[no].entity = -1;
    return 0;
}
API int gridworld_define_event_node(void *game, int no, int op, int *inputs, int n) {        /* :37-49 */
    Env *e = game;
    if (no >= e->nnode) e->nnode = no + 1;
    e->node[no].op = op; e->node[no].nraw = n;
    for (int i = 0; i < n; i++) e->node[no].raw[i] = inputs[i];
    return 0;
}
API int gridworld_add_reward_rule(void *game, int on, int *recv, float *val, int n, bool terminal, bool autov) {
    Env *e = game; (void)autov;                                                               /* :51-69 */
    Rule *r = &e->rule[e->nrule++];
    memset(r, 0, sizeof *r);
    r->on = on; r->nrecv = n; r->terminal = terminal;
    for (int i = 0; i < n; i++) { r->recv[i] = recv[i]; r->val[i] = val[i]; }
    return 0;
}

/* related symbols / inference map per node, then the binding plan per rule: RewardEngine.cc:71-189.
 * std::set / std::map iteration order == symbol number order (pointers into one vector). */
static void add_rel(Node *n, int s) { for (int i = 0; i < n->nrel; i++) if (n->related[i] == s) return; n->related[n->nrel++] = s; }
static void add_inf(Node *n, int s, int o) { for (int i = 0; i < n->ninf; i++) if (n->isub[i] == s) return; n->isub[n->ninf] = s; n->iobj[n->ninf++] = o; }
static void collect(Env *e, int no) {
    Node *n = &e->node[no];
    if (n->nrel) return;
    switch (n->op) {
        case OP_AND: case OP_OR: case OP_NOT:
            for (int q = 0; q < (n->op == OP_NOT ? 1 : 2); q++) {
                collect(e, n->raw[q]);
                Node *c = &e->node[n->raw[q]];
                for (int i = 0; i < c->nrel; i++) add_rel(n, c->related[i]);
                for (int i = 0; i < c->ninf; i++) add_inf(n, c->isub[i], c->iobj[i]);
            }
            break;
        case OP_KILL: case OP_COLLIDE: case OP_ATTACK:
            add_rel(n, n->raw[0]); add_rel(n, n->raw[1]); add_inf(n, n->raw[0], n->raw[1]); break;
        case OP_AT: case OP_IN: case OP_DIE: case OP_IN_A_LINE: add_rel(n, n->raw[0]); break;
        default: die("event op not restated", NULL);
    }
}
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
static void plan_rules(Env *e) {
    for (int i = 0; i < e->nnode; i++) collect(e, i);
    for (int ri = 0; ri < e->nrule; ri++) {
        Rule *r = &e->rule[ri];
        Node on = e->node[r->on];
        qsort(on.related, on.nrel, sizeof(int), cmp_int);
        bool added[32] = {0};
        r->nin = 0;
        for (int i = 0; i < on.nrel; i++) {
            int s = on.related[i];
            if (added[s]) continue;
            for (int q = 0; q < on.ninf; q++)
                if (on.isub[q] == s) { r->in[r->nin] = s; r->inf[r->nin++] = on.iobj[q]; added[s] = added[on.iobj[q]] = true; break; }
        }
        for (int i = 0; i < on.nrel; i++)
            if (!added[on.related[i]]) { r->in[r->nin] = on.related[i]; r->inf[r->nin++] = -1; }
    }
}

API int env_reset(void *game) {                                     /* GridWorld.cc:72-118, Map.cc:23-47 */
    Env *e = game;
    e->id_counter = 0;
    e->file_ct++; e->frame_ct = 0;                                  /* RenderGenerator::next_file (GridWorld.cc:97) */
    e->large = e->w * e->h > 99 * 99;
    e->nsep = e->large ? (e->w * e->h > 1000 * 1000 ? 16 : 8) : 1;
    free(e->cell);
    e->cell = malloc(sizeof(int) * e->w * e->h);
    for (int i = 0; i < e->w * e->h; i++) e->cell[i] = CELL_EMPTY;
    free(e->food);
    e->food = calloc((size_t)e->w * e->h, sizeof(float));
    for (int i = 0; i < e->w; i++) { e->cell[i] = CELL_WALL; e->cell[(e->h - 1) * e->w + i] = CELL_WALL; }
    for (int i = 0; i < e->h; i++) { e->cell[i * e->w] = CELL_WALL; e->cell[i * e->w + e->w - 1] = CELL_WALL; }
    e->npool = 0;
    for (int g = 0; g < e->ngroup; g++) { e->grp[g].n = 0; e->grp[g].dead_ct = 0; }
    e->attack.n = 0;
    for (int b = 0; b <= 16; b++) e->move[b].n = e->turn[b].n = 0;
    if (!e->rules_ready) { plan_rules(e); e->rules_ready = 1; }
    e->reset_done = 1;
    return 0;
}

static void random_blank(Env *e, int w, int h, int *px, int *py) {  /* Map.cc:49-63 */
    for (int tries = 0;; ) {
        int x = (int)rng_draw(e) % (e->w - w);
        int y = (int)rng_draw(e) % (e->h - h);
        if (blank_area(e, x, y, w, h, -1)) { *px = x; *py = y; return; }
        if (tries++ > e->w * e->h) die("cannot find a blank position in a filled map", NULL);
    }
}
static void add_wall(Env *e, int x, int y) {                         /* Map.cc:108-115 */
    if (x < 0 || y < 0 || x >= e->w || y >= e->h) return;
    if (e->cell[y * e->w + x] >= 0 || e->cell[y * e->w + x] == CELL_FOOD) return;   /* BLANK slot with an occupier: refused */
    e->cell[y * e->w + x] = CELL_WALL;
}
/* ---- headings (turn_mode): Map.cc:515-607 ---- */
static void size_for_dir(const Type *t, int dir, int *w, int *h) {   /* get_size_for_dir */
    if (dir == NORTH || dir == SOUTH) { *w = t->width; *h = t->length; } else { *w = t->length; *h = t->width; }
}
static void rela_to_abs(int cx, int cy, int dir, int rx, int ry, int *ax, int *ay) {
    switch (dir) {
        case NORTH: *ax = cx + rx; *ay = cy + ry; break;
        case SOUTH: *ax = cx - rx; *ay = cy - ry; break;
        case WEST: *ax = cx + ry; *ay = cy - rx; break;
        default: *ax = cx - ry; *ay = cy + rx; break;               /* EAST */
    }
}
static void save_to_real(const Type *t, const Agent *a, int *rx, int *ry) {
    switch (a->dir) {
        case NORTH: *rx = a->x; *ry = a->y; break;
        case SOUTH: *rx = a->x + t->width - 1; *ry = a->y + t->length - 1; break;
        case WEST: *rx = a->x; *ry = a->y + t->width - 1; break;
        default: *rx = a->x + t->length - 1; *ry = a->y; break;     /* EAST */
    }
}
static void real_to_save(const Type *t, int rx, int ry, int dir, int *sx, int *sy) {
    switch (dir) {
        case NORTH: *sx = rx; *sy = ry; break;
        case SOUTH: *sx = rx - t->width + 1; *sy = ry - t->length + 1; break;
        case WEST: *sx = rx; *sy = ry - t->width + 1; break;
        default: *sx = rx - t->length + 1; *sy = ry; break;         /* EAST */
    }
}

static void add_agent(Env *e, int g, int x, int y, int dir) {        /* Map.cc:75-97 + Agent ctor GridWorld.h:133-144 */
    Group *G = &e->grp[g];
    const Type *t = &e->type[G->type];
    int bw, bh;
    size_for_dir(t, dir, &bw, &bh);
    if (!blank_area(e, x, y, bw, bh, -1)) return;                    /* occupied: silently ignored, no id used */
    if (e->npool == e->cappool) { e->cappool = e->cappool ? 2 * e->cappool : 1024; e->pool = realloc(e->pool, sizeof(Agent) * e->cappool); }
    int s = e->npool++;
    Agent *a = &e->pool[s];
    memset(a, 0, sizeof *a);
    a->id = e->id_counter++; a->group = g; a->x = x; a->y = y; a->dir = dir;
    a->hp = t->hp; a->action = t->n_action; a->last_op = OP_NULL; a->op_obj = -1;
    a->last_reward = 0; a->next_reward = t->step_reward;
    if (G->n == G->cap) { G->cap = G->cap ? 2 * G->cap : 256; G->slot = realloc(G->slot, sizeof(int) * G->cap); }
    a->index = 0;                          /* GridWorld.h:136: index(0); only clear_dead refreshes it (GridWorld.cc:655) */
    G->slot[G->n++] = s;
    paint(e, x, y, bw, bh, s);
}

API int gridworld_add_agents(void *game, int group, int n, const char *method,
                             const int *px, const int *py, const int *pdir) {                /* GridWorld.cc:180-290 */
    Env *e = game;
    int rnd = !strcmp(method, "random"), cus = !strcmp(method, "custom"), fil = !strcmp(method, "fill");
    if (!rnd && !cus && !fil) die("unsupported method in add_agents : ", method);
    if (group == -1) {
        if (rnd) for (int i = 0; i < n; i++) { int x, y; random_blank(e, 1, 1, &x, &y); add_wall(e, x, y); }
        else if (cus) for (int i = 0; i < n; i++) add_wall(e, px[i], py[i]);
        else for (int x = px[0]; x < px[0] + px[2]; x++) for (int y = px[1]; y < px[1] + px[3]; y++) add_wall(e, x, y);
        return 0;
    }
    const Type *t = &e->type[e->grp[group].type];
    if (rnd) for (int i = 0; i < n; i++) {                           /* :228-243: heading first, then the rotated footprint */
        int dir = e->turn_mode ? (int)(rng_draw(e) % 4u) : NORTH, bw, bh, x, y;
        size_for_dir(t, dir, &bw, &bh);
        random_blank(e, bw, bh, &x, &y);
        add_agent(e, group, x, y, dir);
    }
    else if (cus) for (int i = 0; i < n; i++) {
        if (pdir && pdir[i] >= 4) die("invalid direction", NULL);
        add_agent(e, group, px[i], py[i], e->turn_mode && pdir ? pdir[i] : NORTH);
    }
    else {                                                           /* :264-287 */
        int dir = e->turn_mode ? px[4] : NORTH, bw, bh;
        if (dir < 0 || dir >= 4) die("invalid direction", NULL);
        size_for_dir(t, dir, &bw, &bh);
        for (int x = px[0]; x < px[0] + px[2]; x += bw) for (int y = px[1]; y < px[1] + px[3]; y += bh) add_agent(e, group, x, y, dir);
    }
    return 0;
}

/* channel of group `other` in the observation of group `me`: make_channel_trans, GridWorld.cc:897-913 */
static int chan(const Env *e, int me, int other) {
    int rel = other - me; if (rel < 0) rel += e->ngroup;
    return g2c(e, 0) + rel * (2 + (e->minimap_mode ? 1 : 0));
}

API int env_get_observation(void *game, int g, float **bufs) {      /* GridWorld.cc:292-401 + Map.cc:129-207 */
    Env *e = game;
    Group *G = &e->grp[g];
    const Type *t = &e->type[G->type];
    const int C = g2c(e, e->ngroup), vw = t->view.width, vh = t->view.height, F = feature_size(e, g);
    float *view = bufs[0], *feat = bufs[1];
    memset(view, 0, sizeof(float) * (size_t)G->n * vh * vw * C);
    memset(feat, 0, sizeof(float) * (size_t)G->n * F);
    float *mini = NULL;
    int scale_h = (e->h + vh - 1) / vh, scale_w = (e->w + vw - 1) / vw;
    if (e->minimap_mode) {                                           /* :331-360 */
        mini = calloc((size_t)vh * vw * e->ngroup, sizeof(float));
        for (int j = 0; j < e->ngroup; j++) {
            size_t total = 0;
            for (int k = 0; k < e->grp[j].n; k++) {
                const Agent *b = &e->pool[e->grp[j].slot[k]];
                if (t->can_absorb && b->absorbed) continue;          /* :346 (the observer's type decides) */
                mini[((b->y / scale_h) * vw + b->x / scale_w) * e->ngroup + j]++;
                total++;
            }
            for (int k = 0; k < vh * vw; k++) mini[k * e->ngroup + j] /= total;
        }
    }
    for (int i = 0; i < G->n; i++) {
        const Agent *a = &e->pool[G->slot[i]];
        float *out = view + (size_t)i * vh * vw * C;
        int real_x, real_y, eye_x, eye_y;                           /* Map.cc:138-146: the window lives in the agent's frame */
        save_to_real(t, a, &real_x, &real_y);
        rela_to_abs(real_x, real_y, a->dir, t->width / 2, t->length / 2, &eye_x, &eye_y);
        for (int vy = 0; vy < vh; vy++)
            for (int vx = 0; vx < vw; vx++) {
                int x, y;
                rela_to_abs(eye_x, eye_y, a->dir, t->view.x1 + vx, t->view.y1 + vy, &x, &y);
                if (x < 0 || x >= e->w || y < 0 || y >= e->h) continue;          /* window clipped to the map */
                int c = e->cell[y * e->w + x];
                if (c == CELL_EMPTY || !t->view.in[vy * vw + vx]) continue;
                float *px = out + (size_t)(vy * vw + vx) * C;
                if (c == CELL_WALL) { px[0] = 1; continue; }
                if (c == CELL_FOOD) { px[1] = 1; continue; }        /* food channel, no hp (Map.cc:191-199) */
                const Agent *b = &e->pool[c];
                int ch = chan(e, g, b->group);
                px[ch] = 1;
                px[ch + 1] = b->hp / e->type[e->grp[b->group].type].hp;            /* Map.cc:197 */
            }
        if (mini) {                                                                  /* :371-384 */
            int sx = a->x / scale_w, sy = a->y / scale_h;
            for (int j = 0; j < e->ngroup; j++) {
                int ch = chan(e, g, j) + 2;
                for (int k = 0; k < vh * vw; k++) out[(size_t)k * C + ch] = mini[k * e->ngroup + j];
                out[(size_t)(sy * vw + sx) * C + ch] += 1;
            }
        }
        float *f = feat + (size_t)i * F;                                             /* :386-396 */
        int id = a->id;
        for (int k = 0; k < e->embedding; k++, id >>= 1) f[k] = (float)(id & 1);
        f[e->embedding + a->action] = 1;
        f[e->embedding + t->n_action] = a->last_reward;
        if (e->minimap_mode) {
            f[e->embedding + t->n_action + 1] = (float)a->x / e->w;
            f[e->embedding + t->n_action + 2] = (float)a->y / e->h;
        }
    }
    free(mini);
    return 0;
}

API int env_set_action(void *game, int g, const int *actions) {     /* GridWorld.cc:403-454 */
    Env *e = game;
    Group *G = &e->grp[g];
    const Type *t = &e->type[G->type];
    int bw = (e->w + e->nsep - 1) / e->nsep;
    for (int i = 0; i < G->n; i++) {
        int s = G->slot[i], act = actions[i];
        e->pool[s].action = act;
        if (act < t->attack_base) {
            int b = 16;                                              /* boundary buffer */
            if (e->large) { int xm = e->pool[s].x % bw; if (!(xm < 4 || xm > bw - 4)) b = e->pool[s].x / bw; }
            push(act < t->turn_base ? &e->move[b] : &e->turn[b], s, act);     /* the turn action keeps its raw number (:430-433) */
        } else push(&e->attack, s, act - t->attack_base);
    }
    return 0;
}

static void kill_agent(Env *e, Agent *v) {                          /* Agent::be_attack death branch + Map::remove_agent */
    const Type *tv = &e->type[e->grp[v->group].type];
    v->dead = true;
    v->next_reward = tv->dead_penalty;
    int bw, bh;
    size_for_dir(tv, v->dir, &bw, &bh);
    paint(e, v->x, v->y, bw, bh, CELL_EMPTY);
    e->grp[v->group].dead_ct++;
}

/* reward rules: RewardEngine.cc:216-443 */
static bool eval_node(Env *e, int no) {
    Node *n = &e->node[no];
    switch (n->op) {
        case OP_AND: return eval_node(e, n->raw[0]) && eval_node(e, n->raw[1]);
        case OP_OR: return eval_node(e, n->raw[0]) || eval_node(e, n->raw[1]);
        case OP_NOT: return !eval_node(e, n->raw[0]);
        case OP_KILL: case OP_ATTACK: case OP_COLLIDE: {
            Symbol *s = &e->sym[n->raw[0]], *o = &e->sym[n->raw[1]];
            if (s->index == -2) {
                Group *G = &e->grp[s->group];
                for (int i = 0; i < G->n; i++) { Agent *a = &e->pool[G->slot[i]]; if (!(a->last_op == n->op && a->op_obj == o->entity)) return false; }
                return true;
            }
            Agent *a = &e->pool[s->entity];
            return a->last_op == n->op && a->op_obj == o->entity;
        }
        case OP_AT: case OP_IN: case OP_DIE: {
            Symbol *s = &e->sym[n->raw[0]];
            int lo = 0, hi = 1; Group *G = NULL;
            if (s->index == -2) { G = &e->grp[s->group]; hi = G->n; }
            for (int i = lo; i < hi; i++) {
                Agent *a = &e->pool[G ? G->slot[i] : s->entity];
                bool v = n->op == OP_DIE ? a->dead
                       : n->op == OP_AT ? (a->x == n->raw[1] && a->y == n->raw[2])
                       : (a->x > n->raw[1] && a->x < n->raw[3] && a->y > n->raw[2] && a->y < n->raw[4]);
                if (!v) return false;
            }
            return true;
        }
        case OP_IN_A_LINE: {                                           /* RewardEngine.cc:262-292 */
            Symbol *s = &e->sym[n->raw[0]];
            if (s->index != -2) die("in_a_line needs an 'all' subject (the reference asserts)", NULL);
            Group *G = &e->grp[s->group];
            if (G->n < 2) return true;
            Agent *a0 = &e->pool[G->slot[0]], *a1 = &e->pool[G->slot[1]];
            int dx = a0->x - a1->x, dy = a0->y - a1->y;
            bool vertical;
            if (dx == 0 && dy != 0) vertical = true;
            else if (dx != 0 && dy == 0) vertical = false;
            else return false;
            int base = vertical ? a0->x : a0->y, mn = vertical ? a0->y : a0->x, mx = mn;
            for (int i = 1; i < G->n; i++) {
                Agent *a = &e->pool[G->slot[i]];
                int var = vertical ? a->y : a->x, fix = vertical ? a->x : a->y;
                if (var < mn) mn = var;
                if (var > mx) mx = var;
                if (fix != base) return false;
            }
            return mx - mn + 1 == G->n;
        }
        /* OP_ALIGN reads counter_x / counter_y, which the reference never allocates (GridWorld.cc:31): a null
           dereference there, refused here */
        default: die("event op not restated", NULL);
    }
    return false;
}
static bool bind_check(Env *e, int sym, int entity) {               /* AgentSymbol::bind_with_check :14-23 */
    Symbol *s = &e->sym[sym];
    Agent *a = &e->pool[entity];
    if (s->group != a->group) return false;
    if (s->index != -1 && s->index != a->index) return false;
    s->entity = entity;
    return true;
}
static void calc_rule(Env *e, Rule *r, int now) {                   /* :373-443 */
    if (now == r->nin) {
        if (eval_node(e, r->on)) {
            r->trigger = true;
            for (int i = 0; i < r->nrecv; i++) {
                Symbol *s = &e->sym[r->recv[i]];
                if (s->index == -2) e->grp[s->group].reward += r->val[i];
                else e->pool[s->entity].next_reward += r->val[i];
            }
        }
        return;
    }
    Symbol *s = &e->sym[r->in[now]];
    Group *G = &e->grp[s->group];
    if (s->index == -1) {
        for (int i = 0; i < G->n; i++) {
            int slot = G->slot[i];
            s->entity = slot;
            if (e->pool[slot].involved) continue;
            e->pool[slot].involved = 1;
            if (r->inf[now] >= 0) {
                int obj = e->pool[slot].op_obj;
                if (obj >= 0 && bind_check(e, r->inf[now], obj)) calc_rule(e, r, now + 1);
            } else calc_rule(e, r, now + 1);
            e->pool[slot].involved = 0;
        }
    } else if (s->index == -2) {
        if (r->inf[now] >= 0) {
            if (G->n > 0) { int obj = e->pool[G->slot[0]].op_obj; if (obj >= 0 && bind_check(e, r->inf[now], obj)) calc_rule(e, r, now + 1); }
        } else calc_rule(e, r, now + 1);
    } else if (s->index < G->n) {
        int slot = G->slot[s->index];
        s->entity = slot;
        if (r->inf[now] >= 0 && e->pool[slot].op_obj >= 0 && bind_check(e, r->inf[now], e->pool[slot].op_obj)) calc_rule(e, r, now + 1);
    }
}

API int env_step(void *game, int *done) {                           /* GridWorld.cc:456-631 */
    Env *e = game;
    /* shuffle: :464-468 */
    for (int i = 0; i < e->attack.n; i++) {
        int j = (int)rng_draw(e) % (i + 1);
        Act t = e->attack.v[i]; e->attack.v[i] = e->attack.v[j]; e->attack.v[j] = t;
    }
    /* attack: :475-506, Map::get_attack_obj Map.cc:209-252, Map::do_attack :255-310 */
    const int record = e->first_render_done;                        /* `if (!first_render)`, :484,508: else the old events stay */
    if (record) e->nev = 0;
    for (int i = 0; i < e->attack.n; i++) {
        Agent *a = &e->pool[e->attack.v[i].agent];
        if (a->dead) continue;
        const Type *t = &e->type[e->grp[a->group].type];
        int k = e->attack.v[i].action;
        int rx, ry, tx, ty;
        save_to_real(t, a, &rx, &ry);
        rela_to_abs(rx, ry, a->dir, t->width / 2 + t->attack.dx[k], t->length / 2 + t->attack.dy[k], &tx, &ty);
        if (record) {                                               /* RenderAttackEvent{id, obj_x, obj_y}, on or off the board */
            if (e->nev == e->capev) { e->capev = e->capev ? 2 * e->capev : 256; e->ev = realloc(e->ev, sizeof(int) * 3 * e->capev); }
            e->ev[3 * e->nev] = a->id; e->ev[3 * e->nev + 1] = tx; e->ev[3 * e->nev + 2] = ty; e->nev++;
        }
        int c = (tx >= 0 && tx < e->w && ty >= 0 && ty < e->h) ? e->cell[ty * e->w + tx] : CELL_EMPTY;
        if (c == CELL_FOOD) {                                        /* Map.cc:292-303: eat; any group may (get_attack_obj :245) */
            float *food = &e->food[ty * e->w + tx];
            float add = t->eat_ability < *food ? t->eat_ability : *food;
            float nh = a->hp + add; a->hp = nh < t->hp ? nh : t->hp;
            *food -= add;
            if (*food < 0.1) { e->cell[ty * e->w + tx] = CELL_EMPTY; *food = 0; }
            a->next_reward += 0.0f + t->attack_penalty;
            continue;
        }
        if (c < 0 || (!t->attack_in_group && e->pool[c].group == a->group)) { a->next_reward += t->attack_penalty; continue; }
        Agent *v = &e->pool[c];
        const Type *tv = &e->type[e->grp[v->group].type];
        float reward = 0.0f;
        v->hp -= t->damage;
        if (v->hp < 0.0) {
            kill_agent(e, v);
            a->last_op = OP_KILL; a->op_obj = c;
            float nh = a->hp + tv->kill_supply; a->hp = nh < t->hp ? nh : t->hp;     /* std::min(type.hp, hp+add) */
            if (e->food_mode) { e->cell[ty * e->w + tx] = CELL_FOOD; e->food[ty * e->w + tx] = tv->food_supply; }   /* the attacked cell only */
            reward = tv->kill_reward;
        } else { a->last_op = OP_ATTACK; a->op_obj = c; }
        a->next_reward += reward + t->attack_penalty;
    }
    e->attack.n = 0;
    /* starve: :519-542, Agent::starve GridWorld.h:194-201 */
    for (int g = 0; g < e->ngroup; g++) {
        const Type *t = &e->type[e->grp[g].type];
        for (int i = 0; i < e->grp[g].n; i++) {
            Agent *a = &e->pool[e->grp[g].slot[i]];
            if (a->dead) continue;
            if (t->step_recover > 0) { float nh = a->hp + t->step_recover; a->hp = nh < t->hp ? nh : t->hp; }
            else { a->hp -= -t->step_recover; if (a->hp < 0.0) kill_agent(e, a); }
        }
    }
    /* turn: :544-571, Map::do_turn Map.cc:361-406.  The caller passes the raw action number, so wise = 2 * act - 1 is
       never -1: the "else" rotation formula always applies and new_dir = (dir + wise + 4) % 4.  With turn offsets 0
       the pivot is the agent's real corner, which therefore stays put. */
    for (int bi = 0; bi <= 16 && e->turn_mode; bi++) {
        for (int i = 0; i < e->turn[bi].n; i++) {
            int s = e->turn[bi].v[i].agent;
            Agent *a = &e->pool[s];
            if (a->dead) continue;
            const Type *t = &e->type[e->grp[a->group].type];
            int wise = e->turn[bi].v[i].action * 2 - 1;
            int new_dir = (a->dir + wise + 4) % 4, bw, bh, rx, ry, sx, sy;
            size_for_dir(t, a->dir, &bw, &bh);
            save_to_real(t, a, &rx, &ry);
            real_to_save(t, rx, ry, new_dir, &sx, &sy);
            if (blank_area(e, sx, sy, bh, bw, s)) {
                paint(e, a->x, a->y, bw, bh, CELL_EMPTY);
                a->dir = new_dir;
                paint(e, sx, sy, bh, bw, s);
                a->x = sx; a->y = sy;
            }
        }
        e->turn[bi].n = 0;
    }
    /* move: :573-613, Map::do_move Map.cc:313-358; band buffers in order, then the boundary buffer */
    for (int bi = 0; bi <= 16; bi++) {
        int b = bi < 16 ? bi : 16;
        for (int i = 0; i < e->move[b].n; i++) {
            int s = e->move[b].v[i].agent;
            Agent *a = &e->pool[s];
            if (a->dead || a->absorbed) continue;
            const Type *t = &e->type[e->grp[a->group].type];
            int nx, ny, bw, bh;
            rela_to_abs(a->x, a->y, a->dir, t->move.dx[e->move[b].v[i].action], t->move.dy[e->move[b].v[i].action], &nx, &ny);   /* :587-598 */
            size_for_dir(t, a->dir, &bw, &bh);
            if (blank_area(e, nx, ny, bw, bh, s)) {
                paint(e, a->x, a->y, bw, bh, CELL_EMPTY);
                paint(e, nx, ny, bw, bh, s);
                a->x = nx; a->y = ny;
            } else {
                int o = first_other(e, nx, ny, bw, bh, s);
                if (o >= 0 && e->type[e->grp[e->pool[o].group].type].can_absorb) {   /* Map.cc:341-349 */
                    Agent *obj = &e->pool[o];
                    if (!obj->absorbed) {
                        obj->absorbed = true;
                        obj->hp = obj->hp * 2;
                        a->dead = true;                               /* no dead_penalty, dead_ct untouched */
                        paint(e, a->x, a->y, bw, bh, CELL_EMPTY);
                        a->last_op = OP_COLLIDE; a->op_obj = o;
                    }
                } else if (o >= 0) { a->last_op = OP_COLLIDE; a->op_obj = o; }
            }
        }
        e->move[b].n = 0;
    }
    /* rewards: :681-692 */
    for (int r = 0; r < e->nrule; r++) { e->rule[r].trigger = false; calc_rule(e, &e->rule[r], 0); }
    /* done: :618-630 */
    int live = 0;
    for (int g = 0; g < e->ngroup; g++) if (e->grp[g].n - e->grp[g].dead_ct > 0) live++;
    *done = live < e->ngroup;
    for (int r = 0; r < e->nrule; r++) if (e->rule[r].trigger && e->rule[r].terminal) *done = 1;
    return 0;
}

API int env_get_reward(void *game, int g, float *buf) {             /* GridWorld.cc:694-704 */
    Env *e = game;
    for (int i = 0; i < e->grp[g].n; i++) buf[i] = e->pool[e->grp[g].slot[i]].next_reward + e->grp[g].reward;
    return 0;
}

API int gridworld_clear_dead(void *game) {                          /* GridWorld.cc:633-665; pool slots of the dead are recycled lazily */
    Env *e = game;
    for (int g = 0; g < e->ngroup; g++) {
        Group *G = &e->grp[g];
        const Type *t = &e->type[G->type];
        int pt = 0;
        G->reward = 0;
        for (int i = 0; i < G->n; i++) {
            Agent *a = &e->pool[G->slot[i]];
            if (a->dead) continue;
            a->last_reward = a->next_reward; a->last_op = OP_NULL; a->next_reward = t->step_reward; a->op_obj = -1; a->involved = 0;
            a->index = pt;
            G->slot[pt++] = G->slot[i];
        }
        G->n = pt; G->dead_ct = 0;
    }
    return 0;
}

API int env_get_info(void *game, int g, const char *name, void *buf) {          /* GridWorld.cc:709-894 */
    Env *e = game;
    int *ib = buf; bool *bb = buf;
    if (!strcmp(name, "num")) ib[0] = e->grp[g].n;
    else if (!strcmp(name, "id")) for (int i = 0; i < e->grp[g].n; i++) ib[i] = e->pool[e->grp[g].slot[i]].id;
    else if (!strcmp(name, "pos")) for (int i = 0; i < e->grp[g].n; i++) { ib[2 * i] = e->pool[e->grp[g].slot[i]].x; ib[2 * i + 1] = e->pool[e->grp[g].slot[i]].y; }
    else if (!strcmp(name, "alive")) for (int i = 0; i < e->grp[g].n; i++) bb[i] = !e->pool[e->grp[g].slot[i]].dead;
    else if (!strcmp(name, "action_space")) ib[0] = e->type[e->grp[g].type].n_action;
    else if (!strcmp(name, "view_space")) { const Type *t = &e->type[e->grp[g].type]; ib[0] = t->view.height; ib[1] = t->view.width; ib[2] = g2c(e, e->ngroup); }
    else if (!strcmp(name, "feature_space")) ib[0] = feature_size(e, g);
    else if (!strcmp(name, "attack_base")) ib[0] = e->type[e->grp[g].type].attack_base;
    else if (!strcmp(name, "view2attack")) {
        const Type *t = &e->type[e->grp[g].type];
        for (int i = 0; i < t->view.width * t->view.height; i++) ib[i] = -1;
        for (int i = 0; i < t->attack.count; i++) {                 /* linear index, no bounds check (NDPointer::at): in-buffer */
            long idx = (long)(t->attack.dy[i] - t->view.y1) * t->view.width + (t->attack.dx[i] - t->view.x1);   /* indices wrap to the */
            if (idx >= 0 && idx < (long)t->view.height * t->view.width) ib[idx] = i;                      /* neighbouring row; the rest is dropped */
        }
    } else if (!strcmp(name, "render_window_info")) {               /* GridWorld.cc:797-834 */
        e->first_render_done = 1;
        int x1 = ib[0], y1 = ib[1], x2 = ib[2], y2 = ib[3], ct = 1;
        for (int i = 0; i < e->ngroup; i++) {
            const Type *t = &e->type[e->grp[i].type];
            for (int j = 0; j < e->grp[i].n; j++) {
                const Agent *a = &e->pool[e->grp[i].slot[j]];
                if (a->x < x1 || a->x > x2 || a->y < y1 || a->y > y2) continue;
                if (t->can_absorb && !a->absorbed) continue;
                ib[ct * 4] = a->id; ib[ct * 4 + 1] = a->x; ib[ct * 4 + 2] = a->y; ib[ct * 4 + 3] = i; ct++;
            }
        }
        ib[0] = ct - 1; ib[1] = e->nev;
    } else if (!strcmp(name, "attack_event")) {                     /* GridWorld.cc:835-842 */
        for (int i = 0; i < 3 * e->nev; i++) ib[i] = e->ev[i];
    } else if (!strcmp(name, "groups_info")) {                      /* GridWorld.cc:872-887 (4 colours: callers with <= 4 groups) */
        static const int colors[4][3] = {{192, 64, 64}, {64, 64, 192}, {64, 192, 64}, {64, 64, 64}};
        for (int i = 0; i < e->ngroup; i++) {
            ib[i * 5] = e->type[e->grp[i].type].width; ib[i * 5 + 1] = e->type[e->grp[i].type].length;
            for (int c = 0; c < 3; c++) ib[i * 5 + 2 + c] = colors[i % 4][c];
        }
    } else if (!strcmp(name, "walls_info")) {                       /* GridWorld.cc:785-795, Map::get_wall Map.cc:609-615 */
        int ct = 0;
        for (int i = 0; i < e->w * e->h; i++)
            if (e->cell[i] == CELL_WALL) { ct++; ib[ct * 2] = i % e->w; ib[ct * 2 + 1] = i / e->w; }
        ib[0] = ct;
    } else if (!strcmp(name, "global_minimap")) {                   /* GridWorld.cc:738-762: dead-but-unculled agents count too */
        float *fb = buf;
        int vh = (int)lroundf(fb[0]), vw = (int)lroundf(fb[1]), ng = e->ngroup;
        for (int i = 0; i < vh * vw * ng; i++) fb[i] = 0.0f;
        int sh = (e->h + vh - 1) / vh, sw = (e->w + vw - 1) / vw;
        for (int i = 0; i < ng; i++) {
            int ch = ((i - g) % ng + ng) % ng;
            for (int j = 0; j < e->grp[i].n; j++) {
                const Agent *a = &e->pool[e->grp[i].slot[j]];
                fb[((a->y / sh) * vw + a->x / sw) * ng + ch] += 1.0f;
            }
            for (int j = 0; j < vh * vw; j++) fb[j * ng + ch] /= (float)e->grp[i].n;
        }
    } else if (!strcmp(name, "mean_info")) {                        /* GridWorld.cc:763-784 (every agent must hold a valid action) */
        float *fb = buf;
        int n = e->grp[g].n, na = e->type[e->grp[g].type].n_action;
        float sx = 0, sy = 0;
        int ctr[256] = {0};
        for (int i = 0; i < n; i++) {
            const Agent *a = &e->pool[e->grp[g].slot[i]];
            sx += a->x; sy += a->y;
            if (a->action >= 0 && a->action < na && a->action < 256) ctr[a->action]++;
        }
        fb[0] = sx / n; fb[1] = sy / n;
        for (int i = 0; i < na && i < 256; i++) fb[2 + i] = (float)(1.0 * ctr[i] / n);
    } else die("info name not restated : ", name);
    return 0;
}

/* ---- replay dump: GridWorld::render (GridWorld.cc:939-948), RenderGenerator::gen_config / render_a_frame
 * (RenderGenerator.cc:56-185).  ostream << float prints like "%g". ---- */
static void rgba(FILE *f, const char *key, const int *c, const char *alpha, int last) {
    fprintf(f, "\"%s\": \"rgba(%d,%d,%d,%s)\"%s\n", key, c[0], c[1], c[2], alpha, last ? "" : ",");
}
static void gen_config(Env *e) {
    static const int colors[4][3] = {{192, 64, 64}, {64, 64, 192}, {64, 192, 64}, {64, 64, 64}};
    static const int grey[3] = {127, 127, 127}, dark[3] = {63, 63, 63};
    char path[1200];
    snprintf(path, sizeof path, "%s/config.json", e->render_dir);
    FILE *f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "{\n\"width\": %d,\n\"height\": %d,\n\"static-file\": \"static.map\",\n", e->w, e->h);
    rgba(f, "obstacle-style", grey, "1", 0);
    fprintf(f, "\"dynamic-file-directory\": \".\",\n");
    rgba(f, "attack-style", dark, "0.8", 0);
    fprintf(f, "\"minimap-width\": 300,\n\"minimap-height\": 250,\n\"group\" : [\n");
    for (int i = 0; i < e->ngroup; i++) {
        const Type *t = &e->type[e->grp[i].type];
        const int *c = colors[i % 4];                               /* (the reference reads past its 4 rows for i >= 4) */
        fprintf(f, "{\n\"height\": %d,\n\"width\": %d,\n", t->length, t->width);
        rgba(f, "style", c, "1", 0);
        fprintf(f, "\"anchor\": [0, 0],\n\"max-speed\": %d,\n", (int)t->speed);
        rgba(f, "speed-style", c, "0.01", 0);
        fprintf(f, "\"vision-radius\": %g,\n\"vision-angle\": %g,\n", t->view_radius, t->view_angle);
        rgba(f, "vision-style", c, "0.2", 0);
        fprintf(f, "\"attack-radius\": %g,\n\"attack-angle\": %g,\n", t->attack_radius, t->attack_angle);
        rgba(f, "attack-style", c, "0.1", 0);
        fprintf(f, "\"broadcast-radius\": 1\n%s\n", i == e->ngroup - 1 ? "}" : "},");
    }
    fprintf(f, "]\n}\n");
    fclose(f);
}

API int env_render(void *game) {
    Env *e = game;
    if (!e->first_render_done) { e->first_render_done = 1; if (e->render_dir[0]) gen_config(e); }
    if (!e->render_dir[0]) return 0;                                /* RenderGenerator.cc:109-111 */
    char path[1200];
    snprintf(path, sizeof path, "%s/video_%d.txt", e->render_dir, e->file_ct);
    FILE *f = fopen(path, e->frame_ct == 0 ? "w" : "a");
    if (!f) return 0;
    if (e->frame_ct == 0) {
        int nw = 0;
        for (int i = 0; i < e->w * e->h; i++) nw += e->cell[i] == CELL_WALL;
        fprintf(f, "W %d\n", nw);
        for (int i = 0; i < e->w * e->h; i++) if (e->cell[i] == CELL_WALL) fprintf(f, "%d %d\n", i % e->w, i / e->w);
    }
    int n_agents = 0;
    for (int g = 0; g < e->ngroup; g++) {
        n_agents += e->grp[g].n;
        if (e->type[e->grp[g].type].can_absorb)
            for (int j = 0; j < e->grp[g].n; j++) if (!e->pool[e->grp[g].slot[j]].absorbed) n_agents--;
    }
    fprintf(f, "F %d %d 0\n", n_agents, e->nev);
    static const int dir2angle[4] = {0, 90, 180, 270};
    for (int g = 0; g < e->ngroup; g++) {
        const Type *t = &e->type[e->grp[g].type];
        for (int j = 0; j < e->grp[g].n; j++) {
            const Agent *a = &e->pool[e->grp[g].slot[j]];
            if (t->can_absorb && !a->absorbed) continue;
            int hp = (int)(100 * a->hp / t->hp);
            hp = hp < 0 ? 0 : (hp > 100 ? 100 : hp);
            fprintf(f, "%d %d %d %d %d %d\n", a->id, hp, dir2angle[a->dir], a->x, a->y, g);
        }
    }
    for (int i = 0; i < e->nev; i++) fprintf(f, "0 %d %d %d\n", e->ev[3 * i], e->ev[3 * i + 1], e->ev[3 * i + 2]);
    fclose(f);
    if (e->frame_ct++ > 10000) { e->frame_ct = 0; e->file_ct++; }  /* frame_per_file = 10000 (RenderGenerator.cc:19,181-184) */
    return 0;
}
API int env_render_next_file(void *game) { Env *e = game; e->file_ct++; e->frame_ct = 0; return 0; }
API int gridworld_set_goal(void *game, int g, const char *m, const int *b) {     /* GridWorld.cc:667-679 (deprecated) */
    Env *e = game;
    (void)b;
    if (strcmp(m, "random")) die("invalid goal type in GridWorld::set_goal", NULL);
    /* two draws per agent; Agent::goal itself is never read anywhere in the reference */
    for (int i = 0; i < e->grp[g].n; i++) { rng_draw(e); rng_draw(e); }
    return 0;
}
API int discrete_snake_clear_dead(void *game) { (void)game; die("DiscreteSnake not restated", NULL); return 0; }
API int discrete_snake_add_object(void *game, int a, int b, const char *c, const int *d) { (void)game; (void)a; (void)b; (void)c; (void)d; die("DiscreteSnake not restated", NULL); return 0; }
