/* TEST INFRASTRUCTURE ONLY -- see magent_oracle.h for the scope and how parity is pinned.
 *
 * Plain-C restatement of the reference battle path.  Reference paths are relative to
 * examples/battle_model/src/gridworld/ unless stated.  Battle config: no turn_mode, no food_mode,
 * minimap_mode on, 1x1 agents facing NORTH (GridWorld.cc:264), two groups of one agent type.
 */
#include "magent_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { OP_KILL = 3, OP_COLLIDE = 6, OP_ATTACK = 7, OP_NULL = 11 }; /* grid_def.h:18-24 */
enum { N_GROUP = 2 };

typedef struct { /* Range.h:15-109: rectangle + mask + offset list */
    int width, height, count;
    int x1, y1, x2, y2;
    unsigned char *in;
    int *dx, *dy;
} mo_range;

typedef struct mo_agent { /* GridWorld.h:134-258 */
    int id, dead, group, index;
    int x, y;
    float hp;
    int last_op;
    struct mo_agent *op_obj;
    int last_action;
    float next_reward, last_reward;
} mo_agent;

typedef struct { /* Map.h:23-29 plus the split-out channel id (Map.h:74) */
    int obstacle;
    mo_agent *occupier;
    int channel_id;
} mo_slot;

typedef struct { mo_agent *agent; int action; int band; } mo_act;

typedef struct {
    mo_agent **agents;
    int n, cap, dead_ct;
    float group_reward; /* Group::next_reward, always 0 for the battle rules (GridWorld.h:287) */
} mo_group;

struct mo_env {
    int w, h, embedding_size;
    mo_type type;
    mo_range view, attack, move;
    int n_action, attack_base;
    mo_slot *slots;
    mo_group groups[N_GROUP];
    int id_counter;
    unsigned long rng; /* minstd_rand0 state (GridWorld.h:106, seeded 0 at GridWorld.cc:31) */
    mo_act *attack_buf, *move_buf;
    int n_attack, n_move, cap_attack, cap_move;
    int *inject; int n_inject;
};

/* ---- Range.h:171-215 CircleRange ------------------------------------------------------------ */
static void circle_range(mo_range *r, float radius, float inner_radius, int parity) {
    const double eps = 1e-8;
    int width = 2 * (int)(radius + eps) + parity;
    int center = (int)radius;
    if (width % 2 != parity) width++;
    r->width = r->height = width;
    r->in = (unsigned char *)calloc((size_t)width * width, 1);
    r->dx = (int *)calloc((size_t)width * width, sizeof(int));
    r->dy = (int *)calloc((size_t)width * width, sizeof(int));
    r->count = 0;
    double delta = parity == 0 ? 0.5 : 0;
    for (int i = 0; i < width; i++)
        for (int j = 0; j < width; j++) {
            double dis_x = fabs(j - center + delta), dis_y = fabs(i - center + delta);
            double dis = sqrt(dis_x * dis_x + dis_y * dis_y);
            if (dis < radius + eps && dis > inner_radius - eps) {
                r->in[i * width + j] = 1;
                r->dx[r->count] = j - center;
                r->dy[r->count] = i - center;
                r->count++;
            }
        }
    r->x1 = r->y1 = -center;
    r->x2 = r->y2 = width - center - 1;
}
static void range_free(mo_range *r) { free(r->in); free(r->dx); free(r->dy); }

/* ---- GridWorld.cc:999-1018 channel layout --------------------------------------------------- */
static int group2channel(int group) { return 1 + group * 3; } /* wall | (has,hp,minimap) per group */

void mo_default_type(mo_type *t) { /* python/magent/builtin/config/battle.py:16-29,41-42 */
    t->hp = 10; t->speed = 2; t->view_radius = 6; t->attack_radius = 1.5f;
    t->damage = 2; t->step_recover = 0.1f; t->kill_supply = 0;
    t->step_reward = -0.005f; t->kill_reward = 5; t->dead_penalty = -0.1f; t->attack_penalty = -0.1f;
    t->attack_bonus = 0.2f;
}

mo_env *mo_new(int width, int height, int embedding_size, const mo_type *type) {
    mo_env *e = (mo_env *)calloc(1, sizeof(mo_env));
    e->w = width; e->h = height; e->embedding_size = embedding_size;
    if (type) e->type = *type; else mo_default_type(&e->type);
    /* AgentType.cc:87-112: parity = width % 2 = 1; attack inner radius = width / 2.0f */
    circle_range(&e->view, e->type.view_radius, 0, 1);
    circle_range(&e->attack, e->type.attack_radius, 0.5f, 1);
    circle_range(&e->move, e->type.speed, 0, 1);
    e->attack_base = e->move.count;               /* AgentType.cc:115-122 (no turn actions) */
    e->n_action = e->attack_base + e->attack.count;
    e->rng = 1;                                   /* minstd_rand0::seed(0) -> state 1 */
    return e;
}

static void free_agents(mo_env *e) {
    for (int g = 0; g < N_GROUP; g++) {
        for (int i = 0; i < e->groups[g].n; i++) free(e->groups[g].agents[i]);
        e->groups[g].n = 0; e->groups[g].dead_ct = 0;
    }
}

void mo_free(mo_env *e) {
    if (!e) return;
    free_agents(e);
    for (int g = 0; g < N_GROUP; g++) free(e->groups[g].agents);
    range_free(&e->view); range_free(&e->attack); range_free(&e->move);
    free(e->slots); free(e->attack_buf); free(e->move_buf); free(e->inject);
    free(e);
}

void mo_set_seed(mo_env *e, unsigned long seed) { /* libstdc++ linear_congruential_engine::seed */
    unsigned long s = seed % 2147483647UL;
    e->rng = s == 0 ? 1 : s;
}
unsigned long mo_rng_next(mo_env *e) { e->rng = (e->rng * 16807UL) % 2147483647UL; return e->rng; }
unsigned long mo_rng_state(const mo_env *e) { return e->rng; }

static int add_wall(mo_env *e, int x, int y) { /* Map.cc:108-115 */
    if (x < 0 || y < 0 || x >= e->w || y >= e->h) return 1;
    mo_slot *s = &e->slots[y * e->w + x];
    if (!s->obstacle && s->occupier) return 1;
    s->obstacle = 1; s->channel_id = 0;
    return 0;
}

void mo_reset(mo_env *e) { /* GridWorld.cc:76-124 + Map.cc:23-47; the RNG is NOT reseeded */
    e->id_counter = 0;
    free(e->slots);
    e->slots = (mo_slot *)calloc((size_t)e->w * e->h, sizeof(mo_slot));
    for (int i = 0; i < e->w * e->h; i++) e->slots[i].channel_id = -1;
    for (int i = 0; i < e->w; i++) { add_wall(e, i, 0); add_wall(e, i, e->h - 1); }
    for (int i = 0; i < e->h; i++) { add_wall(e, 0, i); add_wall(e, e->w - 1, i); }
    free_agents(e);
}

/* Map.cc:466-482 for a 1x1 body */
static int is_blank(const mo_env *e, int x, int y, const mo_agent *self) {
    if (x < 0 || y < 0 || x + 1 >= e->w || y + 1 >= e->h) return 0;
    const mo_slot *s = &e->slots[y * e->w + x];
    return !(s->obstacle || (s->occupier && s->occupier != self));
}

int mo_add_agents(mo_env *e, int group, int n, const int *xs, const int *ys) {
    int added = 0;
    if (group == -1) { /* GridWorld.cc:203-211 */
        for (int i = 0; i < n; i++) added += add_wall(e, xs[i], ys[i]) == 0;
        return added;
    }
    mo_group *g = &e->groups[group];
    for (int i = 0; i < n; i++) { /* GridWorld.cc:256-269, Map.cc:75-97, GridWorld.cc:180-187 */
        if (!is_blank(e, xs[i], ys[i], NULL)) continue; /* occupied: warned and ignored */
        mo_agent *a = (mo_agent *)calloc(1, sizeof(mo_agent));
        a->id = e->id_counter++; a->group = group; a->x = xs[i]; a->y = ys[i];
        a->hp = e->type.hp;
        a->last_action = e->n_action;            /* GridWorld.h:145 */
        a->last_reward = 0; a->last_op = OP_NULL; /* GridWorld.h:146-148,173-179 */
        a->next_reward = e->type.step_reward;
        mo_slot *s = &e->slots[a->y * e->w + a->x];
        s->occupier = a; s->channel_id = group2channel(group);
        if (g->n == g->cap) {
            g->cap = g->cap ? 2 * g->cap : 64;
            g->agents = (mo_agent **)realloc(g->agents, sizeof(mo_agent *) * g->cap);
        }
        a->index = 0;                            /* GridWorld.h:139: index(0) until clear_dead */
        g->agents[g->n++] = a;
        added++;
    }
    return added;
}

int mo_get_num(const mo_env *e, int group) { return e->groups[group].n; }
int mo_view_size(const mo_env *e) { return e->view.width; }
int mo_n_channel(const mo_env *e) { (void)e; return group2channel(N_GROUP); }
int mo_feature_size(const mo_env *e) { return e->embedding_size + e->n_action + 1 + 2; }
int mo_n_action(const mo_env *e) { return e->n_action; }
int mo_view_count(const mo_env *e) { return e->view.count; }
void mo_action_table(const mo_env *e, int *dxdy, int *attack_base) {
    for (int i = 0; i < e->move.count; i++) { dxdy[2 * i] = e->move.dx[i]; dxdy[2 * i + 1] = e->move.dy[i]; }
    for (int i = 0; i < e->attack.count; i++) {
        dxdy[2 * (e->attack_base + i)] = e->attack.dx[i];
        dxdy[2 * (e->attack_base + i) + 1] = e->attack.dy[i];
    }
    *attack_base = e->attack_base;
}

/* ---- GridWorld.cc:303-426 + Map.cc:130-218 -------------------------------------------------- */
void mo_get_observation(mo_env *e, int group, float *view, float *feature) {
    const mo_group *g = &e->groups[group];
    const int vw = e->view.width, vh = e->view.height, nc = mo_n_channel(e);
    const int fs = mo_feature_size(e), emb = e->embedding_size;
    const size_t per_view = (size_t)vh * vw * nc;
    memset(view, 0, sizeof(float) * g->n * per_view);          /* :329-330 */
    memset(feature, 0, sizeof(float) * (size_t)g->n * fs);

    /* make_channel_trans (:981-997): own (has,hp,minimap) -> 1,2,3; other -> 4,5,6 */
    int trans[16] = {0};
    { int base = group2channel(0);
      for (int i = 0; i < N_GROUP; i++) { trans[group2channel((group + i) % N_GROUP)] = base; base += 3; } }

    /* minimap (:341-380): count[y/s][x/s] over every agent still in the list, / total */
    const int scale_h = (e->h + vh - 1) / vh, scale_w = (e->w + vw - 1) / vw;
    float *minimap = (float *)calloc((size_t)vh * vw * N_GROUP, sizeof(float));
    for (int i = 0; i < N_GROUP; i++) {
        const mo_group *gi = &e->groups[i];
        size_t total = 0;
        for (int j = 0; j < gi->n; j++) {
            int x = gi->agents[j]->x / scale_w, y = gi->agents[j]->y / scale_h;
            minimap[(y * vw + x) * N_GROUP + i] += 1.0f; total++;
        }
        for (int k = 0; k < vh * vw; k++) minimap[k * N_GROUP + i] /= total; /* float / (float)size_t */
    }

    for (int i = 0; i < g->n; i++) {
        const mo_agent *a = g->agents[i];
        float *buf = view + i * per_view;
        /* Map::extract_view, dir NORTH, offsets 0: window [x+x1, x+x2] x [y+y1, y+y2] clamped */
        int x1 = a->x + e->view.x1, x2 = a->x + e->view.x2, y1 = a->y + e->view.y1, y2 = a->y + e->view.y2;
        int sx = x1 > 0 ? x1 : 0, ex = x2 < e->w - 1 ? x2 : e->w - 1;
        int sy = y1 > 0 ? y1 : 0, ey = y2 < e->h - 1 ? y2 : e->h - 1;
        for (int x = sx; x <= ex; x++)
            for (int y = sy; y <= ey; y++) {
                int vx = x - x1, vy = y - y1;
                const mo_slot *s = &e->slots[y * e->w + x];
                if (s->channel_id != -1 && e->view.in[vy * vw + vx]) {          /* Map.cc:202 */
                    int ch = trans[s->channel_id];
                    buf[(vy * vw + vx) * nc + ch] = 1;
                    if (s->occupier)                                            /* Map.cc:206-209 */
                        buf[(vy * vw + vx) * nc + ch + 1] = s->occupier->hp / e->type.hp;
                }
            }
        /* minimap channels + self marker in BOTH (:396-409); not disc-masked */
        int self_x = a->x / scale_w, self_y = a->y / scale_h;
        for (int j = 0; j < N_GROUP; j++) {
            int mc = trans[group2channel(j)] + 2;
            for (int k = 0; k < vh * vw; k++) buf[k * nc + mc] = minimap[k * N_GROUP + j];
            buf[(self_y * vw + self_x) * nc + mc] += 1;
        }
        /* features (:411-421): id bits LSB first (GridWorld.h:162-171), one-hot last action,
         * last reward, x/W, y/H */
        float *f = feature + (size_t)i * fs;
        for (int b = 0, t = a->id; b < emb; b++, t >>= 1) f[b] = (float)(t & 1);
        f[emb + a->last_action] = 1;
        f[emb + e->n_action] = a->last_reward;
        f[emb + e->n_action + 1] = (float)a->x / e->w;
        f[emb + e->n_action + 2] = (float)a->y / e->h;
    }
    free(minimap);
}

/* ---- GridWorld.cc:430-496: small-map branch :481-495; on a large map (w*h > 99*99, :79-88) a move is filed by
 * the x-band of its agent (:443-455): NUM_SEP_BUFFER bands of `bandwidth` columns, or the boundary buffer when the
 * agent stands within 4 columns of a band edge.  The band is kept next to the move; step() runs band 0 ..
 * NUM_SEP-1, then the boundary buffer (:662-672), each in call order. -------------------------------------------- */
static int move_band(const mo_env *e, int x) {
    if (e->w * e->h <= 99 * 99) return 0;
    const int nsep = e->w * e->h > 1000 * 1000 ? 16 : 8;
    const int bandwidth = (e->w + nsep - 1) / nsep;
    const int xr = x % bandwidth;
    return (xr < 4 || xr > bandwidth - 4) ? nsep : x / bandwidth;
}
static void push(mo_act **buf, int *n, int *cap, mo_agent *a, int action) {
    if (*n == *cap) { *cap = *cap ? 2 * *cap : 256; *buf = (mo_act *)realloc(*buf, sizeof(mo_act) * *cap); }
    (*buf)[*n].agent = a; (*buf)[*n].action = action; (*buf)[*n].band = 0; (*n)++;
}
void mo_set_action(mo_env *e, int group, const int *actions) {
    mo_group *g = &e->groups[group];
    for (int i = 0; i < g->n; i++) {
        mo_agent *a = g->agents[i];
        a->last_action = actions[i];
        if (actions[i] < e->attack_base) {
            push(&e->move_buf, &e->n_move, &e->cap_move, a, actions[i]);
            e->move_buf[e->n_move - 1].band = move_band(e, a->x);
        } else push(&e->attack_buf, &e->n_attack, &e->cap_attack, a, actions[i] - e->attack_base);
    }
}

void mo_inject_attack_order(mo_env *e, const int *perm, int n) {
    free(e->inject);
    e->inject = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    memcpy(e->inject, perm, sizeof(int) * n);
    e->n_inject = n;
}
int mo_attack_count(const mo_env *e) { return e->n_attack; }

static void remove_agent(mo_env *e, mo_agent *a) { /* Map.cc:99-106,516-525 */
    mo_slot *s = &e->slots[a->y * e->w + a->x];
    s->occupier = NULL; s->channel_id = -1;
}

/* ---- GridWorld.cc:498-694 ------------------------------------------------------------------- */
int mo_step(mo_env *e) {
    const mo_type *t = &e->type;
    /* shuffle (:510-515): inside-out Fisher-Yates, one draw per element including i = 0 */
    if (e->inject) {
        mo_act *tmp = (mo_act *)malloc(sizeof(mo_act) * (e->n_attack > 0 ? e->n_attack : 1));
        for (int i = 0; i < e->n_attack; i++) tmp[i] = e->attack_buf[e->inject[i]];
        memcpy(e->attack_buf, tmp, sizeof(mo_act) * e->n_attack);
        free(tmp); free(e->inject); e->inject = NULL; e->n_inject = 0;
    } else {
        for (int i = 0; i < e->n_attack; i++) {
            int j = (int)mo_rng_next(e) % (i + 1);
            mo_act sw = e->attack_buf[i]; e->attack_buf[i] = e->attack_buf[j]; e->attack_buf[j] = sw;
        }
    }
    /* attacks in shuffled order (:524-557; serial == OMP_NUM_THREADS=1) */
    for (int i = 0; i < e->n_attack; i++) {
        mo_agent *a = e->attack_buf[i].agent;
        if (a->dead) continue;                                            /* :527 */
        /* Map::get_attack_obj (Map.cc:220-263) */
        int ox = a->x + e->attack.dx[e->attack_buf[i].action];
        int oy = a->y + e->attack.dy[e->attack_buf[i].action];
        mo_agent *obj = NULL;
        if (ox >= 0 && ox < e->w && oy >= 0 && oy < e->h) obj = e->slots[oy * e->w + ox].occupier;
        if (!obj || obj->group == a->group) {                             /* miss (:537-540) */
            a->next_reward += t->attack_penalty;
            continue;
        }
        /* Map::do_attack (Map.cc:266-321) + Agent::be_attack (GridWorld.h:208-214) */
        float reward = 0.0f;
        obj->hp -= t->damage;
        if (obj->hp < 0.0) { obj->dead = 1; obj->next_reward = t->dead_penalty; }
        if (obj->dead) {
            a->last_op = OP_KILL; a->op_obj = obj;
            remove_agent(e, obj);
            e->groups[obj->group].dead_ct++;
            a->hp = fminf(t->hp, a->hp + t->kill_supply);                 /* add_hp, GridWorld.h:190 */
            reward = t->kill_reward;
        } else {
            a->last_op = OP_ATTACK; a->op_obj = obj;
        }
        a->next_reward += reward + t->attack_penalty;                     /* :556 */
    }
    e->n_attack = 0;

    /* starve (:574-595, GridWorld.h:199-206) */
    for (int g = 0; g < N_GROUP; g++) {
        mo_group *gr = &e->groups[g];
        int starve_ct = 0;
        for (int j = 0; j < gr->n; j++) {
            mo_agent *a = gr->agents[j];
            if (a->dead) continue;
            if (t->step_recover > 0) a->hp = fminf(t->hp, a->hp + t->step_recover);
            else {
                a->hp -= -t->step_recover;
                if (a->hp < 0.0) { a->dead = 1; a->next_reward = t->dead_penalty; }
            }
            if (a->dead) { remove_agent(e, a); starve_ct++; }
        }
        gr->dead_ct += starve_ct;
    }

    /* moves in set_action order (:631-672 + Map::do_move Map.cc:324-369); band by band on a large map */
    const int n_band = e->w * e->h <= 99 * 99 ? 1 : (e->w * e->h > 1000 * 1000 ? 17 : 9);
    for (int band = 0; band < n_band; band++)
    for (int j = 0; j < e->n_move; j++) {
        mo_agent *a = e->move_buf[j].agent;
        if (e->move_buf[j].band != band || a->dead) continue;
        int nx = a->x + e->move.dx[e->move_buf[j].action], ny = a->y + e->move.dy[e->move_buf[j].action];
        if (is_blank(e, nx, ny, a)) {
            mo_slot *olds = &e->slots[a->y * e->w + a->x];
            int ch = olds->channel_id;
            olds->occupier = NULL; olds->channel_id = -1;
            mo_slot *news = &e->slots[ny * e->w + nx];
            news->occupier = a; news->channel_id = ch;
            a->x = nx; a->y = ny;
        } else if (!(nx < 0 || ny < 0 || nx + 1 >= e->w || ny + 1 >= e->h)) { /* Map.cc:498-513 */
            mo_agent *c = e->slots[ny * e->w + nx].occupier;
            if (c && c != a) { a->last_op = OP_COLLIDE; a->op_obj = c; }
        }
    }
    e->n_move = 0;

    /* calc_reward (:744-758) for the two battle rules, via RewardEngine.cc:373-443,216-240:
     * for every agent of the rule's subject group (dead ones included) whose op_obj is in the object
     * group and whose last_op is OP_ATTACK, the subject receives the rule value */
    for (int r = 0; r < N_GROUP; r++) {
        mo_group *gr = &e->groups[r];
        for (int i = 0; i < gr->n; i++) {
            mo_agent *a = gr->agents[i];
            if (a->op_obj && a->op_obj->group == 1 - r && a->last_op == OP_ATTACK)
                a->next_reward += t->attack_bonus;
        }
    }

    /* done (:680-686): any group with no one alive */
    int live = 0;
    for (int g = 0; g < N_GROUP; g++) live += (e->groups[g].n - e->groups[g].dead_ct) > 0;
    return live < N_GROUP;
}

void mo_get_reward(mo_env *e, int group, float *buf) {
    mo_group *g = &e->groups[group];
    for (int i = 0; i < g->n; i++) buf[i] = g->agents[i]->next_reward + g->group_reward;
}
void mo_get_alive(mo_env *e, int group, unsigned char *buf) {
    for (int i = 0; i < e->groups[group].n; i++) buf[i] = !e->groups[group].agents[i]->dead;
}
void mo_get_id(mo_env *e, int group, int *buf) {
    for (int i = 0; i < e->groups[group].n; i++) buf[i] = e->groups[group].agents[i]->id;
}
void mo_get_pos(mo_env *e, int group, int *buf) {
    for (int i = 0; i < e->groups[group].n; i++) {
        buf[2 * i] = e->groups[group].agents[i]->x; buf[2 * i + 1] = e->groups[group].agents[i]->y;
    }
}
void mo_get_hp(mo_env *e, int group, float *buf) {
    for (int i = 0; i < e->groups[group].n; i++) buf[i] = e->groups[group].agents[i]->hp;
}

void mo_get_mean_info(mo_env *e, int group, float *buf) { /* GridWorld.cc:849-870 */
    mo_group *g = &e->groups[group];
    int *counter = (int *)calloc((size_t)e->n_action + 1, sizeof(int));
    float sum_x = 0, sum_y = 0;
    for (int i = 0; i < g->n; i++) {
        sum_x += g->agents[i]->x; sum_y += g->agents[i]->y;
        counter[g->agents[i]->last_action]++;   /* last_action == n_action before the first step */
    }
    size_t n = (size_t)g->n;
    buf[0] = sum_x / n; buf[1] = sum_y / n;
    for (int i = 0; i < e->n_action; i++) buf[2 + i] = (float)(1.0 * counter[i] / n);
    free(counter);
}

void mo_clear_dead(mo_env *e) { /* GridWorld.cc:696-728 + Agent::init_reward GridWorld.h:173-179 */
    for (int gi = 0; gi < N_GROUP; gi++) {
        mo_group *g = &e->groups[gi];
        g->group_reward = 0;
        int pt = 0;
        for (int j = 0; j < g->n; j++) {
            mo_agent *a = g->agents[j];
            if (a->dead) { free(a); continue; }
            a->last_reward = a->next_reward; a->last_op = OP_NULL;
            a->next_reward = e->type.step_reward; a->op_obj = NULL;
            a->index = pt;
            g->agents[pt++] = a;
        }
        g->n = pt; g->dead_ct = 0;
    }
}

void mo_mean_action(const int *acts, int n, int n_action, double *out) { /* senario_battle.py:141 */
    for (int k = 0; k < n_action; k++) out[k] = 0.0;
    for (int i = 0; i < n; i++) out[acts[i]] += 1.0;
    for (int k = 0; k < n_action; k++) out[k] /= (double)n;
}
