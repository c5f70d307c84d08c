/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C, single-threaded CPU restatement of the MAgent GridWorld battle path of
 * maomaomi1/Mean-Field-Multi-Agent-Reinforcement-Learning (examples/battle_model/src/gridworld).
 * Each function in magent_oracle.c cites the reference file:line it follows.
 *
 * Parity pinning: the reference ships no golden vectors or tests for this path
 * (examples/battle_model/src/gridworld/test.cc:6-152 is #if 0).  This oracle is therefore pinned
 * against the reference engine itself, compiled here from its own sources into
 * oracle/_ref/libmagent_ref.so (oracle/Makefile, OMP_NUM_THREADS=1) -- tests/test_oracle_vs_reference.py
 * steps both on identical action streams and compares every output bit for bit -- and against the
 * fixtures under tests/golden/ that were generated from that reference build
 * (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this library.
 */
#ifndef MFMARL_MAGENT_ORACLE_H
#define MFMARL_MAGENT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mo_env mo_env;

/* agent type attributes that matter on the battle path (AgentType.h:21-45) */
typedef struct {
    float hp, speed, view_radius, attack_radius;
    float damage, step_recover, kill_supply;
    float step_reward, kill_reward, dead_penalty, attack_penalty;
    float attack_bonus; /* value of the two `attack` reward rules (config/battle.py:41-42) */
} mo_type;

void mo_default_type(mo_type *t);                       /* config/battle.py:16-29 */

mo_env *mo_new(int width, int height, int embedding_size, const mo_type *type /* NULL = battle */);
void mo_free(mo_env *e);
void mo_set_seed(mo_env *e, unsigned long seed);        /* GridWorld.cc:150-151 */
void mo_reset(mo_env *e);                               /* GridWorld.cc:76-124 */
int  mo_add_agents(mo_env *e, int group, int n, const int *xs, const int *ys); /* GridWorld.cc:256-269; group -1 = walls :203-211 */

int  mo_get_num(const mo_env *e, int group);            /* GridWorld.cc:786-787 */
int  mo_view_size(const mo_env *e);                     /* view width == height */
int  mo_n_channel(const mo_env *e);
int  mo_feature_size(const mo_env *e);                  /* GridWorld.cc:1010-1018 */
int  mo_n_action(const mo_env *e);
int  mo_view_count(const mo_env *e);                    /* cells inside the view disc */
void mo_action_table(const mo_env *e, int *dxdy /* [n_action][2] */, int *attack_base);

void mo_get_observation(mo_env *e, int group, float *view, float *feature); /* GridWorld.cc:303-426 */
void mo_set_action(mo_env *e, int group, const int *actions);               /* GridWorld.cc:430-496 */
/* test hook: the next mo_step uses perm[] (a permutation of 0..A-1, new order -> pre-shuffle index)
 * instead of drawing from the engine RNG.  n must equal the attack count of that step. */
void mo_inject_attack_order(mo_env *e, const int *perm, int n);
int  mo_attack_count(const mo_env *e);
int  mo_step(mo_env *e);                                /* GridWorld.cc:498-694, returns done */
void mo_get_reward(mo_env *e, int group, float *buf);   /* GridWorld.cc:760-770 */
void mo_get_alive(mo_env *e, int group, unsigned char *buf); /* GridWorld.cc:801-806 */
void mo_get_id(mo_env *e, int group, int *buf);         /* GridWorld.cc:788-793 */
void mo_get_pos(mo_env *e, int group, int *buf);        /* GridWorld.cc:794-800 */
void mo_get_hp(mo_env *e, int group, float *buf);       /* not in the reference ABI; for state checks */
void mo_get_mean_info(mo_env *e, int group, float *buf);/* GridWorld.cc:849-870 */
void mo_clear_dead(mo_env *e);                          /* GridWorld.cc:696-728 */

/* RNG exposed for the host-side emulator tests: libstdc++ minstd_rand0 (GridWorld.h:106) */
unsigned long mo_rng_next(mo_env *e);
unsigned long mo_rng_state(const mo_env *e);

/* senario_battle.py:141,255 -- np.mean(one_hot(acts), axis=0) in float64 */
void mo_mean_action(const int *acts, int n, int n_action, double *out);

#ifdef __cplusplus
}
#endif
#endif
