/*
 * cm3_oracle.c - CPU restatement of the reference env steppers.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / the CPU baseline.  The product library
 * (cm3_b200/csrc) never links or calls it and has no CPU fallback.
 *
 * It restates, literally and in float64 like the reference, the algorithms of
 * (all paths relative to /root/reference, commit b5677214):
 *     env/checkers.py                                                   (whole file)
 *     env/multiagent-particle-envs/multiagent/core.py:117-196           (World.step)
 *     env/multiagent-particle-envs/multiagent/environment.py:81-225     (step/_set_action)
 *     env/multiagent-particle-envs/multiagent/scenarios/multi-goal_spread.py:65-154
 * The Checkers world is kept as the reference keeps it - a dense [rows][cols][3] float64
 * array - deliberately NOT as the product's bitboards, so that the two are independent.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself, executed in the build
 * container through oracle/ref_shims.py by oracle/gen_golden.py; the resulting fixtures
 * are committed under tests/golden/ and tests/test_oracle_golden.py replays every one.
 *
 * The only third-party arithmetic on the path is NumPy's IEEE-754 float64 ufuncs
 * (np.sqrt, np.square, np.sum over 2 elements, np.logaddexp; NumPy is unpinned by the
 * reference, 2.3.5 in the container).  np.logaddexp is restated from NumPy's published
 * npy_logaddexp (npymath/npy_math_internal.h.src).
 *
 * Also here: Philox4x32-10 (Salmon et al., SC'11 - the published Random123 algorithm)
 * as the CPU twin of the device action / reset generator, which has no counterpart in
 * the reference (its RNG is host MT19937).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OCK_MAX_AGENTS 8

/* ====================================================================== Checkers */

typedef struct {
    int n_rows, n_columns, n_obs, n_agents, max_steps;
    int agents_r[OCK_MAX_AGENTS]; /* BEFORE expansion, as passed to Checkers.__init__ */
    int agents_c[OCK_MAX_AGENTS];
} ock_config;

typedef struct {
    ock_config cfg;
    int total_rows, total_columns, max_collectible; /* checkers.py:24-28 */
    int B;
    double *world;     /* [B][total_rows][total_columns][3]   checkers.py:267 */
    int *loc;          /* [B][N][2]  agents_location          checkers.py:279 */
    double *collected; /* [B][N][2]  agents_collected         checkers.py:286 */
    int *steps;        /* [B] */
    int *goal_idx;     /* [B][N]  np.where(goals[idx]==1)[0][0]  checkers.py:235 */
} ock_env;

#define W_AT(e, b, r, c, ch) \
    ((e)->world[((((size_t)(b)) * (e)->total_rows + (r)) * (e)->total_columns + (c)) * 3 + (ch)])

ock_env *ock_create(const ock_config *cfg, int B) {
    /* checkers.py:16-17 asserts */
    if (cfg->n_rows % 2 != 1 || cfg->n_columns % 2 != 0) return NULL;
    if (cfg->n_agents < 1 || cfg->n_agents > OCK_MAX_AGENTS || B < 1) return NULL;
    ock_env *e = (ock_env *)calloc(1, sizeof(ock_env));
    e->cfg = *cfg;
    e->total_rows = cfg->n_rows + 2 * cfg->n_obs;        /* checkers.py:24 */
    e->total_columns = cfg->n_columns + 2 * cfg->n_obs + 1; /* checkers.py:25 */
    e->max_collectible = cfg->n_rows * cfg->n_columns;   /* checkers.py:28 */
    e->B = B;
    size_t cells = (size_t)B * e->total_rows * e->total_columns * 3;
    e->world = (double *)calloc(cells, sizeof(double));
    e->loc = (int *)calloc((size_t)B * cfg->n_agents * 2, sizeof(int));
    e->collected = (double *)calloc((size_t)B * cfg->n_agents * 2, sizeof(double));
    e->steps = (int *)calloc(B, sizeof(int));
    e->goal_idx = (int *)calloc((size_t)B * cfg->n_agents, sizeof(int));
    return e;
}

void ock_destroy(ock_env *e) {
    if (!e) return;
    free(e->world); free(e->loc); free(e->collected); free(e->steps); free(e->goal_idx);
    free(e);
}

/* checkers.py:38-63 */
static void ock_populate_world(ock_env *e, int b) {
    const int O = e->cfg.n_obs, R = e->cfg.n_rows, C = e->cfg.n_columns;
    const int TR = e->total_rows, TC = e->total_columns;
    /* invalid cells, :43-46 */
    for (int r = 0; r < TR; ++r) for (int c = 0; c < O; ++c) W_AT(e, b, r, c, 2) = 1.0;
    for (int r = 0; r < O; ++r) for (int c = 0; c < TC; ++c) W_AT(e, b, r, c, 2) = 1.0;
    for (int r = O + R; r < TR; ++r) for (int c = 0; c < TC; ++c) W_AT(e, b, r, c, 2) = 1.0;
    for (int r = O; r < O + R; ++r) for (int c = O + C + 1; c < TC; ++c) W_AT(e, b, r, c, 2) = 1.0;
    /* agent cells are invalid too, :49-51 */
    for (int i = 0; i < e->cfg.n_agents; ++i) {
        const int *l = &e->loc[((size_t)b * e->cfg.n_agents + i) * 2];
        W_AT(e, b, l[0], l[1], 2) = -1.0;
    }
    /* rewards, :54-63 */
    int green_first = 1;
    for (int row = O; row < O + R; ++row) {
        if (green_first) {
            for (int c = O; c < O + C; c += 2) W_AT(e, b, row, c, 0) = -1.0;
            for (int c = O + 1; c < O + C; c += 2) W_AT(e, b, row, c, 1) = -1.0;
            green_first = 0;
        } else {
            for (int c = O; c < O + C; c += 2) W_AT(e, b, row, c, 1) = -1.0;
            for (int c = O + 1; c < O + C; c += 2) W_AT(e, b, row, c, 0) = -1.0;
            green_first = 1;
        }
    }
}

/* checkers.py:112-125 (1-D case) */
static void ock_normalize(const ock_env *e, const int *loc, double *out) {
    out[0] = ((double)loc[0] - e->total_rows / 2.0) / e->total_rows;
    out[1] = ((double)loc[1] - e->total_columns / 2.0) / e->total_columns;
}

typedef struct {
    double *grid;          /* [B][R][C+1][2] */
    double *vec;           /* [B][N][4] */
    double *obs_others;    /* [B][N][2*max(N-1,1)] */
    double *obs_self_t;    /* [B][N][W][W][3] */
    double *obs_self_v;    /* [B][N][4] */
    double *reward;        /* [B] */
    double *local_rewards; /* [B][N] */
    uint8_t *done;         /* [B] */
} ock_outputs;

/* checkers.py:66-94 (global state) and :97-154 (local observation) */
static void ock_observe(const ock_env *e, int b, const ock_outputs *o) {
    const int O = e->cfg.n_obs, R = e->cfg.n_rows, C = e->cfg.n_columns, N = e->cfg.n_agents;
    const int Wd = 2 * O + 1;
    const int L = 2 * (N > 1 ? N - 1 : 1);
    /* get_valid_grid :66-76 -> world[O:O+R, O:O+C+1, 0:2] */
    double *g = o->grid + (size_t)b * R * (C + 1) * 2;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C + 1; ++c)
            for (int ch = 0; ch < 2; ++ch)
                g[(r * (C + 1) + c) * 2 + ch] = W_AT(e, b, O + r, O + c, ch);
    for (int i = 0; i < N; ++i) {
        const int *l = &e->loc[((size_t)b * N + i) * 2];
        const double *col = &e->collected[((size_t)b * N + i) * 2];
        /* get_global_state :90-93 */
        double *v = o->vec + ((size_t)b * N + i) * 4;
        v[0] = l[0]; v[1] = l[1]; v[2] = col[0]; v[3] = col[1];
        /* get_obs :97-109 */
        double *t = o->obs_self_t + ((size_t)b * N + i) * Wd * Wd * 3;
        for (int dr = 0; dr < Wd; ++dr)
            for (int dc = 0; dc < Wd; ++dc)
                for (int ch = 0; ch < 3; ++ch)
                    t[(dr * Wd + dc) * 3 + ch] = W_AT(e, b, l[0] - O + dr, l[1] - O + dc, ch);
        t[(O * Wd + O) * 3 + 2] = 0.0; /* :107 */
        /* obs_self_v :137-139 */
        double *sv = o->obs_self_v + ((size_t)b * N + i) * 4;
        ock_normalize(e, l, sv);
        sv[2] = col[0] / (e->max_collectible / 2.0);
        sv[3] = col[1] / (e->max_collectible / 2.0);
        /* obs_others :146-150 */
        double *oo = o->obs_others + ((size_t)b * N + i) * L;
        if (N == 1) {
            ock_normalize(e, l, oo);
        } else {
            int k = 0;
            for (int j = 0; j < N; ++j) {
                if (j == i) continue;
                ock_normalize(e, &e->loc[((size_t)b * N + j) * 2], oo + 2 * k);
                ++k;
            }
        }
    }
}

/* checkers.py:265-291.  goal_idx[b*N+i] in {0,1}; mask NULL = all envs. */
void ock_reset(ock_env *e, const int *goal_idx, const uint8_t *mask, const ock_outputs *o,
               int nthreads) {
    const int N = e->cfg.n_agents;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int b = 0; b < e->B; ++b) {
        if (mask && !mask[b]) continue;
        memset(&W_AT(e, b, 0, 0, 0), 0,
               sizeof(double) * e->total_rows * e->total_columns * 3); /* :267 */
        e->steps[b] = 0;
        for (int i = 0; i < N; ++i) e->goal_idx[(size_t)b * N + i] = goal_idx[(size_t)b * N + i];
        for (int i = 0; i < N; ++i) {
            int r = e->cfg.agents_r[i] + e->cfg.n_obs; /* :34 */
            int c = e->cfg.agents_c[i] + e->cfg.n_obs; /* :35 */
            if (N == 1) /* :271-276 */
                r = (e->goal_idx[(size_t)b * N] == 0 ? 0 : 2) + e->cfg.n_obs;
            e->loc[((size_t)b * N + i) * 2 + 0] = r;
            e->loc[((size_t)b * N + i) * 2 + 1] = c;
            e->collected[((size_t)b * N + i) * 2 + 0] = 0.0;
            e->collected[((size_t)b * N + i) * 2 + 1] = 0.0;
        }
        ock_populate_world(e, b);
        if (o) {
            ock_observe(e, b, o);
            if (o->done) o->done[b] = 0; /* :291 returns False */
        }
    }
}

/* checkers.py:157-187 */
static double ock_agent_act(ock_env *e, int b, int idx, int action) {
    int *l = &e->loc[((size_t)b * e->cfg.n_agents + idx) * 2];
    const int r = l[0], c = l[1];
    double reward = 0.0;
    if (action == 0) {
    } else if (action == 1 && W_AT(e, b, r - 1, c, 2) == 0.0) {
        W_AT(e, b, r - 1, c, 2) = -1.0; W_AT(e, b, r, c, 2) = 0.0; l[0] = r - 1;
    } else if (action == 2 && W_AT(e, b, r + 1, c, 2) == 0.0) {
        W_AT(e, b, r + 1, c, 2) = -1.0; W_AT(e, b, r, c, 2) = 0.0; l[0] = r + 1;
    } else if (action == 3 && W_AT(e, b, r, c - 1, 2) == 0.0) {
        W_AT(e, b, r, c - 1, 2) = -1.0; W_AT(e, b, r, c, 2) = 0.0; l[1] = c - 1;
    } else if (action == 4 && W_AT(e, b, r, c + 1, 2) == 0.0) {
        W_AT(e, b, r, c + 1, 2) = -1.0; W_AT(e, b, r, c, 2) = 0.0; l[1] = c + 1;
    } else {
        reward = -0.1; /* :184-186 - also any out-of-range action */
    }
    return reward;
}

/* checkers.py:190-225 */
static double ock_get_reward(ock_env *e, int b, int idx, int goal) {
    const int *l = &e->loc[((size_t)b * e->cfg.n_agents + idx) * 2];
    double *col = &e->collected[((size_t)b * e->cfg.n_agents + idx) * 2];
    const int r = l[0], c = l[1];
    double reward = 0.0;
    if (W_AT(e, b, r, c, 0) == -1.0) { /* green */
        W_AT(e, b, r, c, 0) = 1.0;
        reward = (goal == 0) ? 1.0 : -0.5;
        col[0] += 1.0;
    } else if (W_AT(e, b, r, c, 1) == -1.0) { /* orange */
        W_AT(e, b, r, c, 1) = 1.0;
        reward = (goal == 0) ? -0.5 : 1.0;
        col[1] += 1.0;
    }
    return reward;
}

/* checkers.py:228-262.  actions [B][N] int32. */
void ock_step(ock_env *e, const int32_t *actions, const ock_outputs *o, int nthreads) {
    const int N = e->cfg.n_agents;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int b = 0; b < e->B; ++b) {
        double total = 0.0;
        for (int idx = 0; idx < N; ++idx) { /* in index order, :233 */
            double penalty = ock_agent_act(e, b, idx, actions[(size_t)b * N + idx]);
            double rew = penalty + ock_get_reward(e, b, idx, e->goal_idx[(size_t)b * N + idx]);
            o->local_rewards[(size_t)b * N + idx] = rew;
            total += rew; /* np.sum over <8 items is a left-to-right sum, :243 */
        }
        ock_observe(e, b, o);
        o->reward[b] = total;
        e->steps[b] += 1; /* :244 */
        int done;
        if (e->steps[b] == e->cfg.max_steps) { /* :246 equality, not >= */
            done = 1;
        } else if (N == 1) { /* :248-255 */
            const int ch = e->goal_idx[(size_t)b * N];
            double s = 0.0;
            for (int r = 0; r < e->total_rows; ++r)
                for (int c = 0; c < e->total_columns; ++c) s += W_AT(e, b, r, c, ch);
            done = (s == e->max_collectible / 2.0);
        } else { /* :256-260 */
            double s = 0.0;
            for (int r = 0; r < e->total_rows; ++r)
                for (int c = 0; c < e->total_columns; ++c)
                    s += W_AT(e, b, r, c, 0) + W_AT(e, b, r, c, 1);
            done = (s == (double)e->max_collectible);
        }
        o->done[b] = (uint8_t)done;
    }
}

void ock_get_steps(const ock_env *e, int *steps) { memcpy(steps, e->steps, sizeof(int) * e->B); }

/* ====================================================================== Particle */

typedef struct {
    int n_agents, max_steps;
    double dt;             /* core.py:94   0.1  */
    double damping;        /* core.py:96   0.25 */
    double contact_force;  /* core.py:98   1e2  */
    double contact_margin; /* core.py:99   1e-3 */
    double agent_size;     /* multi-goal_spread.py:47  0.15 */
    double mass;           /* core.py:47-51  1.0 */
    double sensitivity;    /* environment.py:211  5.0 */
    double reach_thresh;   /* multi-goal_spread.py:126  0.05 */
} opt_config;

typedef struct {
    opt_config cfg;
    int B;
    double *pos;       /* [B][N][2] agent.state.p_pos */
    double *vel;       /* [B][N][2] agent.state.p_vel */
    double *landmarks; /* [B][N][2] landmark.state.p_pos */
    int *steps;        /* [B] env.steps */
    int64_t *collisions; /* [B] scenario.collisions */
    uint8_t *reached;  /* [B][N] agent.reached */
} opt_env;

typedef struct {
    double *global_state; /* [B][N][4] (vel, pos) rows, environment.py:113-116 */
    double *obs_others;   /* [B][N][4*max(N-1,1)] */
    double *obs_self;     /* [B][N][4] */
    double *reward;       /* [B] */
    double *reward_n;     /* [B][N] */
    uint8_t *done;        /* [B] */
} opt_outputs;

void opt_default_config(opt_config *c, int n_agents, int max_steps) {
    c->n_agents = n_agents; c->max_steps = max_steps;
    c->dt = 0.1; c->damping = 0.25; c->contact_force = 1e+2; c->contact_margin = 1e-3;
    c->agent_size = 0.15; c->mass = 1.0; c->sensitivity = 5.0; c->reach_thresh = 0.05;
}

opt_env *opt_create(const opt_config *cfg, int B) {
    if (cfg->n_agents < 1 || cfg->n_agents > OCK_MAX_AGENTS || B < 1) return NULL;
    opt_env *e = (opt_env *)calloc(1, sizeof(opt_env));
    e->cfg = *cfg; e->B = B;
    size_t n = (size_t)B * cfg->n_agents;
    e->pos = (double *)calloc(n * 2, sizeof(double));
    e->vel = (double *)calloc(n * 2, sizeof(double));
    e->landmarks = (double *)calloc(n * 2, sizeof(double));
    e->steps = (int *)calloc(B, sizeof(int));
    e->collisions = (int64_t *)calloc(B, sizeof(int64_t));
    e->reached = (uint8_t *)calloc(n, 1);
    return e;
}

void opt_destroy(opt_env *e) {
    if (!e) return;
    free(e->pos); free(e->vel); free(e->landmarks); free(e->steps); free(e->collisions);
    free(e->reached); free(e);
}

/* State injection (parity protocol): any pointer may be NULL. */
void opt_set_state(opt_env *e, const double *pos, const double *vel, const double *landmarks,
                   const int *steps, const int64_t *collisions, const uint8_t *reached) {
    size_t n = (size_t)e->B * e->cfg.n_agents;
    if (pos) memcpy(e->pos, pos, n * 2 * sizeof(double));
    if (vel) memcpy(e->vel, vel, n * 2 * sizeof(double));
    if (landmarks) memcpy(e->landmarks, landmarks, n * 2 * sizeof(double));
    if (steps) memcpy(e->steps, steps, e->B * sizeof(int));
    if (collisions) memcpy(e->collisions, collisions, e->B * sizeof(int64_t));
    if (reached) memcpy(e->reached, reached, n);
}

void opt_get_state(const opt_env *e, double *pos, double *vel, double *landmarks, int *steps,
                   int64_t *collisions, uint8_t *reached) {
    size_t n = (size_t)e->B * e->cfg.n_agents;
    if (pos) memcpy(pos, e->pos, n * 2 * sizeof(double));
    if (vel) memcpy(vel, e->vel, n * 2 * sizeof(double));
    if (landmarks) memcpy(landmarks, e->landmarks, n * 2 * sizeof(double));
    if (steps) memcpy(steps, e->steps, e->B * sizeof(int));
    if (collisions) memcpy(collisions, e->collisions, e->B * sizeof(int64_t));
    if (reached) memcpy(reached, e->reached, n);
}

/* NumPy npy_logaddexp (float64), as called by core.py:192 */
static double np_logaddexp(double x, double y) {
    if (x == y) return x + 0.693147180559945309417232121458176568; /* NPY_LOGE2 */
    const double tmp = x - y;
    if (tmp > 0) return x + log1p(exp(-tmp));
    else if (tmp <= 0) return y + log1p(exp(tmp));
    return tmp; /* NaN */
}

/* multi-goal_spread.py:145-154 + environment.py:113-116 */
static void opt_observe(const opt_env *e, int b, const opt_outputs *o) {
    const int N = e->cfg.n_agents;
    const int L = 4 * (N > 1 ? N - 1 : 1);
    const double *pos = e->pos + (size_t)b * N * 2, *vel = e->vel + (size_t)b * N * 2;
    for (int i = 0; i < N; ++i) {
        double *gs = o->global_state + ((size_t)b * N + i) * 4;
        double *os = o->obs_self + ((size_t)b * N + i) * 4;
        gs[0] = os[0] = vel[2 * i]; gs[1] = os[1] = vel[2 * i + 1];
        gs[2] = os[2] = pos[2 * i]; gs[3] = os[3] = pos[2 * i + 1];
        double *oo = o->obs_others + ((size_t)b * N + i) * L;
        int k = 0;
        for (int j = 0; j < N; ++j) {
            if (j == i && N > 1) continue; /* :148-151 (N == 1 keeps self) */
            oo[4 * k + 0] = vel[2 * j] - vel[2 * i];
            oo[4 * k + 1] = vel[2 * j + 1] - vel[2 * i + 1];
            oo[4 * k + 2] = pos[2 * j] - pos[2 * i];
            oo[4 * k + 3] = pos[2 * j + 1] - pos[2 * i + 1];
            ++k;
        }
    }
}

/* environment.py:125-149 after the caller has injected the reset_world() state:
 * velocities zero, reached False, collisions 0, steps 0 (multi-goal_spread.py:84-93). */
void opt_reset_to(opt_env *e, const double *pos, const double *landmarks, const uint8_t *mask,
                  const opt_outputs *o) {
    const int N = e->cfg.n_agents;
    for (int b = 0; b < e->B; ++b) {
        if (mask && !mask[b]) continue;
        for (int k = 0; k < 2 * N; ++k) {
            e->pos[(size_t)b * N * 2 + k] = pos[(size_t)b * N * 2 + k];
            e->vel[(size_t)b * N * 2 + k] = 0.0;
            e->landmarks[(size_t)b * N * 2 + k] = landmarks[(size_t)b * N * 2 + k];
        }
        for (int i = 0; i < N; ++i) e->reached[(size_t)b * N + i] = 0;
        e->collisions[b] = 0;
        e->steps[b] = 0;
        if (o) {
            opt_observe(e, b, o);
            if (o->done) o->done[b] = 0; /* np.any(done_n) with reached == False */
        }
    }
}

/* environment.py:81-123.  actions [B][N] int32. */
void opt_step(opt_env *e, const int32_t *actions, const opt_outputs *o, int nthreads) {
    const int N = e->cfg.n_agents;
    const opt_config *c = &e->cfg;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int b = 0; b < e->B; ++b) {
        double *pos = e->pos + (size_t)b * N * 2, *vel = e->vel + (size_t)b * N * 2;
        const double *lm = e->landmarks + (size_t)b * N * 2;
        double force[OCK_MAX_AGENTS][2];
        /* _set_action, environment.py:177-214 and apply_action_force, core.py:134-140 */
        for (int i = 0; i < N; ++i) {
            double u0 = 0.0, u1 = 0.0;
            const int a = actions[(size_t)b * N + i];
            if (a == 1) u0 = -1.0;
            if (a == 2) u0 = +1.0;
            if (a == 3) u1 = -1.0;
            if (a == 4) u1 = +1.0;
            u0 *= c->sensitivity; u1 *= c->sensitivity;
            force[i][0] = u0 + 0.0; /* u + noise, noise = 0.0 (u_noise None) */
            force[i][1] = u1 + 0.0;
        }
        /* apply_environment_force, core.py:143-155: entity pairs a < b; only agents
         * collide (landmarks have collide=False, multi-goal_spread.py:55) */
        for (int a = 0; a < N; ++a) {
            for (int bb = a + 1; bb < N; ++bb) {
                /* get_collision_force, core.py:180-196 */
                const double dx = pos[2 * a] - pos[2 * bb], dy = pos[2 * a + 1] - pos[2 * bb + 1];
                const double dist = sqrt(dx * dx + dy * dy);
                const double dist_min = c->agent_size + c->agent_size;
                const double k = c->contact_margin;
                const double pen = np_logaddexp(0.0, -(dist - dist_min) / k) * k;
                const double fx = c->contact_force * dx / dist * pen;
                const double fy = c->contact_force * dy / dist * pen;
                force[a][0] = fx + force[a][0];   force[a][1] = fy + force[a][1];
                force[bb][0] = -fx + force[bb][0]; force[bb][1] = -fy + force[bb][1];
            }
        }
        /* integrate_state, core.py:158-169 (max_speed None) */
        for (int i = 0; i < N; ++i) {
            for (int d = 0; d < 2; ++d) {
                double v = vel[2 * i + d] * (1 - c->damping);
                v += (force[i][d] / c->mass) * c->dt;
                vel[2 * i + d] = v;
                pos[2 * i + d] += v * c->dt;
            }
        }
        e->steps[b] += 1; /* environment.py:93 */
        /* per agent: observation -> reward -> done, environment.py:95-104 */
        opt_observe(e, b, o);
        double total = 0.0;
        int all_done = 1;
        for (int i = 0; i < N; ++i) {
            /* reward, multi-goal_spread.py:121-138 */
            double rew = 0.0;
            const double tx = pos[2 * i] - lm[2 * i], ty = pos[2 * i + 1] - lm[2 * i + 1];
            rew -= sqrt(tx * tx + ty * ty);
            e->reached[(size_t)b * N + i] = (rew >= -c->reach_thresh);
            for (int j = 0; j < N; ++j) {
                if (j == i) continue;
                /* is_collision, :114-118 */
                const double dx = pos[2 * j] - pos[2 * i], dy = pos[2 * j + 1] - pos[2 * i + 1];
                const double dist = sqrt(dx * dx + dy * dy);
                if (dist < c->agent_size + c->agent_size) {
                    rew -= 1.0;
                    e->collisions[b] += 1; /* double counted across the pair, :135-137 */
                }
            }
            o->reward_n[(size_t)b * N + i] = rew;
            total += rew; /* np.sum, environment.py:107 */
            if (!e->reached[(size_t)b * N + i]) all_done = 0; /* done(), :140-143 */
        }
        o->reward[b] = total;
        o->done[b] = (e->steps[b] == c->max_steps) || all_done; /* environment.py:118 */
    }
}

/* ====================================================================== Philox */
/* Philox4x32-10, Salmon/Moraes/Dror/Shaw, "Parallel Random Numbers: As Easy as 1, 2, 3"
 * (SC'11).  Constants from the paper / Random123. */
void oracle_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Uniform action stream used by the device rollout kernels when no action buffer is
 * given: counter = (global env id, global step index, stream tag, 0), key = seed;
 * agent i (< 4) takes word i, action = (word * n_actions) >> 32. */
void oracle_philox_actions(uint64_t seed, int64_t env0, int B, int N, int64_t t0, int T,
                           int n_actions, int8_t *actions /* [T][B][N] */) {
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int t = 0; t < T; ++t)
        for (int b = 0; b < B; ++b) {
            const uint64_t env = (uint64_t)(env0 + b), step = (uint64_t)(t0 + t);
            for (int blk = 0; blk * 4 < N; ++blk) {
                const uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32) ^ ((uint32_t)blk << 24),
                                         (uint32_t)step, 0xAC710000u | (uint32_t)(step >> 32)};
                uint32_t r[4];
                oracle_philox4x32_10(ctr, key, r);
                for (int i = 0; i < 4 && blk * 4 + i < N; ++i)
                    actions[((size_t)t * B + b) * N + blk * 4 + i] =
                        (int8_t)(((uint64_t)r[i] * (uint32_t)n_actions) >> 32);
            }
        }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
