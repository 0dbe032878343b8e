/*
 * wab2_oracle.c — CPU restatement of the reference's Environment 2.0 world turn (TEST INFRASTRUCTURE).
 *
 * Follows "/root/reference/Environment 2.0": World.py:93-132 (default_game_update), :243-316
 * (_get_visible_objects), :325-334 (perform_entity_action), :346-377 (reset_world, get_observations),
 * Ostrich.py / Wolf.py / Bush.py (entity state, Bush.take_food :31-39), WAB_Environment2.py:61-134 (creation,
 * reset_environment, take_action), WAB_Environment2_Single.py:36-69 (reset, step). Bug-compatible on purpose
 * (SURVEY Appendix C): the wolf hides GLOBAL LABEL j rather than its victim (World.py:112-115), bushes never
 * become invisible (:129-132 is a chained-assignment no-op), reset_world does not refresh table positions
 * (:353-356), respawn coordinates are drawn from [0, W] inclusive (Single.py:45-46).
 * Randomness: keyed Philox draws, sites 8-10 of oracle/ref_shim/v2.py. Parity status: pinned against the
 * reference run under oracle/ref_shim/v2.py (tests/test_v2_oracle.py) and tests/golden/v2_trace_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void wab_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);   /* wab_oracle.c */

enum { T_OSTRICH = 0, T_WOLF = 1, T_BUSH = 2 };
enum { SITE_V2_CREATE = 8, SITE_V2_RESET = 9, SITE_V2_PICK = 10 };

typedef struct Wab2OracleConfig {
    int32_t width, height, n_ostriches, n_wolves, n_bushes;
    int32_t lookout_view_radius, gatherer_view_radius, wolf_view_radius;   /* WAB_Environment2.py:35-36, :49 */
    int32_t starting_role;                                                   /* :19 */
    double ostrich_starting_food;                                            /* :32 */
    double wolf_starting_food, wolf_food_for_eating_ostrich;                 /* :43-44 */
    double food_per_bush, food_given_per_turn;                               /* :28-29 */
} Wab2OracleConfig;

typedef struct {
    int32_t type;
    int32_t x, y;          /* entity object coordinates, never wrapped (World.py:331-332 only wraps the table) */
    int32_t tx, ty;        /* table columns X, Y */
    int32_t visible;       /* table column Visible */
    double food;
    int32_t role;          /* ostrich role / wolf is_running / bush has_food */
    int32_t status;
} Ent;

typedef struct Wab2OracleWorld {
    Wab2OracleConfig cfg;
    uint64_t seed, env_id;
    int64_t episode;
    int32_t turn, acted, n;
    Ent *e;
} Wab2OracleWorld;

static int32_t keyed_int(const Wab2OracleWorld *w, uint32_t site, int64_t episode, uint32_t turn, uint32_t entity,
                         uint32_t axis, int32_t low, int32_t high) {
    uint32_t ctr[4] = { (uint32_t)w->env_id, (uint32_t)episode, (site << 28) | ((turn & 0xFFFFFu) << 8) | (axis & 0xFFu), entity };
    uint32_t key[2] = { (uint32_t)w->seed, (uint32_t)(w->seed >> 32) }, out[4];
    wab_oracle_philox(ctr, key, out);
    return low + (int32_t)(((uint64_t)out[0] * (uint64_t)(high - low + 1)) >> 32);
}
static int32_t pymod(int32_t a, int32_t m) { int32_t r = a % m; return r < 0 ? r + m : r; }

Wab2OracleWorld *wab2_oracle_create(const Wab2OracleConfig *cfg, uint64_t seed, uint64_t env_id) {
    Wab2OracleWorld *w = (Wab2OracleWorld *)calloc(1, sizeof(*w));
    w->cfg = *cfg; w->seed = seed; w->env_id = env_id;
    w->n = cfg->n_ostriches + cfg->n_wolves + cfg->n_bushes;
    w->e = (Ent *)calloc((size_t)w->n, sizeof(Ent));
    for (int32_t i = 0; i < w->n; ++i) {            /* create_ostriches / wolves / bushes, WAB_Environment2.py:61-110 */
        Ent *e = &w->e[i];
        e->type = i < cfg->n_ostriches ? T_OSTRICH : (i < cfg->n_ostriches + cfg->n_wolves ? T_WOLF : T_BUSH);
        e->x = e->tx = keyed_int(w, SITE_V2_CREATE, 0, 0, (uint32_t)i, 0, 0, cfg->width - 1);
        e->y = e->ty = keyed_int(w, SITE_V2_CREATE, 0, 0, (uint32_t)i, 1, 0, cfg->height - 1);
        e->visible = 1;
        if (e->type == T_OSTRICH) { e->food = cfg->ostrich_starting_food; e->role = cfg->starting_role; e->status = 0; }
        else if (e->type == T_WOLF) { e->food = cfg->wolf_starting_food; e->role = 0; e->status = 0; }
        else { e->food = cfg->food_per_bush; e->role = e->food > 0; e->status = 0; }
    }
    return w;
}
void wab2_oracle_destroy(Wab2OracleWorld *w) { if (w) { free(w->e); free(w); } }

/* reset_environment, WAB_Environment2.py:113-118 */
void wab2_oracle_reset(Wab2OracleWorld *w) {
    w->episode += 1;
    for (int32_t i = 0; i < w->n; ++i) {            /* Single.reset :36-42 -> entity.reset */
        Ent *e = &w->e[i];
        e->x = keyed_int(w, SITE_V2_RESET, w->episode, 0, (uint32_t)i, 0, 0, w->cfg.width);     /* inclusive, :45 */
        e->y = keyed_int(w, SITE_V2_RESET, w->episode, 0, (uint32_t)i, 1, 0, w->cfg.height);    /* :46 */
        if (e->type == T_OSTRICH) { e->food = w->cfg.ostrich_starting_food; e->role = w->cfg.starting_role; e->status = 0; }
        else if (e->type == T_WOLF) { e->food = w->cfg.wolf_starting_food; e->status = 0; e->role = 0; }
        else { e->food = w->cfg.food_per_bush; e->role = e->food > 0; }
    }
    w->acted = 0;
    for (int32_t i = 0; i < w->n; ++i) w->e[i].visible = 1;   /* reset_world :350-358; the X/Y write-back is a no-op */
    w->turn = 0;
}

/* default_game_update, World.py:93-132 */
static void game_update(Wab2OracleWorld *w, int32_t i) {
    Ent *a = &w->e[i];
    if (a->type == T_BUSH) return;
    const int32_t want = a->type == T_WOLF ? T_OSTRICH : T_BUSH;
    int32_t k = 0;
    for (int32_t q = 0; q < w->n; ++q) {
        const Ent *e = &w->e[q];
        k += (e->visible && e->tx == a->tx && e->ty == a->ty && e->type == want);
    }
    if (k == 0) return;
    const int32_t j = keyed_int(w, SITE_V2_PICK, w->episode, (uint32_t)w->turn, (uint32_t)i, 0, 0, k - 1);
    int32_t seen = 0;
    Ent *pick = NULL;
    for (int32_t q = 0; q < w->n && !pick; ++q) {
        Ent *e = &w->e[q];
        if (e->visible && e->tx == a->tx && e->ty == a->ty && e->type == want && seen++ == j) pick = e;
    }
    if (a->type == T_WOLF) {
        a->food += w->cfg.wolf_food_for_eating_ostrich;     /* :113 */
        pick->status = 2;                                   /* :114 */
        w->e[j].visible = 0;                                /* :115 — label j, not the victim */
    } else {
        double got;                                         /* Bush.take_food, Bush.py:31-39 */
        if (pick->food >= w->cfg.food_given_per_turn) { pick->food -= w->cfg.food_given_per_turn; got = w->cfg.food_given_per_turn; }
        else { got = pick->food; pick->food = 0; pick->role = 0; }
        a->food += got;                                     /* :127 */
    }
}

/* take_action, WAB_Environment2.py:125-134 -> Single.step :50-69 -> World.perform_entity_action :325-334 */
void wab2_oracle_take_action(Wab2OracleWorld *w, int32_t i, int32_t action, double *reward, int32_t *done) {
    Ent *e = &w->e[i];
    if (e->type == T_OSTRICH) {                             /* default_ostrich_act :25-43 */
        if (action == 0) e->y += 1; else if (action == 1) e->x += 1; else if (action == 2) e->y -= 1;
        else if (action == 3) e->x -= 1; else if (action == 4) e->role = 0; else if (action == 5) e->role = 1;
    } else if (e->type == T_WOLF) {                         /* default_wolf_act :61-73 */
        if (action == 0) e->y += 1; else if (action == 1) e->x += 1; else if (action == 2) e->y -= 1;
        else if (action == 3) e->x -= 1;
    }
    e->tx = pymod(e->x, w->cfg.width);                      /* :331 */
    e->ty = pymod(e->y, w->cfg.height);                     /* :332 */
    game_update(w, i);                                      /* :333 */
    if (e->type == T_OSTRICH) { *reward = e->status == 0 ? 1 : 0; *done = e->status != 0; }      /* :54-58, Ostrich.is_done */
    else if (e->type == T_WOLF) { *reward = e->food > 10 ? 1 : 0; *done = e->status == 1; }      /* :84-85, Wolf.is_done */
    else { *reward = 0; *done = 1; }                                                             /* :21-22, Bush.is_done */
    if (++w->acted == w->n) { w->turn += 1; w->acted = 0; } /* WAB_Environment2.py:131-133 */
}

/* wrap-aware delta along one axis, World.py:252-291 (only one direction is ever considered: if / elif) */
static int32_t axis_delta(int32_t obj, int32_t ent, int32_t r, int32_t size) {
    int32_t d = obj - ent;
    if (ent < r) {
        if (size - (r - ent) <= obj) {
            int32_t wrap = -ent - (size - obj);
            if (abs(wrap) < abs(d)) d = wrap;               /* min(d, wrap, key=abs): ties keep d */
        }
    } else if (size < ent + r) {
        if (obj <= r - size + ent) {
            int32_t wrap = obj + size - ent;
            if (abs(wrap) < abs(d)) d = wrap;
        }
    }
    return d;
}

/* get_observations, World.py:360-377. planes u8[3][2R+1][2R+1] (R = window radius given by the caller),
 * cell [dx+R][dy+R] = 1 if a visible object of that type is listed at that delta. Returns the number of rows
 * of the reference's DataFrame; internal5 = internal_obs (:17, :50-51, :80-81), bush padded with zeros. */
int32_t wab2_oracle_get_obs(const Wab2OracleWorld *w, int32_t i, int32_t R, uint8_t *planes, double *internal5) {
    const Ent *a = &w->e[i];
    int32_t r = 0;
    if (a->type == T_OSTRICH) r = a->role == 1 ? w->cfg.gatherer_view_radius : w->cfg.lookout_view_radius;
    else if (a->type == T_WOLF) r = w->cfg.wolf_view_radius;
    const int32_t S = 2 * R + 1;
    memset(planes, 0, (size_t)(3 * S * S));
    int32_t rows = 0;
    for (int32_t q = 0; q < w->n; ++q) {
        const Ent *e = &w->e[q];
        const int32_t dx = axis_delta(e->tx, a->tx, r, w->cfg.width);
        const int32_t dy = axis_delta(e->ty, a->ty, r, w->cfg.height);
        if (dx * dx + dy * dy > r * r) continue;            /* :295-297 */
        if (!e->visible) continue;                          /* :300 */
        rows++;
        if (abs(dx) <= R && abs(dy) <= R) planes[(e->type * S + (dx + R)) * S + (dy + R)] = 1;
    }
    internal5[0] = a->x; internal5[1] = a->y; internal5[2] = a->food;
    internal5[3] = a->type == T_BUSH ? 0 : a->role; internal5[4] = a->type == T_BUSH ? 0 : a->status;
    return rows;
}

/* 9 doubles per entity: type, x, y, tx, ty, visible, food, role, status */
void wab2_oracle_get_state(const Wab2OracleWorld *w, double *out) {
    for (int32_t i = 0; i < w->n; ++i) {
        const Ent *e = &w->e[i];
        double *o = out + 9 * i;
        o[0] = e->type; o[1] = e->x; o[2] = e->y; o[3] = e->tx; o[4] = e->ty; o[5] = e->visible; o[6] = e->food;
        o[7] = e->role; o[8] = e->status;
    }
}
int32_t wab2_oracle_turn(const Wab2OracleWorld *w) { return w->turn; }

/* Test hook, the inverse of wab2_oracle_get_state: 9 doubles per entity (the type column is ignored). */
void wab2_oracle_set_state(Wab2OracleWorld *w, const double *in, int32_t turn) {
    for (int32_t i = 0; i < w->n; ++i) {
        Ent *e = &w->e[i];
        const double *o = in + 9 * i;
        e->x = (int32_t)o[1]; e->y = (int32_t)o[2]; e->tx = (int32_t)o[3]; e->ty = (int32_t)o[4]; e->visible = (int32_t)o[5];
        e->food = o[6]; e->role = (int32_t)o[7]; e->status = (int32_t)o[8];
    }
    w->turn = turn; w->acted = 0;
}

/* Batch run for size-independent parity checks (the v2 counterpart of wab_oracle_run): n_envs worlds with ids
 * env_id_base .. +n_envs-1, `episodes` x (reset_environment + `turns` world turns) driven the way the reference
 * driver loop drives one world (Env2Tests.py:46-88): for every entity in id order get_obs(i) (acting entities
 * only: ostriches and wolves; bushes act with 0) then take_action(i, a). actions u8[episodes*turns][A][n_envs]
 * (entity-major, the layout of wab2_turn). Returns in out2:
 *   [0] sum over every acting-entity action of (a + 1) * (sum_k planes[k] * (k + 1) + 7 x + 11 y + 13 food + 17 role
 *       + 19 status + 23 reward + 29 done)        (planes with window radius R, internal obs as integers)
 *   [1] sum over worlds and entities k of (k + 1) * (x + 3 y + 5 X + 7 Y + 11 Visible + 13 food + 17 role + 19 status)
 *       of the final entity tables.
 * The return value is the number of world turns executed. */
#include <omp.h>
int64_t wab2_oracle_run(const Wab2OracleConfig *cfg, uint64_t seed, uint64_t env_id_base, int64_t n_envs, int32_t episodes,
                        int32_t turns, const uint8_t *actions, int32_t R, int32_t n_threads, int64_t *out2) {
    const int32_t A = cfg->n_ostriches + cfg->n_wolves, n = A + cfg->n_bushes, S = 2 * R + 1;
    int64_t cs_turns = 0, cs_state = 0, done_turns = 0;
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel reduction(+ : cs_turns, cs_state, done_turns)
    {
        uint8_t *planes = (uint8_t *)malloc((size_t)(3 * S * S));
        double *st = (double *)malloc(sizeof(double) * 9 * (size_t)n);
#pragma omp for schedule(dynamic, 16)
        for (int64_t e = 0; e < n_envs; ++e) {
            Wab2OracleWorld *w = wab2_oracle_create(cfg, seed, env_id_base + (uint64_t)e);
            for (int32_t ep = 0; ep < episodes; ++ep) {
                wab2_oracle_reset(w);
                for (int32_t t = 0; t < turns; ++t) {
                    const uint8_t *act = actions + ((size_t)(ep * turns + t) * (size_t)A) * (size_t)n_envs;
                    for (int32_t i = 0; i < n; ++i) {
                        int64_t c = 0;
                        if (i < A) {
                            double in5[5];
                            wab2_oracle_get_obs(w, i, R, planes, in5);
                            for (int32_t k = 0; k < 3 * S * S; ++k) c += (int64_t)planes[k] * (k + 1);
                            c += 7 * (int64_t)in5[0] + 11 * (int64_t)in5[1] + 13 * (int64_t)in5[2] + 17 * (int64_t)in5[3] + 19 * (int64_t)in5[4];
                        }
                        double reward; int32_t done;
                        wab2_oracle_take_action(w, i, i < A ? (int32_t)act[(size_t)i * (size_t)n_envs + (size_t)e] : 0, &reward, &done);
                        if (i < A) cs_turns += (int64_t)(i + 1) * (c + 23 * (int64_t)reward + 29 * (int64_t)done);
                    }
                    done_turns += 1;
                }
            }
            wab2_oracle_get_state(w, st);
            for (int32_t k = 0; k < n; ++k) {
                const double *o = st + 9 * k;
                cs_state += (int64_t)(k + 1) * ((int64_t)o[1] + 3 * (int64_t)o[2] + 5 * (int64_t)o[3] + 7 * (int64_t)o[4] + 11 * (int64_t)o[5] +
                                                13 * (int64_t)o[6] + 17 * (int64_t)o[7] + 19 * (int64_t)o[8]);
            }
            wab2_oracle_destroy(w);
        }
        free(planes); free(st);
    }
    out2[0] = cs_turns; out2[1] = cs_state;
    return done_turns;
}
