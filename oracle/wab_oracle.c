/*
 * wab_oracle.c — CPU restatement of the reference hot path (TEST INFRASTRUCTURE, tier B).
 * See wab_oracle.h for scope and parity status. Every function cites the reference lines it follows
 * (/root/reference/wab_env.py). Data structures are deliberately the naive ones of the reference
 * (a record list of every bush cell ever revealed, a wolf list, an fp64 food level); nothing here
 * is shared with the CUDA implementation.
 */
#include "wab_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Philox4x32-10 (Random123) */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void wab_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Philox2x32-10 (Random123) */
void wab_oracle_philox2(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
    uint32_t c0 = ctr[0], c1 = ctr[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p = (uint64_t)0xD256D193u * c0;
        uint32_t n0 = (uint32_t)(p >> 32) ^ key ^ c1;
        c1 = (uint32_t)p; c0 = n0;
        key += PHILOX_W0;
    }
    out[0] = c0; out[1] = c1;
}

enum { SITE_BUSH = 1, SITE_INIT = 2, SITE_SPAWN = 3, SITE_DESP = 4, SITE_START = 5 };

typedef struct { int32_t x, y, food; } BushRec;

struct WabOracleEnv {
    WabOracleConfig cfg;
    uint32_t *thr; uint8_t *mlook, *mgath;
    uint64_t spawn_cdf[32], init_cdf[32];
    uint64_t seed, env_id; int64_t episode;
    int32_t turn;
    int32_t ox, oy; double food; int32_t role, status;      /* self.ostriches row 0 */
    int32_t *wx, *wy; int32_t nw, capw;                     /* self.wolves          */
    BushRec *recs; int32_t nrec, caprec;                    /* self.bushes          */
    int32_t *hslot; int64_t *hgen; uint32_t hmask;          /* (x,y) -> record index */
    uint32_t bush_ka, bush_kb;   /* the episode's bush key: words 2, 3 of the START call (oracle/keyed_rng.py) */
    int32_t *snap;     /* window bush food as of the last update_master_df_and_distances */
    int32_t snap_status;                                    /* ostrich status in that frame */
    int32_t nw_snap;                                        /* wolves present in that frame */
};

/* keyed draw: oracle/keyed_rng.py contract */
static uint32_t keyed_word(const WabOracleEnv *e, uint32_t site, uint32_t turn, uint32_t sub,
                           uint32_t payload, uint32_t lane) {
    uint32_t ctr[4] = { (uint32_t)e->env_id, (uint32_t)e->episode,
                        (site << 28) | ((turn & 0xFFFFFu) << 8) | (sub & 0xFFu), payload };
    uint32_t key[2] = { (uint32_t)e->seed, (uint32_t)(e->seed >> 32) };
    uint32_t out[4];
    wab_oracle_philox(ctr, key, out);
    return out[lane & 3u];
}
static double unit(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }
static uint32_t pack_xy(int32_t x, int32_t y) { return ((uint32_t)x & 0xFFFFu) | (((uint32_t)y & 0xFFFFu) << 16); }

/* binomial-first sites of oracle/keyed_rng.py: the n uniforms U_j the reference compares with p.
 * chosen[] (n bytes) marks the K cells that succeed; U_j = p * V_j for them, p + (1 - p) * V_j otherwise. */
static void binomial_first_choose(const WabOracleEnv *e, uint32_t site, uint32_t turn, int32_t n, const uint64_t *cdf,
                                  uint8_t *chosen) {
    uint64_t v = ((uint64_t)keyed_word(e, site, turn, 0, 0, 0) << 32) | keyed_word(e, site, turn, 0, 0, 1);
    int32_t k = 0;
    for (int32_t t = 0; t < 32; ++t) k += (v >= cdf[t]);
    if (k > n) k = n;
    memset(chosen, 0, (size_t)n);
    for (int32_t i = 0; i < k; ++i) {
        uint32_t r = keyed_word(e, site, turn, 1, (uint32_t)i >> 2, (uint32_t)i & 3u);
        int32_t q = (int32_t)(((uint64_t)r * (uint64_t)(n - i)) >> 32);
        for (int32_t j = 0; j < n; ++j)            /* the q-th index not chosen so far */
            if (!chosen[j] && q-- == 0) { chosen[j] = 1; break; }
    }
}
static double binomial_first_unit(const WabOracleEnv *e, uint32_t site, uint32_t turn, uint32_t j, int hit, double p) {
    double v = unit(keyed_word(e, site, turn, 2, j >> 2, j & 3u));
    return hit ? p * v : p + (1.0 - p) * v;
}

/* ------------------------------------------------------------------ bush record store */
static uint32_t cell_hash(int32_t x, int32_t y) {
    uint32_t h = (uint32_t)x * 0x9E3779B1u ^ ((uint32_t)y * 0x85EBCA77u + 0x165667B1u);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
static int32_t bush_find(const WabOracleEnv *e, int32_t x, int32_t y) {
    uint32_t h = cell_hash(x, y) & e->hmask;
    for (;;) {
        if (e->hgen[h] != e->episode) return -1;
        const BushRec *r = &e->recs[e->hslot[h]];
        if (r->x == x && r->y == y) return e->hslot[h];
        h = (h + 1) & e->hmask;
    }
}
static void bush_add(WabOracleEnv *e, int32_t x, int32_t y, int32_t food) {
    if (e->nrec == e->caprec) {
        e->caprec *= 2;
        e->recs = (BushRec *)realloc(e->recs, sizeof(BushRec) * (size_t)e->caprec);
    }
    if ((uint32_t)(e->nrec + 1) * 2u > e->hmask + 1u) {   /* grow + rehash */
        uint32_t ncap = (e->hmask + 1u) * 2u;
        free(e->hslot); free(e->hgen);
        e->hslot = (int32_t *)malloc(sizeof(int32_t) * ncap);
        e->hgen = (int64_t *)malloc(sizeof(int64_t) * ncap);
        for (uint32_t i = 0; i < ncap; ++i) e->hgen[i] = -1;
        e->hmask = ncap - 1u;
        for (int32_t i = 0; i < e->nrec; ++i) {
            uint32_t h = cell_hash(e->recs[i].x, e->recs[i].y) & e->hmask;
            while (e->hgen[h] == e->episode) h = (h + 1) & e->hmask;
            e->hgen[h] = e->episode; e->hslot[h] = i;
        }
    }
    e->recs[e->nrec].x = x; e->recs[e->nrec].y = y; e->recs[e->nrec].food = food;
    uint32_t h = cell_hash(x, y) & e->hmask;
    while (e->hgen[h] == e->episode) h = (h + 1) & e->hmask;
    e->hgen[h] = e->episode; e->hslot[h] = e->nrec;
    e->nrec++;
}
static void wolf_add(WabOracleEnv *e, int32_t x, int32_t y) {
    if (e->nw == e->capw) {
        e->capw *= 2;
        e->wx = (int32_t *)realloc(e->wx, sizeof(int32_t) * (size_t)e->capw);
        e->wy = (int32_t *)realloc(e->wy, sizeof(int32_t) * (size_t)e->capw);
    }
    e->wx[e->nw] = x; e->wy[e->nw] = y; e->nw++;
}

/* generate_n_bush_values, wab_env.py:631-635: round(U**bush_power * max_berries) through the
 * host-computed monotone threshold table (numpy's pow is not bit-reproducible in C). */
static int32_t bush_value(const WabOracleEnv *e, uint32_t w) {
    int32_t lo = 0, hi = e->cfg.n_bush_thr;           /* count of thresholds <= w */
    while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if (e->thr[mid] <= w) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* 32-bit bush draw of a cell (oracle/keyed_rng.py): two Philox2x32 calls on the cell's 2x2 block */
static uint32_t bush_word(const WabOracleEnv *e, int32_t x, int32_t y) {
    uint32_t k2 = (uint32_t)(e->seed ^ (e->seed >> 32));
    uint32_t lane = ((uint32_t)x & 1u) | (((uint32_t)y & 1u) << 1);
    uint32_t ctr[2] = { pack_xy(x >> 1, y >> 1) ^ e->bush_ka, e->bush_kb }, p[2], q[2];
    wab_oracle_philox2(ctr, k2, p);
    ctr[1] = ~e->bush_kb;
    wab_oracle_philox2(ctr, k2, q);
    uint32_t h = (p[lane >> 1] >> (16u * (lane & 1u))) & 0xFFFFu, l = (q[lane >> 1] >> (16u * (lane & 1u))) & 0xFFFFu;
    return (h << 16) | l;
}

/* generate_bushes, wab_env.py:613-629 (visible_coords :510-525) */
static void generate_bushes(WabOracleEnv *e) {
    const int32_t hw = e->cfg.width / 2, hh = e->cfg.height / 2;
    for (int32_t x = e->ox - hw; x <= e->ox + hw; ++x)
        for (int32_t y = e->oy - hh; y <= e->oy + hh; ++y) {
            if (bush_find(e, x, y) >= 0) continue;                       /* :624 */
            bush_add(e, x, y, bush_value(e, bush_word(e, x, y)));        /* :627-629 */
        }
}

/* initialize_wolves, wab_env.py:578-593 */
static void initialize_wolves(WabOracleEnv *e) {
    const int32_t hw = e->cfg.width / 2, hh = e->cfg.height / 2;
    const double p = e->cfg.chance_wolf_on_square / 2;                  /* :590 */
    const int32_t n = e->cfg.width * e->cfg.height;
    uint8_t *chosen = (uint8_t *)malloc((size_t)n);
    binomial_first_choose(e, SITE_INIT, 0, n, e->init_cdf, chosen);
    for (int32_t x = e->ox - hw; x <= e->ox + hw; ++x)
        for (int32_t y = e->oy - hh; y <= e->oy + hh; ++y) {
            uint32_t c = (uint32_t)((x - e->ox + hw) * e->cfg.height + (y - e->oy + hh));
            if (binomial_first_unit(e, SITE_INIT, 0, c, chosen[c], p) < p) wolf_add(e, x, y);
        }
    free(chosen);
}

/* spawn_wolves, wab_env.py:527-576 */
static void spawn_wolves(WabOracleEnv *e) {
    const int32_t hw = e->cfg.width / 2, hh = e->cfg.height / 2, m = e->cfg.wolf_spawn_margin;
    const double p = e->cfg.chance_wolf_on_square / 2;                  /* :573 */
    const int32_t n = (e->cfg.width + 2 * m) * (e->cfg.height + 2 * m) - e->cfg.width * e->cfg.height;
    uint8_t *chosen = (uint8_t *)malloc((size_t)n);
    binomial_first_choose(e, SITE_SPAWN, (uint32_t)e->turn, n, e->spawn_cdf, chosen);
    uint32_t j = 0;
    for (int32_t x = e->ox - hw - m; x < e->ox + hw + m + 1; ++x)        /* :536-547 */
        for (int32_t y = e->oy - hh - m; y < e->oy + hh + m + 1; ++y) {  /* :549-560 */
            int visible = (x >= e->ox - hw && x <= e->ox + hw && y >= e->oy - hh && y <= e->oy + hh);
            if (visible) continue;                                       /* :566 */
            if (binomial_first_unit(e, SITE_SPAWN, (uint32_t)e->turn, j, chosen[j], p) < p) wolf_add(e, x, y);
            ++j;
        }
    free(chosen);
}

/* update_master_df_and_distances, wab_env.py:504-508: the frame the kill/eat/obs code reads */
static void update_distances(WabOracleEnv *e) {
    const int32_t W = e->cfg.width, H = e->cfg.height, hw = W / 2, hh = H / 2;
    for (int32_t i = 0; i < W; ++i)
        for (int32_t j = 0; j < H; ++j) {
            /* grid index = delta + half, delta = ostrich - object (:59-60, :403-409) */
            int32_t x = e->ox - (i - hw), y = e->oy - (j - hh);
            int32_t r = bush_find(e, x, y);
            e->snap[i * H + j] = (r >= 0 && e->recs[r].food > 0) ? e->recs[r].food : 0;  /* :506 */
        }
    e->snap_status = e->status;
    e->nw_snap = e->nw;
}

/* _get_obs, wab_env.py:359-452 */
static void get_obs(const WabOracleEnv *e, WabOracleObs *obs) {
    const int32_t W = e->cfg.width, H = e->cfg.height, hw = W / 2, hh = H / 2;
    uint8_t *wolf = obs->grids, *bush = obs->grids + W * H, *ost = obs->grids + 2 * W * H;
    memset(obs->grids, 0, (size_t)(3 * W * H));
    for (int32_t k = 0; k < e->nw_snap; ++k) {                           /* :412-428 */
        int32_t dx = e->ox - e->wx[k], dy = e->oy - e->wy[k];
        if (abs(dx) * 2 < W && abs(dy) * 2 < H) wolf[(dx + hw) * H + (dy + hh)] = 1;  /* :418-419 */
    }
    for (int32_t c = 0; c < W * H; ++c) bush[c] = e->snap[c] > 0;        /* :430-444 */
    ost[hw * H + hh] = 1;                                                /* :393-410 */
    if (e->cfg.restrict_view) {                                          /* mask_grid :344-357 */
        const uint8_t *mask = (e->role == 1) ? e->mgath : e->mlook;
        for (int32_t c = 0; c < W * H; ++c)
            if (mask[c] == 1) { wolf[c] = 0; bush[c] = 0; ost[c] = 0; }
    }
    obs->food = (int32_t)ceil(e->food * e->cfg.turns_to_empty_food);    /* :452 */
    obs->role = e->role;                                                 /* :390 */
    obs->status = e->status;                                             /* :387 */
}

WabOracleEnv *wab_oracle_create(const WabOracleConfig *cfg, uint64_t seed, uint64_t env_id) {
    if (!cfg || cfg->width % 2 == 0 || cfg->height % 2 == 0) return NULL;  /* :147-148 */
    if (cfg->n_actions < 1 || cfg->n_actions > WAB_ORACLE_MAX_ACTIONS) return NULL;
    if (cfg->restrict_view && (!cfg->mask_lookout || !cfg->mask_gatherer)) return NULL;
    if (!cfg->spawn_cdf || !cfg->init_cdf) return NULL;
    WabOracleEnv *e = (WabOracleEnv *)calloc(1, sizeof(*e));
    e->cfg = *cfg;
    e->thr = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(cfg->n_bush_thr > 0 ? cfg->n_bush_thr : 1));
    memcpy(e->thr, cfg->bush_thr, sizeof(uint32_t) * (size_t)cfg->n_bush_thr);
    size_t cells = (size_t)cfg->width * (size_t)cfg->height;
    if (cfg->mask_lookout) { e->mlook = (uint8_t *)malloc(cells); memcpy(e->mlook, cfg->mask_lookout, cells); }
    if (cfg->mask_gatherer) { e->mgath = (uint8_t *)malloc(cells); memcpy(e->mgath, cfg->mask_gatherer, cells); }
    e->cfg.bush_thr = e->thr; e->cfg.mask_lookout = e->mlook; e->cfg.mask_gatherer = e->mgath;
    memcpy(e->spawn_cdf, cfg->spawn_cdf, sizeof(e->spawn_cdf)); memcpy(e->init_cdf, cfg->init_cdf, sizeof(e->init_cdf));
    e->cfg.spawn_cdf = e->spawn_cdf; e->cfg.init_cdf = e->init_cdf;
    e->seed = seed; e->env_id = env_id; e->episode = -1;
    e->capw = 8; e->wx = (int32_t *)malloc(sizeof(int32_t) * 8); e->wy = (int32_t *)malloc(sizeof(int32_t) * 8);
    e->caprec = 1024; e->recs = (BushRec *)malloc(sizeof(BushRec) * 1024);
    e->hmask = 4095u;
    e->hslot = (int32_t *)malloc(sizeof(int32_t) * 4096);
    e->hgen = (int64_t *)malloc(sizeof(int64_t) * 4096);
    for (int i = 0; i < 4096; ++i) e->hgen[i] = -1;
    e->snap = (int32_t *)calloc(cells, sizeof(int32_t));
    return e;
}

void wab_oracle_destroy(WabOracleEnv *e) {
    if (!e) return;
    free(e->thr); free(e->mlook); free(e->mgath); free(e->wx); free(e->wy);
    free(e->recs); free(e->hslot); free(e->hgen); free(e->snap); free(e);
}

/* reset, wab_env.py:231-248 */
void wab_oracle_reset(WabOracleEnv *e, WabOracleObs *obs) {
    e->episode++;                      /* keys: one episode per reset() call */
    e->turn = 0;                       /* :232 */
    e->nrec = 0; e->nw = 0;            /* :234-238 (hash entries expire with the episode stamp) */
    /* spawn_ostriches :595-611 */
    e->ox = 0; e->oy = 0; e->status = 0;
    e->bush_ka = keyed_word(e, SITE_START, 0, 0, 0, 2);
    e->bush_kb = keyed_word(e, SITE_START, 0, 0, 0, 3);
    e->food = (e->cfg.starting_food < 0) ? unit(keyed_word(e, SITE_START, 0, 0, 0, 0)) : e->cfg.starting_food;
    e->role = (e->cfg.starting_role < 0) ? (int32_t)(keyed_word(e, SITE_START, 0, 0, 0, 1) >> 31)
                                         : e->cfg.starting_role;
    generate_bushes(e);                /* :244 */
    if (e->cfg.wolves) initialize_wolves(e);   /* :245-246 */
    update_distances(e);               /* :247 */
    if (obs) get_obs(e, obs);          /* :248 */
}

/* step, wab_env.py:250-342 */
int wab_oracle_step(WabOracleEnv *e, int32_t action, WabOracleObs *obs, double *reward_out, int32_t *done_out) {
    if (action < 0 || action >= e->cfg.n_actions) return -1;            /* IndexError at :253 */
    double reward = 0;                                                   /* :251 */
    e->turn += 1;                                                        /* :252 */
    e->ox += e->cfg.action_dx[action];                                   /* :255 */
    e->oy += e->cfg.action_dy[action];                                   /* :256 */
    if (e->cfg.action_role[action] >= 0) e->role = e->cfg.action_role[action];  /* :257-258 */
    generate_bushes(e);                                                  /* :259 */

    /* despawn :262-264 — keep iff U > chance; rank = ordinal among earlier wolves on the cell */
    {
        uint8_t *keep = (uint8_t *)malloc((size_t)e->nw + 1);
        for (int32_t k = 0; k < e->nw; ++k) {      /* ranks over the unfiltered frame */
            uint32_t rank = 0;
            for (int32_t q = 0; q < k; ++q) rank += (e->wx[q] == e->wx[k] && e->wy[q] == e->wy[k]);
            uint32_t w = keyed_word(e, SITE_DESP, (uint32_t)e->turn, rank >> 2, pack_xy(e->wx[k], e->wy[k]), rank & 3u);
            keep[k] = unit(w) > e->cfg.wolf_chance_to_despawn;
        }
        int32_t o = 0;
        for (int32_t k = 0; k < e->nw; ++k)
            if (keep[k]) { e->wx[o] = e->wx[k]; e->wy[o] = e->wy[k]; ++o; }
        e->nw = o;
        free(keep);
    }

    update_distances(e);                                                 /* :266 */
    if (e->cfg.wolves_can_move) {                                        /* :267-289 */
        for (int32_t k = 0; k < e->nw; ++k) {
            int32_t dx = e->ox - e->wx[k], dy = e->oy - e->wy[k];        /* :59-60 */
            int32_t sx = (dx > 0) - (dx < 0), sy = (dy > 0) - (dy < 0);
            int32_t mx = (abs(dx) >= abs(dy)) * sx;                      /* :278-280 */
            int32_t my = (abs(dx) < abs(dy)) * sy;                       /* :281-283 */
            e->wx[k] += mx; e->wy[k] += my;                              /* :285-286 */
        }
        update_distances(e);                                             /* :289 */
    }
    if (!e->cfg.god_mode) {                                              /* :292-297 */
        for (int32_t k = 0; k < e->nw; ++k)
            if (e->wx[k] == e->ox && e->wy[k] == e->oy) { e->status = 2; break; }
    }
    /* eat :300-313 (bush and status as of the frame above) */
    {
        const int32_t H = e->cfg.height, hw = e->cfg.width / 2, hh = H / 2;
        if (e->snap[hw * H + hh] > 0 && (e->role == 1 || e->cfg.lookout_only) && e->snap_status == 0) {
            e->food += 1 / e->cfg.turns_to_fill_food;                    /* :307-309 */
            if (e->food < 0) e->food = 0;                                /* :310 clip(0, 1) */
            if (e->food > 1) e->food = 1;
            e->recs[bush_find(e, e->ox, e->oy)].food -= 1;               /* :312 */
            reward += e->cfg.reward_for_eating;                          /* :313 */
        }
    }
    e->food -= 1 / e->cfg.turns_to_empty_food;                           /* :316 */
    if (e->food <= 0) { e->status = 1; e->food = 0; }                    /* :319-322 */
    if (e->cfg.wolves) spawn_wolves(e);                                  /* :325-326 */

    int32_t done;
    if (e->status == 0) {                                                /* :328-334 */
        if (e->turn >= e->cfg.max_turns) { reward += e->cfg.reward_for_finishing; done = 1; }
        else { reward += e->cfg.reward_per_turn; done = 0; }
    } else if (e->status == 1) { reward += e->cfg.reward_for_starving; done = 1; }   /* :335-337 */
    else { reward += e->cfg.reward_for_being_killed; done = 1; }                     /* :338-340 */
    if (obs) get_obs(e, obs);                                            /* :342 */
    if (reward_out) *reward_out = reward;
    if (done_out) *done_out = done;
    return 0;
}

void wab_oracle_get_state(const WabOracleEnv *e, int32_t *x, int32_t *y, double *food, int32_t *role,
                          int32_t *status, int32_t *turn, int64_t *episode) {
    if (x) *x = e->ox;
    if (y) *y = e->oy;
    if (food) *food = e->food;
    if (role) *role = e->role;
    if (status) *status = e->status;
    if (turn) *turn = e->turn;
    if (episode) *episode = e->episode;
}
int32_t wab_oracle_num_wolves(const WabOracleEnv *e) { return e->nw; }
void wab_oracle_get_wolves(const WabOracleEnv *e, int32_t *out) {
    for (int32_t k = 0; k < e->nw; ++k) { out[2 * k] = e->wx[k]; out[2 * k + 1] = e->wy[k]; }
}
int32_t wab_oracle_num_bushes(const WabOracleEnv *e) { return e->nrec; }
void wab_oracle_get_bushes(const WabOracleEnv *e, int32_t *out) {
    for (int32_t k = 0; k < e->nrec; ++k) {
        out[3 * k] = e->recs[k].x; out[3 * k + 1] = e->recs[k].y; out[3 * k + 2] = e->recs[k].food;
    }
}

/* position-weighted checksum of one observation + step result (size-independent parity property) */
static uint64_t obs_checksum(const WabOracleObs *obs, int32_t cells3, int32_t done) {
    uint64_t s = 0;
    for (int32_t k = 0; k < cells3; ++k) s += (uint64_t)obs->grids[k] * (uint64_t)(k + 1);
    s += 1000ull * (uint64_t)obs->food + 100000ull * (uint64_t)obs->role + 200000ull * (uint64_t)obs->status
       + 400000ull * (uint64_t)done;
    return s;
}

int64_t wab_oracle_run(const WabOracleConfig *cfg, uint64_t seed, int64_t n_envs, int64_t n_steps,
                       const uint8_t *actions, int32_t n_threads, uint64_t *checksum_out) {
    uint64_t total = 0;
    const int32_t cells3 = 3 * cfg->width * cfg->height;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel reduction(+ : total)
    {
        uint8_t *grids = (uint8_t *)malloc((size_t)cells3);
        WabOracleObs obs; obs.grids = grids;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n_envs; ++i) {
            WabOracleEnv *e = wab_oracle_create(cfg, seed, (uint64_t)i);
            wab_oracle_reset(e, &obs);
            for (int64_t t = 0; t < n_steps; ++t) {
                double r; int32_t done;
                wab_oracle_step(e, actions[t * n_envs + i], &obs, &r, &done);
                if (done) wab_oracle_reset(e, &obs);       /* VecEnv auto-reset: post-reset obs */
                total += obs_checksum(&obs, cells3, done);
            }
            wab_oracle_destroy(e);
        }
        free(grids);
    }
    if (checksum_out) *checksum_out = total;
    return n_envs * n_steps;
}

/* Egocentric proximity observations, wab_env.py:637-667 (+ generate_potential_actions :71-84, the EgoCentric
 * _get_obs :951-958): for the five squares the ostrich can reach next (up, right, down, left, stay) the taxicab
 * distance to the nearest wolf (ALL wolves, fresh spawns included) and to the nearest bush record with food > 0 (EVERY
 * cell ever seen this episode, not just the window), as proximity = clip(max_distance - distance, 0, max_distance),
 * max_distance = width // 2 + height // 2 + 1 (:932-934). Without any wolf (bush) the reference takes the distances
 * to be pd.Series([0] * 5) (:648, :664), i.e. proximity max_distance. out10 = wolves[5], bushes[5]. */
void wab_oracle_ego_proximities(const WabOracleEnv *e, int32_t *out10) {
    const int32_t maxd = e->cfg.width / 2 + e->cfg.height / 2 + 1;
    const int32_t cx[5] = { e->ox, e->ox + 1, e->ox, e->ox - 1, e->ox };
    const int32_t cy[5] = { e->oy + 1, e->oy, e->oy - 1, e->oy, e->oy };
    for (int a = 0; a < 5; ++a) {
        int32_t dw = -1, db = -1;
        for (int32_t k = 0; k < e->nw; ++k) {
            const int32_t d = abs(cx[a] - e->wx[k]) + abs(cy[a] - e->wy[k]);
            if (dw < 0 || d < dw) dw = d;
        }
        for (int32_t k = 0; k < e->nrec; ++k) {
            if (e->recs[k].food <= 0) continue;
            const int32_t d = abs(cx[a] - e->recs[k].x) + abs(cy[a] - e->recs[k].y);
            if (db < 0 || d < db) db = d;
        }
        if (dw < 0) dw = 0;
        if (db < 0) db = 0;
        int32_t pw = maxd - dw, pb = maxd - db;
        out10[a] = pw < 0 ? 0 : (pw > maxd ? maxd : pw);
        out10[5 + a] = pb < 0 ? 0 : (pb > maxd ? maxd : pb);
    }
}
