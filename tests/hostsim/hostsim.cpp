// hostsim.cpp — TEST-ONLY host build of the kernel logic header wab_gym_b200/csrc/wab_core.cuh.
//
// Compiles the per-environment step / reset / observation-composition code that the CUDA kernels
// run (the same header, the same Params builder) with g++, so that `-m "not gpu"` tests can compare
// it with the oracle in a container without a GPU. It serialises what the kernel fans out over a
// warp (36 bush blocks + 31 wolf-init groups of a reset; the 363-bit observation string of one env).
// It is NOT a CPU fallback: nothing in the product package loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../wab_gym_b200/csrc/wab_core.cuh"
#include "../../wab_gym_b200/csrc/wab_features.cuh"
#include "../../wab_gym_b200/csrc/wab_params.h"

using namespace wab;

struct HostSim {
    WabConfig cfg;
    Params P;
    std::vector<uint32_t> thr;
    Env E;
    std::vector<uint32_t> wolves, logcell;
    std::vector<uint8_t> logcnt;
    Slots S;
    bool f64;
};

static void expand_obs(const HostSim* h, const uint32_t wm_in[4], const uint32_t bm_in[4], uint32_t role, uint8_t* grids) {
    uint32_t B[12], wm[4], bm[4];
    for (int w = 0; w < 4; ++w) { wm[w] = wm_in[w]; bm[w] = bm_in[w]; }
    apply_view_mask(h->P, role, wm, bm);
    compose_obs(wm, bm, B);
    B[11] = 0u;   // bits 352..362 (last ostrich row) are never set
    for (int b = 0; b < OBS_BYTES; ++b) grids[b] = (uint8_t)((B[b >> 5] >> (b & 31)) & 1u);
}

extern "C" {

HostSim* hostsim_create(const WabConfig* cfg, const uint32_t* bush_thr, int32_t n_bush_thr, uint64_t seed,
                        uint64_t env_id, char* err, int errlen) {
    std::string e;
    if (validate_config(cfg, n_bush_thr, 1, e) != WAB_OK) {
        if (err && errlen > 0) { strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
        return nullptr;
    }
    HostSim* h = new HostSim();
    h->cfg = *cfg;
    h->thr.assign(bush_thr, bush_thr + n_bush_thr);
    params_from_config(*cfg, h->thr.data(), n_bush_thr, seed, 0, h->P);
    h->P.bush_thr = h->thr.data();
    h->f64 = cfg->food_mode == WAB_FOOD_F64;
    memset(&h->E, 0, sizeof(h->E));
    h->E.env_id = (uint32_t)env_id;
    h->E.episode = 0xFFFFFFFFu;
    h->wolves.assign(cfg->wolf_cap, 0); h->logcell.assign(cfg->log_cap, 0); h->logcnt.assign(cfg->log_cap, 0);
    h->S.wolves = h->wolves.data(); h->S.wstride = 1;
    h->S.logcell = h->logcell.data(); h->S.logcnt = h->logcnt.data(); h->S.lstride = 1;
    return h;
}
void hostsim_destroy(HostSim* h) { delete h; }

// returns overflow flag
static uint32_t do_reset(HostSim* h, uint32_t wm[4], uint32_t bm[4]) {
    uint32_t overflow = 0;
    if (h->f64) reset_scalars<true>(h->P, h->E); else reset_scalars<false>(h->P, h->E);
    uint32_t m[4] = {0, 0, 0, 0};
    for (int blk = 0; blk < 36; ++blk) reset_bush_block(h->P, h->E.bk_a, h->E.bk_b, blk, m);
    for (int w = 0; w < 4; ++w) h->E.m[w] = m[w];
    if (h->P.wolves) reset_init_wolves(h->P, h->E, h->S, overflow);
    wolf_plane(h->E, h->S, wm);
    for (int w = 0; w < 4; ++w) bm[w] = h->E.m[w];
    return overflow;
}

void hostsim_reset(HostSim* h, uint8_t* grids, int32_t* food, int32_t* role, int32_t* status, int32_t* overflow) {
    uint32_t wm[4], bm[4];
    uint32_t ov = do_reset(h, wm, bm);
    expand_obs(h, wm, bm, h->E.role, grids);
    *food = (int32_t)food_observation(h->P, h->E, h->f64);
    *role = (int32_t)h->E.role; *status = (int32_t)h->E.status;
    if (overflow) *overflow = (int32_t)ov;
}

// One step with the kernel's semantics (auto-reset when configured: post-reset observation).
void hostsim_step(HostSim* h, int32_t action, uint8_t* grids, int32_t* food, int32_t* role, int32_t* status,
                  float* reward, int32_t* done, int32_t* info, int32_t* overflow) {
    StepOut O;
    memset(&O, 0, sizeof(O));
    const Coop<1> coop{0u, 1u};
    if (h->f64) env_step<true, 1>(h->P, h->E, h->S, (uint32_t)action, O, coop);
    else env_step<false, 1>(h->P, h->E, h->S, (uint32_t)action, O, coop);
    if (O.done && h->P.auto_reset) {
        O.overflow |= do_reset(h, O.wm, O.bm);
        O.food_obs = food_observation(h->P, h->E, h->f64);
        O.role = h->E.role; O.status = h->E.status;
    }
    expand_obs(h, O.wm, O.bm, O.role, grids);
    *food = (int32_t)O.food_obs; *role = (int32_t)O.role; *status = (int32_t)O.status;
    *reward = O.reward; *done = (int32_t)O.done; *info = (int32_t)O.info;
    if (overflow) *overflow = (int32_t)O.overflow;
}

// hidden state: x, y, food (as double), role, status, turn, episode, nw, nlog; wolves (2 ints each)
void hostsim_state(const HostSim* h, int32_t* scal9, double* food, int32_t* wolves_xy, uint32_t* bush_mask) {
    const Env& E = h->E;
    scal9[0] = E.x; scal9[1] = E.y; scal9[2] = E.food_i; scal9[3] = (int32_t)E.role; scal9[4] = (int32_t)E.status;
    scal9[5] = (int32_t)E.turn; scal9[6] = (int32_t)E.episode; scal9[7] = (int32_t)E.nw; scal9[8] = (int32_t)E.nlog;
    *food = h->f64 ? E.food_f : (double)E.food_i / h->P.food_obs_scale;
    for (uint32_t k = 0; k < E.nw; ++k) {
        wolves_xy[2 * k] = unpack_x(h->S.wolves[k]); wolves_xy[2 * k + 1] = unpack_y(h->S.wolves[k]);
    }
    for (int w = 0; w < 4; ++w) bush_mask[w] = E.m[w];
}

// PragmaticObsWrapper features from two observed 11x11 grids (row-major u8), wab_env.py:726-761
void hostsim_features(const uint8_t* wolf_grid, const uint8_t* bush_grid, int32_t food, int32_t role, int32_t status,
                      uint8_t* out28) {
    uint32_t wm[4] = {0, 0, 0, 0}, bm[4] = {0, 0, 0, 0};
    for (int c = 0; c < CELLS; ++c) {
        if (wolf_grid[c]) wm[c >> 5] |= 1u << (c & 31);
        if (bush_grid[c]) bm[c >> 5] |= 1u << (c & 31);
    }
    uint32_t f[7];
    pragmatic_features(wm, bm, (uint32_t)food, (uint32_t)role, (uint32_t)status, f);
    memcpy(out28, f, 28);
}

// Volume probe of the reveal paths (slide_window, reset_bush_block) against the draw contract evaluated cell by
// cell on full 32-bit words. A tie of a cell's high half-word with the threshold's (2^-16 per cell) takes a rare
// branch that whole-episode tests almost never reach; this loop reaches it hundreds of times.
// out3 = {cells whose high half-word tied, mismatching cells, cells checked}
static uint32_t exact_bush(const Params& P, uint32_t ka, uint32_t kb, int32_t wx, int32_t wy, uint32_t* tie) {
    const uint32_t c0 = pack_xy(wx >> 1, wy >> 1) ^ ka;
    uint32_t p[2], q[2];
    philox2(P, c0, kb, p);
    philox2(P, c0, ~kb, q);
    const uint32_t hw = (wy & 1) ? p[1] : p[0], lw = (wy & 1) ? q[1] : q[0];
    const uint32_t h = (wx & 1) ? (hw >> 16) : (hw & 0xFFFFu), l = (wx & 1) ? (lw >> 16) : (lw & 0xFFFFu);
    if (h == (P.thr_bush1 >> 16)) *tie += 1;
    return ((h << 16) | l) >= P.thr_bush1 ? 1u : 0u;
}

void hostsim_reveal_probe(uint64_t seed, uint32_t thr, int64_t iters, int64_t* out3) {
    Params P;
    memset(&P, 0, sizeof(P));
    uint32_t k2 = (uint32_t)(seed ^ (seed >> 32));
    for (int r = 0; r < 10; ++r) { P.rk2[r] = k2; k2 += PHILOX_W0; }
    P.thr_bush1 = thr; P.thr_bush2 = 0xFFFFFFFFu; P.n_bush_thr = 1;
    Slots S;
    memset(&S, 0, sizeof(S));
    uint64_t lcg = seed * 6364136223846793005ull + 1442695040888963407ull;
    auto next = [&lcg]() { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(lcg >> 32); };
    uint32_t ties = 0;
    int64_t bad = 0, checked = 0;
    const Coop<1> coop{0u, 1u};
    for (int64_t it = 0; it < iters; ++it) {
        Env E;
        memset(&E, 0, sizeof(E));
        E.bk_a = next(); E.bk_b = next();
        if (it % 8 == 7) {                       // a reset window around the origin
            uint32_t m[4] = {0, 0, 0, 0};
            for (int blk = 0; blk < 36; ++blk) reset_bush_block(P, E.bk_a, E.bk_b, blk, m);
            for (int i = 0; i < 11; ++i)
                for (int j = 0; j < 11; ++j) {
                    const uint32_t want = exact_bush(P, E.bk_a, E.bk_b, 5 - i, 5 - j, &ties);
                    const int b = 11 * i + j;
                    bad += ((m[b >> 5] >> (b & 31)) & 1u) != want;
                    ++checked;
                }
            bad += (m[3] >> 25) != 0u;
            continue;
        }
        E.x = (int32_t)(next() % 4001u) - 2000; E.y = (int32_t)(next() % 4001u) - 2000;
        const uint32_t d = next() & 3u;
        const int32_t dx = d == 0 ? 1 : d == 1 ? -1 : 0, dy = d == 2 ? 1 : d == 3 ? -1 : 0;
        slide_window<1>(P, E, S, dx, dy, coop);
        uint32_t seen[4] = {0, 0, 0, 0};
        for (int g = 0; g < 11; ++g) {
            const int32_t wx = dx ? E.x + HALF * dx : E.x + HALF - g, wy = dx ? E.y + HALF - g : E.y + HALF * dy;
            const uint32_t want = exact_bush(P, E.bk_a, E.bk_b, wx, wy, &ties);
            const int b = 11 * (5 - (wx - E.x)) + (5 - (wy - E.y));
            bad += ((E.m[b >> 5] >> (b & 31)) & 1u) != want;
            seen[b >> 5] |= 1u << (b & 31);
            ++checked;
        }
        for (int w = 0; w < 4; ++w) bad += (E.m[w] & ~seen[w]) != 0u;      // nothing outside the new line
    }
    out3[0] = ties; out3[1] = bad; out3[2] = checked;
}

void hostsim_philox2(const uint32_t* ctr, uint32_t key, uint32_t* out) {
    Params P;
    memset(&P, 0, sizeof(P));
    for (int r = 0; r < 10; ++r) { P.rk2[r] = key; key += PHILOX_W0; }
    philox2(P, ctr[0], ctr[1], out);
}

void hostsim_philox(const uint32_t* ctr, uint32_t k0, uint32_t k1, uint32_t* out) {
    Params P;
    memset(&P, 0, sizeof(P));
    fill_round_keys(P, k0, k1);
    philox(P, ctr[0], ctr[1], ctr[2], ctr[3], out);
}

}  // extern "C"

// ------------------------------------------------------------------ Environment 2.0 (wab2_core.cuh)
#include "../../wab_gym_b200/csrc/wab2_core.cuh"

struct HostSim2 {
    Params2 P;
    std::vector<uint32_t> words;
    World2 W;
};

extern "C" {

HostSim2* hostsim2_create(const Wab2Config* c, uint64_t seed, uint64_t env_id) {
    HostSim2* h = new HostSim2();
    Params2& P = h->P;
    memset(&P, 0, sizeof(P));
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { P.rk0[r] = k0; P.rk1[r] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
    P.width = c->width; P.height = c->height; P.n_ostriches = c->n_ostriches; P.n_wolves = c->n_wolves; P.n_bushes = c->n_bushes;
    P.n_entities = c->n_ostriches + c->n_wolves + c->n_bushes; P.n_acting = c->n_ostriches + c->n_wolves;
    P.lookout_r = c->lookout_view_radius; P.gatherer_r = c->gatherer_view_radius; P.wolf_r = c->wolf_view_radius;
    P.window_r = c->window_radius; P.starting_role = c->starting_role; P.ostrich_food = c->ostrich_starting_food;
    P.wolf_food = c->wolf_starting_food; P.wolf_eat_gain = c->wolf_food_for_eating_ostrich;
    P.bush_food = c->food_per_bush; P.bush_given = c->food_given_per_turn; P.env_id_base = 0;
    h->words.assign((size_t)3 * P.n_entities, 0u);                       // [2E] table rows and food, then [E] object coords
    h->W.base = h->words.data(); h->W.stride = 1; h->W.env_id = (uint32_t)env_id;
    h->W.obj = h->words.data() + 2 * P.n_entities; h->W.ostride = 1;
    world2_create(P, h->W);
    return h;
}
void hostsim2_destroy(HostSim2* h) { delete h; }
void hostsim2_reset(HostSim2* h) { world2_reset(h->P, h->W); }
int32_t hostsim2_observe(HostSim2* h, int32_t a, uint8_t* planes, int32_t* internal5) {
    const int S = 2 * h->P.window_r + 1, nbits = 3 * S * S;
    std::vector<uint32_t> bits((size_t)(nbits + 31) / 32 + 1, 0u);
    int32_t rows = world2_observe(h->P, h->W, a, bits.data(), 0, internal5);
    for (int b = 0; b < nbits; ++b) planes[b] = (uint8_t)((bits[b >> 5] >> (b & 31)) & 1u);
    return rows;
}
void hostsim2_act(HostSim2* h, int32_t a, int32_t action, float* reward, int32_t* done) {
    uint32_t d = 0;
    world2_act(h->P, h->W, a, (uint32_t)action, *reward, d);
    *done = (int32_t)d;
}
void hostsim2_state(const HostSim2* h, int32_t* out9, int32_t* turn) {
    for (int k = 0; k < h->P.n_entities; ++k) {
        const uint32_t obj = h->words[2 * h->P.n_entities + k], tab = h->words[2 * k], food = h->words[2 * k + 1];
        int32_t* o = out9 + 9 * k;
        o[0] = (int32_t)entity_type(h->P, k); o[1] = unpack_x(obj); o[2] = unpack_y(obj); o[3] = (int32_t)(tab & 0xFFu);
        o[4] = (int32_t)((tab >> 8) & 0xFFu); o[5] = (int32_t)((tab >> 16) & 1u); o[6] = (int32_t)food;
        o[7] = (int32_t)((tab >> 17) & 1u); o[8] = (int32_t)((tab >> 18) & 3u);
    }
    *turn = (int32_t)h->W.turn;
}

}  // extern "C"
