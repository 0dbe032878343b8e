// wab2_core.cuh — per-environment logic of the Environment 2.0 world turn ("/root/reference/Environment 2.0").
//
// One thread owns one world: its E entities live in a strided scratch array (shared memory in the kernel,
// a plain array in tests/hostsim) as three words each, and the entities act one after the other exactly as in
// the reference driver loop (Env2Tests.py:46-88): get_obs(i) -> take_action(i, a). Host/device portable like
// wab_core.cuh. Behaviour follows World.py:93-132 (default_game_update), :243-316 (_get_visible_objects),
// :325-334 (perform_entity_action), :346-377; Bush.py:31-39; WAB_Environment2.py:61-134;
// WAB_Environment2_Single.py:36-69 — including the bugs listed in SURVEY Appendix C.
#pragma once
#include "wab_core.cuh"

#ifndef WAB2_SCAN_UNROLL
#define WAB2_SCAN_UNROLL 4
#endif

namespace wab {
constexpr int kScanUnroll = WAB2_SCAN_UNROLL;

enum : uint32_t { SITE_V2_CREATE = 8, SITE_V2_RESET = 9, SITE_V2_PICK = 10 };
enum : uint32_t { T_OSTRICH = 0, T_WOLF = 1, T_BUSH = 2 };

struct Params2 {
    uint32_t rk0[10], rk1[10];
    int32_t width, height, n_ostriches, n_wolves, n_bushes, n_entities, n_acting;
    int32_t lookout_r, gatherer_r, wolf_r, window_r;       // WAB_Environment2.py:35-36, :49; obs window radius
    int32_t starting_role, ostrich_food, wolf_food, wolf_eat_gain, bush_food, bush_given;
    uint8_t halfwidth[3][16];      // [lookout | gatherer | wolf][|dx|] = floor(sqrt(r^2 - dx^2)): the circle of World.py:295-297
    uint64_t env_id_base;
};

// entity k of this world: three words — object coords (x:i16 | y:i16<<16, never wrapped, World.py:331-332), table row
// (X:8 | Y:8 | Visible:1 | role:1 | status:2), food (integer-valued). Table row and food are what every observation and
// co-location test scans: base[(2*k + f) * stride], f = 0 table row, 1 food (shared memory in the kernel). The object
// coords are touched once per action only and stay where the state lives: obj[k * ostride] (global memory in the
// kernel — a third less shared memory per world, which is what bounds the worlds resident per SM).
struct World2 {
    uint32_t* base; int32_t stride;
    uint32_t* obj; int64_t ostride;
    uint32_t env_id, episode, turn;
};
WAB_HD uint32_t& w2_obj(const World2& W, int k) { return W.obj[(int64_t)k * W.ostride]; }
WAB_HD uint32_t& w2_tab(const World2& W, int k) { return W.base[(2 * k + 0) * W.stride]; }
WAB_HD uint32_t& w2_food(const World2& W, int k) { return W.base[(2 * k + 1) * W.stride]; }
WAB_HD uint32_t tab_pack(uint32_t tx, uint32_t ty, uint32_t vis, uint32_t role, uint32_t status) {
    return tx | (ty << 8) | (vis << 16) | (role << 17) | (status << 18);
}
WAB_HD uint32_t entity_type(const Params2& P, int k) {
    return k < P.n_ostriches ? T_OSTRICH : (k < P.n_ostriches + P.n_wolves ? T_WOLF : T_BUSH);
}
WAB_HD void philox2(const Params2& P, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ P.rk0[r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ P.rk1[r];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// randint(low, high) inclusive, keyed (oracle/ref_shim/v2.py)
WAB_HD int32_t keyed_int(const Params2& P, uint32_t env_id, uint32_t episode, uint32_t site, uint32_t turn, uint32_t entity,
                         uint32_t axis, int32_t low, int32_t high) {
    uint32_t w[4];
    philox2(P, env_id, episode, ctr2(site, turn, axis), entity, w);
    return low + (int32_t)(((uint64_t)w[0] * (uint64_t)(uint32_t)(high - low + 1)) >> 32);
}
WAB_HD int32_t pymod(int32_t a, int32_t m) { int32_t r = a % m; return r < 0 ? r + m : r; }

// create_ostriches / create_wolves / create_bushes, WAB_Environment2.py:61-110
WAB_HD void world2_create(const Params2& P, World2& W) {
    W.episode = 0; W.turn = 0;
    WAB_ROLLED
    for (int k = 0; k < P.n_entities; ++k) {
        const uint32_t t = entity_type(P, k);
        const int32_t x = keyed_int(P, W.env_id, 0, SITE_V2_CREATE, 0, (uint32_t)k, 0, 0, P.width - 1);
        const int32_t y = keyed_int(P, W.env_id, 0, SITE_V2_CREATE, 0, (uint32_t)k, 1, 0, P.height - 1);
        w2_obj(W, k) = pack_xy(x, y);
        const int32_t food = t == T_OSTRICH ? P.ostrich_food : (t == T_WOLF ? P.wolf_food : P.bush_food);
        const uint32_t role = t == T_OSTRICH ? (uint32_t)P.starting_role : (t == T_BUSH ? (food > 0 ? 1u : 0u) : 0u);
        w2_tab(W, k) = tab_pack((uint32_t)x, (uint32_t)y, 1u, role, 0u);
        w2_food(W, k) = (uint32_t)food;
    }
}

// reset_environment, WAB_Environment2.py:113-118: respawn in [0, W] x [0, H] INCLUSIVE (Single.py:45-46), entity
// state back to its start, Visible = True; the table X / Y keep their stale values (World.py:353-356 is a no-op).
WAB_HD void world2_reset(const Params2& P, World2& W) {
    W.episode += 1; W.turn = 0;
    WAB_ROLLED
    for (int k = 0; k < P.n_entities; ++k) {
        const uint32_t t = entity_type(P, k);
        const int32_t x = keyed_int(P, W.env_id, W.episode, SITE_V2_RESET, 0, (uint32_t)k, 0, 0, P.width);
        const int32_t y = keyed_int(P, W.env_id, W.episode, SITE_V2_RESET, 0, (uint32_t)k, 1, 0, P.height);
        w2_obj(W, k) = pack_xy(x, y);
        const int32_t food = t == T_OSTRICH ? P.ostrich_food : (t == T_WOLF ? P.wolf_food : P.bush_food);
        const uint32_t role = t == T_OSTRICH ? (uint32_t)P.starting_role : (t == T_BUSH ? (food > 0 ? 1u : 0u) : 0u);
        const uint32_t tab = w2_tab(W, k);
        w2_tab(W, k) = tab_pack(tab & 0xFFu, (tab >> 8) & 0xFFu, 1u, role, 0u);
        w2_food(W, k) = (uint32_t)food;
    }
}

// wrap-aware delta along one axis, World.py:252-291: only ONE wrap direction is ever considered (if / elif), and
// min(d, wrap, key=abs) keeps d on ties.
WAB_HD int32_t axis_delta(int32_t obj, int32_t ent, int32_t r, int32_t size) {
    const int32_t d = obj - ent;
    const bool low = ent < r;                               // "if entity < radius" branch (World.py:255)
    const bool high = !low && size < ent + r;               // "elif size < entity + radius" branch (:264)
    const bool in_low = low && (size - (r - ent) <= obj);
    const bool in_high = high && (obj <= r - size + ent);
    const int32_t wrap = in_low ? d - size : d + size;      // -ent - (size - obj)  |  obj + size - ent
    const int32_t ad = d < 0 ? -d : d, aw = wrap < 0 ? -wrap : wrap;
    return ((in_low || in_high) && aw < ad) ? wrap : d;     // min(d, wrap, key=abs): ties keep d
}

WAB_HD void set_stream_bit(uint32_t* bits, int32_t pos) {   // streams of neighbouring lanes share boundary words
#if defined(__CUDA_ARCH__)
    atomicOr(bits + (pos >> 5), 1u << (pos & 31));
#else
    bits[pos >> 5] |= 1u << (pos & 31);
#endif
}

// get_observations(a), World.py:360-377: sets bit (bit0 + type*S*S + (dx+R)*S + (dy+R)) of `bits` (stride 1 words)
// for every listed object; internal5 = internal_obs (x, y, food, role | is_running, status). Returns the row count.
WAB_HD int32_t world2_observe(const Params2& P, const World2& W, int a, uint32_t* bits, int32_t bit0, int32_t internal5[5]) {
    const uint32_t atab = w2_tab(W, a);
    const uint32_t at = entity_type(P, a);
    const int32_t ax = (int32_t)(atab & 0xFFu), ay = (int32_t)((atab >> 8) & 0xFFu);
    int32_t r = 0;
    if (at == T_OSTRICH) r = ((atab >> 17) & 1u) ? P.gatherer_r : P.lookout_r;
    else if (at == T_WOLF) r = P.wolf_r;
    const int32_t R = P.window_r, S = 2 * R + 1;
    int32_t rows = 0;
    // independent iterations: partial unrolling overlaps their shared-memory loads and compare chains (the kernel is
    // latency-bound at one thread per world: 65,536 worlds are 14 warps per SM). Measured on config 3: unroll 1 ->
    // 2.12e8, 4 -> 2.31e8, 8 -> 2.34e8 world turns/s; fetching the rows ahead by hand gave nothing more.
#if defined(__CUDA_ARCH__)
#pragma unroll kScanUnroll
#endif
    for (int k = 0; k < P.n_entities; ++k) {
        const uint32_t tab = w2_tab(W, k);
        const int32_t dx = axis_delta((int32_t)(tab & 0xFFu), ax, r, P.width);
        const int32_t dy = axis_delta((int32_t)((tab >> 8) & 0xFFu), ay, r, P.height);
        if (dx * dx + dy * dy > r * r) continue;             // :295-297
        if (!((tab >> 16) & 1u)) continue;                   // :300
        ++rows;
        if (dx >= -R && dx <= R && dy >= -R && dy <= R) {
            const int32_t pos = bit0 + ((int32_t)entity_type(P, k) * S + (dx + R)) * S + (dy + R);
            set_stream_bit(bits, pos);
        }
    }
    const uint32_t obj = w2_obj(W, a);
    internal5[0] = unpack_x(obj); internal5[1] = unpack_y(obj); internal5[2] = (int32_t)w2_food(W, a);
    internal5[3] = at == T_BUSH ? 0 : (int32_t)((atab >> 17) & 1u);
    internal5[4] = at == T_BUSH ? 0 : (int32_t)((atab >> 18) & 3u);
    return rows;
}

// take_action(a, action): entity act (World.py:25-43, :61-73), table update (:331-332), default_game_update (:93-132),
// reward (:54-58, :84-85, :21-22) and done (Ostrich / Wolf / Bush .is_done). The turn counter advances after the last
// entity (WAB_Environment2.py:131-133).
WAB_HD void world2_act(const Params2& P, World2& W, int a, uint32_t action, float& reward, uint32_t& done) {
    const uint32_t t = entity_type(P, a);
    uint32_t obj = w2_obj(W, a), tab = w2_tab(W, a);
    int32_t x = unpack_x(obj), y = unpack_y(obj);
    uint32_t role = (tab >> 17) & 1u;
    if (t != T_BUSH) {
        if (action == 0u) y += 1; else if (action == 1u) x += 1; else if (action == 2u) y -= 1; else if (action == 3u) x -= 1;
        else if (t == T_OSTRICH && action == 4u) role = 0u; else if (t == T_OSTRICH && action == 5u) role = 1u;
    }
    const uint32_t tx = (uint32_t)pymod(x, P.width), ty = (uint32_t)pymod(y, P.height);
    w2_obj(W, a) = pack_xy(x, y);
    tab = tab_pack(tx, ty, (tab >> 16) & 1u, role, (tab >> 18) & 3u);
    w2_tab(W, a) = tab;
    if (t != T_BUSH) {
        const uint32_t want = t == T_WOLF ? T_OSTRICH : T_BUSH;
        const int lo = want == T_OSTRICH ? 0 : P.n_ostriches + P.n_wolves;
        const int hi = want == T_OSTRICH ? P.n_ostriches : P.n_entities;
        const uint32_t cell = (tab & 0xFFFFu) | (1u << 16);              // same X, Y and Visible
        int32_t k = 0;
        WAB_ROLLED
        for (int q = lo; q < hi; ++q) k += ((w2_tab(W, q) & 0x1FFFFu) == cell) ? 1 : 0;
        if (k > 0) {
            const int32_t j = keyed_int(P, W.env_id, W.episode, SITE_V2_PICK, W.turn, (uint32_t)a, 0, 0, k - 1);
            int32_t seen = 0, pick = -1;
            WAB_ROLLED
            for (int q = lo; q < hi && pick < 0; ++q)
                if ((w2_tab(W, q) & 0x1FFFFu) == cell && seen++ == j) pick = q;
            if (t == T_WOLF) {
                w2_food(W, a) += (uint32_t)P.wolf_eat_gain;                                   // :113
                w2_tab(W, pick) = (w2_tab(W, pick) & ~(3u << 18)) | (2u << 18);               // :114 killed
                w2_tab(W, j) &= ~(1u << 16);                                                  // :115 hides LABEL j
            } else {
                int32_t bf = (int32_t)w2_food(W, pick), got;                                   // Bush.take_food
                if (bf >= P.bush_given) { bf -= P.bush_given; got = P.bush_given; }
                else { got = bf; bf = 0; w2_tab(W, pick) &= ~(1u << 17); }
                w2_food(W, pick) = (uint32_t)bf;
                w2_food(W, a) += (uint32_t)got;                                               // :127
            }
        }
    }
    const uint32_t now = w2_tab(W, a);
    if (t == T_OSTRICH) { const uint32_t st = (now >> 18) & 3u; reward = st == 0u ? 1.f : 0.f; done = st != 0u; }
    else if (t == T_WOLF) { reward = (int32_t)w2_food(W, a) > 10 ? 1.f : 0.f; done = (((now >> 18) & 3u) == 1u); }
    else { reward = 0.f; done = 1u; }
    if (a == P.n_entities - 1) W.turn += 1;
}

}  // namespace wab
