// wab2_grid.cuh — Environment 2.0 world turn, one WARP per world (for worlds with 2R+1 <= W, H <= 64).
//
// The thread-per-world kernel (wab2_kernels.cuh) walks every entity for every observation: fine for 33
// entities, hopeless for the 64x64 world of BASELINE config 4 (328 entities, 72 observations per turn). Here the
// world keeps, in shared memory, one occupancy bit plane per entity type (column x = 64 bits over y; a bit is
// set iff a VISIBLE entity of that type has that table position), so an observation is 3 x (2R+1) rotated and
// circle-masked column reads instead of an entity scan, a co-location test is one bit test (plus a ballot
// over the candidates when it hits), and a move is two bit updates (with a ballot recount of the vacated cell).
//
// A turn has three parts:
//   1. PRE-PASS, one lane per entity: everything an entity's action does to ITSELF depends on nothing any other
//      entity does in the same turn (decode, move, role; World.py:25-43, :61-73, :331-332), and neither do its
//      internal observation (World.py:50-51, :80-81) nor an ostrich's reward / done (ostriches act before any
//      wolf can kill them in the turn: ids are ostriches, wolves, bushes). So they are computed 32 entities at a
//      time; the new table row waits in `newtab` until the entity's place in the order.
//   2. ORDERED LOOP over the entities (the only sequential part): entity a's window columns are read from the
//      planes as they are right before it acts (get_obs(a) then take_action(a), Env2Tests.py:46-88) into a
//      staging slot; then its pending row is committed, the planes follow, and default_game_update
//      (World.py:93-132) runs when the plane says somebody is there.
//   3. FLUSH, every kGridGroup observations: the staged column words of the group are turned into the u8 windows
//      by all 32 lanes at once — lane -> (observation, 16-byte chunk), 16 bits from two neighbouring column
//      words, byte expansion through the table, one streaming 16-byte store — and the < 16 ragged bytes at
//      either end of a window as single bytes. A wolf's reward (its own food after its own action) is written
//      by a post-pass, one lane per wolf.
//
// Semantics are those of wab2_core.cuh / the reference (World.py:93-132, :243-316, :325-377), including the
// one visible difference between the reference's wrap rule and a true torus when every window fits the world:
// the strict test `size < entity + radius` (World.py:264, :285) makes the single cell at delta = +radius
// invisible when entity + radius == size.
#pragma once

namespace {

constexpr int kGridGroup = 8;   // observations staged between two flushes

struct GridGeom {        // shared-memory layout of one warp's world, in 32-bit words
    int tab;             // [E] table rows, then [E] food (contiguous, like the state in global memory)
    int newtab;          // [A] pending rows of the pre-pass
    int cols;            // [3][W] u64 occupancy columns (2 words each), then [2][W] u64 "more than one" columns (ostriches, wolves)
    int stage;           // [kGridGroup][stage_words] window columns of the observations waiting for the flush
    int total;
};
__host__ __device__ inline int grid_stage_words(int S) { return 3 * S + 16 / S + 2; }   // zero words behind the last column
__host__ __device__ inline GridGeom grid_geom(int E, int A, int W, int S) {
    GridGeom g;
    g.tab = 0;
    g.newtab = 2 * E;
    g.cols = (2 * E + A + 1) & ~1;
    g.stage = g.cols + 5 * W * 2;
    g.total = (g.stage + kGridGroup * grid_stage_words(S) + 2 + 1) & ~1;   // + 2: the flush reads one column pair past the last slot
    return g;
}

__device__ __forceinline__ uint64_t rot_window(uint64_t col, int s, int H) {   // bits (s + k) mod H of col at position k
    const uint64_t m = (1ull << H) - 1ull;
    return s ? ((col >> s) | (col << (H - s))) & m : col;
}
__device__ __forceinline__ int col_word(int W, uint32_t plane, uint32_t x, uint32_t y) { return (int)(((plane * W + x) << 1) + (y >> 5)); }
// id range of an entity type
__device__ __forceinline__ void type_range(const Params2& P, uint32_t type, int& lo, int& hi) {
    lo = type == T_OSTRICH ? 0 : (type == T_WOLF ? P.n_ostriches : P.n_ostriches + P.n_wolves);
    hi = type == T_OSTRICH ? P.n_ostriches : (type == T_WOLF ? P.n_ostriches + P.n_wolves : P.n_entities);
}
// A visible entity of `type` arrived at (x, y). Planes 3, 4 say "possibly more than one here" for ostriches and wolves
// (sticky until a recount): leaving a cell whose bit is clear needs no scan of the others. Every lane calls with the
// same arguments; lane 0 is the only writer; ends with a warp barrier.
__device__ __forceinline__ void plane_arrive(uint32_t* cols, int W, uint32_t type, uint32_t x, uint32_t y, int lane) {
    const int w = col_word(W, type, x, y);
    const uint32_t bit = 1u << (y & 31), old = cols[w];
    __syncwarp();
    if (lane == 0) {
        cols[w] = old | bit;
        if ((old & bit) && type != T_BUSH) cols[col_word(W, 3u + type, x, y)] |= bit;
    }
    __syncwarp();
}
// A visible entity of `type` left `cell` = X | Y << 8 (or became invisible there); tab[] is already up to date.
__device__ __forceinline__ void plane_leave(const Params2& P, const uint32_t* tab, uint32_t* cols, uint32_t type, uint32_t cell,
                                            int lane) {
    const int W = P.width;
    const uint32_t x = cell & 0xFFu, y = (cell >> 8) & 0xFFu, bit = 1u << (y & 31);
    const int w = col_word(W, type, x, y), wm = col_word(W, type == T_BUSH ? 0u : 3u + type, x, y);
    const bool shared_cell = type == T_BUSH || (cols[wm] & bit);       // bushes only ever "move" in the first turn: always count
    int left = 0;
    if (shared_cell) {
        int lo, hi;
        type_range(P, type, lo, hi);
        const uint32_t want = cell | (1u << 16);
        for (int base = lo; base < hi; base += 32) {
            const int k = base + lane;
            left += __popc(__ballot_sync(FULL, k < hi && (tab[k] & 0x1FFFFu) == want));
        }
    }
    __syncwarp();
    if (lane == 0) {
        if (left == 0) cols[w] &= ~bit;
        if (shared_cell && type != T_BUSH && left < 2) cols[wm] &= ~bit;
    }
    __syncwarp();
}

// Shared memory by 32-bit shared-window address: the ordered loop and the flush touch it on almost every instruction, and
// with generic pointers the compiler rebuilds the window base (S2R SR_CgaCtaId, LEA) in front of most accesses.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// Windows of the `g` observations staged at shared address `s_stage` (slot e = entity a0 + e) -> `dst`, the first byte of
// entity a0's window. The outputs of this kernel are WORLD-major, so the g windows are one run of g * obs_bytes bytes:
// bit p = (type*S + dxi)*S + dyi of window e is byte e * obs_bytes + p of the run, which starts at ANY byte address —
// whole 16-byte chunks go out as one streaming store per lane, the < 16 bytes at either end of the run byte by byte.
// inv_s = ceil(2^20 / S): (p * inv_s) >> 20 == p / S for p < 3 S^2 (p * S < 2^20); inv_ob = ceil(2^32 / obs_bytes):
// __umulhi(b, inv_ob) == b / obs_bytes for b < kGridGroup * obs_bytes (b * obs_bytes < 2^32).
template <bool NARROW>
__device__ __forceinline__ uint32_t window_bits16(uint32_t s_col, int S, uint32_t r) {   // 16 bits from bit r of column word s_col on
    uint32_t h = (lds32(s_col) >> r) | (lds32(s_col + 4u) << (S - r));
    if (NARROW)                                                      // windows narrower than 15: more than two columns in 16 bits
        for (int filled = 2 * S - (int)r, k = 2; filled < 16; filled += S, ++k) h |= lds32(s_col + 4u * k) << filled;
    return h;
}
template <bool NARROW>
__device__ __forceinline__ void grid_flush_group(uint32_t s_stage, uint32_t s_lut, uint8_t* dst, int g, int S, int SW,
                                                 uint32_t inv_s, uint32_t inv_ob, int obs_bytes, int lane) {
    const int total = g * obs_bytes;
    const int off = (int)(reinterpret_cast<uintptr_t>(dst) & 15);
    const int head = off ? min(16 - off, total) : 0;             // bytes before the first whole chunk
    const int n_chunks = (total - head) >> 4;
    uint4* out16 = reinterpret_cast<uint4*>(dst + head) + lane;
    // this lane's chunk: byte head + 16 c of the run = bit p of the window in slot s_slot; 32 chunks = 512 bytes further per round
    uint32_t p = (uint32_t)(head + (lane << 4));
    const uint32_t e0 = __umulhi(p, inv_ob);
    p -= e0 * (uint32_t)obs_bytes;
    uint32_t s_slot = s_stage + e0 * (uint32_t)(SW * 4);
#pragma unroll 1
    for (int c = lane; c < n_chunks; c += 32, out16 += 32) {
        const uint32_t q = (p * inv_s) >> 20, r = p - q * (uint32_t)S;
        uint32_t h = window_bits16<NARROW>(s_slot + q * 4u, S, r);
        const int have = obs_bytes - (int)p;                     // the zero words behind a slot's last column end its bits
        if (have < 16) h |= window_bits16<NARROW>(s_slot + (uint32_t)(SW * 4), S, 0u) << have;
        const uint2 lo = lds64(s_lut + ((h & 0xFFu) << 3)), hi = lds64(s_lut + ((h >> 5) & 0x7F8u));
        __stcs(out16, make_uint4(lo.x, lo.y, hi.x, hi.y));
        p += 512u;
        while (p >= (uint32_t)obs_bytes) { p -= (uint32_t)obs_bytes; s_slot += (uint32_t)(SW * 4); }
    }
    const int tail0 = head + (n_chunks << 4);
    const int pb = lane < 16 ? lane : tail0 + lane - 16;             // ragged head (lanes 0-15) and tail (lanes 16-31)
    if (lane < 16 ? pb < head : pb < total) {
        const uint32_t e = __umulhi((uint32_t)pb, inv_ob), pw = (uint32_t)pb - e * (uint32_t)obs_bytes;
        const uint32_t q = (pw * inv_s) >> 20, r = pw - q * (uint32_t)S;
        dst[pb] = (uint8_t)((lds32(s_stage + (e * (uint32_t)SW + q) * 4u) >> r) & 1u);
    }
}

// One world turn, one warp per world. H64: the world is 64 high (a window column is one funnel shift of the plane column);
// NARROW: the window is narrower than 15 cells.
template <bool H64, bool NARROW>
__global__ void __launch_bounds__(128) wab2_grid_turn_kernel(const __grid_constant__ Params2 P, const State2Ptrs st,
                                                             const uint8_t* __restrict__ actions, const Out2Ptrs out) {
    extern __shared__ uint32_t smem2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int E = P.n_entities, A = P.n_acting, nO = P.n_ostriches, W = P.width, H = H64 ? 64 : P.height;
    const int R = P.window_r, S = 2 * R + 1, obs_bytes = 3 * S * S, SW = grid_stage_words(S);
    const GridGeom g = grid_geom(E, A, W, S);
    uint2* lut = reinterpret_cast<uint2*>(smem2 + wpb * g.total);
    build_lut(lut);
    const int64_t idx = (int64_t)blockIdx.x * wpb + warp, n = st.n;
    if (idx >= n) return;                                   // whole warp
    uint32_t* tab = smem2 + warp * g.total + g.tab;
    uint32_t* food = tab + E;
    uint32_t* newtab = smem2 + warp * g.total + g.newtab;
    uint32_t* cols = smem2 + warp * g.total + g.cols;
    uint32_t* stage = smem2 + warp * g.total + g.stage;
    const uint32_t s_tab = (uint32_t)__cvta_generic_to_shared(tab), s_new = (uint32_t)__cvta_generic_to_shared(newtab),
                   s_cols = (uint32_t)__cvta_generic_to_shared(cols), s_stage = (uint32_t)__cvta_generic_to_shared(stage),
                   s_lut = (uint32_t)__cvta_generic_to_shared(lut);
    uint32_t* gent = st.ent + idx * st.stride_world;        // this world: [obj | table row | food][E], planar
    for (int k = lane; k < 2 * E; k += 32) tab[k] = gent[E + k];
    for (int k = lane; k < 5 * W * 2; k += 32) cols[k] = 0u;
    for (int k = lane; k < kGridGroup * SW + 2; k += 32) stage[k] = 0u;   // the words behind the last column stay zero
    const uint32_t env_id = (uint32_t)(P.env_id_base + (uint64_t)idx), episode = st.episode[idx];
    const uint32_t turn = st.turn[idx];
    // this lane's column of the window: circle mask per radius kind (lookout, gatherer, wolf).
    // |dy| <= m  <=>  dx^2 + dy^2 <= r^2 (World.py:295-297): bits R-m .. R+m of the 2R+1 window column; 0 beyond the radius
    const int rad[3] = {P.lookout_r, P.gatherer_r, P.wolf_r};
    uint32_t cmask[3];
#pragma unroll
    for (int rs = 0; rs < 3; ++rs) {
        const int adx = lane - R < 0 ? R - lane : lane - R;
        cmask[rs] = 0u;
        if (lane < S && adx <= rad[rs]) { const int m = (int)P.halfwidth[rs][adx]; cmask[rs] = ((2u << (2 * m)) - 1u) << (R - m); }
    }
    __syncwarp();
    for (int k = lane; k < E; k += 32) {
        const uint32_t t = tab[k];
        if ((t >> 16) & 1u) {
            const uint32_t ty = entity_type(P, k), x = t & 0xFFu, y = (t >> 8) & 0xFFu, bit = 1u << (y & 31);
            const uint32_t old = atomicOr(cols + col_word(W, ty, x, y), bit);
            if ((old & bit) && ty != T_BUSH) atomicOr(cols + col_word(W, 3u + ty, x, y), bit);
        }
    }
    // world-major outputs: entity e of this world
    int32_t* o_internal = out.internal ? out.internal + idx * A * 5 : nullptr;
    float* o_reward = out.reward + idx * A;
    uint8_t* o_done = out.done + idx * A;

    // ---- 1. pre-pass: one lane per acting entity
    for (int e = lane; e < A; e += 32) {
        const bool ostrich = e < nO;
        const uint32_t ob = gent[e], t = tab[e];
        int32_t x = unpack_x(ob), y = unpack_y(ob);
        uint32_t role = (t >> 17) & 1u;
        if (o_internal) {                                             // internal_obs, World.py:50-51, :80-81
            int32_t* dst = o_internal + e * 5;
            dst[0] = x; dst[1] = y; dst[2] = (int32_t)food[e]; dst[3] = (int32_t)role; dst[4] = (int32_t)((t >> 18) & 3u);
        }
        if (ostrich) {                                                // compute_reward / is_done: status cannot change before it acts
            const uint32_t s2 = (t >> 18) & 3u;
            o_reward[e] = s2 == 0u ? 1.f : 0.f;
            o_done[e] = (uint8_t)(s2 != 0u);
        }
        const uint32_t action = (uint32_t)actions[(int64_t)e * n + idx];   // act, World.py:25-43, :61-73
        if (action == 0u) y += 1; else if (action == 1u) x += 1; else if (action == 2u) y -= 1; else if (action == 3u) x -= 1;
        else if (ostrich && action == 4u) role = 0u; else if (ostrich && action == 5u) role = 1u;
        const uint32_t nob = pack_xy(x, y);
        if (nob != ob) gent[e] = nob;
        uint32_t tx, ty;
        if (turn == 0u) {          // tables may be stale after reset_world: the full wrap of the object coordinates
            tx = (uint32_t)pymod(x, W); ty = (uint32_t)pymod(y, H);
        } else {                   // afterwards the table follows the object one cell at a time
            int nx = (int)(t & 0xFFu) + (x - unpack_x(ob)), ny = (int)((t >> 8) & 0xFFu) + (y - unpack_y(ob));
            nx += nx < 0 ? W : 0; nx -= nx >= W ? W : 0;
            ny += ny < 0 ? H : 0; ny -= ny >= H ? H : 0;
            tx = (uint32_t)nx; ty = (uint32_t)ny;
        }
        newtab[e] = tab_pack(tx, ty, 0u, role, 0u);
    }
    __syncwarp();

    // ---- 2. the acting entities in order: ostriches, then wolves (the type is a compile-time constant of the loop body)
    const uint32_t inv_s = ((1u << 20) + (uint32_t)S - 1u) / (uint32_t)S;
    const uint32_t inv_ob = 0xFFFFFFFFu / (uint32_t)obs_bytes + 1u;   // ceil(2^32 / obs_bytes)
    uint8_t* o_planes = out.planes ? out.planes + idx * A * obs_bytes : nullptr;
    const bool observe = o_planes != nullptr;
    const int dx = lane - R;
    const uint32_t plane_bytes = (uint32_t)W * 8u;                    // one occupancy plane
    bool bush_dirty = turn == 0u;
    auto entities = [&](auto type_tag, const int g0, const int a_begin, const int a_end) {   // g0: first entity of the staged group
        constexpr uint32_t AT = decltype(type_tag)::value;
#pragma unroll 1
        for (int a = a_begin; a < a_end; ++a) {
            const uint32_t ot = lds32(s_tab + 4u * a);
            const int ox = (int)(ot & 0xFFu), oy = (int)((ot >> 8) & 0xFFu);
            if (observe && lane < S) {                               // get_observations(a), World.py:360-377: one lane per column
                const bool gatherer = AT == T_OSTRICH && ((ot >> 17) & 1u);   // offset — the three type planes share x, rotation,
                const int r = AT == T_WOLF ? rad[2] : (gatherer ? rad[1] : rad[0]);          // circle mask and quirks
                uint32_t mask = AT == T_WOLF ? cmask[2] : (gatherer ? cmask[1] : cmask[0]);
                // no division: 0 <= ox < W, 0 <= oy < H and R < W, H, so one conditional add wraps
                const int sy = H64 ? ((oy - R) & 63) : (oy - R + (oy < R ? H : 0));
                int x = ox + dx;
                x += x < 0 ? W : 0;
                x -= x >= W ? W : 0;
                if (dx == r && ox + r == W) mask = 0u;                // World.py:264 strict test: this image is missed
                if (dx == 0 && oy + r == H) mask &= ~(1u << (R + r)); // World.py:285, same on the y axis
                const uint32_t s_col = s_cols + ((uint32_t)x << 3);
                const uint32_t s_out = s_stage + (uint32_t)((a - g0) * SW + lane) * 4u;
#pragma unroll
                for (int type_p = 0; type_p < 3; ++type_p) {
                    const uint2 cw = lds64(s_col + type_p * plane_bytes);
                    uint32_t bits;
                    if (H64) {                                        // bits (sy + k) mod 64, k < 32: one funnel shift
                        const uint32_t lo = sy & 32 ? cw.y : cw.x, hi = sy & 32 ? cw.x : cw.y;
                        bits = __funnelshift_r(lo, hi, (uint32_t)sy & 31u);
                    } else {
                        bits = (uint32_t)rot_window((uint64_t)cw.x | ((uint64_t)cw.y << 32), sy, H);
                    }
                    sts32(s_out + type_p * (uint32_t)(S * 4), bits & mask);
                }
            }
            // ---- take_action(a): the pending row becomes the table row (World.py:331-332); Visible and status are
            // whatever the others made of them meanwhile
            const uint32_t nt = lds32(s_new + 4u * a);
            constexpr uint32_t own = AT == T_OSTRICH ? 0x2FFFFu : 0xFFFFu;   // X, Y (+ an ostrich's role)
            const uint32_t merged = (ot & ~own) | (nt & own);
            const uint32_t tx = merged & 0xFFu, ty = (merged >> 8) & 0xFFu;
            const uint32_t s_cell = s_cols + (tx << 3) + ((ty >> 5) << 2);   // the new cell's word in plane 0
            __syncwarp();                                             // the window columns were read from the planes as they were
            if (merged != ot) {
                if (lane == 0) sts32(s_tab + 4u * a, merged);
                if (((ot >> 16) & 1u) && ((ot ^ merged) & 0xFFFFu)) {
                    const uint32_t bit_n = 1u << (ty & 31u), bit_o = 1u << ((uint32_t)oy & 31u);
                    const uint32_t s_n = s_cell + AT * plane_bytes;
                    const uint32_t s_o = s_cols + AT * plane_bytes + ((uint32_t)ox << 3) + (((uint32_t)oy >> 5) << 2);
                    if (!(lds32(s_o + 3u * plane_bytes) & bit_o)) {   // nobody else was in the old cell: no scan
                        if (lane == 0) {
                            const uint32_t on = lds32(s_n);
                            sts32(s_n, on | bit_n);
                            if (on & bit_n) sts32(s_n + 3u * plane_bytes, lds32(s_n + 3u * plane_bytes) | bit_n);
                            sts32(s_o, lds32(s_o) & ~bit_o);
                        }
                        __syncwarp();
                    } else {
                        plane_arrive(cols, W, AT, tx, ty, lane);      // (its barrier also publishes tab[a])
                        plane_leave(P, tab, cols, AT, ot & 0xFFFFu, lane);
                    }
                } else {
                    __syncwarp();
                }
            }
            // ---- default_game_update (World.py:93-132)
            constexpr uint32_t want = AT == T_WOLF ? T_OSTRICH : T_BUSH;
            if ((lds32(s_cell + want * plane_bytes) >> (ty & 31u)) & 1u) {
                int lo, hi;
                type_range(P, want, lo, hi);
                const uint32_t cellv = (merged & 0xFFFFu) | (1u << 16);
                int k = 0;
                for (int base = lo; base < hi; base += 32) {
                    const int q = base + lane;
                    k += __popc(__ballot_sync(FULL, q < hi && (tab[q] & 0x1FFFFu) == cellv));
                }
                if (k > 0) {
                    const int32_t j = keyed_int(P, env_id, episode, SITE_V2_PICK, turn, (uint32_t)a, 0, 0, k - 1);
                    int pick = -1, seen = 0;
                    for (int base = lo; base < hi && pick < 0; base += 32) {
                        const int q = base + lane;
                        const unsigned b = __ballot_sync(FULL, q < hi && (tab[q] & 0x1FFFFu) == cellv);
                        const int c = __popc(b);
                        if (j < seen + c) pick = base + (int)__fns(b, 0, j - seen + 1);
                        seen += c;
                    }
                    __syncwarp();
                    if (AT == T_WOLF) {
                        const uint32_t jt = tab[j];
                        __syncwarp();              // every lane holds the old row of label j before lane 0 rewrites it
                        if (lane == 0) {
                            food[a] += (uint32_t)P.wolf_eat_gain;                                  // :113
                            tab[pick] = (tab[pick] & ~(3u << 18)) | (2u << 18);                    // :114 killed
                            tab[j] = tab[j] & ~(1u << 16);                                         // :115 hides LABEL j
                        }
                        __syncwarp();
                        if ((jt >> 16) & 1u) plane_leave(P, tab, cols, entity_type(P, j), jt & 0xFFFFu, lane);
                    } else {
                        if (lane == 0) {                                                           // Bush.take_food
                            int32_t bf = (int32_t)food[pick], got;
                            if (bf >= P.bush_given) { bf -= P.bush_given; got = P.bush_given; }
                            else { got = bf; bf = 0; tab[pick] &= ~(1u << 17); }
                            food[pick] = (uint32_t)bf;
                            food[a] += (uint32_t)got;                                              // :127
                        }
                        bush_dirty = true;
                        __syncwarp();
                    }
                }
            }
        }
    };
    for (int g0 = 0; g0 < A; g0 += kGridGroup) {
        const int g1 = min(g0 + kGridGroup, A);
        entities(std::integral_constant<uint32_t, T_OSTRICH>(), g0, g0, min(g1, nO));
        entities(std::integral_constant<uint32_t, T_WOLF>(), g0, max(g0, nO), g1);
        if (observe) {                                                // ---- 3. a full group (or the last observers): windows out
            __syncwarp();
            grid_flush_group<NARROW>(s_stage, s_lut, o_planes + (int64_t)g0 * obs_bytes, g1 - g0, S, SW, inv_s, inv_ob, obs_bytes, lane);
            __syncwarp();
        }
    }
    // bushes never move: their "action" only refreshes a table position left stale by reset_world (World.py:353-356),
    // i.e. it is a no-op except in the first turn of an episode
    if (turn == 0u) {
        uint32_t bush_ob = 0u;
        for (int a = A; a < E; ++a) {
            if (((a - A) & 31) == 0) bush_ob = a + lane < E ? gent[a + lane] : 0u;      // 32 bushes' coordinates per load
            const uint32_t ob = __shfl_sync(FULL, bush_ob, (a - A) & 31), ot = tab[a];
            const uint32_t merged = (ot & ~0xFFFFu) | (uint32_t)pymod(unpack_x(ob), W) | ((uint32_t)pymod(unpack_y(ob), H) << 8);
            __syncwarp();
            if (merged != ot) {
                if (lane == 0) tab[a] = merged;
                if ((ot >> 16) & 1u) {
                    plane_arrive(cols, W, T_BUSH, merged & 0xFFu, (merged >> 8) & 0xFFu, lane);
                    plane_leave(P, tab, cols, T_BUSH, ot & 0xFFFFu, lane);
                } else {
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    // a wolf's reward is its own food after its own action (World.py:84-85); nobody else touches it
    for (int e = nO + lane; e < A; e += 32) {
        o_reward[e] = (int32_t)food[e] > 10 ? 1.f : 0.f;
        o_done[e] = (uint8_t)(((tab[e] >> 18) & 3u) == 1u);
    }
    // table rows and food back; the bushes' only if a bush was eaten from (or moved, in the first turn)
    for (int k = lane; k < 2 * E; k += 32) {
        const int e = k < E ? k : k - E;
        if (e < A || bush_dirty) gent[E + k] = tab[k];
    }
    if (lane == 0) st.turn[idx] = turn + 1u;
}

}  // namespace
