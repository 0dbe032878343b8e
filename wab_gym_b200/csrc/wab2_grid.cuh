// wab2_grid.cuh — Environment 2.0 world turn, one WARP per world (for worlds with 2R+1 <= W, H <= 64).
//
// The thread-per-world kernel (wab2_kernels.cuh) walks every entity for every observation: fine for 33
// entities, hopeless for the 64x64 world of BASELINE config 4 (328 entities, 72 observers per turn). Here the
// world keeps, in shared memory, one occupancy bit plane per entity type (column x = 64 bits over y; a bit is
// set iff a VISIBLE entity of that type has that table position), so an observation is 3 x (2R+1) rotated and
// circle-masked column reads instead of an entity scan, a co-location test is one bit test (plus a ballot
// over the candidates when it hits), and a move is two bit updates (with a ballot recount of the vacated cell).
// Semantics are those of wab2_core.cuh / the reference (World.py:93-132, :243-316, :325-377), including the
// one visible difference between the reference's wrap rule and a true torus when every window fits the world:
// the strict test `size < entity + radius` (World.py:264, :285) makes the single cell at delta = +radius
// invisible when entity + radius == size.
#pragma once

namespace {

struct GridGeom {        // shared-memory layout of one warp's world, in 32-bit words
    int ent;             // [3][E]   obj, tab, food
    int cols;            // [3][W]   u64 occupancy columns (2 words each)
    int stream;          // observation bit stream
    int total;
};
__host__ __device__ inline GridGeom grid_geom(int E, int W, int stream_words) {
    GridGeom g;
    g.ent = 0;
    g.cols = (3 * E + 1) & ~1;
    g.stream = g.cols + 3 * W * 2;
    g.total = (g.stream + stream_words + 1) & ~1;
    return g;
}

__device__ __forceinline__ uint64_t rot_window(uint64_t col, int s, int H) {   // bits (s + k) mod H of col at position k
    if (H == 64) return s ? (col >> s) | (col << (64 - s)) : col;
    const uint64_t m = (1ull << H) - 1ull;
    return s ? ((col >> s) | (col << (H - s))) & m : col;
}
__device__ __forceinline__ void col_set(uint32_t* cols, int W, uint32_t type, uint32_t x, uint32_t y) {
    atomicOr(cols + ((type * W + x) << 1) + (y >> 5), 1u << (y & 31));
}
__device__ __forceinline__ void col_clear(uint32_t* cols, int W, uint32_t type, uint32_t x, uint32_t y) {
    atomicAnd(cols + ((type * W + x) << 1) + (y >> 5), ~(1u << (y & 31)));
}
__device__ __forceinline__ bool col_test(const uint32_t* cols, int W, uint32_t type, uint32_t x, uint32_t y) {
    return (cols[((type * W + x) << 1) + (y >> 5)] >> (y & 31)) & 1u;
}
// id range of an entity type
__device__ __forceinline__ void type_range(const Params2& P, uint32_t type, int& lo, int& hi) {
    lo = type == T_OSTRICH ? 0 : (type == T_WOLF ? P.n_ostriches : P.n_ostriches + P.n_wolves);
    hi = type == T_OSTRICH ? P.n_ostriches : (type == T_WOLF ? P.n_ostriches + P.n_wolves : P.n_entities);
}
// After an entity of `type` left (or became invisible at) cell `cell` = X | Y << 8: clear the plane bit unless
// another visible entity of that type is still there. All lanes call; tab[] must be up to date and synced.
__device__ __forceinline__ void recount_cell(const Params2& P, const uint32_t* tab, uint32_t* cols, uint32_t type,
                                             uint32_t cell, int lane) {
    int lo, hi;
    type_range(P, type, lo, hi);
    const uint32_t want = cell | (1u << 16);
    bool any = false;
    for (int base = lo; base < hi; base += 32) {
        const int k = base + lane;
        any |= __any_sync(FULL, k < hi && (tab[k] & 0x1FFFFu) == want);
    }
    if (!any && lane == 0) col_clear(cols, P.width, type, cell & 0xFFu, (cell >> 8) & 0xFFu);
    __syncwarp();
}

// One world turn, one warp per world.
__global__ void __launch_bounds__(128) wab2_grid_turn_kernel(const __grid_constant__ Params2 P, const State2Ptrs st,
                                                             const uint8_t* __restrict__ actions, const Out2Ptrs out,
                                                             const int stream_words) {
    extern __shared__ uint32_t smem2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int E = P.n_entities, A = P.n_acting, W = P.width, H = P.height;
    const GridGeom g = grid_geom(E, W, stream_words);
    uint2* lut = reinterpret_cast<uint2*>(smem2 + wpb * g.total);
    build_lut(lut);
    const int64_t idx = (int64_t)blockIdx.x * wpb + warp, n = st.n;
    if (idx >= n) return;                                   // whole warp
    uint32_t* obj = smem2 + warp * g.total + g.ent;
    uint32_t* tab = obj + E;
    uint32_t* food = tab + E;
    uint32_t* cols = smem2 + warp * g.total + g.cols;
    uint32_t* stream = smem2 + warp * g.total + g.stream;
    const uint32_t* src = st.ent + idx * st.stride_world;
    for (int k = lane; k < 3 * E; k += 32) obj[(k % 3) * E + k / 3] = src[k];   // world-major [entity][obj, tab, food]: coalesced read, planar in smem
    for (int k = lane; k < 3 * W * 2; k += 32) cols[k] = 0u;
    const uint32_t env_id = (uint32_t)(P.env_id_base + (uint64_t)idx), episode = st.episode[idx];
    uint32_t turn = st.turn[idx];
    __syncwarp();
    for (int k = lane; k < E; k += 32) {
        const uint32_t t = tab[k];
        if ((t >> 16) & 1u) col_set(cols, W, entity_type(P, k), t & 0xFFu, (t >> 8) & 0xFFu);
    }
    __syncwarp();
    const int R = P.window_r, S = 2 * R + 1, obs_bytes = 3 * S * S;
    // bushes never move: their "action" only refreshes a table position left stale by reset_world (World.py:353-356),
    // i.e. it is a no-op except in the first turn of an episode
    const int last = turn == 0u ? E : A;
    for (int a = 0; a < last; ++a) {
        const bool acting = a < A;
        const uint32_t at = entity_type(P, a);
        const int64_t o = (int64_t)a * n + idx;
        uint32_t atab = tab[a];
        if (acting && out.planes) {                                  // get_observations(a), World.py:360-377
            const int64_t first_byte = o * obs_bytes;
            const int off = (int)(first_byte & 15);
#pragma unroll 1
            for (int k = lane; k < stream_words; k += 32) stream[k] = 0u;     // rolled: the unrolled form cost ~70 instructions
            __syncwarp();
            const int ax = (int)(atab & 0xFFu), ay = (int)((atab >> 8) & 0xFFu);
            const int rsel = at == T_WOLF ? 2 : (((atab >> 17) & 1u) ? 1 : 0);
            const int r = rsel == 2 ? P.wolf_r : (rsel == 1 ? P.gatherer_r : P.lookout_r);
            // no division anywhere below: 0 <= ax < W, 0 <= ay < H and R < W, H, so one conditional add wraps
            const int sy = ay - R + (ay < R ? H : 0);
            const uint64_t smask = (1ull << S) - 1ull;
#pragma unroll 1
            for (int dxi = lane; dxi < S; dxi += 32) {               // one lane per column offset: the three type
                const int dx = dxi - R, adx = dx < 0 ? -dx : dx;     // planes share x, rotation, circle mask and quirks
                if (adx > r) continue;
                int x = ax + dx;
                x += x < 0 ? W : 0;
                x -= x >= W ? W : 0;
                const int m = (int)P.halfwidth[rsel][adx];           // |dy| <= m  <=>  dx^2 + dy^2 <= r^2
                uint64_t mask = (((2ull << (2 * m)) - 1ull) << (R - m)) & smask;
                if (dx == r && ax + r == W) mask = 0;                // World.py:264 strict test: this image is missed
                if (dx == 0 && ay + r == H) mask &= ~(1ull << (R + r));   // World.py:285, same on the y axis
#pragma unroll
                for (int type_p = 0; type_p < 3; ++type_p) {
                    const uint2 cw = *reinterpret_cast<const uint2*>(cols + ((type_p * W + x) << 1));
                    const uint64_t bits = rot_window((uint64_t)cw.x | ((uint64_t)cw.y << 32), sy, H) & mask;
                    if (bits) {
                        const int pos = off + (type_p * S + dxi) * S;
                        const uint64_t sh = bits << (pos & 31);
                        atomicOr(stream + (pos >> 5), (uint32_t)sh);
                        if (sh >> 32) atomicOr(stream + (pos >> 5) + 1, (uint32_t)(sh >> 32));
                    }
                }
            }
            __syncwarp();
            stream_flush(stream, lut, out.planes + (first_byte - off), off, off + obs_bytes, lane);
            __syncwarp();
        }
        if (acting && out.internal && lane == 0) {                   // internal_obs, World.py:50-51, :80-81
            int32_t* dst = out.internal + o * 5;
            const uint32_t ob = obj[a];
            dst[0] = unpack_x(ob); dst[1] = unpack_y(ob); dst[2] = (int32_t)food[a];
            dst[3] = (int32_t)((atab >> 17) & 1u); dst[4] = (int32_t)((atab >> 18) & 3u);
        }
        // ---- take_action(a): act, table update (World.py:325-334)
        const uint32_t action = acting ? (uint32_t)actions[o] : 0u;
        int32_t x = unpack_x(obj[a]), y = unpack_y(obj[a]);
        uint32_t role = (atab >> 17) & 1u;
        if (at != T_BUSH) {
            if (action == 0u) y += 1; else if (action == 1u) x += 1; else if (action == 2u) y -= 1; else if (action == 3u) x -= 1;
            else if (at == T_OSTRICH && action == 4u) role = 0u; else if (at == T_OSTRICH && action == 5u) role = 1u;
        }
        uint32_t tx, ty;
        if (turn == 0u) {          // tables may be stale after reset_world: the full wrap of the object coordinates
            tx = (uint32_t)pymod(x, W); ty = (uint32_t)pymod(y, H);
        } else {                   // afterwards the table follows the object one cell at a time
            int nx = (int)(atab & 0xFFu) + (x - unpack_x(obj[a])), ny = (int)((atab >> 8) & 0xFFu) + (y - unpack_y(obj[a]));
            nx += nx < 0 ? W : 0; nx -= nx >= W ? W : 0;
            ny += ny < 0 ? H : 0; ny -= ny >= H ? H : 0;
            tx = (uint32_t)nx; ty = (uint32_t)ny;
        }
        const uint32_t old_cell = atab & 0xFFFFu, new_cell = tx | (ty << 8);
        const uint32_t vis = (atab >> 16) & 1u;
        atab = tab_pack(tx, ty, vis, role, (atab >> 18) & 3u);
        __syncwarp();
        if (lane == 0) { obj[a] = pack_xy(x, y); tab[a] = atab; }
        __syncwarp();
        if (vis && old_cell != new_cell) {
            if (lane == 0) col_set(cols, W, at, tx, ty);
            recount_cell(P, tab, cols, at, old_cell, lane);
        }
        // ---- default_game_update (World.py:93-132)
        if (at != T_BUSH) {
            const uint32_t want = at == T_WOLF ? T_OSTRICH : T_BUSH;
            if (col_test(cols, W, want, tx, ty)) {
                int lo, hi;
                type_range(P, want, lo, hi);
                const uint32_t cellv = new_cell | (1u << 16);
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
                    if (at == T_WOLF) {
                        const uint32_t jt = tab[j];
                        __syncwarp();              // every lane holds the old row of label j before lane 0 rewrites it
                        if (lane == 0) {
                            food[a] += (uint32_t)P.wolf_eat_gain;                                  // :113
                            tab[pick] = (tab[pick] & ~(3u << 18)) | (2u << 18);                    // :114 killed
                            tab[j] = tab[j] & ~(1u << 16);                                         // :115 hides LABEL j
                        }
                        __syncwarp();
                        if ((jt >> 16) & 1u) recount_cell(P, tab, cols, entity_type(P, j), jt & 0xFFFFu, lane);
                    } else {
                        if (lane == 0) {                                                           // Bush.take_food
                            int32_t bf = (int32_t)food[pick], got;
                            if (bf >= P.bush_given) { bf -= P.bush_given; got = P.bush_given; }
                            else { got = bf; bf = 0; tab[pick] &= ~(1u << 17); }
                            food[pick] = (uint32_t)bf;
                            food[a] += (uint32_t)got;                                              // :127
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (acting && lane == 0) {                                   // compute_reward / is_done
            const uint32_t now = tab[a];
            float reward; uint32_t done;
            if (at == T_OSTRICH) { const uint32_t s2 = (now >> 18) & 3u; reward = s2 == 0u ? 1.f : 0.f; done = s2 != 0u; }
            else { reward = (int32_t)food[a] > 10 ? 1.f : 0.f; done = (((now >> 18) & 3u) == 1u); }
            out.reward[o] = reward;
            out.done[o] = (uint8_t)done;
        }
        __syncwarp();
    }
    turn += 1;
    uint32_t* dst = st.ent + idx * st.stride_world;
    for (int k = lane; k < 3 * E; k += 32) dst[k] = obj[(k % 3) * E + k / 3];
    if (lane == 0) st.turn[idx] = turn;
}

}  // namespace
