// wab_features.cuh — device-side equivalent of the reference's PragmaticObsWrapper
// (/root/reference/wab_env.py:670-824): nearest / second-nearest / count per direction for wolves and
// bushes, standing-on-bush, computed from the 121-bit observation planes instead of float64 grids.
// Host/device portable like wab_core.cuh so the `-m "not gpu"` suite can check it against the
// reference's own known-answer tests (wab_env_test.py:9-169).
//
// Quirks reproduced on purpose (SURVEY §8f ①): the scan is row-major and uses `<=`, so among equally
// distant objects the LAST one wins first place and pushes the previous winner to second
// (wab_env.py:778-791); "up" means a smaller FIRST grid index although that axis is x (:792-799,
// :819-822); counts are clipped at 10 (:734, :737); standing_on_bush reads bushes[5][5] (:742).
#pragma once
#include "wab_core.cuh"

namespace wab {

constexpr int FEAT_BYTES = 28;   // nearest_wolf[4] second_wolf[4] n_wolves[4] nearest_bush[4] second_bush[4] n_bushes[4]
                                 // standing_on_bush food role status
constexpr int MAX_DISTANCE = VIEW / 2 + VIEW / 2 + 1;    // 11, wab_env.py:709
constexpr int FLAT_DIM = 2 * (2 * 4 * (MAX_DISTANCE + 1) + 4 * 11) + 2 + 2 + 3 + CELLS;   // + food one-hot added at run time

// rows i < 5 / i > 5 / columns j < 5 / j > 5 of the 11x11 plane as 121-bit masks
constexpr uint32_t dirmask_word(int dir, int w) {
    uint32_t m = 0;
    for (int i = 0; i < VIEW; ++i)
        for (int j = 0; j < VIEW; ++j) {
            const bool in = dir == 0 ? i < HALF : dir == 1 ? j > HALF : dir == 2 ? i > HALF : j < HALF;
            const int c = 11 * i + j;
            if (in && (c >> 5) == w) m |= 1u << (c & 31);
        }
    return m;
}
template <int D, int W> struct DirMask { static constexpr uint32_t v = dirmask_word(D, W); };

WAB_HD uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return a | (b << 8) | (c << 16) | (d << 24); }

// _get_num_things_each_direction (wab_env.py:812-824), clipped at 10 (:734); bytes [up, right, down, left]
WAB_HD uint32_t direction_counts(const uint32_t m[4]) {
    const uint32_t up = popc32(m[0] & DirMask<0, 0>::v) + popc32(m[1] & DirMask<0, 1>::v) + popc32(m[2] & DirMask<0, 2>::v) + popc32(m[3] & DirMask<0, 3>::v);
    const uint32_t ri = popc32(m[0] & DirMask<1, 0>::v) + popc32(m[1] & DirMask<1, 1>::v) + popc32(m[2] & DirMask<1, 2>::v) + popc32(m[3] & DirMask<1, 3>::v);
    const uint32_t dn = popc32(m[0] & DirMask<2, 0>::v) + popc32(m[1] & DirMask<2, 1>::v) + popc32(m[2] & DirMask<2, 2>::v) + popc32(m[3] & DirMask<2, 3>::v);
    const uint32_t le = popc32(m[0] & DirMask<3, 0>::v) + popc32(m[1] & DirMask<3, 1>::v) + popc32(m[2] & DirMask<3, 2>::v) + popc32(m[3] & DirMask<3, 3>::v);
    return pack4(up < 10u ? up : 10u, ri < 10u ? ri : 10u, dn < 10u ? dn : 10u, le < 10u ? le : 10u);
}

WAB_HD uint32_t encode_offsets(int32_t rr, int32_t rc) {   // wab_env.py:792-808; bytes [up, right, down, left]
    const int32_t up = rr < 0 ? -rr : 0, right = rc > 0 ? rc : 0, down = rr > 0 ? rr : 0, left = rc < 0 ? -rc : 0;
    return pack4((uint32_t)(up ? MAX_DISTANCE - up : 0), (uint32_t)(right ? MAX_DISTANCE - right : 0),
                 (uint32_t)(down ? MAX_DISTANCE - down : 0), (uint32_t)(left ? MAX_DISTANCE - left : 0));
}

// _get_nearest_things (wab_env.py:763-810): sequential row-major scan of the set bits
WAB_HD void nearest_two(const uint32_t m[4], uint32_t& first, uint32_t& second) {
    int32_t s1 = MAX_DISTANCE, s2 = MAX_DISTANCE, r1 = 0, c1 = 0, r2 = 0, c2 = 0;
    WAB_ROLLED
    for (int w = 0; w < 4; ++w) {
        uint32_t bits = w == 0 ? m[0] : w == 1 ? m[1] : w == 2 ? m[2] : m[3];
        WAB_ROLLED
        while (bits) {
#if defined(__CUDA_ARCH__)
            const int b = __ffs((int)bits) - 1;
#else
            const int b = __builtin_ctz(bits);
#endif
            bits &= bits - 1u;
            const int c = 32 * w + b;
            const int i = (c * 373) >> 12;                 // c / 11 for 0 <= c < 121
            const int32_t rr = i - HALF, rc = c - 11 * i - HALF;
            const int32_t t = (rr < 0 ? -rr : rr) + (rc < 0 ? -rc : rc);
            if (t <= s1) { s2 = s1; r2 = r1; c2 = c1; s1 = t; r1 = rr; c1 = rc; }
            else if (t <= s2) { s2 = t; r2 = rr; c2 = rc; }
        }
    }
    first = encode_offsets(r1, c1);      // an empty plane leaves (0, 0) -> [0, 0, 0, 0] (:771-772)
    second = encode_offsets(r2, c2);
}

// observation(obs) of PragmaticObsWrapper (wab_env.py:726-761) for one env as 7 little-endian words
// (= the 28 feature bytes); wm / bm are the wolf and bush planes AS OBSERVED (after mask_grid).
WAB_HD void pragmatic_features(const uint32_t wm[4], const uint32_t bm[4], uint32_t food, uint32_t role,
                               uint32_t status, uint32_t out[7]) {
    nearest_two(wm, out[0], out[1]);
    out[2] = direction_counts(wm);
    nearest_two(bm, out[3], out[4]);
    out[5] = direction_counts(bm);
    out[6] = pack4((bm[1] >> 28) & 1u /* bushes[5][5], :742 */, food & 0xFFu, role & 0xFFu, status & 0xFFu);
}

}  // namespace wab
