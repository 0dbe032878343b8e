"""Keyed counter-based randomness contract (TEST INFRASTRUCTURE — oracle side).

This is the *oracle-side* statement of the random-draw contract that replaces the reference's
global ``np.random`` stream (reference draw sites: ``wab_env.py:263`` despawn, ``:572`` spawn,
``:589`` wolf init, ``:597-599`` start food / role, ``:633`` bush value). The reference consumes
draws in CPython set-iteration order, which no parallel implementation can reproduce, so every
draw is instead *keyed by identity*: (seed, env, episode, site, turn, cell or wolf). Both the
shimmed reference (``oracle/ref_shim``), the C restatement (``oracle/wab_oracle.c``) and the CUDA
kernels evaluate this same function; parity is then independent of evaluation order.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.

Contract
--------
``philox4x32-10`` (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
Random123 constants) with

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (env_id, episode, (site << 28) | (turn << 8) | sub, payload)

returns four 32-bit words. A *word draw* is ONE word ``w`` ("lane") and the uniform handed to the
reference is the exact double ``U = w * 2**-32``:

    site   turn  sub      payload                              lane          cite
    DESP   t     rank>>2  (wx & 0xffff) | (wy & 0xffff)<<16      rank & 3      wab_env.py:262-264
    START  0     0        0                                      0 food, 1 role  wab_env.py:596-599
                                                                 2, 3 = the episode's bush key (ka, kb), below

Bush values (wab_env.py:627, 631-635) are the highest-volume draws: 11 newly revealed cells on almost every step
and a cell must get the same value whenever it is revealed. They come from ``philox2x32-10`` (same paper; half the
multiplies of the 4x32 variant) keyed per episode: with B = (x>>1 & 0xffff) | (y>>1 & 0xffff)<<16 the 2x2 block of a
cell and k2 = (seed ^ seed >> 32) & 0xffffffff,

    (p0, p1) = philox2x32_10(counter = (B ^ ka,  kb), key = k2)     high half-words of the block's 4 cells
    (q0, q1) = philox2x32_10(counter = (B ^ ka, ~kb), key = k2)     low half-words
    word(x, y) = h << 16 | l,  half-word index (x&1) | (y&1)<<1 of (p0, p1) resp. (q0, q1)

and U = word * 2**-32 as for every word draw. Whether a cell has a bush is decided by ``h`` alone unless it equals
the top half of the first bush threshold (probability 2**-16), so the low half is evaluated lazily.

The two sites that draw for MANY cells at once with a tiny success probability (48 ring cells per step,
W*H cells per reset, p = chance/2 = 0.0005) are *binomial-first*: instead of n independent word draws,
the n uniforms handed to the reference are generated in the statistically identical order "how many
succeed, which ones, then each value given its outcome":

    primary    call (site, turn, sub 0, payload 0):      v = w0 << 32 | w1            (64 bits)
               K = #{k < 32 : v >= T_k},  T_k = min(ceil(BinomialCDF(n, p; k) * 2**64), 2**64 - 1),  K <= n
    choice     call (site, turn, sub 1, payload i >> 2), word i & 3 -> r_i,  i = 0 .. K-1:
               q_i = (r_i * (n - i)) >> 32 ; cell_i = the q_i-th (ascending) index not chosen so far
    value      call (site, turn, sub 2, payload j >> 2), word j & 3 -> V_j = word * 2**-32, j = 0 .. n-1:
               U_j = p * V_j              if j was chosen      (always <  p)
               U_j = p + (1 - p) * V_j    otherwise            (always >= p)

(K ~ Binomial(n, p), a uniformly random K-subset, U | success ~ U(0, p), U | failure ~ U(p, 1): the joint
law of (U_0 .. U_{n-1}) is exactly n iid uniforms, up to the 2**-64 / 2**-32 granularity of the tables.)
The reference evaluates ``U_j < p`` on all n values; a fast implementation needs ONE Philox call and one
64-bit compare (``v < T_0``, probability (1-p)**n = 97.6 % for the ring) to know that nothing happens.

    site   turn  n                      index                                   cite
    INIT   0     W * H                  c = (x+W//2)*H + (y+H//2)               wab_env.py:588-591
    SPAWN  t     ring cells (48)        j = ring index (below)                  wab_env.py:571-574

``ring index``: cells of the (W+2m)x(H+2m) box around the (already moved) ostrich minus its WxH
view, enumerated box-x-major then box-y, skipping interior cells. ``rank``: ordinal of a wolf
among the wolves standing on the same cell (co-located wolves move identically for ever, so any
assignment of ranks inside a stack yields the same multiset of survivors).
"""
import numpy as np

SITE_BUSH, SITE_INIT, SITE_SPAWN, SITE_DESP, SITE_START = 1, 2, 3, 4, 5

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Inputs broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & _MASK for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


_M2 = np.uint64(0xD256D193)


def philox2x32_10(c0, c1, key):
    """Vectorised Philox2x32-10 (Random123). Returns 2 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c0, c1 = np.broadcast_arrays(c0, c1)
    k = int(key) & 0xFFFFFFFF
    for _ in range(10):
        p = _M2 * c0
        c0, c1 = (p >> _S32) ^ np.uint64(k) ^ c1, p & _MASK
        k = (k + _W0) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32)


def _draw(seed, env_id, episode, site, turn, sub, payload, lane):
    """One 32-bit word per element of the broadcast (payload, lane, sub, turn) arrays."""
    payload = np.asarray(payload, dtype=np.int64) & 0xFFFFFFFF
    lane = np.asarray(lane, dtype=np.int64)
    sub = np.asarray(sub, dtype=np.int64)
    turn = np.asarray(turn, dtype=np.int64)
    c2 = (int(site) << 28) | ((turn & 0xFFFFF) << 8) | (sub & 0xFF)
    out = philox4x32_10(
        np.uint64(int(env_id) & 0xFFFFFFFF),
        np.uint64(int(episode) & 0xFFFFFFFF),
        c2,
        payload,
        int(seed) & 0xFFFFFFFF,
        (int(seed) >> 32) & 0xFFFFFFFF,
    )
    stacked = np.stack(np.broadcast_arrays(*out), axis=-1)
    lane_b = np.broadcast_to(lane, stacked.shape[:-1])
    return np.take_along_axis(stacked, lane_b[..., None], axis=-1)[..., 0].astype(np.uint32)


def _pack_xy(x, y):
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    return (x & 0xFFFF) | ((y & 0xFFFF) << 16)


def bush_key(seed, env_id, episode):
    """(ka, kb): words 2 and 3 of the episode's START call."""
    w = _draw(seed, env_id, episode, SITE_START, 0, 0, np.zeros(2, dtype=np.int64), np.array([2, 3]))
    return int(w[0]), int(w[1])


def bush_words(seed, env_id, episode, x, y):
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    ka, kb = bush_key(seed, env_id, episode)
    k2 = (int(seed) ^ (int(seed) >> 32)) & 0xFFFFFFFF
    block = (_pack_xy(x >> 1, y >> 1) ^ ka) & 0xFFFFFFFF
    lane = (x & 1) | ((y & 1) << 1)

    def half(words):
        w = np.where(lane >= 2, words[1], words[0]).astype(np.uint64)
        return (w >> (np.uint64(16) * (lane & 1).astype(np.uint64))) & np.uint64(0xFFFF)

    hi = half(philox2x32_10(block, kb, k2))
    lo = half(philox2x32_10(block, (~kb) & 0xFFFFFFFF, k2))
    return ((hi << np.uint64(16)) | lo).astype(np.uint32)


BINOMIAL_TABLE = 32   # thresholds kept; configurations whose tail beyond this is not negligible are rejected


def binomial_thresholds(n, p):
    """T_k = min(ceil(CDF(k) * 2**64), 2**64 - 1) for k = 0 .. 31, exact rational arithmetic on the double p."""
    from fractions import Fraction
    from math import comb
    n = int(n)
    p = min(max(Fraction(float(p)), Fraction(0)), Fraction(1))
    out, cdf = [], Fraction(0)
    for k in range(BINOMIAL_TABLE):
        if k <= n:
            cdf += comb(n, k) * p ** k * (1 - p) ** (n - k)
        t = -((-cdf * (1 << 64)).__floor__())            # ceil
        out.append(min(int(t), (1 << 64) - 1))
    return out


def binomial_first(seed, env_id, episode, site, turn, n, p):
    """(K, chosen cell indices in choice order) of a binomial-first site."""
    w = [_draw(seed, env_id, episode, site, turn, 0, 0, lane) for lane in (0, 1)]
    v = (int(w[0]) << 32) | int(w[1])
    k = min(sum(1 for t in binomial_thresholds(n, p) if v >= t), int(n))
    chosen = []
    free = list(range(int(n)))
    for i in range(k):
        r = int(_draw(seed, env_id, episode, site, turn, 1, i >> 2, i & 3))
        chosen.append(free.pop((r * (int(n) - i)) >> 32))
    return k, chosen


def binomial_first_units(seed, env_id, episode, site, turn, n, p, index):
    """The exact doubles U_j handed to the reference for the cells `index` (any order) of a binomial-first site."""
    index = np.asarray(index, dtype=np.int64)
    _, chosen = binomial_first(seed, env_id, episode, site, turn, n, p)
    v = to_unit(_draw(seed, env_id, episode, site, turn, 2, index >> 2, index & 3))
    hit = np.isin(index, np.asarray(chosen, dtype=np.int64))
    p = float(p)
    return np.where(hit, p * v, p + (1.0 - p) * v)


def init_units(seed, env_id, episode, x, y, width, height, p):
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    return binomial_first_units(seed, env_id, episode, SITE_INIT, 0, width * height, p,
                                (x + width // 2) * height + (y + height // 2))


def ring_index(dx, dy, width, height, margin):
    """Index of ring cell at offset (dx, dy) from the ostrich; box-x-major enumeration skipping the view."""
    dx = np.asarray(dx, dtype=np.int64)
    dy = np.asarray(dy, dtype=np.int64)
    hw, hh = width // 2, height // 2
    bx = dx + hw + margin
    by = dy + hh + margin
    bh = height + 2 * margin
    full_before = np.minimum(bx, margin) * bh + np.clip(bx - (width + margin), 0, None) * bh
    mid_before = np.clip(bx - margin, 0, width) * (2 * margin)
    in_mid = (bx >= margin) & (bx < width + margin)
    within = np.where(in_mid, np.where(by < margin, by, by - height), by)
    return full_before + mid_before + within


def ring_size(width, height, margin):
    return (width + 2 * margin) * (height + 2 * margin) - width * height


def spawn_units(seed, env_id, episode, turn, dx, dy, width, height, margin, p):
    return binomial_first_units(seed, env_id, episode, SITE_SPAWN, turn, ring_size(width, height, margin), p,
                                ring_index(dx, dy, width, height, margin))


def despawn_words(seed, env_id, episode, turn, wx, wy):
    """Words for wolves listed in any order; rank = ordinal among earlier wolves on the same cell."""
    wx = np.asarray(wx, dtype=np.int64)
    wy = np.asarray(wy, dtype=np.int64)
    rank = np.zeros(len(wx), dtype=np.int64)
    seen = {}
    for i, cell in enumerate(zip(wx.tolist(), wy.tolist())):
        rank[i] = seen.get(cell, 0)
        seen[cell] = rank[i] + 1
    if len(wx) == 0:
        return np.zeros(0, dtype=np.uint32)
    return _draw(seed, env_id, episode, SITE_DESP, turn, rank >> 2, _pack_xy(wx, wy), rank & 3)


def start_words(seed, env_id, episode):
    """(food word, role word)."""
    w = _draw(seed, env_id, episode, SITE_START, 0, 0, np.zeros(2, dtype=np.int64), np.array([0, 1]))
    return int(w[0]), int(w[1])


def to_unit(words):
    """Exact double U = w * 2**-32 in [0, 1)."""
    return np.asarray(words, dtype=np.float64) * (2.0 ** -32)
