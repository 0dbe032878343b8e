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
    BUSH   0     0        (x>>1 & 0xffff) | (y>>1 & 0xffff)<<16  (x&1)|(y&1)<<1  wab_env.py:627,631-635
    DESP   t     rank>>2  (wx & 0xffff) | (wy & 0xffff)<<16      rank & 3      wab_env.py:262-264
    START  0     0        0                                      0 food, 1 role  wab_env.py:596-599

The two sites that draw for MANY cells per step with a tiny success probability (48 ring cells per
step, 121 cells per reset, p = 0.0005) use a *two-level draw* with 48 bits of resolution:
``U = (h * 2**32 + r) * 2**-48`` where ``h`` is a 16-bit half-word of a PRIMARY call shared by 8
cells and ``r`` a word of a SECONDARY call shared by 4 cells. ``U < p`` is decided by ``h`` alone
unless ``h`` equals the top 16 bits of the threshold (probability 2**-16), so an implementation
evaluates the secondary call lazily; the draw itself is an exact, iid uniform either way.

    site   turn  primary (sub 0)                    secondary (sub 1)          cite
    INIT   0     payload c>>3, half-word c&7        payload c>>2, lane c&3     wab_env.py:588-591   c = (x+W//2)*H + (y+H//2)
    SPAWN  t     payload j>>3, half-word j&7        payload j>>2, lane j&3     wab_env.py:571-574   j = ring index (below)

(half-word k of a call = bits 16*(k&1) .. 16*(k&1)+15 of word k>>1.)

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


def bush_words(seed, env_id, episode, x, y):
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    return _draw(seed, env_id, episode, SITE_BUSH, 0, 0, _pack_xy(x >> 1, y >> 1), (x & 1) | ((y & 1) << 1))


def two_level_units(seed, env_id, episode, site, turn, index):
    """Exact doubles U = (h * 2**32 + r) * 2**-48 for the cells `index` of a two-level site."""
    index = np.asarray(index, dtype=np.int64)
    word = _draw(seed, env_id, episode, site, turn, 0, index >> 3, (index >> 1) & 3).astype(np.uint64)
    h = (word >> (np.uint64(16) * (index & 1).astype(np.uint64))) & np.uint64(0xFFFF)
    r = _draw(seed, env_id, episode, site, turn, 1, index >> 2, index & 3).astype(np.uint64)
    return ((h << np.uint64(32)) | r).astype(np.float64) * (2.0 ** -48)


def init_units(seed, env_id, episode, x, y, width, height):
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    return two_level_units(seed, env_id, episode, SITE_INIT, 0, (x + width // 2) * height + (y + height // 2))


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


def spawn_units(seed, env_id, episode, turn, dx, dy, width, height, margin):
    return two_level_units(seed, env_id, episode, SITE_SPAWN, turn, ring_index(dx, dy, width, height, margin))


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
