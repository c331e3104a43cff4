"""Philox4x32-10 counter RNG in NumPy -- bit-exact twin of `csrc/imb_rng.cuh`.

TEST INFRASTRUCTURE.  Not from the reference (which uses NumPy/torch global RNGs,
data/buffer.py:231, algorithms/base.py:272-282); this is the repo-defined device RNG
("perf mode", SURVEY.md section 7 hard part 4) restated on the CPU so that perf-mode
indices and done masks can be checked bit-exactly.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)

# stream ids (xor-ed into key word 1); must match csrc/imb_rng.cuh
STREAM_ENV_RESET = 0x1001
STREAM_ACT_NOISE = 0x2002
STREAM_REPLAY = 0x3003
STREAM_EXPERT = 0x4004
STREAM_PPO_PERM = 0x5005
STREAM_EXPERT_POLICY = 0x6006


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays/scalars."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & _MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _MASK, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


def key_for(seed: int, stream: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, ((seed >> 32) ^ stream) & 0xFFFFFFFF


def u01(x):
    """uint32 -> float32 in (0,1): ((x >> 8) + 0.5) * 2^-24."""
    return ((np.asarray(x, np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def box_muller(xa, xb):
    """Two uint32 words -> two float32 standard normals (fp32 arithmetic)."""
    u1 = u01(xa)
    u2 = u01(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(6.283185307179586) * u2).astype(np.float32)
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def normals(seed, stream, a, b, n):
    """`n` float32 normals per (a, b) pair: counter = (a, b, chunk, 0), 4 normals/chunk.

    a, b: uint32 arrays of identical shape S.  Returns float32 array S + (n,).
    """
    k0, k1 = key_for(seed, stream)
    a = np.asarray(a, np.uint32)
    b = np.broadcast_to(np.asarray(b, np.uint32), a.shape)
    nchunk = (n + 3) // 4
    out = np.empty(a.shape + (nchunk * 4,), np.float32)
    for j in range(nchunk):
        x0, x1, x2, x3 = philox4x32(a, b, np.uint32(j), np.uint32(0), k0, k1)
        z0, z1 = box_muller(x0, x1)
        z2, z3 = box_muller(x2, x3)
        out[..., 4 * j + 0] = z0
        out[..., 4 * j + 1] = z1
        out[..., 4 * j + 2] = z2
        out[..., 4 * j + 3] = z3
    return out[..., :n]


def randint(seed, stream, draw, n, size):
    """`n` indices in [0,size): element i uses counter (i//4, draw, 0, 0) word i%4,
    idx = (word * size) >> 32  (multiply-shift)."""
    k0, k1 = key_for(seed, stream)
    blk = np.arange((n + 3) // 4, dtype=np.uint32)
    xs = philox4x32(blk, np.uint32(draw & 0xFFFFFFFF), np.uint32((draw >> 32) & 0xFFFFFFFF), np.uint32(0), k0, k1)
    w = np.stack(xs, axis=1).reshape(-1)[:n].astype(np.uint64)
    return ((w * np.uint64(size)) >> np.uint64(32)).astype(np.int64)


def _feistel_round_keys(seed, stream, draw, rounds=4):
    k0, k1 = key_for(seed, stream)
    x = philox4x32(np.uint32(draw & 0xFFFFFFFF), np.uint32((draw >> 32) & 0xFFFFFFFF), np.uint32(0), np.uint32(0), k0, k1)
    return [int(v) for v in x][:rounds]


def _mix32(x):
    x = np.asarray(x, np.uint64) & _MASK
    x = (x ^ (x >> np.uint64(16))) & _MASK
    x = (x * np.uint64(0x7FEB352D)) & _MASK
    x = (x ^ (x >> np.uint64(15))) & _MASK
    x = (x * np.uint64(0x846CA68B)) & _MASK
    x = (x ^ (x >> np.uint64(16))) & _MASK
    return x


def feistel_perm(seed, stream, draw, n):
    """Pseudo-random permutation of range(n): 4-round balanced Feistel on 2*hb bits
    with cycle walking.  Element i -> perm[i].  Twin of imb_feistel_perm (CUDA)."""
    bits = max(2, int(n - 1).bit_length()) if n > 1 else 2
    hb = (bits + 1) // 2
    mask = np.uint64((1 << hb) - 1)
    keys = _feistel_round_keys(seed, stream, draw)
    x = np.arange(n, dtype=np.uint64)
    out = np.empty(n, np.int64)
    pending = np.arange(n)
    cur = x.copy()
    while len(pending):
        l = cur >> np.uint64(hb)
        r = cur & mask
        for k in keys:
            f = _mix32((r ^ np.uint64(k)) & _MASK) & mask
            l, r = r, (l ^ f) & mask
        cur = (l << np.uint64(hb)) | r
        ok = cur < np.uint64(n)
        out[pending[ok]] = cur[ok].astype(np.int64)
        pending = pending[~ok]
        cur = cur[~ok]
    return out
