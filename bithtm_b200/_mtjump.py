"""Jump-ahead polynomials for the device MT19937 generator (host-side precompute).

The legacy ``np.random`` stream the reference draws from (networks.py:87,
projections.py:120,235) is one sequential MT19937 word sequence x[n].  To produce it
on many CTAs at once the device needs the generator state at word offsets D_p ahead
of a window of known words.  Over GF(2) the word sequence obeys the linear recurrence
of the generator's characteristic polynomial phi(t) (degree 19937, 135 terms):

    XOR_{e in PHI} x[n + e] = 0                      for every n >= 1
    x[m + D]  =  XOR_{i : g_D[i] = 1} x[m + i]       with g_D(t) = t^D mod phi(t)

(n >= 1 because the low 31 bits of a freshly seeded key[0] are not part of the state).
``jump_table`` returns g_D for D = D0 + p*J as bit-packed rows of 624 words; the kernel
``ph_rng_chunk`` (csrc/mt19937.cuh) applies row p to a window of D0 = 19937 + 623 words
and obtains the 624 state words that start chunk p.

PHI was obtained by Berlekamp-Massey on one output bit of the generator and is
checked against a simulated stream by ``self_check`` (tests/test_cpu.py).
"""

from __future__ import annotations

import os

import numpy as np

MT_N, MT_M = 624, 397
DEGREE = 19937
# exponents of the characteristic polynomial of MT19937 (135 terms)
PHI = (
    0, 1189, 1416, 1585, 1643, 1870, 2493, 2773, 3000, 3227, 3454, 3681, 3908, 4135, 4362, 4753, 5661, 6337, 6569,
    7129, 7477, 7525, 7583, 7752, 7979, 8206, 9505, 9901, 9969, 10128, 10693, 10761, 10920, 11089, 11147, 11157,
    11215, 11321, 11374, 11384, 11485, 11611, 11712, 11717, 11838, 11881, 11944, 11997, 12277, 12335, 12393, 12504,
    12509, 12620, 12673, 12731, 12736, 12789, 12905, 12958, 12963, 13137, 13185, 13190, 13243, 13301, 13412, 13528,
    13533, 13639, 13697, 13760, 13813, 13866, 14093, 14151, 14209, 14320, 14325, 14436, 14547, 14552, 14605, 14721,
    14774, 14779, 14953, 15001, 15006, 15059, 15117, 15228, 15344, 15349, 15455, 15513, 15576, 15629, 15682, 15909,
    15967, 16025, 16136, 16141, 16252, 16363, 16368, 16421, 16537, 16590, 16595, 16817, 16822, 16875, 16933, 17044,
    17160, 17271, 17329, 17445, 17498, 17725, 17783, 17841, 17952, 18068, 18179, 18237, 18406, 18633, 18691, 18860,
    19087, 19314, 19937,
)
_PHI_LOW = np.array(PHI[:-1], dtype=np.int64)
MAX_SHIFT = DEGREE - PHI[-2]  # 623: a shift this small overflows into phi's low terms only once

WINDOW_WORDS = DEGREE + MT_N - 1   # 20560 words determine any 624 consecutive later words
CHUNK_WORDS = 40 * MAX_SHIFT       # J: stream words one CTA generates after its jump


def _times_t_pow(g: np.ndarray, s: int) -> np.ndarray:
    """g(t) * t^s mod phi(t) for 1 <= s <= 623; g is a uint8 coefficient array of length 19937."""
    assert 1 <= s <= MAX_SHIFT
    high = g[DEGREE - s:]            # coefficients that overflow: exponent e + s - 19937 in [0, s)
    out = np.zeros(DEGREE, dtype=np.uint8)
    out[s:] = g[:DEGREE - s]
    if high.any():
        for e in _PHI_LOW:           # t^19937 = XOR of the low terms of phi
            out[e:e + s] ^= high
    return out


def power_of_t(d: int, start: np.ndarray | None = None) -> np.ndarray:
    """t^d mod phi (times ``start`` when given)."""
    g = start
    if g is None:
        g = np.zeros(DEGREE, dtype=np.uint8)
        g[0] = 1
    while d > 0:
        s = min(d, MAX_SHIFT)
        g = _times_t_pow(g, s)
        d -= s
    return g


def pack(g: np.ndarray) -> np.ndarray:
    """uint8 coefficients [19937] -> uint32 [624], bit i of the polynomial = bit i%32 of word i//32."""
    bits = np.zeros(MT_N * 32, dtype=np.uint8)
    bits[:DEGREE] = g
    return np.packbits(bits, bitorder="little").view(np.uint32).copy()


def jump_table(n_polys: int, cache_dir: str | None = None) -> np.ndarray:
    """uint32 [n_polys][624]: row p = t^(WINDOW_WORDS + p * CHUNK_WORDS) mod phi."""
    n_polys = int(n_polys)
    path = None
    if cache_dir:
        prefix = f"mtjump_w{WINDOW_WORDS}_j{CHUNK_WORDS}_n"
        path = os.path.join(cache_dir, f"{prefix}{n_polys}.npy")
        try:  # any cached table with at least n_polys rows will do (row p does not depend on the count)
            cached = sorted((int(f[len(prefix):-4]), f) for f in os.listdir(cache_dir)
                            if f.startswith(prefix) and f.endswith(".npy") and f[len(prefix):-4].isdigit())
        except OSError:
            cached = []
        for n, f in cached:
            if n >= n_polys:
                try:
                    tab = np.load(os.path.join(cache_dir, f))
                    if tab.shape == (n, MT_N) and tab.dtype == np.uint32:
                        return np.ascontiguousarray(tab[:n_polys])
                except Exception:
                    pass
    tab = np.zeros((n_polys, MT_N), dtype=np.uint32)
    g = power_of_t(WINDOW_WORDS)
    for p in range(n_polys):
        tab[p] = pack(g)
        if p + 1 < n_polys:
            g = power_of_t(CHUNK_WORDS, g)
    if path:
        try:
            tmp = path + f".{os.getpid()}.tmp"
            with open(tmp, "wb") as f:  # np.save would append ".npy" to a bare path
                np.save(f, tab)
            os.replace(tmp, path)
        except OSError:
            pass
    return tab


# ------------------------------------------------------------------------------ skip table (lazy draws)
_MASK = (1 << DEGREE) - 1


def _reduce_int(x: int) -> int:
    """x(t) mod phi(t), polynomials over GF(2) as Python integers (bit i = coefficient of t^i)."""
    while x >> DEGREE:
        hi = x >> DEGREE
        x &= _MASK
        for e in PHI[:-1]:  # t^(19937 + j) = XOR_e t^(e + j)
            x ^= hi << e
    return x


def _square_int(x: int) -> int:
    return _reduce_int(int(format(x, "b"), 4))  # bit i -> bit 2i


def _mul_int(a: int, b: int) -> int:
    """a(t) * b(t) mod phi: the carry-less product as an FFT convolution of the coefficient vectors
    (counts <= 19937 are exact in float64), then the sparse reduction."""
    n = 1 << 16
    av = np.frombuffer(a.to_bytes(n // 8, "little"), dtype=np.uint8)
    bv = np.frombuffer(b.to_bytes(n // 8, "little"), dtype=np.uint8)
    ab = np.unpackbits(av, bitorder="little").astype(np.float64)
    bb = np.unpackbits(bv, bitorder="little").astype(np.float64)
    conv = np.fft.irfft(np.fft.rfft(ab) * np.fft.rfft(bb), n)
    bits = (np.rint(conv).astype(np.int64) & 1).astype(np.uint8)
    return _reduce_int(int.from_bytes(np.packbits(bits, bitorder="little").tobytes(), "little"))


def _pack_int(g: int) -> np.ndarray:
    return np.frombuffer(g.to_bytes(MT_N * 4, "little"), dtype=np.uint32).copy()


def skip_table(n_polys: int, gran: int, cache_dir: str | None = None) -> np.ndarray:
    """uint32 [n_polys][624]: row b - 1 = t^(b * gran) mod phi, b = 1..n_polys -- the jumps of the lazy
    draws (csrc/mt19937.cuh): x[m + a + b * gran + j] = XOR_{i : row[b-1][i]} x[m + a + i + j]."""
    n_polys, gran = int(n_polys), int(gran)
    path = None
    if cache_dir:
        prefix = f"mtskip_g{gran}_n"
        path = os.path.join(cache_dir, f"{prefix}{n_polys}.npy")
        try:
            cached = sorted((int(f[len(prefix):-4]), f) for f in os.listdir(cache_dir)
                            if f.startswith(prefix) and f.endswith(".npy") and f[len(prefix):-4].isdigit())
        except OSError:
            cached = []
        for n, f in cached:
            if n >= n_polys:
                try:
                    tab = np.load(os.path.join(cache_dir, f))
                    if tab.shape == (n, MT_N) and tab.dtype == np.uint32:
                        return np.ascontiguousarray(tab[:n_polys])
                except Exception:
                    pass
    tab = np.zeros((n_polys, MT_N), dtype=np.uint32)
    if gran <= 8 * DEGREE:  # multiplying by t^gran is a shift and a short reduction
        step = None
    else:
        step = 2  # t
        e, base, acc = gran, 2, 1
        while e:  # t^gran by square and multiply
            if e & 1:
                acc = _mul_int(acc, base) if acc != 1 else base
            e >>= 1
            if e:
                base = _square_int(base)
        step = acc
    g = 1
    for p in range(n_polys):
        g = _reduce_int(g << gran) if step is None else (_mul_int(g, step) if g != 1 else step)
        tab[p] = _pack_int(g)
    if path:
        try:
            tmp = path + f".{os.getpid()}.tmp"
            with open(tmp, "wb") as f:
                np.save(f, tab)
            os.replace(tmp, path)
        except OSError:
            pass
    return tab


# ------------------------------------------------------------------------------ self check
def raw_stream(key: np.ndarray, n: int) -> np.ndarray:
    """Untempered MT19937 words x[0..n) continuing key = x[0..624) (block-vectorised)."""
    x = np.zeros(max(n, MT_N) + MT_N, dtype=np.uint32)
    x[:MT_N] = key
    d = MT_N - MT_M  # 227 words can be produced at once

    def twist(u, v):
        y = (u & np.uint32(0x80000000)) | (v & np.uint32(0x7FFFFFFF))
        return (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908B0DF), np.uint32(0))

    i = MT_N
    while i < n:
        j = min(i + d, len(x))
        x[i:j] = x[i - d:j - d] ^ twist(x[i - MT_N:j - MT_N], x[i - MT_N + 1:j - MT_N + 1])
        i = j
    return x[:n]


def apply_jump(row: np.ndarray, window: np.ndarray) -> np.ndarray:
    """The 624 words that follow `window` (WINDOW_WORDS words) at the row's distance."""
    bits = np.unpackbits(row.view(np.uint8), bitorder="little")[:DEGREE]
    idx = np.nonzero(bits)[0]
    out = np.zeros(MT_N, dtype=np.uint32)
    for j in range(MT_N):
        out[j] = np.bitwise_xor.reduce(window[idx + j])
    return out


def self_check(seed: int = 3, polys: int = 3) -> bool:
    key = np.random.RandomState(seed).get_state()[1].astype(np.uint32)
    n = 1 + WINDOW_WORDS + polys * CHUNK_WORDS + MT_N
    x = raw_stream(key, n)
    acc = np.zeros(8, dtype=np.uint32)
    for e in PHI:
        acc ^= x[1 + e:9 + e]
    if acc.any():
        return False
    tab = jump_table(polys)
    window = x[1:1 + WINDOW_WORDS]
    for p in range(polys):
        start = 1 + WINDOW_WORDS + p * CHUNK_WORDS
        if not np.array_equal(apply_jump(tab[p], window), x[start:start + MT_N]):
            return False
    return True
