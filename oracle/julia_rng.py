"""TEST INFRASTRUCTURE (oracle/): Julia 1.4's MersenneTwister + randn, restated so that the reference notebook
`notebooks/Mississippi_sharp_null_sims.ipynb` (cells 11-16: `Random.seed!(4); Ysim = prior_rand(gpNull)`, the two analytic
p-values and the three bootstrap tests that follow) can be replayed without Julia.  Only tests/ and tests/golden/make_golden.py import this.

What is restated [recalled from the published algorithms; pinned by the notebook's printed outputs]:
  * dSFMT-19937 (Saito & Matsumoto, dSFMT 2.2): `dsfmt_init_by_array`, the 128-bit recursion, doubles in [1, 2);
  * Random.MersenneTwister: `seed!(rng, n)` -> `init_by_array(UInt32[n])`, a cache of 1002 doubles refilled by
    `dsfmt_fill_array_close1_open2!` (same sequence as successive generation);
  * `randn(rng)`: 256-layer ziggurat on a 52-bit integer (Random/src/normal.jl), tables built by the randmtzig recurrence
    (last-bit differences in the tables move a draw only if it hits a layer edge exactly, probability 2^-51);
  * `randn!(rng, ::Array{Float64})`: either one `randn` per element, or (bulk=True, the MersenneTwister method) the
    array is first filled with raw [1,2) doubles and then converted element by element, the rare rejection draws
    coming from the generator AFTER the bulk fill.

VALIDATED against the notebook: seeding, the dSFMT recursion and cache, scalar `rand()` and `randn()` over 4.4 million
draws (the exact boot_mlltest count 354 / 10 000 and both boot_chi2test counts depend on every one of them;
`rand(mvn)` of Distributions goes through the global-RNG wrapper, i.e. one `randn` per element, bulk=False).
Round 2 -- also VALIDATED, against notebooks/SimulatedExample.ipynb (Julia 1.4.0; oracle/simulated_example.py replays its
cells 5-18 and reproduces the printed optimum 123.1640 to all seven printed digits, the four fitted hyper-parameters, both
LATE estimates and both printed p-values to 1e-4 relative):
  * `rand(2, n)` = `rand!(rng, ::Array{Float64})` (Random/src/RNGs.jl): n2 = (n-2) div 2 * 2 values come straight from the
    dSFMT stream (`dsfmt_fill_array_close_open!` on the caller's array, bypassing the 1002-double cache) when n2 >= 382,
    the remaining n - n2 through the cache; smaller arrays are copied from the cache;
  * `randn(n)` consumes ONE scalar `randn` per element in Julia 1.4.0 (bulk=False below is what reproduces the notebook;
    the bulk=True variant does not);
  * `rand(["A","B","C"], n)` = `SamplerRangeFast(1:3)`: the low `bit-width(k-1)` bits of a raw 52-bit draw, rejected while
    they exceed k-1.
"""
import math

import numpy as np

_M64 = (1 << 64) - 1
_M32 = (1 << 32) - 1

# dSFMT-19937 parameters
_N = 191
_POS1 = 117
_SL1 = 19
_SR = 12
_MSK1 = 0x000FFAFFFFFFFB3F
_MSK2 = 0x000FFDFFFC90FFFD
_FIX1 = 0x90014964B32F4329
_FIX2 = 0x3B8D12AC548A7C7A
_PCV1 = 0x3D84E1AC0DC82880
_PCV2 = 0x0000000000000001
_LOW_MASK = 0x000FFFFFFFFFFFFF
_HIGH_CONST = 0x3FF0000000000000


class DSFMT:
    """state: (N+1) 128-bit words as pairs of uint64 (u[2i], u[2i+1]); the last word is the 'lung'."""

    def __init__(self, key):
        self.u = [0] * (2 * (_N + 1))
        self._init_by_array([int(k) & _M32 for k in key])

    # -- initialisation -----------------------------------------------------------------------------
    def _init_by_array(self, key):
        size = (_N + 1) * 4
        lag = 11 if size >= 623 else 7 if size >= 68 else 5 if size >= 39 else 3
        mid = (size - lag) // 2
        p = [0x8B8B8B8B] * size

        def f1(x):
            x &= _M32
            return ((x ^ (x >> 27)) * 1664525) & _M32

        def f2(x):
            x &= _M32
            return ((x ^ (x >> 27)) * 1566083941) & _M32

        klen = len(key)
        count = max(klen + 1, size)
        r = f1(p[0] ^ p[mid % size] ^ p[(size - 1) % size])
        p[mid % size] = (p[mid % size] + r) & _M32
        r = (r + klen) & _M32
        p[(mid + lag) % size] = (p[(mid + lag) % size] + r) & _M32
        p[0] = r
        count -= 1
        i, j = 1, 0
        while j < count and j < klen:
            r = f1(p[i] ^ p[(i + mid) % size] ^ p[(i + size - 1) % size])
            p[(i + mid) % size] = (p[(i + mid) % size] + r) & _M32
            r = (r + key[j] + i) & _M32
            p[(i + mid + lag) % size] = (p[(i + mid + lag) % size] + r) & _M32
            p[i] = r
            i = (i + 1) % size
            j += 1
        while j < count:
            r = f1(p[i] ^ p[(i + mid) % size] ^ p[(i + size - 1) % size])
            p[(i + mid) % size] = (p[(i + mid) % size] + r) & _M32
            r = (r + i) & _M32
            p[(i + mid + lag) % size] = (p[(i + mid + lag) % size] + r) & _M32
            p[i] = r
            i = (i + 1) % size
            j += 1
        for _ in range(size):
            r = f2((p[i] + p[(i + mid) % size] + p[(i + size - 1) % size]) & _M32)
            p[(i + mid) % size] ^= r
            r = (r - i) & _M32
            p[(i + mid + lag) % size] ^= r
            p[i] = r
            i = (i + 1) % size
        # little-endian 32 -> 64
        for w in range(2 * (_N + 1)):
            self.u[w] = p[2 * w] | (p[2 * w + 1] << 32)
        for w in range(2 * _N):                                   # initial_mask
            self.u[w] = (self.u[w] & _LOW_MASK) | _HIGH_CONST
        t0 = self.u[2 * _N] ^ _FIX1                                # period_certification
        t1 = self.u[2 * _N + 1] ^ _FIX2
        inner = (t0 & _PCV1) ^ (t1 & _PCV2)
        s = 32
        while s > 0:
            inner ^= inner >> s
            s >>= 1
        if (inner & 1) == 0:
            self.u[2 * _N + 1] ^= 1                                # PCV2 & 1 == 1

    # -- one full pass of the recursion: 2N = 382 new 64-bit outputs ------------------------------------
    def gen_all(self):
        u = self.u
        L0, L1 = u[2 * _N], u[2 * _N + 1]
        for i in range(_N):
            jb = i + _POS1
            if jb >= _N:
                jb -= _N
            t0, t1 = u[2 * i], u[2 * i + 1]
            b0, b1 = u[2 * jb], u[2 * jb + 1]
            n0 = ((t0 << _SL1) & _M64) ^ (L1 >> 32) ^ ((L1 << 32) & _M64) ^ b0
            n1 = ((t1 << _SL1) & _M64) ^ (L0 >> 32) ^ ((L0 << 32) & _M64) ^ b1
            L0, L1 = n0, n1
            u[2 * i] = (L0 >> _SR) ^ (L0 & _MSK1) ^ t0
            u[2 * i + 1] = (L1 >> _SR) ^ (L1 & _MSK2) ^ t1
        u[2 * _N], u[2 * _N + 1] = L0, L1
        return u[: 2 * _N]


class MersenneTwister:
    """Random.MersenneTwister of Julia 1.0-1.6 as far as `seed!`, `rand()` and `randn` use it."""
    CACHE = 1002   # MT_CACHE_F = 501 << 1

    def __init__(self, seed: int):
        key, n = [], int(seed)
        while True:
            key.append(n & _M32)
            n >>= 32
            if n == 0:
                break
        self.d = DSFMT(key)
        self.vals = []
        self.idx = 0
        self._pending = []

    def _gen(self):
        # dsfmt_fill_array_close1_open2!(state, vals, 1002): the same sequence as successive generation; 1002 is not a
        # multiple of 382, the surplus of the last pass stays pending for the next refill
        out = list(self._pending)
        while len(out) < self.CACHE:
            out.extend(self.d.gen_all())
        self.vals, self._pending = out[: self.CACHE], out[self.CACHE:]
        self.idx = 0

    def raw52(self) -> int:
        """next [1,2) double as its 52 mantissa bits (rand(rng, UInt52()))."""
        if self.idx >= len(self.vals):
            self._gen()
        v = self.vals[self.idx]
        self.idx += 1
        return v & _LOW_MASK

    def rand(self) -> float:
        """rand(rng) in [0, 1)."""
        bits = self.raw52() | _HIGH_CONST
        return float(np.array([bits], dtype=np.uint64).view(np.float64)[0]) - 1.0

    def _stream(self, k: int):
        """The next k outputs of the dSFMT stream itself, bypassing the cache (`dsfmt_fill_array_*` on a caller's array:
        the same linear recursion continued, the state afterwards holding the last N outputs); k is even."""
        out = list(self._pending)
        while len(out) < k:
            out.extend(self.d.gen_all())
        self._pending = out[k:]
        return out[:k]

    def rand_array(self, n: int) -> np.ndarray:
        """rand!(rng, Array{Float64}(undef, n)) in [0, 1)  (Random/src/RNGs.jl, `rand!(r::MersenneTwister,
        A::UnsafeView{Float64}, ::SamplerTrivial{<:FloatInterval_64})`)."""
        n2 = (n - 2) // 2 * 2
        if n2 < 382:                                   # dsfmt_get_min_array_size(): _rand_max383! copies from the cache
            raw = [self.raw52() for _ in range(n)]
        else:
            raw = [v & _LOW_MASK for v in self._stream(n2)] + [self.raw52() for _ in range(n - n2)]
        return (np.array(raw, dtype=np.uint64) | np.uint64(_HIGH_CONST)).view(np.float64) - 1.0

    def rand_range(self, k: int) -> int:
        """rand(rng, 1:k) through SamplerRangeFast (the MersenneTwister sampler for BitInteger unit ranges, k-1 < 2^52)."""
        m = k - 1
        mask = (1 << m.bit_length()) - 1
        while True:
            x = self.raw52() & mask
            if x <= m:
                return x + 1


# ---- ziggurat tables (randmtzig recurrence with a 2^51 integer scale) --------------------------------
_ZR = 3.6541528853610087963519472518
_ZV = 0.00492867323399
_SCALE = 2.0 ** 51


def _tables():
    ki, wi, fi = [0] * 256, [0.0] * 256, [0.0] * 256
    x1 = _ZR
    wi[255] = x1 / _SCALE
    fi[255] = math.exp(-0.5 * x1 * x1)
    ki[0] = int(x1 * fi[255] / _ZV * _SCALE)
    wi[0] = _ZV / fi[255] / _SCALE
    fi[0] = 1.0
    for i in range(254, 0, -1):
        x = math.sqrt(-2.0 * math.log(_ZV / x1 + fi[i + 1]))
        ki[i + 1] = int(x / x1 * _SCALE)
        wi[i] = x / _SCALE
        fi[i] = math.exp(-0.5 * x * x)
        x1 = x
    ki[1] = 0
    return ki, wi, fi


_KI, _WI, _FI = _tables()


def _randn_from_bits(rng: MersenneTwister, r: int) -> float:
    rabs = r >> 1
    idx = rabs & 0xFF
    x = (-rabs if (r & 1) else rabs) * _WI[idx]
    if rabs < _KI[idx]:
        return x
    # randn_unlikely
    if idx == 0:
        while True:
            xx = -(1.0 / _ZR) * math.log(rng.rand())
            yy = -math.log(rng.rand())
            if yy + yy > xx * xx:
                return (-_ZR - xx) if ((rabs >> 8) & 1) else (_ZR + xx)
    if (_FI[idx - 1] - _FI[idx]) * rng.rand() + _FI[idx] < math.exp(-0.5 * x * x):
        return x
    return randn(rng)


def randn(rng: MersenneTwister) -> float:
    return _randn_from_bits(rng, rng.raw52())


def randn_array(rng: MersenneTwister, n: int, bulk: bool) -> np.ndarray:
    """randn!(rng, Vector{Float64}(undef, n)): bulk=False one randn per element; bulk=True the MersenneTwister method
    (raw fill first, rejection draws afterwards)."""
    if not bulk:
        return np.array([randn(rng) for _ in range(n)])
    raw = [rng.raw52() for _ in range(n)]
    return np.array([_randn_from_bits(rng, r) for r in raw])
