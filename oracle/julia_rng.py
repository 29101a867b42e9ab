"""TEST INFRASTRUCTURE (oracle): pure-Python restatement of the random stream the reference draws from.

The reference calls Julia's global RNG (`rand()`, `rand(Float64,3)`: Ewald/main.jl:516,
Ewald/auxillary.jl:99,109, Ewald/quaternions.jl:64,176; seed named at Ewald/main.jl:36 and
Monatomic/mainMonatomic.jl:15: `Random.seed!(11234)`).  For Julia 1.x <= 1.6 that is
`MersenneTwister`: dSFMT-19937 (un-vendored C dependency of Julia itself, algorithm published by
Saito & Matsumoto, "A PRNG specialized in double precision floating point numbers using an affine
transition", 2009; parameter set dSFMT-params19937.h), seeded with
`dsfmt_init_by_array(make_seed(seed))`, `rand()` = next close1_open2 value - 1.0.

Pinned (tests/test_oracle.py) against the values the Julia manual prints for
`rand(MersenneTwister(1234), 2)` = [0.5908446386657102, 0.7667970365022592] and the well-known
`Random.seed!(0); rand(4)`.  Only tests/, smoke() and bench.py's cpu_baseline may import this.
"""
from __future__ import annotations

import struct

import numpy as np

N, POS1, SL1, SR = 191, 117, 19, 12
MSK1, MSK2 = 0x000FFAFFFFFFFB3F, 0x000FFDFFFC90FFFD
FIX1, FIX2 = 0x90014964B32F4329, 0x3B8D12AC548A7C7A
PCV1, PCV2 = 0x3D84E1AC0DC82880, 0x0000000000000001
M64, M32 = (1 << 64) - 1, (1 << 32) - 1


def make_seed(n: int) -> list[int]:
    """Random/src/RNGs.jl make_seed(n::Integer): little-endian 32-bit words, at least one."""
    if n < 0:
        raise ValueError("seed must be non-negative")
    key = []
    while True:
        key.append(n & M32)
        n >>= 32
        if n == 0:
            return key


def init_by_array(key: list[int]) -> list[int]:
    """dSFMT.c dsfmt_chk_init_by_array + initial_mask + period_certification → 2(N+1) 64-bit words."""
    size = (N + 1) * 4
    lag = 11 if size >= 623 else 7 if size >= 68 else 5 if size >= 39 else 3
    mid = (size - lag) // 2
    s = [0x8B8B8B8B] * size

    def f1(x):
        return ((x ^ (x >> 27)) * 1664525) & M32

    def f2(x):
        return ((x ^ (x >> 27)) * 1566083941) & M32

    kl = len(key)
    count = max(kl + 1, size)
    r = f1(s[0] ^ s[mid % size] ^ s[(size - 1) % size])
    s[mid % size] = (s[mid % size] + r) & M32
    r = (r + kl) & M32
    s[(mid + lag) % size] = (s[(mid + lag) % size] + r) & M32
    s[0] = r
    count -= 1
    i, j = 1, 0
    while j < count:
        r = f1(s[i] ^ s[(i + mid) % size] ^ s[(i + size - 1) % size])
        s[(i + mid) % size] = (s[(i + mid) % size] + r) & M32
        r = (r + (key[j] if j < kl else 0) + i) & M32
        s[(i + mid + lag) % size] = (s[(i + mid + lag) % size] + r) & M32
        s[i] = r
        i = (i + 1) % size
        j += 1
    for _ in range(size):
        r = f2((s[i] + s[(i + mid) % size] + s[(i + size - 1) % size]) & M32)
        s[(i + mid) % size] ^= r
        r = (r - i) & M32
        s[(i + mid + lag) % size] ^= r
        s[i] = r
        i = (i + 1) % size
    u = [s[2 * k] | (s[2 * k + 1] << 32) for k in range(2 * (N + 1))]
    for k in range(2 * N):
        u[k] = (u[k] & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000
    inner = ((u[2 * N] ^ FIX1) & PCV1) ^ ((u[2 * N + 1] ^ FIX2) & PCV2)
    sh = 32
    while sh > 0:
        inner ^= inner >> sh
        sh >>= 1
    if inner & 1 == 0:
        u[2 * N + 1] ^= 1
    return u


def gen_rand_all(u: list[int]) -> None:
    """dSFMT.c gen_rand_all / do_recursion: regenerate the 2N doubles in place."""
    l0, l1 = u[2 * N], u[2 * N + 1]
    for i in range(N):
        b = (i + POS1) % N
        t0, t1 = u[2 * i], u[2 * i + 1]
        n0 = ((t0 << SL1) & M64) ^ (l1 >> 32) ^ ((l1 << 32) & M64) ^ u[2 * b]
        n1 = ((t1 << SL1) & M64) ^ (l0 >> 32) ^ ((l0 << 32) & M64) ^ u[2 * b + 1]
        l0, l1 = n0, n1
        u[2 * i] = (l0 >> SR) ^ (l0 & MSK1) ^ t0
        u[2 * i + 1] = (l1 >> SR) ^ (l1 & MSK2) ^ t1
    u[2 * N], u[2 * N + 1] = l0, l1


def julia_rand(seed: int, n: int, skip: int = 0) -> np.ndarray:
    """First n Float64 values, after `skip`, of `Random.seed!(seed); rand()`."""
    u = init_by_array(make_seed(seed))
    out = np.empty(skip + n, dtype=np.float64)
    pos = 0
    while pos < skip + n:
        gen_rand_all(u)
        m = min(2 * N, skip + n - pos)
        block = np.frombuffer(struct.pack("<%dQ" % (2 * N), *u[:2 * N]), dtype="<f8")
        out[pos:pos + m] = block[:m] - 1.0
        pos += m
    return out[skip:]
