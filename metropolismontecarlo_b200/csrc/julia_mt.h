// julia_mt.h — the reference's random stream without Julia (host code, part of the driver stand-in).
//
// The reference draws every uniform from Julia's global RNG (`rand()`, `rand(Float64,3)`:
// Ewald/main.jl:516, Ewald/auxillary.jl:99,109, Ewald/quaternions.jl:64,176) and names its seed at
// Ewald/main.jl:36 / Monatomic/mainMonatomic.jl:15 (`Random.seed!(11234)`).  In the Julia the
// reference targets (1.x ≤ 1.6) that RNG is `MersenneTwister` = dSFMT-19937 (Saito & Matsumoto),
// seeded by `dsfmt_init_by_array(make_seed(n))` where make_seed splits n into 32-bit words, and
// `rand()` returns the next close1_open2 double minus 1.0.  Julia refills a cache with
// dsfmt_fill_array_close1_open2 and the small-array path of `rand(Float64,3)` copies consecutive
// cache entries, so the stream of Float64 uniforms is exactly the sequential dSFMT output.
// Pinned by the values Julia's own manual prints for MersenneTwister(1234) (tests/test_oracle.py).
#pragma once
#include <cstdint>
#include <cstring>

namespace julia_mt {

constexpr int N = 191;                      // 128-bit words of state (+1 "lung")
constexpr int POS1 = 117, SL1 = 19, SR = 12;
constexpr uint64_t MSK1 = 0x000ffafffffffb3fULL, MSK2 = 0x000ffdfffc90fffdULL;
constexpr uint64_t FIX1 = 0x90014964b32f4329ULL, FIX2 = 0x3b8d12ac548a7c7aULL;
constexpr uint64_t PCV1 = 0x3d84e1ac0dc82880ULL, PCV2 = 0x0000000000000001ULL;

struct State {
    uint64_t u[2 * (N + 1)];
    int idx;                                // next unread double in u[0 .. 2N)
};

inline uint32_t mix1(uint32_t x) { return (x ^ (x >> 27)) * 1664525u; }
inline uint32_t mix2(uint32_t x) { return (x ^ (x >> 27)) * 1566083941u; }

inline void seed(State &s, uint64_t seed_value)
{
    uint32_t key[2];
    int key_len = 0;
    do { key[key_len++] = (uint32_t)(seed_value & 0xffffffffu); seed_value >>= 32; } while (seed_value != 0);
    constexpr int size = (N + 1) * 4, lag = 11, mid = (size - lag) / 2;   // size = 768 >= 623
    uint32_t w[size];
    std::memset(w, 0x8b, sizeof(w));
    int count = (key_len + 1 > size) ? key_len + 1 : size;
    uint32_t r = mix1(w[0] ^ w[mid % size] ^ w[(size - 1) % size]);
    w[mid % size] += r;
    r += (uint32_t)key_len;
    w[(mid + lag) % size] += r;
    w[0] = r;
    count--;
    int i = 1, j = 0;
    for (; j < count && j < key_len; ++j) {
        r = mix1(w[i] ^ w[(i + mid) % size] ^ w[(i + size - 1) % size]);
        w[(i + mid) % size] += r;
        r += key[j] + (uint32_t)i;
        w[(i + mid + lag) % size] += r;
        w[i] = r;
        i = (i + 1) % size;
    }
    for (; j < count; ++j) {
        r = mix1(w[i] ^ w[(i + mid) % size] ^ w[(i + size - 1) % size]);
        w[(i + mid) % size] += r;
        r += (uint32_t)i;
        w[(i + mid + lag) % size] += r;
        w[i] = r;
        i = (i + 1) % size;
    }
    for (j = 0; j < size; ++j) {
        r = mix2(w[i] + w[(i + mid) % size] + w[(i + size - 1) % size]);
        w[(i + mid) % size] ^= r;
        r -= (uint32_t)i;
        w[(i + mid + lag) % size] ^= r;
        w[i] = r;
        i = (i + 1) % size;
    }
    for (int k = 0; k < 2 * (N + 1); ++k) s.u[k] = (uint64_t)w[2 * k] | ((uint64_t)w[2 * k + 1] << 32);
    for (int k = 0; k < 2 * N; ++k) s.u[k] = (s.u[k] & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;
    uint64_t inner = ((s.u[2 * N] ^ FIX1) & PCV1) ^ ((s.u[2 * N + 1] ^ FIX2) & PCV2);   // period certification
    for (int sh = 32; sh > 0; sh >>= 1) inner ^= inner >> sh;
    if ((inner & 1) == 0) s.u[2 * N + 1] ^= 1;
    s.idx = 2 * N;                          // empty: first draw regenerates the block
}

inline void regenerate(State &s)
{
    uint64_t L0 = s.u[2 * N], L1 = s.u[2 * N + 1];
    for (int i = 0; i < N; ++i) {
        const int b = (i + POS1) % N;
        const uint64_t t0 = s.u[2 * i], t1 = s.u[2 * i + 1];
        const uint64_t n0 = (t0 << SL1) ^ (L1 >> 32) ^ (L1 << 32) ^ s.u[2 * b];
        const uint64_t n1 = (t1 << SL1) ^ (L0 >> 32) ^ (L0 << 32) ^ s.u[2 * b + 1];
        L0 = n0; L1 = n1;
        s.u[2 * i] = (L0 >> SR) ^ (L0 & MSK1) ^ t0;
        s.u[2 * i + 1] = (L1 >> SR) ^ (L1 & MSK2) ^ t1;
    }
    s.u[2 * N] = L0; s.u[2 * N + 1] = L1;
    s.idx = 0;
}

// Julia `rand()` : CloseOpen01 = close1_open2 - 1.0
inline double next(State &s)
{
    if (s.idx >= 2 * N) regenerate(s);
    double d;
    std::memcpy(&d, &s.u[s.idx++], sizeof(d));
    return d - 1.0;
}

}  // namespace julia_mt
