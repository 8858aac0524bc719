// MT19937 streams of the growth path (product code; the oracle has its own independent copy).
//
// The reference draws from two interleaved Mersenne twisters:
//   * Python `random`  : seed(int) = init_by_array(32-bit limbs of |seed|)
//   * legacy numpy RandomState : seed(int < 2^32) = init_genrand(seed)
// both producing doubles as ((a>>5)*2^26 + (b>>6)) / 2^53 from two consecutive 32-bit outputs.
// Host code seeds and consumes the streams during Greenhouse/Forest initialisation; the 624-word
// state + cursor is then handed to the device, which continues the very same streams.
#pragma once
#include <stdint.h>

namespace octa {

struct MTState {
    uint32_t mt[624];
    int32_t idx;      // next word to temper; 624 = regenerate first
    int32_t pad;
};

#if defined(__CUDACC__)
#define OCTA_HD __host__ __device__ __forceinline__
#else
#define OCTA_HD inline
#endif

OCTA_HD uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

OCTA_HD double mt_double(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

inline void mt_init_genrand(MTState& s, uint32_t seed) {
    s.mt[0] = seed;
    for (int i = 1; i < 624; ++i) s.mt[i] = 1812433253u * (s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) + (uint32_t)i;
    s.idx = 624;
    s.pad = 0;
}

inline void mt_init_by_array(MTState& s, const uint32_t* key, int len) {
    mt_init_genrand(s, 19650218u);
    int i = 1, j = 0;
    for (int k = (624 > len ? 624 : len); k; --k) {
        s.mt[i] = (s.mt[i] ^ ((s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= 624) { s.mt[0] = s.mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (int k = 623; k; --k) {
        s.mt[i] = (s.mt[i] ^ ((s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= 624) { s.mt[0] = s.mt[623]; i = 1; }
    }
    s.mt[0] = 0x80000000u;
    s.idx = 624;
}

inline void mt_regen_host(MTState& s) {
    for (int k = 0; k < 624; ++k) {
        uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[(k + 1) % 624] & 0x7fffffffu);
        s.mt[k] = s.mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s.idx = 0;
}

inline uint32_t mt_next_host(MTState& s) {
    if (s.idx >= 624) mt_regen_host(s);
    return mt_temper(s.mt[s.idx++]);
}

inline double mt_next_double_host(MTState& s) {
    uint32_t a = mt_next_host(s), b = mt_next_host(s);
    return mt_double(a, b);
}

}  // namespace octa
