// Host-side permutation source of the reference stream (ref:975-976): numpy.random.shuffle applied
// row after row, drawn from the GLOBAL legacy MT19937 state.  numpy's legacy shuffle of a 1-D
// array is Fisher-Yates from the top, j = interval(i) by masked rejection on ONE 32-bit tempered
// MT19937 output per trial.  This is the same generator and the same
// rejection rule in C, so that a [rows, S] block of index arrays costs a few nanoseconds per
// draw instead of a Python-level call per row; the caller hands the state in and out
// (numpy.random.get_state / set_state), so the stream continues exactly where numpy would be.
// tests/test_host_cpu.py checks it draw for draw against numpy.
#include <stdint.h>

#include "../../include/sctda_b200.h"

namespace {

struct MT {
    uint32_t* key;   // 624 words
    int pos;
};

inline void mt_gen(MT& s) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    uint32_t* mt = s.key;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    for (; kk < 623; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    s.pos = 0;
}

inline uint32_t mt_next(MT& s) {
    if (s.pos == 624) mt_gen(s);
    uint32_t y = s.key[s.pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

}  // namespace

extern "C" int32_t sctda_mt19937_shuffle_rows(uint32_t* key624, int32_t* pos, int32_t rows, int32_t S, int32_t* out) {
    if (!key624 || !pos || rows < 0 || S < 0 || (rows > 0 && S > 0 && !out) || *pos < 0 || *pos > 624) return SCTDA_ERR_INVALID;
    MT s{key624, *pos};
    for (int32_t r = 0; r < rows; ++r) {
        int32_t* row = out + (int64_t)r * S;
        for (int32_t i = 0; i < S; ++i) row[i] = i;
        for (int32_t i = S - 1; i >= 1; --i) {
            uint32_t mask = (uint32_t)i;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            uint32_t v;
            do { v = mt_next(s) & mask; } while (v > (uint32_t)i);
            const int32_t t = row[i]; row[i] = row[v]; row[v] = t;
        }
    }
    *pos = s.pos;
    return SCTDA_OK;
}
