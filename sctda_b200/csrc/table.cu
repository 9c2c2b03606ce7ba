// Host-side table ingest: the numeric fields of `<t>.all.tsv` /
// `<t>.mapper.tsv`.  The reference parses the table with one Python float() call per field
// (/root/reference/scTDA/main.py:839-885: 2e9 calls at 1e5 cells x 2e4 genes); here the body of the
// file is split into rows and fields and converted by all host cores.
//
// The parser is CONSERVATIVE: it converts only fields of the plain form
//     [+-] digits [. digits] [ (e|E) [+-] digits ]           (at least one digit in the mantissa)
// and reports every other field (identifiers, library names, empty fields, nan / inf, digit
// separators, surrounding blanks, ...) back to the caller by position, so that Python's float()
// decides them exactly as the reference does.  Conversions are correctly rounded: up to 19
// significant digits with a decimal exponent of magnitude <= 22 go through the exact fast path
// (an integer below 2^53 times or divided by an exactly representable power of ten is one rounding),
// everything else through strtod.
//
// Output is gene-major, Y[column][row] (the image of `dicgenes`): rows are parsed 32 at a time into
// a tile and written out transposed, so that every column receives runs of 32 doubles.
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/sctda_b200.h"

namespace {

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// true and *out set when [p, e) is a plain decimal number; false for anything else
inline bool parse_plain(const char* p, const char* e, double* out) {
    const char* s = p;
    if (s == e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') { neg = (*s == '-'); ++s; }
    uint64_t mant = 0;
    int nd = 0, dropped = 0, frac = 0;
    bool any = false, inexact = false;
    for (; s < e && *s >= '0' && *s <= '9'; ++s) {
        any = true;
        if (nd < 19) { mant = mant * 10 + (uint64_t)(*s - '0'); if (mant) ++nd; }
        else { ++dropped; inexact |= (*s != '0'); }
    }
    if (s < e && *s == '.') {
        ++s;
        for (; s < e && *s >= '0' && *s <= '9'; ++s) {
            any = true;
            if (nd < 19) { mant = mant * 10 + (uint64_t)(*s - '0'); if (mant) ++nd; ++frac; }
            else inexact |= (*s != '0');
        }
    }
    if (!any) return false;
    int ex = 0;
    if (s < e && (*s == 'e' || *s == 'E')) {
        ++s;
        bool eneg = false;
        if (s < e && (*s == '+' || *s == '-')) { eneg = (*s == '-'); ++s; }
        if (s == e) return false;
        int nde = 0;
        for (; s < e && *s >= '0' && *s <= '9'; ++s) { if (nde < 6) { ex = ex * 10 + (*s - '0'); ++nde; } else return false; }
        if (eneg) ex = -ex;
    }
    if (s != e) return false;
    const int e10 = ex + dropped - frac;
    if (!inexact && mant < (1ull << 53) && e10 >= -22 && e10 <= 22) {
        double v = (double)mant;
        v = e10 >= 0 ? v * kPow10[e10] : v / kPow10[-e10];
        *out = neg ? -v : v;
        return true;
    }
    if (mant == 0 && !inexact) { *out = neg ? -0.0 : 0.0; return true; }
    const size_t len = (size_t)(e - p);                       // slow path: correctly rounded by the C library
    char tmp[128];
    if (len >= sizeof(tmp)) return false;
    memcpy(tmp, p, len);
    tmp[len] = 0;
    char* endp = nullptr;
    const double v = strtod(tmp, &endp);
    if (endp != tmp + len) return false;
    *out = v;
    return true;
}

}  // namespace

extern "C" int32_t sctda_parse_table_f64(const char* body, int64_t nbytes, int32_t sep, int32_t strip_bar, int64_t nrows,
                                        int64_t ncols, double* out, int64_t* bad_pos, int32_t* bad_len, int64_t bad_cap,
                                        int64_t* n_bad, int64_t* bad_row, int32_t nthreads) {
    if (!body || nbytes < 0 || nrows < 0 || ncols < 1 || !out || !n_bad || !bad_row || (bad_cap > 0 && (!bad_pos || !bad_len)))
        return SCTDA_ERR_INVALID;
    *n_bad = 0;
    *bad_row = -1;
    if (nrows == 0) return SCTDA_OK;
    // row starts
    std::vector<int64_t> start((size_t)nrows + 1);
    {
        int64_t r = 0, pos = 0;
        while (r < nrows && pos <= nbytes) {
            start[(size_t)r++] = pos;
            const char* nl = (const char*)memchr(body + pos, '\n', (size_t)(nbytes - pos));
            pos = nl ? (int64_t)(nl - body) + 1 : nbytes + 1;
        }
        if (r < nrows) { *bad_row = r; return SCTDA_ERR_INVALID; }           // fewer rows than announced
        start[(size_t)nrows] = pos;
    }
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    constexpr int64_t RB = 32;                                              // rows per tile
    const int64_t nblocks = (nrows + RB - 1) / RB;
    if ((int64_t)nt > nblocks) nt = (int)nblocks;
    std::atomic<int64_t> next(0), bad_count(0), first_bad_row(-1);
    auto work = [&]() {
        std::vector<double> tile((size_t)(RB * ncols));
        std::vector<int64_t> lpos;                                          // this block's unparsed fields, row-major
        std::vector<int32_t> llen;
        for (;;) {
            const int64_t blk = next.fetch_add(1);
            if (blk >= nblocks) break;
            const int64_t r0 = blk * RB, r1 = (r0 + RB < nrows) ? r0 + RB : nrows;
            lpos.clear();
            llen.clear();
            for (int64_t r = r0; r < r1; ++r) {
                const char* p = body + start[(size_t)r];
                const char* le = body + start[(size_t)r + 1] - 1;           // the '\n' (or one past the buffer)
                if (le > body + nbytes) le = body + nbytes;
                if (le > p && le[-1] == '\r') --le;                          // "\r\n"
                double* trow = tile.data() + (size_t)((r - r0) * ncols);
                int64_t c = 0;
                bool ended = false;
                while (c < ncols) {
                    const char* fe = (const char*)memchr(p, sep, (size_t)(le - p));
                    if (!fe) fe = le;
                    const char* ne = fe;                                     // end of the numeric part
                    if (strip_bar) {
                        const char* bar = (const char*)memchr(p, '|', (size_t)(fe - p));
                        if (bar) ne = bar;
                    }
                    double v = 0.0;
                    if (!parse_plain(p, ne, &v)) {
                        v = 0.0;
                        lpos.push_back((int64_t)(r * ncols + c));
                        lpos.push_back((int64_t)(p - body));
                        llen.push_back((int32_t)(ne - p));
                    }
                    trow[c++] = v;
                    if (fe == le) { ended = true; break; }
                    p = fe + 1;
                }
                if (c != ncols || !ended) {                                  // too few or too many fields
                    int64_t cur = first_bad_row.load();
                    while ((cur < 0 || r < cur) && !first_bad_row.compare_exchange_weak(cur, r)) {}
                    for (; c < ncols; ++c) trow[c] = 0.0;
                }
            }
            const int64_t nr = r1 - r0;
            for (int64_t c = 0; c < ncols; ++c) {                            // transposed write-out
                double* dst = out + (size_t)(c * nrows + r0);
                for (int64_t i = 0; i < nr; ++i) dst[i] = tile[(size_t)(i * ncols + c)];
            }
            if (!llen.empty()) {
                const int64_t n = (int64_t)llen.size();
                const int64_t at = bad_count.fetch_add(n);
                for (int64_t i = 0; i < n; ++i) {
                    if (at + i < bad_cap) {
                        bad_pos[2 * (at + i)] = lpos[(size_t)(2 * i)];        // (flat field index, byte offset)
                        bad_pos[2 * (at + i) + 1] = lpos[(size_t)(2 * i + 1)];
                        bad_len[at + i] = llen[(size_t)i];
                    }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    *n_bad = bad_count.load();
    *bad_row = first_bad_row.load();
    if (*bad_row >= 0) return SCTDA_ERR_INVALID;
    return SCTDA_OK;
}

// ------------------------------------------------------------------------------------------------
// Table egress: the numeric fields of Preprocess.save (/root/reference/scTDA/main.py:663-728), i.e.
// str(numpy.log2(1.0 + 1000000.0 * t / cv)) under Python 2 for every (cell, gene): 12 significant
// digits ('%.12g'), '.0' appended when the result has neither '.' nor an exponent, 'nan' / 'inf' /
// '-inf' as Python prints them.  Rows are formatted by all host cores into per-row byte ranges.
// ------------------------------------------------------------------------------------------------
static inline int format_py2(double v, char* out) {                    // at most 24 bytes, no terminator needed
    if (v != v) { memcpy(out, "nan", 3); return 3; }
    if (v > 1.7976931348623157e308) { memcpy(out, "inf", 3); return 3; }
    if (v < -1.7976931348623157e308) { memcpy(out, "-inf", 4); return 4; }
    char buf[40];
    int n = snprintf(buf, sizeof(buf), "%.12g", v);
    bool plain = true;
    for (int i = 0; i < n; ++i)
        if (buf[i] == '.' || buf[i] == 'e') { plain = false; break; }
    memcpy(out, buf, (size_t)n);
    if (plain) { out[n] = '.'; out[n + 1] = '0'; n += 2; }
    return n;
}

extern "C" int32_t sctda_format_table_f64(const double* counts, int64_t ld, int64_t nrows, int64_t ncols,
                                          const double* denom, int32_t log2_tpm, int32_t sep, char* out, int64_t cap,
                                          int64_t* row_end, int32_t nthreads) {
    if (!counts || nrows < 0 || ncols < 1 || ld < ncols || !out || !row_end || cap < nrows * ncols * 25) return SCTDA_ERR_INVALID;
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if ((int64_t)nt > nrows) nt = (int)(nrows > 0 ? nrows : 1);
    // every row owns the fixed slot [r * ncols * 25, (r + 1) * ncols * 25) of `out`; row_end[r] = bytes used
    // (the caller concatenates prefix + slot[:row_end[r]] + '\n'); 24 bytes per value + separator
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t r = next.fetch_add(1);
            if (r >= nrows) break;
            char* p = out + (size_t)(r * ncols * 25);
            char* p0 = p;
            const double* row = counts + (size_t)(r * ld);
            const double cv = denom ? denom[r] : 1.0;
            for (int64_t c = 0; c < ncols; ++c) {
                double v = row[c];
                if (log2_tpm) v = log2(1.0 + 1000000.0 * v / cv);           // ref:692, 726: the operation order of the reference
                p += format_py2(v, p);
                if (c + 1 < ncols) *p++ = (char)sep;
            }
            row_end[r] = (int64_t)(p - p0);
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return SCTDA_OK;
}

