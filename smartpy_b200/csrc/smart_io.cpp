// smart_io.cpp -- bulk text I/O of the Monte-Carlo sample database (include/smart_b200_io.h).
//
// Host code only.  The reference prints one line per sample with Python's '%.6e' of every value's
// float32 rounding (smartpy/montecarlo/montecarlo.py:211-231) and reads the file back through
// csv.DictReader (:233-262); here blocks of rows are formatted / parsed on all host cores.  The C
// library's "%.6e" and strtod are correctly rounded, like Python's float formatting and float():
// the bytes written and the float32 values read are the reference's (tests/test_host_logic.py
// compares them with numpy.savetxt / numpy.loadtxt).
#include "smart_b200_io.h"
#include "smart_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>

int smart_internal_fail(int code, const char *msg);   // error channel of the library (smart_kernels.cu)

namespace {

constexpr int kMaxChars = 14;   // "-d.dddddde+dd" and one separator

int worker_count(int32_t asked, int64_t n_rows)
{
    int n = asked > 0 ? asked : static_cast<int>(std::thread::hardware_concurrency());
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    const int64_t by_rows = (n_rows + 1023) / 1024;        // a thread is not worth less than ~1000 rows
    if (n > by_rows) n = static_cast<int>(by_rows < 1 ? 1 : by_rows);
    return n;
}

// 10^0 .. 10^22 are exact in binary64; the larger ones carry half an ulp of error
const double kPow10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11, 1e12, 1e13,
                         1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22, 1e23, 1e24, 1e25, 1e26, 1e27,
                         1e28, 1e29, 1e30, 1e31, 1e32, 1e33, 1e34, 1e35, 1e36, 1e37, 1e38, 1e39, 1e40, 1e41,
                         1e42, 1e43, 1e44, 1e45, 1e46, 1e47, 1e48, 1e49, 1e50, 1e51, 1e52};

// One value as Python's '%.6e' % float(v) prints it (correctly rounded, ties to even); returns the
// number of characters.  Fast path: a float32 has 24 significant bits, so v * 10^(6 - k) computed in
// binary64 (k = its decimal exponent; at most four roundings, 5e-16 relative) rounds to the right
// seven digits unless it lies within 1e-13 relative of a tie -- exact ties do occur (1234567.5 is a
// float32) -- and those go to the C library's exact conversion.
inline int format_value(float v, char *dst)
{
    if (v != v) {
        memcpy(dst, "nan", 3);
        return 3;
    }
    if (isinf(v)) {
        if (v < 0) {
            memcpy(dst, "-inf", 4);
            return 4;
        }
        memcpy(dst, "inf", 3);
        return 3;
    }
    char *p = dst;
    double x = static_cast<double>(v);
    if (signbit(v)) {
        *p++ = '-';
        x = -x;
    }
    if (x == 0.0) {
        memcpy(p, "0.000000e+00", 12);
        return static_cast<int>(p - dst) + 12;
    }
    int e2;
    frexp(x, &e2);                                           // x = f * 2^e2, 0.5 <= f < 1
    int k = static_cast<int>(floor((e2 - 1) * 0.30102999566398120));   // floor(log10(2^(e2-1))) <= log10(x)
    // (x is below 2^e2 <= 10^(k + 1.31): at most one step up; the table values beyond 10^22 are
    // rounded, but no float32 lies within a binary64 ulp of a power of ten it is not equal to)
    while (k < 38 && x >= (k + 1 >= 0 ? kPow10[k + 1] : 1.0 / kPow10[-(k + 1)])) ++k;
    const int q = 6 - k;                                     // scale to [1e6, 1e7)
    double y;
    if (q >= 0) y = q <= 22 ? x * kPow10[q] : (x * kPow10[22]) * kPow10[q - 22];
    else y = x / kPow10[-q];
    double d = nearbyint(y);                                 // ties to even (default rounding mode)
    const double frac = fabs(y - floor(y) - 0.5);
    if (frac < y * 1e-13 || !(y >= 999999.0 && y <= 10000001.0)) {   // too close to a tie (or k misjudged): exact path
        const int n = snprintf(p, kMaxChars + 2, "%.6e", x);
        return static_cast<int>(p - dst) + n;
    }
    if (d >= 10000000.0) {                                   // 9.9999996 -> 1.000000e+01
        d = 1000000.0;
        ++k;
    } else if (d < 1000000.0) {                              // (k one too high: cannot happen with the loop above)
        const int n = snprintf(p, kMaxChars + 2, "%.6e", x);
        return static_cast<int>(p - dst) + n;
    }
    unsigned digits = static_cast<unsigned>(d);
    char buf[7];
    for (int i = 6; i >= 0; --i) {
        buf[i] = static_cast<char>('0' + digits % 10);
        digits /= 10;
    }
    *p++ = buf[0];
    *p++ = '.';
    memcpy(p, buf + 1, 6);
    p += 6;
    *p++ = 'e';
    int ek = k;
    if (ek < 0) {
        *p++ = '-';
        ek = -ek;
    } else {
        *p++ = '+';
    }
    *p++ = static_cast<char>('0' + ek / 10);
    *p++ = static_cast<char>('0' + ek % 10);
    return static_cast<int>(p - dst);
}

// One number of a database line.  Fast path for the format the writer produces, [-]d.dddddde[+-]dd:
// seven digits and a power of ten of at most 22 are both exact in binary64, so one multiplication or
// division gives the correctly rounded value strtod would (Clinger); anything else goes to strtod.
inline double parse_value(char *p, char **end)
{
    char *s = p;
    bool neg = false;
    if (*s == '-') {
        neg = true;
        ++s;
    }
    if (s[0] >= '0' && s[0] <= '9' && s[1] == '.' && s[8] == 'e' && (s[9] == '+' || s[9] == '-') &&
        s[10] >= '0' && s[10] <= '9' && s[11] >= '0' && s[11] <= '9' && !(s[12] >= '0' && s[12] <= '9')) {
        unsigned m = static_cast<unsigned>(s[0] - '0');
        bool ok = true;
        for (int i = 2; i < 8; ++i) {
            ok = ok && s[i] >= '0' && s[i] <= '9';
            m = m * 10 + static_cast<unsigned>(s[i] - '0');
        }
        const int e = (s[10] - '0') * 10 + (s[11] - '0');
        const int q = (s[9] == '-' ? -e : e) - 6;            // value = m * 10^q
        if (ok && q >= -22 && q <= 22) {
            const double val = q >= 0 ? m * kPow10[q] : m / kPow10[-q];
            *end = s + 12;
            return neg ? -val : val;
        }
    }
    return strtod(p, end);
}

}  // namespace

extern "C" {

int64_t smart_csv_bound(int64_t n_rows, int32_t n_cols)
{
    if (n_rows < 0 || n_cols < 1) return 0;
    return n_rows * static_cast<int64_t>(n_cols) * kMaxChars;
}

int64_t smart_csv_format_f32(const float *table, int64_t n_rows, int32_t n_cols, int64_t ld, char *out,
                             int64_t out_cap, int32_t n_threads)
{
    if (!table || !out || n_rows < 0 || n_cols < 1 || ld < n_cols)
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_csv_format_f32: bad argument");
    if (out_cap < smart_csv_bound(n_rows, n_cols))
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_csv_format_f32: out_cap below smart_csv_bound()");
    if (n_rows == 0) return 0;
    const int workers = worker_count(n_threads, n_rows);
    const int64_t per = (n_rows + workers - 1) / workers;
    const int64_t row_cap = static_cast<int64_t>(n_cols) * kMaxChars;
    // Every worker formats its rows where the longest possible text of the rows before them would
    // end (so no two workers ever touch the same bytes), then the pieces are closed up in order.
    std::vector<int64_t> length(workers, 0);
    std::vector<std::thread> pool;
    auto work = [&](int w) {
        const int64_t r0 = w * per, r1 = std::min(n_rows, r0 + per);
        char *dst = out + r0 * row_cap;
        char *p = dst;
        char tmp[kMaxChars + 4];
        for (int64_t r = r0; r < r1; ++r) {
            const float *row = table + r * ld;
            for (int32_t c = 0; c < n_cols; ++c) {
                const int n = format_value(row[c], tmp);
                memcpy(p, tmp, n);
                p += n;
                *p++ = c + 1 < n_cols ? ',' : '\n';
            }
        }
        length[w] = p - dst;
    };
    for (int w = 1; w < workers; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto &t : pool) t.join();
    int64_t total = length[0];
    for (int w = 1; w < workers; ++w) {
        memmove(out + total, out + w * per * row_cap, static_cast<size_t>(length[w]));
        total += length[w];
    }
    return total;
}

int64_t smart_csv_parse_f32(const char *text, int64_t n_bytes, int32_t n_cols_in_file, const int32_t *wanted,
                            int32_t n_wanted, float *out, int64_t max_rows, int32_t n_threads)
{
    if (!text || !wanted || !out || n_bytes < 0 || n_cols_in_file < 1 || n_wanted < 1 || max_rows < 0)
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_csv_parse_f32: bad argument");
    for (int32_t k = 0; k < n_wanted; ++k)
        if (wanted[k] < 0 || wanted[k] >= n_cols_in_file)
            return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_csv_parse_f32: wanted column out of range");
    // line starts (a last line without '\n' counts; empty lines end the data)
    std::vector<int64_t> start;
    int64_t pos = 0;
    while (pos < n_bytes && static_cast<int64_t>(start.size()) < max_rows) {
        const char *nl = static_cast<const char *>(memchr(text + pos, '\n', static_cast<size_t>(n_bytes - pos)));
        const int64_t end = nl ? nl - text : n_bytes;
        if (end == pos || (end == pos + 1 && text[pos] == '\r')) break;
        start.push_back(pos);
        pos = end + 1;
    }
    const int64_t n_rows = static_cast<int64_t>(start.size());
    if (n_rows == 0) return 0;
    start.push_back(pos);                                   // one past the last line
    const int workers = worker_count(n_threads, n_rows);
    const int64_t per = (n_rows + workers - 1) / workers;
    std::vector<int> bad(workers, 0);
    std::vector<std::thread> pool;
    auto work = [&](int w) {
        const int64_t r0 = w * per, r1 = std::min(n_rows, r0 + per);
        std::vector<char> line;
        std::vector<double> value(n_cols_in_file);
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t len = start[r + 1] - start[r];
            line.assign(text + start[r], text + start[r] + len);
            line.insert(line.end(), 16, '\0');              // strtod needs a terminated string (and the fast path looks 12 ahead)
            char *p = line.data();
            for (int32_t c = 0; c < n_cols_in_file; ++c) {
                char *end = nullptr;
                value[c] = parse_value(p, &end);
                if (end == p) {
                    bad[w] = 1;
                    return;
                }
                p = end;
                while (*p == ' ' || *p == '\r') ++p;
                if (c + 1 < n_cols_in_file) {
                    if (*p != ',') {
                        bad[w] = 1;
                        return;
                    }
                    ++p;
                } else if (*p != '\n' && *p != '\0') {
                    bad[w] = 1;
                    return;
                }
            }
            float *dst = out + r * n_wanted;
            for (int32_t k = 0; k < n_wanted; ++k) dst[k] = static_cast<float>(value[wanted[k]]);
        }
    };
    for (int w = 1; w < workers; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto &t : pool) t.join();
    for (int w = 0; w < workers; ++w)
        if (bad[w])
            return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_csv_parse_f32: a line does not hold the expected numbers");
    return n_rows;
}

}  // extern "C"
