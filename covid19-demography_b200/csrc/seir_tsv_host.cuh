// seir_tsv_host.cuh -- the bytes pandas' to_csv(sep="\t", index=False, header=False, na_rep="NA") writes for a float64 matrix
// (parameter_sweep.py:716-729 writes 19 such files per job, simulate_agepolicies.py:646-716 twelve), host only.
//
// pandas prints a float64 as Python's repr: the SHORTEST digit string that reads back as the same double, laid out in
// fixed notation while the decimal point position decpt (value = 0.d1d2.. x 10^decpt) satisfies -4 < decpt <= 16 -- with ".0"
// appended to an integer value -- and as d.ddde+XX (two exponent digits at least) otherwise (CPython, format_float_short, code 'r').
// std::to_chars yields the shortest digits; the layout is done here.  NaN is "NA", infinities "inf" / "-inf".
// A sweep writes thousands of jobs of a few thousand numbers each: formatted in Python this was more than the device run of the job.
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace seir_host {

// writes repr(x) at p (at most 26 bytes), returns the end
inline char *float_repr(double x, char *p)
{
    if (std::isnan(x)) { memcpy(p, "NA", 2); return p + 2; }
    if (std::isinf(x)) {
        if (x < 0) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    char sci[40];
    const auto res = std::to_chars(sci, sci + sizeof sci, x, std::chars_format::scientific);  // [-]d[.ddd]e[+-]XX, shortest
    const char *s = sci, *end = res.ptr;
    if (*s == '-') { *p++ = '-'; ++s; }
    char digits[24];
    int nd = 0;
    const char *q = s;
    for (; q < end && *q != 'e'; ++q)
        if (*q != '.') digits[nd++] = *q;
    int e10 = 0;
    {
        const char *r = q + 1;  // behind the 'e'
        const bool neg = *r == '-';
        if (*r == '-' || *r == '+') ++r;
        for (; r < end; ++r) e10 = e10 * 10 + (*r - '0');
        if (neg) e10 = -e10;
    }
    const int decpt = e10 + 1;
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {                       // 0.000ddd
            *p++ = '0'; *p++ = '.';
            for (int k = 0; k < -decpt; ++k) *p++ = '0';
            memcpy(p, digits, (size_t)nd); p += nd;
        } else if (nd <= decpt) {               // ddd000.0
            memcpy(p, digits, (size_t)nd); p += nd;
            for (int k = nd; k < decpt; ++k) *p++ = '0';
            *p++ = '.'; *p++ = '0';
        } else {                                // dd.ddd
            memcpy(p, digits, (size_t)decpt); p += decpt;
            *p++ = '.';
            memcpy(p, digits + decpt, (size_t)(nd - decpt)); p += nd - decpt;
        }
    } else {                                    // d.ddde+XX
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, digits + 1, (size_t)(nd - 1)); p += nd - 1; }
        *p++ = 'e';
        int e = decpt - 1;
        if (e < 0) { *p++ = '-'; e = -e; } else *p++ = '+';
        if (e >= 100) { *p++ = (char)('0' + e / 100); e %= 100; *p++ = (char)('0' + e / 10); *p++ = (char)('0' + e % 10); }
        else { *p++ = (char)('0' + e / 10); *p++ = (char)('0' + e % 10); }
    }
    return p;
}

}  // namespace seir_host

extern "C" int seir_format_tsv(const double *a, int64_t rows, int64_t cols, char *out, int64_t cap, int64_t *written)
{
    if (rows < 0 || cols < 0 || (rows * cols > 0 && !a) || !out || !written) return fail("seir_format_tsv: bad argument");
    if (cap < rows * cols * 27 + rows + 1) return fail("seir_format_tsv: the output buffer is too small (27 bytes per value + 1 per row)");
    char *p = out;
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t c = 0; c < cols; ++c) {
            if (c) *p++ = '\t';
            p = seir_host::float_repr(a[r * cols + c], p);
        }
        *p++ = '\n';
    }
    *written = (int64_t)(p - out);
    return 0;
}
