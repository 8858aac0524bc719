// Cell formatters of the graph CSV shared by the host writer (octa_csv.cu) and the device writer (octa_csv_dev.cu): integer
// arithmetic only (128-bit products of the binary significand with powers of ten), no libc, so host and device produce the
// same bytes.  What the cells must look like is described at the top of octa_csv.cu (numpy array2string defaults for the two
// position cells, Python's repr(float) for the radius).  Every function returns the number of characters written, or -1 when
// the value is outside the range it covers -- the caller then uses the generic host formatter (octa_csv.cu), which handles
// everything.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define OCTA_CF_HD __host__ __device__ __forceinline__
#else
#define OCTA_CF_HD inline
#endif

namespace octa {
namespace csvfmt {

typedef unsigned __int128 u128;

OCTA_CF_HD uint64_t dbits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t b;
    __builtin_memcpy(&b, &x, 8);
    return b;
#endif
}
OCTA_CF_HD bool dsign(double x) { return (dbits(x) >> 63) != 0; }

OCTA_CF_HD u128 pow10_128(int k) {          // 10^k, 0 <= k <= 38
    u128 r = 1;
    for (int i = 0; i < k; ++i) r *= 10u;
    return r;
}

// |x| * 10^8 rounded to the nearest integer, ties to even, EXACTLY: the digits numpy's Dragon4 prints for precision=8 in
// positional mode (see octa_csv.cu).  Valid for 2^-60 < |x| < 2^20; false otherwise.
OCTA_CF_HD bool fixed8(double x, uint64_t* q_out) {
    const double ax = fabs(x);
    if (!(ax < 1048576.0)) return false;
    if (ax == 0.0) { *q_out = 0; return true; }
    {
        // Fast path.  p = fl(ax * 1e8) differs from the exact product T by at most ulp(p)/2 <= p * 2^-53, and p - floor(p) is exact
        // (p < 2^47).  round-half-even(T) can differ from rounding p only if p lies within that error of a half-integer.
        const double p = ax * 1e8;
        const double fl = floor(p);
        const double fr = p - fl;
        if (fabs(fr - 0.5) > p * 1.2e-16 + 1e-300) { *q_out = (uint64_t)fl + (fr > 0.5 ? 1u : 0u); return true; }
    }
    const uint64_t bits = dbits(ax);
    const int be = (int)(bits >> 52);
    if (be == 0) return false;                             // subnormal: generic path
    const uint64_t m = (bits & 0xfffffffffffffull) | (1ull << 52);   // ax = m * 2^(be-1075)
    const int s = 1075 - be;                               // ax = m / 2^s
    if (s <= 0 || s > 113) return false;
    const u128 N = (u128)m * 100000000u;
    const u128 one = 1;
    u128 q = N >> s;
    const u128 rem = N & ((one << s) - 1), half = one << (s - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
    *q_out = (uint64_t)q;
    return true;
}

// integer / fraction digit strings of q = round(|x| * 1e8); fraction trimmed of trailing zeros (8 characters are written)
OCTA_CF_HD void fixed8_digits(uint64_t q, bool neg, char* ip, int* ilen, char* fp, int* flen) {
    uint64_t ipart = 0;
    uint32_t f = (uint32_t)q;
    if (q >= 100000000u) { ipart = q / 100000000u; f = (uint32_t)(q - ipart * 100000000u); }
    int k = 0;
    if (neg) ip[k++] = '-';
    if (ipart < 10) ip[k++] = (char)('0' + ipart);
    else {
        char tmp[24];
        int n = 0;
        do { tmp[n++] = (char)('0' + ipart % 10); ipart /= 10; } while (ipart);
        while (n) ip[k++] = tmp[--n];
    }
    *ilen = k;
    for (int d = 7; d >= 0; --d) { fp[d] = (char)('0' + f % 10u); f /= 10u; }
    int fl = 8;
    while (fl > 0 && fp[fl - 1] == '0') --fl;
    *flen = fl;
}

// 9 significant digits of |x| (d.dddddddd x 10^e10), exactly rounded on the binary value (ties to even): what Dragon4 prints
// for dragon4_scientific(precision=8).  Valid for 1e-11 < |x| < 1e8; false otherwise.
OCTA_CF_HD bool sci9(double x, uint64_t* q_out, int* e10_out) {
    const double ax = fabs(x);
    if (!(ax > 1e-11 && ax < 1e8)) return false;
    const uint64_t bits = dbits(ax);
    const int be = (int)(bits >> 52);
    if (be == 0) return false;
    const uint64_t m = (bits & 0xfffffffffffffull) | (1ull << 52);
    const int s = 1075 - be;                               // ax = m / 2^s
    const int e = be - 1022;                               // ax = fr * 2^e with fr in [0.5, 1)
    int e10 = (int)floor((e - 1) * 0.30102999566398120);   // floor(log10(ax)) or one less
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int k = 8 - e10;                             // scale by 10^k
        if (k < 0 || k > 19 || s <= 0 || s > 120) return false;
        const u128 N = (u128)m * pow10_128(k);
        const u128 one = 1;
        u128 q = N >> s;
        const u128 rem = N & ((one << s) - 1), half = one << (s - 1);
        if (rem > half || (rem == half && (q & 1))) ++q;
        if (q >= 1000000000u) {
            const u128 lo = (u128)1000000000u << s;
            if (N >= lo) { ++e10; continue; }              // genuinely >= 10^(e10+1)
            *q_out = 100000000u; *e10_out = e10 + 1;       // rounding carried into a 10th digit
            return true;
        }
        if (q < 100000000u) { --e10; continue; }
        *q_out = (uint64_t)q; *e10_out = e10;
        return true;
    }
    return false;
}

// str(ndarray float64[3]) -- numpy array2string defaults -- into p (>= 96 bytes free).  -1: generic formatter needed.
OCTA_CF_HD int array3(char* p, const double* v) {
    double mx = 0, mn = 0;
    bool any = false;
    for (int i = 0; i < 3; ++i) {
        if (!(v[i] == v[i]) || fabs(v[i]) > 1.7e308) return -1;
        const double a = fabs(v[i]);
        if (a != 0.0) { if (!any) { mx = mn = a; any = true; } else { if (a > mx) mx = a; if (a < mn) mn = a; } }
    }
    const bool exp_format = any && (mx >= 1.e8 || mn < 0.0001 || mx / mn > 1000.);
    int k = 0;
    if (!exp_format) {
        uint64_t q[3];
        if (!(fixed8(v[0], &q[0]) && fixed8(v[1], &q[1]) && fixed8(v[2], &q[2]))) return -1;
        char ipb[3][24], fpb[3][8];
        int il[3], fl[3], pl = 0, pr = 0;
        for (int i = 0; i < 3; ++i) {
            fixed8_digits(q[i], dsign(v[i]), ipb[i], &il[i], fpb[i], &fl[i]);
            if (il[i] > pl) pl = il[i];
            if (fl[i] > pr) pr = fl[i];
        }
        p[k++] = '[';
        for (int i = 0; i < 3; ++i) {
            if (i) p[k++] = ' ';
            for (int z = il[i]; z < pl; ++z) p[k++] = ' ';
            for (int z = 0; z < il[i]; ++z) p[k++] = ipb[i][z];
            p[k++] = '.';
            for (int z = 0; z < fl[i]; ++z) p[k++] = fpb[i][z];
            for (int z = fl[i]; z < pr; ++z) p[k++] = ' ';
        }
        p[k++] = ']';
        return k;
    }
    uint64_t q[3];
    int ex[3];
    for (int i = 0; i < 3; ++i) {
        if (v[i] == 0.0) { q[i] = 0; ex[i] = 0; }
        else if (!sci9(v[i], &q[i], &ex[i])) return -1;
    }
    char dg[3][10];
    int fl[3], pl = 1, prec = 0, exp_size = 2;
    for (int i = 0; i < 3; ++i) {
        uint64_t t = q[i];
        for (int d = 8; d >= 0; --d) { dg[i][d] = (char)('0' + t % 10); t /= 10; }
        int f = 8;
        while (f > 0 && dg[i][f] == '0') --f;          // trimmed fraction length (digits 1..f)
        fl[i] = f;
        if (f > prec) prec = f;
        if (dsign(v[i])) pl = 2;
        int a = ex[i] < 0 ? -ex[i] : ex[i], nd = 1;
        while (a >= 10) { a /= 10; ++nd; }
        if (nd > exp_size) exp_size = nd;
    }
    p[k++] = '[';
    for (int i = 0; i < 3; ++i) {
        if (i) p[k++] = ' ';
        const bool neg = dsign(v[i]);
        if (pl == 2 && !neg) p[k++] = ' ';
        if (neg) p[k++] = '-';
        p[k++] = dg[i][0];
        p[k++] = '.';
        for (int d = 1; d <= prec; ++d) p[k++] = d <= fl[i] ? dg[i][d] : '0';
        p[k++] = 'e';
        p[k++] = ex[i] < 0 ? '-' : '+';
        char eb[12];
        int a = ex[i] < 0 ? -ex[i] : ex[i], nd = 0;
        do { eb[nd++] = (char)('0' + a % 10); a /= 10; } while (a);
        for (int z = nd; z < exp_size; ++z) p[k++] = '0';
        while (nd) p[k++] = eb[--nd];
    }
    p[k++] = ']';
    return k;
}

// Python repr(float) for 1e-4 <= x < 1 (every vessel radius): "0." + leading zeros + the SHORTEST digit string that reads back
// as x.  For n = 1, 2, ... the correctly rounded n-digit decimal q / 10^k of x = m / 2^s is formed exactly (q = round-half-even of
// m 10^k / 2^s); it reads back as x iff it lies within half an ulp of x, i.e. iff 2 |q 2^s - m 10^k| < 10^k (<= for an even
// significand: ties parse to even), and if ANY n-digit decimal does, the closest one does.  Powers of two (asymmetric rounding
// interval) and values outside the range return -1.  p needs 32 bytes.
OCTA_CF_HD int repr_unit(char* p, double x) {
    if (!(x >= 1e-4 && x < 1.0)) return -1;
    const uint64_t bits = dbits(x);
    const int be = (int)(bits >> 52);
    const uint64_t frac = bits & 0xfffffffffffffull;
    if (frac == 0) return -1;                              // power of two: the interval below x is half as wide
    const uint64_t m = frac | (1ull << 52);
    const int s = 1075 - be;                               // x = m / 2^s, 53 <= s <= 66
    // e10 = floor(log10(x)) in [-4, -1]
    int e10 = x >= 0.1 ? -1 : (x >= 0.01 ? -2 : (x >= 0.001 ? -3 : -4));
    const u128 one = 1;
    const u128 half_mask = (one << s) - 1, half = one << (s - 1);
    uint64_t digits = 0;
    int nd = 0;
    // most radii need 16 or 17 digits: probe 15 first, then walk in the direction that decides
    auto probe = [&](int n, uint64_t* q_out) -> int {      // 1: n digits read back as x, 0: they do not, -1: out of range
        const int k = n - 1 - e10;                         // q / 10^k has n significant digits
        if (k < 1 || k > 21) return -1;
        const u128 P = pow10_128(k);
        const u128 N = (u128)m * P;                        // < 2^53 * 10^21 < 2^123
        u128 q = N >> s;
        const u128 rem = N & half_mask;
        bool up = rem > half || (rem == half && (q & 1));
        u128 diff = up ? ((one << s) - rem) : rem;         // |q 2^s - N|
        if (up) ++q;
        const u128 twice = diff << 1;
        const bool ok = twice < P || (twice == P && (m & 1) == 0);
        *q_out = (uint64_t)q;
        return ok ? 1 : 0;
    };
    uint64_t q = 0;
    int r = probe(15, &q);
    if (r < 0) return -1;
    if (r == 1) {
        digits = q; nd = 15;
        for (int n = 14; n >= 1; --n) {
            uint64_t q2;
            const int r2 = probe(n, &q2);
            if (r2 != 1) break;
            digits = q2; nd = n;
        }
    } else {
        r = probe(16, &q);
        if (r < 0) return -1;
        if (r == 1) { digits = q; nd = 16; }
        else {
            r = probe(17, &q);
            if (r != 1) return -1;
            digits = q; nd = 17;
        }
    }
    // rounding up may have produced 10^nd (e.g. 0.0999999... -> 0.1): one digit more in front
    {
        uint64_t lim = 1;
        for (int i = 0; i < nd; ++i) lim *= 10u;
        if (digits >= lim) { digits /= 10u; e10 += 1; if (e10 > -1) return -1; }
        while (nd > 1 && digits % 10u == 0) { digits /= 10u; --nd; }
    }
    int k = 0;
    p[k++] = '0'; p[k++] = '.';
    for (int z = 0; z < -e10 - 1; ++z) p[k++] = '0';
    char tmp[20];
    for (int d = nd - 1; d >= 0; --d) { tmp[d] = (char)('0' + digits % 10u); digits /= 10u; }
    for (int d = 0; d < nd; ++d) p[k++] = tmp[d];
    return k;
}

}  // namespace csvfmt
}  // namespace octa
