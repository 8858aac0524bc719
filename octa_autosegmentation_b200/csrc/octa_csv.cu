// Byte-exact writer (and parser) of the reference's graph CSV (host code).
//
// generate_vessel_graph.py:59-66 writes, through csv.writer (default dialect -> "\r\n"):
//     node1,node2,radius
//     str(ndarray float64[3]),str(ndarray float64[3]),repr(float)
// The two array cells follow numpy's array2string defaults (numpy/_core/arrayprint.py FloatingFormat,
// precision=8, floatmode='maxprec'): exponential notation iff max >= 1e8 or min < 1e-4 or
// max/min > 1000 over the non-zero |x|; positional mode prints every element as the shortest
// round-trip digit string cut to <= 8 fractional digits, trailing zeros trimmed but the '.' kept,
// right-aligned on the integer part and space-padded to the longest fraction; exponential mode
// prints every element with as many fractional digits as the longest one needs (zero padded), a sign
// and >= 2 exponent digits, non-negative numbers getting a blank where a '-' would be.  The radius
// cell is Python's repr(float): shortest round-trip digits, fixed notation for 1e-4 <= |x| < 1e16.
// Every consumer of the file (visualize_vessel_graphs.py:72-75, data_transforms.py:377-381,
// tree2img.py:73-76) reads the cells back with  s[1:-1].split(" ")  /  float(s); octa_parse_csv
// does the same.
#include <charconv>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "octa_common.h"

namespace {

struct Sci {            // value = (-1)^neg * d[0].d[1..] * 10^e10
    bool neg;
    std::string d;      // decimal digits, no leading zeros (except the value 0 -> "0")
    int e10;
};

Sci shortest(double x) {
    Sci s;
    s.neg = signbit(x);
    if (x == 0.0) { s.d = "0"; s.e10 = 0; return s; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), fabs(x), std::chars_format::scientific);
    *r.ptr = 0;
    // d[.ddd]e[+-]XX
    const char* e = strchr(buf, 'e');
    s.d.assign(1, buf[0]);
    if (buf[1] == '.') s.d.append(buf + 2, (size_t)(e - (buf + 2)));
    s.e10 = atoi(e + 1);
    return s;
}

// positional digits with at most `prec` fractional digits, trailing zeros trimmed; returns int / frac parts
void positional(double x, int prec, std::string* ip, std::string* fp) {
    const Sci s = shortest(x);
    const int n = (int)s.d.size();
    const int nfrac = n - 1 - s.e10;
    ip->clear(); fp->clear();
    if (nfrac <= prec) {
        if (s.e10 >= 0) {
            ip->assign(s.d, 0, std::min(n, s.e10 + 1));
            if (s.e10 + 1 > n) ip->append(s.e10 + 1 - n, '0');
            if (n > s.e10 + 1) fp->assign(s.d, s.e10 + 1, std::string::npos);
        } else {
            *ip = "0";
            fp->assign(-s.e10 - 1, '0');
            fp->append(s.d);
        }
    } else {
        char buf[512];
        snprintf(buf, sizeof(buf), "%.*f", prec, fabs(x));   // exact, round-half-even on the binary value (as Dragon4's cutoff)
        const char* dot = strchr(buf, '.');
        ip->assign(buf, (size_t)(dot - buf));
        fp->assign(dot + 1);
    }
    while (!fp->empty() && fp->back() == '0') fp->pop_back();
    if (s.neg) ip->insert(ip->begin(), '-');
}

// scientific digits d.ddd with at most `prec` fractional digits (trimmed); exponent returned separately
void scientific(double x, int prec, std::string* ip, std::string* fp, int* e10) {
    const Sci s = shortest(x);
    ip->clear(); fp->clear();
    if ((int)s.d.size() - 1 <= prec) {
        ip->assign(1, s.d[0]);
        fp->assign(s.d, 1, std::string::npos);
        *e10 = s.e10;
    } else {
        char buf[64];
        snprintf(buf, sizeof(buf), "%.*e", prec, fabs(x));
        const char* e = strchr(buf, 'e');
        ip->assign(1, buf[0]);
        fp->assign(buf + 2, (size_t)(e - (buf + 2)));
        *e10 = atoi(e + 1);
    }
    while (!fp->empty() && fp->back() == '0') fp->pop_back();
    if (s.neg) ip->insert(ip->begin(), '-');
}

// |x| * 10^8 rounded to the nearest integer, ties to even, EXACTLY (128-bit integer arithmetic on the binary value):
// the digits numpy's Dragon4 prints for precision=8 in positional mode.  When the shortest round-trip
// representation has <= 8 fractional digits it is a multiple of 1e-8 closer to x than half an ulp, so it is also
// what this rounding returns (after trimming zeros) -- no separate shortest-digits search is needed.
// Valid for 2^-60 < |x| < 2^20 (always true for unit-cube coordinates); returns false otherwise.
bool fixed8(double x, uint64_t* q_out) {
    const double ax = fabs(x);
    if (!(ax < 1048576.0)) return false;
    if (ax == 0.0) { *q_out = 0; return true; }
    {
        // Fast path.  p = fl(ax * 1e8) differs from the exact product T by at most ulp(p)/2 <= p * 2^-53, and p - floor(p) is exact
        // (p < 2^47).  round-half-even(T) can differ from rounding p only if p lies within that error of a half-integer; an
        // integer boundary between T and p does not matter (both sides round to that integer).
        const double p = ax * 1e8;
        const double fl = floor(p);
        const double fr = p - fl;
        if (fabs(fr - 0.5) > p * 1.2e-16 + 1e-300) { *q_out = (uint64_t)fl + (fr > 0.5 ? 1u : 0u); return true; }
    }
    uint64_t bits;
    memcpy(&bits, &ax, 8);
    const int be = (int)(bits >> 52);
    if (be == 0) return false;                             // subnormal: leave to the generic path
    const uint64_t m = (bits & 0xfffffffffffffull) | (1ull << 52);   // ax = m * 2^(be-1075)
    const int s = 1075 - be;                               // ax = m / 2^s
    if (s <= 0 || s > 113) return false;
    const unsigned __int128 N = (unsigned __int128)m * 100000000u;
    const unsigned __int128 one = 1;
    unsigned __int128 q = N >> s;
    const unsigned __int128 rem = N & ((one << s) - 1), half = one << (s - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
    *q_out = (uint64_t)q;
    return true;
}

static const char DIG2[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";

// writes the integer / fraction digit strings of q = round(|x| * 1e8); fraction trimmed of trailing zeros
inline void fixed8_digits(uint64_t q, bool neg, char* ip, int* ilen, char* fp, int* flen) {
    uint64_t ipart = 0;
    uint32_t f = (uint32_t)q;
    if (q >= 100000000u) { ipart = q / 100000000u; f = (uint32_t)(q - ipart * 100000000u); }   // (|x| < 1 is the common case)
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
    const uint32_t hi = f / 10000u, lo = f % 10000u;
    memcpy(fp, DIG2 + 2 * (hi / 100u), 2); memcpy(fp + 2, DIG2 + 2 * (hi % 100u), 2);
    memcpy(fp + 4, DIG2 + 2 * (lo / 100u), 2); memcpy(fp + 6, DIG2 + 2 * (lo % 100u), 2);
    int fl = 8;
    while (fl > 0 && fp[fl - 1] == '0') --fl;
    *flen = fl;
}

// 9 significant digits of |x| (d.dddddddd x 10^e10), exactly rounded on the binary value (ties to even): what
// Dragon4 prints for dragon4_scientific(precision=8).  Valid for 1e-30 < |x| < 1e8; returns false otherwise.
bool sci9(double x, uint64_t* q_out, int* e10_out) {
    static const uint64_t P10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                     1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
                                     100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
                                     1000000000000000000ull, 10000000000000000000ull};
    const double ax = fabs(x);
    if (!(ax > 1e-11 && ax < 1e8)) return false;
    uint64_t bits;
    memcpy(&bits, &ax, 8);
    const int be = (int)(bits >> 52);
    if (be == 0) return false;
    const uint64_t m = (bits & 0xfffffffffffffull) | (1ull << 52);
    const int s = 1075 - be;                               // ax = m / 2^s, s in (26, 90)
    const int e = be - 1022;                               // ax = fr * 2^e with fr in [0.5, 1)
    int e10 = (int)floor((e - 1) * 0.30102999566398120);   // floor(log10(ax)) or one less
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int k = 8 - e10;                             // scale by 10^k
        if (k < 0 || k > 19 || s <= 0 || s > 120) return false;
        const unsigned __int128 N = (unsigned __int128)m * P10[k];
        const unsigned __int128 one = 1;
        unsigned __int128 q = N >> s;
        const unsigned __int128 rem = N & ((one << s) - 1), half = one << (s - 1);
        if (rem > half || (rem == half && (q & 1))) ++q;
        if (q >= 1000000000u) {
            // either e10 was one too small, or rounding carried into a 10th digit (d = 9.99999999(5+) -> 1.00000000 e+1)
            const unsigned __int128 lo = (unsigned __int128)1000000000u << s;
            if (N >= lo) { ++e10; continue; }              // genuinely >= 10^(e10+1)
            *q_out = 100000000u; *e10_out = e10 + 1;
            return true;
        }
        if (q < 100000000u) { --e10; continue; }
        *q_out = (uint64_t)q; *e10_out = e10;
        return true;
    }
    return false;
}

// Positional [a b c] cell written straight into `p` (>= 64 bytes free); nullptr when the generic formatter is needed
// (exponential notation, |x| >= 2^20).  Same layout rules as append_array3 below.
inline char* fast_array3(char* p, const double* v) {
    double mx = 0, mn = 0;
    bool any = false;
    for (int i = 0; i < 3; ++i) {
        const double a = fabs(v[i]);
        if (a != 0.0) { if (!any) { mx = mn = a; any = true; } else { if (a > mx) mx = a; if (a < mn) mn = a; } }
    }
    if (any && (mx >= 1.e8 || mn < 0.0001 || mx / mn > 1000.)) return nullptr;
    uint64_t q[3];
    if (!(fixed8(v[0], &q[0]) && fixed8(v[1], &q[1]) && fixed8(v[2], &q[2]))) return nullptr;
    char ipb[3][24], fpb[3][8];
    int il[3], fl[3], pl = 0, pr = 0;
    for (int i = 0; i < 3; ++i) {
        fixed8_digits(q[i], signbit(v[i]), ipb[i], &il[i], fpb[i], &fl[i]);
        if (il[i] > pl) pl = il[i];
        if (fl[i] > pr) pr = fl[i];
    }
    *p++ = '[';
    for (int i = 0; i < 3; ++i) {
        if (i) *p++ = ' ';
        for (int k = il[i]; k < pl; ++k) *p++ = ' ';
        memcpy(p, ipb[i], (size_t)il[i]); p += il[i];
        *p++ = '.';
        memcpy(p, fpb[i], 8); p += fl[i];                 // (8 bytes copied, fl[i] kept: the tail is overwritten next)
        for (int k = fl[i]; k < pr; ++k) *p++ = ' ';
    }
    *p++ = ']';
    return p;
}

// Python repr(float) for the fixed-notation range, straight into `p` (>= 40 bytes free); nullptr outside it
inline char* fast_repr(char* p, double x) {
    const double a = fabs(x);
    if (!(a >= 1e-4 && a < 1e16)) return nullptr;
    auto r = std::to_chars(p, p + 40, x, std::chars_format::fixed);
    bool dot = false;
    for (char* c = p; c != r.ptr; ++c) if (*c == '.') { dot = true; break; }
    p = r.ptr;
    if (!dot) { *p++ = '.'; *p++ = '0'; }
    return p;
}

void append_array3(std::string& out, const double* v) {
    // FloatingFormat.fillFormat
    double mx = 0, mn = 0;
    bool any = false;
    for (int i = 0; i < 3; ++i) {
        const double a = fabs(v[i]);
        if (a != 0.0) { if (!any) { mx = mn = a; any = true; } else { if (a > mx) mx = a; if (a < mn) mn = a; } }
    }
    const bool exp_format = any && (mx >= 1.e8 || mn < 0.0001 || mx / mn > 1000.);
    if (!exp_format) {
        uint64_t q[3];
        if (fixed8(v[0], &q[0]) && fixed8(v[1], &q[1]) && fixed8(v[2], &q[2])) {
            char ipb[3][24], fpb[3][8];
            int il[3], fl[3], pl = 0, pr = 0;
            for (int i = 0; i < 3; ++i) {
                fixed8_digits(q[i], signbit(v[i]) && q[i] != 0 ? true : signbit(v[i]), ipb[i], &il[i], fpb[i], &fl[i]);
                if (il[i] > pl) pl = il[i];
                if (fl[i] > pr) pr = fl[i];
            }
            char line[128];
            int k = 0;
            line[k++] = '[';
            for (int i = 0; i < 3; ++i) {
                if (i) line[k++] = ' ';
                for (int p = il[i]; p < pl; ++p) line[k++] = ' ';
                memcpy(line + k, ipb[i], il[i]); k += il[i];
                line[k++] = '.';
                memcpy(line + k, fpb[i], fl[i]); k += fl[i];
                for (int p = fl[i]; p < pr; ++p) line[k++] = ' ';
            }
            line[k++] = ']';
            out.append(line, k);
            return;
        }
    }
    if (exp_format) {
        uint64_t q[3];
        int ex[3];
        bool ok = true;
        for (int i = 0; i < 3 && ok; ++i) {
            if (v[i] == 0.0) { q[i] = 0; ex[i] = 0; } else ok = sci9(v[i], &q[i], &ex[i]);
        }
        if (ok) {
            char dg[3][10];
            int fl[3], pl = 1, prec = 0, exp_size = 2;
            for (int i = 0; i < 3; ++i) {
                uint64_t t = q[i];
                for (int d = 8; d >= 0; --d) { dg[i][d] = (char)('0' + t % 10); t /= 10; }
                int f = 8;
                while (f > 0 && dg[i][f] == '0') --f;          // trimmed fraction length (digits 1..f)
                fl[i] = f;
                if (f > prec) prec = f;
                if (signbit(v[i])) pl = 2;
                int a = abs(ex[i]), nd = 1;
                while (a >= 10) { a /= 10; ++nd; }
                if (nd > exp_size) exp_size = nd;
            }
            char line[160];
            int k = 0;
            line[k++] = '[';
            for (int i = 0; i < 3; ++i) {
                if (i) line[k++] = ' ';
                const bool neg = signbit(v[i]);
                if (pl == 2 && !neg) line[k++] = ' ';
                if (neg) line[k++] = '-';
                line[k++] = dg[i][0];
                line[k++] = '.';
                for (int d = 1; d <= prec; ++d) line[k++] = d <= fl[i] ? dg[i][d] : '0';
                line[k++] = 'e';                                // "e%c%0*d" by hand (snprintf dominated this path)
                line[k++] = ex[i] < 0 ? '-' : '+';
                {
                    char eb[12];
                    int a = abs(ex[i]), nd = 0;
                    do { eb[nd++] = (char)('0' + a % 10); a /= 10; } while (a);
                    for (int z = nd; z < exp_size; ++z) line[k++] = '0';
                    while (nd) line[k++] = eb[--nd];
                }
            }
            line[k++] = ']';
            out.append(line, k);
            return;
        }
    }
    std::string ip[3], fp[3];
    out.push_back('[');
    if (!exp_format) {
        size_t pl = 0, pr = 0;
        for (int i = 0; i < 3; ++i) { positional(v[i], 8, &ip[i], &fp[i]); pl = std::max(pl, ip[i].size()); pr = std::max(pr, fp[i].size()); }
        for (int i = 0; i < 3; ++i) {
            if (i) out.push_back(' ');
            out.append(pl - ip[i].size(), ' ');
            out += ip[i];
            out.push_back('.');
            out += fp[i];
            out.append(pr - fp[i].size(), ' ');
        }
    } else {
        int ex[3];
        size_t pl = 0, prec = 0;
        int exp_size = 2;
        for (int i = 0; i < 3; ++i) {
            scientific(v[i], 8, &ip[i], &fp[i], &ex[i]);
            pl = std::max(pl, ip[i].size());
            prec = std::max(prec, fp[i].size());
            int a = abs(ex[i]), nd = 1;
            while (a >= 10) { a /= 10; ++nd; }
            exp_size = std::max(exp_size, nd);
        }
        char eb[16];
        for (int i = 0; i < 3; ++i) {
            if (i) out.push_back(' ');
            out.append(pl - ip[i].size(), ' ');
            out += ip[i];
            out.push_back('.');
            out += fp[i];
            out.append(prec - fp[i].size(), '0');
            snprintf(eb, sizeof(eb), "e%c%0*d", ex[i] < 0 ? '-' : '+', exp_size, abs(ex[i]));
            out += eb;
        }
    }
    out.push_back(']');
}

// Python repr(float)
void append_repr(std::string& out, double x) {
    const double axr = fabs(x);
    if (axr >= 1e-4 && axr < 1e16) {
        // fixed notation range of float_repr_style 'short': shortest round-trip digits, plain positional
        char b[40];
        auto r = std::to_chars(b, b + sizeof(b), x, std::chars_format::fixed);
        bool dot = false;
        for (char* c = b; c != r.ptr; ++c) if (*c == '.') { dot = true; break; }
        out.append(b, r.ptr - b);
        if (!dot) out += ".0";
        return;
    }
    if (x != x) { out += "nan"; return; }
    if (isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
    const Sci s = shortest(x);
    if (s.neg) out.push_back('-');
    const int n = (int)s.d.size();
    const int decpt = s.e10 + 1;          // digits before the decimal point
    if (x == 0.0) { out += "0.0"; return; }
    if (decpt > -4 && decpt <= 16) {      // float_repr_style 'short', format code 'r'
        if (decpt <= 0) { out += "0."; out.append(-decpt, '0'); out += s.d; }
        else if (decpt >= n) { out += s.d; out.append(decpt - n, '0'); out += ".0"; }
        else { out.append(s.d, 0, decpt); out.push_back('.'); out.append(s.d, decpt, std::string::npos); }
    } else {
        out.push_back(s.d[0]);
        if (n > 1) { out.push_back('.'); out.append(s.d, 1, std::string::npos); }
        char eb[16];
        snprintf(eb, sizeof(eb), "e%c%02d", s.e10 < 0 ? '-' : '+', abs(s.e10));
        out += eb;
    }
}

}  // namespace

extern "C" int octa_format_csv(const double* edges7, int64_t n_edges, char* buf, size_t cap, size_t* len) {
    OCTA_ARG_CHECK(n_edges >= 0 && len && (n_edges == 0 || edges7), "bad arguments");
    // Rows are written straight into the caller's buffer while it has room for a worst-case row; the (rare) cells the fast
    // writers decline -- exponential notation, huge values -- and the case of a too small / missing buffer go through strings.
    static const char header[] = "node1,node2,radius\r\n";
    constexpr size_t ROW_MAX = 320;
    std::string spill;                                      // rows that did not fit the caller's buffer
    char* p = buf;
    char* const end = buf ? buf + cap : nullptr;
    bool direct = buf && cap >= sizeof(header) - 1 + ROW_MAX;
    if (direct) { memcpy(p, header, sizeof(header) - 1); p += sizeof(header) - 1; }
    else spill.assign(header, sizeof(header) - 1);
    std::string tmp;
    for (int64_t i = 0; i < n_edges; ++i) {
        const double* e = edges7 + 7 * i;
        for (int k = 0; k < 7; ++k)
            if (!(e[k] == e[k]) || isinf(e[k])) { octa::set_error("octa_format_csv: non-finite value in edge %lld", (long long)i); return OCTA_E_ARG; }
        if (direct && (size_t)(end - p) < ROW_MAX) { direct = false; spill.reserve((size_t)(n_edges - i) * 100); }
        if (direct) {
            char* r = fast_array3(p, e);
            if (!r) { tmp.clear(); append_array3(tmp, e); memcpy(p, tmp.data(), tmp.size()); r = p + tmp.size(); }
            p = r; *p++ = ',';
            r = fast_array3(p, e + 3);
            if (!r) { tmp.clear(); append_array3(tmp, e + 3); memcpy(p, tmp.data(), tmp.size()); r = p + tmp.size(); }
            p = r; *p++ = ',';
            r = fast_repr(p, e[6]);
            if (!r) { tmp.clear(); append_repr(tmp, e[6]); memcpy(p, tmp.data(), tmp.size()); r = p + tmp.size(); }
            p = r; *p++ = '\r'; *p++ = '\n';
        } else {
            append_array3(spill, e);
            spill.push_back(',');
            append_array3(spill, e + 3);
            spill.push_back(',');
            append_repr(spill, e[6]);
            spill += "\r\n";
        }
    }
    const size_t written = buf ? (size_t)(p - buf) : 0;
    *len = written + spill.size();
    if (!buf || !spill.empty()) {
        if (buf && cap >= *len) { memcpy(buf + written, spill.data(), spill.size()); return OCTA_OK; }
        octa::set_error("octa_format_csv: buffer too small (%zu needed)", *len);
        return OCTA_E_NOMEM;
    }
    return OCTA_OK;
}

// Reads a graph CSV the way the reference's consumers do: header line, then rows whose first two cells are
// "[x y z]" (split on blanks, empty tokens dropped) and whose third is a float.  Returns the number of rows
// (also when edges7_out is too small / NULL, so callers can size the buffer), or a negative error code.
// One number of a cell, the way Python's float() reads what numpy / repr wrote: std::from_chars (correctly rounded like
// strtod, several times faster, bounded by `e` instead of relying on a terminator); whatever it declines -- a leading '+',
// hex floats, values that overflow or underflow double -- goes to strtod as before.
static inline bool parse_number(const char** q, const char* e, double* out) {
    const std::from_chars_result r = std::from_chars(*q, e, *out, std::chars_format::general);
    if (r.ec == std::errc()) { *q = r.ptr; return true; }
    char* stop = nullptr;
    *out = strtod(*q, &stop);
    if (stop == *q) return false;
    *q = stop;
    return true;
}

extern "C" int64_t octa_parse_csv(const char* text, size_t len, double* edges7_out, int64_t cap_edges) {
    if (!text) { octa::set_error("octa_parse_csv: null text"); return OCTA_E_ARG; }
    const char* p = text;
    const char* end = text + len;
    auto next_line = [&](const char** b, const char** e) {
        if (p >= end) return false;
        *b = p;
        while (p < end && *p != '\n') ++p;
        *e = p;
        if (p < end) ++p;
        while (*e > *b && ((*e)[-1] == '\r')) --*e;
        return true;
    };
    const char *b, *e;
    if (!next_line(&b, &e)) return 0;   // header
    int64_t n = 0;
    while (next_line(&b, &e)) {
        if (b == e) continue;
        double v[7];
        const char* q = b;
        int k = 0;
        for (int cell = 0; cell < 2; ++cell) {
            while (q < e && *q != '[') ++q;
            if (q >= e) { octa::set_error("octa_parse_csv: malformed row %lld", (long long)n); return OCTA_E_ARG; }
            ++q;
            for (int c = 0; c < 3; ++c) {
                while (q < e && *q == ' ') ++q;
                if (!parse_number(&q, e, &v[k++])) { octa::set_error("octa_parse_csv: malformed number in row %lld", (long long)n); return OCTA_E_ARG; }
            }
            while (q < e && *q != ']') ++q;
            if (q < e) ++q;
        }
        while (q < e && *q != ',') ++q;
        if (q >= e) { octa::set_error("octa_parse_csv: missing radius in row %lld", (long long)n); return OCTA_E_ARG; }
        ++q;
        while (q < e && *q == ' ') ++q;
        if (!parse_number(&q, e, &v[6])) v[6] = 0.0;      // (strtod's "no conversion" value, as before)
        if (edges7_out && n < cap_edges) memcpy(edges7_out + 7 * n, v, sizeof(v));
        ++n;
    }
    return n;
}
