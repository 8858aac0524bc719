// Device/host float64 helpers of the growth kernels.  They mirror, operation for operation, the
// numpy / CPython expressions of the reference (file:line in the comments refer to
// /root/reference/vessel_graph_generation).  This translation unit is compiled with -fmad=false;
// fused multiply-adds appear only where the reference's BLAS kernels use them (probed on the
// build container, see DESIGN.md "arithmetic fidelity"), through explicit fma().
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OCTA_HDI __host__ __device__ __forceinline__
#define OCTA_HDN __host__ __device__
#else
#define OCTA_HDI inline
#define OCTA_HDN
#endif

namespace octa {

constexpr double RAD2DEG = 57.29577951308232;     // numpy npy_rad2deg: x * (180/pi)
constexpr double DEG2RAD = 0.017453292519943295;  // numpy npy_deg2rad: x * (pi/180)

// np.dot of two 1-D float64 vectors of length 3 / 2 (OpenBLAS ddot kernel association)
OCTA_HDI double ddot3(const double* a, const double* b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
OCTA_HDI double ddot2(const double* a, const double* b) { return fma(a[1], b[1], a[0] * b[0]); }
// one row of np.dot(u, V.T) when V has >= 2 rows (OpenBLAS dgemv kernel association)
OCTA_HDI double gemv3(const double* u, const double* v) { return fma(u[2], v[2], fma(u[0], v[0], u[1] * v[1])); }
OCTA_HDI double norm3(const double* a) { return sqrt(ddot3(a, a)); }           // np.linalg.norm(v), 1-D
OCTA_HDI double norm2(const double* a) { return sqrt(ddot2(a, a)); }
OCTA_HDI double norm3_axis(const double* a) { return sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }  // norm(V, axis=1)
OCTA_HDI double clamp11(double c) { return c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c); }

// utilities.py:47-50 get_angle_between_two_vectors on 2-vectors
OCTA_HDI double angle_between_two(const double* u, const double* v) {
    const double c = ddot2(u, v) / norm2(u) / norm2(v);
    return RAD2DEG * acos(clamp11(c));
}

// numpy pairwise summation of a contiguous 1-D array read through an accessor (PW_BLOCKSIZE = 128)
template <class F>
OCTA_HDN double pairwise_sum(F get, long lo, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; ++i) res += get(lo + i);
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = get(lo + j);
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += get(lo + i + j);
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += get(lo + i);
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(get, lo, n2) + pairwise_sum(get, lo + n2, n - n2);
    }
}

// greenhouse.py:309-317
OCTA_HDI double oxygen_distance(double radius, double param_scale) {
    const double c_oxygen = 203.9e-3;
    const double kap = 0.02 * c_oxygen;
    const double r0 = 3.5e-3;
    const double c1 = kap * (radius * param_scale / r0) * exp(1 - (radius * param_scale / r0));
    return c1 * 6 / param_scale;
}

// Murray angles, greenhouse.py:206,214-215 / :264-266.  CPython evaluates x**4 and x**2 through libm pow; here they
// are products (<= 1 ULP from the correctly rounded power, like glibc's pow) -- the angles only steer directions.
OCTA_HDI void murray_angles(double r1, double r2, double kappa, double* phi1, double* phi2) {
    const double rp = pow(pow(r1, kappa) + pow(r2, kappa), 1 / kappa);
    const double rp2 = rp * rp, r12 = r1 * r1, r22 = r2 * r2;
    const double rp4 = rp2 * rp2, r14 = r12 * r12, r24 = r22 * r22;
    *phi1 = RAD2DEG * acos((rp4 + r14 - r24) / (2 * rp2 * r12));
    *phi2 = RAD2DEG * acos((rp4 + r24 - r14) / (2 * rp2 * r22));
}

// ------------------------------------------------------------------------------------------
// CPython hashing of a tuple of three float64 (Objects/tupleobject.c xxHash variant over
// Python/pyhash.c _Py_HashDouble); needed for the iteration order of the `to_add` set,
// greenhouse.py:100-111.
// ------------------------------------------------------------------------------------------
OCTA_HDI int64_t py_hash_double(double v) {
    const uint64_t MOD = ((uint64_t)1 << 61) - 1;
    if (v == 0.0) return 0;
    int e;
    double m = frexp(v, &e);
    int sign = 1;
    if (m < 0) { sign = -1; m = -m; }
    uint64_t x = 0;
    while (m != 0.0) {
        x = ((x << 28) & MOD) | x >> (61 - 28);
        m *= 268435456.0;
        e -= 28;
        const uint64_t y = (uint64_t)m;
        m -= (double)y;
        x += y;
        if (x >= MOD) x -= MOD;
    }
    e = e >= 0 ? e % 61 : 61 - 1 - ((-1 - e) % 61);
    x = ((x << e) & MOD) | x >> (61 - e);
    int64_t h = (int64_t)x * sign;
    if (h == -1) h = -2;
    return h;
}

OCTA_HDI int64_t py_hash_tuple3(double a, double b, double c) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    const double p[3] = {a, b, c};
    for (int i = 0; i < 3; ++i) {
        const uint64_t lane = (uint64_t)py_hash_double(p[i]);
        acc += lane * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += 3 ^ (P5 ^ 3527539ULL);
    if (acc == (uint64_t)-1) return 1546275796;
    return (int64_t)acc;
}

}  // namespace octa
