/* CPU ORACLE (test infrastructure, NOT part of the product): sequential C++ restatement of the
 * reference's space-colonization growth.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Pinned against the real reference: tests/golden/graph_*.csv were produced by the unmodified
 * /root/reference modules through oracle/ref_harness.py (seeded), and tests/test_oracle_growth.py
 * compares this file's output with them (topology + radii bit-exact; positions to the printed digits).
 *
 * Follows (file:line relative to /root/reference/vessel_graph_generation unless noted):
 *   greenhouse.py:17-51    parameter scaling, FAZ radius draw, per-mode (re)initialisation
 *   simulation_space.py:36-67,89-98   76^2 validity mask, candidate sinks, bounds test
 *   forest.py:68-181 (stumps) and :38-66 (nerve)   root stumps
 *   arterial_tree.py:9-44,58-67,174-184,218-229    node flags, Murray radius walk, BFS export
 *   element_mesh.py:97-211  list-order element store; exact NN (+ max_dist), radius query in cKDTree
 *                           result order (= ascending position in tree.indices; build cloned from
 *                           scipy/spatial/ckdtree/src/build.cxx semantics)
 *   greenhouse.py:57-137   develop_forest loop, O2 -> CO2 conversion through a CPython `set`
 *                          (iteration order emulated: Objects/setobject.c, tuple/float hashes)
 *   greenhouse.py:139-155  simulation_space_expansion
 *   greenhouse.py:157-307  grow_vessels (leaf: elongate/bifurcate; inter-node: sprout)
 *   greenhouse.py:309-366  oxygen distance, sample_oxygen_sinks, assign_attraction_points_to_node
 *   generate_vessel_graph.py:45-56   edge export order
 * RNG: two MT19937 streams -- Python `random` (init_by_array) and legacy numpy RandomState
 * (init_genrand, polar gauss, masked-rejection randint), SURVEY Appendix A1.
 *
 * Arithmetic notes.  All float64.  The reference's small dot products go through OpenBLAS
 * (ddot / dgemv / dsyrk); their FMA association on the container's CPU was probed
 * (ddot n=3: fma(a2,b2,fma(a1,b1,a0*b0)); dgemv with >=2 rows: fma(a2,b2,fma(a0,b0,a1*b1));
 * dsyrk: running fma from 0) and is mirrored here so the goldens are met as closely as possible.
 * numpy's arccos/exp are SVML-derived and differ from libm by <=1 ULP in ~9 % / 4.5 % of calls;
 * that is below the 8 printed digits except for a print-boundary lottery.
 * The 3x3 eigenproblem of greenhouse.py:229 uses a Jacobi solver and fixes the eigenvector
 * sign with the caller-supplied `eig_hook` (the Python side passes numpy.linalg.eig, i.e. the same
 * LAPACK dgeev as the reference) when present.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#define FMA(a, b, c) __builtin_fma((a), (b), (c))

extern "C" {

struct OGMode {
    int I, N;
    double eps_n, eps_s, eps_k, delta_art, delta_ven, gamma_art, gamma_ven, phi, omega, kappa, delta_sigma;
    int reinit;      /* mode["name"] != modes[0]["name"]  (greenhouse.py:84) */
    int first_mode;  /* mode == modes[0]                   (greenhouse.py:95) */
};

struct OGConfig {
    double d, r, faz_bound0, faz_bound1, rotation_radius, faz_center[2], nerve_center[2], nerve_radius, param_scale;
    double size[3];
    int n_modes;
    OGMode modes[8];
    int forest_type; /* 0 = stumps, 1 = nerve */
    int n_trees;
    int n_walls;
    int walls[6];    /* enabled source walls in config order: 0=x0 1=x1 2=y0 3=y1 4=z0 5=z1 */
    int ball_order;  /* 0 = cKDTree order (exact), 1 = list-index order (sensitivity experiments) */
    int venous;      /* 1 = grow a venous forest as well (generate_vessel_graph.py:34) */
    /* SimulationSpace.oxygen_sample_geometry_path (simulation_space.py:26-34): C-order bool mask, or NULL */
    const unsigned char* geometry;
    int geom_dims[3];
};

typedef void (*og_eig_hook)(const double* cov9, double* w3, double* v9);
typedef void (*og_trace_hook)(int t, long n_art, long n_oxy, long n_ven, long n_co2, long py_draws, long np_draws);

struct OGStats {
    long n_art_nodes, n_ven_nodes, n_oxy_left, n_co2_left, py_draws, np_u32, nn_queries, ball_queries,
        bifurcations, sprouts, elongations, walk_steps, sum_A, sum_M, sum_P, sum_S,
        multi_balls, reordered_balls, interacting_groups, kd_builds, inter_evals, inter_r1_changed, max_dict, max_list,
        flag_iters_cons, flag_iters_perm, diff_iters, perm_groups;   /* ball-order sensitivity instrumentation */
};
}

namespace {

/* ------------------------------------------------------------------ MT19937 ---------------- */
struct MT {
    uint32_t mt[624];
    int idx;
    long drawn;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
        drawn = 0;
    }
    void init_by_array(const uint32_t* key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        int k = 624 > len ? 624 : len;
        for (; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= len) j = 0;
        }
        for (k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
        idx = 624;
    }
    void regen() {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx = 0;
    }
    uint32_t u32() {
        if (idx >= 624) regen();
        uint32_t y = mt[idx++];
        ++drawn;
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double dbl() {
        uint32_t a = u32() >> 5, b = u32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

/* Python `random` */
struct PyRandom {
    MT g;
    void seed(uint64_t s) {
        uint32_t key[2] = {(uint32_t)(s & 0xffffffffu), (uint32_t)(s >> 32)};
        g.init_by_array(key, key[1] ? 2 : 1);
        g.drawn = 0;
    }
    double random() { return g.dbl(); }
    double uniform(double a, double b) { return a + (b - a) * g.dbl(); }
    int randbelow(int n) { /* _randbelow_with_getrandbits */
        int k = 0;
        for (int v = n; v; v >>= 1) ++k;
        uint32_t r = g.u32() >> (32 - k);
        while ((int)r >= n) r = g.u32() >> (32 - k);
        return (int)r;
    }
};

/* legacy numpy RandomState */
struct NpRandom {
    MT g;
    void seed(uint32_t s) { g.init_genrand(s); }
    double dbl() { return g.dbl(); }
    double uniform(double lo, double hi) { return lo + (hi - lo) * g.dbl(); }
    double normal(double loc, double scale) { /* legacy_gauss, polar method; cache never reused on this path */
        double x1, x2, r2;
        do {
            x1 = 2.0 * g.dbl() - 1.0;
            x2 = 2.0 * g.dbl() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        double f = sqrt(-2.0 * log(r2) / r2);
        return loc + scale * (f * x2);
    }
    uint32_t randint(uint32_t n) { /* randint(0, n): masked rejection on 32-bit draws */
        uint32_t rng = n - 1, mask = rng;
        if (rng == 0) return 0;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = (g.u32() & mask)) > rng) {}
        return v;
    }
};

/* ------------------------------------------------------------------ numpy-like math --------- */
const double RAD2DEG = 57.29577951308232;   /* 180/pi, numpy npy_rad2deg */
const double DEG2RAD = 0.017453292519943295; /* pi/180 */

inline double ddot3(const double* a, const double* b) { return FMA(a[2], b[2], FMA(a[1], b[1], a[0] * b[0])); }
inline double ddot2(const double* a, const double* b) { return FMA(a[1], b[1], a[0] * b[0]); }
inline double gemv3(const double* u, const double* v) { return FMA(u[2], v[2], FMA(u[0], v[0], u[1] * v[1])); }
inline double norm3(const double* a) { return sqrt(ddot3(a, a)); }
inline double norm2(const double* a) { return sqrt(ddot2(a, a)); }
inline double norm3_axis(const double* a) { return sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); } /* add.reduce */
inline double clamp11(double c) { return c < -1 ? -1 : (c > 1 ? 1 : c); }

/* numpy pairwise summation of a contiguous 1-D array (npy pairwise_sum, PW_BLOCKSIZE 128) */
double pairwise_sum(const double* a, long n) {
    if (n < 8) {
        double res = 0.;   /* numpy: res = 0.; for i: res += a[i]  (first add exact) */
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

/* utilities.py:47-50 */
double angle_between_two(const double* u, const double* v) { /* 2-vectors */
    double c = ddot2(u, v) / norm2(u) / norm2(v);
    return RAD2DEG * acos(clamp11(c)) ;
}

/* Jacobi eigen-decomposition of a symmetric 3x3 (fallback when no LAPACK hook is given) */
void jacobi3(const double* A9, double* w, double* V9) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = A9[3 * i + j];
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) {
        w[i] = a[i][i];
        for (int k = 0; k < 3; ++k) V9[3 * k + i] = v[k][i];
    }
}

/* ------------------------------------------------------------------ CPython hashing / set ---- */
const uint64_t HASH_MOD = ((uint64_t)1 << 61) - 1;

int64_t py_hash_double(double v) {
    if (v == 0.0) return 0;
    int e;
    double m = frexp(v, &e);
    int sign = 1;
    if (m < 0) { sign = -1; m = -m; }
    uint64_t x = 0;
    while (m) {
        x = ((x << 28) & HASH_MOD) | x >> (61 - 28);
        m *= 268435456.0;
        e -= 28;
        uint64_t y = (uint64_t)m;
        m -= (double)y;
        x += y;
        if (x >= HASH_MOD) x -= HASH_MOD;
    }
    e = e >= 0 ? e % 61 : 61 - 1 - ((-1 - e) % 61);
    x = ((x << e) & HASH_MOD) | x >> (61 - e);
    int64_t h = (int64_t)x * sign;
    if (h == -1) h = -2;
    return h;
}

int64_t py_hash_tuple3(const double* p) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    for (int i = 0; i < 3; ++i) {
        uint64_t lane = (uint64_t)py_hash_double(p[i]);
        acc += lane * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += 3 ^ (P5 ^ 3527539ULL);
    if (acc == (uint64_t)-1) return 1546275796;
    return (int64_t)acc;
}

/* CPython 3.12 set of (hash, key-id); keys are distinct ids with equality == id equality */
struct PySet {
    struct Entry { int64_t hash; int key; };
    std::vector<Entry> table;
    size_t mask, fill, used;
    PySet() { table.assign(8, Entry{0, -1}); mask = 7; fill = used = 0; }
    static void insert_clean(std::vector<Entry>& t, size_t mask, int key, int64_t hash) {
        size_t perturb = (size_t)hash, i = (size_t)hash & mask;
        while (1) {
            size_t e = i;
            if (t[e].key < 0) { t[e] = Entry{hash, key}; return; }
            if (i + 9 <= mask) {
                for (int j = 0; j < 9; ++j) {
                    ++e;
                    if (t[e].key < 0) { t[e] = Entry{hash, key}; return; }
                }
            }
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    void resize(size_t minused) {
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        std::vector<Entry> nt(newsize, Entry{0, -1});
        for (const Entry& en : table)
            if (en.key >= 0) insert_clean(nt, newsize - 1, en.key, en.hash);
        table.swap(nt);
        mask = newsize - 1;
        fill = used;
    }
    /* returns final slot (or -1 for a duplicate); *resized set when the table grew; examined slots appended */
    long add(int key, int64_t hash, std::vector<size_t>* examined = nullptr, bool* resized = nullptr) {
        size_t perturb = (size_t)hash, i = (size_t)hash & mask;
        while (1) {
            size_t e = i;
            int probes = (i + 9 <= mask) ? 9 : 0;
            do {
                if (table[e].key < 0) {
                    table[e] = Entry{hash, key};
                    ++fill; ++used;
                    if (fill * 5 >= mask * 3) { resize(used > 50000 ? used * 2 : used * 4); if (resized) *resized = true; }
                    return (long)e;
                }
                if (table[e].hash == hash && table[e].key == key) return -1;
                if (examined) examined->push_back(e);
                ++e;
            } while (probes--);
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    template <class F> void for_each(F f) const {
        for (const Entry& en : table) if (en.key >= 0) f(en.key);
    }
};

/* instrumentation: does the table grow before the LAST key of this group is inserted (index order)? */
static bool trial_mid_resize(const PySet& before, const std::vector<int>& fresh, const double* xyz) {
    PySet t = before;
    for (size_t q = 0; q + 1 < fresh.size(); ++q) {
        bool r = false;
        t.add(fresh[q], py_hash_tuple3(&xyz[3 * fresh[q]]), nullptr, &r);
        if (r) return true;
    }
    return false;
}

/* ------------------------------------------------------------------ cKDTree index order ----- */
struct KdOrder {
    const double* pts; /* n x 3 */
    std::vector<long> idx;
    std::vector<int> rank; /* rank[i] = position of point i in idx */
    void build_rec(long start, long end) {
        if (end - start <= 16) return;
        double mins[3], maxes[3];
        for (int k = 0; k < 3; ++k) mins[k] = maxes[k] = pts[idx[start] * 3 + k];
        for (long j = start + 1; j < end; ++j)
            for (int k = 0; k < 3; ++k) {
                double t = pts[idx[j] * 3 + k];
                maxes[k] = maxes[k] > t ? maxes[k] : t;
                mins[k] = mins[k] < t ? mins[k] : t;
            }
        int d = 0;
        double size = 0;
        for (int k = 0; k < 3; ++k)
            if (maxes[k] - mins[k] > size) { d = k; size = maxes[k] - mins[k]; }
        if (maxes[d] == mins[d]) return;
        long n = end - start, half = n / 2;
        const double* P = pts;
        std::nth_element(idx.begin() + start, idx.begin() + start + half, idx.begin() + end, [P, d](long a, long b) {
            double pa = P[a * 3 + d], pb = P[b * 3 + d];
            return pa == pb ? a < b : pa < pb;
        });
        long p = start + half;
        double split = pts[idx[p] * 3 + d];
        p = start;
        long q = end - 1;
        while (p <= q) {
            if (pts[idx[p] * 3 + d] < split) ++p;
            else if (pts[idx[q] * 3 + d] >= split) --q;
            else { std::swap(idx[p], idx[q]); ++p; --q; }
        }
        if (p == start) {
            long j = start;
            split = pts[idx[j] * 3 + d];
            for (long i = start + 1; i < end; ++i)
                if (pts[idx[i] * 3 + d] < split) { j = i; split = pts[idx[j] * 3 + d]; }
            std::swap(idx[start], idx[j]);
            p = start + 1;
        } else if (p == end) {
            long j = end - 1;
            split = pts[idx[j] * 3 + d];
            for (long i = start; i < end - 1; ++i)
                if (pts[idx[i] * 3 + d] > split) { j = i; split = pts[idx[j] * 3 + d]; }
            std::swap(idx[end - 1], idx[j]);
            p = end - 1;
        }
        build_rec(start, p);
        build_rec(p, end);
    }
    void build(const double* points, long n) {
        pts = points;
        idx.resize(n);
        for (long i = 0; i < n; ++i) idx[i] = i;
        if (n > 0) build_rec(0, n);
        rank.resize(n);
        for (long i = 0; i < n; ++i) rank[idx[i]] = (int)i;
    }
};

/* ------------------------------------------------------------------ point store with 2-D buckets
 * element_mesh.py KD_Tree semantics: elements in list order, exact NN / radius queries.  The
 * bucket grid only prunes candidates; every reported distance is the exact float64 expression
 * sqrt(((dx*dx)+dy*dy)+dz*dz) of cKDTree. */
struct PointList {
    std::vector<double> xyz; /* 3 per element, list order */
    std::vector<int> id;     /* payload (node id, or sink serial) */
    static const int G = 48;
    std::vector<std::vector<int>> cell; /* positions in list */
    bool dirty = true;
    long size() const { return (long)id.size(); }
    void push(const double* p, int payload) {
        xyz.insert(xyz.end(), p, p + 3);
        id.push_back(payload);
        dirty = true;
    }
    static int cidx(double v) {
        int c = (int)floor(v * G);
        return c < 0 ? 0 : (c >= G ? G - 1 : c);
    }
    void rebuild() {
        cell.assign((size_t)G * G, std::vector<int>());
        for (long i = 0; i < size(); ++i) cell[cidx(xyz[3 * i]) * G + cidx(xyz[3 * i + 1])].push_back((int)i);
        dirty = false;
    }
    void remove_positions(const std::vector<char>& kill) { /* stable */
        long w = 0;
        for (long i = 0; i < size(); ++i)
            if (!kill[i]) {
                if (w != i) {
                    xyz[3 * w] = xyz[3 * i]; xyz[3 * w + 1] = xyz[3 * i + 1]; xyz[3 * w + 2] = xyz[3 * i + 2];
                    id[w] = id[i];
                }
                ++w;
            }
        xyz.resize(3 * w);
        id.resize(w);
        dirty = true;
    }
    static inline double d2(const double* a, const double* b) {
        double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        return (dx * dx + dy * dy) + dz * dz;
    }
    /* element_mesh.py:150-165: global NN, accepted iff dist <= max_dist.  Returns list position or -1. */
    long nearest_within(const double* p, double max_dist) {
        if (size() == 0) return -1;
        if (dirty) rebuild();
        double r = max_dist;
        int x0 = cidx(p[0] - r), x1 = cidx(p[0] + r), y0 = cidx(p[1] - r), y1 = cidx(p[1] + r);
        long best = -1;
        double bd = INFINITY;
        for (int cx = x0; cx <= x1; ++cx)
            for (int cy = y0; cy <= y1; ++cy)
                for (int i : cell[cx * G + cy]) {
                    double d = d2(&xyz[3 * i], p);
                    if (d < bd || (d == bd && i < best)) { bd = d; best = i; }
                }
        if (best < 0) return -1;
        return sqrt(bd) <= max_dist ? best : -1;
    }
    /* element_mesh.py:121-137: all elements with d^2 <= r^2 (cKDTree compares squared distances) */
    void ball(const double* p, double r, std::vector<int>& out) {
        out.clear();
        if (size() == 0) return;
        if (dirty) rebuild();
        int x0 = cidx(p[0] - r), x1 = cidx(p[0] + r), y0 = cidx(p[1] - r), y1 = cidx(p[1] + r);
        double r2 = r * r;
        for (int cx = x0; cx <= x1; ++cx)
            for (int cy = y0; cy <= y1; ++cy)
                for (int i : cell[cx * G + cy])
                    if (d2(&xyz[3 * i], p) <= r2) out.push_back(i);
        std::sort(out.begin(), out.end());
    }
};

/* ------------------------------------------------------------------ trees -------------------- */
struct Node {
    double pos[3];
    double radius, kappa;
    int parent, child[2], nchild, tree;
};

struct Forest {
    std::vector<Node> nodes;
    std::vector<int> roots;
    int add(const double* p, double radius, int parent, double kappa, int tree) {
        Node n;
        memcpy(n.pos, p, sizeof(n.pos));
        n.radius = radius; n.kappa = kappa; n.parent = parent; n.nchild = 0; n.child[0] = n.child[1] = -1; n.tree = tree;
        nodes.push_back(n);
        int id = (int)nodes.size() - 1;
        if (parent >= 0) nodes[parent].child[nodes[parent].nchild++] = id;
        return id;
    }
};

struct Sim {
    OGConfig cfg;
    PyRandom py;
    NpRandom np;
    og_eig_hook eig_hook = nullptr;
    OGStats st;
    /* greenhouse state */
    double sigma_t, param_scale, d, r, FAZ_radius, rotation_radius, FAZ_center[2];
    int I, N;
    double eps_n, eps_s, eps_k, delta_art, delta_ven, gamma_art, gamma_ven, phi, omega, kappa, delta_sigma;
    double orig_scale[6];
    /* simulation space */
    double shape[3];
    int GSZ = 76; /* geometry_size */
    std::vector<int> valid_voxels; /* triples (i, j, k); k = 0 without a geometry file (the reference's positions miss the z dim there) */
    double ss_FAZ_center[2], ss_FAZ_radius;
    Forest F[2]; /* 0 arterial, 1 venous */
    PointList node_mesh[2], active_mesh[2];
    PointList oxy, co2;
    int sink_serial = 0;

    /* greenhouse.py:34-51 */
    void init_params(const OGMode& m) {
        I = m.I; N = m.N;
        eps_n = m.eps_n; eps_s = m.eps_s; eps_k = m.eps_k; delta_art = m.delta_art; delta_ven = m.delta_ven;
        gamma_art = m.gamma_art; gamma_ven = m.gamma_ven; phi = m.phi; omega = m.omega; kappa = m.kappa;
        delta_sigma = m.delta_sigma;
        sigma_t = 1;
        const double p[5] = {eps_k, eps_n, eps_s, delta_art, delta_ven};
        for (int i = 0; i < 5; ++i) orig_scale[i] = p[i] / param_scale;
        orig_scale[5] = d;
    }

    /* greenhouse.py:17-32 + simulation_space.py:16-54 */
    void init_greenhouse() {
        param_scale = cfg.param_scale;
        d = cfg.d / param_scale;
        r = cfg.r / param_scale;
        FAZ_radius = np.normal(cfg.faz_bound0 / param_scale, cfg.faz_bound1 / param_scale);
        rotation_radius = cfg.rotation_radius / param_scale;
        FAZ_center[0] = cfg.faz_center[0]; FAZ_center[1] = cfg.faz_center[1];
        const double nc[2] = {cfg.nerve_center[0] / param_scale, cfg.nerve_center[1] / param_scale};
        const double nr = cfg.nerve_radius / param_scale;
        if (cfg.geometry) { /* simulation_space.py:29-34: shape = geometry.shape / max(geometry.shape); valid = argwhere(geometry) */
            GSZ = std::max(cfg.geom_dims[0], std::max(cfg.geom_dims[1], cfg.geom_dims[2]));
            for (int k = 0; k < 3; ++k) shape[k] = (double)cfg.geom_dims[k] / (double)GSZ;
            valid_voxels.clear(); /* argwhere of the 3-D mask, C order: triples (i, j, k) */
            for (int i = 0; i < cfg.geom_dims[0]; ++i)
                for (int j = 0; j < cfg.geom_dims[1]; ++j)
                    for (int k = 0; k < cfg.geom_dims[2]; ++k)
                        if (cfg.geometry[((size_t)i * cfg.geom_dims[1] + j) * cfg.geom_dims[2] + k]) {
                            valid_voxels.push_back(i); valid_voxels.push_back(j); valid_voxels.push_back(k);
                        }
            init_params(cfg.modes[0]);
            return;
        }
        GSZ = 76;
        for (int k = 0; k < 3; ++k) shape[k] = cfg.size[k];
        const int GS = 76;
        ss_FAZ_center[0] = FAZ_center[0] * GS; ss_FAZ_center[1] = FAZ_center[1] * GS;
        ss_FAZ_radius = FAZ_radius * GS * 0.5;
        const int nx = (int)ceil(shape[0] * GS), ny = (int)ceil(shape[1] * GS);
        const bool nerve = (nc[0] - nr <= 1) && (nc[1] - nr <= 1); /* simulation_space.py:46 */
        const double ncv[2] = {nc[0] * GS, nc[1] * GS}, nrv = nr * GS;
        valid_voxels.clear();
        for (int i = 0; i < nx; ++i)       /* y_coords (axis 0) */
            for (int j = 0; j < ny; ++j) { /* x_coords (axis 1) */
                double a = (double)j - ss_FAZ_center[0], b = (double)i - ss_FAZ_center[1];
                bool ok = a * a + b * b > ss_FAZ_radius * ss_FAZ_radius;
                if (nerve) {
                    double c = (double)j - ncv[0], e = (double)i - ncv[1];
                    ok = ok && (c * c + e * e > nrv * nrv);
                }
                if (ok) { valid_voxels.push_back(i); valid_voxels.push_back(j); valid_voxels.push_back(0); }
            }
        init_params(cfg.modes[0]);
    }

    /* simulation_space.py:89-98 */
    bool is_valid_position(const double* p) const {
        for (int k = 0; k < 3; ++k)
            if (p[k] >= shape[k] || p[k] < 0) return false;
        if (cfg.geometry) { /* geometry[(pos * geometry_size).astype(uint16)] > 0 */
            const int vi = (int)(uint16_t)(p[0] * (double)GSZ), vj = (int)(uint16_t)(p[1] * (double)GSZ), vk = (int)(uint16_t)(p[2] * (double)GSZ);
            if (vi >= cfg.geom_dims[0] || vj >= cfg.geom_dims[1] || vk >= cfg.geom_dims[2]) abort(); /* numpy IndexError */
            return cfg.geometry[((size_t)vi * cfg.geom_dims[1] + vj) * cfg.geom_dims[2] + vk] != 0;
        }
        double a = p[0] - ss_FAZ_center[0], b = p[1] - ss_FAZ_center[1]; /* zip-truncated eukledian_dist */
        double s = 0 + pow(a, 2.0);
        s = s + pow(b, 2.0);
        return sqrt(s) > ss_FAZ_radius;
    }

    /* forest.py:68-181 */
    void init_stumps(int f) {
        const double d0 = d, r0 = r;
        for (int t = 0; t < cfg.n_trees; ++t) {
            int wall = cfg.walls[py.randbelow(cfg.n_walls)];
            double pos[3], dir[3];
            auto rng_dir = [&](double p, double size) { /* np.random.uniform(-1 if p-d0>0 else 0, 1 if p+d0<size else 0) */
                double lo = (p - d0 > 0) ? -1.0 : 0.0, hi = (p + d0 < size) ? 1.0 : 0.0;
                return np.uniform(lo, hi);
            };
            /* simulation_space.py:69-76, fixed geometry: random.choice over argwhere of the wall plane.  ax_index is
               `0 if first else self.shape[axis]-1` with the NORMALISED shape, i.e. 0.0 for a full-length axis: the far
               walls sample plane 0 as well. */
            auto fixed_wall = [&](int axis, double* a_out, double* b_out) {
                /* np.take(geometry, ax_index, axis) takes plane int(ax_index) = 0 (|shape[axis]-1| < 1 truncates to 0); argwhere of
                   that 2-D plane lists the two remaining axes in C order; the float ax_index only lands in the coordinate
                   that `del pos_3d[along_axis]` drops, but _vox_2_unit_pos still draws three uniforms. */
                const int u = axis == 0 ? 1 : 0, v = axis == 2 ? 1 : 2; /* remaining axes, in order */
                std::vector<int> cells;
                for (int p = 0; p < cfg.geom_dims[u]; ++p)
                    for (int q = 0; q < cfg.geom_dims[v]; ++q) {
                        int idx[3]; idx[axis] = 0; idx[u] = p; idx[v] = q;
                        if (cfg.geometry[((size_t)idx[0] * cfg.geom_dims[1] + idx[1]) * cfg.geom_dims[2] + idx[2]]) { cells.push_back(p); cells.push_back(q); }
                    }
                if (cells.empty()) abort(); /* random.choice of an empty sequence raises IndexError */
                const int c = py.randbelow((int)cells.size() / 2);
                double idx3[3]; idx3[axis] = 0.0; idx3[u] = (double)cells[2 * c]; idx3[v] = (double)cells[2 * c + 1];
                double p3[3];
                for (int k = 0; k < 3; ++k) p3[k] = (idx3[k] + (0.0 + 1.0 * np.dbl())) / (double)GSZ; /* _vox_2_unit_pos */
                *a_out = p3[u];
                *b_out = p3[v];
            };
            if (wall == 0 || wall == 1) {
                double y, z;
                if (cfg.geometry) fixed_wall(0, &y, &z);
                else { y = np.uniform(0, shape[1]); z = np.uniform(0, shape[2]); }
                pos[0] = wall == 0 ? 0.0 : shape[0] - 1e-6; pos[1] = y; pos[2] = z;
                dir[0] = wall == 0 ? np.uniform(0.1, 1) : np.uniform(-1, -0.1);
                dir[1] = rng_dir(y, shape[1]);
                dir[2] = rng_dir(z, shape[2]);
            } else if (wall == 2 || wall == 3) {
                double x, z;
                if (cfg.geometry) fixed_wall(1, &x, &z);
                else { x = np.uniform(0, shape[0]); z = np.uniform(0, shape[2]); }
                pos[0] = x; pos[1] = wall == 2 ? 0.0 : shape[1] - 1e-6; pos[2] = z;
                dir[0] = rng_dir(x, shape[0]);
                dir[1] = wall == 2 ? np.uniform(0.1, 1) : np.uniform(-1, -0.1);
                dir[2] = rng_dir(z, shape[2]);
            } else if (cfg.geometry) { /* forest.py:152-176: z0 / z1, both with first=True */
                double x, y;
                fixed_wall(2, &x, &y);
                pos[0] = x; pos[1] = y; pos[2] = wall == 4 ? 0.0 : shape[2] - 1e-6;
                dir[0] = rng_dir(x, shape[0]);
                dir[1] = rng_dir(y, shape[1]);
                dir[2] = wall == 4 ? np.uniform(0.1, 1) : np.uniform(-1, -0.1);
            } else {
                /* without a geometry file z0/z1 walls reference self.valid_pixels, which does not exist (simulation_space.py:83) */
                abort();
            }
            double nrm = norm3(dir);
            double child[3];
            for (int k = 0; k < 3; ++k) child[k] = pos[k] + dir[k] / nrm * d0; /* forest.py:101,104 */
            int root = F[f].add(pos, r0, -1, 4.0, t);
            F[f].roots.push_back(root);
            F[f].add(child, r0, root, 4.0, t); /* arterial_tree.py:218 default kappa=4 */
        }
    }

    /* forest.py:38-66 */
    void init_nerve(int f) {
        const double d0 = d, r0 = r;
        const double nc[2] = {cfg.nerve_center[0] / param_scale, cfg.nerve_center[1] / param_scale};
        const double nr = cfg.nerve_radius / param_scale;
        for (int t = 0; t < cfg.n_trees; ++t) {
            double alpha = 2 * M_PI * py.random();
            double rr = nr * sqrt(py.random());
            double pos[3] = {rr * cos(alpha) + nc[1], rr * sin(alpha) + nc[0], py.random() * shape[2]};
            double dir[3] = {py.random() - 0.5, 0, 0};
            dir[1] = py.random() - 0.5;
            double nrm = norm3(dir);
            double child[3];
            for (int k = 0; k < 3; ++k) child[k] = pos[k] + dir[k] / nrm * d0;
            int root = F[f].add(pos, r0, -1, 4.0, t);
            F[f].roots.push_back(root);
            F[f].add(child, r0, root, 4.0, t);
        }
    }

    /* arterial_tree.py:174-184 */
    void optimize_radius_to_root(int f, int n) {
        Forest& fo = F[f];
        while (true) {
            Node& nd = fo.nodes[n];
            if (nd.parent < 0 || nd.nchild == 0) return;
            double s = 0;
            for (int c = 0; c < nd.nchild; ++c) s = s + pow(fo.nodes[nd.child[c]].radius, nd.kappa);
            double rp = pow(s, 1 / nd.kappa);
            ++st.walk_steps;
            if (nd.radius == rp) return;
            nd.radius = rp;
            n = nd.parent;
        }
    }

    /* greenhouse.py:309-317 */
    double oxygen_distance(double radius) const {
        const double c_oxygen = 203.9e-3;
        const double kap = 0.02 * c_oxygen;
        const double r0 = 3.5e-3;
        double c1 = kap * (radius * param_scale / r0) * exp(1 - (radius * param_scale / r0));
        return c1 * 6 / param_scale;
    }

    /* simulation_space.py:57-67 + greenhouse.py:319-341 */
    void sample_oxygen_sinks(int n_try, double eps_n_eff, double eps_s_) {
        const uint32_t L = (uint32_t)(valid_voxels.size() / 3);
        std::vector<uint32_t> vi(n_try);
        for (int i = 0; i < n_try; ++i) vi[i] = np.randint(L);
        std::vector<double> cand;
        cand.reserve(3 * n_try);
        for (int i = 0; i < n_try; ++i) {
            double u0 = np.dbl(), u1 = np.dbl(), u2 = np.dbl(); /* uniform(0,1) = 0 + (1-0)*x */
            double p[3] = {((double)valid_voxels[3 * vi[i]] + (0.0 + 1.0 * u0)) / (double)GSZ,
                           ((double)valid_voxels[3 * vi[i] + 1] + (0.0 + 1.0 * u1)) / (double)GSZ,
                           ((double)valid_voxels[3 * vi[i] + 2] + (0.0 + 1.0 * u2)) / (double)GSZ};
            if (is_valid_position(p)) cand.insert(cand.end(), p, p + 3);
        }
        std::vector<double> added;
        std::vector<int> hits;
        const long ncand = (long)cand.size() / 3;
        st.sum_P += node_mesh[0].size(); /* per call, SURVEY 8(d) byte accounting */
        st.sum_S += oxy.size();
        for (long c = 0; c < ncand; ++c) {
            const double* p = &cand[3 * c];
            node_mesh[0].ball(p, eps_n_eff, hits);
            ++st.ball_queries;
            bool ok = true;
            for (int h : hits) {
                const Node& nd = F[0].nodes[node_mesh[0].id[h]];
                double s = 0 + pow(p[0] - nd.pos[0], 2.0);
                s = s + pow(p[1] - nd.pos[1], 2.0);
                s = s + pow(p[2] - nd.pos[2], 2.0);
                if (!(sqrt(s) > oxygen_distance(nd.radius))) { ok = false; break; }
            }
            if (!ok) continue;
            ++st.nn_queries;
            if (oxy.nearest_within(p, eps_s_) >= 0) continue;
            for (size_t a = 0; a < added.size() && ok; a += 3) {
                double df[3] = {p[0] - added[a], p[1] - added[a + 1], p[2] - added[a + 2]};
                if (!(norm3_axis(df) > eps_s_)) ok = false;
            }
            if (!ok) continue;
            added.insert(added.end(), p, p + 3);
        }
        for (size_t a = 0; a < added.size(); a += 3) oxy.push(&added[a], sink_serial++);
    }

    /* greenhouse.py:343-366 + 157-307 */
    void grow_vessels(int f, PointList& att_mesh, double gamma, double delta, bool first_mode, int t,
                      std::vector<int>& new_nodes) {
        new_nodes.clear();
        Forest& fo = F[f];
        PointList& am = active_mesh[f];
        /* assignment: dict in first-attractor order */
        std::vector<int> order;                 /* node ids in dict order */
        std::vector<std::vector<int>> lists;    /* attractor list positions per dict entry */
        std::vector<int> slot(fo.nodes.size(), -1);
        const long A = att_mesh.size();
        st.sum_A += A;
        st.sum_M += am.size();
        for (long a = 0; a < A; ++a) {
            ++st.nn_queries;
            long pos = am.nearest_within(&att_mesh.xyz[3 * a], delta);
            if (pos < 0) continue;
            int nid = am.id[pos];
            if (slot[nid] < 0) { slot[nid] = (int)order.size(); order.push_back(nid); lists.emplace_back(); }
            lists[slot[nid]].push_back((int)a);
        }
        std::vector<char> deactivate(fo.nodes.size(), 0);
        bool any_deact = false;
        std::vector<double> radius_at_start(fo.nodes.size());
        for (size_t i = 0; i < fo.nodes.size(); ++i) radius_at_start[i] = fo.nodes[i].radius;
        if ((long)order.size() > st.max_dict) st.max_dict = (long)order.size();
        for (auto& l : lists) if ((long)l.size() > st.max_list) st.max_list = (long)l.size();
        std::vector<double> ang, angp, unit, sel;
        for (size_t oi = 0; oi < order.size(); ++oi) {
            const int nid = order[oi];
            const std::vector<int>& al = lists[oi];
            const long n = (long)al.size();
            Node nd = fo.nodes[nid]; /* copy: vector may reallocate on add */
            const double vtc[2] = {FAZ_center[0] - nd.pos[0], FAZ_center[1] - nd.pos[1]};
            const double dist_to_center = norm2(vtc);
            auto rel = [&](long i, double* o) {
                const double* a = &att_mesh.xyz[3 * al[i]];
                o[0] = a[0] - nd.pos[0]; o[1] = a[1] - nd.pos[1]; o[2] = a[2] - nd.pos[2];
            };
            /* utilities.py:42-45 get_angle_between_vectors(u, V) */
            auto angles_to = [&](const double* u, std::vector<double>& out) {
                out.resize(n);
                double nu = norm3(u);
                for (long i = 0; i < n; ++i) {
                    double v[3];
                    rel(i, v);
                    double dt = (n == 1) ? ddot3(u, v) : gemv3(u, v);
                    double C = dt / nu / norm3_axis(v);
                    out[i] = RAD2DEG * acos(clamp11(C));
                }
            };
            const bool is_leaf = nd.nchild == 0;
            const bool is_inter = nd.parent >= 0 && nd.nchild == 1;
            if (is_leaf) {
                const Node& par = fo.nodes[nd.parent];
                double v[3] = {nd.pos[0] - par.pos[0], nd.pos[1] - par.pos[1], nd.pos[2] - par.pos[2]};
                angles_to(v, ang);
                const double lim = gamma / 2 > 0 ? gamma / 2 : 0;
                std::vector<long> keep;
                for (long i = 0; i < n; ++i) if (ang[i] <= lim) keep.push_back(i);
                if (keep.empty()) continue;
                double avg[3] = {0, 0, 0};
                bool firstv = true;
                for (long i : keep) {
                    double w[3];
                    rel(i, w);
                    double nw = norm3(w);
                    if (firstv) { avg[0] = w[0] / nw; avg[1] = w[1] / nw; avg[2] = w[2] / nw; firstv = false; }
                    else { avg[0] += w[0] / nw; avg[1] += w[1] / nw; avg[2] += w[2] / nw; }
                }
                sel.clear();
                for (long i : keep) sel.push_back(ang[i]);
                /* np.std: _var with pairwise sums */
                const long m = (long)sel.size();
                double mean = pairwise_sum(sel.data(), m) / (double)m;
                std::vector<double> dv(m);
                for (long i = 0; i < m; ++i) { double x = sel[i] - mean; dv[i] = x * x; }
                double sd = sqrt(pairwise_sum(dv.data(), m) / (double)m);
                bool bif = false;
                if (sd > phi) {
                    if (FAZ_radius == 0) bif = true;
                    else {
                        double u = py.uniform(0, 1);
                        ++st.py_draws;
                        if (pow(dist_to_center / (2 * FAZ_radius), 5.0) > u) bif = angle_between_two(vtc, avg) > 90;
                    }
                }
                if (bif) {
                    ++st.bifurcations;
                    const double r1 = r, r2 = r;
                    const double rp = pow(pow(r1, kappa) + pow(r2, kappa), 1 / kappa);
                    const double phi1 = RAD2DEG * acos((pow(rp, 4.0) + pow(r1, 4.0) - pow(r2, 4.0)) / (2 * pow(rp, 2.0) * pow(r1, 2.0)));
                    const double phi2 = RAD2DEG * acos((pow(rp, 4.0) + pow(r2, 4.0) - pow(r1, 4.0)) / (2 * pow(rp, 2.0) * pow(r2, 2.0)));
                    /* c = np.mean(atts, axis=0): sequential over rows */
                    double c[3] = {0, 0, 0};
                    for (size_t q = 0; q < keep.size(); ++q) {
                        const double* a = &att_mesh.xyz[3 * al[keep[q]]];
                        if (q == 0) { c[0] = a[0]; c[1] = a[1]; c[2] = a[2]; }
                        else { c[0] += a[0]; c[1] += a[1]; c[2] += a[2]; }
                    }
                    for (int k = 0; k < 3; ++k) c[k] /= (double)m;
                    double dpc[3] = {c[0] - nd.pos[0], c[1] - nd.pos[1], c[2] - nd.pos[2]};
                    double ndpc = norm3(dpc);
                    if (ndpc != 0.0) { double n2 = norm3(dpc); for (int k = 0; k < 3; ++k) dpc[k] /= n2; }
                    /* X = (atts - c).T ; np.cov(X) */
                    std::vector<double> X(3 * m);
                    for (long q = 0; q < m; ++q) {
                        const double* a = &att_mesh.xyz[3 * al[keep[q]]];
                        for (int k = 0; k < 3; ++k) X[k * m + q] = a[k] - c[k];
                    }
                    for (int k = 0; k < 3; ++k) {
                        double s = 0;
                        for (long q = 0; q < m; ++q) s = (q == 0) ? X[k * m] : s + X[k * m + q];
                        double av = s / (double)m;
                        for (long q = 0; q < m; ++q) X[k * m + q] -= av;
                    }
                    double cov[9], fact = 1.0 / (double)(m - 1);
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) {
                            double s = 0;
                            for (long q = 0; q < m; ++q) s = FMA(X[i * m + q], X[j * m + q], s);
                            cov[3 * i + j] = s * fact;
                        }
                    double w[3], V[9];
                    if (eig_hook) eig_hook(cov, w, V); else jacobi3(cov, w, V);
                    int am_i = 0;
                    for (int k = 1; k < 3; ++k) if (w[k] > w[am_i]) am_i = k;
                    double dl[3] = {V[am_i], V[3 + am_i], V[6 + am_i]};
                    double c1 = cos(DEG2RAD * phi1), s1 = sin(DEG2RAD * phi1), c2 = cos(DEG2RAD * phi2), s2 = sin(DEG2RAD * phi2);
                    double g1[3], g2[3], p1[3], p2[3];
                    for (int k = 0; k < 3; ++k) { g1[k] = c1 * dpc[k] + s1 * dl[k]; g2[k] = c2 * dpc[k] - s2 * dl[k]; }
                    double n1 = norm3(g1), n2 = norm3(g2);
                    for (int k = 0; k < 3; ++k) { p1[k] = nd.pos[k] + g1[k] / n1 * d; p2[k] = nd.pos[k] + g2[k] / n2 * d; }
                    new_nodes.push_back(fo.add(p1, r1, nid, kappa, nd.tree));
                    new_nodes.push_back(fo.add(p2, r2, nid, kappa, nd.tree));
                    optimize_radius_to_root(f, nid);
                    deactivate[nid] = 1; any_deact = true;
                } else {
                    ++st.elongations;
                    double nv = norm3(v), na = norm3(avg), g[3];
                    for (int k = 0; k < 3; ++k) g[k] = omega * (v[k] / nv) + (1 - omega) * (avg[k] / na);
                    if (rotation_radius > 0 && t > 15) {
                        double ng = norm3(g);
                        for (int k = 0; k < 3; ++k) g[k] /= ng;
                        double cv[2] = {FAZ_center[0] - nd.pos[0], FAZ_center[1] - nd.pos[1]};
                        double ncv = norm2(cv);
                        cv[0] /= ncv; cv[1] /= ncv;
                        double np2[2] = {FAZ_center[0] - (nd.pos[0] + d * g[0]), FAZ_center[1] - (nd.pos[1] + d * g[1])};
                        double dist_new = norm2(np2);
                        double floorw = first_mode ? 0.0 : 0.01;
                        double cand = rotation_radius - dist_new;
                        double weight = cand > floorw ? cand : floorw; /* max(a, b): b only if b > a */
                        weight = sqrt(weight);
                        double ort[3] = {-cv[1], cv[0], 0};
                        if (angle_between_two(g, ort) > 90) { ort[0] = -1 * ort[0]; ort[1] = -1 * ort[1]; ort[2] = -1 * ort[2]; }
                        double outv[3] = {-cv[0], -cv[1], 0};
                        for (int k = 0; k < 3; ++k) g[k] = ((1 - weight) * g[k] + 0.7 * weight * ort[k]) + 0.3 * weight * outv[k];
                    }
                    double ng = norm3(g), pk[3];
                    for (int k = 0; k < 3; ++k) pk[k] = nd.pos[k] + d * (g[k] / ng);
                    new_nodes.push_back(fo.add(pk, r, nid, kappa, nd.tree));
                }
            } else if (is_inter) {
                const Node& ch = fo.nodes[nd.child[0]];
                const Node& par = fo.nodes[nd.parent];
                const double r1 = ch.radius, r2 = r;
                ++st.inter_evals;
                if (r1 != radius_at_start[nd.child[0]]) ++st.inter_r1_changed;
                const double rp = pow(pow(r1, kappa) + pow(r2, kappa), 1 / kappa);
                const double phi1 = RAD2DEG * acos((pow(rp, 4.0) + pow(r1, 4.0) - pow(r2, 4.0)) / (2 * pow(rp, 2.0) * pow(r1, 2.0)));
                const double phi2 = RAD2DEG * acos((pow(rp, 4.0) + pow(r2, 4.0) - pow(r1, 4.0)) / (2 * pow(rp, 2.0) * pow(r2, 2.0)));
                double dseg[3] = {ch.pos[0] - nd.pos[0], ch.pos[1] - nd.pos[1], ch.pos[2] - nd.pos[2]};
                double pseg[3] = {nd.pos[0] - par.pos[0], nd.pos[1] - par.pos[1], nd.pos[2] - par.pos[2]};
                angles_to(dseg, ang);
                angles_to(pseg, angp);
                std::vector<long> keep;
                for (long i = 0; i < n; ++i)
                    if ((phi1 + phi2 - gamma / 2 <= ang[i]) && (ang[i] <= (phi1 + phi2 + gamma / 2)) && (angp[i] <= phi2 + gamma / 2))
                        keep.push_back(i);
                if (keep.empty()) continue;
                double avg[3] = {0, 0, 0};
                bool firstv = true;
                for (long i : keep) {
                    double w[3];
                    rel(i, w);
                    double nw = norm3(w);
                    if (firstv) { avg[0] = w[0] / nw; avg[1] = w[1] / nw; avg[2] = w[2] / nw; firstv = false; }
                    else { avg[0] += w[0] / nw; avg[1] += w[1] / nw; avg[2] += w[2] / nw; }
                }
                double nds = norm3(dseg), dv[3] = {dseg[0] / nds, dseg[1] / nds, dseg[2] / nds};
                double cr[3] = {dv[1] * avg[2] - dv[2] * avg[1], dv[2] * avg[0] - dv[0] * avg[2], dv[0] * avg[1] - dv[1] * avg[0]};
                if (cr[0] == 0 && cr[1] == 0 && cr[2] == 0) continue;
                double u = py.uniform(0, 1);
                ++st.py_draws;
                if (pow(dist_to_center / (2 * FAZ_radius), 5.0) <= u && angle_between_two(vtc, avg) <= 90) continue;
                ++st.sprouts;
                double ncr = norm3(cr), ax[3] = {cr[0] / ncr, cr[1] / ncr, cr[2] / ncr};
                double theta = phi2;
                double ct = cos(DEG2RAD * theta), sth = sin(DEG2RAD * theta);
                double kxv[3] = {ax[1] * dv[2] - ax[2] * dv[1], ax[2] * dv[0] - ax[0] * dv[2], ax[0] * dv[1] - ax[1] * dv[0]};
                double kdv = ddot3(ax, dv);
                double vv[3];
                for (int k = 0; k < 3; ++k) vv[k] = (dv[k] * ct + kxv[k] * sth) + ax[k] * kdv * (1 - ct);
                double nvv = norm3(vv), na = norm3(avg), g[3];
                for (int k = 0; k < 3; ++k) g[k] = omega * (vv[k] / nvv) + (1 - omega) * (avg[k] / na);
                double ng = norm3(g), pk[3];
                for (int k = 0; k < 3; ++k) pk[k] = nd.pos[k] + d * (g[k] / ng);
                new_nodes.push_back(fo.add(pk, r, nid, kappa, nd.tree));
                optimize_radius_to_root(f, nid);
                deactivate[nid] = 1; any_deact = true;
            }
        }
        if (any_deact) {
            std::vector<char> kill(am.size(), 0);
            for (long i = 0; i < am.size(); ++i) kill[i] = deactivate[am.id[i]];
            am.remove_positions(kill);
        }
    }

    void extend_meshes(int f, const std::vector<int>& nn) {
        for (int id : nn) { node_mesh[f].push(F[f].nodes[id].pos, id); active_mesh[f].push(F[f].nodes[id].pos, id); }
    }

    /* greenhouse.py:139-147 */
    void expansion() {
        sigma_t = sigma_t + delta_sigma;
        eps_k = orig_scale[0] / sigma_t; eps_n = orig_scale[1] / sigma_t; eps_s = orig_scale[2] / sigma_t;
        delta_art = orig_scale[3] / sigma_t; delta_ven = orig_scale[4] / sigma_t; d = orig_scale[5] / sigma_t;
        double floor_d = 0.04 / param_scale;
        d = d > floor_d ? d : floor_d;
    }

    void develop(og_trace_hook trace) {
        for (int f = 0; f < (cfg.venous ? 2 : 1); ++f) {
            /* list(forest.get_nodes()): per tree, level order */
            for (size_t tr = 0; tr < F[f].roots.size(); ++tr) {
                int root = F[f].roots[tr];
                node_mesh[f].push(F[f].nodes[root].pos, root);
                active_mesh[f].push(F[f].nodes[root].pos, root);
                int ch = F[f].nodes[root].child[0];
                node_mesh[f].push(F[f].nodes[ch].pos, ch);
                active_mesh[f].push(F[f].nodes[ch].pos, ch);
            }
        }
        int t = 0;
        std::vector<int> new_nodes, hits;
        KdOrder kd;
        for (int mi = 0; mi < cfg.n_modes; ++mi) {
            const OGMode& mode = cfg.modes[mi];
            if (mode.reinit) init_params(mode);
            if (I <= 0) continue;
            const int t_end = t + I;
            const int t_start = t;
            for (t = t_start; t < t_end; ++t) {
                sample_oxygen_sinks(N, eps_n > eps_k ? eps_n : eps_k, eps_s);
                grow_vessels(0, oxy, gamma_art, delta_art, mode.first_mode != 0, t, new_nodes);
                extend_meshes(0, new_nodes);
                /* step 3, greenhouse.py:99-112 */
                if (!new_nodes.empty()) {
                    std::vector<char> to_remove(oxy.size(), 0);
                    PySet to_add, shadow; /* shadow: list-index order, instrumentation only */
                    bool kd_built = false, it_cons = false, it_perm = false;
                    std::vector<int> hidx;
                    for (int nid : new_nodes) {
                        oxy.ball(F[0].nodes[nid].pos, eps_k, hits);
                        ++st.ball_queries;
                        hidx = hits;
                        if (hits.size() > 1) {
                            ++st.multi_balls;
                            if (cfg.ball_order == 0) {
                                if (!kd_built) { kd.build(oxy.xyz.data(), oxy.size()); kd_built = true; ++st.kd_builds; }
                                std::sort(hits.begin(), hits.end(), [&](int a, int b) { return kd.rank[a] < kd.rank[b]; });
                                if (hits != hidx) ++st.reordered_balls;
                            }
                        }
                        for (int h : hits) {
                            to_remove[h] = 1;
                            if (cfg.venous) {
                                ++st.nn_queries;
                                if (node_mesh[1].nearest_within(&oxy.xyz[3 * h], eps_k) < 0)
                                    to_add.add(h, py_hash_tuple3(&oxy.xyz[3 * h]));
                            }
                        }
                        /* conservative order-sensitivity flag on the index-order shadow set */
                        {
                            std::vector<std::vector<size_t>> ex;
                            std::vector<long> fin;
                            std::vector<int> fresh;
                            bool resized = false;
                            const PySet before = shadow;
                            for (int h : hidx)
                                if (cfg.venous && node_mesh[1].nearest_within(&oxy.xyz[3 * h], eps_k) < 0) {
                                    ex.emplace_back();
                                    long slot = shadow.add(h, py_hash_tuple3(&oxy.xyz[3 * h]), &ex.back(), &resized);
                                    if (slot < 0) ex.pop_back(); else { fin.push_back(slot); fresh.push_back(h); }
                                }
                            bool inter = false;
                            if (fin.size() > 1) {
                                if (resized) inter = true;
                                for (size_t a = 0; a < fin.size() && !inter; ++a)
                                    for (size_t b = 0; b < fin.size() && !inter; ++b)
                                        if (a != b)
                                            for (size_t sl : ex[b]) if ((long)sl == fin[a]) { inter = true; break; }
                            }
                            if (inter) {
                                ++st.interacting_groups;
                                it_cons = true;
                                /* permutation closure: does ANY order of this ball's fresh keys change the table? */
                                bool differs = fresh.size() > 4 || (resized && trial_mid_resize(before, fresh, oxy.xyz.data()));
                                if (!differs) {
                                    ++st.perm_groups;
                                    std::vector<int> perm = fresh;
                                    std::sort(perm.begin(), perm.end());
                                    do {
                                        PySet trial = before;
                                        for (int h : perm) trial.add(h, py_hash_tuple3(&oxy.xyz[3 * h]));
                                        if (trial.table.size() != shadow.table.size()) { differs = true; break; }
                                        for (size_t z = 0; z < trial.table.size(); ++z)
                                            if (trial.table[z].key != shadow.table[z].key) { differs = true; break; }
                                    } while (!differs && std::next_permutation(perm.begin(), perm.end()));
                                }
                                if (differs) it_perm = true;
                            }
                        }
                    }
                    if (it_cons) ++st.flag_iters_cons;
                    if (it_perm) ++st.flag_iters_perm;
                    if (cfg.ball_order == 0) {
                        std::vector<int> oa, ob;
                        to_add.for_each([&](int h) { oa.push_back(h); });
                        shadow.for_each([&](int h) { ob.push_back(h); });
                        if (oa != ob) ++st.diff_iters;
                    }
                    to_add.for_each([&](int h) { co2.push(&oxy.xyz[3 * h], oxy.id[h]); });
                    oxy.remove_positions(to_remove);
                }
                if (cfg.venous) {
                    grow_vessels(1, co2, gamma_ven, delta_ven, mode.first_mode != 0, t, new_nodes);
                    extend_meshes(1, new_nodes);
                    if (!new_nodes.empty()) {
                        std::vector<char> to_remove(co2.size(), 0);
                        for (int nid : new_nodes) {
                            co2.ball(F[1].nodes[nid].pos, eps_k, hits);
                            ++st.ball_queries;
                            for (int h : hits) to_remove[h] = 1;
                        }
                        co2.remove_positions(to_remove);
                    }
                }
                expansion();
                if (trace) trace(t, node_mesh[0].size(), oxy.size(), node_mesh[1].size(), co2.size(), st.py_draws, np.g.drawn);
            }
            t = t_end - 1; /* `for t in range(t, t+self.I)` leaves t at the last value (greenhouse.py:90) */
        }
    }

    /* generate_vessel_graph.py:45-56: per tree LevelOrderIter, root excluded */
    long export_edges(int f, double* out, long cap) const {
        long n = 0;
        const Forest& fo = F[f];
        std::vector<int> level, next;
        for (int root : fo.roots) {
            level.assign(1, root);
            while (!level.empty()) {
                next.clear();
                for (int id : level) {
                    const Node& nd = fo.nodes[id];
                    if (nd.parent >= 0) {
                        if (n < cap) {
                            const Node& pa = fo.nodes[nd.parent];
                            double* o = out + 7 * n;
                            o[0] = nd.pos[0]; o[1] = nd.pos[1]; o[2] = nd.pos[2];
                            o[3] = pa.pos[0]; o[4] = pa.pos[1]; o[5] = pa.pos[2];
                            o[6] = nd.radius;
                        }
                        ++n;
                    }
                    for (int c = 0; c < nd.nchild; ++c) next.push_back(nd.child[c]);
                }
                level.swap(next);
            }
        }
        return n;
    }
};

}  // namespace

extern "C" {

/* Runs one seeded sample.  edges7_out receives arterial rows then venous rows (cap rows available).
 * Returns total rows (may exceed cap -> call again with a larger buffer), or <0 on error. */
/* final sink lists of the calling thread's last run: what Greenhouse.save_stats scatters (greenhouse.py:401-418) */
static thread_local std::vector<double> g_last_sinks[2];

long og_last_sinks(int which, double* xyz_out, long cap) {
    if (which < 0 || which > 1) return -1;
    const long n = (long)g_last_sinks[which].size() / 3;
    if (xyz_out && cap >= n) memcpy(xyz_out, g_last_sinks[which].data(), sizeof(double) * 3 * (size_t)n);
    return n;
}

long growth_oracle_run(const OGConfig* cfg, uint64_t seed, double* edges7_out, long cap, long* n_art_edges,
                       long* n_ven_edges, OGStats* stats, og_eig_hook eig, og_trace_hook trace) {
    if (!cfg || cfg->n_modes < 1 || cfg->n_modes > 8 || cfg->n_walls < 0 || cfg->n_walls > 6) return -1;
    Sim* s = new Sim();
    s->cfg = *cfg;
    memset(&s->st, 0, sizeof(s->st));
    s->eig_hook = eig;
    s->py.seed(seed);
    s->np.seed((uint32_t)seed);
    s->init_greenhouse();
    for (int f = 0; f < (cfg->venous ? 2 : 1); ++f) {
        if (cfg->forest_type == 0) s->init_stumps(f); else s->init_nerve(f);
    }
    s->develop(trace);
    long na = s->export_edges(0, edges7_out, cap);
    long nv = cfg->venous ? s->export_edges(1, edges7_out + 7 * (na < cap ? na : cap), cap - (na < cap ? na : cap)) : 0;
    if (n_art_edges) *n_art_edges = na;
    if (n_ven_edges) *n_ven_edges = nv;
    s->st.n_art_nodes = (long)s->F[0].nodes.size();
    s->st.n_ven_nodes = (long)s->F[1].nodes.size();
    s->st.n_oxy_left = s->oxy.size();
    s->st.n_co2_left = s->co2.size();
    s->st.np_u32 = s->np.g.drawn;
    if (stats) *stats = s->st;
    g_last_sinks[0] = s->oxy.xyz;
    g_last_sinks[1] = s->co2.xyz;
    delete s;
    return na + nv;
}

/* exposed for unit tests against CPython / numpy */
int64_t og_hash_tuple3(const double* p) { return py_hash_tuple3(p); }
long og_set_order(const double* pts, long n, long* order_out) {
    PySet s;
    for (long i = 0; i < n; ++i) {
        /* equality: identical coordinates == same key (first occurrence wins) */
        long key = i;
        for (long j = 0; j < i; ++j)
            if (pts[3 * j] == pts[3 * i] && pts[3 * j + 1] == pts[3 * i + 1] && pts[3 * j + 2] == pts[3 * i + 2]) { key = j; break; }
        s.add((int)key, py_hash_tuple3(pts + 3 * i));
    }
    long k = 0;
    s.for_each([&](int key) { order_out[k++] = key; });
    return k;
}
void og_kd_indices(const double* pts, long n, long* idx_out) {
    KdOrder kd;
    kd.build(pts, n);
    for (long i = 0; i < n; ++i) idx_out[i] = kd.idx[i];
}
void og_py_random(uint64_t seed, long n, double* out) {
    PyRandom r; r.seed(seed);
    for (long i = 0; i < n; ++i) out[i] = r.random();
}
void og_np_stream(uint32_t seed, double* normal_out, long n_int, uint32_t bound, uint32_t* ints_out, long n_dbl, double* dbl_out) {
    NpRandom r; r.seed(seed);
    *normal_out = r.normal(0.25, 0.5);
    for (long i = 0; i < n_int; ++i) ints_out[i] = r.randint(bound);
    for (long i = 0; i < n_dbl; ++i) dbl_out[i] = r.uniform(0, 1);
}
int og_py_choice(uint64_t seed, int n, long count, int* out) {
    PyRandom r; r.seed(seed);
    for (long i = 0; i < count; ++i) out[i] = r.randbelow(n);
    return 0;
}
}
