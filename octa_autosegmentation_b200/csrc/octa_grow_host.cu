// Host driver of the growth path: per-sample initialisation (Greenhouse.__init__, SimulationSpace,
// Forest stumps -- a few dozen RNG draws, done on the host and handed to the device together with
// the two MT19937 states), the per-iteration launch sequence, and the export:
//   * exact Murray radii: the device keeps radii with CUDA's pow (<= 2 ULP), which only steers
//     angles; the radii that are PRINTED (repr, 17 digits) are recomputed here by replaying the
//     branch events with the C library's pow -- the same libm CPython's float.__pow__ calls
//     (arterial_tree.py:180) -- in creation order, multi-threaded over graphs;
//   * edge rows in the reference's order (per tree, level order, children in attach order,
//     generate_vessel_graph.py:45-56).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <mutex>
#include <sched.h>
#include <thread>
#include <vector>
#include "octa_common.h"
#include "octa_grow.cuh"

namespace octa {

void grow_timing_begin(cudaStream_t st);
void grow_timing_report();
bool grow_timing_enabled();
struct GrowEvents { cudaEvent_t start, sinks, kd, killa; };
void launch_begin(int dslot, const GrowShape& S, const IterP& P0, int n_sm, cudaStream_t st, cudaStream_t side, const GrowEvents& ev);
void launch_iteration(int dslot, const GrowShape& S, const int commit_smem[2], const IterP& P, const IterP* Pnext, int n_sm, cudaStream_t st,
                      cudaStream_t side, const GrowEvents& ev);
int max_ctx_slots();
int upload_dev_table(int dslot, const GrowDev& D, cudaStream_t st);
int prepare_kernels(const GrowShape& S);

namespace {

struct HostGraphInit {
    MTState np_mt, py_mt;
    double faz_radius;
    std::vector<unsigned char> valid_ij;
    int n_valid;
    // forests: nodes in creation order
    std::vector<double> pos[2];   // 3 per node
    std::vector<int> parent[2];
};

double np_uniform(MTState& s, double lo, double hi) { return lo + (hi - lo) * mt_next_double_host(s); }

// legacy_gauss (polar method); the cached second variate is never consumed on this path
double np_normal(MTState& s, double loc, double scale) {
    double x1, x2, r2;
    do {
        x1 = 2.0 * mt_next_double_host(s) - 1.0;
        x2 = 2.0 * mt_next_double_host(s) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = sqrt(-2.0 * log(r2) / r2);
    return loc + scale * (f * x2);
}

int py_randbelow(MTState& s, int n) {
    int k = 0;
    for (int v = n; v; v >>= 1) ++k;
    uint32_t r = mt_next_host(s) >> (32 - k);
    while ((int)r >= n) r = mt_next_host(s) >> (32 - k);
    return (int)r;
}

inline double hnorm3(const double* a) { return sqrt(fma(a[2], a[2], fma(a[1], a[1], a[0] * a[0]))); }   // ddot-based norm

// greenhouse.py:17-32, simulation_space.py:16-54, forest.py:38-181
int init_graph(const OctaGrowConfig& c, int geom_nvalid, uint64_t seed, HostGraphInit* h) {
    const uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    mt_init_by_array(h->py_mt, key, key[1] ? 2 : 1);
    mt_init_genrand(h->np_mt, (uint32_t)seed);
    const double ps = c.param_scale;
    const double d0 = c.d / ps, r0 = c.r / ps;
    (void)r0;
    h->faz_radius = np_normal(h->np_mt, c.faz_radius_bound[0] / ps, c.faz_radius_bound[1] / ps);
    const double nc[2] = {c.nerve_center[0] / ps, c.nerve_center[1] / ps};
    const double nr = c.nerve_radius / ps;
    const double fc[2] = {c.faz_center[0] * GEOMETRY_SIZE, c.faz_center[1] * GEOMETRY_SIZE};
    const double fr = h->faz_radius * GEOMETRY_SIZE * 0.5;
    const int nx = (int)ceil(c.size[0] * GEOMETRY_SIZE), ny = (int)ceil(c.size[1] * GEOMETRY_SIZE);
    if (nx > GEOMETRY_SIZE || ny > GEOMETRY_SIZE || nx < 1 || ny < 1) { set_error("simulation space larger than the unit square"); return OCTA_E_ARG; }
    const bool nerve = (nc[0] - nr <= 1) && (nc[1] - nr <= 1);
    const double ncv[2] = {nc[0] * GEOMETRY_SIZE, nc[1] * GEOMETRY_SIZE}, nrv = nr * GEOMETRY_SIZE;
    h->valid_ij.clear();
    // fixed geometry: valid_voxels = argwhere(mask) (simulation_space.py:34) is the same for every graph -- the context holds
    // one device copy (GrowDev::geom_valid) and only its length travels per graph
    const int gn = c.geometry ? std::max(c.geom_dims[0], std::max(c.geom_dims[1], c.geom_dims[2])) : 0;   // geometry_size
    if (!gn)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            const double a = (double)j - fc[0], b = (double)i - fc[1];
            bool ok = a * a + b * b > fr * fr;
            if (nerve) { const double e = (double)j - ncv[0], f = (double)i - ncv[1]; ok = ok && (e * e + f * f > nrv * nrv); }
            if (ok) { h->valid_ij.push_back((unsigned char)i); h->valid_ij.push_back((unsigned char)j); }
        }
    h->n_valid = gn ? geom_nvalid : (int)h->valid_ij.size() / 2;
    if (h->n_valid == 0) { set_error("no valid sampling voxel"); return OCTA_E_ARG; }
    for (int f = 0; f < 2; ++f) {
        h->pos[f].clear(); h->parent[f].clear();
        for (int t = 0; t < c.n_trees; ++t) {
            double pos[3], dir[3];
            if (c.forest_type == 0) {
                if (c.n_walls <= 0) { set_error("no source wall enabled"); return OCTA_E_ARG; }
                const int wall = c.walls[py_randbelow(h->py_mt, c.n_walls)];
                auto rng_dir = [&](double p, double size) {
                    const double lo = (p - d0 > 0) ? -1.0 : 0.0, hi = (p + d0 < size) ? 1.0 : 0.0;
                    return np_uniform(h->np_mt, lo, hi);
                };
                // fixed geometry (simulation_space.py:69-76): random.choice over argwhere of the wall plane, then
                // _vox_2_unit_pos (three np.random.uniform(0,1) draws).  The plane index is `0 if first else shape[axis]-1`
                // with the NORMALISED shape, a float in (-1, 0] that np.take truncates to 0: the far walls sample plane 0
                // as well, and the float only lands in the coordinate that `del pos_3d[along_axis]` drops.  argwhere of
                // the 2-D plane lists the two remaining axes (u < v) in C order.
                auto fixed_wall = [&](int axis, double* a_out, double* b_out) -> bool {
                    const int u = axis == 0 ? 1 : 0, v = axis == 2 ? 1 : 2;
                    const int* gd = c.geom_dims;
                    std::vector<int> cells;
                    for (int p = 0; p < gd[u]; ++p)
                        for (int q = 0; q < gd[v]; ++q) {
                            int idx[3]; idx[axis] = 0; idx[u] = p; idx[v] = q;
                            if (c.geometry[((size_t)idx[0] * gd[1] + idx[1]) * gd[2] + idx[2]]) { cells.push_back(p); cells.push_back(q); }
                        }
                    if (cells.empty()) return false;
                    const int pick = py_randbelow(h->py_mt, (int)cells.size() / 2);
                    double idx3[3]; idx3[axis] = 0.0; idx3[u] = (double)cells[2 * pick]; idx3[v] = (double)cells[2 * pick + 1];
                    double p3[3];
                    for (int k = 0; k < 3; ++k) p3[k] = (idx3[k] + np_uniform(h->np_mt, 0, 1)) / (double)gn;
                    *a_out = p3[u];
                    *b_out = p3[v];
                    return true;
                };
                if (gn && wall >= 4) {      // forest.py:152-176: z0 / z1 (both ask for first=True); usable only with a geometry file
                    double x, y;
                    if (!fixed_wall(2, &x, &y)) { set_error("geometry mask: the wall plane has no valid voxel"); return OCTA_E_ARG; }
                    pos[0] = x; pos[1] = y; pos[2] = wall == 4 ? 0.0 : c.size[2] - 1e-6;
                    dir[0] = rng_dir(x, c.size[0]);
                    dir[1] = rng_dir(y, c.size[1]);
                    dir[2] = wall == 4 ? np_uniform(h->np_mt, 0.1, 1) : np_uniform(h->np_mt, -1, -0.1);
                } else if (gn) {
                    double a, z;
                    if (!fixed_wall(wall < 2 ? 0 : 1, &a, &z)) { set_error("geometry mask: the wall plane has no valid voxel"); return OCTA_E_ARG; }
                    if (wall < 2) {
                        pos[0] = wall == 0 ? 0.0 : c.size[0] - 1e-6; pos[1] = a; pos[2] = z;
                        dir[0] = wall == 0 ? np_uniform(h->np_mt, 0.1, 1) : np_uniform(h->np_mt, -1, -0.1);
                        dir[1] = rng_dir(a, c.size[1]);
                        dir[2] = rng_dir(z, c.size[2]);
                    } else {
                        pos[0] = a; pos[1] = wall == 2 ? 0.0 : c.size[1] - 1e-6; pos[2] = z;
                        dir[0] = rng_dir(a, c.size[0]);
                        dir[1] = wall == 2 ? np_uniform(h->np_mt, 0.1, 1) : np_uniform(h->np_mt, -1, -0.1);
                        dir[2] = rng_dir(z, c.size[2]);
                    }
                } else if (wall == 0 || wall == 1) {
                    const double y = np_uniform(h->np_mt, 0, c.size[1]), z = np_uniform(h->np_mt, 0, c.size[2]);
                    pos[0] = wall == 0 ? 0.0 : c.size[0] - 1e-6; pos[1] = y; pos[2] = z;
                    dir[0] = wall == 0 ? np_uniform(h->np_mt, 0.1, 1) : np_uniform(h->np_mt, -1, -0.1);
                    dir[1] = rng_dir(y, c.size[1]);
                    dir[2] = rng_dir(z, c.size[2]);
                } else if (wall == 2 || wall == 3) {
                    const double x = np_uniform(h->np_mt, 0, c.size[0]), z = np_uniform(h->np_mt, 0, c.size[2]);
                    pos[0] = x; pos[1] = wall == 2 ? 0.0 : c.size[1] - 1e-6; pos[2] = z;
                    dir[0] = rng_dir(x, c.size[0]);
                    dir[1] = wall == 2 ? np_uniform(h->np_mt, 0.1, 1) : np_uniform(h->np_mt, -1, -0.1);
                    dir[2] = rng_dir(z, c.size[2]);
                } else {
                    // without a geometry file the reference's z0/z1 branch dereferences an attribute that does not exist (simulation_space.py:83)
                    set_error("source walls z0/z1 need SimulationSpace.oxygen_sample_geometry_path (the reference raises AttributeError without it)");
                    return OCTA_E_ARG;
                }
            } else {
                const double alpha = 2 * M_PI * mt_next_double_host(h->py_mt);
                const double rr = nr * sqrt(mt_next_double_host(h->py_mt));
                pos[0] = rr * cos(alpha) + nc[1];
                pos[1] = rr * sin(alpha) + nc[0];
                pos[2] = mt_next_double_host(h->py_mt) * c.size[2];
                dir[0] = mt_next_double_host(h->py_mt) - 0.5;
                dir[1] = mt_next_double_host(h->py_mt) - 0.5;
                dir[2] = 0;
            }
            const double nrm = hnorm3(dir);
            double child[3];
            for (int k = 0; k < 3; ++k) child[k] = pos[k] + dir[k] / nrm * d0;
            const int root = (int)h->parent[f].size();
            h->pos[f].insert(h->pos[f].end(), pos, pos + 3); h->parent[f].push_back(-1);
            h->pos[f].insert(h->pos[f].end(), child, child + 3); h->parent[f].push_back(root);
        }
    }
    return OCTA_OK;
}

// Mode parameters as the reference sees them: init_params_from_config runs only for a mode whose NAME differs from the first
// mode's (greenhouse.py:84-85), so a later mode that reuses that name keeps gamma / phi / omega / kappa (and eps / delta / I / N)
// of whatever mode was initialised last.  eff[mi] = index of the mode whose parameters are live during mode mi.
void effective_modes(const OctaGrowConfig& c, int eff[8]) {
    int cur = 0;
    for (int mi = 0; mi < c.n_modes && mi < 8; ++mi) {
        if (c.modes[mi].reinit) cur = mi;
        eff[mi] = cur;
    }
}

// greenhouse.py:34-51 / :83-90 / :139-147 -> one IterP per iteration
void build_schedule(const OctaGrowConfig& c, std::vector<IterP>* out) {
    int eff[8] = {0};
    effective_modes(c, eff);
    const double ps = c.param_scale;
    double d = c.d / ps;
    const double r = c.r / ps;
    double eps_n, eps_s, eps_k, delta_art, delta_ven, sigma_t, orig[6], delta_sigma;
    int I, N;
    auto init_params = [&](const OctaGrowMode& m) {
        I = m.I; N = m.N;
        eps_n = m.eps_n; eps_s = m.eps_s; eps_k = m.eps_k; delta_art = m.delta_art; delta_ven = m.delta_ven;
        delta_sigma = m.delta_sigma;
        sigma_t = 1;
        const double p[5] = {eps_k, eps_n, eps_s, delta_art, delta_ven};
        for (int i = 0; i < 5; ++i) orig[i] = p[i] / ps;
        orig[5] = d;
    };
    init_params(c.modes[0]);
    int t = 0, iter = 0;
    for (int mi = 0; mi < c.n_modes; ++mi) {
        const OctaGrowMode& m0 = c.modes[mi];
        if (m0.reinit) init_params(m0);
        const OctaGrowMode& m = c.modes[eff[mi]];       // gamma / phi / omega / kappa change with init_params only
        if (I <= 0) continue;
        const int t_end = t + I;
        for (; t < t_end; ++t) {
            IterP P;
            memset(&P, 0, sizeof(P));
            P.eps_n_eff = eps_n > eps_k ? eps_n : eps_k; P.eps_s = eps_s; P.eps_k = eps_k;
            P.delta[0] = delta_art; P.delta[1] = delta_ven; P.gamma[0] = m.gamma_art; P.gamma[1] = m.gamma_ven;
            P.phi = m.phi; P.omega = m.omega; P.kappa = m.kappa; P.d = d; P.r = r;
            P.rotation_radius = c.rotation_radius / ps; P.faz_cx = c.faz_center[0]; P.faz_cy = c.faz_center[1];
            P.param_scale = ps;
            for (int k = 0; k < 3; ++k) P.shape[k] = c.size[k];
            P.N = N; P.t = t; P.first_mode = m0.first_mode; P.mode_idx = mi; P.iter = iter++;
            P.geom_gs = c.geometry ? std::max(c.geom_dims[0], std::max(c.geom_dims[1], c.geom_dims[2])) : 0;
            for (int k = 0; k < 3; ++k) P.geom_dims[k] = c.geometry ? c.geom_dims[k] : 0;
            for (int q = 0; q < 8; ++q) P.kap_tab[q] = q < c.n_modes ? c.modes[eff[q]].kappa : 4.0;
            P.kap_tab[8] = 4.0;
            for (int q = 0; q < 9; ++q) P.leafc_tab[q] = pow(P.r, P.kap_tab[q]);
            out->push_back(P);
            sigma_t = sigma_t + delta_sigma;
            eps_k = orig[0] / sigma_t; eps_n = orig[1] / sigma_t; eps_s = orig[2] / sigma_t;
            delta_art = orig[3] / sigma_t; delta_ven = orig[4] / sigma_t; d = orig[5] / sigma_t;
            const double floor_d = 0.04 / ps;
            d = d > floor_d ? d : floor_d;
        }
        t = t_end - 1;
    }
}

struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(char* b) : base(b) {}
    template <class T> T* take(size_t n) {
        size_t o = off;
        off = align_up(off + sizeof(T) * n, 256);
        return base ? (T*)(base + o) : nullptr;
    }
};

void carve(Carver& c, const GrowShape& S, GrowDev* D) {
    const size_t GN = (size_t)S.G * S.capN, GS = (size_t)S.G * S.capS, GC = (size_t)S.G * S.Nmax, G = S.G;
    for (int f = 0; f < 2; ++f) {
        D->nx[f] = c.take<double>(GN); D->ny[f] = c.take<double>(GN); D->nz[f] = c.take<double>(GN);
        D->ncon[f] = c.take<double>(GN);
        D->npar[f] = c.take<int>(GN); D->nch0[f] = c.take<int>(GN); D->nch1[f] = c.take<int>(GN);
        D->nnch[f] = c.take<unsigned char>(GN); D->nmeta[f] = c.take<unsigned char>(GN); D->deact[f] = c.take<unsigned char>(GN);
        D->n_nodes[f] = c.take<int>(G); D->n_prev[f] = c.take<int>(G);
        D->n_act[f] = c.take<int>(G);
        D->sx[f] = c.take<double>(GS); D->sy[f] = c.take<double>(GS); D->sz[f] = c.take<double>(GS);
        D->n_s[f] = c.take<int>(G);
    }
    const size_t GG = (size_t)S.G * (S.capN > S.capS ? S.capN : S.capS);
    for (int w = 0; w < 5; ++w) {
        D->gx[w] = c.take<double>(GG); D->gy[w] = c.take<double>(GG); D->gz[w] = c.take<double>(GG); D->gr[w] = c.take<double>(GG);
        D->gi[w] = c.take<int>(GG); D->gcell[w] = c.take<int>(G * (GRID * GRID + 1));
    }
    D->np_mt = c.take<MTState>(G); D->py_mt = c.take<MTState>(G);
    D->py_buf = c.take<unsigned int>(G * S.pycap); D->py_n = c.take<int>(G); D->py_pos = c.take<int>(G);
    D->py_draws = c.take<long long>(G);
    D->faz_radius = c.take<double>(G); D->n_valid = c.take<int>(G); D->valid_ij = c.take<unsigned char>(G * MAX_VALID * 2);
    D->geom_mask = c.take<unsigned char>(S.geom_cells > MAX_VALID ? (size_t)S.geom_cells : (size_t)MAX_VALID);
    D->geom_valid = c.take<unsigned short>(3 * (size_t)(S.geom_nvalid > 0 ? S.geom_nvalid : 1));
    D->vi = c.take<unsigned int>(GC); D->ubuf = c.take<unsigned int>(6 * GC);
    D->cx = c.take<double>(GC); D->cy = c.take<double>(GC); D->cz = c.take<double>(GC);
    D->n_cand = c.take<int>(G); D->cpass = c.take<unsigned char>(GC); D->cstate32 = c.take<unsigned int>(GC);
    D->plist = c.take<int>(GC);
    D->assign = c.take<int>(GS);
    D->first = c.take<int>(GN); D->cnt = c.take<int>(GN); D->slot = c.take<int>(GN); D->cur = c.take<int>(GN);
    D->dict_node = c.take<int>(GN); D->n_dict = c.take<int>(G); D->list_off = c.take<int>(G * (S.capN + 1));
    D->list = c.take<int>(GS); D->sc_idx = c.take<int>(GS); D->sc_ang = c.take<double>(GS); D->sc_inter = c.take<double>(5 * GS);
    D->prop = c.take<Proposal>(GN); D->adec = c.take<ActDec>(GN); D->newl = c.take<int4>(GN);
    D->alist = c.take<int>(GN); D->cbits = c.take<unsigned int>((size_t)S.G * 4 * ((S.capN + 31) / 32));
    D->hitj = c.take<int>(GS); D->hl = c.take<int>(GS); D->ta = c.take<int>(GS); D->seq = c.take<int>(GS);
    D->veto = c.take<unsigned char>(GS);
    D->kd_idx = c.take<int>(GS); D->kd_posL = c.take<int>(GS); D->kd_posR = c.take<int>(GS); D->kd_rank = c.take<int>(GS);
    D->kd_nodes = c.take<int>(GS);
    D->kd_flag = c.take<int>(G); D->kill_T = c.take<int>(G); D->kill_H = c.take<int>(G); D->kd_list = c.take<int>(2 * G); D->kd_nflag = c.take<int>(2);
    D->seqhash = c.take<long long>(GS);
    D->set_hash = c.take<long long>(G * 2 * SET_TBL); D->set_key = c.take<int>(G * 2 * SET_TBL);
    D->err = c.take<int>(G); D->trace = c.take<int>(G * 4096 * 4); D->counters = c.take<long long>(G * 8); D->dbg = c.take<long long>(G * 8);
}

// exact radii + export of one forest (host).  Inputs are strided views into the pinned D2H staging buffers.
void finalize_forest(const OctaGrowConfig& c, int n, const double* px, const double* py, const double* pz,
                     const int* parent, const unsigned char* meta, double* out7, int64_t cap, int64_t* n_out) {
    const double r = c.r / c.param_scale;
    std::vector<double> rad(n, r), kap(n);
    std::vector<int> c0(n, -1), c1(n, -1);
    std::vector<unsigned char> nch(n, 0);
    int eff[8] = {0};
    effective_modes(c, eff);
    for (int i = 0; i < n; ++i) { const int m = meta[i] >> 1; kap[i] = (meta[i] != 0xff && m < c.n_modes) ? c.modes[eff[m]].kappa : 4.0; }
    // Final Murray radii with libm pow (exactly CPython's float.__pow__, arterial_tree.py:180).  The reference walks
    // to the root after every branch event (122 k walk steps per graph); the radius a node ENDS with is the value of
    // the last walk through it, i.e. f(children's final radii) -- the last branch event below a node updates it after
    // all its children are final, and a walk that stops early (recomputed value unchanged, :181-182) leaves ancestors
    // whose inputs did not change.  Nodes without any branch event below them keep the creation radius (elongation
    // never recomputes, greenhouse.py:258).  One pass in reverse creation order (children have larger ids) therefore
    // gives the same bits with ~10x fewer pow calls; tests compare against the oracle's literal event replay.
    for (int i = 0; i < n; ++i) {
        const int p = parent[i];
        if (p < 0) continue;
        if (nch[p] == 0) c0[p] = i; else c1[p] = i;
        ++nch[p];
    }
    // libm pow through a small direct-mapped cache keyed by both operands: the same (base, exponent) pair recurs all over a
    // forest (leaf radius, twigs of equal shape), and pow is most of the host time of a sample (~30 k calls, 2 ms).  A hit
    // returns what the call returned before: the bits cannot change.
    struct PowCache {
        enum { N = 2048 };
        double x[N], k[N], v[N];
        PowCache() { for (int i = 0; i < N; ++i) { x[i] = -1.0; k[i] = 0.0; v[i] = 0.0; } }
        double operator()(double a, double b) {
            uint64_t ua, ub;
            memcpy(&ua, &a, 8); memcpy(&ub, &b, 8);
            uint64_t h = (ua ^ (ub * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
            const int slot = (int)(h >> 53);
            if (x[slot] == a && k[slot] == b) return v[slot];
            const double r = pow(a, b);
            x[slot] = a; k[slot] = b; v[slot] = r;
            return r;
        }
    };
    static thread_local PowCache pw;
    std::vector<unsigned char> ev(n, 0);
    for (int i = n - 1; i >= 0; --i) {
        const int p = parent[i];
        if (meta[i] != 0xff && (meta[i] & 1) && p >= 0) ev[p] = 1;          // a walk started at the parent of this node
        if (ev[i] && p >= 0 && nch[i] > 0) {
            double s = 0 + pw(rad[c0[i]], kap[i]);
            if (nch[i] > 1) s = s + pw(rad[c1[i]], kap[i]);
            rad[i] = pw(s, 1 / kap[i]);
        }
        if (ev[i] && p >= 0) ev[p] = 1;
    }
    // per tree (roots in creation order), level order, children in attach order
    int64_t k = 0;
    std::vector<int> level, next;
    for (int root = 0; root < n; ++root) {
        if (parent[root] >= 0) continue;
        level.assign(1, root);
        while (!level.empty()) {
            next.clear();
            for (int id : level) {
                if (parent[id] >= 0) {
                    if (k < cap) {
                        double* o = out7 + 7 * k;
                        const int pa = parent[id];
                        o[0] = px[id]; o[1] = py[id]; o[2] = pz[id];
                        o[3] = px[pa]; o[4] = py[pa]; o[5] = pz[pa];
                        o[6] = rad[id];
                    }
                    ++k;
                }
                if (nch[id] > 0) next.push_back(c0[id]);
                if (nch[id] > 1) next.push_back(c1[id]);
            }
            level.swap(next);
        }
    }
    *n_out = k;
}

// constant-memory slots of the live contexts' pointer tables
std::mutex g_slot_mutex;
unsigned int g_slots_used = 0;
int acquire_slot() {
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    for (int i = 0; i < max_ctx_slots() && i < 32; ++i)
        if (!(g_slots_used & (1u << i))) { g_slots_used |= 1u << i; return i; }
    return -1;
}
void release_slot(int i) {
    if (i < 0) return;
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    g_slots_used &= ~(1u << i);
}

// persistent growth context: device state for up to G graphs of one configuration
struct GrowCtx {
    OctaGrowConfig cfg;
    GrowShape S;
    GrowDev D;
    char* dbase = nullptr;
    size_t dbytes = 0;
    std::vector<IterP> sched;
    int n_sm = 148;
    char* stage = nullptr;      // pinned host staging
    size_t stage_bytes = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    GrowEvents ev = {nullptr, nullptr, nullptr, nullptr};
    // The growth loop is a chain of short latency-bound launches: both of its streams get the highest priority, so that
    // throughput kernels of other streams (the voxelizer of the previous batch) do not sit in front of it.
    cudaStream_t main = nullptr, side = nullptr;
    int dslot = -1;             // constant-memory slot of this context's pointer table
    std::vector<unsigned char> geom;   // copy of the fixed sampling geometry (cfg.geometry points here)
    std::vector<unsigned short> geom_valid;   // np.argwhere(geom) as (i, j, k) triples (uploaded once to GrowDev::geom_valid)
    int last_n = 0;             // graphs of the last finished run (octa_grow_sinks reads their final sink lists)
    // Envelope of the node count of each forest after iteration i, over the batches this context has grown so far.  The next run
    // sizes k_commit's shared-memory tree mirror per launch from it: most of the schedule needs a fraction of the 224 KB, and a
    // CTA that holds less shared memory lets other graphs' kernels (other loops in flight) share its SM.  A tree that outgrows the
    // prediction runs the same code on the global arrays (bit-identical), so this is a performance hint, never a result.
    std::vector<int> hist_nodes[2];
    // The whole growth loop of a batch (launch_begin + every launch_iteration: ~15 kernels x 250 iterations on two streams with
    // their event edges) captured ONCE as a CUDA graph and re-launched for every later batch of the same size: one
    // cudaGraphLaunch instead of ~3 750 kernel launches per batch.  Iteration parameters travel by value inside the nodes (the
    // schedule is a property of the context), the per-batch state (seeds -> stumps, RNG states) is uploaded before the launch.
    // `cs` = k_commit's shared-memory size per (iteration, forest) baked into the nodes; the graph is re-captured when a tree
    // of an earlier batch outgrew one of them (they are a performance hint: an outgrown mirror runs on the global arrays).
    struct LoopGraph {
        cudaGraphExec_t exec = nullptr;
        int n_graphs = 0;
        bool from_history = false;
        std::vector<int> cs;
        uint64_t kernels = 0;
    } lg;
    ~GrowCtx() {
        if (lg.exec) cudaGraphExecDestroy(lg.exec);
        release_slot(dslot);
        if (ev_done) cudaEventDestroy(ev_done);
        for (cudaEvent_t e : thr_ev) if (e) cudaEventDestroy(e);
        if (main) cudaStreamDestroy(main);
        if (side) cudaStreamDestroy(side);
        for (cudaEvent_t e : {ev.start, ev.sinks, ev.kd, ev.killa}) if (e) cudaEventDestroy(e);
        if (dbase) cudaFree(dbase);
        if (stage) cudaFreeHost(stage);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
    // wait for `st` without spinning: several growth loops (and ranks) share the host cores
    static constexpr int THR_RING = 64;
    cudaEvent_t thr_ev[THR_RING] = {};       // launch throttle (grow_run_impl)
    cudaEvent_t ev_done = nullptr;
    cudaError_t wait(cudaStream_t st) {
        if (!ev_done) {
            cudaError_t e = cudaEventCreateWithFlags(&ev_done, cudaEventBlockingSync | cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaEventRecord(ev_done, st);
        return e != cudaSuccess ? e : cudaEventSynchronize(ev_done);
    }
    int ensure_stage(size_t bytes) {
        if (bytes <= stage_bytes) return OCTA_OK;
        if (stage) cudaFreeHost(stage);
        stage = nullptr; stage_bytes = 0;
        bytes += bytes / 4;                      // headroom: batch sizes drift by a few percent, pinned reallocation costs ~17 ms
        OCTA_CUDA_CHECK(cudaMallocHost(&stage, bytes));
        stage_bytes = bytes;
        return OCTA_OK;
    }
};

int check_config(const OctaGrowConfig* cfg) {
    OCTA_ARG_CHECK(cfg, "config is null");
    OCTA_ARG_CHECK(cfg->n_modes >= 1 && cfg->n_modes <= 8, "n_modes must be in [1, 8]");
    OCTA_ARG_CHECK(cfg->n_trees >= 1 && cfg->n_trees <= 64, "N_trees must be in [1, 64]");
    OCTA_ARG_CHECK(cfg->param_scale > 0, "param_scale must be positive");
    return OCTA_OK;
}

}  // namespace
}  // namespace octa

using namespace octa;

extern "C" int octa_grow_create(const OctaGrowConfig* cfg, int max_graphs, void** handle) {
    int rc = check_config(cfg);
    if (rc) return rc;
    OCTA_ARG_CHECK(handle && max_graphs > 0 && max_graphs <= 4096, "bad arguments");
    if (octa_device_count() <= 0) { set_error("octa_grow_create: no CUDA device (there is no CPU fallback)"); return OCTA_E_CUDA; }
    size_t geom_cells = 0;
    if (cfg->geometry) {      // any 3-D mask (simulation_space.py:29-34); voxel indices pass through uint16 in the reference (:108)
        const int* gd = cfg->geom_dims;
        OCTA_ARG_CHECK(gd[0] >= 1 && gd[1] >= 1 && gd[2] >= 1 && gd[0] <= 65535 && gd[1] <= 65535 && gd[2] <= 65535,
                       "geometry mask: every dimension must be between 1 and 65535");
        geom_cells = (size_t)gd[0] * gd[1] * gd[2];
        OCTA_ARG_CHECK(geom_cells <= ((size_t)1 << 26), "geometry mask: more than 2^26 voxels");
    }
    GrowCtx* ctx = new GrowCtx();
    ctx->dslot = acquire_slot();
    if (ctx->dslot < 0) { delete ctx; set_error("octa_grow_create: too many live growth contexts (max %d)", max_ctx_slots()); return OCTA_E_NOMEM; }
    ctx->cfg = *cfg;
    if (cfg->geometry) {      // own copy of the mask; the space becomes geom_dims / max(geom_dims) (simulation_space.py:31-33)
        const int* gd = cfg->geom_dims;
        const int gs = std::max(gd[0], std::max(gd[1], gd[2]));
        ctx->geom.assign(cfg->geometry, cfg->geometry + geom_cells);
        ctx->cfg.geometry = ctx->geom.data();
        for (int k = 0; k < 3; ++k) ctx->cfg.size[k] = (double)gd[k] / (double)gs;
        for (int i = 0; i < gd[0]; ++i)
            for (int j = 0; j < gd[1]; ++j)
                for (int k = 0; k < gd[2]; ++k)
                    if (ctx->geom[((size_t)i * gd[1] + j) * gd[2] + k]) {
                        ctx->geom_valid.push_back((unsigned short)i); ctx->geom_valid.push_back((unsigned short)j); ctx->geom_valid.push_back((unsigned short)k);
                    }
        if (ctx->geom_valid.empty()) { delete ctx; set_error("geometry mask: no valid sampling voxel"); return OCTA_E_ARG; }
    }
    build_schedule(ctx->cfg, &ctx->sched);
    if (ctx->sched.size() > 4096) { delete ctx; set_error("too many iterations (max 4096)"); return OCTA_E_ARG; }
    int Nmax = 1;
    long total_try = 0;
    for (const IterP& p : ctx->sched) { Nmax = std::max(Nmax, p.N); total_try += p.N; }
    GrowShape& S = ctx->S;
    S.G = max_graphs;
    S.Nmax = Nmax;
    S.geom_cells = (int)geom_cells;
    S.geom_nvalid = (int)(ctx->geom_valid.size() / 3);
    S.capN = cfg->cap_nodes > 0 ? cfg->cap_nodes : (int)std::min<long>(1 << 20, std::max<long>(4096, align_up((size_t)(total_try / 16 + 4096), 1024)));
    S.capS = cfg->cap_sinks > 0 ? cfg->cap_sinks : (int)std::min<long>(1 << 20, std::max<long>(4096, align_up((size_t)(total_try / 12 + 4096), 1024)));
    S.pycap = 2 * S.capN + 4 * 624;
    {
        const char* bo = getenv("OCTA_BALL_ORDER");      // "index" = list-index order (diagnostics); default exact
        S.exact_ball_order = (bo && strcmp(bo, "index") == 0) ? 0 : (bo && strcmp(bo, "always") == 0) ? 1 : 2;
    }
    S.kill_rcap = 4096;
    if (const char* e = getenv("OCTA_KILL_RCAP")) { const int v = atoi(e); if (v >= 0 && v <= 4096) S.kill_rcap = v; }   // tests: force the scan fallback
    S.commit_smem = 224 * 1024;                       // of the 227 KB a CTA may own on sm_100
    if (const char* e = getenv("OCTA_COMMIT_SMEM")) {  // tests: a small budget forces k_commit onto the global-memory tree view
        const int v = atoi(e);
        if (v >= 1024 && v <= 224 * 1024) S.commit_smem = v;
    }
    if (2 * cfg->n_trees > S.capN) { delete ctx; set_error("cap_nodes too small"); return OCTA_E_ARG; }
    Carver sizing(nullptr);
    carve(sizing, S, &ctx->D);
    cudaError_t ce = cudaMalloc(&ctx->dbase, sizing.off);
    if (ce != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", sizing.off, cudaGetErrorString(ce)); delete ctx; return OCTA_E_NOMEM; }
    ctx->dbytes = sizing.off;
    Carver real(ctx->dbase);
    carve(real, S, &ctx->D);
    if (cudaMemset(ctx->dbase, 0, ctx->dbytes) != cudaSuccess ||
        cudaMemset(ctx->D.first, 0x7f, sizeof(int) * (size_t)S.G * S.capN) != cudaSuccess) {
        set_error("cudaMemset failed"); delete ctx; return OCTA_E_CUDA;
    }
    cudaDeviceSynchronize();                          // (the context's own streams do not wait for the default stream)
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (!ctx->geom.empty() && (cudaMemcpy(ctx->D.geom_mask, ctx->geom.data(), ctx->geom.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
                               cudaMemcpy(ctx->D.geom_valid, ctx->geom_valid.data(), ctx->geom_valid.size() * sizeof(unsigned short), cudaMemcpyHostToDevice) != cudaSuccess)) {
        set_error("upload of the geometry mask failed"); delete ctx; return OCTA_E_CUDA;
    }
    if (prepare_kernels(S) != 0) { cudaGetLastError(); set_error("cudaFuncSetAttribute(k_commit) failed"); delete ctx; return OCTA_E_CUDA; }
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);      // hi = numerically lowest = greatest priority
        cudaStreamCreateWithPriority(&ctx->main, cudaStreamNonBlocking, hi);
        cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, hi);
    }
    for (cudaEvent_t* e : {&ctx->ev.start, &ctx->ev.sinks, &ctx->ev.kd, &ctx->ev.killa}) cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    cudaEventCreate(&ctx->e0);
    cudaEventCreate(&ctx->e1);
    *handle = ctx;
    return OCTA_OK;
}

extern "C" void octa_grow_destroy(void* handle) { delete (GrowCtx*)handle; }

static int grow_run_impl(void* handle, const uint64_t* seeds, int n_graphs, double* edges7_out, int64_t cap_edges,
                         int64_t* n_art_edges, int64_t* n_ven_edges, OctaGrowStats* stats, int32_t* trace,
                         double* device_ms, int64_t* packed_offsets) {
    // packed_offsets != NULL: rows are packed back to back (cap_edges = total capacity) and packed_offsets[0..n] receives
    // the row offsets; otherwise graph g owns the fixed slab edges7_out + g*cap_edges*7.
    GrowCtx* ctx = (GrowCtx*)handle;
    OCTA_ARG_CHECK(ctx && seeds && n_graphs > 0 && n_graphs <= ctx->S.G, "bad arguments (n_graphs must not exceed the context size)");
    OCTA_ARG_CHECK(edges7_out && cap_edges > 0 && n_art_edges && n_ven_edges, "output buffers missing");
    for (int i = 0; i < n_graphs; ++i) OCTA_ARG_CHECK(seeds[i] <= 0xffffffffull, "seeds must fit 32 bits (np.random.seed)");
    const OctaGrowConfig& cfg = ctx->cfg;
    GrowShape S = ctx->S;
    S.G = n_graphs;                                   // slabs are graph-major: a smaller batch uses the leading slabs
    GrowDev D = ctx->D;
    cudaStream_t st = ctx->main;
    // optional host-phase wall clock (OCTA_GROW_HOST_TIMING=1, diagnostics)
    static const bool host_timing = [] { const char* e = getenv("OCTA_GROW_HOST_TIMING"); return e && e[0] == '1'; }();
    auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tw[8] = {0};
    tw[0] = wall();
    // ---- host initialisation (Greenhouse.__init__ + Forest x2), staged and uploaded with strided copies
    std::vector<HostGraphInit> init(n_graphs);
    for (int g = 0; g < n_graphs; ++g) {
        int rc = init_graph(cfg, ctx->S.geom_nvalid, seeds[g], &init[g]);
        if (rc) return rc;
    }
    tw[1] = wall();
    const int n0 = 2 * cfg.n_trees;
    const double r0 = cfg.r / cfg.param_scale;
    const size_t G = n_graphs;
    {
        // staging layout: per array a dense [G][n0] block
        size_t off = 0;
        auto blk = [&](size_t elem) { size_t o = off; off = align_up(off + elem * G * n0, 256); return o; };
        size_t o_x[2], o_y[2], o_z[2], o_r[2], o_p[2], o_c0[2], o_c1[2], o_n[2], o_m[2];
        for (int f = 0; f < 2; ++f) {
            o_x[f] = blk(8); o_y[f] = blk(8); o_z[f] = blk(8); o_r[f] = blk(8);
            o_p[f] = blk(4); o_c0[f] = blk(4); o_c1[f] = blk(4); o_n[f] = blk(1); o_m[f] = blk(1);
        }
        const size_t o_mt_np = off; off = align_up(off + sizeof(MTState) * G, 256);
        const size_t o_mt_py = off; off = align_up(off + sizeof(MTState) * G, 256);
        const size_t o_faz = off; off = align_up(off + 8 * G, 256);
        const size_t o_nv = off; off = align_up(off + 4 * G, 256);
        const size_t o_cnt = off; off = align_up(off + 4 * G, 256);
        const size_t o_valid = off; off = align_up(off + (size_t)MAX_VALID * 2 * G, 256);
        int rc = ctx->ensure_stage(off);
        if (rc) return rc;
        char* sg = ctx->stage;
        for (int g = 0; g < n_graphs; ++g) {
            const HostGraphInit& h = init[g];
            for (int f = 0; f < 2; ++f) {
                double* hx = (double*)(sg + o_x[f]) + (size_t)g * n0; double* hy = (double*)(sg + o_y[f]) + (size_t)g * n0;
                double* hz = (double*)(sg + o_z[f]) + (size_t)g * n0; double* hr = (double*)(sg + o_r[f]) + (size_t)g * n0;
                int* hp = (int*)(sg + o_p[f]) + (size_t)g * n0; int* hc0 = (int*)(sg + o_c0[f]) + (size_t)g * n0;
                int* hc1 = (int*)(sg + o_c1[f]) + (size_t)g * n0;
                unsigned char* hn = (unsigned char*)(sg + o_n[f]) + (size_t)g * n0; unsigned char* hm = (unsigned char*)(sg + o_m[f]) + (size_t)g * n0;
                for (int i = 0; i < n0; ++i) { hc0[i] = -1; hc1[i] = -1; hn[i] = 0; }
                for (int i = 0; i < n0; ++i) {
                    hx[i] = h.pos[f][3 * i]; hy[i] = h.pos[f][3 * i + 1]; hz[i] = h.pos[f][3 * i + 2];
                    hr[i] = pow(r0, 4.0);   // ncon: every initial node hangs below a kappa-4 node
                    hp[i] = h.parent[f][i]; hm[i] = 0xff;
                    if (hp[i] >= 0) { hc0[hp[i]] = i; hn[hp[i]] = 1; }
                }
            }
            ((MTState*)(sg + o_mt_np))[g] = h.np_mt;
            ((MTState*)(sg + o_mt_py))[g] = h.py_mt;
            ((double*)(sg + o_faz))[g] = h.faz_radius;
            ((int*)(sg + o_nv))[g] = h.n_valid;
            ((int*)(sg + o_cnt))[g] = n0;
            memset(sg + o_valid + (size_t)g * MAX_VALID * 2, 0, (size_t)MAX_VALID * 2);
            memcpy(sg + o_valid + (size_t)g * MAX_VALID * 2, h.valid_ij.data(), h.valid_ij.size());
        }
        auto up2d = [&](void* dst, size_t dpitch, const void* src, size_t elem) {
            return cudaMemcpy2DAsync(dst, dpitch, src, elem * n0, elem * n0, G, cudaMemcpyHostToDevice, st);
        };
        for (int f = 0; f < 2; ++f) {
            const size_t cn = S.capN;
            OCTA_CUDA_CHECK(up2d(D.nx[f], 8 * cn, sg + o_x[f], 8)); OCTA_CUDA_CHECK(up2d(D.ny[f], 8 * cn, sg + o_y[f], 8));
            OCTA_CUDA_CHECK(up2d(D.nz[f], 8 * cn, sg + o_z[f], 8)); OCTA_CUDA_CHECK(up2d(D.ncon[f], 8 * cn, sg + o_r[f], 8));
            OCTA_CUDA_CHECK(up2d(D.npar[f], 4 * cn, sg + o_p[f], 4));
            OCTA_CUDA_CHECK(up2d(D.nch0[f], 4 * cn, sg + o_c0[f], 4)); OCTA_CUDA_CHECK(up2d(D.nch1[f], 4 * cn, sg + o_c1[f], 4));
            OCTA_CUDA_CHECK(up2d(D.nnch[f], cn, sg + o_n[f], 1)); OCTA_CUDA_CHECK(up2d(D.nmeta[f], cn, sg + o_m[f], 1));
            OCTA_CUDA_CHECK(cudaMemsetAsync(D.deact[f], 0, G * cn, st));
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.n_nodes[f], sg + o_cnt, 4 * G, cudaMemcpyHostToDevice, st));
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.n_act[f], sg + o_cnt, 4 * G, cudaMemcpyHostToDevice, st));
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.n_prev[f], sg + o_cnt, 4 * G, cudaMemcpyHostToDevice, st));
            OCTA_CUDA_CHECK(cudaMemsetAsync(D.n_s[f], 0, 4 * G, st));
        }
        OCTA_CUDA_CHECK(cudaMemcpyAsync(D.np_mt, sg + o_mt_np, sizeof(MTState) * G, cudaMemcpyHostToDevice, st));
        OCTA_CUDA_CHECK(cudaMemcpyAsync(D.py_mt, sg + o_mt_py, sizeof(MTState) * G, cudaMemcpyHostToDevice, st));
        OCTA_CUDA_CHECK(cudaMemcpyAsync(D.faz_radius, sg + o_faz, 8 * G, cudaMemcpyHostToDevice, st));
        OCTA_CUDA_CHECK(cudaMemcpyAsync(D.n_valid, sg + o_nv, 4 * G, cudaMemcpyHostToDevice, st));
        OCTA_CUDA_CHECK(cudaMemcpyAsync(D.valid_ij, sg + o_valid, (size_t)MAX_VALID * 2 * G, cudaMemcpyHostToDevice, st));
        OCTA_CUDA_CHECK(cudaMemsetAsync(D.py_n, 0, 4 * G, st)); OCTA_CUDA_CHECK(cudaMemsetAsync(D.py_pos, 0, 4 * G, st));
        OCTA_CUDA_CHECK(cudaMemsetAsync(D.py_draws, 0, 8 * G, st)); OCTA_CUDA_CHECK(cudaMemsetAsync(D.counters, 0, 64 * G, st)); OCTA_CUDA_CHECK(cudaMemsetAsync(D.dbg, 0, 64 * G, st));
        OCTA_CUDA_CHECK(cudaMemsetAsync(D.err, 0, 4 * G, st));
        OCTA_CUDA_CHECK(cudaMemsetAsync(D.kd_flag, 0, 4 * G, st)); OCTA_CUDA_CHECK(cudaMemsetAsync(D.kd_nflag, 0, 8, st));
        OCTA_CUDA_CHECK(cudaMemsetAsync(D.cbits, 0, sizeof(unsigned int) * G * 4 * ((S.capN + 31) / 32), st));
        OCTA_CUDA_CHECK(cudaMemsetAsync(D.trace, 0, sizeof(int) * G * 4096 * 4, st));
        OCTA_CUDA_CHECK(ctx->wait(st));      // the staging buffer is reused for the read-back
    }
    tw[2] = wall();
    OCTA_CUDA_CHECK(cudaEventRecord(ctx->e0, st));
    grow_timing_begin(st);
    if (upload_dev_table(ctx->dslot, D, st) != 0) { cudaGetLastError(); set_error("cudaMemcpyToSymbol(pointer table) failed"); return OCTA_E_CUDA; }
    const size_t n_it = ctx->sched.size();
    static const bool adapt = [] { const char* e = getenv("OCTA_COMMIT_ADAPT"); return !(e && e[0] == '0'); }();
    // k_commit's shared-memory size per (iteration, forest): the mirror of the tree it will see, from the envelope of the
    // context's previous batches (+12 % and 256 nodes of margin); `raw` = the same without margin
    auto mirror_bytes = [](size_t n) { return ((n * 17 + 3) & ~(size_t)3) + 16 * ((n + 31) >> 5) + 64; };   // k_commit's own formula
    bool have_hist = adapt && n_it > 0 && n_it <= 4096 && ctx->hist_nodes[0].size() == n_it && ctx->hist_nodes[1].size() == n_it;
    std::vector<int> cs_all(2 * n_it, S.commit_smem), cs_raw(2 * n_it, 0);
    for (size_t i = 0; i < n_it && have_hist; ++i)
        for (int f = 0; f < 2; ++f) {
            const size_t nb = (size_t)ctx->hist_nodes[f][i > 0 ? i - 1 : 0];                  // nodes before this call
            const size_t want = (mirror_bytes(nb * 9 / 8 + 256) + 1023) & ~(size_t)1023;
            cs_all[2 * i + f] = (int)std::min<size_t>((size_t)S.commit_smem, std::max<size_t>(want, 8 * 1024));
            cs_raw[2 * i + f] = (int)std::min<size_t>((size_t)S.commit_smem, mirror_bytes(nb));
        }
    // Launch throttle: the host issues an iteration in ~0.1 ms but the device needs ~1-2 ms for it, so an unthrottled thread runs
    // up to the depth of the launch queue ahead and then SPINS inside cudaLaunchKernel (measured: 70 us per launch call, 290 ms of a
    // core per loop, from every grower thread of every rank).  Every THROTTLE_STEP iterations an event is recorded; before going
    // on, the thread SLEEPS (blocking-sync event) until the device is within THROTTLE_AHEAD iterations.  OCTA_GROW_THROTTLE=0: off.
    static const int throttle = [] { const char* e = getenv("OCTA_GROW_THROTTLE"); return e ? atoi(e) : 24; }();
    constexpr int THROTTLE_STEP = 4;
    auto issue_loop = [&](const std::vector<int>& cs, bool capturing) {
        launch_begin(ctx->dslot, S, ctx->sched[0], ctx->n_sm, st, ctx->side, ctx->ev);
        const bool thr = throttle > 0 && !capturing;
        const int ring = GrowCtx::THR_RING;
        for (size_t i = 0; i < n_it; ++i) {
            launch_iteration(ctx->dslot, S, &cs[2 * i], ctx->sched[i], i + 1 < n_it ? &ctx->sched[i + 1] : nullptr, ctx->n_sm, st, ctx->side, ctx->ev);
            if (thr && (i % THROTTLE_STEP) == THROTTLE_STEP - 1) {
                const size_t k = i / THROTTLE_STEP;
                if (!ctx->thr_ev[k % ring]) cudaEventCreateWithFlags(&ctx->thr_ev[k % ring], cudaEventBlockingSync | cudaEventDisableTiming);
                cudaEventRecord(ctx->thr_ev[k % ring], st);
                const size_t behind = (size_t)(throttle + THROTTLE_STEP - 1) / THROTTLE_STEP;
                if (k >= behind) cudaEventSynchronize(ctx->thr_ev[(k - behind) % ring]);
            }
        }
    };
    // OCTA_GROW_GRAPH: 0 (default) = stream launches, 1 = CUDA graph from the context's second batch on (the first one also
    // establishes the mirror sizes), 2 = CUDA graph always.  Measured on one B200 (profiles/README.md, round 2): the graph frees
    // the host (cudaGraphLaunch of the 4 001 kernels: 0.03 ms instead of ~190 ms of launch calls per loop) but the device
    // executes the same loop ~10 % SLOWER from the graph (315 vs 285 ms for 32 graphs alone), and 8 loops in flight reach
    // 426 graphs/s instead of 498 -- so it is opt-in, for hosts with few cores per GPU.
    static const int graph_mode = [] { const char* e = getenv("OCTA_GROW_GRAPH"); return e ? atoi(e) : 0; }();
    const bool use_graph = n_it > 0 && !grow_timing_enabled() && (graph_mode >= 2 || (graph_mode == 1 && have_hist));
    if (use_graph) {
        GrowCtx::LoopGraph& lg = ctx->lg;
        bool stale = !lg.exec || lg.n_graphs != n_graphs || lg.cs.size() != cs_all.size() || (have_hist && !lg.from_history);
        for (size_t q = 0; q < cs_all.size() && !stale; ++q) stale = cs_raw[q] > lg.cs[q];    // a tree outgrew its captured mirror
        if (stale) {
            if (lg.exec) { cudaGraphExecDestroy(lg.exec); lg.exec = nullptr; }
            OCTA_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            t_capturing = true; t_captured = 0;          // captured, not launched: counted per graph launch below
            issue_loop(cs_all, true);
            t_capturing = false;
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (ce != cudaSuccess || !graph) { cudaGetLastError(); set_error("capture of the growth loop failed: %s", cudaGetErrorString(ce)); return OCTA_E_CUDA; }
            ce = cudaGraphInstantiate(&lg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { lg.exec = nullptr; cudaGetLastError(); set_error("cudaGraphInstantiate(growth loop) failed: %s", cudaGetErrorString(ce)); return OCTA_E_CUDA; }
            lg.n_graphs = n_graphs; lg.from_history = have_hist; lg.cs = cs_all; lg.kernels = t_captured;
        }
        const double tg0 = wall();
        OCTA_CUDA_CHECK(cudaGraphLaunch(lg.exec, st));
        count_launch((int)lg.kernels);
        if (host_timing) fprintf(stderr, "[octa grow host] %s cudaGraphLaunch of %llu kernels: %.2f ms\n", stale ? "capture + instantiate done;" : "", (unsigned long long)lg.kernels, wall() - tg0);
    } else if (n_it > 0) {
        issue_loop(cs_all, false);
    }
    OCTA_CUDA_CHECK(cudaEventRecord(ctx->e1, st));
    tw[3] = wall();
    // ---- read back: counts first, then strided copies of the live prefix of every node array
    std::vector<int> err(n_graphs), nn[2], ns[2];
    std::vector<long long> draws(n_graphs), counters((size_t)n_graphs * 8), dbg((size_t)n_graphs * 8);
    for (int f = 0; f < 2; ++f) {
        nn[f].resize(n_graphs); ns[f].resize(n_graphs);
        OCTA_CUDA_CHECK(cudaMemcpyAsync(nn[f].data(), D.n_nodes[f], 4 * G, cudaMemcpyDeviceToHost, st));
        OCTA_CUDA_CHECK(cudaMemcpyAsync(ns[f].data(), D.n_s[f], 4 * G, cudaMemcpyDeviceToHost, st));
    }
    OCTA_CUDA_CHECK(cudaMemcpyAsync(err.data(), D.err, 4 * G, cudaMemcpyDeviceToHost, st));
    OCTA_CUDA_CHECK(cudaMemcpyAsync(draws.data(), D.py_draws, 8 * G, cudaMemcpyDeviceToHost, st));
    OCTA_CUDA_CHECK(cudaMemcpyAsync(counters.data(), D.counters, 64 * G, cudaMemcpyDeviceToHost, st));
    OCTA_CUDA_CHECK(cudaMemcpyAsync(dbg.data(), D.dbg, 64 * G, cudaMemcpyDeviceToHost, st));
    OCTA_CUDA_CHECK(ctx->wait(st));
    OCTA_CUDA_CHECK(cudaGetLastError());
    tw[4] = wall();
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->e0, ctx->e1);
    if (device_ms) *device_ms = ms;
    grow_timing_report();
    if (trace) OCTA_CUDA_CHECK(cudaMemcpy(trace, D.trace, sizeof(int) * G * 4096 * 4, cudaMemcpyDeviceToHost));
    if (n_it > 0 && n_it <= 4096) {      // node counts per iteration of this batch -> envelope for the next run's mirror sizes
        std::vector<int> tr((size_t)G * n_it * 4);
        OCTA_CUDA_CHECK(cudaMemcpy2D(tr.data(), n_it * 16, D.trace, (size_t)4096 * 16, n_it * 16, G, cudaMemcpyDeviceToHost));
        for (int f = 0; f < 2; ++f) {
            std::vector<int>& h = ctx->hist_nodes[f];
            if (h.size() != n_it) h.assign(n_it, 0);
            for (int g = 0; g < n_graphs; ++g)
                for (size_t i = 0; i < n_it; ++i) h[i] = std::max(h[i], tr[((size_t)g * n_it + i) * 4 + 2 * f]);
            for (size_t i = 1; i < n_it; ++i) h[i] = std::max(h[i], h[i - 1]);
        }
    }
    size_t mx[2] = {1, 1};
    for (int f = 0; f < 2; ++f) for (int g = 0; g < n_graphs; ++g) mx[f] = std::max<size_t>(mx[f], (size_t)nn[f][g]);
    size_t so[2][5], off = 0;
    for (int f = 0; f < 2; ++f) {
        const size_t elem[5] = {8, 8, 8, 4, 1};
        for (int a = 0; a < 5; ++a) { so[f][a] = off; off = align_up(off + elem[a] * mx[f] * G, 256); }
    }
    {
        int rc = ctx->ensure_stage(off);
        if (rc) return rc;
    }
    char* sg = ctx->stage;
    for (int f = 0; f < 2; ++f) {
        const size_t cn = S.capN, w = mx[f];
        auto dn2d = [&](size_t o, const void* src, size_t elem) {
            return cudaMemcpy2DAsync(sg + o, elem * w, src, elem * cn, elem * w, G, cudaMemcpyDeviceToHost, st);
        };
        OCTA_CUDA_CHECK(dn2d(so[f][0], D.nx[f], 8)); OCTA_CUDA_CHECK(dn2d(so[f][1], D.ny[f], 8)); OCTA_CUDA_CHECK(dn2d(so[f][2], D.nz[f], 8));
        OCTA_CUDA_CHECK(dn2d(so[f][3], D.npar[f], 4)); OCTA_CUDA_CHECK(dn2d(so[f][4], D.nmeta[f], 1));
    }
    OCTA_CUDA_CHECK(ctx->wait(st));
    std::vector<int64_t> row0(n_graphs + 1, 0);
    if (packed_offsets) {
        // every non-root node is one row; roots = N_trees per forest
        for (int g = 0; g < n_graphs; ++g)
            row0[g + 1] = row0[g] + std::max(0, nn[0][g] - cfg.n_trees) + std::max(0, nn[1][g] - cfg.n_trees);
        for (int g = 0; g <= n_graphs; ++g) packed_offsets[g] = row0[g];
        if (row0[n_graphs] > cap_edges) {
            set_error("octa_grow_run_packed: edge buffer too small (%lld rows needed, %lld available)",
                      (long long)row0[n_graphs], (long long)cap_edges);
            return OCTA_E_NOMEM;
        }
    }
    tw[5] = wall();
    // ---- exact radii + edge rows, multi-threaded over graphs
    // (the cores this process may run on: ranks of a multi-GPU job are pinned to their own slice of the box)
    unsigned ncores = std::thread::hardware_concurrency();
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) ncores = (unsigned)CPU_COUNT(&set);
    }
    unsigned nthreads = std::max(1u, std::min<unsigned>(ncores, (unsigned)n_graphs));
    std::vector<std::thread> pool;
    for (unsigned wk = 0; wk < nthreads; ++wk)
        pool.emplace_back([&, wk]() {
            for (int g = (int)wk; g < n_graphs; g += (int)nthreads) {
                double* out = packed_offsets ? edges7_out + (size_t)row0[g] * 7 : edges7_out + (size_t)g * cap_edges * 7;
                const int64_t cap_g = packed_offsets ? row0[g + 1] - row0[g] : cap_edges;
                int64_t cnt[2] = {0, 0};
                int64_t used = 0;
                for (int f = 0; f < 2; ++f) {
                    const size_t w = mx[f];
                    finalize_forest(cfg, nn[f][g], (const double*)(sg + so[f][0]) + g * w, (const double*)(sg + so[f][1]) + g * w,
                                    (const double*)(sg + so[f][2]) + g * w, (const int*)(sg + so[f][3]) + g * w,
                                    (const unsigned char*)(sg + so[f][4]) + g * w, out + 7 * used, cap_g - used, &cnt[f]);
                    used = std::min<int64_t>(cap_g, used + cnt[f]);
                }
                n_art_edges[g] = cnt[0];
                n_ven_edges[g] = cnt[1];
            }
        });
    for (auto& t : pool) t.join();
    tw[6] = wall();
    if (host_timing)
        fprintf(stderr, "[octa grow host] init %.1f  stage+upload %.1f  launch issue %.1f  wait device %.1f  download %.1f  radii+rows (%u threads) %.1f  | total %.1f ms\n",
                tw[1] - tw[0], tw[2] - tw[1], tw[3] - tw[2], tw[4] - tw[3], tw[5] - tw[4], nthreads, tw[6] - tw[5], tw[6] - tw[0]);
    int worst = 0;
    for (int g = 0; g < n_graphs; ++g) {
        if (stats) {
            OctaGrowStats& s = stats[g];
            s.n_art_nodes = nn[0][g]; s.n_ven_nodes = nn[1][g]; s.n_oxy_left = ns[0][g]; s.n_co2_left = ns[1][g];
            s.py_draws = draws[g];
            s.sum_A = counters[(size_t)g * 8 + 0]; s.sum_M = counters[(size_t)g * 8 + 1];
            s.sum_P = counters[(size_t)g * 8 + 2]; s.sum_S = counters[(size_t)g * 8 + 3];
            for (int q = 0; q < 4; ++q) s.commit_cycles[q] = counters[(size_t)g * 8 + 4 + q];
            for (int q = 0; q < 8; ++q) s.replay_detail[q] = dbg[(size_t)g * 8 + q];
            s.err = err[g]; s.n_iters = (int)ctx->sched.size();
        }
        if (err[g] && !worst) worst = err[g];
        if (!packed_offsets && n_art_edges[g] + n_ven_edges[g] > cap_edges && !worst) worst = 100;
        if (packed_offsets && n_art_edges[g] + n_ven_edges[g] != row0[g + 1] - row0[g] && !worst) worst = 101;
    }
    ctx->last_n = worst ? 0 : n_graphs;
    if (worst) {
        set_error("octa_grow_run: simulation error code %d (1 node capacity, 2 sink capacity, 3 rng buffer, "
                  "5 set table, 6 sample outside the geometry mask array, 12/13 eigen solver, 100 edge buffer too small)", worst);
        return OCTA_E_STATE;
    }
    return OCTA_OK;
}

extern "C" int octa_grow_run(void* handle, const uint64_t* seeds, int n_graphs, double* edges7_out, int64_t cap_edges,
                             int64_t* n_art_edges, int64_t* n_ven_edges, OctaGrowStats* stats, int32_t* trace,
                             double* device_ms) {
    return grow_run_impl(handle, seeds, n_graphs, edges7_out, cap_edges, n_art_edges, n_ven_edges, stats, trace, device_ms, nullptr);
}

extern "C" int octa_grow_run_packed(void* handle, const uint64_t* seeds, int n_graphs, double* edges7_out,
                                    int64_t cap_total_edges, int64_t* edge_offsets, int64_t* n_art_edges,
                                    int64_t* n_ven_edges, OctaGrowStats* stats, int32_t* trace, double* device_ms) {
    OCTA_ARG_CHECK(edge_offsets, "edge_offsets is null");
    return grow_run_impl(handle, seeds, n_graphs, edges7_out, cap_total_edges, n_art_edges, n_ven_edges, stats, trace, device_ms,
                         edge_offsets);
}

extern "C" int octa_grow_sinks(void* handle, int graph, int which, double* xyz_out, int64_t cap, int64_t* n) {
    GrowCtx* ctx = (GrowCtx*)handle;
    OCTA_ARG_CHECK(ctx && n && (which == 0 || which == 1), "bad arguments");
    OCTA_ARG_CHECK(graph >= 0 && graph < ctx->last_n, "graph index outside the context's last run");
    const GrowDev& D = ctx->D;
    int cnt = 0;
    OCTA_CUDA_CHECK(cudaMemcpy(&cnt, D.n_s[which] + graph, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    if (!xyz_out || cap < cnt) { set_error("octa_grow_sinks: buffer holds %lld rows, %d needed", (long long)(xyz_out ? cap : 0), cnt); return OCTA_E_NOMEM; }
    if (cnt == 0) return OCTA_OK;
    std::vector<double> soa((size_t)3 * cnt);
    const size_t o = (size_t)graph * ctx->S.capS;
    OCTA_CUDA_CHECK(cudaMemcpy(soa.data(), D.sx[which] + o, 8 * (size_t)cnt, cudaMemcpyDeviceToHost));
    OCTA_CUDA_CHECK(cudaMemcpy(soa.data() + cnt, D.sy[which] + o, 8 * (size_t)cnt, cudaMemcpyDeviceToHost));
    OCTA_CUDA_CHECK(cudaMemcpy(soa.data() + 2 * (size_t)cnt, D.sz[which] + o, 8 * (size_t)cnt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; ++i) { xyz_out[3 * i] = soa[i]; xyz_out[3 * i + 1] = soa[cnt + i]; xyz_out[3 * i + 2] = soa[2 * (size_t)cnt + i]; }
    return OCTA_OK;
}

extern "C" int octa_grow_batch_host(const OctaGrowConfig* cfg, const uint64_t* seeds, int n_graphs, double* edges7_out,
                                    int64_t cap_edges, int64_t* n_art_edges, int64_t* n_ven_edges,
                                    OctaGrowStats* stats, int32_t* trace, double* device_ms) {
    void* h = nullptr;
    int rc = octa_grow_create(cfg, n_graphs, &h);
    if (rc) return rc;
    rc = octa_grow_run(h, seeds, n_graphs, edges7_out, cap_edges, n_art_edges, n_ven_edges, stats, trace, device_ms);
    octa_grow_destroy(h);
    return rc;
}
