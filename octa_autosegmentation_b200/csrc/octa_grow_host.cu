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
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "octa_common.h"
#include "octa_grow.cuh"

namespace octa {

void launch_iteration(const GrowDev& D, const GrowShape& S, const IterP& P, int n_sm, cudaStream_t st);

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
int init_graph(const OctaGrowConfig& c, uint64_t seed, HostGraphInit* h) {
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
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            const double a = (double)j - fc[0], b = (double)i - fc[1];
            bool ok = a * a + b * b > fr * fr;
            if (nerve) { const double e = (double)j - ncv[0], f = (double)i - ncv[1]; ok = ok && (e * e + f * f > nrv * nrv); }
            if (ok) { h->valid_ij.push_back((unsigned char)i); h->valid_ij.push_back((unsigned char)j); }
        }
    h->n_valid = (int)h->valid_ij.size() / 2;
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
                if (wall == 0 || wall == 1) {
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
                    // the reference's z0/z1 branch dereferences an attribute that does not exist (simulation_space.py:83)
                    set_error("source walls z0/z1 are not usable in the reference either (AttributeError)");
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

// greenhouse.py:34-51 / :83-90 / :139-147 -> one IterP per iteration
void build_schedule(const OctaGrowConfig& c, std::vector<IterP>* out) {
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
        const OctaGrowMode& m = c.modes[mi];
        if (m.reinit) init_params(m);
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
            P.N = N; P.t = t; P.first_mode = m.first_mode; P.mode_idx = mi; P.iter = iter++;
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
        D->nrad[f] = c.take<double>(GN); D->nkap[f] = c.take<double>(GN);
        D->npar[f] = c.take<int>(GN); D->nch0[f] = c.take<int>(GN); D->nch1[f] = c.take<int>(GN);
        D->nnch[f] = c.take<unsigned char>(GN); D->nmeta[f] = c.take<unsigned char>(GN); D->deact[f] = c.take<unsigned char>(GN); D->dirty[f] = c.take<unsigned char>(GN);
        D->n_nodes[f] = c.take<int>(G); D->n_prev[f] = c.take<int>(G);
        D->act[f] = c.take<int>(GN); D->n_act[f] = c.take<int>(G);
        D->ax[f] = c.take<double>(GN); D->ay[f] = c.take<double>(GN); D->az[f] = c.take<double>(GN);
        D->sx[f] = c.take<double>(GS); D->sy[f] = c.take<double>(GS); D->sz[f] = c.take<double>(GS);
        D->n_s[f] = c.take<int>(G);
    }
    const size_t GG = (size_t)S.G * (S.capN > S.capS ? S.capN : S.capS);
    for (int w = 0; w < 4; ++w) {
        D->gx[w] = c.take<double>(GG); D->gy[w] = c.take<double>(GG); D->gz[w] = c.take<double>(GG); D->gr[w] = c.take<double>(GG);
        D->gi[w] = c.take<int>(GG); D->gcell[w] = c.take<int>(G * (GRID * GRID + 1));
    }
    D->np_mt = c.take<MTState>(G); D->py_mt = c.take<MTState>(G);
    D->py_buf = c.take<unsigned int>(G * S.pycap); D->py_n = c.take<int>(G); D->py_pos = c.take<int>(G);
    D->py_draws = c.take<long long>(G);
    D->faz_radius = c.take<double>(G); D->n_valid = c.take<int>(G); D->valid_ij = c.take<unsigned char>(G * MAX_VALID * 2);
    D->vi = c.take<unsigned int>(GC); D->ubuf = c.take<unsigned int>(6 * GC);
    D->cx = c.take<double>(GC); D->cy = c.take<double>(GC); D->cz = c.take<double>(GC);
    D->n_cand = c.take<int>(G); D->cpass = c.take<unsigned char>(GC); D->cstate = c.take<unsigned char>(GC);
    D->plist = c.take<int>(GC);
    D->assign = c.take<int>(GS);
    D->first = c.take<int>(GN); D->cnt = c.take<int>(GN); D->slot = c.take<int>(GN); D->slot_call = c.take<int>(GN); D->cur = c.take<int>(GN); D->rtag = c.take<int>(GN);
    D->dict_node = c.take<int>(GN); D->n_dict = c.take<int>(G); D->list_off = c.take<int>(G * (S.capN + 1));
    D->list = c.take<int>(GS); D->sc_idx = c.take<int>(GS); D->sc_ang = c.take<double>(GS);
    D->prop = c.take<Proposal>(GN); D->alist = c.take<int>(GN); D->n_alist = c.take<int>(G);
    D->hitj = c.take<int>(GS); D->hl = c.take<int>(GS); D->ta = c.take<int>(GS); D->seq = c.take<int>(GS);
    D->veto = c.take<unsigned char>(GS);
    D->set_hash = c.take<long long>(G * 2 * SET_TBL); D->set_key = c.take<int>(G * 2 * SET_TBL);
    D->err = c.take<int>(G); D->trace = c.take<int>(G * 4096 * 4); D->counters = c.take<long long>(G * 8);
}

// exact radii + export of one graph (host)
struct GraphExport {
    std::vector<double> pos[2];
    std::vector<int> parent[2];
    std::vector<unsigned char> meta[2];
};

void finalize_forest(const OctaGrowConfig& c, const std::vector<double>& pos, const std::vector<int>& parent,
                     const std::vector<unsigned char>& meta, double* out7, int64_t cap, int64_t* n_out) {
    const int n = (int)parent.size();
    const double r = c.r / c.param_scale;
    std::vector<double> rad(n, r), kap(n);
    std::vector<int> c0(n, -1), c1(n, -1);
    std::vector<unsigned char> nch(n, 0);
    for (int i = 0; i < n; ++i) { const int m = meta[i] >> 1; kap[i] = (m >= 0 && m < c.n_modes && meta[i] != 0xff) ? c.modes[m].kappa : 4.0; }
    // replay in creation order (arterial_tree.py:174-184 with libm pow, exactly CPython's float.__pow__)
    for (int i = 0; i < n; ++i) {
        const int p = parent[i];
        if (p < 0) continue;
        if (nch[p] == 0) c0[p] = i; else c1[p] = i;
        ++nch[p];
        if (meta[i] != 0xff && (meta[i] & 1)) {
            int q = p;
            while (true) {
                if (parent[q] < 0 || nch[q] == 0) break;
                double s = 0 + pow(rad[c0[q]], kap[q]);
                if (nch[q] > 1) s = s + pow(rad[c1[q]], kap[q]);
                const double rp = pow(s, 1 / kap[q]);
                if (rad[q] == rp) break;
                rad[q] = rp;
                q = parent[q];
            }
        }
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
                        o[0] = pos[3 * id]; o[1] = pos[3 * id + 1]; o[2] = pos[3 * id + 2];
                        o[3] = pos[3 * pa]; o[4] = pos[3 * pa + 1]; o[5] = pos[3 * pa + 2];
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

}  // namespace
}  // namespace octa

using namespace octa;

extern "C" int octa_grow_batch_host(const OctaGrowConfig* cfg, const uint64_t* seeds, int n_graphs, double* edges7_out,
                                    int64_t cap_edges, int64_t* n_art_edges, int64_t* n_ven_edges,
                                    OctaGrowStats* stats, int32_t* trace, double* device_ms) {
    OCTA_ARG_CHECK(cfg && seeds && n_graphs > 0 && n_graphs <= 4096, "bad arguments");
    OCTA_ARG_CHECK(cfg->n_modes >= 1 && cfg->n_modes <= 8, "n_modes must be in [1, 8]");
    OCTA_ARG_CHECK(cfg->n_trees >= 1 && cfg->n_trees <= 64, "N_trees must be in [1, 64]");
    OCTA_ARG_CHECK(cfg->param_scale > 0, "param_scale must be positive");
    OCTA_ARG_CHECK(edges7_out && cap_edges > 0 && n_art_edges && n_ven_edges, "output buffers missing");
    for (int i = 0; i < n_graphs; ++i) OCTA_ARG_CHECK(seeds[i] <= 0xffffffffull, "seeds must fit 32 bits (np.random.seed)");
    if (octa_device_count() <= 0) { set_error("octa_grow_batch_host: no CUDA device (there is no CPU fallback)"); return OCTA_E_CUDA; }
    std::vector<IterP> sched;
    build_schedule(*cfg, &sched);
    OCTA_ARG_CHECK(sched.size() <= 4096, "too many iterations (max 4096)");
    int Nmax = 1;
    for (const IterP& p : sched) Nmax = std::max(Nmax, p.N);
    GrowShape S;
    S.G = n_graphs;
    S.Nmax = Nmax;
    long total_try = 0;
    for (const IterP& p : sched) total_try += p.N;
    S.capN = cfg->cap_nodes > 0 ? cfg->cap_nodes : (int)std::min<long>(1 << 20, std::max<long>(4096, align_up((size_t)(total_try / 16 + 4096), 1024)));
    S.capS = cfg->cap_sinks > 0 ? cfg->cap_sinks : (int)std::min<long>(1 << 20, std::max<long>(4096, align_up((size_t)(total_try / 12 + 4096), 1024)));
    S.pycap = 2 * S.capN + 4 * 624;
    // host initialisation
    std::vector<HostGraphInit> init(n_graphs);
    for (int g = 0; g < n_graphs; ++g) {
        int rc = init_graph(*cfg, seeds[g], &init[g]);
        if (rc) return rc;
        if ((int)init[g].parent[0].size() > S.capN) { set_error("cap_nodes too small"); return OCTA_E_ARG; }
    }
    GrowDev D;
    Carver sizing(nullptr);
    carve(sizing, S, &D);
    char* dbase = nullptr;
    cudaError_t ce = cudaMalloc(&dbase, sizing.off);
    if (ce != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", sizing.off, cudaGetErrorString(ce)); return OCTA_E_NOMEM; }
    struct Guard { char* p; ~Guard() { cudaFree(p); } } guard{dbase};
    Carver real(dbase);
    carve(real, S, &D);
    cudaStream_t st = nullptr;
    OCTA_CUDA_CHECK(cudaMemsetAsync(dbase, 0, sizing.off, st));
    OCTA_CUDA_CHECK(cudaMemsetAsync(D.first, 0x7f, sizeof(int) * (size_t)S.G * S.capN, st));
    const double r0 = cfg->r / cfg->param_scale;
    {
        std::vector<double> hx, hy, hz, hr, hk;
        std::vector<int> hp, hc0, hact;
        std::vector<unsigned char> hn, hm;
        for (int g = 0; g < n_graphs; ++g) {
            const HostGraphInit& h = init[g];
            const size_t nb = (size_t)g * S.capN;
            for (int f = 0; f < 2; ++f) {
                const int n = (int)h.parent[f].size();
                hx.assign(n, 0); hy.assign(n, 0); hz.assign(n, 0); hr.assign(n, r0); hk.assign(n, 4.0);
                hp.assign(n, -1); hc0.assign(n, -1); hact.assign(n, 0); hn.assign(n, 0); hm.assign(n, 0xff);
                for (int i = 0; i < n; ++i) {
                    hx[i] = h.pos[f][3 * i]; hy[i] = h.pos[f][3 * i + 1]; hz[i] = h.pos[f][3 * i + 2];
                    hp[i] = h.parent[f][i];
                    if (hp[i] >= 0) { hc0[hp[i]] = i; hn[hp[i]] = 1; }
                    hact[i] = i;
                }
                auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st); };
                OCTA_CUDA_CHECK(up(D.nx[f] + nb, hx.data(), 8 * n)); OCTA_CUDA_CHECK(up(D.ny[f] + nb, hy.data(), 8 * n));
                OCTA_CUDA_CHECK(up(D.nz[f] + nb, hz.data(), 8 * n)); OCTA_CUDA_CHECK(up(D.nrad[f] + nb, hr.data(), 8 * n));
                OCTA_CUDA_CHECK(up(D.nkap[f] + nb, hk.data(), 8 * n)); OCTA_CUDA_CHECK(up(D.npar[f] + nb, hp.data(), 4 * n));
                OCTA_CUDA_CHECK(up(D.nch0[f] + nb, hc0.data(), 4 * n)); OCTA_CUDA_CHECK(up(D.nnch[f] + nb, hn.data(), n));
                OCTA_CUDA_CHECK(up(D.nmeta[f] + nb, hm.data(), n));
                OCTA_CUDA_CHECK(up(D.act[f] + nb, hact.data(), 4 * n));
                OCTA_CUDA_CHECK(up(D.ax[f] + nb, hx.data(), 8 * n)); OCTA_CUDA_CHECK(up(D.ay[f] + nb, hy.data(), 8 * n));
                OCTA_CUDA_CHECK(up(D.az[f] + nb, hz.data(), 8 * n));
                OCTA_CUDA_CHECK(up(D.n_nodes[f] + g, &n, 4)); OCTA_CUDA_CHECK(up(D.n_act[f] + g, &n, 4));
                OCTA_CUDA_CHECK(up(D.n_prev[f] + g, &n, 4));
                OCTA_CUDA_CHECK(cudaStreamSynchronize(st));   // staging vectors are reused
            }
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.np_mt + g, &h.np_mt, sizeof(MTState), cudaMemcpyHostToDevice, st));
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.py_mt + g, &h.py_mt, sizeof(MTState), cudaMemcpyHostToDevice, st));
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.faz_radius + g, &h.faz_radius, 8, cudaMemcpyHostToDevice, st));
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.n_valid + g, &h.n_valid, 4, cudaMemcpyHostToDevice, st));
            OCTA_CUDA_CHECK(cudaMemcpyAsync(D.valid_ij + (size_t)g * MAX_VALID * 2, h.valid_ij.data(), h.valid_ij.size(),
                                            cudaMemcpyHostToDevice, st));
        }
        OCTA_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (!trace) D.trace = nullptr;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaEvent_t e0, e1;
    OCTA_CUDA_CHECK(cudaEventCreate(&e0));
    OCTA_CUDA_CHECK(cudaEventCreate(&e1));
    OCTA_CUDA_CHECK(cudaEventRecord(e0, st));
    for (const IterP& P : sched) launch_iteration(D, S, P, n_sm, st);
    OCTA_CUDA_CHECK(cudaEventRecord(e1, st));
    OCTA_CUDA_CHECK(cudaStreamSynchronize(st));
    OCTA_CUDA_CHECK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (device_ms) *device_ms = ms;
    // read back
    std::vector<int> err(n_graphs), nn[2], ns[2];
    std::vector<long long> draws(n_graphs), counters((size_t)n_graphs * 8);
    for (int f = 0; f < 2; ++f) {
        nn[f].resize(n_graphs); ns[f].resize(n_graphs);
        OCTA_CUDA_CHECK(cudaMemcpy(nn[f].data(), D.n_nodes[f], 4 * n_graphs, cudaMemcpyDeviceToHost));
        OCTA_CUDA_CHECK(cudaMemcpy(ns[f].data(), D.n_s[f], 4 * n_graphs, cudaMemcpyDeviceToHost));
    }
    OCTA_CUDA_CHECK(cudaMemcpy(err.data(), D.err, 4 * n_graphs, cudaMemcpyDeviceToHost));
    OCTA_CUDA_CHECK(cudaMemcpy(draws.data(), D.py_draws, 8 * n_graphs, cudaMemcpyDeviceToHost));
    OCTA_CUDA_CHECK(cudaMemcpy(counters.data(), D.counters, 8 * 8 * n_graphs, cudaMemcpyDeviceToHost));
    if (trace) OCTA_CUDA_CHECK(cudaMemcpy(trace, D.trace, sizeof(int) * (size_t)n_graphs * 4096 * 4, cudaMemcpyDeviceToHost));
    std::vector<GraphExport> ex(n_graphs);
    std::vector<double> tmp;
    for (int g = 0; g < n_graphs; ++g) {
        const size_t nb = (size_t)g * S.capN;
        for (int f = 0; f < 2; ++f) {
            const int n = nn[f][g];
            ex[g].pos[f].resize(3 * (size_t)n); ex[g].parent[f].resize(n); ex[g].meta[f].resize(n);
            tmp.resize(3 * (size_t)n);
            OCTA_CUDA_CHECK(cudaMemcpy(tmp.data(), D.nx[f] + nb, 8 * n, cudaMemcpyDeviceToHost));
            OCTA_CUDA_CHECK(cudaMemcpy(tmp.data() + n, D.ny[f] + nb, 8 * n, cudaMemcpyDeviceToHost));
            OCTA_CUDA_CHECK(cudaMemcpy(tmp.data() + 2 * (size_t)n, D.nz[f] + nb, 8 * n, cudaMemcpyDeviceToHost));
            for (int i = 0; i < n; ++i) { ex[g].pos[f][3 * i] = tmp[i]; ex[g].pos[f][3 * i + 1] = tmp[n + i]; ex[g].pos[f][3 * i + 2] = tmp[2 * (size_t)n + i]; }
            OCTA_CUDA_CHECK(cudaMemcpy(ex[g].parent[f].data(), D.npar[f] + nb, 4 * n, cudaMemcpyDeviceToHost));
            OCTA_CUDA_CHECK(cudaMemcpy(ex[g].meta[f].data(), D.nmeta[f] + nb, n, cudaMemcpyDeviceToHost));
        }
    }
    // exact radii + edge export, multi-threaded over graphs
    const OctaGrowConfig c = *cfg;
    unsigned nthreads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)n_graphs));
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nthreads; ++w)
        pool.emplace_back([&, w]() {
            for (int g = (int)w; g < n_graphs; g += (int)nthreads) {
                double* out = edges7_out + (size_t)g * cap_edges * 7;
                int64_t na = 0, nv = 0;
                finalize_forest(c, ex[g].pos[0], ex[g].parent[0], ex[g].meta[0], out, cap_edges, &na);
                const int64_t used = na < cap_edges ? na : cap_edges;
                finalize_forest(c, ex[g].pos[1], ex[g].parent[1], ex[g].meta[1], out + 7 * used, cap_edges - used, &nv);
                n_art_edges[g] = na;
                n_ven_edges[g] = nv;
            }
        });
    for (auto& t : pool) t.join();
    int worst = 0;
    for (int g = 0; g < n_graphs; ++g) {
        if (stats) {
            OctaGrowStats& s = stats[g];
            s.n_art_nodes = nn[0][g]; s.n_ven_nodes = nn[1][g]; s.n_oxy_left = ns[0][g]; s.n_co2_left = ns[1][g];
            s.py_draws = draws[g];
            s.sum_A = counters[(size_t)g * 8 + 0]; s.sum_M = counters[(size_t)g * 8 + 1];
            s.sum_P = counters[(size_t)g * 8 + 2]; s.sum_S = counters[(size_t)g * 8 + 3];
            s.err = err[g]; s.n_iters = (int)sched.size();
        }
        if (err[g] && !worst) worst = err[g];
        if (n_art_edges[g] + n_ven_edges[g] > cap_edges && !worst) worst = 100;
    }
    if (worst) {
        set_error("octa_grow_batch_host: simulation error code %d (1 node capacity, 2 sink capacity, 3 rng buffer, "
                  "4 recheck queue, 5 set table, 12/13 eigen solver, 100 edge buffer too small)", worst);
        return OCTA_E_STATE;
    }
    return OCTA_OK;
}
