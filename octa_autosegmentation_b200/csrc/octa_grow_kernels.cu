// K1-K5: space-colonization growth kernels for sm_100a (one launch = one phase for the whole batch).
//
// Replaces, per iteration of Greenhouse.develop_forest (greenhouse.py:90-125):
//   k_sample        simulation_space.py:57-67 get_candidate_sinks (legacy numpy MT19937 stream on device)
//   k_grid_build    (ours) bucket grids over nodes / sinks / active nodes; replaces the cKDTree rebuilds (element_mesh.py:97-101)
//   k_sink_tests    greenhouse.py:337-338   static rejection tests (i) arterial oxygen range, (ii) sink spacing
//   k_sink_greedy   greenhouse.py:339-341   order-dependent spacing among new sinks (lexicographically-first MIS)
//   k_assign        greenhouse.py:343-366   nearest ACTIVE node of every attractor (<= delta)
//   k_group         greenhouse.py:357-365   dict order (= first-attractor order) and per-node attractor lists
//   k_eval          greenhouse.py:174-306   per-node growth proposal (all float64 math, parallel)
//   k_commit        greenhouse.py:191,235-239,289,303-306 + arterial_tree.py:174-184
//                   sequential replay in dict order: Python-RNG draws, node creation, Murray walk
//   k_kill          greenhouse.py:99-123    kill-radius prune (bucket grids), O2 -> CO2 through a CPython-set emulation with an
//                   order test; k_kdbuild_list builds cKDTree's index permutation for the graphs whose result depends on the
//                   order inside a ball (element_mesh.py:136-137) and finishes their kill
// Bit-level conventions are documented in octa_grow_math.cuh.  This TU is compiled with -fmad=false.
#include "octa_common.h"
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>
#include <cuda/atomic>
#include <vector>
#include "octa_eig3.h"
#include "octa_grow.cuh"
#include "octa_grow_math.cuh"
#include "octa_kdorder_par.cuh"
#include "octa_pyset.cuh"

namespace octa {

// The table of device pointers (GrowDev, ~0.9 KB) lives in constant memory, one slot per growth context, instead of
// travelling with each of the ~3 750 launches of a run: kernels receive the slot index.
constexpr int MAX_CTX_SLOTS = 16;
__constant__ GrowDev c_dev[MAX_CTX_SLOTS];

// ------------------------------------------------------------------------------------------
// block-wide helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_scan_incl(int v, int* total) {
    __shared__ int ws[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? ws[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        ws[lane] = w;
        if (lane == 31) ws[32] = w;
    }
    __syncthreads();
    if (warp > 0) x += ws[warp - 1];
    *total = ws[32];
    __syncthreads();
    return x;
}

// MT19937 state regeneration by a whole block (>= 256 threads); mt in shared memory
__device__ void mt_regen_block(uint32_t* mt) {
    const uint32_t UP = 0x80000000u, LOW = 0x7fffffffu, MAG = 0x9908b0dfu;
    const int tid = threadIdx.x;
    for (int phase = 0; phase < 3; ++phase) {
        const int base = phase * 227;
        const int cnt = phase == 2 ? 170 : 227;   // 454..623
        uint32_t v[3];
        int nk = 0;
        for (int q = tid; q < cnt; q += blockDim.x) {
            const int k = base + q;
            const uint32_t nxt = (k == 623) ? mt[0] : mt[k + 1];
            const uint32_t y = (mt[k] & UP) | (nxt & LOW);
            const uint32_t src = (k < 227) ? mt[k + 397] : mt[k - 227];
            if (nk < 3) v[nk] = src ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
            ++nk;
        }
        __syncthreads();
        nk = 0;
        for (int q = tid; q < cnt; q += blockDim.x) { mt[base + q] = v[nk < 3 ? nk : 2]; ++nk; }
        __syncthreads();
    }
}

__device__ __forceinline__ double dist2(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return (dx * dx + dy * dy) + dz * dz;   // cKDTree / np.linalg.norm(axis=1) association
}

// x**y for x > 0 through exp/log: ~1e-15*|y ln x| relative error, used only for the in-loop (steering) radii
__device__ __forceinline__ double fast_pow(double x, double y) { return exp(y * log(x)); }

// sqrt(d2) <= r, avoiding the square root outside a narrow band around the threshold
__device__ __forceinline__ bool within_sqrt(double d2, double r, double r2) {
    if (d2 < r2 * (1.0 - 1e-12)) return true;
    if (d2 > r2 * (1.0 + 1e-12)) return false;
    return sqrt(d2) <= r;
}

// ------------------------------------------------------------------------------------------
// k_sample: one CTA per graph
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sample_body(const GrowDev& D, const GrowShape& S, const IterP& P, int g) {
    __shared__ uint32_t mt[624];
    __shared__ int s_idx, s_p;
    const int tid = threadIdx.x;
    if (D.err[g]) return;
    MTState* st = D.np_mt + g;
    for (int i = tid; i < 624; i += blockDim.x) mt[i] = st->mt[i];
    if (tid == 0) s_idx = st->idx;
    __syncthreads();
    const int N = P.N;
    const uint32_t L = (uint32_t)D.n_valid[g];
    const uint32_t rng = L - 1;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    unsigned int* vi = D.vi + (size_t)g * S.Nmax;
    unsigned int* ub = D.ubuf + (size_t)g * 6 * S.Nmax;
    // 1. randint(0, L, N): masked rejection over the 32-bit stream
    int count = 0;
    while (count < N) {
        if (s_idx >= 624) { mt_regen_block(mt); if (tid == 0) s_idx = 0; __syncthreads(); }
        const int idx = s_idx;
        const int avail = 624 - idx;
        uint32_t x = 0;
        int flag = 0;
        if (tid < avail) {
            x = (rng == 0) ? 0u : (mt_temper(mt[idx + tid]) & mask);
            flag = (rng == 0) ? 1 : (x <= rng);
        }
        int total;
        const int incl = block_scan_incl(flag, &total);
        const int need = N - count;
        if (flag && incl <= need) vi[count + incl - 1] = x;
        if (flag && incl == need) s_p = tid;
        __syncthreads();
        int consumed;
        if (total >= need) { consumed = s_p + 1; count = N; } else { consumed = avail; count += total; }
        __syncthreads();
        if (tid == 0) s_idx = idx + ((rng == 0) ? 0 : consumed);   // rng == 0 draws nothing (numpy)
        __syncthreads();
    }
    // 2. uniform(0,1,(N,3)): 6N stream words
    int written = 0;
    while (written < 6 * N) {
        if (s_idx >= 624) { mt_regen_block(mt); if (tid == 0) s_idx = 0; __syncthreads(); }
        const int idx = s_idx;
        const int avail = 624 - idx;
        const int take = avail < 6 * N - written ? avail : 6 * N - written;
        if (tid < take) ub[written + tid] = mt_temper(mt[idx + tid]);
        written += take;
        __syncthreads();
        if (tid == 0) s_idx = idx + take;
        __syncthreads();
    }
    for (int i = tid; i < 624; i += blockDim.x) st->mt[i] = mt[i];
    if (tid == 0) st->idx = s_idx;
    // 3. candidates, filtered by is_valid_position (simulation_space.py:89-98), stable order
    const unsigned char* vij = D.valid_ij + (size_t)g * MAX_VALID * 2;
    double* cx = D.cx + (size_t)g * S.Nmax; double* cy = D.cy + (size_t)g * S.Nmax; double* cz = D.cz + (size_t)g * S.Nmax;
    const double fzc0 = P.faz_cx * GEOMETRY_SIZE, fzc1 = P.faz_cy * GEOMETRY_SIZE;
    const double fzr = D.faz_radius[g] * GEOMETRY_SIZE * 0.5;
    int ncand = 0;
    for (int base = 0; base < N; base += blockDim.x) {
        const int i = base + tid;
        int valid = 0;
        double px = 0, py = 0, pz = 0;
        if (i < N) {
            const uint32_t v = vi[i];
            const double u0 = mt_double(ub[6 * i], ub[6 * i + 1]), u1 = mt_double(ub[6 * i + 2], ub[6 * i + 3]),
                         u2 = mt_double(ub[6 * i + 4], ub[6 * i + 5]);
            const double gs = P.geom_gs ? (double)P.geom_gs : (double)GEOMETRY_SIZE;      // simulation_space.py:30 / :41
            double vx, vy, vz = 0.0;      // the voxel: argwhere row of the mask (3-D), or a pixel of the FAZ mask (positions miss the z dim)
            if (P.geom_gs) { const unsigned short* t = D.geom_valid + 3 * (size_t)v; vx = (double)t[0]; vy = (double)t[1]; vz = (double)t[2]; }
            else { vx = (double)vij[2 * v]; vy = (double)vij[2 * v + 1]; }
            px = (vx + u0) / gs;
            py = (vy + u1) / gs;
            pz = (vz + u2) / gs;
            valid = !(px >= P.shape[0] || py >= P.shape[1] || pz >= P.shape[2] || px < 0 || py < 0 || pz < 0);
            if (valid && P.geom_gs) {
                // fixed geometry (simulation_space.py:95-96): geometry[(pos * geometry_size).astype(uint16)] > 0
                const int vi_ = (int)(unsigned short)(px * gs), vj_ = (int)(unsigned short)(py * gs), vk_ = (int)(unsigned short)(pz * gs);
                if (vi_ >= P.geom_dims[0] || vj_ >= P.geom_dims[1] || vk_ >= P.geom_dims[2]) { D.err[g] = 6; valid = 0; }   // numpy would raise IndexError
                else valid = D.geom_mask[((size_t)vi_ * P.geom_dims[1] + vj_) * P.geom_dims[2] + vk_] != 0;
            } else if (valid) {   // zip-truncated eukledian_dist(pos, FAZ_center[voxel units]) > FAZ_radius[voxel units]
                const double a = px - fzc0, b = py - fzc1;
                valid = sqrt(a * a + b * b) > fzr;
            }
        }
        int total;
        const int incl = block_scan_incl(valid, &total);
        if (valid) { const int o = ncand + incl - 1; cx[o] = px; cy[o] = py; cz[o] = pz; }
        ncand += total;
    }
    if (tid == 0) {
        D.n_cand[g] = ncand;
        D.counters[g * 8 + 2] += D.n_nodes[0][g];   // sum_P
        D.counters[g * 8 + 3] += D.n_s[0][g];       // sum_S
    }
}

// ------------------------------------------------------------------------------------------
// uniform bucket grid (GRID x GRID cells over the unit square, points outside clamp to the border cells):
// prunes the pair scans of k_sink_tests / k_assign; every reported distance is still the exact float64
// expression, so results are identical to a brute-force scan (= cKDTree's exact queries).
// Points are counting-sorted by cell with their coordinates copied next to each other, so one row of
// cells is one contiguous range of the sorted arrays.
// ------------------------------------------------------------------------------------------
constexpr int TILE = 128;

__device__ __forceinline__ int grid_cell(double v) {
    int c = (int)floor(v * (double)GRID);
    return c < 0 ? 0 : (c >= GRID ? GRID - 1 : c);
}

// which: 0 = all arterial nodes (+radius), 1 = O2 sinks, 2 / 3 = active arterial / venous nodes, 4 = all venous nodes (k_kill's veto)
__device__ __forceinline__ void grid_build_body(const GrowDev& D, const GrowShape& S, const IterP& P, int which, int g) {
    __shared__ int hist[GRID * GRID];
    __shared__ int cursor[GRID * GRID];
    const int tid = threadIdx.x;
    if (D.err[g]) return;
    // The ACTIVE node set (element_mesh.py NodeKdTree of active nodes) is the node array minus the nodes that
    // branched: list order = creation order with deletions, so node ids preserve the list's relative order and
    // no separate list has to be maintained.
    const double *px, *py, *pz, *pr = nullptr;
    const unsigned char* skip = nullptr;
    int n;
    size_t cap;
    if (which == 0) { cap = S.capN; px = D.nx[0] + g * cap; py = D.ny[0] + g * cap; pz = D.nz[0] + g * cap; pr = D.ncon[0] + g * cap; n = D.n_nodes[0][g]; }
    else if (which == 1) { cap = S.capS; px = D.sx[0] + g * cap; py = D.sy[0] + g * cap; pz = D.sz[0] + g * cap; n = D.n_s[0][g]; }
    else if (which == 4) { cap = S.capN; px = D.nx[1] + g * cap; py = D.ny[1] + g * cap; pz = D.nz[1] + g * cap; n = D.n_nodes[1][g]; }   // every venous node
    else { const int f = which - 2; cap = S.capN; px = D.nx[f] + g * cap; py = D.ny[f] + g * cap; pz = D.nz[f] + g * cap; n = D.n_nodes[f][g]; skip = D.deact[f] + g * cap; }
    const size_t gcap = S.capN > S.capS ? S.capN : S.capS;
    double* gx = D.gx[which] + g * gcap; double* gy = D.gy[which] + g * gcap; double* gz = D.gz[which] + g * gcap;
    double* grd = D.gr[which] + g * gcap;
    int* gi = D.gi[which] + g * gcap;
    int* cs = D.gcell[which] + (size_t)g * (GRID * GRID + 1);
    for (int c = tid; c < GRID * GRID; c += blockDim.x) hist[c] = 0;
    __syncthreads();
    int live = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        if (skip && skip[i]) continue;
        ++live;
        atomicAdd(&hist[grid_cell(py[i]) * GRID + grid_cell(px[i])], 1);
    }
    (void)live;
    __syncthreads();
    int run = 0;
    for (int base = 0; base < GRID * GRID; base += blockDim.x) {
        const int c = base + tid;
        const int v = hist[c];
        int total;
        const int incl = block_scan_incl(v, &total);
        cs[c] = run + incl - v;
        cursor[c] = run + incl - v;
        run += total;
    }
    if (tid == 0) { cs[GRID * GRID] = run; if (which == 2 || which == 3) D.n_act[which - 2][g] = run; }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        if (skip && skip[i]) continue;
        const double x = px[i], y = py[i];
        const int pos = atomicAdd(&cursor[grid_cell(y) * GRID + grid_cell(x)], 1);
        gx[pos] = x; gy[pos] = y; gz[pos] = pz[i]; gi[pos] = i;
        if (pr) {
            // radius from the Murray contribution: ncon = radius ** kappa(parent); leaves and unbranched chains carry r itself
            const int par = D.npar[0][g * cap + i];
            double rad = P.r;
            if (par >= 0) {
                const int m = D.nmeta[0][g * cap + par];
                const int kmp = m == 0xff ? 8 : (m >> 1);
                const double cv = pr[i];
                if (cv != P.leafc_tab[kmp]) rad = fast_pow(cv, 1.0 / P.kap_tab[kmp]);
            }
            grd[pos] = rad;
        }
    }
}

// grid.y = 2 with which = 3: the grids of the active venous nodes (blockIdx.y = 0) and of all venous nodes (blockIdx.y = 1) side by side
__global__ void __launch_bounds__(1024) k_grid_build(int dslot, GrowShape S, IterP P, int which) { grid_build_body(c_dev[dslot], S, P, which + (int)blockIdx.y, blockIdx.x); }

// k_prepare: everything the sampling of an iteration needs, in ONE launch: blockIdx.y = 0..2 -> the bucket grids of
// the arterial nodes (+radius), the O2 sinks and the active arterial nodes, blockIdx.y = 3 -> the candidate sampler.
// 4*G CTAs run side by side instead of four dependent single-wave launches.  (The grid of the active venous nodes is
// rebuilt by k_grid_build right after the venous commit.)
__global__ void __launch_bounds__(1024, 2) k_prepare(int dslot, GrowShape S, IterP P) {
    const GrowDev& D = c_dev[dslot];
    if (blockIdx.y < 3) grid_build_body(D, S, P, (int)blockIdx.y, blockIdx.x);
    else sample_body(D, S, P, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// k_sink_tests: thread per candidate over the bucket grids of the arterial nodes and the O2 sinks
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) k_sink_tests(int dslot, GrowShape S, IterP P) {
    const GrowDev& D = c_dev[dslot];
    const int tiles = (S.Nmax + TILE - 1) / TILE;
    const double epsn2 = P.eps_n_eff * P.eps_n_eff, epss2 = P.eps_s * P.eps_s;
    const size_t gcap = S.capN > S.capS ? S.capN : S.capS;
    for (int w = blockIdx.x; w < S.G * tiles; w += gridDim.x) {
        const int g = w / tiles, tile = w - g * tiles;
        if (D.err[g]) continue;
        const int nc = D.n_cand[g];
        const int c = tile * TILE + threadIdx.x;
        if (c >= nc) continue;
        const double px = D.cx[(size_t)g * S.Nmax + c], py = D.cy[(size_t)g * S.Nmax + c], pz = D.cz[(size_t)g * S.Nmax + c];
        bool pass = true;
        // (i) every arterial node within eps must be farther than its oxygen range (greenhouse.py:337)
        {
            const double* gx = D.gx[0] + g * gcap; const double* gy = D.gy[0] + g * gcap; const double* gz = D.gz[0] + g * gcap;
            const double* gr = D.gr[0] + g * gcap;
            const int* cs = D.gcell[0] + (size_t)g * (GRID * GRID + 1);
            const int x0 = grid_cell(px - P.eps_n_eff), x1 = grid_cell(px + P.eps_n_eff);
            const int y0 = grid_cell(py - P.eps_n_eff), y1 = grid_cell(py + P.eps_n_eff);
            for (int cy = y0; cy <= y1 && pass; ++cy) {
                const int beg = cs[cy * GRID + x0], end = cs[cy * GRID + x1 + 1];
                for (int q = beg; q < end; ++q) {
                    const double d2 = dist2(px, py, pz, gx[q], gy[q], gz[q]);
                    if (d2 <= epsn2 && !(sqrt(d2) > oxygen_distance(gr[q], P.param_scale))) { pass = false; break; }
                }
            }
        }
        // (ii) no existing sink within eps_s (greenhouse.py:338)
        if (pass) {
            const double* gx = D.gx[1] + g * gcap; const double* gy = D.gy[1] + g * gcap; const double* gz = D.gz[1] + g * gcap;
            const int* cs = D.gcell[1] + (size_t)g * (GRID * GRID + 1);
            const int x0 = grid_cell(px - P.eps_s), x1 = grid_cell(px + P.eps_s);
            const int y0 = grid_cell(py - P.eps_s), y1 = grid_cell(py + P.eps_s);
            for (int cy = y0; cy <= y1 && pass; ++cy) {
                const int beg = cs[cy * GRID + x0], end = cs[cy * GRID + x1 + 1];
                for (int q = beg; q < end; ++q)
                    if (within_sqrt(dist2(px, py, pz, gx[q], gy[q], gz[q]), P.eps_s, epss2)) { pass = false; break; }
            }
        }
        D.cpass[(size_t)g * S.Nmax + c] = pass ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// k_sink_greedy: one CTA per graph; lexicographically-first maximal independent set in rounds
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 2) k_sink_greedy(int dslot, GrowShape S, IterP P) {
    const GrowDev& D = c_dev[dslot];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (D.err[g]) return;
    const int nc = D.n_cand[g];
    const double* cx = D.cx + (size_t)g * S.Nmax; const double* cy = D.cy + (size_t)g * S.Nmax; const double* cz = D.cz + (size_t)g * S.Nmax;
    const unsigned char* cp = D.cpass + (size_t)g * S.Nmax;
    int* pl = D.plist + (size_t)g * S.Nmax;
    // decision state per passing candidate: shared memory (read by every other candidate every round)
    // (32-bit cells accessed through block-scope relaxed atomics: plain LDS / STS, and the memory model -- and compute-sanitizer --
    // see the cross-thread reads inside a round as what they are)
    __shared__ unsigned int s_state[4096];
    unsigned int* state_p = (S.Nmax <= 4096) ? s_state : (D.cstate32 + (size_t)g * S.Nmax);
    auto st_load = [&](int k) { return cuda::atomic_ref<unsigned int, cuda::thread_scope_block>(state_p[k]).load(cuda::memory_order_relaxed); };
    auto st_store = [&](int k, unsigned int v) { cuda::atomic_ref<unsigned int, cuda::thread_scope_block>(state_p[k]).store(v, cuda::memory_order_relaxed); };
    int np_ = 0;
    for (int base = 0; base < nc; base += blockDim.x) {
        const int i = base + tid;
        const int f = (i < nc) ? cp[i] : 0;
        int total;
        const int incl = block_scan_incl(f, &total);
        if (f) { pl[np_ + incl - 1] = i; st_store(np_ + incl - 1, 0u); }
        np_ += total;
    }
    __syncthreads();
    const double eps2 = P.eps_s * P.eps_s;
    // state: 0 undecided, 1 accepted, 2 rejected.  Candidate k is rejected as soon as an earlier ACCEPTED
    // candidate lies within eps_s, accepted once every earlier candidate within eps_s is rejected.
    // Threads read state[j] of other threads inside a round without a barrier: a state only ever moves 0 -> 1 or 0 -> 2, and either value seen gives a decision the
    // sequential greedy would also reach, so the fixed point is unique; early visibility just saves rounds.
    // The passing candidates are binned into a 32 x 32 grid (one row of cells = one contiguous range of `s_sorted`): a candidate
    // only looks at the earlier candidates of the cells within eps_s instead of at all of them (the test itself is unchanged, and
    // the decision depends on the SET of earlier neighbours and their states, not on the order they are visited in).
    constexpr int G2 = 32;
    __shared__ unsigned short s_sorted[8192];
    __shared__ int s_cstart[G2 * G2 + 1];
    __shared__ int s_cursor[G2 * G2];
    const bool binned = np_ <= 8192;
    auto cell2 = [](double v) { int c = (int)floor(v * (double)G2); return c < 0 ? 0 : (c >= G2 ? G2 - 1 : c); };
    if (binned) {
        for (int c = tid; c < G2 * G2; c += blockDim.x) s_cursor[c] = 0;
        __syncthreads();
        for (int k = tid; k < np_; k += blockDim.x) { const int ik = pl[k]; atomicAdd(&s_cursor[cell2(cy[ik]) * G2 + cell2(cx[ik])], 1); }
        __syncthreads();
        int run = 0;
        for (int base = 0; base < G2 * G2; base += blockDim.x) {
            const int c = base + tid;
            const int v = c < G2 * G2 ? s_cursor[c] : 0;
            int total;
            const int incl = block_scan_incl(v, &total);
            if (c < G2 * G2) { s_cstart[c] = run + incl - v; s_cursor[c] = run + incl - v; }
            run += total;
        }
        if (tid == 0) s_cstart[G2 * G2] = run;
        __syncthreads();
        for (int k = tid; k < np_; k += blockDim.x) { const int ik = pl[k]; s_sorted[atomicAdd(&s_cursor[cell2(cy[ik]) * G2 + cell2(cx[ik])], 1)] = (unsigned short)k; }
        __syncthreads();
    }
    const double er = P.eps_s * (1.0 + 1e-9) + 1e-12;             // the cell range may only be too wide
    for (int round = 0; round < nc + 2; ++round) {
        int undecided = 0;
        for (int k = tid; k < np_; k += blockDim.x) {
            if (st_load(k) != 0u) continue;
            const int ik = pl[k];
            const double px = cx[ik], py = cy[ik], pz = cz[ik];
            bool acc = false, und = false;
            if (binned) {
                const int x0 = cell2(px - er), x1 = cell2(px + er), y0 = cell2(py - er), y1 = cell2(py + er);
                for (int yy = y0; yy <= y1 && !acc; ++yy) {
                    const int beg = s_cstart[yy * G2 + x0], end = s_cstart[yy * G2 + x1 + 1];
                    for (int q = beg; q < end; ++q) {
                        const int j = s_sorted[q];
                        if (j >= k) continue;
                        const unsigned int sj = st_load(j);
                        if (sj == 2) continue;
                        const int ij = pl[j];
                        const double d2 = dist2(px, py, pz, cx[ij], cy[ij], cz[ij]);
                        if (within_sqrt(d2, P.eps_s, eps2)) {   // not (norm > eps_s)
                            if (sj == 1) { acc = true; break; }
                            und = true;
                        }
                    }
                }
            } else
            for (int j = 0; j < k; ++j) {
                const unsigned int sj = st_load(j);
                if (sj == 2) continue;
                const int ij = pl[j];
                const double d2 = dist2(px, py, pz, cx[ij], cy[ij], cz[ij]);
                if (within_sqrt(d2, P.eps_s, eps2)) {   // not (norm > eps_s)
                    if (sj == 1) { acc = true; break; }
                    und = true;
                }
            }
            if (acc) st_store(k, 2u);
            else if (!und) st_store(k, 1u);
            else undecided = 1;
        }
        __threadfence_block();
        if (!__syncthreads_or(undecided)) break;
    }
    __syncthreads();
    // append the accepted candidates, in candidate order, to the oxygen sink list
    int n0 = D.n_s[0][g];
    const size_t sb = (size_t)g * S.capS;
    for (int base = 0; base < np_; base += blockDim.x) {
        const int k = base + tid;
        const int f = (k < np_) ? (st_load(k) == 1u) : 0;
        int total;
        const int incl = block_scan_incl(f, &total);
        if (f) {
            const int o = n0 + incl - 1;
            if (o < S.capS) { const int ik = pl[k]; D.sx[0][sb + o] = cx[ik]; D.sy[0][sb + o] = cy[ik]; D.sz[0][sb + o] = cz[ik]; }
        }
        n0 += total;
    }
    if (tid == 0) {
        if (n0 > S.capS) { D.err[g] = 2; n0 = S.capS; }
        D.n_s[0][g] = n0;
    }
}

// ------------------------------------------------------------------------------------------
// k_assign: thread per attractor; exact nearest ACTIVE node within delta through the bucket grid
// (ties -> lowest list position, as a scan in list order would give)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) k_assign(int dslot, GrowShape S, IterP P, int f) {
    const GrowDev& D = c_dev[dslot];
    const int tiles = (S.capS + TILE - 1) / TILE;
    const double delta = P.delta[f];
    const size_t gcap = S.capN > S.capS ? S.capN : S.capS;
    for (int w = blockIdx.x; w < S.G * tiles; w += gridDim.x) {
        const int g = w / tiles, tile = w - g * tiles;
        if (D.err[g]) continue;
        const int A = D.n_s[f][g];
        if (tile == 0 && threadIdx.x == 0) { D.counters[g * 8 + 0] += A; D.counters[g * 8 + 1] += D.n_act[f][g]; }
        const int a = tile * TILE + threadIdx.x;
        if (a >= A) continue;
        const size_t sb = (size_t)g * S.capS, nb = (size_t)g * S.capN;
        const double px = D.sx[f][sb + a], py = D.sy[f][sb + a], pz = D.sz[f][sb + a];
        const double* gx = D.gx[2 + f] + g * gcap; const double* gy = D.gy[2 + f] + g * gcap; const double* gz = D.gz[2 + f] + g * gcap;
        const int* gi = D.gi[2 + f] + g * gcap;
        const int* cs = D.gcell[2 + f] + (size_t)g * (GRID * GRID + 1);
        const int x0 = grid_cell(px - delta), x1 = grid_cell(px + delta), y0 = grid_cell(py - delta), y1 = grid_cell(py + delta);
        // rows of cells are visited outwards from the attractor's own row, and a row whose nearest edge is already farther than the
        // best node found so far is skipped (its points cannot win, nor tie: the bound is strict and slightly loosened) -- the
        // argmin with ties to the lowest list position is what a scan of every row in the range finds
        double best = INFINITY;
        int bi = -1;
        const int cyc = grid_cell(py);
        for (int step = 0; step <= 2 * (GRID - 1); ++step) {
            const int cy = (step & 1) ? cyc + ((step + 1) >> 1) : cyc - (step >> 1);
            if (cy < y0 || cy > y1) { if (cyc - ((step + 1) >> 1) < y0 && cyc + ((step + 1) >> 1) > y1) break; continue; }
            double dyr = 0.0;                                   // distance of py from the row's band (border rows also hold the outside)
            if (cy < cyc) dyr = py - (double)(cy + 1) / (double)GRID;
            else if (cy > cyc) dyr = (double)cy / (double)GRID - py;
            dyr -= 1e-12;
            if (dyr > 0.0 && dyr * dyr > best) continue;
            const int beg = cs[cy * GRID + x0], end = cs[cy * GRID + x1 + 1];
            for (int q = beg; q < end; ++q) {
                const double d2 = dist2(gx[q], gy[q], gz[q], px, py, pz);
                const int li = gi[q];
                if (d2 < best || (d2 == best && li < bi)) { best = d2; bi = li; }
            }
        }
        D.assign[sb + a] = (bi >= 0 && sqrt(best) <= delta) ? bi : -1;
    }
}

// ------------------------------------------------------------------------------------------
// k_group: one CTA per graph
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_group(int dslot, GrowShape S, IterP P, int f) {
    const GrowDev& D = c_dev[dslot];
    __shared__ uint32_t mt[624];
    __shared__ int s_idx;
    const int g = blockIdx.x, tid = threadIdx.x;
    if (D.err[g]) return;
    const int call_id = 2 * P.iter + f + 1;
    const int A = D.n_s[f][g];
    const size_t sb = (size_t)g * S.capS, nb = (size_t)g * S.capN;
    const int* asg = D.assign + sb;
    int* first = D.first + nb; int* cnt = D.cnt + nb; int* slot = D.slot + nb; int* cur = D.cur + nb;
    int* dict = D.dict_node + nb; int* loff = D.list_off + (size_t)g * (S.capN + 1); int* lst = D.list + sb;
    for (int a = tid; a < A; a += blockDim.x) {
        const int nd = asg[a];
        if (nd >= 0) { atomicMin(&first[nd], a); atomicAdd(&cnt[nd], 1); }
    }
    __syncthreads();
    int nd_ = 0;
    for (int base = 0; base < A; base += blockDim.x) {
        const int a = base + tid;
        int isf = 0, nd = -1;
        if (a < A) { nd = asg[a]; isf = (nd >= 0 && first[nd] == a); }
        int total;
        const int incl = block_scan_incl(isf, &total);
        if (isf) { const int rk = nd_ + incl - 1; dict[rk] = nd; slot[nd] = rk; }
        nd_ += total;
    }
    __syncthreads();
    int off = 0;
    for (int base = 0; base < nd_; base += blockDim.x) {
        const int e = base + tid;
        const int c = (e < nd_) ? cnt[dict[e]] : 0;
        int total;
        const int incl = block_scan_incl(c, &total);
        if (e < nd_) { loff[e] = off + incl - c; cur[e] = 0; }
        off += total;
    }
    if (tid == 0) { loff[nd_] = off; D.n_dict[g] = nd_; }
    __syncthreads();
    for (int a = tid; a < A; a += blockDim.x) {
        const int nd = asg[a];
        if (nd >= 0) { const int e = slot[nd]; const int pos = atomicAdd(&cur[e], 1); lst[loff[e] + pos] = a; }
    }
    __syncthreads();
    for (int e = tid; e < nd_; e += blockDim.x) {
        int* l = lst + loff[e];
        const int n = loff[e + 1] - loff[e];
        for (int i = 1; i < n; ++i) {       // lists are short (mean 2, max ~150): insertion sort
            const int v = l[i];
            int j = i - 1;
            while (j >= 0 && l[j] > v) { l[j + 1] = l[j]; --j; }
            l[j + 1] = v;
        }
        const int nd = dict[e];
        first[nd] = 0x7fffffff;
        cnt[nd] = 0;
    }
    // top up the Python-`random` word buffer: k_commit may consume one double per dict entry
    unsigned int* pb = D.py_buf + (size_t)g * S.pycap;
    int pos = D.py_pos[g], n = D.py_n[g];
    const int need = 2 * nd_ + 2;
    __syncthreads();
    if (n - pos < need) {
        const int rem = n - pos;
        for (int base = 0; base < rem; base += blockDim.x) {   // move the unread tail to the front
            const int i = base + tid;
            unsigned int v = 0;
            if (i < rem) v = pb[pos + i];
            __syncthreads();
            if (i < rem) pb[i] = v;
            __syncthreads();
        }
        MTState* st = D.py_mt + g;
        for (int i = tid; i < 624; i += blockDim.x) mt[i] = st->mt[i];
        if (tid == 0) s_idx = st->idx;
        __syncthreads();
        int have = rem;
        while (have < need && have + 624 <= S.pycap) {
            if (s_idx >= 624) { mt_regen_block(mt); if (tid == 0) s_idx = 0; __syncthreads(); }
            const int idx = s_idx, avail = 624 - idx;
            if (tid < avail) pb[have + tid] = mt_temper(mt[idx + tid]);
            have += avail;
            __syncthreads();
            if (tid == 0) s_idx = 624;
            __syncthreads();
        }
        for (int i = tid; i < 624; i += blockDim.x) st->mt[i] = mt[i];
        if (tid == 0) { st->idx = s_idx; D.py_pos[g] = 0; D.py_n[g] = have; if (have < need) D.err[g] = 3; }
    }
}

// ------------------------------------------------------------------------------------------
// per-node growth proposals
// ------------------------------------------------------------------------------------------
struct NodeCtx {
    double pos[3], par[3], ch[3];
    double vtc[2], dist_to_center;
    int nch, parent;
};

// utilities.py:42-45: angle (degrees) between u and every (att - pos); n = list length (selects the BLAS path)
__device__ __forceinline__ double angle_to(const double* u, double nu, const double* v, int n) {
    const double dt = (n == 1) ? ddot3(u, v) : gemv3(u, v);
    const double C = dt / nu / norm3_axis(v);
    return RAD2DEG * acos(clamp11(C));
}

// leaf branch, greenhouse.py:177-258
__device__ void eval_leaf(const GrowDev& D, const GrowShape& S, const IterP& P, int g, int f, int e, const NodeCtx& nc,
                          const int* lst, int n, Proposal* pr) {
    const size_t sb = (size_t)g * S.capS;
    const double* sx = D.sx[f] + sb; const double* sy = D.sy[f] + sb; const double* sz = D.sz[f] + sb;
    double* sang = D.sc_ang + sb + (lst - (D.list + sb));
    int* sidx = D.sc_idx + sb + (lst - (D.list + sb));
    const double v[3] = {nc.pos[0] - nc.par[0], nc.pos[1] - nc.par[1], nc.pos[2] - nc.par[2]};
    const double nv = norm3(v);
    const double lim = P.gamma[f] / 2 > 0 ? P.gamma[f] / 2 : 0;
    int m = 0;
    double avg[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const int a = lst[i];
        const double w[3] = {sx[a] - nc.pos[0], sy[a] - nc.pos[1], sz[a] - nc.pos[2]};
        const double ang = angle_to(v, nv, w, n);
        if (ang <= lim) {
            const double nw = norm3(w);
            if (m == 0) { avg[0] = w[0] / nw; avg[1] = w[1] / nw; avg[2] = w[2] / nw; }
            else { avg[0] += w[0] / nw; avg[1] += w[1] / nw; avg[2] += w[2] / nw; }
            sang[m] = ang;
            sidx[m] = a;
            ++m;
        }
    }
    pr->type = P_NONE;
    if (m == 0) return;
    // np.std(angles): population std via numpy's pairwise sums
    const double mean = pairwise_sum([&](long i) { return sang[i]; }, 0, m) / (double)m;
    const double var = pairwise_sum([&](long i) { const double x = sang[i] - mean; return x * x; }, 0, m) / (double)m;
    const double sd = sqrt(var);
    const double faz = D.faz_radius[g];
    // elongation, greenhouse.py:242-258 (always prepared: it is the fall-back of a failed bifurcation test)
    {
        const double na = norm3(avg);
        double gv[3];
        for (int k = 0; k < 3; ++k) gv[k] = P.omega * (v[k] / nv) + (1 - P.omega) * (avg[k] / na);
        if (P.rotation_radius > 0 && P.t > 15) {
            const double ng = norm3(gv);
            for (int k = 0; k < 3; ++k) gv[k] /= ng;
            double cv[2] = {P.faz_cx - nc.pos[0], P.faz_cy - nc.pos[1]};
            const double ncv = norm2(cv);
            cv[0] /= ncv; cv[1] /= ncv;
            const double np2[2] = {P.faz_cx - (nc.pos[0] + P.d * gv[0]), P.faz_cy - (nc.pos[1] + P.d * gv[1])};
            const double dist_new = norm2(np2);
            const double floorw = P.first_mode ? 0.0 : 0.01;
            const double cand = P.rotation_radius - dist_new;
            double weight = cand > floorw ? cand : floorw;
            weight = sqrt(weight);
            double ort[3] = {-cv[1], cv[0], 0};
            if (angle_between_two(gv, ort) > 90) { ort[0] = -1 * ort[0]; ort[1] = -1 * ort[1]; ort[2] = -1 * ort[2]; }
            const double outv[3] = {-cv[0], -cv[1], 0};
            for (int k = 0; k < 3; ++k) gv[k] = ((1 - weight) * gv[k] + 0.7 * weight * ort[k]) + 0.3 * weight * outv[k];
        }
        const double ng = norm3(gv);
        for (int k = 0; k < 3; ++k) pr->p[k] = nc.pos[k] + P.d * (gv[k] / ng);
    }
    pr->type = P_LEAF_ELONG;
    if (!(sd > P.phi)) return;
    // bifurcation candidate, greenhouse.py:191-239
    pr->type = (faz == 0) ? P_LEAF_BIF : P_LEAF_DRAW;
    if (faz != 0) {
        pr->ratio5 = pow(nc.dist_to_center / (2 * faz), 5.0);
        pr->cond = angle_between_two(nc.vtc, avg) > 90;
    }
    double phi1, phi2;
    murray_angles(P.r, P.r, P.kappa, &phi1, &phi2);
    double c[3] = {0, 0, 0};
    for (int q = 0; q < m; ++q) {
        const int a = sidx[q];
        if (q == 0) { c[0] = sx[a]; c[1] = sy[a]; c[2] = sz[a]; }
        else { c[0] += sx[a]; c[1] += sy[a]; c[2] += sz[a]; }
    }
    for (int k = 0; k < 3; ++k) c[k] /= (double)m;
    double dpc[3] = {c[0] - nc.pos[0], c[1] - nc.pos[1], c[2] - nc.pos[2]};
    if (norm3(dpc) != 0.0) { const double n2 = norm3(dpc); for (int k = 0; k < 3; ++k) dpc[k] /= n2; }
    // np.cov((atts - c).T): row means (sequential), centred rows, syrk-style running-fma products, * 1/(m-1)
    double av[3] = {0, 0, 0};
    for (int q = 0; q < m; ++q) {
        const int a = sidx[q];
        const double x[3] = {sx[a] - c[0], sy[a] - c[1], sz[a] - c[2]};
        if (q == 0) { av[0] = x[0]; av[1] = x[1]; av[2] = x[2]; } else { av[0] += x[0]; av[1] += x[1]; av[2] += x[2]; }
    }
    for (int k = 0; k < 3; ++k) av[k] /= (double)m;
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = 0; q < m; ++q) {
        const int a = sidx[q];
        const double x[3] = {(sx[a] - c[0]) - av[0], (sy[a] - c[1]) - av[1], (sz[a] - c[2]) - av[2]};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) cov[3 * i + j] = fma(x[i], x[j], cov[3 * i + j]);
    }
    const double fact = 1.0 / (double)(m - 1);
    for (int i = 0; i < 9; ++i) cov[i] *= fact;
    double dl[3];
    const int est = eig3::principal_axis(cov, dl);
    if (est != 0) D.err[g] = 10 + est;
    const double c1 = cos(DEG2RAD * phi1), s1 = sin(DEG2RAD * phi1), c2 = cos(DEG2RAD * phi2), s2 = sin(DEG2RAD * phi2);
    double g1[3], g2[3];
    for (int k = 0; k < 3; ++k) { g1[k] = c1 * dpc[k] + s1 * dl[k]; g2[k] = c2 * dpc[k] - s2 * dl[k]; }
    const double n1 = norm3(g1), n2 = norm3(g2);
    for (int k = 0; k < 3; ++k) { pr->b1[k] = nc.pos[k] + g1[k] / n1 * P.d; pr->b2[k] = nc.pos[k] + g2[k] / n2 * P.d; }
}

// inter-node branch, greenhouse.py:259-306, for a given distal radius r1
// `cache_mode`: 1 = store the radius-independent per-attractor terms (two angles, unit vector) for a later
// re-evaluation, 2 = re-evaluate from that cache (k_commit, when the distal radius changed inside the call), 0 = neither.
__device__ void eval_inter(const GrowDev& D, const GrowShape& S, const IterP& P, int g, int f, const NodeCtx& nc,
                           const int* lst, int n, double r1, double c_used, Proposal* pr, int cache_mode) {
    const size_t sb = (size_t)g * S.capS;
    const double* sx = D.sx[f] + sb; const double* sy = D.sy[f] + sb; const double* sz = D.sz[f] + sb;
    double* cache = D.sc_inter + (sb + (lst - (D.list + sb))) * 5;
    pr->type = P_INTER_EMPTY;
    pr->c_used = c_used;
    double phi1, phi2;
    murray_angles(r1, P.r, P.kappa, &phi1, &phi2);
    const double dseg[3] = {nc.ch[0] - nc.pos[0], nc.ch[1] - nc.pos[1], nc.ch[2] - nc.pos[2]};
    const double pseg[3] = {nc.pos[0] - nc.par[0], nc.pos[1] - nc.par[1], nc.pos[2] - nc.par[2]};
    const double nd_ = norm3(dseg), np_ = norm3(pseg);
    const double g2 = P.gamma[f] / 2;
    int m = 0;
    double avg[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        double ad, ap, un[3];
        if (cache_mode == 2) {
            ad = cache[5 * i]; ap = cache[5 * i + 1]; un[0] = cache[5 * i + 2]; un[1] = cache[5 * i + 3]; un[2] = cache[5 * i + 4];
        } else {
            const int a = lst[i];
            const double w[3] = {sx[a] - nc.pos[0], sy[a] - nc.pos[1], sz[a] - nc.pos[2]};
            ad = angle_to(dseg, nd_, w, n);
            ap = angle_to(pseg, np_, w, n);
            const double nw = norm3(w);
            un[0] = w[0] / nw; un[1] = w[1] / nw; un[2] = w[2] / nw;
            if (cache_mode == 1) { cache[5 * i] = ad; cache[5 * i + 1] = ap; cache[5 * i + 2] = un[0]; cache[5 * i + 3] = un[1]; cache[5 * i + 4] = un[2]; }
        }
        if ((phi1 + phi2 - g2 <= ad) && (ad <= (phi1 + phi2 + g2)) && (ap <= phi2 + g2)) {
            if (m == 0) { avg[0] = un[0]; avg[1] = un[1]; avg[2] = un[2]; }
            else { avg[0] += un[0]; avg[1] += un[1]; avg[2] += un[2]; }
            ++m;
        }
    }
    if (m == 0) return;
    const double dv[3] = {dseg[0] / nd_, dseg[1] / nd_, dseg[2] / nd_};
    const double cr[3] = {dv[1] * avg[2] - dv[2] * avg[1], dv[2] * avg[0] - dv[0] * avg[2], dv[0] * avg[1] - dv[1] * avg[0]};
    if (cr[0] == 0 && cr[1] == 0 && cr[2] == 0) return;
    pr->type = P_INTER_DRAW;
    pr->ratio5 = pow(nc.dist_to_center / (2 * D.faz_radius[g]), 5.0);
    pr->cond = angle_between_two(nc.vtc, avg) <= 90;
    const double ncr = norm3(cr);
    const double ax[3] = {cr[0] / ncr, cr[1] / ncr, cr[2] / ncr};
    const double ct = cos(DEG2RAD * phi2), sth = sin(DEG2RAD * phi2);
    const double kxv[3] = {ax[1] * dv[2] - ax[2] * dv[1], ax[2] * dv[0] - ax[0] * dv[2], ax[0] * dv[1] - ax[1] * dv[0]};
    const double kdv = ddot3(ax, dv);
    double vv[3];
    for (int k = 0; k < 3; ++k) vv[k] = (dv[k] * ct + kxv[k] * sth) + ax[k] * kdv * (1 - ct);
    const double nvv = norm3(vv), na = norm3(avg);
    double gv[3];
    for (int k = 0; k < 3; ++k) gv[k] = P.omega * (vv[k] / nvv) + (1 - P.omega) * (avg[k] / na);
    const double ng = norm3(gv);
    for (int k = 0; k < 3; ++k) pr->p[k] = nc.pos[k] + P.d * (gv[k] / ng);
}

__device__ __forceinline__ void load_ctx(const GrowDev& D, const GrowShape& S, const IterP& P, int g, int f, int nd, NodeCtx* nc) {
    const size_t nb = (size_t)g * S.capN;
    nc->pos[0] = D.nx[f][nb + nd]; nc->pos[1] = D.ny[f][nb + nd]; nc->pos[2] = D.nz[f][nb + nd];
    nc->parent = D.npar[f][nb + nd];
    nc->nch = D.nnch[f][nb + nd];
    if (nc->parent >= 0) { const int p = nc->parent; nc->par[0] = D.nx[f][nb + p]; nc->par[1] = D.ny[f][nb + p]; nc->par[2] = D.nz[f][nb + p]; }
    if (nc->nch >= 1) { const int c = D.nch0[f][nb + nd]; nc->ch[0] = D.nx[f][nb + c]; nc->ch[1] = D.ny[f][nb + c]; nc->ch[2] = D.nz[f][nb + c]; }
    nc->vtc[0] = P.faz_cx - nc->pos[0]; nc->vtc[1] = P.faz_cy - nc->pos[1];
    nc->dist_to_center = norm2(nc->vtc);
}

__global__ void __launch_bounds__(128) k_eval(int dslot, GrowShape S, IterP P, int f) {
    const GrowDev& D = c_dev[dslot];
    const int g = blockIdx.y;
    if (D.err[g]) return;
    const int nd_ = D.n_dict[g];
    const size_t sb = (size_t)g * S.capS, nb = (size_t)g * S.capN;
    const int* loff = D.list_off + (size_t)g * (S.capN + 1);
    // Leaf entries (elongate / bifurcate: covariance + eigenvector) and inter-node entries (sprout) run different code of
    // thousands of float64 instructions each; in dict order they are mixed, so every warp would run both.  The entries of a chunk
    // of 128 are dealt to the thread slots by kind instead: leaves from slot 0, inter-nodes from the next multiple of 32 -- a warp
    // sees one kind.  (An entry's proposal depends on the entry alone: which thread computes it does not matter.)
    __shared__ unsigned char s_perm[256];
    __shared__ int s_wcnt[2][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int chunk = blockIdx.x * 128; chunk < nd_; chunk += gridDim.x * 128) {          // (block-uniform)
        const int e0 = chunk + tid;
        int kind = 2;                                                                      // 0 leaf, 1 inter-node, 2 none
        if (e0 < nd_) kind = D.nnch[f][nb + D.dict_node[nb + e0]] == 0 ? 0 : 1;
        const unsigned m0 = __ballot_sync(0xffffffffu, kind == 0), m1 = __ballot_sync(0xffffffffu, kind == 1);
        if (lane == 0) { s_wcnt[0][warp] = __popc(m0); s_wcnt[1][warp] = __popc(m1); }
        __syncthreads();
        int b0 = 0, b1 = 0, n0 = 0, n1 = 0;
        for (int w = 0; w < 4; ++w) { if (w < warp) { b0 += s_wcnt[0][w]; b1 += s_wcnt[1][w]; } n0 += s_wcnt[0][w]; n1 += s_wcnt[1][w]; }
        const int inter0 = (n0 + 31) & ~31;
        const unsigned lt = (1u << lane) - 1u;
        if (kind == 0) s_perm[b0 + __popc(m0 & lt)] = (unsigned char)tid;
        else if (kind == 1) s_perm[inter0 + b1 + __popc(m1 & lt)] = (unsigned char)tid;
        __syncthreads();
      for (int slot = tid; slot < inter0 + n1; slot += 128) {
        if (slot >= n0 && slot < inter0) continue;
        const int e = chunk + s_perm[slot];
        const int nd = D.dict_node[nb + e];
        Proposal pr;
        pr.type = P_NONE; pr.cond = 0; pr.ratio5 = 0; pr.c_used = 0;
        NodeCtx nc;
        load_ctx(D, S, P, g, f, nd, &nc);
        const int* lst = D.list + sb + loff[e];
        const int n = loff[e + 1] - loff[e];
        if (nc.nch == 0) eval_leaf(D, S, P, g, f, e, nc, lst, n, &pr);
        else if (nc.parent >= 0 && nc.nch == 1) {
            // distal radius from the child's Murray contribution (GrowDev::ncon)
            const double cv = D.ncon[f][nb + D.nch0[f][nb + nd]];
            const int m = D.nmeta[f][nb + nd];
            eval_inter(D, S, P, g, f, nc, lst, n, fast_pow(cv, 1.0 / P.kap_tab[m == 0xff ? 8 : (m >> 1)]), cv, &pr, 1);
        }
        D.prop[nb + e] = pr;
      }
        __syncthreads();                                                                   // s_perm / s_wcnt are reused by the next chunk
    }
}

// ------------------------------------------------------------------------------------------
// k_commit: one CTA per graph.
//   prologue (all threads)  mirrors the forest's topology and Murray contributions (GrowDev::ncon) in shared memory
//                           and compacts the dict entries the replay has to look at into 32-byte decision records;
//   replay   (thread 0)     walks them in dict order: Python-RNG draws, branch decisions, node ids, tree links --
//                           shared-memory traffic only, except for the re-evaluation of an inter-node;
//   epilogue (all threads)  writes the new nodes' SoA fields, refreshes the contributions still dirty bottom-up
//                           (a node is computed by whichever chain reaches it last) and writes them back.
// Radii inside a call are LAZY (arterial_tree.py:174-184 is order-insensitive up to rounding): a branch event only
// marks its ancestor chain dirty, stopping at the first already-dirty node.  An inter-node entry refreshes the dirty
// part of its distal subtree when its turn comes and is re-evaluated if the distal contribution differs from the one
// k_eval used.  With contributions instead of radii a refresh step is an addition; pow appears only across creation
// modes of different kappa.
// `SM` = the tree fits the shared-memory mirror (16-bit node ids); otherwise the same code runs on the global arrays.
// ------------------------------------------------------------------------------------------
template <bool SM>
struct TreeView {
    // SM: shared-memory mirror, 16-bit links (0xffff = none); !SM: the global SoA arrays themselves
    typename std::conditional<SM, unsigned short, int>::type *par, *c0, *c1;
    double* C;
    typename std::conditional<SM, unsigned short, int>::type *slot;   // dict rank of the inter-node entries of this call
    unsigned char* km;           // SM: creation-mode index; !SM: nmeta (decoded on access)
    unsigned int *dirty, *arr;   // by node: contribution is stale / one child of a bifurcation has arrived (epilogue)
    unsigned int *inter, *tag;   // by node: is an inter-node dict entry of this call; by dict rank: re-check pending
    int n_before;
    __device__ __forceinline__ int get_par(int n) const { if (SM) { const int v = par[n]; return v == 0xffff ? -1 : v; } return par[n]; }
    __device__ __forceinline__ int get_c0(int n) const { if (SM) { const int v = c0[n]; return v == 0xffff ? -1 : v; } return c0[n]; }
    __device__ __forceinline__ int get_c1(int n) const { if (SM) { const int v = c1[n]; return v == 0xffff ? -1 : v; } return c1[n]; }
    __device__ __forceinline__ void set_c0(int n, int v) { c0[n] = v; }
    __device__ __forceinline__ void set_c1(int n, int v) { c1[n] = v; }
    __device__ __forceinline__ int kmode(int n) const { if (SM) return km[n]; const int m = km[n]; return m == 0xff ? 8 : (m >> 1); }
    __device__ __forceinline__ bool is_dirty(int n) const { return n < n_before && ((dirty[n >> 5] >> (n & 31)) & 1u); }
    __device__ __forceinline__ void set_dirty(int n) { dirty[n >> 5] |= 1u << (n & 31); }
    __device__ __forceinline__ void clr_dirty(int n) { dirty[n >> 5] &= ~(1u << (n & 31)); }
    __device__ __forceinline__ bool is_inter(int n) const { return (inter[n >> 5] >> (n & 31)) & 1u; }
    __device__ __forceinline__ bool is_tagged(int rk) const { return (tag[rk >> 5] >> (rk & 31)) & 1u; }
    __device__ __forceinline__ void set_tag(int rk) { tag[rk >> 5] |= 1u << (rk & 31); }
    __device__ __forceinline__ void clr_tag(int rk) { tag[rk >> 5] &= ~(1u << (rk & 31)); }
    __device__ __forceinline__ int next_tag(int from, int lim) const {      // first tagged rank in [from, lim), or -1
        int q = from;
        while (q < lim) {
            const unsigned int wv = tag[q >> 5] >> (q & 31);
            if (wv) { const int r = q + __ffs(wv) - 1; return r < lim ? r : -1; }
            q = (q | 31) + 1;
        }
        return -1;
    }
};

// contribution of node n (children c0, c1; at least one) to its parent; vol = read children through volatile loads
template <bool SM, bool VOL>
__device__ __forceinline__ double node_contribution(const TreeView<SM>& T, const IterP& P, int n, int c0, int c1) {
    const int kmn = T.kmode(n);
    auto cv = [&](int c) -> double {        // nodes created in this call are leaves of radius r
        if (c >= T.n_before) return P.leafc_tab[kmn];
        if (VOL) return cuda::atomic_ref<double, cuda::thread_scope_block>(T.C[c]).load(cuda::memory_order_relaxed);
        return T.C[c];
    };
    double s = cv(c0);
    if (c1 >= 0) s = s + cv(c1);
    const int p = T.get_par(n);
    const double kn = P.kap_tab[kmn], kp = P.kap_tab[T.kmode(p)];
    return kp == kn ? s : fast_pow(s, kp / kn);
}

template <bool SM>
__device__ void commit_body(const GrowDev& D, const GrowShape& S, const IterP& P, int f, int* s_dyn) {
    const int g = blockIdx.x, tid = threadIdx.x;
    const size_t sb = (size_t)g * S.capS, nb = (size_t)g * S.capN;
    const int nd_ = D.n_dict[g];
    Proposal* prop = D.prop + nb;
    ActDec* adec = D.adec + nb;      // (re-pointed to shared memory below when it fits)
    int4* newl = D.newl + nb;
    const int* dict = D.dict_node + nb;
    __shared__ int s_nnew, s_err, s_nstart, s_marked;
    const int n_before = D.n_nodes[f][g];
    const int nwords = (n_before + 31) >> 5;
    const int budget_words = S.commit_smem / 4;
    TreeView<SM> T;
    T.n_before = n_before;
    int off_words = 0;
    if (SM) {
        // layout: C [n] f64 | par, c0, c1, slot [n] u16 | km [n] u8 | dirty, arr, inter, tag [nwords] u32
        T.C = reinterpret_cast<double*>(s_dyn);
        unsigned short* h = reinterpret_cast<unsigned short*>(T.C + n_before);
        T.par = (decltype(T.par))h; T.c0 = (decltype(T.c0))(h + n_before); T.c1 = (decltype(T.c1))(h + 2 * n_before);
        T.slot = (decltype(T.slot))(h + 3 * n_before);
        T.km = reinterpret_cast<unsigned char*>(h + 4 * n_before);
        const size_t bytes = (((size_t)n_before * 17 + 3) & ~(size_t)3);
        T.dirty = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(s_dyn) + bytes);
        T.arr = T.dirty + nwords; T.inter = T.arr + nwords; T.tag = T.inter + nwords;
        off_words = (int)((bytes / 4 + 4 * nwords + 3) & ~(size_t)3);
    } else {
        T.C = D.ncon[f] + nb;
        T.par = (decltype(T.par))(D.npar[f] + nb); T.c0 = (decltype(T.c0))(D.nch0[f] + nb); T.c1 = (decltype(T.c1))(D.nch1[f] + nb);
        T.slot = (decltype(T.slot))(D.slot + nb);
        T.km = D.nmeta[f] + nb;
        const size_t cw = (S.capN + 31) / 32;
        T.dirty = D.cbits + (size_t)g * 4 * cw;
        T.arr = T.dirty + cw; T.inter = T.arr + cw; T.tag = T.inter + cw;
    }
    // the decision records and the Python-stream words of this call also live in shared memory when they fit
    const bool dec_smem = (size_t)nd_ * sizeof(ActDec) + (size_t)(2 * nd_ + 2) * 4 + 64 <= (size_t)(budget_words - off_words) * 4;
    ActDec* adec_s = reinterpret_cast<ActDec*>(s_dyn + off_words);
    unsigned int* pb_s = reinterpret_cast<unsigned int*>(adec_s + nd_);
    if (dec_smem) adec = adec_s;
    const long long t_start = clock64();
    if (SM) {
        const int* gpar = D.npar[f] + nb; const int* gc0 = D.nch0[f] + nb; const int* gc1 = D.nch1[f] + nb;
        const double* gC = D.ncon[f] + nb; const unsigned char* gm = D.nmeta[f] + nb;
#pragma unroll 4
        for (int i = tid; i < n_before; i += blockDim.x) {
            T.C[i] = gC[i];
            T.par[i] = (unsigned short)gpar[i]; T.c0[i] = (unsigned short)gc0[i]; T.c1[i] = (unsigned short)gc1[i];   // -1 -> 0xffff
            const int m = gm[i];
            T.km[i] = (unsigned char)(m == 0xff ? 8 : (m >> 1));
        }
    }
    for (int i = tid; i < nwords; i += blockDim.x) { T.dirty[i] = 0; T.arr[i] = 0; T.inter[i] = 0; T.tag[i] = 0; }
    if (tid == 0) { s_nstart = 0; s_marked = 0; }
    __syncthreads();
    for (int e = tid; e < nd_; e += blockDim.x) {
        const int t = prop[e].type;
        if (t == P_INTER_DRAW || t == P_INTER_EMPTY) {
            const int nd = dict[e];
            atomicOr(&T.inter[nd >> 5], 1u << (nd & 31));
            if (SM) T.slot[nd] = (unsigned short)e;
        }
    }
    const int ppos0 = D.py_pos[g];
    if (dec_smem) {
        const unsigned int* pbg = D.py_buf + (size_t)g * S.pycap + ppos0;
        for (int i = tid; i < 2 * nd_ + 2; i += blockDim.x) pb_s[i] = pbg[i];
    }
    // ---- prologue: decision records of the entries that act without a re-evaluation
    int na = 0;
    for (int base = 0; base < nd_; base += blockDim.x) {
        const int e = base + tid;
        int fl = 0, t = 0;
        if (e < nd_) { t = prop[e].type; fl = (t == P_LEAF_ELONG || t == P_LEAF_DRAW || t == P_LEAF_BIF || t == P_INTER_DRAW); }
        int total;
        const int incl = block_scan_incl(fl, &total);
        if (fl) { ActDec a; a.e = e; a.nd = dict[e]; a.type = t; a.cond = prop[e].cond; a.ratio5 = prop[e].ratio5; a.c_used = prop[e].c_used; adec[na + incl - 1] = a; }
        na += total;
    }
    __syncthreads();
    const long long t_replay = clock64();
    // ---- replay
    if (tid == 0) {
        unsigned char* DEACT = D.deact[f] + nb;
        const int* loff = D.list_off + (size_t)g * (S.capN + 1);
        // word stream of Python's `random`: shared-memory copy of this call's window, or the global buffer
        const unsigned int* pb = dec_smem ? pb_s - ppos0 : D.py_buf + (size_t)g * S.pycap;
        int ppos = ppos0;
        int n_nodes = n_before, nnew = 0, err = 0;
        const int id_limit = SM ? (S.capN < 0xfffe ? S.capN : 0xfffe) : S.capN;
        long long draws = 0;
        long long dbg_walk = 0, dbg_recheck = 0, dbg_events = 0, dbg_steps = 0, dbg_reevals = 0, dbg_entries = 0;
        auto next_uniform = [&]() { const double u = mt_double(pb[ppos], pb[ppos + 1]); ppos += 2; ++draws; return u; };
        auto add_node = [&](int e, int which, int parent, int parent_nch, int walk_after) -> bool {
            if (n_nodes >= id_limit) { err = 1; return false; }
            const int id = n_nodes++;
            if (parent_nch == 0) T.set_c0(parent, id); else T.set_c1(parent, id);
            newl[nnew++] = make_int4(e, which | (walk_after << 2), parent, id);
            return true;
        };
        int cur_rank = -1, outstanding = 0;
        auto mark_walk = [&](int n) {
            ++dbg_events;
            if (SM) {
                // software-pipelined: the parent link and the dirty word of the node under the cursor were loaded one
                // step ahead, so a step costs one shared-memory round trip instead of four dependent ones
                int p = T.par[n];
                unsigned int dw = T.dirty[n >> 5];
                while (p != 0xffff && !((dw >> (n & 31)) & 1u)) {          // (the root is never marked)
                    ++dbg_steps;
                    const int pp = T.par[p];
                    const unsigned int iw = T.inter[p >> 5];
                    T.dirty[n >> 5] = dw | (1u << (n & 31));
                    const unsigned int dwp = T.dirty[p >> 5];
                    if ((iw >> (p & 31)) & 1u) {                             // p's distal contribution changes: re-check it when its turn comes
                        const int rk = T.slot[p];
                        if (rk > cur_rank && !T.is_tagged(rk)) { T.set_tag(rk); ++outstanding; }
                    }
                    n = p; p = pp; dw = dwp;
                }
                return;
            }
            while (true) {
                ++dbg_steps;
                const int p = T.get_par(n);
                if (p < 0 || T.is_dirty(n)) return;          // (the root is never marked)
                T.set_dirty(n);
                if (T.is_inter(p)) {                         // p's distal contribution changes: re-check it when its turn comes
                    const int rk = T.slot[p];
                    if (rk > cur_rank && !T.is_tagged(rk)) { T.set_tag(rk); ++outstanding; }
                }
                n = p;
            }
        };
        auto refresh_subtree = [&](int top) {     // post-order over the dirty part of subtree(top); no stack needed
            int n = top;
            while (true) {
                const int c0 = T.get_c0(n), c1 = T.get_c1(n);
                if (c0 >= 0 && T.is_dirty(c0)) { n = c0; continue; }
                if (c1 >= 0 && T.is_dirty(c1)) { n = c1; continue; }
                T.C[n] = node_contribution<SM, false>(T, P, n, c0, c1);
                T.clr_dirty(n);
                if (n == top) return;
                n = T.get_par(n);
            }
        };
        int ai = 0, scan = 0;
        while (!err) {
            // next entry in dict order: the next decision record, unless an entry tagged for a re-check comes first.
            // Tags always point past cur_rank and `scan` moves monotonically: each dict slot is inspected at most once.
            const int ea = (ai < na) ? adec[ai].e : 0x7fffffff;
            int e = -1;
            if (outstanding > 0) {
                if (scan <= cur_rank) scan = cur_rank + 1;
                const int lim = ea < nd_ ? ea : nd_;
                e = T.next_tag(scan, lim);
                if (e < 0) scan = lim;
            }
            ActDec a;
            bool tagged;
            if (e >= 0) {                       // tagged entry that is not (or not yet) a decision record
                tagged = true; --outstanding; T.clr_tag(e); scan = e + 1;
                a.e = e; a.nd = dict[e]; a.type = prop[e].type; a.cond = prop[e].cond; a.ratio5 = prop[e].ratio5; a.c_used = prop[e].c_used;
            } else if (ea != 0x7fffffff) {
                a = adec[ai++];
                e = a.e;
                tagged = outstanding > 0 && T.is_tagged(e);
                if (tagged) { --outstanding; T.clr_tag(e); }
            } else {
                break;
            }
            ++dbg_entries;
            cur_rank = e;
            const int nd = a.nd;
            if (a.type == P_INTER_DRAW || a.type == P_INTER_EMPTY) {
                if (tagged) {
                    const long long t0 = clock64();
                    // an earlier entry of this call branched below this node: bring its distal contribution up to date
                    const int cd = T.get_c0(nd);
                    if (T.is_dirty(cd)) refresh_subtree(cd);
                    const double cv = T.C[cd];
                    if (cv != a.c_used) {
                        ++dbg_reevals;
                        NodeCtx nc;
                        load_ctx(D, S, P, g, f, nd, &nc);
                        Proposal pr;
                        pr.cond = 0; pr.ratio5 = 0;
                        eval_inter(D, S, P, g, f, nc, D.list + sb + loff[e], loff[e + 1] - loff[e], fast_pow(cv, 1.0 / P.kap_tab[T.kmode(nd)]), cv, &pr, 2);
                        prop[e] = pr;
                        a.type = pr.type; a.cond = pr.cond; a.ratio5 = pr.ratio5;
                    }
                    dbg_recheck += clock64() - t0;
                }
                if (a.type != P_INTER_DRAW) continue;
                const double u = next_uniform();
                if (a.ratio5 <= u && a.cond) continue;
                if (!add_node(e, 0, nd, 1, 1)) break;
                { const long long t1 = clock64(); mark_walk(nd); dbg_walk += clock64() - t1; }
                DEACT[nd] = 1;
            } else if (a.type == P_LEAF_ELONG) {
                if (!add_node(e, 0, nd, 0, 0)) break;
            } else if (a.type == P_LEAF_DRAW || a.type == P_LEAF_BIF) {
                bool bif = true;
                if (a.type == P_LEAF_DRAW) { const double u = next_uniform(); bif = (a.ratio5 > u) && a.cond; }
                if (bif) {
                    if (!add_node(e, 1, nd, 0, 0)) break;
                    if (!add_node(e, 2, nd, 1, 1)) break;
                    { const long long t1 = clock64(); mark_walk(nd); dbg_walk += clock64() - t1; }
                    DEACT[nd] = 1;
                } else {
                    if (!add_node(e, 0, nd, 0, 0)) break;
                }
            }
        }
        if (D.dbg) { long long* q = D.dbg + g * 8; q[1] += dbg_walk; q[2] += dbg_recheck; q[3] += dbg_entries; q[4] += dbg_events; q[5] += dbg_steps; q[6] += dbg_reevals; q[7] += na; }
        if (err) D.err[g] = err;
        D.py_pos[g] = ppos;
        D.py_draws[g] += draws;
        D.n_prev[f][g] = n_before;
        D.n_nodes[f][g] = n_nodes;
        s_nnew = nnew;
        s_err = err;
        s_marked = dbg_events > 0;
    }
    __syncthreads();
    if (s_err) return;
    const long long t_epi = clock64();

    // ---- epilogue 1: SoA fields of the new nodes and of their parents' links
    for (int k = tid; k < s_nnew; k += blockDim.x) {
        const int4 nn = newl[k];
        const int e = nn.x, which = nn.y & 3, walk = nn.y >> 2, parent = nn.z, id = nn.w;
        const Proposal& pr = prop[e];
        const double* p = which == 0 ? pr.p : (which == 1 ? pr.b1 : pr.b2);
        D.nx[f][nb + id] = p[0]; D.ny[f][nb + id] = p[1]; D.nz[f][nb + id] = p[2];
        D.ncon[f][nb + id] = P.leafc_tab[T.kmode(parent)];
        D.npar[f][nb + id] = parent; D.nch0[f][nb + id] = -1; D.nch1[f][nb + id] = -1; D.nnch[f][nb + id] = 0;
        D.deact[f][nb + id] = 0;
        D.nmeta[f][nb + id] = (unsigned char)((P.mode_idx << 1) | walk);
        const int c0 = T.get_c0(parent), c1 = T.get_c1(parent);
        D.nch0[f][nb + parent] = c0; D.nch1[f][nb + parent] = c1; D.nnch[f][nb + parent] = (unsigned char)((c0 >= 0) + (c1 >= 0));
    }
    // ---- epilogue 2: bottom-up refresh of every contribution still dirty.  A chain starts at a dirty node without
    // dirty children and climbs; at a bifurcation the first child to arrive stops and the second one (or the only
    // dirty one) adds both contributions and goes on.
    if (s_marked) {
        int* starters = D.alist + nb;
        for (int w = tid; w < nwords; w += blockDim.x) {
            unsigned int bits = T.dirty[w];
            while (bits) {
                const int n = (w << 5) + __ffs(bits) - 1;
                bits &= bits - 1;
                const int c0 = T.get_c0(n), c1 = T.get_c1(n);
                const bool d0 = c0 >= 0 && T.is_dirty(c0), d1 = c1 >= 0 && T.is_dirty(c1);
                if (!d0 && !d1) starters[atomicAdd(&s_nstart, 1)] = n;
                else if (c1 >= 0 && (d0 != d1)) atomicOr(&T.arr[n >> 5], 1u << (n & 31));
            }
        }
        __syncthreads();
        const int nstart = s_nstart;
        // Chains of different threads meet at bifurcations: the first child to arrive publishes its contribution and stops
        // (release), the second one takes over both (acquire).  Contributions and the arrival bitmap are accessed through
        // block-scope atomics, so the hand-off is ordered by the memory model itself (compute-sanitizer racecheck agrees).
        for (int k = tid; k < nstart; k += blockDim.x) {
            int n = starters[k];
            double val = node_contribution<SM, true>(T, P, n, T.get_c0(n), T.get_c1(n));
            cuda::atomic_ref<double, cuda::thread_scope_block>(T.C[n]).store(val, cuda::memory_order_relaxed);
            while (true) {
                const int p = T.get_par(n);
                if (p < 0 || !T.is_dirty(p)) break;                  // (the root is never marked)
                const unsigned int pbit = 1u << (p & 31);
                if (T.get_c1(p) >= 0) {
                    cuda::atomic_ref<unsigned int, cuda::thread_scope_block> arrived(T.arr[p >> 5]);
                    if (!(arrived.fetch_or(pbit, cuda::memory_order_acq_rel) & pbit)) break;   // first arrival: the sibling's chain continues
                }
                val = node_contribution<SM, true>(T, P, p, T.get_c0(p), T.get_c1(p));
                cuda::atomic_ref<double, cuda::thread_scope_block>(T.C[p]).store(val, cuda::memory_order_relaxed);
                n = p;
            }
        }
        __syncthreads();
        if (SM) {
            double* gC = D.ncon[f] + nb;
            for (int i = tid; i < n_before; i += blockDim.x) gC[i] = T.C[i];
        }
    }
    const long long t_act = clock64();
    if (tid == 0) {
        const long long t_end = clock64();          // per-phase cycle counters (reported through OctaGrowStats)
        D.counters[g * 8 + 4] += t_replay - t_start; D.counters[g * 8 + 5] += t_epi - t_replay;
        D.counters[g * 8 + 6] += t_act - t_epi; D.counters[g * 8 + 7] += t_end - t_act;
    }
}

__global__ void __launch_bounds__(512) k_commit(int dslot, GrowShape S, IterP P, int f) {
    const GrowDev& D = c_dev[dslot];
    extern __shared__ __align__(16) int s_dyn[];
    const int g = blockIdx.x;
    if (D.err[g]) return;
    const int n_before = D.n_nodes[f][g];
    const size_t need = (((size_t)n_before * 17 + 3) & ~(size_t)3) + 16 * (size_t)((n_before + 31) >> 5) + 64;
    // room for every node this call can add must remain within 16-bit ids
    if (need <= (size_t)S.commit_smem && n_before + 2 * D.n_dict[g] < 0xfffe) commit_body<true>(D, S, P, f, s_dyn);
    else commit_body<false>(D, S, P, f, s_dyn);
}

// ------------------------------------------------------------------------------------------
// k_kill: one CTA per graph.  f = 0: satisfied O2 sinks -> CO2 sources (set order); f = 1: CO2 removal
// ------------------------------------------------------------------------------------------
// Stable removal of every hit from the sink list f (element_mesh.py:195-211) + per-iteration trace
// H >= 0: the removed positions are hl[0..H) in ascending order (H <= KILL_RCAP; staged in `rs`, shared memory): the new place of
// an element is its position minus the number of removed positions before it (binary search), ONE barrier per 4 * blockDim
// elements instead of a block scan (three barriers) per blockDim elements, and nothing before the first removed position moves.
// H < 0: positions with hitj >= 0 are removed, found with block scans (fallback).
constexpr int KILL_RCAP = 4096;
__device__ void kill_compact(const GrowDev& D, const GrowShape& S, const IterP& P, int f, int g, int Sn, bool removed, int H = -1,
                             const int* hl = nullptr, int* rs = nullptr) {
    const int tid = threadIdx.x;
    const size_t sb = (size_t)g * S.capS;
    double* sx = D.sx[f] + sb; double* sy = D.sy[f] + sb; double* sz = D.sz[f] + sb;
    const int* hitj = D.hitj + sb;
    if (removed && H >= 0) {
        __syncthreads();
        for (int k = tid; k < H; k += blockDim.x) rs[k] = hl[k];
        __syncthreads();
        if (H > 0) {
            const int first = rs[0];
            const int CH = 4 * blockDim.x;
            for (int base = first; base < Sn; base += CH) {
                double px[4], py[4], pz[4];
                int dst[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = base + q * blockDim.x + tid;
                    dst[q] = -1;
                    if (i < Sn) {
                        int lo = 0, hi = H;                     // removed positions < i
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (rs[mid] < i) lo = mid + 1; else hi = mid; }
                        if (!(lo < H && rs[lo] == i)) { dst[q] = i - lo; px[q] = sx[i]; py[q] = sy[i]; pz[q] = sz[i]; }
                    }
                }
                __syncthreads();                                // every read of this chunk precedes its writes
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (dst[q] >= 0) { sx[dst[q]] = px[q]; sy[dst[q]] = py[q]; sz[dst[q]] = pz[q]; }
            }
            if (tid == 0) D.n_s[f][g] = Sn - H;
        }
    } else if (removed) {
        int w = 0;
        for (int base = 0; base < Sn; base += blockDim.x) {
            const int i = base + tid;
            int keep = 0;
            double px = 0, py = 0, pz = 0;
            if (i < Sn) { keep = hitj[i] < 0; px = sx[i]; py = sy[i]; pz = sz[i]; }
            int total;
            const int incl = block_scan_incl(keep, &total);
            if (keep) { const int o = w + incl - 1; sx[o] = px; sy[o] = py; sz[o] = pz; }
            w += total;
            __syncthreads();
        }
        if (tid == 0) D.n_s[f][g] = w;
    }
    __syncthreads();
    if (tid == 0 && D.trace && P.iter < 4096) {       // (the sampler of the next iteration may already run beside the venous phase)
        int* tr = D.trace + ((size_t)g * 4096 + P.iter) * 4;
        if (f == 0) { tr[0] = D.n_nodes[0][g]; tr[1] = D.n_s[0][g]; }
        else { tr[2] = D.n_nodes[1][g]; tr[3] = D.n_s[1][g]; }
    }
}

// O2 -> CO2 conversion of the T sinks ta[0..T) (hit, not vetoed): insertion sequence by (first new node that hits it, rank inside
// the ball) -- the order of `for node in new_nodes: for oxy in ball(node)` --, CPython set emulation, append in slot order.
// kdrank == nullptr: list-index order inside a ball.  Returns true (on every thread) if DETECT found the result order-sensitive;
// nothing has been appended then.
template <bool DETECT>
__device__ bool kill_convert(const GrowDev& D, const GrowShape& S, int g, int T, const int* kdrank, KillShared* ks) {
    const int tid = threadIdx.x;
    const size_t sb = (size_t)g * S.capS;
    const double* sx = D.sx[0] + sb; const double* sy = D.sy[0] + sb; const double* sz = D.sz[0] + sb;
    const int* hitj = D.hitj + sb;
    const int* ta = D.ta + sb;
    int* seq = D.seq + sb;
    // order test of pyset_run_multi: hashes and ball ids of the sequence staged in the (now idle) node-coordinate arrays, its
    // tables in the space of the hash tables
    const bool multi = DETECT && T <= MS_MAXT;
    long long* ms_hs = reinterpret_cast<long long*>(ks->nxs);
    int* ms_bl = reinterpret_cast<int*>(ks->nys);
    static_assert(sizeof(ks->nxs) >= sizeof(long long) * MS_MAXT && sizeof(ks->nys) >= sizeof(int) * MS_MAXT, "staging space");
    static_assert(sizeof(ks->th) >= sizeof(short) * MS_TABLES * MS_TBL && SM_TBL >= MS_TBL, "table space");
    for (int q = tid; q < T; q += blockDim.x) {
        const int i = ta[q];
        const int ji = hitj[i];
        const int ki = kdrank ? kdrank[i] : i;
        int rank = 0;
        for (int q2 = 0; q2 < T; ++q2) {
            const int i2 = ta[q2];
            const int j2 = hitj[i2];
            const int k2 = kdrank ? kdrank[i2] : i2;
            rank += (j2 < ji) || (j2 == ji && k2 < ki);
        }
        seq[rank] = i;
        const long long h = py_hash_tuple3(sx[i], sy[i], sz[i]);      // CPython hash of the sink tuple, in parallel
        D.seqhash[sb + rank] = h;
        if (multi) { ms_hs[rank] = h; ms_bl[rank] = ji; }
    }
    for (int i = tid; i < 8; i += blockDim.x) { ks->tk[0][i] = -1; ks->th[0][i] = 0; }
    __syncthreads();
    long long* gth = D.set_hash + (size_t)g * 2 * SET_TBL;
    int* gtk = D.set_key + (size_t)g * 2 * SET_TBL;
    if (tid == 0 && multi) {
        int mask = 7;
        const int why = pyset_run_multi(ms_hs, ms_bl, T, reinterpret_cast<short*>(ks->th), &mask);
        if (!why) {
            const short* tab = reinterpret_cast<const short*>(ks->th);
            for (int z = 0; z <= mask; ++z) ks->tk[0][z] = tab[z] >= 0 ? seq[tab[z]] : -1;
        }
        ks->tabinfo[0] = 0; ks->tabinfo[1] = mask; ks->tabinfo[2] = 0; ks->tabinfo[3] = why ? 1 : 0;
    } else if (tid == 0) {
        PySetDev ps;
        ps.sh = ks; ps.gth = gth; ps.gtk = gtk;
        ps.init();
        const bool flagged = pyset_run<DETECT>(ps, seq, D.seqhash + sb, T, hitj, gtk);
        ks->tabinfo[0] = ps.cur; ks->tabinfo[1] = (int)ps.mask; ks->tabinfo[2] = ps.err; ks->tabinfo[3] = flagged ? 1 : 0;
        if (ps.err) D.err[g] = ps.err;
    }
    __syncthreads();
    if (ks->tabinfo[3]) return true;
    // append to the CO2 list in slot order (all threads: stable compaction of the occupied slots)
    if (!ks->tabinfo[2] && T > 0) {
        const int cur = ks->tabinfo[0], nslots = ks->tabinfo[1] + 1;
        const int* curk = cur < 2 ? ks->tk[cur] : gtk + (size_t)(cur - 2) * SET_TBL;
        int nco2 = D.n_s[1][g];
        for (int base = 0; base < nslots; base += blockDim.x) {
            const int zslot = base + tid;
            const int key = zslot < nslots ? curk[zslot] : -1;
            int total;
            const int incl = block_scan_incl(key >= 0, &total);
            if (key >= 0) {
                const int o = nco2 + incl - 1;
                if (o < S.capS) { D.sx[1][sb + o] = sx[key]; D.sy[1][sb + o] = sy[key]; D.sz[1][sb + o] = sz[key]; }
            }
            nco2 += total;
        }
        if (tid == 0) { if (nco2 > S.capS) { D.err[g] = 2; nco2 = S.capS; } D.n_s[1][g] = nco2; }
    }
    __syncthreads();
    return false;
}

// S.exact_ball_order: 0 = list-index order inside a ball (diagnostics), 1 = cKDTree order, permutation built for every graph in
// every iteration (k_kdbuild on the side stream), 2 = cKDTree order ON DEMAND: this kernel runs the conversion in list-index
// order with the order-sensitivity test of pyset_run; a graph whose result could depend on the order is put on the iteration's
// work list instead of being finished, k_kdbuild_list builds the permutation of the listed graphs only and finishes their kill.
__global__ void __launch_bounds__(1024) k_kill(int dslot, GrowShape S, IterP P, int f) {
    const GrowDev& D = c_dev[dslot];
    __shared__ KillShared ks;
    const int g = blockIdx.x, tid = threadIdx.x;
    if (D.err[g]) return;
    const size_t sb = (size_t)g * S.capS, nb = (size_t)g * S.capN;
    const int n0 = D.n_prev[f][g], n1 = D.n_nodes[f][g];
    const int nn = n1 - n0;
    const int Sn = D.n_s[f][g];
    double* sx = D.sx[f] + sb; double* sy = D.sy[f] + sb; double* sz = D.sz[f] + sb;
    int* hitj = D.hitj + sb;
    const double epsk2 = P.eps_k * P.eps_k;
    const size_t gcap = S.capN > S.capS ? S.capN : S.capS;
    // positions hit by this call, in arrival order (the hash-table space is idle until kill_convert), then ranked into `hl`
    __shared__ int s_nh;
    int* s_hl = reinterpret_cast<int*>(ks.th);
    static_assert(sizeof(ks.th) >= sizeof(int) * KILL_RCAP, "hit list space");
    int* hl = D.hl + sb;
    int H = -1;                      // hits in `hl` (sorted), -1: too many for the list, the scans decide
    int Hc = -1;                     // what kill_compact gets: H if `hl` lists every removed position, else -1
    if (tid == 0) s_nh = 0;
    __syncthreads();
    if (nn > 0) {
      if (f == 0) {
        // Ball test through the bucket grid of the O2 sinks (built by k_prepare of this iteration BEFORE the sampler appended its
        // accepted sinks: the list has only grown since, so the grid's indices are still list positions); the sinks appended
        // after the build are scanned directly.  hitj = FIRST new node that hits a sink = atomicMin over the hitting nodes --
        // the same value the sink-major scan below finds, with the same float64 distance expression.
        for (int i = tid; i < Sn; i += blockDim.x) hitj[i] = -1;
        __syncthreads();
        const double* gx = D.gx[1] + g * gcap; const double* gy = D.gy[1] + g * gcap; const double* gz = D.gz[1] + g * gcap;
        const int* gi = D.gi[1] + g * gcap;
        const int* cs = D.gcell[1] + (size_t)g * (GRID * GRID + 1);
        const int n_grid = cs[GRID * GRID];                       // sinks the grid knows (<= Sn)
        const double er = P.eps_k * (1.0 + 1e-9) + 1e-12;         // the cell range may only be too wide, never too narrow
        const int rows = 2 * ((int)ceil(er * (double)GRID) + 1) + 1;
        unsigned int* uhit = reinterpret_cast<unsigned int*>(hitj);
        for (int item = tid; item < nn * rows; item += blockDim.x) {
            const int j = item / rows, r = item - j * rows;
            const double qx = D.nx[0][nb + n0 + j], qy = D.ny[0][nb + n0 + j], qz = D.nz[0][nb + n0 + j];
            const int y0 = grid_cell(qy - er), y1 = grid_cell(qy + er);
            const int cy = y0 + r;
            if (cy > y1) continue;
            const int beg = cs[cy * GRID + grid_cell(qx - er)], end = cs[cy * GRID + grid_cell(qx + er) + 1];
            for (int q = beg; q < end; ++q)
                if (dist2(gx[q], gy[q], gz[q], qx, qy, qz) <= epsk2) {                                      // cKDTree ball: d^2 <= r^2
                    const int idx = gi[q];
                    if (atomicMin(&uhit[idx], (unsigned int)j) == 0xffffffffu) {                             // first hit of this sink
                        const int slot = atomicAdd(&s_nh, 1);
                        if (slot < KILL_RCAP) s_hl[slot] = idx;
                    }
                }
        }
        for (int i = n_grid + tid; i < Sn; i += blockDim.x) {     // appended after the grid was built
            const double px = sx[i], py = sy[i], pz = sz[i];
            for (int j = 0; j < nn; ++j)
                if (dist2(px, py, pz, D.nx[0][nb + n0 + j], D.ny[0][nb + n0 + j], D.nz[0][nb + n0 + j]) <= epsk2) {
                    hitj[i] = j;
                    const int slot = atomicAdd(&s_nh, 1);
                    if (slot < KILL_RCAP) s_hl[slot] = i;
                    break;
                }
        }
      } else {
        for (int i = tid; i < Sn; i += blockDim.x) hitj[i] = -1;
        for (int base = 0; base < nn; base += 512) {
            const int cntn = nn - base < 512 ? nn - base : 512;
            __syncthreads();
            for (int j = tid; j < cntn; j += blockDim.x) { ks.nxs[j] = D.nx[f][nb + n0 + base + j]; ks.nys[j] = D.ny[f][nb + n0 + base + j]; ks.nzs[j] = D.nz[f][nb + n0 + base + j]; }
            __syncthreads();
            for (int i = tid; i < Sn; i += blockDim.x) {
                if (hitj[i] >= 0) continue;
                const double px = sx[i], py = sy[i], pz = sz[i];
                for (int j = 0; j < cntn; ++j)
                    if (dist2(px, py, pz, ks.nxs[j], ks.nys[j], ks.nzs[j]) <= epsk2) {   // cKDTree ball: d^2 <= r^2
                        hitj[i] = base + j;
                        const int slot = atomicAdd(&s_nh, 1);
                        if (slot < KILL_RCAP) s_hl[slot] = i;
                        break;
                    }
            }
        }
      }
        __syncthreads();
        if (s_nh <= S.kill_rcap) {   // hits in list order: rank of each position among the hit positions (tens of hits)
            H = s_nh;
            for (int k = tid; k < H; k += blockDim.x) {
                const int v = s_hl[k];
                int r = 0;
                for (int m = 0; m < H; ++m) r += s_hl[m] < v;
                hl[r] = v;
            }
        }
        __syncthreads();
        if (f == 0) {
            if (H < 0) {             // hits in list order by block scans
                H = 0;
                for (int base = 0; base < Sn; base += blockDim.x) {
                    const int i = base + tid;
                    const int fl = (i < Sn) ? (hitj[i] >= 0) : 0;
                    int total;
                    const int incl = block_scan_incl(fl, &total);
                    if (fl) hl[H + incl - 1] = i;
                    H += total;
                }
                __syncthreads();
                if (H > S.kill_rcap) H = -H - 1;    // (compaction by scans; the veto below uses the count)
            }
            const bool listed = H >= 0;
            if (!listed) H = -H - 1;
            Hc = listed ? H : -1;
            // veto: a venous node within eps_k (greenhouse.py:106-109).  Thread per venous node against the hits staged in
            // shared memory (a warp per hit walking all venous nodes was one dependent L2 round trip per 32 nodes).
            // Through the bucket grid of ALL venous nodes (k_grid_build after the venous commit): item = (hit, row of cells).
            unsigned char* veto = D.veto + sb;
            {
                const double* vgx = D.gx[4] + g * gcap; const double* vgy = D.gy[4] + g * gcap; const double* vgz = D.gz[4] + g * gcap;
                const int* vcs = D.gcell[4] + (size_t)g * (GRID * GRID + 1);
                const double er = P.eps_k * (1.0 + 1e-9) + 1e-12;
                const int rows = 2 * ((int)ceil(er * (double)GRID) + 1) + 1;
                for (int k = tid; k < H; k += blockDim.x) veto[k] = 0;
                __syncthreads();
                for (int item = tid; item < H * rows; item += blockDim.x) {
                    const int k = item / rows, r = item - k * rows;
                    const int i = hl[k];
                    const double px = sx[i], py = sy[i], pz = sz[i];
                    const int y0 = grid_cell(py - er), y1 = grid_cell(py + er);
                    const int cy = y0 + r;
                    if (cy > y1) continue;
                    const int beg = vcs[cy * GRID + grid_cell(px - er)], end = vcs[cy * GRID + grid_cell(px + er) + 1];
                    for (int q = beg; q < end; ++q)
                        if (within_sqrt(dist2(vgx[q], vgy[q], vgz[q], px, py, pz), P.eps_k, epsk2)) { veto[k] = 1; break; }
                }
            }
            __syncthreads();
            int T = 0;
            for (int base = 0; base < H; base += blockDim.x) {
                const int h = base + tid;
                const int fl = (h < H) ? !veto[h] : 0;
                int total;
                const int incl = block_scan_incl(fl, &total);
                if (fl) D.ta[sb + T + incl - 1] = hl[h];
                T += total;
            }
            __syncthreads();
            // Within one ball the reference receives the hits in cKDTree order (ascending position in tree.indices,
            // element_mesh.py:136-137).
            if (S.exact_ball_order == 2) {
                if (kill_convert<true>(D, S, g, T, nullptr, &ks)) {
                    if (tid == 0) {          // finished by k_kdbuild_list
                        D.kill_T[g] = T;
                        D.kill_H[g] = listed ? H : -1;
                        D.kd_flag[g] = 1;
                        D.kd_list[(P.iter & 1) * S.G + atomicAdd(&D.kd_nflag[P.iter & 1], 1)] = g;
                        if (D.dbg) D.dbg[g * 8 + 0] += 1;     // (diagnostics: graph-iterations that needed the exact order)
                    }
                    return;
                }
            } else {
                kill_convert<false>(D, S, g, T, S.exact_ball_order ? D.kd_rank + sb : nullptr, &ks);
            }
        }
    }
    kill_compact(D, S, P, f, g, Sn, nn > 0, f == 0 ? Hc : H, hl, s_hl);
}

// ------------------------------------------------------------------------------------------
// k_kdbuild: one CTA per graph; cKDTree index permutation of the O2 sink list as it stands after sampling
// (= the tree the reference queries in step 3, greenhouse.py:101-102).  Runs on a side stream, concurrently with
// the arterial growth kernels, which do not modify the sink list.
// ------------------------------------------------------------------------------------------
__device__ void kdbuild_graph(const GrowDev& D, const GrowShape& S, int smem_bytes, int g, char* s_kd, int* s_ws, double* s_wd) {
    const int tid = threadIdx.x;
    const size_t sb = (size_t)g * S.capS;
    const int Sn = D.n_s[0][g];
    int* rk = D.kd_rank + sb;
    // shared-memory resident build (12 bytes per sink) when the list fits, else / on bail-out the global-memory one
    if ((size_t)Sn * 12 + 64 <= (size_t)smem_bytes && Sn < 65536 &&
        kdsm::build_ranks_block(D.sx[0] + sb, D.sy[0] + sb, D.sz[0] + sb, Sn, rk, s_kd, D.kd_nodes + sb, D.kd_nodes + sb + S.capS / 2, s_ws, s_wd))
        return;
    __syncthreads();
    int* kidx = D.kd_idx + sb;
    kdpar::build_indices_block(D.sx[0] + sb, D.sy[0] + sb, D.sz[0] + sb, Sn, kidx, D.kd_posL + sb, D.kd_posR + sb,
                               D.kd_nodes + sb, D.kd_nodes + sb + S.capS / 2, s_ws, s_wd);
    __syncthreads();
    for (int i = tid; i < Sn; i += blockDim.x) rk[kidx[i]] = i;
}

__global__ void __launch_bounds__(1024) k_kdbuild(int dslot, GrowShape S, int smem_bytes) {
    const GrowDev& D = c_dev[dslot];
    extern __shared__ __align__(16) char s_kd[];
    __shared__ int s_ws[kdpar::WS_INTS];
    __shared__ double s_wd[kdpar::WD_DOUBLES];
    const int g = blockIdx.x;
    if (D.err[g]) return;
    kdbuild_graph(D, S, smem_bytes, g, s_kd, s_ws, s_wd);
}

// on-demand variant (S.exact_ball_order == 2): the CTAs walk the work list k_kill filled in this iteration
// The build of one graph keeps ONE SM's issue slots 46 % busy for ~350 us and sits on the critical path of the iteration, while
// a subtree of the kd tree is independent of its siblings: the split form runs the first KD_TOP_LEVELS levels in one CTA
// (k_kd_top), the subtrees below them in one CTA each (k_kd_sub: 2^KD_TOP_LEVELS CTAs per graph, each with its own shared memory
// and all 32 warps), and k_kd_fix finishes the kill (and redoes a graph by the one-CTA build if a partial build bailed out).
// Per graph the scratch arrays double as hand-over space: kd_posL = the permutation after the top levels (16-bit), kd_posR =
// {number of subtree roots or -1, 3 ints per root, ..., [13] = redo flag}.
constexpr int KD_TOP_LEVELS = 2;
constexpr int KD_SUBS = 1 << KD_TOP_LEVELS;

__global__ void __launch_bounds__(1024) k_kd_top(int dslot, GrowShape S, IterP P, int smem_bytes) {
    const GrowDev& D = c_dev[dslot];
    extern __shared__ __align__(16) char s_kd[];
    __shared__ int s_ws[kdpar::WS_INTS];
    __shared__ double s_wd[kdpar::WD_DOUBLES];
    const int parity = P.iter & 1;
    const int n = D.kd_nflag[parity];
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int g = D.kd_list[parity * S.G + q];
        const size_t sb = (size_t)g * S.capS;
        const int Sn = D.n_s[0][g];
        int* top = D.kd_posR + sb;
        if (threadIdx.x == 0) top[13] = 0;
        if ((size_t)Sn * 12 + 64 <= (size_t)smem_bytes && Sn < 65536)
            kdsm::build_top_block(D.sx[0] + sb, D.sy[0] + sb, D.sz[0] + sb, Sn, KD_TOP_LEVELS, D.kd_rank + sb,
                                  reinterpret_cast<unsigned short*>(D.kd_posL + sb), top, s_kd, D.kd_nodes + sb, D.kd_nodes + sb + S.capS / 2, s_ws, s_wd);
        else if (threadIdx.x == 0) top[0] = -1;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_kd_sub(int dslot, GrowShape S, IterP P, int smem_bytes) {
    const GrowDev& D = c_dev[dslot];
    extern __shared__ __align__(16) char s_kd[];
    __shared__ int s_ws[kdpar::WS_INTS];
    __shared__ double s_wd[kdpar::WD_DOUBLES];
    const int parity = P.iter & 1;
    const int n = D.kd_nflag[parity];
    const int sub = blockIdx.x;
    for (int q = blockIdx.y; q < n; q += gridDim.y) {
        const int g = D.kd_list[parity * S.G + q];
        const size_t sb = (size_t)g * S.capS;
        int* top = D.kd_posR + sb;
        const int cnt = top[0];
        if (cnt < 0 || sub >= cnt) continue;                 // (block-uniform)
        const int s = top[1 + 3 * sub], e = top[2 + 3 * sub];
        bool ok = (size_t)(e - s) * 12 + 64 <= (size_t)smem_bytes;
        if (ok) ok = kdsm::build_sub_block(D.sx[0] + sb, D.sy[0] + sb, D.sz[0] + sb, s, e, D.kd_rank + sb,
                                           reinterpret_cast<const unsigned short*>(D.kd_posL + sb), s_kd,
                                           D.kd_nodes + sb + (size_t)sub * (S.capS / KD_SUBS), D.kd_nodes + sb + (size_t)sub * (S.capS / KD_SUBS) + S.capS / (2 * KD_SUBS),
                                           s_ws, s_wd);
        if (!ok && threadIdx.x == 0) top[13] = 1;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_kd_fix(int dslot, GrowShape S, IterP P, int smem_bytes) {
    const GrowDev& D = c_dev[dslot];
    extern __shared__ __align__(16) char s_kd[];
    __shared__ int s_ws[kdpar::WS_INTS];
    __shared__ double s_wd[kdpar::WD_DOUBLES];
    const int parity = P.iter & 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) D.kd_nflag[parity ^ 1] = 0;      // the next iteration's list starts empty
    const int n = D.kd_nflag[parity];
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int g = D.kd_list[parity * S.G + q];
        const size_t sb = (size_t)g * S.capS;
        const int* top = D.kd_posR + sb;
        const bool redo = top[0] < 0 || top[13];            // (block-uniform) a partial build bailed out: the one-CTA build
        __syncthreads();                                    // (its scratch includes `top`: every thread has read the flags)
        if (redo) {
            kdbuild_graph(D, S, smem_bytes, g, s_kd, s_ws, s_wd);
            __threadfence_block();
        }
        __syncthreads();
        if (!D.err[g] && D.kd_flag[g]) {
            KillShared* ks = reinterpret_cast<KillShared*>(s_kd);
            kill_convert<false>(D, S, g, D.kill_T[g], D.kd_rank + sb, ks);
            kill_compact(D, S, P, 0, g, D.n_s[0][g], true, D.kill_H[g], D.hl + sb, reinterpret_cast<int*>(ks->th));
            __syncthreads();
            if (threadIdx.x == 0) D.kd_flag[g] = 0;
        }
        __syncthreads();
    }
}

// ... and finishes the arterial kill of each listed graph right behind its own build (the conversion with the exact order inside
// every ball, then the compaction of the sink list): one launch instead of two, and a graph does not wait for the slowest build
// of the batch.  The set tables of the conversion live in the build's shared memory, which is idle by then.
__global__ void __launch_bounds__(1024) k_kdbuild_list(int dslot, GrowShape S, IterP P, int smem_bytes) {
    const GrowDev& D = c_dev[dslot];
    extern __shared__ __align__(16) char s_kd[];
    __shared__ int s_ws[kdpar::WS_INTS];
    __shared__ double s_wd[kdpar::WD_DOUBLES];
    static_assert(sizeof(KillShared) <= 64 * 1024, "KillShared must fit the build's shared memory");
    const int parity = P.iter & 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) D.kd_nflag[parity ^ 1] = 0;      // the next iteration's list starts empty
    const int n = D.kd_nflag[parity];
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int g = D.kd_list[parity * S.G + q];
        kdbuild_graph(D, S, smem_bytes, g, s_kd, s_ws, s_wd);
        __threadfence_block();
        __syncthreads();
        if (!D.err[g] && D.kd_flag[g]) {
            KillShared* ks = reinterpret_cast<KillShared*>(s_kd);
            const size_t sb = (size_t)g * S.capS;
            kill_convert<false>(D, S, g, D.kill_T[g], D.kd_rank + sb, ks);
            kill_compact(D, S, P, 0, g, D.n_s[0][g], true, D.kill_H[g], D.hl + sb, reinterpret_cast<int*>(ks->th));
            __syncthreads();
            if (threadIdx.x == 0) D.kd_flag[g] = 0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// launch wrappers (called from octa_grow_host.cu)
// ------------------------------------------------------------------------------------------
size_t commit_smem_bytes(const GrowShape& S) { return (size_t)S.commit_smem; }

int max_ctx_slots() { return MAX_CTX_SLOTS; }

// copy a context's pointer table into its constant-memory slot (ordered on `st`; the source may be a stack variable)
int upload_dev_table(int dslot, const GrowDev& D, cudaStream_t st) {
    return (int)cudaMemcpyToSymbolAsync(c_dev, &D, sizeof(GrowDev), sizeof(GrowDev) * (size_t)dslot, cudaMemcpyHostToDevice, st);
}

constexpr int KD_SMEM_BYTES = 208 * 1024;        // + ~17 KB static: scan / reduction scratch
constexpr int KD_SUB_SMEM_BYTES = 144 * 1024;    // a subtree below the top levels: a quarter of the list (12 000 sinks fit; larger: one-CTA build)

int prepare_kernels(const GrowShape& S) {
    cudaError_t e = cudaFuncSetAttribute(k_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)commit_smem_bytes(S));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_kdbuild, cudaFuncAttributeMaxDynamicSharedMemorySize, KD_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_kdbuild_list, cudaFuncAttributeMaxDynamicSharedMemorySize, KD_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_kd_top, cudaFuncAttributeMaxDynamicSharedMemorySize, KD_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_kd_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, KD_SUB_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_kd_fix, cudaFuncAttributeMaxDynamicSharedMemorySize, KD_SMEM_BYTES);
    // Diagnostics (OCTA_CARVEOUT=1): every growth kernel asks for the largest shared-memory carve-out, so that CTAs of different
    // kernels never wait for an SM to change its L1 / shared split.  Measured SLOWER (462 -> 443 graphs/s): the gather kernels
    // (k_assign, k_sink_tests, k_kill) lose their L1, which costs more than the extra co-residency gains.  Off by default.
    static const bool max_carve = [] { const char* v = getenv("OCTA_CARVEOUT"); return v && v[0] == '1'; }();
    if (max_carve) {
        const void* ks[] = {(const void*)k_grid_build, (const void*)k_prepare, (const void*)k_sink_tests, (const void*)k_sink_greedy,
                            (const void*)k_assign, (const void*)k_group, (const void*)k_eval, (const void*)k_commit, (const void*)k_kill,
                            (const void*)k_kdbuild};
        for (const void* k : ks)
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    }
    return (int)e;
}

// Optional per-kernel timing (OCTA_GROW_TIMING=1): an event after every launch; grow_timing_report() sums the
// intervals per kernel kind once the stream is idle.  Diagnostics only (tools/grow_probe.py).
namespace {
struct Timing {
    std::vector<cudaEvent_t> ev;
    std::vector<int> kind;
    size_t used = 0;
    cudaEvent_t next(int k) {
        if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); kind.push_back(0); }
        kind[used] = k;
        return ev[used++];
    }
};
Timing g_timing[2];        // [0] main stream, [1] side stream
bool g_timing_on = false, g_timing_init = false;
const char* const kTimingNames[] = {"start", "k_prepare", "k_sink_tests", "k_sink_greedy", "k_assign[a]", "k_group[a]", "k_eval[a]",
                                    "k_commit[a]", "k_kill[a]", "k_assign[v]", "k_group[v]", "k_eval[v]", "k_commit[v]", "k_kill[v]",
                                    "k_kdbuild_list (+fix)", "(side waits)"};
constexpr int N_KINDS = 16;
inline void tick(cudaStream_t st, int k, int which = 0) { if (g_timing_on) cudaEventRecord(g_timing[which].next(k), st); }
}  // namespace

bool grow_timing_enabled() {
    if (!g_timing_init) { g_timing_init = true; const char* e = getenv("OCTA_GROW_TIMING"); g_timing_on = e && e[0] == '1'; }
    return g_timing_on;
}

void grow_timing_begin(cudaStream_t st) {
    grow_timing_enabled();
    g_timing[0].used = g_timing[1].used = 0;
    tick(st, 0);
}

void grow_timing_report() {       // call after the streams have been synchronised
    if (!g_timing_on) return;
    for (int w = 0; w < 2; ++w) {
        const Timing& T = g_timing[w];
        if (T.used < 2) continue;
        double sum[N_KINDS] = {0};
        for (size_t i = 1; i < T.used; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, T.ev[i - 1], T.ev[i]);
            sum[T.kind[i]] += ms;
        }
        double tot = 0;
        for (int k = 1; k < N_KINDS; ++k) tot += sum[k];
        fprintf(stderr, "[octa grow timing] %s stream, total %.1f ms:", w ? "side" : "main", tot);
        for (int k = 1; k < N_KINDS; ++k) if (sum[k] > 0) fprintf(stderr, " %s %.1f", kTimingNames[k], sum[k]);
        fprintf(stderr, "\n");
    }
}

// Two pipelines per batch.  The sampling of iteration i+1 (bucket grids, candidate sampler, sink tests, greedy
// acceptance, cKDTree permutation) reads the arterial forest and the O2 sink list only, both final once the arterial
// kill of iteration i is done -- so it runs on the side stream BESIDE the venous phase of iteration i and the
// arterial growth of iteration i+1:
//   main: [wait sinks] assign/group/eval/commit [a]  [wait kd] kill[a]  assign/group/eval/commit [v] grid(ven) kill[v]
//   side:                                     [wait kill[a]] prepare tests greedy (sinks) kdbuild (kd)   -> iteration i+1
struct GrowEvents { cudaEvent_t start, sinks, kd, killa; };

void launch_sampling(int dslot, const GrowShape& S, const IterP& P, int n_sm, cudaStream_t side, const GrowEvents& ev) {
    tick(side, 15, 1);            // (the time since the previous side-stream event was spent waiting for the main stream)
    k_prepare<<<dim3(S.G, 4), 1024, 0, side>>>(dslot, S, P);
    tick(side, 1, 1);
    k_sink_tests<<<n_sm * 8, TILE, 0, side>>>(dslot, S, P);
    tick(side, 2, 1);
    k_sink_greedy<<<S.G, 1024, 0, side>>>(dslot, S, P);
    tick(side, 3, 1);
    cudaEventRecord(ev.sinks, side);
    count_launch(3);
    if (S.exact_ball_order == 1) {
        k_kdbuild<<<S.G, 1024, KD_SMEM_BYTES, side>>>(dslot, S, KD_SMEM_BYTES);
        tick(side, 14, 1);
        count_launch(1);
    }
    cudaEventRecord(ev.kd, side);
}

void launch_begin(int dslot, const GrowShape& S, const IterP& P0, int n_sm, cudaStream_t st, cudaStream_t side, const GrowEvents& ev) {
    // uploads of the initial state were issued on `st`
    k_grid_build<<<dim3(S.G, 2), 1024, 0, st>>>(dslot, S, P0, 3);
    count_launch(1);
    cudaEventRecord(ev.start, st);
    cudaStreamWaitEvent(side, ev.start, 0);
    launch_sampling(dslot, S, P0, n_sm, side, ev);
}

void launch_iteration(int dslot, const GrowShape& S, const int commit_smem[2], const IterP& P, const IterP* Pnext, int n_sm, cudaStream_t st,
                      cudaStream_t side, const GrowEvents& ev) {
    cudaStreamWaitEvent(st, ev.sinks, 0);
    for (int f = 0; f < 2; ++f) {
        k_assign<<<n_sm * 8, TILE, 0, st>>>(dslot, S, P, f);
        tick(st, 4 + 5 * f);
        k_group<<<S.G, 1024, 0, st>>>(dslot, S, P, f);
        tick(st, 5 + 5 * f);
        k_eval<<<dim3(16, S.G), 128, 0, st>>>(dslot, S, P, f);
        tick(st, 6 + 5 * f);
        // 256 threads x 128 registers = half the register file, and a mirror sized for this iteration (octa_grow_host.cu): two
        // k_commit CTAs, or k_commit and another loop's kernels, share an SM -- the replay itself is one thread
        static const int commit_threads = [] { const char* e = getenv("OCTA_COMMIT_THREADS"); const int v = e ? atoi(e) : 0; return (v == 128 || v == 256 || v == 512) ? v : 256; }();
        GrowShape Sc = S;
        Sc.commit_smem = commit_smem[f];
        k_commit<<<S.G, commit_threads, (size_t)commit_smem[f], st>>>(dslot, Sc, P, f);
        tick(st, 7 + 5 * f);
        if (f == 0) cudaStreamWaitEvent(st, ev.kd, 0);
        else { k_grid_build<<<dim3(S.G, 2), 1024, 0, st>>>(dslot, S, P, 3); count_launch(1); }
        // 512 threads x 63 registers: two k_kill CTAs (or k_kill + another loop's kernel) share an SM; the pipeline is bound by SM
        // slots held by one-CTA-per-graph kernels, not by their parallel phases (measured: 444 -> 453 graphs/s vs 1024 threads)
        static const int kill_threads = [] { const char* e = getenv("OCTA_KILL_THREADS"); const int v = e ? atoi(e) : 0; return (v == 256 || v == 512 || v == 1024) ? v : 512; }();
        k_kill<<<S.G, kill_threads, 0, st>>>(dslot, S, P, f);
        tick(st, 8 + 5 * f);
        count_launch(5);
        if (f == 0 && S.exact_ball_order == 2) {
            // exact cKDTree order on demand: permutation + conversion redone for the graphs k_kill listed (about a third of the
            // graph-iterations of the docker config); the CTAs of k_kdbuild_list own a whole SM's shared memory, so the launch is
            // half a batch wide and walks the list
            // OCTA_KD_SPLIT=1: the three-kernel form (top levels, four subtree CTAs per graph, fix).  Measured on one B200, 64 graphs:
            // a loop ALONE takes 263 instead of 299 ms (the build leaves the critical path 36 % faster), but with 7 loops in flight
            // -- where aggregate SM time and launches count, not the latency of one loop -- 628 instead of 645 graphs/s.  Default:
            // the one-kernel form; the split is for latency-bound use (single batches).
            static const bool kd_split = [] { const char* e = getenv("OCTA_KD_SPLIT"); return e && e[0] == '1'; }();
            if (kd_split) {
                k_kd_top<<<(S.G + 1) / 2, 1024, KD_SMEM_BYTES, st>>>(dslot, S, P, KD_SMEM_BYTES);
                k_kd_sub<<<dim3(KD_SUBS, (S.G + 1) / 2), 1024, KD_SUB_SMEM_BYTES, st>>>(dslot, S, P, KD_SUB_SMEM_BYTES);
                k_kd_fix<<<(S.G + 1) / 2, 1024, KD_SMEM_BYTES, st>>>(dslot, S, P, KD_SMEM_BYTES);
                tick(st, 14);
                count_launch(3);
            } else {
                k_kdbuild_list<<<(S.G + 1) / 2, 1024, KD_SMEM_BYTES, st>>>(dslot, S, P, KD_SMEM_BYTES);
                tick(st, 14);
                count_launch(1);
            }
        }
        if (f == 0) {
            cudaEventRecord(ev.killa, st);
            if (Pnext) {
                cudaStreamWaitEvent(side, ev.killa, 0);
                launch_sampling(dslot, S, *Pnext, n_sm, side, ev);
            }
        }
    }
}

}  // namespace octa
