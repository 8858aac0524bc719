// K6: anti-aliased capsule voxelizer for sm_100a.
//
// Replaces vessel_graph_generation/tree2img.py:176-280 (voxelize_forest) and :151-172
// (getCrossSlice, mode='cuboid') of the reference.  Semantics mirrored (closed form):
//   S = max(dims); D_i = max(ceil(S/76 + 0.03 S), dims_i); c = (D - dims)/2
//   per edge: p1 = node1*S + c, p2 = node2*S + c, R = radius*S           (tree2img.py:243-245)
//   bbox_i = [max(0, floor(min(p1,p2)_i - R*sqrt2)), min(D_i, ceil(max(p1,p2)_i + R*sqrt2 + 1)))  (:152-166)
//   voxel centre v+0.5; t = ((v-p2).(p1-p2)) / |p1-p2|^2                   (:259-261)
//   0<t<1 : I1 = 1 - (|v-(p2+t(p1-p2))| - (R-sqrt3/2))/sqrt3              (:262-271)
//   always: I2 = 1 - (min(|v-p1|,|v-p2|) - (R-sqrt3/2))/sqrt3              (:273-278)
//   img = max(img, I);  out = uint16(255*clip(img,0,1))                    (:279-280)
// All arithmetic is IEEE float64 with the reference's operation order and no FMA contraction
// (this file is compiled with -fmad=false); since quantisation is monotone, max is taken on the
// quantised value, which makes the accumulation order-independent and exact.
//
// Design (B200): the volume is cut into tiles [TX=16][TY=16][TZ<=64]; one CTA owns one tile of one graph and
// max-accumulates it in shared memory as u16 cells laid out exactly like the volume (27 KB for z = 53: four
// CTAs per SM).  The (y,z) rows of the clipped edge boxes are dealt out to lanes; a lane culls its row in
// fp32, the warp compacts the surviving voxels and evaluates them in float64 with all lanes busy (see
// rasterize notes below).  The finished tile leaves with TMA bulk copies shared -> global
// (cp.async.bulk, one per x plane) -- the HBM traffic is the algorithmic 2 bytes/voxel, no float scratch volume
// and no global atomics.  Edges are binned to tiles by three small kernels (prep/count, scan, fill).  Edges
// that overlap more than KBIG tiles go to a per-graph "big" list that every tile tests, which bounds the
// workspace at (sizeof(VoxEdge) + 4*KBIG + 4) bytes per edge for ANY input.
#include "octa_common.h"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int KBIG = 64;          // max tiles an edge may be listed in before it becomes a "big" edge
constexpr int TILE_Y = 8;            // 16 x 8 x Z tiles, four warps each (128 columns): 1.7 % faster than 16 x 16 tiles with eight warps (less waiting at the tile barriers)
constexpr int TILE_Z_MAX = 64;
constexpr int VOX_THREADS = 256;

struct VoxEdge {   // per edge, tile independent (vox_prep_kernel)
    double p1[3];
    double p2[3];
    double R;
    int lo[3];     // inclusive
    int hi[3];     // exclusive; lo == hi on any axis -> edge contributes nothing
    // derived constants of the column kernel: float64 fast path ...
    double ss, inv_ss, c0, tguard;
    // ... and float32 tier: axis vector, 1/|s|^2, 1/|s_xy|^2, reach = R + sqrt3/2, (reach + slack)^2, guard band of 255*I
    float f[3], finv, inv2d, reach, thr, eps;
    int zc0, zc1;  // z range outside of which the contribution is certainly zero, cut to [lo[2], hi[2])
};

struct VoxGeom {
    int D[3];        // output volume dims (image_dim)
    int T[3];        // tile dims
    int nt[3];       // tiles per axis
    int ntiles;      // nt[0]*nt[1]*nt[2]
    double S;        // scale_factor
    double c[3];     // pos_correction
    double zfix;     // image_dim[2]//2 as double (ignore_z)
    int ignore_z;
    int slowcap;     // usable entries of the column kernel's deferred-cell queue (<= SLOWCAP; tests shrink it to reach the overflow path)
    double min_radius, max_radius;
};

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

// clamp a double to int range before conversion (floor/ceil results of wild inputs)
__device__ __forceinline__ int d2i_sat(double v) {
    if (!(v > -2.0e9)) return -2000000000;
    if (v > 2.0e9) return 2000000000;
    return (int)v;
}

// ---------------------------------------------------------------------------------------------
// prep: one thread per edge.  mode 0 = count tiles, mode 1 = fill tile lists.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_range(const VoxEdge& e, const VoxGeom& g, int tlo[3], int thi[3], int& n) {
    n = 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (e.hi[a] <= e.lo[a]) { n = 0; tlo[a] = 0; thi[a] = -1; continue; }
        tlo[a] = e.lo[a] / g.T[a];
        thi[a] = (e.hi[a] - 1) / g.T[a];
        if (n) n *= (thi[a] - tlo[a] + 1);
    }
    if (e.hi[0] <= e.lo[0] || e.hi[1] <= e.lo[1] || e.hi[2] <= e.lo[2]) n = 0;
}

__global__ void vox_prep_kernel(const double* __restrict__ edges7, const int64_t* __restrict__ edge_offsets,
                                int n_graphs, VoxGeom g, VoxEdge* __restrict__ prep,
                                int* __restrict__ tile_count, int* __restrict__ big_count,
                                int* __restrict__ big_idx) {
    const int64_t n_edges = edge_offsets[n_graphs];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    // graph of this edge: binary search in the (small) offsets array
    int lo = 0, hi = n_graphs;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (edge_offsets[mid] <= i) lo = mid; else hi = mid;
    }
    const int gr = lo;
    const double* e7 = edges7 + i * 7;
    VoxEdge e;
    const double radius = e7[6];
    const bool keep = !(radius < g.min_radius || radius > g.max_radius);   // tree2img.py:227
    e.R = radius * g.S;                                                    // :243
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        e.p1[a] = e7[a] * g.S + g.c[a];                                    // :244 (mul, then add; -fmad=false)
        e.p2[a] = e7[3 + a] * g.S + g.c[a];                                // :245
    }
    if (g.ignore_z) { e.p1[2] = g.zfix; e.p2[2] = g.zfix; }                // :247-249
    const double off = e.R * 1.4142135623730951;                           // :152  (radius/voxel_size)*sqrt(2)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double s = e.p1[a], t = e.p2[a];
        if (s > t) { double tmp = s; s = t; t = tmp; }                     // :155-160
        int l = d2i_sat(floor(s - off));                                   // :161
        int h = d2i_sat(ceil(t + off + 1.0));                              // :162
        e.lo[a] = imax(0, l);
        e.hi[a] = imin(g.D[a], h);
        if (!keep || !(e.hi[a] > e.lo[a])) { e.lo[a] = 0; e.hi[a] = 0; }
    }
    if (!keep) { e.lo[0] = e.hi[0] = 0; }
    {
        const double SQRT3 = 1.7320508075688772, INV_SQRT3 = 0.57735026918962576;
        const double s0 = e.p1[0] - e.p2[0], s1 = e.p1[1] - e.p2[1], s2 = e.p1[2] - e.p2[2];
        const double ss = (s0 * s0 + s1 * s1) + s2 * s2;
        e.ss = ss;
        e.inv_ss = ss > 0.0 ? 1.0 / ss : 0.0;
        e.c0 = 1.0 + (e.R - SQRT3 / 2) * INV_SQRT3;      // I = c0 - d/sqrt3
        e.tguard = 1e-9 * ss;
        e.f[0] = (float)s0; e.f[1] = (float)s1; e.f[2] = (float)s2;
        e.finv = (float)e.inv_ss;
        const double s2d = s0 * s0 + s1 * s1;
        e.inv2d = s2d > 0.0 ? (float)(1.0 / s2d) : 0.f;
        const double reach = e.R + SQRT3 / 2;
        e.reach = (float)reach;
        // |voxel - p2| inside the box is at most |s| + R*sqrt2 + 2 per axis: the float32 tier works on numbers of that
        // size with a few roundings each (the tile-relative offsets are carried as two floats), so 255*I = 147.2*(reach-d)
        // is good to a few ulp of that magnitude; 16 ulp + 1e-3 is the guard band around the integers
        const double ext = fabs(s0) + fabs(s1) + fabs(s2) + 3.0 * (e.R * 1.4142135623730951 + 2.0) + reach;
        const double eps = 1e-3 + 147.3 * ext * (16.0 / 16777216.0);
        e.eps = (float)eps;
        const double rs = reach + eps / 100.0;
        e.thr = (float)(rs * rs * (1.0 + 1e-6));
        double zl = e.p1[2] < e.p2[2] ? e.p1[2] : e.p2[2], zh = e.p1[2] < e.p2[2] ? e.p2[2] : e.p1[2];
        e.zc0 = imax(e.lo[2], d2i_sat(floor(zl - rs - 0.51)));
        e.zc1 = imin(e.hi[2], d2i_sat(ceil(zh + rs - 0.49)) + 1);
        if (e.zc1 < e.zc0) e.zc1 = e.zc0;
    }
    prep[i] = e;
    int tlo[3], thi[3], n;
    tile_range(e, g, tlo, thi, n);
    if (n == 0) return;
    if (n > KBIG) {
        int pos = atomicAdd(&big_count[gr], 1);
        big_idx[edge_offsets[gr] + pos] = (int)(i - edge_offsets[gr]);
        return;
    }
    int* tc = tile_count + (size_t)gr * g.ntiles;
    for (int tx = tlo[0]; tx <= thi[0]; ++tx)
        for (int ty = tlo[1]; ty <= thi[1]; ++ty)
            for (int tz = tlo[2]; tz <= thi[2]; ++tz)
                atomicAdd(&tc[(tx * g.nt[1] + ty) * g.nt[2] + tz], 1);
}

// one CTA per graph: exclusive scan of that graph's tile counts -> starts (+ cursor copy)
__global__ void vox_scan_kernel(const int* __restrict__ tile_count, int* __restrict__ tile_start,
                                int* __restrict__ tile_cursor, int ntiles) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int gr = blockIdx.x;
    const int* cnt = tile_count + (size_t)gr * ntiles;
    int* st = tile_start + (size_t)gr * (ntiles + 1);
    int* cur = tile_cursor + (size_t)gr * ntiles;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < ntiles; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = (i < ntiles) ? cnt[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < nw) ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;   // inclusive
        }
        __syncthreads();
        int prefix = carry + (warp ? warp_sums[warp - 1] : 0) + (x - v);
        if (i < ntiles) { st[i] = prefix; cur[i] = prefix; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) st[ntiles] = carry;
}

__global__ void vox_fill_kernel(const int64_t* __restrict__ edge_offsets, int n_graphs, VoxGeom g,
                                const VoxEdge* __restrict__ prep, int* __restrict__ tile_cursor,
                                int* __restrict__ tile_edges) {
    const int64_t n_edges = edge_offsets[n_graphs];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    int lo = 0, hi = n_graphs;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (edge_offsets[mid] <= i) lo = mid; else hi = mid;
    }
    const int gr = lo;
    const VoxEdge e = prep[i];
    int tlo[3], thi[3], n;
    tile_range(e, g, tlo, thi, n);
    if (n == 0 || n > KBIG) return;
    int* cur = tile_cursor + (size_t)gr * g.ntiles;
    int* lst = tile_edges + (size_t)KBIG * edge_offsets[gr];
    const int local = (int)(i - edge_offsets[gr]);
    for (int tx = tlo[0]; tx <= thi[0]; ++tx)
        for (int ty = tlo[1]; ty <= thi[1]; ++ty)
            for (int tz = tlo[2]; tz <= thi[2]; ++tz) {
                int pos = atomicAdd(&cur[(tx * g.nt[1] + ty) * g.nt[2] + tz], 1);
                lst[pos] = local;
            }
}

// ---------------------------------------------------------------------------------------------
// main kernel: one CTA = one tile of one graph
// ---------------------------------------------------------------------------------------------
// exact reference chain for one voxel (tree2img.py:259-280); used only inside the guard bands of the fast path
__device__ __noinline__ uint32_t exact_voxel_q(const VoxEdge& e, double vx, double vy, double vz) {
    const double s0 = e.p1[0] - e.p2[0], s1 = e.p1[1] - e.p2[1], s2 = e.p1[2] - e.p2[2];
    const double ss = (s0 * s0 + s1 * s1) + s2 * s2;
    const double SQRT3 = 1.7320508075688772;          // np.linalg.norm([1,1,1]) (:214)
    const double rr = e.R - SQRT3 / 2;                // (radius - voxel_diag/2)
    const double u0 = vx - e.p2[0], u1 = vy - e.p2[1], u2 = vz - e.p2[2];
    const double t = ((u0 * s0 + u1 * s1) + u2 * s2) / ss;
    const double q0 = vx - e.p1[0], q1 = vy - e.p1[1], q2 = vz - e.p1[2];
    const double dcap = fmin(sqrt((q0 * q0 + q1 * q1) + q2 * q2), sqrt((u0 * u0 + u1 * u1) + u2 * u2));
    double I = 1.0 - ((dcap - rr) / SQRT3);
    if (t > 0.0 && t < 1.0) {
        const double e0 = vx - (e.p2[0] + t * s0), e1 = vy - (e.p2[1] + t * s1), e2 = vz - (e.p2[2] + t * s2);
        const double dl = sqrt((e0 * e0 + e1 * e1) + e2 * e2);
        const double I1 = 1.0 - ((dl - rr) / SQRT3);
        I = fmax(I, I1);
    }
    if (!(I > 0.0)) return 0;
    const double cl = I < 1.0 ? I : 1.0;
    return (uint32_t)(255.0 * cl);
}

// Per-edge constants staged in shared memory for one pass of EPASS edges of a tile.
constexpr int EPASS = 32;
constexpr int QCAP = 512;       // survivor queue entries per warp: 32 rows x at most 16 voxels
struct EdgeSm {
    double p1[3], p2[3], s[3];
    double R, ss, inv_ss, c0, tguard;
    float a[3], f[3], finv, thr;
    int b0[3], n[3];
    int rowbase;            // exclusive prefix of (y,z) rows over the pass
};

// Work distribution: the (y,z) rows of all clipped edge boxes of a pass are numbered consecutively and dealt out
// to the lanes of the CTA's warps, so every warp gets the same amount of work whatever the number and size of the
// edges in the tile.  Three tiers per voxel:
//   1. fp32 broad phase in box-local coordinates: a lane walks its row along x (<= 16 voxels) and produces a bit
//      mask of the voxels whose exact contribution can be > 0 (12 % survive);
//   2. the warp COMPACTS the survivors of its 32 rows into a shared-memory queue and evaluates them with all lanes
//      busy: fast float64 (fma, reciprocal instead of division); 255*I is accurate to ~1e-11, so its floor
//      equals the reference's unless 255*I lies within 1e-7 of an integer or t within 1e-9 of {0,1};
//   3. inside those guard bands (probability ~1e-7) the reference's exact operation chain decides.
// (Without the compaction the float64 tier ran at the survivors' lane density: almost every x step of a warp had
// at least one surviving lane and paid the full float64 cost for it.)
__device__ __forceinline__ int setup_edge(const VoxEdge& e, const int t0[3], const int t1[3], EdgeSm* o) {
    int rows = 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        o->b0[a] = imax(e.lo[a], t0[a]);
        o->n[a] = imin(e.hi[a], t1[a]) - o->b0[a];
        if (o->n[a] <= 0) rows = 0;
        o->p1[a] = e.p1[a]; o->p2[a] = e.p2[a];
        o->s[a] = e.p1[a] - e.p2[a];
    }
    o->rowbase = rows ? o->n[1] * o->n[2] : 0;
    o->R = e.R;
    const double ss = (o->s[0] * o->s[0] + o->s[1] * o->s[1]) + o->s[2] * o->s[2];
    o->ss = ss;
    o->inv_ss = ss > 0.0 ? 1.0 / ss : 0.0;
    const double SQRT3 = 1.7320508075688772, INV_SQRT3 = 0.57735026918962576;
    o->c0 = 1.0 + (e.R - SQRT3 / 2) * INV_SQRT3;      // I = c0 - d/sqrt3
    o->tguard = 1e-9 * ss;
    float ext = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        o->a[a] = (float)(e.p2[a] - (double)o->b0[a] - 0.5);
        o->f[a] = (float)o->s[a];
        ext += fabsf(o->a[a]) + fabsf(o->f[a]) + (float)(o->n[a] > 0 ? o->n[a] : 0);
    }
    const float fss = (float)ss;
    o->finv = fss > 0.f ? 1.0f / fss : 0.f;
    const float reach = (float)e.R + 0.8660254f + 0.02f + 8e-6f * ext;
    o->thr = reach * reach;
    return o->rowbase;
}

// tier 1: bit ix of the result is set when voxel (b0x + ix, row) may receive a positive contribution
__device__ __forceinline__ uint32_t cull_row(const EdgeSm& E, int iy, int iz) {
    const float f0 = E.f[0], f1 = E.f[1], f2 = E.f[2], a0 = E.a[0], finv = E.finv, thr = E.thr;
    const float w1 = (float)iy - E.a[1], w2 = (float)iz - E.a[2];
    const float dot12 = w1 * f1 + w2 * f2;
    const int nx = E.n[0];
    uint32_t mask = 0;
    float fx = 0.f;                                   // (float)ix without a conversion per step (small integers are exact)
#pragma unroll 4
    for (int ix = 0; ix < nx; ++ix, fx += 1.0f) {
        const float w0 = fx - a0;
        float tf = fmaf(w0, f0, dot12) * finv;
        tf = fminf(fmaxf(tf, 0.f), 1.f);
        const float d0 = fmaf(-tf, f0, w0), d1 = fmaf(-tf, f1, w1), d2 = fmaf(-tf, f2, w2);
        if (!(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)) > thr)) mask |= 1u << ix;
    }
    return mask;
}

// tiers 2 and 3 for one surviving voxel; returns the quantised contribution
__device__ __forceinline__ uint32_t eval_voxel(const EdgeSm& E, int ix, int iy, int iz) {
    const double vx = (double)(E.b0[0] + ix) + 0.5, vy = (double)(E.b0[1] + iy) + 0.5, vz = (double)(E.b0[2] + iz) + 0.5;
    const double s0 = E.s[0], s1 = E.s[1], s2 = E.s[2], ss = E.ss;
    const double u0 = vx - E.p2[0], u1 = vy - E.p2[1], u2 = vz - E.p2[2];
    const double dot = fma(u2, s2, fma(u1, s1, u0 * s0));
    uint32_t q = 0;
    bool exact = fabs(dot) < E.tguard || fabs(dot - ss) < E.tguard;
    if (!exact) {
        double dd;
        if (dot > 0.0 && dot < ss) {
            const double t = dot * E.inv_ss;
            const double e0 = fma(-t, s0, u0), e1 = fma(-t, s1, u1), e2 = fma(-t, s2, u2);
            dd = fma(e2, e2, fma(e1, e1, e0 * e0));
        } else {
            const double q0 = vx - E.p1[0], q1 = vy - E.p1[1], q2 = vz - E.p1[2];
            dd = fmin(fma(u2, u2, fma(u1, u1, u0 * u0)), fma(q2, q2, fma(q1, q1, q0 * q0)));
        }
        const double val = 255.0 * fma(-sqrt(dd), 0.57735026918962576, E.c0);
        if (val < -1e-7) return 0;
        const double fl = floor(val);
        const double fr = val - fl;
        exact = fr < 1e-7 || fr > 1.0 - 1e-7 || !(val == val);
        q = val >= 255.0 ? 255u : (uint32_t)fl;
    }
    if (exact) {
        VoxEdge e;
#pragma unroll
        for (int a = 0; a < 3; ++a) { e.p1[a] = E.p1[a]; e.p2[a] = E.p2[a]; }
        e.R = E.R;
        q = exact_voxel_q(e, vx, vy, vz);
    }
    return q;
}

// max-accumulate a 16-bit tile cell (values 0..255; contention is rare, most updates stop at the first load)
__device__ __forceinline__ void tile_max(unsigned short* p, uint32_t q) {
    unsigned short old = *(volatile unsigned short*)p;
    while (old < q) {
        const unsigned short seen = atomicCAS(p, old, (unsigned short)q);
        if (seen == old) break;
        old = seen;
    }
}

__global__ void __launch_bounds__(VOX_THREADS, 4)
vox_tile_kernel(const VoxEdge* __restrict__ prep, const int64_t* __restrict__ edge_offsets, VoxGeom g,
                const int* __restrict__ tile_start, const int* __restrict__ tile_edges,
                const int* __restrict__ big_count, const int* __restrict__ big_idx,
                uint16_t* __restrict__ out) {
    // dynamic shared memory: the tile accumulators, u16 [TX][TY][TZ] in the volume's own layout
    extern __shared__ __align__(128) unsigned short acc[];
    // grid = (tiles along y and z, tiles along x, graphs)
    const int gr = blockIdx.z;
    const int tx_i = blockIdx.y, ty_i = blockIdx.x / g.nt[2], tz_i = blockIdx.x - ty_i * g.nt[2];
    const int tile = (tx_i * g.nt[1] + ty_i) * g.nt[2] + tz_i;
    const int T[3] = {g.T[0], g.T[1], g.T[2]};
    const int t0[3] = {tx_i * T[0], ty_i * T[1], tz_i * T[2]};
    const int t1[3] = {imin(t0[0] + T[0], g.D[0]), imin(t0[1] + T[1], g.D[1]), imin(t0[2] + T[2], g.D[2])};
    const int tile_elems = T[0] * T[1] * T[2];
    const int64_t e_base = edge_offsets[gr];
    const VoxEdge* ge = prep + e_base;
    uint16_t* vol = out + (size_t)gr * g.D[0] * g.D[1] * g.D[2];
    const int ny = t1[1] - t0[1], nz = t1[2] - t0[2];
    // the (y,z) plane of one x of the tile is one contiguous run both in shared memory and in the volume
    const bool plane_contig = (nz == g.D[2] && T[2] == g.D[2]);
    const int plane_len = ny * nz;

    const int* st = tile_start + (size_t)gr * (g.ntiles + 1);
    const int beg = st[tile], end = st[tile + 1];
    const int nbig = big_count[gr];

    if (!(end > beg || nbig > 0)) {
        // empty tile: stream zeros without touching shared memory
        for (int x = t0[0]; x < t1[0]; ++x) {
            if (plane_contig) {
                const size_t base = ((size_t)x * g.D[1] + t0[1]) * g.D[2];
                if (((base | (size_t)plane_len) & 7) == 0) {
                    uint4* p = reinterpret_cast<uint4*>(vol + base);
                    for (int i = threadIdx.x; i < plane_len / 8; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
                } else {
                    for (int i = threadIdx.x; i < plane_len; i += blockDim.x) vol[base + i] = 0;
                }
            } else {
                for (int i = threadIdx.x; i < plane_len; i += blockDim.x) {
                    const int y = i / nz, z = i - y * nz;
                    vol[((size_t)x * g.D[1] + t0[1] + y) * g.D[2] + t0[2] + z] = 0;
                }
            }
        }
        return;
    }

    {
        uint4* acc4 = reinterpret_cast<uint4*>(acc);
        for (int i = threadIdx.x; i < tile_elems / 8; i += blockDim.x) acc4[i] = make_uint4(0, 0, 0, 0);
        for (int i = (tile_elems / 8) * 8 + threadIdx.x; i < tile_elems; i += blockDim.x) acc[i] = 0;
    }
    __shared__ EdgeSm es[EPASS];
    __shared__ unsigned short s_queue[(VOX_THREADS / 32) * QCAP];
    __shared__ int s_total, s_rowbase[EPASS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned short* queue = s_queue + warp * QCAP;
    const int* lst = tile_edges + (size_t)KBIG * e_base;
    const int* bl = big_idx + e_base;
    const int nlist = end - beg, nall = nlist + nbig;
    const int ystride = T[2], xstride = T[1] * T[2];
    for (int pass = 0; pass < nall; pass += EPASS) {
        const int cnt = imin(EPASS, nall - pass);
        if (warp == 0) {                   // EPASS == 32: warp 0 sets the pass up and scans the row counts, ONE barrier per pass
            int v = 0;                     // (the first pass shares it with the zeroing of the accumulators above)
            if (lane < cnt) {
                const int k = pass + lane;
                const VoxEdge e = ge[k < nlist ? lst[beg + k] : bl[k - nlist]];
                v = setup_edge(e, t0, t1, &es[lane]);
            }
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane < cnt) es[lane].rowbase = x - v;
            s_rowbase[lane] = x - v;
            if (lane == 31) s_total = x;
        }
        __syncthreads();
        const int total = s_total;
        int lo = 0;                        // last edge with rowbase <= item (items of a lane only move forward)
        for (int base = warp * 32; base < total; base += VOX_THREADS) {      // warp-uniform trip count
            const int item = base + lane;
            uint32_t mask = 0, rowinfo = 0;
            if (item < total) {
                while (lo + 1 < cnt && s_rowbase[lo + 1] <= item) ++lo;
                const EdgeSm& E = es[lo];
                const int row = item - E.rowbase, eny = E.n[1];      // rows numbered y-fastest: neighbouring lanes update
                int iz = (int)((float)row * (1.0f / (float)eny));    // different 32-bit words of the u16 accumulators
                int iy = row - iz * eny;
                if (iy < 0) { --iz; iy += eny; } else if (iy >= eny) { ++iz; iy -= eny; }
                mask = cull_row(E, iy, iz);
                rowinfo = ((uint32_t)lo << 16) | ((uint32_t)iy << 8) | (uint32_t)iz;
            }
            // compaction: queue entry = (source lane << 4) | ix
            const int c = __popc(mask);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            int w = incl - c;
            while (mask) {
                const int ix = __ffs(mask) - 1;
                mask &= mask - 1;
                queue[w++] = (unsigned short)((lane << 4) | ix);
            }
            __syncwarp();
            for (int jb = 0; jb < tot; jb += 32) {
                const int j = jb + lane;
                const bool on = j < tot;
                const uint32_t ent = on ? queue[j] : 0u;
                const uint32_t ri = __shfl_sync(0xffffffffu, rowinfo, ent >> 4);
                if (on) {
                    const EdgeSm& E = es[ri >> 16];
                    const int ix = ent & 15, iy = (ri >> 8) & 255, iz = ri & 255;
                    const uint32_t q = eval_voxel(E, ix, iy, iz);
                    if (q) tile_max(acc + (E.b0[0] + ix - t0[0]) * xstride + (E.b0[1] + iy - t0[1]) * ystride + (E.b0[2] + iz - t0[2]), q);
                }
            }
            __syncwarp();
        }
        if (pass + EPASS >= nall) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (see the bulk stores below)
        __syncthreads();
    }

    // stream the tile out: the accumulators already are the volume's u16 cells, 2 algorithmic bytes per voxel
    const size_t base0 = ((size_t)t0[0] * g.D[1] + t0[1]) * g.D[2];
    const size_t xpitch = (size_t)g.D[1] * g.D[2];
    if (plane_contig && (((base0 | xpitch | (size_t)plane_len | (size_t)xstride) & 7) == 0)) {
        // TMA bulk copies shared -> global, one per x plane (16-byte aligned runs), issued by one thread; every thread fenced
        // its accumulator writes for the async proxy before the closing barrier of the last pass
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)plane_len * 2u;
            for (int x = 0; x < t1[0] - t0[0]; ++x) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(acc + (size_t)x * xstride);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(vol + base0 + (size_t)x * xpitch), "r"(src), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    for (int x = t0[0]; x < t1[0]; ++x) {
        const unsigned short* slab = acc + (size_t)(x - t0[0]) * xstride;
        if (plane_contig) {
            const size_t base = ((size_t)x * g.D[1] + t0[1]) * g.D[2];
            for (int i = threadIdx.x; i < plane_len; i += blockDim.x) vol[base + i] = slab[i];
        } else {
            for (int i = threadIdx.x; i < plane_len; i += blockDim.x) {
                const int y = i / nz, z = i - y * nz;
                vol[((size_t)x * g.D[1] + t0[1] + y) * g.D[2] + t0[2] + z] = slab[y * T[2] + z];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel, column ownership (default): one CTA = one tile of one graph, one THREAD = one (x,y) column of the tile
// ---------------------------------------------------------------------------------------------
// vox_tile_kernel above deals (y,z) rows of the edge boxes to lanes, so two lanes can meet in one voxel and every update is
// a 16-bit compare-and-swap (emulated: a 32-bit CAS loop with byte permutes), and it walks every row of every box voxel
// by voxel to find the 18 % of the candidates that contribute.  Here a thread OWNS the z column of one (x,y) of the tile:
//   * no atomics: a column is updated by plain 16-bit loads and stores of its owner;
//   * culling is per (column, edge), not per candidate voxel: box test, then the distance of the column from the
//     xy projection of the capsule axis (a lower bound of the 3-D distance) -- vessels are thin (median radius one voxel
//     at 1216^2), so a 16 x 16 box of an edge keeps ~28 of its 256 columns;
//   * a surviving column walks the z range the capsule can reach (6 voxels on average) in float32: the offsets of the
//     column from the edge's end point are formed from two-float tile-relative constants, 255*I = 147.2*(reach - d) is
//     accurate to a few 1e-3 (VoxEdge::eps), so its floor equals the reference's unless it lies within eps of an integer;
//   * those voxels (< 1 %) take the float64 fast path (fma, reciprocal: 255*I to ~1e-11), and inside 1e-7 of an integer or
//     with t within 1e-9 of {0,1} the reference's exact operation chain decides (exact_voxel_q) -- as in the row kernel.
// The tile is accumulated as u16 cells in the volume's own layout and leaves with TMA bulk copies, as before.
struct ColEdge {             // per edge of a pass, tile-local
    float ah[3], al[3];      // p2 - (tile origin + 0.5) as hi + lo
    float f[3], finv, inv2d, reach, thr, eps;
    short x0, nx, y0, ny, z0, nz;
    float rfxy;              // 1 / |s_xy| (0: the edge is parallel to z)
    int idx;
};
constexpr int SLOWCAP = 160; // deferred float64 evaluations per CTA (~15 per tile are usual; beyond: the cell is marked and recomputed)

__device__ __noinline__ uint32_t slow_voxel_q(const VoxEdge* __restrict__ ep, int vx_i, int vy_i, int vz_i) {
    const VoxEdge& E = *ep;
    const double vx = (double)vx_i + 0.5, vy = (double)vy_i + 0.5, vz = (double)vz_i + 0.5;
    const double s0 = E.p1[0] - E.p2[0], s1 = E.p1[1] - E.p2[1], s2 = E.p1[2] - E.p2[2], ss = E.ss;
    const double u0 = vx - E.p2[0], u1 = vy - E.p2[1], u2 = vz - E.p2[2];
    const double dot = fma(u2, s2, fma(u1, s1, u0 * s0));
    uint32_t q = 0;
    bool exact = fabs(dot) < E.tguard || fabs(dot - ss) < E.tguard;
    if (!exact) {
        double dd;
        if (dot > 0.0 && dot < ss) {
            const double t = dot * E.inv_ss;
            const double e0 = fma(-t, s0, u0), e1 = fma(-t, s1, u1), e2 = fma(-t, s2, u2);
            dd = fma(e2, e2, fma(e1, e1, e0 * e0));
        } else {
            const double q0 = vx - E.p1[0], q1 = vy - E.p1[1], q2 = vz - E.p1[2];
            dd = fmin(fma(u2, u2, fma(u1, u1, u0 * u0)), fma(q2, q2, fma(q1, q1, q0 * q0)));
        }
        const double val = 255.0 * fma(-sqrt(dd), 0.57735026918962576, E.c0);
        if (val < -1e-7) return 0;
        const double fl = floor(val);
        const double fr = val - fl;
        exact = fr < 1e-7 || fr > 1.0 - 1e-7 || !(val == val);
        q = val >= 255.0 ? 255u : (uint32_t)fl;
    }
    if (exact) q = exact_voxel_q(E, vx, vy, vz);
    return q;
}

// cells marked 0xffff (their deferred evaluation did not fit the queue): exact maximum over every edge of the tile
__device__ __noinline__ void resolve_marked(const VoxEdge* __restrict__ ge, const int* __restrict__ lst, int nlist,
                                            const int* __restrict__ bl, int nbig, unsigned short* acc, int tile_elems,
                                            int t0x, int t0y, int t0z, int xstride, int ystride) {
    for (int c = threadIdx.x; c < tile_elems; c += blockDim.x) {
        if (acc[c] != 0xffffu) continue;
        const int cx = c / xstride, r = c - cx * xstride, cy = r / ystride, z = r - cy * ystride;
        const int vx = t0x + cx, vy = t0y + cy, vz = t0z + z;
        uint32_t best = 0;
        for (int k = 0; k < nlist + nbig; ++k) {
            const VoxEdge* e = ge + (k < nlist ? lst[k] : bl[k - nlist]);
            if (vx < e->lo[0] || vx >= e->hi[0] || vy < e->lo[1] || vy >= e->hi[1] || vz < e->lo[2] || vz >= e->hi[2]) continue;
            const uint32_t q = slow_voxel_q(e, vx, vy, vz);
            best = q > best ? q : best;
        }
        acc[c] = (unsigned short)best;
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
vox_col_kernel(const VoxEdge* __restrict__ prep, const int64_t* __restrict__ edge_offsets, VoxGeom g,
               const int* __restrict__ tile_start, const int* __restrict__ tile_edges,
               const int* __restrict__ big_count, const int* __restrict__ big_idx,
               uint16_t* __restrict__ out) {
    extern __shared__ __align__(128) unsigned short acc[];
    const int gr = blockIdx.z;
    const int tx_i = blockIdx.y, ty_i = blockIdx.x / g.nt[2], tz_i = blockIdx.x - ty_i * g.nt[2];
    const int tile = (tx_i * g.nt[1] + ty_i) * g.nt[2] + tz_i;
    const int t0x = tx_i * g.T[0], t0y = ty_i * g.T[1], t0z = tz_i * g.T[2];
    const int64_t e_base = edge_offsets[gr];
    const VoxEdge* ge = prep + e_base;
    uint16_t* vol = out + (size_t)gr * g.D[0] * g.D[1] * g.D[2];
    const int* st = tile_start + (size_t)gr * (g.ntiles + 1);
    const int beg = st[tile], end = st[tile + 1];
    const int nbig = big_count[gr];

    if (!(end > beg || nbig > 0)) {
        // empty tile: stream zeros without touching shared memory
        const int t1x = imin(t0x + g.T[0], g.D[0]);
        const int ny = imin(t0y + g.T[1], g.D[1]) - t0y, nz = imin(t0z + g.T[2], g.D[2]) - t0z;
        const bool plane_contig = (nz == g.D[2] && g.T[2] == g.D[2]);
        const int plane_len = ny * nz;
        for (int x = t0x; x < t1x; ++x) {
            if (plane_contig) {
                const size_t base = ((size_t)x * g.D[1] + t0y) * g.D[2];
                if (((base | (size_t)plane_len) & 7) == 0) {
                    uint4* p = reinterpret_cast<uint4*>(vol + base);
                    for (int i = threadIdx.x; i < plane_len / 8; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
                } else {
                    for (int i = threadIdx.x; i < plane_len; i += blockDim.x) vol[base + i] = 0;
                }
            } else {
                for (int i = threadIdx.x; i < plane_len; i += blockDim.x) {
                    const int y = i / nz, z = i - y * nz;
                    vol[((size_t)x * g.D[1] + t0y + y) * g.D[2] + t0z + z] = 0;
                }
            }
        }
        return;
    }

    const int tile_elems = g.T[0] * g.T[1] * g.T[2];
    {
        uint4* acc4 = reinterpret_cast<uint4*>(acc);
        for (int i = threadIdx.x; i < tile_elems / 8; i += blockDim.x) acc4[i] = make_uint4(0, 0, 0, 0);
        for (int i = (tile_elems / 8) * 8 + threadIdx.x; i < tile_elems; i += blockDim.x) acc[i] = 0;
    }
    __shared__ ColEdge ce[EPASS];
    __shared__ float4 s_hit[NT / 32][32];         // per warp: (w0, w1, dxy, cell offset of the column) of the hit columns
    __shared__ int2 s_slow[SLOWCAP];                        // deferred float64 evaluations of the CTA: (edge, cell)
    __shared__ int s_nslow;
    if (threadIdx.x == 0) s_nslow = 0;                      // (ordered before its first use by the barrier of the first pass)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    float4* hitw = s_hit[warp];
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(acc);
    const int* lst = tile_edges + (size_t)KBIG * e_base;
    const int* bl = big_idx + e_base;
    const int nlist = end - beg, nall = nlist + nbig;
    const int ystride = g.T[2], xstride = g.T[1] * g.T[2];
    // columns of the tile -> threads.  16 x 16 tiles: a warp owns an 8 x 4 block of columns
    const int ncols = g.T[0] * g.T[1];
    const bool blocked = (g.T[0] == 16 && g.T[1] == 16 && NT == 256);
    for (int pass = 0; pass < nall; pass += EPASS) {
        const int cnt = imin(EPASS, nall - pass);
        if (pass > 0) __syncthreads();                 // the previous pass has been read by every thread
        if (threadIdx.x < cnt) {
            const int k = pass + threadIdx.x;
            const int idx = k < nlist ? lst[beg + k] : bl[k - nlist];
            const VoxEdge& e = ge[idx];
            ColEdge c;
            c.idx = idx;
            const int t0[3] = {t0x, t0y, t0z};
            int lo3[3], n3[3];
            bool any = true;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int elo = a == 2 ? e.zc0 : e.lo[a], ehi = a == 2 ? e.zc1 : e.hi[a];
                lo3[a] = imax(elo, t0[a]) - t0[a];
                n3[a] = imin(imin(ehi, t0[a] + g.T[a]), g.D[a]) - t0[a] - lo3[a];
                if (n3[a] <= 0) any = false;
                const double av = e.p2[a] - (double)t0[a] - 0.5;
                c.ah[a] = (float)av;
                c.al[a] = (float)(av - (double)c.ah[a]);
                c.f[a] = e.f[a];
            }
            c.x0 = (short)lo3[0]; c.nx = (short)(any ? n3[0] : 0);
            c.y0 = (short)lo3[1]; c.ny = (short)(any ? n3[1] : 0);
            c.z0 = (short)lo3[2]; c.nz = (short)(any ? n3[2] : 0);
            c.rfxy = sqrtf(e.inv2d);
            c.finv = e.finv; c.inv2d = e.inv2d; c.reach = e.reach; c.thr = e.thr; c.eps = e.eps;
            ce[threadIdx.x] = c;
        }
        __syncthreads();                               // (also: the accumulators are cleared)
        for (int cb = warp * 32; cb < ncols; cb += NT) {      // warp-uniform
            int cx, cy, bx0, by0, bnx, bny;            // this lane's column; the box of the warp's 32 columns
            if (blocked) {
                bx0 = (warp & 1) << 3; by0 = (warp >> 1) << 2; bnx = 8; bny = 4;
                cx = bx0 + (lane & 7); cy = by0 + (lane >> 3);
            } else {
                const int col = cb + lane;
                cx = col / g.T[1]; cy = col - cx * g.T[1];
                const int c1 = imin(cb + 31, ncols - 1);
                bx0 = cb / g.T[1]; bnx = c1 / g.T[1] - bx0 + 1;
                by0 = bnx > 1 ? 0 : cb - bx0 * g.T[1]; bny = bnx > 1 ? g.T[1] : c1 - cb + 1;
            }
            const float coff = __int_as_float(cx * xstride + cy * ystride);
            const float fcx = (float)cx, fcy = (float)cy;
            // edges of the pass whose box meets the warp's columns at all (lane k looks at edge k)
            unsigned todo;
            {
                bool ov = false;
                if (lane < cnt) {
                    const ColEdge& E = ce[lane];
                    ov = E.x0 < bx0 + bnx && bx0 < E.x0 + E.nx && E.y0 < by0 + bny && by0 < E.y0 + E.ny;
                    if (ov) {
                        // the box of a thin diagonal edge is mostly empty: also require the capsule's xy shadow to come within
                        // the half diagonal of the warp's block of its centre (triangle inequality; 0.01 covers the float32 error)
                        const float mx = (float)bx0 + 0.5f * (float)(bnx - 1), my = (float)by0 + 0.5f * (float)(bny - 1);
                        const float w0 = (mx - E.ah[0]) - E.al[0], w1 = (my - E.ah[1]) - E.al[1];
                        const float t2 = __saturatef(fmaf(w1, E.f[1], w0 * E.f[0]) * E.inv2d);
                        const float e0 = fmaf(-t2, E.f[0], w0), e1 = fmaf(-t2, E.f[1], w1);
                        const float hx = 0.5f * (float)(bnx - 1), hy = 0.5f * (float)(bny - 1);
                        const float lim = sqrtf(E.thr) + sqrtf(hx * hx + hy * hy) + 0.01f;
                        ov = !(fmaf(e1, e1, e0 * e0) > lim * lim);
                    }
                }
                todo = __ballot_sync(0xffffffffu, ov);
            }
            while (todo) {
                const int k = __ffs(todo) - 1;
                todo &= todo - 1;
                const ColEdge& E = ce[k];
                const float f0 = E.f[0], f1 = E.f[1];
                bool hit = (unsigned)(cx - E.x0) < (unsigned)E.nx && (unsigned)(cy - E.y0) < (unsigned)E.ny;
                float w0 = 0.f, w1 = 0.f, dxy = 0.f;
                int zl = 0, zn = 0;
                if (hit) {
                    w0 = (fcx - E.ah[0]) - E.al[0]; w1 = (fcy - E.ah[1]) - E.al[1];
                    dxy = fmaf(w1, f1, w0 * f0);
                    const float tp = dxy * E.inv2d;
                    const float t2 = __saturatef(tp);
                    const float e0 = fmaf(-t2, f0, w0), e1 = fmaf(-t2, f1, w1);
                    hit = !(fmaf(e1, e1, e0 * e0) > E.thr);              // else: the column misses the capsule's xy shadow
                    if (hit) {
                        // z cells of THIS column the capsule can reach.  A cell within r (r^2 = thr) of the point q(t*) of the
                        // axis has |xy - q_xy(t*)| <= r, so t* lies within sqrt(r^2 - p^2) / |s_xy| of the column's projection
                        // parameter tp (p = distance of the column from the axis LINE in xy), and |z - q_z(t*)| <= sqrt(r^2 - p^2).
                        const float p0 = fmaf(-tp, f0, w0), p1 = fmaf(-tp, f1, w1);
                        float hh;
                        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(hh) : "f"(fmaxf(E.thr - fmaf(p1, p1, p0 * p0), 0.f)));
                        hh += 0.02f;
                        float tlo = 0.f, thi = 1.f;
                        if (E.inv2d > 0.f) { const float dt = fmaf(hh, E.rfxy, 1e-4f); tlo = __saturatef(tp - dt); thi = __saturatef(tp + dt); }
                        const float a2 = E.ah[2] + E.al[2], f2e = E.f[2];
                        const float za = fmaf(f2e, tlo, a2), zb = fmaf(f2e, thi, a2);
                        zl = max((int)ceilf(fminf(za, zb) - hh), (int)E.z0);
                        zn = min((int)floorf(fmaxf(za, zb) + hh), (int)E.z0 + (int)E.nz - 1) - zl + 1;
                        hit = zn > 0;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (!m) continue;
                // Within one edge all (column, z) cells are distinct and the columns belong to this warp: the cells of the hit
                // columns are dealt to ALL lanes (plain loads and stores, no atomics), whatever the shape of the hit set.  A hit
                // column brings its own z run (zl, zn); the runs are dealt with the stride of the longest one.
                const int nzr = (int)__reduce_max_sync(0xffffffffu, (unsigned)(hit ? zn : 0));
                if (hit) hitw[__popc(m & lt)] = make_float4(w0, w1, dxy, __int_as_float(__float_as_int(coff) | (zl << 14) | (zn << 21)));
                __syncwarp();
                const int total = __popc(m) * nzr, rcp = ((1 << 20) + nzr - 1) / nzr, eidx = E.idx;
                const float f2 = E.f[2], finv = E.finv, reach = E.reach, eps = E.eps, thr = E.thr, ah2 = E.ah[2], al2 = E.al[2];
                const float hi255 = 255.f + eps, om = 1.f - eps;
#pragma unroll 2
                for (int j = lane; j < total; j += 32) {
                    const int h = (int)(((unsigned)j * (unsigned)rcp) >> 20);
                    const float4 hv = hitw[h];
                    const int pk = __float_as_int(hv.w), dz = j - h * nzr;
                    if (dz >= (pk >> 21)) continue;
                    const int z = ((pk >> 14) & 127) + dz;
                    const float w2 = ((float)z - ah2) - al2;
                    const float t = __saturatef(fmaf(w2, f2, hv.z) * finv);
                    const float d0 = fmaf(-t, f0, hv.x), d1 = fmaf(-t, f1, hv.y), d2 = fmaf(-t, f2, w2);
                    const float dd = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
                    if (dd > thr) continue;
                    float d;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(dd));
                    const float v = (reach - d) * 147.22431864335458f;       // 255 * I
                    const float fl = floorf(v), fr = v - fl;
                    const int cell = (pk & 16383) + z;
                    // certain: 255*I >= 255 + eps (clipped to 255), or at least eps away from every integer
                    const bool top = v >= hi255;
                    if (!top && (!(fr >= eps) || fr > om)) {               // within eps of an integer (or NaN): float64 decides
                        if (!(v <= -eps)) {
                            const int slot = atomicAdd(&s_nslow, 1);
                            if (slot < g.slowcap) s_slow[slot] = make_int2(eidx, cell);
                            else { const unsigned short mark = 0xffffu; asm volatile("st.shared.u16 [%0], %1;" ::"r"(acc_s + 2u * cell), "h"(mark)); }
                        }
                        continue;
                    }
                    if (v < 1.f) continue;                                  // quantises to 0
                    const unsigned short q = top ? (unsigned short)255 : (unsigned short)(int)fl;
                    unsigned short old;
                    const uint32_t ca = acc_s + 2u * cell;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(old) : "r"(ca));
                    if (q > old) asm volatile("st.shared.u16 [%0], %1;" ::"r"(ca), "h"(q));
                }
                __syncwarp();                                              // hitw is rewritten by the next edge
            }
        }
    }
    __syncthreads();
    {   // the deferred cells of the CTA: float64 fast path / exact chain; two entries may name the same cell -> CAS max
        const int ns = imin(s_nslow, g.slowcap);
        for (int i = threadIdx.x; i < ns; i += NT) {
            const int2 ent = s_slow[i];
            const int cx = ent.y / xstride, r = ent.y - cx * xstride, cy = r / ystride, z = r - cy * ystride;
            const uint32_t v = slow_voxel_q(ge + ent.x, t0x + cx, t0y + cy, t0z + z);
            if (v) tile_max(acc + ent.y, v);
        }
        if (s_nslow > g.slowcap) {                    // (never seen in real graphs) marked cells: recomputed from every edge of the tile
            __syncthreads();
            resolve_marked(ge, lst + beg, nlist, bl, nbig, acc, tile_elems, t0x, t0y, t0z, xstride, ystride);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (the bulk stores below read the accumulators)
    __syncthreads();

    const int t1x = imin(t0x + g.T[0], g.D[0]);
    const int ny = imin(t0y + g.T[1], g.D[1]) - t0y, nz = imin(t0z + g.T[2], g.D[2]) - t0z;
    const bool plane_contig = (nz == g.D[2] && g.T[2] == g.D[2]);
    const int plane_len = ny * nz;
    const size_t base0 = ((size_t)t0x * g.D[1] + t0y) * g.D[2];
    const size_t xpitch = (size_t)g.D[1] * g.D[2];
    if (plane_contig && (((base0 | xpitch | (size_t)plane_len | (size_t)xstride) & 7) == 0)) {
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)plane_len * 2u;
            for (int x = 0; x < t1x - t0x; ++x) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(acc + (size_t)x * xstride);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(vol + base0 + (size_t)x * xpitch), "r"(src), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    for (int x = t0x; x < t1x; ++x) {
        const unsigned short* slab = acc + (size_t)(x - t0x) * xstride;
        if (plane_contig) {
            const size_t base = ((size_t)x * g.D[1] + t0y) * g.D[2];
            for (int i = threadIdx.x; i < plane_len; i += blockDim.x) vol[base + i] = slab[i];
        } else {
            for (int i = threadIdx.x; i < plane_len; i += blockDim.x) {
                const int y = i / nz, z = i - y * nz;
                vol[((size_t)x * g.D[1] + t0y + y) * g.D[2] + t0z + z] = slab[y * g.T[2] + z];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_geom(const int dims[3], const OctaVoxOpts* opts, VoxGeom* g) {
    OCTA_ARG_CHECK(dims && dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "volume dimensions must be positive");
    int S = dims[0] > dims[1] ? dims[0] : dims[1];
    if (dims[2] > S) S = dims[2];
    OCTA_ARG_CHECK(S <= 16384, "volume dimension too large");
    // tree2img.py:206-211
    const double MAX_RADIUS = 0.015;
    const double sf = (double)S;
    const int min_dim = (int)ceil((1.0 / 76) * sf + 2 * MAX_RADIUS * sf);
    g->S = sf;
    for (int a = 0; a < 3; ++a) {
        g->D[a] = dims[a] > min_dim ? dims[a] : min_dim;
        g->c[a] = (double)(g->D[a] - dims[a]) / 2;
    }
    g->zfix = (double)(g->D[2] / 2);
    g->ignore_z = opts ? opts->ignore_z : 0;
    g->min_radius = opts ? opts->min_radius : 0.0;
    g->max_radius = opts ? opts->max_radius : 1.0;
    g->slowcap = 160;
    if (const char* ev = getenv("OCTA_VOX_SLOWCAP")) { const int v = atoi(ev); if (v >= 0 && v <= 160) g->slowcap = v; }   // test knob
    g->T[1] = TILE_Y;
    if (const char* ev = getenv("OCTA_VOX_TILE_Y")) { const int v = atoi(ev); if (v == 8 || v == 16 || v == 32) g->T[1] = v; }   // tuning knob
    g->T[2] = g->D[2] < TILE_Z_MAX ? g->D[2] : TILE_Z_MAX;
    g->T[0] = 16;                  // <= 16: a row's survivors are a 16-bit mask
    while (g->T[0] > 1 && (size_t)g->T[0] * g->T[1] * g->T[2] * sizeof(uint16_t) > 56 * 1024) g->T[0] >>= 1;
    for (int a = 0; a < 3; ++a) g->nt[a] = (g->D[a] + g->T[a] - 1) / g->T[a];
    g->ntiles = g->nt[0] * g->nt[1] * g->nt[2];
    return OCTA_OK;
}

struct VoxWorkspace {
    VoxEdge* prep;
    int64_t* edge_offsets;
    int* tile_count;
    int* tile_start;
    int* tile_cursor;
    int* big_count;
    int* big_idx;
    int* tile_edges;
    size_t bytes;
};

VoxWorkspace carve(void* base, int n_graphs, int64_t n_edges, int ntiles) {
    VoxWorkspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = octa::align_up(off + bytes, 256); return (char*)base + o; };
    w.prep = (VoxEdge*)take(sizeof(VoxEdge) * (size_t)(n_edges > 0 ? n_edges : 1));
    w.edge_offsets = (int64_t*)take(sizeof(int64_t) * (size_t)(n_graphs + 1));
    w.tile_count = (int*)take(sizeof(int) * (size_t)n_graphs * ntiles);
    w.big_count = (int*)take(sizeof(int) * (size_t)n_graphs);   // directly after tile_count: one memset
    w.tile_start = (int*)take(sizeof(int) * (size_t)n_graphs * (ntiles + 1));
    w.tile_cursor = (int*)take(sizeof(int) * (size_t)n_graphs * ntiles);
    w.big_idx = (int*)take(sizeof(int) * (size_t)(n_edges > 0 ? n_edges : 1));
    w.tile_edges = (int*)take(sizeof(int) * (size_t)KBIG * (size_t)(n_edges > 0 ? n_edges : 1));
    w.bytes = off;
    return w;
}

}  // namespace

extern "C" int octa_voxelize_out_dims(const int dims[3], int out_dims[3]) {
    VoxGeom g;
    int rc = make_geom(dims, nullptr, &g);
    if (rc) return rc;
    OCTA_ARG_CHECK(out_dims, "out_dims is null");
    for (int a = 0; a < 3; ++a) out_dims[a] = g.D[a];
    return OCTA_OK;
}

extern "C" size_t octa_voxelize_workspace_bytes(int n_graphs, int64_t n_edges, const int dims[3]) {
    VoxGeom g;
    if (n_graphs <= 0 || n_edges < 0 || make_geom(dims, nullptr, &g)) return 0;
    return carve(nullptr, n_graphs, n_edges, g.ntiles).bytes;
}

extern "C" int octa_voxelize_batch_dev(const double* edges7_dev, const int64_t* edge_offsets_host, int n_graphs,
                                       const int dims[3], const OctaVoxOpts* opts, uint16_t* out_dev,
                                       void* workspace_dev, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    OCTA_ARG_CHECK(n_graphs > 0 && n_graphs <= 65535, "n_graphs must be in [1, 65535]");
    OCTA_ARG_CHECK(edge_offsets_host && out_dev && workspace_dev, "null pointer");
    OCTA_ARG_CHECK(edge_offsets_host[0] == 0, "edge_offsets[0] must be 0");
    for (int i = 0; i < n_graphs; ++i)
        OCTA_ARG_CHECK(edge_offsets_host[i + 1] >= edge_offsets_host[i], "edge_offsets must be non-decreasing");
    const int64_t n_edges = edge_offsets_host[n_graphs];
    OCTA_ARG_CHECK(n_edges == 0 || edges7_dev, "edges pointer is null");
    OCTA_ARG_CHECK(n_edges < (int64_t)1 << 31, "too many edges");
    VoxGeom g;
    int rc = make_geom(dims, opts, &g);
    if (rc) return rc;
    VoxWorkspace w = carve(workspace_dev, n_graphs, n_edges, g.ntiles);
    if (w.bytes > workspace_bytes) {
        octa::set_error("octa_voxelize_batch_dev: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return OCTA_E_NOMEM;
    }
    OCTA_CUDA_CHECK(cudaMemcpyAsync(w.edge_offsets, edge_offsets_host, sizeof(int64_t) * (n_graphs + 1),
                                    cudaMemcpyHostToDevice, stream));
    // tile_count and big_count are adjacent (256-byte aligned carve) -> clear both
    OCTA_CUDA_CHECK(cudaMemsetAsync(w.tile_count, 0, (char*)w.tile_start - (char*)w.tile_count, stream));
    if (n_edges > 0) {
        const int threads = 128;
        const int blocks = (int)((n_edges + threads - 1) / threads);
        vox_prep_kernel<<<blocks, threads, 0, stream>>>(edges7_dev, w.edge_offsets, n_graphs, g, w.prep,
                                                        w.tile_count, w.big_count, w.big_idx);
        octa::count_launch();
    }
    vox_scan_kernel<<<n_graphs, 1024, 0, stream>>>(w.tile_count, w.tile_start, w.tile_cursor, g.ntiles);
    octa::count_launch();
    if (n_edges > 0) {
        const int threads = 128;
        const int blocks = (int)((n_edges + threads - 1) / threads);
        vox_fill_kernel<<<blocks, threads, 0, stream>>>(w.edge_offsets, n_graphs, g, w.prep, w.tile_cursor,
                                                        w.tile_edges);
        octa::count_launch();
    }
    const size_t smem = octa::align_up((size_t)g.T[0] * g.T[1] * g.T[2] * sizeof(uint16_t), 16);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(vox_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(vox_col_kernel<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(vox_col_kernel<256, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // six tiles per SM
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(vox_col_kernel<128, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(vox_col_kernel<128, 12>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        smem_set = smem;
    }
    // OCTA_VOX_KERNEL=rows selects the row-dealing kernel of round 1 (kept for A/B measurements; bit-identical results)
    const char* vk = getenv("OCTA_VOX_KERNEL");
    const bool use_rows = vk && vk[0] == 'r';
    dim3 grid((unsigned)(g.nt[1] * g.nt[2]), (unsigned)g.nt[0], (unsigned)n_graphs);
    if (use_rows)
        vox_tile_kernel<<<grid, VOX_THREADS, smem, stream>>>(w.prep, w.edge_offsets, g, w.tile_start, w.tile_edges,
                                                             w.big_count, w.big_idx, out_dev);
    else if (g.T[0] * g.T[1] <= 128)       // half-size tiles (OCTA_VOX_TILE_Y=8): four warps per tile
        vox_col_kernel<128, 12><<<grid, 128, smem, stream>>>(w.prep, w.edge_offsets, g, w.tile_start, w.tile_edges,
                                                              w.big_count, w.big_idx, out_dev);
    else
        vox_col_kernel<256, 6><<<grid, 256, smem, stream>>>(w.prep, w.edge_offsets, g, w.tile_start, w.tile_edges,
                                                             w.big_count, w.big_idx, out_dev);
    octa::count_launch();
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}

extern "C" int octa_voxelize_host(const double* edges7, int64_t n_edges, const int dims[3], const OctaVoxOpts* opts,
                                  uint16_t* out) {
    OCTA_ARG_CHECK(n_edges >= 0 && out, "bad arguments");
    OCTA_ARG_CHECK(n_edges == 0 || edges7, "edges pointer is null");
    if (octa_device_count() <= 0) {
        octa::set_error("octa_voxelize_host: no CUDA device (there is no CPU fallback)");
        return OCTA_E_CUDA;
    }
    VoxGeom g;
    int rc = make_geom(dims, opts, &g);
    if (rc) return rc;
    const size_t vol_bytes = (size_t)g.D[0] * g.D[1] * g.D[2] * sizeof(uint16_t);
    const size_t ws_bytes = octa_voxelize_workspace_bytes(1, n_edges, dims);
    double* d_edges = nullptr;
    uint16_t* d_out = nullptr;
    void* d_ws = nullptr;
    auto cleanup = [&]() { cudaFree(d_edges); cudaFree(d_out); cudaFree(d_ws); };
    cudaError_t ce;
    if ((ce = cudaMalloc(&d_edges, sizeof(double) * 7 * (size_t)(n_edges ? n_edges : 1))) != cudaSuccess ||
        (ce = cudaMalloc(&d_out, vol_bytes)) != cudaSuccess || (ce = cudaMalloc(&d_ws, ws_bytes)) != cudaSuccess) {
        octa::set_error("octa_voxelize_host: cudaMalloc failed: %s", cudaGetErrorString(ce));
        cleanup();
        return OCTA_E_NOMEM;
    }
    if (n_edges) ce = cudaMemcpy(d_edges, edges7, sizeof(double) * 7 * (size_t)n_edges, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { octa::set_error("H2D failed: %s", cudaGetErrorString(ce)); cleanup(); return OCTA_E_CUDA; }
    const int64_t offs[2] = {0, n_edges};
    rc = octa_voxelize_batch_dev(d_edges, offs, 1, dims, opts, d_out, d_ws, ws_bytes, nullptr);
    if (rc == OCTA_OK) {
        ce = cudaMemcpy(out, d_out, vol_bytes, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) { octa::set_error("D2H failed: %s", cudaGetErrorString(ce)); rc = OCTA_E_CUDA; }
    }
    cleanup();
    return rc;
}
