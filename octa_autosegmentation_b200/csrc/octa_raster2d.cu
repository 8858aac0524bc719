// K7: 2-D anti-aliased label rasterizer for sm_100a.
//
// Replaces vessel_graph_generation/tree2img.py:12-114 (rasterize_forest), whose arithmetic lives in matplotlib's Agg backend:
// every kept edge is a round-capped stroke of width 1.3 * radius * max(W, H) points = * 100/72 pixels (tree2img.py:82-86, dpi
// :51) drawn white on black, pixel (row, col) <-> (pos[ax0]*H, pos[ax1]*W), ax = {0,1,2} \ {MIP_axis} (:46,:85), anti-aliased,
// blended "over" in list order in 8 bits, read back as gray (:104-113).  csrc/octa_aggcells.cuh restates that pipeline stage
// by stage (centre-line clip and snap, inscribed-polygon caps, the 24.8 fixed-point cover/area cells of Agg's scanline
// rasterizer, calculate_alpha, fixed_blender_rgba_plain); the sequential CPU restatement of the same pipeline in the test tree (agg_oracle.c)
// reproduces all 500 label PNGs the reference ships bit for bit, and this kernel equals it pixel for pixel
// (tests/test_raster2d_gpu.py).  All pixel arithmetic is integer.
//
// Design: coverage is a property of ONE stroke, blending is a property of ONE pixel, so the work is split that way.
//   r2d_cover_kernel   one warp per stroke: centre line, outline vertices (lane = vertex, float64 trig), Agg's clipper ->
//                      24.8 integer lines in shared memory; the (line, scanline) pairs are dealt to the lanes, each evaluates
//                      its piece of the line in closed form (Agg's scanline DDA is a floor division) and adds the (cover,
//                      area) cell contributions into the stroke's cell window in shared memory (integer atomics: order
//                      free); lane = row then sweeps its row (running cover -> calculate_alpha) and the stroke leaves as an
//                      8-bit ALPHA MAP over its cell bounding box (<= 32 x 32 cells: 98.7 % of the strokes at 1216^2, all at
//                      304^2; mean 177 bytes) -- or, for larger strokes, as its line list.  It also bins the stroke to tiles.
//   r2d_tile_kernel    tile ownership like K6: one CTA owns a 32x32 pixel tile of one graph, warp = pixel row, lane = column.
//                      The tile's stroke list (ranked back into edge order: blending does not commute) is walked in order;
//                      a stroke with an alpha map costs a row test per warp and one byte load + blend per covered pixel.
//                      Larger strokes are accumulated per row from their lines (cover / area cells of the warp's row in
//                      shared memory, warp scan), outlines with more than 32 lines are evaluated on the fly by the lanes.
// HBM traffic: 56 bytes/edge in, 1 byte/pixel out, + the alpha maps (written once, read by the ~1.6 tiles a stroke touches).
#include "octa_common.h"
#include <math.h>
#include "octa_aggcells.cuh"

namespace {

using namespace octa;

constexpr int RT = 32;          // tile width (pixels) = warp width
constexpr int RTH = 16;         // tile height (pixel rows = warps of the CTA)
constexpr int RKBIG = 64;       // an edge listed in more tiles than this goes to the graph's "big" list, tested by every tile
constexpr int SB = 32;          // strokes per batch of the tile kernel
constexpr int LSTRIDE = 32;     // integer lines a stroke's slot can hold
constexpr int AW = 32;          // alpha maps cover at most AW x AW cells
constexpr int SLOT_BYTES = AW * AW;     // per-stroke slot: alpha map (kind 1) or LSTRIDE lines (kind 2)
constexpr int R2D_THREADS = RT * RTH;
constexpr int COVER_WARPS = 4;  // strokes per CTA of r2d_cover_kernel

enum : int { K_NONE = 0, K_AMAP = 1, K_LINES = 2, K_WIDE = 3 };

struct REdge {                  // clipped / snapped centre line + half width (pixels), tile range, what the slot holds
    agg::Stroke s;
    int lo[2], hi[2];           // tile range, inclusive; hi < lo -> nothing to draw
    int kind;
    int bx0, by0, bw, bh;       // K_AMAP: cell window of the alpha map;  K_LINES / K_WIDE: bx0 = number of lines, by0 / bh = first / last scanline
    int ncap;                   // arc steps per cap
};

struct RGeom {
    int H, W, ntx, nty, ntiles;
    int ax0, ax1;
    double scale, min_radius, max_radius;
};

__device__ __forceinline__ int graph_of(const int64_t* offs, int n_graphs, int64_t i) {
    int lo = 0, hi = n_graphs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

struct CoverSm {                // per warp
    agg::Line lines[agg::MAX_LINES];
    int pre[agg::MAX_LINES + 1];        // exclusive prefix of the scanlines each line crosses
    int cover[AW][AW + 1], area[AW][AW + 1];
    unsigned char amap[AW * AW];
};

struct CellAdd {                // contribution of a line piece to cell ex of row `r` of the stroke's window
    int* cover; int* area; int bx0, bw;
    __device__ __forceinline__ void operator()(int ex, int c, int a) {
        const int x = ex - bx0;
        if (x >= 0 && x < bw) { atomicAdd(cover + x, c); atomicAdd(area + x, a); }     // (cells right of the window: column W, never rendered)
    }
};

__global__ void __launch_bounds__(COVER_WARPS * 32)
r2d_cover_kernel(const double* __restrict__ edges7, const int64_t* __restrict__ offs, int n_graphs, RGeom g,
                 REdge* __restrict__ prep, unsigned char* __restrict__ slots, int* __restrict__ tile_count,
                 int* __restrict__ big_count, int* __restrict__ big_idx, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char cover_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CoverSm& S = reinterpret_cast<CoverSm*>(cover_smem)[warp];
    const int64_t n_edges = offs[n_graphs];
    const int64_t i = (int64_t)blockIdx.x * COVER_WARPS + warp;
    if (i >= n_edges) return;
    const int gr = graph_of(offs, n_graphs, i);
    REdge q;
    q.lo[0] = q.lo[1] = 1; q.hi[0] = q.hi[1] = 0;
    q.s.x0 = q.s.y0 = q.s.x1 = q.s.y1 = q.s.w = 0;
    q.kind = K_NONE; q.bx0 = q.by0 = q.bw = q.bh = 0; q.ncap = 0;
    const bool ok = agg::prepare_stroke(edges7 + 7 * i, g.ax0, g.ax1, g.H, g.W, g.scale, g.min_radius, g.max_radius, &q.s) && (q.s.w == q.s.w);
    if (!ok) { if (lane == 0) prep[i] = q; return; }
    int nc = agg::cap_steps(q.s.w);
    if (nc > agg::MAX_CAP_SEG) { nc = agg::MAX_CAP_SEG; if (lane == 0) atomicExch(err, 1); }      // stroke wider than ~800 px: reported
    if (nc < 0) nc = 0;
    q.ncap = nc;
    const int nv = 2 * (nc + 2);
    const agg::Clip clip = {0.0, 0.0, (double)g.W, (double)g.H};
    // ---- outline -> clipped integer lines (lane = vertex; the edge to the next vertex goes through the clipper)
    int nl_total = 0;
    int cx0 = 0x7fffffff, cx1 = -0x7fffffff, cy0 = 0x7fffffff, cy1 = -0x7fffffff;
    for (int base = 0; base < nv; base += 32) {
        const int k = base + lane;
        double vx = 0, vy = 0, wx, wy;
        if (k < nv) agg::stroke_vertex(q.s, nc, k, &vx, &vy);
        if (nv <= 32) {
            const int nxt = lane + 1 == nv ? 0 : lane + 1;
            wx = __shfl_sync(0xffffffffu, vx, nxt); wy = __shfl_sync(0xffffffffu, vy, nxt);
        } else if (k < nv) {
            agg::stroke_vertex(q.s, nc, k + 1 == nv ? 0 : k + 1, &wx, &wy);
        }
        agg::Line tmp[3];
        int nl = 0;
        if (k < nv) nl = agg::clip_edge(clip, vx, vy, wx, wy, tmp);
        int incl = nl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        for (int t = 0; t < nl; ++t) {
            const int pos = nl_total + incl - nl + t;
            if (pos < agg::MAX_LINES) S.lines[pos] = tmp[t];
            cx0 = min(cx0, min(tmp[t].x1, tmp[t].x2) >> agg::SUB_SHIFT); cx1 = max(cx1, max(tmp[t].x1, tmp[t].x2) >> agg::SUB_SHIFT);
            cy0 = min(cy0, min(tmp[t].y1, tmp[t].y2) >> agg::SUB_SHIFT); cy1 = max(cy1, max(tmp[t].y1, tmp[t].y2) >> agg::SUB_SHIFT);
        }
        nl_total += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (nl_total > agg::MAX_LINES) nl_total = agg::MAX_LINES;      // (a convex outline gains at most 4 lines from the clipper)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx0 = min(cx0, __shfl_xor_sync(0xffffffffu, cx0, o)); cx1 = max(cx1, __shfl_xor_sync(0xffffffffu, cx1, o));
        cy0 = min(cy0, __shfl_xor_sync(0xffffffffu, cy0, o)); cy1 = max(cy1, __shfl_xor_sync(0xffffffffu, cy1, o));
    }
    // rendered cells: columns 0 .. W-1 (the clipper parks everything right of the canvas in column W), rows 0 .. H-1
    cx0 = max(cx0, 0); cy0 = max(cy0, 0); cx1 = min(cx1, g.W - 1); cy1 = min(cy1, g.H - 1);
    if (nl_total == 0 || cx1 < cx0 || cy1 < cy0) { if (lane == 0) prep[i] = q; return; }
    __syncwarp();
    const int bw = cx1 - cx0 + 1, bh = cy1 - cy0 + 1;
    unsigned char* slot = slots + (size_t)i * SLOT_BYTES;
    if (bw <= AW && bh <= AW) {
        // ---- alpha map of the stroke over its cell window
        q.kind = K_AMAP; q.bx0 = cx0; q.by0 = cy0; q.bw = bw; q.bh = bh;
        for (int t = lane; t < bh * (AW + 1); t += 32) { (&S.cover[0][0])[t] = 0; (&S.area[0][0])[t] = 0; }
        // scanlines each line crosses inside the window -> prefix; pairs (line, scanline) dealt to the lanes
        int run = 0;
        for (int base = 0; base < nl_total; base += 32) {
            const int li = base + lane;
            int cnt = 0;
            if (li < nl_total) {
                const int e1 = S.lines[li].y1 >> agg::SUB_SHIFT, e2 = S.lines[li].y2 >> agg::SUB_SHIFT;
                const int a0 = max(min(e1, e2), cy0), a1 = min(max(e1, e2), cy1);
                cnt = a1 >= a0 ? a1 - a0 + 1 : 0;
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            if (li < nl_total) S.pre[li] = run + incl - cnt;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) S.pre[nl_total] = run;
        __syncwarp();
        for (int p = lane; p < run; p += 32) {
            int lo = 0, hi = nl_total;                      // last line with pre <= p
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.pre[mid] <= p) lo = mid; else hi = mid; }
            const agg::Line L = S.lines[lo];
            const int e1 = L.y1 >> agg::SUB_SHIFT, e2 = L.y2 >> agg::SUB_SHIFT;
            const int ey = max(min(e1, e2), cy0) + (p - S.pre[lo]);
            CellAdd add = {S.cover[ey - cy0], S.area[ey - cy0], cx0, bw};
            agg::line_row(L, ey, add);
        }
        __syncwarp();
        // lane = row: running cover along the row -> calculate_alpha (padded rows: no bank conflicts)
        if (lane < bh) {
            int cov = 0;
            for (int x = 0; x < bw; ++x) {
                cov += S.cover[lane][x];
                S.amap[lane * bw + x] = (unsigned char)agg::calc_alpha((cov << (agg::SUB_SHIFT + 1)) - S.area[lane][x]);
            }
        }
        __syncwarp();
        const int nbytes = bw * bh;
        for (int t = lane * 4; t < nbytes; t += 128)        // (slots are 1 KB aligned; the tail of the last word is padding)
            *reinterpret_cast<unsigned int*>(slot + t) = *reinterpret_cast<const unsigned int*>(S.amap + t);
    } else if (nl_total <= LSTRIDE) {
        q.kind = K_LINES; q.bx0 = nl_total; q.by0 = cy0; q.bh = cy1;
        if (lane < nl_total) reinterpret_cast<agg::Line*>(slot)[lane] = S.lines[lane];
    } else {
        q.kind = K_WIDE; q.bx0 = nl_total; q.by0 = cy0; q.bh = cy1;
    }
    if (lane != 0) return;
    // tiles that contain rendered cells of the stroke (a cell's coverage depends on outline edges to its LEFT in the row, all
    // inside the window: tiles from the window's first column on are enough)
    q.lo[0] = cx0 / RT; q.hi[0] = cx1 / RT; q.lo[1] = cy0 / RTH; q.hi[1] = cy1 / RTH;
    prep[i] = q;
    const int n = (q.hi[0] - q.lo[0] + 1) * (q.hi[1] - q.lo[1] + 1);
    if (n > RKBIG) { const int pos = atomicAdd(&big_count[gr], 1); big_idx[offs[gr] + pos] = (int)(i - offs[gr]); return; }
    int* tc = tile_count + (size_t)gr * g.ntiles;
    for (int ty = q.lo[1]; ty <= q.hi[1]; ++ty)
        for (int tx = q.lo[0]; tx <= q.hi[0]; ++tx) atomicAdd(&tc[ty * g.ntx + tx], 1);
}

__global__ void r2d_scan_kernel(const int* __restrict__ tile_count, int* __restrict__ tile_start, int* __restrict__ cursor, int ntiles) {
    __shared__ int ws[32];
    __shared__ int carry;
    const int gr = blockIdx.x;
    const int* cnt = tile_count + (size_t)gr * ntiles;
    int* st = tile_start + (size_t)gr * (ntiles + 1);
    int* cur = cursor + (size_t)gr * ntiles;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < ntiles; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? cnt[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nw ? ws[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            ws[lane] = w;
        }
        __syncthreads();
        const int prefix = carry + (warp ? ws[warp - 1] : 0) + (x - v);
        if (i < ntiles) { st[i] = prefix; cur[i] = prefix; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) st[ntiles] = carry;
}

__global__ void r2d_fill_kernel(const int64_t* __restrict__ offs, int n_graphs, RGeom g, const REdge* __restrict__ prep,
                                int* __restrict__ cursor, int* __restrict__ tile_edges) {
    const int64_t n_edges = offs[n_graphs];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const int gr = graph_of(offs, n_graphs, i);
    const REdge& q = prep[i];
    const int lo0 = q.lo[0], hi0 = q.hi[0], lo1 = q.lo[1], hi1 = q.hi[1];
    if (hi0 < lo0 || hi1 < lo1) return;
    const int n = (hi0 - lo0 + 1) * (hi1 - lo1 + 1);
    if (n > RKBIG) return;
    int* cur = cursor + (size_t)gr * g.ntiles;
    int* lst = tile_edges + (size_t)RKBIG * offs[gr];
    const int local = (int)(i - offs[gr]);
    for (int ty = lo1; ty <= hi1; ++ty)
        for (int tx = lo0; tx <= hi0; ++tx) lst[atomicAdd(&cur[ty * g.ntx + tx], 1)] = local;
}

// shared-memory state of one tile
struct TileSm {
    int list_in[R2D_THREADS], list_sorted[R2D_THREADS];     // the tile's strokes, as filled / in edge order
    agg::Line lines[SB][LSTRIDE];                           // K_LINES strokes of the batch
    int kind[SB], b0[SB], b1[SB], b2[SB], b3[SB];           // kind + (bx0, by0, bw, bh) or (nlines, first row, -, last row)
    int cover[RTH][RT], area[RTH][RT], left[RTH];
    int batch_ids[SB], batch_n;
    int n_sorted;
};

struct RowAdd {     // (cover, area) contribution of a line piece to cell ex of this warp's row
    int* cover; int* area; int* left; int tx0;
    __device__ __forceinline__ void operator()(int ex, int c, int a) {
        if (ex < tx0) { if (c) atomicAdd(left, c); }
        else if (ex < tx0 + RT) { atomicAdd(cover + (ex - tx0), c); atomicAdd(area + (ex - tx0), a); }
    }
};

// strokes without an alpha map (window larger than AW x AW cells: 1.3 % of the strokes at 1216^2): the warp accumulates the
// cover / area cells of ITS row from the stroke's lines (outlines with more than LSTRIDE lines: the lanes compute the outline
// edges they need themselves), scans the covers and returns the pixel's coverage
__device__ __noinline__ int row_alpha_from_lines(TileSm& T, const REdge& e, int s, int kind, int row, int lane, int py, int tx0, int H, int W) {
    RowAdd add = {T.cover[row], T.area[row], &T.left[row], tx0};
    if (kind == K_LINES) {
        if (lane < T.b0[s]) agg::line_row(T.lines[s][lane], py, add);
    } else {
        const agg::Clip clip = {0.0, 0.0, (double)W, (double)H};
        const int nv = 2 * (e.ncap + 2);
        for (int k = lane; k < nv; k += 32) {
            double ax, ay, bx, by;
            agg::stroke_vertex(e.s, e.ncap, k, &ax, &ay);
            agg::stroke_vertex(e.s, e.ncap, k + 1 == nv ? 0 : k + 1, &bx, &by);
            agg::Line tmp[3];
            const int m = agg::clip_edge(clip, ax, ay, bx, by, tmp);
            for (int q = 0; q < m; ++q) agg::line_row(tmp[q], py, add);
        }
    }
    __syncwarp();
    const int c = T.cover[row][lane], a = T.area[row][lane], l = T.left[row];
    __syncwarp();
    T.cover[row][lane] = 0; T.area[row][lane] = 0;
    if (lane == 0) T.left[row] = 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    __syncwarp();
    return agg::calc_alpha(((l + incl) << (agg::SUB_SHIFT + 1)) - a);
}

__global__ void __launch_bounds__(R2D_THREADS, 3)
r2d_tile_kernel(const REdge* __restrict__ prep, const unsigned char* __restrict__ slots, const int64_t* __restrict__ offs, RGeom g,
                const int* __restrict__ tile_start, int* __restrict__ tile_edges, const int* __restrict__ big_count,
                const int* __restrict__ big_idx, uint8_t* __restrict__ out, const int64_t* __restrict__ layer_split) {
    extern __shared__ __align__(16) unsigned char r2d_smem[];
    TileSm& T = *reinterpret_cast<TileSm*>(r2d_smem);
    const int gr = blockIdx.y, tile = blockIdx.x;
    const int tx = tile % g.ntx, ty = tile / g.ntx;
    const int tid = threadIdx.x, lane = tid & 31, row = tid >> 5;
    const int64_t eb = offs[gr];
    const REdge* ge = prep + eb;
    const int* st = tile_start + (size_t)gr * (g.ntiles + 1);
    const int beg = st[tile], end = st[tile + 1];
    int* lst = tile_edges + (size_t)RKBIG * eb;
    const int* bl = big_idx + eb;
    const int nbig = big_count[gr];
    const int n = end - beg;
    const int py = ty * RTH + row, px = tx * RT + lane;
    if (n == 0 && nbig == 0) {                      // empty tile: black
        if (px < g.W && py < g.H) out[((size_t)gr * g.H + py) * g.W + px] = 0;
        return;
    }
    // ---- the tile's strokes in edge order.  Usual case: the list (+ the big edges that overlap the tile) fits one entry per
    // thread and is ranked in shared memory (ids are distinct: rank = number of smaller ids).
    T.cover[row][lane] = 0; T.area[row][lane] = 0;
    if (lane == 0) T.left[row] = 0;
    if (tid == 0) T.n_sorted = 0;
    __syncthreads();
    const bool in_smem = n + nbig <= R2D_THREADS;
    if (in_smem) {
        int v = -1;
        if (tid < n) v = lst[beg + tid];
        else if (tid < n + nbig) {
            const int b = bl[tid - n];
            const REdge& e = ge[b];
            if (!(tx < e.lo[0] || tx > e.hi[0] || ty < e.lo[1] || ty > e.hi[1])) v = b;
        }
        T.list_in[tid] = v;
        __syncthreads();
        if (v >= 0) {
            int rank = 0;
            for (int j = 0; j < n + nbig; ++j) { const int u = T.list_in[j]; rank += (u >= 0 && u < v); }
            T.list_sorted[rank] = v;
            atomicAdd(&T.n_sorted, 1);
        }
        __syncthreads();
    } else {
        // more strokes in one tile than threads: odd-even transposition of the list in place (global memory)
        for (int pass = 0; pass < n; ++pass) {
            for (int i = (pass & 1) + 2 * tid; i + 1 < n; i += 2 * blockDim.x) {
                const int a = lst[beg + i], b = lst[beg + i + 1];
                if (a > b) { lst[beg + i] = b; lst[beg + i + 1] = a; }
            }
            __syncthreads();
        }
    }
    const int n_sorted = in_smem ? T.n_sorted : n;
    // Two layers (generate_vessel_graph.py:80-85: the arterial and the venous forest are rasterized separately and combined with
    // np.maximum): edges with a local index >= split start a fresh canvas; the pixel leaves as the maximum of both.
    const int split = layer_split ? (int)layer_split[gr] : 0x7fffffff;
    unsigned char gray = 0, gray_first = 0;
    bool second = false;
    int head = 0, bprev = -1;        // (global path) next list entry / last big edge taken
    while (true) {
        // ---- batch assembly
        if (in_smem) {
            if (tid == 0) T.batch_n = min(SB, n_sorted - head);
            if (tid < SB && head + tid < n_sorted) T.batch_ids[tid] = T.list_sorted[head + tid];
            head += SB;
        } else if (tid == 0) {
            // merge of the sorted list with the (unsorted, tiny) big list, filtered by tile overlap
            int k = 0;
            while (k < SB) {
                int bbest = 0x7fffffff;
                for (int j = 0; j < nbig; ++j) {
                    const int v = bl[j];
                    if (v > bprev && v < bbest) {
                        const REdge& e = ge[v];
                        if (!(tx < e.lo[0] || tx > e.hi[0] || ty < e.lo[1] || ty > e.hi[1])) bbest = v;
                    }
                }
                const int lbest = head < n ? lst[beg + head] : 0x7fffffff;
                if (bbest == 0x7fffffff && lbest == 0x7fffffff) break;
                if (lbest < bbest) { T.batch_ids[k++] = lbest; ++head; } else { T.batch_ids[k++] = bbest; bprev = bbest; }
            }
            T.batch_n = k;
        }
        __syncthreads();
        const int bn = T.batch_n;
        if (bn <= 0) break;
        // ---- the batch's stroke records (and the line lists of the larger strokes) -> shared memory: thread = (stroke, slot)
        for (int it = tid; it < SB * LSTRIDE; it += R2D_THREADS) {
            const int s = it >> 5, l = it & 31;
            if (s < bn) {
                const int id = T.batch_ids[s];
                const REdge& e = ge[id];
                const int kind = e.kind;
                if (kind == K_LINES && l < e.bx0) T.lines[s][l] = reinterpret_cast<const agg::Line*>(slots + (size_t)(eb + id) * SLOT_BYTES)[l];
                if (l == 0) { T.kind[s] = kind; T.b0[s] = e.bx0; T.b1[s] = e.by0; T.b2[s] = e.bw; T.b3[s] = e.bh; }
            }
        }
        __syncthreads();
        // ---- warp = pixel row, lane = column.  Lane s first tests stroke s against the warp's row (SB == 32): the warp then
        // visits only the strokes that touch its row, in order.
        int sid = 0x7fffffff, skind = K_NONE, sb0 = 0, sb1 = 0, sb2 = 0, sb3 = 0;
        if (lane < bn) { sid = T.batch_ids[lane]; skind = T.kind[lane]; sb0 = T.b0[lane]; sb1 = T.b1[lane]; sb2 = T.b2[lane]; sb3 = T.b3[lane]; }
        const bool touches = skind == K_AMAP ? (py >= sb1 && py < sb1 + sb3) : (skind != K_NONE && py >= sb1 && py <= sb3);
        unsigned int todo = __ballot_sync(0xffffffffu, touches);
        const unsigned int late = __ballot_sync(0xffffffffu, lane < bn && sid >= split);       // strokes of the second layer
        const int s_split = late ? __ffs(late) - 1 : 32;
        while (todo) {
            const int s = __ffs(todo) - 1;
            todo &= todo - 1;
            if (!second && s >= s_split) { second = true; gray_first = gray; gray = 0; }
            const int kind = __shfl_sync(0xffffffffu, skind, s), id = __shfl_sync(0xffffffffu, sid, s);
            const int x0 = __shfl_sync(0xffffffffu, sb0, s), y0 = __shfl_sync(0xffffffffu, sb1, s), bw = __shfl_sync(0xffffffffu, sb2, s);
            int alpha = 0;
            if (kind == K_AMAP) {
                const int rx = px - x0;
                if (rx >= 0 && rx < bw) alpha = slots[(size_t)(eb + id) * SLOT_BYTES + (py - y0) * bw + rx];
            } else {
                alpha = row_alpha_from_lines(T, ge[id], s, kind, row, lane, py, tx * RT, g.H, g.W);
            }
            gray = agg::blend_cover(gray, alpha);
        }
        if (!second && s_split < 32) { second = true; gray_first = gray; gray = 0; }
        __syncthreads();
    }
    if (px < g.W && py < g.H) out[((size_t)gr * g.H + py) * g.W + px] = gray > gray_first ? gray : gray_first;
}

int make_rgeom(int H, int W, int mip_axis, const OctaVoxOpts* opts, RGeom* g) {
    OCTA_ARG_CHECK(H > 0 && W > 0 && H <= 16384 && W <= 16384, "bad image resolution");
    OCTA_ARG_CHECK(mip_axis >= 0 && mip_axis <= 2, "MIP axis must be 0, 1 or 2");
    g->H = H; g->W = W;
    g->ntx = (W + RT - 1) / RT; g->nty = (H + RTH - 1) / RTH; g->ntiles = g->ntx * g->nty;
    int ax[2], k = 0;
    for (int a = 0; a < 3; ++a) if (a != mip_axis) ax[k++] = a;                     // tree2img.py:46
    g->ax0 = ax[0]; g->ax1 = ax[1];
    g->scale = (double)(W > H ? W : H);                                            // :50
    g->min_radius = opts ? opts->min_radius : 0.0;
    g->max_radius = opts ? opts->max_radius : 1.0;
    return OCTA_OK;
}

struct RWork { REdge* prep; unsigned char* slots; int64_t* offs; int64_t* split; int *tile_count, *big_count, *err, *tile_start, *cursor, *big_idx, *tile_edges; size_t bytes; };

RWork rcarve(void* base, int n_graphs, int64_t n_edges, int ntiles) {
    RWork w;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = octa::align_up(off + b, 256); return (char*)base + o; };
    const size_t ne = (size_t)(n_edges > 0 ? n_edges : 1);
    w.prep = (REdge*)take(sizeof(REdge) * ne);
    w.slots = (unsigned char*)take((size_t)SLOT_BYTES * ne);
    w.offs = (int64_t*)take(sizeof(int64_t) * (n_graphs + 1));
    w.split = (int64_t*)take(sizeof(int64_t) * n_graphs);
    w.tile_count = (int*)take(sizeof(int) * (size_t)n_graphs * ntiles);
    w.big_count = (int*)take(sizeof(int) * n_graphs);
    w.err = (int*)take(sizeof(int));
    w.tile_start = (int*)take(sizeof(int) * (size_t)n_graphs * (ntiles + 1));
    w.cursor = (int*)take(sizeof(int) * (size_t)n_graphs * ntiles);
    w.big_idx = (int*)take(sizeof(int) * ne);
    w.tile_edges = (int*)take(sizeof(int) * RKBIG * ne);
    w.bytes = off;
    return w;
}

}  // namespace

extern "C" size_t octa_raster2d_workspace_bytes(int n_graphs, int64_t n_edges, int H, int W) {
    RGeom g;
    if (n_graphs <= 0 || n_edges < 0 || make_rgeom(H, W, 2, nullptr, &g)) return 0;
    return rcarve(nullptr, n_graphs, n_edges, g.ntiles).bytes;
}

static int raster2d_batch(const double* edges7_dev, const int64_t* edge_offsets_host, const int64_t* layer_split_host, int n_graphs,
                          int H, int W, int mip_axis, const OctaVoxOpts* opts, uint8_t* out_dev, void* workspace_dev,
                          size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    OCTA_ARG_CHECK(n_graphs > 0 && n_graphs <= 65535 && edge_offsets_host && out_dev && workspace_dev, "bad arguments");
    OCTA_ARG_CHECK(edge_offsets_host[0] == 0, "edge_offsets[0] must be 0");
    for (int i = 0; i < n_graphs; ++i) OCTA_ARG_CHECK(edge_offsets_host[i + 1] >= edge_offsets_host[i], "edge_offsets must be non-decreasing");
    const int64_t n_edges = edge_offsets_host[n_graphs];
    OCTA_ARG_CHECK(n_edges == 0 || edges7_dev, "edges pointer is null");
    RGeom g;
    int rc = make_rgeom(H, W, mip_axis, opts, &g);
    if (rc) return rc;
    RWork w = rcarve(workspace_dev, n_graphs, n_edges, g.ntiles);
    if (w.bytes > workspace_bytes) { octa::set_error("octa_raster2d_batch_dev: workspace too small (%zu < %zu)", workspace_bytes, w.bytes); return OCTA_E_NOMEM; }
    OCTA_CUDA_CHECK(cudaMemcpyAsync(w.offs, edge_offsets_host, sizeof(int64_t) * (n_graphs + 1), cudaMemcpyHostToDevice, stream));
    if (layer_split_host) {
        for (int i = 0; i < n_graphs; ++i)
            OCTA_ARG_CHECK(layer_split_host[i] >= 0 && layer_split_host[i] <= edge_offsets_host[i + 1] - edge_offsets_host[i], "layer_split out of range");
        OCTA_CUDA_CHECK(cudaMemcpyAsync(w.split, layer_split_host, sizeof(int64_t) * n_graphs, cudaMemcpyHostToDevice, stream));
    }
    OCTA_CUDA_CHECK(cudaMemsetAsync(w.tile_count, 0, (char*)w.tile_start - (char*)w.tile_count, stream));      // counts, big counts, error flag
    const int threads = 128, blocks = (int)((n_edges + threads - 1) / threads);
    if (n_edges > 0) {
        static bool cover_attr = false;
        if (!cover_attr) {
            OCTA_CUDA_CHECK(cudaFuncSetAttribute(r2d_cover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(CoverSm) * COVER_WARPS)));
            cover_attr = true;
        }
        const int64_t pblocks = (n_edges + COVER_WARPS - 1) / COVER_WARPS;      // one warp per edge
        r2d_cover_kernel<<<(unsigned)pblocks, COVER_WARPS * 32, sizeof(CoverSm) * COVER_WARPS, stream>>>(
            edges7_dev, w.offs, n_graphs, g, w.prep, w.slots, w.tile_count, w.big_count, w.big_idx, w.err);
        octa::count_launch();
    }
    r2d_scan_kernel<<<n_graphs, 1024, 0, stream>>>(w.tile_count, w.tile_start, w.cursor, g.ntiles);
    octa::count_launch();
    if (n_edges > 0) { r2d_fill_kernel<<<blocks, threads, 0, stream>>>(w.offs, n_graphs, g, w.prep, w.cursor, w.tile_edges); octa::count_launch(); }
    static bool attr_set = false;
    if (!attr_set) {
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(r2d_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSm)));
        attr_set = true;
    }
    r2d_tile_kernel<<<dim3((unsigned)g.ntiles, (unsigned)n_graphs), R2D_THREADS, sizeof(TileSm), stream>>>(
        w.prep, w.slots, w.offs, g, w.tile_start, w.tile_edges, w.big_count, w.big_idx, out_dev, layer_split_host ? w.split : nullptr);
    octa::count_launch();
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}

extern "C" int octa_raster2d_batch_dev(const double* edges7_dev, const int64_t* edge_offsets_host, int n_graphs, int H, int W,
                                       int mip_axis, const OctaVoxOpts* opts, uint8_t* out_dev, void* workspace_dev,
                                       size_t workspace_bytes, void* stream_) {
    return raster2d_batch(edges7_dev, edge_offsets_host, nullptr, n_graphs, H, W, mip_axis, opts, out_dev, workspace_dev, workspace_bytes, stream_);
}

extern "C" int octa_raster2d_batch_layers_dev(const double* edges7_dev, const int64_t* edge_offsets_host, const int64_t* layer_split_host,
                                              int n_graphs, int H, int W, int mip_axis, const OctaVoxOpts* opts, uint8_t* out_dev,
                                              void* workspace_dev, size_t workspace_bytes, void* stream_) {
    OCTA_ARG_CHECK(layer_split_host, "layer_split is null");
    return raster2d_batch(edges7_dev, edge_offsets_host, layer_split_host, n_graphs, H, W, mip_axis, opts, out_dev, workspace_dev, workspace_bytes, stream_);
}

extern "C" int octa_raster2d_host(const double* edges7, int64_t n_edges, int H, int W, int mip_axis, const OctaVoxOpts* opts,
                                  uint8_t* out) {
    OCTA_ARG_CHECK(n_edges >= 0 && out && (n_edges == 0 || edges7), "bad arguments");
    if (octa_device_count() <= 0) { octa::set_error("octa_raster2d_host: no CUDA device (there is no CPU fallback)"); return OCTA_E_CUDA; }
    const size_t ws = octa_raster2d_workspace_bytes(1, n_edges, H, W);
    if (!ws) { octa::set_error("octa_raster2d_host: bad image resolution"); return OCTA_E_ARG; }
    double* d_e = nullptr; uint8_t* d_o = nullptr; void* d_w = nullptr;
    auto cleanup = [&]() { cudaFree(d_e); cudaFree(d_o); cudaFree(d_w); };
    cudaError_t ce;
    if ((ce = cudaMalloc(&d_e, sizeof(double) * 7 * (size_t)(n_edges ? n_edges : 1))) != cudaSuccess ||
        (ce = cudaMalloc(&d_o, (size_t)H * W)) != cudaSuccess || (ce = cudaMalloc(&d_w, ws)) != cudaSuccess) {
        octa::set_error("octa_raster2d_host: cudaMalloc failed: %s", cudaGetErrorString(ce)); cleanup(); return OCTA_E_NOMEM;
    }
    if (n_edges && (ce = cudaMemcpy(d_e, edges7, sizeof(double) * 7 * (size_t)n_edges, cudaMemcpyHostToDevice)) != cudaSuccess) {
        octa::set_error("H2D failed: %s", cudaGetErrorString(ce)); cleanup(); return OCTA_E_CUDA;
    }
    const int64_t offs[2] = {0, n_edges};
    int rc = octa_raster2d_batch_dev(d_e, offs, 1, H, W, mip_axis, opts, d_o, d_w, ws, nullptr);
    if (rc == OCTA_OK && (ce = cudaMemcpy(out, d_o, (size_t)H * W, cudaMemcpyDeviceToHost)) != cudaSuccess) {
        octa::set_error("D2H failed: %s", cudaGetErrorString(ce)); rc = OCTA_E_CUDA;
    }
    if (rc == OCTA_OK) {        // strokes wider than the outline buffers allow (half width > ~400 px) are reported, never silently approximated
        const RGeom* unused = nullptr; (void)unused;
        RGeom g; make_rgeom(H, W, mip_axis, opts, &g);
        RWork w = rcarve(d_w, 1, n_edges, g.ntiles);
        int flag = 0;
        if (cudaMemcpy(&flag, w.err, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess && flag) {
            octa::set_error("octa_raster2d_host: a stroke is wider than the rasterizer supports (half width > 400 px)"); rc = OCTA_E_ARG;
        }
    }
    cleanup();
    return rc;
}

// ---- CPU-callable restatement of the tile algorithm (same header code, sequential): test hook for the `-m "not gpu"` suite.
// Rasterizes ONE graph on the host exactly the way r2d_tile_kernel does (per tile, per stroke, per row, closed-form row pieces).
extern "C" int octa_test_raster2d_rows_host(const double* edges7, int64_t n_edges, int H, int W, int mip_axis, double min_radius,
                                            double max_radius, uint8_t* out) {
    OctaVoxOpts o = {min_radius, max_radius, 0, 0};
    RGeom g;
    int rc = make_rgeom(H, W, mip_axis, &o, &g);
    if (rc) return rc;
    for (size_t i = 0; i < (size_t)H * W; ++i) out[i] = 0;
    const agg::Clip clip = {0.0, 0.0, (double)W, (double)H};
    static double vx[agg::MAX_VERT], vy[agg::MAX_VERT];
    static agg::Line lines[agg::MAX_LINES];
    struct HostAdd {
        int* cover; int* area; int* left; int x0, x1;
        void operator()(int ex, int c, int a) {
            if (ex < x0) *left += c;
            else if (ex < x1) { cover[ex - x0] += c; area[ex - x0] += a; }
        }
    };
    int* cover = new int[W + 1]; int* area = new int[W + 1];
    for (int64_t e = 0; e < n_edges; ++e) {
        agg::Stroke s;
        if (!agg::prepare_stroke(edges7 + 7 * e, g.ax0, g.ax1, H, W, g.scale, min_radius, max_radius, &s)) continue;
        int nc = agg::cap_steps(s.w);
        if (nc > agg::MAX_CAP_SEG) { delete[] cover; delete[] area; return OCTA_E_ARG; }
        const int nv = 2 * (nc + 2);
        for (int k = 0; k < nv; ++k) agg::stroke_vertex(s, nc, k, &vx[k], &vy[k]);
        int nl = 0, ymin = 0x7fffffff, ymax = -0x7fffffff;
        for (int k = 0; k < nv; ++k) {
            const int k2 = k + 1 == nv ? 0 : k + 1;
            const int m = agg::clip_edge(clip, vx[k], vy[k], vx[k2], vy[k2], lines + nl);
            for (int q = 0; q < m; ++q) {
                const int e1 = lines[nl + q].y1 >> agg::SUB_SHIFT, e2 = lines[nl + q].y2 >> agg::SUB_SHIFT;
                ymin = e1 < ymin ? e1 : ymin; ymin = e2 < ymin ? e2 : ymin; ymax = e1 > ymax ? e1 : ymax; ymax = e2 > ymax ? e2 : ymax;
            }
            nl += m;
        }
        // split the row into 32-pixel tiles like the kernel does (left = cover of everything to the left of the tile)
        for (int py = ymin < 0 ? 0 : ymin; py <= ymax && py < H; ++py)
            for (int tx0 = 0; tx0 < W; tx0 += RT) {
                int left = 0;
                for (int x = 0; x < RT; ++x) cover[x] = area[x] = 0;
                HostAdd add = {cover, area, &left, tx0, tx0 + RT};
                for (int li = 0; li < nl; ++li) agg::line_row(lines[li], py, add);
                int run = left;
                for (int x = 0; x < RT && tx0 + x < W; ++x) {
                    run += cover[x];
                    uint8_t* p = out + (size_t)py * W + tx0 + x;
                    *p = agg::blend_cover(*p, agg::calc_alpha((run << (agg::SUB_SHIFT + 1)) - area[x]));
                }
            }
    }
    delete[] cover; delete[] area;
    return OCTA_OK;
}
