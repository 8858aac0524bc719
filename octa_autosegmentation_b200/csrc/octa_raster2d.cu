// K7: 2-D anti-aliased label rasterizer for sm_100a.
//
// Replaces vessel_graph_generation/tree2img.py:12-114 (rasterize_forest), whose arithmetic lives in matplotlib's Agg backend:
// every kept edge is a round-capped stroke of width 1.3 * radius * max(W, H) points = * 100/72 pixels (tree2img.py:82-86, dpi
// :51) drawn white on black, pixel (row, col) <-> (pos[ax0]*H, pos[ax1]*W), ax = {0,1,2} \ {MIP_axis} (:46,:85), anti-aliased,
// blended "over" in list order in 8 bits, read back as gray (:104-113).  csrc/octa_aggcells.cuh restates that pipeline stage
// by stage (centre-line clip and snap, inscribed-polygon caps, the 24.8 fixed-point cover/area cells of Agg's scanline
// rasterizer, calculate_alpha, fixed_blender_rgba_plain); the CPU restatement of the same pipeline, oracle/agg_oracle.c,
// reproduces all 500 label PNGs the reference ships bit for bit, and this kernel equals it pixel for pixel
// (tests/test_raster2d_gpu.py).  All pixel arithmetic is integer.
//
// Design: tile ownership like K6.  One CTA owns a 32x32 pixel tile of one graph; its stroke list (binned by prep / scan /
// fill, ranked back into edge order: blending does not commute) is processed in batches of SB strokes:
//   1. outline vertices of the batch (float64 trig, one thread per vertex) -> shared memory
//   2. one thread per outline edge: Agg's clipper -> 24.8 integer lines in shared memory
//   3. ONE WARP PER PIXEL ROW walks the batch's strokes in order: the lanes take the stroke's lines, evaluate the piece of each
//      line inside the row in closed form (Agg's scanline DDA is a floor division) and add its (cover, area) cell
//      contributions with shared-memory atomics (integer sums: order-free); a warp scan of the covers gives every pixel's
//      coverage, lane = column blends its own pixel.  Rows never interact, so there is no barrier inside a batch.
// HBM traffic is the algorithmic 1 byte/pixel + 56 bytes/edge (+ the tile lists).
#include "octa_common.h"
#include <math.h>
#include "octa_aggcells.cuh"

namespace {

using namespace octa;

constexpr int RT = 32;          // tile edge (pixels) = warp width
constexpr int RKBIG = 64;       // an edge listed in more tiles than this goes to the graph's "big" list, tested by every tile
constexpr int SB = 8;           // strokes per batch
constexpr int R2D_THREADS = RT * RT;

struct REdge {                  // clipped / snapped centre line + half width (pixels), tile range
    agg::Stroke s;
    int lo[2], hi[2];           // tile range, inclusive; hi < lo -> nothing to draw
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

__global__ void r2d_prep_kernel(const double* __restrict__ edges7, const int64_t* __restrict__ offs, int n_graphs, RGeom g,
                                REdge* __restrict__ prep, int* __restrict__ tile_count, int* __restrict__ big_count,
                                int* __restrict__ big_idx) {
    const int64_t n_edges = offs[n_graphs];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const int gr = graph_of(offs, n_graphs, i);
    REdge q;
    q.lo[0] = q.lo[1] = 1; q.hi[0] = q.hi[1] = 0;
    q.s.x0 = q.s.y0 = q.s.x1 = q.s.y1 = q.s.w = 0;
    if (!agg::prepare_stroke(edges7 + 7 * i, g.ax0, g.ax1, g.H, g.W, g.scale, g.min_radius, g.max_radius, &q.s) || !(q.s.w == q.s.w)) {
        prep[i] = q;
        return;
    }
    // the outline is inscribed in the capsule of half width w; 1/64 px of slack covers the 1/256 vertex rounding
    const double reach = q.s.w + 0.015625;
    const double bx0 = fmin(q.s.x0, q.s.x1) - reach, bx1 = fmax(q.s.x0, q.s.x1) + reach;
    const double by0 = fmin(q.s.y0, q.s.y1) - reach, by1 = fmax(q.s.y0, q.s.y1) + reach;
    const int px0 = (int)fmax(0.0, floor(bx0)), px1 = (int)fmin((double)g.W - 1, floor(bx1));
    const int py0 = (int)fmax(0.0, floor(by0)), py1 = (int)fmin((double)g.H - 1, floor(by1));
    if (!(bx1 >= 0) || !(by1 >= 0) || px0 > px1 || py0 > py1) { prep[i] = q; return; }
    // a cell's coverage depends on every outline edge to its LEFT in the row, all of which lie inside the bounding box: tiles
    // from the box's first column on are enough
    q.lo[0] = px0 / RT; q.hi[0] = px1 / RT; q.lo[1] = py0 / RT; q.hi[1] = py1 / RT;
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
    const REdge q = prep[i];
    if (q.hi[0] < q.lo[0] || q.hi[1] < q.lo[1]) return;
    const int n = (q.hi[0] - q.lo[0] + 1) * (q.hi[1] - q.lo[1] + 1);
    if (n > RKBIG) return;
    int* cur = cursor + (size_t)gr * g.ntiles;
    int* lst = tile_edges + (size_t)RKBIG * offs[gr];
    const int local = (int)(i - offs[gr]);
    for (int ty = q.lo[1]; ty <= q.hi[1]; ++ty)
        for (int tx = q.lo[0]; tx <= q.hi[0]; ++tx) lst[atomicAdd(&cur[ty * g.ntx + tx], 1)] = local;
}

// shared-memory state of one tile
struct TileSm {
    int list_in[R2D_THREADS], list_sorted[R2D_THREADS];     // the tile's strokes, as filled / in edge order
    double vx[SB][agg::MAX_VERT], vy[SB][agg::MAX_VERT];
    agg::Line lines[SB][agg::MAX_LINES];
    int nvert[SB], ncap[SB], nlines[SB], ymin[SB], ymax[SB];
    agg::Stroke stroke[SB];
    int cover[RT][RT], area[RT][RT], left[RT];
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

__global__ void __launch_bounds__(R2D_THREADS)
r2d_tile_kernel(const REdge* __restrict__ prep, const int64_t* __restrict__ offs, RGeom g, const int* __restrict__ tile_start,
                int* __restrict__ tile_edges, const int* __restrict__ big_count, const int* __restrict__ big_idx,
                uint8_t* __restrict__ out, int* __restrict__ err) {
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
        // more than 1024 strokes in one tile: odd-even transposition of the list in place (global memory)
        for (int pass = 0; pass < n; ++pass) {
            for (int i = (pass & 1) + 2 * tid; i + 1 < n; i += 2 * blockDim.x) {
                const int a = lst[beg + i], b = lst[beg + i + 1];
                if (a > b) { lst[beg + i] = b; lst[beg + i + 1] = a; }
            }
            __syncthreads();
        }
    }
    const int n_sorted = in_smem ? T.n_sorted : n;
    unsigned char gray = 0;
    const int py = ty * RT + row, px = tx * RT + lane;
    const agg::Clip clip = {0.0, 0.0, (double)g.W, (double)g.H};
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
        // ---- 1. stroke parameters and outline vertices
        if (tid < bn) {
            const agg::Stroke s = ge[T.batch_ids[tid]].s;
            int nc = agg::cap_steps(s.w);
            if (nc > agg::MAX_CAP_SEG) { nc = agg::MAX_CAP_SEG; atomicExch(err, 1); }     // stroke wider than ~800 px: reported
            if (nc < 0) nc = 0;
            T.stroke[tid] = s; T.ncap[tid] = nc; T.nvert[tid] = 2 * (nc + 2);
            T.nlines[tid] = 0; T.ymin[tid] = 0x7fffffff; T.ymax[tid] = -0x7fffffff;
        }
        __syncthreads();
        for (int it = tid; it < bn * agg::MAX_VERT; it += R2D_THREADS) {
            const int s = it / agg::MAX_VERT, k = it - s * agg::MAX_VERT;
            if (k < T.nvert[s]) agg::stroke_vertex(T.stroke[s], T.ncap[s], k, &T.vx[s][k], &T.vy[s][k]);
        }
        __syncthreads();
        // ---- 2. outline edges through Agg's clipper -> integer lines
        for (int it = tid; it < bn * agg::MAX_VERT; it += R2D_THREADS) {
            const int s = it / agg::MAX_VERT, k = it - s * agg::MAX_VERT;
            const int nv = T.nvert[s];
            if (k >= nv) continue;
            const int k2 = k + 1 == nv ? 0 : k + 1;
            agg::Line tmp[3];
            const int nl = agg::clip_edge(clip, T.vx[s][k], T.vy[s][k], T.vx[s][k2], T.vy[s][k2], tmp);
            if (nl > 0) {
                const int pos = atomicAdd(&T.nlines[s], nl);
                for (int q = 0; q < nl; ++q) {
                    if (pos + q < agg::MAX_LINES) T.lines[s][pos + q] = tmp[q];
                    const int e1 = tmp[q].y1 >> agg::SUB_SHIFT, e2 = tmp[q].y2 >> agg::SUB_SHIFT;
                    atomicMin(&T.ymin[s], min(e1, e2));
                    atomicMax(&T.ymax[s], max(e1, e2));
                }
            }
        }
        __syncthreads();
        // ---- 3. one warp per pixel row
        for (int s = 0; s < bn; ++s) {
            if (py < T.ymin[s] || py > T.ymax[s]) continue;
            const int nl = min(T.nlines[s], agg::MAX_LINES);
            RowAdd add = {T.cover[row], T.area[row], &T.left[row], tx * RT};
            for (int li = lane; li < nl; li += 32) agg::line_row(T.lines[s][li], py, add);
            __syncwarp();
            const int c = T.cover[row][lane], a = T.area[row][lane], l = T.left[row];
            __syncwarp();
            T.cover[row][lane] = 0; T.area[row][lane] = 0;
            if (lane == 0) T.left[row] = 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const int alpha = agg::calc_alpha(((l + incl) << (agg::SUB_SHIFT + 1)) - a);
            gray = agg::blend_cover(gray, alpha);
            __syncwarp();
        }
        __syncthreads();
    }
    if (px < g.W && py < g.H) out[((size_t)gr * g.H + py) * g.W + px] = gray;
}

int make_rgeom(int H, int W, int mip_axis, const OctaVoxOpts* opts, RGeom* g) {
    OCTA_ARG_CHECK(H > 0 && W > 0 && H <= 16384 && W <= 16384, "bad image resolution");
    OCTA_ARG_CHECK(mip_axis >= 0 && mip_axis <= 2, "MIP axis must be 0, 1 or 2");
    g->H = H; g->W = W;
    g->ntx = (W + RT - 1) / RT; g->nty = (H + RT - 1) / RT; g->ntiles = g->ntx * g->nty;
    int ax[2], k = 0;
    for (int a = 0; a < 3; ++a) if (a != mip_axis) ax[k++] = a;                     // tree2img.py:46
    g->ax0 = ax[0]; g->ax1 = ax[1];
    g->scale = (double)(W > H ? W : H);                                            // :50
    g->min_radius = opts ? opts->min_radius : 0.0;
    g->max_radius = opts ? opts->max_radius : 1.0;
    return OCTA_OK;
}

struct RWork { REdge* prep; int64_t* offs; int *tile_count, *big_count, *err, *tile_start, *cursor, *big_idx, *tile_edges; size_t bytes; };

RWork rcarve(void* base, int n_graphs, int64_t n_edges, int ntiles) {
    RWork w;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = octa::align_up(off + b, 256); return (char*)base + o; };
    const size_t ne = (size_t)(n_edges > 0 ? n_edges : 1);
    w.prep = (REdge*)take(sizeof(REdge) * ne);
    w.offs = (int64_t*)take(sizeof(int64_t) * (n_graphs + 1));
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

extern "C" int octa_raster2d_batch_dev(const double* edges7_dev, const int64_t* edge_offsets_host, int n_graphs, int H, int W,
                                       int mip_axis, const OctaVoxOpts* opts, uint8_t* out_dev, void* workspace_dev,
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
    OCTA_CUDA_CHECK(cudaMemsetAsync(w.tile_count, 0, (char*)w.tile_start - (char*)w.tile_count, stream));      // counts, big counts, error flag
    const int threads = 128, blocks = (int)((n_edges + threads - 1) / threads);
    if (n_edges > 0) { r2d_prep_kernel<<<blocks, threads, 0, stream>>>(edges7_dev, w.offs, n_graphs, g, w.prep, w.tile_count, w.big_count, w.big_idx); octa::count_launch(); }
    r2d_scan_kernel<<<n_graphs, 1024, 0, stream>>>(w.tile_count, w.tile_start, w.cursor, g.ntiles);
    octa::count_launch();
    if (n_edges > 0) { r2d_fill_kernel<<<blocks, threads, 0, stream>>>(w.offs, n_graphs, g, w.prep, w.cursor, w.tile_edges); octa::count_launch(); }
    static bool attr_set = false;
    if (!attr_set) {
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(r2d_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSm)));
        attr_set = true;
    }
    r2d_tile_kernel<<<dim3((unsigned)g.ntiles, (unsigned)n_graphs), R2D_THREADS, sizeof(TileSm), stream>>>(
        w.prep, w.offs, g, w.tile_start, w.tile_edges, w.big_count, w.big_idx, out_dev, w.err);
    octa::count_launch();
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
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
