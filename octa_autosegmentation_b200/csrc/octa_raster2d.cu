// K7: 2-D anti-aliased label rasterizer for sm_100a.
//
// Replaces vessel_graph_generation/tree2img.py:12-114 (rasterize_forest), whose arithmetic lives in
// matplotlib's Agg backend: every kept edge is a round-capped stroke of width
//     1.3 * radius * max(W, H)  points  =  * dpi/72 = 100/72 pixels              (tree2img.py:82-86, dpi :51)
// drawn white on black into a W x H canvas with y inverted, pixel (row, col) <-> (pos[ax0]*H, pos[ax1]*W),
// ax = {0,1,2} \ {MIP_axis} (:46,:85), anti-aliased, composited "over" in list order, read back as 8-bit
// gray (:104-113).  matplotlib is not part of this image and the reference pins no version, so this
// path is "parity unpinned": the model below is validated statistically against the
// label PNGs the reference ships (tests/test_raster2d_gpu.py, IoU / vessel fraction).
//
// Coverage model: box-filtered capsule -- with d the distance of the pixel centre to the segment and
// r the half stroke width in pixels, coverage = clamp(min(d+1/2, r) - max(d-1/2, -r), 0, 1), which is the
// exact pixel/strip overlap for axis-aligned strokes and within a few percent otherwise.  Compositing:
// T = prod(1 - a_i) in list order, gray = round(255 (1 - T)).
//
// Design: same tile ownership as K6 -- one CTA owns a 32x32 pixel tile of one graph, one thread owns
// one pixel and walks the tile's edge list (binned by prep/scan/fill kernels, each list sorted back
// into edge order so the result is deterministic), then writes its byte once: HBM traffic is the
// algorithmic 1 byte/pixel + 56 bytes/edge.
#include "octa_common.h"
#include <math.h>

namespace {

constexpr int RT = 32;          // tile edge (pixels)
constexpr int RKBIG = 64;

struct REdge {                  // pixel-space capsule
    float x1, y1, x2, y2, r;
    int lo[2], hi[2];           // tile range, inclusive; hi < lo -> skipped
};

struct RGeom {
    int H, W, ntx, nty, ntiles;
    int ax0, ax1;
    double scale, min_radius, max_radius;
};

__global__ void r2d_prep_kernel(const double* __restrict__ edges7, const int64_t* __restrict__ offs, int n_graphs, RGeom g,
                                REdge* __restrict__ prep, int* __restrict__ tile_count, int* __restrict__ big_count,
                                int* __restrict__ big_idx) {
    const int64_t n_edges = offs[n_graphs];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    int lo = 0, hi = n_graphs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= i) lo = mid; else hi = mid; }
    const int gr = lo;
    const double* e = edges7 + 7 * i;
    REdge q;
    const double radius = e[6];
    const bool keep = !(radius < g.min_radius || radius > g.max_radius);            // tree2img.py:67
    // thickness = 1.3 * radius * scale_factor [points]; 1 pt = 100/72 px; half width in px
    const double half = 0.5 * ((radius * 1.3) * g.scale) * (100.0 / 72.0);
    const double x1 = e[g.ax1] * g.W, y1 = e[g.ax0] * g.H, x2 = e[3 + g.ax1] * g.W, y2 = e[3 + g.ax0] * g.H;
    q.x1 = (float)x1; q.y1 = (float)y1; q.x2 = (float)x2; q.y2 = (float)y2; q.r = (float)half;
    const double reach = half + 0.75;
    const double bx0 = fmin(x1, x2) - reach, bx1 = fmax(x1, x2) + reach, by0 = fmin(y1, y2) - reach, by1 = fmax(y1, y2) + reach;
    int px0 = (int)fmax(0.0, floor(bx0)), px1 = (int)fmin((double)g.W - 1, floor(bx1));
    int py0 = (int)fmax(0.0, floor(by0)), py1 = (int)fmin((double)g.H - 1, floor(by1));
    if (!keep || !(bx1 >= 0) || !(by1 >= 0) || px0 > px1 || py0 > py1 || !(half == half)) { q.lo[0] = q.lo[1] = 1; q.hi[0] = q.hi[1] = 0; prep[i] = q; return; }
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
    int lo = 0, hi = n_graphs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= i) lo = mid; else hi = mid; }
    const int gr = lo;
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

__device__ __forceinline__ float capsule_cover(const REdge& e, float px, float py) {
    const float sx = e.x2 - e.x1, sy = e.y2 - e.y1;
    const float ux = px - e.x1, uy = py - e.y1;
    const float ss = sx * sx + sy * sy;
    float t = ss > 0.f ? (ux * sx + uy * sy) / ss : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float dx = ux - t * sx, dy = uy - t * sy;
    const float d = sqrtf(dx * dx + dy * dy);
    const float c = fminf(d + 0.5f, e.r) - fmaxf(d - 0.5f, -e.r);
    return fminf(fmaxf(c, 0.f), 1.f);
}

__global__ void __launch_bounds__(RT * RT)
r2d_tile_kernel(const REdge* __restrict__ prep, const int64_t* __restrict__ offs, RGeom g, const int* __restrict__ tile_start,
                int* __restrict__ tile_edges, const int* __restrict__ big_count, const int* __restrict__ big_idx,
                uint8_t* __restrict__ out) {
    const int gr = blockIdx.y, tile = blockIdx.x;
    const int tx = tile % g.ntx, ty = tile / g.ntx;
    const int64_t eb = offs[gr];
    const REdge* ge = prep + eb;
    const int* st = tile_start + (size_t)gr * (g.ntiles + 1);
    const int beg = st[tile], end = st[tile + 1];
    int* lst = tile_edges + (size_t)RKBIG * eb;
    const int* bl = big_idx + eb;
    const int nbig = big_count[gr];
    // restore list order (atomics filled the lists in arbitrary order; the float product below is applied in edge order).
    // Lists of up to one entry per thread are ranked in shared memory (edge ids are distinct: rank = number of smaller ids,
    // two barriers); longer ones fall back to an odd-even transposition in place (one barrier and one global round trip per pass,
    // which dominated this kernel when it was the only path).
    __shared__ int s_in[RT * RT], s_sorted[RT * RT];
    const int n = end - beg;
    const bool in_smem = n <= RT * RT;
    if (in_smem) {
        int v = 0;
        if ((int)threadIdx.x < n) { v = lst[beg + threadIdx.x]; s_in[threadIdx.x] = v; }
        __syncthreads();
        if ((int)threadIdx.x < n) {
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += s_in[j] < v;
            s_sorted[rank] = v;
        }
        __syncthreads();
    } else {
        for (int pass = 0; pass < n; ++pass) {
            for (int i = (pass & 1) + 2 * threadIdx.x; i + 1 < n; i += 2 * blockDim.x) {
                const int a = lst[beg + i], b = lst[beg + i + 1];
                if (a > b) { lst[beg + i] = b; lst[beg + i + 1] = a; }
            }
            __syncthreads();
        }
    }
    const int px = tx * RT + (threadIdx.x % RT), py = ty * RT + (threadIdx.x / RT);
    if (px >= g.W || py >= g.H) return;
    const float cx = (float)px + 0.5f, cy = (float)py + 0.5f;
    float T = 1.f;
    if (in_smem) { for (int k = 0; k < n; ++k) T *= 1.f - capsule_cover(ge[s_sorted[k]], cx, cy); }
    else { for (int k = beg; k < end; ++k) T *= 1.f - capsule_cover(ge[lst[k]], cx, cy); }
    // edges spanning more than RKBIG tiles (rare): the product commutes, so they are applied after the tile
    // list, in ascending edge order (selection over the unsorted, tiny list keeps the result deterministic)
    int prev = -1;
    for (int k = 0; k < nbig; ++k) {
        int best = 0x7fffffff;
        for (int j = 0; j < nbig; ++j) { const int v = bl[j]; if (v > prev && v < best) best = v; }
        prev = best;
        const REdge e = ge[best];
        if (tx < e.lo[0] || tx > e.hi[0] || ty < e.lo[1] || ty > e.hi[1]) continue;
        T *= 1.f - capsule_cover(e, cx, cy);
    }
    out[((size_t)gr * g.H + py) * g.W + px] = (uint8_t)__float2int_rn(255.f * (1.f - T));
}

int make_rgeom(int H, int W, int mip_axis, const OctaVoxOpts* opts, RGeom* g) {
    OCTA_ARG_CHECK(H > 0 && W > 0 && H <= 32768 && W <= 32768, "bad image resolution");
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

struct RWork { REdge* prep; int64_t* offs; int *tile_count, *big_count, *tile_start, *cursor, *big_idx, *tile_edges; size_t bytes; };

RWork rcarve(void* base, int n_graphs, int64_t n_edges, int ntiles) {
    RWork w;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = octa::align_up(off + b, 256); return (char*)base + o; };
    const size_t ne = (size_t)(n_edges > 0 ? n_edges : 1);
    w.prep = (REdge*)take(sizeof(REdge) * ne);
    w.offs = (int64_t*)take(sizeof(int64_t) * (n_graphs + 1));
    w.tile_count = (int*)take(sizeof(int) * (size_t)n_graphs * ntiles);
    w.big_count = (int*)take(sizeof(int) * n_graphs);
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
    OCTA_CUDA_CHECK(cudaMemsetAsync(w.tile_count, 0, (char*)w.tile_start - (char*)w.tile_count, stream));
    const int threads = 128, blocks = (int)((n_edges + threads - 1) / threads);
    if (n_edges > 0) { r2d_prep_kernel<<<blocks, threads, 0, stream>>>(edges7_dev, w.offs, n_graphs, g, w.prep, w.tile_count, w.big_count, w.big_idx); octa::count_launch(); }
    r2d_scan_kernel<<<n_graphs, 1024, 0, stream>>>(w.tile_count, w.tile_start, w.cursor, g.ntiles);
    octa::count_launch();
    if (n_edges > 0) { r2d_fill_kernel<<<blocks, threads, 0, stream>>>(w.offs, n_graphs, g, w.prep, w.cursor, w.tile_edges); octa::count_launch(); }
    r2d_tile_kernel<<<dim3((unsigned)g.ntiles, (unsigned)n_graphs), RT * RT, 0, stream>>>(w.prep, w.offs, g, w.tile_start, w.tile_edges,
                                                                                      w.big_count, w.big_idx, out_dev);
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
    cleanup();
    return rc;
}
