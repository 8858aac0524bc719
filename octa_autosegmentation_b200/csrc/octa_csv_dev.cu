// Device writer of the graph CSV (generate_vessel_graph.py:59-66) for a whole batch: the same bytes octa_format_csv writes on
// the host, formatted by the cell formatters of octa_csvfmt.cuh (integer arithmetic only, shared with the host build and tested
// there against numpy / CPython).  The host writer costs 3.7 ms of a core per docker-config graph -- at 650 graphs/s that is
// 2.4 cores per GPU, the largest host cost of the end-to-end path; on the device a batch of 64 files (80 MB of text) takes
// well under a millisecond.
//   csv_rows_kernel   thread per row: the three cells into a 144-byte slot + the row length; a cell the formatters decline
//                     marks the graph for the host writer
//   csv_scan_kernel   CTA per graph: exclusive scan of its row lengths (+ the 20-byte header); then the scan over the graphs
//   csv_pack_kernel   warp per row: slot -> final position (coalesced within the row), headers
#include "octa_common.h"
#include "octa_csvfmt.cuh"

namespace {

constexpr int SLOT = 144;          // >= longest row: 2 x 56 (cells) + 24 (radius) + 4 (commas, CRLF)
constexpr int HEADER_LEN = 20;     // "node1,node2,radius\r\n"

struct CsvWs {
    char* slots;        // [E][SLOT]
    int* lens;          // [E]
    int* row_off;       // [E]   offset of the row inside its file
    int64_t* edge_off;  // [G+1] device copy of the edge offsets
    long long* glen;    // [G]
    size_t bytes;
};

CsvWs carve(void* base, int n_graphs, int64_t n_edges) {
    CsvWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = octa::align_up(off + bytes, 256); return (char*)base + o; };
    const size_t E = (size_t)(n_edges > 0 ? n_edges : 1);
    w.slots = take(E * SLOT);
    w.lens = (int*)take(E * sizeof(int));
    w.row_off = (int*)take(E * sizeof(int));
    w.edge_off = (int64_t*)take(sizeof(int64_t) * (size_t)(n_graphs + 1));
    w.glen = (long long*)take(sizeof(long long) * (size_t)(n_graphs + 1));
    w.bytes = off;
    return w;
}

__global__ void __launch_bounds__(128) csv_rows_kernel(const double* __restrict__ edges7, const int64_t* __restrict__ edge_off,
                                                       int n_graphs, char* __restrict__ slots, int* __restrict__ lens,
                                                       int32_t* __restrict__ fallback) {
    const int64_t n_edges = edge_off[n_graphs];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    double e[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) e[k] = edges7[i * 7 + k];
    __align__(16) char row[SLOT];
    int k = 0;
    bool ok = true;
    int c = octa::csvfmt::array3(row, e);
    if (c < 0) ok = false; else { k = c; row[k++] = ','; }
    if (ok) { c = octa::csvfmt::array3(row + k, e + 3); if (c < 0) ok = false; else { k += c; row[k++] = ','; } }
    if (ok) { c = octa::csvfmt::repr_unit(row + k, e[6]); if (c < 0) ok = false; else { k += c; row[k++] = '\r'; row[k++] = '\n'; } }
    if (!ok) {
        int lo = 0, hi = n_graphs;               // graph of this row
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (edge_off[mid] <= i) lo = mid; else hi = mid; }
        fallback[lo] = 1;
        k = 0;
    }
    lens[i] = k;
    uint4* dst = reinterpret_cast<uint4*>(slots + (size_t)i * SLOT);
    const uint4* src = reinterpret_cast<const uint4*>(row);
    for (int q = 0; q < (k + 15) / 16; ++q) dst[q] = src[q];
}

// one CTA per graph: row offsets inside the file (header first) and the file length
__global__ void __launch_bounds__(1024) csv_scan_kernel(const int* __restrict__ lens, const int64_t* __restrict__ edge_off,
                                                        int* __restrict__ row_off, long long* __restrict__ glen) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int g = blockIdx.x;
    const int64_t r0 = edge_off[g], r1 = edge_off[g + 1];
    if (threadIdx.x == 0) carry = HEADER_LEN;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t base = r0; base < r1; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int v = i < r1 ? lens[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nw ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int prefix = carry + (warp ? warp_sums[warp - 1] : 0) + (x - v);
        if (i < r1) row_off[i] = prefix;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) glen[g] = carry;
}

// file offsets: exclusive scan of the file lengths (n_graphs is small: one thread)
__global__ void csv_offsets_kernel(const long long* __restrict__ glen, int n_graphs, int64_t* __restrict__ text_off) {
    if (threadIdx.x || blockIdx.x) return;
    long long run = 0;
    for (int g = 0; g < n_graphs; ++g) { text_off[g] = run; run += glen[g]; }
    text_off[n_graphs] = run;
}

__global__ void __launch_bounds__(256) csv_pack_kernel(const char* __restrict__ slots, const int* __restrict__ lens,
                                                       const int* __restrict__ row_off, const int64_t* __restrict__ edge_off,
                                                       const int64_t* __restrict__ text_off, const int32_t* __restrict__ fallback,
                                                       int n_graphs, char* __restrict__ text, size_t text_cap) {
    const int64_t n_edges = edge_off[n_graphs];
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // warp = row; the first n_graphs warps also write headers
    if (w < n_graphs && !fallback[w]) {
        const char hdr[HEADER_LEN + 1] = "node1,node2,radius\r\n";
        const int64_t o = text_off[w];
        if (lane < HEADER_LEN && (size_t)(o + lane) < text_cap) text[o + lane] = hdr[lane];
    }
    if (w >= n_edges) return;
    int lo = 0, hi = n_graphs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (edge_off[mid] <= w) lo = mid; else hi = mid; }
    if (fallback[lo]) return;
    const int n = lens[w];
    const int64_t o = text_off[lo] + row_off[w];
    if ((size_t)(o + n) > text_cap) return;
    const char* src = slots + (size_t)w * SLOT;
    for (int q = lane; q < n; q += 32) text[o + q] = src[q];
}

}  // namespace

extern "C" size_t octa_format_csv_workspace_bytes(int n_graphs, int64_t n_edges) {
    if (n_graphs <= 0 || n_edges < 0) return 0;
    return carve(nullptr, n_graphs, n_edges).bytes;
}

extern "C" size_t octa_format_csv_text_cap(int n_graphs, int64_t n_edges) {
    if (n_graphs <= 0 || n_edges < 0) return 0;
    return (size_t)n_graphs * HEADER_LEN + (size_t)n_edges * SLOT;
}

extern "C" int octa_format_csv_batch_dev(const double* edges7_dev, const int64_t* edge_offsets_host, int n_graphs, char* text_dev,
                                         size_t text_cap, int64_t* text_offsets_dev, int32_t* fallback_dev, void* workspace_dev,
                                         size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    OCTA_ARG_CHECK(n_graphs > 0 && n_graphs <= 65535, "n_graphs must be in [1, 65535]");
    OCTA_ARG_CHECK(edge_offsets_host && text_dev && text_offsets_dev && fallback_dev && workspace_dev, "null pointer");
    OCTA_ARG_CHECK(edge_offsets_host[0] == 0, "edge_offsets[0] must be 0");
    for (int i = 0; i < n_graphs; ++i)
        OCTA_ARG_CHECK(edge_offsets_host[i + 1] >= edge_offsets_host[i], "edge_offsets must be non-decreasing");
    const int64_t n_edges = edge_offsets_host[n_graphs];
    OCTA_ARG_CHECK(n_edges == 0 || edges7_dev, "edges pointer is null");
    OCTA_ARG_CHECK(n_edges < ((int64_t)1 << 31) / SLOT * 8, "too many edges");
    if (octa_device_count() <= 0) { octa::set_error("octa_format_csv_batch_dev: no CUDA device (there is no CPU fallback)"); return OCTA_E_CUDA; }
    const CsvWs w = carve(workspace_dev, n_graphs, n_edges);
    if (w.bytes > workspace_bytes) { octa::set_error("octa_format_csv_batch_dev: workspace too small (%zu < %zu)", workspace_bytes, w.bytes); return OCTA_E_NOMEM; }
    // (a text buffer below the bound is allowed: rows that would not fit are dropped and the caller sees offsets beyond text_cap)
    OCTA_CUDA_CHECK(cudaMemcpyAsync(w.edge_off, edge_offsets_host, sizeof(int64_t) * (n_graphs + 1), cudaMemcpyHostToDevice, stream));
    OCTA_CUDA_CHECK(cudaMemsetAsync(fallback_dev, 0, sizeof(int32_t) * n_graphs, stream));
    if (n_edges > 0) {
        csv_rows_kernel<<<(unsigned)((n_edges + 127) / 128), 128, 0, stream>>>(edges7_dev, w.edge_off, n_graphs, w.slots, w.lens, fallback_dev);
        octa::count_launch();
    }
    csv_scan_kernel<<<n_graphs, 1024, 0, stream>>>(w.lens, w.edge_off, w.row_off, w.glen);
    csv_offsets_kernel<<<1, 32, 0, stream>>>(w.glen, n_graphs, text_offsets_dev);
    const int64_t warps = n_edges > n_graphs ? n_edges : n_graphs;
    csv_pack_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(w.slots, w.lens, w.row_off, w.edge_off, text_offsets_dev, fallback_dev,
                                                                               n_graphs, text_dev, text_cap);
    octa::count_launch(3);
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}
