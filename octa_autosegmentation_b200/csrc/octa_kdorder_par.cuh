// Block-parallel construction of cKDTree's `indices` permutation (see octa_kdorder.h for the sequential
// restatement and what it mirrors).  One CTA (1024 threads) builds the permutation of one graph's sink list.
//
// Parallelisation.  The tree is processed level by level.  A node is handled by a GROUP of warps:
// 32/16/8/4 warps on levels 0..3 (synchronised with named barriers), one warp per node below.  Inside a
// node, std::nth_element's introselect loop is replayed round by round; the Hoare partition of a round is
// a deterministic function of the input order and is computed with prefix sums:
//   cut   = first + 1 + #{x < pivot}
//   l_k   = k-th position (ascending) left of cut holding an element > pivot
//   r_k   = k-th position (descending) right of cut holding an element < pivot
//   swap idx[l_k] <-> idx[r_k] for all k           (exactly the swaps __unguarded_partition performs)
// so the resulting permutation is bit-identical to the sequential library routine.  Pivot selection
// (median of three), the <= 3 element insertion sort and the depth-limit fallback (heap select, never
// reached for random inputs) run on the group's first thread through the sequential code.
#pragma once
#include "octa_kdorder.h"

namespace octa {
namespace kdpar {

constexpr int KD_THREADS = 1024;

__device__ __forceinline__ void grp_sync(int nwarps, int bar_id) {
    if (nwarps == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nwarps * 32) : "memory");
}

// inclusive scan of v over the threads of a group (thread order); `ws` = 33 ints of shared memory owned by the group
__device__ __forceinline__ int grp_scan(int v, int gt, int nwarps, int bar_id, int* ws, int* total) {
    const int lane = gt & 31, warp = gt >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (nwarps == 1) {
        *total = __shfl_sync(0xffffffffu, x, 31);
        return x;
    }
    if (lane == 31) ws[warp] = x;
    grp_sync(nwarps, bar_id);
    if (warp == 0) {
        int w = lane < nwarps ? ws[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        ws[lane] = w;
        if (lane == 31) ws[32] = w;
    }
    grp_sync(nwarps, bar_id);
    if (warp > 0) x += ws[warp - 1];
    *total = ws[32];
    grp_sync(nwarps, bar_id);
    return x;
}

// std::nth_element on idx[first, last) around nth, by a group of `nwarps` warps.  posL/posR: scratch of (last-first) ints.
static __device__ void nth_element_group(const double* __restrict__ key, int* idx, int first, int nth, int last, int gt, int nwarps,
                                  int bar_id, int* ws, int* bc /*4 ints of shared memory*/, int* posL, int* posR) {
    const int G = nwarps * 32;
    const kd::Less less{key};
    int depth_limit = kd::floor_log2(last - first) * 2;
    while (last - first > 3) {
        if (depth_limit == 0) {
            if (gt == 0) { kd::heap_select(idx + first, idx + nth + 1, idx + last, less); kd::swap_i(idx + first, idx + nth); }
            grp_sync(nwarps, bar_id);
            return;
        }
        --depth_limit;
        if (gt == 0) {
            int* mid = idx + first + (last - first) / 2;
            kd::move_median_to_first(idx + first, idx + first + 1, mid, idx + last - 1, less);
        }
        grp_sync(nwarps, bar_id);
        const int piv = idx[first];
        const double pk = key[piv];
        const int lo = first + 1, len = last - lo;
        // every thread owns a contiguous slice; slices of more than 32 elements are handled in super-chunks
        int cut;
        {
            int c = 0;
            for (int i = lo + gt; i < last; i += G) { const int a = idx[i]; const double ka = key[a]; c += (ka == pk ? a < piv : ka < pk); }
            int tot;
            grp_scan(c, gt, nwarps, bar_id, ws, &tot);
            cut = lo + tot;
        }
        int carryL = 0, carryR = 0;
        for (int sc = 0; sc < len; sc += G * 32) {
            const int slen = (len - sc) < G * 32 ? (len - sc) : G * 32;
            const int S = (slen + G - 1) / G;                       // slice length (<= 32)
            const int b = lo + sc + gt * S;
            const int e = (b + S < lo + sc + slen) ? b + S : lo + sc + slen;
            unsigned int mask = 0;                                   // bit j: element b+j is misplaced
            int nL = 0, nR = 0;
            for (int i = b; i < e; ++i) {
                const int a = idx[i];
                const double ka = key[a];
                const bool lt = (ka == pk ? a < piv : ka < pk);
                if (i < cut) { if (!lt) { mask |= 1u << (i - b); ++nL; } }
                else if (lt) { mask |= 1u << (i - b); ++nR; }
            }
            int tot;
            const int incl = grp_scan(nL | (nR << 16), gt, nwarps, bar_id, ws, &tot);
            int rL = carryL + (int)((unsigned)incl & 0xffffu) - nL, rR = carryR + (int)((unsigned)incl >> 16) - nR;
            for (int i = b; i < e; ++i)
                if ((mask >> (i - b)) & 1u) { if (i < cut) posL[rL++] = i; else posR[rR++] = i; }
            carryL += (int)((unsigned)tot & 0xffffu);
            carryR += (int)((unsigned)tot >> 16);
        }
        grp_sync(nwarps, bar_id);
        const int K = carryL;                                        // == carryR
        for (int k = gt; k < K; k += G) {
            const int i = posL[k], j = posR[K - 1 - k];
            const int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
        }
        grp_sync(nwarps, bar_id);
        if (cut <= nth) first = cut; else last = cut;
    }
    if (gt == 0) kd::insertion_sort(idx + first, idx + last, less);
    grp_sync(nwarps, bar_id);
    (void)bc;
}

// One tree node on idx[start,end) by a group: bounding box, split dimension, nth_element.  Returns the split
// position (same value on every thread of the group) or -1 for a leaf.
static __device__ int build_node_group(const double* const xyz[3], int* idx, int start, int end, int gt, int nwarps, int bar_id,
                                int* ws, double* wd /*6*32 doubles of shared memory*/, int* posL, int* posR) {
    if (end - start <= kd::LEAFSIZE) return -1;
    const int G = nwarps * 32;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = start + gt; i < end; i += G) {
        const int a = idx[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double t = xyz[k][a]; mn[k] = fmin(mn[k], t); mx[k] = fmax(mx[k], t); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if (nwarps > 1) {
        const int lane = gt & 31, warp = gt >> 5;
        if (lane == 0) for (int k = 0; k < 3; ++k) { wd[k * 32 + warp] = mn[k]; wd[(3 + k) * 32 + warp] = mx[k]; }
        grp_sync(nwarps, bar_id);
        for (int k = 0; k < 3; ++k) {
            double a = INFINITY, b = -INFINITY;
            for (int w = 0; w < nwarps; ++w) { a = fmin(a, wd[k * 32 + w]); b = fmax(b, wd[(3 + k) * 32 + w]); }
            mn[k] = a; mx[k] = b;
        }
        grp_sync(nwarps, bar_id);
    }
    int d = 0;
    double size = 0;
    for (int k = 0; k < 3; ++k)
        if (mx[k] - mn[k] > size) { d = k; size = mx[k] - mn[k]; }
    if (mx[d] == mn[d]) return -1;
    const int n = end - start;
    nth_element_group(xyz[d], idx, start, start + n / 2, end, gt, nwarps, bar_id, ws, nullptr, posL + start, posR + start);
    // (cKDTree's own partition loop and slide fix-ups are no-ops for distinct coordinates)
    return start + n / 2;
}

// Whole build by one CTA of KD_THREADS threads.  node_s/node_e: two ping-pong lists of at most n/8+2 ranges each.
static __device__ void build_indices_block(const double* x, const double* y, const double* z, int n, int* idx, int* posL, int* posR,
                                    int* node_a, int* node_b) {
    __shared__ int s_ws[32 * 33];
    __shared__ double s_wd[8 * 6 * 32];
    __shared__ int s_cnt;
    const double* const xyz[3] = {x, y, z};
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) idx[i] = i;
    int* cur = node_a;
    int* nxt = node_b;
    if (tid == 0) { cur[0] = 0; cur[1] = n; s_cnt = 0; }
    __syncthreads();
    int ncur = 1;
    for (int level = 0; ncur > 0; ++level) {
        // group size for this level: 32 warps >> level, at least one warp
        int nwarps = 32 >> level;
        if (nwarps < 4) nwarps = 1;                  // below 4 warps per node: one warp per node
        const int ngroups = 32 / nwarps;
        const int grp = (tid >> 5) / nwarps;
        const int gt = tid - grp * nwarps * 32;
        const int bar_id = 1 + grp;                  // <= 8 multi-warp groups -> ids 1..8
        for (int base = 0; base < ncur; base += ngroups) {
            const int k = base + grp;
            if (k < ncur) {
                const int s = cur[2 * k], e = cur[2 * k + 1];
                const int p = build_node_group(xyz, idx, s, e, gt, nwarps, bar_id, s_ws + grp * 33, s_wd + (grp & 7) * 192,
                                               posL, posR);
                if (p >= 0 && gt == 0) {
                    const int o = atomicAdd(&s_cnt, 2);
                    nxt[2 * o] = s; nxt[2 * o + 1] = p; nxt[2 * o + 2] = p; nxt[2 * o + 3] = e;
                }
            }
        }
        __syncthreads();
        ncur = s_cnt;
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        int* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
}

}  // namespace kdpar
}  // namespace octa
