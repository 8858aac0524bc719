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
constexpr int WS_INTS = 32 * 33;       // scan scratch (ints) and reduction scratch (doubles) a build needs in shared memory
constexpr int WD_DOUBLES = 8 * 6 * 32;

static __device__ void build_indices_block(const double* x, const double* y, const double* z, int n, int* idx, int* posL, int* posR,
                                    int* node_a, int* node_b, int* s_ws /*WS_INTS*/, double* s_wd /*WD_DOUBLES*/) {
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

// ------------------------------------------------------------------------------------------------------------
// Shared-memory resident build (the one k_kdbuild uses when the sink list fits: 12 bytes per point).
//
// Same tree walk and the same replay of std::nth_element as above, but the permutation (16-bit ids) and the key of
// the node's split dimension travel TOGETHER through the partition swaps, so a round never leaves shared memory,
// and a round needs three group barriers instead of nine:
//   * the median-of-three is evaluated redundantly by every thread and applied logically (position m holds the
//     old first element) -- the physical writes happen in the swap phase;
//   * one pass builds each thread's "< pivot" bit mask (contiguous slice, odd length: conflict-free 8-byte loads);
//     ONE prefix sum of the per-thread counts gives the cut and every rank: the k-th misplaced element on the
//     left (ascending) has k = (i - lo) - #{< pivot before i}, the k-th on the right (descending) has
//     k = total - #{< pivot up to and including i};
//   * right-hand owners publish their positions by rank, left-hand owners fetch their partner and swap.
// Returns false when the introselect depth limit is hit or a slice would exceed the 64-bit mask (never for random
// inputs); the caller then runs the global-memory version above.
// ------------------------------------------------------------------------------------------------------------
namespace kdsm {

constexpr int THREADS = 1024;

struct View {
    double* key;              // [n] key of the owning node's split dimension, permuted together with idx
    unsigned short* idx;      // [n] the permutation
    unsigned short* posR;     // [n] scratch: positions of the right-hand misplaced elements by descending rank
};

__device__ __forceinline__ bool pair_less(double ka, int ia, double kb, int ib) { return ka == kb ? ia < ib : ka < kb; }

// inclusive scan over a group; one barrier (ws: 2 x 32 ints owned by the group, alternating)
__device__ __forceinline__ int scan1(int v, int gt, int nwarps, int bar_id, int* ws, int& phase, int* total) {
    const int lane = gt & 31, warp = gt >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (nwarps == 1) { *total = __shfl_sync(0xffffffffu, x, 31); return x; }
    int* w = ws + (phase & 1) * 32;
    ++phase;
    if (lane == 31) w[warp] = x;
    kdpar::grp_sync(nwarps, bar_id);
    int s = lane < nwarps ? w[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
    }
    *total = __shfl_sync(0xffffffffu, s, nwarps - 1);
    const int before = __shfl_sync(0xffffffffu, s, warp > 0 ? warp - 1 : 0);
    return x + (warp > 0 ? before : 0);
}

// std::__heap_select / __adjust_heap / __push_heap (bits/stl_heap.h) on (key, id) pairs -- the introselect fallback,
// run by one thread (it is reached on a handful of elements, after 2*log2(size) unlucky partitions)
struct Pair { double k; int i; };
__device__ __forceinline__ Pair pget(const View& M, int p) { Pair r; r.k = M.key[p]; r.i = M.idx[p]; return r; }
__device__ __forceinline__ void pset(const View& M, int p, const Pair& v) { M.key[p] = v.k; M.idx[p] = (unsigned short)v.i; }
__device__ __forceinline__ bool pless(const Pair& a, const Pair& b) { return pair_less(a.k, a.i, b.k, b.i); }

static __device__ void adjust_heap_sm(const View& M, int base, long hole, long len, Pair value) {
    const long top = hole;
    long second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (pless(pget(M, base + second), pget(M, base + second - 1))) --second;
        pset(M, base + hole, pget(M, base + second));
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        pset(M, base + hole, pget(M, base + second - 1));
        hole = second - 1;
    }
    long parent = (hole - 1) / 2;                            // __push_heap
    while (hole > top && pless(pget(M, base + parent), value)) {
        pset(M, base + hole, pget(M, base + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    pset(M, base + hole, value);
}

static __device__ void heap_select_sm(const View& M, int first, int middle, int last) {
    const long len = middle - first;
    if (len >= 2) {                                          // __make_heap
        long parent = (len - 2) / 2;
        while (true) {
            adjust_heap_sm(M, first, parent, len, pget(M, first + parent));
            if (parent == 0) break;
            --parent;
        }
    }
    for (int i = middle; i < last; ++i)
        if (pless(pget(M, i), pget(M, first))) {             // __pop_heap(first, middle, i)
            const Pair value = pget(M, i);
            pset(M, i, pget(M, first));
            adjust_heap_sm(M, first, 0, len, value);
        }
}

// n-th (0-based) set bit of a 32-bit mask counted from the bottom / from the top
__device__ __forceinline__ int nth_bit_up(unsigned int mask, int n) {
    for (int r = 0; r < n; ++r) mask &= mask - 1;
    return __ffs(mask) - 1;
}
__device__ __forceinline__ int nth_bit_down(unsigned int mask, int n) { return 31 - nth_bit_up(__brev(mask), n); }

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// The tail of std::nth_element once the range fits one element per lane (<= 32): ONE warp keeps the elements in
// registers (lane l = position base + l) and replays the remaining introselect rounds with ballots and shuffles --
// no shared-memory traffic, no barriers, ~100 instructions per round.  Same operations as the rounds above:
// median-of-three to the front, Hoare partition (k-th misplaced element from the left <-> k-th from the right),
// insertion sort of the last <= 3 elements, heap select when the depth limit runs out.
static __device__ void nth_element_warp(const View& M, int first, int nth, int last, int lane, int depth_limit) {
    const int base = first, cnt = last - first;
    double k = 0; int ix = 0;
    if (lane < cnt) { k = M.key[base + lane]; ix = M.idx[base + lane]; }
    int f = 0, l = cnt;                                       // current range, relative to base
    const int nrel = nth - base;
    while (l - f > 3) {
        if (depth_limit == 0) {
            if (lane < cnt) { M.key[base + lane] = k; M.idx[base + lane] = (unsigned short)ix; }
            __syncwarp();
            if (lane == 0) {
                heap_select_sm(M, base + f, nth + 1, base + l);
                const Pair a = pget(M, base + f), b = pget(M, nth);
                pset(M, base + f, b); pset(M, nth, a);
            }
            __syncwarp();
            return;
        }
        --depth_limit;
        const int pa = f + 1, pb = f + (l - f) / 2, pc = l - 1;
        const double kf = shfl_d(k, f), ka = shfl_d(k, pa), kb = shfl_d(k, pb), kc = shfl_d(k, pc);
        const int if_ = __shfl_sync(0xffffffffu, ix, f), ia = __shfl_sync(0xffffffffu, ix, pa), ib = __shfl_sync(0xffffffffu, ix, pb),
                  ic = __shfl_sync(0xffffffffu, ix, pc);
        int m;
        if (pair_less(ka, ia, kb, ib)) {
            if (pair_less(kb, ib, kc, ic)) m = pb;
            else if (pair_less(ka, ia, kc, ic)) m = pc;
            else m = pa;
        } else if (pair_less(ka, ia, kc, ic)) m = pa;
        else if (pair_less(kb, ib, kc, ic)) m = pc;
        else m = pb;
        const double pk = (m == pa) ? ka : (m == pb ? kb : kc);
        const int pi = (m == pa) ? ia : (m == pb ? ib : ic);
        if (lane == f) { k = pk; ix = pi; } else if (lane == m) { k = kf; ix = if_; }
        const bool in = lane > f && lane < l;
        const unsigned int range = __ballot_sync(0xffffffffu, in);
        const unsigned int lt = __ballot_sync(0xffffffffu, in && pair_less(k, ix, pk, pi));
        const int cut = f + 1 + __popc(lt);
        const unsigned int leftm = cut >= 32 ? 0xffffffffu : ((1u << cut) - 1u);
        const unsigned int LM = ~lt & range & leftm;          // > pivot, left of the cut
        const unsigned int RM = lt & ~leftm;                  // < pivot, at or right of the cut
        const unsigned int me = 1u << lane;
        int src = lane;
        if (LM & me) src = nth_bit_down(RM, __popc(LM & (me - 1u)));
        else if (RM & me) src = nth_bit_up(LM, __popc(RM & ~(me | (me - 1u))));
        k = shfl_d(k, src);
        ix = __shfl_sync(0xffffffffu, ix, src);
        if (cut <= nrel) f = cut; else l = cut;
    }
    // std::__insertion_sort on <= 3 elements == ascending order of the (distinct) pairs
    {
        const int n3 = l - f;
        const double k0 = shfl_d(k, f), k1 = shfl_d(k, f + 1 < 32 ? f + 1 : 31), k2 = shfl_d(k, f + 2 < 32 ? f + 2 : 31);
        const int i0 = __shfl_sync(0xffffffffu, ix, f), i1 = __shfl_sync(0xffffffffu, ix, f + 1 < 32 ? f + 1 : 31),
                  i2 = __shfl_sync(0xffffffffu, ix, f + 2 < 32 ? f + 2 : 31);
        if (lane >= f && lane < l && n3 >= 2) {
            // rank of each of the n3 values
            const double kk[3] = {k0, k1, k2};
            const int ii[3] = {i0, i1, i2};
            const int want = lane - f;                        // this lane receives the value of rank `want`
            for (int q = 0; q < n3; ++q) {
                int r = 0;
                for (int t = 0; t < n3; ++t) if (t != q && pair_less(kk[t], ii[t], kk[q], ii[q])) ++r;
                if (r == want) { k = kk[q]; ix = ii[q]; }
            }
        }
    }
    if (lane < cnt) { M.key[base + lane] = k; M.idx[base + lane] = (unsigned short)ix; }
    __syncwarp();
}

static __device__ bool nth_element_sm(const View& M, int first, int nth, int last, int gt, int nwarps, int bar_id, int* ws, int& phase) {
    const int G = nwarps * 32;
    int depth_limit = kd::floor_log2(last - first) * 2;
    while (last - first > 3) {
        if (last - first <= 32) {                            // the tail: one warp, registers only
            if (gt < 32) nth_element_warp(M, first, nth, last, gt, depth_limit);
            kdpar::grp_sync(nwarps, bar_id);
            return true;
        }
        if (depth_limit == 0) {
            if (gt == 0) {
                heap_select_sm(M, first, nth + 1, last);
                const Pair a = pget(M, first), b = pget(M, nth);
                pset(M, first, b); pset(M, nth, a);
            }
            kdpar::grp_sync(nwarps, bar_id);
            return true;
        }
        --depth_limit;
        // std::__move_median_to_first(first, first+1, mid, last-1), evaluated by every thread
        const int pa = first + 1, pb = first + (last - first) / 2, pc = last - 1;
        const double kf = M.key[first]; const int if_ = M.idx[first];
        const double ka = M.key[pa], kb = M.key[pb], kc = M.key[pc];
        const int ia = M.idx[pa], ib = M.idx[pb], ic = M.idx[pc];
        int m;
        if (pair_less(ka, ia, kb, ib)) {
            if (pair_less(kb, ib, kc, ic)) m = pb;
            else if (pair_less(ka, ia, kc, ic)) m = pc;
            else m = pa;
        } else if (pair_less(ka, ia, kc, ic)) m = pa;
        else if (pair_less(kb, ib, kc, ic)) m = pc;
        else m = pb;
        const double pk = (m == pa) ? ka : (m == pb ? kb : kc);
        const int pi = (m == pa) ? ia : (m == pb ? ib : ic);
        const int lo = first + 1, len = last - lo;
        int S = (len + G - 1) / G;
        S |= 1;                                              // odd slice length: the strided 8-byte loads do not collide
        if (S > 64) return false;
        const int b = lo + gt * S < last ? lo + gt * S : last;
        const int e = b + S < last ? b + S : last;
        unsigned long long mask = 0;                         // bit j: element b+j is < pivot
        for (int i = b; i < e; ++i) mask |= (unsigned long long)pair_less(M.key[i], M.idx[i], pk, pi) << (i - b);
        if (m >= b && m < e) {                               // position m logically holds the old first element
            const unsigned long long bit = 1ull << (m - b);
            mask = pair_less(kf, if_, pk, pi) ? (mask | bit) : (mask & ~bit);
        }
        const int c = __popcll(mask);
        int total;
        const int incl = scan1(c, gt, nwarps, bar_id, ws, phase, &total);
        const int cut = lo + total;
        const int before = incl - c;                         // elements < pivot before this slice
        const int ne = e - b;
        const unsigned long long valid = ne >= 64 ? ~0ull : ((1ull << ne) - 1ull);
        const int cb = cut - b;                              // slice-relative cut
        const unsigned long long left = cb <= 0 ? 0ull : (cb >= 64 ? ~0ull : ((1ull << cb) - 1ull));
        // right-hand misplaced elements (< pivot, at or after the cut) publish their position by descending rank
        for (unsigned long long rm = mask & ~left & valid; rm; rm &= rm - 1) {
            const int j = __ffsll((long long)rm) - 1;
            const int upto = before + __popcll(mask & ((2ull << j) - 1ull));      // elements < pivot up to and including it
            M.posR[lo + (total - upto)] = (unsigned short)(b + j);
        }
        kdpar::grp_sync(nwarps, bar_id);
        // left-hand misplaced elements (> pivot, before the cut) fetch their partner and swap (key and id together)
        for (unsigned long long lm = ~mask & left & valid; lm; lm &= lm - 1) {
            const int jb = __ffsll((long long)lm) - 1;
            const int i = b + jb;
            const int lt_before = before + __popcll(mask & ((1ull << jb) - 1ull));
            const int j = M.posR[lo + (i - lo) - lt_before];
            double ki = M.key[i], kj = M.key[j];
            int ii = M.idx[i], ij = M.idx[j];
            if (i == m) { ki = kf; ii = if_; }
            if (j == m) { kj = kf; ij = if_; }
            M.key[i] = kj; M.idx[i] = (unsigned short)ij;
            M.key[j] = ki; M.idx[j] = (unsigned short)ii;
        }
        if (gt == 0) {
            const bool ltm = pair_less(kf, if_, pk, pi);
            const bool swapped = (m < cut) ? !ltm : ltm;
            if (!swapped) { M.key[m] = kf; M.idx[m] = (unsigned short)if_; }
            M.key[first] = pk; M.idx[first] = (unsigned short)pi;
        }
        kdpar::grp_sync(nwarps, bar_id);
        if (cut <= nth) first = cut; else last = cut;
    }
    if (gt == 0) {                                           // std::__insertion_sort on <= 3 elements
        for (int i = first + 1; i < last; ++i) {
            const double k = M.key[i]; const int ix = M.idx[i];
            int p = i;
            while (p > first && pair_less(k, ix, M.key[p - 1], M.idx[p - 1])) { M.key[p] = M.key[p - 1]; M.idx[p] = M.idx[p - 1]; --p; }
            M.key[p] = k; M.idx[p] = (unsigned short)ix;
        }
    }
    kdpar::grp_sync(nwarps, bar_id);
    return true;
}

// One tree node on [start,end): bounding box, split dimension, key refill, nth_element.
// Returns the split position, -1 for a leaf, -2 for "fall back to the global-memory build".  *dim: in = dimension the
// key array currently holds for this range (-1: none), out = split dimension.
static __device__ int build_node_sm(const View& M, const double* const xyz[3], int start, int end, int* dim, int gt, int nwarps,
                                    int bar_id, int* ws, int& phase, double* wd /*6*32 doubles owned by the group*/) {
    if (end - start <= kd::LEAFSIZE) return -1;
    const int G = nwarps * 32;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = start + gt; i < end; i += G) {
        const int a = M.idx[i];
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
        kdpar::grp_sync(nwarps, bar_id);
        // every warp folds the per-warp partials itself (one value per lane, shuffle reduction)
        for (int k = 0; k < 3; ++k) {
            double a = lane < nwarps ? wd[k * 32 + lane] : INFINITY, b = lane < nwarps ? wd[(3 + k) * 32 + lane] : -INFINITY;
            for (int o = 16; o > 0; o >>= 1) {
                a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
                b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
            }
            mn[k] = a; mx[k] = b;
        }
        kdpar::grp_sync(nwarps, bar_id);
    }
    int d = 0;
    double size = 0;
    for (int k = 0; k < 3; ++k)
        if (mx[k] - mn[k] > size) { d = k; size = mx[k] - mn[k]; }
    if (mx[d] == mn[d]) return -1;
    if (*dim != d) {
        const double* kd_ = xyz[d];
        for (int i = start + gt; i < end; i += G) M.key[i] = kd_[M.idx[i]];
        kdpar::grp_sync(nwarps, bar_id);
    }
    *dim = d;
    const int n = end - start;
    if (!nth_element_sm(M, start, start + n / 2, end, gt, nwarps, bar_id, ws, phase)) return -2;
    return start + n / 2;
}

// Whole build by one CTA of THREADS threads; writes rank[point] = position in cKDTree's `indices`.
// smem: 12*n bytes (+ padding), nodes_a/nodes_b: global scratch of 3*(n/8+4) ints each.  Returns false (on every
// thread) if the caller has to fall back.
// Levels of the build over the node list `cur` (ncur nodes of 3 ints: start, end, dimension the key array holds), at most
// max_levels of them: a node is handled by 32 / 16 / 8 / 4 warps on the first four levels of the call, by one warp below.
// Returns the number of nodes left for the next level (their list is `cur` again), or -1 when the caller has to fall back.
static __device__ int run_levels(const View& M, const double* const xyz[3], int*& cur, int*& nxt, int ncur, int max_levels,
                                 int* s_ws, double* s_wd, int* s_cnt, int* s_fail) {
    const int tid = threadIdx.x;
    int phase = 0;
    for (int level = 0; ncur > 0 && level < max_levels; ++level) {
        int nwarps = 32 >> level;
        if (level > 5 || nwarps < 4) nwarps = 1;     // below 4 warps per node: one warp per node
        const int ngroups = 32 / nwarps;
        const int grp = (tid >> 5) / nwarps;
        const int gt = tid - grp * nwarps * 32;
        const int bar_id = 1 + grp;                  // <= 8 multi-warp groups -> ids 1..8
        phase = 0;                                   // (groups were re-formed: restart the scan buffers' alternation)
        for (int base = 0; base < ncur; base += ngroups) {
            const int k = base + grp;
            if (k < ncur) {
                const int s = cur[3 * k], e = cur[3 * k + 1];
                int dim = cur[3 * k + 2];
                const int p = build_node_sm(M, xyz, s, e, &dim, gt, nwarps, bar_id, s_ws + (grp & 7) * 64, phase, s_wd + (grp & 7) * 192);
                if (p == -2) { if (gt == 0) *s_fail = 1; }
                else if (p >= 0 && gt == 0) {
                    // children that are leaves need no visit
                    const int nl = (p - s > kd::LEAFSIZE) ? 1 : 0, nr = (e - p > kd::LEAFSIZE) ? 1 : 0;
                    if (nl + nr) {
                        int o = atomicAdd(s_cnt, nl + nr);
                        if (nl) { nxt[3 * o] = s; nxt[3 * o + 1] = p; nxt[3 * o + 2] = dim; ++o; }
                        if (nr) { nxt[3 * o] = p; nxt[3 * o + 1] = e; nxt[3 * o + 2] = dim; }
                    }
                }
            }
        }
        __syncthreads();
        ncur = *s_cnt;
        const int fail = *s_fail;
        __syncthreads();
        if (fail) return -1;
        if (tid == 0) *s_cnt = 0;
        int* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
    return ncur;
}

// Whole build by one CTA of THREADS threads; writes rank[point] = position in cKDTree's `indices`.
// smem: 12*n bytes (+ padding), nodes_a/nodes_b: global scratch of 3*(n/8+4) ints each.  Returns false (on every
// thread) if the caller has to fall back.
static __device__ bool build_ranks_block(const double* x, const double* y, const double* z, int n, int* rank, char* smem,
                                         int* node_a, int* node_b, int* s_ws /*kdpar::WS_INTS*/, double* s_wd /*kdpar::WD_DOUBLES*/) {
    __shared__ int s_cnt, s_fail;
    View M;
    M.key = reinterpret_cast<double*>(smem);
    M.idx = reinterpret_cast<unsigned short*>(M.key + n);
    M.posR = M.idx + n;
    const double* const xyz[3] = {x, y, z};
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) M.idx[i] = (unsigned short)i;
    int* cur = node_a;
    int* nxt = node_b;
    if (tid == 0) { cur[0] = 0; cur[1] = n; cur[2] = -1; s_cnt = 0; s_fail = 0; }
    __syncthreads();
    if (run_levels(M, xyz, cur, nxt, 1, 1 << 30, s_ws, s_wd, &s_cnt, &s_fail) < 0) return false;
    for (int i = tid; i < n; i += blockDim.x) rank[M.idx[i]] = i;
    return true;
}

// The same build split over several CTAs (the build of one graph keeps the issue slots of its SM 46 % busy for 350 us, and it
// sits on the critical path of the iteration): build_top_block runs the first `top_levels` levels, leaves the permutation so
// far in idx16 (global), provisional ranks in `rank`, and the nodes of the next level in top[1..] (top[0] = their number, -1 =
// fall back to the one-CTA build); build_sub_block finishes ONE of those nodes -- a subtree is independent of its siblings --
// in its own shared memory and writes the final ranks of its range.
static __device__ void build_top_block(const double* x, const double* y, const double* z, int n, int top_levels, int* rank,
                                       unsigned short* idx16, int* top /*1 + 3 * 2^top_levels ints*/, char* smem, int* node_a, int* node_b,
                                       int* s_ws, double* s_wd) {
    __shared__ int s_cnt, s_fail;
    View M;
    M.key = reinterpret_cast<double*>(smem);
    M.idx = reinterpret_cast<unsigned short*>(M.key + n);
    M.posR = M.idx + n;
    const double* const xyz[3] = {x, y, z};
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) M.idx[i] = (unsigned short)i;
    int* cur = node_a;
    int* nxt = node_b;
    if (tid == 0) { cur[0] = 0; cur[1] = n; cur[2] = -1; s_cnt = 0; s_fail = 0; }
    __syncthreads();
    const int left = run_levels(M, xyz, cur, nxt, 1, top_levels, s_ws, s_wd, &s_cnt, &s_fail);
    if (left < 0) { if (tid == 0) top[0] = -1; return; }
    for (int i = tid; i < n; i += blockDim.x) { const unsigned short a = M.idx[i]; idx16[i] = a; rank[a] = i; }
    if (tid == 0) top[0] = left;
    for (int i = tid; i < 3 * left; i += blockDim.x) top[1 + i] = cur[i];
}

static __device__ bool build_sub_block(const double* x, const double* y, const double* z, int start, int end, int* rank,
                                       const unsigned short* idx16, char* smem, int* node_a, int* node_b, int* s_ws, double* s_wd) {
    __shared__ int s_cnt, s_fail;
    const int n = end - start;
    View M;                                          // (positions are absolute: the arrays are addressed with their offset)
    M.key = reinterpret_cast<double*>(smem) - start;
    M.idx = reinterpret_cast<unsigned short*>(reinterpret_cast<double*>(smem) + n) - start;
    M.posR = M.idx + n;
    const double* const xyz[3] = {x, y, z};
    const int tid = threadIdx.x;
    for (int i = start + tid; i < end; i += blockDim.x) M.idx[i] = idx16[i];
    int* cur = node_a;
    int* nxt = node_b;
    if (tid == 0) { cur[0] = start; cur[1] = end; cur[2] = -1; s_cnt = 0; s_fail = 0; }
    __syncthreads();
    if (run_levels(M, xyz, cur, nxt, 1, 1 << 30, s_ws, s_wd, &s_cnt, &s_fail) < 0) return false;
    for (int i = start + tid; i < end; i += blockDim.x) rank[M.idx[i]] = i;
    return true;
}

}  // namespace kdsm
}  // namespace octa
