// Result ORDER of scipy.spatial.cKDTree.query_ball_point as used by the reference
// (element_mesh.py:136-137, unsorted single-point query): hits come back in ascending position of
// `tree.indices`, the permutation left behind by cKDTree's balanced build (scipy/spatial/ckdtree/src/build.cxx):
//   build(start, end): stop at <= leafsize (16) points; tight bounding box of the node's points; split dimension =
//   first dimension of strictly largest extent; std::nth_element(idx+start, idx+start+n/2, idx+end, cmp) with
//   cmp(a,b) = x[a][d] == x[b][d] ? a < b : x[a][d] < x[b][d]; children [start, start+n/2) and [start+n/2, end).
// The order inside a leaf is whatever the nth_element calls of its ancestors left behind, so this header restates
// libstdc++'s std::nth_element (bits/stl_algo.h: __introselect -> __unguarded_partition_pivot ->
// __move_median_to_first / __unguarded_partition, __heap_select on depth exhaustion, __insertion_sort on <= 3)
// operation for operation.  Written from the published algorithm; checked against scipy's tree.indices
// (tests/test_kdorder.py).  Host/device code; the device uses it per (small) range, see octa_grow_kernels.cu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OCTA_KD_HD __host__ __device__
#else
#define OCTA_KD_HD
#endif

namespace octa {
namespace kd {

constexpr int LEAFSIZE = 16;

// comparator on one coordinate with index tie-break
struct Less {
    const double* v;   // coordinate `d` of every point (SoA)
    OCTA_KD_HD bool operator()(int a, int b) const {
        const double pa = v[a], pb = v[b];
        return pa == pb ? a < b : pa < pb;
    }
};

OCTA_KD_HD inline void swap_i(int* a, int* b) { const int t = *a; *a = *b; *b = t; }

OCTA_KD_HD inline void move_median_to_first(int* result, int* a, int* b, int* c, const Less& less) {
    if (less(*a, *b)) {
        if (less(*b, *c)) swap_i(result, b);
        else if (less(*a, *c)) swap_i(result, c);
        else swap_i(result, a);
    } else if (less(*a, *c)) swap_i(result, a);
    else if (less(*b, *c)) swap_i(result, c);
    else swap_i(result, b);
}

OCTA_KD_HD inline int* unguarded_partition(int* first, int* last, int* pivot, const Less& less) {
    while (true) {
        while (less(*first, *pivot)) ++first;
        --last;
        while (less(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swap_i(first, last);
        ++first;
    }
}

OCTA_KD_HD inline void push_heap_(int* first, long hole, long top, int value, const Less& less) {
    long parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

OCTA_KD_HD inline void adjust_heap(int* first, long hole, long len, int value, const Less& less) {
    const long top = hole;
    long second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (less(first[second], first[second - 1])) --second;
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    push_heap_(first, hole, top, value, less);
}

OCTA_KD_HD inline void heap_select(int* first, int* middle, int* last, const Less& less) {
    const long len = middle - first;
    if (len >= 2) {                     // __make_heap
        long parent = (len - 2) / 2;
        while (true) {
            const int value = first[parent];
            adjust_heap(first, parent, len, value, less);
            if (parent == 0) break;
            --parent;
        }
    }
    for (int* i = middle; i < last; ++i)
        if (less(*i, *first)) {         // __pop_heap(first, middle, i)
            const int value = *i;
            *i = *first;
            adjust_heap(first, 0, len, value, less);
        }
}

OCTA_KD_HD inline void insertion_sort(int* first, int* last, const Less& less) {
    if (first == last) return;
    for (int* i = first + 1; i != last; ++i) {
        const int val = *i;
        if (less(val, *first)) {
            for (int* p = i; p != first; --p) *p = *(p - 1);
            *first = val;
        } else {                        // __unguarded_linear_insert
            int* next = i - 1;
            int* cur = i;
            while (less(val, *next)) { *cur = *next; cur = next; --next; }
            *cur = val;
        }
    }
}

OCTA_KD_HD inline int floor_log2(long n) { int k = 0; while (n > 1) { n >>= 1; ++k; } return k; }

// std::nth_element(first, nth, last, less)
OCTA_KD_HD inline void nth_element(int* first, int* nth, int* last, const Less& less) {
    if (first == last || nth == last) return;
    int depth_limit = floor_log2(last - first) * 2;
    while (last - first > 3) {
        if (depth_limit == 0) {
            heap_select(first, nth + 1, last, less);
            swap_i(first, nth);
            return;
        }
        --depth_limit;
        int* mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1, less);
        int* cut = unguarded_partition(first + 1, last, first, less);
        if (cut <= nth) first = cut; else last = cut;
    }
    insertion_sort(first, last, less);
}

// One node of cKDTree's build on idx[start, end): returns the split position p (children [start,p), [p,end)) or -1
// for a leaf.  xyz: the three coordinate arrays.
OCTA_KD_HD inline int build_node(const double* const xyz[3], int* idx, int start, int end) {
    if (end - start <= LEAFSIZE) return -1;
    double mins[3], maxes[3];
    for (int k = 0; k < 3; ++k) mins[k] = maxes[k] = xyz[k][idx[start]];
    for (int j = start + 1; j < end; ++j)
        for (int k = 0; k < 3; ++k) {
            const double t = xyz[k][idx[j]];
            maxes[k] = maxes[k] > t ? maxes[k] : t;
            mins[k] = mins[k] < t ? mins[k] : t;
        }
    int d = 0;
    double size = 0;
    for (int k = 0; k < 3; ++k)
        if (maxes[k] - mins[k] > size) { d = k; size = maxes[k] - mins[k]; }
    if (maxes[d] == mins[d]) return -1;
    const Less less{xyz[d]};
    const int n = end - start;
    nth_element(idx + start, idx + start + n / 2, idx + end, less);
    int p = start + n / 2;
    const double split = xyz[d][idx[p]];
    // cKDTree's own partition loop: a no-op after nth_element unless values tie with the split value
    p = start;
    int q = end - 1;
    while (p <= q) {
        if (xyz[d][idx[p]] < split) ++p;
        else if (xyz[d][idx[q]] >= split) --q;
        else { swap_i(idx + p, idx + q); ++p; --q; }
    }
    if (p == start) {                   // slide midpoint: no point below the split
        int j = start;
        double s = xyz[d][idx[j]];
        for (int i = start + 1; i < end; ++i)
            if (xyz[d][idx[i]] < s) { j = i; s = xyz[d][idx[j]]; }
        swap_i(idx + start, idx + j);
        p = start + 1;
    } else if (p == end) {
        int j = end - 1;
        double s = xyz[d][idx[j]];
        for (int i = start; i < end - 1; ++i)
            if (xyz[d][idx[i]] > s) { j = i; s = xyz[d][idx[j]]; }
        swap_i(idx + end - 1, idx + j);
        p = end - 1;
    }
    return p;
}

// sequential build of the whole permutation (explicit stack; left subtree first like the recursion)
OCTA_KD_HD inline void build_indices_seq(const double* x, const double* y, const double* z, int n, int* idx) {
    const double* const xyz[3] = {x, y, z};
    for (int i = 0; i < n; ++i) idx[i] = i;
    int stack_s[64], stack_e[64], sp = 0;
    stack_s[0] = 0; stack_e[0] = n; sp = 1;
    while (sp > 0) {
        --sp;
        const int s = stack_s[sp], e = stack_e[sp];
        const int p = build_node(xyz, idx, s, e);
        if (p < 0) continue;
        // (the two children are independent; order of processing does not matter for the permutation)
        stack_s[sp] = p; stack_e[sp] = e; ++sp;
        stack_s[sp] = s; stack_e[sp] = p; ++sp;
    }
}

}  // namespace kd
}  // namespace octa
