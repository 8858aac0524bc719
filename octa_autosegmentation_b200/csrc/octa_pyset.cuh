// CPython 3.12 `set` emulation (Objects/setobject.c) for the O2 -> CO2 conversion of greenhouse.py:100-111, and the test that
// decides whether the exact cKDTree ball order is needed.  Host/device shared so the CPU test suite can drive it against
// real CPython sets (octa_testhooks.cu).
#pragma once
#include <stddef.h>
#include "octa_grow.cuh"

#ifdef __CUDACC__
#define OCTA_PS_HD __host__ __device__
#else
#define OCTA_PS_HD
#endif

namespace octa {

OCTA_PS_HD inline void pyset_insert_clean(long long* th, int* tk, size_t mask, int key, long long hash) {
    size_t perturb = (size_t)hash, i = (size_t)hash & mask;
    while (true) {
        size_t e = i;
        if (tk[e] < 0) { tk[e] = key; th[e] = hash; return; }
        if (i + 9 <= mask) {
            for (int j = 0; j < 9; ++j) { ++e; if (tk[e] < 0) { tk[e] = key; th[e] = hash; return; } }
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

// CPython set emulation (Objects/setobject.c, 3.12) run by ONE thread -> iteration order = slot order.  Tables of up to
// SM_TBL slots live in shared memory (the usual case: tens of insertions); larger ones spill to global memory.
constexpr int SM_TBL = 1024;
struct KillShared {
    double nxs[512], nys[512], nzs[512];
    long long th[2][SM_TBL];
    int tk[2][SM_TBL];
    int tabinfo[4];                 // which table holds the result (0/1 smem, 2/3 global), mask, err, order-sensitive flag
};

struct PySetDev {
    KillShared* sh;
    long long* gth; int* gtk;       // global tables 2/3 (SET_TBL slots each)
    int cur;                        // 0/1: shared tables, 2/3: global tables
    size_t mask, fill, used;
    long long* curh; int* curk;
    int err;
    OCTA_PS_HD long long* tabh(int t) const { return t < 2 ? sh->th[t] : gth + (size_t)(t - 2) * SET_TBL; }
    OCTA_PS_HD int* tabk(int t) const { return t < 2 ? sh->tk[t] : gtk + (size_t)(t - 2) * SET_TBL; }
    OCTA_PS_HD void init() {       // (slots 0..7 of table 0 were cleared by the block)
        cur = 0; mask = 7; fill = 0; used = 0; curh = tabh(0); curk = tabk(0); err = 0;
    }
    // set_add_entry + set_table_resize.  Returns the slot the key was written to in the table that was current when the call
    // started (-1: duplicate); *resized: the table grew after the insertion; *same_ball: an examined slot holds a key whose
    // ball id (hitj) equals `ball` (only tracked when hitj != nullptr).
    OCTA_PS_HD int add(int key, long long hash, bool* resized, const int* hitj = nullptr, int ball = -1, bool* same_ball = nullptr) {
        size_t perturb = (size_t)hash, i = (size_t)hash & mask;
        while (true) {
            size_t e = i;
            int probes = (i + 9 <= mask) ? 9 : 0;
            do {
                if (curk[e] < 0) {
                    curk[e] = key; curh[e] = hash;
                    ++fill; ++used;
                    if (fill * 5 >= mask * 3) {
                        const size_t minused = used > 50000 ? used * 2 : used * 4;
                        size_t newsize = 8;
                        while (newsize <= minused) newsize <<= 1;
                        if (newsize > (size_t)SET_TBL) { err = 5; return (int)e; }
                        const int alt = newsize <= (size_t)SM_TBL ? (cur == 0 ? 1 : 0) : (cur == 2 ? 3 : 2);
                        long long* alth = tabh(alt); int* altk = tabk(alt);
                        for (size_t z = 0; z < newsize; ++z) { altk[z] = -1; alth[z] = 0; }
                        for (size_t z = 0; z <= mask; ++z)
                            if (curk[z] >= 0) pyset_insert_clean(alth, altk, newsize - 1, curk[z], curh[z]);
                        cur = alt; curh = alth; curk = altk;
                        mask = newsize - 1;
                        fill = used;
                        if (resized) *resized = true;
                    }
                    return (int)e;
                }
                if (curh[e] == hash && curk[e] == key) return -1;
                if (hitj && hitj[curk[e]] == ball) *same_ball = true;
                ++e;
            } while (probes--);
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
};

// Insertion of the sequence seq[0..T) (hashes sh[0..T)).  DETECT: the sequence is in list-index order inside every ball
// (greenhouse.py:100-110 receives the hits of one ball in cKDTree order, which costs a whole kd build to know); returns true
// when the final table could depend on the order inside some ball, i.e. when the exact order is needed:
//   * a ball whose keys never examine a slot held by a key of the same ball, with no table growth before its last key,
//     leaves the same table for every order (each key lands on the first free slot of its own probe sequence, and those
//     slots are distinct);
//   * otherwise every order of the ball's keys (<= 4 keys: <= 24 orders) is replayed from the state before the ball and the
//     resulting tables are compared; all equal -> the order is irrelevant, go on.
// Keys are distinct sinks and each belongs to exactly one ball (hitj = first new node that hits it).
template <bool DETECT>
OCTA_PS_HD inline bool pyset_run(PySetDev& ps, const int* seq, const long long* sh, int T, const int* hitj, int* scratch_keys) {
    int q = 0;
    while (q < T && !ps.err) {
        if (!DETECT) { ps.add(seq[q], sh[q], nullptr); ++q; continue; }
        const int ball = hitj[seq[q]];
        int q2 = q + 1;
        while (q2 < T && hitj[seq[q2]] == ball) ++q2;
        const int m = q2 - q;
        if (m == 1) { ps.add(seq[q], sh[q], nullptr); q = q2; continue; }
        // pass in index order, with tracking
        const int cur0 = ps.cur;
        const size_t mask0 = ps.mask, fill0 = ps.fill, used0 = ps.used;
        int wslot[4];                        // slots written in table cur0 (before any growth)
        int nlog = 0, grew_at = -1;
        bool inter = false;
        for (int k = 0; k < m && !ps.err; ++k) {
            bool grew = false;
            const bool before_growth = grew_at < 0;
            const int slot = ps.add(seq[q + k], sh[q + k], &grew, hitj, ball, &inter);
            if (before_growth && k < 4) wslot[nlog++] = slot;
            if (grew) { if (grew_at >= 0) grew_at = -2; else grew_at = k; }        // -2: grew twice
        }
        if (ps.err) return false;
        if (!inter && (grew_at == -1 || grew_at == m - 1)) { q = q2; continue; }
        if (m > 4 || grew_at == -2 || cur0 >= 2 || ps.cur >= 2) return true;          // not worth a closure: ask for the exact order
        // reference result: slots of the ball's keys (no growth inside the ball), or the whole new table
        const bool whole = grew_at >= 0 && grew_at < m - 1;
        int ref_slot[4];
        const int nslots = (int)ps.mask + 1;
        if (whole) { for (int z = 0; z < nslots; ++z) scratch_keys[z] = ps.curk[z]; }
        else { for (int k = 0; k < nlog; ++k) ref_slot[k] = wslot[k]; }
        const int cur1 = ps.cur;
        int perm[4] = {0, 1, 2, 3};
        bool differs = false;
        while (!differs) {
            // next permutation of perm[0..m) (lexicographic); done when none is left
            int i = m - 2;
            while (i >= 0 && perm[i] > perm[i + 1]) --i;
            if (i < 0) break;
            int j = m - 1;
            while (perm[j] < perm[i]) --j;
            { const int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
            for (int a = i + 1, b = m - 1; a < b; ++a, --b) { const int t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
            // back to the state before the ball: undo the writes into table cur0 (a growth rewrites the other table completely)
            int* k0 = ps.tabk(cur0);
            for (int k = 0; k < nlog; ++k) k0[wslot[k]] = -1;
            ps.cur = cur0; ps.mask = mask0; ps.fill = fill0; ps.used = used0; ps.curh = ps.tabh(cur0); ps.curk = k0;
            nlog = 0;
            bool g2 = false;
            int slot_of[4];
            for (int k = 0; k < m; ++k) {
                const bool before_growth = !g2;
                const int slot = ps.add(seq[q + perm[k]], sh[q + perm[k]], &g2);
                if (before_growth) wslot[nlog++] = slot;
                slot_of[perm[k]] = slot;
            }
            if (ps.cur != cur1) { differs = true; break; }     // (cannot happen: growth depends on counts only)
            if (whole) { for (int z = 0; z < nslots; ++z) if (scratch_keys[z] != ps.curk[z]) { differs = true; break; } }
            else { for (int k = 0; k < m; ++k) if (slot_of[k] != ref_slot[k]) { differs = true; break; } }
        }
        if (differs) return true;
        q = q2;
    }
    return false;
}

}  // namespace octa
