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

// ------------------------------------------------------------------------------------------------------------------------
// Multi-state order test (T <= MS_MAXT keys, tables of up to MS_TBL slots).  pyset_run<true> above decides ball by ball and
// gives up as soon as two orders of a ball leave different tables -- but most of those differences sit in the small early
// tables (8 and 32 slots: the first 19 keys) and are erased by the next growth, which re-inserts the keys in slot order into a
// table four times the size.  This test carries every table the unknown orders can have produced so far (at most MS_ALT of
// them; all share mask and fill, growth depends on counts only), merges equal ones, and asks for the exact order only when
//   * more than one table is alive and no growth is left (final fill = T is known), or at the end,
//   * more than MS_ALT tables are alive, or a ball of more than 4 keys would have to be permuted.
// Tables hold POSITIONS in the insertion sequence as 16-bit integers (key = seq[pos], hash = hs[pos], ball = bl[pos]).
constexpr int MS_TBL = 512;
constexpr int MS_MAXT = 306;            // the 307th key grows the table to 2048 slots
constexpr int MS_ALT = 6;
constexpr int MS_TABLES = 2 * MS_ALT + 2;

OCTA_PS_HD inline int ms_insert(short* tab, unsigned mask, int pos, long long hash, const int* bl, int ball, bool* same_ball) {
    unsigned long long perturb = (unsigned long long)hash;
    unsigned i = (unsigned)perturb & mask;
    while (true) {
        unsigned e = i;
        int probes = (i + 9 <= mask) ? 9 : 0;
        do {
            const int k = tab[e];
            if (k < 0) { tab[e] = (short)pos; return (int)e; }
            if (bl && bl[k] == ball) *same_ball = true;
            ++e;
        } while (probes--);
        perturb >>= 5;
        i = (i * 5u + 1u + (unsigned)perturb) & mask;       // (only the bits under the mask matter)
    }
}

// tables are 4-byte aligned and hold an even number of cells: copies and comparisons go word by word
OCTA_PS_HD inline void ms_copy(short* dst, const short* src, unsigned n) {
    unsigned* d = reinterpret_cast<unsigned*>(dst);
    const unsigned* c = reinterpret_cast<const unsigned*>(src);
    for (unsigned z = 0; z < n / 2; ++z) d[z] = c[z];
}
OCTA_PS_HD inline bool ms_equal(const short* a, const short* b, unsigned n) {
    const unsigned* x = reinterpret_cast<const unsigned*>(a);
    const unsigned* y = reinterpret_cast<const unsigned*>(b);
    for (unsigned z = 0; z < n / 2; ++z) if (x[z] != y[z]) return false;
    return true;
}

// positions q + perm[0..m) into `tab` (set_add_entry + set_table_resize; `tmp` receives a grown table first)
OCTA_PS_HD inline void ms_apply(short* tab, short* tmp, unsigned* mask, unsigned* fill, const long long* hs, const int* bl, int q,
                                const int* perm, int m, bool track, bool* inter, bool* grew_inside, bool* grew) {
    for (int k = 0; k < m; ++k) {
        const int pos = q + (perm ? perm[k] : k);          // perm == nullptr: index order
        ms_insert(tab, *mask, pos, hs[pos], track ? bl : nullptr, bl[q], inter);
        ++*fill;
        if (*fill * 5 >= *mask * 3) {
            unsigned newsize = 8;
            while (newsize <= *fill * 4) newsize <<= 1;
            unsigned* t32 = reinterpret_cast<unsigned*>(tmp);
            for (unsigned z = 0; z < newsize / 2; ++z) t32[z] = 0xffffffffu;
            for (unsigned z = 0; z <= *mask; ++z)
                if (tab[z] >= 0) ms_insert(tmp, newsize - 1, tab[z], hs[tab[z]], nullptr, 0, nullptr);
            ms_copy(tab, tmp, newsize);
            *mask = newsize - 1;
            *grew = true;
            if (k < m - 1) *grew_inside = true;
        }
    }
}

// hs / bl: hash and ball id per sequence position (T <= MS_MAXT), tabs: MS_TABLES tables of MS_TBL shorts.  Returns 0 and the
// final table in tabs[0 .. *mask_out] (positions, -1 = empty) when the result does not depend on the order inside any ball,
// else a reason code > 0 (1: ball of > 4 keys, 2: too many alive tables, 3: alive tables differ with no growth left).
// The usual ball (no growth inside it, no key examining a slot of a key of the same ball) is inserted IN PLACE into every alive
// table; tables are only copied where a growth falls inside a ball or the orders of a ball have to be tried.
OCTA_PS_HD inline int pyset_run_multi(const long long* hs, const int* bl, int T, short* tabs, int* mask_out) {
    short* A = tabs;                               // alive tables
    short* B = tabs + (size_t)MS_ALT * MS_TBL;     // tables after the current ball
    short* cand = tabs + (size_t)2 * MS_ALT * MS_TBL;
    short* tmp = cand + MS_TBL;
    unsigned mask = 7, fill = 0;
    int alive = 1;
    for (int z = 0; z < 8; ++z) A[z] = -1;
    int q = 0;
    while (q < T) {
        const int ball = bl[q];
        int q2 = q + 1;
        while (q2 < T && bl[q2] == ball) ++q2;
        const int m = q2 - q;
        unsigned mask1 = mask, fill1 = fill;
        bool expand = (fill + (unsigned)m) * 5 >= mask * 3;          // a growth falls into this ball: tables are rebuilt anyway
        if (!expand) {
            // in place, index order, with the log of the written slots (to take the ball back if its orders have to be tried)
            int slots[MS_ALT][4];
            unsigned inter_mask = 0;
            for (int s = 0; s < alive; ++s) {
                short* tab = A + (size_t)s * MS_TBL;
                bool inter = false;
                for (int k = 0; k < m; ++k) {
                    const int e = ms_insert(tab, mask, q + k, hs[q + k], m > 1 ? bl : nullptr, ball, &inter);
                    if (m <= 4) slots[s][k] = e;
                }
                if (inter) inter_mask |= 1u << s;
            }
            fill1 = fill + (unsigned)m;
            if (inter_mask) {
                if (m > 4) return 1;
                for (int s = 0; s < alive; ++s) {                    // take the ball back where its order may matter
                    if (!((inter_mask >> s) & 1u)) continue;
                    short* tab = A + (size_t)s * MS_TBL;
                    for (int k = 0; k < m; ++k) tab[slots[s][k]] = -1;
                }
                expand = true;
            }
            if (!expand) { mask = mask1; fill = fill1; q = q2; continue; }
            // fall through to the general path: tables without `inter` already hold the ball (marked in done_mask)
            int nb = 0;
            auto push = [&](const short* t, unsigned n) -> bool {
                for (int u = 0; u < nb; ++u) if (ms_equal(B + (size_t)u * MS_TBL, t, n)) return true;
                if (nb == MS_ALT) return false;
                ms_copy(B + (size_t)nb * MS_TBL, t, n);
                ++nb;
                return true;
            };
            for (int s = 0; s < alive; ++s) {
                const short* src = A + (size_t)s * MS_TBL;
                if (!((inter_mask >> s) & 1u)) { if (!push(src, mask + 1)) return 2; continue; }
                int perm[4] = {0, 1, 2, 3};
                while (true) {
                    ms_copy(cand, src, mask + 1);
                    unsigned mk = mask, fl = fill;
                    bool i2 = false, g2 = false, g3 = false;
                    ms_apply(cand, tmp, &mk, &fl, hs, bl, q, perm, m, false, &i2, &g2, &g3);
                    if (!push(cand, mk + 1)) return 2;
                    int i = m - 2;
                    while (i >= 0 && perm[i] > perm[i + 1]) --i;
                    if (i < 0) break;
                    int j = m - 1;
                    while (perm[j] < perm[i]) --j;
                    { const int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
                    for (int a = i + 1, b = m - 1; a < b; ++a, --b) { const int t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
                }
            }
            { short* t = A; A = B; B = t; }
            alive = nb;
        } else if (m == 1) {
            bool grew_any = false;
            for (int s = 0; s < alive; ++s) {
                unsigned mk = mask, fl = fill;
                bool gi = false, gr = false;
                ms_apply(A + (size_t)s * MS_TBL, tmp, &mk, &fl, hs, bl, q, nullptr, 1, false, nullptr, &gi, &gr);
                mask1 = mk; fill1 = fl; grew_any = gr;
            }
            if (grew_any && alive > 1) {            // a growth may have erased the differences: merge equal tables
                int keep = 0;
                for (int s = 0; s < alive; ++s) {
                    bool dup = false;
                    for (int u = 0; u < keep && !dup; ++u) dup = ms_equal(A + (size_t)u * MS_TBL, A + (size_t)s * MS_TBL, mask1 + 1);
                    if (!dup) {
                        if (keep != s) ms_copy(A + (size_t)keep * MS_TBL, A + (size_t)s * MS_TBL, mask1 + 1);
                        ++keep;
                    }
                }
                alive = keep;
            }
        } else {
            int nb = 0;
            auto push = [&](const short* t, unsigned n) -> bool {        // false: too many alive tables
                for (int u = 0; u < nb; ++u) if (ms_equal(B + (size_t)u * MS_TBL, t, n)) return true;
                if (nb == MS_ALT) return false;
                ms_copy(B + (size_t)nb * MS_TBL, t, n);
                ++nb;
                return true;
            };
            for (int s = 0; s < alive; ++s) {
                const short* src = A + (size_t)s * MS_TBL;
                // index order with tracking: no key of the ball examines a slot held by a key of the same ball and no growth
                // before its last key -> each key lands on the first free slot of its own probe sequence, whatever the order
                ms_copy(cand, src, mask + 1);
                unsigned mk = mask, fl = fill;
                bool inter = false, gi = false, gr = false;
                ms_apply(cand, tmp, &mk, &fl, hs, bl, q, nullptr, m, true, &inter, &gi, &gr);
                mask1 = mk; fill1 = fl;
                if (!push(cand, mk + 1)) return 2;
                if (!inter && !gi) continue;
                if (m > 4) return 1;
                int perm[4] = {0, 1, 2, 3};
                while (true) {
                    int i = m - 2;
                    while (i >= 0 && perm[i] > perm[i + 1]) --i;
                    if (i < 0) break;
                    int j = m - 1;
                    while (perm[j] < perm[i]) --j;
                    { const int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
                    for (int a = i + 1, b = m - 1; a < b; ++a, --b) { const int t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
                    ms_copy(cand, src, mask + 1);
                    mk = mask; fl = fill;
                    bool i2 = false, g2 = false, g3 = false;
                    ms_apply(cand, tmp, &mk, &fl, hs, bl, q, perm, m, false, &i2, &g2, &g3);
                    if (!push(cand, mk + 1)) return 2;
                }
            }
            { short* t = A; A = B; B = t; }
            alive = nb;
        }
        mask = mask1; fill = fill1;
        if (alive > 1 && !((unsigned)T * 5 >= mask * 3)) return 3;      // no growth left that could merge them
        q = q2;
    }
    if (alive > 1) return 3;
    if (A != tabs) ms_copy(tabs, A, mask + 1);
    *mask_out = (int)mask;
    return 0;
}

}  // namespace octa
