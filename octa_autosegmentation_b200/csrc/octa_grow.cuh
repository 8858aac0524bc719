// Shared declarations of the growth kernels (octa_grow_kernels.cu) and their host driver
// (octa_grow_host.cu).  Layout in HBM: structure-of-arrays, one slab of `cap` entries per graph, so a
// warp walking consecutive nodes / sinks of one graph issues fully coalesced 8-byte loads and the
// whole working set of a 64-graph batch (~100 MB) stays resident in the 126 MB L2.
#pragma once
#include <stdint.h>
#include "octa_rng.h"

namespace octa {

constexpr int GEOMETRY_SIZE = 76;                 // simulation_space.py:8
constexpr int MAX_VALID = GEOMETRY_SIZE * GEOMETRY_SIZE;
constexpr int GRID = 64;                          // bucket grid cells per axis over the unit square
constexpr int SET_TBL = 16384;                    // slots per CPython-set emulation table (x2: resize target)

// per-iteration parameters (identical for every graph of the batch; computed on the host exactly
// like greenhouse.py:34-51 / :139-147 evolve them)
struct IterP {
    double eps_n_eff, eps_s, eps_k, delta[2], gamma[2], phi, omega, kappa, d, r, rotation_radius;
    double faz_cx, faz_cy, param_scale, shape[3];
    int N, t, first_mode, mode_idx, iter;
    int geom_gs;         // fixed sampling geometry: geometry_size = max(mask.shape) (simulation_space.py:30), 0 = none
    int geom_dims[3];    // shape of the mask (GrowDev::geom_mask, C order)
    double kap_tab[9];   // kappa of the node's creation mode (arterial_tree.py:32); [8] = 4, the add_node default of the stumps
    double leafc_tab[9]; // r ** kap_tab[q]: what a fresh leaf contributes to a parent of creation mode q (see GrowDev::ncon)
};

enum PropType : int { P_NONE = 0, P_LEAF_ELONG = 1, P_LEAF_DRAW = 2, P_LEAF_BIF = 3, P_INTER_DRAW = 4, P_INTER_EMPTY = 5 };

struct Proposal {
    int type;
    int cond;          // leaf: angle(vtc, avg) > 90 ; inter: angle(vtc, avg) <= 90
    double ratio5;     // (dist_to_center / (2 FAZ_radius)) ** 5
    double c_used;     // inter-node: ncon value of the distal child the evaluation was made with
    double p[3];       // elongation / sprout position
    double b1[3], b2[3];  // bifurcation children
};

struct __align__(16) ActDec {  // decision record of one dict entry the replay has to look at
    int e, nd, type, cond;
    double ratio5, c_used;
};

struct GrowShape {
    int G, capN, capS, Nmax, pycap;
    int exact_ball_order;   // cKDTree's index permutation for the O2->CO2 insertion order: 2 = built on demand (exact, default), 1 = built every iteration (exact), 0 = list-index order instead (diagnostics)
    int commit_smem; // bytes of dynamic shared memory of k_commit (tree mirror + decision records)
    int kill_rcap;   // hits per call k_kill keeps as a sorted list (<= KILL_RCAP = 4096; beyond: block scans over the sink list; tests shrink it)
    int geom_cells, geom_nvalid;   // fixed sampling geometry: voxels of the mask and how many of them are set (0 = no mask)
};

struct GrowDev {
    // vessel nodes, [f][g*capN + i]
    double *nx[2], *ny[2], *nz[2];
    // Steering radii are kept as Murray CONTRIBUTIONS: ncon[n] = radius(n) ** kappa(parent(n)), so that a parent's
    // update (arterial_tree.py:174-184, r_p = (sum r_c**k)**(1/k)) is a plain sum of its children's ncon and a
    // pow is needed only where parent and grandparent were created in modes of different kappa, and when a
    // radius itself is wanted (k_eval / k_commit: the distal radius of an inter-node).  These values only steer
    // directions; the radii that get PRINTED are recomputed on the host with libm pow (octa_grow_host.cu).
    double *ncon[2];
    int *npar[2], *nch0[2], *nch1[2];
    unsigned char *nnch[2], *nmeta[2], *deact[2];      // deact: the node branched and left the ACTIVE set
    int *n_nodes[2], *n_prev[2];
    // number of ACTIVE nodes (node array minus deact marks), refreshed by k_prepare for the byte accounting
    int *n_act[2];
    // sinks: [0] oxygen sinks, [1] CO2 sources
    double *sx[2], *sy[2], *sz[2];
    int *n_s[2];
    // bucket grids: [0] arterial nodes (+radius), [1] O2 sinks, [2]/[3] active arterial / venous nodes, [4] all venous nodes
    double *gx[5], *gy[5], *gz[5], *gr[5];
    int *gi[5], *gcell[5];
    // RNG streams
    MTState *np_mt, *py_mt;
    unsigned int* py_buf;
    int *py_n, *py_pos;
    long long* py_draws;
    // per-graph constants
    double* faz_radius;
    int* n_valid;
    unsigned char* valid_ij;
    unsigned char* geom_mask;   // bytes of SimulationSpace.oxygen_sample_geometry_path, C order [d0][d1][d2] (one copy per context), or unused
    unsigned short* geom_valid; // np.argwhere(mask): (i, j, k) per set voxel, C order, shared by every graph of the context
    // scratch
    unsigned int *vi, *ubuf;
    double *cx, *cy, *cz;
    int* n_cand;
    unsigned char *cpass;
    unsigned int* cstate32;   // k_sink_greedy decision cells when they do not fit in shared memory
    int* plist;
    int *assign, *first, *cnt, *slot, *cur;
    int *dict_node, *n_dict, *list_off, *list, *sc_idx;
    double *sc_ang;
    double *sc_inter;     // per (inter-node, attractor): angle to distal / proximal segment and unit vector (5 doubles)
    Proposal* prop;
    ActDec* adec;         // decision records (global fallback when they do not fit in k_commit's shared memory)
    int4* newl;
    int *alist;           // k_commit scratch: start nodes of the bottom-up refresh
    unsigned int* cbits;  // k_commit scratch, [g][4][capN/32]: dirty / arrival / inter / tag bitmaps when the tree does not fit in shared memory
    int *hitj, *hl, *ta, *seq;
    int *kd_idx, *kd_posL, *kd_posR, *kd_rank, *kd_nodes;
    // on-demand exact ball order: per graph "k_kill left the arterial kill to k_kdbuild_list" + its T; per iteration parity the list
    // of those graphs and its length
    int *kd_flag, *kill_T, *kill_H, *kd_list, *kd_nflag;   // kill_H: hits of the arterial kill (sorted list in `hl`), -1 = not available
    unsigned char* veto;
    long long* seqhash;
    long long* set_hash;
    int* set_key;
    int* err;
    int* trace;     // [g][iter][4]
    long long* dbg;       // [g][8] k_commit replay breakdown (cycles: -, walks, rechecks; counts: entries, events, walk steps, re-evaluations, records)
    long long* counters;  // [g][8] byte-accounting counters (sum_A, sum_M, sum_P, sum_S, ...)
};

}  // namespace octa
