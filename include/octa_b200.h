/* octa_b200.h -- C ABI of the B200-native vessel-graph hot path.
 *
 * The reference (aiforvision/OCTA-autosegmentation @ 9cdc313) is pure Python and has no FFI of its
 * own; the seams this library sits behind are the Python call sites listed per entry point below
 * (file:line relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add at each of them.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in any signature;
 *   - `*_dev` entry points take DEVICE pointers and a `void* stream` (a cudaStream_t, 0 = legacy
 *     default stream) and are asynchronous w.r.t. the host;
 *   - `*_host` entry points take HOST pointers, do their own H2D / D2H and return when the result
 *     is in the caller's buffer;
 *   - every function returns 0 on success, a negative OCTA_E_* code otherwise; the message of the
 *     last error on the calling thread is available through octa_last_error();
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point returns
 *     OCTA_E_CUDA.
 */
#ifndef OCTA_B200_H
#define OCTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCTA_ABI_VERSION 1

#define OCTA_OK 0
#define OCTA_E_ARG (-1)     /* bad argument */
#define OCTA_E_CUDA (-2)    /* CUDA runtime / launch failure, or no device */
#define OCTA_E_NOMEM (-3)   /* workspace too small / allocation failure */
#define OCTA_E_STATE (-4)   /* simulation overflowed a fixed-capacity buffer */

int octa_abi_version(void);
const char* octa_last_error(void);
/* number of kernel launches issued by this library on the calling process so far */
uint64_t octa_launch_count(void);
/* CUDA device count visible to the library (0 when there is none; never fails) */
int octa_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Voxelizer -- replaces vessel_graph_generation/tree2img.py:176-280 `voxelize_forest` (+ its helper
 * getCrossSlice(mode='cuboid') :151-172), called from generate_vessel_graph.py:69-72 and
 * visualize_vessel_graphs.py:79.
 *
 * An edge is 7 float64: node1.xyz, node2.xyz, radius (unit-cube coordinates), row-major E x 7.
 * The radius filter of tree2img.py:226-228 is applied here; the sequential dropout / blackdict logic
 * (:220-224,:238-240) consumes the caller's Python RNG and stays on the host side of the boundary.
 * ---------------------------------------------------------------------------------------------- */
typedef struct OctaVoxOpts {
    double min_radius; /* tree2img.py:179  default 0 */
    double max_radius; /* tree2img.py:180  default 1 */
    int ignore_z;      /* tree2img.py:247-249 */
    int reserved;
} OctaVoxOpts;

/* image_dim of tree2img.py:206-210: out_dims[i] = max(ceil(S/76 + 0.03*S), dims[i]), S = max(dims). */
int octa_voxelize_out_dims(const int dims[3], int out_dims[3]);

/* Bytes of device scratch octa_voxelize_batch_dev needs for n_graphs graphs / n_edges edges in total. */
size_t octa_voxelize_workspace_bytes(int n_graphs, int64_t n_edges, const int dims[3]);

/* Batched device entry point.  Graph g owns edges [edge_offsets[g], edge_offsets[g+1]) of `edges7`
 * and the volume out + g * X*Y*Z' (uint16, C order [X][Y][Z'], values 0..255 exactly as
 * `(255*np.clip(img,0,1)).astype(np.uint16)`, tree2img.py:279-280).  `edge_offsets` is a HOST array
 * of n_graphs+1 entries.  All edges of one graph are max-combined into one volume, which equals the
 * reference's np.maximum(art_mat, ven_mat) (generate_vessel_graph.py:70-72). */
int octa_voxelize_batch_dev(const double* edges7_dev, const int64_t* edge_offsets_host, int n_graphs,
                            const int dims[3], const OctaVoxOpts* opts, uint16_t* out_dev,
                            void* workspace_dev, size_t workspace_bytes, void* stream);

/* Single-graph host entry point (host buffers in, host buffer out; copies are inside). */
int octa_voxelize_host(const double* edges7, int64_t n_edges, const int dims[3], const OctaVoxOpts* opts,
                       uint16_t* out);

/* ------------------------------------------------------------------------------------------------
 * Test hooks (host-only code paths of host/device-shared building blocks; used by the CPU tests).
 * ---------------------------------------------------------------------------------------------- */
/* 3x3 symmetric eigenproblem with LAPACK dgeev's ordering and sign (greenhouse.py:229). cov9/v9 row-major;
 * column k of v9 is the unit eigenvector of w3[k].  Returns 0 ok, 1 complex/equal pair left, 2 no convergence. */
int octa_test_eig3(const double* cov9, double* w3, double* v9);
int octa_test_eig3_debug(const double* cov9, double* w3, double* v9, double* dbg48);
/* d_l of greenhouse.py:230-233: real part of the eigenvector of argmax(w); 0 ok, 2 no convergence, 3 complex principal pair */
int octa_test_principal_axis(const double* cov9, double* dl3);
/* CPython hash((np.float64 x, y, z)) (greenhouse.py:100-111 set ordering) */
int64_t octa_test_hash_tuple3(const double* xyz);

#ifdef __cplusplus
}
#endif
#endif /* OCTA_B200_H */
