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

#define OCTA_ABI_VERSION 3

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
 * 2-D label rasterizer -- replaces vessel_graph_generation/tree2img.py:12-114 `rasterize_forest` (matplotlib Agg
 * LineCollection, round caps, anti-aliased), called from generate_vessel_graph.py:79-86,
 * visualize_vessel_graphs.py:95 and data/data_transforms.py:384.  Output: uint8 gray [H][W] per graph
 * (H = image_resolution[1], W = image_resolution[0]); pixel row = pos[ax0]*H, col = pos[ax1]*W with
 * ax = {0,1,2} minus mip_axis.  Uses OctaVoxOpts.min_radius / max_radius (tree2img.py:66-68); `ignore_z` is unused.
 * All pixel arithmetic is the integer scanline arithmetic of Agg 2.4 (cover / area cells in 24.8 fixed point, 8-bit coverage,
 * 8-bit "over" blending in list order): the output equals oracle/agg_oracle.c pixel for pixel, and through PIL's
 * convert("1") the 500 label PNGs the reference ships bit for bit.
 * octa_raster2d_batch_layers_dev: the first layer_split[g] edges of graph g and the remaining ones are rasterized on separate
 * canvases and combined with max -- generate_vessel_graph.py:80-85 (arterial / venous forest, np.maximum).
 * Strokes wider than ~800 px are not supported (octa_raster2d_host reports them; the batch call sets an error flag in
 * the workspace).
 * ---------------------------------------------------------------------------------------------- */
size_t octa_raster2d_workspace_bytes(int n_graphs, int64_t n_edges, int H, int W);
int octa_raster2d_batch_dev(const double* edges7_dev, const int64_t* edge_offsets_host, int n_graphs, int H, int W,
                            int mip_axis, const OctaVoxOpts* opts, uint8_t* out_dev, void* workspace_dev,
                            size_t workspace_bytes, void* stream);
int octa_raster2d_batch_layers_dev(const double* edges7_dev, const int64_t* edge_offsets_host, const int64_t* layer_split_host,
                                   int n_graphs, int H, int W, int mip_axis, const OctaVoxOpts* opts, uint8_t* out_dev,
                                   void* workspace_dev, size_t workspace_bytes, void* stream);
int octa_raster2d_host(const double* edges7, int64_t n_edges, int H, int W, int mip_axis, const OctaVoxOpts* opts,
                       uint8_t* out);

/* ------------------------------------------------------------------------------------------------
 * Growth -- replaces generate_vessel_graph.py:24-56 (`main` up to the edge lists):
 *   Greenhouse(config['Greenhouse'])                      greenhouse.py:17-51
 *   Forest(config['Forest'], ...) x2 (arterial, venous)   forest.py:15-181
 *   greenhouse.develop_forest()                           greenhouse.py:57-137
 *   edge lists in LevelOrderIter order, root excluded     generate_vessel_graph.py:45-56
 * for a BATCH of independent samples.  Sample i is seeded the way the oracle harness seeds the reference:
 * random.seed(seeds[i]); np.random.seed(seeds[i]) immediately before Greenhouse(...).
 * ---------------------------------------------------------------------------------------------- */
typedef struct OctaGrowMode {        /* one entry of config['Greenhouse']['modes'] (raw YAML values) */
    int32_t I, N;
    double eps_n, eps_s, eps_k, delta_art, delta_ven, gamma_art, gamma_ven, phi, omega, kappa, delta_sigma;
    int32_t reinit;      /* mode['name'] != modes[0]['name']   greenhouse.py:84 */
    int32_t first_mode;  /* mode == modes[0]                   greenhouse.py:95 */
} OctaGrowMode;

typedef struct OctaGrowConfig {
    double d, r, faz_radius_bound[2], rotation_radius, faz_center[2], nerve_center[2], nerve_radius, param_scale;
    double size[3];                  /* SimulationSpace no_voxel_x/y/z */
    int32_t n_modes;
    OctaGrowMode modes[8];
    int32_t forest_type;             /* 0 = 'stumps', 1 = 'nerve' */
    int32_t n_trees;
    int32_t n_walls;
    int32_t walls[6];                /* enabled source walls in config order: 0=x0 1=x1 2=y0 3=y1 */
    int32_t cap_nodes, cap_sinks;    /* per-graph capacities; 0 = automatic */
    /* SimulationSpace.oxygen_sample_geometry_path (simulation_space.py:26-34,69-76,95-96): the loaded .npy as a C-order
     * 0/1 byte mask [geom_dims[0]][geom_dims[1]][geom_dims[2]], or NULL.  With a mask, `size` is ignored (the space is
     * geom_dims / max(geom_dims)) and the source walls z0 / z1 become usable (forest.py:152-176).  Any 3-D mask: every dimension
     * 1..65535 (voxel indices pass through uint16 in the reference, simulation_space.py:108), at most 2^26 voxels.  Copied at create time. */
    const unsigned char* geometry;
    int32_t geom_dims[3];
} OctaGrowConfig;

typedef struct OctaGrowStats {
    int64_t n_art_nodes, n_ven_nodes, n_oxy_left, n_co2_left, py_draws;
    int64_t sum_A, sum_M, sum_P, sum_S;   /* byte-accounting counters of SURVEY.md 8(d) */
    int64_t commit_cycles[4];        /* SM cycles of k_commit's prologue / sequential replay / refresh / active-list phases */
    int64_t replay_detail[8];        /* replay breakdown: cycles in tag scans, walks, rechecks; counts of entries, events, walk steps, tags, decision records */
    int32_t err;                     /* 0 ok; 1 node capacity, 2 sink capacity, 3 rng buffer, 5 set table, 6 geometry mask index, 1x eig */
    int32_t n_iters;
} OctaGrowStats;

/* Host-buffer entry point.  edges7_out: n_graphs * cap_edges_per_graph * 7 doubles; graph g's rows start at
 * edges7_out + g*cap_edges_per_graph*7, arterial rows first (n_art_edges[g]) then venous (n_ven_edges[g]).
 * trace (optional): n_graphs * 4096 * 4 int32, per iteration (art nodes, O2 sinks, venous nodes, CO2 sources).
 * device_ms (optional): device time of the growth loop measured with CUDA events. */
int octa_grow_batch_host(const OctaGrowConfig* cfg, const uint64_t* seeds, int n_graphs, double* edges7_out,
                         int64_t cap_edges_per_graph, int64_t* n_art_edges, int64_t* n_ven_edges,
                         OctaGrowStats* stats, int32_t* trace, double* device_ms);

/* Persistent variant: device state for up to max_graphs samples of one configuration is allocated once
 * (octa_grow_create) and reused by every octa_grow_run (same arguments as octa_grow_batch_host, n_graphs <= max_graphs);
 * the context belongs to the CUDA device that was current at creation and to one host thread at a time. */
int octa_grow_create(const OctaGrowConfig* cfg, int max_graphs, void** handle);
int octa_grow_run(void* handle, const uint64_t* seeds, int n_graphs, double* edges7_out, int64_t cap_edges_per_graph,
                  int64_t* n_art_edges, int64_t* n_ven_edges, OctaGrowStats* stats, int32_t* trace, double* device_ms);
/* Same, but the rows of all graphs are packed back to back into edges7_out (capacity cap_total_edges rows, e.g. a pinned
 * host buffer that is then copied to the device in one piece); edge_offsets[0..n_graphs] receives the row offsets. */
int octa_grow_run_packed(void* handle, const uint64_t* seeds, int n_graphs, double* edges7_out, int64_t cap_total_edges,
                         int64_t* edge_offsets, int64_t* n_art_edges, int64_t* n_ven_edges, OctaGrowStats* stats,
                         int32_t* trace, double* device_ms);
/* Final sink lists of graph `graph` (0 <= graph < n_graphs) of the context's LAST successful run -- what Greenhouse.save_stats
 * plots (greenhouse.py:401-418: oxy_mesh / co2_mesh .get_all_elements()).  which: 0 = oxygen sinks, 1 = CO2 sources; rows
 * (x, y, z) in list order.  *n receives the count, also when xyz_out is NULL or cap rows are too few (OCTA_E_NOMEM then).
 * The per-iteration node / sink counts of the same plot set are the `trace` argument of the run calls. */
int octa_grow_sinks(void* handle, int graph, int which, double* xyz_out, int64_t cap, int64_t* n);
void octa_grow_destroy(void* handle);

/* ------------------------------------------------------------------------------------------------
 * Graph CSV -- replaces the writer of generate_vessel_graph.py:59-66 (csv.writer rows of
 * [str(ndarray), str(ndarray), float]: numpy array2string + repr(float), "\r\n") and the row parser every
 * consumer uses (tree2img.py:73-76 / :233-236, visualize_vessel_graphs.py:72-75, data_transforms.py:377-381).
 * ---------------------------------------------------------------------------------------------- */
/* Formats header + n_edges rows into buf (cap bytes).  *len receives the byte count (also when buf is NULL or
 * too small, in which case OCTA_E_NOMEM is returned). */
int octa_format_csv(const double* edges7, int64_t n_edges, char* buf, size_t cap, size_t* len);
/* Parses CSV text; returns the number of data rows (>= 0) or a negative OCTA_E_* code.  Rows beyond cap_edges
 * are counted but not stored. */
int64_t octa_parse_csv(const char* text, size_t len, double* edges7_out, int64_t cap_edges);
/* The same files for a whole batch, written ON THE DEVICE (the host writer costs 3.7 ms of a core per docker-config graph, more
 * than the GPU needs to grow it): edges7_dev [E][7] with graph g owning the rows edge_offsets_host[g] .. [g+1].  text_dev
 * (text_cap bytes; octa_format_csv_text_cap gives the bound) receives the files back to back, text_offsets_dev[0..n_graphs]
 * their byte offsets, fallback_dev[g] != 0 marks a graph with a cell the device formatters do not cover (magnitudes outside the
 * unit cube's, a radius outside [1e-4, 1) or a power of two): its text region is undefined and the caller formats that graph with
 * octa_format_csv.  Everything is enqueued on `stream`; nothing is synchronised. */
size_t octa_format_csv_workspace_bytes(int n_graphs, int64_t n_edges);
size_t octa_format_csv_text_cap(int n_graphs, int64_t n_edges);
int octa_format_csv_batch_dev(const double* edges7_dev, const int64_t* edge_offsets_host, int n_graphs, char* text_dev,
                              size_t text_cap, int64_t* text_offsets_dev, int32_t* fallback_dev, void* workspace_dev,
                              size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GAN contrast adaptation (BASELINE config #5) -- replaces the generator forward behind
 * test.py:58-90 (GanSegModel.inference, models/gan_seg_model.py:65-79 -> ResnetGenerator.forward,
 * models/networks.py:350-443; resnetGenerator9 :502-503) and the pixel transforms of
 * docker/trained_models/GAN/config.yml:49-92 (ScaleIntensityd, background Rotate90d+Flipd,
 * AddRandomBackgroundNoised data/data_transforms.py:498-516, CastToTyped) plus the uint8 writer
 * utils/visualizer.py:338.  The 3x3 convolutions run on tcgen05 tensor cores (bf16 x bf16 -> fp32).
 * ---------------------------------------------------------------------------------------------- */
typedef struct OctaGanWeights {      /* HOST pointers, float32, the reference state-dict layouts [Cout][Cin][kh][kw] */
    const float* stem_w;             /* model.1.weight  [64][1][7][7]                                          */
    const float* conv_w[22];         /* model.4, model.8, model.12..20 .conv_block.1 / .5 (interleaved), model.22, model.26 */
    const float* head_w;             /* model.30.weight [1][64][7][7]                                          */
    float head_b;                    /* model.30.bias; the other biases cancel in InstanceNorm2d(affine=False)  */
    int reserved;
} OctaGanWeights;
/* Uploads the weights (3x3 kernels as bf16 [Cout][tap][Cin]) and allocates activations for up to max_images images of
 * H x W (multiples of 4) per call: about 100 MB per 304 x 304 image. */
int octa_gan_create(const OctaGanWeights* w, int max_images, int H, int W, void** handle);
/* x_dev float32 [n][H][W] in [0,1] -> y_dev float32 [n][H][W] (sigmoid output) and/or y_u8_dev = uint8(y * 255);
 * either output may be NULL.  Asynchronous on `stream`. */
int octa_gan_forward_dev(void* handle, const float* x_dev, int n_images, float* y_dev, uint8_t* y_u8_dev, void* stream);
void octa_gan_destroy(void* handle);
/* speckle_dev[i] = the float64 array of `np.random.seed(seeds[i]); np.random.uniform(0, 1, (H, W))` (data_transforms.py:510) */
int octa_gan_speckle_dev(const uint32_t* seeds_dev, int n_images, int H, int W, double* speckle_dev, void* stream);
/* x = float32(max(scale(raster), scale(background)^T * speckle)) per image (scale = ScaleIntensity to [0,1]; the
 * transpose is Rotate90d(k=1) followed by Flipd(0)).  background_dev and speckle_dev may both be NULL (no noise).
 * minmax_ws_dev: 4 ints per image of scratch. */
int octa_gan_input_dev(const uint8_t* raster_dev, const uint8_t* background_dev, const double* speckle_dev, int n_images, int H, int W,
                       int* minmax_ws_dev, float* x_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Test hooks (host-only code paths of host/device-shared building blocks; used by the CPU tests).
 * ---------------------------------------------------------------------------------------------- */
/* 3x3 symmetric eigenproblem with LAPACK dgeev's ordering and sign (greenhouse.py:229). cov9/v9 row-major;
 * column k of v9 is the unit eigenvector of w3[k].  Returns 0 ok, 1 complex/equal pair left, 2 no convergence. */
int octa_test_eig3(const double* cov9, double* w3, double* v9);
int octa_test_eig3_debug(const double* cov9, double* w3, double* v9, double* dbg48);
/* d_l of greenhouse.py:230-233: real part of the eigenvector of argmax(w); 0 ok, 2 no convergence, 3 complex principal pair */
int octa_test_principal_axis(const double* cov9, double* dl3);
/* cKDTree `tree.indices` permutation of n points given as SoA (element_mesh.py:97-101,136-137: ball-result order) */
void octa_test_kd_indices(const double* x, const double* y, const double* z, int n, int* idx_out);
/* the same permutation built by one CTA on the GPU (the block-parallel code path of the growth kernels) */
int octa_test_kd_indices_gpu(const double* x, const double* y, const double* z, int n, int* idx_out);
/* the shared-memory resident build k_kdbuild prefers (n <= ~17 k): rank_out[idx_out[i]] = i; -1 everywhere if it bailed out */
int octa_test_kd_ranks_gpu_smem(const double* x, const double* y, const double* z, int n, int* rank_out);
/* one 3x3 convolution (padding 1, zero or reflect) through the tcgen05 kernel of the GAN path, HOST tensors in torch layouts:
 * x [n][Cin][H][W], w [Cout][Cin][3][3] -> y [n][Cout][H][W] (inputs rounded to bf16, fp32 accumulation, bf16 result) */
int octa_test_gan_conv3_host(const float* x, const float* w, int n, int H, int W, int cin, int cout, int reflect, float* y);
/* CPython hash((np.float64 x, y, z)) (greenhouse.py:100-111 set ordering) */
int64_t octa_test_hash_tuple3(const double* xyz);

#ifdef __cplusplus
}
#endif
#endif /* OCTA_B200_H */
