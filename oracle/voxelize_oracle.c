/* CPU ORACLE (test infrastructure, NOT part of the product): plain-C restatement of the reference's
 * capsule voxelizer.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  Pinned against the real reference: tests/golden/vox_*.npz were
 * produced by /root/reference/vessel_graph_generation/tree2img.py:voxelize_forest via
 * oracle/make_golden.py, and tests/test_oracle_voxelize.py checks this file against them bit for bit.
 *
 * Follows, line by line:
 *   tree2img.py:206-216   image_dim / pos_correction / zero volume
 *   tree2img.py:225-249   per-edge radius filter, scaling, ignore_z
 *   tree2img.py:151-172   getCrossSlice(mode='cuboid') bounding box
 *   tree2img.py:256-278   perpendicular pass (0<t<1) and end-cap pass, img = max(img, .)
 *   tree2img.py:279-280   (255*clip(img,0,1)).astype(uint16)
 * The dropout / blackdict branch (:220-224,:238-240) is host logic on the Python side of both the
 * oracle and the product (it consumes Python's global RNG) and is exercised in the tests there.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libvox_oracle.so oracle/voxelize_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int vox_oracle_out_dims(const int dims[3], int out[3]) {
    int S = dims[0];
    if (dims[1] > S) S = dims[1];
    if (dims[2] > S) S = dims[2];
    const double MAX_RADIUS = 0.015;
    int min_dim = (int)ceil((1.0 / 76) * (double)S + 2 * MAX_RADIUS * (double)S); /* :209 */
    for (int a = 0; a < 3; ++a) out[a] = dims[a] > min_dim ? dims[a] : min_dim;   /* :210 */
    return 0;
}

/* edges7: E x 7 doubles (node1 xyz, node2 xyz, radius).  out: uint16 [D0][D1][D2]. */
int vox_oracle(const double* edges7, long n_edges, const int dims[3], double min_radius, double max_radius,
               int ignore_z, uint16_t* out) {
    int D[3];
    vox_oracle_out_dims(dims, D);
    int S = dims[0];
    if (dims[1] > S) S = dims[1];
    if (dims[2] > S) S = dims[2];
    const double sf = (double)S;
    double c[3];
    for (int a = 0; a < 3; ++a) c[a] = (double)(D[a] - dims[a]) / 2; /* :211 */
    const size_t nvox = (size_t)D[0] * D[1] * D[2];
    double* img = (double*)calloc(nvox, sizeof(double)); /* :216 */
    if (!img) return -1;
    const double voxel_diag = sqrt(3.0); /* :214 */
    for (long k = 0; k < n_edges; ++k) {
        const double* e = edges7 + 7 * k;
        double radius = e[6];
        if (radius < min_radius || radius > max_radius) continue; /* :227 */
        radius *= sf;                                             /* :243 */
        double cur[3], prox[3];
        for (int a = 0; a < 3; ++a) {
            cur[a] = e[a] * sf + c[a];      /* :244 */
            prox[a] = e[3 + a] * sf + c[a]; /* :245 */
        }
        if (ignore_z) { cur[2] = (double)(D[2] / 2); prox[2] = (double)(D[2] / 2); } /* :247-249 */
        /* getCrossSlice, :152-166 */
        const double voxel_offset = radius * sqrt(2.0);
        long lo[3], hi[3];
        int empty = 0;
        for (int a = 0; a < 3; ++a) {
            double s = cur[a], t = prox[a];
            if (s > t) { double tmp = s; s = t; t = tmp; }
            double l = floor(s - voxel_offset), h = ceil(t + voxel_offset + 1);
            lo[a] = l < 0 ? 0 : (long)l;
            hi[a] = h > (double)D[a] ? D[a] : (long)h;
            if (hi[a] <= lo[a]) empty = 1;
        }
        if (empty) continue; /* :254 */
        const double seg[3] = {cur[0] - prox[0], cur[1] - prox[1], cur[2] - prox[2]}; /* :259 */
        const double ss = (seg[0] * seg[0] + seg[1] * seg[1]) + seg[2] * seg[2];
        const double rr = radius - voxel_diag / 2;
        for (long x = lo[0]; x < hi[0]; ++x)
            for (long y = lo[1]; y < hi[1]; ++y)
                for (long z = lo[2]; z < hi[2]; ++z) {
                    const double v[3] = {(double)x + .5, (double)y + .5, (double)z + .5}; /* :256 */
                    const double u[3] = {v[0] - prox[0], v[1] - prox[1], v[2] - prox[2]}; /* :260 */
                    const double t = ((u[0] * seg[0] + u[1] * seg[1]) + u[2] * seg[2]) / ss; /* :261 */
                    double* px = &img[((size_t)x * D[1] + y) * D[2] + z];
                    if (t > 0 && t < 1) { /* :262 */
                        double d0 = v[0] - (prox[0] + t * seg[0]), d1 = v[1] - (prox[1] + t * seg[1]),
                               d2 = v[2] - (prox[2] + t * seg[2]);                 /* :265-266 */
                        double dist = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
                        double contrib = 1 - ((dist - rr) / voxel_diag);           /* :269 */
                        if (contrib > *px) *px = contrib;                          /* :271 */
                    }
                    const double q[3] = {v[0] - cur[0], v[1] - cur[1], v[2] - cur[2]};
                    double dc = sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
                    double dp = sqrt((u[0] * u[0] + u[1] * u[1]) + u[2] * u[2]);
                    double dist = dc < dp ? dc : dp;                               /* :273-276 */
                    double contrib = 1 - ((dist - rr) / voxel_diag);
                    if (contrib > *px) *px = contrib;                              /* :278 */
                }
    }
    for (size_t i = 0; i < nvox; ++i) { /* :279-280 */
        double v = img[i];
        v = v < 0 ? 0 : (v > 1 ? 1 : v);
        out[i] = (uint16_t)(255 * v);
    }
    free(img);
    return 0;
}
