// Host-callable hooks that expose host/device-shared building blocks of the product to the CPU test
// suite (no GPU needed): the dgeev-faithful 3x3 eigen solver, CPython hashing, ...
#include "octa_common.h"
#include "octa_eig3.h"
#include "octa_grow_math.cuh"

extern "C" int octa_test_eig3(const double* cov9, double* w3, double* v9) {
    return octa::eig3::dgeev3_sym(cov9, w3, v9);
}

extern "C" int64_t octa_test_hash_tuple3(const double* p) { return octa::py_hash_tuple3(p[0], p[1], p[2]); }

#include <vector>
#include "octa_pyset.cuh"
// The set emulation + order-sensitivity test k_kill runs (host build of the same code).  xyz: T sink tuples in insertion order,
// ball[q]: id of the ball (new node) that inserts tuple q (non-decreasing).  order_out: iteration order of the resulting set
// (indices into xyz); returns 1 if `detect` and the result was flagged as possibly depending on the order inside a ball
// (order_out is then not filled), 0 otherwise, < 0 on error.
extern "C" int octa_test_pyset(const double* xyz, const int* ball, int T, int detect, int* order_out, int* n_out) {
    static octa::KillShared ks;
    std::vector<long long> gth((size_t)2 * octa::SET_TBL), sh(T);
    std::vector<int> gtk((size_t)2 * octa::SET_TBL), seq(T);
    for (int q = 0; q < T; ++q) { seq[q] = q; sh[q] = octa::py_hash_tuple3(xyz[3 * q], xyz[3 * q + 1], xyz[3 * q + 2]); }
    for (int i = 0; i < 8; ++i) { ks.tk[0][i] = -1; ks.th[0][i] = 0; }
    octa::PySetDev ps;
    ps.sh = &ks; ps.gth = gth.data(); ps.gtk = gtk.data();
    ps.init();
    if (detect == 2 && T <= octa::MS_MAXT) {          // the multi-state test k_kill runs for T <= MS_MAXT
        std::vector<short> tabs((size_t)octa::MS_TABLES * octa::MS_TBL);
        int mask = 0;
        const int why = octa::pyset_run_multi(sh.data(), ball, T, tabs.data(), &mask);
        if (why) return why;
        int n = 0;
        for (int z = 0; z <= mask; ++z) if (tabs[z] >= 0) order_out[n++] = seq[tabs[z]];
        *n_out = n;
        return 0;
    }
    const bool flagged = detect ? octa::pyset_run<true>(ps, seq.data(), sh.data(), T, ball, gtk.data())
                                : octa::pyset_run<false>(ps, seq.data(), sh.data(), T, ball, gtk.data());
    if (ps.err) return -ps.err;
    if (flagged) return 1;
    int n = 0;
    const int* ck = ps.tabk(ps.cur);
    for (size_t z = 0; z <= ps.mask; ++z) if (ck[z] >= 0) order_out[n++] = ck[z];
    *n_out = n;
    return 0;
}

extern "C" int octa_test_eig3_debug(const double* cov9, double* w3, double* v9, double* dbg36) {
    return octa::eig3::dgeev3_sym(cov9, w3, v9, dbg36);
}

extern "C" int octa_test_principal_axis(const double* cov9, double* dl3) { return octa::eig3::principal_axis(cov9, dl3); }

#include "octa_kdorder.h"
extern "C" void octa_test_kd_indices(const double* x, const double* y, const double* z, int n, int* idx_out) {
    octa::kd::build_indices_seq(x, y, z, n, idx_out);
}

#include "octa_kdorder_par.cuh"
namespace {
__global__ void __launch_bounds__(1024) kd_test_kernel(const double* x, const double* y, const double* z, int n, int* idx,
                                                       int* posL, int* posR, int* nodes, int mode) {
    __shared__ int s_ws[octa::kdpar::WS_INTS];
    __shared__ double s_wd[octa::kdpar::WD_DOUBLES];
    extern __shared__ __align__(16) char s_dyn[];
    if (mode == 1) {   // shared-memory resident build; idx receives the RANK of every point (-1 everywhere on bail-out)
        if (!octa::kdsm::build_ranks_block(x, y, z, n, idx, s_dyn, nodes, nodes + 3 * (n / 8 + 4), s_ws, s_wd))
            for (int i = threadIdx.x; i < n; i += blockDim.x) idx[i] = -1;
        return;
    }
    octa::kdpar::build_indices_block(x, y, z, n, idx, posL, posR, nodes, nodes + n / 2 + 8, s_ws, s_wd);
}
}  // namespace

// GPU build of the same permutation by one CTA (the code paths k_kdbuild uses); host buffers in/out.
// mode 0: global-memory build -> indices; mode 1: shared-memory resident build -> RANKS (rank[indices[i]] = i).
static int kd_indices_gpu(const double* x, const double* y, const double* z, int n, int* out, int mode) {
    OCTA_ARG_CHECK(n >= 0 && out, "bad arguments");
    if (octa_device_count() <= 0) { octa::set_error("no CUDA device"); return OCTA_E_CUDA; }
    if (n == 0) return OCTA_OK;
    size_t smem = 0;
    if (mode == 1) {
        smem = (size_t)n * 12 + 64;
        OCTA_ARG_CHECK(smem <= 208 * 1024 && n < 65536, "too many points for the shared-memory build");
        OCTA_CUDA_CHECK(cudaFuncSetAttribute(kd_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    }
    double* d = nullptr;
    int* w = nullptr;
    OCTA_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * 3 * (size_t)n));
    OCTA_CUDA_CHECK(cudaMalloc(&w, sizeof(int) * (4 * (size_t)n + 64)));
    cudaMemcpy(d, x, 8 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(d + n, y, 8 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(d + 2 * (size_t)n, z, 8 * (size_t)n, cudaMemcpyHostToDevice);
    kd_test_kernel<<<1, 1024, smem>>>(d, d + n, d + 2 * (size_t)n, n, w, w + n, w + 2 * (size_t)n, w + 3 * (size_t)n, mode);
    octa::count_launch();
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce == cudaSuccess) ce = cudaMemcpy(out, w, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(w);
    if (ce != cudaSuccess) { octa::set_error("kd_test_kernel: %s", cudaGetErrorString(ce)); return OCTA_E_CUDA; }
    return OCTA_OK;
}

extern "C" int octa_test_kd_indices_gpu(const double* x, const double* y, const double* z, int n, int* idx_out) {
    return kd_indices_gpu(x, y, z, n, idx_out, 0);
}

extern "C" int octa_test_kd_ranks_gpu_smem(const double* x, const double* y, const double* z, int n, int* rank_out) {
    return kd_indices_gpu(x, y, z, n, rank_out, 1);
}

// The cell formatters the DEVICE CSV writer uses (host build of the same code): text into out (>= 128 bytes), returns the length
// or -1 when the formatter declines (generic host path).
#include "octa_csvfmt.cuh"
extern "C" int octa_test_csvfmt_array3(const double* v3, char* out) { return octa::csvfmt::array3(out, v3); }
extern "C" int octa_test_csvfmt_repr(double x, char* out) { return octa::csvfmt::repr_unit(out, x); }
// batch forms for the randomized tests: n values -> lengths (or -1) and texts at stride 128
extern "C" void octa_test_csvfmt_repr_many(const double* x, int64_t n, char* out, int* len) {
    for (int64_t i = 0; i < n; ++i) len[i] = octa::csvfmt::repr_unit(out + 128 * i, x[i]);
}
extern "C" void octa_test_csvfmt_array3_many(const double* v, int64_t n, char* out, int* len) {
    for (int64_t i = 0; i < n; ++i) len[i] = octa::csvfmt::array3(out + 128 * i, v + 3 * i);
}
