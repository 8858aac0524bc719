// Host-callable hooks that expose host/device-shared building blocks of the product to the CPU test
// suite (no GPU needed): the dgeev-faithful 3x3 eigen solver, CPython hashing, ...
#include "octa_common.h"
#include "octa_eig3.h"
#include "octa_grow_math.cuh"

extern "C" int octa_test_eig3(const double* cov9, double* w3, double* v9) {
    return octa::eig3::dgeev3_sym(cov9, w3, v9);
}

extern "C" int64_t octa_test_hash_tuple3(const double* p) { return octa::py_hash_tuple3(p[0], p[1], p[2]); }

extern "C" int octa_test_eig3_debug(const double* cov9, double* w3, double* v9, double* dbg36) {
    return octa::eig3::dgeev3_sym(cov9, w3, v9, dbg36);
}

extern "C" int octa_test_principal_axis(const double* cov9, double* dl3) { return octa::eig3::principal_axis(cov9, dl3); }

#include "octa_kdorder.h"
extern "C" void octa_test_kd_indices(const double* x, const double* y, const double* z, int n, int* idx_out) {
    octa::kd::build_indices_seq(x, y, z, n, idx_out);
}
