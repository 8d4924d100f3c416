// Host build of the product's secular_core.h so the numerics can be unit-tested without a GPU.
#include "../../symmetric_eigenvalue_b200/csrc/secular_core.h"
extern "C" int secular_solve_all(int k, const double* d, const double* w, double rho,
                                 int* origin, double* tau, int* iters) {
    double sumw = 0; for (int j = 0; j < k; ++j) sumw += w[j];
    cuppen::SerialLanes L;
    int maxit = 0;
    for (int i = 0; i < k; ++i) {
        cuppen::SecularRoot r = cuppen::secular_solve(L, k, d, w, rho, sumw, i);
        origin[i] = r.origin; tau[i] = r.tau; iters[i] = r.iters;
        if (r.iters > maxit) maxit = r.iters;
    }
    return maxit;
}
