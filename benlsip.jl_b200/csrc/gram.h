// gram.h -- FP64 tensor-core (DMMA) Gram kernel  G = J'J  and Gram-apply (K12; not in the reference).
#pragma once
#include <cuda_runtime.h>

namespace bnl {
// J row-major M x ld.  G: ld x ld row-major (symmetric, full square written), summed over local rows.
// workspace: nsplit * ld * ld doubles.
cudaError_t gram_launch(const double* J, long long M, int ld, double* G, double* workspace, int nsplit, cudaStream_t st);
int gram_pick_split(long long M, int ld, int sm_count);
// y[0..n) = G v ; y[ld] = v'Gv
cudaError_t gram_apply(const double* G, int n, int ld, const double* v, double* y, cudaStream_t st);
cudaError_t gram_gemv(const double* G, int n, int ld, const double* v, double* y, cudaStream_t st);  // y[0..n) = G v only
double gram_flops(long long M, int ld);
}  // namespace bnl
