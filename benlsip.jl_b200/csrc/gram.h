// gram.h -- FP64 tensor-core (DMMA) Gram kernel  G = J'J  (K12; not in the reference).
#pragma once
#include <cuda_runtime.h>

namespace bnl {
// J row-major M x ld.  G: ld x ld (symmetric, full square written), summed over local rows.
// workspace: nsplit * ld * ld doubles.  Returns the number of split-K slices used via *nsplit_out.
cudaError_t gram_launch(const double* J, long long M, int ld, double* G, double* workspace, int nsplit,
                        cudaStream_t st);
int gram_pick_split(long long M, int ld, int sm_count);
}  // namespace bnl
