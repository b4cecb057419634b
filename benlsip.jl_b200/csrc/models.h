// models.h -- built-in device-side residual / Jacobian generators (K11; definitions: oracle/models.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rowgeom.h"

namespace bnl {

struct ModelArgs {
    int model_id;
    long long M, M_total, row0;
    RowGeom geo;       // row-chunk geometry of the reductions (residual sum of squares)
    int n, ld;
    uint32_t seed;
    double noise;
    const double* cs;  // GLM column scale (device, length ld, zero padded)
};

// y = model(x_true) + noise  (once, at bind time)
cudaError_t model_setup_y(const ModelArgs& a, const double* x_true, double* y, cudaStream_t st);
// r = model(x) - y ; partial[gi*G + b] = sum r_i^2 over the row chunk (gi, b) (rowgeom.h)
cudaError_t model_residual(const ModelArgs& a, const double* x, const double* y, double* r, double* partial, cudaStream_t st);
// J (row-major M x ld, zero padded) = d model / d x at x
cudaError_t model_jacobian(const ModelArgs& a, const double* x, double* J, cudaStream_t st);

// GLM only, ld <= 1024: the same J, plus the per-(chunk, warp) partials of J'r in the layout of the streaming kernels' partials
// (KCH, RB = the streaming plan's): finishing them with group_reduce / group_sum gives g = J'r bit-identical to a J'w pass.
cudaError_t model_jacobian_jtr(const ModelArgs& a, const double* x, const double* r, double* J, double* partial, long long pstride,
                               int KCH, int RB, cudaStream_t st);

}  // namespace bnl
