// matvec.h -- host interface of the streaming Jacobian kernels (matvec.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace bnl {

enum { MODE_JTJV = 0, MODE_JV = 1, MODE_JTW = 2 };

struct MvArgs {
    const double* J;   // row-major M x ld
    long long M;       // local rows
    int ld, R, NS, TG;
    const double* v;   // length ld (zero padded)      [JTJV, JV]
    const double* w;   // length M                     [JTW]
    double* t_out;     // length M or null             [JV]
    double* partial;   // [grid][pstride]
    long long pstride;
};

struct MvPlan {
    int ld, TG, KCH, RB, R, NS, grid;
    long long pstride;
    size_t smem_bytes;
    bool supported;
    bool warp_team;
};

MvPlan mv_make_plan(long long M, int n, int sm_count, size_t smem_optin_bytes);

// Launches the streaming kernel + the fixed-order partial reduction on `stream`.
// out: length ld+1 doubles; out[0..ld) = J'(Jv) or J'w, out[ld] = sum (Jv)_i^2 (JTJV, JV modes).
// p2p != nullptr: the partial reduction also pushes to every peer's mailbox and `out` holds the ALL-REDUCED result.
struct P2PArgs;
cudaError_t mv_launch(int mode, const MvPlan& p, const double* J, long long M, const double* v, const double* w,
                      double* t_out, double* partial, double* out, cudaStream_t stream, const P2PArgs* p2p = nullptr,
                      unsigned long long epoch = 0);

}  // namespace bnl
