// matvec.h -- host interface of the streaming Jacobian kernels (matvec.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "p2p.h"
#include "rowgeom.h"

namespace bnl {

enum { MODE_JTJV = 0, MODE_JV = 1, MODE_JTW = 2 };

struct MvArgs {
    const double* J;   // row-major M_loc x ld
    RowGeom geo;       // chunk geometry: CTA b owns the chunks (g, b) of the local groups
    int ld, R, NS, TG;
    const double* v;   // length ld (zero padded)      [JTJV, JV]
    const double* w;   // length M_loc                 [JTW]
    double* t_out;     // length M_loc or null         [JV, JTJV]: t = J v
    double* partial;   // [ng][G][T][pstride]
    long long pstride;
};

struct MvPlan {
    int ld, TG, T, KCH, RB, R, NS;
    long long pstride;
    size_t smem_bytes;
    bool supported;
    bool warp_team;
};

MvPlan mv_make_plan(int n, size_t smem_optin_bytes);
inline size_t mv_partial_doubles(const MvPlan& p, const RowGeom& geo) { return (size_t)geo.ng * geo.G * p.T * p.pstride; }

// Launches the streaming kernel (grid = geo.G CTAs) on `stream`.  It leaves one partial per (chunk, team) in
// partial[ng][G][T][pstride]: [0..ld) = column sums of J'(Jv) / J'w, [ld] = sum (Jv)_i^2 (JTJV, JV).  The caller finishes with
// the fixed reduction tree (group_reduce + group_sum, p2p.h): columns [0, ld] for JTJV, [ld, ld] for JV, [0, ld) for JTW.
cudaError_t mv_launch(int mode, const MvPlan& p, const RowGeom& geo, const double* J, const double* v, const double* w,
                      double* t_out, double* partial, cudaStream_t stream);

}  // namespace bnl
