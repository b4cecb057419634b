// cauchy_loop.h -- host interface of the persistent breakpoint-loop kernel (cauchy_loop.cu).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "p2p.h"
#include "rowgeom.h"
#include "vecops.h"

namespace bnl {

constexpr int kCLMaxP = 8;  // nonlinear constraints the loop carries itself (C-part of the Hessian); above: literal search

// exit status of one launch (Scal::cl_status)
enum {
    CL_ADVANCE = 0,         // internal: keep going
    CL_DONE_NOSTEP = 1,     // phi' >= 0 (:620)
    CL_DONE_INTERIOR = 2,   // interior minimiser, step applied with the LITERAL -phi'/phi'' (:622-626)
    CL_DONE_EXHAUSTED = 3,  // count(fixvars) reached n - m (:615)
    CL_NEED_LITERAL = 4,    // evaluate Hd = H*d literally (hess_mul + vk_cauchy_eval) and re-enter with use_literal = 1
    CL_ERR_BOUNDS = 5,      // next_breakpoint found no breakpoint (ind = -1): BoundsError in the reference
    CL_TIMEOUT = 6,         // a peer did not answer
    CL_WANT_TRANSPOSE = 7   // the search is long: build the tile-transposed copy of J, then re-enter (nothing was advanced)
};

constexpr int kJtTile = 16;      // rows per tile of the transposed copy: one column's 16 rows = one 128-byte line
constexpr int kJtTrigger = 128;  // breakpoints of one search after which the copy is built (it costs ~2 passes = ~130 breakpoints)

struct CauchyLoopArgs {
    VecCtx c;
    RowGeom geo;
    const double* J;
    const double* Jt;   // tile-transposed copy of J (Jt[((i/16)*ld + j)*16 + i%16] = J[i][j]) or null
    int ld;
    double* t;          // M_loc: J d
    double* u;          // M_loc: J s_c
    double* partial2;   // [ng][G][2]
    P2PArgs p2p;
    int multi;          // nranks > 1: exchange the group sums through the LL mailbox
    unsigned long long ll_epoch0;
    double delta;
    double guard;       // relative width of the rounding band
    int first;          // 1: fresh search (u = 0, t = J d just computed)
    int use_literal;    // 1: the first decision uses sd->phi_p / sd->phi_pp (literal evaluation)
    int nmm;            // n - m_lin
    int q0;             // breakpoints this search has already advanced (earlier launches)
    int want_jt;        // 1: exit with CL_WANT_TRANSPOSE when the search reaches kJtTrigger breakpoints (no copy attempted yet)
    unsigned int* arrive;  // sync block (cauchy_loop_sync_bytes()): arrive counter at +0, broadcast record at +64
    void* bcast;
};

size_t cauchy_loop_sync_bytes();
// Jt (size ceil16(M) * ld doubles) = tile-transposed copy of the row-major M x ld matrix J
cudaError_t transpose16_launch(const double* J, long long M, int ld, double* Jt, cudaStream_t st);
cudaError_t cauchy_loop_launch(const CauchyLoopArgs& a, int sm_count, cudaStream_t st);

}  // namespace bnl
