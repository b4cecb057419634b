// dense.h -- small dense factor / solve kernels for the general projection (m_lin > 0):
// cholesky_aug_aat, update_chol!, left_mul, left_mul_tr, projection_nullspace!/projection_subspace!
// (src/polyhedral_constraints.jl:35-136).  Single-CTA kernels: the systems are (m_lin + q)^2, replicated.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace bnl {

struct DenseCtx {
    int n, ld, m;          // m = m_lin
    int cap;               // leading dimension / capacity of L (>= n)
    const double* A;       // m x ld row-major (zero padded)
    double* LA;            // m x m   column-major lower factor of A A'
    double* L;             // cap x cap column-major lower factor of A~ A~'  (dim m+q)
    double* G;             // m x cap column-major:  L_A \ A[:,fix]
    double* ywork;         // cap
    double* Lr;            // m x m column-major lower factor of A_free A_free'  (reduced-space projection)
    long long* fixidx;     // ascending indices of fixed variables (cap)
    int* q_dev;            // number of fixed variables
    Scal* sd;
    Scal* sh;
};

void dk_chol_aat(const DenseCtx& c, cudaStream_t st);                              // LA = cholesky(A*A').L  (basic_tralcnlss.jl:206)
void dk_rebuild(const DenseCtx& c, const unsigned char* fix, cudaStream_t st);     // update_chol!  :62-68
void dk_project(const DenseCtx& c, const double* r, double* v, bool negate, cudaStream_t st);  // projection! :158-170

// left_mul (A~ x = [A x; x[fix]], :86-98) and left_mul_tr (A~' y, :72-84) on their own (the reference's tests call them)
void dk_left_mul(const DenseCtx& c, const double* x, double* y, cudaStream_t st);      // y: m + q (needs dk_rebuild's fixidx)
void dk_left_mul_tr(const DenseCtx& c, const double* y, double* x, cudaStream_t st);   // x: n

// Reduced-space form of the same projection (default on the solve path).  A~A~' = [AA' A_F; A_F' I] and
// S = I - G'G is a rank-m downdate of the identity, so P(r) = r - A~'(A~A~')^{-1}A~ r is equivalently
//     v_F = 0,   v_free = r_free - A_free' (A_free A_free')^{-1} A_free r_free
// which needs only the m x m Cholesky factor of A_free A_free' (O(m^2 n) per active-set change instead of the
// reference's O(q^3) rebuild, src/polyhedral_constraints.jl:51 flags that cost).  Same mathematics, same failure
// condition (A_free rank deficient <=> PosDefException), different rounding.
void dk_rs_rebuild(const DenseCtx& c, const unsigned char* fix, cudaStream_t st);
// the variable sd->bp_ind just left the free set: rank-one downdate of the m x m factor, O(m^2)
bool dk_rs_downdate_fits(const DenseCtx& c);  // the m x m factor fits in shared memory (m <= 160)
void dk_rs_downdate(const DenseCtx& c, cudaStream_t st);
// one Cauchy breakpoint in one launch: s += theta d ; fix[ind] = 1 ; downdate ; d = P(-g)   (theta, ind from the device scalars)
void dk_rs_breakpoint(const DenseCtx& c, double* s, double* d, const double* g, unsigned char* fix, cudaStream_t st);
void dk_rs_project(const DenseCtx& c, const unsigned char* fix, const double* r, double* v, bool negate, cudaStream_t st);

}  // namespace bnl
