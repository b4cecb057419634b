// vecops.h -- host interface of the fused O(n) kernels (vecops.cu).  All kernels are single-CTA
// (n <= a few thousand: latency-bound) and write their scalar results to a device `Scal`
// and to its pinned, host-mapped mirror.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "rowgeom.h"

namespace bnl {

struct VecCtx {
    int n, ld, m_lin, p;
    // replicated n-vectors (length ld, zero padded)
    double *x, *g, *s, *d, *hv, *r, *v, *pdir, *w, *gm, *xn, *xlow, *xupp, *wl, *wu, *t1, *t2;
    unsigned char *fix, *at;
    Scal *sd, *sh;  // device scalars, host-mapped mirror
    double atol_active, atol_negcurve, atol_boundary, kappa2;
    // nonlinear-constraint Jacobian (p x ld row-major) and mu*C, p-vectors
    double *C, *muC, *cv, *pvec;
    double mu;
};

// mask == true: m_lin == 0, projection is the exact mask v = fix ? 0 : r and is fused in.
// mask == false: the caller has already put P(.) in the indicated buffer with the general projection.
void vk_active_reset(const VecCtx& c, const double* xa, const double* sa /*nullable: use xa+sa*/, cudaStream_t st);
void vk_cauchy_init(const VecCtx& c, bool mask, cudaStream_t st);                 // s = 0; d = P(-g) [mask]
void vk_cauchy_eval(const VecCtx& c, double delta, cudaStream_t st);              // phi_p, phi_pp, theta, ind
void vk_cauchy_advance(const VecCtx& c, bool mask, int breakpoint, cudaStream_t st);
void vk_gminor_nrg(const VecCtx& c, bool mask, cudaStream_t st);                  // gm = hv + g; nrg_g, nrg_gm [mask]
void vk_norm_to(const VecCtx& c, const double* v, int which, cudaStream_t st);    // which: 0 nrg_g, 1 nrg_gm, 2 pix, 3 norm_g, 4 norm_s
void vk_cg_init(const VecCtx& c, bool mask, double delta, bool bounds_given, cudaStream_t st);  // bounds_given: keep c.wl / c.wu
void vk_cg_step(const VecCtx& c, bool mask, int phase, cudaStream_t st);          // phase 0: all (mask) / a ; 1: b (general)
void vk_minor_finish(const VecCtx& c, cudaStream_t st);                           // linesearch + w *= alpha + s += w
void vk_minor_post(const VecCtx& c, bool mask, double delta, cudaStream_t st);    // gm = hv+g; active_bounds; add_active; nrg
void vk_dot_gs(const VecCtx& c, cudaStream_t st);                                 // gs = g.s
void vk_trial_point(const VecCtx& c, cudaStream_t st);                            // xn = x + s; norm_s
void vk_pix(const VecCtx& c, bool mask, cudaStream_t st);                         // pix = ||P(-g)|| [mask], norm_g
void vk_hess_c(const VecCtx& c, const double* v, double* hv, bool add_to_hv, cudaStream_t st);  // cv=C v; Cv_sumsq; hv += C'(muC v)
void vk_add_Ct(const VecCtx& c, const double* pv, double* gout, cudaStream_t st);  // gout += C' pv
void vk_scale_C(const VecCtx& c, cudaStream_t st);                                 // muC = mu * C
// per-chunk partials of dot(r,r) (rowgeom.h): partial[ng*G]; the caller finishes with group_reduce / group_sum
void vk_sumsq_chunks(const double* r, const RowGeom& geo, double* partial, cudaStream_t st);
void vk_transpose_in(const double* src_colmajor, long long rows, int cols, long long lds, double* dst_rowmajor,
                     int ldd, cudaStream_t st);
void vk_pack_fix(const unsigned char* fix, int n, unsigned long long* words, cudaStream_t st);
void vk_unpack_fix(const unsigned long long* words, int n, unsigned char* fix, Scal* sd, Scal* sh, cudaStream_t st);
void vk_set_flags(const VecCtx& c, const long long* idx, int count, cudaStream_t st);  // fix[idx] = 1, recount
void vk_list_flags(const unsigned char* flags, int n, long long* idx_out, int* count_out, cudaStream_t st);
void vk_mask_project(const VecCtx& c, const double* src, double* dst, cudaStream_t st);  // dst = fix ? 0 : src
// built-in nonlinear constraint  c(x) = x'x - rho2  (p = 1):  value -> sd->c0 ;  jac_nlcons(x) = 2x' -> C row 0
void vk_sphere_value(const VecCtx& c, const double* x, double rho2, cudaStream_t st);
void vk_sphere_jac(const VecCtx& c, const double* x, cudaStream_t st);
void vk_publish(Scal* sd, Scal* sh, cudaStream_t st);
void vk_active_flags(const VecCtx& c, const double* x, const double* s, double delta, cudaStream_t st);  // at[] only

}  // namespace bnl
