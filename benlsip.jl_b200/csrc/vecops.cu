// vecops.cu -- fused O(n) kernels of the step computation: bound projection / active-set masking fused
// into the Cauchy-step and CG-step updates, all dots / norms / arg-min scans in-kernel, fixed reduction trees.
// Each kernel cites the reference lines it restates (paths relative to the reference repo).
#include "vecops.h"

namespace bnl {
namespace {

constexpr int kVT = 1024;  // threads of the single CTA

__device__ __forceinline__ void publish(const Scal* sd, Scal* sh) {
    // copy the device scalars to the host-mapped mirror (tiny; visible to the host after stream sync)
    __syncthreads();
    const int nw = sizeof(Scal) / 8;
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(sd);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(sh);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) d[i] = s[i];
}

// active_bounds!  src/polyhedral_constraints.jl:203-215 (overwrites fixvars from x; also used on x+s, :452)
__global__ void k_active_reset(VecCtx c, const double* xa, const double* sa) {
    __shared__ int shi[32];
    int cnt = 0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        const double xi = sa ? (xa[i] + sa[i]) : xa[i];
        const bool f = (xi - c.xlow[i] <= c.atol_active) || (c.xupp[i] - xi <= c.atol_active);
        c.fix[i] = f ? 1 : 0;
        cnt += f ? 1 : 0;
    }
    cnt = block_sum_int(cnt, shi);
    if (threadIdx.x == 0) c.sd->nb_fix = cnt;
    publish(c.sd, c.sh);
}

// cauchy_step prologue, src/basic_tralcnlss.jl:587-592: s_c = 0; d = projection(lincons, -g)
template <bool MASK>
__global__ void k_cauchy_init(VecCtx c) {
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        c.s[i] = 0.0;
        if (MASK) c.d[i] = c.fix[i] ? 0.0 : -c.g[i];
    }
}

// :609-611 + next_breakpoint :536-562.  phi_p = dot(s_c,Hd) + dot(g,d); phi_pp = dot(d,Hd)
__global__ void k_cauchy_eval(VecCtx c, double delta) {
    __shared__ double shd[32];
    __shared__ long long shl[32];
    double a = 0.0, b = 0.0, e = 0.0;
    double th = INFINITY;
    long long ind = -1;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        const double di = c.d[i], hi = c.hv[i], si = c.s[i];
        a = fma(si, hi, a);
        b = fma(c.g[i], di, b);
        e = fma(di, hi, e);
        if (!c.fix[i]) {
            double tt = INFINITY;
            if (di < 0.0) {
                const double dl = fmax(c.xlow[i] - c.x[i], -delta);  // d_l :603
                tt = (dl - si) / di;
            } else if (di > 0.0) {
                const double du = fmin(c.xupp[i] - c.x[i], delta);  // d_u :602
                tt = (du - si) / di;
            }
            if (tt < th) {  // strict <, ascending i within a thread
                th = tt;
                ind = i;
            }
        }
    }
    a = block_sum(a, shd);
    b = block_sum(b, shd);
    e = block_sum(e, shd);
    block_argmin(th, ind, shd, shl);
    if (threadIdx.x == 0) {
        c.sd->phi_p = a + b;
        c.sd->phi_a = a;
        c.sd->phi_b = b;
        c.sd->phi_pp = e;
        c.sd->theta = th;
        c.sd->bp_ind = ind;
        c.sd->bp_dind = (ind >= 0) ? c.d[ind] : 0.0;
    }
    publish(c.sd, c.sh);
}

// :622-635.  breakpoint == 0: s_c += (-phi_p/phi_pp) d.   breakpoint == 1: s_c += theta d; add_active!(ind); d = P(-g)
template <bool MASK>
__global__ void k_cauchy_advance(VecCtx c, int breakpoint) {
    const double step = breakpoint ? c.sd->theta : (-c.sd->phi_p / c.sd->phi_pp);
    const long long ind = c.sd->bp_ind;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        c.s[i] = c.s[i] + step * c.d[i];
        if (breakpoint) {
            if (i == ind) c.fix[i] = 1;
            if (MASK) c.d[i] = (c.fix[i] || i == ind) ? 0.0 : -c.g[i];
        }
    }
    if (breakpoint && threadIdx.x == 0) {
        // fixvars[ind] was free (next_breakpoint only scans free variables) => count grows by one
        c.sd->nb_fix = c.sd->nb_fix + 1;
    }
    publish(c.sd, c.sh);
}

// g_minor = H*s + g (:412,:437) and the two reduced-gradient norms (:420-421, :869-875) for the mask projection
template <bool MASK>
__global__ void k_gminor_nrg(VecCtx c) {
    __shared__ double shd[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        const double gmi = c.hv[i] + c.g[i];
        c.gm[i] = gmi;
        if (MASK && !c.fix[i]) {
            a = fma(c.g[i], c.g[i], a);
            b = fma(gmi, gmi, b);
        }
    }
    if (MASK) {
        a = block_sum(a, shd);
        b = block_sum(b, shd);
        if (threadIdx.x == 0) {
            c.sd->nrg_g = sqrt(a);
            c.sd->nrg_gm = sqrt(b);
        }
    }
    publish(c.sd, c.sh);
}

__global__ void k_norm_to(VecCtx c, const double* v, int which) {
    __shared__ double shd[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) a = fma(v[i], v[i], a);
    a = block_sum(a, shd);
    if (threadIdx.x == 0) {
        const double nv = sqrt(a);
        switch (which) {
            case 0: c.sd->nrg_g = nv; break;
            case 1: c.sd->nrg_gm = nv; break;
            case 2: c.sd->pix = nv; break;
            case 3: c.sd->norm_g = nv; break;
            default: c.sd->norm_s = nv; break;
        }
    }
    publish(c.sd, c.sh);
}

// minor_iterate :660-665 (trap T1: finite w_l/w_u on the FIXED variables) + projected_cg prologue :702-718
// bounds_given != 0: projected_cg called with the caller's own w_l / w_u (:690-697) already in c.wl / c.wu
template <bool MASK>
__global__ void k_cg_init(VecCtx c, double delta, int bounds_given) {
    __shared__ double shd[32];
    double rtv = 0.0, vv = 0.0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        const bool f = c.fix[i] != 0;
        const double xm = c.x[i] + c.s[i];  // x_minor :660
        if (!bounds_given) {
            c.wu[i] = f ? fmin(c.xupp[i] - xm, delta) : INFINITY;
            c.wl[i] = f ? fmax(c.xlow[i] - xm, -delta) : -INFINITY;
        }
        c.w[i] = 0.0;
        const double ri = c.gm[i];
        c.r[i] = ri;
        const double vi = MASK ? (f ? 0.0 : ri) : c.v[i];
        if (MASK) c.v[i] = vi;
        c.pdir[i] = -vi;
        rtv = fma(ri, vi, rtv);
        vv = fma(vi, vi, vv);
    }
    rtv = block_sum(rtv, shd);
    vv = block_sum(vv, shd);
    if (threadIdx.x == 0) {
        c.sd->rtv = rtv;
        c.sd->tol_cg = c.kappa2 * sqrt(vv);  // :710
        c.sd->cg_neg_curv = 0;
        c.sd->cg_outside = 0;
        c.sd->cg_solved = 0;
        c.sd->cg_iter = 1;
    }
    publish(c.sd, c.sh);
}

// One projected_cg iteration after Hp = H*p is in hv (:723-749).
// PHASE 0 (mask): the whole iteration.  General projection: PHASE 0 stops after r += alpha*Hp (the caller then
// projects r -> v) and PHASE 1 finishes (:743-748).
template <bool MASK, int PHASE>
__global__ void k_cg_step(VecCtx c) {
    __shared__ double shd[32];
    if (PHASE == 0) {
        double pHp = 0.0, rtv = 0.0, gam = INFINITY;
        for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
            const double pi = c.pdir[i];
            pHp = fma(pi, c.hv[i], pHp);
            rtv = fma(c.r[i], c.v[i], rtv);
            if (pi <= -c.atol_boundary)
                gam = fmin(gam, (c.wl[i] - c.w[i]) / pi);  // factor_to_boundary :793-809
            else if (pi >= c.atol_boundary)
                gam = fmin(gam, (c.wu[i] - c.w[i]) / pi);
        }
        pHp = block_sum(pHp, shd);
        rtv = block_sum(rtv, shd);
        gam = block_min(gam, shd);
        double step = 0.0;
        int full = 0;
        if (pHp <= c.atol_negcurve) {  // :725
            if (fabs(pHp) > c.atol_negcurve) step = gam;  // :727-730 (dead for PSD H, trap T2)
            if (threadIdx.x == 0) c.sd->cg_neg_curv = 1;
        } else {
            const double alpha = rtv / pHp;  // :733
            if (alpha > gam) {               // :735-737
                step = gam;
                if (threadIdx.x == 0) c.sd->cg_outside = 1;
            } else {
                step = alpha;
                full = 1;
            }
            if (threadIdx.x == 0) {
                c.sd->alpha = alpha;
                c.sd->rtv = rtv;
            }
        }
        if (threadIdx.x == 0) {
            c.sd->pHp = pHp;
            c.sd->gamma = gam;
        }
        double rtv_next = 0.0;
        for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
            if (step != 0.0) c.w[i] = c.w[i] + step * c.pdir[i];  // :729,:737,:739
            if (full) {
                const double ri = c.r[i] + step * c.hv[i];  // :740
                c.r[i] = ri;
                if (MASK) {
                    const double vi = c.fix[i] ? 0.0 : ri;  // projection! :741
                    c.v[i] = vi;
                    rtv_next = fma(ri, vi, rtv_next);
                }
            }
        }
        if (MASK && full) {
            rtv_next = block_sum(rtv_next, shd);
            const double beta = rtv_next / rtv;  // :744
            for (int i = threadIdx.x; i < c.n; i += blockDim.x) c.pdir[i] = -c.v[i] + beta * c.pdir[i];  // :745
            if (threadIdx.x == 0) {
                c.sd->beta = beta;
                c.sd->rtv = rtv_next;
                c.sd->cg_solved = (fabs(rtv_next) < c.sd->tol_cg) ? 1 : 0;  // :747
                c.sd->cg_iter = c.sd->cg_iter + 1;
            }
        }
    } else {
        // general projection, second half: v = P(r) is in c.v
        const double rtv = c.sd->rtv;
        double rtv_next = 0.0;
        for (int i = threadIdx.x; i < c.n; i += blockDim.x) rtv_next = fma(c.r[i], c.v[i], rtv_next);
        rtv_next = block_sum(rtv_next, shd);
        const double beta = rtv_next / rtv;
        for (int i = threadIdx.x; i < c.n; i += blockDim.x) c.pdir[i] = -c.v[i] + beta * c.pdir[i];
        __syncthreads();
        if (threadIdx.x == 0) {
            c.sd->beta = beta;
            c.sd->rtv = rtv_next;
            c.sd->cg_solved = (fabs(rtv_next) < c.sd->tol_cg) ? 1 : 0;
            c.sd->cg_iter = c.sd->cg_iter + 1;
        }
    }
    publish(c.sd, c.sh);
}

// linesearch :766-791 (skipped on negative curvature :669) ; w .= alpha*w :671 ; s .+= w :436
// wHw = jv_sumsq (hv[ld], all-reduced) + mu * Cv_sumsq
__global__ void k_minor_finish(VecCtx c) {
    __shared__ double shd[32];
    const bool neg = c.sd->cg_neg_curv != 0;
    double alpha = 1.0;
    if (!neg) {
        double gw = 0.0, allowed = INFINITY;
        for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
            const double wi = c.w[i];
            gw = fma(c.gm[i], wi, gw);
            if (!c.fix[i]) {
                if (wi < 0.0)
                    allowed = fmin(allowed, c.wl[i] / wi);
                else if (wi > 0.0)
                    allowed = fmin(allowed, c.wu[i] / wi);
            }
        }
        gw = block_sum(gw, shd);
        allowed = block_min(allowed, shd);
        const double wHw = c.hv[c.ld] + c.mu * c.sd->Cv_sumsq;  // vthv :92-96
        const double alpha_opt = (wHw > 0.0) ? (-gw / wHw) : INFINITY;
        alpha = fmin(alpha_opt, allowed);
        if (threadIdx.x == 0) {
            c.sd->wHw = wHw;
            c.sd->gw = gw;
            c.sd->alpha_ls = alpha;
        }
    }
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        double wi = c.w[i];
        if (!neg) {
            wi = alpha * wi;
            c.w[i] = wi;
        }
        c.s[i] = c.s[i] + wi;
    }
    publish(c.sd, c.sh);
}

// :437-448: g_minor = H*s+g; active_indx = active_bounds(lincons,x,s,delta) (polyhedral_constraints.jl:219-237);
// if m + |active_indx| <= n: add_active! and (mask) the two reduced-gradient norms.
template <bool MASK>
__global__ void k_minor_post(VecCtx c, double delta) {
    __shared__ double shd[32];
    __shared__ int shi[32];
    int cnt = 0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        c.gm[i] = c.hv[i] + c.g[i];
        const double sl = fmax(c.xlow[i] - c.x[i], -delta);
        const double su = fmin(c.xupp[i] - c.x[i], delta);
        const double si = c.s[i];
        const bool at = (si - sl <= c.atol_active) || (su - si <= c.atol_active);
        c.at[i] = at ? 1 : 0;
        cnt += at ? 1 : 0;
    }
    cnt = block_sum_int(cnt, shi);
    const bool add = (c.m_lin + cnt <= c.n);  // :441
    int nfix = 0;
    double a = 0.0, b = 0.0;
    if (add) {
        for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
            const unsigned char f = (c.fix[i] | c.at[i]) ? 1 : 0;
            c.fix[i] = f;
            nfix += f;
            if (MASK && !f) {
                a = fma(c.g[i], c.g[i], a);
                const double gmi = c.gm[i];
                b = fma(gmi, gmi, b);
            }
        }
        nfix = block_sum_int(nfix, shi);
        if (MASK) {
            a = block_sum(a, shd);
            b = block_sum(b, shd);
        }
    }
    if (threadIdx.x == 0) {
        c.sd->n_at_bound = cnt;
        if (add) {
            c.sd->nb_fix = nfix;
            if (MASK) {
                c.sd->nrg_g = sqrt(a);
                c.sd->nrg_gm = sqrt(b);
            }
        }
    }
    publish(c.sd, c.sh);
}

// active_bounds only (flags into c.at, count into n_at_bound) -- fine-grained ABI
__global__ void k_active_flags(VecCtx c, const double* x, const double* s, double delta) {
    __shared__ int shi[32];
    int cnt = 0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        const double sl = fmax(c.xlow[i] - x[i], -delta);
        const double su = fmin(c.xupp[i] - x[i], delta);
        const bool at = (s[i] - sl <= c.atol_active) || (su - s[i] <= c.atol_active);
        c.at[i] = at ? 1 : 0;
        cnt += at ? 1 : 0;
    }
    cnt = block_sum_int(cnt, shi);
    if (threadIdx.x == 0) c.sd->n_at_bound = cnt;
    publish(c.sd, c.sh);
}

__global__ void k_publish(Scal* sd, Scal* sh) { publish(sd, sh); }
__global__ void k_sphere_value(VecCtx c, const double* x, double rho2) {
    __shared__ double shd[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) a = fma(x[i], x[i], a);
    a = block_sum(a, shd);
    if (threadIdx.x == 0) c.sd->c0 = a - rho2;
    publish(c.sd, c.sh);
}
__global__ void k_sphere_jac(VecCtx c, const double* x) {
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) c.C[i] = 2.0 * x[i];
}
__global__ void k_mask_project(VecCtx c, const double* src, double* dst) {
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) dst[i] = c.fix[i] ? 0.0 : src[i];
}

__global__ void k_dot_gs(VecCtx c) {  // dot(g,s) :458
    __shared__ double shd[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) a = fma(c.g[i], c.s[i], a);
    a = block_sum(a, shd);
    if (threadIdx.x == 0) {
        c.sd->gs = a;
        c.sd->jv_sumsq = c.hv[c.ld];  // dot(Jv,Jv) of the preceding vthv pass
    }
    publish(c.sd, c.sh);
}

__global__ void k_trial_point(VecCtx c) {  // x_next = x+s :351 ; norm(s) :356
    __shared__ double shd[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        const double si = c.s[i];
        c.xn[i] = c.x[i] + si;
        a = fma(si, si, a);
    }
    a = block_sum(a, shd);
    if (threadIdx.x == 0) c.sd->norm_s = sqrt(a);
    publish(c.sd, c.sh);
}

template <bool MASK>
__global__ void k_pix(VecCtx c) {  // criticality_measure :839-844 ; norm(g) for initial_tr :817-819
    __shared__ double shd[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        const double gi = c.g[i];
        b = fma(gi, gi, b);
        if (MASK && !c.fix[i]) a = fma(gi, gi, a);
    }
    a = block_sum(a, shd);
    b = block_sum(b, shd);
    if (threadIdx.x == 0) {
        if (MASK) c.sd->pix = sqrt(a);
        c.sd->norm_g = sqrt(b);
    }
    publish(c.sd, c.sh);
}

// C-part of AlHessian: cv = C v; Cv_sumsq = dot(Cv,Cv) (:94-95); hv (+)= C' ((mu*C) v) (:104-105)
__global__ void k_hess_c(VecCtx c, const double* v, double* hv, int add_to_hv) {
    __shared__ double shd[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < c.p; i += nw) {
        double a = 0.0, b = 0.0;
        for (int j = lane; j < c.n; j += 32) {
            a = fma(c.C[(size_t)i * c.ld + j], v[j], a);
            b = fma(c.muC[(size_t)i * c.ld + j], v[j], b);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            c.cv[i] = a;
            c.pvec[i] = b;
        }
    }
    __syncthreads();
    double q = 0.0;
    for (int i = threadIdx.x; i < c.p; i += blockDim.x) q = fma(c.cv[i], c.cv[i], q);
    q = block_sum(q, shd);
    if (threadIdx.x == 0) c.sd->Cv_sumsq = q;
    if (hv != nullptr) {
        for (int j = threadIdx.x; j < c.n; j += blockDim.x) {
            double a = 0.0;
            for (int i = 0; i < c.p; ++i) a = fma(c.C[(size_t)i * c.ld + j], c.pvec[i], a);
            hv[j] = add_to_hv ? (hv[j] + a) : a;
        }
    }
    publish(c.sd, c.sh);
}

__global__ void k_add_Ct(VecCtx c, const double* pv, double* gout) {  // g += Cx' * y_bar :45,:74
    for (int j = threadIdx.x; j < c.n; j += blockDim.x) {
        double a = 0.0;
        for (int i = 0; i < c.p; ++i) a = fma(c.C[(size_t)i * c.ld + j], pv[i], a);
        gout[j] = gout[j] + a;
    }
}

__global__ void k_scale_C(VecCtx c) {
    const size_t tot = (size_t)c.p * c.ld;
    for (size_t i = threadIdx.x; i < tot; i += blockDim.x) c.muC[i] = c.mu * c.C[i];
}

// dot(rx,rx) :44,:59 -- one partial per row chunk (rowgeom.h): CTA (gi, b) sums its chunk in a fixed thread pattern
__global__ void __launch_bounds__(256) k_sumsq_chunks(const double* __restrict__ r, RowGeom geo, double* __restrict__ partial) {
    __shared__ double shd[32];
    const int gi = blockIdx.x / geo.G, b = blockIdx.x % geo.G;
    const long long lb = geo.local_begin(gi, b), le = geo.local_end(gi, b);
    double a = 0.0;
    for (long long i = lb + threadIdx.x; i < le; i += blockDim.x) a = fma(r[i], r[i], a);
    a = block_sum(a, shd);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}

// column-major (Julia) rows x cols, leading dim lds  ->  row-major rows x ldd (zero padded)
__global__ void k_transpose_in(const double* __restrict__ src, long long rows, int cols, long long lds,
                               double* __restrict__ dst, int ldd) {
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + threadIdx.x;
        const int cc = c0 + j;
        tile[j][threadIdx.x] = (r < rows && cc < cols) ? src[(size_t)cc * lds + r] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + j;
        const int cc = c0 + threadIdx.x;
        if (r < rows && cc < ldd) dst[(size_t)r * ldd + cc] = tile[threadIdx.x][j];
    }
}

__global__ void k_pack_fix(const unsigned char* fix, int n, unsigned long long* words) {
    const int nw = (n + 63) / 64;
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {
        unsigned long long x = 0;
        for (int b = 0; b < 64; ++b) {
            const int i = w * 64 + b;
            if (i < n && fix[i]) x |= (1ull << b);
        }
        words[w] = x;
    }
}
__global__ void k_unpack_fix(const unsigned long long* words, int n, unsigned char* fix, Scal* sd, Scal* sh) {
    __shared__ int shi[32];
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned char f = (words[i >> 6] >> (i & 63)) & 1ull;
        fix[i] = f;
        cnt += f;
    }
    cnt = block_sum_int(cnt, shi);
    if (threadIdx.x == 0) sd->nb_fix = cnt;
    publish(sd, sh);
}
__global__ void k_set_flags(VecCtx c, const long long* idx, int count) {
    __shared__ int shi[32];
    for (int k = threadIdx.x; k < count; k += blockDim.x) c.fix[idx[k]] = 1;
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) cnt += c.fix[i];
    cnt = block_sum_int(cnt, shi);
    if (threadIdx.x == 0) c.sd->nb_fix = cnt;
    publish(c.sd, c.sh);
}
// findall(flags): ascending indices.  Single CTA, chunked ballot scan.
__global__ void k_list_flags(const unsigned char* flags, int n, long long* idx_out, int* count_out) {
    __shared__ int warp_cnt[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const bool f = (i < n) && flags[i];
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (f) idx_out[off + __popc(m & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += warp_cnt[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = base;
}

}  // namespace

void vk_active_reset(const VecCtx& c, const double* xa, const double* sa, cudaStream_t st) {
    k_active_reset<<<1, kVT, 0, st>>>(c, xa, sa);
}
void vk_cauchy_init(const VecCtx& c, bool mask, cudaStream_t st) {
    if (mask)
        k_cauchy_init<true><<<1, kVT, 0, st>>>(c);
    else
        k_cauchy_init<false><<<1, kVT, 0, st>>>(c);
}
void vk_cauchy_eval(const VecCtx& c, double delta, cudaStream_t st) { k_cauchy_eval<<<1, kVT, 0, st>>>(c, delta); }
void vk_cauchy_advance(const VecCtx& c, bool mask, int breakpoint, cudaStream_t st) {
    if (mask)
        k_cauchy_advance<true><<<1, kVT, 0, st>>>(c, breakpoint);
    else
        k_cauchy_advance<false><<<1, kVT, 0, st>>>(c, breakpoint);
}
void vk_gminor_nrg(const VecCtx& c, bool mask, cudaStream_t st) {
    if (mask)
        k_gminor_nrg<true><<<1, kVT, 0, st>>>(c);
    else
        k_gminor_nrg<false><<<1, kVT, 0, st>>>(c);
}
void vk_norm_to(const VecCtx& c, const double* v, int which, cudaStream_t st) { k_norm_to<<<1, kVT, 0, st>>>(c, v, which); }
void vk_cg_init(const VecCtx& c, bool mask, double delta, bool bounds_given, cudaStream_t st) {
    if (mask)
        k_cg_init<true><<<1, kVT, 0, st>>>(c, delta, bounds_given ? 1 : 0);
    else
        k_cg_init<false><<<1, kVT, 0, st>>>(c, delta, bounds_given ? 1 : 0);
}
void vk_cg_step(const VecCtx& c, bool mask, int phase, cudaStream_t st) {
    if (mask)
        k_cg_step<true, 0><<<1, kVT, 0, st>>>(c);
    else if (phase == 0)
        k_cg_step<false, 0><<<1, kVT, 0, st>>>(c);
    else
        k_cg_step<false, 1><<<1, kVT, 0, st>>>(c);
}
void vk_minor_finish(const VecCtx& c, cudaStream_t st) { k_minor_finish<<<1, kVT, 0, st>>>(c); }
void vk_minor_post(const VecCtx& c, bool mask, double delta, cudaStream_t st) {
    if (mask)
        k_minor_post<true><<<1, kVT, 0, st>>>(c, delta);
    else
        k_minor_post<false><<<1, kVT, 0, st>>>(c, delta);
}
void vk_dot_gs(const VecCtx& c, cudaStream_t st) { k_dot_gs<<<1, kVT, 0, st>>>(c); }
void vk_trial_point(const VecCtx& c, cudaStream_t st) { k_trial_point<<<1, kVT, 0, st>>>(c); }
void vk_pix(const VecCtx& c, bool mask, cudaStream_t st) {
    if (mask)
        k_pix<true><<<1, kVT, 0, st>>>(c);
    else
        k_pix<false><<<1, kVT, 0, st>>>(c);
}
void vk_hess_c(const VecCtx& c, const double* v, double* hv, bool add_to_hv, cudaStream_t st) {
    k_hess_c<<<1, kVT, 0, st>>>(c, v, hv, add_to_hv ? 1 : 0);
}
void vk_add_Ct(const VecCtx& c, const double* pv, double* gout, cudaStream_t st) { k_add_Ct<<<1, kVT, 0, st>>>(c, pv, gout); }
void vk_scale_C(const VecCtx& c, cudaStream_t st) { k_scale_C<<<1, kVT, 0, st>>>(c); }
void vk_sumsq_chunks(const double* r, const RowGeom& geo, double* partial, cudaStream_t st) {
    k_sumsq_chunks<<<geo.ng * geo.G, 256, 0, st>>>(r, geo, partial);
}
void vk_transpose_in(const double* src, long long rows, int cols, long long lds, double* dst, int ldd, cudaStream_t st) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((ldd + 31) / 32));
    dim3 block(32, 8);
    k_transpose_in<<<grid, block, 0, st>>>(src, rows, cols, lds, dst, ldd);
}
void vk_pack_fix(const unsigned char* fix, int n, unsigned long long* words, cudaStream_t st) {
    k_pack_fix<<<1, 256, 0, st>>>(fix, n, words);
}
void vk_unpack_fix(const unsigned long long* words, int n, unsigned char* fix, Scal* sd, Scal* sh, cudaStream_t st) {
    k_unpack_fix<<<1, kVT, 0, st>>>(words, n, fix, sd, sh);
}
void vk_set_flags(const VecCtx& c, const long long* idx, int count, cudaStream_t st) {
    k_set_flags<<<1, kVT, 0, st>>>(c, idx, count);
}
void vk_list_flags(const unsigned char* flags, int n, long long* idx_out, int* count_out, cudaStream_t st) {
    k_list_flags<<<1, kVT, 0, st>>>(flags, n, idx_out, count_out);
}
void vk_mask_project(const VecCtx& c, const double* src, double* dst, cudaStream_t st) { k_mask_project<<<1, kVT, 0, st>>>(c, src, dst); }
void vk_sphere_value(const VecCtx& c, const double* x, double rho2, cudaStream_t st) { k_sphere_value<<<1, kVT, 0, st>>>(c, x, rho2); }
void vk_sphere_jac(const VecCtx& c, const double* x, cudaStream_t st) { k_sphere_jac<<<1, kVT, 0, st>>>(c, x); }
void vk_publish(Scal* sd, Scal* sh, cudaStream_t st) { k_publish<<<1, 64, 0, st>>>(sd, sh); }
void vk_active_flags(const VecCtx& c, const double* x, const double* s, double delta, cudaStream_t st) {
    k_active_flags<<<1, kVT, 0, st>>>(c, x, s, delta);
}

}  // namespace bnl
