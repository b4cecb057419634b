// dense.cu -- see dense.h.  Straightforward single-CTA column algorithms; exactness matters more than speed
// here (systems are small and replicated on every GPU).
#include "dense.h"

namespace bnl {
namespace {

constexpr int kDT = 1024;

// In-place lower Cholesky of the dim x dim column-major matrix S (leading dim lds).  Sets *fail on a pivot <= 0.
__device__ void chol_inplace(double* S, int dim, int lds, int* fail) {
    for (int k = 0; k < dim; ++k) {
        __syncthreads();
        const double akk = S[(size_t)k * lds + k];
        if (!(akk > 0.0)) {
            if (threadIdx.x == 0) *fail = 1;
            return;  // uniform: every thread reads the same akk
        }
        const double lkk = sqrt(akk);
        __syncthreads();
        for (int i = k + threadIdx.x; i < dim; i += blockDim.x) {
            const double v = S[(size_t)k * lds + i];
            S[(size_t)k * lds + i] = (i == k) ? lkk : v / lkk;
        }
        __syncthreads();
        // trailing update: S[i][j] -= L[i][k] * L[j][k] for k < j <= i
        const int rem = dim - k - 1;
        const long long tot = (long long)rem * rem;
        for (long long e = threadIdx.x; e < tot; e += blockDim.x) {
            const int jj = (int)(e / rem), ii = (int)(e % rem);
            if (ii >= jj) {
                const int i = k + 1 + ii, j = k + 1 + jj;
                S[(size_t)j * lds + i] -= S[(size_t)k * lds + i] * S[(size_t)k * lds + j];
            }
        }
    }
    __syncthreads();
}

__global__ void k_chol_aat(DenseCtx c) {
    const int m = c.m;
    // LA = A A' (lower), then factor
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int j = e / m, i = e % m;
        double s = 0.0;
        if (i >= j) {
            for (int k = 0; k < c.n; ++k) s = fma(c.A[(size_t)i * c.ld + k], c.A[(size_t)j * c.ld + k], s);
        }
        c.LA[(size_t)j * m + i] = s;
    }
    __syncthreads();
    chol_inplace(c.LA, m, m, &c.sd->chol_fail);
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {  // zero the strict upper triangle
        const int j = e / m, i = e % m;
        if (i < j) c.LA[(size_t)j * m + i] = 0.0;
    }
}

// cholesky_aug_aat (src/polyhedral_constraints.jl:35-59): fixidx = findall(fix); G = L_A \ A[:,fix];
// S = I - G'G; L = [L_A 0; G' chol(S)]
__global__ void k_rebuild(DenseCtx c, const unsigned char* fix) {
    __shared__ int warp_cnt[32];
    __shared__ int base;
    const int m = c.m, cap = c.cap;
    // ---- findall(fix), ascending ----
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int start = 0; start < c.n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const bool f = (i < c.n) && fix[i];
        const unsigned msk = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[warp] = __popc(msk);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (f) c.fixidx[off + __popc(msk & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += warp_cnt[w];
            base += t;
        }
        __syncthreads();
    }
    const int q = base;
    if (threadIdx.x == 0) {
        *c.q_dev = q;
        c.sd->nb_fix = q;
    }
    // ---- G = L_A \ A[:,fix]  (one thread per column, forward substitution) ----
    for (int k = threadIdx.x; k < q; k += blockDim.x) {
        const long long col = c.fixidx[k];
        for (int i = 0; i < m; ++i) {
            double s = c.A[(size_t)i * c.ld + col];
            for (int j = 0; j < i; ++j) s -= c.LA[(size_t)j * m + i] * c.G[(size_t)k * m + j];
            c.G[(size_t)k * m + i] = s / c.LA[(size_t)i * m + i];
        }
    }
    __syncthreads();
    // ---- assemble L: top-left L_A, bottom-left G', bottom-right S = I - G'G (lower) ----
    const int mpp = m + q;
    for (long long e = threadIdx.x; e < (long long)mpp * mpp; e += blockDim.x) {
        const int j = (int)(e / mpp), i = (int)(e % mpp);
        double v = 0.0;
        if (i >= j) {
            if (j < m) {
                v = (i < m) ? c.LA[(size_t)j * m + i] : c.G[(size_t)(i - m) * m + j];
            } else {
                double s = (i == j) ? 1.0 : 0.0;
                for (int k = 0; k < m; ++k) s -= c.G[(size_t)(i - m) * m + k] * c.G[(size_t)(j - m) * m + k];
                v = s;
            }
        }
        c.L[(size_t)j * cap + i] = v;
    }
    __syncthreads();
    chol_inplace(c.L + (size_t)m * cap + m, q, cap, &c.sd->chol_fail);
    // publish nb_fix / chol_fail
    __syncthreads();
    const int nwords = sizeof(Scal) / 8;
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(c.sd);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(c.sh);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) d[i] = s[i];
}

// projection! (:158-170): y = L \ (A~ r); w = L' \ y; v = r - A~' w   (r optionally negated first)
__global__ void k_project(DenseCtx c, const double* r, double* v, int negate) {
    const int m = c.m, cap = c.cap;
    const int q = *c.q_dev;
    const int mpp = m + q;
    const double sgn = negate ? -1.0 : 1.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* y = c.ywork;
    // left_mul :86-98
    for (int i = warp; i < m; i += nw) {
        double s = 0.0;
        for (int j = lane; j < c.n; j += 32) s = fma(c.A[(size_t)i * c.ld + j], sgn * r[j], s);
        s = warp_sum(s);
        if (lane == 0) y[i] = s;
    }
    for (int k = threadIdx.x; k < q; k += blockDim.x) y[m + k] = sgn * r[c.fixidx[k]];
    __syncthreads();
    // forward substitution  L y' = y  (column-oriented)
    for (int k = 0; k < mpp; ++k) {
        const double yk = y[k] / c.L[(size_t)k * cap + k];
        __syncthreads();
        if (threadIdx.x == 0) y[k] = yk;
        for (int i = k + 1 + threadIdx.x; i < mpp; i += blockDim.x) y[i] -= c.L[(size_t)k * cap + i] * yk;
        __syncthreads();
    }
    // backward substitution  L' w = y  (row-oriented on L: w_k = (y_k - sum_{i>k} L[i][k] w_i) / L[k][k])
    __shared__ double shd[32];
    for (int k = mpp - 1; k >= 0; --k) {
        double s = 0.0;
        for (int i = k + 1 + threadIdx.x; i < mpp; i += blockDim.x) s = fma(c.L[(size_t)k * cap + i], y[i], s);
        s = block_sum(s, shd);
        if (threadIdx.x == 0) y[k] = (y[k] - s) / c.L[(size_t)k * cap + k];
        __syncthreads();
    }
    // left_mul_tr :72-84 and v = r - A~' w
    for (int j = threadIdx.x; j < c.n; j += blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s = fma(c.A[(size_t)i * c.ld + j], y[i], s);
        v[j] = s;  // temporarily A' w
    }
    __syncthreads();
    for (int k = threadIdx.x; k < q; k += blockDim.x) v[c.fixidx[k]] += y[m + k];
    __syncthreads();
    for (int j = threadIdx.x; j < c.n; j += blockDim.x) v[j] = sgn * r[j] - v[j];
}

__global__ void k_left_mul(DenseCtx c, const double* x, double* y) {
    const int m = c.m, q = *c.q_dev;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < m; i += nw) {
        double s = 0.0;
        for (int j = lane; j < c.n; j += 32) s = fma(c.A[(size_t)i * c.ld + j], x[j], s);
        s = warp_sum(s);
        if (lane == 0) y[i] = s;
    }
    for (int k = threadIdx.x; k < q; k += blockDim.x) y[m + k] = x[c.fixidx[k]];
}
__global__ void k_left_mul_tr(DenseCtx c, const double* y, double* x) {
    const int m = c.m, q = *c.q_dev;
    for (int j = threadIdx.x; j < c.n; j += blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s = fma(c.A[(size_t)i * c.ld + j], y[i], s);
        x[j] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < q; k += blockDim.x) x[c.fixidx[k]] += y[m + k];
}

// ---- reduced-space projection --------------------------------------------------------------------------------
// Lr = cholesky(A_free A_free').L   (m x m, column-major, global memory; factorisation by warp 0)
__global__ void k_rs_rebuild(DenseCtx c, const unsigned char* fix) {
    const int m = c.m;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int j = e / m, i = e % m;
        double s = 0.0;
        if (i >= j) {
            const double* ai = c.A + (size_t)i * c.ld;
            const double* aj = c.A + (size_t)j * c.ld;
            for (int k = 0; k < c.n; ++k)
                if (!fix[k]) s = fma(ai[k], aj[k], s);
        }
        c.Lr[(size_t)j * m + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int k = 0; k < m; ++k) {
            const double akk = c.Lr[(size_t)k * m + k];
            if (!(akk > 0.0)) {
                if (lane == 0) c.sd->chol_fail = 1;
                break;
            }
            const double lkk = sqrt(akk);
            __syncwarp();
            for (int i = k + lane; i < m; i += 32) {
                const double v = c.Lr[(size_t)k * m + i];
                c.Lr[(size_t)k * m + i] = (i == k) ? lkk : v / lkk;
            }
            __syncwarp();
            for (int j = k + 1; j < m; ++j) {
                const double ljk = c.Lr[(size_t)k * m + j];
                for (int i = j + lane; i < m; i += 32) c.Lr[(size_t)j * m + i] -= c.Lr[(size_t)k * m + i] * ljk;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    const int nwords = sizeof(Scal) / 8;
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(c.sd);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(c.sh);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) d[i] = s[i];
}

// One variable left the free set (add_active!(ind), src/polyhedral_constraints.jl:240-249): A_free A_free' loses the rank-one
// term a a' with a = A[:,ind], so its Cholesky factor is DOWNDATED in O(m^2) (hyperbolic rotations, LINPACK dchdd) instead of
// being rebuilt from the O(m^2 n) products -- the reference rebuilds its (m+q)^2 factor from scratch here and flags the cost
// itself (:51).  Executed by warp 0 of the CTA on a shared-memory copy of the factor (sL: m*m doubles, sa: m doubles).
__device__ void rs_downdate_dev(const DenseCtx& c, long long ind, double* sL, double* sa) {
    const int m = c.m;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) sL[e] = c.Lr[e];
    for (int i = threadIdx.x; i < m; i += blockDim.x) sa[i] = c.A[(size_t)i * c.ld + ind];
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int k = 0; k < m; ++k) {
            const double lkk = sL[(size_t)k * m + k], ak = sa[k];
            const double r2 = lkk * lkk - ak * ak;
            if (!(r2 > 0.0)) {  // A_free lost full row rank: PosDefException in the reference (:57)
                if (lane == 0) c.sd->chol_fail = 1;
                break;
            }
            const double r = sqrt(r2), cc = r / lkk, ss = ak / lkk;
            __syncwarp();
            if (lane == 0) sL[(size_t)k * m + k] = r;
            for (int i = k + 1 + lane; i < m; i += 32) {
                const double lik = (sL[(size_t)k * m + i] - ss * sa[i]) / cc;
                sL[(size_t)k * m + i] = lik;
                sa[i] = cc * sa[i] - ss * lik;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) c.Lr[e] = sL[e];
    __syncthreads();
}

__device__ void publish_scal(const DenseCtx& c) {
    __syncthreads();
    const int nwords = sizeof(Scal) / 8;
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(c.sd);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(c.sh);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) d[i] = s[i];
}

__global__ void k_rs_downdate(DenseCtx c) {
    extern __shared__ double dsm[];
    const long long ind = c.sd->bp_ind;
    if (ind >= 0) rs_downdate_dev(c, ind, dsm, dsm + (size_t)c.m * c.m);
    publish_scal(c);
}

// v = P(+-r) in the reduced space (whole CTA): t = A_free r_free ; Lr y = t ; Lr' w = y ; v_free = r_free - A_free' w ; v_F = 0
__device__ void rs_project_dev(const DenseCtx& c, const unsigned char* fix, const double* r, double* v, int negate) {
    const int m = c.m;
    const double sgn = negate ? -1.0 : 1.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* y = c.ywork;
    for (int i = warp; i < m; i += nw) {
        double s = 0.0;
        for (int j = lane; j < c.n; j += 32)
            if (!fix[j]) s = fma(c.A[(size_t)i * c.ld + j], sgn * r[j], s);
        s = warp_sum(s);
        if (lane == 0) y[i] = s;
    }
    __syncthreads();
    if (warp == 0) {
        // Lr y' = t ; Lr' w = y'   (warp 0, column-oriented forward, row-oriented backward)
        for (int k = 0; k < m; ++k) {
            const double yk = y[k] / c.Lr[(size_t)k * m + k];
            __syncwarp();
            if (lane == 0) y[k] = yk;
            for (int i = k + 1 + lane; i < m; i += 32) y[i] -= c.Lr[(size_t)k * m + i] * yk;
            __syncwarp();
        }
        for (int k = m - 1; k >= 0; --k) {
            double s = 0.0;
            for (int i = k + 1 + lane; i < m; i += 32) s = fma(c.Lr[(size_t)k * m + i], y[i], s);
            s = warp_sum(s);
            __syncwarp();
            if (lane == 0) y[k] = (y[k] - s) / c.Lr[(size_t)k * m + k];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c.n; j += blockDim.x) {
        double out = 0.0;
        if (!fix[j]) {
            double s = 0.0;
            for (int i = 0; i < m; ++i) s = fma(c.A[(size_t)i * c.ld + j], y[i], s);
            out = sgn * r[j] - s;
        }
        v[j] = out;
    }
}

__global__ void k_rs_project(DenseCtx c, const unsigned char* fix, const double* r, double* v, int negate) {
    rs_project_dev(c, fix, r, v, negate);
}

// One Cauchy breakpoint of the general path in ONE launch (src/basic_tralcnlss.jl:628-632): s_c += theta d ; add_active!(ind)
// (flag + factor downdate) ; d = P(-g).  Same arithmetic, in the same order, as k_cauchy_advance<false> + k_rs_downdate +
// k_rs_project launched one after the other; the factor stays in shared memory between the downdate and the two triangular
// solves of the projection.
__global__ void k_rs_breakpoint(DenseCtx c, double* s, double* d, const double* g, unsigned char* fix) {
    extern __shared__ double dsm[];
    const int m = c.m;
    double* sL = dsm;                      // m x m factor
    double* sa = dsm + (size_t)m * m;      // m: the leaving column, then the projection's m-vector
    const double step = c.sd->theta;
    const long long ind = c.sd->bp_ind;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
        s[i] = s[i] + step * d[i];
        if (i == ind) fix[i] = 1;
    }
    if (threadIdx.x == 0) c.sd->nb_fix = c.sd->nb_fix + 1;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) sL[e] = c.Lr[e];
    if (ind >= 0)
        for (int i = threadIdx.x; i < m; i += blockDim.x) sa[i] = c.A[(size_t)i * c.ld + ind];
    __syncthreads();
    if (ind >= 0 && warp == 0) {  // downdate (rs_downdate_dev's arithmetic)
        for (int k = 0; k < m; ++k) {
            const double lkk = sL[(size_t)k * m + k], ak = sa[k];
            const double r2 = lkk * lkk - ak * ak;
            if (!(r2 > 0.0)) {
                if (lane == 0) c.sd->chol_fail = 1;
                break;
            }
            const double r = sqrt(r2), cc = r / lkk, ss = ak / lkk;
            __syncwarp();
            if (lane == 0) sL[(size_t)k * m + k] = r;
            for (int i = k + 1 + lane; i < m; i += 32) {
                const double lik = (sL[(size_t)k * m + i] - ss * sa[i]) / cc;
                sL[(size_t)k * m + i] = lik;
                sa[i] = cc * sa[i] - ss * lik;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) c.Lr[e] = sL[e];
    // d = P(-g) (rs_project_dev's arithmetic with negate = 1, factor and m-vector in shared memory)
    double* y = sa;
    for (int i = warp; i < m; i += nw) {
        double t = 0.0;
#pragma unroll 8
        for (int j = lane; j < c.n; j += 32)
            if (!fix[j]) t = fma(__ldg(c.A + (size_t)i * c.ld + j), -1.0 * g[j], t);
        t = warp_sum(t);
        if (lane == 0) y[i] = t;
    }
    __syncthreads();
    if (warp == 0) {
        for (int k = 0; k < m; ++k) {
            const double yk = y[k] / sL[(size_t)k * m + k];
            __syncwarp();
            if (lane == 0) y[k] = yk;
            for (int i = k + 1 + lane; i < m; i += 32) y[i] -= sL[(size_t)k * m + i] * yk;
            __syncwarp();
        }
        for (int k = m - 1; k >= 0; --k) {
            double t = 0.0;
            for (int i = k + 1 + lane; i < m; i += 32) t = fma(sL[(size_t)k * m + i], y[i], t);
            t = warp_sum(t);
            __syncwarp();
            if (lane == 0) y[k] = (y[k] - t) / sL[(size_t)k * m + k];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c.n; j += blockDim.x) {
        double out = 0.0;
        if (!fix[j]) {
            double t = 0.0;
#pragma unroll 8
            for (int i = 0; i < m; ++i) t = fma(__ldg(c.A + (size_t)i * c.ld + j), y[i], t);
            out = -1.0 * g[j] - t;
        }
        d[j] = out;
    }
    publish_scal(c);
}

}  // namespace

void dk_left_mul(const DenseCtx& c, const double* x, double* y, cudaStream_t st) { k_left_mul<<<1, kDT, 0, st>>>(c, x, y); }
void dk_left_mul_tr(const DenseCtx& c, const double* y, double* x, cudaStream_t st) { k_left_mul_tr<<<1, kDT, 0, st>>>(c, y, x); }
void dk_rs_rebuild(const DenseCtx& c, const unsigned char* fix, cudaStream_t st) { k_rs_rebuild<<<1, kDT, 0, st>>>(c, fix); }
static size_t rs_smem(const DenseCtx& c) { return ((size_t)c.m * c.m + c.m) * sizeof(double); }
bool dk_rs_downdate_fits(const DenseCtx& c) {  // the factor must fit in shared memory (m <= 160); above: rebuild instead
    static bool opted = false;
    if (rs_smem(c) <= 48 * 1024) return true;
    if (rs_smem(c) > 200 * 1024) return false;
    if (!opted) {
        opted = cudaFuncSetAttribute(k_rs_downdate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess &&
                cudaFuncSetAttribute(k_rs_breakpoint, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
    }
    return opted;
}
void dk_rs_downdate(const DenseCtx& c, cudaStream_t st) { k_rs_downdate<<<1, 256, rs_smem(c), st>>>(c); }
void dk_rs_breakpoint(const DenseCtx& c, double* s, double* d, const double* g, unsigned char* fix, cudaStream_t st) {
    k_rs_breakpoint<<<1, kDT, rs_smem(c), st>>>(c, s, d, g, fix);
}
void dk_rs_project(const DenseCtx& c, const unsigned char* fix, const double* r, double* v, bool negate, cudaStream_t st) {
    k_rs_project<<<1, kDT, 0, st>>>(c, fix, r, v, negate ? 1 : 0);
}

void dk_chol_aat(const DenseCtx& c, cudaStream_t st) { k_chol_aat<<<1, kDT, 0, st>>>(c); }
void dk_rebuild(const DenseCtx& c, const unsigned char* fix, cudaStream_t st) { k_rebuild<<<1, kDT, 0, st>>>(c, fix); }
void dk_project(const DenseCtx& c, const double* r, double* v, bool negate, cudaStream_t st) {
    k_project<<<1, kDT, 0, st>>>(c, r, v, negate ? 1 : 0);
}

}  // namespace bnl
