// models.cu -- device generators for the synthetic problem families (the reference ships none; user
// callbacks `residuals` / `jac_res` are evaluated in new_point / evaluate_al / first_derivatives,
// src/basic_tralcnlss.jl:32-77).  Executable specification: oracle/models.py (same 32-bit counter hash).
//   GLM    : a_ij = sym(seed,i,j)*cs_j ; r_i = phi(a_i.x) - y_i ; J_ij = phi'(a_i.x) a_ij ; phi(z) = z + 0.1 sin z
//   EXPSUM : channel c = i mod C ; r_i = a_c exp(-b_c t_i) - y_i ; J dense with two non-zeros per row
// J rows are written straight to HBM in the row-major panel layout the streaming kernels read (8*M*ld bytes).
#include "common.cuh"
#include "models.h"

namespace bnl {
namespace {

constexpr int kWarpsPerCta = 8;

// ---------------------------------------------------------------- GLM ------------------------------------
// one warp per row; lane owns column pairs (2*lane + 64k, +1)
__device__ __forceinline__ double glm_row_dot(const ModelArgs& a, uint32_t rk, const double* __restrict__ x, int lane) {
    double z = 0.0;
    const int NC = a.ld >> 1;
    for (int c = lane; c < NC; c += 32) {
        const int j = 2 * c;
        const double2 cs2 = __ldg(reinterpret_cast<const double2*>(a.cs) + c);
        const double2 x2 = __ldg(reinterpret_cast<const double2*>(x) + c);
        const double a0 = usym(hash_rc(rk, (uint32_t)j)) * cs2.x;
        const double a1 = usym(hash_rc(rk, (uint32_t)(j + 1))) * cs2.y;
        z = fma(a0, x2.x, z);
        z = fma(a1, x2.y, z);
    }
    return warp_sum(z);
}

// WHAT: 0 setup y, 1 residual (+ sumsq partial), 2 jacobian
template <int WHAT>
__global__ void __launch_bounds__(kWarpsPerCta * 32) glm_kernel(ModelArgs a, const double* __restrict__ x,
                                                                const double* __restrict__ yin, double* __restrict__ out,
                                                                double* __restrict__ J, double* __restrict__ partial) {
    __shared__ double shd[32];
    const int lane = threadIdx.x & 31;
    // WHAT == 1 carries a row reduction (sum r_i^2): CTA (gi, b) owns one row chunk and a warp takes every 8th row of it, so
    // the partial only depends on chunk-local indices (rowgeom.h).  The element-wise kernels stride over all local rows.
    long long i0, i1, istep;
    if (WHAT == 1) {
        const int cg = blockIdx.x / a.geo.G, cb = blockIdx.x % a.geo.G;
        i0 = a.geo.local_begin(cg, cb) + (threadIdx.x >> 5);
        i1 = a.geo.local_end(cg, cb);
        istep = kWarpsPerCta;
    } else {
        i0 = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
        i1 = a.M;
        istep = (long long)gridDim.x * kWarpsPerCta;
    }
    double ss = 0.0;
    for (long long i = i0; i < i1; i += istep) {
        const unsigned long long gi = (unsigned long long)(a.row0 + i);
        const uint32_t rk = rowkey(a.seed, gi);
        const double z = glm_row_dot(a, rk, x, lane);
        if (WHAT == 0) {
            if (lane == 0) out[i] = z + 0.1 * sin(z) + a.noise * usym(hash_rc(rowkey(a.seed + 1u, gi), 0u));
        } else if (WHAT == 1) {
            if (lane == 0) {
                const double r = z + 0.1 * sin(z) - yin[i];
                out[i] = r;
                ss = fma(r, r, ss);
            }
        } else {
            const double dphi = 1.0 + 0.1 * cos(z);
            const int NC = a.ld >> 1;
            double2* Jrow = reinterpret_cast<double2*>(J + (size_t)i * a.ld);
            for (int c = lane; c < NC; c += 32) {
                const int j = 2 * c;
                const double2 cs2 = __ldg(reinterpret_cast<const double2*>(a.cs) + c);
                double2 o;
                o.x = dphi * (usym(hash_rc(rk, (uint32_t)j)) * cs2.x);
                o.y = dphi * (usym(hash_rc(rk, (uint32_t)(j + 1))) * cs2.y);
                Jrow[c] = o;  // padding columns: cs = 0 => exact zeros
            }
        }
    }
    if (WHAT == 1) {
        ss = block_sum(ss, shd);
        if (threadIdx.x == 0) partial[blockIdx.x] = ss;
    }
}

// ---------------------------------------------------------------- EXPSUM ---------------------------------
__device__ __forceinline__ double expsum_t(const ModelArgs& a, long long gi, int C) {
    const long long denom = (a.M_total + C - 1) / C;
    return ((double)(gi / C) + 0.5) / (double)denom;
}

template <int WHAT>
__global__ void expsum_rows_kernel(ModelArgs a, const double* __restrict__ x, const double* __restrict__ yin,
                                   double* __restrict__ out, double* __restrict__ partial) {
    __shared__ double shd[32];
    const int C = a.n / 2;
    double ss = 0.0;
    long long i0, i1, stride;
    if (WHAT == 1) {  // one row chunk per CTA (rowgeom.h)
        const int cg = blockIdx.x / a.geo.G, cb = blockIdx.x % a.geo.G;
        i0 = a.geo.local_begin(cg, cb) + threadIdx.x;
        i1 = a.geo.local_end(cg, cb);
        stride = blockDim.x;
    } else {
        i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        i1 = a.M;
        stride = (long long)gridDim.x * blockDim.x;
    }
    for (long long i = i0; i < i1; i += stride) {
        const long long gi = a.row0 + i;
        const int c = (int)(gi % C);
        const double t = expsum_t(a, gi, C);
        const double m = x[c] * exp(-x[C + c] * t);
        if (WHAT == 0) {
            out[i] = m + a.noise * usym(hash_rc(rowkey(a.seed + 1u, (unsigned long long)gi), 0u));
        } else {
            const double r = m - yin[i];
            out[i] = r;
            ss = fma(r, r, ss);
        }
    }
    if (WHAT == 1) {
        ss = block_sum(ss, shd);
        if (threadIdx.x == 0) partial[blockIdx.x] = ss;
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) expsum_jac_kernel(ModelArgs a, const double* __restrict__ x,
                                                                       double* __restrict__ J) {
    const int C = a.n / 2;
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * kWarpsPerCta;
    const int NC = a.ld >> 1;
    for (long long i = gw; i < a.M; i += nw) {
        const long long gi = a.row0 + i;
        const int c = (int)(gi % C);
        const double t = expsum_t(a, gi, C);
        const double e = exp(-x[C + c] * t);
        const double de = -(x[c] * t) * e;
        double2* Jrow = reinterpret_cast<double2*>(J + (size_t)i * a.ld);
        for (int cc = lane; cc < NC; cc += 32) {
            const int j = 2 * cc;
            double2 o;
            o.x = (j == c) ? e : ((j == C + c) ? de : 0.0);
            o.y = (j + 1 == c) ? e : ((j + 1 == C + c) ? de : 0.0);
            Jrow[cc] = o;
        }
    }
}

int rows_grid(long long M, int nblocks_cap) {
    long long g = (M + kWarpsPerCta - 1) / kWarpsPerCta;
    if (g > nblocks_cap) g = nblocks_cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

cudaError_t model_setup_y(const ModelArgs& a, const double* x_true, double* y, cudaStream_t st) {
    if (a.model_id == 1) {
        glm_kernel<0><<<rows_grid(a.M, 148 * 8), kWarpsPerCta * 32, 0, st>>>(a, x_true, nullptr, y, nullptr, nullptr);
    } else if (a.model_id == 2) {
        expsum_rows_kernel<0><<<rows_grid(a.M / 32 + 1, 148 * 8), 256, 0, st>>>(a, x_true, nullptr, y, nullptr);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t model_residual(const ModelArgs& a, const double* x, const double* y, double* r, double* partial, cudaStream_t st) {
    const int grid = a.geo.ng * a.geo.G;  // one CTA per row chunk; partial[grid] goes through the fixed reduction tree (p2p.h)
    if (a.model_id == 1) {
        glm_kernel<1><<<grid, kWarpsPerCta * 32, 0, st>>>(a, x, y, r, nullptr, partial);
    } else if (a.model_id == 2) {
        expsum_rows_kernel<1><<<grid, 256, 0, st>>>(a, x, y, r, partial);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t model_jacobian(const ModelArgs& a, const double* x, double* J, cudaStream_t st) {
    if (a.model_id == 1) {
        glm_kernel<2><<<rows_grid(a.M, 148 * 8), kWarpsPerCta * 32, 0, st>>>(a, x, nullptr, nullptr, J, nullptr);
    } else if (a.model_id == 2) {
        expsum_jac_kernel<<<rows_grid(a.M, 148 * 8), kWarpsPerCta * 32, 0, st>>>(a, x, J);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace bnl
