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
        const double a0 = usym_fast(hash_rc(rk, (uint32_t)j)) * cs2.x;
        const double a1 = usym_fast(hash_rc(rk, (uint32_t)(j + 1))) * cs2.y;
        z = fma(a0, x2.x, z);
        z = fma(a1, x2.y, z);
    }
    return warp_sum(z);
}

// jac_res(x): one warp per row, J_ij = phi'(a_i.x) a_ij written straight to HBM in the row-major panel layout (coalesced 16-byte
// stores); element-wise, no reduction
__global__ void __launch_bounds__(kWarpsPerCta * 32) glm_jac_kernel(ModelArgs a, const double* __restrict__ x, double* __restrict__ J) {
    const int lane = threadIdx.x & 31;
    const long long i0 = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const long long istep = (long long)gridDim.x * kWarpsPerCta;
    const int NC = a.ld >> 1;
    for (long long i = i0; i < a.M; i += istep) {
        const unsigned long long gi = (unsigned long long)(a.row0 + i);
        const uint32_t rk = rowkey(a.seed, gi);
        const double z = glm_row_dot(a, rk, x, lane);
        const double dphi = 1.0 + 0.1 * cos(z);
        double2* Jrow = reinterpret_cast<double2*>(J + (size_t)i * a.ld);
        for (int c = lane; c < NC; c += 32) {
            const int j = 2 * c;
            const double2 cs2 = __ldg(reinterpret_cast<const double2*>(a.cs) + c);
            double2 o;
            o.x = dphi * (usym_fast(hash_rc(rk, (uint32_t)j)) * cs2.x);
            o.y = dphi * (usym_fast(hash_rc(rk, (uint32_t)(j + 1))) * cs2.y);
            Jrow[c] = o;  // padding columns: cs = 0 => exact zeros
        }
    }
}

// jac_res(x) FUSED with J'r (first_derivatives, src/basic_tralcnlss.jl:72-74: g = Jx'*rx right after Jx = jac_res(x)): while a
// row of J is still in registers its contribution r_i J_i. is accumulated, so the accepted step does not stream the 8*M*n bytes
// back in.  The kernel walks the rows exactly like the streaming J'w kernel (matvec.cu, MODE_JTW): CTA b owns the chunks (g, b),
// warp w of 8 takes the RB-row stages w, w+8, ... of a chunk, lane u owns the double2 column chunks u + 32k, and acc += J*r is
// the same FMA in the same order -- the partials, and therefore g, are BIT-IDENTICAL to the unfused pass (tested).
template <int KCH>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) glm_jac_jtr_kernel(ModelArgs a, const double* __restrict__ x,
                                                                        const double* __restrict__ r, double* __restrict__ J,
                                                                        double* __restrict__ partial, long long pstride, int RB) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = blockIdx.x / a.geo.G, cb = blockIdx.x % a.geo.G;
    const long long lb = a.geo.local_begin(cg, cb), le = a.geo.local_end(cg, cb);
    const long long nst = (le - lb + RB - 1) / RB;
    const int NC = a.ld >> 1;
    double2 acc[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) acc[k] = make_double2(0.0, 0.0);
    for (long long st = warp; st < nst; st += kWarpsPerCta) {
        for (int q = 0; q < RB; ++q) {
            const long long i = lb + st * RB + q;
            if (i >= le) break;
            const unsigned long long gi = (unsigned long long)(a.row0 + i);
            const uint32_t rk = rowkey(a.seed, gi);
            const double z = glm_row_dot(a, rk, x, lane);  // same arithmetic as glm_jac_kernel
            const double dphi = 1.0 + 0.1 * cos(z);
            const double ri = __ldg(r + i);
            double2* Jrow = reinterpret_cast<double2*>(J + (size_t)i * a.ld);
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const int c = lane + 32 * k;
                if (c < NC) {
                    const int j = 2 * c;
                    const double2 cs2 = __ldg(reinterpret_cast<const double2*>(a.cs) + c);
                    double2 o;
                    o.x = dphi * (usym_fast(hash_rc(rk, (uint32_t)j)) * cs2.x);
                    o.y = dphi * (usym_fast(hash_rc(rk, (uint32_t)(j + 1))) * cs2.y);
                    Jrow[c] = o;
                    acc[k].x = fma(o.x, ri, acc[k].x);
                    acc[k].y = fma(o.y, ri, acc[k].y);
                }
            }
        }
    }
    double* pout = partial + ((size_t)blockIdx.x * kWarpsPerCta + warp) * pstride;  // [gi][b][team = warp][*]
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int c = lane + 32 * k;
        if (c < NC) reinterpret_cast<double2*>(pout)[c] = acc[k];
    }
}

// ---- GLM residual / data kernels: ONE THREAD PER ROW -----------------------------------------------------------------
// z_i = a_i . x needs n hashes per row and nothing from memory but x and cs, which every thread of a warp reads at the same j:
// they sit in shared memory and are broadcast (one LDS per warp and column), the hash and the FMAs run without shuffles.
// Four interleaved partial sums (columns j mod 4) in fixed order.  WHAT: 0 data y (element-wise), 1 residual + sum of squares
// per row chunk (CTA (gi, b) owns chunk (gi, b): rowgeom.h).
template <int WHAT>
__global__ void __launch_bounds__(256) glm_rows_kernel(ModelArgs a, const double* __restrict__ x, const double* __restrict__ yin,
                                                       double* __restrict__ out, double* __restrict__ partial) {
    extern __shared__ double2 sxc[];  // [ld] (x_j, cs_j)
    __shared__ double shd[32];
    for (int j = threadIdx.x; j < a.ld; j += blockDim.x) sxc[j] = make_double2(j < a.n ? x[j] : 0.0, a.cs[j]);
    __syncthreads();
    long long i0, i1, istep;
    if (WHAT == 1) {
        const int cg = blockIdx.x / a.geo.G, cb = blockIdx.x % a.geo.G;
        i0 = a.geo.local_begin(cg, cb) + threadIdx.x;
        i1 = a.geo.local_end(cg, cb);
        istep = blockDim.x;
    } else {
        i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        i1 = a.M;
        istep = (long long)gridDim.x * blockDim.x;
    }
    double ss = 0.0;
    const int n4 = a.ld & ~3;  // ld is a multiple of 16; padding columns have cs = 0
    for (long long i = i0; i < i1; i += istep) {
        const unsigned long long gi = (unsigned long long)(a.row0 + i);
        const uint32_t rk = rowkey(a.seed, gi);
        double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
        uint32_t key = rk;  // rk + j * 0x9E3779B9
#pragma unroll 2
        for (int j = 0; j < n4; j += 4) {
            const double2 p0 = sxc[j], p1 = sxc[j + 1], p2 = sxc[j + 2], p3 = sxc[j + 3];
            z0 = fma(usym_fast(mix32(key)) * p0.y, p0.x, z0);
            z1 = fma(usym_fast(mix32(key + 0x9E3779B9u)) * p1.y, p1.x, z1);
            z2 = fma(usym_fast(mix32(key + 2u * 0x9E3779B9u)) * p2.y, p2.x, z2);
            z3 = fma(usym_fast(mix32(key + 3u * 0x9E3779B9u)) * p3.y, p3.x, z3);
            key += 4u * 0x9E3779B9u;
        }
        const double z = (z0 + z1) + (z2 + z3);
        if (WHAT == 0) {
            out[i] = z + 0.1 * sin(z) + a.noise * usym(hash_rc(rowkey(a.seed + 1u, gi), 0u));
        } else {
            const double r = z + 0.1 * sin(z) - yin[i];
            out[i] = r;
            ss = fma(r, r, ss);
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

// ---------------------------------------------------------------- EXPSUM_DENSE ---------------------------
// BASELINE config[1] as SURVEY 8d words it: ONE sum of K = n/2 exponentials observed on one time grid,
//   model_i(x) = sum_k a_k exp(-b_k t_i),  t_i = (i + 0.5) / M_total,  x = [a_0..a_{K-1}, b_0..b_{K-1}]
//   J_{i,k} = exp(-b_k t_i),  J_{i,K+k} = -a_k t_i exp(-b_k t_i)      (a dense M x n Jacobian)
template <int WHAT>
__global__ void __launch_bounds__(256) expsum_dense_rows_kernel(ModelArgs a, const double* __restrict__ x, const double* __restrict__ yin,
                                                                double* __restrict__ out, double* __restrict__ partial) {
    extern __shared__ double2 sab[];  // [K] (a_k, b_k)
    __shared__ double shd[32];
    const int K = a.n / 2;
    for (int k = threadIdx.x; k < K; k += blockDim.x) sab[k] = make_double2(x[k], x[K + k]);
    __syncthreads();
    long long i0, i1, istep;
    if (WHAT == 1) {  // one row chunk per CTA (rowgeom.h)
        const int cg = blockIdx.x / a.geo.G, cb = blockIdx.x % a.geo.G;
        i0 = a.geo.local_begin(cg, cb) + threadIdx.x;
        i1 = a.geo.local_end(cg, cb);
        istep = blockDim.x;
    } else {
        i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        i1 = a.M;
        istep = (long long)gridDim.x * blockDim.x;
    }
    double ss = 0.0;
    for (long long i = i0; i < i1; i += istep) {
        const long long gi = a.row0 + i;
        const double t = ((double)gi + 0.5) / (double)a.M_total;
        double m = 0.0;
        for (int k = 0; k < K; ++k) m = fma(sab[k].x, exp(-sab[k].y * t), m);
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

__global__ void __launch_bounds__(kWarpsPerCta * 32) expsum_dense_jac_kernel(ModelArgs a, const double* __restrict__ x,
                                                                             double* __restrict__ J) {
    const int K = a.n / 2;
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * kWarpsPerCta;
    for (long long i = gw; i < a.M; i += nw) {
        const long long gi = a.row0 + i;
        const double t = ((double)gi + 0.5) / (double)a.M_total;
        double* Jrow = J + (size_t)i * a.ld;
        for (int j = lane; j < a.ld; j += 32) {
            double v = 0.0;  // padding columns
            if (j < K)
                v = exp(-x[K + j] * t);
            else if (j < a.n)
                v = -(x[j - K] * t) * exp(-x[j] * t);
            Jrow[j] = v;
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

static cudaError_t glm_rows_smem(size_t bytes) {  // (x_j, cs_j) pairs in shared memory: above 48 KB (n > 3072) needs the opt-in
    static size_t granted = 48 * 1024;
    if (bytes <= granted) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(glm_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(glm_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) granted = bytes;
    return e;
}

cudaError_t model_setup_y(const ModelArgs& a, const double* x_true, double* y, cudaStream_t st) {
    if (a.model_id == 1) {
        cudaError_t e = glm_rows_smem((size_t)a.ld * sizeof(double2));
        if (e != cudaSuccess) return e;
        glm_rows_kernel<0><<<rows_grid(a.M / 32 + 1, 148 * 8), 256, (size_t)a.ld * sizeof(double2), st>>>(a, x_true, nullptr, y, nullptr);
    } else if (a.model_id == 2) {
        expsum_rows_kernel<0><<<rows_grid(a.M / 32 + 1, 148 * 8), 256, 0, st>>>(a, x_true, nullptr, y, nullptr);
    } else if (a.model_id == 3) {
        expsum_dense_rows_kernel<0><<<rows_grid(a.M / 32 + 1, 148 * 8), 256, (size_t)(a.n / 2) * sizeof(double2), st>>>(a, x_true, nullptr, y, nullptr);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t model_residual(const ModelArgs& a, const double* x, const double* y, double* r, double* partial, cudaStream_t st) {
    const int grid = a.geo.ng * a.geo.G;  // one CTA per row chunk; partial[grid] goes through the fixed reduction tree (p2p.h)
    if (a.model_id == 1) {
        cudaError_t e = glm_rows_smem((size_t)a.ld * sizeof(double2));
        if (e != cudaSuccess) return e;
        glm_rows_kernel<1><<<grid, 256, (size_t)a.ld * sizeof(double2), st>>>(a, x, y, r, partial);
    } else if (a.model_id == 2) {
        expsum_rows_kernel<1><<<grid, 256, 0, st>>>(a, x, y, r, partial);
    } else if (a.model_id == 3) {
        expsum_dense_rows_kernel<1><<<grid, 256, (size_t)(a.n / 2) * sizeof(double2), st>>>(a, x, y, r, partial);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t model_jacobian(const ModelArgs& a, const double* x, double* J, cudaStream_t st) {
    if (a.model_id == 1) {
        glm_jac_kernel<<<rows_grid(a.M, 148 * 8), kWarpsPerCta * 32, 0, st>>>(a, x, J);
    } else if (a.model_id == 2) {
        expsum_jac_kernel<<<rows_grid(a.M, 148 * 8), kWarpsPerCta * 32, 0, st>>>(a, x, J);
    } else if (a.model_id == 3) {
        expsum_dense_jac_kernel<<<rows_grid(a.M, 148 * 8), kWarpsPerCta * 32, 0, st>>>(a, x, J);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// jac_res(x) with the J'r partials of the row-chunk geometry (GLM, ld <= 1024): partial[ng][G][8][pstride], columns [0, ld)
cudaError_t model_jacobian_jtr(const ModelArgs& a, const double* x, const double* r, double* J, double* partial, long long pstride,
                               int KCH, int RB, cudaStream_t st) {
    if (a.model_id != 1) return cudaErrorInvalidValue;
    const int grid = a.geo.ng * a.geo.G;
    switch (KCH) {
        case 1: glm_jac_jtr_kernel<1><<<grid, kWarpsPerCta * 32, 0, st>>>(a, x, r, J, partial, pstride, RB); break;
        case 2: glm_jac_jtr_kernel<2><<<grid, kWarpsPerCta * 32, 0, st>>>(a, x, r, J, partial, pstride, RB); break;
        case 4: glm_jac_jtr_kernel<4><<<grid, kWarpsPerCta * 32, 0, st>>>(a, x, r, J, partial, pstride, RB); break;
        case 8: glm_jac_jtr_kernel<8><<<grid, kWarpsPerCta * 32, 0, st>>>(a, x, r, J, partial, pstride, RB); break;
        case 16: glm_jac_jtr_kernel<16><<<grid, kWarpsPerCta * 32, 0, st>>>(a, x, r, J, partial, pstride, RB); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace bnl
