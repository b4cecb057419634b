// common.cuh -- shared device helpers for libbenlsip_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbenlsip_b200 is written for sm_100a (B200) only"
#endif

namespace bnl {

constexpr int kColAlign = 16;  // row stride of J and length of every n-vector is padded to 16 doubles (128 B)

__host__ __device__ inline int pad_cols(int n) { return (n + kColAlign - 1) / kColAlign * kColAlign; }

// ---- Scalars shared between the O(n) kernels and the host control flow -------------------------------
// One copy lives in device memory (read by kernels), one in pinned mapped host memory (read by the host
// after a stream sync).  Kernels write both.
struct Scal {
    // cauchy_step
    double phi_p, phi_pp, theta;
    long long bp_ind;
    double bp_dind;            // d[bp_ind] at the time of the scan (incremental Cauchy mode)
    // projected_cg
    double pHp, rtv, alpha, gamma, beta, tol_cg;
    int cg_neg_curv, cg_outside, cg_solved, cg_iter;
    // inner_step
    double nrg_g, nrg_gm;      // ||P(-g)||, ||P(-g_minor)||
    int n_at_bound;            // |active_bounds(...)|
    int nb_fix;                // count(fixvars)
    double gs;                 // dot(g, s)
    double alpha_ls;           // linesearch result
    double wHw, gw;
    // solve_subproblem
    double norm_g, norm_s, pix;
    double sumsq_r;            // dot(rx,rx) (global, after all-reduce)
    double jv_sumsq;           // dot(Jv,Jv) (global)
    double Cv_sumsq;           // dot(Cv,Cv)
    double c0;                 // built-in nonlinear constraint value c(x) (p = 1)
    int chol_fail;             // device Cholesky hit a non-positive pivot
    int p2p_timeout;           // peer-memory all-reduce gave up waiting for a rank
    // persistent Cauchy breakpoint loop (cauchy_loop.cu)
    int cl_status, cl_breakpoints;
    long long cl_rounds;
    double phi_a, phi_b;       // the two terms of phi' = dot(s_c,Hd) + dot(g,d) (the guard band is relative to |a| + |b|)
};

// ---- block-wide deterministic reductions (fixed tree => run-to-run bit-identical) --------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole CTA; result returned to every thread.  `sh` needs >= 32 doubles. blockDim multiple of 32.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sh from a previous use
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += sh[i];  // fixed order, every thread computes the same value
    return t;
}

// min with lowest-index tie-break (reference: strict `<` scan, src/basic_tralcnlss.jl:555)
__device__ __forceinline__ void argmin_combine(double& v, long long& i, double v2, long long i2) {
    if (v2 < v || (v2 == v && i2 >= 0 && (i < 0 || i2 < i))) {
        v = v2;
        i = i2;
    }
}

__device__ __forceinline__ void block_argmin(double& v, long long& idx, double* shv, long long* shi) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        long long i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        argmin_combine(v, idx, v2, i2);
    }
    __syncthreads();
    if (lane == 0) {
        shv[w] = v;
        shi[w] = idx;
    }
    __syncthreads();
    double bv = shv[0];
    long long bi = shi[0];
    for (int i = 1; i < nw; ++i) argmin_combine(bv, bi, shv[i], shi[i]);
    v = bv;
    idx = bi;
}

__device__ __forceinline__ double block_min(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = sh[0];
    for (int i = 1; i < nw; ++i) t = fmin(t, sh[i]);
    return t;
}

__device__ __forceinline__ int block_sum_int(int v, int* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int t = 0;
    for (int i = 0; i < nw; ++i) t += sh[i];
    return t;
}

// ---- counter-based hash shared with oracle/models.py -------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t rowkey(uint32_t seed, unsigned long long i) {
    return mix32((uint32_t)i ^ mix32(seed));
}
__host__ __device__ __forceinline__ uint32_t hash_rc(uint32_t rk, uint32_t j) { return mix32(rk + j * 0x9E3779B9u); }
__host__ __device__ __forceinline__ double u01(uint32_t h) { return (double)h * 2.3283064365386963e-10; }          // 2^-32
__host__ __device__ __forceinline__ double usym(uint32_t h) { return (double)h * 4.6566128730773926e-10 - 1.0; }   // 2^-31
#ifdef __CUDACC__
// usym(h) without an integer -> FP64 conversion: the double with exponent 0 and mantissa h<<20 is v = 1 + h*2^-32 exactly, and
// 2 v - 3 = h*2^-31 - 1 exactly (32 fractional bits): bit-identical to usym(), one FMA + two integer ops.
__device__ __forceinline__ double usym_fast(uint32_t h) {
    const double v = __hiloint2double((int)(0x3FF00000u | (h >> 12)), (int)(h << 20));
    return fma(2.0, v, -3.0);
}
#endif

}  // namespace bnl
