// gram.cu -- G = J'J as a dense FP64 contraction on the tensor cores (DMMA: mma.sync.m8n8k4.f64; there is no
// FP64 tcgen05/wgmma kind).  Not in the reference (its Hessian is matrix-free, src/basic_tralcnlss.jl:6-10;
// J'J is formed only in test/structures.jl:11): an opt-in capability for preconditioning / Gram-apply modes.
//
// Tiling: CTA = 128x128 output tile (upper-triangular tile pairs only) x one split-K slice of rows; 8 warps,
// warp tile 64x32 = 8x4 m8n8k4 tiles (64 FP64 accumulators / thread).  J rows are staged in shared memory
// (row stride 132 doubles => conflict-free 64-bit operand loads).  Split-K partials are summed in fixed order.
#include "common.cuh"
#include "gram.h"

namespace bnl {
namespace {

constexpr int TB = 128;       // output tile edge
constexpr int KT = 16;        // rows per smem chunk
constexpr int LDS_ = TB + 4;  // padded smem row stride (doubles)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) gram_tile_kernel(const double* __restrict__ J, long long M, int ld, int ntile,
                                                          int nsplit, double* __restrict__ ws) {
    __shared__ double sA[KT * LDS_];
    __shared__ double sB[KT * LDS_];
    // tile pair (ta <= tb) from the linear upper-triangular index
    int t = blockIdx.x, ta = 0;
    while (t >= ntile - ta) {
        t -= ntile - ta;
        ++ta;
    }
    const int tb = ta + t;
    const int split = blockIdx.y;
    const long long rows_per = ((M + nsplit - 1) / nsplit + KT - 1) / KT * KT;
    const long long r_begin = (long long)split * rows_per;
    long long r_end = r_begin + rows_per;
    if (r_end > M) r_end = M;
    const int a0 = ta * TB, b0 = tb * TB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 1) * 64;   // warp tile origin inside the CTA tile (a direction)
    const int wn = (warp >> 1) * 32;  // (b direction)
    const int lk = lane & 3, lm = lane >> 2;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (long long r0 = r_begin; r0 < r_end; r0 += KT) {
        __syncthreads();
        // stage KT rows x 128 cols of the two column blocks (double2, coalesced); rows past r_end are zero
        for (int e = tid; e < KT * (TB / 2); e += 256) {
            const int k = e / (TB / 2), c2 = e % (TB / 2);
            const long long r = r0 + k;
            double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
            if (r < r_end) {
                if (a0 + 2 * c2 < ld) va = *reinterpret_cast<const double2*>(J + (size_t)r * ld + a0 + 2 * c2);
                if (b0 + 2 * c2 < ld) vb = *reinterpret_cast<const double2*>(J + (size_t)r * ld + b0 + 2 * c2);
            }
            *reinterpret_cast<double2*>(&sA[k * LDS_ + 2 * c2]) = va;
            *reinterpret_cast<double2*>(&sB[k * LDS_ + 2 * c2]) = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = sA[(kk + lk) * LDS_ + wm + i * 8 + lm];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = sB[(kk + lk) * LDS_ + wn + j * 8 + lm];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
    double* out = ws + (size_t)split * ld * ld;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ga = a0 + wm + i * 8 + lm;
            const int gb = b0 + wn + j * 8 + 2 * lk;
            if (ga < ld && gb < ld) {
                out[(size_t)ga * ld + gb] = acc[i][j][0];
                if (gb + 1 < ld) out[(size_t)ga * ld + gb + 1] = acc[i][j][1];
            }
        }
}

// G[a][b] = sum_split ws[split][min][max] in fixed order; mirrors the upper-triangular tiles.
__global__ void gram_reduce_kernel(const double* __restrict__ ws, int ld, int nsplit, double* __restrict__ G) {
    const size_t tot = (size_t)ld * ld;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(e / ld), b = (int)(e % ld);
        const int ta = a / TB, tb = b / TB;
        const size_t src = (ta <= tb) ? ((size_t)a * ld + b) : ((size_t)b * ld + a);
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += ws[(size_t)k * tot + src];
        G[e] = s;
    }
}

}  // namespace

int gram_pick_split(long long M, int ld, int sm_count) {
    const int ntile = (ld + TB - 1) / TB;
    const int npairs = ntile * (ntile + 1) / 2;
    int nsplit = (2 * sm_count * 2 + npairs - 1) / npairs;  // >= 2 waves of 2 CTAs/SM
    long long maxsplit = (M + KT - 1) / KT;
    if (nsplit > maxsplit) nsplit = (int)maxsplit;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 64) nsplit = 64;
    return nsplit;
}

cudaError_t gram_launch(const double* J, long long M, int ld, double* G, double* workspace, int nsplit, cudaStream_t st) {
    const int ntile = (ld + TB - 1) / TB;
    const int npairs = ntile * (ntile + 1) / 2;
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)nsplit * ld * ld * sizeof(double), st);
    if (e != cudaSuccess) return e;
    dim3 grid(npairs, nsplit);
    gram_tile_kernel<<<grid, 256, 0, st>>>(J, M, ld, ntile, nsplit, workspace);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    gram_reduce_kernel<<<296, 256, 0, st>>>(workspace, ld, nsplit, G);
    return cudaGetLastError();
}

}  // namespace bnl
