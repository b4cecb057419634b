// gram.cu -- G = J'J as a dense FP64 contraction on the tensor cores (DMMA: mma.sync.m8n8k4.f64; there is no
// FP64 tcgen05/wgmma kind), and the Gram-apply kernels H*v = G v, v'Hv = v'Gv.
// Not in the reference (its Hessian is matrix-free, src/basic_tralcnlss.jl:6-10; J'J is formed only in
// test/structures.jl:11): an opt-in mode that replaces the ~2 J passes of EVERY Hessian apply by one Gram formation
// per Jacobian (2*M*n^2 flops on the FP64 tensor pipe) + n x n gemvs that live in L2.
//
// Tiling: CTA = 128x128 output tile (upper-triangular tile pairs only) x one split-K slice of rows; 8 warps,
// warp tile 64x32 = 8x4 m8n8k4 tiles (64 FP64 accumulators / thread).  J rows are staged by a 3-stage cp.async
// (LDGSTS) ring of 32-row chunks; smem row stride 132 doubles => conflict-free 64-bit operand loads.
// Split-K partials are summed in fixed order (deterministic), the lower triangle is mirrored.
#include "common.cuh"
#include "gram.h"

namespace bnl {
namespace {

constexpr int TB = 128;       // output tile edge
constexpr int KT = 32;        // rows per smem chunk
constexpr int LDS_ = TB + 4;  // padded smem row stride (doubles)
constexpr int NSTG = 3;       // cp.async stages
constexpr int kGramSmem = NSTG * 2 * KT * LDS_ * (int)sizeof(double);

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(256, 1) gram_tile_kernel(const double* __restrict__ J, long long M, int ld, int ntile,
                                                          int nsplit, double* __restrict__ ws) {
    extern __shared__ __align__(16) double gsm[];
    double* sA = gsm;                          // [NSTG][KT][LDS_]
    double* sB = gsm + NSTG * KT * LDS_;       // [NSTG][KT][LDS_]
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
    const long long nchunk = (r_end > r_begin) ? (r_end - r_begin + KT - 1) / KT : 0;

    auto load_chunk = [&](long long ch, int stg) {
        // KT rows x 128 cols of both column blocks, 16-byte cp.async each; out-of-range => zero fill (src bytes 0)
        const long long r0 = r_begin + ch * KT;
#pragma unroll
        for (int e = tid; e < KT * (TB / 2); e += 256) {
            const int k = e / (TB / 2), c2 = e % (TB / 2);
            const long long r = r0 + k;
            const bool rok = r < r_end;
            const bool aok = rok && (a0 + 2 * c2 < ld);
            const bool bok = rok && (b0 + 2 * c2 < ld);
            const double* ga = J + (aok ? ((size_t)r * ld + a0 + 2 * c2) : 0);
            const double* gb = J + (bok ? ((size_t)r * ld + b0 + 2 * c2) : 0);
            cp_async16(&sA[(stg * KT + k) * LDS_ + 2 * c2], ga, aok ? 16 : 0);
            cp_async16(&sB[(stg * KT + k) * LDS_ + 2 * c2], gb, bok ? 16 : 0);
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int s = 0; s < NSTG - 1; ++s) {
        if (s < nchunk) load_chunk(s, s);
        cp_async_commit();
    }
    for (long long ch = 0; ch < nchunk; ++ch) {
        cp_async_wait<NSTG - 2>();
        __syncthreads();  // chunk ch landed for everyone; everyone finished reading chunk ch-1's buffer
        if (ch + NSTG - 1 < nchunk) load_chunk(ch + NSTG - 1, (int)((ch + NSTG - 1) % NSTG));
        cp_async_commit();
        const double* cA = sA + (size_t)(ch % NSTG) * KT * LDS_;
        const double* cB = sB + (size_t)(ch % NSTG) * KT * LDS_;
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = cA[(kk + lk) * LDS_ + wm + i * 8 + lm];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = cB[(kk + lk) * LDS_ + wn + j * 8 + lm];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
    double* out = ws + (size_t)split * ld * ld;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ga = a0 + wm + i * 8 + lm;
            const int gb = b0 + wn + j * 8 + 2 * lk;
            if (ga < ld && gb + 1 < ld) {
                *reinterpret_cast<double2*>(&out[(size_t)ga * ld + gb]) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else if (ga < ld && gb < ld) {
                out[(size_t)ga * ld + gb] = acc[i][j][0];
            }
        }
}

// G[a][b] = sum_split ws[split][min-tile][max-tile] in fixed order; mirrors the upper-triangular tiles.
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

// y = G v (one warp per row, fixed order) -- G is ld x ld row-major symmetric, lives in L2 between applies
__global__ void __launch_bounds__(256) gram_gemv_kernel(const double* __restrict__ G, int n, int ld, const double* __restrict__ v,
                                                        double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const double2* g2 = reinterpret_cast<const double2*>(G + (size_t)row * ld);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    double s0 = 0.0, s1 = 0.0;
    for (int c = lane; c < (ld >> 1); c += 32) {
        const double2 a = g2[c], b = v2[c];
        s0 = fma(a.x, b.x, s0);
        s1 = fma(a.y, b.y, s1);
    }
    double s = warp_sum(s0 + s1);
    if (lane == 0) y[row] = s;
}
// slot = v . y  (single CTA, fixed order) -- dot(Jv,Jv) = v'Gv in Gram mode
__global__ void dot_to_slot_kernel(const double* __restrict__ v, const double* __restrict__ y, int n, double* __restrict__ slot) {
    __shared__ double shd[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a = fma(v[i], y[i], a);
    a = block_sum(a, shd);
    if (threadIdx.x == 0) *slot = a;
}

}  // namespace

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

int gram_pick_split(long long M, int ld, int sm_count) {
    const int ntile = (ld + TB - 1) / TB;
    const int npairs = ntile * (ntile + 1) / 2;
    // make npairs * nsplit a multiple of the SM count (whole waves) when that is cheap, else ~2 waves
    int nsplit = sm_count / gcd_i(npairs, sm_count);
    const size_t ws_cap = (size_t)2 << 30;  // keep the split-K workspace under 2 GiB
    while (nsplit > 1 && (size_t)nsplit * ld * ld * sizeof(double) > ws_cap) nsplit = (nsplit + 1) / 2;
    if (nsplit > 64) nsplit = 64;
    long long maxsplit = (M + KT - 1) / KT;
    if (nsplit > maxsplit) nsplit = (int)maxsplit;
    if (nsplit < 1) nsplit = 1;
    return nsplit;
}

cudaError_t gram_launch(const double* J, long long M, int ld, double* G, double* workspace, int nsplit, cudaStream_t st) {
    const int ntile = (ld + TB - 1) / TB;
    const int npairs = ntile * (ntile + 1) / 2;
    cudaError_t e = cudaFuncSetAttribute(gram_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmem);
    if (e != cudaSuccess) return e;
    // tiles of the strict lower triangle are never written: the reduce kernel only reads upper tiles
    dim3 grid(npairs, nsplit);
    gram_tile_kernel<<<grid, 256, kGramSmem, st>>>(J, M, ld, ntile, nsplit, workspace);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    gram_reduce_kernel<<<296, 256, 0, st>>>(workspace, ld, nsplit, G);
    return cudaGetLastError();
}

cudaError_t gram_apply(const double* G, int n, int ld, const double* v, double* y, cudaStream_t st) {
    gram_gemv_kernel<<<(n + 7) / 8, 256, 0, st>>>(G, n, ld, v, y);
    dot_to_slot_kernel<<<1, 1024, 0, st>>>(v, y, n, y + ld);
    return cudaGetLastError();
}

cudaError_t gram_gemv(const double* G, int n, int ld, const double* v, double* y, cudaStream_t st) {
    gram_gemv_kernel<<<(n + 7) / 8, 256, 0, st>>>(G, n, ld, v, y);
    return cudaGetLastError();
}

double gram_flops(long long M, int ld) {
    const int ntile = (ld + TB - 1) / TB;
    const double npairs = ntile * (ntile + 1) / 2.0;
    return 2.0 * (double)M * npairs * TB * TB;  // flops actually issued (upper-triangular 128-tiles)
}

}  // namespace bnl
