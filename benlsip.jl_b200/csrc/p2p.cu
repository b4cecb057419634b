// p2p.cu -- see p2p.h and rowgeom.h: the fixed reduction tree over (team, chunk, group) and its NVLink exchange.
#include "p2p.h"

namespace bnl {
namespace {

// One CTA = 32 columns x kChains chains of one local group.  Thread (x, k) sums the chunks b = k, k + kChains, ... of its
// column (teams of a chunk first, in order), the kChains chain sums are added in order by the k == 0 thread.
template <int T>
__global__ void __launch_bounds__(32 * kChains) group_reduce_kernel(const double* __restrict__ P, int G, long long pstride,
                                                                    int col0, int ncols, int g0, P2PArgs a,
                                                                    unsigned long long epoch, int mode) {
    __shared__ double sh[kChains][33];
    const int x = threadIdx.x, k = threadIdx.y;
    const int j = col0 + blockIdx.x * 32 + x;
    const int gi = blockIdx.y;
    double s = 0.0;
    if (j < ncols) {
        const double* base = P + ((size_t)gi * G) * T * pstride + j;
        for (int b = k; b < G; b += kChains) {
            const double* pb = base + (size_t)b * T * pstride;
            double v[T];
#pragma unroll
            for (int t = 0; t < T; ++t) v[t] = __ldcg(pb + (size_t)t * pstride);
            double cs = v[0];
#pragma unroll
            for (int t = 1; t < T; ++t) cs += v[t];
            s += cs;
        }
    }
    sh[k][x] = s;
    __syncthreads();
    if (k == 0 && j < ncols) {
        double tot = sh[0][x];
#pragma unroll
        for (int kk = 1; kk < kChains; ++kk) tot += sh[kk][x];
        const size_t off = p2p_slot(epoch, g0 + gi, j);
        if (mode == GR_PUSH) {
            for (int r = 0; r < a.nranks; ++r) a.mbox[r][off] = tot;
        } else {
            a.mbox[a.rank][off] = tot;
        }
    }
    if (mode == GR_PUSH) p2p_push_finish(a, epoch);
}

__global__ void group_sum_kernel(P2PArgs a, unsigned long long epoch, double* __restrict__ out, int col0, int ncols, int wait) {
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    if (wait && threadIdx.x < a.nranks) {
        const unsigned long long* f = a.flag[a.rank] + (size_t)(epoch & 1ull) * kP2PMaxRanks + threadIdx.x;
        const long long t0 = clock64();
        unsigned long long v;
        while (true) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer died or the call sequence diverged
                s_fail = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (s_fail && threadIdx.x == 0) *a.timeout_flag = 1;
    const double* mb = a.mbox[a.rank] + p2p_slot(epoch, 0, 0);
    for (int j = col0 + threadIdx.x; j < ncols; j += blockDim.x) {
        double s = __ldcg(mb + j);  // group 0, then groups 1..7 in order (L2 loads: no stale L1 lines of peer-written data)
#pragma unroll
        for (int g = 1; g < kGroups; ++g) s += __ldcg(mb + (size_t)g * kP2PWidth + j);
        out[j] = s;
    }
}

}  // namespace

cudaError_t group_reduce(const double* P, int G, int T, long long pstride, int col0, int ncols, int g0, int ng,
                         const P2PArgs& a, unsigned long long epoch, int mode, cudaStream_t st) {
    if (ncols > kP2PWidth || ncols <= col0 || ng < 1) return cudaErrorInvalidValue;
    dim3 grid((ncols - col0 + 31) / 32, ng), block(32, kChains);
    if (T == 1)
        group_reduce_kernel<1><<<grid, block, 0, st>>>(P, G, pstride, col0, ncols, g0, a, epoch, mode);
    else if (T == 8)
        group_reduce_kernel<8><<<grid, block, 0, st>>>(P, G, pstride, col0, ncols, g0, a, epoch, mode);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t group_sum(const P2PArgs& a, unsigned long long epoch, double* out, int col0, int ncols, int wait, cudaStream_t st) {
    group_sum_kernel<<<1, 1024, 0, st>>>(a, epoch, out, col0, ncols, wait);
    return cudaGetLastError();
}

}  // namespace bnl
