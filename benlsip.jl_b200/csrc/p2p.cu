// p2p.cu -- see p2p.h.
#include "p2p.h"

namespace bnl {
namespace {

__global__ void p2p_push_kernel(P2PArgs a, unsigned long long epoch, const double* __restrict__ buf, int count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) p2p_push_value(a, epoch, j, buf[j]);
    p2p_push_finish(a, epoch);
}

__global__ void p2p_wait_sum_kernel(P2PArgs a, unsigned long long epoch, double* __restrict__ out, int col0, int ncols) {
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    if (threadIdx.x < a.nranks) {
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
    const double* mb = a.mbox[a.rank] + (size_t)(epoch & 1ull) * a.nranks * kP2PWidth;
    for (int j = col0 + threadIdx.x; j < ncols; j += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < a.nranks; ++r) s += __ldcg(mb + (size_t)r * kP2PWidth + j);  // fixed rank order, L2 (no stale L1)
        out[j] = s;
    }
}

}  // namespace

cudaError_t p2p_wait_sum(const P2PArgs& a, unsigned long long epoch, double* out, int col0, int ncols, cudaStream_t st) {
    p2p_wait_sum_kernel<<<1, 1024, 0, st>>>(a, epoch, out, col0, ncols);
    return cudaGetLastError();
}

cudaError_t p2p_allreduce(const P2PArgs& a, unsigned long long epoch, double* buf, int count, cudaStream_t st) {
    if (count > kP2PWidth) return cudaErrorInvalidValue;
    p2p_push_kernel<<<(count + 255) / 256, 256, 0, st>>>(a, epoch, buf, count);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return p2p_wait_sum(a, epoch, buf, 0, count, st);
}

}  // namespace bnl
