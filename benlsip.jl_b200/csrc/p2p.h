// p2p.h -- one-shot all-reduce over NVLink peer memory for the path's only collective (n+1 doubles per Hessian
// apply: latency-bound).  Every rank owns a mailbox [2 parities][nranks][W] + flags in its own HBM, mapped
// into every peer with CUDA IPC.  "push": a rank stores its n+1 partial sums straight into the mailbox slot
// [parity][my_rank] of EVERY peer (st.global over NVLink), fences, then raises flag[parity][my_rank] = epoch on every
// peer.  "wait+sum": a rank spins on its own flags until all ranks have raised them, then adds the nranks mailbox rows
// in fixed rank order => the result is bit-identical on all ranks (the replicated control flow relies on that).
// The push is fused into the kernel that reduces the per-CTA partials of the streaming kernel (matvec.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bnl {

constexpr int kP2PMaxRanks = 16;
constexpr int kP2PWidth = 8192 + 16;  // doubles per mailbox row (ld <= 8192, + the ||Jv||^2 slot)

struct P2PArgs {
    int nranks, rank;
    double* mbox[kP2PMaxRanks];              // mailbox base of every rank (peer-mapped; [rank] is local)
    unsigned long long* flag[kP2PMaxRanks];  // flags base of every rank
    unsigned int* done_counter;              // local: last-CTA detection of the push kernel
    int* timeout_flag;                       // local: set if a wait gave up
};

inline size_t p2p_buffer_bytes(int nranks) {
    return (size_t)2 * nranks * kP2PWidth * sizeof(double) + (size_t)2 * kP2PMaxRanks * sizeof(unsigned long long) + 256;
}
inline unsigned long long* p2p_flags_of(double* mbox_base, int nranks) {
    return reinterpret_cast<unsigned long long*>(mbox_base + (size_t)2 * nranks * kP2PWidth);
}

// generic two-kernel form: buf[0..count) -> all-reduced in place
cudaError_t p2p_allreduce(const P2PArgs& a, unsigned long long epoch, double* buf, int count, cudaStream_t st);
// second half only (after a fused reduce+push): out[col0..ncols) = sum over ranks
cudaError_t p2p_wait_sum(const P2PArgs& a, unsigned long long epoch, double* out, int col0, int ncols, cudaStream_t st);

#ifdef __CUDACC__
// device helpers shared with matvec.cu
__device__ __forceinline__ void p2p_push_value(const P2PArgs& a, unsigned long long epoch, int col, double v) {
    const size_t off = ((size_t)(epoch & 1ull) * a.nranks + a.rank) * kP2PWidth + col;
    for (int r = 0; r < a.nranks; ++r) a.mbox[r][off] = v;
}
// call by ALL threads of the CTA after their p2p_push_value calls; the last CTA of the grid raises the flags
__device__ __forceinline__ void p2p_push_finish(const P2PArgs& a, unsigned long long epoch) {
    __shared__ unsigned int s_ticket;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.done_counter, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x - 1) {
        if (threadIdx.x == 0) *a.done_counter = 0u;
        __threadfence_system();
        if (threadIdx.x < a.nranks) {
            unsigned long long* f = a.flag[threadIdx.x] + (size_t)(epoch & 1ull) * kP2PMaxRanks + a.rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        }
    }
}
#endif

}  // namespace bnl
