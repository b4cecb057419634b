// p2p.h -- the path's only collective: combining the per-group sums of the row reductions (rowgeom.h).
//
// Every rank owns a "mailbox" [2 parities][kGroups][kP2PWidth] doubles + flags in its own HBM.  With N > 1 ranks the
// mailboxes are mapped into every peer with CUDA IPC and a rank stores the sums of ITS groups straight into the slot
// [parity][group] of EVERY rank's mailbox (st.global over NVLink), fences, then raises flag[parity][my_rank] = epoch on every
// peer.  "wait+sum": a rank spins on its own flags until all ranks have raised them, then adds the kGroups rows in group
// order.  Every rank adds the same 8 vectors in the same order, and with N = 1 the same 8 vectors are produced locally, so
// the result is bit-identical across ranks AND across N = 1, 2, 4, 8.  The push is fused into the kernel that reduces the
// per-chunk partials (group_reduce).  When peers cannot be mapped, the group sums travel by ncclAllGather instead (same sums,
// same order).
//
// The incremental Cauchy loop (cauchy_loop.cu) exchanges 2 doubles per group and breakpoint; that is pure latency, so it uses
// an "LL" mailbox: every 8-byte word carries 4 bytes of data + a 4-byte epoch tag, stored with single 8-byte stores (atomic),
// and the receiver polls the words themselves -- no fence, no separate flag: one NVLink store latency per exchange.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rowgeom.h"

namespace bnl {

constexpr int kP2PMaxRanks = 8;
constexpr int kP2PWidth = 8192 + 16;  // doubles per mailbox row (ld <= 8192, + the ||Jv||^2 slot)
constexpr int kLLVals = 4;            // doubles per group in the LL mailbox

enum { GR_LOCAL = 0, GR_PUSH = 1 };

struct P2PArgs {
    int nranks, rank;
    double* mbox[kP2PMaxRanks];              // mailbox base of every rank (peer-mapped; [rank] is local)
    unsigned long long* flag[kP2PMaxRanks];  // flags base of every rank: [2][kP2PMaxRanks]
    unsigned long long* ll[kP2PMaxRanks];    // LL mailbox base of every rank: [2][kGroups][kLLVals][2] words
    unsigned int* done_counter;              // local: last-CTA detection of the push kernel
    int* timeout_flag;                       // local: set if a wait gave up
};

inline size_t p2p_mbox_doubles() { return (size_t)2 * kGroups * kP2PWidth; }
inline size_t p2p_buffer_bytes() {
    return p2p_mbox_doubles() * sizeof(double) + (size_t)2 * kP2PMaxRanks * sizeof(unsigned long long) +
           (size_t)2 * kGroups * kLLVals * 2 * sizeof(unsigned long long) + 256;
}
inline size_t p2p_alloc_bytes() { return p2p_buffer_bytes() > ((size_t)4 << 20) ? p2p_buffer_bytes() : ((size_t)4 << 20); }
inline unsigned long long* p2p_flags_of(double* mbox_base) { return reinterpret_cast<unsigned long long*>(mbox_base + p2p_mbox_doubles()); }
inline unsigned long long* p2p_ll_of(double* mbox_base) { return p2p_flags_of(mbox_base) + 2 * kP2PMaxRanks; }

// Partials P[ng][G][T][pstride] -> group sums (fixed tree, rowgeom.h) -> mailbox slot of each local group (pushed to every
// peer when mode == GR_PUSH).  Columns [col0, ncols).
cudaError_t group_reduce(const double* P, int G, int T, long long pstride, int col0, int ncols, int g0, int ng,
                         const P2PArgs& a, unsigned long long epoch, int mode, cudaStream_t st);
// out[col0..ncols) = sum over the kGroups mailbox rows in group order (after waiting for every rank's flag when wait != 0)
cudaError_t group_sum(const P2PArgs& a, unsigned long long epoch, double* out, int col0, int ncols, int wait, cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ size_t p2p_slot(unsigned long long epoch, int group, int col) {
    return ((size_t)(epoch & 1ull) * kGroups + group) * kP2PWidth + col;
}
// call by ALL threads of ALL CTAs of the grid after their stores; the last CTA of the grid raises the flags
__device__ __forceinline__ void p2p_push_finish(const P2PArgs& a, unsigned long long epoch) {
    __shared__ unsigned int s_ticket;
    const unsigned int nthreads = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(a.done_counter, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x * gridDim.y - 1) {
        if (tid == 0) *a.done_counter = 0u;
        __threadfence_system();
        if (tid < (unsigned)a.nranks && tid < nthreads) {
            unsigned long long* f = a.flag[tid] + (size_t)(epoch & 1ull) * kP2PMaxRanks + a.rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        }
    }
}

// ---- LL mailbox (device helpers, used by the controller CTA of the Cauchy loop) ----
__device__ __forceinline__ size_t ll_word(unsigned long long epoch, int group, int val, int half) {
    return (((size_t)(epoch & 1ull) * kGroups + group) * kLLVals + val) * 2 + half;
}
__device__ __forceinline__ void ll_store(const P2PArgs& a, int peer, unsigned long long epoch, int group, int val, double x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const unsigned long long tag = (epoch & 0xffffffffull) << 32;
    unsigned long long* base = a.ll[peer];
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(base + ll_word(epoch, group, val, 0)), "l"(w0) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(base + ll_word(epoch, group, val, 1)), "l"(w1) : "memory");
}
// polls the local LL mailbox until both words of (group, val) carry this epoch's tag; false on timeout (~20 s)
__device__ __forceinline__ bool ll_load(const P2PArgs& a, unsigned long long epoch, int group, int val, double* out) {
    const unsigned long long* base = a.ll[a.rank];
    const unsigned long long want = epoch & 0xffffffffull;
    unsigned long long w0, w1;
    const long long t0 = clock64();
    while (true) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(base + ll_word(epoch, group, val, 0)) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(base + ll_word(epoch, group, val, 1)) : "memory");
        if ((w0 >> 32) == want && (w1 >> 32) == want) break;
        if (clock64() - t0 > 40000000000ll) return false;
    }
    *out = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    return true;
}
#endif

}  // namespace bnl
