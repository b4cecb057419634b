// matvec.cu -- the HBM-bound kernels of the hot path: one templated streaming kernel over a row-major
// Jacobian panel ring, in three modes:
//
//   MODE_JTJV : out = J'(J v)  in ONE pass over J (+ sum_i (Jv)_i^2)   replaces  Base.:*(H,v)
//               src/basic_tralcnlss.jl:102-106 (two DGEMVs = two passes in the reference)
//   MODE_JV   : t = J v (optional store) and sum_i t_i^2                replaces  vthv  :92-96, H.J*v :93,:103
//   MODE_JTW  : out = J' w                                             replaces  Jx'*rx :45,:74,:893
//
// Data layout: J is ROW-major in HBM, M_loc x ld doubles, ld = n padded to 16 (zero padding), so a panel
// of R consecutive rows is ONE contiguous R*ld*8-byte range: a single 1-D TMA bulk copy
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes -> SASS UBLKCP) brings it into a
// shared-memory ring of NS stages guarded by "full" mbarriers (no dedicated producer warp: team leaders
// re-arm their own slots).
//
// Consumers: thread-owns-columns.  A "team" of TG threads shares one row (TG = 32 for ld <= 1024: eight
// independent warp teams per SM; TG = 256 above); thread u of the team owns the double2 column chunks
// u + k*TG (k < KCH), keeps v and the J' accumulator for them in registers, reads its part of RB rows from
// smem with conflict-free LDS.128, then the slot is released at once (the rows live in registers from here
// on).  The RB row dot products are reduced with a transposing shuffle butterfly (warp teams: nothing else;
// 256-thread team: + one named barrier); J'.t is accumulated from the same registers.  Per CTA the n column
// sums (and sum t^2) go to partial[cta][*]; a second tiny kernel sums the partials in fixed CTA order =>
// deterministic, no FP64 atomics.
//
// Algorithmic bytes per launch (what roofline.achieved uses): 8*M_loc*ld (+ O(n)); J is read exactly once.
#include "common.cuh"
#include "matvec.h"
#include "p2p.h"

namespace bnl {

namespace {

constexpr int kThreads = 256;  // 8 warps, all consumers; team leaders double as TMA issuers
constexpr int kMaxRB = 4;                  // rows per reduction batch (template RB = 4 or 2)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16 B aligned).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                             uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int MODE, int KCH, int RB, bool WARP_TEAM>
__global__ void __launch_bounds__(kThreads, 1) mv_stream_kernel(const MvArgs a) {
    // A "team" is the set of TG threads that shares one row; every team streams its own sequence of stages
    // (RB rows each) through its own private ring of NSt slots, so T = 256/TG independent
    // load -> dot -> reduce -> accumulate chains are in flight per SM and no thread is a dedicated producer:
    // the team leader re-arms a slot (mbarrier expect_tx + TMA bulk copy of the stage NSt ahead) as soon as
    // the team has pulled the slot into registers.
    //   WARP_TEAM: TG = 32 (ld <= 1024): the row dot product is a pure shuffle butterfly, no barrier at all.
    //   else     : TG = 256 (ld > 1024): one team, cross-warp reduce through smem + one named barrier.
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ld = a.ld;
    const int NC = ld >> 1;  // double2 chunks per row
    constexpr int TG = WARP_TEAM ? 32 : kThreads;
    constexpr int T = kThreads / TG;  // teams
    const int NSt = a.NS / T;         // slots per team
    const size_t stage_doubles = (size_t)RB * ld;
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* red = stages + (size_t)a.NS * stage_doubles;                     // [2][kMaxRB][8]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 2 * kMaxRB * 8);  // [NS]

    const int tid = threadIdx.x;
    const long long total_stages = (a.M + RB - 1) / RB;

    if (tid == 0) {
        for (int s = 0; s < a.NS; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int team = tid / TG;
    const int u = tid - team * TG;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    double* my_stages = stages + (size_t)team * NSt * stage_doubles;
    uint64_t* my_full = full_bar + team * NSt;
    // global stage index of this team's j-th stage: st(j) = blockIdx.x + (team + j*T) * gridDim.x
    const long long st0 = blockIdx.x + (long long)team * gridDim.x;
    const long long st_step = (long long)T * gridDim.x;
    uint64_t pol = 0;

    auto issue = [&](long long j) {  // leader only: arm slot j % NSt with stage st(j)
        const long long st = st0 + j * st_step;
        if (st < total_stages) {
            const int slot = (int)(j % NSt);
            const long long row0 = st * RB;
            const long long rows = (a.M - row0 < RB) ? (a.M - row0) : RB;
            const uint32_t bytes = (uint32_t)(rows * ld * sizeof(double));
            mbar_arrive_expect_tx(&my_full[slot], bytes);
            tma_bulk_g2s(my_stages + (size_t)slot * stage_doubles, a.J + row0 * ld, bytes, &my_full[slot], pol);
        }
    };
    if (u == 0) {
        pol = policy_evict_first();
        for (int j = 0; j < NSt; ++j) issue(j);
    }

    double2 vv[KCH];
    double2 acc[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int c = u + k * TG;
        acc[k] = make_double2(0.0, 0.0);
        if (MODE != MODE_JTW)
            vv[k] = (c < NC) ? reinterpret_cast<const double2*>(a.v)[c] : make_double2(0.0, 0.0);
        else
            vv[k] = make_double2(0.0, 0.0);
    }
    double tsq = 0.0;
    int batch_parity = 0;

    long long j = 0;
    for (long long st = st0; st < total_stages; st += st_step, ++j) {
        const int slot = (int)(j % NSt);
        const uint32_t use = (uint32_t)(j / NSt);
        const long long row0 = st * RB;
        const int rows_valid = (int)((a.M - row0 < RB) ? (a.M - row0) : RB);
        mbar_wait(&my_full[slot], use & 1u);
        const double2* sbase = reinterpret_cast<const double2*>(my_stages + (size_t)slot * stage_doubles);

        double2 jr[RB][KCH];
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const bool valid = q < rows_valid;
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const int c = u + k * TG;
                jr[q][k] = (valid && c < NC) ? sbase[(size_t)q * NC + c] : make_double2(0.0, 0.0);
            }
        }

        double t[RB];
        if (MODE == MODE_JTW) {
#pragma unroll
            for (int q = 0; q < RB; ++q) t[q] = (q < rows_valid) ? __ldg(a.w + row0 + q) : 0.0;
            // every thread of the team has issued its reads of the slot (in-order issue): re-arm it
            if (WARP_TEAM)
                __syncwarp();
            else
                named_bar_sync(1, kThreads);
            if (u == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(j + NSt);
            }
        } else {
            double part[RB];
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // 4 independent FMA chains
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    if (k & 1) {
                        s2 = fma(jr[q][k].x, vv[k].x, s2);
                        s3 = fma(jr[q][k].y, vv[k].y, s3);
                    } else {
                        s0 = fma(jr[q][k].x, vv[k].x, s0);
                        s1 = fma(jr[q][k].y, vv[k].y, s1);
                    }
                }
                part[q] = (s0 + s1) + (s2 + s3);
            }
            if constexpr (WARP_TEAM) {
                // Re-arm the slot as early as provably safe: part[] depends on every J value this lane loaded, the warp
                // issues in order, and an LDS completes for all lanes at once => once the FMAs producing part[] have
                // issued, every read of the slot has completed.  The empty asm pins the re-arm below those FMAs.
                double dep = part[0];
#pragma unroll
                for (int q = 1; q < RB; ++q) dep += part[q];
                asm volatile("" ::"d"(dep) : "memory");
                if (u == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(j + NSt);
                }
            }
            // transposing butterfly: RB values x 32 lanes -> a lane holds one row's warp sum
            double kx;
            if constexpr (RB == 4) {  // 6 shuffles; lane holds row (2*bit4 + bit3)
                const bool hi16 = (lane & 16) != 0;
                double x0 = hi16 ? part[0] : part[2];
                double x1 = hi16 ? part[1] : part[3];
                double k0 = hi16 ? part[2] : part[0];
                double k1 = hi16 ? part[3] : part[1];
                k0 += __shfl_xor_sync(0xffffffffu, x0, 16);
                k1 += __shfl_xor_sync(0xffffffffu, x1, 16);
                const bool hi8 = (lane & 8) != 0;
                double sx = hi8 ? k0 : k1;
                kx = hi8 ? k1 : k0;
                kx += __shfl_xor_sync(0xffffffffu, sx, 8);
            } else if constexpr (RB == 2) {  // 5 shuffles; lanes 0-15 -> row 0, lanes 16-31 -> row 1
                const bool hi16 = (lane & 16) != 0;
                double sx = hi16 ? part[0] : part[RB - 1];
                kx = hi16 ? part[RB - 1] : part[0];
                kx += __shfl_xor_sync(0xffffffffu, sx, 16);
                kx += __shfl_xor_sync(0xffffffffu, kx, 8);
            } else {  // RB == 1: plain butterfly, every lane ends with the row sum
                kx = part[0];
                kx += __shfl_xor_sync(0xffffffffu, kx, 16);
                kx += __shfl_xor_sync(0xffffffffu, kx, 8);
            }
            kx += __shfl_xor_sync(0xffffffffu, kx, 4);
            kx += __shfl_xor_sync(0xffffffffu, kx, 2);
            kx += __shfl_xor_sync(0xffffffffu, kx, 1);
            constexpr int LSH = (RB == 4) ? 3 : ((RB == 2) ? 4 : 5);  // row q's sum sits in lanes with (lane >> LSH) == q
            if constexpr (WARP_TEAM) {
#pragma unroll
                for (int q = 0; q < RB; ++q) t[q] = __shfl_sync(0xffffffffu, kx, q << LSH);
            } else {
                double* rbuf = red + batch_parity * (kMaxRB * 8);
                if ((lane & ((1 << LSH) - 1)) == 0 && (lane >> LSH) < RB) rbuf[(lane >> LSH) * 8 + warp] = kx;
                named_bar_sync(1, kThreads);
                if (u == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(j + NSt);
                }
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const double* rq = rbuf + q * 8;
                    t[q] = ((rq[0] + rq[1]) + (rq[2] + rq[3])) + ((rq[4] + rq[5]) + (rq[6] + rq[7]));
                }
                batch_parity ^= 1;
            }
            if (u == 0) {
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    tsq = fma(t[q], t[q], tsq);
                    if (MODE == MODE_JV && a.t_out != nullptr && q < rows_valid) a.t_out[row0 + q] = t[q];
                }
            }
        }
        if (MODE != MODE_JV) {
#pragma unroll
            for (int q = 0; q < RB; ++q) {
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    acc[k].x = fma(jr[q][k].x, t[q], acc[k].x);
                    acc[k].y = fma(jr[q][k].y, t[q], acc[k].y);
                }
            }
        }
    }

    // ---- per-CTA result: combine teams in fixed order, write partial[cta][0..ld) (+ sum t^2 at [ld]) ----
    double* pout = a.partial + (size_t)blockIdx.x * a.pstride;
    if (T == 1) {
        if (MODE != MODE_JV) {
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const int c = u + k * TG;
                if (c < NC) reinterpret_cast<double2*>(pout)[c] = acc[k];
            }
        }
        if (u == 0) pout[ld] = tsq;
    } else {
        // every armed stage was waited on by its team => all bulk copies have landed => the ring is reusable
        __syncthreads();
        double2* comb = reinterpret_cast<double2*>(stages);  // [T][NC] double2 (T*ld*8 bytes <= ring: NSt >= 1, RB >= 1)
        double* tsq_s = red;                                 // [T]
        if (MODE != MODE_JV) {
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const int c = u + k * TG;
                if (c < NC) comb[(size_t)team * NC + c] = acc[k];
            }
        }
        if (u == 0) tsq_s[team] = tsq;
        __syncthreads();
        if (MODE != MODE_JV) {
            for (int c = tid; c < NC; c += kThreads) {
                double2 s = make_double2(0.0, 0.0);
#pragma unroll
                for (int gg = 0; gg < T; ++gg) {
                    const double2 x = comb[(size_t)gg * NC + c];
                    s.x += x.x;
                    s.y += x.y;
                }
                reinterpret_cast<double2*>(pout)[c] = s;
            }
        }
        if (tid == 0) {
            double s = 0.0;
            for (int gg = 0; gg < T; ++gg) s += tsq_s[gg];
            pout[ld] = s;
        }
    }
}

template <int MODE>
cudaError_t launch_mode(const MvArgs& a, int kch, int rb, bool warp_team, int grid, size_t smem, cudaStream_t stream) {
#define BNL_LAUNCH(K, B, W)                                                                                    \
    {                                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(mv_stream_kernel<MODE, K, B, W>,                                  \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        if (e != cudaSuccess) return e;                                                                        \
        mv_stream_kernel<MODE, K, B, W><<<grid, kThreads, smem, stream>>>(a);                                  \
        return cudaGetLastError();                                                                             \
    }
    if (warp_team) {
        if (kch == 1 && rb == 4) BNL_LAUNCH(1, 4, true)
        if (kch == 2 && rb == 4) BNL_LAUNCH(2, 4, true)
        if (kch == 4 && rb == 4) BNL_LAUNCH(4, 4, true)
        if (kch == 8 && rb == 2) BNL_LAUNCH(8, 2, true)
        if (kch == 16 && rb == 1) BNL_LAUNCH(16, 1, true)
    } else {
        if (kch == 4 && rb == 4) BNL_LAUNCH(4, 4, false)
        if (kch == 8 && rb == 2) BNL_LAUNCH(8, 2, false)
        if (kch == 16 && rb == 1) BNL_LAUNCH(16, 1, false)
    }
    return cudaErrorInvalidValue;
#undef BNL_LAUNCH
}

// Sum partial[c][j] over CTAs c in fixed order.  One thread per column; columns [col0, ncols).
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nparts, long long pstride, int col0,
                                       int ncols, double* __restrict__ out) {
    const int j = col0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double s = 0.0;
    int c = 0;
    for (; c + 4 <= nparts; c += 4) {
        const double a0 = partial[(size_t)(c + 0) * pstride + j];
        const double a1 = partial[(size_t)(c + 1) * pstride + j];
        const double a2 = partial[(size_t)(c + 2) * pstride + j];
        const double a3 = partial[(size_t)(c + 3) * pstride + j];
        s += a0;
        s += a1;
        s += a2;
        s += a3;
    }
    for (; c < nparts; ++c) s += partial[(size_t)c * pstride + j];
    out[j] = s;
}

// reduce_partials_kernel + the push half of the peer-memory all-reduce in ONE kernel: the column sums of this rank
// go straight into every peer's mailbox over NVLink (p2p.h); out is not written here (p2p_wait_sum does it).
__global__ void reduce_push_kernel(const double* __restrict__ partial, int nparts, long long pstride, int col0, int ncols,
                                   P2PArgs p2p, unsigned long long epoch) {
    const int j = col0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ncols) {
        double s = 0.0;
        int c = 0;
        for (; c + 4 <= nparts; c += 4) {
            const double a0 = partial[(size_t)(c + 0) * pstride + j];
            const double a1 = partial[(size_t)(c + 1) * pstride + j];
            const double a2 = partial[(size_t)(c + 2) * pstride + j];
            const double a3 = partial[(size_t)(c + 3) * pstride + j];
            s += a0;
            s += a1;
            s += a2;
            s += a3;
        }
        for (; c < nparts; ++c) s += partial[(size_t)c * pstride + j];
        p2p_push_value(p2p, epoch, j, s);
    }
    p2p_push_finish(p2p, epoch);
}

}  // namespace

// ---- host-side planning ---------------------------------------------------------------------------------
MvPlan mv_make_plan(long long M, int n, int sm_count, size_t smem_optin_bytes) {
    MvPlan p{};
    p.ld = pad_cols(n);
    const int NC = p.ld / 2;
    p.supported = true;
    if (NC <= 512) {  // ld <= 1024: warp teams, a lane owns KCH <= 16 double2 chunks of the row
        p.warp_team = true;
        p.TG = 32;
        int kch = (NC + 31) / 32, k2 = 1;
        while (k2 < kch) k2 <<= 1;
        p.KCH = k2;
        p.RB = (k2 <= 4) ? 4 : (k2 == 8 ? 2 : 1);  // RB*KCH <= 16 double2 of J in registers
    } else {  // 1024 < ld <= 8192: one 256-thread team
        p.warp_team = false;
        p.TG = kThreads;
        const int kch = (NC + kThreads - 1) / kThreads;
        if (kch <= 4) {
            p.KCH = 4;
            p.RB = 4;
        } else if (kch <= 8) {
            p.KCH = 8;
            p.RB = 2;
        } else if (kch <= 16) {  // 4096 < ld <= 8192: one 64 KB row per stage
            p.KCH = 16;
            p.RB = 1;
        } else {
            p.supported = false;
            return p;  // ld > 8192: outside the streaming kernels' range
        }
    }
    p.R = p.RB;  // a stage is one RB-row batch of one team
    const int T = kThreads / p.TG;
    const size_t stage_bytes = (size_t)p.RB * p.ld * sizeof(double);
    const size_t fixed = 2 * kMaxRB * 8 * sizeof(double) + 2 * 32 * sizeof(uint64_t) + 256;
    size_t budget = smem_optin_bytes > fixed ? smem_optin_bytes - fixed : 0;
    if (budget > 200 * 1024) budget = 200 * 1024;
    int ns = (int)(budget / stage_bytes);
    if (ns > 32) ns = 32;
    ns = (ns / T) * T;  // every team owns NS/T private slots
    if (ns < 2 * T) p.supported = false;
    p.NS = ns;
    p.smem_bytes = (size_t)ns * stage_bytes + 2 * kMaxRB * 8 * sizeof(double) + (size_t)ns * sizeof(uint64_t);
    long long total_stages = (M + p.R - 1) / p.R;
    long long grid = total_stages < sm_count ? total_stages : sm_count;
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.pstride = p.ld + kColAlign;
    return p;
}

cudaError_t mv_launch(int mode, const MvPlan& p, const double* J, long long M, const double* v, const double* w,
                      double* t_out, double* partial, double* out, cudaStream_t stream, const P2PArgs* p2p,
                      unsigned long long epoch) {
    if (!p.supported) return cudaErrorInvalidValue;
    MvArgs a{};
    a.J = J;
    a.M = M;
    a.ld = p.ld;
    a.R = p.R;
    a.NS = p.NS;
    a.TG = p.TG;
    a.v = v;
    a.w = w;
    a.t_out = t_out;
    a.partial = partial;
    a.pstride = p.pstride;
    cudaError_t e;
    switch (mode) {
        case MODE_JTJV: e = launch_mode<MODE_JTJV>(a, p.KCH, p.RB, p.warp_team, p.grid, p.smem_bytes, stream); break;
        case MODE_JV: e = launch_mode<MODE_JV>(a, p.KCH, p.RB, p.warp_team, p.grid, p.smem_bytes, stream); break;
        case MODE_JTW: e = launch_mode<MODE_JTW>(a, p.KCH, p.RB, p.warp_team, p.grid, p.smem_bytes, stream); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    const int ncols = p.ld + 1;
    const int col0 = (mode == MODE_JV) ? p.ld : 0;  // JV only produces the sum-of-squares slot
    if (p2p != nullptr) {
        // fused local reduce + NVLink push; then wait for all ranks and sum in rank order (bit-identical everywhere)
        reduce_push_kernel<<<(ncols - col0 + 127) / 128, 128, 0, stream>>>(partial, p.grid, p.pstride, col0, ncols, *p2p, epoch);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return p2p_wait_sum(*p2p, epoch, out, col0, ncols, stream);
    }
    reduce_partials_kernel<<<(ncols - col0 + 127) / 128, 128, 0, stream>>>(partial, p.grid, p.pstride, col0, ncols, out);
    return cudaGetLastError();
}

}  // namespace bnl
