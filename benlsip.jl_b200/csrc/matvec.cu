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
// 256-thread team: + one named barrier); J'.t is accumulated from the same registers.
//
// Row geometry (rowgeom.h): CTA b owns the chunks (g, b) of this rank's groups and walks them one after the other; the TMA
// ring keeps running across chunk boundaries, but the accumulators are flushed to partial[g][b][team][*] at every chunk end,
// so a partial only depends on chunk-local row indices.  group_reduce / group_sum (p2p.h) then add teams, chunks and groups
// in one fixed tree => deterministic, no FP64 atomics, and bit-identical for 1, 2, 4 or 8 GPUs.
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
    const int ng = a.geo.ng;
    __shared__ long long s_cb[kGroups], s_ce[kGroups], s_nst[kGroups];  // chunk (gi, blockIdx.x): local rows, stages
    __shared__ long long s_pst[kThreads / 32];                          // producer cursors (team leaders only)
    __shared__ int s_pg[kThreads / 32];

    if (tid == 0) {
        for (int s = 0; s < a.NS; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < ng) {
        const long long cb = a.geo.local_begin(tid, blockIdx.x), ce = a.geo.local_end(tid, blockIdx.x);
        s_cb[tid] = cb;
        s_ce[tid] = ce;
        s_nst[tid] = (ce - cb + RB - 1) / RB;
    }
    __syncthreads();

    const int team = tid / TG;
    const int u = tid - team * TG;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    double* my_stages = stages + (size_t)team * NSt * stage_doubles;
    uint64_t* my_full = full_bar + team * NSt;
    uint64_t pol = 0;

    // leader only: arm `slot` with the team's next stage in (chunk, stage) order; the cursor lives in shared memory
    auto issue_next = [&](int slot) {
        int pg = s_pg[team];
        long long pst = s_pst[team];
        while (pg < ng && pst >= s_nst[pg]) {
            ++pg;
            pst = team;
        }
        if (pg < ng) {
            const long long row0 = s_cb[pg] + pst * RB;
            const long long left = s_ce[pg] - row0;
            const long long rows = left < RB ? left : RB;
            const uint32_t bytes = (uint32_t)(rows * ld * sizeof(double));
            mbar_arrive_expect_tx(&my_full[slot], bytes);
            tma_bulk_g2s(my_stages + (size_t)slot * stage_doubles, a.J + row0 * ld, bytes, &my_full[slot], pol);
            pst += T;
        }
        s_pg[team] = pg;
        s_pst[team] = pst;
    };
    if (u == 0) {
        pol = policy_evict_first();
        s_pg[team] = 0;
        s_pst[team] = team;
        for (int j = 0; j < NSt; ++j) issue_next(j);
    }

    double2 vv[KCH];
    double2 acc[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int c = u + k * TG;
        if (MODE != MODE_JTW)
            vv[k] = (c < NC) ? reinterpret_cast<const double2*>(a.v)[c] : make_double2(0.0, 0.0);
        else
            vv[k] = make_double2(0.0, 0.0);
    }
    int batch_parity = 0;
    int slot = 0;
    uint32_t phase = 0;

    for (int gi = 0; gi < ng; ++gi) {
#pragma unroll
        for (int k = 0; k < KCH; ++k) acc[k] = make_double2(0.0, 0.0);
        double tsq = 0.0;
        const long long cb = s_cb[gi], ce = s_ce[gi], nst = s_nst[gi];
        for (long long st = team; st < nst; st += T) {
            const long long row0 = cb + st * RB;
            const int rows_valid = (int)((ce - row0 < RB) ? (ce - row0) : RB);
            mbar_wait(&my_full[slot], phase);
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
                    issue_next(slot);
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
                    // Re-arm the slot as early as safe: part[] depends on every J value this lane loaded (the empty asm pins
                    // the re-arm below the FMAs that consumed them), and the warp barrier makes sure EVERY lane is past its
                    // reads of the slot before the leader lets the bulk copy overwrite it (lanes may leave the mbarrier
                    // spin loop on different iterations under independent thread scheduling).
                    double dep = part[0];
#pragma unroll
                    for (int q = 1; q < RB; ++q) dep += part[q];
                    asm volatile("" ::"d"(dep) : "memory");
                    __syncwarp();
                    if (u == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue_next(slot);
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
                        issue_next(slot);
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
                        if (a.t_out != nullptr && q < rows_valid) a.t_out[row0 + q] = t[q];  // t = J v rows (JV and JTJV modes)
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
            if (++slot == NSt) {
                slot = 0;
                phase ^= 1u;
            }
        }
        // ---- chunk end: flush this team's partial for chunk (gi, blockIdx.x): [0..ld) column sums, [ld] sum t^2 ----
        double* pout = a.partial + (((size_t)gi * a.geo.G + blockIdx.x) * T + team) * a.pstride;
        if (MODE != MODE_JV) {
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const int c = u + k * TG;
                if (c < NC) reinterpret_cast<double2*>(pout)[c] = acc[k];
            }
        }
        if (MODE != MODE_JTW && u == 0) pout[ld] = tsq;
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

}  // namespace

// ---- host-side planning ---------------------------------------------------------------------------------
MvPlan mv_make_plan(int n, size_t smem_optin_bytes) {
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
    p.T = T;
    p.pstride = p.ld + kColAlign;
    return p;
}

cudaError_t mv_launch(int mode, const MvPlan& p, const RowGeom& geo, const double* J, const double* v, const double* w,
                      double* t_out, double* partial, cudaStream_t stream) {
    if (!p.supported) return cudaErrorInvalidValue;
    MvArgs a{};
    a.J = J;
    a.geo = geo;
    a.ld = p.ld;
    a.R = p.R;
    a.NS = p.NS;
    a.TG = p.TG;
    a.v = v;
    a.w = w;
    a.t_out = t_out;
    a.partial = partial;
    a.pstride = p.pstride;
    switch (mode) {
        case MODE_JTJV: return launch_mode<MODE_JTJV>(a, p.KCH, p.RB, p.warp_team, geo.G, p.smem_bytes, stream);
        case MODE_JV: return launch_mode<MODE_JV>(a, p.KCH, p.RB, p.warp_team, geo.G, p.smem_bytes, stream);
        case MODE_JTW: return launch_mode<MODE_JTW>(a, p.KCH, p.RB, p.warp_team, geo.G, p.smem_bytes, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace bnl
