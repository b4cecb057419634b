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
// shared-memory ring of NS stages; full/empty mbarriers decouple the producer warp from 8 consumer warps.
//
// Consumers: thread-owns-columns.  A "row group" of TG threads (TG = min(256, pow2ceil(ld/2)), >= 32)
// shares one row; thread u of the group owns the double2 column chunks u + k*TG (k < KCH), keeps v and the
// J' accumulator for them in registers, reads its part of 4 rows from smem with conflict-free LDS.128,
// then the slot is released at once (the row lives in registers from here on).  The 4 row dot products are
// reduced with a 6-shuffle transposing butterfly + one named barrier per batch; J'.t is accumulated from
// the same registers.  Per CTA the n column sums (and sum t^2) go to partial[cta][*]; a second tiny kernel
// sums the partials in fixed CTA order => deterministic, no FP64 atomics.
//
// Algorithmic bytes per launch (what roofline.achieved uses): 8*M_loc*ld (+ O(n)); J is read exactly once.
#include "common.cuh"
#include "matvec.h"

namespace bnl {

namespace {

constexpr int kConsumers = 256;
constexpr int kThreads = kConsumers + 32;  // + 1 producer warp
constexpr int kMaxRB = 4;                  // rows per reduction batch (template RB = 4 or 2)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
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

template <int MODE, int KCH, int RB>
__global__ void __launch_bounds__(kThreads, 1) mv_stream_kernel(const MvArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ld = a.ld;
    const int NC = ld >> 1;  // double2 chunks per row
    const int R = a.R;       // rows per stage
    const int NS = a.NS;
    const size_t stage_doubles = (size_t)R * ld;
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* red = stages + (size_t)NS * stage_doubles;                     // [2][kMaxRB][8]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 2 * kMaxRB * 8);  // [NS]
    uint64_t* empty_bar = full_bar + NS;                               // [NS]

    const int tid = threadIdx.x;
    const long long total_stages = (a.M + R - 1) / R;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kConsumers) {
        // ===================== producer warp: one elected lane issues the bulk copies ==================
        if (tid == kConsumers) {
            const uint64_t pol = policy_evict_first();
            long long it = 0;
            for (long long st = blockIdx.x; st < total_stages; st += gridDim.x, ++it) {
                const int slot = (int)(it % NS);
                const uint32_t use = (uint32_t)(it / NS);
                mbar_wait(&empty_bar[slot], (use & 1u) ^ 1u);
                const long long row0 = st * R;
                const long long rows = (a.M - row0 < R) ? (a.M - row0) : R;
                const uint32_t bytes = (uint32_t)(rows * ld * sizeof(double));
                mbar_arrive_expect_tx(&full_bar[slot], bytes);
                tma_bulk_g2s(stages + (size_t)slot * stage_doubles, a.J + row0 * ld, bytes, &full_bar[slot], pol);
            }
        }
        return;
    }

    // ============================ consumers =============================================================
    const int TG = a.TG;           // threads per row group
    const int G = kConsumers / TG;  // row groups
    const int g = tid / TG;
    const int u = tid - g * TG;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int WPG = TG >> 5;  // warps per group
    const int RG = R / G;     // rows per group per stage (R is a multiple of G)

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

    long long it = 0;
    for (long long st = blockIdx.x; st < total_stages; st += gridDim.x, ++it) {
        const int slot = (int)(it % NS);
        const uint32_t use = (uint32_t)(it / NS);
        const long long row0 = st * R;
        const int rows_valid = (int)((a.M - row0 < R) ? (a.M - row0) : R);
        mbar_wait(&full_bar[slot], use & 1u);
        const double2* sbase = reinterpret_cast<const double2*>(stages + (size_t)slot * stage_doubles);

        const int nbatch = (RG + RB - 1) / RB;
        for (int b = 0; b < nbatch; ++b) {
            double2 jr[RB][KCH];
            int ridx[RB];
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const int rg = b * RB + q;
                const int r = g + G * rg;
                const bool valid = (rg < RG) && (r < rows_valid);
                ridx[q] = valid ? r : -1;
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const int c = u + k * TG;
                    jr[q][k] = (valid && c < NC) ? sbase[(size_t)r * NC + c] : make_double2(0.0, 0.0);
                }
            }
            if (b == nbatch - 1) {  // all of this thread's reads of the slot are done: hand it back
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[slot]);
            }

            double t[RB];
            if (MODE == MODE_JTW) {
#pragma unroll
                for (int q = 0; q < RB; ++q) t[q] = (ridx[q] >= 0) ? __ldg(a.w + row0 + ridx[q]) : 0.0;
            } else {
                double part[RB];
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < KCH; ++k) {
                        s = fma(jr[q][k].x, vv[k].x, s);
                        s = fma(jr[q][k].y, vv[k].y, s);
                    }
                    part[q] = s;
                }
                // transposing butterfly: RB values x 32 lanes -> lane holds one row's warp sum
                double kx;
                if constexpr (RB == 4) {  // 6 shuffles; lane holds row (2*bit4 + bit3)
                    const bool hi16 = (lane & 16) != 0;
                    double s0 = hi16 ? part[0] : part[2];
                    double s1 = hi16 ? part[1] : part[3];
                    double k0 = hi16 ? part[2] : part[0];
                    double k1 = hi16 ? part[3] : part[1];
                    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                    const bool hi8 = (lane & 8) != 0;
                    double sx = hi8 ? k0 : k1;
                    kx = hi8 ? k1 : k0;
                    kx += __shfl_xor_sync(0xffffffffu, sx, 8);
                } else {  // RB == 2: 5 shuffles; lanes 0-7 -> row 0, lanes 16-23 -> row 1 (rows 2,3 of the slot unused)
                    const bool hi16 = (lane & 16) != 0;
                    double sx = hi16 ? part[0] : part[RB - 1];
                    kx = hi16 ? part[RB - 1] : part[0];
                    kx += __shfl_xor_sync(0xffffffffu, sx, 16);
                    kx += __shfl_xor_sync(0xffffffffu, kx, 8);
                }
                kx += __shfl_xor_sync(0xffffffffu, kx, 4);
                kx += __shfl_xor_sync(0xffffffffu, kx, 2);
                kx += __shfl_xor_sync(0xffffffffu, kx, 1);
                // writer lanes: RB==4 -> lanes 0,8,16,24 (row = lane>>3); RB==2 -> lanes 0,16 (row = lane>>4)
                const int wrow = (RB == 4) ? (lane >> 3) : (lane >> 4);
                const bool writer = (RB == 4) ? ((lane & 7) == 0) : ((lane & 15) == 0);
                double* rbuf = red + batch_parity * (kMaxRB * 8);
                if (writer) rbuf[wrow * 8 + warp] = kx;
                if (WPG == 1)
                    __syncwarp();
                else
                    named_bar_sync(1 + g, TG);
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    double s = 0.0;
                    for (int w = 0; w < WPG; ++w) s += rbuf[q * 8 + g * WPG + w];
                    t[q] = s;
                }
                batch_parity ^= 1;
                if (u == 0) {
#pragma unroll
                    for (int q = 0; q < RB; ++q) {
                        tsq = fma(t[q], t[q], tsq);
                        if (MODE == MODE_JV && a.t_out != nullptr && ridx[q] >= 0) a.t_out[row0 + ridx[q]] = t[q];
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
    }

    // ---- per-CTA result: combine row groups in fixed order, write partial[cta][0..ld] (+ tsq at [ld]) ----
    double* pout = a.partial + (size_t)blockIdx.x * a.pstride;
    if (G == 1) {
        if (MODE != MODE_JV) {
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const int c = u + k * TG;
                if (c < NC) reinterpret_cast<double2*>(pout)[c] = acc[k];
            }
        }
        if (u == 0) pout[ld] = tsq;
    } else {
        // all bulk copies this CTA issued have completed (every stage was waited on) => ring is reusable
        named_bar_sync(10, kConsumers);
        double2* comb = reinterpret_cast<double2*>(stages);  // [G][NC] double2, G*ld*8 <= stage bytes (R >= G)
        double* tsq_s = red;                                 // [G]
        if (MODE != MODE_JV) {
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const int c = u + k * TG;
                if (c < NC) comb[(size_t)g * NC + c] = acc[k];
            }
        }
        if (u == 0) tsq_s[g] = tsq;
        named_bar_sync(10, kConsumers);
        if (MODE != MODE_JV) {
            for (int c = tid; c < NC; c += kConsumers) {
                double2 s = make_double2(0.0, 0.0);
                for (int gg = 0; gg < G; ++gg) {
                    const double2 x = comb[(size_t)gg * NC + c];
                    s.x += x.x;
                    s.y += x.y;
                }
                reinterpret_cast<double2*>(pout)[c] = s;
            }
        }
        if (tid == 0) {
            double s = 0.0;
            for (int gg = 0; gg < G; ++gg) s += tsq_s[gg];
            pout[ld] = s;
        }
    }
}

template <int MODE>
cudaError_t launch_mode(const MvArgs& a, int kch, int rb, int grid, size_t smem, cudaStream_t stream) {
#define BNL_LAUNCH(K, B)                                                                                       \
    {                                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(mv_stream_kernel<MODE, K, B>,                                     \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        if (e != cudaSuccess) return e;                                                                        \
        mv_stream_kernel<MODE, K, B><<<grid, kThreads, smem, stream>>>(a);                                     \
        return cudaGetLastError();                                                                             \
    }
    if (kch == 1 && rb == 4) BNL_LAUNCH(1, 4)
    if (kch == 2 && rb == 4) BNL_LAUNCH(2, 4)
    if (kch == 4 && rb == 4) BNL_LAUNCH(4, 4)
    if (kch == 8 && rb == 2) BNL_LAUNCH(8, 2)
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

}  // namespace

// ---- host-side planning ---------------------------------------------------------------------------------
MvPlan mv_make_plan(long long M, int n, int sm_count, size_t smem_optin_bytes) {
    MvPlan p{};
    p.ld = pad_cols(n);
    const int NC = p.ld / 2;
    int tg = 32;
    while (tg < NC && tg < kConsumers) tg <<= 1;
    p.TG = tg;
    const int G = kConsumers / tg;
    int kch = (NC + tg - 1) / tg;
    int kch_t = 1;
    while (kch_t < kch) kch_t <<= 1;
    p.KCH = kch_t;  // supported: 1,2,4,8  => ld <= 4096
    p.RB = (kch_t <= 4) ? 4 : 2;  // register budget: RB*KCH double2 of J per thread
    p.supported = (kch_t <= 8);
    const size_t row_bytes = (size_t)p.ld * sizeof(double);
    // ~32 KB stages (64 KB once a row exceeds 8 KB), R a multiple of G, at most 64 rows
    const size_t target = (p.ld <= 1024) ? 32768 : 65536;
    long long R = (long long)(target / row_bytes);
    if (R < 1) R = 1;
    if (R > 64) R = 64;
    R = (R / G) * G;
    if (R < G) R = G;
    p.R = (int)R;
    const size_t stage_bytes = (size_t)p.R * row_bytes;
    const size_t fixed = 2 * kMaxRB * 8 * sizeof(double) + 2 * 8 * sizeof(uint64_t) + 256;
    size_t budget = smem_optin_bytes > fixed ? smem_optin_bytes - fixed : 0;
    if (budget > 200 * 1024) budget = 200 * 1024;
    int ns = (int)(budget / stage_bytes);
    if (ns > 8) ns = 8;
    if (ns < 2) p.supported = false;
    p.NS = ns;
    p.smem_bytes = (size_t)ns * stage_bytes + 2 * kMaxRB * 8 * sizeof(double) + 2 * (size_t)ns * sizeof(uint64_t);
    long long total_stages = (M + p.R - 1) / p.R;
    long long grid = total_stages < sm_count ? total_stages : sm_count;
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.pstride = p.ld + kColAlign;
    return p;
}

cudaError_t mv_launch(int mode, const MvPlan& p, const double* J, long long M, const double* v, const double* w,
                      double* t_out, double* partial, double* out, cudaStream_t stream) {
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
        case MODE_JTJV: e = launch_mode<MODE_JTJV>(a, p.KCH, p.RB, p.grid, p.smem_bytes, stream); break;
        case MODE_JV: e = launch_mode<MODE_JV>(a, p.KCH, p.RB, p.grid, p.smem_bytes, stream); break;
        case MODE_JTW: e = launch_mode<MODE_JTW>(a, p.KCH, p.RB, p.grid, p.smem_bytes, stream); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    const int ncols = p.ld + 1;
    const int col0 = (mode == MODE_JV) ? p.ld : 0;  // JV only produces the sum-of-squares slot
    reduce_partials_kernel<<<(ncols - col0 + 127) / 128, 128, 0, stream>>>(partial, p.grid, p.pstride, col0, ncols, out);
    return cudaGetLastError();
}

}  // namespace bnl
