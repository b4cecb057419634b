// cauchy_loop.cu -- the breakpoint loop of cauchy_step (src/basic_tralcnlss.jl:615-636) as ONE persistent cooperative kernel.
//
// The reference recomputes Hd = H*d after every breakpoint (:633) although d only lost one component and the search only
// uses two scalars of it, phi'' = d'Hd and phi' = s_c'Hd + g'd (:634-635).  With t = J d and u = J s_c kept as M-vectors in
// HBM, phi'' = ||t||^2 (+ mu ||C d||^2), phi' = u.t (+ mu (C s_c).(C d)) + g.d, and a breakpoint is
//     u += theta t ;  t -= d_ind J[:,ind]
// -- one strided column of J and two vector streams instead of a pass over J.  That is the same algebra with different
// rounding, so the loop is GUARDED: it only takes a decision on its own when the decision is outside a rounding band
// (relative width `guard`, ~1e4 x the actual error); whenever the reference would stop at an interior minimiser (whose step
// length -phi'/phi'' enters the iterate) or a decision falls inside the band, the kernel returns CL_NEED_LITERAL, the host
// evaluates Hd = H*d literally (one fused pass over J + k_cauchy_eval) and re-enters the loop with those values.  Every
// number that reaches the iterate is therefore the literal one: the Cauchy point is bit-identical to the literal search.
//
// Grid: G worker CTAs (CTA b owns the chunks (g, b) of the local groups, rowgeom.h) + 1 controller CTA.
//   worker     : apply the previous breakpoint to its rows of t, u; write per-chunk partials of (t.t, u.t); arrive.
//   controller : while the workers stream, advance s_c / d / fixvars for the decided breakpoint and scan for the next one
//                (next_breakpoint :536-562, exactly the arithmetic of k_cauchy_eval); then wait for the arrivals, add the
//                partials in the fixed tree, exchange the 2 x 8 group sums with the peers (LL mailbox, p2p.h), decide,
//                broadcast (status, theta, ind, d_ind).
// No host round trip per breakpoint; ~20 us per breakpoint at 8 GPUs instead of ~75 us.
#include "cauchy_loop.h"

namespace bnl {
namespace {

constexpr int kCLThreads = 512;

struct __align__(16) CLBcast {
    unsigned long long seq;
    double theta, dind;
    long long ind;
    int status, pad;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kCLThreads) k_cauchy_loop(CauchyLoopArgs a) {
    __shared__ double shd[32];
    __shared__ long long shl[32];
    __shared__ double s_grp[kGroups][2];
    __shared__ double s_cd[kCLMaxP], s_cs[kCLMaxP];
    __shared__ CLBcast s_bc;
    __shared__ int s_fail;
    const VecCtx& c = a.c;
    const int tid = threadIdx.x;
    const int G = a.geo.G, ng = a.geo.ng;
    CLBcast* bc = reinterpret_cast<CLBcast*>(a.bcast);

    if (blockIdx.x < (unsigned)G) {
        // ------------------------------------------------ worker ------------------------------------------------
        const int b = blockIdx.x;
        double* __restrict__ tp = a.t;
        double* __restrict__ up = a.u;
        double theta = 0.0, dind = 0.0;
        long long ind = -1;
        // a fresh search entered with the literal scalars of its first interval never ran the it == 0 pass that zeroes u
        const unsigned long long zero_u_at = (a.first && a.use_literal) ? 1ull : ~0ull;
        for (unsigned long long it = 0;; ++it) {
            const bool skip = (it == 0 && a.use_literal);  // the first decision of a (re-)entry uses the literal scalars
            if (!skip) {
                for (int gi = 0; gi < ng; ++gi) {
                    const long long lb = a.geo.local_begin(gi, b), le = a.geo.local_end(gi, b);
                    double tt = 0.0, ut = 0.0;
                    if (it == 0) {
                        for (long long i = lb + tid; i < le; i += kCLThreads) {
                            const double ti = tp[i], ui = a.first ? 0.0 : up[i];
                            if (a.first) up[i] = 0.0;
                            tt = fma(ti, ti, tt);
                            ut = fma(ui, ti, ut);
                        }
                    } else {
                        // u += theta t ; t -= d_ind J[:,ind].  The column read is one 8-byte element per 8 KB row: pure latency,
                        // so four rows are in flight per thread (loads first, then the arithmetic in row order: the per-thread
                        // summation order does not depend on the unrolling).
                        const bool zero_u = (it == zero_u_at);  // u = J s_c = 0 before the first breakpoint
                        // element (r, ind): from the tile-transposed copy when it exists (16 consecutive rows of a column share a
                        // 128-byte line, the granularity HBM serves a strided read at: profiles/r2_cauchy_loop_ncu.md), else from
                        // the row-major Jacobian (one line per row)
                        const bool tr = a.Jt != nullptr;
                        const double* __restrict__ Jc = tr ? a.Jt + (size_t)ind * kJtTile : a.J + ind;
                        const size_t tile_stride = (size_t)a.ld * kJtTile;
                        auto jelem = [&](long long r) -> double {
                            return tr ? __ldg(Jc + (size_t)(r >> 4) * tile_stride + (r & 15)) : __ldg(Jc + (size_t)r * a.ld);
                        };
                        long long i = lb + tid;
                        for (; i + 3 * kCLThreads < le; i += 4 * kCLThreads) {
                            double tv[4], uv[4], jv[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const long long r = i + (long long)q * kCLThreads;
                                tv[q] = tp[r];
                                uv[q] = zero_u ? 0.0 : up[r];
                                jv[q] = jelem(r);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const long long r = i + (long long)q * kCLThreads;
                                const double ui = fma(theta, tv[q], uv[q]);
                                const double ti = fma(-dind, jv[q], tv[q]);
                                tp[r] = ti;
                                up[r] = ui;
                                tt = fma(ti, ti, tt);
                                ut = fma(ui, ti, ut);
                            }
                        }
                        for (; i < le; i += kCLThreads) {
                            const double t0 = tp[i], u0 = zero_u ? 0.0 : up[i];
                            const double ui = fma(theta, t0, u0);
                            const double ti = fma(-dind, jelem(i), t0);
                            tp[i] = ti;
                            up[i] = ui;
                            tt = fma(ti, ti, tt);
                            ut = fma(ui, ti, ut);
                        }
                    }
                    tt = block_sum(tt, shd);
                    ut = block_sum(ut, shd);
                    if (tid == 0) {
                        double* p2 = a.partial2 + ((size_t)gi * G + b) * 2;
                        p2[0] = tt;
                        p2[1] = ut;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(a.arrive, 1u);
                while (ld_acquire_u64(&bc->seq) < it + 1) {
                }
                s_bc.theta = *reinterpret_cast<volatile double*>(&bc->theta);
                s_bc.dind = *reinterpret_cast<volatile double*>(&bc->dind);
                s_bc.ind = *reinterpret_cast<volatile long long*>(&bc->ind);
                s_bc.status = *reinterpret_cast<volatile int*>(&bc->status);
            }
            __syncthreads();
            if (s_bc.status != CL_ADVANCE) return;
            theta = s_bc.theta;
            dind = s_bc.dind;
            ind = s_bc.ind;
            __syncthreads();
        }
    }

    // -------------------------------------------------- controller --------------------------------------------------
    const int lane = tid & 31, warp = tid >> 5;
    int status = CL_ADVANCE;
    int breakpoints = 0;
    unsigned long long rounds = 0;  // LL exchange rounds performed (all ranks perform the same number)
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (unsigned long long it = 0;; ++it) {
        // ---- next_breakpoint :536-562 + g.d (+ C-part) on the current s_c, d, fixvars: overlaps the workers' streaming ----
        double gd = 0.0, th = INFINITY;
        long long ind = -1;
        for (int i = tid; i < c.n; i += kCLThreads) {
            const double di = c.d[i], si = c.s[i];
            gd = fma(c.g[i], di, gd);
            if (!c.fix[i]) {
                double tt = INFINITY;
                if (di < 0.0) {
                    const double dl = fmax(c.xlow[i] - c.x[i], -a.delta);  // d_l :603
                    tt = (dl - si) / di;
                } else if (di > 0.0) {
                    const double du = fmin(c.xupp[i] - c.x[i], a.delta);  // d_u :602
                    tt = (du - si) / di;
                }
                if (tt < th) {  // strict <, ascending i within a thread
                    th = tt;
                    ind = i;
                }
            }
        }
        gd = block_sum(gd, shd);
        block_argmin(th, ind, shd, shl);
        const double dind = (ind >= 0) ? c.d[ind] : 0.0;
        double c_pp = 0.0, c_p = 0.0;  // mu ||C d||^2 and mu (C s_c).(C d)
        if (c.p > 0) {
            for (int r = warp; r < c.p; r += kCLThreads / 32) {
                double x1 = 0.0, x2 = 0.0;
                for (int j = lane; j < c.n; j += 32) {
                    const double cij = c.C[(size_t)r * c.ld + j];
                    x1 = fma(cij, c.d[j], x1);
                    x2 = fma(cij, c.s[j], x2);
                }
                x1 = warp_sum(x1);
                x2 = warp_sum(x2);
                if (lane == 0) {
                    s_cd[r] = x1;
                    s_cs[r] = x2;
                }
            }
            __syncthreads();
            for (int r = 0; r < c.p; ++r) {
                c_pp = fma(s_cd[r], s_cd[r], c_pp);
                c_p = fma(s_cs[r], s_cd[r], c_p);
            }
            c_pp *= c.mu;
            c_p *= c.mu;
        }
        const int nb_fix = c.sd->nb_fix;
        if (it == 0 && nb_fix >= a.nmm) {  // the reference's loop condition fails on entry (:615)
            status = CL_DONE_EXHAUSTED;
        }
        // ---- wait for the workers' partials of this iteration ----
        if (tid == 0) {
            const unsigned int target = (unsigned int)G * (unsigned int)(it + 1);
            while (ld_acquire_u32(a.arrive) < target) {
            }
        }
        __syncthreads();
        double phi_p = 0.0, phi_pp = 0.0, scale = 0.0;
        const bool literal = (it == 0 && a.use_literal);
        if (status == CL_ADVANCE && !literal) {
            // fixed tree: warp gi adds the G chunk partials of local group gi (lane-strided chains, then the xor tree)
            if (warp < ng) {
                double x0 = 0.0, x1 = 0.0;
                for (int b = lane; b < G; b += 32) {
                    const double* p2 = a.partial2 + ((size_t)warp * G + b) * 2;
                    x0 += __ldcg(p2);
                    x1 += __ldcg(p2 + 1);
                }
                x0 = warp_sum(x0);
                x1 = warp_sum(x1);
                if (lane == 0) {
                    s_grp[a.geo.g0 + warp][0] = x0;
                    s_grp[a.geo.g0 + warp][1] = x1;
                }
            }
            __syncthreads();
            if (a.multi) {
                const unsigned long long ep = a.ll_epoch0 + rounds + 1;
                // push this rank's group sums to every rank (itself included), then collect all kGroups x 2 values
                const int nsend = a.p2p.nranks * ng * 2;
                if (tid < nsend) {
                    const int peer = tid / (ng * 2), rem = tid % (ng * 2), gi = rem >> 1, v = rem & 1;
                    ll_store(a.p2p, peer, ep, a.geo.g0 + gi, v, s_grp[a.geo.g0 + gi][v]);
                }
                __syncthreads();
                if (tid < kGroups * 2) {
                    double x;
                    if (ll_load(a.p2p, ep, tid >> 1, tid & 1, &x))
                        s_grp[tid >> 1][tid & 1] = x;
                    else
                        s_fail = 1;
                }
                __syncthreads();
            }
            ++rounds;
            double tt = s_grp[0][0], ut = s_grp[0][1];
#pragma unroll
            for (int g = 1; g < kGroups; ++g) {
                tt += s_grp[g][0];
                ut += s_grp[g][1];
            }
            phi_pp = tt + c_pp;           // dot(d,Hd)   = ||J d||^2 + mu ||C d||^2
            phi_p = (ut + c_p) + gd;      // dot(s_c,Hd) + dot(g,d)
            scale = fabs(ut) + fabs(c_p) + fabs(gd);
            if (s_fail) status = CL_TIMEOUT;
        } else if (status == CL_ADVANCE) {
            phi_p = c.sd->phi_p;  // literal values from k_cauchy_eval (:609-611 / :634-635)
            phi_pp = c.sd->phi_pp;
        }
        // ---- decision :618-636 ----
        double step = 0.0;
        bool advance = false;
        if (status == CL_ADVANCE) {
            if (literal) {
                const double delta_t = (phi_pp > 0.0) ? -phi_p / phi_pp : 0.0;  // :618
                if (phi_p >= 0.0) {
                    status = CL_DONE_NOSTEP;  // :620
                } else if (phi_p < 0.0 && phi_pp > 0.0 && delta_t < th) {
                    step = delta_t;  // :622-626
                    status = CL_DONE_INTERIOR;
                } else {
                    advance = true;
                }
            } else {
                if (!(fabs(phi_p) > a.guard * scale)) {
                    status = CL_NEED_LITERAL;  // the sign of phi' is inside the rounding band (also NaN)
                } else if (phi_p > 0.0) {
                    status = CL_DONE_NOSTEP;
                } else if (!(phi_pp > 0.0)) {
                    status = CL_NEED_LITERAL;
                } else {
                    const double delta_t = -phi_p / phi_pp;
                    const double relb = a.guard * scale / fabs(phi_p) + a.guard;
                    if (delta_t > th * (1.0 + relb))
                        advance = true;  // clearly beyond the next breakpoint
                    else
                        status = CL_NEED_LITERAL;  // interior minimiser (needs the literal step length) or inside the band
                }
            }
            if (advance && a.want_jt && a.q0 + breakpoints >= kJtTrigger) {
                advance = false;  // a long search: the host builds the tile-transposed copy of J, then re-enters here
                status = CL_WANT_TRANSPOSE;
            }
            if (advance) {
                if (ind < 0)
                    status = CL_ERR_BOUNDS;  // add_active!(ind = -1): BoundsError in the reference (:544, :631)
                else if (nb_fix + 1 >= a.nmm)
                    status = CL_DONE_EXHAUSTED;  // the loop condition (:615) fails after this breakpoint
            }
        }
        // ---- broadcast, then advance the n-vectors (the workers stream meanwhile) ----
        if (tid == 0) {
            bc->theta = th;
            bc->dind = dind;
            bc->ind = ind;
            bc->status = status;
            st_release_u64(&bc->seq, it + 1);
        }
        if (advance && status != CL_ERR_BOUNDS) {  // :628-632 (mask projection: d = P(-g) only loses component ind)
            for (int i = tid; i < c.n; i += kCLThreads) {
                c.s[i] = c.s[i] + th * c.d[i];
                if (i == ind) {
                    c.fix[i] = 1;
                    c.d[i] = 0.0;
                }
            }
            if (tid == 0) c.sd->nb_fix = nb_fix + 1;
            ++breakpoints;
        } else if (status == CL_DONE_INTERIOR) {
            for (int i = tid; i < c.n; i += kCLThreads) c.s[i] = c.s[i] + step * c.d[i];  // :625
        }
        __syncthreads();
        if (status != CL_ADVANCE) break;
    }
    if (tid == 0) {
        c.sd->cl_status = status;
        c.sd->cl_breakpoints = breakpoints;
        c.sd->cl_rounds = (long long)rounds;
    }
    __syncthreads();
    const int nw = sizeof(Scal) / 8;
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(c.sd);
    unsigned long long* dsh = reinterpret_cast<unsigned long long*>(c.sh);
    for (int i = tid; i < nw; i += kCLThreads) dsh[i] = s[i];
}

// Jt[((i/16)*ld + j)*16 + i%16] = J[i][j]: one CTA per 16-row tile; 32 columns at a time go through shared memory, so both the
// reads (256 B per row) and the writes (the 16 x 32 sub-tile is one contiguous 4 KB block of Jt) are coalesced.
__global__ void __launch_bounds__(512) k_transpose16(const double* __restrict__ J, long long M, int ld, double* __restrict__ Jt) {
    __shared__ double tile[kJtTile][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // load: row ty, column tx
    const int rr = threadIdx.x & 15, cc = threadIdx.x >> 4;        // store: column cc, row rr
    for (long long tb = blockIdx.x; tb * kJtTile < M; tb += gridDim.x) {
        const long long r = tb * kJtTile + ty;
        for (int c0 = 0; c0 < ld; c0 += 32) {
            tile[ty][tx] = (r < M && c0 + tx < ld) ? J[(size_t)r * ld + c0 + tx] : 0.0;
            __syncthreads();
            if (c0 + cc < ld) Jt[((size_t)tb * ld + c0 + cc) * kJtTile + rr] = tile[rr][cc];
            __syncthreads();
        }
    }
}

}  // namespace

cudaError_t transpose16_launch(const double* J, long long M, int ld, double* Jt, cudaStream_t st) {
    long long tiles = (M + kJtTile - 1) / kJtTile;
    int grid = (int)(tiles < 148 * 16 ? (tiles > 0 ? tiles : 1) : 148 * 16);
    k_transpose16<<<grid, 512, 0, st>>>(J, M, ld, Jt);
    return cudaGetLastError();
}

size_t cauchy_loop_sync_bytes() { return 256; }

cudaError_t cauchy_loop_launch(const CauchyLoopArgs& a_in, int sm_count, cudaStream_t st) {
    static int max_blocks_per_sm = -1;
    if (max_blocks_per_sm < 0) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_cauchy_loop, kCLThreads, 0);
        if (e != cudaSuccess) return e;
        max_blocks_per_sm = nb;
    }
    if ((long long)max_blocks_per_sm * sm_count < a_in.geo.G + 1) return cudaErrorCooperativeLaunchTooLarge;
    CauchyLoopArgs a = a_in;
    // arrive counter and broadcast block start from zero at every launch
    cudaError_t e = cudaMemsetAsync(a.arrive, 0, cauchy_loop_sync_bytes(), st);
    if (e != cudaSuccess) return e;
    void* params[] = {&a};
    return cudaLaunchCooperativeKernel((const void*)k_cauchy_loop, dim3(a.geo.G + 1), dim3(kCLThreads), params, 0, st);
}

}  // namespace bnl
