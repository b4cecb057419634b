// capi.cu -- the extern "C" layer of libbenlsip_b200.so (include/benlsip_b200.h): argument checks, host<->device staging
// of the caller's arrays, status codes; the control flow it calls lives in solver.cu.
#include "solver_internal.h"

using namespace bnl_host;

extern "C" {

int bnl_version(void) { return 100; }

int bnl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* bnl_status_string(int s) {
    switch (s) {
        case BNL_OK: return "ok";
        case BNL_EINVAL: return "invalid argument";
        case BNL_EDIM: return "DimensionMismatch";
        case BNL_ECUDA: return "CUDA error";
        case BNL_ENCCL: return "NCCL error";
        case BNL_EOOM: return "out of device memory";
        case BNL_ENOTPD: return "PosDefException";
        case BNL_EBOUNDS: return "BoundsError";
        case BNL_EASSERT: return "AssertionError";
        case BNL_ENODEV: return "no sm_100 CUDA device (libbenlsip_b200 has no CPU path)";
        case BNL_ECALLBACK: return "user callback failed";
    }
    return "unknown";
}

void bnl_default_params(bnl_params* p) {
    const double sqrt_eps = 1.4901161193847656e-08;  // sqrt(eps(Float64))
    p->eta1 = 0.25;
    p->eta2 = 0.75;
    p->gamma1 = 0.0625;
    p->gamma2 = 2.0;
    p->kappa2 = 0.1;
    p->kappa3 = 0.1;
    p->tr_factor = 0.1;
    p->atol_active = sqrt_eps;
    p->atol_negcurve = sqrt_eps;
    p->atol_boundary = 1e-10;
    p->max_minor_iter = 50;
    p->max_inner_iter = 500;
}

void bnl_default_outer_params(bnl_outer_params* p) {
    const double sqrt_eps = 1.4901161193847656e-08;
    p->mu0 = 10.0;
    p->tau = 100.0;
    p->omega0 = 1.0;
    p->eta0 = 1.0;
    p->feas_tol = sqrt_eps;
    p->crit_tol = sqrt_eps;
    p->k_crit = 1.0;
    p->k_feas = 0.1;
    p->beta_crit = 1.0;
    p->beta_feas = 0.9;
    p->max_outer_iter = 500;
    p->reserved = 0;
}

int bnl_create(int device, bnl_handle* out) {
    if (!out) return BNL_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return BNL_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BNL_ENODEV;
    if (prop.major < 10) return BNL_ENODEV;  // sm_100a cubin only: nothing else can run
    if (cudaSetDevice(device) != cudaSuccess) return BNL_ECUDA;
    S* h = new S();
    h->device = device;
    h->prop = prop;
    bnl_default_params(&h->prm);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&h->sd, sizeof(Scal)) != cudaSuccess ||
        cudaHostAlloc(&h->sh, sizeof(Scal), cudaHostAllocMapped) != cudaSuccess) {
        delete h;
        return BNL_ECUDA;
    }
    cudaMemset(h->sd, 0, sizeof(Scal));
    memset(h->sh, 0, sizeof(Scal));
    // group mailbox of the row reductions (p2p.h): local until bnl_comm_init maps it into the peers
    // (>= 4 MB: an allocation of its own, never a sub-allocation of a shared 2 MB block -- it is exported through CUDA IPC)
    if (cudaMalloc(&h->p2p_buf, p2p_alloc_bytes()) != cudaSuccess || cudaMemset(h->p2p_buf, 0, p2p_alloc_bytes()) != cudaSuccess ||
        cudaMalloc(&h->p2p_counter, 256) != cudaSuccess || cudaMemset(h->p2p_counter, 0, 256) != cudaSuccess) {
        bnl_destroy(h);
        return BNL_EOOM;
    }
    p2p_local_setup(h);
    {
        const char* env = getenv("BNL_CAUCHY");
        h->cauchy_mode = (env && env[0] == 'l') ? BNL_CAUCHY_LITERAL : BNL_CAUCHY_INCREMENTAL;
        const char* jt = getenv("BNL_JT");
        h->jt_disabled = jt && jt[0] == '0';
        const char* fj = getenv("BNL_FUSE_JTR");
        h->fuse_jtr = !(fj && fj[0] == '0');
        const char* gg = getenv("BNL_GRAM_GUARD");
        if (gg && atof(gg) > 0.0) h->gram_guard = atof(gg);
        const char* gd = getenv("BNL_CAUCHY_GUARD");
        if (gd && atof(gd) > 0.0) h->cauchy_guard = atof(gd);
        const char* rp = getenv("BNL_REUSE_POINT");
        h->reuse_point = !(rp && rp[0] == '0');
    }
    cudaEventCreate(&h->ev_t0);
    cudaEventCreate(&h->ev_t1);
    *out = h;
    return BNL_OK;
}

void bnl_destroy(bnl_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int r2 = 0; r2 < kP2PMaxRanks; ++r2)
        if (h->p2p_opened[r2]) cudaIpcCloseMemHandle(h->p2p_opened[r2]);
    cudaFree(h->p2p_buf);
    cudaFree(h->p2p_counter);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    free_problem(h);
    for (auto& e : h->ev_busy) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto& e : h->ev_free) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->pin) cudaFreeHost(h->pin);
    for (int b = 0; b < 2; ++b) {
        if (h->pin2[b]) cudaFreeHost(h->pin2[b]);
        if (h->pin2_ev[b]) cudaEventDestroy(h->pin2_ev[b]);
    }
    cudaFree(h->sd);
    cudaFreeHost(h->sh);
    cudaStreamDestroy(h->stream);
    delete h;
}

const char* bnl_last_error(bnl_handle h) { return h ? h->err.c_str() : "null handle"; }

int bnl_set_params(bnl_handle h, const bnl_params* p) {
    if (!valid(h) || !p) return BNL_EINVAL;
    // @assert (0 < eta1 <= eta2 < 1) && (0 < gamma1 < 1 < gamma2)   src/basic_tralcnlss.jl:200
    if (!((0 < p->eta1) && (p->eta1 <= p->eta2) && (p->eta2 < 1) && (0 < p->gamma1) && (p->gamma1 < 1) && (1 < p->gamma2)))
        return h->fail(BNL_EASSERT, "AssertionError: Invalid trust region updates paramaters");
    h->prm = *p;
    h->t0_valid = false;
    sync_params_to_ctx(h);
    return BNL_OK;
}

int bnl_comm_unique_id(void* id128) {
    if (!id128) return BNL_EINVAL;
    if (!g_nccl.load()) return BNL_ENCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return BNL_ENCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return BNL_OK;
}

static void comm_teardown(S* h) {
    cudaStreamSynchronize(h->stream);
    for (int r2 = 0; r2 < kP2PMaxRanks; ++r2) {
        if (h->p2p_opened[r2]) cudaIpcCloseMemHandle(h->p2p_opened[r2]);
        h->p2p_opened[r2] = nullptr;
    }
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    h->comm = nullptr;
    h->p2p_on = false;
    h->nranks = 1;
    h->rank = 0;
    p2p_local_setup(h);
}

int bnl_comm_init(bnl_handle h, int nranks, int rank, const void* id128) {
    if (!valid(h) || nranks < 1 || rank < 0 || rank >= nranks) return BNL_EINVAL;
    if (kGroups % nranks != 0 || nranks > kP2PMaxRanks)
        return h->fail(BNL_EINVAL, "nranks = %d: the row geometry has %d groups, nranks must be 1, 2, 4 or 8", nranks, kGroups);
    CK(cudaSetDevice(h->device));
    h->pc_valid = false;
    comm_teardown(h);  // a second call replaces the previous communicator / peer mappings
    h->comm_set = true;
    if (nranks == 1) {
        if (h->problem_set) {
            RET(resolve_geometry(h));
            RET(alloc_row_buffers(h));
        }
        return BNL_OK;
    }
    if (!id128) return BNL_EINVAL;
    if (!g_nccl.load()) return h->fail(BNL_ENCCL, "cannot dlopen libnccl.so.2");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (r != ncclSuccess) return h->fail(BNL_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    h->nranks = nranks;
    h->rank = rank;
    auto min_over_ranks = [&](int v) -> int {  // also a barrier
        int* dv = nullptr;
        if (cudaMalloc(&dv, sizeof(int)) != cudaSuccess) return 0;
        cudaMemcpy(dv, &v, sizeof(int), cudaMemcpyHostToDevice);
        g_nccl.AllReduce(dv, dv, 1, ncclInt, ncclMin, h->comm, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaMemcpy(&v, dv, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(dv);
        return v;
    };
    // ---- map every rank's mailbox into every peer (CUDA IPC; handles exchanged with ncclAllGather) ----
    const char* env = getenv("BNL_P2P_ALLREDUCE");
    // every rank must take the same branch BEFORE the all-gather below: agree on "want" first (min over ranks)
    const int want = min_over_ranks((!(env && env[0] == '0') && g_nccl.AllGather != nullptr) ? 1 : 0);
    int ok = want;
    if (want) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        char* dh = nullptr;
        std::vector<cudaIpcMemHandle_t> all(nranks);
        cudaIpcMemHandle_t mine;
        int mine_ok = cudaMemset(h->p2p_buf, 0, p2p_buffer_bytes()) == cudaSuccess && cudaMemset(h->p2p_counter, 0, 256) == cudaSuccess &&
                      cudaIpcGetMemHandle(&mine, h->p2p_buf) == cudaSuccess && cudaMalloc(&dh, (size_t)nranks * 64) == cudaSuccess;
        if (!mine_ok) memset(&mine, 0, sizeof mine);
        if (dh) {
            cudaMemcpy(dh + (size_t)rank * 64, &mine, 64, cudaMemcpyHostToDevice);
            // every rank enters the collective, whatever happened locally
            ok = g_nccl.AllGather(dh + (size_t)rank * 64, dh, 64, ncclChar, h->comm, h->stream) == ncclSuccess &&
                 cudaStreamSynchronize(h->stream) == cudaSuccess &&
                 cudaMemcpy(all.data(), dh, (size_t)nranks * 64, cudaMemcpyDeviceToHost) == cudaSuccess && mine_ok;
        } else {
            ok = 0;
        }
        ok = min_over_ranks(ok);  // nobody opens handles unless every rank exported one
        for (int r2 = 0; ok && r2 < nranks; ++r2) {
            void* base = h->p2p_buf;
            if (r2 != rank) {
                ok = cudaIpcOpenMemHandle(&base, all[r2], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                if (ok) h->p2p_opened[r2] = base;
            }
            h->p2p.mbox[r2] = static_cast<double*>(base);
            h->p2p.flag[r2] = p2p_flags_of(static_cast<double*>(base));
            h->p2p.ll[r2] = p2p_ll_of(static_cast<double*>(base));
        }
        cudaGetLastError();
        if (dh) cudaFree(dh);
    }
    // all ranks must agree (min over ranks); this all-reduce is also the barrier after everyone's memset
    ok = min_over_ranks(ok);
    h->p2p.nranks = nranks;
    h->p2p.rank = rank;
    h->p2p.mbox[rank] = h->p2p_buf;
    h->p2p.flag[rank] = p2p_flags_of(h->p2p_buf);
    h->p2p.ll[rank] = p2p_ll_of(h->p2p_buf);
    h->p2p_on = ok != 0;
    h->p2p_epoch = 0;
    h->ll_epoch = 0;
    if (h->problem_set) {
        RET(resolve_geometry(h));
        RET(alloc_row_buffers(h));
    }
    return BNL_OK;
}

int bnl_shard_rows(int64_t M_total, int32_t nranks, int32_t rank, int64_t* row0, int64_t* M_local) {
    RowGeom g{};
    if (!geom_make(M_total, nranks, rank, &g)) return BNL_EINVAL;
    if (row0) *row0 = g.row0;
    if (M_local) *M_local = g.local_rows();
    return BNL_OK;
}

int bnl_comm_info(bnl_handle h, int32_t* nranks, int32_t* rank, int32_t* p2p_allreduce) {
    if (!valid(h)) return BNL_EINVAL;
    if (nranks) *nranks = h->nranks;
    if (rank) *rank = h->rank;
    if (p2p_allreduce) *p2p_allreduce = h->p2p_on ? 1 : 0;
    return BNL_OK;
}

int bnl_set_problem(bnl_handle h, int64_t M_local, int64_t M_total, int64_t row0, int32_t n, int32_t m_lin, int32_t p,
                    const double* A, const double* xlow, const double* xupp) {
    if (!valid(h)) return BNL_EINVAL;
    if (M_local < 0 || M_total < M_local || row0 < 0 || n <= 0 || m_lin < 0 || p < 0 || m_lin > n)
        return h->fail(BNL_EDIM, "DimensionMismatch: M_local=%lld M_total=%lld n=%d m_lin=%d p=%d", (long long)M_local,
                       (long long)M_total, n, m_lin, p);
    if (m_lin > 0 && !A) return h->fail(BNL_EINVAL, "A is NULL but m_lin > 0");
    CK(cudaSetDevice(h->device));
    free_problem(h);
    h->M = M_local;
    h->M_total = M_total;
    h->row0 = row0;
    h->n = n;
    h->ld = pad_cols(n);
    h->m_lin = m_lin;
    h->p = p;
    h->mask = (m_lin == 0);
    size_t optin = h->prop.sharedMemPerBlockOptin;
    h->plan = mv_make_plan(n, optin);
    if (!h->plan.supported) return h->fail(BNL_EDIM, "n = %d unsupported by the streaming kernels (n <= 8192)", n);
    RET(resolve_geometry(h));
    const size_t ld = h->ld;
    const size_t vlen = ld + kColAlign;  // +16: slot [ld] carries ||Jv||^2
    const int nvec = 17;
    CK(cudaMalloc(&h->vecpool, nvec * vlen * sizeof(double)));
    CK(cudaMemset(h->vecpool, 0, nvec * vlen * sizeof(double)));
    double** slots[nvec] = {&h->vc.x,  &h->vc.g,  &h->vc.s,    &h->vc.d,    &h->vc.hv, &h->vc.r,  &h->vc.v,  &h->vc.pdir, &h->vc.w,
                            &h->vc.gm, &h->vc.xn, &h->vc.xlow, &h->vc.xupp, &h->vc.wl, &h->vc.wu, &h->vc.t1, &h->vc.t2};
    for (int i = 0; i < nvec; ++i) *slots[i] = h->vecpool + (size_t)i * vlen;
    CK(cudaMalloc(&h->flagpool, 2 * vlen));
    CK(cudaMemset(h->flagpool, 0, 2 * vlen));
    h->vc.fix = h->flagpool;
    h->vc.at = h->flagpool + vlen;
    h->vc.n = n;
    h->vc.ld = h->ld;
    h->vc.m_lin = m_lin;
    h->vc.p = p;
    h->vc.sd = h->sd;
    h->vc.sh = h->sh;
    h->vc.mu = 0.0;
    sync_params_to_ctx(h);
    const size_t pc = std::max(p, 1);
    double* cpool = nullptr;
    CK(cudaMalloc(&cpool, (2 * pc * ld + 2 * pc + 16) * sizeof(double)));
    CK(cudaMemset(cpool, 0, (2 * pc * ld + 2 * pc + 16) * sizeof(double)));
    h->vc.C = cpool;
    h->vc.muC = cpool + pc * ld;
    h->vc.cv = cpool + 2 * pc * ld;
    h->vc.pvec = h->vc.cv + pc;
    // (cpool is owned through vc.C: freed with the problem)
    RET(alloc_row_buffers(h));
    CK(cudaMalloc(&h->d_words, ((n + 63) / 64 + 1) * sizeof(unsigned long long)));
    CK(cudaMalloc(&h->d_idx, vlen * sizeof(long long)));
    CK(cudaMalloc(&h->d_count, sizeof(int)));
    CK(cudaMalloc(&h->J, std::max<size_t>((size_t)M_local * ld, 16) * sizeof(double)));
    CK(cudaMalloc(&h->r, std::max<size_t>(M_local, 16) * sizeof(double)));
    CK(cudaMalloc(&h->r_trial, std::max<size_t>(M_local, 16) * sizeof(double)));
    CK(cudaMemset(h->sd, 0, sizeof(Scal)));
    memset(h->sh, 0, sizeof(Scal));
    // bounds
    std::vector<double> lo(n, -kInf), up(n, kInf);
    if (xlow) std::copy(xlow, xlow + n, lo.begin());
    if (xupp) std::copy(xupp, xupp + n, up.begin());
    RET(put_vec(h, lo.data(), h->vc.xlow, n));
    RET(put_vec(h, up.data(), h->vc.xupp, n));
    h->h_xlow = lo;
    h->h_xupp = up;
    // linear equalities: A (column-major m_lin x n) -> row-major m_lin x ld; chol_aat = cholesky(A*A') :206
    if (m_lin > 0) {
        double* dA = nullptr;
        CK(cudaMalloc(&dA, (size_t)m_lin * ld * sizeof(double)));
        CK(cudaMemset(dA, 0, (size_t)m_lin * ld * sizeof(double)));
        h->dc.A = dA;  // owned by the handle from here on (freed with the problem even if a later step fails)
        RET(upload_colmajor(h, A, m_lin, n, m_lin, dA, h->ld));
        h->dc.n = n;
        h->dc.ld = h->ld;
        h->dc.m = m_lin;
        h->dc.cap = n;
        h->dc.A = dA;
        CK(cudaMalloc(&h->dc.LA, (size_t)m_lin * m_lin * sizeof(double)));
        CK(cudaMalloc(&h->dc.L, (size_t)n * n * sizeof(double)));
        CK(cudaMalloc(&h->dc.G, (size_t)m_lin * n * sizeof(double)));
        CK(cudaMalloc(&h->dc.ywork, (size_t)(n + 16) * sizeof(double)));
        CK(cudaMalloc(&h->dc.Lr, (size_t)m_lin * m_lin * sizeof(double)));
        {
            const char* env = getenv("BNL_LITERAL_PROJECTION");
            h->literal_proj = env && env[0] == '1';
        }
        CK(cudaMalloc(&h->dc.fixidx, (size_t)(n + 16) * sizeof(long long)));
        CK(cudaMalloc(&h->dc.q_dev, sizeof(int)));
        CK(cudaMemset(h->dc.q_dev, 0, sizeof(int)));
        h->dc.sd = h->sd;
        h->dc.sh = h->sh;
        dk_chol_aat(h->dc, h->stream);
        dk_rebuild(h->dc, h->vc.fix, h->stream);  // lincons.chol = chol_aat (no fixed variables yet)
        dk_rs_rebuild(h->dc, h->vc.fix, h->stream);
        RET(sync(h));
        if (h->sh->chol_fail) {
            cudaMemsetAsync(&h->sd->chol_fail, 0, sizeof(int), h->stream);
            return h->fail(BNL_ENOTPD, "PosDefException: cholesky(A*A') failed (basic_tralcnlss.jl:206)");
        }
    }
    h->problem_set = true;
    return BNL_OK;
}

int bnl_use_builtin_model(bnl_handle h, int32_t model_id, const double* params, int32_t nparams, uint32_t seed) {
    if (!valid(h) || !h->problem_set) return BNL_EINVAL;
    if (h->p > 1) return h->fail(BNL_EINVAL, "built-in models support at most one (built-in) nonlinear constraint");
    CK(cudaSetDevice(h->device));
    h->pc_valid = false;
    const int n = h->n;
    h->seed = seed;
    h->noise = (nparams > 0 && params) ? params[0] : 1e-3;
    h->cond_exp = (nparams > 1 && params) ? params[1] : 0.0;
    h->m_x0.assign(n, 0.0);
    h->m_xlow.assign(n, 0.0);
    h->m_xupp.assign(n, 0.0);
    h->m_xtrue.assign(n, 0.0);
    std::vector<double> cs(h->ld, 0.0);
    if (model_id == BNL_MODEL_GLM) {
        // oracle/models.py: glm_col_scale, glm_x_true
        for (int j = 0; j < n; ++j) {
            cs[j] = std::pow(10.0, -h->cond_exp * (double)j / (double)n) / std::sqrt((double)n);
            const double sgn = usym(hash_rc(rowkey(seed + 2u, 0ull), (uint32_t)j));
            h->m_xtrue[j] = (j % 10 == 0) ? (sgn < 0 ? -1.25 : 1.25) : 0.9 * sgn;
            h->m_xlow[j] = -1.0;
            h->m_xupp[j] = 1.0;
            h->m_x0[j] = 0.0;
        }
    } else if (model_id == BNL_MODEL_EXPSUM) {
        if (n % 2) return h->fail(BNL_EDIM, "EXPSUM needs an even n");
        const int C = n / 2;
        for (int c = 0; c < C; ++c) {
            h->m_xtrue[c] = 1.0 + u01(hash_rc(rowkey(seed, 0ull), (uint32_t)c));
            h->m_xtrue[C + c] = 0.5 + 3.0 * u01(hash_rc(rowkey(seed, 1ull), (uint32_t)c));
        }
        for (int j = 0; j < n; ++j) {
            const double xt = h->m_xtrue[j];
            if (j % 8 == 0) {
                h->m_xlow[j] = xt;
                h->m_xupp[j] = xt + 0.5;
            } else {
                h->m_xlow[j] = xt - 0.25;
                h->m_xupp[j] = xt + 0.25;
            }
            h->m_x0[j] = 0.5 * (h->m_xlow[j] + h->m_xupp[j]);
        }
    } else if (model_id == BNL_MODEL_EXPSUM_DENSE) {
        // oracle/models.py: DenseExpSumProblem (one K-term exponential sum on one time grid, SURVEY 8d cfg2 as worded)
        if (n % 2) return h->fail(BNL_EDIM, "EXPSUM_DENSE needs an even n");
        if (n > 4096) return h->fail(BNL_EDIM, "EXPSUM_DENSE supports n <= 4096");
        const int K = n / 2;
        for (int k = 0; k < K; ++k) {
            h->m_xtrue[k] = 1.0 + u01(hash_rc(rowkey(seed, 0ull), (uint32_t)k));
            h->m_xtrue[K + k] = 0.5 * (double)k + u01(hash_rc(rowkey(seed, 1ull), (uint32_t)k));
        }
        for (int j = 0; j < n; ++j) {
            const double xt = h->m_xtrue[j];
            if (j % 8 == 0) {
                h->m_xlow[j] = xt;
                h->m_xupp[j] = xt + 0.5;
            } else {
                h->m_xlow[j] = xt - 0.25;
                h->m_xupp[j] = xt + 0.25;
            }
            h->m_x0[j] = 0.5 * (h->m_xlow[j] + h->m_xupp[j]);
        }
    } else {
        return h->fail(BNL_EINVAL, "unknown builtin model %d", model_id);
    }
    h->model_id = model_id;
    h->cb_res = h->cb_jac = h->cb_nl = h->cb_jnl = nullptr;
    if (!h->d_cs) CK(cudaMalloc(&h->d_cs, (size_t)(h->ld + kColAlign) * sizeof(double)));
    if (!h->d_xtrue) CK(cudaMalloc(&h->d_xtrue, (size_t)(h->ld + kColAlign) * sizeof(double)));
    CK(cudaMemset(h->d_xtrue, 0, (size_t)(h->ld + kColAlign) * sizeof(double)));
    RET(put_vec(h, cs.data(), h->d_cs, h->ld));
    RET(put_vec(h, h->m_xtrue.data(), h->d_xtrue, n));
    if (!h->ydata) CK(cudaMalloc(&h->ydata, std::max<size_t>(h->M, 16) * sizeof(double)));
    CK(model_setup_y(margs(h), h->d_xtrue, h->ydata, h->stream));
    KLAUNCH();
    // the model's own box replaces whatever bnl_set_problem was given
    RET(put_vec(h, h->m_xlow.data(), h->vc.xlow, n));
    RET(put_vec(h, h->m_xupp.data(), h->vc.xupp, n));
    h->h_xlow = h->m_xlow;
    h->h_xupp = h->m_xupp;
    RET(sync(h));
    h->have_J = false;
    return BNL_OK;
}

int bnl_use_callbacks(bnl_handle h, bnl_callback residuals, bnl_callback jac_res, bnl_callback nlconstraints,
                      bnl_callback jac_nlcons, void* ctx) {
    if (!valid(h) || !h->problem_set || !residuals || !jac_res) return BNL_EINVAL;
    if (h->p > 0 && (!nlconstraints || !jac_nlcons)) return h->fail(BNL_EINVAL, "p > 0 needs nlconstraints and jac_nlcons");
    h->model_id = 0;
    h->pc_valid = false;
    h->cb_res = residuals;
    h->cb_jac = jac_res;
    h->cb_nl = nlconstraints;
    h->cb_jnl = jac_nlcons;
    h->cb_ctx = ctx;
    h->have_J = false;
    return BNL_OK;
}

int bnl_use_builtin_nlcons(bnl_handle h, int32_t kind, const double* params, int32_t nparams) {
    if (!valid(h) || !h->problem_set) return BNL_EINVAL;
    if (kind != BNL_NLCONS_SPHERE || nparams < 1 || !params) return h->fail(BNL_EINVAL, "unknown built-in nonlinear constraint");
    if (h->p != 1) return h->fail(BNL_EDIM, "the sphere constraint needs p == 1");
    h->nl_kind = kind;
    h->nl_rho2 = params[0];
    h->pc_valid = false;
    return BNL_OK;
}

int bnl_model_set_truth(bnl_handle h, const double* x_true, const double* x0) {
    if (!valid(h) || h->model_id == 0 || !x_true) return BNL_EINVAL;
    CK(cudaSetDevice(h->device));
    h->pc_valid = false;
    std::copy(x_true, x_true + h->n, h->m_xtrue.begin());
    if (x0) std::copy(x0, x0 + h->n, h->m_x0.begin());
    RET(put_vec(h, h->m_xtrue.data(), h->d_xtrue, h->n));
    CK(model_setup_y(margs(h), h->d_xtrue, h->ydata, h->stream));  // y = model(x_true) + noise
    KLAUNCH();
    RET(sync(h));
    h->have_J = false;
    return BNL_OK;
}

int bnl_model_vectors(bnl_handle h, double* x0, double* xlow, double* xupp, double* x_true) {
    if (!valid(h) || h->model_id == 0) return BNL_EINVAL;
    const size_t nb = (size_t)h->n * sizeof(double);
    if (x0) memcpy(x0, h->m_x0.data(), nb);
    if (xlow) memcpy(xlow, h->m_xlow.data(), nb);
    if (xupp) memcpy(xupp, h->m_xupp.data(), nb);
    if (x_true) memcpy(x_true, h->m_xtrue.data(), nb);
    return BNL_OK;
}

// Every entry point but the two solve calls forgets the point of the last solve's end (solver_internal.h: pc_valid): the reuse
// at a subproblem restart only ever spans consecutive solve_subproblem calls / the outer iterations of one bnl_tralcnllss.
#define ENTER()                                                          \
    if (!valid(h)) return BNL_EINVAL;                                    \
    if (!h->problem_set) return h->fail(BNL_EINVAL, "bnl_set_problem first"); \
    CK(cudaSetDevice(h->device));                                        \
    const bool pc_was_valid = h->pc_valid;                               \
    (void)pc_was_valid;                                                  \
    h->pc_valid = false;

int bnl_upload_jacobian(bnl_handle h, const double* J_colmajor, int64_t ldj) {
    ENTER();
    if (!J_colmajor || ldj < h->M) return h->fail(BNL_EDIM, "DimensionMismatch: ldj < M");
    RET(upload_colmajor(h, J_colmajor, h->M, h->n, ldj, h->J, h->ld));
    h->have_J = true;
    h->gram_valid = false;
    h->t0_valid = false;
    h->jt_valid = h->jt_attempted = false;
    return BNL_OK;
}

int bnl_upload_nlcons_jacobian(bnl_handle h, const double* C_colmajor, int64_t ldc) {
    ENTER();
    if (h->p == 0) return BNL_OK;
    if (!C_colmajor || ldc < h->p) return h->fail(BNL_EDIM, "DimensionMismatch: ldc < p");
    RET(upload_colmajor(h, C_colmajor, h->p, h->n, ldc, h->vc.C, h->ld));
    vk_scale_C(h->vc, h->stream);
    RET(sync(h));
    return BNL_OK;
}

int bnl_set_mu(bnl_handle h, double mu) {
    ENTER();
    h->vc.mu = mu;
    if (h->p > 0) vk_scale_C(h->vc, h->stream);
    RET(sync(h));
    return BNL_OK;
}

int bnl_eval_jacobian(bnl_handle h, const double* x) {
    ENTER();
    RET(put_vec(h, x, h->vc.t2, h->n));
    RET(eval_jacobian(h, h->vc.t2));
    if (h->p > 0) vk_scale_C(h->vc, h->stream);
    RET(sync(h));
    return BNL_OK;
}

int bnl_residuals(bnl_handle h, const double* x, double* r_local, double* sumsq) {
    ENTER();
    RET(put_vec(h, x, h->vc.t2, h->n));
    std::vector<double> cdummy;
    RET(eval_residual(h, h->vc.t2, h->r_trial, cdummy));
    vk_publish(h->sd, h->sh, h->stream);
    RET(sync(h));
    if (sumsq) *sumsq = h->sh->sumsq_r;
    if (r_local && h->M > 0) RET(get_vec(h, h->r_trial, r_local, h->M));
    return BNL_OK;
}

int bnl_nlcons(bnl_handle h, const double* x, double* c, double* C_colmajor) {
    ENTER();
    if (h->p == 0) return BNL_OK;
    if (h->model_id != 0) {
        if (h->nl_kind != BNL_NLCONS_SPHERE) return h->fail(BNL_EINVAL, "no built-in nonlinear constraint bound");
        RET(put_vec(h, x, h->vc.t2, h->n));
        vk_sphere_value(h->vc, h->vc.t2, h->nl_rho2, h->stream);
        RET(sync(h));
        if (c) c[0] = h->sh->c0;
        if (C_colmajor)
            for (int j = 0; j < h->n; ++j) C_colmajor[j] = 2.0 * x[j];  // p = 1: column-major 1 x n
    } else {
        if (!h->cb_nl || !h->cb_jnl) return h->fail(BNL_EINVAL, "no nonlinear-constraint callbacks bound");
        if (c && h->cb_nl(x, c, h->cb_ctx) != 0) return h->fail(BNL_ECALLBACK, "nlconstraints callback failed");
        if (C_colmajor && h->cb_jnl(x, C_colmajor, h->cb_ctx) != 0) return h->fail(BNL_ECALLBACK, "jac_nlcons callback failed");
    }
    return BNL_OK;
}

int bnl_gradient(bnl_handle h, const double* x, double* g) {  // jac_res(x)' * residuals(x)   :893
    ENTER();
    if (!x || !g) return BNL_EINVAL;
    RET(put_vec(h, x, h->vc.t2, h->n));
    std::vector<double> cdummy;
    RET(eval_residual(h, h->vc.t2, h->r_trial, cdummy));
    RET(eval_jacobian(h, h->vc.t2));
    RET(jtw_dev(h, h->r_trial, h->vc.hv));
    RET(sync(h));
    RET(get_vec(h, h->vc.hv, g, h->n));
    return BNL_OK;
}

int bnl_hess_mul(bnl_handle h, const double* v, double* Hv) {
    ENTER();
    if (!v || !Hv) return BNL_EINVAL;
    RET(put_vec(h, v, h->vc.t2, h->n));
    RET(hess_mul(h, h->vc.t2, h->vc.hv));
    RET(sync(h));
    RET(get_vec(h, h->vc.hv, Hv, h->n));
    return BNL_OK;
}

int bnl_vthv(bnl_handle h, const double* v, double* out) {
    ENTER();
    if (!v || !out) return BNL_EINVAL;
    RET(put_vec(h, v, h->vc.t2, h->n));
    RET(vthv_dev(h, h->vc.t2));
    vk_dot_gs(h->vc, h->stream);  // publishes jv_sumsq (and an unrelated g.s)
    RET(sync(h));
    *out = h->sh->jv_sumsq + h->vc.mu * h->sh->Cv_sumsq;
    return BNL_OK;
}

int bnl_jv(bnl_handle h, const double* v, double* Jv_local) {
    ENTER();
    if (!v) return BNL_EINVAL;
    if (!h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound");
    if (!h->tvec) CK(cudaMalloc(&h->tvec, std::max<size_t>(h->M, 16) * sizeof(double)));
    RET(put_vec(h, v, h->vc.t2, h->n));
    CK(mv_launch(MODE_JV, h->plan, h->geo, h->J, h->vc.t2, nullptr, h->tvec, h->partial, h->stream));
    RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, h->ld, h->ld + 1, h->vc.hv));
    RET(sync(h));
    h->st.jv++;
    if (Jv_local && h->M > 0) RET(get_vec(h, h->tvec, Jv_local, h->M));
    return BNL_OK;
}

int bnl_jtw(bnl_handle h, const double* w_local, double* JTw) {
    ENTER();
    if (!w_local || !JTw) return BNL_EINVAL;
    if (!h->tvec) CK(cudaMalloc(&h->tvec, std::max<size_t>(h->M, 16) * sizeof(double)));
    if (h->M > 0) RET(put_vec(h, w_local, h->tvec, h->M));
    RET(jtw_dev(h, h->tvec, h->vc.hv));
    RET(sync(h));
    RET(get_vec(h, h->vc.hv, JTw, h->n));
    return BNL_OK;
}

int bnl_gram(bnl_handle h, double* G_colmajor, double* ms) {
    ENTER();
    if (!h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound");
    const size_t ld = h->ld;
    const double before = h->st.gram_ms;
    RET(form_gram(h));
    RET(sync(h));
    if (ms) *ms = h->st.gram_ms - before;
    if (G_colmajor) {
        std::vector<double> tmp(ld * ld);
        RET(get_vec(h, h->gram, tmp.data(), ld * ld));
        for (int j = 0; j < h->n; ++j)
            for (int i = 0; i < h->n; ++i) G_colmajor[(size_t)j * h->n + i] = tmp[(size_t)i * ld + j];
    }
    return BNL_OK;
}

int bnl_set_cauchy_mode(bnl_handle h, int32_t mode) {
    if (!valid(h) || (mode != BNL_CAUCHY_LITERAL && mode != BNL_CAUCHY_INCREMENTAL)) return BNL_EINVAL;
    h->cauchy_mode = mode;
    return BNL_OK;
}

int bnl_set_hessian_mode(bnl_handle h, int32_t mode) {
    if (!valid(h) || (mode != BNL_HESSIAN_MATRIX_FREE && mode != BNL_HESSIAN_GRAM)) return BNL_EINVAL;
    h->hess_mode = mode;
    return BNL_OK;
}

int bnl_project(bnl_handle h, const double* r, double* v) {
    ENTER();
    if (!r || !v) return BNL_EINVAL;
    RET(put_vec(h, r, h->vc.t2, h->n));
    if (h->mask) {
        vk_mask_project(h->vc, h->vc.t2, h->vc.t1, h->stream);  // exact mask
        RET(sync(h));
        RET(get_vec(h, h->vc.t1, v, h->n));
    } else {
        RET(project_general(h, h->vc.t2, h->vc.t1, false));
        RET(sync(h));
        RET(get_vec(h, h->vc.t1, v, h->n));
    }
    return BNL_OK;
}

// left_mul(lincons, x) -> y (length m_lin + nb_fix)  and  left_mul_tr(lincons, y) -> x   (src/polyhedral_constraints.jl:72-98)
static int left_mul_common(bnl_handle h, const double* in, double* out, bool transpose) {
    if (!in || !out) return BNL_EINVAL;
    int q = 0;
    RET(bnl_get_fixvars(h, nullptr, &q));
    const int mpp = h->m_lin + q;
    std::vector<double> host(std::max(h->n, mpp), 0.0);
    if (h->mask) {  // no linear equalities: A~ = rows of the identity
        std::vector<uint64_t> words((h->n + 63) / 64);
        RET(bnl_get_fixvars(h, words.data(), &q));
        int k = 0;
        if (!transpose) {
            for (int i = 0; i < h->n; ++i)
                if ((words[i >> 6] >> (i & 63)) & 1ull) out[k++] = in[i];
        } else {
            for (int i = 0; i < h->n; ++i) out[i] = ((words[i >> 6] >> (i & 63)) & 1ull) ? in[k++] : 0.0;
        }
        return BNL_OK;
    }
    dk_rebuild(h->dc, h->vc.fix, h->stream);  // refresh the ascending index list of fixed variables
    if (!transpose) {
        RET(put_vec(h, in, h->vc.t2, h->n));
        dk_left_mul(h->dc, h->vc.t2, h->dc.ywork, h->stream);
        RET(sync(h));
        RET(get_vec(h, h->dc.ywork, out, mpp));
    } else {
        RET(put_vec(h, in, h->dc.ywork, mpp));
        dk_left_mul_tr(h->dc, h->dc.ywork, h->vc.t2, h->stream);
        RET(sync(h));
        RET(get_vec(h, h->vc.t2, out, h->n));
    }
    return BNL_OK;
}
int bnl_left_mul(bnl_handle h, const double* x, double* y) {
    ENTER();
    return left_mul_common(h, x, y, false);
}
int bnl_left_mul_tr(bnl_handle h, const double* y, double* x) {
    ENTER();
    return left_mul_common(h, y, x, true);
}

int bnl_active_bounds_reset(bnl_handle h, const double* x) {
    ENTER();
    RET(put_vec(h, x, h->vc.t2, h->n));
    vk_active_reset(h->vc, h->vc.t2, nullptr, h->stream);
    RET(rebuild_chol(h));
    RET(sync(h));
    RET(check_chol(h));
    return BNL_OK;
}

int bnl_active_bounds(bnl_handle h, const double* x, const double* s, double delta, int64_t* idx, int32_t* count) {
    ENTER();
    RET(put_vec(h, x, h->vc.t2, h->n));
    RET(put_vec(h, s, h->vc.t1, h->n));
    vk_active_flags(h->vc, h->vc.t2, h->vc.t1, delta, h->stream);
    vk_list_flags(h->vc.at, h->n, h->d_idx, h->d_count, h->stream);
    RET(sync(h));
    const int cnt = h->sh->n_at_bound;
    if (count) *count = cnt;
    if (idx && cnt > 0) {
        std::vector<long long> tmp(cnt);
        CK(cudaMemcpy(tmp.data(), h->d_idx, cnt * sizeof(long long), cudaMemcpyDeviceToHost));
        for (int i = 0; i < cnt; ++i) idx[i] = tmp[i];
    }
    return BNL_OK;
}

int bnl_add_active(bnl_handle h, const int64_t* idx, int32_t count) {
    ENTER();
    if (count < 0 || (count > 0 && !idx)) return BNL_EINVAL;
    for (int i = 0; i < count; ++i)
        if (idx[i] < 0 || idx[i] >= h->n) return h->fail(BNL_EBOUNDS, "BoundsError: add_active! index %lld", (long long)idx[i]);
    if (count > 0) {
        std::vector<long long> tmp(idx, idx + count);
        CK(cudaMemcpy(h->d_idx, tmp.data(), count * sizeof(long long), cudaMemcpyHostToDevice));
    }
    vk_set_flags(h->vc, h->d_idx, count, h->stream);
    RET(sync(h));
    if (h->m_lin + h->sh->nb_fix > h->n) return h->fail(BNL_EASSERT, "AssertionError: m + count(fixvars) <= n (polyhedral_constraints.jl:43)");
    RET(rebuild_chol(h));
    RET(sync(h));
    RET(check_chol(h));
    return BNL_OK;
}

int bnl_set_fixvars(bnl_handle h, const uint64_t* words) {
    ENTER();
    const int nw = (h->n + 63) / 64;
    CK(cudaMemcpy(h->d_words, words, nw * sizeof(uint64_t), cudaMemcpyHostToDevice));
    vk_unpack_fix(h->d_words, h->n, h->vc.fix, h->sd, h->sh, h->stream);
    RET(sync(h));
    if (h->m_lin + h->sh->nb_fix > h->n) return h->fail(BNL_EASSERT, "AssertionError: m + count(fixvars) <= n");
    RET(rebuild_chol(h));
    RET(sync(h));
    RET(check_chol(h));
    return BNL_OK;
}

int bnl_get_fixvars(bnl_handle h, uint64_t* words, int32_t* nb_fix) {
    ENTER();
    const int nw = (h->n + 63) / 64;
    vk_pack_fix(h->vc.fix, h->n, h->d_words, h->stream);
    vk_publish(h->sd, h->sh, h->stream);
    RET(sync(h));
    if (words) CK(cudaMemcpy(words, h->d_words, nw * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (nb_fix) {
        int c = 0;
        if (words)
            for (int i = 0; i < nw; ++i) c += __builtin_popcountll(words[i]);
        else
            c = h->sh->nb_fix;
        *nb_fix = c;
    }
    return BNL_OK;
}

int bnl_get_chol(bnl_handle h, double* L_colmajor, int32_t* dim) {
    ENTER();
    int q = 0;
    if (h->mask) {
        RET(bnl_get_fixvars(h, nullptr, &q));
        // m_lin == 0: the factor of A~A~' is exactly I_q -- nothing is stored
        if (dim) *dim = q;
        if (L_colmajor)
            for (int j = 0; j < q; ++j)
                for (int i = 0; i < q; ++i) L_colmajor[(size_t)j * q + i] = (i == j) ? 1.0 : 0.0;
        return BNL_OK;
    }
    // lincons.chol is only materialised on request: the solve path uses the reduced-space factor (dense.h)
    dk_rebuild(h->dc, h->vc.fix, h->stream);
    RET(sync(h));
    RET(check_chol(h));
    CK(cudaMemcpy(&q, h->dc.q_dev, sizeof(int), cudaMemcpyDeviceToHost));
    const int mpp = h->m_lin + q;
    if (dim) *dim = mpp;
    if (L_colmajor) {
        std::vector<double> tmp((size_t)h->dc.cap * mpp);
        CK(cudaMemcpy(tmp.data(), h->dc.L, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int j = 0; j < mpp; ++j)
            for (int i = 0; i < mpp; ++i) L_colmajor[(size_t)j * mpp + i] = (i >= j) ? tmp[(size_t)j * h->dc.cap + i] : 0.0;
    }
    return BNL_OK;
}

int bnl_cauchy_step(bnl_handle h, const double* x, const double* g, double delta, double* s_c) {
    ENTER();
    RET(put_vec(h, x, h->vc.x, h->n));
    RET(put_vec(h, g, h->vc.g, h->n));
    RET(cauchy_step(h, delta));
    RET(sync(h));
    RET(get_vec(h, h->vc.s, s_c, h->n));
    return BNL_OK;
}

int bnl_projected_cg(bnl_handle h, const double* x, const double* s, const double* g_minor, double delta, double* w,
                     int32_t* cg_status, int32_t* iters) {
    ENTER();
    RET(put_vec(h, x, h->vc.x, h->n));
    RET(put_vec(h, s, h->vc.s, h->n));
    RET(put_vec(h, g_minor, h->vc.gm, h->n));
    int status = 0, it = 0;
    RET(minor_iterate(h, delta, &status, &it, false, false));
    RET(sync(h));
    RET(get_vec(h, h->vc.w, w, h->n));
    if (cg_status) *cg_status = status;
    if (iters) *iters = it;
    return BNL_OK;
}

// projected_cg(g_minor, H, w_l, w_u, lincons, kappa2) with the CALLER's step bounds (src/basic_tralcnlss.jl:690-697): reaches
// the alpha > gamma / bound_hit branch (:735-737) that minor_iterate's own bounds (trap T1) never trigger.
int bnl_projected_cg_bounds(bnl_handle h, const double* g_minor, const double* w_l, const double* w_u, double* w,
                            int32_t* cg_status, int32_t* iters) {
    ENTER();
    if (!g_minor || !w_l || !w_u || !w) return BNL_EINVAL;
    RET(put_vec(h, g_minor, h->vc.gm, h->n));
    RET(put_vec(h, w_l, h->vc.wl, h->n));
    RET(put_vec(h, w_u, h->vc.wu, h->n));
    int status = 0, it = 0;
    RET(minor_iterate(h, 0.0, &status, &it, false, true));
    RET(sync(h));
    RET(get_vec(h, h->vc.w, w, h->n));
    if (cg_status) *cg_status = status;
    if (iters) *iters = it;
    return BNL_OK;
}

// linesearch(g_model, H, w, w_l, w_u, fix_bounds) (src/basic_tralcnlss.jl:766-791) with the current fixvars -> alpha
int bnl_linesearch(bnl_handle h, const double* g_model, const double* w, const double* w_l, const double* w_u, double* alpha) {
    ENTER();
    if (!g_model || !w || !w_l || !w_u || !alpha) return BNL_EINVAL;
    RET(put_vec(h, g_model, h->vc.gm, h->n));
    RET(put_vec(h, w, h->vc.w, h->n));
    RET(put_vec(h, w_l, h->vc.wl, h->n));
    RET(put_vec(h, w_u, h->vc.wu, h->n));
    CK(cudaMemsetAsync(&h->sd->cg_neg_curv, 0, sizeof(int), h->stream));
    RET(vthv_dev(h, h->vc.w));          // :775
    vk_minor_finish(h->vc, h->stream);  // alpha = min(alpha_opt, alpha_allowed) :776-790 (also scales w and adds it to s: unused here)
    RET(sync(h));
    *alpha = h->sh->alpha_ls;
    return BNL_OK;
}

int bnl_inner_step(bnl_handle h, const double* x, const double* g, double delta, double* s, double* pred) {
    ENTER();
    h->t0_valid = false;
    RET(put_vec(h, x, h->vc.x, h->n));
    RET(put_vec(h, g, h->vc.g, h->n));
    double pr = 0.0;
    RET(inner_step(h, delta, &pr));
    RET(get_vec(h, h->vc.s, s, h->n));
    if (pred) *pred = pr;
    return BNL_OK;
}

int bnl_new_point(bnl_handle h, const double* x, const double* y, double mu, double* mx, double* g, double* cx) {
    ENTER();
    RET(put_vec(h, x, h->vc.x, h->n));
    std::vector<double> yv(h->p, 0.0);
    if (h->p > 0 && y) std::copy(y, y + h->p, yv.begin());
    double m = 0.0;
    RET(new_point(h, yv, mu, &m));
    if (mx) *mx = m;
    if (g) RET(get_vec(h, h->vc.g, g, h->n));
    if (cx && h->p > 0) std::copy(h->h_cx.begin(), h->h_cx.end(), cx);
    return BNL_OK;
}

static int solve_subproblem_host(bnl_handle h, const double* x0, const double* y, double mu, double omega_tol, double* x,
                                 double* cx, double* pix, FILE* log) {
    if (!x0) return BNL_EINVAL;
    h->t0_carry = h->t0_valid;
    h->t0_valid = false;
    cudaEvent_t e0 = h->ev_t0, e1 = h->ev_t1;
    RET(put_vec(h, x0, h->vc.x, h->n));
    CK(cudaEventRecord(e0, h->stream));
    std::vector<double> yv(h->p, 0.0);
    if (h->p > 0 && y) std::copy(y, y + h->p, yv.begin());
    double px = kInf;
    int rc = solve_subproblem_dev(h, yv, mu, omega_tol, &px, log, x0);
    cudaEventRecord(e1, h->stream);
    cudaStreamSynchronize(h->stream);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    h->st.solve_ms += t;
    if (rc != BNL_OK) return rc;
    if (h->reuse_point && h->model_id != 0) {
        // the loop leaves r = residuals(x), J = jac_res(x), d_jtr = J'r and cx of its final x in place (a rejected step touches
        // none of them, an accepted one refreshes all): remember which x that is, bit for bit
        h->pc_x.resize(h->n);
        RET(get_vec(h, h->vc.x, h->pc_x.data(), h->n));
        h->pc_cx = h->h_cx;
        h->pc_sumsq = h->acc_sumsq;
        h->pc_valid = true;
        if (x) std::copy(h->pc_x.begin(), h->pc_x.end(), x);
    } else if (x) {
        RET(get_vec(h, h->vc.x, x, h->n));
    }
    if (cx && h->p > 0) std::copy(h->h_cx.begin(), h->h_cx.end(), cx);
    if (pix) *pix = px;
    return BNL_OK;
}

int bnl_solve_subproblem(bnl_handle h, const double* x0, const double* y, double mu, double omega_tol, double* x,
                         double* cx, double* pix) {
    ENTER();
    h->pc_valid = pc_was_valid;
    return solve_subproblem_host(h, x0, y, mu, omega_tol, x, cx, pix, nullptr);
}

int bnl_get_stats(bnl_handle h, bnl_stats* out) {
    if (!valid(h) || !out) return BNL_EINVAL;
    *out = h->st;
    return BNL_OK;
}
int bnl_reset_stats(bnl_handle h) {
    if (!valid(h)) return BNL_EINVAL;
    h->st = bnl_stats{};
    h->ilog.clear();
    return BNL_OK;
}
int bnl_get_inner_log(bnl_handle h, bnl_inner_record* out, int32_t capacity, int32_t* count) {
    if (!valid(h)) return BNL_EINVAL;
    const int nrec = (int)h->ilog.size();
    if (count) *count = nrec;
    if (out)
        for (int i = 0; i < std::min(nrec, capacity); ++i) out[i] = h->ilog[i];
    return BNL_OK;
}

int bnl_device_info(bnl_handle h, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* free_bytes,
                    int64_t* total_bytes) {
    if (!valid(h)) return BNL_EINVAL;
    CK(cudaSetDevice(h->device));
    size_t f = 0, t = 0;
    CK(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = h->prop.multiProcessorCount;
    if (cc_major) *cc_major = h->prop.major;
    if (cc_minor) *cc_minor = h->prop.minor;
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return BNL_OK;
}

int bnl_time_kernel(bnl_handle h, int32_t kind, int32_t reps, double* avg_ms, double* bytes_per_launch) {
    ENTER();
    if (reps < 1) return BNL_EINVAL;
    if (kind <= 2 && !h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound");
    if ((kind == 3 || kind == 4 || kind == 6) && h->model_id == 0) return h->fail(BNL_EINVAL, "builtin model needed");
    if (kind == 5 && !h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound");
    const double Jbytes = 8.0 * (double)h->M * (double)h->ld;
    double bytes = 0.0;
    cudaEvent_t e0 = h->ev_t0, e1 = h->ev_t1;
    // the vectors used: x as v (any data), r as w
    for (int rep = -1; rep < reps; ++rep) {  // one untimed warm-up
        if (rep == 0) CK(cudaEventRecord(e0, h->stream));
        switch (kind) {
            case 0:
                CK(mv_launch(MODE_JTJV, h->plan, h->geo, h->J, h->vc.x, nullptr, nullptr, h->partial, h->stream));
                RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, 0, h->ld + 1, h->vc.t1));
                bytes = Jbytes + 16.0 * h->n;
                break;
            case 1:
                CK(mv_launch(MODE_JV, h->plan, h->geo, h->J, h->vc.x, nullptr, nullptr, h->partial, h->stream));
                RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, h->ld, h->ld + 1, h->vc.t1));
                bytes = Jbytes + 8.0 * h->n;
                break;
            case 2:
                CK(mv_launch(MODE_JTW, h->plan, h->geo, h->J, nullptr, h->r, nullptr, h->partial, h->stream));
                RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, 0, h->ld, h->vc.t1));
                bytes = Jbytes + 8.0 * h->M + 8.0 * h->n;
                break;
            case 3:
                CK(model_residual(margs(h), h->vc.x, h->ydata, h->r_trial, h->rpartial, h->stream));
                RET(row_reduce(h, h->rpartial, 1, 1, 0, 1, &h->sd->sumsq_r));
                bytes = 16.0 * h->M;
                break;
            case 4:
                CK(model_jacobian(margs(h), h->vc.x, h->J, h->stream));
                bytes = Jbytes;
                break;
            case 6:  // Jacobian generation fused with J'r (GLM, ld <= 1024)
                if (!(h->model_id == BNL_MODEL_GLM && h->plan.warp_team && h->plan.T == 8)) return h->fail(BNL_EINVAL, "fused generator: GLM with ld <= 1024");
                CK(model_jacobian_jtr(margs(h), h->vc.x, h->r, h->J, h->partial, h->plan.pstride, h->plan.KCH, h->plan.RB, h->stream));
                RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, 0, h->ld, h->vc.t1));
                bytes = Jbytes;
                break;
            case 5:
                if (!h->gram) RET(form_gram(h));
                CK(gram_launch(h->J, h->M, h->ld, h->gram, h->gram_ws, h->gram_nsplit, h->stream));
                bytes = gram_flops(h->M, h->ld);  // FLOPs, not bytes, for this kind
                break;
            default: return h->fail(BNL_EINVAL, "kind");
        }
    }
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    if (avg_ms) *avg_ms = (double)t / reps;
    if (bytes_per_launch) *bytes_per_launch = bytes;
    return BNL_OK;
}

// ---- benlsip.out log format (src/misc.jl:1-80): Julia's Printf prints non-finite values as NaN / Inf / -Inf ----------------
static std::string jl_e(int prec, double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Inf" : "-Inf";
    char buf[64];
    snprintf(buf, sizeof buf, "%.*e", prec, v);
    return buf;
}
static void log_header(FILE* io, S* h, const bnl_outer_params& op) {  // print_tralcnllss_header, src/misc.jl:1-45
    const std::string stars(64, '*'), blank = "*" + std::string(62, ' ') + "*";
    int nlow = 0, nupp = 0;
    for (int i = 0; i < h->n; ++i) {
        nlow += std::isfinite(h->h_xlow[i]) ? 1 : 0;
        nupp += std::isfinite(h->h_xupp[i]) ? 1 : 0;
    }
    fprintf(io, "\n\n%s\n%s\n", stars.c_str(), blank.c_str());
    fprintf(io, "*%sBEnlsip.jl v-DEV%s*\n%s\n", std::string(23, ' ').c_str(), std::string(23, ' ').c_str(), blank.c_str());
    fprintf(io, "*                   Better version of ENLSIP                   *\n%s\n%s\n", blank.c_str(), stars.c_str());
    fprintf(io, "\nProblem dimensions\n");
    fprintf(io, "Number of parameters.................: %5i\n", h->n);
    fprintf(io, "Number of residuals..................: %5lli\n", (long long)h->M_total);
    fprintf(io, "Number of nonlinear constraints......: %5i\n", h->p);
    fprintf(io, "Number of linear constraints.........: %5i\n", h->m_lin);
    fprintf(io, "Number of lower bounds...............: %5i\n", nlow);
    fprintf(io, "Number of upper bounds...............: %5i\n", nupp);
    fprintf(io, "\nAlgorithm parameters\n");
    // the reference passes (feas_tol, crit_tol) into the (crit_tol, feas_tol) slots (src/basic_tralcnlss.jl:213-226): kept
    fprintf(io, "Optimality tolerance.................................: %.6e\n", op.feas_tol);
    fprintf(io, "Nonlinear constraints feasibility tolerance..........: %.6e\n", op.crit_tol);
    fprintf(io, "Increase penalty parameter factor....................: %5f\n", op.tau);
    fprintf(io, "Step acceptance treshold.............................: %5f\n", h->prm.eta1);
    fprintf(io, "Great step acceptance treshold.......................: %5f\n", h->prm.eta2);
    fprintf(io, "Trust region increase factor.........................: %5f\n", h->prm.gamma2);
    fprintf(io, "Trust region decrease factor.........................: %5f\n", h->prm.gamma1);
    fprintf(io, "\n\n\n");
}
static void log_outer(FILE* io, int k, double objective, double nl_feas, double mu, double pix, double omega, bool first) {
    // print_outer_iter_header, src/misc.jl:47-68
    const std::string bar(80, '=');
    fprintf(io, "\n%s\n                          Outer iter %d\n  objective    nl feasibility     \xce\xbc      criticality   tolerance\n",
            bar.c_str(), k);
    if (first)
        fprintf(io, "%s   %s  %s        -         %s", jl_e(7, objective).c_str(), jl_e(6, nl_feas).c_str(), jl_e(2, mu).c_str(),
                jl_e(2, omega).c_str());
    else
        fprintf(io, "%s   %s  %s     %s     %s", jl_e(7, objective).c_str(), jl_e(6, nl_feas).c_str(), jl_e(2, mu).c_str(),
                jl_e(2, pix).c_str(), jl_e(2, omega).c_str());
    fprintf(io, "\n%s\niter     AL value       ||s||        \xce\x94          \xcf\x81\n", bar.c_str());
}

// ---- tralcnllss :167-298 (outer loop inside the library) ---------------------------------
// Every rank of a sharded run must make the same call; log_path may differ per rank (typically non-NULL on rank 0 only): the
// objective of the log line (:292) is evaluated on every rank regardless, as the reference does, so the collective sequence
// never depends on who logs.
int bnl_tralcnllss(bnl_handle h, const double* x0, const bnl_outer_params* op_in, const char* log_path, double* x_out,
                   double* y_out, double* final_mu, double* final_pix) {
    ENTER();
    if (!x0) return h->fail(BNL_EINVAL, "x0 is NULL");
    bnl_outer_params op;
    if (op_in)
        op = *op_in;
    else
        bnl_default_outer_params(&op);
    FILE* log = nullptr;
    if (log_path) {
        log = fopen(log_path, "w");
        if (!log) return h->fail(BNL_EINVAL, "cannot open log file %s", log_path);
    }
    const int n = h->n, p = h->p;
    std::vector<double> x(x0, x0 + n), y(p, 0.0), cx(p, 0.0), xn(n), cxn(p, 0.0);
    double mu = op.mu0;
    double omega = op.omega0 / std::pow(mu, op.k_crit), eta = op.eta0 / std::pow(mu, op.k_feas);  // :153-163
    int rc = BNL_OK;
    // rx = residuals(x); cx = nlconstraints(x)  :209-210
    double ss0 = 0.0;
    rc = bnl_residuals(h, x.data(), nullptr, &ss0);
    if (rc == BNL_OK && p > 0) rc = bnl_nlcons(h, x.data(), cx.data(), nullptr);
    if (log && rc == BNL_OK) log_header(log, h, op);  // :213-226
    // least_squares_multipliers :887-903 : y = -(CC')^{-1} C J'r  (p small: host arithmetic on device-computed J'r and C)
    if (p > 0 && rc == BNL_OK) {
        std::vector<double> g(n);
        // g = jac_res(x)' * residuals(x)   :893
        rc = put_vec(h, x.data(), h->vc.x, n);
        if (rc == BNL_OK) rc = eval_residual(h, h->vc.x, h->r, cx);
        if (rc == BNL_OK) rc = eval_jacobian(h, h->vc.x);
        if (rc == BNL_OK) rc = jtw_dev(h, h->r, h->vc.hv);
        if (rc == BNL_OK) rc = sync(h);
        if (rc == BNL_OK) rc = get_vec(h, h->vc.hv, g.data(), n);
        std::vector<double> Cjac((size_t)p * n);
        if (rc == BNL_OK) rc = bnl_nlcons(h, x.data(), nullptr, Cjac.data());
        if (rc == BNL_OK) {
            const double* C = Cjac.data();  // column-major p x n
            std::vector<double> CCt((size_t)p * p, 0.0), b(p, 0.0);
            for (int i = 0; i < p; ++i) {
                for (int j = 0; j < p; ++j) {
                    double s = 0.0;
                    for (int k = 0; k < n; ++k) s += C[(size_t)k * p + i] * C[(size_t)k * p + j];
                    CCt[(size_t)j * p + i] = s;
                }
                double s = 0.0;
                for (int k = 0; k < n; ++k) s += C[(size_t)k * p + i] * g[k];
                b[i] = -s;
            }
            for (int k = 0; k < p && rc == BNL_OK; ++k) {  // Cholesky
                double d = CCt[(size_t)k * p + k];
                for (int t = 0; t < k; ++t) d -= CCt[(size_t)t * p + k] * CCt[(size_t)t * p + k];
                if (!(d > 0)) {
                    rc = h->fail(BNL_ENOTPD, "PosDefException: cholesky(C*C') (basic_tralcnlss.jl:895)");
                    break;
                }
                CCt[(size_t)k * p + k] = std::sqrt(d);
                for (int i = k + 1; i < p; ++i) {
                    double s = CCt[(size_t)k * p + i];
                    for (int t = 0; t < k; ++t) s -= CCt[(size_t)t * p + i] * CCt[(size_t)t * p + k];
                    CCt[(size_t)k * p + i] = s / CCt[(size_t)k * p + k];
                }
            }
            if (rc == BNL_OK) {
                for (int i = 0; i < p; ++i) {
                    double s = b[i];
                    for (int t = 0; t < i; ++t) s -= CCt[(size_t)t * p + i] * y[t];
                    y[i] = s / CCt[(size_t)i * p + i];
                }
                for (int i = p - 1; i >= 0; --i) {
                    double s = y[i];
                    for (int t = i + 1; t < p; ++t) s -= CCt[(size_t)i * p + t] * y[t];
                    y[i] = s / CCt[(size_t)i * p + i];
                }
            }
        }
    }
    // MixedConstraints(A, chol_aat; l, u) :231 -- fixvars .= false
    if (rc == BNL_OK) {
        std::vector<uint64_t> zero((n + 63) / 64, 0);
        rc = bnl_set_fixvars(h, zero.data());
    }
    bool first_order_critical = false;
    int outer_iter = 1;
    double pix = kInf;
    if (log && rc == BNL_OK) {  // :237-245
        double nc = 0.0;
        for (double v : cx) nc += v * v;
        log_outer(log, outer_iter, ss0, std::sqrt(nc), mu, 0.0, omega, true);
    }
    while (rc == BNL_OK && !first_order_critical && outer_iter <= op.max_outer_iter) {  // :246
        rc = solve_subproblem_host(h, x.data(), y.data(), mu, omega, xn.data(), cxn.data(), &pix, log);
        if (rc != BNL_OK) break;
        double feas = 0.0;
        for (double v : cxn) feas += v * v;
        feas = std::sqrt(feas);
        if (feas <= eta) {  // :273
            x = xn;
            cx = cxn;
            first_order_critical = (pix <= op.crit_tol) && (feas <= op.feas_tol);
            if (!first_order_critical) {
                for (int i = 0; i < p; ++i) y[i] = y[i] + mu * cx[i];  // first_order_multipliers :905-911
                omega /= std::pow(mu, op.beta_crit);
                eta /= std::pow(mu, op.beta_feas);
            }
        } else {  // :284-289
            mu *= op.tau;
            omega = op.omega0 / std::pow(mu, op.k_crit);
            eta = op.eta0 / std::pow(mu, op.k_feas);
        }
        ++outer_iter;
        h->st.outer_iters++;
        double ss = 0.0;
        if (point_hit(h, x.data()))  // x is the point the subproblem just ended at: its dot(rx,rx) is at hand
            ss = h->pc_sumsq;
        else
            rc = bnl_residuals(h, x.data(), nullptr, &ss);  // objective = dot(rx,rx) :292 -- on every rank, logging or not
        if (rc != BNL_OK) break;
        if (log) log_outer(log, outer_iter, ss, feas, mu, pix, omega, false);  // :293
    }
    if (log) fclose(log);
    if (rc != BNL_OK) return rc;
    if (x_out) std::copy(x.begin(), x.end(), x_out);
    if (y_out && p > 0) std::copy(y.begin(), y.end(), y_out);
    if (final_mu) *final_mu = mu;
    if (final_pix) *final_pix = pix;
    return BNL_OK;
}

}  // extern "C"
