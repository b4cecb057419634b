// solver.cu -- host control flow of the inner Gauss-Newton trust-region subproblem solve and the C ABI
// (include/benlsip_b200.h).  The control flow mirrors, function by function, the reference's
// src/basic_tralcnlss.jl (solve_subproblem :303-378, inner_step :394-460, cauchy_step :574-639,
// minor_iterate :649-675, projected_cg :690-764, linesearch :766-791) -- every numeric operation runs in a
// CUDA kernel; the host only branches on scalars the kernels publish to pinned mapped memory.
// There is no CPU arithmetic path and no fallback: without an sm_100 device bnl_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/benlsip_b200.h"
#include "common.cuh"
#include "dense.h"
#include "gram.h"
#include "matvec.h"
#include "models.h"
#include "p2p.h"
#include "vecops.h"

using namespace bnl;

namespace {

// ---- NCCL through dlopen: no link-time dependency; a single-GPU user never loads it ---------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
    }
};
NcclApi g_nccl;

constexpr double kInf = std::numeric_limits<double>::infinity();

}  // namespace

struct EvPair {
    cudaEvent_t a, b;
    int cls;
};

struct bnl_solver {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string err;
    bnl_params prm{};
    bool problem_set = false;

    long long M = 0, M_total = 0, row0 = 0;
    int n = 0, ld = 0, m_lin = 0, p = 0;
    bool mask = true;
    bool literal_proj = false;  // BNL_LITERAL_PROJECTION=1: the reference's block factor on the solve path too

    double *J = nullptr, *r = nullptr, *r_trial = nullptr, *ydata = nullptr, *tvec = nullptr;
    double* vecpool = nullptr;
    unsigned char* flagpool = nullptr;
    VecCtx vc{};
    DenseCtx dc{};
    double* partial = nullptr;
    double* sumsq_partial = nullptr;
    int sumsq_blocks = 148 * 8;
    unsigned long long* d_words = nullptr;
    long long* d_idx = nullptr;
    int* d_count = nullptr;
    Scal *sd = nullptr, *sh = nullptr;
    MvPlan plan{};
    double* gram = nullptr;     // G = J'J (ld x ld), all-reduced
    double* gram_ws = nullptr;  // split-K workspace
    int gram_nsplit = 0;
    int hess_mode = 0;          // BNL_HESSIAN_MATRIX_FREE / BNL_HESSIAN_GRAM
    bool gram_valid = false;

    // model binding
    int model_id = 0;
    uint32_t seed = 0;
    double noise = 0.0, cond_exp = 0.0;
    double* d_cs = nullptr;
    double* d_xtrue = nullptr;
    std::vector<double> m_x0, m_xlow, m_xupp, m_xtrue;
    bnl_callback cb_res = nullptr, cb_jac = nullptr, cb_nl = nullptr, cb_jnl = nullptr;
    void* cb_ctx = nullptr;
    bool have_J = false;
    int nl_kind = 0;       // built-in nonlinear constraint (BNL_NLCONS_*), 0 = none / callbacks
    double nl_rho2 = 0.0;

    // host staging
    double* pin = nullptr;
    size_t pin_doubles = 0;
    double* pin2[2] = {nullptr, nullptr};  // double-buffered staging for matrix uploads
    cudaEvent_t pin2_ev[2] = {nullptr, nullptr};
    std::vector<double> h_x, h_cx, h_cx_next, h_ybar, h_tmp;

    // comm
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory all-reduce (p2p.h)
    bool p2p_on = false;
    P2PArgs p2p{};
    double* p2p_buf = nullptr;
    unsigned int* p2p_counter = nullptr;
    unsigned long long p2p_epoch = 0;
    void* p2p_opened[kP2PMaxRanks] = {nullptr};

    bnl_stats st{};
    std::vector<bnl_inner_record> ilog;
    std::vector<EvPair> ev_busy, ev_free;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // handle-owned pair for whole-call timings (no leak on error paths)

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return h->fail(e_ == cudaErrorMemoryAllocation ? BNL_EOOM : BNL_ECUDA, "%s:%d %s: %s", __FILE__, \
                           __LINE__, #call, cudaGetErrorString(e_));                                         \
    } while (0)
#define RET(call)                  \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != BNL_OK) return rc_; \
    } while (0)
#define KLAUNCH() (h->st.kernel_launches++)

namespace {

typedef bnl_solver S;

int sync(S* h) {
    if (h->p2p_on) vk_publish(h->sd, h->sh, h->stream);  // make a peer-wait timeout visible even when no O(n) kernel followed
    CK(cudaStreamSynchronize(h->stream));
    if (h->p2p_on && h->sh->p2p_timeout) return h->fail(BNL_ENCCL, "peer-memory all-reduce timed out waiting for a rank");
    // harvest finished event pairs
    for (size_t i = 0; i < h->ev_busy.size();) {
        EvPair& e = h->ev_busy[i];
        if (cudaEventQuery(e.b) == cudaSuccess) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e.a, e.b);
            switch (e.cls) {
                case 0: h->st.hess_mul_ms += ms; break;
                case 1: h->st.vthv_ms += ms; break;
                case 2: h->st.jtw_ms += ms; break;
                case 3: h->st.res_eval_ms += ms; break;
                case 4: h->st.jac_eval_ms += ms; break;
                case 5: h->st.gram_ms += ms; break;
            }
            h->ev_free.push_back(e);
            h->ev_busy[i] = h->ev_busy.back();
            h->ev_busy.pop_back();
        } else {
            ++i;
        }
    }
    return BNL_OK;
}

struct EvScope {  // records a CUDA-event pair around a kernel class on the launching stream
    S* h;
    EvPair e{};
    bool ok = false;
    EvScope(S* h_, int cls) : h(h_) {
        if (!h->ev_free.empty()) {
            e = h->ev_free.back();
            h->ev_free.pop_back();
            ok = true;
        } else if (h->ev_busy.size() < 4096) {
            ok = cudaEventCreate(&e.a) == cudaSuccess && cudaEventCreate(&e.b) == cudaSuccess;
        }
        e.cls = cls;
        if (ok) cudaEventRecord(e.a, h->stream);
    }
    ~EvScope() {
        if (ok) {
            cudaEventRecord(e.b, h->stream);
            h->ev_busy.push_back(e);
        }
    }
};

int ensure_pin(S* h, size_t doubles) {
    if (h->pin_doubles >= doubles) return BNL_OK;
    if (h->pin) cudaFreeHost(h->pin);
    h->pin = nullptr;
    h->pin_doubles = 0;
    CK(cudaHostAlloc(&h->pin, doubles * sizeof(double), cudaHostAllocDefault));
    h->pin_doubles = doubles;
    return BNL_OK;
}

// host n-vector -> device vector (through pinned staging; sync so the staging buffer can be reused)
int put_vec(S* h, const double* src, double* dst, size_t count) {
    RET(ensure_pin(h, std::max<size_t>(count, 1 << 16)));
    memcpy(h->pin, src, count * sizeof(double));
    CK(cudaMemcpyAsync(dst, h->pin, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BNL_OK;
}
int get_vec(S* h, const double* src_dev, double* dst, size_t count) {
    RET(ensure_pin(h, std::max<size_t>(count, 1 << 16)));
    CK(cudaMemcpyAsync(h->pin, src_dev, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(dst, h->pin, count * sizeof(double));
    return BNL_OK;
}

int allreduce(S* h, double* buf, size_t count) {
    if (h->nranks <= 1) return BNL_OK;
    if (h->p2p_on && count <= (size_t)kP2PWidth) {
        CK(p2p_allreduce(h->p2p, ++h->p2p_epoch, buf, (int)count, h->stream));
        h->st.kernel_launches += 2;
        h->st.allreduces++;
        h->st.p2p_allreduces++;
        return BNL_OK;
    }
    ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess) return h->fail(BNL_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    h->st.allreduces++;
    return BNL_OK;
}

// ---- Gram mode (K12): G = J'J on the FP64 tensor cores, all-reduced over the row shards ---------------------
int form_gram(S* h) {
    const size_t ld = h->ld;
    if (!h->gram) {
        h->gram_nsplit = gram_pick_split(h->M, h->ld, h->prop.multiProcessorCount);
        CK(cudaMalloc(&h->gram, ld * ld * sizeof(double)));
        CK(cudaMalloc(&h->gram_ws, (size_t)h->gram_nsplit * ld * ld * sizeof(double)));
    }
    {
        EvScope ev(h, 5);
        CK(gram_launch(h->J, h->M, h->ld, h->gram, h->gram_ws, h->gram_nsplit, h->stream));
    }
    h->st.kernel_launches += 2;
    h->st.j_passes += 1;  // every J element is staged once per 128-column tile pair from L2; HBM sees ~1 pass per tile row
    RET(allreduce(h, h->gram, ld * ld));
    h->st.gram_count++;
    h->gram_valid = true;
    return BNL_OK;
}

// ---- AlHessian -------------------------------------------------------------------------------------------
// Base.:*(H,v) :102-106.  dv: device, length ld.  out: device, length >= ld+1 (out[ld] = ||Jv||^2, global).
int hess_mul(S* h, const double* dv, double* out) {
    if (!h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound (bnl_eval_jacobian / bnl_upload_jacobian first)");
    if (h->hess_mode == BNL_HESSIAN_GRAM) {
        if (!h->gram_valid) RET(form_gram(h));
        EvScope ev(h, 0);
        CK(gram_apply(h->gram, h->n, h->ld, dv, out, h->stream));
        h->st.kernel_launches += 2;
    } else {
        const bool fused = h->p2p_on;  // reduce + NVLink push in one kernel, then wait+sum
        {
            EvScope ev(h, 0);
            CK(mv_launch(MODE_JTJV, h->plan, h->J, h->M, dv, nullptr, nullptr, h->partial, out, h->stream,
                         fused ? &h->p2p : nullptr, fused ? ++h->p2p_epoch : 0));
        }
        h->st.kernel_launches += fused ? 3 : 2;
        h->st.j_passes += 1;
        if (fused) {
            h->st.allreduces++;
            h->st.p2p_allreduces++;
        } else {
            RET(allreduce(h, out, (size_t)h->ld + 1));
        }
    }
    if (h->p > 0) {
        vk_hess_c(h->vc, dv, out, true, h->stream);
        KLAUNCH();
    }
    h->st.hess_mul++;
    h->st.jv++;
    h->st.jtw++;
    return BNL_OK;
}

// vthv(H,v) :92-96 -> leaves ||Jv||^2 in vc.hv[ld] (global) and Cv_sumsq in the scalars
int vthv_dev(S* h, const double* dv) {
    if (!h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound");
    if (h->hess_mode == BNL_HESSIAN_GRAM) {
        if (!h->gram_valid) RET(form_gram(h));
        EvScope ev(h, 1);
        CK(gram_apply(h->gram, h->n, h->ld, dv, h->vc.t1, h->stream));  // t1[ld] = v'Gv
        CK(cudaMemcpyAsync(h->vc.hv + h->ld, h->vc.t1 + h->ld, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        h->st.kernel_launches += 2;
    } else {
        const bool fused = h->p2p_on;
        {
            EvScope ev(h, 1);
            CK(mv_launch(MODE_JV, h->plan, h->J, h->M, dv, nullptr, nullptr, h->partial, h->vc.hv, h->stream,
                         fused ? &h->p2p : nullptr, fused ? ++h->p2p_epoch : 0));
        }
        h->st.kernel_launches += fused ? 3 : 2;
        h->st.j_passes += 1;
        if (fused) {
            h->st.allreduces++;
            h->st.p2p_allreduces++;
        } else {
            RET(allreduce(h, h->vc.hv + h->ld, 1));
        }
    }
    if (h->p > 0) {
        vk_hess_c(h->vc, dv, nullptr, false, h->stream);
        KLAUNCH();
    }
    h->st.vthv++;
    h->st.jv++;
    return BNL_OK;
}

// J' w  (w: device, local rows) -> out (length ld+1), all-reduced
int jtw_dev(S* h, const double* dw, double* out) {
    if (!h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound");
    const bool fused = h->p2p_on;
    {
        EvScope ev(h, 2);
        CK(mv_launch(MODE_JTW, h->plan, h->J, h->M, nullptr, dw, nullptr, h->partial, out, h->stream,
                     fused ? &h->p2p : nullptr, fused ? ++h->p2p_epoch : 0));
    }
    h->st.kernel_launches += fused ? 3 : 2;
    h->st.j_passes += 1;
    if (fused) {
        h->st.allreduces++;
        h->st.p2p_allreduces++;
    } else {
        RET(allreduce(h, out, (size_t)h->ld));
    }
    h->st.jtw++;
    return BNL_OK;
}

// ---- projection / active set (general path hooks) --------------------------------------------------------
int rebuild_chol(S* h) {  // update_chol! :62-68 (general path only; for m_lin == 0 the factor is I: nothing to do)
    if (h->mask) return BNL_OK;
    if (h->literal_proj)
        dk_rebuild(h->dc, h->vc.fix, h->stream);  // the reference's (m+q)^2 block factor, O(q^3)
    else
        dk_rs_rebuild(h->dc, h->vc.fix, h->stream);  // m x m factor of A_free A_free' (dense.h)
    KLAUNCH();
    h->st.chol_rebuilds++;
    return BNL_OK;
}
int check_chol(S* h) {  // after a sync
    if (!h->mask && h->sh->chol_fail) {
        cudaMemsetAsync(&h->sd->chol_fail, 0, sizeof(int), h->stream);
        return h->fail(BNL_ENOTPD, "PosDefException: cholesky of I - G'G failed (polyhedral_constraints.jl:57)");
    }
    return BNL_OK;
}
// v = P(+-r) into dst (general path); mask path is fused into the vec kernels, except for the fine-grained ABI
int project_general(S* h, const double* src, double* dst, bool negate) {
    if (h->literal_proj)
        dk_project(h->dc, src, dst, negate, h->stream);
    else
        dk_rs_project(h->dc, h->vc.fix, src, dst, negate, h->stream);
    KLAUNCH();
    return BNL_OK;
}

// ---- model evaluation ------------------------------------------------------------------------------------
ModelArgs margs(S* h) {
    ModelArgs a{};
    a.model_id = h->model_id;
    a.M = h->M;
    a.M_total = h->M_total;
    a.row0 = h->row0;
    a.n = h->n;
    a.ld = h->ld;
    a.seed = h->seed;
    a.noise = h->noise;
    a.cs = h->d_cs;
    return a;
}

// residuals(x) (+ nlconstraints(x) in callback mode).  dx: device x; rbuf: device M.  Leaves global dot(r,r)
// in sd->sumsq_r (after all-reduce).  c_out: host p-vector.
int eval_residual(S* h, const double* dx, double* rbuf, std::vector<double>& c_out) {
    c_out.assign(h->p, 0.0);
    if (h->model_id != 0) {
        EvScope ev(h, 3);
        CK(model_residual(margs(h), dx, h->ydata, rbuf, h->sumsq_partial, h->sumsq_blocks, &h->sd->sumsq_r, h->stream));
        h->st.kernel_launches += 2;
        if (h->p > 0) {  // built-in nlconstraints(x): the p-vector lives on the host, like the reference's closure result
            if (h->nl_kind != BNL_NLCONS_SPHERE) return h->fail(BNL_EINVAL, "p > 0 with a built-in model needs bnl_use_builtin_nlcons");
            vk_sphere_value(h->vc, dx, h->nl_rho2, h->stream);
            KLAUNCH();
            CK(cudaStreamSynchronize(h->stream));
            c_out[0] = h->sh->c0;
        }
    } else {
        if (!h->cb_res) return h->fail(BNL_EINVAL, "no model bound");
        h->h_x.resize(h->n);
        RET(get_vec(h, dx, h->h_x.data(), h->n));
        RET(ensure_pin(h, std::max<size_t>((size_t)h->M, 1 << 16)));
        if (h->cb_res(h->h_x.data(), h->pin, h->cb_ctx) != 0) return h->fail(BNL_ECALLBACK, "residuals callback failed");
        CK(cudaMemcpyAsync(rbuf, h->pin, (size_t)h->M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        vk_sumsq(rbuf, h->M, h->sumsq_partial, h->sumsq_blocks, &h->sd->sumsq_r, h->stream);
        h->st.kernel_launches += 2;
        if (h->p > 0) {
            if (!h->cb_nl) return h->fail(BNL_EINVAL, "p > 0 but no nlconstraints callback");
            if (h->cb_nl(h->h_x.data(), c_out.data(), h->cb_ctx) != 0) return h->fail(BNL_ECALLBACK, "nlconstraints callback failed");
        }
    }
    RET(allreduce(h, &h->sd->sumsq_r, 1));
    h->st.res_eval++;
    return BNL_OK;
}

// Pageable host memory -> device through TWO pinned staging buffers: the memcpy into one buffer overlaps the async
// H2D DMA out of the other (cudaMemcpyAsync on the solver's stream, one event per buffer).
int stage_upload(S* h, const double* src, double* dst, size_t count) {
    const size_t chunk = (size_t)1 << 19;  // 4 MB per staging buffer
    if (!h->pin2[0]) {
        for (int b = 0; b < 2; ++b) {
            CK(cudaHostAlloc(&h->pin2[b], chunk * sizeof(double), cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&h->pin2_ev[b], cudaEventDisableTiming));
        }
    }
    size_t k = 0;
    for (size_t off = 0; off < count; off += chunk, ++k) {
        const int b = (int)(k & 1);
        const size_t cnt = std::min(chunk, count - off);
        CK(cudaEventSynchronize(h->pin2_ev[b]));  // the DMA that last read this buffer (this call or an earlier one) is done
        memcpy(h->pin2[b], src + off, cnt * sizeof(double));
        CK(cudaMemcpyAsync(dst + off, h->pin2[b], cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaEventRecord(h->pin2_ev[b], h->stream));
    }
    return BNL_OK;
}

// Column-major (Julia) host matrix -> row-major device matrix: staged upload into a temporary column-major device
// buffer, then one transpose kernel.  Only host-supplied matrices (callback mode, small configs) pay for the temporary.
int upload_colmajor(S* h, const double* src, long long rows, int cols, long long lds, double* dst_rowmajor, int ldd) {
    const size_t total = (size_t)rows * cols;
    double* tmp = nullptr;
    CK(cudaMalloc(&tmp, std::max<size_t>(total, 1) * sizeof(double)));
    int rc = BNL_OK;
    if (lds == rows) {
        rc = stage_upload(h, src, tmp, total);
    } else {
        for (int c = 0; c < cols && rc == BNL_OK; ++c) rc = stage_upload(h, src + (size_t)c * lds, tmp + (size_t)c * rows, (size_t)rows);
    }
    if (rc == BNL_OK) {
        vk_transpose_in(tmp, rows, cols, rows, dst_rowmajor, ldd, h->stream);
        KLAUNCH();
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = h->fail(BNL_ECUDA, "upload_colmajor: %s", cudaGetErrorString(e));
    }
    cudaFree(tmp);
    return rc;
}

// jac_res(x), jac_nlcons(x): fills J (and C in callback mode)
int eval_jacobian(S* h, const double* dx) {
    if (h->model_id != 0) {
        {
            EvScope ev(h, 4);
            CK(model_jacobian(margs(h), dx, h->J, h->stream));
            KLAUNCH();
        }
        if (h->p > 0) {  // built-in jac_nlcons(x)
            if (h->nl_kind != BNL_NLCONS_SPHERE) return h->fail(BNL_EINVAL, "p > 0 with a built-in model needs bnl_use_builtin_nlcons");
            vk_sphere_jac(h->vc, dx, h->stream);
            vk_scale_C(h->vc, h->stream);
            h->st.kernel_launches += 2;
        }
    } else {
        if (!h->cb_jac) return h->fail(BNL_EINVAL, "no model bound");
        h->h_x.resize(h->n);
        RET(get_vec(h, dx, h->h_x.data(), h->n));
        h->h_tmp.resize((size_t)h->M * h->n);
        if (h->cb_jac(h->h_x.data(), h->h_tmp.data(), h->cb_ctx) != 0) return h->fail(BNL_ECALLBACK, "jac_res callback failed");
        RET(upload_colmajor(h, h->h_tmp.data(), h->M, h->n, h->M, h->J, h->ld));
        if (h->p > 0) {
            if (!h->cb_jnl) return h->fail(BNL_EINVAL, "p > 0 but no jac_nlcons callback");
            h->h_tmp.resize((size_t)h->p * h->n);
            if (h->cb_jnl(h->h_x.data(), h->h_tmp.data(), h->cb_ctx) != 0) return h->fail(BNL_ECALLBACK, "jac_nlcons callback failed");
            RET(upload_colmajor(h, h->h_tmp.data(), h->p, h->n, h->p, h->vc.C, h->ld));
            vk_scale_C(h->vc, h->stream);
            KLAUNCH();
        }
    }
    h->have_J = true;
    h->gram_valid = false;
    h->st.jac_eval++;
    if (h->hess_mode == BNL_HESSIAN_GRAM) RET(form_gram(h));
    return BNL_OK;
}

// g = Jx'*rx + Cx'*y_bar  (:45, :74)
int gradient(S* h, const double* rbuf, const std::vector<double>& ybar) {
    RET(jtw_dev(h, rbuf, h->vc.hv));
    CK(cudaMemcpyAsync(h->vc.g, h->vc.hv, (size_t)h->ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (h->p > 0) {
        RET(put_vec(h, ybar.data(), h->vc.pvec, h->p));
        vk_add_Ct(h->vc, h->vc.pvec, h->vc.g, h->stream);
        KLAUNCH();
    }
    return BNL_OK;
}

double al_value(S* h, double sumsq, const std::vector<double>& y, const std::vector<double>& c, double mu) {
    // mx = 0.5*dot(rx,rx) + dot(y,cx) + 0.5*mu*dot(cx,cx)  (:44, :59) -- p-vectors live on the host (callbacks)
    double yc = 0.0, cc = 0.0;
    for (int i = 0; i < h->p; ++i) {
        yc += y[i] * c[i];
        cc += c[i] * c[i];
    }
    return 0.5 * sumsq + yc + 0.5 * mu * cc;
}

// ---- cauchy_step :574-639 --------------------------------------------------------------------------------
int cauchy_step(S* h, double delta) {
    VecCtx& c = h->vc;
    vk_active_reset(c, c.x, nullptr, h->stream);  // :591
    KLAUNCH();
    RET(rebuild_chol(h));
    vk_cauchy_init(c, h->mask, h->stream);  // s_c = 0 ; d = P(-g) :592
    KLAUNCH();
    if (!h->mask) RET(project_general(h, c.g, c.d, true));
    RET(hess_mul(h, c.d, c.hv));  // :609
    vk_cauchy_eval(c, delta, h->stream);
    KLAUNCH();
    RET(sync(h));
    RET(check_chol(h));
    bool min_found = false;
    const int nmm = h->n - h->m_lin;
    while (!min_found && h->sh->nb_fix < nmm) {  // :615
        const double phi_p = h->sh->phi_p, phi_pp = h->sh->phi_pp, theta = h->sh->theta;
        const double delta_t = (phi_pp > 0) ? -phi_p / phi_pp : 0.0;  // :618
        if (phi_p >= 0) {
            min_found = true;
        } else if (phi_p < 0 && phi_pp > 0 && delta_t < theta) {
            vk_cauchy_advance(c, h->mask, 0, h->stream);  // :625
            KLAUNCH();
            min_found = true;
        } else {
            if (h->sh->bp_ind < 0) return h->fail(BNL_EBOUNDS, "BoundsError: next_breakpoint found no breakpoint (ind = -1)");
            vk_cauchy_advance(c, h->mask, 1, h->stream);  // :628-632
            KLAUNCH();
            RET(rebuild_chol(h));
            if (!h->mask) RET(project_general(h, c.g, c.d, true));
            RET(hess_mul(h, c.d, c.hv));  // :633
            vk_cauchy_eval(c, delta, h->stream);
            KLAUNCH();
            RET(sync(h));
            RET(check_chol(h));
            h->st.breakpoints++;
        }
    }
    return BNL_OK;
}

// reduced-gradient norms of g and g_minor with the current active set (:420-421, :446-447)
int nrg_general(S* h) {
    VecCtx& c = h->vc;
    RET(project_general(h, c.g, c.t1, true));
    vk_norm_to(c, c.t1, 0, h->stream);
    RET(project_general(h, c.gm, c.t1, true));
    vk_norm_to(c, c.t1, 1, h->stream);
    h->st.kernel_launches += 2;
    return BNL_OK;
}

// ---- minor_iterate :649-675 with projected_cg :690-764 and linesearch :766-791 ---------------------------
int minor_iterate(S* h, double delta, int* status_out, int* iters_out, bool apply_linesearch_and_accumulate) {
    VecCtx& c = h->vc;
    if (!h->mask) RET(project_general(h, c.gm, c.v, false));  // v = projection(lincons, r), r = g_minor :706
    vk_cg_init(c, h->mask, delta, h->stream);
    KLAUNCH();
    RET(sync(h));
    const int max_iter = 2 * (h->n - h->m_lin - h->sh->nb_fix);  // :714
    int iter = 1, nhp = 0;
    bool approx = false, outside = false, neg = false;
    while (!approx && !outside && !neg && iter <= max_iter) {  // :720
        RET(hess_mul(h, c.pdir, c.hv));                        // :722
        vk_cg_step(c, h->mask, 0, h->stream);
        KLAUNCH();
        if (!h->mask) {
            RET(sync(h));
            if (!h->sh->cg_neg_curv && !h->sh->cg_outside) {
                RET(project_general(h, c.r, c.v, false));  // :741
                vk_cg_step(c, false, 1, h->stream);
                KLAUNCH();
            }
        }
        RET(sync(h));
        neg = h->sh->cg_neg_curv != 0;
        outside = h->sh->cg_outside != 0;
        approx = h->sh->cg_solved != 0;
        iter = h->sh->cg_iter;
        ++nhp;
        h->st.cg_iters++;
    }
    int status;
    if (approx)
        status = BNL_CG_SOLVED;
    else if (outside)
        status = BNL_CG_BOUND_HIT;
    else if (neg)
        status = BNL_CG_NEGATIVE_CURVATURE;
    else if (iter == max_iter)
        status = BNL_CG_MAX_ITER;
    else
        status = BNL_CG_NOTHING;  // trap T3
    *status_out = status;
    if (iters_out) *iters_out = nhp;
    if (apply_linesearch_and_accumulate) {
        if (status != BNL_CG_NEGATIVE_CURVATURE) RET(vthv_dev(h, c.w));  // linesearch :775
        vk_minor_finish(c, h->stream);                                    // alpha, w *= alpha (:671), s += w (:436)
        KLAUNCH();
    }
    return BNL_OK;
}

// ---- inner_step :394-460 ---------------------------------------------------------------------------------
int inner_step(S* h, double delta, double* pred_out) {
    VecCtx& c = h->vc;
    RET(cauchy_step(h, delta));       // :410
    RET(hess_mul(h, c.s, c.hv));      // g_minor = H*s+g :412
    vk_gminor_nrg(c, h->mask, h->stream);
    KLAUNCH();
    if (!h->mask) RET(nrg_general(h));
    RET(sync(h));
    bool approx_solved = h->sh->nrg_gm <= h->prm.kappa3 * h->sh->nrg_g;  // :423
    const int allowed = h->n - h->m_lin - h->sh->nb_fix;                 // :425 (1-arg max, trap T6)
    const int max_minor = std::min(h->prm.max_minor_iter, allowed);
    bool cg_stop = false;
    int j = 1;
    while (j <= max_minor && !approx_solved && !cg_stop) {  // :430
        int status = 0;
        RET(minor_iterate(h, delta, &status, nullptr, true));  // :434-436
        cg_stop = (status == BNL_CG_NEGATIVE_CURVATURE);
        RET(hess_mul(h, c.s, c.hv));  // :437
        vk_minor_post(c, h->mask, delta, h->stream);
        KLAUNCH();
        RET(sync(h));
        if (h->m_lin + h->sh->n_at_bound <= h->n) {  // :441
            if (!h->mask) {
                RET(rebuild_chol(h));
                RET(nrg_general(h));
                RET(sync(h));
                RET(check_chol(h));
            }
            approx_solved = h->sh->nrg_gm <= h->prm.kappa3 * h->sh->nrg_g;  // :448
        } else {  // :450-452
            approx_solved = true;
            vk_active_reset(c, c.x, c.s, h->stream);
            KLAUNCH();
            RET(rebuild_chol(h));
        }
        ++j;
        h->st.minor_iters++;
    }
    RET(vthv_dev(h, c.s));  // :458
    vk_dot_gs(c, h->stream);
    KLAUNCH();
    RET(sync(h));
    RET(check_chol(h));
    *pred_out = h->sh->gs + 0.5 * (h->sh->jv_sumsq + h->vc.mu * h->sh->Cv_sumsq);
    return BNL_OK;
}

int set_mu(S* h, double mu) {
    h->vc.mu = mu;
    if (h->p > 0 && h->have_J) {
        vk_scale_C(h->vc, h->stream);
        KLAUNCH();
    }
    return BNL_OK;
}

// new_point :32-49 at the x in vc.x
int new_point(S* h, const std::vector<double>& y, double mu, double* mx_out) {
    VecCtx& c = h->vc;
    h->vc.mu = mu;
    RET(eval_residual(h, c.x, h->r, h->h_cx));
    RET(eval_jacobian(h, c.x));
    if (h->p > 0) {
        vk_scale_C(h->vc, h->stream);
        KLAUNCH();
    }
    h->h_ybar.resize(h->p);
    for (int i = 0; i < h->p; ++i) h->h_ybar[i] = y[i] + mu * h->h_cx[i];  // :43
    RET(gradient(h, h->r, h->h_ybar));                                     // :45
    vk_publish(h->sd, h->sh, h->stream);
    KLAUNCH();
    RET(sync(h));
    *mx_out = al_value(h, h->sh->sumsq_r, y, h->h_cx, mu);
    return BNL_OK;
}

}  // namespace

namespace {

// ---- solve_subproblem :303-378 ---------------------------------------------------------------------------
// x0 must already be in vc.x; y on the host.  Leaves x in vc.x, cx in h->h_cx.
int solve_subproblem_dev(S* h, const std::vector<double>& y, double mu, double omega_tol, double* pix_out, FILE* log) {
    VecCtx& c = h->vc;
    double mx = 0.0;
    RET(new_point(h, y, mu, &mx));  // :332
    vk_pix(c, true, h->stream);     // norm(g) for initial_tr (pix slot ignored here)
    KLAUNCH();
    RET(sync(h));
    double delta = h->prm.tr_factor * h->sh->norm_g;  // initial_tr :817-819
    double pix = kInf;
    int k = 1;
    bool solved = false;
    while (!solved && k <= h->prm.max_inner_iter) {  // :339
        double pred = 0.0;
        RET(inner_step(h, delta, &pred));  // :341
        vk_trial_point(c, h->stream);      // x_next = x+s :351
        KLAUNCH();
        RET(eval_residual(h, c.xn, h->r_trial, h->h_cx_next));  // evaluate_al :352
        vk_publish(h->sd, h->sh, h->stream);
        KLAUNCH();
        RET(sync(h));
        const double mx_next = al_value(h, h->sh->sumsq_r, y, h->h_cx_next, mu);
        const double ared = mx_next - mx;
        const double rho = ared / pred;  // NaN when pred == 0 (trap T8)
        bnl_inner_record rec{};
        rec.k = k;
        rec.mx = mx;
        rec.norm_s = h->sh->norm_s;
        rec.delta = delta;
        rec.rho = rho;
        rec.pred = pred;
        if (log) fprintf(log, "%4d   %.6e   %.2e   %.2e   %.2e\n", k, mx, rec.norm_s, delta, rho);  // misc.jl:70-80
        if (rho > h->prm.eta1) {  // :358-363
            CK(cudaMemcpyAsync(c.x, c.xn, (size_t)h->ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            std::swap(h->r, h->r_trial);
            h->h_cx = h->h_cx_next;
            mx = mx_next;
            RET(eval_jacobian(h, c.x));  // first_derivatives :72
            for (int i = 0; i < h->p; ++i) h->h_ybar[i] = y[i] + mu * h->h_cx[i];
            RET(gradient(h, h->r, h->h_ybar));  // :74
        }
        // update_tr :821-837 (NaN rho leaves delta unchanged)
        if (rho > h->prm.eta2)
            delta = h->prm.gamma2 * delta;
        else if (rho < h->prm.eta1)
            delta = h->prm.gamma1 * delta;
        // criticality_measure :369 with whatever active set inner_step left behind (trap T7)
        if (h->mask) {
            vk_pix(c, true, h->stream);
            KLAUNCH();
        } else {
            RET(project_general(h, c.g, c.t1, true));
            vk_norm_to(c, c.t1, 2, h->stream);
            KLAUNCH();
        }
        RET(sync(h));
        pix = h->sh->pix;
        rec.pix = pix;
        rec.nb_fix = h->sh->nb_fix;
        if (h->ilog.size() < (1u << 20)) h->ilog.push_back(rec);
        solved = pix < omega_tol;  // :373
        ++k;
        h->st.inner_iters++;
    }
    *pix_out = pix;
    return BNL_OK;
}

int free_problem(S* h) {
    cudaFree(h->J);
    cudaFree(h->r);
    cudaFree(h->r_trial);
    cudaFree(h->ydata);
    cudaFree(h->tvec);
    cudaFree(h->vecpool);
    cudaFree(h->flagpool);
    cudaFree(h->partial);
    cudaFree(h->sumsq_partial);
    cudaFree(h->d_words);
    cudaFree(h->d_idx);
    cudaFree(h->d_count);
    cudaFree(h->d_cs);
    cudaFree(h->d_xtrue);
    cudaFree(h->gram);
    cudaFree(h->gram_ws);
    h->gram_ws = nullptr;
    h->gram_valid = false;
    cudaFree(h->vc.C);
    cudaFree((void*)h->dc.A);
    cudaFree(h->dc.LA);
    cudaFree(h->dc.L);
    cudaFree(h->dc.G);
    cudaFree(h->dc.ywork);
    cudaFree(h->dc.Lr);
    cudaFree(h->dc.fixidx);
    cudaFree(h->dc.q_dev);
    h->J = h->r = h->r_trial = h->ydata = h->tvec = h->vecpool = h->partial = h->sumsq_partial = nullptr;
    h->flagpool = nullptr;
    h->d_words = nullptr;
    h->d_idx = nullptr;
    h->d_count = nullptr;
    h->d_cs = h->d_xtrue = h->gram = nullptr;
    h->dc = DenseCtx{};
    h->vc = VecCtx{};
    h->problem_set = false;
    h->have_J = false;
    h->model_id = 0;
    h->nl_kind = 0;
    return BNL_OK;
}

void sync_params_to_ctx(S* h) {
    h->vc.atol_active = h->prm.atol_active;
    h->vc.atol_negcurve = h->prm.atol_negcurve;
    h->vc.atol_boundary = h->prm.atol_boundary;
    h->vc.kappa2 = h->prm.kappa2;
}

bool valid(S* h) { return h != nullptr; }

}  // namespace

// ============================================ C ABI ======================================================
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

int bnl_comm_init(bnl_handle h, int nranks, int rank, const void* id128) {
    if (!valid(h) || nranks < 1 || rank < 0 || rank >= nranks) return BNL_EINVAL;
    if (nranks == 1) {
        h->nranks = 1;
        h->rank = 0;
        return BNL_OK;
    }
    if (!id128) return BNL_EINVAL;
    if (!g_nccl.load()) return h->fail(BNL_ENCCL, "cannot dlopen libnccl.so.2");
    CK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (r != ncclSuccess) return h->fail(BNL_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    h->nranks = nranks;
    h->rank = rank;
    // ---- peer-memory all-reduce over NVLink (p2p.h): exchange CUDA-IPC handles with ncclAllGather ----
    const char* env = getenv("BNL_P2P_ALLREDUCE");
    const bool want = !(env && env[0] == '0') && nranks <= kP2PMaxRanks && g_nccl.AllGather != nullptr;
    int ok = want ? 1 : 0;
    char* dh = nullptr;
    std::vector<cudaIpcMemHandle_t> all(nranks);
    if (want) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        const size_t bytes = std::max<size_t>(p2p_buffer_bytes(nranks), (size_t)4 << 20);
        cudaIpcMemHandle_t mine;
        ok = ok && cudaMalloc(&h->p2p_buf, bytes) == cudaSuccess && cudaMemset(h->p2p_buf, 0, bytes) == cudaSuccess &&
             cudaMalloc(&h->p2p_counter, 256) == cudaSuccess && cudaMemset(h->p2p_counter, 0, 256) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine, h->p2p_buf) == cudaSuccess && cudaMalloc(&dh, (size_t)nranks * 64) == cudaSuccess;
        if (ok) {
            cudaMemcpy(dh + (size_t)rank * 64, &mine, 64, cudaMemcpyHostToDevice);
            ok = g_nccl.AllGather(dh + (size_t)rank * 64, dh, 64, ncclChar, h->comm, h->stream) == ncclSuccess &&
                 cudaStreamSynchronize(h->stream) == cudaSuccess &&
                 cudaMemcpy(all.data(), dh, (size_t)nranks * 64, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        for (int r2 = 0; ok && r2 < nranks; ++r2) {
            void* base = h->p2p_buf;
            if (r2 != rank) {
                ok = cudaIpcOpenMemHandle(&base, all[r2], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                if (ok) h->p2p_opened[r2] = base;
            }
            h->p2p.mbox[r2] = static_cast<double*>(base);
            h->p2p.flag[r2] = p2p_flags_of(static_cast<double*>(base), nranks);
        }
        cudaGetLastError();
    }
    // all ranks must agree (min over ranks); this all-reduce is also the barrier after everyone's memset
    int* dok = nullptr;
    if (cudaMalloc(&dok, sizeof(int)) == cudaSuccess) {
        cudaMemcpy(dok, &ok, sizeof(int), cudaMemcpyHostToDevice);
        g_nccl.AllReduce(dok, dok, 1, ncclInt, ncclMin, h->comm, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaMemcpy(&ok, dok, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(dok);
    } else {
        ok = 0;
    }
    if (dh) cudaFree(dh);
    h->p2p.nranks = nranks;
    h->p2p.rank = rank;
    h->p2p.done_counter = h->p2p_counter;
    h->p2p.timeout_flag = &h->sd->p2p_timeout;
    h->p2p_on = ok != 0;
    h->p2p_epoch = 0;
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
    h->plan = mv_make_plan(M_local, n, h->prop.multiProcessorCount, optin);
    if (!h->plan.supported) return h->fail(BNL_EDIM, "n = %d unsupported by the streaming kernels (n <= 8192)", n);
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
    CK(cudaMalloc(&h->partial, (size_t)h->plan.grid * h->plan.pstride * sizeof(double)));
    CK(cudaMemset(h->partial, 0, (size_t)h->plan.grid * h->plan.pstride * sizeof(double)));
    CK(cudaMalloc(&h->sumsq_partial, h->sumsq_blocks * sizeof(double)));
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
    RET(sync(h));
    h->have_J = false;
    return BNL_OK;
}

int bnl_use_callbacks(bnl_handle h, bnl_callback residuals, bnl_callback jac_res, bnl_callback nlconstraints,
                      bnl_callback jac_nlcons, void* ctx) {
    if (!valid(h) || !h->problem_set || !residuals || !jac_res) return BNL_EINVAL;
    if (h->p > 0 && (!nlconstraints || !jac_nlcons)) return h->fail(BNL_EINVAL, "p > 0 needs nlconstraints and jac_nlcons");
    h->model_id = 0;
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
    return BNL_OK;
}

int bnl_model_set_truth(bnl_handle h, const double* x_true, const double* x0) {
    if (!valid(h) || h->model_id == 0 || !x_true) return BNL_EINVAL;
    CK(cudaSetDevice(h->device));
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

#define ENTER()                                                          \
    if (!valid(h)) return BNL_EINVAL;                                    \
    if (!h->problem_set) return h->fail(BNL_EINVAL, "bnl_set_problem first"); \
    CK(cudaSetDevice(h->device));

int bnl_upload_jacobian(bnl_handle h, const double* J_colmajor, int64_t ldj) {
    ENTER();
    if (!J_colmajor || ldj < h->M) return h->fail(BNL_EDIM, "DimensionMismatch: ldj < M");
    RET(upload_colmajor(h, J_colmajor, h->M, h->n, ldj, h->J, h->ld));
    h->have_J = true;
    h->gram_valid = false;
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
    CK(mv_launch(MODE_JV, h->plan, h->J, h->M, h->vc.t2, nullptr, h->tvec, h->partial, h->vc.hv, h->stream));
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
    RET(minor_iterate(h, delta, &status, &it, false));
    RET(sync(h));
    RET(get_vec(h, h->vc.w, w, h->n));
    if (cg_status) *cg_status = status;
    if (iters) *iters = it;
    return BNL_OK;
}

int bnl_inner_step(bnl_handle h, const double* x, const double* g, double delta, double* s, double* pred) {
    ENTER();
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
    cudaEvent_t e0 = h->ev_t0, e1 = h->ev_t1;
    RET(put_vec(h, x0, h->vc.x, h->n));
    CK(cudaEventRecord(e0, h->stream));
    std::vector<double> yv(h->p, 0.0);
    if (h->p > 0 && y) std::copy(y, y + h->p, yv.begin());
    double px = kInf;
    int rc = solve_subproblem_dev(h, yv, mu, omega_tol, &px, log);
    cudaEventRecord(e1, h->stream);
    cudaStreamSynchronize(h->stream);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    h->st.solve_ms += t;
    if (rc != BNL_OK) return rc;
    if (x) RET(get_vec(h, h->vc.x, x, h->n));
    if (cx && h->p > 0) std::copy(h->h_cx.begin(), h->h_cx.end(), cx);
    if (pix) *pix = px;
    return BNL_OK;
}

int bnl_solve_subproblem(bnl_handle h, const double* x0, const double* y, double mu, double omega_tol, double* x,
                         double* cx, double* pix) {
    ENTER();
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
    if ((kind == 3 || kind == 4) && h->model_id == 0) return h->fail(BNL_EINVAL, "builtin model needed");
    if (kind == 5 && !h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound");
    const double Jbytes = 8.0 * (double)h->M * (double)h->ld;
    double bytes = 0.0;
    cudaEvent_t e0 = h->ev_t0, e1 = h->ev_t1;
    // the vectors used: x as v (any data), r as w
    for (int rep = -1; rep < reps; ++rep) {  // one untimed warm-up
        if (rep == 0) CK(cudaEventRecord(e0, h->stream));
        switch (kind) {
            case 0:
                CK(mv_launch(MODE_JTJV, h->plan, h->J, h->M, h->vc.x, nullptr, nullptr, h->partial, h->vc.t1, h->stream));
                bytes = Jbytes + 16.0 * h->n;
                break;
            case 1:
                CK(mv_launch(MODE_JV, h->plan, h->J, h->M, h->vc.x, nullptr, nullptr, h->partial, h->vc.t1, h->stream));
                bytes = Jbytes + 8.0 * h->n;
                break;
            case 2:
                CK(mv_launch(MODE_JTW, h->plan, h->J, h->M, nullptr, h->r, nullptr, h->partial, h->vc.t1, h->stream));
                bytes = Jbytes + 8.0 * h->M + 8.0 * h->n;
                break;
            case 3:
                CK(model_residual(margs(h), h->vc.x, h->ydata, h->r_trial, h->sumsq_partial, h->sumsq_blocks, &h->sd->sumsq_r,
                                  h->stream));
                bytes = 16.0 * h->M;
                break;
            case 4:
                CK(model_jacobian(margs(h), h->vc.x, h->J, h->stream));
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

// ---- tralcnllss :167-298 (outer loop inside the library) ---------------------------------
int bnl_tralcnllss(bnl_handle h, const double* x0, const bnl_outer_params* op_in, const char* log_path, double* x_out,
                   double* y_out, double* final_mu, double* final_pix) {
    ENTER();
    bnl_outer_params op;
    if (op_in)
        op = *op_in;
    else
        bnl_default_outer_params(&op);
    FILE* log = log_path ? fopen(log_path, "w") : nullptr;
    const int n = h->n, p = h->p;
    std::vector<double> x(x0, x0 + n), y(p, 0.0), cx(p, 0.0), xn(n), cxn(p, 0.0);
    double mu = op.mu0;
    double omega = op.omega0 / std::pow(mu, op.k_crit), eta = op.eta0 / std::pow(mu, op.k_feas);  // :153-163
    int rc = BNL_OK;
    // least_squares_multipliers :887-903 : y = -(CC')^{-1} C J'r  (p small: host arithmetic on device-computed J'r and C)
    if (p > 0) {
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
    if (log && rc == BNL_OK) {
        double ss = 0.0;
        bnl_residuals(h, x.data(), nullptr, &ss);
        double nc = 0.0;
        for (double v : cx) nc += v * v;
        fprintf(log, "\n%s\n                          Outer iter %d\n  objective    nl feasibility     \xce\xbc      criticality   tolerance\n",
                std::string(80, '=').c_str(), outer_iter);
        fprintf(log, "%.7e   %.6e  %.2e        -         %.2e", ss, std::sqrt(nc), mu, omega);
        fprintf(log, "\n%s\niter     AL value       ||s||        \xce\x94          \xcf\x81\n", std::string(80, '=').c_str());
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
        if (log) {
            double ss = 0.0;
            bnl_residuals(h, x.data(), nullptr, &ss);  // objective :292
            fprintf(log, "\n%s\n                          Outer iter %d\n  objective    nl feasibility     \xce\xbc      criticality   tolerance\n",
                    std::string(80, '=').c_str(), outer_iter);
            fprintf(log, "%.7e   %.6e  %.2e     %.2e     %.2e", ss, feas, mu, pix, omega);
            fprintf(log, "\n%s\niter     AL value       ||s||        \xce\x94          \xcf\x81\n", std::string(80, '=').c_str());
        }
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
