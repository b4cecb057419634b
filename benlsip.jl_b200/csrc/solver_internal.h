// solver_internal.h -- state and internal interface of the host-side solver (shared by solver.cu, the control flow,
// and capi.cu, the extern "C" layer).  Not installed: the public interface is include/benlsip_b200.h.
#pragma once
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
#include "cauchy_loop.h"
#include "common.cuh"
#include "dense.h"
#include "gram.h"
#include "matvec.h"
#include "models.h"
#include "p2p.h"
#include "rowgeom.h"
#include "vecops.h"


using namespace bnl;


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
extern NcclApi g_nccl;

constexpr double kInf = std::numeric_limits<double>::infinity();


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
    RowGeom geo{};              // row-chunk geometry of every row reduction (rowgeom.h)
    double* partial = nullptr;  // [ng][G][T][pstride] partials of the streaming kernels
    double* rpartial = nullptr; // [ng][G][4] partials of the scalar row reductions
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
    int cauchy_mode = 0;        // BNL_CAUCHY_LITERAL / BNL_CAUCHY_INCREMENTAL
    double *inc_t = nullptr, *inc_u = nullptr;   // t = J d, u = J s_c (incremental Cauchy search)
    double* inc_t0 = nullptr;                    // J P(-g) of the current (x, g, J): reused by the searches after rejected steps
    double* hd0 = nullptr;                       // H P(-g) of the same state (n-vector + norm slot)
    bool t0_valid = false;
    bool t0_carry = false;                       // t0_valid as the previous solve_subproblem left it (see new_point)
    double* Jt = nullptr;                        // tile-transposed copy of J for long Cauchy searches (built lazily, may stay null)
    bool jt_valid = false;                       // Jt holds the current J
    bool jt_attempted = false;                   // a copy was attempted for the current J (uniform across ranks)
    bool jt_disabled = false;                    // BNL_JT=0, or the allocation failed once
    bool jtr_valid = false;                      // hv = J'r of the Jacobian just generated (fused generator)
    bool fuse_jtr = true;                        // BNL_FUSE_JTR=0 disables the fused generator
    // The point the last solve_subproblem ended at (built-in device models only: pure functions of x).  tralcnllss restarts
    // the next subproblem from exactly that x whenever the feasibility test passes (:273-283), and residuals(x), jac_res(x)
    // and Jx'*rx do not depend on (y, mu, omega): r, J and J'r are still in HBM.  pc_valid says that h->r, h->J, d_jtr,
    // pc_sumsq and pc_cx belong to the host vector pc_x; every entry point other than the two solve calls clears it.
    bool reuse_point = true;                     // BNL_REUSE_POINT=0: re-evaluate at every subproblem start like the reference
    bool pc_valid = false;
    bool jtr_cached = false;                     // gradient(): take J'r from d_jtr (set by new_point on a hit)
    std::vector<double> pc_x, pc_cx;
    double pc_sumsq = 0.0, acc_sumsq = 0.0;      // dot(rx,rx) of pc_x / of the current accepted iterate inside a solve
    double* d_jtr = nullptr;                     // J'r of the last gradient() (ld doubles)
    unsigned int* cl_sync = nullptr;             // arrive counter + broadcast record of the persistent loop kernel
    double cauchy_guard = 1e-9;                  // relative width of the loop's rounding band
    double gram_guard = 1e-7;                    // the same for breakpoints evaluated on the Gram matrix (general projection)

    // model binding
    int model_id = 0;
    uint32_t seed = 0;
    double noise = 0.0, cond_exp = 0.0;
    double* d_cs = nullptr;
    double* d_xtrue = nullptr;
    std::vector<double> m_x0, m_xlow, m_xupp, m_xtrue;
    std::vector<double> h_xlow, h_xupp;  // host copy of the bounds in force (log header: count(isfinite, x_l))
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
    bool comm_set = false;      // bnl_comm_init was called (the geometry then follows (nranks, rank))
    // group mailbox (p2p.h): local at N = 1, mapped into every peer (CUDA IPC) at N > 1
    bool p2p_on = false;        // N > 1 and all peers mapped: group sums travel as NVLink stores, else by ncclAllGather
    P2PArgs p2p{};
    double* p2p_buf = nullptr;
    unsigned int* p2p_counter = nullptr;
    unsigned long long p2p_epoch = 0;
    unsigned long long ll_epoch = 0;
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


namespace bnl_host {

typedef bnl_solver S;

struct EvScope {  // records a CUDA-event pair around a kernel class on the launching stream
    S* h;
    EvPair e{};
    bool ok = false;
    EvScope(S* h_, int cls);
    ~EvScope();
};

int sync(S* h);
int ensure_pin(S* h, size_t doubles);
int put_vec(S* h, const double* src, double* dst, size_t count);
int get_vec(S* h, const double* src_dev, double* dst, size_t count);
int allreduce(S* h, double* buf, size_t count);
// finishes a row reduction: partials P[ng][G][T][pstride] -> out[col0..ncols) summed over ALL ranks' rows (fixed tree)
int row_reduce(S* h, const double* P, int T, long long pstride, int col0, int ncols, double* out);
int resolve_geometry(S* h);
int alloc_row_buffers(S* h);
void p2p_local_setup(S* h);
int form_gram(S* h);
int hess_mul(S* h, const double* dv, double* out, double* t_out = nullptr);  // t_out: also store J v (local rows)
int vthv_dev(S* h, const double* dv);
int jtw_dev(S* h, const double* dw, double* out);
int rebuild_chol(S* h);
int downdate_chol(S* h);
int check_chol(S* h);
int project_general(S* h, const double* src, double* dst, bool negate);
bnl::ModelArgs margs(S* h);
int eval_residual(S* h, const double* dx, double* rbuf, std::vector<double>& c_out);
int upload_colmajor(S* h, const double* src, long long rows, int cols, long long lds, double* dst_rowmajor, int ldd);
int eval_jacobian(S* h, const double* dx, const double* r_for_gradient = nullptr);  // r given: J'r is accumulated on the fly
int gradient(S* h, const double* rbuf, const std::vector<double>& ybar);
int cauchy_step(S* h, double delta);
int minor_iterate(S* h, double delta, int* status_out, int* iters_out, bool apply_linesearch_and_accumulate, bool bounds_given);
int inner_step(S* h, double delta, double* pred_out);
int new_point(S* h, const std::vector<double>& y, double mu, double* mx_out, const double* x_host = nullptr);
bool point_hit(const S* h, const double* x_host);  // x_host is bit for bit the point r, J, J'r in HBM were evaluated at
int solve_subproblem_dev(S* h, const std::vector<double>& y, double mu, double omega_tol, double* pix_out, FILE* log,
                         const double* x0_host = nullptr);
int free_problem(S* h);
void sync_params_to_ctx(S* h);
inline bool valid(S* h) { return h != nullptr; }

}  // namespace bnl_host
