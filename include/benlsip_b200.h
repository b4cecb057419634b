/*
 * benlsip_b200.h -- C ABI of libbenlsip_b200.so: a B200-native (sm_100a) replacement for the inner
 * Gauss-Newton trust-region subproblem solve of pierre-borie/BEnlsip.jl.
 *
 * The reference has NO FFI / plugin layer (pure Julia).  The boundary this library
 * offers is therefore the set of Julia methods on the hot path, one entry point per method, so a Julia
 * shim (julia/BEnlsipB200.jl) can forward each method to a `ccall`.
 * Each declaration cites the reference method it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - plain pointers and sizes only; all host arrays are caller-owned, FP64, Julia (column-major) layout;
 *     the library copies during the call and never retains host pointers;
 *   - indices crossing the boundary are 0-based;  `fixvars` crosses as UInt64 words in Julia
 *     `BitVector.chunks` layout (bit i&63 of word i>>6);
 *   - every function returns BNL_OK (0) or a negative bnl_status; bnl_last_error(h) gives the text.
 *     The Julia shim maps codes back to the exception types the reference throws
 *     (AssertionError, PosDefException, BoundsError, DimensionMismatch).
 *   - one handle = one solver instance bound to one GPU; calls on a handle must be serialised by the
 *     caller (the reference is not re-entrant either: src/basic_tralcnlss.jl:4, lincons mutated in place).
 *   - there is NO CPU fallback: bnl_create fails with BNL_ENODEV when no sm_100 device is visible.
 *   - multi-GPU: one process (one handle) per GPU; rows of J / r are sharded (bnl_shard_rows), every O(n)
 *     quantity is replicated; the only collective is the exchange of the n+1 per-group sums of a Hessian
 *     apply (NVLink stores into peer-mapped mailboxes fused into the reduction kernel; ncclAllGather as the
 *     fallback; ncclAllReduce only for the n^2 Gram of the opt-in Gram mode).  Results are bit-identical for
 *     1, 2, 4 and 8 GPUs.
 */
#ifndef BENLSIP_B200_H
#define BENLSIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bnl_solver* bnl_handle;

typedef enum bnl_status {
    BNL_OK = 0,
    BNL_EINVAL = -1,  /* bad argument / bad state (ArgumentError)                                   */
    BNL_EDIM = -2,    /* DimensionMismatch                                                           */
    BNL_ECUDA = -3,   /* CUDA runtime error                                                          */
    BNL_ENCCL = -4,   /* NCCL error / NCCL not loadable                                              */
    BNL_EOOM = -5,    /* device allocation failed                                                    */
    BNL_ENOTPD = -6,  /* PosDefException: cholesky() of A~A~' failed (src/polyhedral_constraints.jl:57) */
    BNL_EBOUNDS = -7, /* BoundsError: add_active!(ind=-1) (src/basic_tralcnlss.jl:544,:631)          */
    BNL_EASSERT = -8, /* AssertionError (src/basic_tralcnlss.jl:200; polyhedral_constraints.jl:43,:110,:128) */
    BNL_ENODEV = -9,  /* no usable sm_100 GPU: the library has no CPU path                           */
    BNL_ECALLBACK = -10
} bnl_status;

/* CG_status enum, src/basic_tralcnlss.jl:12; BNL_CG_NOTHING is Julia's `nothing` (SURVEY trap T3). */
enum { BNL_CG_SOLVED = 0, BNL_CG_BOUND_HIT = 1, BNL_CG_NEGATIVE_CURVATURE = 2, BNL_CG_MAX_ITER = 3, BNL_CG_NOTHING = -1 };

/* Keyword arguments of `tralcnllss` that reach the hot path (src/basic_tralcnlss.jl:177-197) plus the
 * constants the reference hard-wires (:817 tr_factor; polyhedral_constraints.jl:207,:224 atol_active;
 * :697 atol_negcurve; :798 atol_boundary).  bnl_default_params fills in the reference defaults.        */
typedef struct bnl_params {
    double eta1, eta2, gamma1, gamma2; /* 0.25 0.75 0.0625 2 */
    double kappa2, kappa3;             /* 0.1 0.1 */
    double tr_factor;                  /* 0.1 */
    double atol_active;                /* sqrt(eps) */
    double atol_negcurve;              /* sqrt(eps) */
    double atol_boundary;              /* 1e-10 */
    int32_t max_minor_iter;            /* 50  (nb_minor_step) */
    int32_t max_inner_iter;            /* 500 (k_max) */
} bnl_params;

/* Outer-loop keyword arguments (src/basic_tralcnlss.jl:177-197), used by bnl_tralcnllss only. */
typedef struct bnl_outer_params {
    double mu0, tau, omega0, eta0, feas_tol, crit_tol, k_crit, k_feas, beta_crit, beta_feas;
    int32_t max_outer_iter;
    int32_t reserved;
} bnl_outer_params;

/* Work counters (J-pass accounting of SURVEY.md section 3) and device times from CUDA events. */
typedef struct bnl_stats {
    int64_t outer_iters, inner_iters, minor_iters, cg_iters, breakpoints;
    int64_t hess_mul, vthv, jtw, jv, res_eval, jac_eval, chol_rebuilds, allreduces;
    double hess_mul_ms;   /* sum of CUDA-event durations of the fused J'(Jv) kernel launches           */
    double vthv_ms, jtw_ms, res_eval_ms, jac_eval_ms;
    double solve_ms;      /* CUDA-event duration of the last bnl_solve_subproblem / bnl_tralcnllss      */
    int64_t kernel_launches;
    int64_t j_passes;     /* streaming passes over the Jacobian actually executed                     */
    int64_t gram_count;   /* Gram formations (BNL_HESSIAN_GRAM)                                        */
    double gram_ms;
    int64_t p2p_allreduces; /* all-reduces done by the fused NVLink peer-memory kernels instead of NCCL       */
    int64_t inc_breakpoints; /* breakpoints handled by the device-side incremental Cauchy loop (no J pass)          */
    int64_t cauchy_loop_launches; /* launches of the persistent breakpoint-loop kernel                               */
    int64_t cauchy_literal_evals; /* literal Hd = H*d evaluations the guarded loop asked for (:633-635)              */
    int64_t t0_reuses;            /* Cauchy searches after a rejected step that reused t = J P(-g) instead of a J pass */
    int64_t chol_downdates;       /* O(m^2) rank-one downdates of the projection factor (one per Cauchy breakpoint, m_lin > 0) */
    double chol_ms;               /* CUDA-event time of the factor rebuilds + downdates                                */
    int64_t fused_jtr;            /* Jacobian generations that produced J'r on the fly (no J'w pass for the gradient)  */
    int64_t gram_breakpoints;     /* Cauchy breakpoints (m_lin > 0) whose Hd came from the Gram matrix (guarded)        */
    int64_t jt_builds;            /* tile-transposed copies of J built for long Cauchy searches (>= 128 breakpoints)     */
    int64_t point_reuses;         /* subproblems that started at the very x the previous one ended at (built-in models): r, J,
                                     J'r taken from HBM instead of re-evaluated (:332-336); BNL_REUSE_POINT=0 disables          */
} bnl_stats;

/* One line of the reference's inner-iteration log (print_inner_iter, src/misc.jl:70-80) + extras. */
typedef struct bnl_inner_record {
    int32_t k, nb_fix;
    double mx, norm_s, delta, rho, pix, pred;
    double omega_tol;                  /* the subproblem tolerance pix is tested against (:373) */
    int64_t breakpoints_cum, cg_cum;   /* Cauchy breakpoints / projected-CG iterations since bnl_reset_stats */
} bnl_inner_record;

/* How Base.:*(H,v) / vthv are evaluated.  MATRIX_FREE (default) is the reference's J'(Jv) (src/basic_tralcnlss.jl:102-106),
 * fused into one HBM pass.  GRAM forms G = J'J once per Jacobian on the FP64 tensor cores (DMMA) and applies H from G:
 * same mathematics, different rounding (kappa(G) = kappa(J)^2) => opt-in, validated separately.        */
enum { BNL_HESSIAN_MATRIX_FREE = 0, BNL_HESSIAN_GRAM = 1 };

/* How cauchy_step (src/basic_tralcnlss.jl:574-639) obtains phi' and phi'' after a breakpoint.  LITERAL: a fresh Hd = H*d
 * per breakpoint (:633), one pass over J each.  INCREMENTAL (default; used when the projection is the bound mask, i.e.
 * m_lin == 0, with at most 8 nonlinear constraints, in matrix-free mode): d only loses one component per breakpoint, so
 * t = J d and u = J s_c are updated in place (one strided column of J + two M-vector streams) by ONE persistent device
 * kernel that walks the breakpoints without host round trips.  The loop is guarded: a decision inside a rounding band, and
 * every interior minimiser (whose step length enters the iterate), is re-evaluated with the literal Hd = H*d, so the Cauchy
 * point is bit-identical to the literal search's.  With linear equality constraints (m_lin > 0) the same guard is applied to
 * breakpoints evaluated on G = J'J (formed once per Jacobian on the FP64 tensor cores when a search walks more than 8
 * breakpoints): again every number that reaches the iterate is literal.  Environment: BNL_CAUCHY=literal selects LITERAL at
 * bnl_create; BNL_CAUCHY_GUARD / BNL_GRAM_GUARD set the relative widths of the rounding bands (1e-9 / 1e-7).                */
enum { BNL_CAUCHY_LITERAL = 0, BNL_CAUCHY_INCREMENTAL = 1 };

/* Built-in device-side models (definitions in oracle/models.py, the executable spec). */
enum { BNL_MODEL_GLM = 1, BNL_MODEL_EXPSUM = 2 /* n/2 channel-separated decays */, BNL_MODEL_EXPSUM_DENSE = 3 /* one n/2-term sum */ };
/* Built-in nonlinear equality constraint (p = 1) for the device models: c(x) = x'x - rho2, params = {rho2}. */
enum { BNL_NLCONS_SPHERE = 1 };

/* User callbacks (the reference's `residuals, jac_res, nlconstraints, jac_nlcons` closures,
 * src/basic_tralcnlss.jl:167-176), e.g. Julia `@cfunction`.  Matrices are written column-major.
 * Called on the calling thread only.  Return 0 on success.                                             */
typedef int (*bnl_callback)(const double* x, double* out, void* ctx);

/* ---- lifetime ---------------------------------------------------------------------------------- */
int bnl_version(void);
int bnl_device_count(void);
int bnl_create(int device, bnl_handle* out);
void bnl_destroy(bnl_handle h);
const char* bnl_last_error(bnl_handle h);
const char* bnl_status_string(int status);
void bnl_default_params(bnl_params* p);
void bnl_default_outer_params(bnl_outer_params* p);
int bnl_set_params(bnl_handle h, const bnl_params* p);

/* ---- multi-GPU (row sharding).  id is an ncclUniqueId (128 bytes). ------------------- */
int bnl_comm_unique_id(void* id128);
int bnl_comm_init(bnl_handle h, int nranks, int rank, const void* id128);
/* p2p_allreduce = 1 when the n+1-double all-reduce runs as fused NVLink peer-memory kernels (CUDA IPC mailboxes;
 * default when all ranks could map each other; BNL_P2P_ALLREDUCE=0 forces NCCL). */
int bnl_comm_info(bnl_handle h, int32_t* nranks, int32_t* rank, int32_t* p2p_allreduce);
/* The rows rank `rank` of `nranks` (1, 2, 4 or 8) must own: every sum over residual rows is taken over a fixed geometry of
 * 8 groups x G chunks that depends only on M_total, so results are bit-identical for any supported GPU count; a rank owns
 * whole groups.  Pure host function (no device needed).                                                                 */
int bnl_shard_rows(int64_t M_total, int32_t nranks, int32_t rank, int64_t* row0, int64_t* M_local);

/* ---- problem: MixedConstraints(A, chol_aat; l, u), src/polyhedral_constraints.jl:9-18, and
 *      chol_aat = cholesky(A*A'), src/basic_tralcnlss.jl:206.  M_local rows [row0,row0+M_local) of
 *      M_total live on this GPU.  A is m_lin x n column-major (may be NULL when m_lin == 0).            */
int bnl_set_problem(bnl_handle h, int64_t M_local, int64_t M_total, int64_t row0, int32_t n, int32_t m_lin,
                    int32_t p, const double* A, const double* xlow, const double* xupp);

/* ---- model binding ---------------------------------------------------------------------------------
 * builtin: params = {noise, cond_exp} for GLM, {noise} for EXPSUM; data generated on the device.
 * callbacks: residuals -> M_local, jac_res -> M_local x n, nlconstraints -> p, jac_nlcons -> p x n.  */
int bnl_use_builtin_model(bnl_handle h, int32_t model_id, const double* params, int32_t nparams, uint32_t seed);
int bnl_use_callbacks(bnl_handle h, bnl_callback residuals, bnl_callback jac_res, bnl_callback nlconstraints,
                      bnl_callback jac_nlcons, void* ctx);
int bnl_model_vectors(bnl_handle h, double* x0, double* xlow, double* xupp, double* x_true); /* builtin only */
int bnl_use_builtin_nlcons(bnl_handle h, int32_t kind, const double* params, int32_t nparams);
int bnl_model_set_truth(bnl_handle h, const double* x_true, const double* x0 /*or NULL*/); /* regenerates the data y */

/* ---- AlHessian (src/basic_tralcnlss.jl:6-10): the handle holds the current (J, C, mu) -------------- */
int bnl_upload_jacobian(bnl_handle h, const double* J_colmajor, int64_t ldj);  /* pinned async H2D + transpose */
int bnl_upload_nlcons_jacobian(bnl_handle h, const double* C_colmajor, int64_t ldc);
int bnl_set_mu(bnl_handle h, double mu);
int bnl_eval_jacobian(bnl_handle h, const double* x);            /* jac_res(x), jac_nlcons(x) via model/callbacks */
int bnl_residuals(bnl_handle h, const double* x, double* r_local /*or NULL*/, double* sumsq /*global*/);
int bnl_nlcons(bnl_handle h, const double* x, double* c, double* C_colmajor); /* nlconstraints(x), jac_nlcons(x) :41-42 */
int bnl_gradient(bnl_handle h, const double* x, double* g);      /* jac_res(x)'*residuals(x)  :893 */
int bnl_hess_mul(bnl_handle h, const double* v, double* Hv);     /* Base.:*(H,v)   :102-106 */
int bnl_vthv(bnl_handle h, const double* v, double* out);        /* vthv(H,v)      :92-96   */
int bnl_jv(bnl_handle h, const double* v, double* Jv_local);     /* H.J*v          :93,:103 */
int bnl_jtw(bnl_handle h, const double* w_local, double* JTw);   /* H.J'*w         :105,:45 */
int bnl_gram(bnl_handle h, double* G_colmajor /*n x n or NULL*/, double* ms); /* J'J: K12, not in the reference */
int bnl_set_hessian_mode(bnl_handle h, int32_t mode);
int bnl_set_cauchy_mode(bnl_handle h, int32_t mode);

/* ---- MixedConstraints methods (src/polyhedral_constraints.jl) --------------------------------------- */
int bnl_project(bnl_handle h, const double* r, double* v);                 /* projection!          :158-170 */
int bnl_left_mul(bnl_handle h, const double* x, double* y);                /* left_mul: y = [A x; x[fix]]  :86-98 (y: m_lin + nb_fix) */
int bnl_left_mul_tr(bnl_handle h, const double* y, double* x);             /* left_mul_tr: x = A~' y       :72-84 */
int bnl_active_bounds_reset(bnl_handle h, const double* x);                /* active_bounds!       :203-215 */
int bnl_active_bounds(bnl_handle h, const double* x, const double* s, double delta, int64_t* idx,
                      int32_t* count);                                     /* active_bounds        :219-237 */
int bnl_add_active(bnl_handle h, const int64_t* idx, int32_t count);       /* add_active!          :240-261 */
int bnl_set_fixvars(bnl_handle h, const uint64_t* words);                  /* fixvars .= ...; update_chol! :62-68 */
int bnl_get_fixvars(bnl_handle h, uint64_t* words, int32_t* nb_fix);       /* lincons.fixvars, nb_fix :31 */
int bnl_get_chol(bnl_handle h, double* L_colmajor, int32_t* dim);          /* lincons.chol.L */

/* ---- step computation and the subproblem solve (the hot path) -------------------------------------- */
/* cauchy_step(x,g,H,chol_aat,lincons,delta) :574-639 */
int bnl_cauchy_step(bnl_handle h, const double* x, const double* g, double delta, double* s_c);
/* projected_cg(g_minor,H,w_l,w_u,lincons,kappa2) :690-764 (w_l/w_u built as in minor_iterate :662-665) */
int bnl_projected_cg(bnl_handle h, const double* x, const double* s, const double* g_minor, double delta,
                     double* w, int32_t* cg_status, int32_t* iters);
/* projected_cg(g_minor,H,w_l,w_u,lincons,kappa2) :690-764 with the caller's own w_l / w_u and the current fixvars */
int bnl_projected_cg_bounds(bnl_handle h, const double* g_minor, const double* w_l, const double* w_u, double* w,
                            int32_t* cg_status, int32_t* iters);
/* linesearch(g_model,H,w,w_l,w_u,lincons.fixvars) :766-791 -> alpha */
int bnl_linesearch(bnl_handle h, const double* g_model, const double* w, const double* w_l, const double* w_u, double* alpha);
/* inner_step(x,g,H,chol_aat,lincons,delta,nb_minor_step,kappa2,kappa3) :394-460 -> (s, model_reduction) */
int bnl_inner_step(bnl_handle h, const double* x, const double* g, double delta, double* s, double* pred);
/* new_point(x,y,mu,...) :32-49 -> mx, g (J, C kept in the handle as H) */
int bnl_new_point(bnl_handle h, const double* x, const double* y, double mu, double* mx, double* g, double* cx);
/* solve_subproblem(x0,y,mu,...,omega_tol,...) :303-378 -> (x, cx, pix).
 * With a built-in device model, a call whose x0 equals -- bit for bit -- the x the previous bnl_solve_subproblem on this handle
 * returned (tralcnllss does exactly that whenever its feasibility test passes, :273-283) does not re-evaluate residuals(x),
 * jac_res(x), Jx'*rx (:332-336): they are still in device memory and do not depend on (y, mu, omega_tol).  Results are
 * bit-identical either way; any other entry point in between, or BNL_REUSE_POINT=0 at bnl_create, forces the re-evaluation.
 * User callbacks are always called.                                                                                       */
int bnl_solve_subproblem(bnl_handle h, const double* x0, const double* y, double mu, double omega_tol, double* x,
                         double* cx, double* pix);
/* tralcnllss(x0,...) :167-298 -> (x, y); SURVEY 8f rank 1 (outer loop inside the library; optional:
 * the Julia / Python host may keep the outer loop and call bnl_solve_subproblem instead).
 * log_path NULL = no log; otherwise the reference's benlsip.out format (src/misc.jl).                  */
int bnl_tralcnllss(bnl_handle h, const double* x0, const bnl_outer_params* op, const char* log_path, double* x,
                   double* y, double* final_mu, double* final_pix);

/* ---- introspection --------------------------------------------------------------------------------- */
int bnl_get_stats(bnl_handle h, bnl_stats* out);
int bnl_reset_stats(bnl_handle h);
int bnl_get_inner_log(bnl_handle h, bnl_inner_record* out, int32_t capacity, int32_t* count);
/* Device-side microbenchmarks for the roofline report: `reps` back-to-back launches of one kernel class
 * timed with CUDA events on the library's stream; kind: 0 fused J'(Jv), 1 Jv (norm only), 2 J'w,
 * 3 residual eval, 4 Jacobian generation, 5 Gram (DMMA; FLOPs are returned in bytes_per_launch), 6 Jacobian generation fused
 * with J'r (GLM, n <= 1024).
 * Returns average ms per launch and algorithmic bytes per launch. */
int bnl_time_kernel(bnl_handle h, int32_t kind, int32_t reps, double* avg_ms, double* bytes_per_launch);
int bnl_device_info(bnl_handle h, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* free_bytes,
                    int64_t* total_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BENLSIP_B200_H */
