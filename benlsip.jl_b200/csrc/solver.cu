// solver.cu -- host control flow of the inner Gauss-Newton trust-region subproblem solve and the C ABI
// (include/benlsip_b200.h).  The control flow mirrors, function by function, the reference's
// src/basic_tralcnlss.jl (solve_subproblem :303-378, inner_step :394-460, cauchy_step :574-639,
// minor_iterate :649-675, projected_cg :690-764, linesearch :766-791) -- every numeric operation runs in a
// CUDA kernel; the host only branches on scalars the kernels publish to pinned mapped memory.
// There is no CPU arithmetic path and no fallback: without an sm_100 device bnl_create fails.
#include "solver_internal.h"

NcclApi g_nccl;

namespace bnl_host {


int sync(S* h) {
    if (h->p2p_on) vk_publish(h->sd, h->sh, h->stream);  // make a peer-wait timeout visible even when no O(n) kernel followed
    CK(cudaStreamSynchronize(h->stream));
    if (h->p2p_on && h->sh->p2p_timeout) return h->fail(BNL_ENCCL, "peer-memory exchange timed out waiting for a rank");
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
                case 6: h->st.chol_ms += ms; break;
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

EvScope::EvScope(S* h_, int cls) : h(h_) {
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
EvScope::~EvScope() {
    if (ok) {
        cudaEventRecord(e.b, h->stream);
        h->ev_busy.push_back(e);
    }
}

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

// NCCL all-reduce: only the n^2 Gram (opt-in mode) still uses it; every reduction of the default path goes through row_reduce.
int allreduce(S* h, double* buf, size_t count) {
    if (h->nranks <= 1) return BNL_OK;
    if (!h->comm) return h->fail(BNL_ENCCL, "no NCCL communicator (bnl_comm_init)");
    ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess) return h->fail(BNL_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    h->st.allreduces++;
    return BNL_OK;
}

// The fixed reduction tree of rowgeom.h: teams -> chunks -> groups (this rank's groups), exchange of the group sums with the
// peers (NVLink stores into every rank's mailbox, or ncclAllGather when peers cannot be mapped), then the 8 group sums in
// order.  Bit-identical on every rank and for N = 1, 2, 4, 8.
int row_reduce(S* h, const double* P, int T, long long pstride, int col0, int ncols, double* out) {
    const unsigned long long epoch = ++h->p2p_epoch;
    const RowGeom& g = h->geo;
    if (h->nranks <= 1) {
        CK(group_reduce(P, g.G, T, pstride, col0, ncols, g.g0, g.ng, h->p2p, epoch, GR_LOCAL, h->stream));
        CK(group_sum(h->p2p, epoch, out, col0, ncols, 0, h->stream));
    } else if (h->p2p_on) {
        CK(group_reduce(P, g.G, T, pstride, col0, ncols, g.g0, g.ng, h->p2p, epoch, GR_PUSH, h->stream));
        CK(group_sum(h->p2p, epoch, out, col0, ncols, 1, h->stream));
        h->st.allreduces++;
        h->st.p2p_allreduces++;
    } else {
        CK(group_reduce(P, g.G, T, pstride, col0, ncols, g.g0, g.ng, h->p2p, epoch, GR_LOCAL, h->stream));
        double* region = h->p2p_buf + (size_t)(epoch & 1ull) * kGroups * kP2PWidth;  // [kGroups][kP2PWidth]: rank r owns rows [r*ng, (r+1)*ng)
        const size_t cnt = (size_t)g.ng * kP2PWidth;
        if (!h->comm || !g_nccl.AllGather) return h->fail(BNL_ENCCL, "no NCCL communicator (bnl_comm_init)");
        ncclResult_t r = g_nccl.AllGather(region + (size_t)g.g0 * kP2PWidth, region, cnt, ncclDouble, h->comm, h->stream);
        if (r != ncclSuccess) return h->fail(BNL_ENCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        CK(group_sum(h->p2p, epoch, out, col0, ncols, 0, h->stream));
        h->st.allreduces++;
    }
    h->st.kernel_launches += 2;
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
int hess_mul(S* h, const double* dv, double* out, double* t_out) {
    if (!h->have_J) return h->fail(BNL_EINVAL, "no Jacobian bound (bnl_eval_jacobian / bnl_upload_jacobian first)");
    if (h->hess_mode == BNL_HESSIAN_GRAM) {
        if (!h->gram_valid) RET(form_gram(h));
        EvScope ev(h, 0);
        CK(gram_apply(h->gram, h->n, h->ld, dv, out, h->stream));
        h->st.kernel_launches += 2;
    } else {
        {
            EvScope ev(h, 0);
            CK(mv_launch(MODE_JTJV, h->plan, h->geo, h->J, dv, nullptr, t_out, h->partial, h->stream));
        }
        KLAUNCH();
        h->st.j_passes += 1;
        RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, 0, h->ld + 1, out));
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
        {
            EvScope ev(h, 1);
            CK(mv_launch(MODE_JV, h->plan, h->geo, h->J, dv, nullptr, nullptr, h->partial, h->stream));
        }
        KLAUNCH();
        h->st.j_passes += 1;
        RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, h->ld, h->ld + 1, h->vc.hv));
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
    {
        EvScope ev(h, 2);
        CK(mv_launch(MODE_JTW, h->plan, h->geo, h->J, nullptr, dw, nullptr, h->partial, h->stream));
    }
    KLAUNCH();
    h->st.j_passes += 1;
    RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, 0, h->ld, out));
    h->st.jtw++;
    return BNL_OK;
}

// ---- projection / active set (general path hooks) --------------------------------------------------------
int rebuild_chol(S* h) {  // update_chol! :62-68 (general path only; for m_lin == 0 the factor is I: nothing to do)
    if (h->mask) return BNL_OK;
    EvScope ev(h, 6);
    if (h->literal_proj)
        dk_rebuild(h->dc, h->vc.fix, h->stream);  // the reference's (m+q)^2 block factor, O(q^3)
    else
        dk_rs_rebuild(h->dc, h->vc.fix, h->stream);  // m x m factor of A_free A_free' (dense.h)
    KLAUNCH();
    h->st.chol_rebuilds++;
    return BNL_OK;
}
// add_active!(lincons, chol_aat, ind) for the breakpoint variable sd->bp_ind: O(m^2) downdate of the reduced-space factor
// (the literal block factor, BNL_LITERAL_PROJECTION=1, is rebuilt as in the reference)
int downdate_chol(S* h) {
    if (h->mask) return BNL_OK;
    if (h->literal_proj || !dk_rs_downdate_fits(h->dc)) return rebuild_chol(h);
    EvScope ev(h, 6);
    dk_rs_downdate(h->dc, h->stream);
    KLAUNCH();
    h->st.chol_downdates++;
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
    a.geo = h->geo;
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
    if (rbuf == h->r) h->pc_valid = false;
    c_out.assign(h->p, 0.0);
    if (h->model_id != 0) {
        EvScope ev(h, 3);
        CK(model_residual(margs(h), dx, h->ydata, rbuf, h->rpartial, h->stream));
        KLAUNCH();
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
        vk_sumsq_chunks(rbuf, h->geo, h->rpartial, h->stream);
        KLAUNCH();
        if (h->p > 0) {
            if (!h->cb_nl) return h->fail(BNL_EINVAL, "p > 0 but no nlconstraints callback");
            if (h->cb_nl(h->h_x.data(), c_out.data(), h->cb_ctx) != 0) return h->fail(BNL_ECALLBACK, "nlconstraints callback failed");
        }
    }
    RET(row_reduce(h, h->rpartial, 1, 1, 0, 1, &h->sd->sumsq_r));  // dot(rx,rx) over all ranks' rows
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
int eval_jacobian(S* h, const double* dx, const double* r_for_gradient) {
    h->jtr_valid = false;
    h->pc_valid = false;
    if (h->model_id != 0) {
        // first_derivatives (:72-74) computes Jx'*rx right after Jx: when the residual of this x is at hand the GLM generator
        // accumulates J'r while it writes J (bit-identical to a separate J'w pass) and the pass is not streamed back in
        const bool fuse = r_for_gradient != nullptr && h->fuse_jtr && h->model_id == BNL_MODEL_GLM && h->plan.warp_team && h->plan.T == 8;
        {
            EvScope ev(h, 4);
            if (fuse)
                CK(model_jacobian_jtr(margs(h), dx, r_for_gradient, h->J, h->partial, h->plan.pstride, h->plan.KCH, h->plan.RB, h->stream));
            else
                CK(model_jacobian(margs(h), dx, h->J, h->stream));
            KLAUNCH();
        }
        if (fuse) {
            RET(row_reduce(h, h->partial, h->plan.T, h->plan.pstride, 0, h->ld, h->vc.hv));
            h->jtr_valid = true;
            h->st.fused_jtr++;
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
    h->t0_valid = false;
    h->jt_valid = h->jt_attempted = false;
    h->st.jac_eval++;
    if (h->hess_mode == BNL_HESSIAN_GRAM) RET(form_gram(h));
    return BNL_OK;
}

// g = Jx'*rx + Cx'*y_bar  (:45, :74).  The J'r part is kept in d_jtr: a subproblem that restarts from this point takes it from there.
int gradient(S* h, const double* rbuf, const std::vector<double>& ybar) {
    h->t0_valid = false;
    const size_t bytes = (size_t)h->ld * sizeof(double);
    if (!h->d_jtr) CK(cudaMalloc(&h->d_jtr, (size_t)(h->ld + kColAlign) * sizeof(double)));
    if (h->jtr_cached) {  // new_point at the point of the last solve's end: J'r of this (J, r) is still in d_jtr
        h->jtr_cached = false;
        h->st.jtw++;
    } else {
        if (h->jtr_valid) {  // J'r came out of the Jacobian generation (eval_jacobian): no pass
            h->jtr_valid = false;
            h->st.jtw++;
        } else {
            RET(jtw_dev(h, rbuf, h->vc.hv));
        }
        CK(cudaMemcpyAsync(h->d_jtr, h->vc.hv, bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaMemcpyAsync(h->vc.g, h->d_jtr, bytes, cudaMemcpyDeviceToDevice, h->stream));
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

// ---- cauchy_step :574-639 with the breakpoint loop on the device (cauchy_loop.cu) ---------------------------------
// The first interval is the reference's own: Hd = H*d (:609) in one fused pass that also leaves t = J d in HBM, literal
// phi', phi'' (:610-611).  A search without breakpoints therefore costs exactly what the literal one costs.  From the first
// breakpoint on, ONE persistent kernel walks the breakpoints on t = J d, u = J s_c; whenever a number that reaches the
// iterate is needed (interior minimiser) or a decision is inside the rounding band, Hd = H*d is evaluated literally
// (:633-635) and the loop is re-entered with those values => the Cauchy point equals the literal search's bit for bit.
// After a REJECTED step (x, g, J untouched) both Hd and t of the first interval are reused: no pass at all.
int cauchy_step_incremental(S* h, double delta) {
    VecCtx& c = h->vc;
    if (!h->inc_t) {
        const size_t mb = std::max<size_t>(h->M, 16) * sizeof(double);
        CK(cudaMalloc(&h->inc_t, mb));
        CK(cudaMalloc(&h->inc_u, mb));
        CK(cudaMalloc(&h->inc_t0, mb));
        CK(cudaMalloc(&h->hd0, (size_t)(h->ld + kColAlign) * sizeof(double)));
        CK(cudaMalloc(&h->cl_sync, cauchy_loop_sync_bytes()));
        h->t0_valid = false;
    }
    vk_active_reset(c, c.x, nullptr, h->stream);  // :591
    vk_cauchy_init(c, true, h->stream);           // s_c = 0 ; d = P(-g) :592
    h->st.kernel_launches += 2;
    const size_t hv_bytes = (size_t)(h->ld + kColAlign) * sizeof(double);
    if (h->t0_valid) {
        // a rejected step left x, g and J untouched (:358-366): d = P(-g) is the same vector bit for bit, and so are
        // Hd = H*d (:609) and t = J d -- reuse both instead of a pass over J
        CK(cudaMemcpyAsync(h->inc_t, h->inc_t0, (size_t)h->M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(c.hv, h->hd0, hv_bytes, cudaMemcpyDeviceToDevice, h->stream));
        h->st.t0_reuses++;
        h->st.hess_mul++;  // logical applies of the reference (:609), no pass here
        h->st.jv++;
        h->st.jtw++;
    } else {
        RET(hess_mul(h, c.d, c.hv, h->inc_t));  // Hd = H*d :609 -- the fused pass also leaves t = J d in HBM
        CK(cudaMemcpyAsync(h->inc_t0, h->inc_t, (size_t)h->M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(h->hd0, c.hv, hv_bytes, cudaMemcpyDeviceToDevice, h->stream));
        h->t0_valid = true;
    }
    vk_cauchy_eval(c, delta, h->stream);  // phi', phi'' of the first interval: the literal ones (:610-611)
    KLAUNCH();
    CauchyLoopArgs a{};
    a.c = c;
    a.geo = h->geo;
    a.J = h->J;
    a.ld = h->ld;
    a.t = h->inc_t;
    a.u = h->inc_u;
    a.partial2 = h->rpartial;
    a.p2p = h->p2p;
    a.multi = h->nranks > 1 ? 1 : 0;
    a.delta = delta;
    a.guard = h->cauchy_guard;
    a.nmm = h->n - h->m_lin;
    a.arrive = h->cl_sync;
    a.bcast = reinterpret_cast<char*>(h->cl_sync) + 64;
    a.first = 1;
    a.use_literal = 1;  // the first interval is decided with the literal scalars: a search without breakpoints costs one pass
    int search_bp = 0;
    for (;;) {
        a.ll_epoch0 = h->ll_epoch;
        a.q0 = search_bp;
        a.Jt = h->jt_valid ? h->Jt : nullptr;
        a.want_jt = h->jt_attempted ? 0 : 1;
        CK(cauchy_loop_launch(a, h->prop.multiProcessorCount, h->stream));
        h->st.kernel_launches++;
        h->st.cauchy_loop_launches++;
        RET(sync(h));
        h->ll_epoch += (unsigned long long)h->sh->cl_rounds;
        h->st.breakpoints += h->sh->cl_breakpoints;
        h->st.inc_breakpoints += h->sh->cl_breakpoints;
        search_bp += h->sh->cl_breakpoints;
        switch (h->sh->cl_status) {
            case CL_WANT_TRANSPOSE: {
                // A long search (>= kJtTrigger breakpoints): every further breakpoint reads one column of J, and HBM serves a strided
                // 8-byte read with a 128-byte line.  The tile-transposed copy (16 consecutive rows of a column per line) makes the
                // column read 16 x denser; it costs two passes' worth of traffic once per Jacobian and the same bytes again in HBM,
                // so it is only built here, and only if the memory is there.  Same values, same arithmetic: bit-identical results.
                h->jt_attempted = true;  // every rank takes this branch at the same breakpoint, whether or not its copy succeeds
                if (!h->jt_disabled) {
                    const size_t bytes = (size_t)((h->M + kJtTile - 1) / kJtTile) * kJtTile * h->ld * sizeof(double);
                    if (!h->Jt && cudaMalloc(&h->Jt, std::max<size_t>(bytes, 16)) != cudaSuccess) {
                        cudaGetLastError();
                        h->Jt = nullptr;
                        h->jt_disabled = true;
                    }
                    if (h->Jt) {
                        CK(transpose16_launch(h->J, h->M, h->ld, h->Jt, h->stream));
                        KLAUNCH();
                        h->jt_valid = true;
                        h->st.jt_builds++;
                    }
                }
                a.first = 0;
                a.use_literal = 0;  // re-decide the same breakpoint from the (unchanged) t, u
                break;
            }
            case CL_DONE_NOSTEP:
            case CL_DONE_INTERIOR:
            case CL_DONE_EXHAUSTED: return BNL_OK;
            case CL_NEED_LITERAL:
                RET(hess_mul(h, c.d, c.hv));          // :609 / :633
                vk_cauchy_eval(c, delta, h->stream);  // :610-611 / :634-635 -> sd->phi_p, sd->phi_pp
                KLAUNCH();
                h->st.cauchy_literal_evals++;
                a.first = 0;
                a.use_literal = 1;
                break;
            case CL_ERR_BOUNDS: return h->fail(BNL_EBOUNDS, "BoundsError: next_breakpoint found no breakpoint (ind = -1)");
            case CL_TIMEOUT: return h->fail(BNL_ENCCL, "peer-memory exchange timed out waiting for a rank (Cauchy loop)");
            default: return h->fail(BNL_ECUDA, "cauchy loop: unexpected status %d", h->sh->cl_status);
        }
    }
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
            RET(downdate_chol(h));  // add_active!(ind): one column leaves A_free
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

// ---- cauchy_step :574-639 for the general projection (m_lin > 0), breakpoints evaluated on the Gram matrix, GUARDED --------
// Whenever the trust region is active the reference's search walks up to n faces of the box, one Hd = H*d (:633) and one factor
// update each.  Here the first kGramSwitch breakpoints of a search are the reference's own; if the search is still advancing,
// G = J'J is formed ONCE per Jacobian on the FP64 tensor cores (gram.cu) and the following breakpoints take Hd from G (an
// L2-resident n x n gemv instead of a pass over J).  G*d rounds differently from J'(Jd), so -- exactly like the device loop of the
// bound-only case (cauchy_loop.cu) -- such a value only ever takes a decision that is outside a rounding band; an interior
// minimiser (whose step length enters the iterate) and every in-band decision are re-evaluated with the literal Hd.  The Cauchy
// point is therefore bit-identical to the literal search's, whatever the Gram matrix's own rounding (or reduction order across
// GPUs) is.
int cauchy_step_gram_guarded(S* h, double delta) {
    constexpr int kGramSwitch = 8;
    VecCtx& c = h->vc;
    vk_active_reset(c, c.x, nullptr, h->stream);  // :591
    KLAUNCH();
    RET(rebuild_chol(h));
    vk_cauchy_init(c, h->mask, h->stream);  // s_c = 0 ; d = P(-g) :592
    KLAUNCH();
    if (!h->mask) RET(project_general(h, c.g, c.d, true));
    RET(hess_mul(h, c.d, c.hv));  // :609 (literal)
    vk_cauchy_eval(c, delta, h->stream);
    KLAUNCH();
    RET(sync(h));
    RET(check_chol(h));
    bool literal = true;  // are the scalars in sh the literal ones?
    int nbp = 0;
    const int nmm = h->n - h->m_lin;
    auto literal_eval = [&]() -> int {
        RET(hess_mul(h, c.d, c.hv));
        vk_cauchy_eval(c, delta, h->stream);
        KLAUNCH();
        RET(sync(h));
        h->st.cauchy_literal_evals++;
        literal = true;
        return BNL_OK;
    };
    bool min_found = false;
    while (!min_found && h->sh->nb_fix < nmm) {  // :615
        if (!literal) {
            // guard: may this Gram-derived (phi', phi'') take the decision "advance" on its own?
            const double phi_p = h->sh->phi_p, phi_pp = h->sh->phi_pp, theta = h->sh->theta;
            const double scale = std::fabs(h->sh->phi_a) + std::fabs(h->sh->phi_b);
            bool clear_advance = false;
            if (std::fabs(phi_p) > h->gram_guard * scale && phi_p < 0.0 && phi_pp > 0.0) {
                const double relb = h->gram_guard * scale / std::fabs(phi_p) + h->gram_guard;
                clear_advance = (-phi_p / phi_pp) > theta * (1.0 + relb);
            }
            if (!clear_advance) RET(literal_eval());  // stop, interior minimiser, or inside the band: the reference's own numbers decide
        }
        const double phi_p = h->sh->phi_p, phi_pp = h->sh->phi_pp, theta = h->sh->theta;
        const double delta_t = (phi_pp > 0) ? -phi_p / phi_pp : 0.0;  // :618
        if (literal && phi_p >= 0) {
            min_found = true;
        } else if (literal && phi_p < 0 && phi_pp > 0 && delta_t < theta) {
            vk_cauchy_advance(c, h->mask, 0, h->stream);  // :625, literal step length
            KLAUNCH();
            min_found = true;
        } else {
            if (h->sh->bp_ind < 0) return h->fail(BNL_EBOUNDS, "BoundsError: next_breakpoint found no breakpoint (ind = -1)");
            if (!h->literal_proj && dk_rs_downdate_fits(h->dc)) {
                // :628-632 in one launch: s_c += theta d ; add_active!(ind) (flag + O(m^2) factor downdate) ; d = P(-g)
                EvScope ev(h, 6);
                dk_rs_breakpoint(h->dc, c.s, c.d, c.g, c.fix, h->stream);
                KLAUNCH();
                h->st.chol_downdates++;
            } else {
                vk_cauchy_advance(c, h->mask, 1, h->stream);
                KLAUNCH();
                RET(downdate_chol(h));
                RET(project_general(h, c.g, c.d, true));
            }
            ++nbp;
            h->st.breakpoints++;
            if (nbp >= kGramSwitch || h->gram_valid) {
                if (!h->gram_valid) RET(form_gram(h));
                CK(gram_gemv(h->gram, h->n, h->ld, c.d, c.hv, h->stream));  // Hd ~ G d  (k_cauchy_eval takes the dots itself)
                KLAUNCH();
                if (h->p > 0) {
                    vk_hess_c(h->vc, c.d, c.hv, true, h->stream);
                    KLAUNCH();
                }
                h->st.gram_breakpoints++;
                literal = false;
            } else {
                RET(hess_mul(h, c.d, c.hv));  // :633 (literal)
                literal = true;
            }
            vk_cauchy_eval(c, delta, h->stream);
            KLAUNCH();
            RET(sync(h));
            RET(check_chol(h));
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
int minor_iterate(S* h, double delta, int* status_out, int* iters_out, bool apply_linesearch_and_accumulate, bool bounds_given) {
    VecCtx& c = h->vc;
    if (!h->mask) RET(project_general(h, c.gm, c.v, false));  // v = projection(lincons, r), r = g_minor :706
    vk_cg_init(c, h->mask, delta, bounds_given, h->stream);
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
    // bound-only problems (mask projection; the loop carries up to kCLMaxP nonlinear-constraint rows itself): device-side
    // breakpoint loop; everything else, and Gram mode (where H*d is an L2-resident gemv anyway): the literal search
    // (with several ranks the loop exchanges its two scalars through the peer-mapped LL mailbox: without peer mapping, literal)
    if (h->cauchy_mode == BNL_CAUCHY_INCREMENTAL && h->mask && h->p <= kCLMaxP && h->hess_mode == BNL_HESSIAN_MATRIX_FREE &&
        (h->nranks == 1 || h->p2p_on))
        RET(cauchy_step_incremental(h, delta));
    else if (h->cauchy_mode == BNL_CAUCHY_INCREMENTAL && !h->mask && h->hess_mode == BNL_HESSIAN_MATRIX_FREE)
        RET(cauchy_step_gram_guarded(h, delta));  // general projection: long searches take Hd from the Gram matrix, guarded
    else
        RET(cauchy_step(h, delta));  // :410
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
        RET(minor_iterate(h, delta, &status, nullptr, true, false));  // :434-436
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
    // vthv(H,s) :458 = ||J s||^2 + mu ||C s||^2.  The last Hessian apply above was H*s of this very s (:412 / :437): the fused
    // kernel left ||J s||^2 in hv[ld] (same per-row arithmetic and reduction tree as the J*v-only mode) and k_hess_c left
    // ||C s||^2, so the reference's extra pass over J is not repeated.
    h->st.vthv++;
    h->st.jv++;
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

bool point_hit(const S* h, const double* x_host) {
    return x_host && h->reuse_point && h->pc_valid && h->model_id != 0 && h->have_J && (int)h->pc_x.size() == h->n &&
           std::memcmp(h->pc_x.data(), x_host, (size_t)h->n * sizeof(double)) == 0;
}

// new_point :32-49 at the x in vc.x (x_host: the same vector on the host, when the caller has it)
int new_point(S* h, const std::vector<double>& y, double mu, double* mx_out, const double* x_host) {
    VecCtx& c = h->vc;
    h->vc.mu = mu;
    const bool hit = point_hit(h, x_host);
    h->pc_valid = false;  // from here on h->r / J follow the iterates of this solve
    double sumsq = 0.0;
    if (hit) {
        // residuals(x), jac_res(x), Jx'*rx of a built-in (pure) model at the very point the previous subproblem ended at:
        // r, J (with its Gram matrix / tile-transposed copy, if formed) and J'r are in HBM, dot(rx,rx) and cx on the host --
        // the values a re-evaluation would reproduce bit for bit (every kernel on the path is deterministic)
        h->h_cx = h->pc_cx;
        sumsq = h->pc_sumsq;
        if (h->p > 0) {  // built-in jac_nlcons(x): O(n), regenerated (the scaling below depends on mu)
            if (h->nl_kind != BNL_NLCONS_SPHERE) return h->fail(BNL_EINVAL, "p > 0 with a built-in model needs bnl_use_builtin_nlcons");
            vk_sphere_jac(h->vc, c.x, h->stream);
            vk_scale_C(h->vc, h->stream);
            h->st.kernel_launches += 2;
        }
        if (h->hess_mode == BNL_HESSIAN_GRAM && !h->gram_valid) RET(form_gram(h));
        h->jtr_cached = h->d_jtr != nullptr;
        h->jtr_valid = false;
        h->st.point_reuses++;
    } else {
        RET(eval_residual(h, c.x, h->r, h->h_cx));
        RET(eval_jacobian(h, c.x, h->r));
    }
    if (h->p > 0) {
        vk_scale_C(h->vc, h->stream);
        KLAUNCH();
    }
    h->h_ybar.resize(h->p);
    for (int i = 0; i < h->p; ++i) h->h_ybar[i] = y[i] + mu * h->h_cx[i];  // :43
    RET(gradient(h, h->r, h->h_ybar));                                     // :45
    // Without nonlinear constraints g = J'r and H = J'J are those of the previous subproblem's end as well: if its last step was
    // rejected, Hd = H*P(-g) and t = J P(-g) of the first Cauchy interval (:609) are still the right ones (cauchy_step_incremental)
    if (hit && h->p == 0 && h->t0_carry) h->t0_valid = true;
    h->t0_carry = false;
    vk_publish(h->sd, h->sh, h->stream);
    KLAUNCH();
    RET(sync(h));
    if (!hit) sumsq = h->sh->sumsq_r;
    h->acc_sumsq = sumsq;
    *mx_out = al_value(h, sumsq, y, h->h_cx, mu);
    return BNL_OK;
}


// ---- solve_subproblem :303-378 ---------------------------------------------------------------------------
// x0 must already be in vc.x; y on the host.  Leaves x in vc.x, cx in h->h_cx.
int solve_subproblem_dev(S* h, const std::vector<double>& y, double mu, double omega_tol, double* pix_out, FILE* log,
                         const double* x0_host) {
    VecCtx& c = h->vc;
    double mx = 0.0;
    RET(new_point(h, y, mu, &mx, x0_host));  // :332
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
        const double sumsq_next = h->sh->sumsq_r;
        const double mx_next = al_value(h, sumsq_next, y, h->h_cx_next, mu);
        const double ared = mx_next - mx;
        const double rho = ared / pred;  // NaN when pred == 0 (trap T8)
        bnl_inner_record rec{};
        rec.k = k;
        rec.mx = mx;
        rec.norm_s = h->sh->norm_s;
        rec.delta = delta;
        rec.rho = rho;
        rec.pred = pred;
        if (log) {  // print_inner_iter, misc.jl:70-80 (Julia's Printf spells non-finite values NaN / Inf)
            auto e = [](int prec, double v) -> std::string {
                if (std::isnan(v)) return "NaN";
                if (std::isinf(v)) return v > 0 ? "Inf" : "-Inf";
                char buf[64];
                snprintf(buf, sizeof buf, "%.*e", prec, v);
                return buf;
            };
            fprintf(log, "%4d   %s   %s   %s   %s\n", k, e(6, mx).c_str(), e(2, rec.norm_s).c_str(), e(2, delta).c_str(), e(2, rho).c_str());
            fflush(log);  // a long solve can be followed (and a cut-off one read) from its log
        }
        if (rho > h->prm.eta1) {  // :358-363
            CK(cudaMemcpyAsync(c.x, c.xn, (size_t)h->ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            std::swap(h->r, h->r_trial);
            h->h_cx = h->h_cx_next;
            mx = mx_next;
            h->acc_sumsq = sumsq_next;
            RET(eval_jacobian(h, c.x, h->r));  // first_derivatives :72 (+ the J'r of :74 on the fly)
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
        rec.omega_tol = omega_tol;
        rec.breakpoints_cum = h->st.breakpoints;
        rec.cg_cum = h->st.cg_iters;
        if (h->ilog.size() < (1u << 20)) h->ilog.push_back(rec);
        solved = pix < omega_tol;  // :373
        ++k;
        h->st.inner_iters++;
    }
    *pix_out = pix;
    return BNL_OK;
}

void p2p_local_setup(S* h) {
    h->p2p = P2PArgs{};
    h->p2p.nranks = 1;
    h->p2p.rank = 0;
    h->p2p.mbox[0] = h->p2p_buf;
    h->p2p.flag[0] = p2p_flags_of(h->p2p_buf);
    h->p2p.ll[0] = p2p_ll_of(h->p2p_buf);
    h->p2p.done_counter = h->p2p_counter;
    h->p2p.timeout_flag = &h->sd->p2p_timeout;
}

// The row-chunk geometry of this handle (rowgeom.h).  After bnl_comm_init it follows (nranks, rank) and the caller's rows
// must be exactly that rank's shard (bnl_shard_rows); before, it is inferred from (row0, M_local): the largest whole-group
// range that matches (a single-GPU problem owns all 8 groups).
int resolve_geometry(S* h) {
    RowGeom g{};
    if (h->comm_set) {
        if (!geom_make(h->M_total, h->nranks, h->rank, &g)) return h->fail(BNL_EINVAL, "bad (nranks, rank)");
        if (g.row0 != h->row0 || g.local_rows() != h->M)
            return h->fail(BNL_EDIM, "DimensionMismatch: rows [%lld, %lld) are not the shard of rank %d of %d (use bnl_shard_rows: [%lld, %lld))",
                           h->row0, h->row0 + h->M, h->rank, h->nranks, g.row0, g.row0 + g.local_rows());
    } else {
        bool found = false;
        for (int nr = 1; nr <= kGroups && !found; nr <<= 1)
            for (int r = 0; r < nr && !found; ++r)
                found = geom_make(h->M_total, nr, r, &g) && g.row0 == h->row0 && g.local_rows() == h->M;
        if (!found)
            return h->fail(BNL_EDIM, "DimensionMismatch: rows [%lld, %lld) of %lld are not a whole-group shard (use bnl_shard_rows)", h->row0,
                           h->row0 + h->M, h->M_total);
    }
    h->geo = g;
    return BNL_OK;
}

int alloc_row_buffers(S* h) {
    cudaFree(h->partial);
    cudaFree(h->rpartial);
    h->partial = h->rpartial = nullptr;
    const size_t np = std::max<size_t>(mv_partial_doubles(h->plan, h->geo), 16);
    CK(cudaMalloc(&h->partial, np * sizeof(double)));
    CK(cudaMemset(h->partial, 0, np * sizeof(double)));
    CK(cudaMalloc(&h->rpartial, std::max<size_t>((size_t)h->geo.ng * h->geo.G * 4, 16) * sizeof(double)));
    // a handle that owns only some groups and has no peers (a shard examined on its own): the other groups' mailbox rows
    // must read as zero
    if (h->nranks <= 1) CK(cudaMemset(h->p2p_buf, 0, p2p_buffer_bytes()));
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
    cudaFree(h->rpartial);
    cudaFree(h->d_words);
    cudaFree(h->d_idx);
    cudaFree(h->d_count);
    cudaFree(h->d_cs);
    cudaFree(h->d_xtrue);
    cudaFree(h->gram);
    cudaFree(h->gram_ws);
    cudaFree(h->inc_t);
    cudaFree(h->inc_u);
    cudaFree(h->cl_sync);
    cudaFree(h->inc_t0);
    cudaFree(h->hd0);
    cudaFree(h->d_jtr);
    h->d_jtr = nullptr;
    h->pc_valid = h->jtr_cached = false;
    cudaFree(h->Jt);
    h->Jt = nullptr;
    h->jt_valid = h->jt_attempted = false;
    h->inc_t = h->inc_u = h->inc_t0 = h->hd0 = nullptr;
    h->t0_valid = false;
    h->cl_sync = nullptr;
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
    h->J = h->r = h->r_trial = h->ydata = h->tvec = h->vecpool = h->partial = h->rpartial = nullptr;
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


}  // namespace bnl_host
