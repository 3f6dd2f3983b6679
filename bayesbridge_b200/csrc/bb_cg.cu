// The prior-preconditioned conjugate-gradient sampler (reference: reg_coef_sampler/cg_sampler.py:20-151)
// and the CG loop it delegates to scipy.sparse.linalg.cg (scipy 1.18.1, _isolve/iterative.py:cg).
// The recurrences below follow scipy's update order exactly:
//     r = b - A x0
//     loop: if ||r|| < atol: stop ; rho = r.r ; p = r + (rho/rho_prev) p ; q = A p ;
//           alpha = rho / (p.q) ; x += alpha p ; r -= alpha q
// with A(x) = D.x + s.X'(omega.(X(s.x))), D = (s.prior_prec_sqrt)^2  (cg_sampler.py:105-112).
// All vector state stays on the device; the host only reads back {iter, done} between chunks.
#include "bb_internal.cuh"

int bb_dense_tdot(bb_mat* m, const double* w, const int* done_flag);

__device__ __forceinline__ double tdot_entry(const double* __restrict__ traw, const double* __restrict__ c, i64 j, int icpt) {
    const double sw = traw[0];
    if (j < icpt) return sw;
    // sparse_matrix.py:126-128   result = X.T.dot(v); result -= sum(v) * column_offset
    return __dsub_rn(traw[1 + (j - icpt)], __dmul_rn(sw, c[j - icpt]));
}

// u_n = sqrt(omega) .* eps1  (cg_sampler.py:66), eps1 injected or Philox
__global__ void k_rhs_noise(const double* __restrict__ omega, double omega_scalar, const double* __restrict__ eps1,
                            int philox, uint64_t seed, uint64_t offset, i64 row_offset, i64 n, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double e;
        if (philox) { RandStream rs; rs.init(seed, offset, (uint64_t)(row_offset + i), STREAM_EPS1); e = rs.normal(); }
        else e = eps1[i];
        double om = omega ? omega[i] : omega_scalar;
        out[i] = sqrt(om) * e;
    }
}

// y_gaussian weighting for z = X'(omega .* y): logit -> kappa ; linear -> omega * y
__global__ void k_z_weight(const double* __restrict__ n_trial, const double* __restrict__ n_success, int is_linear,
                           const double* __restrict__ omega, double omega_scalar, i64 n, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        if (is_linear) out[i] = (omega ? omega[i] : omega_scalar) * n_success[i];
        else out[i] = n_success[i] - 0.5 * n_trial[i];
    }
}

// b = s.(z + v), v = Tdot(sqrt(omega) eps1) + pps.eps2 ; D = (s.pps)^2 ; x = x0 / s ; partial b.b
__global__ void k_cg_init(const double* __restrict__ traw, const double* __restrict__ c, int icpt, i64 P,
                          const double* __restrict__ z, const double* __restrict__ pps, const double* __restrict__ s,
                          const double* __restrict__ x0, const double* __restrict__ eps2,
                          int philox, uint64_t seed, uint64_t offset,
                          double* __restrict__ b, double* __restrict__ D, double* __restrict__ x, double* __restrict__ red_bb) {
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double e2;
        if (philox) { RandStream rs; rs.init(seed, offset, (uint64_t)j, STREAM_EPS2); e2 = rs.normal(); }
        else e2 = eps2[j];
        double t = tdot_entry(traw, c, j, icpt);
        double v = __dadd_rn(t, __dmul_rn(pps[j], e2));
        double bj = __dmul_rn(s[j], __dadd_rn(z[j], v));
        b[j] = bj;
        double sp = __dmul_rn(s[j], pps[j]);
        D[j] = __dmul_rn(sp, sp);
        x[j] = x0[j] / s[j];
        acc += bj * bj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_bb[blockIdx.x] = acc;
}

__global__ void k_cg_scalars(CgScalars* st, const double* __restrict__ red_bb, int nred, double atol, int maxiter,
                             unsigned long long* ps_bar) {
    double bb = warp_sum_partials(red_bb, nred);
    if (threadIdx.x == 0) {
        st->bar_base = 0ull;        // grid-barrier bookkeeping of the fused P-side kernel restarts with every solve
        *ps_bar = 0ull;
        double bn = sqrt(bb);
        st->bnorm = bn;
        // cg_sampler.py:75 rtol = atol/||b|| ; scipy: atol_eff = max(0, rtol*||b||)
        st->atol_eff = (bn > 0.0) ? (atol / bn) * bn : atol;
        st->iter = 0;
        st->done = (bn == 0.0) ? 3 : 0;
        st->maxiter = maxiter;
        st->rho[0] = 0.0; st->rho[1] = 0.0;
        st->rnorm = 0.0;
    }
}

// q = D.p + s.Tdot ; partial p.q ; bumps the iteration counter when count_iter.
// With a peer-memory exchange (pub != nullptr) the all-reduce of [sum w; X'w] is done HERE: the block waits for
// every rank's publication and each thread sums the ranks' partials in rank order (compute + collective fused).
__global__ void k_cg_q(CgScalars* st, const double* __restrict__ traw, const double* __restrict__ c, int icpt, i64 P,
                       const double* __restrict__ pvec, const double* __restrict__ s, const double* __restrict__ D,
                       double* __restrict__ q, double* __restrict__ red_pq, int count_iter,
                       const P2PView* __restrict__ pub_ptr) {
    if (st->done) return;
    __shared__ double sm[33];
    __shared__ int p2p_ok;
    P2PView pub;
    double sw = 0.0;
    if (pub_ptr != nullptr) {
        pub = *pub_ptr;
        if (!p2p_wait_all(pub, &p2p_ok)) return;
        sw = p2p_sum(pub, 0);
    }
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double t;
        if (pub_ptr != nullptr) {
            t = (j < icpt) ? sw : __dsub_rn(p2p_sum(pub, 1 + (j - icpt)), __dmul_rn(sw, c[j - icpt]));
        } else {
            t = tdot_entry(traw, c, j, icpt);
        }
        double pj = pvec[j];
        double qj = __dadd_rn(__dmul_rn(D[j], pj), __dmul_rn(s[j], t));
        q[j] = qj;
        acc += pj * qj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) {
        red_pq[blockIdx.x] = acc;
        if (blockIdx.x == 0 && count_iter) st->iter = st->iter + 1;
    }
}

// r = b - q ; partial r.r      (initial residual)
__global__ void k_cg_resid(const CgScalars* st, const double* __restrict__ b, const double* __restrict__ q, i64 P,
                           double* __restrict__ r, double* __restrict__ red_rr) {
    if (st->done) return;
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double rj = __dsub_rn(b[j], q[j]);
        r[j] = rj;
        acc += rj * rj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_rr[blockIdx.x] = acc;
}

// convergence test, rho, search direction, and the scaled gather vector of the next product
__global__ void k_cg_dir(CgScalars* st, const double* __restrict__ red_rr, int nred, i64 P, int icpt,
                         const double* __restrict__ r, double* __restrict__ pvec, const double* __restrict__ s,
                         const double* __restrict__ c, double* __restrict__ sv, double* __restrict__ red_shift) {
    if (st->done) return;
    __shared__ double sm[33];
    const int it = st->iter;
    const bool lead = (blockIdx.x == 0 && threadIdx.x == 0);
    const double rho = warp_sum_partials(red_rr, nred);
    const double rn = sqrt(rho);
    // scipy: `for iteration in range(maxiter): if norm(r) < atol: return x, 0 ...` -- after maxiter
    // updates the loop ends without another convergence test.
    if (it >= st->maxiter) { if (lead) { st->done = 2; st->rnorm = rn; } return; }
    if (rn < st->atol_eff) { if (lead) { st->done = 1; st->rnorm = rn; } return; }
    const double beta = (it > 0) ? rho / st->rho[(it + 1) & 1] : 0.0;
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double pj = (it > 0) ? __dadd_rn(__dmul_rn(beta, pvec[j]), r[j]) : r[j];
        pvec[j] = pj;
        double x = __dmul_rn(s[j], pj);
        sv[j] = x;
        acc += (j < icpt) ? x : -c[j - icpt] * x;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_shift[blockIdx.x] = acc;
    if (lead) { st->rho[it & 1] = rho; st->rnorm = rn; }
}

// alpha = rho/(p.q) ; x += alpha p ; r -= alpha q ; partial r.r
__global__ void k_cg_update(const CgScalars* st, const double* __restrict__ red_pq, int nred, i64 P,
                            double* __restrict__ x, double* __restrict__ r, const double* __restrict__ pvec,
                            const double* __restrict__ q, double* __restrict__ red_rr) {
    if (st->done) return;
    __shared__ double sm[33];
    const int it = st->iter - 1;            // k_cg_q already counted this iteration
    const double pq = warp_sum_partials(red_pq, nred);
    const double alpha = st->rho[it & 1] / pq;
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        x[j] = __dadd_rn(x[j], __dmul_rn(alpha, pvec[j]));
        double rj = __dsub_rn(r[j], __dmul_rn(alpha, q[j]));
        r[j] = rj;
        acc += rj * rj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_rr[blockIdx.x] = acc;
}

__global__ void k_cg_final(const CgScalars* st, const double* __restrict__ s, const double* __restrict__ x, i64 P,
                           double* __restrict__ coef) {
    const bool zero_rhs = (st->done == 3);   // scipy: `if bnrm2 == 0: return b, 0`
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x)
        coef[j] = zero_rhs ? 0.0 : __dmul_rn(s[j], x[j]);
}

static int P_grid(i64 P) {
    i64 g = (P + 1023) / 1024; if (g < 1) g = 1; if (g > RED_MAX) g = RED_MAX; return (int)g;
}
static int N_grid(i64 n) { return P_grid(n); }

// one application of the operator to the vector whose scaled image is already in m->sv
// (k_prepare / k_cg_dir wrote sv and the shift partials): q = D.v + s.X'(omega.(X sv))
static int apply_operator(bb_mat* m, const double* vP, int count_iter) {
    bb_ctx* ctx = m->ctx;
    const int* done = &m->cg->done;
    BB_TRY(bb_op_dot_flag(m, 1, done));
    P2PView view;
    const bool p2p = (ctx->nranks > 1) && bb_p2p_view(ctx, m->p + 1, &view);
    BB_TRY(bb_op_tdot_flag(m, m->w_n, true, done, /*fuse_reduce_into_consumer=*/p2p));
    // red_pq holds one partial per block and k_cg_update sums P_grid(P) of them: keep that grid (<= RED_MAX);
    k_cg_q<<<P_grid(m->P), 256, 0, ctx->stream>>>(m->cg, m->traw, m->col_offset, m->add_intercept, m->P, vP,
                                                  m->s, m->D, m->q, m->red + RED_PQ * RED_MAX, count_iter,
                                                  p2p ? m->p2p_view_dev : nullptr);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

static int cg_iteration(bb_mat* m) {
    bb_ctx* ctx = m->ctx;
    const int gP = P_grid(m->P);
    k_cg_dir<<<gP, 256, 0, ctx->stream>>>(m->cg, m->red + RED_RR * RED_MAX, gP, m->P, m->add_intercept, m->r, m->pvec, m->s,
                                          m->col_offset, m->sv, m->red + RED_SHIFT * RED_MAX);
    BB_LAUNCHED(ctx);
    BB_TRY(apply_operator(m, m->pvec, 1));
    k_cg_update<<<gP, 256, 0, ctx->stream>>>(m->cg, m->red + RED_PQ * RED_MAX, gP, m->P, m->x, m->r, m->pvec, m->q,
                                             m->red + RED_RR * RED_MAX);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// Fused form: the stop test / direction update of iteration k+1 is the tail of iteration k's P-side kernel, so one
// iteration is  X (s.p) -> omega. -> X' -> [collect + all-reduce + q + update + test + direction]  = 4 launches.
static int cg_iteration_fused(bb_mat* m) {
    const int* done = &m->cg->done;
    const bool precollect = bb_pside_precollect(m);
    if (!m->is_sparse && m->dense_stream) {
        BB_TRY(bb_dense_fused(m, done));          // X read once: 8 n p bytes instead of 16 n p
        if (precollect) BB_TRY(bb_op_collect_local(m, done));
        return bb_pside_enqueue(m);
    }
    BB_TRY(bb_op_dot_flag(m, 1, done));
    if (precollect) BB_TRY(bb_op_tdot_local(m, m->w_n, done));
    else if (m->is_sparse) BB_TRY(bb_launch_spmv(m, &m->ftdot, m->w_n, done, /*skip_overflow_add=*/bb_pside_folds_overflow(m)));
    else BB_TRY(bb_dense_tdot(m, m->w_n, done));
    return bb_pside_enqueue(m);
}

// compute z = X'(omega.y_gaussian) into m->z (device)
static int compute_z(bb_mat* m) {
    bb_ctx* ctx = m->ctx;
    if (!m->has_outcome) { bb_set_error("z == NULL requires bb_set_outcome"); return BB_ERR_STATE; }
    if (!m->is_linear && m->zk_valid) {
        BB_CUDA(cudaMemcpyAsync(m->z, m->zk, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return BB_OK;
    }
    k_z_weight<<<N_grid(m->n), 256, 0, ctx->stream>>>(m->n_trial, m->n_success, m->is_linear,
                                                      m->use_omega_scalar ? nullptr : m->omega, m->omega_scalar, m->n, m->u_n);
    BB_LAUNCHED(ctx);
    BB_TRY(bb_op_tdot(m, m->u_n));
    BB_TRY(bb_op_tdot_finish(m, m->z));
    if (!m->is_linear) {
        BB_CUDA(cudaMemcpyAsync(m->zk, m->z, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        m->zk_valid = 1;
    }
    return BB_OK;
}

// The solver proper: everything between "inputs are on the device" and "coef is in m->out_P".
// Inputs in device memory: m->pps, m->x0, m->s, m->z, the observation precisions (vector or scalar), and -- for
// injected noise -- m->eps_n / m->eps_P.  No host synchronisation except the {iter, done} polls.
static int cg_core(bb_mat* m, double atol, int maxiter, int philox, uint64_t seed, uint64_t offset) {
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const double* om = m->use_omega_scalar ? nullptr : m->omega;
    // the captured iteration graph bakes pointer arguments: rebuild it when omega switches between
    // the vector and the scalar representation; the scalar VALUE travels through device memory
    if (m->cg_graph && m->cg_graph_scalar_mode != m->use_omega_scalar) {
        cudaGraphExecDestroy(m->cg_graph);
        m->cg_graph = nullptr;
    }
    m->cg_graph_scalar_mode = m->use_omega_scalar;
    BB_CUDA(cudaMemcpyAsync(m->omega_scalar_dev, &m->omega_scalar, sizeof(double), cudaMemcpyHostToDevice, st));
    // right-hand side
    k_rhs_noise<<<N_grid(m->n), 256, 0, st>>>(om, m->omega_scalar, m->eps_n, philox, seed, offset, m->row_offset, m->n, m->u_n);
    BB_LAUNCHED(ctx);
    BB_TRY(bb_op_tdot(m, m->u_n));
    const int gP = P_grid(m->P);
    k_cg_init<<<gP, 256, 0, st>>>(m->traw, m->col_offset, m->add_intercept, m->P, m->z, m->pps, m->s, m->x0, m->eps_P,
                                  philox, seed, offset, m->b, m->D, m->x, m->red + RED_BB * RED_MAX);
    BB_LAUNCHED(ctx);
    k_cg_scalars<<<1, 32, 0, st>>>(m->cg, m->red + RED_BB * RED_MAX, gP, atol, maxiter, m->ps_bar);
    BB_LAUNCHED(ctx);
    // initial residual r = b - A x  (exactly b when x0 == 0)
    BB_TRY(bb_op_prepare_flag(m, m->x, m->s, &m->cg->done));
    BB_TRY(apply_operator(m, m->x, 0));
    k_cg_resid<<<gP, 256, 0, st>>>(m->cg, m->b, m->q, m->P, m->r, m->red + RED_RR * RED_MAX);
    BB_LAUNCHED(ctx);

    // iterations, enqueued in chunks; the device skips work once done != 0.  One CUDA graph holds
    // CG_GRAPH_ITERS iterations (fewer graph launches; surplus iterations are early-exit kernels).
    const int CG_GRAPH_ITERS = 4;
    int total = 0;
    int first = (ctx->opt_cg_chunk > 0) ? (int)ctx->opt_cg_chunk : (m->last_n_iter > 0 ? m->last_n_iter + 1 : 8);
    int chunk = first;
    const bool use_graph = ctx->opt_use_graph != 0;
    const bool fused = bb_pside_available(m);
    if (m->cg_graph && m->cg_graph_fused != (fused ? 1 : 0)) {
        cudaGraphExecDestroy(m->cg_graph);
        m->cg_graph = nullptr;
    }
    m->cg_graph_fused = fused ? 1 : 0;
    if (fused) {
        BB_TRY(bb_pside_prepare(m));
        // stop test and direction of iteration 0; every later one is the tail of the fused kernel
        k_cg_dir<<<gP, 256, 0, st>>>(m->cg, m->red + RED_RR * RED_MAX, gP, m->P, m->add_intercept, m->r, m->pvec, m->s,
                                     m->col_offset, m->sv, m->red + RED_SHIFT * RED_MAX);
        BB_LAUNCHED(ctx);
    }
    // unfused: launch k starts with the stop test of iteration k, so maxiter + 1 launches may be needed
    const int max_launches = fused ? (maxiter > 0 ? maxiter : 1) : maxiter + 1;
    for (;;) {
        int todo = chunk;
        if (total + todo > max_launches) todo = max_launches - total;
        if (todo < 1) todo = 1;
        if (use_graph) {
            if (!m->cg_graph) {
                cudaGraph_t g = nullptr;
                const i64 l0 = ctx->launches;
                BB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rc = BB_OK;
                for (int k = 0; k < CG_GRAPH_ITERS && rc == BB_OK; ++k) rc = fused ? cg_iteration_fused(m) : cg_iteration(m);
                cudaError_t e = cudaStreamEndCapture(st, &g);
                m->cg_graph_launches = (int)(ctx->launches - l0);
                ctx->launches = l0;
                if (rc != BB_OK) { if (g) cudaGraphDestroy(g); return rc; }
                if (e != cudaSuccess) { bb_set_error("graph capture: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
                e = cudaGraphInstantiate(&m->cg_graph, g, 0);
                cudaGraphDestroy(g);
                if (e != cudaSuccess) { bb_set_error("graph instantiate: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
            }
            const int ngraphs = (todo + CG_GRAPH_ITERS - 1) / CG_GRAPH_ITERS;
            for (int k = 0; k < ngraphs; ++k) BB_CUDA(cudaGraphLaunch(m->cg_graph, st));
            ctx->launches += (i64)ngraphs * m->cg_graph_launches;
            todo = ngraphs * CG_GRAPH_ITERS;
        } else {
            for (int k = 0; k < todo; ++k) BB_TRY(fused ? cg_iteration_fused(m) : cg_iteration(m));
        }
        total += todo;
        BB_CUDA(cudaMemcpyAsync(m->cg_host, m->cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        if (m->cg_host->done != 0) break;
        if (total >= max_launches) break;   // cannot happen: the last allowed launch sets done=2
        chunk = (ctx->opt_cg_chunk > 0) ? (int)ctx->opt_cg_chunk : CG_GRAPH_ITERS;
    }
    k_cg_final<<<gP, 256, 0, st>>>(m->cg, m->s, m->x, m->P, m->out_P);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// a peer that never published (lost rank, mismatched call sequence) raises the exchange's error flag instead of hanging
// the GPU: surface it as soon as the solve returns
static int check_exchange_health(bb_ctx* ctx) {
    if (ctx->nranks <= 1 || !ctx->p2p_ready) return BB_OK;
    int ready = 0, err = 0;
    BB_TRY(bb_comm_p2p_status(ctx, &ready, &err));
    if (err != 0) { bb_set_error("peer-memory exchange timed out waiting for a rank (the ranks must issue the same sequence of calls)"); return BB_ERR_STATE; }
    return BB_OK;
}

extern "C" int bb_cg_sample(bb_mat* m, const double* omega, const double* prior_prec_sqrt,
                            const double* z, const double* x0, const double* precond_scale,
                            double atol, int maxiter, int noise_mode,
                            const double* eps1, const double* eps2, uint64_t seed, uint64_t offset,
                            double* coef_out, int* n_iter, int* info, double* stats) {
    BB_ARG(m && prior_prec_sqrt && x0 && precond_scale && coef_out, "null pointer");
    BB_ARG(noise_mode == BB_NOISE_PHILOX || (eps1 && eps2), "BB_NOISE_INJECT needs eps1 and eps2");
    BB_ARG(maxiter >= 0, "maxiter");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const size_t Pb = (size_t)m->P * sizeof(double), nb = (size_t)m->n * sizeof(double);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (stats) { BB_CUDA(cudaEventCreate(&ev0)); BB_CUDA(cudaEventCreate(&ev1)); BB_CUDA(cudaEventRecord(ev0, st)); }

    if (omega) { BB_CUDA(cudaMemcpyAsync(m->omega, omega, nb, cudaMemcpyHostToDevice, st)); m->use_omega_scalar = 0; }
    BB_CUDA(cudaMemcpyAsync(m->pps, prior_prec_sqrt, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->x0, x0, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->s, precond_scale, Pb, cudaMemcpyHostToDevice, st));
    if (z) BB_CUDA(cudaMemcpyAsync(m->z, z, Pb, cudaMemcpyHostToDevice, st));
    else BB_TRY(compute_z(m));
    const int philox = (noise_mode == BB_NOISE_PHILOX);
    if (!philox) {
        BB_CUDA(cudaMemcpyAsync(m->eps_n, eps1, nb, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(m->eps_P, eps2, Pb, cudaMemcpyHostToDevice, st));
    }
    BB_TRY(cg_core(m, atol, maxiter, philox, seed, offset));
    BB_CUDA(cudaMemcpyAsync(coef_out, m->out_P, Pb, cudaMemcpyDeviceToHost, st));
    if (stats) BB_CUDA(cudaEventRecord(ev1, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    BB_TRY(check_exchange_health(ctx));
    const int done = m->cg_host->done;
    m->last_n_iter = m->cg_host->iter;
    if (n_iter) *n_iter = m->cg_host->iter;
    if (info) *info = (done == 1 || done == 3) ? 0 : maxiter;
    if (stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        stats[0] = m->cg_host->bnorm; stats[1] = m->cg_host->rnorm; stats[2] = (double)ms;
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
    return BB_OK;
}


// =====================================================================================================
// Device-resident P-side Gibbs state.  The local scales lambda and the running summaries
// (mean, second moment) of the prior-scaled coefficients live on the device; the vectors the CG sampler needs
// (prior_prec_sqrt, initial guess, preconditioner scale) are formed there, and after the draw the summaries are
// updated and the few sums the host needs (for tau and the log-posterior) are reduced there.  Every element-wise
// expression mirrors the host code's rounding (reg_coef_sampler.py:75-96,194-201; cg_sampler.py:123-138;
// reg_coef_posterior_summarizer.py:12-41,93-124), so the first draw of a chain is bit-identical to the host path.
__device__ __forceinline__ double prior_scale_dev(double gscale, double lscale, double slab) {
    double raw = __dmul_rn(gscale, lscale);
    if (!isinf(slab)) {
        double r = raw / slab;
        raw = raw / sqrt(__dadd_rn(1.0, __dmul_rn(r, r)));
    }
    return raw;
}

__global__ void k_state_pre(i64 P, int k, double gscale, double slab, long long n_avg,
                            const double* __restrict__ prior_sd, const double* __restrict__ lscale,
                            const double* __restrict__ mean, const double* __restrict__ square,
                            double* __restrict__ pps, double* __restrict__ x0, double* __restrict__ s) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        if (j < k) {
            pps[j] = 1.0 / prior_sd[j];
            x0[j] = mean[j];
            double sd = 1.0;
            if (n_avg > 1) {                       // reg_coef_posterior_summarizer.py:105-124
                double nn = (double)n_avg;
                double var = __dmul_rn(nn / (nn - 1.0), __dsub_rn(square[j], __dmul_rn(mean[j], mean[j])));
                double w = (nn - 1.0) / (nn - 1.0 + 5.0);
                sd = sqrt(__dadd_rn(__dmul_rn(w, var), __dmul_rn(1.0 - w, 1.0)));
            }
            s[j] = __dmul_rn(2.0, sd);
        } else {
            double sc = prior_scale_dev(gscale, lscale[j - k], slab);
            double pj = 1.0 / sc;
            pps[j] = pj;
            x0[j] = __dmul_rn(mean[j], sc);
            s[j] = 1.0 / pj;                       // cg_sampler.py:129  prior_prec_sqrt ** -1
        }
    }
}

// summaries <- new draw; partial sums: [0] sum_{j>=k} |b_j|^alpha, [1] #nonzero among j>=k, [2] sum (b_j/slab)^2,
// [3] sum_{j<k} (b_j/prior_sd_j)^2
__global__ void k_state_post(i64 P, int k, double gscale, double slab, double w, double alpha,
                             const double* __restrict__ prior_sd, const double* __restrict__ lscale,
                             const double* __restrict__ coef, double* __restrict__ mean, double* __restrict__ square,
                             double* __restrict__ red /* [4][RED_MAX] */) {
    __shared__ double sm[33];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const bool use_slab = !isinf(slab);
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        const double b = coef[j];
        double theta = b;
        if (j >= k) {
            theta = b / prior_scale_dev(gscale, lscale[j - k], slab);
            double ab = fabs(b);
            a0 += (alpha == 0.5) ? sqrt(ab) : pow(ab, alpha);
            a1 += (b != 0.0) ? 1.0 : 0.0;
        } else {
            double r = b / prior_sd[j];
            a3 += r * r;
        }
        if (use_slab) { double r = b / slab; a2 += r * r; }
        // mean <- w*theta + (1-w)*mean ; square <- w*theta^2 + (1-w)*square
        mean[j] = __dadd_rn(__dmul_rn(theta, w), __dmul_rn(mean[j], 1.0 - w));
        square[j] = __dadd_rn(__dmul_rn(__dmul_rn(theta, theta), w), __dmul_rn(square[j], 1.0 - w));
    }
    a0 = block_sum(a0, sm); a1 = block_sum(a1, sm); a2 = block_sum(a2, sm); a3 = block_sum(a3, sm);
    if (threadIdx.x == 0) {
        red[0 * RED_MAX + blockIdx.x] = a0; red[1 * RED_MAX + blockIdx.x] = a1;
        red[2 * RED_MAX + blockIdx.x] = a2; red[3 * RED_MAX + blockIdx.x] = a3;
    }
}

__global__ void k_state_sums(const double* __restrict__ red, int nred, double* __restrict__ sums) {
    for (int q = 0; q < 4; ++q) {
        double t = warp_sum_partials(red + q * RED_MAX, nred);
        if (threadIdx.x == 0) sums[q] = t;
    }
}

extern "C" int bb_state_init(bb_mat* m, int n_unshrunk, const double* prior_sd_unshrunk, double slab_size) {
    BB_ARG(m && n_unshrunk >= 0 && n_unshrunk <= m->P && (n_unshrunk == 0 || prior_sd_unshrunk), "mat/n_unshrunk/prior_sd");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    if (!m->st_lscale) {
        const size_t Pb = (size_t)(m->P + 1) * sizeof(double);
        BB_CUDA(cudaMalloc((void**)&m->st_lscale, Pb + 8 * sizeof(double)));      // + the three counters of the sharded draw
        BB_CUDA(cudaMalloc((void**)&m->st_mean, Pb));
        BB_CUDA(cudaMalloc((void**)&m->st_square, Pb));
        BB_CUDA(cudaMalloc((void**)&m->st_prior_sd, Pb));
        BB_CUDA(cudaMalloc((void**)&m->st_sums, 8 * sizeof(double)));
    }
    m->st_k = n_unshrunk;
    m->st_slab = slab_size;
    m->st_n_avg = 0;
    if (n_unshrunk > 0)
        BB_CUDA(cudaMemcpyAsync(m->st_prior_sd, prior_sd_unshrunk, (size_t)n_unshrunk * sizeof(double), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaStreamSynchronize(st));
    m->st_ready = 0;
    return BB_OK;
}

extern "C" int bb_state_set(bb_mat* m, const double* lscale, const double* mean, const double* square, int64_t n_averaged) {
    BB_ARG(m && lscale && mean && square && n_averaged >= 0, "null pointer");
    if (!m->st_lscale) { bb_set_error("bb_state_init first"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    const size_t Pb = (size_t)m->P * sizeof(double);
    BB_CUDA(cudaMemcpyAsync(m->st_lscale, lscale, (size_t)(m->P - m->st_k) * sizeof(double), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->st_mean, mean, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->st_square, square, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaStreamSynchronize(st));
    m->st_n_avg = n_averaged;
    m->st_ready = 1;
    return BB_OK;
}

extern "C" int bb_state_get(bb_mat* m, double* lscale, double* mean, double* square, int64_t* n_averaged) {
    BB_ARG(m != nullptr, "mat");
    if (!m->st_ready) { bb_set_error("device state not set"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    const size_t Pb = (size_t)m->P * sizeof(double);
    if (lscale) BB_CUDA(cudaMemcpyAsync(lscale, m->st_lscale, (size_t)(m->P - m->st_k) * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (mean) BB_CUDA(cudaMemcpyAsync(mean, m->st_mean, Pb, cudaMemcpyDeviceToHost, st));
    if (square) BB_CUDA(cudaMemcpyAsync(square, m->st_square, Pb, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    if (n_averaged) *n_averaged = m->st_n_avg;
    return BB_OK;
}

// One coefficient update with the P-side inputs formed on the device (device Philox noise only).
//   omega: host pointer, or NULL to use the resident precisions;  sums_out[4]: see k_state_post.
extern "C" int bb_cg_sample_resident(bb_mat* m, const double* omega, double gscale, double bridge_exp,
                                     double atol, int maxiter, uint64_t seed, uint64_t offset,
                                     double* coef_out, int* n_iter, int* info, double* sums_out) {
    BB_ARG(m && coef_out && sums_out && maxiter >= 0 && gscale > 0.0, "null pointer / maxiter / gscale");
    if (!m->st_ready) { bb_set_error("device state not set (bb_state_init / bb_state_set)"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const size_t Pb = (size_t)m->P * sizeof(double), nb = (size_t)m->n * sizeof(double);
    if (omega) { BB_CUDA(cudaMemcpyAsync(m->omega, omega, nb, cudaMemcpyHostToDevice, st)); m->use_omega_scalar = 0; }
    const int gP = P_grid(m->P);
    k_state_pre<<<gP, 256, 0, st>>>(m->P, m->st_k, gscale, m->st_slab, m->st_n_avg, m->st_prior_sd, m->st_lscale,
                                    m->st_mean, m->st_square, m->pps, m->x0, m->s);
    BB_LAUNCHED(ctx);
    BB_TRY(compute_z(m));
    BB_TRY(cg_core(m, atol, maxiter, 1, seed, offset));
    const double w = 1.0 / (1.0 + (double)m->st_n_avg);
    k_state_post<<<gP, 256, 0, st>>>(m->P, m->st_k, gscale, m->st_slab, w, bridge_exp, m->st_prior_sd, m->st_lscale,
                                     m->out_P, m->st_mean, m->st_square, m->red + RED_STATE * RED_MAX);
    BB_LAUNCHED(ctx);
    k_state_sums<<<1, 32, 0, st>>>(m->red + RED_STATE * RED_MAX, gP, m->st_sums);
    BB_LAUNCHED(ctx);
    BB_CUDA(cudaMemcpyAsync(coef_out, m->out_P, Pb, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(sums_out, m->st_sums, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    BB_TRY(check_exchange_health(ctx));
    m->st_n_avg += 1;
    const int done = m->cg_host->done;
    m->last_n_iter = m->cg_host->iter;
    if (n_iter) *n_iter = m->cg_host->iter;
    if (info) *info = (done == 1 || done == 3) ? 0 : maxiter;
    return BB_OK;
}
