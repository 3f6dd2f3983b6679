// Device-side timing of one kernel class on resident data (CUDA events on the launching stream).
#include "bb_internal.cuh"

__global__ void k_flush(double* buf, i64 n, double v) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) buf[i] = v;
}
__global__ void k_fill_test(double* buf, i64 n, double scale) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        buf[i] = scale * (1.0 + (double)(i % 7));
}

static int flush_l2(bb_ctx* ctx) {
    const size_t bytes = (size_t)512 << 20;   // > 126 MB L2
    if (!ctx->flush_buf) {
        BB_CUDA(cudaMalloc(&ctx->flush_buf, bytes));
        ctx->flush_bytes = bytes;
    }
    k_flush<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>((double*)ctx->flush_buf, (i64)(ctx->flush_bytes / 8), 1.0);
    return BB_OK;
}

extern "C" int bb_time_kernel(bb_mat* m, const char* what, int reps, int do_flush, double* ms_out) {
    BB_ARG(m && what && ms_out && reps > 0, "mat/what/ms_out/reps");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    int kind = -1;
    if (!strcmp(what, "dot")) kind = 0;
    else if (!strcmp(what, "tdot")) kind = 1;
    else if (!strcmp(what, "op")) kind = 2;
    else if (!strcmp(what, "spmv_dot")) kind = 3;     // the SpMV kernel alone (+ its fix-up)
    else if (!strcmp(what, "spmv_tdot")) kind = 4;
    else if (!strcmp(what, "exchange")) kind = 5;     // all-reduce of the (p+1)-vector (every rank must call)
    else if (!strcmp(what, "fused_op")) kind = 6;     // dense: omega.(X sv) and its X' product in ONE pass over X (+ collect)
    BB_ARG(kind >= 0, "what must be dot | tdot | op | spmv_dot | spmv_tdot | exchange | fused_op");
    BB_ARG(kind < 3 || kind >= 5 || m->is_sparse, "spmv_* needs a sparse matrix");
    BB_ARG(kind != 6 || (!m->is_sparse && m->dense_stream), "fused_op needs a dense matrix served by the streaming kernel");
    if (kind == 5) {
        // back-to-back exchanges, timed as one region (the per-exchange latency is what matters in the CG loop)
        cudaEvent_t a, b;
        BB_CUDA(cudaEventCreate(&a));
        BB_CUDA(cudaEventCreate(&b));
        for (int r = 0; r < 3; ++r) BB_TRY(bb_allreduce_dev(ctx, m->traw, m->p + 1));
        BB_CUDA(cudaStreamSynchronize(st));
        BB_CUDA(cudaEventRecord(a, st));
        for (int r = 0; r < reps; ++r) BB_TRY(bb_allreduce_dev(ctx, m->traw, m->p + 1));
        BB_CUDA(cudaEventRecord(b, st));
        BB_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        BB_CUDA(cudaEventElapsedTime(&ms, a, b));
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        *ms_out = ms / reps;
        return BB_OK;
    }
    // deterministic, non-trivial inputs
    k_fill_test<<<256, 256, 0, st>>>(m->v_P, m->P, 1e-3);
    k_fill_test<<<256, 256, 0, st>>>(m->eps_n, m->n, 1e-3);
    k_fill_test<<<256, 256, 0, st>>>(m->sv, m->P, 1e-3);
    if (!m->use_omega_scalar) {
        // keep whatever omega is resident; if never set it is zeros, which is still valid timing input
    }
    cudaEvent_t e0, e1;
    BB_CUDA(cudaEventCreate(&e0));
    BB_CUDA(cudaEventCreate(&e1));
    double total = 0.0;
    for (int r = -2; r < reps; ++r) {      // two untimed warm-ups
        if (do_flush) BB_TRY(flush_l2(ctx));
        BB_CUDA(cudaEventRecord(e0, st));
        if (kind == 0) {
            BB_TRY(bb_op_prepare(m, m->v_P, nullptr));
            BB_TRY(bb_op_dot(m, 0));
        } else if (kind == 1) {
            BB_TRY(bb_op_tdot(m, m->eps_n));
        } else if (kind == 2) {
            BB_TRY(bb_op_prepare(m, m->v_P, nullptr));
            BB_TRY(bb_op_dot(m, 1));
            BB_TRY(bb_op_tdot_flag(m, m->w_n, true, nullptr, false));
        } else if (kind == 6) {
            BB_TRY(bb_op_prepare(m, m->v_P, nullptr));
            BB_TRY(bb_dense_fused(m, nullptr));
            BB_TRY(bb_op_collect_local(m, nullptr));
        } else if (kind == 3) {
            BB_TRY(bb_launch_spmv(m, &m->fdot, m->sv + m->add_intercept, nullptr));   // the (16-byte aligned) vector dot gathers from
        } else {
            BB_TRY(bb_launch_spmv(m, &m->ftdot, m->eps_n, nullptr));
        }
        BB_CUDA(cudaEventRecord(e1, st));
        BB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        BB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 0) total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = total / reps;
    return BB_OK;
}
