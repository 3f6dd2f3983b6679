// The P-side of one CG iteration as ONE kernel, with the all-reduce of [sum w; X'w] inside it.
//
// Reference: the body of scipy.sparse.linalg.cg's loop (called from reg_coef_sampler/cg_sampler.py:77-80) after the
// two matrix products of the operator (cg_sampler.py:105-112):
//     t = [sum w; X'w]  (summed over the row shards)      q = D.p + s.t        alpha = rho / (p.q)
//     x += alpha p ; r -= alpha q ; rho' = r.r            stop test ; p = r + (rho'/rho) p ; sv = s.p
// In the unfused path this is k_sell_ovf_add + k_tdot_collect + ncclAllReduce + k_cg_q + k_cg_update + k_cg_dir: six
// launches of 5-30 us each around 100 KB ... 1 MB of data.  Here a small persistent grid (<= 128 CTAs, all co-resident)
// walks the phases with grid barriers in between; every thread owns the same coefficient indices in every phase, so
// only the reductions and the exchanged vector cross CTAs.
//
// Exchange (nranks > 1): a deterministic two-shot all-reduce over NVLink peer memory, PUSH only.  Rank r owns chunk r
// of the vector.  Phase A: every rank stores chunk c of its local partial vector straight into rank c's inbox (slot
// = sender) and then raises its flag at every peer.  Phase B: rank r adds the nranks slots of its chunk in RANK ORDER
// and stores the sum into every rank's result buffer, then raises its second flag everywhere.  Phase C: everybody
// waits for the nranks result flags.  (nranks-1)/nranks of the vector leaves and enters every GPU twice, nothing is
// read over the link, and the result is bit-identical on all ranks, which keeps the replicated recurrences and the
// stopping decision in lock-step.  Buffers are reused by consecutive exchanges without a further handshake: a rank
// can only push into a peer's inbox for exchange k+1 after it has passed phase C of exchange k, which needs the
// peer's phase-B flag, which the peer raised after reading its inbox; the same argument covers the result buffer.
//
// Default form of the same data flow: flag-in-data lines (ll_store / ll_load_wait below, option pside_ll).  A value is
// announced by the tag inside its own 16-byte line, so phases A-C have no flags, fences or grid barriers; a reader spins on
// exactly the lines it needs.  Buffer reuse without a handshake, line by line: (inbox) rank a stores lines of exchange k+1
// only after its p.q grid barrier of iteration k, i.e. after ALL its threads have read their result lines of exchange k;
// result line j of exchange k exists only after owner c has read the N inbox lines of element j -- so every inbox line of
// exchange k has been consumed by then.  (result region) owner c stores result line j of exchange k+1 after reading the
// inbox line of j sent by rank q for k+1, which q sent after its p.q barrier of iteration k, i.e. after reading result line
// j of exchange k.  Tags are exchange numbers, so a stale line never matches.
#include "bb_internal.cuh"

constexpr int PS_THREADS = 512;
constexpr int PS_MAX_CTAS = 128;
constexpr int PS_MAX_RANKS = 8;           // flag words 8..15 (inbox) and 16..23 (result) of the 32-word flag area

struct PsideArgs {
    CgScalars* st;
    // local [sum w; X'w]: slab partial sums (sparse) or row-block partial sums (dense), plus partial sums of w
    double* part; int nslab; i64 p;
    const double* red_w; int nred_w;
    const int* ovf_piece; const int* ovf_first; int n_ovf; i64 V;     // overflow fragments of the sliced format
    double* traw;                                                    // [p + 1] result when nranks == 1
    int precollected;                                                // traw already holds the local vector (k_tdot_collect ran)
    // CG vectors
    const double* c; int icpt; i64 P;
    double* pvec; double* q; double* x; double* r; double* sv;
    const double* s; const double* D;
    double* red_pq; double* red_rr; double* red_shift; int nshift;
    unsigned long long* bar;
    const P2PView* view;                                             // nullptr: no exchange
    int lean_barrier;                                                // option pside_barrier
    int ll;                                                          // exchange with flag-in-data lines (option pside_ll)
};

__device__ __forceinline__ unsigned long long ps_ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the grid are co-resident (grid <= PS_MAX_CTAS <= number of SMs, one stream): a counting barrier.
// Arrival is a release reduction WITHOUT a return value (the CTA's writes, ordered before it by the block barrier, are
// visible to whoever observes the count), so the poll starts at once instead of after the atomic's round trip; the poll is
// an acquire load.  (Round-2 measurement: the fused kernel is four to five of these per CG iteration, each followed by a
// fixed-order sum of the partials -- at the N = 8 shard size it is the largest single kernel, profiles/r02_launches.md.)
__device__ __forceinline__ void ps_grid_barrier(unsigned long long* ctr, unsigned long long target, int lean = 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (lean) {
            asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
            while (ps_ld_acquire_gpu(ctr) < target) { }
        } else {                                  // the first form (option pside_barrier = 0), kept for the A/B
            __threadfence();
            atomicAdd(ctr, 1ull);
            while (ps_ld_acquire_gpu(ctr) < target) { }
            __threadfence();
        }
    }
    __syncthreads();
}

// fixed-order sum of `count` partials written by other CTAs of this launch (L1 bypassed); all 32 lanes call
__device__ __forceinline__ double ps_sum_partials(const double* buf, int count) {
    const int lane = threadIdx.x & 31;
    double t = 0.0;
    for (int i = lane; i < count; i += 32) t += __ldcg(buf + i);
    return warp_sum(t);
}

// threads 0..n-1 wait until flags[0..n) >= want; a lost peer raises the error flag instead of hanging the GPU
__device__ __forceinline__ void ps_wait_flags(const unsigned long long* flags, int n, unsigned long long want, P2PState* st) {
    if ((int)threadIdx.x < n) {
        unsigned long long spins = 0;
        while (ld_acquire_sys_u64(flags + threadIdx.x) < want) {
            if (++spins > (1ull << 26)) { st->error = 1u; break; }
            __nanosleep(20);
        }
    }
    __syncthreads();
}

// ---- flag-in-data ("LL") lines: the exchange without flag round trips -------------------------------------------------
// An exchanged double travels as one 16-byte line {lo32, tag, hi32, tag}, tag = low 32 bits of the exchange number, written
// with ONE 16-byte store.  8-byte halves of such a store are never torn, so a reader that sees `tag` in both halves has the
// whole value: the data announce themselves, and neither a grid barrier + system fence + flag store on the sending side nor a
// flag wait on the receiving side is needed.  (The technique of NCCL's LL protocol, here with fp64 payloads.)  Twice the
// bytes over NVLink (3.2 MB per rank and exchange at p = 1e5) for two flag round trips and two grid barriers less.
__device__ __forceinline__ void ll_store(double* line, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};"
                 ::"l"(line), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag) : "memory");
}
// spins until the line carries `tag`; a lost peer raises the error flag (and yields 0) instead of hanging the GPU
__device__ __forceinline__ double ll_load_wait(const double* line, unsigned tag, P2PState* st) {
    unsigned x, f0, y, f1;
    unsigned long long spins = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(f0), "=r"(y), "=r"(f1) : "l"(line) : "memory");
        if (f0 == tag && f1 == tag) break;
        if (++spins > (1ull << 24)) { st->error = 1u; return 0.0; }
    }
    return __longlong_as_double((long long)(((unsigned long long)y << 32) | (unsigned long long)x));
}

__global__ void __launch_bounds__(PS_THREADS)
k_cg_pside(const PsideArgs a) {
    pdl_trigger();
    pdl_wait();           // everything this kernel reads was written by the kernels before it
    CgScalars* st = a.st;
    if (st->done) return;
    __shared__ double sm[33];
    const int G = (int)gridDim.x, tid = (int)threadIdx.x, lane = tid & 31;
    const i64 gtid = (i64)blockIdx.x * PS_THREADS + tid, gthreads = (i64)G * PS_THREADS;
    const bool lead = (blockIdx.x == 0 && tid == 0);
    unsigned long long target = st->bar_base;
    const int it0 = st->iter;
    const double rho_old = st->rho[it0 & 1];
    const double atol_eff = st->atol_eff;
    const int maxiter = st->maxiter;
    const i64 L = a.p + 1;

    int N = 1, me = 0;
    i64 Cw = L, cap = 0;
    double* const* peer = nullptr;
    P2PState* pst = nullptr;
    unsigned long long want = 0;
    if (a.view != nullptr) {
        const P2PView pv = *a.view;
        N = pv.nranks; me = pv.rank; peer = pv.peer_base; pst = pv.st; cap = pv.cap;
        Cw = (((L + N - 1) / N) + 1) & ~(i64)1;
        want = pst->seq2 + 1ull;
        // an earlier exchange of this solve timed out (a rank is gone): stop the solve at once instead of spinning through
        // every remaining iteration; the host raises as soon as the call returns (check_exchange_health)
        if (pst->error != 0u) {
            if (lead) st->done = 2;
            return;
        }
    }
    const i64 inbox_off = 32 + 2 * cap, result_off = 32 + 3 * cap;

    // ---- phase A0: fold the overflow fragments of long columns into their slots (bb_sell.cu) ----
    if (a.n_ovf > 0 && !a.precollected) {
        const i64 gwarp = gtid >> 5, gwarps = gthreads >> 5;
        for (i64 i = gwarp; i < a.n_ovf; i += gwarps) {
            const int f0 = a.ovf_first[i], f1 = a.ovf_first[i + 1];
            const double t = warp_sum_partials(a.part + a.V + f0, f1 - f0);
            if (lane == 0) a.part[a.ovf_piece[i]] += t;
        }
        target += (unsigned long long)G;
        ps_grid_barrier(a.bar, target, a.lean_barrier);
    }

    // ---- phase A1: local [sum w; X'w]; with an exchange, pushed chunk by chunk into the owners' inboxes ----
    const bool ll = (N > 1) && (a.ll != 0);
    const unsigned tag = (unsigned)want;
    if (a.precollected) {
        // many slabs (few ranks): a separate, much wider launch has already summed them into traw
        if (N > 1) {
            for (i64 j = gtid; j < L; j += gthreads) {
                const i64 owner = j / Cw;
                const i64 e = (i64)me * Cw + (j - owner * Cw);
                if (ll) ll_store(peer[owner] + inbox_off + 2 * e, a.traw[j], tag);
                else peer[owner][inbox_off + e] = a.traw[j];
            }
        }
    } else {
        if (blockIdx.x == 0 && tid < 32) {
            const double sw = warp_sum_partials(a.red_w, a.nred_w);
            if (lane == 0) {
                if (N == 1) a.traw[0] = sw;
                else if (ll) ll_store(peer[0] + inbox_off + 2 * ((i64)me * Cw), sw, tag);
                else peer[0][inbox_off + (i64)me * Cw] = sw;
            }
        }
        for (i64 j = 1 + gtid; j < L; j += gthreads) {
            const double* col = a.part + (j - 1);
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            int rr = 0;
            for (; rr + 3 < a.nslab; rr += 4) {
                t0 += __ldcg(col + (i64)rr * a.p); t1 += __ldcg(col + (i64)(rr + 1) * a.p);
                t2 += __ldcg(col + (i64)(rr + 2) * a.p); t3 += __ldcg(col + (i64)(rr + 3) * a.p);
            }
            for (; rr < a.nslab; ++rr) t0 += __ldcg(col + (i64)rr * a.p);
            const double t = (t0 + t1) + (t2 + t3);
            if (N == 1) {
                a.traw[j] = t;
            } else {
                const i64 owner = j / Cw;
                const i64 e = (i64)me * Cw + (j - owner * Cw);
                if (ll) ll_store(peer[owner] + inbox_off + 2 * e, t, tag);
                else peer[owner][inbox_off + e] = t;
            }
        }
    }
    if (!(a.precollected && N == 1) && !ll) {
        target += (unsigned long long)G;
        ps_grid_barrier(a.bar, target, a.lean_barrier);
    }

    const double* tvec = a.traw;
    if (ll) {
        // ---- phase B: reduce the own chunk in rank order as its lines arrive, push the sums to every rank ----
        double* own = peer[me];
        const i64 lo = (i64)me * Cw;
        i64 cnt = L - lo; if (cnt > Cw) cnt = Cw; if (cnt < 0) cnt = 0;
        for (i64 jj = gtid; jj < cnt; jj += gthreads) {
            double acc = 0.0;
            for (int sdr = 0; sdr < N; ++sdr) acc += ll_load_wait(own + inbox_off + 2 * ((i64)sdr * Cw + jj), tag, pst);
            for (int qq = 0; qq < N; ++qq) ll_store(peer[qq] + result_off + 2 * (lo + jj), acc, tag);
        }
        tvec = own + result_off;          // lines: entry j of the reduced vector is at tvec + 2 j, read with ll_load_wait
    } else if (N > 1) {
        double* own = peer[me];
        unsigned long long* own_flags = reinterpret_cast<unsigned long long*>(own);
        if (lead) {
            __threadfence_system();
            for (int qq = 0; qq < N; ++qq)
                st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(peer[qq]) + 8 + me, want);
        }
        // ---- phase B: reduce the own chunk in rank order, push the sums to every rank ----
        ps_wait_flags(own_flags + 8, N, want, pst);
        const i64 lo = (i64)me * Cw;
        i64 cnt = L - lo; if (cnt > Cw) cnt = Cw; if (cnt < 0) cnt = 0;
        for (i64 jj = gtid; jj < cnt; jj += gthreads) {
            double acc = 0.0;
            for (int sdr = 0; sdr < N; ++sdr) acc += __ldcv(own + inbox_off + (i64)sdr * Cw + jj);
            for (int qq = 0; qq < N; ++qq) peer[qq][result_off + lo + jj] = acc;
        }
        target += (unsigned long long)G;
        ps_grid_barrier(a.bar, target, a.lean_barrier);
        if (lead) {
            __threadfence_system();
            for (int qq = 0; qq < N; ++qq)
                st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(peer[qq]) + 16 + me, want);
        }
        // ---- phase C: wait for every chunk ----
        ps_wait_flags(own_flags + 16, N, want, pst);
        tvec = own + result_off;
    }

    // ---- q = D.p + s.t ; p.q ----
    const double sw = ll ? ll_load_wait(tvec, tag, pst) : __ldcv(tvec);
    double acc = 0.0;
    for (i64 j = gtid; j < a.P; j += gthreads) {
        // sparse_matrix.py:126-128   result = X.T.dot(v); result -= sum(v) * column_offset
        const double tj = (j < a.icpt) ? sw : (ll ? ll_load_wait(tvec + 2 * (1 + (j - a.icpt)), tag, pst) : __ldcv(tvec + 1 + (j - a.icpt)));
        const double t = (j < a.icpt) ? sw : __dsub_rn(tj, __dmul_rn(sw, a.c[j - a.icpt]));
        const double pj = a.pvec[j];
        const double qj = __dadd_rn(__dmul_rn(a.D[j], pj), __dmul_rn(a.s[j], t));
        a.q[j] = qj;
        acc += pj * qj;
    }
    acc = block_sum(acc, sm);
    if (tid == 0) a.red_pq[blockIdx.x] = acc;
    target += (unsigned long long)G;
    ps_grid_barrier(a.bar, target, a.lean_barrier);
    const double pq = ps_sum_partials(a.red_pq, G);

    // ---- alpha = rho/(p.q) ; x += alpha p ; r -= alpha q ; r.r ----
    const double alpha = rho_old / pq;
    acc = 0.0;
    for (i64 j = gtid; j < a.P; j += gthreads) {
        a.x[j] = __dadd_rn(a.x[j], __dmul_rn(alpha, a.pvec[j]));
        const double rj = __dsub_rn(a.r[j], __dmul_rn(alpha, a.q[j]));
        a.r[j] = rj;
        acc += rj * rj;
    }
    acc = block_sum(acc, sm);
    if (tid == 0) a.red_rr[blockIdx.x] = acc;
    target += (unsigned long long)G;
    ps_grid_barrier(a.bar, target, a.lean_barrier);
    const double rho = ps_sum_partials(a.red_rr, G);

    // ---- stop test of the next iteration, search direction, scaled gather vector of the next product ----
    const int it = it0 + 1;
    const double rn = sqrt(rho);
    int done = 0;
    if (it >= maxiter) done = 2;            // scipy: after maxiter updates the loop ends without another test
    else if (rn < atol_eff) done = 1;
    if (!done) {
        const double beta = rho / rho_old;
        acc = 0.0;
        for (i64 j = gtid; j < a.P; j += gthreads) {
            const double pj = __dadd_rn(__dmul_rn(beta, a.pvec[j]), a.r[j]);
            a.pvec[j] = pj;
            const double xs = __dmul_rn(a.s[j], pj);
            a.sv[j] = xs;
            acc += (j < a.icpt) ? xs : -a.c[j - a.icpt] * xs;
        }
        acc = block_sum(acc, sm);
        if (tid == 0) a.red_shift[blockIdx.x] = acc;
        if (blockIdx.x == 0)                  // k_dot_finish sums `nshift` partials: the ones this grid does not write are zero
            for (int k = G + tid; k < a.nshift; k += PS_THREADS) a.red_shift[k] = 0.0;
    }
    if (lead) {
        st->iter = it;
        st->rnorm = rn;
        if (done) st->done = done; else st->rho[it & 1] = rho;
        st->bar_base = target;
        if (pst != nullptr) pst->seq2 = want;
    }
}

// ---- host side ---------------------------------------------------------------------------------
bool bb_pside_precollect(bb_mat* m);
int bb_pside_grid(bb_ctx* ctx, i64 P) {
    i64 g = (ctx->opt_pside_ctas > 0) ? ctx->opt_pside_ctas : (P + 1023) / 1024;
    i64 cap = PS_MAX_CTAS;
    if (cap > ctx->sm_count) cap = ctx->sm_count;
    if (g > cap) g = cap;
    i64 gP = (P + 1023) / 1024;            // k_dot_finish reads P_grid(P) shift partials: never write more than that
    if (gP < 1) gP = 1;
    if (g > gP) g = gP;
    if (g < 1) g = 1;
    return (int)g;
}

bool bb_pside_available(bb_mat* m) {
    bb_ctx* ctx = m->ctx;
    if (ctx->opt_cg_fused == 0) return false;
    if (ctx->nranks == 1) return true;
    P2PView v;
    return ctx->nranks <= PS_MAX_RANKS && bb_p2p_view2(ctx, m->p + 1, &v);
}

// The fused kernel folds the overflow fragments of long columns itself (one warp per column) only while there are few
// of them; a matrix with many long columns (BASELINE config 3: columns of up to 5e4 nnz) keeps the wide k_sell_ovf_add launch.
bool bb_pside_folds_overflow(bb_mat* m) {
    if (!m->is_sparse || m->ftdot.variant != 1 || m->ftdot.n_ovf_pieces == 0) return false;
    if (bb_pside_precollect(m)) return false;
    if (m->ctx->opt_pside_fold_ovf == 0) return false;
    if (m->ctx->opt_pside_fold_ovf == 1) return true;
    const i64 warps = (i64)bb_pside_grid(m->ctx, m->P) * (PS_THREADS / 32);
    return (i64)m->ftdot.n_ovf_pieces <= 4 * warps;
}

bool bb_pside_precollect(bb_mat* m) {
    const int nslab = m->is_sparse ? m->ftdot.nslab : m->dense_nblk;
    const i64 lim = m->ctx->opt_pside_collect_max > 0 ? m->ctx->opt_pside_collect_max : 8;
    return nslab > lim;
}

// uploads the exchange view (device copy read by the kernel); call outside graph capture
int bb_pside_prepare(bb_mat* m) {
    bb_ctx* ctx = m->ctx;
    if (ctx->nranks == 1) return BB_OK;
    P2PView view;
    if (!bb_p2p_view2(ctx, m->p + 1, &view)) { bb_set_error("fused CG iteration: peer-memory exchange not attached"); return BB_ERR_STATE; }
    if (m->p2p_view_valid != 1 + view.variant) {
        BB_CUDA(cudaStreamSynchronize(ctx->stream));
        BB_CUDA(cudaMemcpy(m->p2p_view_dev, &view, sizeof(P2PView), cudaMemcpyHostToDevice));
        m->p2p_view_valid = 1 + view.variant;
    }
    return BB_OK;
}

// enqueue the fused P-side kernel of one CG iteration (after the Tdot product has been launched)
int bb_pside_enqueue(bb_mat* m) {
    bb_ctx* ctx = m->ctx;
    PsideArgs a;
    memset(&a, 0, sizeof(a));
    a.st = m->cg;
    if (m->is_sparse) {
        SlabFmt* f = &m->ftdot;
        a.part = f->part; a.nslab = f->nslab;
        if (bb_pside_folds_overflow(m)) {
            a.ovf_piece = f->ovf_piece; a.ovf_first = f->ovf_first; a.n_ovf = f->n_ovf_pieces; a.V = (i64)f->nslab * f->n_seg;
        }
    } else {
        a.part = m->dense_part; a.nslab = m->dense_nblk;
    }
    a.p = m->p;
    // Summing many slab partials per column wants far more threads than this small grid has: above PS_COLLECT_MAX
    // slabs the wide k_tdot_collect launch does it (and the overflow fold before it), below it is a phase of this kernel.
    if (bb_pside_precollect(m)) {
        a.precollected = 1;
        a.n_ovf = 0;
    }
    a.red_w = m->red + RED_W * RED_MAX; a.nred_w = m->nred_w;
    a.traw = m->traw;
    a.c = m->col_offset; a.icpt = m->add_intercept; a.P = m->P;
    a.pvec = m->pvec; a.q = m->q; a.x = m->x; a.r = m->r; a.sv = m->sv; a.s = m->s; a.D = m->D;
    a.red_pq = m->red + RED_PQ * RED_MAX; a.red_rr = m->red + RED_RR * RED_MAX; a.red_shift = m->red + RED_SHIFT * RED_MAX;
    i64 gP = (m->P + 1023) / 1024; if (gP < 1) gP = 1; if (gP > RED_MAX) gP = RED_MAX;
    a.nshift = (int)gP;
    a.bar = m->ps_bar;
    a.lean_barrier = (ctx->opt_pside_barrier != 0) ? 1 : 0;
    a.view = nullptr;
    a.ll = 0;
    if (ctx->nranks > 1) {
        if (m->p2p_view_valid == 0) { bb_set_error("fused CG iteration: bb_pside_prepare first"); return BB_ERR_STATE; }
        a.view = m->p2p_view_dev;
        // flag-in-data lines need twice the room: nranks chunks of Cw lines in the inbox, p + 1 lines in the result region
        const i64 Lx = m->p + 1, Nr = ctx->nranks, Cwx = (((Lx + Nr - 1) / Nr) + 1) & ~(i64)1;
        a.ll = (ctx->opt_pside_ll != 0 && 2 * Nr * Cwx <= bb_p2p_capacity(ctx) && 2 * Lx <= bb_p2p_capacity(ctx)) ? 1 : 0;
    }
    static BBDeviceOnce attr_set = {{0, 0, 0, 0}};
    if (attr_set.first(ctx->device)) BB_CUDA(bb_prefer_max_smem(ctx, k_cg_pside));
    BB_CUDA(bb_launch(ctx, true, k_cg_pside, dim3(bb_pside_grid(ctx, m->P)), dim3(PS_THREADS), 0, a));
    BB_LAUNCHED(ctx);
    return BB_OK;
}
