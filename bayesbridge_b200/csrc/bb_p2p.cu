// One-shot all-reduce over NVLink peer memory (no NCCL on the hot path).
//
// Every rank owns one cudaMalloc'ed exchange buffer that the other ranks of the box map through CUDA IPC:
//     [ flags: 32 x u64 = 256 B ][ slot 0: cap doubles ][ slot 1: cap doubles ][ inbox: cap ][ result: cap ]
// (inbox / result and flag words 8..23 belong to the two-shot exchange of the fused CG iteration, bb_pside.cu)
// A reduction of `count` doubles is two kernels on the library stream:
//   publish : copy the local partial vector into slot (seq & 1); the LAST block to finish makes the data visible
//             system-wide and stores seq+1 into flags[my_rank] of EVERY peer (st.release.sys over NVLink);
//   reduce  : one thread per block spins until all nranks flags in the local buffer have reached seq+1
//             (ld.acquire.sys), then every thread sums the nranks partials in RANK ORDER with uncached peer loads
//             -> the result is bit-identical on every rank, which is what keeps the replicated CG recurrences
//             (and the stopping decision) in lock-step without any further exchange.
// Two slots are enough: a rank can only start publication k+2 after it has consumed publication k+1, which every
// peer has published only after it finished reading publication k.  The spin has an iteration cap: a lost peer
// raises an error flag instead of hanging the GPU.
#include "bb_internal.cuh"
#include <stdlib.h>

__global__ void k_p2p_publish(const double* __restrict__ src, i64 count, P2PView v) {
    double* slot = p2p_publish_slot(v);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) slot[i] = src[i];
    p2p_publish_done(v);
}

__global__ void k_p2p_reduce(double* __restrict__ dst, i64 count, P2PView v) {
    __shared__ int ok;
    if (!p2p_wait_all(v, &ok)) return;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x)
        dst[i] = p2p_sum(v, i);
}

// ---- host side ------------------------------------------------------------------------------------
struct bb_p2p {
    double* own;                 // exchange buffer of this rank
    double** peers_host;         // mapped base pointers (own for self)
    double** peers_dev;
    P2PState* state;
    i64 cap;
    int nranks, rank;
};

extern "C" int bb_comm_p2p_export(bb_ctx* c, int64_t capacity, char* handle_out_64) {
    BB_ARG(c && handle_out_64 && capacity > 0, "ctx/handle/capacity");
    BB_ARG(c->nranks > 1, "bb_comm_init first");
    BB_CUDA(cudaSetDevice(c->device));
    if (c->p2p) { bb_set_error("p2p exchange already initialised"); return BB_ERR_STATE; }
    bb_p2p* p = (bb_p2p*)calloc(1, sizeof(bb_p2p));
    p->cap = (capacity + 31) & ~(i64)31;
    p->nranks = c->nranks;
    p->rank = c->rank;
    size_t bytes = (size_t)(32 + 4 * p->cap) * sizeof(double);
    BB_CUDA(cudaMalloc((void**)&p->own, bytes));
    BB_CUDA(cudaMemset(p->own, 0, bytes));
    BB_CUDA(cudaMalloc((void**)&p->state, sizeof(P2PState)));
    BB_CUDA(cudaMemset(p->state, 0, sizeof(P2PState)));
    BB_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    BB_CUDA(cudaIpcGetMemHandle(&h, p->own));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    memcpy(handle_out_64, &h, 64);
    c->p2p = p;
    return BB_OK;
}

extern "C" int bb_comm_p2p_attach(bb_ctx* c, const char* handles_nranks_x_64) {
    BB_ARG(c && handles_nranks_x_64, "ctx/handles");
    bb_p2p* p = (bb_p2p*)c->p2p;
    BB_ARG(p != nullptr, "bb_comm_p2p_export first");
    BB_CUDA(cudaSetDevice(c->device));
    p->peers_host = (double**)calloc((size_t)p->nranks, sizeof(double*));
    for (int q = 0; q < p->nranks; ++q) {
        if (q == p->rank) { p->peers_host[q] = p->own; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles_nranks_x_64 + 64 * q, 64);
        void* ptr = nullptr;
        BB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->peers_host[q] = (double*)ptr;
    }
    BB_CUDA(cudaMalloc((void**)&p->peers_dev, (size_t)p->nranks * sizeof(double*)));
    BB_CUDA(cudaMemcpy(p->peers_dev, p->peers_host, (size_t)p->nranks * sizeof(double*), cudaMemcpyHostToDevice));
    c->p2p_ready = 1;
    return BB_OK;
}

int bb_p2p_free(bb_ctx* c) {
    bb_p2p* p = (bb_p2p*)c->p2p;
    if (!p) return BB_OK;
    if (p->peers_host) {
        for (int q = 0; q < p->nranks; ++q)
            if (q != p->rank && p->peers_host[q]) cudaIpcCloseMemHandle(p->peers_host[q]);
        free(p->peers_host);
    }
    if (p->peers_dev) cudaFree(p->peers_dev);
    if (p->state) cudaFree(p->state);
    if (p->own) cudaFree(p->own);
    free(p);
    c->p2p = nullptr;
    c->p2p_ready = 0;
    return BB_OK;
}

i64 bb_p2p_capacity(bb_ctx* c) {
    bb_p2p* p = (bb_p2p*)c->p2p;
    return (p && c->p2p_ready) ? p->cap : 0;
}

bool bb_p2p_view(bb_ctx* c, i64 count, P2PView* out) {
    bb_p2p* p = (bb_p2p*)c->p2p;
    if (!p || !c->p2p_ready || (c->opt_allreduce_p2p == 0 && !c->comm_local) || count > p->cap) return false;
    out->peer_base = p->peers_dev;
    out->st = p->state;
    out->cap = p->cap;
    out->nranks = p->nranks;
    out->rank = p->rank;
    out->variant = (int)c->opt_p2p_variant;
    return true;
}

// the two-shot exchange needs room for nranks chunks of ceil(count/nranks) (+ rounding) doubles
bool bb_p2p_view2(bb_ctx* c, i64 count, P2PView* out) {
    bb_p2p* p = (bb_p2p*)c->p2p;
    if (!p || !c->p2p_ready || p->nranks > 8 || count + 4 * (i64)p->nranks > p->cap) return false;
    out->peer_base = p->peers_dev;
    out->st = p->state;
    out->cap = p->cap;
    out->nranks = p->nranks;
    out->rank = p->rank;
    out->variant = (int)c->opt_p2p_variant;
    return true;
}

static int reduce_grid(i64 count) {
    i64 g = (count + 1023) / 1024;
    if (g < 1) g = 1;
    if (g > 64) g = 64;
    return (int)g;
}

int bb_p2p_reduce_into(bb_ctx* c, double* dst, i64 count) {
    P2PView v;
    if (!bb_p2p_view(c, count, &v)) { bb_set_error("p2p exchange not available"); return BB_ERR_STATE; }
    k_p2p_reduce<<<reduce_grid(count), 256, 0, c->stream>>>(dst, count, v);
    BB_LAUNCHED(c);
    return BB_OK;
}

// in-place sum over ranks of dbuf[0..count) on ctx->stream; false => caller falls back to NCCL
bool bb_p2p_allreduce(bb_ctx* c, double* dbuf, i64 count, const int* done_flag, int* rc_out) {
    (void)done_flag;
    P2PView v;
    *rc_out = BB_OK;
    if (!bb_p2p_view(c, count, &v)) return false;
    k_p2p_publish<<<reduce_grid(count), 256, 0, c->stream>>>(dbuf, count, v);
    k_p2p_reduce<<<reduce_grid(count), 256, 0, c->stream>>>(dbuf, count, v);
    c->launches += 2;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { bb_set_error("p2p allreduce launch: %s", cudaGetErrorString(e)); *rc_out = BB_ERR_CUDA; }
    return true;
}

extern "C" int bb_comm_p2p_status(bb_ctx* c, int* ready, int* error) {
    BB_ARG(c != nullptr, "ctx");
    bb_p2p* p = (bb_p2p*)c->p2p;
    if (ready) *ready = (p && c->p2p_ready) ? 1 : 0;
    if (error) {
        *error = 0;
        if (p && p->state) {
            P2PState h;
            BB_CUDA(cudaSetDevice(c->device));
            BB_CUDA(cudaStreamSynchronize(c->stream));
            BB_CUDA(cudaMemcpy(&h, p->state, sizeof(P2PState), cudaMemcpyDeviceToHost));
            *error = (int)h.error;
        }
    }
    return BB_OK;
}
