// Context, error reporting, options and the NCCL communicator (loaded with dlopen so that a
// single-GPU process never needs libnccl).
#include "bb_internal.cuh"
#include <dlfcn.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void bb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* bb_last_error(void) { return g_err; }
extern "C" int bb_version(void) { return 100; }

extern "C" int bb_device_count(int* count) {
    BB_ARG(count != nullptr, "count");
    BB_CUDA(cudaGetDeviceCount(count));
    return BB_OK;
}

extern "C" int bb_init(int device, bb_ctx** out) {
    BB_ARG(out != nullptr, "out");
    BB_CUDA(cudaSetDevice(device));
    bb_ctx* c = (bb_ctx*)calloc(1, sizeof(bb_ctx));
    c->device = device;
    BB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    BB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        bb_set_error("libbbgpu is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
        free(c);
        return BB_ERR_CUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->opt_spmv_stage = 1;
    c->opt_slab_width = 0;
    c->opt_bank_permute = 2;     // most-loaded-bank-first matching (bb_sell.cu); 1: greedy; 0: canonical order
    c->opt_spmv_variant = 1;
    c->opt_rowwise_max_nnz = 1 << 21;   // measured: C1 (1e5 nnz) 8.5 / 10.4 us vs 13.7 / 13.0 us sliced; C3 (1e7 nnz) 39 / 59 us vs 39 / 32 us
    c->opt_spmv_bulk = 1;
    c->opt_cg_chunk = 0;
    c->opt_use_graph = 1;
    c->opt_cg_fused = 1;
    c->opt_pside_ctas = 0;
    c->opt_pside_barrier = 1;
    c->opt_pside_ll = 1;
    c->opt_pside_fold_ovf = -1;
    c->opt_dense_stream = 1;
    c->opt_pdl = 1;
    c->opt_sell_lpt = 1;
    c->opt_uniform_carveout = 0;   // measured (profiles/r02_pdl_ab.md): no gain at the N = 8 shard size, the dot SpMV loses 1-5 % (less L1 for its stores)
    c->opt_allreduce_p2p = 1;
    c->opt_p2p_variant = 3;      // one system fence + relaxed flag stores, parallel flag polls (fastest measured at N=8)
    c->nranks = 1;
    c->rank = 0;
    BB_CUDA(cudaEventCreate(&c->tev0));
    BB_CUDA(cudaEventCreate(&c->tev1));
    *out = c;
    return BB_OK;
}

extern "C" int bb_destroy(bb_ctx* c) {
    if (!c) return BB_OK;
    cudaSetDevice(c->device);
    bb_p2p_free(c);
    if (c->nccl_comm && c->nccl_handle) {
        typedef int (*destroy_t)(void*);
        destroy_t f = (destroy_t)dlsym(c->nccl_handle, "ncclCommDestroy");
        if (f) f(c->nccl_comm);
    }
    if (c->flush_buf) cudaFree(c->flush_buf);
    for (int i = 0; i < 4; ++i) if (c->scratch[i]) cudaFree(c->scratch[i]);
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->tev0);
    cudaEventDestroy(c->tev1);
    cudaStreamDestroy(c->stream);
    free(c);
    return BB_OK;
}

extern "C" int bb_sync(bb_ctx* c) {
    BB_ARG(c != nullptr, "ctx");
    BB_CUDA(cudaStreamSynchronize(c->stream));
    return BB_OK;
}

static i64* option_slot(bb_ctx* c, const char* name) {
    if (!strcmp(name, "spmv_stage")) return &c->opt_spmv_stage;
    if (!strcmp(name, "slab_width")) return &c->opt_slab_width;
    if (!strcmp(name, "bank_permute")) return &c->opt_bank_permute;
    if (!strcmp(name, "spmv_variant")) return &c->opt_spmv_variant;
    if (!strcmp(name, "rowwise_max_nnz")) return &c->opt_rowwise_max_nnz;
    if (!strcmp(name, "spmv_bulk")) return &c->opt_spmv_bulk;
    if (!strcmp(name, "sell_lmax")) return &c->opt_sell_lmax;
    if (!strcmp(name, "cg_chunk")) return &c->opt_cg_chunk;
    if (!strcmp(name, "use_graph")) return &c->opt_use_graph;
    if (!strcmp(name, "cg_fused")) return &c->opt_cg_fused;
    if (!strcmp(name, "pside_ctas")) return &c->opt_pside_ctas;
    if (!strcmp(name, "pside_barrier")) return &c->opt_pside_barrier;
    if (!strcmp(name, "pside_ll")) return &c->opt_pside_ll;
    if (!strcmp(name, "pside_fold_ovf")) return &c->opt_pside_fold_ovf;
    if (!strcmp(name, "pside_collect_max")) return &c->opt_pside_collect_max;
    if (!strcmp(name, "dense_stream")) return &c->opt_dense_stream;
    if (!strcmp(name, "sell_slice_cost")) return &c->opt_sell_slice_cost;
    if (!strcmp(name, "sell_partition")) return &c->opt_sell_partition;
    if (!strcmp(name, "sell_lpt")) return &c->opt_sell_lpt;
    if (!strcmp(name, "pdl")) return &c->opt_pdl;
    if (!strcmp(name, "uniform_carveout")) return &c->opt_uniform_carveout;
    if (!strcmp(name, "allreduce_p2p")) return &c->opt_allreduce_p2p;
    if (!strcmp(name, "p2p_variant")) return &c->opt_p2p_variant;
    return nullptr;
}

extern "C" int bb_set_option(bb_ctx* c, const char* name, int64_t value) {
    BB_ARG(c && name, "ctx/name");
    i64* s = option_slot(c, name);
    BB_ARG(s != nullptr, "unknown option");
    *s = value;
    return BB_OK;
}

extern "C" int bb_get_option(bb_ctx* c, const char* name, int64_t* value) {
    BB_ARG(c && name && value, "ctx/name/value");
    if (!strcmp(name, "sm_count")) { *value = c->sm_count; return BB_OK; }
    if (!strcmp(name, "nranks")) { *value = c->nranks; return BB_OK; }
    if (!strcmp(name, "rank")) { *value = c->rank; return BB_OK; }
    i64* s = option_slot(c, name);
    BB_ARG(s != nullptr, "unknown option");
    *value = *s;
    return BB_OK;
}

extern "C" int bb_get_launch_count(bb_ctx* c, int64_t* launches) {
    BB_ARG(c && launches, "ctx/launches");
    *launches = c->launches;
    return BB_OK;
}

extern "C" int bb_reset_launch_count(bb_ctx* c) {
    BB_ARG(c != nullptr, "ctx");
    c->launches = 0;
    return BB_OK;
}

extern "C" int bb_get_device_ms(bb_ctx* c, double* ms) {
    BB_ARG(c && ms, "ctx/ms");
    *ms = c->dev_ms;
    return BB_OK;
}

extern "C" int bb_reset_device_ms(bb_ctx* c) {
    BB_ARG(c != nullptr, "ctx");
    c->dev_ms = 0.0;
    return BB_OK;
}

int bb_ctx_pinned(bb_ctx* c, size_t bytes, double** out) {
    if (bytes > c->pinned_bytes) {
        if (c->pinned) cudaFreeHost(c->pinned);
        c->pinned = nullptr;
        c->pinned_bytes = 0;
        size_t want = bytes + bytes / 4 + 4096;
        BB_CUDA(cudaMallocHost((void**)&c->pinned, want));
        c->pinned_bytes = want;
    }
    *out = c->pinned;
    return BB_OK;
}

int bb_ctx_scratch(bb_ctx* c, int slot, size_t bytes, void** out) {
    if (bytes > c->scratch_bytes[slot]) {
        if (c->scratch[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->scratch[slot]); }
        c->scratch[slot] = nullptr;
        c->scratch_bytes[slot] = 0;
        size_t want = bytes + bytes / 4 + 4096;
        BB_CUDA(cudaMalloc(&c->scratch[slot], want));
        c->scratch_bytes[slot] = want;
    }
    *out = c->scratch[slot];
    return BB_OK;
}

// ---- NCCL through dlopen -----------------------------------------------------------------
// Minimal mirror of the NCCL ABI we use (nccl.h: ncclUniqueId is 128 opaque bytes;
// ncclDouble == 8 in ncclDataType_t; ncclSum == 0 in ncclRedOp_t).
struct bb_nccl_id { char internal[128]; };
typedef int (*nccl_get_id_t)(bb_nccl_id*);
typedef int (*nccl_init_rank_t)(void** comm, int nranks, bb_nccl_id id, int rank);
typedef int (*nccl_allreduce_t)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t s);
typedef const char* (*nccl_errstr_t)(int);

static void* open_nccl(const char* path) {
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) bb_set_error("cannot dlopen NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
    return h;
}

extern "C" int bb_comm_unique_id(const char* nccl_lib_path, char* id_out_128) {
    BB_ARG(id_out_128 != nullptr, "id_out_128");
    void* h = open_nccl(nccl_lib_path);
    if (!h) return BB_ERR_NCCL;
    nccl_get_id_t f = (nccl_get_id_t)dlsym(h, "ncclGetUniqueId");
    if (!f) { bb_set_error("ncclGetUniqueId not found"); return BB_ERR_NCCL; }
    bb_nccl_id id;
    int rc = f(&id);
    if (rc != 0) { bb_set_error("ncclGetUniqueId failed: %d", rc); return BB_ERR_NCCL; }
    memcpy(id_out_128, id.internal, 128);
    return BB_OK;
}

extern "C" int bb_comm_init(bb_ctx* c, const char* nccl_lib_path, int nranks, int rank, const char* id_128) {
    BB_ARG(c && id_128, "ctx/id");
    BB_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "nranks/rank");
    c->nranks = nranks;
    c->rank = rank;
    if (nranks == 1) return BB_OK;
    BB_CUDA(cudaSetDevice(c->device));
    void* h = open_nccl(nccl_lib_path);
    if (!h) return BB_ERR_NCCL;
    c->nccl_handle = h;
    nccl_init_rank_t f = (nccl_init_rank_t)dlsym(h, "ncclCommInitRank");
    if (!f) { bb_set_error("ncclCommInitRank not found"); return BB_ERR_NCCL; }
    bb_nccl_id id;
    memcpy(id.internal, id_128, 128);
    int rc = f(&c->nccl_comm, nranks, id, rank);
    if (rc != 0) {
        nccl_errstr_t es = (nccl_errstr_t)dlsym(h, "ncclGetErrorString");
        bb_set_error("ncclCommInitRank failed: %s", es ? es(rc) : "?");
        c->nccl_comm = nullptr;
        return BB_ERR_NCCL;
    }
    return BB_OK;
}

extern "C" int bb_comm_init_local(bb_ctx* c, int nranks, int rank) {
    BB_ARG(c != nullptr, "ctx");
    BB_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "nranks/rank");
    c->nranks = nranks;
    c->rank = rank;
    c->comm_local = 1;
    return BB_OK;
}

int bb_allreduce_dev(bb_ctx* c, double* dbuf, i64 count) {
    if (c->nranks == 1) return BB_OK;
    {
        int rc = BB_OK;
        if (bb_p2p_allreduce(c, dbuf, count, nullptr, &rc)) return rc;
    }
    if (c->comm_local) {
        // no NCCL behind this communicator: reduce in chunks of the exchange capacity
        const i64 cap = bb_p2p_capacity(c);
        if (cap <= 0) { bb_set_error("local communicator: attach the peer-memory exchange first (bb_comm_p2p_export/attach)"); return BB_ERR_STATE; }
        for (i64 off = 0; off < count; off += cap) {
            int rc = BB_OK;
            if (!bb_p2p_allreduce(c, dbuf + off, (count - off < cap) ? count - off : cap, nullptr, &rc)) {
                bb_set_error("local communicator: peer-memory exchange unavailable"); return BB_ERR_STATE;
            }
            if (rc != BB_OK) return rc;
        }
        return BB_OK;
    }
    if (!c->nccl_comm) { bb_set_error("communicator not initialised"); return BB_ERR_STATE; }
    static nccl_allreduce_t f = nullptr;
    if (!f) f = (nccl_allreduce_t)dlsym(c->nccl_handle, "ncclAllReduce");
    if (!f) { bb_set_error("ncclAllReduce not found"); return BB_ERR_NCCL; }
    int rc = f(dbuf, dbuf, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream);
    if (rc != 0) { bb_set_error("ncclAllReduce failed: %d", rc); return BB_ERR_NCCL; }
    c->launches++;
    return BB_OK;
}

extern "C" int bb_comm_allreduce_host(bb_ctx* c, double* buf, int64_t count) {
    BB_ARG(c && buf && count >= 0, "ctx/buf/count");
    if (c->nranks == 1 || count == 0) return BB_OK;
    BB_CUDA(cudaSetDevice(c->device));
    double* d = nullptr;
    BB_CUDA(cudaMalloc((void**)&d, (size_t)count * sizeof(double)));
    BB_CUDA(cudaMemcpyAsync(d, buf, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int rc = bb_allreduce_dev(c, d, count);
    if (rc == BB_OK) {
        cudaError_t e = cudaMemcpyAsync(buf, d, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { bb_set_error("allreduce_host copy: %s", cudaGetErrorString(e)); rc = BB_ERR_CUDA; }
    }
    cudaFree(d);
    return rc;
}
