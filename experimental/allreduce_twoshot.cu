// PROTOTYPE + micro-benchmark: deterministic TWO-SHOT all-reduce of a (p+1)-vector of doubles over NVLink peer memory.
//
// STATUS: compiles for sm_100a; HAS NOT RUN ON A GPU YET.  Not built into libbbgpu.so.  Single process, N GPUs with
// peer access (the library itself is one process per GPU and maps the same buffers through CUDA IPC, bb_p2p.cu; the
// device code below does not care which of the two produced the peer pointers).
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/ar2 experimental/allreduce_twoshot.cu
//     /tmp/ar2 [count = 100001] [iterations = 200]          (gpurun --gpus 2 ... 8)
//
// Why: per CG iteration every rank holds its local [sum w ; X_g' w] (p+1 doubles, 800 KB at p = 100k) and needs the
// sum over ranks, bit-identical everywhere.  The library's ONE-SHOT exchange (every rank reads all N slots: N x the
// NVLink traffic of a reduce-scatter + all-gather) measured 54 us at N = 8 against 29 us for ncclAllReduce
// (DESIGN.md section 6).  Two-shot: rank r sums chunk r of all N slots in rank order (reads 1/N of every peer) and
// stores the result into every rank's result buffer; a consumer then reads only local memory.  Per rank and call
// that is (N-1)/N of the vector in and out instead of (N-1) vectors in.  The sum of every element is formed by one
// rank in the fixed order 0..N-1, so the result is bit-identical on all ranks and identical to the one-shot form.
//
// Protocol (monotone sequence numbers, no resets; slots and result buffers double-buffered by sequence parity):
//   publish(k)  rank q: part[k&1] <- local vector; fence.sys; flag1[q] := k+1 on every peer
//   reduce(k)   rank r: wait flag1[*] >= k+1; for i in chunk r: s = sum_q part_q[k&1][i] (rank order);
//               res_t[k&1][i] := s for every rank t; last block: fence.sys; flag2[r] := k+1 on every peer
//   consume(k)  rank t: wait flag2[*] >= k+1; read res[k&1] (local)
// Reuse is safe with two buffers: a rank publishes k+2 only after consume(k+1), which needs every peer's reduce(k+1),
// which follows that peer's reduce(k) in stream order -- so nobody still reads part[k&1] of sequence k; likewise a
// peer's reduce(k+2) needs this rank's publish(k+2), issued after this rank's consume(k+1) > consume(k).
// In the library publish() is the tail of k_tdot_collect and consume() the head of k_cg_q (as for the one-shot form).
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef long long i64;
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)
constexpr int MAXR = 8;

struct Rank {                 // one per device, passed by value
    double* part[MAXR];       // part[q]: publication slots of rank q     [2][cap]
    double* res[MAXR];        // res[t] : result buffers of rank t         [2][cap]
    u64* flag1[MAXR];         // flag1[t][q]: on rank t, "rank q published"  (u64[MAXR])
    u64* flag2[MAXR];         // flag2[t][r]: on rank t, "rank r's chunk has landed"
    unsigned* counter;        // local: blocks done (publish, reduce)
    unsigned* error;          // local: spin cap hit
    i64 cap;
    int nranks, rank;
};

__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(u64* p, u64 v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// thread q of the block polls flag[q]; returns false (and raises the error flag) when the spin cap is hit
__device__ __forceinline__ bool wait_all(const u64* flags, int nranks, u64 want, unsigned* error) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < nranks) {
        u64 spins = 0;
        while (ld_acquire_sys(flags + threadIdx.x) < want) {
            if (++spins > (1ull << 24)) { ok = 0; *error = 1u; break; }
            __nanosleep(20);
        }
    }
    __syncthreads();
    return ok != 0;
}

// the last block of the grid to arrive runs `signal` (after a system fence): flag[me] := value on every peer
__device__ __forceinline__ void last_block_signals(unsigned* counter, u64* const* flag_of_rank, int nranks, int me, u64 value) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(counter, 1u);
        is_last = prev == gridDim.x - 1;
        if (is_last) { *counter = 0u; __threadfence(); }
    }
    __syncthreads();
    if (is_last && (int)threadIdx.x < nranks) {
        __threadfence_system();
        st_relaxed_sys(flag_of_rank[threadIdx.x] + me, value);
    }
}

__global__ void k_publish(Rank R, const double* __restrict__ src, i64 count, u64 seq) {
    double* slot = R.part[R.rank] + (seq & 1ull) * R.cap;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) slot[i] = src[i];
    last_block_signals(R.counter, R.flag1, R.nranks, R.rank, seq + 1ull);
}

// ---- two-shot ------------------------------------------------------------------------------------------------
__global__ void k_reduce_scatter_bcast(Rank R, i64 count, u64 seq) {
    if (!wait_all(R.flag1[R.rank], R.nranks, seq + 1ull, R.error)) return;
    const i64 chunk = (count + R.nranks - 1) / R.nranks;
    const i64 lo = chunk * R.rank, hi = min(count, lo + chunk);
    const i64 off = (seq & 1ull) * R.cap;
    for (i64 i = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (i64)gridDim.x * blockDim.x) {
        double x[MAXR];
#pragma unroll
        for (int q = 0; q < MAXR; ++q) x[q] = q < R.nranks ? __ldcv(R.part[q] + off + i) : 0.0;   // loads in flight together
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < MAXR; ++q) if (q < R.nranks) s += x[q];                                  // rank order
#pragma unroll
        for (int t = 0; t < MAXR; ++t) if (t < R.nranks) R.res[t][off + i] = s;
    }
    last_block_signals(R.counter + 1, R.flag2, R.nranks, R.rank, seq + 1ull);
}

__global__ void k_consume(Rank R, double* __restrict__ dst, i64 count, u64 seq) {
    if (!wait_all(R.flag2[R.rank], R.nranks, seq + 1ull, R.error)) return;
    const double* res = R.res[R.rank] + (seq & 1ull) * R.cap;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) dst[i] = __ldcv(res + i);
}

// ---- one-shot (the protocol bb_p2p.cu implements today), for the A/B ---------------------------------------------
__global__ void k_oneshot_reduce(Rank R, double* __restrict__ dst, i64 count, u64 seq) {
    if (!wait_all(R.flag1[R.rank], R.nranks, seq + 1ull, R.error)) return;
    const i64 off = (seq & 1ull) * R.cap;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) {
        double x[MAXR];
#pragma unroll
        for (int q = 0; q < MAXR; ++q) x[q] = q < R.nranks ? __ldcv(R.part[q] + off + i) : 0.0;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < MAXR; ++q) if (q < R.nranks) s += x[q];
        dst[i] = s;
    }
}
// one-shot needs a second handshake before a slot is reused after TWO publications; with the sequence of this
// benchmark (publish k, reduce k, publish k+1, ...) a rank can publish k+2 while a slow peer still reads slot k&1 of
// sequence k only if it ran two reductions ahead, which the flag wait of reduce(k+1) on that peer's publish(k+1)
// excludes (the peer publishes k+1 after finishing reduce(k)).  Same argument as in bb_p2p.cu.

int main(int argc, char** argv) {
    const i64 count = argc > 1 ? atoll(argv[1]) : 100001;
    const int iters = argc > 2 ? atoi(argv[2]) : 200;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const int N = std::min(ndev, MAXR);
    if (N < 2) { printf("needs >= 2 GPUs (found %d)\n", ndev); return 0; }
    const i64 cap = (count + 31) & ~(i64)31;
    std::vector<Rank> R((size_t)N);
    std::vector<double*> src((size_t)N), dst((size_t)N);
    std::vector<cudaStream_t> st((size_t)N);
    std::vector<std::vector<double>> host((size_t)N, std::vector<double>((size_t)count));
    for (int d = 0; d < N; ++d) {
        CK(cudaSetDevice(d));
        for (int q = 0; q < N; ++q) if (q != d) {
            int can = 0; CK(cudaDeviceCanAccessPeer(&can, d, q));
            if (!can) { printf("no peer access %d -> %d\n", d, q); return 0; }
            cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            (void)cudaGetLastError();
        }
        CK(cudaStreamCreateWithFlags(&st[(size_t)d], cudaStreamNonBlocking));
    }
    std::vector<double*> part((size_t)N), res((size_t)N); std::vector<u64*> f1((size_t)N), f2((size_t)N);
    for (int d = 0; d < N; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaMalloc((void**)&part[(size_t)d], 2 * cap * sizeof(double)));
        CK(cudaMalloc((void**)&res[(size_t)d], 2 * cap * sizeof(double)));
        CK(cudaMalloc((void**)&f1[(size_t)d], MAXR * sizeof(u64))); CK(cudaMemset(f1[(size_t)d], 0, MAXR * sizeof(u64)));
        CK(cudaMalloc((void**)&f2[(size_t)d], MAXR * sizeof(u64))); CK(cudaMemset(f2[(size_t)d], 0, MAXR * sizeof(u64)));
        CK(cudaMalloc((void**)&src[(size_t)d], count * sizeof(double)));
        CK(cudaMalloc((void**)&dst[(size_t)d], count * sizeof(double)));
        for (i64 i = 0; i < count; ++i) host[(size_t)d][(size_t)i] = std::sin(0.37 * (double)i + d) * std::pow(10.0, (double)((i + d) % 7) - 3.0);
        CK(cudaMemcpy(src[(size_t)d], host[(size_t)d].data(), count * sizeof(double), cudaMemcpyHostToDevice));
    }
    for (int d = 0; d < N; ++d) {
        CK(cudaSetDevice(d));
        Rank& r = R[(size_t)d];
        memset(&r, 0, sizeof(r));
        for (int q = 0; q < N; ++q) { r.part[q] = part[(size_t)q]; r.res[q] = res[(size_t)q]; r.flag1[q] = f1[(size_t)q]; r.flag2[q] = f2[(size_t)q]; }
        CK(cudaMalloc((void**)&r.counter, 2 * sizeof(unsigned))); CK(cudaMemset(r.counter, 0, 2 * sizeof(unsigned)));
        CK(cudaMalloc((void**)&r.error, sizeof(unsigned))); CK(cudaMemset(r.error, 0, sizeof(unsigned)));
        r.cap = cap; r.nranks = N; r.rank = d;
        CK(cudaDeviceSynchronize());
    }
    std::vector<double> want((size_t)count);
    for (i64 i = 0; i < count; ++i) { double s = 0.0; for (int d = 0; d < N; ++d) s += host[(size_t)d][(size_t)i]; want[(size_t)i] = s; }
    const int gpub = (int)std::min<i64>(64, (count + 1023) / 1024), grs = (int)std::max<i64>(1, std::min<i64>(64, (count / N + 255) / 256));   // ~1 element per thread: latency-bound
    u64 seq = 0;
    for (int mode = 0; mode < 2; ++mode) {           // 0: two-shot, 1: one-shot
        std::vector<cudaEvent_t> e0((size_t)N), e1((size_t)N);
        for (int d = 0; d < N; ++d) { CK(cudaSetDevice(d)); CK(cudaEventCreate(&e0[(size_t)d])); CK(cudaEventCreate(&e1[(size_t)d])); CK(cudaMemsetAsync(dst[(size_t)d], 0, count * sizeof(double), st[(size_t)d])); }
        // One CUDA graph per device and phase (sequence numbers are baked into the nodes): the host would otherwise be
        // the bottleneck (3 launches x N devices per call from one thread).  Warm-up graph first, then the timed one.
        auto run_phase = [&](int ncalls, bool timed) {
            std::vector<cudaGraphExec_t> exec((size_t)N);
            for (int d = 0; d < N; ++d) {
                CK(cudaSetDevice(d));
                cudaGraph_t graph;
                CK(cudaStreamBeginCapture(st[(size_t)d], cudaStreamCaptureModeRelaxed));
                for (int it = 0; it < ncalls; ++it) {
                    const u64 sq = seq + (u64)it;
                    k_publish<<<gpub, 256, 0, st[(size_t)d]>>>(R[(size_t)d], src[(size_t)d], count, sq);
                    if (mode == 0) {
                        k_reduce_scatter_bcast<<<grs, 256, 0, st[(size_t)d]>>>(R[(size_t)d], count, sq);
                        k_consume<<<gpub, 256, 0, st[(size_t)d]>>>(R[(size_t)d], dst[(size_t)d], count, sq);
                    } else {
                        k_oneshot_reduce<<<gpub, 256, 0, st[(size_t)d]>>>(R[(size_t)d], dst[(size_t)d], count, sq);
                    }
                }
                CK(cudaStreamEndCapture(st[(size_t)d], &graph));
                CK(cudaGraphInstantiate(&exec[(size_t)d], graph, 0));
                CK(cudaGraphDestroy(graph));
            }
            for (int d = 0; d < N; ++d) {
                CK(cudaSetDevice(d));
                if (timed) CK(cudaEventRecord(e0[(size_t)d], st[(size_t)d]));
                CK(cudaGraphLaunch(exec[(size_t)d], st[(size_t)d]));
                if (timed) CK(cudaEventRecord(e1[(size_t)d], st[(size_t)d]));
            }
            for (int d = 0; d < N; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[(size_t)d])); CK(cudaGraphExecDestroy(exec[(size_t)d])); }
            seq += (u64)ncalls;
        };
        run_phase(5, false);
        run_phase(iters, true);
        double worst = 0.0;
        for (int d = 0; d < N; ++d) {
            CK(cudaSetDevice(d));
            float ms; CK(cudaEventElapsedTime(&ms, e0[(size_t)d], e1[(size_t)d]));
            worst = std::max(worst, (double)ms);
        }
        bool ok = true;
        std::vector<double> got((size_t)count), got0;
        for (int d = 0; d < N; ++d) {
            CK(cudaSetDevice(d));
            unsigned err = 0; CK(cudaMemcpy(&err, R[(size_t)d].error, sizeof(err), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(got.data(), dst[(size_t)d], count * sizeof(double), cudaMemcpyDeviceToHost));
            if (d == 0) got0 = got;
            ok = ok && err == 0 && memcmp(got.data(), want.data(), count * sizeof(double)) == 0 && memcmp(got.data(), got0.data(), count * sizeof(double)) == 0;
        }
        printf("%s all-reduce, %d ranks, %lld doubles: %.1f us per call (max over ranks, %d calls), result %s\n",
               mode == 0 ? "two-shot" : "one-shot", N, count, worst * 1e3 / iters, iters, ok ? "bit-identical on all ranks and equal to the rank-ordered host sum: PASS" : "FAIL");
    }
    return 0;
}
