// Micro-benchmark: how fast can 148 persistent CTAs x 32 warps stream an index array from HBM, depending on HOW the
// bytes of a warp are laid out?  (Decides the storage order of the sliced SpMV format, bb_sell.cu.)
//   mode 0: every warp owns one contiguous strip, rows of 256 B (8 B per lane)      -- the layout of k_sell_spmv today
//   mode 1: the 32 warps of a CTA interleave their rows: row k of warp w sits at (k*32 + w) * 256 B of the CTA's section
//   mode 2: like 0 with rows of 512 B (16 B per lane)
//   mode 3: like 1 with rows of 512 B
//   mode 4: plain grid-stride read with many small CTAs (the shape of a copy kernel), 16 B per thread
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o experimental/_build/stream_bench experimental/stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint2 ldg8(const uint2* p) {
    uint2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v;
}
__device__ __forceinline__ uint4 ldg16(const uint4* p) {
    uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}

template <int R, bool INTERLEAVE>
__global__ void __launch_bounds__(1024, 1) k_strip8(const uint2* __restrict__ data, long long rows_per_warp, unsigned* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long cta_base = (long long)blockIdx.x * 32 * rows_per_warp;      // in rows
    const uint2* p;
    long long stride;
    if (INTERLEAVE) { p = data + (cta_base + warp) * 32 + lane; stride = 32 * 32; }
    else { p = data + (cta_base + (long long)warp * rows_per_warp) * 32 + lane; stride = 32; }
    uint2 a[R], b[R];
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) a[k] = ldg8(p + k * stride);
    for (long long r = 0; r < rows_per_warp; r += 2 * R) {
#pragma unroll
        for (int k = 0; k < R; ++k) if (r + R + k < rows_per_warp) b[k] = ldg8(p + (r + R + k) * stride);
#pragma unroll
        for (int k = 0; k < R; ++k) acc += a[k].x ^ a[k].y;
#pragma unroll
        for (int k = 0; k < R; ++k) if (r + 2 * R + k < rows_per_warp) a[k] = ldg8(p + (r + 2 * R + k) * stride);
#pragma unroll
        for (int k = 0; k < R; ++k) if (r + R + k < rows_per_warp) acc += b[k].x ^ b[k].y;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int R, bool INTERLEAVE>
__global__ void __launch_bounds__(1024, 1) k_strip16(const uint4* __restrict__ data, long long rows_per_warp, unsigned* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long cta_base = (long long)blockIdx.x * 32 * rows_per_warp;
    const uint4* p;
    long long stride;
    if (INTERLEAVE) { p = data + (cta_base + warp) * 32 + lane; stride = 32 * 32; }
    else { p = data + (cta_base + (long long)warp * rows_per_warp) * 32 + lane; stride = 32; }
    uint4 a[R], b[R];
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) a[k] = ldg16(p + k * stride);
    for (long long r = 0; r < rows_per_warp; r += 2 * R) {
#pragma unroll
        for (int k = 0; k < R; ++k) if (r + R + k < rows_per_warp) b[k] = ldg16(p + (r + R + k) * stride);
#pragma unroll
        for (int k = 0; k < R; ++k) acc += a[k].x ^ a[k].y ^ a[k].z ^ a[k].w;
#pragma unroll
        for (int k = 0; k < R; ++k) if (r + 2 * R + k < rows_per_warp) a[k] = ldg16(p + (r + 2 * R + k) * stride);
#pragma unroll
        for (int k = 0; k < R; ++k) if (r + R + k < rows_per_warp) acc += b[k].x ^ b[k].y ^ b[k].z ^ b[k].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void k_gridstride(const uint4* __restrict__ data, long long n16, unsigned* __restrict__ out) {
    unsigned acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        uint4 v = ldg16(data + i); acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void k_fill(unsigned* p, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (unsigned)(i * 2654435761u);
}

int main(int argc, char** argv) {
    const long long bytes_target = (argc > 1 ? atoll(argv[1]) : 256) << 20;     // MiB streamed per launch
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long long warps = (long long)sms * 32;
    long long rows8 = bytes_target / (warps * 256); rows8 -= rows8 % 32;
    long long rows16 = bytes_target / (warps * 512); rows16 -= rows16 % 32;
    const long long bytes = warps * rows8 * 256;
    unsigned *buf, *out, *flush;
    CK(cudaMalloc(&buf, bytes + 4096)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&flush, 512ll << 20));
    k_fill<<<1024, 256>>>(buf, bytes / 4); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto launch, long long nbytes) {
        float best = 1e9f, tot = 0.f; const int reps = 10;
        for (int r = -2; r < reps; ++r) {
            k_fill<<<1024, 256>>>(flush, (512ll << 20) / 4);          // evict the 126 MB L2
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 0) { tot += ms; if (ms < best) best = ms; }
        }
        printf("%-44s %7.1f us avg %7.1f us best  %6.0f GB/s (avg)\n", name, 1e3 * tot / reps, 1e3 * best, nbytes / (tot / reps) / 1e6);
    };
    printf("SMs %d, %.1f MB per launch\n", sms, bytes / 1e6);
    run("0: warp strips, 8 B/lane, ring 8+8", [&] { k_strip8<8, false><<<sms, 1024>>>((const uint2*)buf, rows8, out); }, bytes);
    run("0: warp strips, 8 B/lane, ring 16+16", [&] { k_strip8<16, false><<<sms, 1024>>>((const uint2*)buf, rows8, out); }, bytes);
    run("1: CTA-interleaved rows, 8 B/lane, ring 8+8", [&] { k_strip8<8, true><<<sms, 1024>>>((const uint2*)buf, rows8, out); }, bytes);
    run("1: CTA-interleaved rows, 8 B/lane, ring 16+16", [&] { k_strip8<16, true><<<sms, 1024>>>((const uint2*)buf, rows8, out); }, bytes);
    run("2: warp strips, 16 B/lane, ring 4+4", [&] { k_strip16<4, false><<<sms, 1024>>>((const uint4*)buf, rows16, out); }, warps * rows16 * 512);
    run("2: warp strips, 16 B/lane, ring 8+8", [&] { k_strip16<8, false><<<sms, 1024>>>((const uint4*)buf, rows16, out); }, warps * rows16 * 512);
    run("3: CTA-interleaved rows, 16 B/lane, ring 4+4", [&] { k_strip16<4, true><<<sms, 1024>>>((const uint4*)buf, rows16, out); }, warps * rows16 * 512);
    run("3: CTA-interleaved rows, 16 B/lane, ring 8+8", [&] { k_strip16<8, true><<<sms, 1024>>>((const uint4*)buf, rows16, out); }, warps * rows16 * 512);
    run("4: grid-stride, 16 B/thread, 148*8 CTAs x 256", [&] { k_gridstride<<<sms * 8, 256>>>((const uint4*)buf, bytes / 16, out); }, bytes);
    run("4: grid-stride, 16 B/thread, 148*16 CTAs x 512", [&] { k_gridstride<<<sms * 16, 512>>>((const uint4*)buf, bytes / 16, out); }, bytes);
    return 0;
}
