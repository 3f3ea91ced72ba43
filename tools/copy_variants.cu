// copy_variants.cu — tuning aid (not part of the product): which access pattern reaches the chip's copy
// ceiling?  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/copy_variants tools/copy_variants.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum Cache { DEF = 0, NC_NA = 1, CS = 2 };

template <int C> __device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 r;
    if constexpr (C == DEF) r = *p;
    else if constexpr (C == NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int C> __device__ __forceinline__ void st(uint4* p, uint4 v) {
    if constexpr (C == DEF) *p = v;
    else if constexpr (C == NC_NA) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// one tile per CTA (massively oversubscribed grid, like an elementwise kernel)
template <int U, int C, bool STORE>
__global__ void tile_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, unsigned long long nvec, unsigned long long* sink) {
    unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x * U + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) if (base + (unsigned long long)j * blockDim.x < nvec) v[j] = ld<C>(src + base + (unsigned long long)j * blockDim.x);
    unsigned acc = 0;
#pragma unroll
    for (int j = 0; j < U; ++j) if (base + (unsigned long long)j * blockDim.x < nvec) {
        if constexpr (STORE) st<C>(dst + base + (unsigned long long)j * blockDim.x, v[j]); else acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    if (!STORE && acc == 0x12345678u) atomicAdd(sink, 1ull);
}

// persistent: grid-stride over tiles
template <int U, int C, bool STORE>
__global__ void persist_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, unsigned long long nvec, unsigned long long* sink) {
    const unsigned long long tile = (unsigned long long)blockDim.x * U, ntiles = nvec / tile;
    unsigned acc = 0;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        unsigned long long base = t * tile + threadIdx.x;
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = ld<C>(src + base + (unsigned long long)j * blockDim.x);
#pragma unroll
        for (int j = 0; j < U; ++j) { if constexpr (STORE) st<C>(dst + base + (unsigned long long)j * blockDim.x, v[j]); else acc ^= v[j].x ^ v[j].w; }
    }
    if (!STORE && acc == 0x12345678u) atomicAdd(sink, 1ull);
}

// persistent, each CTA owns one contiguous span (no interleaving between CTAs)
template <int U, int C, bool STORE>
__global__ void span_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, unsigned long long nvec, unsigned long long* sink) {
    const unsigned long long tile = (unsigned long long)blockDim.x * U, ntiles = nvec / tile;
    const unsigned long long per = (ntiles + gridDim.x - 1) / gridDim.x, t0 = blockIdx.x * per, t1 = min(ntiles, t0 + per);
    unsigned acc = 0;
    for (unsigned long long t = t0; t < t1; ++t) {
        unsigned long long base = t * tile + threadIdx.x;
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = ld<C>(src + base + (unsigned long long)j * blockDim.x);
#pragma unroll
        for (int j = 0; j < U; ++j) { if constexpr (STORE) st<C>(dst + base + (unsigned long long)j * blockDim.x, v[j]); else acc ^= v[j].x ^ v[j].w; }
    }
    if (!STORE && acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <typename F> float timeit(F f, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
    const unsigned long long n = 1ull << 30, nvec = n / 16;
    uint4 *src, *dst; unsigned long long* sink;
    CK(cudaMalloc(&src, n)); CK(cudaMalloc(&dst, n)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(src, 1, n)); CK(cudaMemset(dst, 0, n));
    int sms = 148;
    auto rep = [&](const char* name, float ms, bool store) { printf("{\"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f}\n", name, ms, (store ? 2.0 : 1.0) * n / ms / 1e6); };
    rep("cudaMemcpyAsync D2D", timeit([&] { cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice); }), true);
#define TILE(U, B, C, S, NAME) rep(NAME, timeit([&] { tile_kernel<U, C, S><<<(unsigned)((nvec + (unsigned long long)B * U - 1) / ((unsigned long long)B * U)), B>>>(src, dst, nvec, sink); }), S)
#define PERS(U, B, C, S, G, NAME) rep(NAME, timeit([&] { persist_kernel<U, C, S><<<sms * G, B>>>(src, dst, nvec, sink); }), S)
#define SPAN(U, B, C, S, G, NAME) rep(NAME, timeit([&] { span_kernel<U, C, S><<<sms * G, B>>>(src, dst, nvec, sink); }), S)
    TILE(4, 128, DEF, true, "copy tile U4 B128 default");
    TILE(4, 128, NC_NA, true, "copy tile U4 B128 nc/no_allocate");
    TILE(4, 128, CS, true, "copy tile U4 B128 .cs");
    TILE(8, 256, DEF, true, "copy tile U8 B256 default");
    TILE(8, 256, NC_NA, true, "copy tile U8 B256 nc/no_allocate");
    TILE(4, 256, DEF, true, "copy tile U4 B256 default");
    TILE(2, 256, DEF, true, "copy tile U2 B256 default");
    TILE(16, 256, DEF, true, "copy tile U16 B256 default");
    PERS(8, 256, NC_NA, true, 4, "copy persist U8 B256 G4 nc (current LDG engine)");
    PERS(8, 256, DEF, true, 4, "copy persist U8 B256 G4 default");
    PERS(4, 256, DEF, true, 8, "copy persist U4 B256 G8 default");
    PERS(8, 256, DEF, true, 8, "copy persist U8 B256 G8 default");
    PERS(4, 128, DEF, true, 16, "copy persist U4 B128 G16 default");
    SPAN(8, 256, DEF, true, 4, "copy span U8 B256 G4 default");
    SPAN(8, 256, DEF, true, 8, "copy span U8 B256 G8 default");
    TILE(4, 128, DEF, false, "read tile U4 B128 default");
    TILE(8, 256, DEF, false, "read tile U8 B256 default");
    TILE(8, 256, NC_NA, false, "read tile U8 B256 nc");
    PERS(8, 256, NC_NA, false, 4, "read persist U8 B256 G4 nc (current scan)");
    PERS(8, 256, DEF, false, 8, "read persist U8 B256 G8 default");
    SPAN(8, 256, DEF, false, 8, "read span U8 B256 G8 default");
    rep("cudaMemsetAsync (write only)", timeit([&] { cudaMemsetAsync(dst, 0, n); }), false);
    return 0;
}
