// peer_variants.cu — measurement aid (not part of the product): which SM-issued access shape carries a slot
// payload over NVLink fastest when BOTH directions of a GPU pair are busy (the ring bench's situation)?
//
// Round 1 (profiles/r01_push_pull.jsonl) left the SM paths at 640–671 (pull) / 703–708 (push) GB/s per direction
// against 772 for the copy engine.  The model that fits those numbers: one wire carries ≈ 880 GB/s raw; a pull
// costs 1.125 wire bytes per payload byte on the response wire plus 0.19 on the request wire (128-byte requests),
// a push costs ≈ 1.23 on the write wire (128-byte write packets).  If that is right, only LARGER packets help:
// this tool tries 256-bit LDG/STG (sm_100's LDG.E.256 / STG.E.256), deeper / larger TMA bulk stages, cache
// operators, grid shapes — each as push (kernel on the source GPU) and pull (kernel on the sink GPU), one and two
// directions busy — and, under ncu (NCU=1: one launch per variant, one direction), nvltx/nvlrx bytes per payload byte.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Iinclude -o tools/_build/peer_variants tools/peer_variants.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../lzy_b200/csrc/tma_kernels.cuh"

using namespace lzyb;

#define CK(x)                                                                  \
    do {                                                                       \
        cudaError_t e = (x);                                                   \
        if (e != cudaSuccess) {                                                \
            printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); \
            fflush(stdout);                                                    \
            exit(1);                                                           \
        }                                                                      \
    } while (0)

static unsigned long long N = 1ull << 30;
static int REPS = 10;

// ---- 256-bit access helpers (PTX 8.8, sm_100+) ------------------------------------------------
struct u8x32 {
    uint4 lo, hi;
};
__device__ __forceinline__ u8x32 ld256(const uint8_t* p) {
    u8x32 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(uint8_t* p, const u8x32& v) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.lo.x), "r"(v.lo.y),
                 "r"(v.lo.z), "r"(v.lo.w), "r"(v.hi.x), "r"(v.hi.y), "r"(v.hi.z), "r"(v.hi.w)
                 : "memory");
}
__device__ __forceinline__ void st128_cs(uint4* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st128_wt(uint4* p, const uint4& v) {
    asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit copy + digest; tile = 256 threads × UNROLL × 32 B
template <int UNROLL, bool FOLD>
__global__ void __launch_bounds__(256) copy256_kernel(const __grid_constant__ CopyParams a) {
    const unsigned long long nv = a.nbytes >> 5;  // 32-byte vectors
    const unsigned long long tile = 256ull * UNROLL;
    const unsigned long long ntiles = nv / tile;
    unsigned long long sum = 0ull, xr = 0ull;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long base = t * tile + threadIdx.x;
        u8x32 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) v[j] = ld256(a.src + ((base + 256ull * j) << 5));
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const unsigned long long i = base + 256ull * j;  // 32-byte vector index = 4 words
            if constexpr (FOLD) {
                const unsigned long long key = (a.word_base + 4ull * i + 1ull) * LZYB_K_POS;
                fold16(v[j].lo, key, sum, xr);
                fold16(v[j].hi, key + 2ull * LZYB_K_POS, sum, xr);
            }
            st256(a.dst + (i << 5), v[j]);
        }
    }
    block_epilogue(a, sum, xr);
}

// 128-bit copy with a chosen store operator (0 = no_allocate (product), 1 = .cs, 2 = .wt)
template <int STOP>
__global__ void __launch_bounds__(256) copy128_kernel(const __grid_constant__ CopyParams a) {
    constexpr int UNROLL = 8;
    const unsigned long long nvec = a.nbytes >> 4;
    const unsigned long long tile = 256ull * UNROLL;
    const unsigned long long ntiles = nvec / tile;
    unsigned long long sum = 0ull, xr = 0ull;
    uint4* dstv = reinterpret_cast<uint4*>(a.dst);
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long base = t * tile + threadIdx.x;
        uint4 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) v[j] = ld_stream(reinterpret_cast<const uint4*>(a.src) + base + 256ull * j);
        unsigned long long key = (a.word_base + 2ull * base + 1ull) * LZYB_K_POS;
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            fold16(v[j], key, sum, xr);
            key += 2ull * 256 * LZYB_K_POS;
            if constexpr (STOP == 0) st_stream(dstv + base + 256ull * j, v[j]);
            else if constexpr (STOP == 1) st128_cs(dstv + base + 256ull * j, v[j]);
            else st128_wt(dstv + base + 256ull * j, v[j]);
        }
    }
    block_epilogue(a, sum, xr);
}

// compare sink with source (runs on the sink GPU, reads the source over NVLink)
__global__ void compare_kernel(const uint4* a, const uint4* b, unsigned long long nvec, unsigned long long* bad) {
    unsigned long long local = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nvec;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        uint4 x = a[i], y = b[i];
        if (x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w) ++local;
    }
    if (local) atomicAdd(bad, local);
}

struct Side {
    int dev;
    uint8_t *a, *b;  // a: source payload, b: sink
    SlotCtl* ctl;
    unsigned long long* bad;
    cudaStream_t st;
    cudaEvent_t e0, e1;
};

struct Variant {
    const char* name;
    int id;
};

enum {
    V_LDG128_2T = 0,   // product LDG kernel, 2 tiles/CTA
    V_LDG128_CS,       // .cs stores
    V_LDG128_WT,       // .wt stores
    V_256_U4_1T,       // 256-bit, 32 KiB tile per CTA, one tile per CTA
    V_256_U8_1T,       // 256-bit, 64 KiB tile per CTA
    V_256_U4_P4,       // 256-bit persistent 4 CTAs/SM
    V_256_U4_P8,       // 256-bit persistent 8 CTAs/SM
    V_256_NOFOLD,      // 256-bit without digest arithmetic
    V_TMA_3x32_2,      // product ring
    V_TMA_3x64_1,      // 64 KiB stages, 1 CTA/SM
    V_TMA_6x32_1,
    V_TMA_2x96_1,      // 96 KiB stages ×2
    V_TMA_8x8_3,
    V_TMA_12x16_1,
    V_CE,
    V_COUNT
};
static const char* kNames[V_COUNT] = {
    "ldg128 2 tiles/CTA (product)", "ldg128 st.cs", "ldg128 st.wt", "ldst256 32K tile/CTA", "ldst256 64K tile/CTA",
    "ldst256 persistent 4/SM", "ldst256 persistent 8/SM", "ldst256 no digest", "tma 3x32K 2/SM (product)", "tma 3x64K 1/SM",
    "tma 6x32K 1/SM", "tma 2x96K 1/SM", "tma 8x8K 3/SM", "tma 12x16K 1/SM", "copy engine"};

template <int STAGES, int SB>
static void tma_go(const CopyParams& p, int grid, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * SB + 2 * STAGES * sizeof(uint64_t);
    static bool set[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!set[dev]) {
        CK(cudaFuncSetAttribute(tma_stream_digest_kernel<STAGES, SB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set[dev] = true;
    }
    tma_stream_digest_kernel<STAGES, SB, true><<<grid, TMA_THREADS, smem, st>>>(p);
}

static void launch(int v, const uint8_t* src, uint8_t* dst, SlotCtl* ctl, cudaStream_t st, int sms) {
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = src; p.dst = dst; p.nbytes = N; p.ctl = ctl; p.total_nbytes = N; p.fin = FIN_NONE;
    const unsigned long long t128 = (N >> 4) / (256ull * 8);
    switch (v) {
        case V_LDG128_2T: stream_digest_kernel<16, 8, true><<<(unsigned)(t128 / 2), 256, 0, st>>>(p); break;
        case V_LDG128_CS: copy128_kernel<1><<<(unsigned)(t128 / 2), 256, 0, st>>>(p); break;
        case V_LDG128_WT: copy128_kernel<2><<<(unsigned)(t128 / 2), 256, 0, st>>>(p); break;
        case V_256_U4_1T: copy256_kernel<4, true><<<(unsigned)((N >> 5) / (256ull * 4)), 256, 0, st>>>(p); break;
        case V_256_U8_1T: copy256_kernel<8, true><<<(unsigned)((N >> 5) / (256ull * 8)), 256, 0, st>>>(p); break;
        case V_256_U4_P4: copy256_kernel<4, true><<<sms * 4, 256, 0, st>>>(p); break;
        case V_256_U4_P8: copy256_kernel<4, true><<<sms * 8, 256, 0, st>>>(p); break;
        case V_256_NOFOLD: copy256_kernel<4, false><<<(unsigned)((N >> 5) / (256ull * 4)), 256, 0, st>>>(p); break;
        case V_TMA_3x32_2: tma_go<3, 32768>(p, sms * 2, st); break;
        case V_TMA_3x64_1: tma_go<3, 65536>(p, sms, st); break;
        case V_TMA_6x32_1: tma_go<6, 32768>(p, sms, st); break;
        case V_TMA_2x96_1: tma_go<2, 98304>(p, sms, st); break;
        case V_TMA_8x8_3: tma_go<8, 8192>(p, sms * 3, st); break;
        case V_TMA_12x16_1: tma_go<12, 16384>(p, sms, st); break;
        default: cudaMemcpyAsync(dst, src, N, cudaMemcpyDefault, st); break;
    }
}

int main(int argc, char** argv) {
    const bool ncu = getenv("NCU") != nullptr;
    if (getenv("MIB")) N = (unsigned long long)atoll(getenv("MIB")) << 20;
    if (ncu) REPS = 1;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("{\"skip\": \"needs 2 GPUs\"}\n"); return 0; }
    Side s[2];
    int sms = 0;
    for (int i = 0; i < 2; ++i) {
        s[i].dev = i;
        CK(cudaSetDevice(i));
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, i, 1 - i));
        if (!can) { printf("{\"skip\": \"no peer access\"}\n"); return 0; }
        CK(cudaDeviceEnablePeerAccess(1 - i, 0));
        CK(cudaMalloc(&s[i].a, N)); CK(cudaMalloc(&s[i].b, N)); CK(cudaMalloc(&s[i].ctl, sizeof(SlotCtl)));
        CK(cudaMalloc(&s[i].bad, 8));
        CK(cudaMemset(s[i].b, 0, N)); CK(cudaMemset(s[i].ctl, 0, sizeof(SlotCtl))); CK(cudaMemset(s[i].bad, 0, 8));
        fill_pattern_kernel<<<1184, 256>>>(s[i].a, N, 11 + i);
        CK(cudaStreamCreateWithFlags(&s[i].st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&s[i].e0)); CK(cudaEventCreate(&s[i].e1));
        CK(cudaDeviceSynchronize());
    }
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    for (int v = 0; v < V_COUNT; ++v)
        for (int push = 0; push < 2; ++push)
            for (int dirs = 1; dirs <= 2; ++dirs) {
                if (ncu && dirs == 2) continue;
                if (ncu && v == V_CE) continue;
                // copy engine: "push" = issued on the source device, "pull" = on the sink device
                float ms[2] = {0, 0};
                for (int k = 0; k < dirs; ++k) { CK(cudaSetDevice(1 - k)); CK(cudaMemsetAsync(s[1 - k].b, 0, N, s[1 - k].st)); CK(cudaStreamSynchronize(s[1 - k].st)); }
                for (int pass = ncu ? 1 : 0; pass < 2; ++pass) {  // pass 0 = warm-up
                    for (int k = 0; k < dirs; ++k) {
                        const int run = push ? k : 1 - k;  // transfer k: source on GPU k, sink on GPU 1-k
                        CK(cudaSetDevice(run));
                        CK(cudaEventRecord(s[run].e0, s[run].st));
                    }
                    for (int r = 0; r < (pass ? REPS : 2); ++r)
                        for (int k = 0; k < dirs; ++k) {
                            const int run = push ? k : 1 - k;
                            CK(cudaSetDevice(run));
                            launch(v, s[k].a, s[1 - k].b, s[run].ctl, s[run].st, sms);
                        }
                    for (int k = 0; k < dirs; ++k) {
                        const int run = push ? k : 1 - k;
                        CK(cudaSetDevice(run));
                        CK(cudaEventRecord(s[run].e1, s[run].st));
                    }
                    for (int k = 0; k < dirs; ++k) {
                        const int run = push ? k : 1 - k;
                        CK(cudaSetDevice(run));
                        CK(cudaEventSynchronize(s[run].e1));
                        CK(cudaEventElapsedTime(&ms[k], s[run].e0, s[run].e1));
                    }
                    CK(cudaGetLastError());
                }
                // bit-exact check of what landed
                unsigned long long bad_total = 0;
                for (int k = 0; k < dirs; ++k) {
                    CK(cudaSetDevice(1 - k));
                    CK(cudaMemsetAsync(s[1 - k].bad, 0, 8, s[1 - k].st));
                    compare_kernel<<<1184, 256, 0, s[1 - k].st>>>((const uint4*)s[k].a, (const uint4*)s[1 - k].b, N >> 4, s[1 - k].bad);
                    unsigned long long hb = 0;
                    CK(cudaMemcpyAsync(&hb, s[1 - k].bad, 8, cudaMemcpyDeviceToHost, s[1 - k].st));
                    CK(cudaStreamSynchronize(s[1 - k].st));
                    bad_total += hb;
                }
                const float worst = dirs == 2 ? std::max(ms[0], ms[1]) : ms[0];
                printf("{\"variant\": \"%s\", \"mode\": \"%s\", \"directions\": %d, \"ms_per_payload\": %.4f, \"GBps_per_dir\": %.1f, "
                       "\"payload_MiB\": %llu, \"mismatched_vectors\": %llu}\n",
                       kNames[v], push ? "push" : "pull", dirs, worst / REPS, (double)N * REPS / worst / 1e6, N >> 20, bad_total);
                fflush(stdout);
            }
    return 0;
}
