// push_variants.cu — measurement aid (not part of the product): what does the COMPLETION SIGNAL of a push cost?
//
// The first push_kernel (profiles/r02_push_bench_v1.jsonl) delivered 584 GB/s one way and 471 with both directions
// busy where a bare SM push does 714 / 703: each CTA ended every 256 KiB chunk with __syncthreads + red.release.sys
// (SASS: MEMBAR.ALL.SYS + RED), and a system-scope membar behind stores that sit in a saturated NVLink egress queue
// waits for that queue.  Variants (all 256-bit LDG/STG, persistent grid, whole chunks per CTA, ascending):
//   sig 0  no signal at all                       — the ceiling of this loop structure
//   sig 1  relaxed red after the barrier          — what the counter traffic alone costs (NOT a correct protocol)
//   sig 2  red.release.sys after the barrier      — round-2 v1
//   sig 3  a dedicated signal warp: copy warps never fence; they arrive on a named barrier after a chunk's last store
//          and go on with the next chunk, the signal warp does fence.acq_rel.sys + st.relaxed.sys of the chunk's
//          epoch word, at most two chunks behind
// × chunk size (tiles of 32 KiB per signal) × CTAs/SM, with and without the sink's landed-scan running, one and two
// directions.  Every run is checked bit for bit on the sink.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Iinclude -o tools/_build/push_variants tools/push_variants.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ldg_push_kernels.cuh"
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
static int REPS = 8;

struct VParams {
    const uint8_t* src;
    uint8_t* dst;
    unsigned long long nbytes;
    unsigned int* flags;
    unsigned int chunk_shift;
    unsigned int epoch;
};

__device__ __forceinline__ void copy_tile256(const uint8_t* s, uint8_t* d, int tid) {
    v256 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ld256_stream(s + ((tid + 256 * j) << 5));
#pragma unroll
    for (int j = 0; j < 4; ++j) st256_stream(d + ((tid + 256 * j) << 5), v[j]);
}
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// SIG 0..2: 256 threads, thread 0 signals after __syncthreads
template <int SIG>
__global__ void __launch_bounds__(256) push_v(const __grid_constant__ VParams a) {
    const unsigned long long ntiles = a.nbytes / PUSH_TILE_BYTES;  // tool: payload is a multiple of the tile
    const unsigned long long tpc = 1ull << a.chunk_shift;
    const unsigned long long nchunks = (ntiles + tpc - 1) >> a.chunk_shift;
    for (unsigned long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const unsigned long long t0 = c << a.chunk_shift;
        const unsigned long long t1 = (t0 + tpc < ntiles) ? t0 + tpc : ntiles;
#pragma unroll 1
        for (unsigned long long t = t0; t < t1; ++t) copy_tile256(a.src + t * PUSH_TILE_BYTES, a.dst + t * PUSH_TILE_BYTES, threadIdx.x);
        if constexpr (SIG != 0) {
            __syncthreads();
            if (threadIdx.x == 0) {
                if constexpr (SIG == 1) st_relaxed_sys(a.flags + c, a.epoch);
                else {
                    fence_sys();
                    st_relaxed_sys(a.flags + c, a.epoch);
                }
            }
        }
    }
}

// SIG 3: 256 copy threads + one signal warp (288 threads)
__global__ void __launch_bounds__(288, 4) push_ws(const __grid_constant__ VParams a) {
    const unsigned long long ntiles = a.nbytes / PUSH_TILE_BYTES;
    const unsigned long long tpc = 1ull << a.chunk_shift;
    const unsigned long long nchunks = (ntiles + tpc - 1) >> a.chunk_shift;
    const unsigned long long mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;
    if (threadIdx.x < 256) {
        for (unsigned long long k = 0; k < mine; ++k) {
            const unsigned long long c = blockIdx.x + k * gridDim.x;
            const unsigned long long t0 = c << a.chunk_shift;
            const unsigned long long t1 = (t0 + tpc < ntiles) ? t0 + tpc : ntiles;
#pragma unroll 1
            for (unsigned long long t = t0; t < t1; ++t) copy_tile256(a.src + t * PUSH_TILE_BYTES, a.dst + t * PUSH_TILE_BYTES, threadIdx.x);
            if (k >= 2) named_sync(3 + (int)(k & 1), 288);   // the signal for chunk k-2 is out: its barrier slot is free again
            named_arrive(1 + (int)(k & 1), 288);             // chunk k's stores are issued; go on with chunk k+1
        }
    } else {
        for (unsigned long long k = 0; k < mine; ++k) {
            named_sync(1 + (int)(k & 1), 288);
            if (threadIdx.x == 256) {
                fence_sys();
                st_relaxed_sys(a.flags + (blockIdx.x + k * gridDim.x), a.epoch);
            }
            __syncwarp();
            if (k + 2 < mine) named_arrive(3 + (int)(k & 1), 288);
        }
    }
}

// SIG 4: TMA push.  Three warps per CTA: a loader lane (cp.async.bulk global → shared ring), a storer lane
// (cp.async.bulk shared → the peer frame; stage recycling by wait_group.read, completion by wait_group), and a signaller
// warp whose lanes take completed tiles in turn: fence.acq_rel.sys + red.add on the counter of the tile's chunk.  No
// payload byte passes through a register and no SM load/store unit sees the NVLink back-pressure.  Tiles are dealt
// round-robin (tile t = blockIdx + k·grid), so the landing front advances grid × stage bytes at a time; a chunk's counter
// reaches tiles-per-chunk when all of its tiles (copied by different CTAs) have landed.  The flag word carries a count,
// the scan waits for `need` and (in the product) resets it.
template <int STAGES, int SB, int DEPTH>
__global__ void __launch_bounds__(96) push_tma(const __grid_constant__ VParams a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * SB);
    uint64_t* empty = full + STAGES;
    volatile unsigned int* done = reinterpret_cast<volatile unsigned int*>(empty + STAGES);  // tiles of this CTA known complete
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long ntiles = a.nbytes / SB;
    const unsigned long long mine = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;
    const unsigned int tiles_per_chunk_log2 = a.chunk_shift;  // here: log2(stage-sized tiles per counter)
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        *done = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0)
            for (unsigned long long k = 0; k < mine; ++k) {
                const int s = (int)(k % STAGES);
                const uint32_t round = (uint32_t)(k / STAGES);
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1u);
                const unsigned long long t = blockIdx.x + k * gridDim.x;
                mbar_expect_tx(&full[s], SB);
                bulk_load(smem + (size_t)s * SB, a.src + t * SB, SB, &full[s]);
            }
    } else if (warp == 1) {
        if (lane == 0) {
            for (unsigned long long k = 0; k < mine; ++k) {
                const int s = (int)(k % STAGES);
                mbar_wait(&full[s], (uint32_t)(k / STAGES) & 1u);
                const unsigned long long t = blockIdx.x + k * gridDim.x;
                bulk_store(a.dst + t * SB, smem + (size_t)s * SB, SB);
                bulk_commit();
                if (k > 0) {
                    bulk_wait_read<1>();
                    mbar_arrive(&empty[(int)((k - 1) % STAGES)]);
                }
                if (k >= DEPTH) {  // at most DEPTH stores of this CTA are in flight; older ones have landed
                    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(DEPTH) : "memory");
                    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32((const void*)done)), "r"((unsigned)(k - DEPTH + 1)) : "memory");
                }
            }
            if (mine > 0) {
                bulk_wait_read<0>();
                mbar_arrive(&empty[(int)((mine - 1) % STAGES)]);
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32((const void*)done)), "r"((unsigned)mine) : "memory");
            }
        }
    } else {
        for (unsigned long long i = lane; i < mine; i += 32) {
            unsigned int d;
            do {
                asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(d) : "r"(smem_u32((const void*)done)) : "memory");
                if (d <= i) __nanosleep(32);
            } while (d <= i);
            const unsigned long long t = blockIdx.x + i * gridDim.x;
            fence_sys();
            asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(a.flags + (t >> tiles_per_chunk_log2)), "r"(1u) : "memory");
        }
    }
}

// the sink's scan, epoch flavour, PARTS CTAs per chunk
struct SParams {
    const uint8_t* src;
    unsigned long long nbytes;
    const unsigned int* flags;
    unsigned int chunk_shift, epoch, parts_log2, count_mode;
    unsigned long long* acc;  // sum of folds, so the loads are not optimised away
    unsigned long long* stamps;  // globaltimer (ns) at which each chunk's flag was seen; [nchunks] = kernel start
    unsigned int diag;           // 0 = normal, 1 = do not wait for flags (read whatever is there), 2 = wait but do not read
};
__global__ void __launch_bounds__(256) scan_v(const __grid_constant__ SParams a) {
    const unsigned long long ntiles = a.nbytes / PUSH_TILE_BYTES;
    const unsigned long long tpc = 1ull << a.chunk_shift;
    const unsigned long long nchunks = (ntiles + tpc - 1) >> a.chunk_shift;
    const unsigned long long chunk_bytes = tpc * PUSH_TILE_BYTES;
    const unsigned long long nitems = nchunks << a.parts_log2;
    unsigned long long sum = 0, xr = 0;
    for (unsigned long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        const unsigned long long c = it >> a.parts_log2, part = it & ((1ull << a.parts_log2) - 1);
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            if (a.stamps && c == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); a.stamps[nchunks] = g; }
            const unsigned long long first = c << a.chunk_shift;
            const unsigned int need = a.count_mode ? (unsigned int)(ntiles - first < tpc ? ntiles - first : tpc) : a.epoch;
            for (; a.diag != 1;) {
                const unsigned int v = ld_acquire_sys(a.flags + c);
                if (a.count_mode ? v >= need : v == need) break;
                __nanosleep(64);
                if (clock64() - t0 > PUSH_TIMEOUT_CYCLES) break;
            }
            if (a.count_mode) const_cast<unsigned int*>(a.flags)[c] = 0u;
            if (a.stamps) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); a.stamps[c] = g; }
        }
        __syncthreads();
        if (a.diag == 2) continue;
        const unsigned long long b0 = c * chunk_bytes, b1 = (b0 + chunk_bytes < a.nbytes) ? b0 + chunk_bytes : a.nbytes;
        const unsigned long long pbytes = (b1 - b0) >> a.parts_log2;
        const uint4* sv = reinterpret_cast<const uint4*>(a.src + b0 + part * pbytes);
        const unsigned long long nvec = pbytes >> 4, vec0 = (b0 + part * pbytes) >> 4;
        for (unsigned long long i = threadIdx.x; i < nvec; i += 256ull * 8) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (i + 256ull * j < nvec) ? ld_coherent(sv + i + 256ull * j) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i + 256ull * j < nvec) fold16(v[j], (2ull * (vec0 + i + 256ull * j) + 1ull) * LZYB_K_POS, sum, xr);
        }
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) atomicAdd(a.acc, sum ^ xr);
}

__global__ void compare_kernel(const uint4* a, const uint4* b, unsigned long long nvec, unsigned long long* bad) {
    unsigned long long local = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nvec; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint4 x = a[i], y = b[i];
        if (x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w) ++local;
    }
    if (local) atomicAdd(bad, local);
}

struct Side {
    uint8_t *a, *b;
    unsigned int* flags;
    unsigned long long *bad, *acc, *stamps;
    cudaStream_t push_st, scan_st;
    cudaEvent_t e0, e1, p0, p1, sync_ev, sync_ev2;
};

int main() {
    if (getenv("MIB")) N = (unsigned long long)atoll(getenv("MIB")) << 20;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("{\"skip\": \"needs 2 GPUs\"}\n"); return 0; }
    Side s[2];
    int sms = 0;
    for (int i = 0; i < 2; ++i) {
        CK(cudaSetDevice(i));
        CK(cudaDeviceEnablePeerAccess(1 - i, 0));
        CK(cudaMalloc(&s[i].a, N)); CK(cudaMalloc(&s[i].b, N)); CK(cudaMalloc(&s[i].flags, 1 << 20));
        CK(cudaMalloc(&s[i].bad, 8)); CK(cudaMalloc(&s[i].acc, 8)); CK(cudaMalloc(&s[i].stamps, 8 * 40000)); CK(cudaMemset(s[i].stamps, 0, 8 * 40000));
        CK(cudaMemset(s[i].b, 0, N)); CK(cudaMemset(s[i].flags, 0, 1 << 20)); CK(cudaMemset(s[i].acc, 0, 8));
        fill_pattern_kernel<<<1184, 256>>>(s[i].a, N, 11 + i);
        CK(cudaStreamCreateWithFlags(&s[i].push_st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s[i].scan_st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&s[i].e0)); CK(cudaEventCreate(&s[i].e1)); CK(cudaEventCreate(&s[i].p0)); CK(cudaEventCreate(&s[i].p1));
        CK(cudaEventCreateWithFlags(&s[i].sync_ev, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s[i].sync_ev2, cudaEventDisableTiming));
        CK(cudaDeviceSynchronize());
    }
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned int epoch = 0;

    struct Cfg { int sig, shift, cps, scan, parts_log2, scan_ctas, stages, depth, diag; };
    // round 4: WHY is the sink's scan late when both directions are busy?  diag 1 = scan reads without waiting, diag 2 = scan
    // waits without reading; and the arrival time of every chunk's flag at the sink (globaltimer), against a linear landing
    const Cfg cfgs[] = {
        {4, 3, 1, 1, 0, 296, 6, 4, 0}, {4, 3, 1, 1, 0, 296, 6, 4, 1}, {4, 3, 1, 1, 0, 296, 6, 4, 2},
        {2, 1, 4, 1, 0, 296, 0, 0, 0}, {2, 1, 4, 1, 0, 296, 0, 0, 1}, {2, 1, 4, 1, 0, 296, 0, 0, 2},
        {4, 3, 1, 1, 0, 74, 6, 4, 0}, {4, 3, 1, 1, 0, 74, 6, 4, 2},
    };
    for (int i = 0; i < 2; ++i) {
        CK(cudaSetDevice(i));
        CK(cudaFuncSetAttribute(push_tma<3, 32768, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768 + 64));
        CK(cudaFuncSetAttribute(push_tma<3, 32768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768 + 64));
        CK(cudaFuncSetAttribute(push_tma<6, 32768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768 + 128));
        CK(cudaFuncSetAttribute(push_tma<6, 32768, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768 + 128));
    }
    for (const Cfg& cf : cfgs)
        for (int dirs = 1; dirs <= 2; ++dirs) {
            for (int i = 0; i < 2; ++i) { CK(cudaSetDevice(i)); CK(cudaMemset(s[i].flags, 0, 1 << 20)); }
            float ms[2] = {0, 0}, pms[2] = {0, 0};
            unsigned long long bad_total = 0;
            for (int pass = 0; pass < 2; ++pass) {
                const int reps = pass ? REPS : 2;
                for (int k = 0; k < dirs; ++k) {  // transfer k: source GPU k → sink GPU 1-k
                    CK(cudaSetDevice(k)); CK(cudaEventRecord(s[k].p0, s[k].push_st));
                    CK(cudaSetDevice(1 - k)); CK(cudaEventRecord(s[1 - k].e0, s[1 - k].scan_st));
                }
                for (int r = 0; r < reps; ++r) {
                    ++epoch;
                    for (int k = 0; k < dirs; ++k) {
                        const int snk = 1 - k;
                        if (cf.scan) {
                            CK(cudaSetDevice(snk));
                            SParams sp{s[snk].b, N, s[snk].flags, (unsigned)cf.shift, epoch, (unsigned)cf.parts_log2, cf.sig == 4 ? 1u : 0u, s[snk].acc, (pass && r == reps - 1) ? s[snk].stamps : nullptr, (unsigned)cf.diag};
                            scan_v<<<cf.scan_ctas, 256, 0, s[snk].scan_st>>>(sp);
                        }
                        CK(cudaSetDevice(k));
                        VParams vp{s[k].a, s[snk].b, N, s[snk].flags, (unsigned)cf.shift, epoch};
                        const unsigned long long nchunks = (N / PUSH_TILE_BYTES + (1ull << cf.shift) - 1) >> cf.shift;
                        const int grid = (int)std::min<unsigned long long>((unsigned long long)sms * cf.cps, nchunks);
                        if (cf.sig == 4) {
                            const int g4 = sms * cf.cps;
                            if (cf.stages == 3) {
                                constexpr size_t sm3 = 3 * 32768 + 2 * 3 * 8 + 16;
                                if (cf.depth == 2) push_tma<3, 32768, 2><<<g4, 96, sm3, s[k].push_st>>>(vp);
                                else push_tma<3, 32768, 4><<<g4, 96, sm3, s[k].push_st>>>(vp);
                            } else {
                                constexpr size_t sm6 = 6 * 32768 + 2 * 6 * 8 + 16;
                                if (cf.depth == 4) push_tma<6, 32768, 4><<<g4, 96, sm6, s[k].push_st>>>(vp);
                                else push_tma<6, 32768, 8><<<g4, 96, sm6, s[k].push_st>>>(vp);
                            }
                            continue;
                        }
                        switch (cf.sig) {
                            case 0: push_v<0><<<grid, 256, 0, s[k].push_st>>>(vp); break;
                            case 1: push_v<1><<<grid, 256, 0, s[k].push_st>>>(vp); break;
                            case 2: push_v<2><<<grid, 256, 0, s[k].push_st>>>(vp); break;
                            default: push_ws<<<grid, 288, 0, s[k].push_st>>>(vp); break;
                        }
                    }
                    // one transfer at a time per direction, like the product: the next push waits for this one's verdict
                    // (with a scan), and without a scan the sink stream simply follows the push
                    for (int k = 0; k < dirs; ++k) {
                        const int snk = 1 - k;
                        if (cf.scan) {
                            CK(cudaSetDevice(snk));
                            CK(cudaEventRecord(s[snk].sync_ev, s[snk].scan_st));
                            CK(cudaSetDevice(k));
                            CK(cudaStreamWaitEvent(s[k].push_st, s[snk].sync_ev, 0));
                        } else {
                            CK(cudaSetDevice(k));
                            CK(cudaEventRecord(s[k].sync_ev2, s[k].push_st));
                            CK(cudaSetDevice(snk));
                            CK(cudaStreamWaitEvent(s[snk].scan_st, s[k].sync_ev2, 0));
                        }
                    }
                }
                for (int k = 0; k < dirs; ++k) {
                    CK(cudaSetDevice(k)); CK(cudaEventRecord(s[k].p1, s[k].push_st));
                    CK(cudaSetDevice(1 - k)); CK(cudaEventRecord(s[1 - k].e1, s[1 - k].scan_st));
                }
                for (int k = 0; k < dirs; ++k) {
                    CK(cudaSetDevice(k)); CK(cudaEventSynchronize(s[k].p1)); CK(cudaEventElapsedTime(&pms[k], s[k].p0, s[k].p1));
                    CK(cudaSetDevice(1 - k)); CK(cudaEventSynchronize(s[1 - k].e1)); CK(cudaEventElapsedTime(&ms[k], s[1 - k].e0, s[1 - k].e1));
                }
                CK(cudaGetLastError());
            }
            for (int k = 0; k < dirs; ++k) {
                const int snk = 1 - k;
                CK(cudaSetDevice(snk));
                CK(cudaMemsetAsync(s[snk].bad, 0, 8, s[snk].scan_st));
                compare_kernel<<<1184, 256, 0, s[snk].scan_st>>>((const uint4*)s[k].a, (const uint4*)s[snk].b, N >> 4, s[snk].bad);
                unsigned long long hb = 0;
                CK(cudaMemcpyAsync(&hb, s[snk].bad, 8, cudaMemcpyDeviceToHost, s[snk].scan_st));
                CK(cudaStreamSynchronize(s[snk].scan_st));
                bad_total += hb;
                CK(cudaMemsetAsync(s[snk].b, 0, N, s[snk].scan_st));
                CK(cudaStreamSynchronize(s[snk].scan_st));
            }
            char lag[256] = "null";
            if (cf.scan && cf.diag != 1) {
                const unsigned long long nch = (N / PUSH_TILE_BYTES + (1ull << cf.shift) - 1) >> cf.shift;
                std::vector<unsigned long long> st(nch + 1);
                CK(cudaSetDevice(1));
                CK(cudaMemcpy(st.data(), s[1].stamps, 8 * (nch + 1), cudaMemcpyDeviceToHost));
                const double t0 = (double)st[nch];
                // when the flag of the chunk at 25 / 50 / 75 / 90 / 100 % of the payload was seen, in µs after the scan started
                snprintf(lag, sizeof(lag), "[%.0f, %.0f, %.0f, %.0f, %.0f]", (st[nch / 4] - t0) / 1e3, (st[nch / 2] - t0) / 1e3,
                         (st[3 * nch / 4] - t0) / 1e3, (st[nch * 9 / 10] - t0) / 1e3, (st[nch - 1] - t0) / 1e3);
            }
            const float worst = dirs == 2 ? std::max(ms[0], ms[1]) : ms[0];
            const float pworst = dirs == 2 ? std::max(pms[0], pms[1]) : pms[0];
            printf("{\"signal\": %d, \"chunk_KiB\": %d, \"ctas_per_sm\": %d, \"sink_scan\": %d, \"scan_ctas\": %d, \"stages\": %d, \"depth\": %d, \"directions\": %d, "
                   "\"ms_sink\": %.4f, \"GBps_per_dir\": %.1f, \"ms_push_stream\": %.4f, \"mismatched_vectors\": %llu, \"diag\": %d, \"flag_seen_us_at_25_50_75_90_100pct\": %s}\n",
                   cf.sig, 32 << cf.shift, cf.cps, cf.scan, cf.scan ? cf.scan_ctas : 0, cf.stages, cf.depth, dirs, worst / REPS, (double)N * REPS / worst / 1e6,
                   pworst / REPS, bad_total, cf.diag, lag);
            fflush(stdout);
        }
    return 0;
}
