// ldg_push_kernels.cuh — MEASUREMENT AID, not part of the product: the first push carrier of round 2 (LDG/STG.256 push kernel on
// the source GPU + a landed-scan kernel on the sink GPU), kept here so that tools/push_variants.cu — the sweep that showed why
// it was replaced (profiles/r02_push_variants*.jsonl) — still builds.  The product carrier is lzy_b200/csrc/push_kernels.cuh.
//
// (original header follows)
// push_kernels.cuh — the NVLink carrier of a slot transfer between two GPUs (sm_100a).
//
// Replaces the reference's network hop OutputSlot.ReadThread → gRPC → SlotInputTransfer.transferChunkTo
// (lzy/slots/.../OutputSlot.java:225-273, transfers/SlotInputTransfer.java:43-99) when producer and consumer sit
// on different GPUs of the box.
//
// Why a PUSH.  Measured with NVLink counters on a B200 pair (profiles/r02_peer_variants_ncu.csv): every SM-issued
// access crosses the link in 128-byte packets, whatever the instruction (LDG/STG.128, LDG/STG.256, UBLKCP):
//     pull:  1.125 wire bytes per payload byte on the response wire + 0.1875–0.25 on the REQUEST wire
//     push:  1.1875 wire bytes per payload byte on the write wire   + 0.003 (acks) on the other
// In a ring every GPU sends and receives at once, so a pull's requests compete with the payload the same GPU is
// serving: 884 GB/s raw ÷ (1.125 + 0.1875) = 673 GB/s per direction (measured 668–672), while posted writes
// keep 703–714.  The copy engine holds 772 but cannot digest; round 1 therefore ran engine copies with
// trailing scans (709–713 in the ring).  The pair of kernels below reaches the same figure with no copy engine:
//
//   push_kernel         runs on the SOURCE GPU: 256-bit loads from local HBM, 256-bit stores into the sink frame
//                       through the peer mapping (SASS LDG.E.ENL2.256 / STG.E.ENL2.256), persistent grid
//                       (4 CTAs/SM), each CTA owning whole chunks (≤ 8 tiles of 32 KiB) in ascending order; after a
//                       chunk's last store the CTA bumps the sink's progress counter for that chunk with
//                       red.release.sys (ordered after the stores of all its threads by the barrier before it).
//   landed_scan_kernel  runs on the SINK GPU at the same time: CTA c waits (ld.acquire.sys on its own HBM) until
//                       chunk c has landed, resets the counter (counters are zero at rest), folds the chunk into
//                       the slot digest straight from local HBM, and the last CTA verifies against the producer's
//                       digest and publishes the frame header + mailbox exactly like every other finalizing launch.
//
// The scan trails the landing by one chunk (256 KiB at 1 GiB), so verification costs ≈ 8 µs of tail instead of a
// second pass; the payload crosses NVLink once, is written to the sink's HBM once and read from it once.
// Counters live in a 16 KiB trailer of the sink frame so that one CUDA-IPC handle exposes frame and counters to a
// producer in another process.
#pragma once

#include "../lzy_b200/csrc/kernels.cuh"

namespace lzyb {

constexpr int PUSH_TILE_BYTES = 32768;        // 256 threads × 4 × 32 B
constexpr int PUSH_CTAS_PER_SM = 4;
constexpr int PUSH_TRAILER_BYTES = 16384;     // progress counters at the end of every large frame
constexpr int PUSH_FLAG_WORDS = PUSH_TRAILER_BYTES / 4;
constexpr int PUSH_MAX_CHUNK_TILES_LOG2 = 3;  // a chunk = up to 8 tiles (256 KiB) per counter through 1 GiB; finer for short
                                              // transfers (so the grid still fills), coarser beyond 1 GiB (4096 counters)
constexpr long long PUSH_TIMEOUT_CYCLES = 4000000000ll;  // ≈ 2 s: a dead producer must not hang the sink GPU

struct PushParams {
    const uint8_t* src;          // source bytes of this chunk (local HBM of the launching GPU)
    uint8_t* dst;                // sink bytes of this chunk (peer mapping), same alignment class as src
    unsigned long long nbytes;
    unsigned int* flags;         // the sink frame's counters (peer mapping)
    unsigned int chunk_shift;    // log2(tiles per counter)
    unsigned int pad;
};

// log2(tiles per chunk) for a transfer of nbytes — pure arithmetic, so producer and consumer agree without talking:
// the largest s ≤ 3 that still leaves ≥ 1024 chunks (a short transfer keeps enough CTAs busy), then as coarse as
// needed to fit the PUSH_FLAG_WORDS counters of the trailer
__host__ __device__ inline unsigned int push_chunk_shift(unsigned long long nbytes) {
    const unsigned long long ntiles = (nbytes + PUSH_TILE_BYTES - 1) / PUSH_TILE_BYTES;
    unsigned int s = 0;
    while (s < PUSH_MAX_CHUNK_TILES_LOG2 && (ntiles >> (s + 1)) >= 1024ull) ++s;
    while (((ntiles + (1ull << s) - 1) >> s) > (unsigned long long)PUSH_FLAG_WORDS) ++s;
    return s;
}

struct v256 {
    uint4 lo, hi;
};
__device__ __forceinline__ v256 ld256_stream(const uint8_t* p) {
    v256 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st256_stream(uint8_t* p, const v256& v) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.lo.x), "r"(v.lo.y),
                 "r"(v.lo.z), "r"(v.lo.w), "r"(v.hi.x), "r"(v.hi.y), "r"(v.hi.z), "r"(v.hi.w)
                 : "memory");
}

// V256: src and dst are both 32-byte aligned (frames are; a chunk offset that is only a multiple of 16 is not)
template <bool V256>
__global__ void __launch_bounds__(256, PUSH_CTAS_PER_SM) push_kernel(const __grid_constant__ PushParams a) {
    const unsigned long long ntiles = (a.nbytes + PUSH_TILE_BYTES - 1) / PUSH_TILE_BYTES;
    const unsigned long long tiles_per_chunk = 1ull << a.chunk_shift;
    const unsigned long long nchunks = (ntiles + tiles_per_chunk - 1) >> a.chunk_shift;
    // a CTA owns whole chunks (ascending, grid-strided): one release per chunk, and chunks land roughly in order
    for (unsigned long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const unsigned long long t0 = c << a.chunk_shift;
        const unsigned long long t1 = (t0 + tiles_per_chunk < ntiles) ? t0 + tiles_per_chunk : ntiles;
#pragma unroll 1
        for (unsigned long long t = t0; t < t1; ++t) {
            const unsigned long long base = t * PUSH_TILE_BYTES;
            const unsigned long long left = a.nbytes - base;
            const uint8_t* s = a.src + base;
            uint8_t* d = a.dst + base;
            if (left >= PUSH_TILE_BYTES) {
                if constexpr (V256) {
                    v256 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = ld256_stream(s + ((threadIdx.x + 256 * j) << 5));
#pragma unroll
                    for (int j = 0; j < 4; ++j) st256_stream(d + ((threadIdx.x + 256 * j) << 5), v[j]);
                } else {
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = ld_stream(reinterpret_cast<const uint4*>(s) + threadIdx.x + 256 * j);
#pragma unroll
                    for (int j = 0; j < 8; ++j) st_stream(reinterpret_cast<uint4*>(d) + threadIdx.x + 256 * j, v[j]);
                }
            } else {  // the ragged last tile: 16-byte vectors, then the byte tail
                const unsigned long long nvec = left >> 4;
                for (unsigned long long i = threadIdx.x; i < nvec; i += 256)
                    st_stream(reinterpret_cast<uint4*>(d) + i, ld_stream(reinterpret_cast<const uint4*>(s) + i));
                if (threadIdx.x == 0)
                    for (unsigned long long b = nvec << 4; b < left; ++b) d[b] = s[b];
            }
        }
        __syncthreads();  // every thread's stores of this chunk are issued, and ordered before thread 0's release
        if (threadIdx.x == 0) red_release_sys_add(a.flags + c, 1u);
    }
}

// The sink's half.  CopyParams as for every scan (src = the sink's own payload bytes of this chunk, dst = nullptr);
// signal_flags = the frame's counters, chunk_shift as the producer uses it.  Grid ≤ SM count: every CTA is resident,
// so a CTA that waits for a late chunk never keeps another from running.
__global__ void __launch_bounds__(256) landed_scan_kernel(const __grid_constant__ CopyParams a) {
    constexpr int UNROLL = 8;
    const unsigned long long ntiles = (a.nbytes + PUSH_TILE_BYTES - 1) / PUSH_TILE_BYTES;
    const unsigned long long tiles_per_chunk = 1ull << a.chunk_shift;
    const unsigned long long nchunks = (ntiles + tiles_per_chunk - 1) >> a.chunk_shift;
    const unsigned long long chunk_bytes = tiles_per_chunk * PUSH_TILE_BYTES;
    unsigned long long sum = 0ull, xr = 0ull;
    for (unsigned long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            bool landed = true;
            while (ld_acquire_sys(a.signal_flags + c) == 0u) {  // the pusher bumps a chunk's counter once, after its last store
                __nanosleep(64);
                if (clock64() - t0 > PUSH_TIMEOUT_CYCLES) {  // producer never delivered: scan what is there, the digest check fails the transfer
                    landed = false;
                    break;
                }
            }
            if (landed) a.signal_flags[c] = 0u;  // sole consumer of this counter: zero at rest for the frame's next life
        }
        __syncthreads();
        const unsigned long long b0 = c * chunk_bytes;
        const unsigned long long b1 = (b0 + chunk_bytes < a.nbytes) ? b0 + chunk_bytes : a.nbytes;
        const uint4* sv = reinterpret_cast<const uint4*>(a.src + b0);
        const unsigned long long nvec = (b1 - b0) >> 4;
        const unsigned long long vec0 = b0 >> 4;  // vector index of the chunk's first vector within this transfer chunk
        unsigned long long i = threadIdx.x;
        for (; i + 256ull * (UNROLL - 1) < nvec; i += 256ull * UNROLL) {
            uint4 v[UNROLL];
#pragma unroll
            for (int j = 0; j < UNROLL; ++j) v[j] = ld_coherent(sv + i + 256ull * j);
            unsigned long long key = (a.word_base + 2ull * (vec0 + i) + 1ull) * LZYB_K_POS;
#pragma unroll
            for (int j = 0; j < UNROLL; ++j) {
                fold16(v[j], key, sum, xr);
                key += 2ull * 256 * LZYB_K_POS;
            }
        }
        for (; i < nvec; i += 256) fold16(ld_coherent(sv + i), (a.word_base + 2ull * (vec0 + i) + 1ull) * LZYB_K_POS, sum, xr);
        const int tail = static_cast<int>((b1 - b0) & 15ull);  // only the payload's very last chunk can have one
        if (tail && threadIdx.x == 0) {
            const volatile uint8_t* tp = a.src + b0 + (nvec << 4);
            unsigned long long w[2] = {0ull, 0ull};
            for (int k = 0; k < tail; ++k) w[k >> 3] |= static_cast<unsigned long long>(tp[k]) << (8 * (k & 7));
            const unsigned long long first_word = a.word_base + 2ull * (vec0 + nvec);
            for (int k = 0; k < (tail + 7) >> 3; ++k) {
                const unsigned long long m = lzyb_word_term(w[k], first_word + k);
                sum += m;
                xr ^= m;
            }
        }
    }
    block_epilogue(a, sum, xr);
}

}  // namespace lzyb
