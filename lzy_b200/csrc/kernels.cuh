// kernels.cuh — sm_100a kernels of the slot data plane (device code + launch parameter blocks).
//
// What these replace in the reference (all CPU byte movers):
//   pipe→tmpfile IOUtils.copyLarge            lzy/slots/.../backends/OutputPipeBackend.java:48-50
//   file→2 MiB heap buffer→ByteString→gRPC    lzy/slots/.../OutputSlot.java:241-251
//   gRPC chunk→FileChannel.write @offset      lzy/slots/.../transfers/SlotInputTransfer.java:95-98
//   8096-byte Python FS copy loops            pylzy/lzy/storage/async_/fs.py:15-38
//   streaming MD5 of the serialized bytes     pylzy/lzy/api/v1/utils/hashing.py:36-38
//
// One kernel family does all of it: stream a byte range with 16-byte vector accesses, fold every
// 64-bit word into the position-keyed digest (include/lzyb_digest.h), optionally store it to the
// destination frame, reduce per warp by shuffle, publish two atomics per block, and let the last
// block to finish write the 128-byte frame header / verify against the producer's header.
// The work is pure byte movement: HBM-bound on one GPU, NVLink-bound between GPUs; there is no
// GEMM-shaped work and tensor cores are deliberately not involved.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lzy_b200.h"
#include "../../include/lzyb_digest.h"

namespace lzyb {

// Per-slot accumulator block in device memory (64 B, one L2 sector pair).
struct SlotCtl {
    unsigned long long sum;   // Σ word terms delivered so far
    unsigned long long xr;    // XOR of word terms delivered so far
    unsigned int ticket;      // blocks finished in the current launch
    unsigned int pad0;
    unsigned long long pad1[5];
};
static_assert(sizeof(SlotCtl) == 64, "SlotCtl must be 64 bytes");

// Per-slot mailbox in pinned, device-mapped host memory: the last block of a finalizing launch
// writes here so the host learns digest + verdict without a D2H memcpy.
struct Mailbox {
    volatile unsigned long long digest[2];
    volatile unsigned long long sum;
    volatile unsigned long long xr;
    volatile unsigned int status;    // 0 = pending, 1 = ok, 2 = digest mismatch, 3 = an awaited push never arrived
    volatile unsigned int seq;       // launch sequence number echoed back
    unsigned long long pad[3];
};
static_assert(sizeof(Mailbox) == 64, "Mailbox must be 64 bytes");

enum FinalizeMode : int {
    FIN_NONE = 0,     // chunk in the middle of a transfer: only accumulate
    FIN_PUBLISH = 1,  // pack / commit / host-ring frame: digest → header
    FIN_VERIFY = 2    // end of a pull: publish only if the digest equals the producer's
};

struct CopyParams {
    const uint8_t* src;         // source bytes of this chunk (any address space reachable by LDG)
    uint8_t* dst;               // destination bytes of this chunk, 16-B aligned; nullptr = scan only
    unsigned long long nbytes;  // bytes in this chunk
    unsigned long long word_base;  // index of the chunk's first 64-bit word within the payload
    SlotCtl* ctl;
    Mailbox* mbox;              // device-visible alias of the pinned mailbox (may be nullptr)
    lzyb_frame_header_t* dst_hdr;        // header to write on finalize (device or mapped host)
    unsigned long long expect[2];        // producer's published digest for FIN_VERIFY
    unsigned long long total_nbytes;     // payload size for lzyb_digest_finalize
    int fin;
    unsigned int seq;
    // 2-D gather (rows of row_bytes, src_pitch apart); rows == 0 → flat copy
    unsigned long long rows, row_bytes, src_pitch;
    // device-chained fan-out (chain_digest_kernel): progress counters, one per chunk of 2^chunk_shift tiles
    const unsigned int* wait_flags;   // upstream stage's counters (peer memory); nullptr = the source is final
    unsigned int* signal_flags;       // this stage's counters; nullptr = nobody reads from this stage
    unsigned int chunk_shift;
    unsigned int sys_fence;           // push into another process's frame: every CTA fences its stores at system scope before its ticket
    // cross-process push (push_kernels.cuh): the sink frame's trailer; the last CTA writes verdict + digest, then the sequence word
    unsigned long long* arrive;
    unsigned long long arrive_seq;
    lzyb_frame_header_t hdr_template;    // magic/version/dtype/shape for FIN_PUBLISH
};

// ---- memory access helpers: streaming data is touched once, keep it out of L1 ----------------
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// Assemble 16 logical bytes from a source whose alignment is only ALIGN bytes.
template <int ALIGN>
__device__ __forceinline__ uint4 load16(const uint8_t* p) {
    if constexpr (ALIGN >= 16) {
        return ld_stream(reinterpret_cast<const uint4*>(p));
    } else if constexpr (ALIGN == 8) {
        const uint2* q = reinterpret_cast<const uint2*>(p);
        uint2 a = __ldg(q), b = __ldg(q + 1);
        return make_uint4(a.x, a.y, b.x, b.y);
    } else if constexpr (ALIGN == 4) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        return make_uint4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    } else {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t b0 = __ldg(p + 4 * k), b1 = __ldg(p + 4 * k + 1), b2 = __ldg(p + 4 * k + 2),
                     b3 = __ldg(p + 4 * k + 3);
            w[k] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Fold one 16-byte vector (two 64-bit words) whose first word has position key `key0`
// (= (word_index + 1) * LZYB_K_POS) into the running partials.
__device__ __forceinline__ void fold16(const uint4& v, unsigned long long key0, unsigned long long& sum,
                                       unsigned long long& xr) {
    unsigned long long w0 = (static_cast<unsigned long long>(v.y) << 32) | v.x;
    unsigned long long w1 = (static_cast<unsigned long long>(v.w) << 32) | v.z;
    unsigned long long m0 = lzyb_mix64(w0 ^ key0);
    unsigned long long m1 = lzyb_mix64(w1 ^ (key0 + LZYB_K_POS));
    sum += m0 + m1;
    xr ^= m0 ^ m1;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_xor(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block epilogue shared by every streaming kernel: warp-shuffle reduce, two atomics per block,
// ticket, and the finalize step executed by the last block to arrive.
__device__ __forceinline__ void block_epilogue(const CopyParams& a, unsigned long long sum,
                                               unsigned long long xr) {
    __shared__ unsigned long long s_sum[32], s_xr[32];
    __shared__ unsigned int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    sum = warp_sum(sum);
    xr = warp_xor(xr);
    if (lane == 0) {
        s_sum[warp] = sum;
        s_xr[warp] = xr;
    }
    __syncthreads();
    if (warp == 0) {
        sum = lane < nwarps ? s_sum[lane] : 0ull;
        xr = lane < nwarps ? s_xr[lane] : 0ull;
        sum = warp_sum(sum);
        xr = warp_xor(xr);
        if (lane == 0) {
            atomicAdd(&a.ctl->sum, sum);
            atomicXor(&a.ctl->xr, xr);
            __threadfence();
            unsigned int t = atomicAdd(&a.ctl->ticket, 1u);
            s_last = (t == gridDim.x - 1) ? 1u : 0u;
        }
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    // ---- last block, one thread: everything every block published is visible after the fence --
    __threadfence();
    a.ctl->ticket = 0;
    if (a.fin == FIN_NONE) return;  // mid-transfer chunk: partials stay in ctl, nothing to publish
    // read the totals and leave the accumulators zeroed: a SlotCtl is always zero at rest
    unsigned long long tsum = atomicExch(&a.ctl->sum, 0ull), txr = atomicExch(&a.ctl->xr, 0ull);
    uint64_t dg[2];
    lzyb_digest_finalize(tsum, txr, a.total_nbytes, dg);
    unsigned int status = 1u;
    if (a.fin == FIN_VERIFY && (a.expect[0] != dg[0] || a.expect[1] != dg[1])) status = 2u;
    if (status == 1u && a.dst_hdr) {
        lzyb_frame_header_t h = a.hdr_template;
        h.digest[0] = dg[0];
        h.digest[1] = dg[1];
        h.sum = tsum;
        h.xr = txr;
        h.flags |= 1u;
        *a.dst_hdr = h;
    }
    if (a.mbox) {
        a.mbox->digest[0] = dg[0];
        a.mbox->digest[1] = dg[1];
        a.mbox->sum = tsum;
        a.mbox->xr = txr;
        a.mbox->seq = a.seq;
        __threadfence_system();
        a.mbox->status = status;
    }
    if (a.arrive) {  // the consumer's GPU is spinning on arrive[0] (arrival_wait_kernel): everything else first
        a.arrive[1] = status;
        a.arrive[2] = dg[0];
        a.arrive[3] = dg[1];
        a.arrive[4] = tsum;
        a.arrive[5] = txr;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.arrive), "l"(a.arrive_seq) : "memory");
    }
}

// Tail of a chunk: the (< 16) bytes after the last full vector, folded as up to two zero-padded
// 64-bit words and copied bytewise.  Executed by one thread of block 0.
__device__ __forceinline__ void fold_tail(const CopyParams& a, const uint8_t* src_tail, uint8_t* dst_tail,
                                          unsigned long long first_word, int tail,
                                          unsigned long long& sum, unsigned long long& xr) {
    unsigned long long w[2] = {0ull, 0ull};
    for (int k = 0; k < tail; ++k) {
        uint8_t b = src_tail[k];
        if (dst_tail) dst_tail[k] = b;
        w[k >> 3] |= static_cast<unsigned long long>(b) << (8 * (k & 7));
    }
    const int nwords = (tail + 7) >> 3;
    for (int k = 0; k < nwords; ++k) {
        unsigned long long m = lzyb_word_term(w[k], first_word + k);
        sum += m;
        xr ^= m;
    }
}

// ---- kernel 1: flat streaming copy/scan with fused digest (LDG.128 / STG.128) -----------------
// Grid is persistent-style: gridDim.x = SMs × resident CTAs; each CTA walks tiles of
// blockDim.x × UNROLL vectors, issuing all UNROLL loads before the first use (MLP per thread).
template <int ALIGN, int UNROLL, bool STORE>
__global__ void __launch_bounds__(256)
stream_digest_kernel(const __grid_constant__ CopyParams a) {
    const unsigned long long nvec = a.nbytes >> 4;
    const unsigned long long tile = static_cast<unsigned long long>(blockDim.x) * UNROLL;
    const unsigned long long ntiles = nvec / tile;
    const unsigned long long key_step = 2ull * blockDim.x * LZYB_K_POS;  // key delta between unroll slots
    unsigned long long sum = 0ull, xr = 0ull;
    uint4* dstv = reinterpret_cast<uint4*>(a.dst);

    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long base = t * tile + threadIdx.x;
        uint4 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j)
            v[j] = load16<ALIGN>(a.src + ((base + static_cast<unsigned long long>(j) * blockDim.x) << 4));
        unsigned long long key = (a.word_base + 2ull * base + 1ull) * LZYB_K_POS;
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            fold16(v[j], key, sum, xr);
            key += key_step;
            if constexpr (STORE) st_stream(dstv + base + static_cast<unsigned long long>(j) * blockDim.x, v[j]);
        }
    }
    // remainder vectors (less than one tile), spread over the grid
    for (unsigned long long i = ntiles * tile + static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
         i < nvec; i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        uint4 v = load16<ALIGN>(a.src + (i << 4));
        fold16(v, (a.word_base + 2ull * i + 1ull) * LZYB_K_POS, sum, xr);
        if constexpr (STORE) st_stream(dstv + i, v);
    }
    const int tail = static_cast<int>(a.nbytes & 15ull);
    if (tail && blockIdx.x == 0 && threadIdx.x == 0) {
        fold_tail(a, a.src + (nvec << 4), STORE ? a.dst + (nvec << 4) : nullptr, a.word_base + 2ull * nvec, tail,
                  sum, xr);
    }
    if (a.sys_fence) __threadfence_system();  // pushing into another process's frame: stores performed before this CTA's ticket
    block_epilogue(a, sum, xr);
}

// ---- kernel 2: 2-D gather pack (rows of row_bytes, src_pitch apart) ---------------------------
// row_bytes is a multiple of 16 and every row start is ALIGN-aligned (checked on the host);
// a work item is (row, column block) so consecutive threads read consecutive vectors of a row.
template <int ALIGN, int UNROLL>
__global__ void __launch_bounds__(256)
gather2d_digest_kernel(const __grid_constant__ CopyParams a) {
    const unsigned long long vpr = a.row_bytes >> 4;  // vectors per row
    const unsigned long long tile = static_cast<unsigned long long>(blockDim.x) * UNROLL;
    const unsigned long long cpr = (vpr + tile - 1) / tile;  // column blocks per row
    const unsigned long long nitems = a.rows * cpr;
    unsigned long long sum = 0ull, xr = 0ull;
    uint4* dstv = reinterpret_cast<uint4*>(a.dst);
    for (unsigned long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        const unsigned long long r = it / cpr, cb = it - r * cpr;
        const uint8_t* srow = a.src + r * a.src_pitch;
        const unsigned long long c0 = cb * tile + threadIdx.x;
        uint4 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const unsigned long long c = c0 + static_cast<unsigned long long>(j) * blockDim.x;
            if (c < vpr) v[j] = load16<ALIGN>(srow + (c << 4));
        }
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const unsigned long long c = c0 + static_cast<unsigned long long>(j) * blockDim.x;
            if (c < vpr) {
                const unsigned long long o = r * vpr + c;  // vector index in the packed payload
                fold16(v[j], (a.word_base + 2ull * o + 1ull) * LZYB_K_POS, sum, xr);
                st_stream(dstv + o, v[j]);
            }
        }
    }
    block_epilogue(a, sum, xr);
}

// ---- kernel 2b: one stage of a device-chained fan-out ---------------------------------------------
// Every stage of the chain (one per GPU) runs this kernel over the whole payload at the same time.
// A CTA owns one 32 KiB tile; before touching it, it waits until the upstream stage has published the
// chunk that contains the tile (a counter of finished tiles, polled over NVLink with acquire
// semantics); after storing it, it bumps its own counter with release semantics at system scope.  The
// payload therefore streams through all GPUs tile by tile inside ONE launch per GPU — no per-chunk
// launches, no host events — and the transfer of chunk c to GPU k+1 overlaps GPU k still receiving
// chunk c+1.  Loads are coherent (not .nc): the source is being written while we run.
__device__ __forceinline__ uint4 ld_coherent(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int CHAIN_UNROLL = 8;
constexpr int CHAIN_SPAN_TILES = 8;                       // a CTA owns 8 consecutive 32 KiB tiles = 256 KiB
constexpr long long CHAIN_TIMEOUT_CYCLES = 4000000000ll;  // ≈ 2 s: a dead upstream must not hang the GPU

// progress counters count finished *spans*; chunk_shift is log2(spans per counter)
__global__ void __launch_bounds__(256) chain_digest_kernel(const __grid_constant__ CopyParams a) {
    const unsigned long long nvec = a.nbytes >> 4;
    const unsigned long long span = 256ull * CHAIN_UNROLL * CHAIN_SPAN_TILES;  // vectors per span
    unsigned long long nspans = (nvec + span - 1) / span;
    if (nspans == 0) nspans = 1;  // a payload shorter than one vector still has its byte tail
    unsigned long long sum = 0ull, xr = 0ull;
    uint4* dstv = reinterpret_cast<uint4*>(a.dst);
    const uint4* srcv = reinterpret_cast<const uint4*>(a.src);
    for (unsigned long long sp = blockIdx.x; sp < nspans; sp += gridDim.x) {
        const unsigned long long c = sp >> a.chunk_shift;
        if (a.wait_flags) {
            if (threadIdx.x == 0) {
                const unsigned long long first = c << a.chunk_shift;
                const unsigned long long in_chunk =
                    (nspans - first) < (1ull << a.chunk_shift) ? (nspans - first) : (1ull << a.chunk_shift);
                const long long t0 = clock64();
                while (ld_acquire_sys(a.wait_flags + c) < static_cast<unsigned int>(in_chunk)) {
                    __nanosleep(128);
                    // upstream never delivered: carry on, the digest check will fail the transfer
                    if (clock64() - t0 > CHAIN_TIMEOUT_CYCLES) break;
                }
            }
            __syncthreads();
        }
#pragma unroll 1
        for (int tl = 0; tl < CHAIN_SPAN_TILES; ++tl) {
            const unsigned long long base = sp * span + static_cast<unsigned long long>(tl) * 256 * CHAIN_UNROLL + threadIdx.x;
            if (base - threadIdx.x >= nvec) break;
            uint4 v[CHAIN_UNROLL];
#pragma unroll
            for (int j = 0; j < CHAIN_UNROLL; ++j) {
                const unsigned long long i = base + static_cast<unsigned long long>(j) * 256;
                if (i < nvec) v[j] = ld_coherent(srcv + i);
            }
            unsigned long long key = (a.word_base + 2ull * base + 1ull) * LZYB_K_POS;
#pragma unroll
            for (int j = 0; j < CHAIN_UNROLL; ++j) {
                const unsigned long long i = base + static_cast<unsigned long long>(j) * 256;
                if (i < nvec) {
                    fold16(v[j], key, sum, xr);
                    st_stream(dstv + i, v[j]);
                }
                key += 2ull * 256 * LZYB_K_POS;
            }
        }
        if (sp == nspans - 1 && (a.nbytes & 15) && threadIdx.x == 0) {
            fold_tail(a, a.src + (nvec << 4), a.dst + (nvec << 4), a.word_base + 2ull * nvec,
                      static_cast<int>(a.nbytes & 15ull), sum, xr);
        }
        if (a.signal_flags) {
            __syncthreads();  // every thread's stores of this span are issued (and ordered before thread 0's release)
            if (threadIdx.x == 0) red_release_sys_add(a.signal_flags + c, 1u);
        }
    }
    block_epilogue(a, sum, xr);
}

// ---- kernel 3: finalize-only (EOS of a chunked transfer, or a zero-byte payload) -------------
__global__ void finalize_kernel(const __grid_constant__ CopyParams a) { block_epilogue(a, 0ull, 0ull); }

// ---- kernel 4: deterministic pattern fill (BASELINE config C3's payload) ----------------------
// word i of the buffer = mix64((seed + i + 1) * LZYB_K_POS): counter-based, so any range can be
// generated (and checked by the oracle) independently.
__global__ void __launch_bounds__(256) fill_pattern_kernel(uint8_t* dst, unsigned long long nbytes,
                                                            unsigned long long seed) {
    const unsigned long long nvec = nbytes >> 4;
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        unsigned long long w0 = lzyb_mix64((seed + 2ull * i + 1ull) * LZYB_K_POS);
        unsigned long long w1 = lzyb_mix64((seed + 2ull * i + 2ull) * LZYB_K_POS);
        st_stream(d + i, make_uint4(static_cast<uint32_t>(w0), static_cast<uint32_t>(w0 >> 32),
                                    static_cast<uint32_t>(w1), static_cast<uint32_t>(w1 >> 32)));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (unsigned long long b = nvec << 4; b < nbytes; ++b) {
            unsigned long long w = lzyb_mix64((seed + (b >> 3) + 1ull) * LZYB_K_POS);
            dst[b] = static_cast<uint8_t>(w >> (8 * (b & 7)));
        }
    }
}

// ---- kernel 5: flip one byte (fault injection: LZYB_FAULT_CORRUPT_ONCE) ------------------------
__global__ void corrupt_byte_kernel(uint8_t* p) { *p ^= 0x5a; }

}  // namespace lzyb
