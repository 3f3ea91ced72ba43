// tma_kernels.cuh — TMA bulk-copy variant of the streaming copy/scan + digest kernel (sm_100a).
//
// Large contiguous slot payloads do not need a register round trip: a producer lane issues
// `cp.async.bulk.shared::cluster.global` (SASS: UBLKCP) into a ring of shared-memory stages and
// signals an mbarrier with the transaction byte count; consumer warps fold the landed stage into
// the digest with conflict-free LDS.128 reads while a storer lane sends the very same stage back
// out with `cp.async.bulk.global.shared::cta` — the payload bytes never pass through registers,
// so LSU issue slots are spent on the digest only.  Stage recycling uses the usual full/empty
// mbarrier pair; the empty barrier collects one arrival per consumer warp plus one from the
// storer (after `cp.async.bulk.wait_group.read` says the stage has been read out).
//
// Same contract as stream_digest_kernel (kernels.cuh): CopyParams in, partials to SlotCtl,
// last block finalizes.  Replaces the same reference byte movers (OutputSlot.java:241-251,
// SlotInputTransfer.java:95-98).
#pragma once

#include "kernels.cuh"

namespace lzyb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

constexpr int TMA_CONSUMER_WARPS = 8;
constexpr int TMA_THREADS = (TMA_CONSUMER_WARPS + 2) * 32;  // + producer warp + storer warp

// Dynamic shared memory: STAGES × STAGE_BYTES of payload staging, then the barriers.
// FOLD = false is a measurement-only variant (lzyb_set_debug_flags bit 0): the same ring without the
// digest arithmetic, i.e. the ceiling of the copy structure itself.
template <int STAGES, int STAGE_BYTES, bool STORE, bool FOLD = true>
__global__ void __launch_bounds__(TMA_THREADS)
tma_stream_digest_kernel(const __grid_constant__ CopyParams a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(STAGES) * STAGE_BYTES);
    uint64_t* empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long ntiles = a.nbytes / STAGE_BYTES;  // full stages in this chunk
    // tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... belong to this CTA
    const unsigned long long my_tiles =
        ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TMA_CONSUMER_WARPS + (STORE ? 1 : 0));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned long long sum = 0ull, xr = 0ull;

    if (warp == 0) {
        // ===== producer: one lane feeds the ring =====
        if (lane == 0) {
            for (unsigned long long k = 0; k < my_tiles; ++k) {
                const int s = static_cast<int>(k % STAGES);
                const uint32_t round = static_cast<uint32_t>(k / STAGES);
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1u);
                const unsigned long long t = blockIdx.x + k * gridDim.x;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                bulk_load(smem + static_cast<size_t>(s) * STAGE_BYTES, a.src + t * STAGE_BYTES, STAGE_BYTES,
                          &full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== storer: send each landed stage back out, release it once it has been read =====
        if constexpr (STORE) {
            if (lane == 0) {
                for (unsigned long long k = 0; k < my_tiles; ++k) {
                    const int s = static_cast<int>(k % STAGES);
                    mbar_wait(&full[s], static_cast<uint32_t>(k / STAGES) & 1u);
                    const unsigned long long t = blockIdx.x + k * gridDim.x;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bulk_store(a.dst + t * STAGE_BYTES, smem + static_cast<size_t>(s) * STAGE_BYTES, STAGE_BYTES);
                    bulk_commit();
                    if (k > 0) {  // the previous store has drained its stage once ≤ 1 group is pending
                        bulk_wait_read<1>();
                        mbar_arrive(&empty[static_cast<int>((k - 1) % STAGES)]);
                    }
                }
                if (my_tiles > 0) {
                    bulk_wait_read<0>();
                    mbar_arrive(&empty[static_cast<int>((my_tiles - 1) % STAGES)]);
                    // a consumer in another process learns completion from a flag, not from this stream: this CTA's
                    // bulk stores must have been performed, not merely read out of shared memory, before its ticket
                    if (a.sys_fence) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
            }
        }
    } else {
        // ===== consumers: digest straight out of shared memory =====
        const int ctid = threadIdx.x - 64;                       // 0 .. 255
        constexpr int CTHREADS = TMA_CONSUMER_WARPS * 32;
        constexpr int VEC_PER_STAGE = STAGE_BYTES / 16;
        constexpr int ITERS = VEC_PER_STAGE / CTHREADS;
        static_assert(VEC_PER_STAGE % CTHREADS == 0, "stage must be a multiple of 4 KiB");
        for (unsigned long long k = 0; k < my_tiles; ++k) {
            const int s = static_cast<int>(k % STAGES);
            mbar_wait(&full[s], static_cast<uint32_t>(k / STAGES) & 1u);
            const unsigned long long t = blockIdx.x + k * gridDim.x;
            const uint4* sv = reinterpret_cast<const uint4*>(smem + static_cast<size_t>(s) * STAGE_BYTES);
            const unsigned long long vbase = t * VEC_PER_STAGE + ctid;  // vector index within the chunk
            unsigned long long key = (a.word_base + 2ull * vbase + 1ull) * LZYB_K_POS;
            if constexpr (FOLD) {
                uint4 v[ITERS];
#pragma unroll
                for (int j = 0; j < ITERS; ++j) v[j] = sv[ctid + j * CTHREADS];
#pragma unroll
                for (int j = 0; j < ITERS; ++j) {
                    fold16(v[j], key, sum, xr);
                    key += 2ull * CTHREADS * LZYB_K_POS;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // remainder after the last full stage (< STAGE_BYTES): plain vector path, block 0 only
        if (blockIdx.x == 0) {
            const unsigned long long done = ntiles * STAGE_BYTES;
            const unsigned long long rem = a.nbytes - done;
            const unsigned long long nvec = rem >> 4;
            const uint8_t* rs = a.src + done;
            uint4* rd = reinterpret_cast<uint4*>(a.dst + done);
            for (unsigned long long i = ctid; i < nvec; i += CTHREADS) {
                uint4 v = ld_stream(reinterpret_cast<const uint4*>(rs) + i);
                fold16(v, (a.word_base + 2ull * ((done >> 4) + i) + 1ull) * LZYB_K_POS, sum, xr);
                if constexpr (STORE) st_stream(rd + i, v);
            }
            const int tail = static_cast<int>(rem & 15ull);
            if (tail && ctid == 0) {
                fold_tail(a, rs + (nvec << 4), STORE ? a.dst + done + (nvec << 4) : nullptr,
                          a.word_base + 2ull * ((done >> 4) + nvec), tail, sum, xr);
            }
        }
    }
    if (a.sys_fence) __threadfence_system();  // generic stores of the remainder path and, cumulatively, the storer's bulk stores
    block_epilogue(a, sum, xr);
}

}  // namespace lzyb
