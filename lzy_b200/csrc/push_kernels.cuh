// push_kernels.cuh — how a slot frame crosses NVLink between two GPUs (sm_100a): a producer-issued PUSH.
//
// Replaces the reference's network hop OutputSlot.ReadThread → gRPC → SlotInputTransfer.transferChunkTo
// (lzy/slots/.../OutputSlot.java:225-273, transfers/SlotInputTransfer.java:43-99) when producer and consumer sit
// on different GPUs of the box.
//
// What the measurements of round 2 say (profiles/r02_peer_variants*.{jsonl,csv}, r02_push_variants*.jsonl, B200 pair):
//
//  * every SM- or TMA-issued access crosses the link in 128-byte packets, whatever the instruction (LDG/STG.128,
//    LDG/STG.256, UBLKCP).  NVLink counters per payload byte:
//        pull:  1.125 on the response wire + 0.1875–0.25 on the REQUEST wire        push:  1.1875 on the write wire + 0.003
//    In a ring every GPU sends and receives at once, so a pull's requests compete with the payload the same GPU is
//    serving: 884 GB/s raw ÷ (1.125 + 0.1875) = 673 GB/s per direction (measured 668–672); posted writes keep 700–714
//    in both situations.  (The copy engine holds 772 with larger packets — and cannot digest.)  ⇒ the PRODUCER's GPU
//    runs the kernel and stores through the peer mapping.
//  * a GPU whose NVLink egress is saturated by its own stores serves its local memory traffic badly: a read-only
//    scan that runs at 4.9 TB/s on an idle B200 gets 0.9–1.3 TB/s next to a push (LDG- or TMA-issued alike), and a
//    grid of pollers costs the push itself 3–12 %.  A sink-side scan that trails the landing — the first design of this
//    round: LDG push + landed-scan kernel, 700 GB/s one way, 430–675 both ways — therefore loses exactly in the
//    situation that matters.  ⇒ the digest is folded where the bytes already are: in the push kernel, from the
//    shared-memory stages the TMA engine has just filled, before the same stages are stored to the peer.
//
// So the carrier is ONE kernel, and it is the TMA ring of tma_kernels.cuh launched on the SOURCE GPU with its
// destination in the sink's frame: loader lane (cp.async.bulk global → smem), 8 consumer warps folding the digest
// from smem, storer lane (cp.async.bulk smem → peer global).  Payload bytes never pass through a register; the
// kernel's last CTA compares the digest of what it pushed with the digest the producer committed (a frame that changed
// after its commit — or a stale mapping — is a peer fault, LZYB_E_PEER) and writes the sink's frame header.
// Measured: 714 GB/s one way, 700 GB/s per direction with both directions busy — the SM-push ceiling of the packet
// format.  What the digest covers: commit → the bytes handed to the wire.  The wire itself is CRC-protected with
// replay and HBM is ECC-protected; a caller who wants the digest taken from the LANDED bytes selects the consumer pull
// (lzyb_set_copy_engine(2): the same ring launched on the sink GPU, 671 GB/s per direction in a ring).
//
// Between worker PROCESSES the sink frame is mapped through CUDA IPC and the consumer cannot wait on the producer's
// stream, so completion travels in band: the pushing kernel's CTAs fence their stores at system scope, the last
// CTA writes header + verdict and then a sequence word into the frame's trailer with release semantics, and the
// consumer's GPU runs `arrival_wait_kernel` — ONE thread — which spins on that word (ld.acquire.sys on its own HBM),
// then hands digest and verdict to the slot's mailbox like every other finalizing launch.  One poller costs nothing.
#pragma once

#include "kernels.cuh"

namespace lzyb {

constexpr int PUSH_TRAILER_BYTES = 256;  // the last bytes of every large frame: arrival word(s) of a cross-process push
constexpr long long PUSH_TIMEOUT_CYCLES = 8000000000ll;  // ≈ 4 s: a producer that never pushes must not hang the sink GPU

// trailer layout (64-bit words): [0] sequence of the last push that COMPLETED into this frame, [1] its verdict
// (1 = digest as committed, 2 = mismatch), [2..3] the digest it pushed, [4..5] its raw partials (sum, xor)
struct ArrivalWait {
    const volatile unsigned long long* trailer;  // the sink frame's own trailer (local HBM)
    unsigned long long seq;             // what the consumer told the producer to write
    Mailbox* mbox;
    unsigned int mbox_seq;
    unsigned int pad;
};

__global__ void arrival_wait_kernel(const __grid_constant__ ArrivalWait a) {
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    unsigned long long seen;
    bool arrived = true;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(const_cast<const unsigned long long*>(a.trailer)) : "memory");
        if (seen == a.seq) break;
        __nanosleep(200);
        if (clock64() - t0 > PUSH_TIMEOUT_CYCLES) {
            arrived = false;
            break;
        }
    }
    unsigned long long verdict = 2ull, d0 = 0ull, d1 = 0ull, sm = 0ull, xr = 0ull;
    if (arrived) {
        verdict = a.trailer[1];
        d0 = a.trailer[2];
        d1 = a.trailer[3];
        sm = a.trailer[4];
        xr = a.trailer[5];
    }
    a.mbox->digest[0] = d0;
    a.mbox->digest[1] = d1;
    a.mbox->sum = sm;
    a.mbox->xr = xr;
    a.mbox->seq = a.mbox_seq;
    __threadfence_system();
    a.mbox->status = !arrived ? 3u : (verdict == 1ull ? 1u : 2u);  // 3 = the producer's push never arrived
}

}  // namespace lzyb
