// capi.cu — implementation of the C ABI declared in include/lzy_b200.h: slot registry, HBM frame
// pool, kernel launch policy, CUDA-IPC peers, pinned host ring, fault injection.
//
// Reference seams implemented here (see the header for per-function citations):
//   OutputSlotBackend / InputSlotBackend   lzy/slots/src/main/java/ai/lzy/slots/backends/*.java
//   InputTransfer.transferChunkTo          lzy/slots/src/main/java/ai/lzy/slots/transfers/InputTransfer.java:8-11
//   SlotsService slot map                  lzy/slots/src/main/java/ai/lzy/slots/SlotsService.java:21-30
//
// There is no CPU data path in this file: without a CUDA device lzyb_init returns LZYB_E_NOGPU and
// every data call returns LZYB_E_STATE.
#include <cuda_runtime.h>
#include <unistd.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "host/host_common.h"
#include "kernels.cuh"
#include "push_kernels.cuh"
#include "tma_kernels.cuh"

using namespace lzyb;

namespace {

constexpr uint64_t HEADER_BYTES = 128;
static_assert(sizeof(lzyb_frame_header_t) == 128, "frame header must be 128 bytes");
static_assert(sizeof(lzyb_ipc_desc_t) == 256, "ipc descriptor must be 256 bytes");
constexpr int CTL_SLAB = 8192;        // slots that can be live per device
constexpr int MBOX_SLAB = 65536;      // mailboxes (pinned, all devices)
constexpr int GRID_AUTO = -1000;      // lzyb_set_grid: let launch_stream pick the grid shape
constexpr int HOST_CLASS_MIN_LOG2 = 16, HOST_CLASSES = 9;  // pinned host blocks: 64 KiB … 16 MiB
constexpr uint64_t BOUNCE_BYTES = 4ull << 20;      // D2H reads of at most this into pageable memory go through a pinned bounce buffer
constexpr uint64_t PUSH_MIN_DEFAULT = 4ull << 20;  // in-process GPU→GPU transfers at least this long are pushed from the source GPU (auto policy)
// TMA ring shapes that can be selected at run time (lzyb_set_tma_config): {stages, stage bytes, CTAs per SM}
struct TmaCfg {
    int stages, stage_bytes, ctas_per_sm;
};
constexpr TmaCfg kTmaCfgs[] = {{6, 16384, 2}, {3, 32768, 2}, {6, 32768, 1}, {12, 16384, 1}, {8, 8192, 3}, {4, 16384, 3},
                              {2, 8192, 1}};  // the last one is the stress shape: the shallowest ring, every stage recycled at once
constexpr int kNumTmaCfgs = sizeof(kTmaCfgs) / sizeof(kTmaCfgs[0]);
constexpr size_t tma_smem(int stages, int stage_bytes) { return static_cast<size_t>(stages) * stage_bytes + 2 * stages * sizeof(uint64_t); }

std::atomic<int> g_debug_flags{0};

template <int STAGES, int STAGE_BYTES>
void launch_tma(const CopyParams& p, bool store, int grid, cudaStream_t st, bool& attr_set) {
    constexpr size_t smem = tma_smem(STAGES, STAGE_BYTES);
    if (!attr_set) {  // the opt-in shared-memory size is a per-device function attribute
        cudaFuncSetAttribute(tma_stream_digest_kernel<STAGES, STAGE_BYTES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(tma_stream_digest_kernel<STAGES, STAGE_BYTES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (g_debug_flags.load() & 1) {  // measurement only: no digest arithmetic
        static bool nf_attr[16] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!nf_attr[dev & 15]) {
            cudaFuncSetAttribute(tma_stream_digest_kernel<STAGES, STAGE_BYTES, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            nf_attr[dev & 15] = true;
        }
        tma_stream_digest_kernel<STAGES, STAGE_BYTES, true, false><<<grid, TMA_THREADS, smem, st>>>(p);
        return;
    }
    if (store)
        tma_stream_digest_kernel<STAGES, STAGE_BYTES, true><<<grid, TMA_THREADS, smem, st>>>(p);
    else
        tma_stream_digest_kernel<STAGES, STAGE_BYTES, false><<<grid, TMA_THREADS, smem, st>>>(p);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    lzyb_host::vfail(code, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            cudaGetLastError(); /* reported here: do not leave it for this thread's next launch check */ \
            return fail(LZYB_E_LOCAL, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                         \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

struct DeviceCtx {
    int ordinal = 0;  // CUDA device ordinal
    int sm_count = 0;
    cudaStream_t stream = nullptr;       // library stream (blocking calls, broadcasts)
    std::mutex mu;
    std::multimap<size_t, void*> pool;   // free frames by capacity
    size_t pooled_bytes = 0;
    SlotCtl* ctl_slab = nullptr;
    std::vector<int> ctl_free;
    bool tma_attr[8] = {false, false, false, false, false, false, false, false};
    // DMA pull (launch_dma_verify): scans trail the copy engine on their own stream; events are recycled
    cudaStream_t verify_stream = nullptr;
    cudaStream_t push_stream = nullptr;  // this GPU as a SOURCE: push kernels of in-process transfers are issued here
    // "is this GPU serving a transfer right now?"  sourcing_ev[j] lives on managed device j and is re-recorded on the stream
    // that carries a transfer OUT of this GPU (its own push stream, or the pulling consumer's stream on device j)
    std::mutex src_mu;
    std::vector<cudaEvent_t> sourcing_ev;
    // small device→host reads into PAGEABLE memory (an ndarray argument of a host-side op, a pickle): cudaMemcpy into
    // pageable memory costs ≈ 0.19 ms per MiB-sized call on these hosts (driver-side staging); a pinned bounce buffer and
    // a memcpy cost a third of that (profiles/r02_profile_api_c1.txt)
    std::mutex bounce_mu;
    uint8_t* bounce = nullptr;
    std::mutex dma_mu;  // one copy-engine pull is issued at a time per sink GPU (event ring, verify-stream order)
    std::vector<cudaEvent_t> dma_events;
    size_t dma_ev_next = 0;
    cudaEvent_t next_dma_event() {
        if (dma_events.size() < 256) {
            cudaEvent_t e = nullptr;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            dma_events.push_back(e);
            return e;
        }
        return dma_events[dma_ev_next++ % dma_events.size()];
    }
};

struct Slot {
    lzyb_handle_t handle = 0;
    std::string id;
    int dev = 0;                 // index into g.devs
    uint64_t nbytes = 0;
    uint64_t cap = 0;            // bytes of the HBM frame (header + padded payload)
    uint8_t* buf = nullptr;      // header at buf, payload at buf + HEADER_BYTES
    SlotCtl* ctl = nullptr;
    int ctl_idx = -1;
    Mailbox* mbox = nullptr;
    int mbox_idx = -1;
    int state = LZYB_ST_EMPTY;
    bool is_input = false;
    bool remote = false;         // imported over CUDA IPC: we do not own buf
    void* ipc_base = nullptr;
    bool digest_pending = false; // a finalizing launch is in flight (or done) for the payload
    bool digest_ready = false;   // the verdict of that launch has been harvested: digest / sum / xr below are valid
    bool ctl_dirty = false;      // partial sums sit in ctl (mid-transfer)
    uint64_t delivered = 0;
    uint64_t digest[2] = {0, 0};
    uint64_t sum = 0, xr = 0;
    uint32_t dtype = LZYB_DT_BYTES, ndim = 0;
    uint64_t shape[LZYB_MAX_DIMS] = {0};
    unsigned seq = 0;
    cudaEvent_t last_ev = nullptr;
    cudaStream_t last_stream = nullptr;
    int fault_kind = LZYB_FAULT_NONE;
    uint64_t fault_arg = 0;
    bool movable = false;        // the owner declared a sole consumer: a whole-payload same-GPU pull moves the frame
    uint64_t push_seq = 0;       // input slot exported as a push sink: the sequence word the producer must write on completion
    bool quarantined = false;    // a push into this frame was declared lost: it may still land, so the frame is never reused
    std::atomic<int> refs{1};    // lzyb_slot_retain / lzyb_slot_close
    // asynchronous readers of this frame inside this process (consumer stream → event recorded after its last
    // pull launch): close waits for them before the frame goes back to the pool
    std::vector<std::pair<cudaStream_t, cudaEvent_t>> readers;
    std::mutex mu;
    uint8_t* payload() const { return buf + HEADER_BYTES; }
};
using SlotPtr = std::shared_ptr<Slot>;

struct RingFrame {
    uint64_t off, bytes;
    bool live;
};

struct Global {
    bool inited = false;
    int first = 0;
    std::vector<std::unique_ptr<DeviceCtx>> devs;
    std::shared_mutex reg_mu;
    std::unordered_map<lzyb_handle_t, SlotPtr> slots;
    std::unordered_map<std::string, lzyb_handle_t> by_id;
    std::atomic<uint64_t> next_handle{1};
    Mailbox* mbox_slab = nullptr;
    std::mutex mbox_mu;
    std::vector<int> mbox_free;
    // host ring
    uint8_t* ring = nullptr;
    uint64_t ring_bytes = 0, ring_head = 0, ring_used = 0;
    std::deque<RingFrame> ring_frames;
    std::mutex ring_mu;
    // CUDA-IPC mappings are expensive to open and the producer's pool recycles the same frames:
    // keep every mapping open until shutdown, keyed by (device, 64-byte handle)
    std::mutex ipc_mu;
    std::map<std::pair<int, std::string>, void*> ipc_cache;
    // policy
    std::atomic<int> engine{0};
    std::atomic<int> tma_cfg{1};  // 3 × 32 KiB, 2 CTAs/SM: best of the sweep in profiles/r01_microbench_local.jsonl
    std::atomic<int> ctas_per_sm{GRID_AUTO};
    std::atomic<int> global_fault{LZYB_FAULT_NONE};
    std::atomic<int> ctl_reserve{0};  // test hook (LZYB_FAULT_SLAB_RESERVE): pretend this many SlotCtl entries are taken
    std::atomic<uint64_t> dma_last_bytes{0};  // last sub-chunk of a copy-engine pull (0 = total/32 in 4..32 MiB)
    std::atomic<int> dma_growth{6};            // each earlier sub-chunk is this many times larger
    std::atomic<uint64_t> push_min{PUSH_MIN_DEFAULT};
    std::atomic<int> lazy_commit{0};  // lzyb_set_lazy_commit
    std::atomic<int> tma_waves{1};    // lzyb_set_transfer_waves
    // pinned host blocks (lzyb_host_alloc): free lists by class, outstanding blocks by address
    std::mutex host_mu;
    std::vector<void*> host_free[HOST_CLASSES];
    std::unordered_map<void*, int> host_out;
    uint64_t host_bytes = 0;  // outstanding + pooled
    uint64_t host_limit = 512ull << 20;
    std::atomic<uint64_t> st_push{0}, st_moves{0};
    std::atomic<uint64_t> next_push_seq{1};  // arrival words never repeat within this process
    // stats
    std::atomic<uint64_t> st_dma{0};
    std::atomic<uint64_t> st_launch{0}, st_pack{0}, st_pull{0}, st_scan{0}, st_frame{0}, st_bpack{0},
        st_bpull_l{0}, st_bpull_p{0}, st_bframe{0}, st_h2d{0}, st_d2h{0}, st_live{0}, st_hbm_live{0};
};
Global g;

uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

uint64_t frame_capacity(uint64_t nbytes) {
    uint64_t need = HEADER_BYTES + round_up(nbytes ? nbytes : 16, 256);
    if (need <= (1ull << 20)) {  // small: power-of-two classes from 4 KiB
        uint64_t c = 4096;
        while (c < need) c <<= 1;
        return c;
    }
    // large: 2 MiB granules (one TLB page each); the last 16 KiB are the progress counters of push transfers
    return round_up(need + PUSH_TRAILER_BYTES, 2ull << 20);
}
bool has_trailer(uint64_t cap) { return cap > (1ull << 20); }

SlotPtr find(lzyb_handle_t h) {
    std::shared_lock<std::shared_mutex> lk(g.reg_mu);
    auto it = g.slots.find(h);
    return it == g.slots.end() ? nullptr : it->second;
}

int alloc_frame(DeviceCtx& d, uint64_t cap, uint8_t** out) {
    {
        std::lock_guard<std::mutex> lk(d.mu);
        auto it = d.pool.find(cap);
        if (it != d.pool.end()) {
            *out = static_cast<uint8_t*>(it->second);
            d.pooled_bytes -= cap;
            d.pool.erase(it);
            return LZYB_OK;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e == cudaErrorMemoryAllocation) {  // give pooled frames back to the driver and retry once
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(d.mu);
        for (auto& kv : d.pool) cudaFree(kv.second);
        d.pool.clear();
        d.pooled_bytes = 0;
        e = cudaMalloc(&p, cap);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LZYB_E_LOCAL, "cudaMalloc(%llu) failed on GPU %d: %s", (unsigned long long)cap, d.ordinal,
                    cudaGetErrorString(e));
    }
    if (has_trailer(cap)) {  // counters are zero at rest; pooled frames keep that invariant (landed_scan_kernel resets what it consumes)
        e = cudaMemset(static_cast<uint8_t*>(p) + cap - PUSH_TRAILER_BYTES, 0, PUSH_TRAILER_BYTES);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(p);
            return fail(LZYB_E_LOCAL, "cudaMemset of a frame trailer failed on GPU %d: %s", d.ordinal, cudaGetErrorString(e));
        }
    }
    *out = static_cast<uint8_t*>(p);
    return LZYB_OK;
}

void free_frame(DeviceCtx& d, uint8_t* p, uint64_t cap) {
    std::lock_guard<std::mutex> lk(d.mu);
    d.pool.emplace(cap, p);
    d.pooled_bytes += cap;
}

cudaStream_t pick_stream(DeviceCtx& d, void* stream) { return stream ? static_cast<cudaStream_t>(stream) : d.stream; }

void note(Slot& s, cudaStream_t st) {
    if (!s.last_ev) cudaEventCreateWithFlags(&s.last_ev, cudaEventDisableTiming);
    cudaEventRecord(s.last_ev, st);
    s.last_stream = st;
}

void fill_template(const Slot& s, lzyb_frame_header_t* h) {
    memset(h, 0, sizeof(*h));
    memcpy(h->magic, "LZYB200", 8);
    h->version = LZYB_ABI_VERSION;
    h->dtype = s.dtype;
    h->nbytes = s.nbytes;
    h->ndim = s.ndim;
    memcpy(h->shape, s.shape, sizeof(h->shape));
}

int src_alignment(const void* p, uint64_t pitch) {
    uintptr_t v = reinterpret_cast<uintptr_t>(p) | static_cast<uintptr_t>(pitch);
    if ((v & 15) == 0) return 16;
    if ((v & 7) == 0) return 8;
    if ((v & 3) == 0) return 4;
    return 1;
}

template <int ALIGN>
void launch_flat(const CopyParams& p, bool store, int grid, cudaStream_t st) {
    if (store)
        stream_digest_kernel<ALIGN, 8, true><<<grid, 256, 0, st>>>(p);
    else
        stream_digest_kernel<ALIGN, 8, false><<<grid, 256, 0, st>>>(p);
}

// One launch of the streaming family.  Chooses the engine (LDG/STG vectors or TMA bulk) and a
// grid that is a multiple of the SM count for large chunks (persistent-style walk) and shrinks
// to the work for small ones.
int launch_stream(DeviceCtx& d, const CopyParams& p, bool store, cudaStream_t st, bool peer = false) {
    const int align = src_alignment(p.src, 0);
    const uint64_t nvec = p.nbytes >> 4;
    int engine = g.engine.load();
    const bool tma_ok = align == 16 && p.nbytes >= (1ull << 20) && (!store || (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0);
    // auto (measured, profiles/r01_summary.md §"engine policy"): a source on another GPU is NVLink-bound and the
    // TMA ring keeps the most bytes in flight with the fewest instructions (670 vs 634 GB/s per direction when
    // both directions are busy); a same-GPU copy is HBM-bound and an oversubscribed LDG/STG grid wins
    // (6.47 TB/s vs 5.97 TB/s for the persistent ring).
    if (engine == 0 || engine >= 3) engine = (peer && tma_ok) ? 2 : 1;  // 3 and 4 only change how peer transfers are carried
    if (engine == 2 && !tma_ok) engine = 1;
    if (engine == 2) {
        const int ci = g.tma_cfg.load();
        const TmaCfg cfg = kTmaCfgs[ci];
        uint64_t tiles = p.nbytes / cfg.stage_bytes;
        // waves > 1: more CTAs than are resident at once, each with 1/waves of the tiles — CTAs retire all along the transfer,
        // so a kernel on a higher-priority stream (an op's pack) gets SMs within a CTA's lifetime instead of the transfer's
        int grid = d.sm_count * cfg.ctas_per_sm * (peer ? g.tma_waves.load() : 1);
        if (g_debug_flags.load() & 2) grid = 4;  // stress: few CTAs, each cycling its ring thousands of times
        if ((uint64_t)grid > tiles) grid = (int)(tiles ? tiles : 1);
        switch (ci) {
            case 0: launch_tma<6, 16384>(p, store, grid, st, d.tma_attr[0]); break;
            case 1: launch_tma<3, 32768>(p, store, grid, st, d.tma_attr[1]); break;
            case 2: launch_tma<6, 32768>(p, store, grid, st, d.tma_attr[2]); break;
            case 3: launch_tma<12, 16384>(p, store, grid, st, d.tma_attr[3]); break;
            case 4: launch_tma<8, 8192>(p, store, grid, st, d.tma_attr[4]); break;
            case 5: launch_tma<4, 16384>(p, store, grid, st, d.tma_attr[5]); break;
            default: launch_tma<2, 8192>(p, store, grid, st, d.tma_attr[6]); break;
        }
    } else {
        const uint64_t tile = 256ull * 8;
        uint64_t want = (nvec + tile - 1) / tile;
        int cps = g.ctas_per_sm.load();
        if (cps == GRID_AUTO) {
            if (!store) cps = -8;                              // read-only scan: 8 tiles per CTA
            else if (p.nbytes >= (128ull << 20)) cps = -2;     // large copy: 2 tiles per CTA, oversubscribed
            else cps = 4;                                      // small copy: persistent, launch-latency bound anyway
        }
        // cps == 0: one tile per CTA (oversubscribed grid; the block scheduler staggers the CTAs, which keeps
        // DRAM reads and writes interleaved instead of phase-locked: profiles/r01_copy_variants.jsonl)
        uint64_t cap = cps > 0 ? (uint64_t)d.sm_count * cps : (1ull << 30);
        if (cps <= 0) {  // -T: T tiles per CTA (0 and -1: one)
            const uint64_t t = cps < -1 ? (uint64_t)(-cps) : 1;
            want = (want + t - 1) / t;
        }
        int grid = (int)(want < cap ? (want ? want : 1) : cap);
        switch (align) {
            case 16: launch_flat<16>(p, store, grid, st); break;
            case 8: launch_flat<8>(p, store, grid, st); break;
            case 4: launch_flat<4>(p, store, grid, st); break;
            default: launch_flat<1>(p, store, grid, st); break;
        }
    }
    g.st_launch++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LZYB_E_LOCAL, "kernel launch failed: %s", cudaGetErrorString(e));
    return LZYB_OK;
}

// Pull from a frame on another GPU with the copy engine, verified by trailing scans.
//
// Measured on a B200 pair (tools/push_pull.cu, profiles/r01_push_pull.jsonl), 1 GiB, both directions of the
// pair busy: SM loads over NVLink 640 (LDG) / 670 (TMA ring) GB/s per direction — the read requests of one
// direction share the wire with the other direction's payload; SM stores (push) 705; the copy engine 772,
// which is also its one-direction figure.  So the bytes ride the copy engine and the SMs do what only they can:
// the digest.  The chunk is cut into sub-chunks; the scan of sub-chunk j (stream_digest_kernel, STORE = false,
// reading the sink's own HBM at ~6 TB/s) runs on the device's verify stream while the engine copies j + 1.
// Every engine copy costs ~14 µs of fixed latency (measured: uniform 32/64/128/256 MiB sub-chunks lose exactly
// that per copy, and more copy streams do not hide it), and the scan of the final sub-chunk is pure tail, so
// the schedule is geometric: a small last sub-chunk (total/32, 4..32 MiB), each earlier one `growth` (6) times
// larger — a scan is ~8x faster than the copy that follows it, so it always finishes in that copy's shadow —
// and the first takes the rest: 1 GiB = 800 + 192 + 32 MiB, three copies and ~11 µs of tail.  The last scan
// carries the chunk's fin mode (FIN_VERIFY writes the header and the mailbox as in the fused kernels).
int launch_stream(DeviceCtx& d, const CopyParams& p, bool store, cudaStream_t st, bool peer);

std::vector<uint64_t> dma_schedule(uint64_t nbytes) {
    const uint64_t ragged = nbytes & 15ull, total = nbytes - ragged;
    const uint64_t growth = std::max<uint64_t>(1, g.dma_growth.load());
    uint64_t last = g.dma_last_bytes.load();
    if (last == 0) last = std::min<uint64_t>(std::max<uint64_t>((total / 32) & ~15ull, 4ull << 20), 32ull << 20);
    std::vector<uint64_t> rev;  // sizes from the end of the chunk towards its start
    uint64_t remaining = total, c = last;
    while (remaining > 0) {
        if (remaining <= c + c / 2 || rev.size() >= 60) {
            rev.push_back(remaining);
            break;
        }
        rev.push_back(c);
        remaining -= c;
        c *= growth;
    }
    if (rev.empty()) rev.push_back(0);
    rev.front() += ragged;  // the payload's ragged tail belongs to the last sub-chunk
    return std::vector<uint64_t>(rev.rbegin(), rev.rend());
}

int launch_dma_verify(DeviceCtx& d, const CopyParams& whole, cudaStream_t st) {
    std::lock_guard<std::mutex> issue(d.dma_mu);
    cudaStream_t vs = d.verify_stream;
    // Earlier chunks of this transfer (and the recycling of the sink frame) are ordered on `st`; every scan
    // waits for an event recorded on `st` after its copy, so it is ordered after all of that as well.
    cudaEvent_t ev = nullptr;
    const std::vector<uint64_t> sizes = dma_schedule(whole.nbytes);
    uint64_t off = 0;
    for (size_t j = 0; j < sizes.size(); ++j) {
        const uint64_t n = sizes[j];
        const bool last = j + 1 == sizes.size();
        if (n) {
            CU(cudaMemcpyAsync(whole.dst + off, whole.src + off, n, cudaMemcpyDefault, st));
            g.st_dma++;
        }
        ev = d.next_dma_event();
        CU(cudaEventRecord(ev, st));
        CU(cudaStreamWaitEvent(vs, ev, 0));
        CopyParams p = whole;
        p.src = whole.dst + off;  // scan what landed
        p.dst = nullptr;
        p.nbytes = n;
        p.word_base = whole.word_base + (off >> 3);
        if (!last) {
            p.fin = FIN_NONE;
            p.mbox = nullptr;
            p.dst_hdr = nullptr;
        }
        int rc = launch_stream(d, p, false, vs, false);
        if (rc != LZYB_OK) return rc;
        off += n;
    }
    ev = d.next_dma_event();
    CU(cudaEventRecord(ev, vs));
    CU(cudaStreamWaitEvent(st, ev, 0));  // `st` continues only with a verified frame
    return LZYB_OK;
}

// ---- the NVLink carrier: the streaming kernel launched on the SOURCE GPU (push_kernels.cuh) -----------------------
unsigned long long* frame_trailer(uint8_t* buf, uint64_t cap) { return reinterpret_cast<unsigned long long*>(buf + cap - PUSH_TRAILER_BYTES); }

cudaEvent_t add_reader(Slot& s, cudaStream_t st) {
    cudaEvent_t ev = nullptr;
    for (auto& r : s.readers)
        if (r.first == st) ev = r.second;
    if (!ev) {
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        s.readers.emplace_back(st, ev);
    }
    cudaEventRecord(ev, st);
    return ev;
}

// Bookkeeping for the auto policy: a transfer out of `sd` was just queued on `stream`, a stream of managed device `rec`
// (current device = that device).
void mark_sourcing(DeviceCtx& sd, int rec, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(sd.src_mu);
    if (sd.sourcing_ev.size() < g.devs.size()) sd.sourcing_ev.resize(g.devs.size(), nullptr);
    if (!sd.sourcing_ev[rec]) cudaEventCreateWithFlags(&sd.sourcing_ev[rec], cudaEventDisableTiming);
    cudaEventRecord(sd.sourcing_ev[rec], stream);
}
bool is_sourcing(DeviceCtx& d) {
    std::lock_guard<std::mutex> lk(d.src_mu);
    for (cudaEvent_t e : d.sourcing_ev)
        if (e && cudaEventQuery(e) == cudaErrorNotReady) return true;
    cudaGetLastError();
    return false;
}

// An in-process transfer between two GPUs.  `p` is the chunk as lzyb_transfer_chunk built it (src on the source GPU,
// dst / header on the sink GPU, ctl on the sink GPU — reached through the peer mapping —, mailbox in pinned host memory);
// the caller's current device is the sink's.  The kernel runs on the source GPU's push stream; the sink stream `st`
// continues only after it (cross-device event), so everything queued on `st` afterwards sees a complete, verified frame.
int launch_push_inproc(DeviceCtx& sd, DeviceCtx& dd, Slot& src, Slot& dst, const CopyParams& p, cudaStream_t st) {
    cudaEvent_t done = nullptr;
    {
        // the push may start once everything queued on the sink stream so far has run and the producer's payload is final
        std::lock_guard<std::mutex> issue(dd.dma_mu);
        cudaEvent_t ev = dd.next_dma_event();
        CU(cudaEventRecord(ev, st));
        DeviceGuard sg(sd.ordinal);
        CU(cudaStreamWaitEvent(sd.push_stream, ev, 0));
        if (src.last_ev) CU(cudaStreamWaitEvent(sd.push_stream, src.last_ev, 0));
        int rc = launch_stream(sd, p, true, sd.push_stream, true);
        if (rc != LZYB_OK) return rc;
        g.st_push++;
        mark_sourcing(sd, src.dev, sd.push_stream);
        add_reader(src, sd.push_stream);         // neither frame is recycled under a push that is still running
        done = add_reader(dst, sd.push_stream);
    }
    CU(cudaStreamWaitEvent(st, done, 0));
    return LZYB_OK;
}

void arm_mailbox(Slot& s, CopyParams& p) {
    s.seq++;
    s.mbox->status = 0;
    s.mbox->seq = 0;
    p.mbox = s.mbox;
    p.seq = s.seq;
}

// wait for the finalizing launch of `s` and fold its verdict into the slot; returns the status word
int harvest(Slot& s) {
    if (s.last_ev) {
        cudaError_t e = cudaEventSynchronize(s.last_ev);
        if (e != cudaSuccess) return fail(LZYB_E_LOCAL, "wait on slot %s failed: %s", s.id.c_str(), cudaGetErrorString(e));
    }
    if (s.mbox->seq != s.seq || s.mbox->status == 0)
        return fail(LZYB_E_LOCAL, "slot %s: finalize did not report (seq %u/%u status %u)", s.id.c_str(), s.mbox->seq,
                    s.seq, s.mbox->status);
    s.digest[0] = s.mbox->digest[0];
    s.digest[1] = s.mbox->digest[1];
    s.sum = s.mbox->sum;
    s.xr = s.mbox->xr;
    s.ctl_dirty = false;
    s.digest_ready = s.mbox->status == 1u;
    return static_cast<int>(s.mbox->status);
}

// A slot that was made READY in stream order (lzyb_set_lazy_commit) learns its digest here, the first time somebody
// needs it: the host then waits for the pack / scan that has been running since.
int ensure_digest(Slot& s) {
    if (s.digest_ready || !s.digest_pending || s.state != LZYB_ST_READY) return LZYB_OK;
    DeviceGuard dg(g.devs[s.dev]->ordinal);
    int st = harvest(s);
    if (st < 0) return st;
    if (st != 1) return fail(LZYB_E_LOCAL, "slot %s: unexpected finalize status %d", s.id.c_str(), st);
    return LZYB_OK;
}

int new_slot(const char* slot_id, int gpu, uint64_t nbytes, bool is_input, lzyb_handle_t* out) {
    if (!g.inited) return fail(LZYB_E_STATE, "lzyb_init has not been called (or found no GPU)");
    if (!slot_id || !out) return fail(LZYB_E_ARG, "null argument");
    int di = gpu - g.first;
    if (di < 0 || di >= (int)g.devs.size()) return fail(LZYB_E_ARG, "GPU %d is not managed by this process", gpu);
    DeviceCtx& d = *g.devs[di];
    if (is_input && g.global_fault.load() == LZYB_FAULT_LOCAL_OPEN) {
        g.global_fault = LZYB_FAULT_NONE;
        return fail(LZYB_E_LOCAL, "injected fault: cannot open input slot %s", slot_id);
    }
    DeviceGuard dg(d.ordinal);
    auto s = std::make_shared<Slot>();
    s->id = slot_id;
    s->dev = di;
    s->nbytes = nbytes;
    s->cap = frame_capacity(nbytes);
    s->is_input = is_input;
    int rc = alloc_frame(d, s->cap, &s->buf);
    if (rc != LZYB_OK) return rc;
    {
        std::lock_guard<std::mutex> lk(d.mu);
        if (!d.ctl_free.empty() && (int)d.ctl_free.size() > g.ctl_reserve.load()) {
            s->ctl_idx = d.ctl_free.back();
            d.ctl_free.pop_back();
            s->ctl = d.ctl_slab + s->ctl_idx;
        }
    }
    if (s->ctl_idx < 0) {  // decided under d.mu, undone outside it: free_frame takes d.mu itself
        free_frame(d, s->buf, s->cap);
        return fail(LZYB_E_LOCAL, "too many live slots on GPU %d", gpu);
    }
    {
        std::lock_guard<std::mutex> lk(g.mbox_mu);
        if (!g.mbox_free.empty()) {
            s->mbox_idx = g.mbox_free.back();
            g.mbox_free.pop_back();
            s->mbox = g.mbox_slab + s->mbox_idx;
            memset((void*)s->mbox, 0, sizeof(Mailbox));
        }
    }
    if (s->mbox_idx < 0) {  // give back what this slot already took
        free_frame(d, s->buf, s->cap);
        std::lock_guard<std::mutex> lk(d.mu);
        d.ctl_free.push_back(s->ctl_idx);
        return fail(LZYB_E_LOCAL, "out of mailboxes");
    }
    s->handle = g.next_handle++;
    {
        std::unique_lock<std::shared_mutex> lk(g.reg_mu);
        g.slots[s->handle] = s;
        g.by_id[s->id] = s->handle;
    }
    g.st_live++;
    g.st_hbm_live += s->cap;
    *out = s->handle;
    return LZYB_OK;
}

}  // namespace

extern "C" {

int lzyb_abi_version(void) { return LZYB_ABI_VERSION; }
const char* lzyb_last_error(void) { return lzyb_host::err_buf(); }
int lzyb_is_initialized(void) { return g.inited ? 1 : 0; }
int lzyb_set_last_error(const char* message) {
    fail(0, "%s", message ? message : "");
    return LZYB_OK;
}
int lzyb_device_count(void) { return g.inited ? (int)g.devs.size() : 0; }

const char* lzyb_strerror(int code) {
    switch (code) {
        case LZYB_OK: return "ok";
        case LZYB_EOF: return "end of stream";
        case LZYB_E_PEER: return "peer fault (ReadException)";
        case LZYB_E_LOCAL: return "local fault (IOException)";
        case LZYB_E_ARG: return "bad argument";
        case LZYB_E_STATE: return "wrong state";
        case LZYB_E_NOGPU: return "no CUDA device";
        case LZYB_E_NOT_FOUND: return "NOT_FOUND";
        case LZYB_E_PRECONDITION: return "FAILED_PRECONDITION";
        case LZYB_E_UNAVAILABLE: return "UNAVAILABLE";
        case LZYB_E_INTERNAL: return "INTERNAL";
        case LZYB_E_UNIMPLEMENTED: return "UNIMPLEMENTED";
        case LZYB_E_INVALID: return "INVALID_ARGUMENT";
        case LZYB_E_EXISTS: return "ALREADY_EXISTS";
        default: return "unknown";
    }
}

int lzyb_init(int first_gpu, int n_gpus, uint64_t host_ring_bytes) {
    if (g.inited) return LZYB_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(LZYB_E_NOGPU, "no CUDA device visible (%s); liblzy_b200 has no CPU fallback",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    if (first_gpu < 0 || first_gpu >= count) return fail(LZYB_E_ARG, "first_gpu %d out of range (%d devices)", first_gpu, count);
    if (n_gpus <= 0 || first_gpu + n_gpus > count) n_gpus = count - first_gpu;
    int prev = 0;
    cudaGetDevice(&prev);
    g.first = first_gpu;
    for (int i = 0; i < n_gpus; ++i) {
        auto d = std::make_unique<DeviceCtx>();
        d->ordinal = first_gpu + i;
        CU(cudaSetDevice(d->ordinal));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, d->ordinal));
        if (prop.major < 10) {
            return fail(LZYB_E_NOGPU, "GPU %d is sm_%d%d; this library carries sm_100a code only", d->ordinal, prop.major,
                        prop.minor);
        }
        d->sm_count = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&d->verify_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&d->push_stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&d->ctl_slab, sizeof(SlotCtl) * CTL_SLAB));
        CU(cudaMemset(d->ctl_slab, 0, sizeof(SlotCtl) * CTL_SLAB));
        for (int k = CTL_SLAB - 1; k >= 0; --k) d->ctl_free.push_back(k);
        g.devs.push_back(std::move(d));
    }
    // all-to-all peer access between the managed devices (NVSwitch: every pair is one hop)
    for (int i = 0; i < n_gpus; ++i) {
        cudaSetDevice(first_gpu + i);
        for (int j = 0; j < n_gpus; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, first_gpu + i, first_gpu + j);
            if (can) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(first_gpu + j, 0);
                if (pe != cudaSuccess) cudaGetLastError();  // already enabled (e.g. by torch) is fine
            }
        }
    }
    cudaSetDevice(first_gpu);
    CU(cudaHostAlloc((void**)&g.mbox_slab, sizeof(Mailbox) * MBOX_SLAB, cudaHostAllocPortable | cudaHostAllocMapped));
    memset((void*)g.mbox_slab, 0, sizeof(Mailbox) * MBOX_SLAB);
    for (int k = MBOX_SLAB - 1; k >= 0; --k) g.mbox_free.push_back(k);
    g.ring_bytes = host_ring_bytes ? round_up(host_ring_bytes, 4096) : (64ull << 20);
    CU(cudaHostAlloc((void**)&g.ring, g.ring_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    cudaSetDevice(prev);
    g.inited = true;
    return LZYB_OK;
}

int lzyb_shutdown(void) {
    if (!g.inited) return LZYB_OK;
    std::vector<lzyb_handle_t> hs;
    {
        std::shared_lock<std::shared_mutex> lk(g.reg_mu);
        for (auto& kv : g.slots) hs.push_back(kv.first);
    }
    for (auto h : hs) lzyb_slot_close(h);
    {
        std::lock_guard<std::mutex> lk(g.ipc_mu);
        for (auto& kv : g.ipc_cache) {
            cudaSetDevice(kv.first.first);
            cudaIpcCloseMemHandle(kv.second);
        }
        g.ipc_cache.clear();
    }
    for (auto& d : g.devs) {
        cudaSetDevice(d->ordinal);
        cudaDeviceSynchronize();
        for (auto& kv : d->pool) cudaFree(kv.second);
        d->pool.clear();
        cudaFree(d->ctl_slab);
        for (cudaEvent_t e : d->dma_events) cudaEventDestroy(e);
        cudaStreamDestroy(d->verify_stream);
        cudaStreamDestroy(d->push_stream);
        for (cudaEvent_t e : d->sourcing_ev)
            if (e) cudaEventDestroy(e);
        if (d->bounce) cudaFreeHost(d->bounce);
        cudaStreamDestroy(d->stream);
    }
    g.devs.clear();
    {
        std::lock_guard<std::mutex> lk(g.host_mu);
        for (int c = 0; c < HOST_CLASSES; c++) {
            for (void* p : g.host_free[c]) {
                cudaFreeHost(p);
                g.host_bytes -= 1ull << (HOST_CLASS_MIN_LOG2 + c);
            }
            g.host_free[c].clear();
        }
    }
    cudaFreeHost((void*)g.mbox_slab);
    cudaFreeHost(g.ring);
    g.mbox_slab = nullptr;
    g.ring = nullptr;
    g.mbox_free.clear();
    g.ring_frames.clear();
    g.ring_head = g.ring_used = 0;
    g.inited = false;
    return LZYB_OK;
}

int lzyb_slot_create(const char* slot_id, int gpu, uint64_t nbytes, lzyb_handle_t* out) {
    return new_slot(slot_id, gpu, nbytes, false, out);
}
int lzyb_input_open(const char* slot_id, int gpu, uint64_t nbytes, lzyb_handle_t* out) {
    return new_slot(slot_id, gpu, nbytes, true, out);
}

int lzyb_slot_find(const char* slot_id, lzyb_handle_t* out) {
    if (!slot_id || !out) return fail(LZYB_E_ARG, "null argument");
    std::shared_lock<std::shared_mutex> lk(g.reg_mu);
    auto it = g.by_id.find(slot_id);
    if (it == g.by_id.end()) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", slot_id);
    *out = it->second;
    return LZYB_OK;
}

int lzyb_slot_device_ptr(lzyb_handle_t h, void** dptr) {
    auto s = find(h);
    if (!s || !dptr) return fail(LZYB_E_ARG, "unknown slot handle %llu", (unsigned long long)h);
    if (s->state == LZYB_ST_CLOSED) return fail(LZYB_E_STATE, "slot %s is closed", s->id.c_str());
    *dptr = s->payload();
    return LZYB_OK;
}

int lzyb_slot_set_meta(lzyb_handle_t h, uint32_t dtype, uint32_t ndim, const uint64_t* shape) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    if (ndim > LZYB_MAX_DIMS) return fail(LZYB_E_ARG, "ndim %u > %d", ndim, LZYB_MAX_DIMS);
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != LZYB_ST_EMPTY || s->digest_pending) return fail(LZYB_E_STATE, "slot %s: meta must be set before data", s->id.c_str());
    s->dtype = dtype;
    s->ndim = ndim;
    memset(s->shape, 0, sizeof(s->shape));
    for (uint32_t i = 0; i < ndim; ++i) s->shape[i] = shape[i];
    return LZYB_OK;
}

static int pack_common(lzyb_handle_t h, const void* src, uint64_t rows, uint64_t row_bytes, uint64_t pitch, void* stream) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->is_input || s->remote) return fail(LZYB_E_STATE, "slot %s is not a local output slot", s->id.c_str());
    if (s->state != LZYB_ST_EMPTY || s->digest_pending) return fail(LZYB_E_STATE, "slot %s already holds data", s->id.c_str());
    const uint64_t nbytes = rows ? rows * row_bytes : row_bytes;
    if (nbytes != s->nbytes) return fail(LZYB_E_ARG, "slot %s holds %llu bytes, pack of %llu", s->id.c_str(),
                                         (unsigned long long)s->nbytes, (unsigned long long)nbytes);
    if (nbytes && !src) return fail(LZYB_E_ARG, "null source");
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = static_cast<const uint8_t*>(src);
    p.dst = s->payload();
    p.nbytes = nbytes;
    p.ctl = s->ctl;
    p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(s->buf);
    p.total_nbytes = s->nbytes;
    p.fin = FIN_PUBLISH;
    fill_template(*s, &p.hdr_template);
    arm_mailbox(*s, p);
    int rc;
    if (rows && pitch != row_bytes) {
        if (row_bytes & 15) return fail(LZYB_E_ARG, "2-D pack needs row_bytes %% 16 == 0 (got %llu)", (unsigned long long)row_bytes);
        p.rows = rows;
        p.row_bytes = row_bytes;
        p.src_pitch = pitch;
        const int align = src_alignment(src, pitch);
        const uint64_t tile = 256ull * 4;
        uint64_t items = rows * ((row_bytes / 16 + tile - 1) / tile);
        int gcps = g.ctas_per_sm.load();
        int grid = d.sm_count * (gcps > 0 ? gcps : 8);
        if ((uint64_t)grid > items) grid = (int)(items ? items : 1);
        switch (align) {
            case 16: gather2d_digest_kernel<16, 4><<<grid, 256, 0, st>>>(p); break;
            case 8: gather2d_digest_kernel<8, 4><<<grid, 256, 0, st>>>(p); break;
            case 4: gather2d_digest_kernel<4, 4><<<grid, 256, 0, st>>>(p); break;
            default: gather2d_digest_kernel<1, 4><<<grid, 256, 0, st>>>(p); break;
        }
        g.st_launch++;
        cudaError_t e = cudaGetLastError();
        rc = e == cudaSuccess ? LZYB_OK : fail(LZYB_E_LOCAL, "gather launch failed: %s", cudaGetErrorString(e));
    } else {
        rc = launch_stream(d, p, true, st);
    }
    if (rc != LZYB_OK) return rc;
    g.st_pack++;
    g.st_bpack += nbytes;
    s->digest_pending = true;
    note(*s, st);
    return LZYB_OK;
}

int lzyb_slot_pack(lzyb_handle_t h, const void* src_dptr, uint64_t nbytes, void* stream) {
    return pack_common(h, src_dptr, 0, nbytes, nbytes, stream);
}
int lzyb_slot_pack_2d(lzyb_handle_t h, const void* src_dptr, uint64_t rows, uint64_t row_bytes, uint64_t src_pitch, void* stream) {
    if (rows == 0) return pack_common(h, src_dptr, 0, 0, 0, stream);
    return pack_common(h, src_dptr, rows, row_bytes, src_pitch, stream);
}

static int scan_publish(Slot& s, DeviceCtx& d, cudaStream_t st) {
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = s.payload();
    p.dst = nullptr;
    p.nbytes = s.nbytes;
    p.ctl = s.ctl;
    p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(s.buf);
    p.total_nbytes = s.nbytes;
    p.fin = FIN_PUBLISH;
    fill_template(s, &p.hdr_template);
    arm_mailbox(s, p);
    int rc = launch_stream(d, p, false, st);
    if (rc != LZYB_OK) return rc;
    g.st_scan++;
    s.digest_pending = true;
    note(s, st);
    return LZYB_OK;
}

int lzyb_slot_pack_host(lzyb_handle_t h, const void* src_host, uint64_t nbytes, void* stream) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->is_input || s->remote) return fail(LZYB_E_STATE, "slot %s is not a local output slot", s->id.c_str());
    if (s->state != LZYB_ST_EMPTY || s->digest_pending) return fail(LZYB_E_STATE, "slot %s already holds data", s->id.c_str());
    if (nbytes != s->nbytes) return fail(LZYB_E_ARG, "slot %s holds %llu bytes, pack of %llu", s->id.c_str(),
                                         (unsigned long long)s->nbytes, (unsigned long long)nbytes);
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    if (nbytes) CU(cudaMemcpyAsync(s->payload(), src_host, nbytes, cudaMemcpyHostToDevice, st));
    g.st_h2d += nbytes;
    return scan_publish(*s, d, st);
}

int lzyb_slot_pack_host_part(lzyb_handle_t h, uint64_t offset, const void* src_host, uint64_t len, int last, void* stream) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->is_input || s->remote) return fail(LZYB_E_STATE, "slot %s is not a local output slot", s->id.c_str());
    if (s->state != LZYB_ST_EMPTY || s->digest_pending) return fail(LZYB_E_STATE, "slot %s already holds data", s->id.c_str());
    if (offset & 15) return fail(LZYB_E_ARG, "part offset must be a multiple of 16");
    if (offset != s->delivered) return fail(LZYB_E_ARG, "part offset %llu != bytes written %llu", (unsigned long long)offset,
                                            (unsigned long long)s->delivered);
    if (len && !src_host) return fail(LZYB_E_ARG, "null source");
    const uint64_t padded = round_up(len, 16);
    const uint64_t end = last ? s->nbytes : offset + padded;
    if (offset + len > s->nbytes || end > s->nbytes || (last && offset + padded < s->nbytes))
        return fail(LZYB_E_ARG, "part [%llu, +%llu) does not fit the %llu-byte payload", (unsigned long long)offset,
                    (unsigned long long)len, (unsigned long long)s->nbytes);
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    if (len) CU(cudaMemcpyAsync(s->payload() + offset, src_host, len, cudaMemcpyHostToDevice, st));
    if (end > offset + len) CU(cudaMemsetAsync(s->payload() + offset + len, 0, end - (offset + len), st));
    g.st_h2d += len;
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = s->payload() + offset;
    p.nbytes = end - offset;
    p.word_base = offset >> 3;
    p.ctl = s->ctl;
    p.total_nbytes = s->nbytes;
    p.fin = FIN_NONE;
    if (last) {
        p.fin = FIN_PUBLISH;
        p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(s->buf);
        fill_template(*s, &p.hdr_template);
        arm_mailbox(*s, p);
    }
    int rc = launch_stream(d, p, false, st);
    if (rc != LZYB_OK) return rc;
    g.st_scan++;
    s->delivered = end;
    s->ctl_dirty = !last;
    if (last) s->digest_pending = true;
    note(*s, st);
    return LZYB_OK;
}

int lzyb_slot_commit(lzyb_handle_t h, uint64_t digest_out[2], void* stream) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == LZYB_ST_CLOSED || s->state == LZYB_ST_FAILED) return fail(LZYB_E_STATE, "slot %s is closed/failed", s->id.c_str());
    if (s->state != LZYB_ST_READY) {
        if (s->is_input) return fail(LZYB_E_STATE, "slot %s is an input slot; finish the transfer instead", s->id.c_str());
        DeviceCtx& d = *g.devs[s->dev];
        DeviceGuard dg(d.ordinal);
        if (!s->digest_pending) {  // payload was written in place through lzyb_slot_device_ptr
            int rc = scan_publish(*s, d, pick_stream(d, stream));
            if (rc != LZYB_OK) return rc;
        }
        if (!digest_out && g.lazy_commit.load()) {
            // complete in stream order: everything that reads the frame is queued behind the pack (same-GPU hand-overs,
            // op bodies on the worker's stream, stream-waits of copies and transfers); the digest is harvested by whoever
            // asks for it first (ensure_digest).  A pack cannot fail its own digest, so READY is not a guess.
            s->state = LZYB_ST_READY;
            s->delivered = s->nbytes;
            return LZYB_OK;
        }
        int st = harvest(*s);
        if (st < 0) return st;
        if (st != 1) return fail(LZYB_E_LOCAL, "slot %s: unexpected finalize status %d", s->id.c_str(), st);
        s->state = LZYB_ST_READY;
        s->delivered = s->nbytes;
    }
    if (digest_out) {
        int rc = ensure_digest(*s);
        if (rc != LZYB_OK) return rc;
        digest_out[0] = s->digest[0];
        digest_out[1] = s->digest[1];
    }
    return LZYB_OK;
}

int64_t lzyb_slot_read(lzyb_handle_t h, uint64_t offset, void* dst_host, uint64_t len) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != LZYB_ST_READY) return fail(LZYB_E_STATE, "slot %s is not ready", s->id.c_str());
    if (s->is_input) return fail(LZYB_E_UNIMPLEMENTED, "unexpected `read` call on input slot %s", s->id.c_str());
    if (offset >= s->nbytes) return LZYB_EOF;
    uint64_t n = s->nbytes - offset;
    if (n > len) n = len;
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    if (s->last_ev && s->last_stream != d.stream) CU(cudaStreamWaitEvent(d.stream, s->last_ev, 0));  // the frame's last writer
    CU(cudaMemcpyAsync(dst_host, s->payload() + offset, n, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    g.st_d2h += n;
    return (int64_t)n;
}

int lzyb_copy_to_host(lzyb_handle_t h, uint64_t offset, void* dst_host, uint64_t len, void* stream) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == LZYB_ST_CLOSED) return fail(LZYB_E_STATE, "slot %s is closed", s->id.c_str());
    if (offset + len > s->nbytes) return fail(LZYB_E_ARG, "range beyond payload");
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    if (s->last_ev && s->last_stream != st) CU(cudaStreamWaitEvent(st, s->last_ev, 0));  // the frame's last writer
    bool bounced = false;
    if (len && len <= BOUNCE_BYTES) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, dst_host) != cudaSuccess) cudaGetLastError();
        else if (attr.type == cudaMemoryTypeUnregistered) {
            std::lock_guard<std::mutex> bl(d.bounce_mu);
            if (!d.bounce && cudaHostAlloc((void**)&d.bounce, BOUNCE_BYTES, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                d.bounce = nullptr;
            }
            if (d.bounce) {
                CU(cudaMemcpyAsync(d.bounce, s->payload() + offset, len, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                memcpy(dst_host, d.bounce, len);
                bounced = true;
            }
        }
    }
    if (len && !bounced) CU(cudaMemcpyAsync(dst_host, s->payload() + offset, len, cudaMemcpyDeviceToHost, st));
    if (!stream && !bounced) CU(cudaStreamSynchronize(st));
    g.st_d2h += len;
    return LZYB_OK;
}

int lzyb_slot_info(lzyb_handle_t h, lzyb_slot_info_t* out) {
    auto s = find(h);
    if (!s || !out) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (int rc = ensure_digest(*s); rc != LZYB_OK) return rc;
    memset(out, 0, sizeof(*out));
    out->nbytes = s->nbytes;
    out->gpu = g.devs[s->dev]->ordinal;
    out->state = s->state;
    out->is_input = s->is_input;
    out->is_remote = s->remote;
    out->delivered = s->delivered;
    out->digest[0] = s->digest[0];
    out->digest[1] = s->digest[1];
    out->dtype = s->dtype;
    out->ndim = s->ndim;
    memcpy(out->shape, s->shape, sizeof(out->shape));
    return LZYB_OK;
}

int lzyb_slot_meta(lzyb_handle_t h, lzyb_slot_info_t* out) {
    auto s = find(h);
    if (!s || !out) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    memset(out, 0, sizeof(*out));
    out->nbytes = s->nbytes;
    out->gpu = g.devs[s->dev]->ordinal;
    out->state = s->state;
    out->is_input = s->is_input;
    out->is_remote = s->remote;
    out->delivered = s->delivered;
    if (s->digest_ready) {
        out->digest[0] = s->digest[0];
        out->digest[1] = s->digest[1];
    }
    out->dtype = s->dtype;
    out->ndim = s->ndim;
    memcpy(out->shape, s->shape, sizeof(out->shape));
    return LZYB_OK;
}

int lzyb_slot_wait(lzyb_handle_t h, void* stream) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->last_ev) return LZYB_OK;
    DeviceGuard dg(g.devs[s->dev]->ordinal);
    if (!stream) {
        CU(cudaEventSynchronize(s->last_ev));
    } else if (s->last_stream != static_cast<cudaStream_t>(stream)) {
        CU(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), s->last_ev, 0));
    }
    return LZYB_OK;
}

int lzyb_host_alloc(uint64_t nbytes, void** out) {
    if (!out || !nbytes) return fail(LZYB_E_ARG, "null argument");
    if (!g.inited) return fail(LZYB_E_STATE, "lzyb_init has not been called (or found no GPU)");
    int c = 0;
    while (c < HOST_CLASSES && (1ull << (HOST_CLASS_MIN_LOG2 + c)) < nbytes) c++;
    if (c == HOST_CLASSES) return fail(LZYB_E_LOCAL, "no pinned block class for %llu bytes", (unsigned long long)nbytes);
    uint64_t size = 1ull << (HOST_CLASS_MIN_LOG2 + c);
    std::lock_guard<std::mutex> lk(g.host_mu);
    void* p = nullptr;
    if (!g.host_free[c].empty()) {
        p = g.host_free[c].back();
        g.host_free[c].pop_back();
    } else {
        if (g.host_bytes + size > g.host_limit)
            return fail(LZYB_E_LOCAL, "pinned host pool limit reached (%llu bytes in use)", (unsigned long long)g.host_bytes);
        if (cudaHostAlloc(&p, size, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return fail(LZYB_E_LOCAL, "cudaHostAlloc of %llu bytes failed", (unsigned long long)size);
        }
        g.host_bytes += size;
    }
    g.host_out[p] = c;
    *out = p;
    return LZYB_OK;
}

int lzyb_host_free(void* block) {
    std::lock_guard<std::mutex> lk(g.host_mu);
    auto it = g.host_out.find(block);
    if (it == g.host_out.end()) return fail(LZYB_E_ARG, "not an outstanding pinned block");
    int c = it->second;
    g.host_out.erase(it);
    if (g.inited) {
        g.host_free[c].push_back(block);
    } else {  // the library was shut down while the caller still held the block
        cudaFreeHost(block);
        g.host_bytes -= 1ull << (HOST_CLASS_MIN_LOG2 + c);
    }
    return LZYB_OK;
}

int lzyb_set_host_pool_limit(uint64_t bytes) {
    std::lock_guard<std::mutex> lk(g.host_mu);
    g.host_limit = bytes;
    return LZYB_OK;
}

int lzyb_set_transfer_waves(int waves) {
    if (waves < 1 || waves > 64) return fail(LZYB_E_ARG, "waves must be in [1, 64]");
    g.tma_waves = waves;
    return LZYB_OK;
}

int lzyb_set_lazy_commit(int on) {
    g.lazy_commit = on ? 1 : 0;
    return LZYB_OK;
}

int lzyb_slot_header(lzyb_handle_t h, lzyb_frame_header_t* out) {
    auto s = find(h);
    if (!s || !out) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == LZYB_ST_CLOSED) return fail(LZYB_E_STATE, "slot %s is closed", s->id.c_str());
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    if (s->last_ev) CU(cudaEventSynchronize(s->last_ev));
    CU(cudaMemcpy(out, s->buf, sizeof(*out), cudaMemcpyDeviceToHost));
    return LZYB_OK;
}

int lzyb_slot_retain(lzyb_handle_t h) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    s->refs++;
    return LZYB_OK;
}

int lzyb_slot_set_movable(lzyb_handle_t h, int movable) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->is_input || s->remote) return fail(LZYB_E_STATE, "slot %s is not a local output slot", s->id.c_str());
    s->movable = movable != 0;
    return LZYB_OK;
}

int lzyb_slot_close(lzyb_handle_t h) {
    SlotPtr s;
    {
        std::unique_lock<std::shared_mutex> lk(g.reg_mu);
        auto it = g.slots.find(h);
        if (it == g.slots.end()) return LZYB_OK;  // idempotent (InputSlot.close, InputSlot.java:110-113)
        s = it->second;
        if (--s->refs > 0) return LZYB_OK;        // another holder (the HBM frame store) keeps the frame
        g.slots.erase(it);
        auto bi = g.by_id.find(s->id);
        if (bi != g.by_id.end() && bi->second == h) g.by_id.erase(bi);
    }
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    if (s->last_ev) {
        cudaEventSynchronize(s->last_ev);  // cancel = let in-flight work drain before the frame is reused
        cudaEventDestroy(s->last_ev);
        s->last_ev = nullptr;
    }
    for (auto& r : s->readers) {  // pulls other slots of this process queued against this frame
        cudaEventSynchronize(r.second);
        cudaEventDestroy(r.second);
    }
    s->readers.clear();
    if (s->remote) {
        // the mapping stays in g.ipc_cache (closed at shutdown)
    } else if (s->buf && s->quarantined) {
        g.st_hbm_live -= s->cap;  // deliberately not pooled and not freed: a lost push may still write here
    } else if (s->buf) {
        free_frame(d, s->buf, s->cap);
        g.st_hbm_live -= s->cap;
    }
    if (s->ctl_idx >= 0) {
        if (s->ctl_dirty) cudaMemset(s->ctl, 0, sizeof(SlotCtl));
        std::lock_guard<std::mutex> dl(d.mu);
        d.ctl_free.push_back(s->ctl_idx);
    }
    if (s->mbox_idx >= 0) {
        std::lock_guard<std::mutex> ml(g.mbox_mu);
        g.mbox_free.push_back(s->mbox_idx);
    }
    s->buf = nullptr;
    s->state = LZYB_ST_CLOSED;
    g.st_live--;
    return LZYB_OK;
}

// ---- transfers -------------------------------------------------------------------------------

int64_t lzyb_transfer_chunk(lzyb_handle_t dst_h, lzyb_handle_t src_h, uint64_t offset, uint64_t max_len, void* stream) {
    auto dst = find(dst_h);
    auto src = find(src_h);
    if (!dst || !src) return fail(LZYB_E_ARG, "unknown slot handle");
    if (dst == src) return fail(LZYB_E_ARG, "source and sink are the same slot");
    std::scoped_lock lk(dst->mu, src->mu);
    if (!dst->is_input) return fail(LZYB_E_STATE, "slot %s is not an input slot", dst->id.c_str());
    if (dst->state == LZYB_ST_CLOSED || dst->state == LZYB_ST_FAILED) return fail(LZYB_E_LOCAL, "sink %s already closed", dst->id.c_str());
    if (dst->state == LZYB_ST_READY) return LZYB_EOF;
    if (src->state != LZYB_ST_READY) return fail(LZYB_E_PEER, "peer %s has no complete data", src->id.c_str());
    if (src->nbytes != dst->nbytes) return fail(LZYB_E_PEER, "peer %s holds %llu bytes, sink expects %llu", src->id.c_str(),
                                                (unsigned long long)src->nbytes, (unsigned long long)dst->nbytes);
    if (offset != dst->delivered) return fail(LZYB_E_ARG, "offset %llu != delivered %llu (transferChunkTo repositions the sink)",
                                             (unsigned long long)offset, (unsigned long long)dst->delivered);
    if (dst->digest_pending) return LZYB_EOF;  // final chunk already queued
    if (offset == dst->nbytes && dst->nbytes != 0) return LZYB_EOF;
    uint64_t len = dst->nbytes - offset;
    if (max_len < len) {
        len = max_len & ~15ull;
        if (len == 0) return fail(LZYB_E_ARG, "max_len %llu is below the 16-byte transfer granule", (unsigned long long)max_len);
    }
    // fault injection at the backend seam (FailingInputStream, SlotsTest.java:914-943)
    if (src->fault_kind == LZYB_FAULT_PEER_AFTER) {
        if (offset >= src->fault_arg) return fail(LZYB_E_PEER, "injected fault: Cannot read from slot peer %s", src->id.c_str());
        if (offset + len > src->fault_arg) len = (src->fault_arg - offset) & ~15ull;
        if (len == 0) return fail(LZYB_E_PEER, "injected fault: Cannot read from slot peer %s", src->id.c_str());
    }
    DeviceCtx& d = *g.devs[dst->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    const bool last = offset + len == dst->nbytes;
    // ---- move instead of copy (InputSlotBackend.toOutput's ownership hand-over, FileInputBackend.java:38-45, applied
    // to the producer→consumer edge): the producer's owner declared a sole consumer (lzyb_slot_set_movable), nobody else
    // holds the frame (refs == 1), both slots sit on one GPU and the consumer asks for the whole payload.  The two
    // slots swap frames: the consumer owns the committed frame (header and digest as the producer published them — the
    // bytes never left this GPU, there is nothing to re-verify), the producer keeps the consumer's empty one and reads
    // as MOVED from now on.  Any second holder (frame store, proxy, another consumer's retain) keeps the copy path.
    if (src->movable && src->refs.load() == 1 && !src->remote && !dst->remote && src->dev == dst->dev && offset == 0 &&
        len == dst->nbytes && dst->delivered == 0 && src->cap == dst->cap && src->fault_kind == LZYB_FAULT_NONE) {
        if (src->last_ev && src->last_stream != st) cudaStreamWaitEvent(st, src->last_ev, 0);
        for (auto& r : src->readers) cudaStreamWaitEvent(st, r.second, 0);  // earlier readers of the frame finish first
        std::swap(src->buf, dst->buf);
        dst->dtype = src->dtype;
        dst->ndim = src->ndim;
        memcpy(dst->shape, src->shape, sizeof(dst->shape));
        if (src->digest_ready) {
            dst->digest[0] = src->digest[0];
            dst->digest[1] = src->digest[1];
            dst->sum = src->sum;
            dst->xr = src->xr;
            dst->digest_pending = false;
            dst->digest_ready = true;
        } else {
            // the producer was committed in stream order (lzyb_set_lazy_commit) and its pack may still be running: the
            // pending finalize — mailbox, sequence number, the event of the launch — travels with the frame
            std::swap(src->mbox, dst->mbox);
            std::swap(src->mbox_idx, dst->mbox_idx);
            std::swap(src->seq, dst->seq);
            std::swap(src->last_ev, dst->last_ev);
            std::swap(src->last_stream, dst->last_stream);
            dst->digest_pending = true;
            dst->digest_ready = false;
            src->digest_pending = false;
        }
        dst->delivered = dst->nbytes;
        dst->state = LZYB_ST_READY;
        src->state = LZYB_ST_MOVED;
        src->movable = false;
        if (dst->digest_ready) note(*dst, st);  // (a pending finalize keeps the event of its own launch: harvest waits for that)
        g.st_moves++;
        g.st_bpull_l += len;
        return (int64_t)len;
    }
    if (int rc = ensure_digest(*src); rc != LZYB_OK) return rc;  // the copy verifies against the producer's digest
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = src->payload() + offset;
    p.dst = dst->payload() + offset;
    p.nbytes = len;
    p.word_base = offset >> 3;
    p.ctl = dst->ctl;
    p.total_nbytes = dst->nbytes;
    p.fin = FIN_NONE;
    if (last) {
        p.fin = FIN_VERIFY;
        p.expect[0] = src->digest[0];
        p.expect[1] = src->digest[1];
        p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(dst->buf);
        dst->dtype = src->dtype;
        dst->ndim = src->ndim;
        memcpy(dst->shape, src->shape, sizeof(dst->shape));
        fill_template(*dst, &p.hdr_template);
        arm_mailbox(*dst, p);
    }
    // the producer's payload must be final on this stream's timeline
    if (src->last_ev && src->last_stream != st) cudaStreamWaitEvent(st, src->last_ev, 0);
    const bool peer = src->remote || src->dev != dst->dev;
    const int engine = g.engine.load();
    // Who runs the kernel (measured, push_kernels.cuh): with ONE direction of a GPU pair busy a consumer pull is the
    // fastest SM carrier (770–780 GB/s: read responses cost 1.125 wire bytes per payload byte), but once the sink GPU is
    // itself serving a transfer the pull's requests share the wire with that payload (657–671 per direction) and a
    // producer push — posted writes, 1.19 wire bytes, nothing on the reverse wire — keeps 684–700.  auto therefore pushes
    // exactly when the sink GPU is sourcing a transfer right now (a ring, a bidirectional exchange) and pulls otherwise;
    // short transfers always pull (one launch on one GPU, no cross-device hand-shake).  A source mapped from another
    // process can only be pulled from here (its owner pushes through lzyb_slot_push).  The copy engine (3) is a named
    // alternative, never auto's choice: it holds 716–737 GB/s with larger packets than an SM can issue.
    const bool in_process_peer = peer && !src->remote && !dst->remote;
    const bool push = in_process_peer && (engine == 4 || (engine == 0 && len >= g.push_min.load() && is_sourcing(d)));
    const bool dma = !push && peer && engine == 3;
    int rc;
    if (push) {
        rc = launch_push_inproc(*g.devs[src->dev], d, *src, *dst, p, st);
    } else {
        rc = dma ? launch_dma_verify(d, p, st) : launch_stream(d, p, true, st, peer);
        if (rc == LZYB_OK && in_process_peer) mark_sourcing(*g.devs[src->dev], dst->dev, st);
    }
    if (rc != LZYB_OK) return rc;
    g.st_pull++;
    if (src->dev == dst->dev && !src->remote)
        g.st_bpull_l += len;
    else
        g.st_bpull_p += len;
    dst->delivered += len;
    dst->state = LZYB_ST_FILLING;
    dst->ctl_dirty = !last;
    if (last) dst->digest_pending = true;
    note(*dst, st);
    if (!stream) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(LZYB_E_LOCAL, "transfer failed: %s", cudaGetErrorString(e));
    } else if (!src->remote) {
        add_reader(*src, st);  // an asynchronous reader: the source frame must not be recycled under it
    }
    return (int64_t)len;
}

static int input_reset_locked(Slot& s) {
    DeviceCtx& d = *g.devs[s.dev];
    DeviceGuard dg(d.ordinal);
    if (s.last_ev) cudaEventSynchronize(s.last_ev);
    if (s.ctl_dirty) {
        CU(cudaMemset(s.ctl, 0, sizeof(SlotCtl)));
        s.ctl_dirty = false;
    }
    for (auto& r : s.readers) cudaEventSynchronize(r.second);  // a push into this frame that is still running finishes first
    s.delivered = 0;
    s.digest_pending = false;
    s.state = LZYB_ST_EMPTY;
    return LZYB_OK;
}

static int finish_common(Slot& dst, const uint64_t expect[2], const Slot* src, void* stream) {
    if (!dst.is_input) return fail(LZYB_E_STATE, "slot %s is not an input slot", dst.id.c_str());
    if (dst.state == LZYB_ST_READY) return g.lazy_commit.load() ? LZYB_OK : ensure_digest(dst);
    if (dst.state == LZYB_ST_CLOSED || dst.state == LZYB_ST_FAILED) return fail(LZYB_E_LOCAL, "sink %s already closed", dst.id.c_str());
    if (dst.delivered != dst.nbytes) return fail(LZYB_E_STATE, "slot %s: %llu of %llu bytes delivered", dst.id.c_str(),
                                                 (unsigned long long)dst.delivered, (unsigned long long)dst.nbytes);
    DeviceCtx& d = *g.devs[dst.dev];
    DeviceGuard dg(d.ordinal);
    if (!dst.digest_pending) {  // zero-byte payload or host-fed chunks: finalize on its own
        cudaStream_t st = pick_stream(d, stream);
        CopyParams p;
        memset(&p, 0, sizeof(p));
        p.ctl = dst.ctl;
        p.total_nbytes = dst.nbytes;
        p.fin = expect ? FIN_VERIFY : FIN_PUBLISH;
        if (expect) {
            p.expect[0] = expect[0];
            p.expect[1] = expect[1];
        }
        p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(dst.buf);
        if (src) {
            dst.dtype = src->dtype;
            dst.ndim = src->ndim;
            memcpy(dst.shape, src->shape, sizeof(dst.shape));
        }
        fill_template(dst, &p.hdr_template);
        arm_mailbox(dst, p);
        finalize_kernel<<<1, 32, 0, st>>>(p);
        g.st_launch++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(LZYB_E_LOCAL, "finalize launch failed: %s", cudaGetErrorString(e));
        dst.digest_pending = true;
        note(dst, st);
    }
    int st = harvest(dst);
    if (st < 0) return st;
    if (st == 3) {
        // the producer may still be alive and push late — into this very frame.  It is therefore neither reset for another
        // attempt nor returned to the pool: the slot fails, its frame is retired with it (a consumer that wants to retry
        // opens a fresh input slot).
        dst.state = LZYB_ST_FAILED;
        dst.quarantined = true;
        dst.digest_pending = false;
        return fail(LZYB_E_PEER, "slot %s: the producer's push did not arrive within the sink's timeout", dst.id.c_str());
    }
    const bool foreign = st == 1 && expect && (dst.digest[0] != expect[0] || dst.digest[1] != expect[1]);  // the verdict of another frame
    if (st == 2 || foreign) {
        const unsigned long long g0 = dst.digest[0], g1 = dst.digest[1];
        input_reset_locked(dst);
        return fail(LZYB_E_PEER, "digest mismatch on slot %s: delivered bytes differ from what the producer committed "
                                 "(%s; delivered %016llx%016llx, expected %016llx%016llx)", dst.id.c_str(),
                    foreign ? "verified, but against another digest" : "verdict of the digesting kernel", g0, g1,
                    expect ? (unsigned long long)expect[0] : 0ull, expect ? (unsigned long long)expect[1] : 0ull);
    }
    dst.state = LZYB_ST_READY;
    return LZYB_OK;
}

int lzyb_input_finish(lzyb_handle_t dst_h, lzyb_handle_t src_h, void* stream) {
    auto dst = find(dst_h);
    auto src = find(src_h);
    if (!dst || !src) return fail(LZYB_E_ARG, "unknown slot handle");
    std::scoped_lock lk(dst->mu, src->mu);
    if (dst->state != LZYB_ST_READY)
        if (int rc = ensure_digest(*src); rc != LZYB_OK) return rc;
    return finish_common(*dst, src->digest, src.get(), stream);
}

int lzyb_input_finish_expect(lzyb_handle_t dst_h, const uint64_t digest[2], void* stream) {
    auto dst = find(dst_h);
    if (!dst) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(dst->mu);
    return finish_common(*dst, digest, nullptr, stream);  // digest == NULL: publish what was computed
}

int lzyb_input_adopt_external(lzyb_handle_t dst_h, const uint64_t digest[2], uint32_t dtype, uint32_t ndim,
                              const uint64_t* shape, void* stream) {
    auto dst = find(dst_h);
    if (!dst) return fail(LZYB_E_ARG, "unknown slot handle");
    if (ndim > LZYB_MAX_DIMS) return fail(LZYB_E_ARG, "ndim %u > %d", ndim, LZYB_MAX_DIMS);
    std::lock_guard<std::mutex> lk(dst->mu);
    if (!dst->is_input) return fail(LZYB_E_STATE, "slot %s is not an input slot", dst->id.c_str());
    if (dst->state != LZYB_ST_EMPTY || dst->delivered != 0) return fail(LZYB_E_STATE, "slot %s already holds data", dst->id.c_str());
    DeviceCtx& d = *g.devs[dst->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    dst->dtype = dtype;
    dst->ndim = ndim;
    memset(dst->shape, 0, sizeof(dst->shape));
    for (uint32_t i = 0; i < ndim && shape; ++i) dst->shape[i] = shape[i];
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = dst->payload();
    p.nbytes = dst->nbytes;
    p.ctl = dst->ctl;
    p.total_nbytes = dst->nbytes;
    p.fin = digest ? FIN_VERIFY : FIN_PUBLISH;
    if (digest) {
        p.expect[0] = digest[0];
        p.expect[1] = digest[1];
    }
    p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(dst->buf);
    fill_template(*dst, &p.hdr_template);
    arm_mailbox(*dst, p);
    int rc = launch_stream(d, p, false, st);
    if (rc != LZYB_OK) return rc;
    g.st_scan++;
    dst->delivered = dst->nbytes;
    dst->digest_pending = true;
    note(*dst, st);
    int status = harvest(*dst);
    if (status < 0) return status;
    if (status == 2) {
        input_reset_locked(*dst);
        return fail(LZYB_E_PEER, "digest mismatch on slot %s after an external fill", dst->id.c_str());
    }
    dst->state = LZYB_ST_READY;
    return LZYB_OK;
}

int64_t lzyb_input_write_host(lzyb_handle_t dst_h, uint64_t offset, const void* src_host, uint64_t len, void* stream) {
    auto dst = find(dst_h);
    if (!dst) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(dst->mu);
    if (!dst->is_input) return fail(LZYB_E_STATE, "slot %s is not an input slot", dst->id.c_str());
    if (dst->state == LZYB_ST_CLOSED || dst->state == LZYB_ST_FAILED) return fail(LZYB_E_LOCAL, "sink %s already closed", dst->id.c_str());
    if (offset != dst->delivered) return fail(LZYB_E_ARG, "offset %llu != delivered %llu", (unsigned long long)offset,
                                             (unsigned long long)dst->delivered);
    if (offset == dst->nbytes) return LZYB_EOF;
    if (offset + len > dst->nbytes) len = dst->nbytes - offset;
    if (offset + len != dst->nbytes && (len & 15)) return fail(LZYB_E_ARG, "intermediate chunks must be multiples of 16 bytes");
    DeviceCtx& d = *g.devs[dst->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    CU(cudaMemcpyAsync(dst->payload() + offset, src_host, len, cudaMemcpyHostToDevice, st));
    g.st_h2d += len;
    // digest the landed bytes in place (scan), accumulating into the slot's partials
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = dst->payload() + offset;
    p.nbytes = len;
    p.word_base = offset >> 3;
    p.ctl = dst->ctl;
    p.total_nbytes = dst->nbytes;
    p.fin = FIN_NONE;
    int rc = launch_stream(d, p, false, st);
    if (rc != LZYB_OK) return rc;
    g.st_scan++;
    dst->delivered += len;
    dst->state = LZYB_ST_FILLING;
    dst->ctl_dirty = true;
    note(*dst, st);
    if (!stream) CU(cudaStreamSynchronize(st));
    return (int64_t)len;
}

int lzyb_input_to_output(lzyb_handle_t h) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == LZYB_ST_CLOSED) return fail(LZYB_E_STATE, "Already closed");
    if (!s->is_input) return LZYB_OK;
    if (s->state != LZYB_ST_READY) return fail(LZYB_E_STATE, "slot %s is not ready", s->id.c_str());
    s->is_input = false;
    return LZYB_OK;
}

int lzyb_input_reset(lzyb_handle_t h) {
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->is_input) return fail(LZYB_E_STATE, "slot %s is not an input slot", s->id.c_str());
    return input_reset_locked(*s);
}

// ---- CUDA IPC peers --------------------------------------------------------------------------

int lzyb_slot_export(lzyb_handle_t h, lzyb_ipc_desc_t* out) {
    auto s = find(h);
    if (!s || !out) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != LZYB_ST_READY) return fail(LZYB_E_STATE, "slot %s is not committed", s->id.c_str());
    if (s->remote) return fail(LZYB_E_STATE, "slot %s is itself an import", s->id.c_str());
    if (int rc = ensure_digest(*s); rc != LZYB_OK) return rc;
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    memset(out, 0, sizeof(*out));
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, s->buf));
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out->ipc_handle, &mh, 64);
    out->buffer_bytes = s->cap;
    out->nbytes = s->nbytes;
    out->gpu = d.ordinal;
    out->pid = (int32_t)getpid();
    out->digest[0] = s->digest[0];
    out->digest[1] = s->digest[1];
    out->sum = s->sum;
    out->xr = s->xr;
    out->dtype = s->dtype;
    out->ndim = s->ndim;
    memcpy(out->shape, s->shape, sizeof(out->shape));
    return LZYB_OK;
}

int lzyb_peer_import(const lzyb_ipc_desc_t* desc, int onto_gpu, lzyb_handle_t* out) {
    if (!g.inited) return fail(LZYB_E_STATE, "lzyb_init has not been called");
    if (!desc || !out) return fail(LZYB_E_ARG, "null argument");
    int di = onto_gpu - g.first;
    if (di < 0 || di >= (int)g.devs.size()) return fail(LZYB_E_ARG, "GPU %d is not managed by this process", onto_gpu);
    if (desc->pid == (int32_t)getpid()) return fail(LZYB_E_ARG, "descriptor comes from this process; use the slot handle");
    // the descriptor crossed a process boundary: nothing in it is trusted to size an access
    if (desc->ndim > LZYB_MAX_DIMS) return fail(LZYB_E_PEER, "peer descriptor: ndim %u > %d", desc->ndim, LZYB_MAX_DIMS);
    if (desc->buffer_bytes < HEADER_BYTES || desc->nbytes > desc->buffer_bytes - HEADER_BYTES)
        return fail(LZYB_E_PEER, "peer descriptor: payload of %llu bytes does not fit its %llu-byte frame",
                    (unsigned long long)desc->nbytes, (unsigned long long)desc->buffer_bytes);
    DeviceCtx& d = *g.devs[di];
    DeviceGuard dg(d.ordinal);
    cudaIpcMemHandle_t mh;
    memcpy(&mh, desc->ipc_handle, 64);
    void* base = nullptr;
    {
        std::lock_guard<std::mutex> lk(g.ipc_mu);
        auto key = std::make_pair(d.ordinal, std::string(reinterpret_cast<const char*>(desc->ipc_handle), 64));
        auto it = g.ipc_cache.find(key);
        if (it != g.ipc_cache.end()) {
            base = it->second;
        } else {
            cudaError_t e = cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(LZYB_E_PEER, "cannot map peer frame (pid %d gpu %d): %s", desc->pid, desc->gpu, cudaGetErrorString(e));
            }
            g.ipc_cache.emplace(key, base);
        }
    }
    auto s = std::make_shared<Slot>();
    char name[64];
    snprintf(name, sizeof(name), "ipc:%d:%d:%p", desc->pid, desc->gpu, base);
    s->id = name;
    s->dev = di;
    s->nbytes = desc->nbytes;
    s->cap = desc->buffer_bytes;
    s->buf = static_cast<uint8_t*>(base);
    s->ipc_base = base;
    s->remote = true;
    s->state = LZYB_ST_READY;
    s->digest_ready = true;
    s->delivered = desc->nbytes;
    s->digest[0] = desc->digest[0];
    s->digest[1] = desc->digest[1];
    s->sum = desc->sum;
    s->xr = desc->xr;
    s->dtype = desc->dtype;
    s->ndim = desc->ndim;
    memcpy(s->shape, desc->shape, sizeof(s->shape));
    s->handle = g.next_handle++;
    {
        std::unique_lock<std::shared_mutex> lk(g.reg_mu);
        g.slots[s->handle] = s;
    }
    g.st_live++;
    *out = s->handle;
    return LZYB_OK;
}

// ---- producer-issued push between processes ---------------------------------------------------------------------

int lzyb_sink_export(lzyb_handle_t h, lzyb_ipc_desc_t* out) {
    auto s = find(h);
    if (!s || !out) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->is_input || s->remote) return fail(LZYB_E_STATE, "slot %s is not a local input slot", s->id.c_str());
    if (s->state != LZYB_ST_EMPTY || s->delivered != 0 || s->digest_pending)
        return fail(LZYB_E_STATE, "slot %s cannot receive a push (state %d, %llu bytes delivered)", s->id.c_str(), s->state,
                    (unsigned long long)s->delivered);
    if (!has_trailer(s->cap)) return fail(LZYB_E_ARG, "slot %s is too small to be pushed into (no trailer): pull it", s->id.c_str());
    DeviceCtx& d = *g.devs[s->dev];
    DeviceGuard dg(d.ordinal);
    memset(out, 0, sizeof(*out));
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, s->buf));
    memcpy(out->ipc_handle, &mh, 64);
    out->buffer_bytes = s->cap;
    out->nbytes = s->nbytes;
    out->gpu = d.ordinal;
    out->pid = (int32_t)getpid();
    s->push_seq = g.next_push_seq++ | ((uint64_t)(uint32_t)getpid() << 40);  // never 0, never repeated for this frame
    memcpy(out->reserved, &s->push_seq, sizeof(uint64_t));  // reserved[0..8): the arrival word the producer must write
    return LZYB_OK;
}

static int map_peer_frame(DeviceCtx& d, const lzyb_ipc_desc_t* desc, void** base) {
    cudaIpcMemHandle_t mh;
    memcpy(&mh, desc->ipc_handle, 64);
    std::lock_guard<std::mutex> lk(g.ipc_mu);
    auto key = std::make_pair(d.ordinal, std::string(reinterpret_cast<const char*>(desc->ipc_handle), 64));
    auto it = g.ipc_cache.find(key);
    if (it != g.ipc_cache.end()) {
        *base = it->second;
        return LZYB_OK;
    }
    cudaError_t e = cudaIpcOpenMemHandle(base, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LZYB_E_PEER, "cannot map peer frame (pid %d gpu %d): %s", desc->pid, desc->gpu, cudaGetErrorString(e));
    }
    g.ipc_cache.emplace(key, *base);
    return LZYB_OK;
}

int lzyb_slot_push(lzyb_handle_t src_h, const lzyb_ipc_desc_t* sink, uint64_t offset, uint64_t len, void* stream) {
    auto src = find(src_h);
    if (!src || !sink) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(src->mu);
    if (src->state != LZYB_ST_READY || src->remote) return fail(LZYB_E_STATE, "slot %s is not a committed local slot", src->id.c_str());
    if (int rc = ensure_digest(*src); rc != LZYB_OK) return rc;
    if (sink->pid == (int32_t)getpid()) return fail(LZYB_E_ARG, "sink belongs to this process; use lzyb_transfer_chunk");
    if (sink->nbytes != src->nbytes) return fail(LZYB_E_PEER, "sink expects %llu bytes, slot %s holds %llu", (unsigned long long)sink->nbytes,
                                                  src->id.c_str(), (unsigned long long)src->nbytes);
    if (!has_trailer(sink->buffer_bytes) || sink->nbytes > sink->buffer_bytes - HEADER_BYTES - PUSH_TRAILER_BYTES)
        return fail(LZYB_E_PEER, "sink descriptor: %llu payload bytes do not fit its %llu-byte frame", (unsigned long long)sink->nbytes,
                    (unsigned long long)sink->buffer_bytes);
    // a push carries the whole payload: the verdict is the digest of everything that went over the wire against the
    // digest this slot committed.  (Partial, offset-resumed deliveries between processes take the pull path.)
    if (offset != 0 || (len != 0 && len != src->nbytes)) return fail(LZYB_E_ARG, "a push carries the whole payload (offset 0, len 0 or nbytes)");
    uint64_t seq = 0;
    memcpy(&seq, sink->reserved, sizeof(uint64_t));
    if (seq == 0) return fail(LZYB_E_PEER, "sink descriptor carries no arrival word (not produced by lzyb_sink_export)");
    DeviceCtx& d = *g.devs[src->dev];
    DeviceGuard dg(d.ordinal);
    void* base = nullptr;
    int rc = map_peer_frame(d, sink, &base);
    if (rc != LZYB_OK) return rc;
    cudaStream_t st = pick_stream(d, stream);
    if (src->last_ev && src->last_stream != st) CU(cudaStreamWaitEvent(st, src->last_ev, 0));
    // the slot's own accumulator block (zero at rest since its commit) collects this push's partials: pushes of one slot
    // take turns on the device
    for (auto& r : src->readers)
        if (r.first != st) CU(cudaStreamWaitEvent(st, r.second, 0));
    uint8_t* sb = static_cast<uint8_t*>(base);
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = src->payload();
    p.dst = sb + HEADER_BYTES;
    p.nbytes = src->nbytes;
    p.ctl = src->ctl;
    p.total_nbytes = src->nbytes;
    p.fin = FIN_VERIFY;
    p.expect[0] = src->digest[0];
    p.expect[1] = src->digest[1];
    p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(sb);
    fill_template(*src, &p.hdr_template);
    p.sys_fence = 1;
    p.arrive = frame_trailer(sb, sink->buffer_bytes);
    p.arrive_seq = seq;
    rc = launch_stream(d, p, true, st, true);
    if (rc != LZYB_OK) return rc;
    g.st_push++;
    add_reader(*src, st);
    g.st_bpull_p += src->nbytes;
    if (!stream) CU(cudaStreamSynchronize(st));
    return LZYB_OK;
}

int64_t lzyb_input_expect_push(lzyb_handle_t dst_h, const lzyb_ipc_desc_t* producer, uint64_t offset, uint64_t len, void* stream) {
    auto dst = find(dst_h);
    if (!dst || !producer) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(dst->mu);
    if (!dst->is_input || dst->remote) return fail(LZYB_E_STATE, "slot %s is not a local input slot", dst->id.c_str());
    if (dst->state == LZYB_ST_CLOSED || dst->state == LZYB_ST_FAILED) return fail(LZYB_E_LOCAL, "sink %s already closed", dst->id.c_str());
    if (dst->state == LZYB_ST_READY || dst->digest_pending) return LZYB_EOF;
    if (dst->push_seq == 0) return fail(LZYB_E_STATE, "slot %s was not exported as a push sink (lzyb_sink_export)", dst->id.c_str());
    if (producer->nbytes != dst->nbytes) return fail(LZYB_E_PEER, "peer holds %llu bytes, sink expects %llu", (unsigned long long)producer->nbytes,
                                                      (unsigned long long)dst->nbytes);
    if (producer->ndim > LZYB_MAX_DIMS) return fail(LZYB_E_PEER, "peer descriptor: ndim %u > %d", producer->ndim, LZYB_MAX_DIMS);
    if (offset != 0 || (len != 0 && len != dst->nbytes)) return fail(LZYB_E_ARG, "a push carries the whole payload (offset 0, len 0 or nbytes)");
    DeviceCtx& d = *g.devs[dst->dev];
    DeviceGuard dg(d.ordinal);
    cudaStream_t st = pick_stream(d, stream);
    dst->dtype = producer->dtype;
    dst->ndim = producer->ndim;
    memcpy(dst->shape, producer->shape, sizeof(dst->shape));
    dst->seq++;
    dst->mbox->status = 0;
    dst->mbox->seq = 0;
    ArrivalWait w;
    memset(&w, 0, sizeof(w));
    w.trailer = frame_trailer(dst->buf, dst->cap);
    w.seq = dst->push_seq;
    w.mbox = dst->mbox;
    w.mbox_seq = dst->seq;
    arrival_wait_kernel<<<1, 32, 0, st>>>(w);
    g.st_launch++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LZYB_E_LOCAL, "arrival-wait launch failed: %s", cudaGetErrorString(e));
    g.st_pull++;
    dst->delivered = dst->nbytes;
    dst->state = LZYB_ST_FILLING;
    dst->ctl_dirty = false;
    dst->digest_pending = true;
    note(*dst, st);
    return (int64_t)dst->nbytes;
}

// ---- fan-out: chunk-pipelined pull chain -------------------------------------------------------

static int broadcast_host_driven(lzyb_handle_t src_h, const lzyb_handle_t* dsts, int n, uint64_t chunk_bytes) {
    if (n <= 0 || !dsts) return fail(LZYB_E_ARG, "no consumers");
    if (chunk_bytes == 0) chunk_bytes = 32ull << 20;
    chunk_bytes = round_up(chunk_bytes, 4096);
    auto src = find(src_h);
    if (!src) return fail(LZYB_E_ARG, "unknown source handle");
    {
        std::lock_guard<std::mutex> lk(src->mu);
        if (int rc = ensure_digest(*src); rc != LZYB_OK) return rc;
    }
    const uint64_t total = src->nbytes;
    std::vector<SlotPtr> chain;
    for (int k = 0; k < n; ++k) {
        auto s = find(dsts[k]);
        if (!s) return fail(LZYB_E_ARG, "unknown consumer handle #%d", k);
        chain.push_back(s);
    }
    // stage k pulls chunk c from stage k-1 as soon as that chunk has landed there: every launch
    // goes to its own GPU's library stream and waits on the event of the upstream launch.
    const uint64_t nchunks = total ? (total + chunk_bytes - 1) / chunk_bytes : 0;
    std::vector<std::vector<cudaEvent_t>> ev(n, std::vector<cudaEvent_t>(nchunks, nullptr));
    int rc = LZYB_OK;
    for (uint64_t c = 0; c < nchunks && rc == LZYB_OK; ++c) {
        const uint64_t off = c * chunk_bytes;
        const uint64_t len = (off + chunk_bytes > total) ? total - off : chunk_bytes;
        for (int k = 0; k < n; ++k) {
            DeviceCtx& d = *g.devs[chain[k]->dev];
            DeviceGuard dg(d.ordinal);
            if (k > 0) cudaStreamWaitEvent(d.stream, ev[k - 1][c], 0);
            lzyb_handle_t from = k == 0 ? src_h : dsts[k - 1];
            // upstream stage is mid-download: pull from its frame directly (same bytes, same offsets)
            int64_t moved;
            if (k == 0) {
                moved = lzyb_transfer_chunk(dsts[k], from, off, len, d.stream);
            } else {
                // temporarily present the upstream input slot as a readable peer
                std::scoped_lock lk(chain[k]->mu, chain[k - 1]->mu);
                Slot& dd = *chain[k];
                Slot& up = *chain[k - 1];
                CopyParams p;
                memset(&p, 0, sizeof(p));
                p.src = up.payload() + off;
                p.dst = dd.payload() + off;
                p.nbytes = len;
                p.word_base = off >> 3;
                p.ctl = dd.ctl;
                p.total_nbytes = total;
                p.fin = FIN_NONE;
                const bool last = off + len == total;
                if (last) {
                    p.fin = FIN_VERIFY;
                    p.expect[0] = src->digest[0];
                    p.expect[1] = src->digest[1];
                    p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(dd.buf);
                    dd.dtype = src->dtype;
                    dd.ndim = src->ndim;
                    memcpy(dd.shape, src->shape, sizeof(dd.shape));
                    fill_template(dd, &p.hdr_template);
                    arm_mailbox(dd, p);
                }
                int lr = launch_stream(d, p, true, d.stream, up.dev != dd.dev);
                if (lr != LZYB_OK) {
                    rc = lr;
                    break;
                }
                g.st_pull++;
                g.st_bpull_p += len;
                dd.delivered += len;
                dd.state = LZYB_ST_FILLING;
                dd.ctl_dirty = !last;
                if (last) dd.digest_pending = true;
                note(dd, d.stream);
                moved = (int64_t)len;
            }
            if (moved < 0) {
                rc = (int)moved;
                break;
            }
            cudaEventCreateWithFlags(&ev[k][c], cudaEventDisableTiming);
            cudaEventRecord(ev[k][c], d.stream);
        }
    }
    if (rc == LZYB_OK) {
        for (int k = 0; k < n && rc == LZYB_OK; ++k) rc = lzyb_input_finish(dsts[k], src_h, nullptr);
    }
    for (auto& row : ev)
        for (auto e : row)
            if (e) cudaEventDestroy(e);
    return rc;
}


// Device-chained fan-out: one chain_digest_kernel launch per consumer GPU, all running at once; the
// stages hand tiles to each other through progress counters in peer memory.
static int broadcast_device_chained(lzyb_handle_t src_h, const lzyb_handle_t* dsts, int n) {
    auto src = find(src_h);
    if (!src) return fail(LZYB_E_ARG, "unknown source handle");
    std::vector<SlotPtr> chain;
    for (int k = 0; k < n; ++k) {
        auto s = find(dsts[k]);
        if (!s) return fail(LZYB_E_ARG, "unknown consumer handle #%d", k);
        chain.push_back(s);
    }
    const uint64_t total = src->nbytes;
    if (src->state != LZYB_ST_READY) return fail(LZYB_E_PEER, "peer %s has no complete data", src->id.c_str());
    {
        std::lock_guard<std::mutex> lk(src->mu);
        if (int rc = ensure_digest(*src); rc != LZYB_OK) return rc;
    }
    for (auto& s : chain) {
        if (!s->is_input || s->state != LZYB_ST_EMPTY || s->delivered != 0 || s->nbytes != total)
            return fail(LZYB_E_STATE, "consumer %s is not an empty input slot of %llu bytes", s->id.c_str(), (unsigned long long)total);
    }
    constexpr unsigned CHUNK_SHIFT = 2;  // 4 spans of 256 KiB = 1 MiB per progress counter
    const uint64_t nvec = total >> 4, span = 256ull * CHAIN_UNROLL * CHAIN_SPAN_TILES;
    uint64_t ntiles = (nvec + span - 1) / span;  // spans
    if (ntiles == 0) ntiles = 1;
    const uint64_t nchunks = (ntiles + (1ull << CHUNK_SHIFT) - 1) >> CHUNK_SHIFT;
    const uint64_t flag_bytes = frame_capacity(nchunks * sizeof(unsigned int));
    std::vector<uint8_t*> flags(n, nullptr);
    int rc = LZYB_OK;
    for (int k = 0; k < n && rc == LZYB_OK; ++k) {
        DeviceCtx& d = *g.devs[chain[k]->dev];
        DeviceGuard dg(d.ordinal);
        rc = alloc_frame(d, flag_bytes, &flags[k]);
        if (rc == LZYB_OK && cudaMemsetAsync(flags[k], 0, nchunks * sizeof(unsigned int), d.stream) != cudaSuccess)
            rc = fail(LZYB_E_LOCAL, "cannot clear progress counters");
    }
    for (int k = 0; k < n && rc == LZYB_OK; ++k) {  // counters are zero everywhere before anybody polls
        DeviceCtx& d = *g.devs[chain[k]->dev];
        DeviceGuard dg(d.ordinal);
        if (cudaStreamSynchronize(d.stream) != cudaSuccess) rc = fail(LZYB_E_LOCAL, "stream sync failed");
    }
    for (int k = 0; k < n && rc == LZYB_OK; ++k) {
        Slot& dd = *chain[k];
        Slot& up = k == 0 ? *src : *chain[k - 1];
        std::lock_guard<std::mutex> lk(dd.mu);
        DeviceCtx& d = *g.devs[dd.dev];
        DeviceGuard dg(d.ordinal);
        CopyParams p;
        memset(&p, 0, sizeof(p));
        p.src = up.payload();
        p.dst = dd.payload();
        p.nbytes = total;
        p.ctl = dd.ctl;
        p.total_nbytes = total;
        p.fin = FIN_VERIFY;
        p.expect[0] = src->digest[0];
        p.expect[1] = src->digest[1];
        p.dst_hdr = reinterpret_cast<lzyb_frame_header_t*>(dd.buf);
        dd.dtype = src->dtype;
        dd.ndim = src->ndim;
        memcpy(dd.shape, src->shape, sizeof(dd.shape));
        fill_template(dd, &p.hdr_template);
        arm_mailbox(dd, p);
        p.wait_flags = k == 0 ? nullptr : reinterpret_cast<const unsigned int*>(flags[k - 1]);
        p.signal_flags = k == n - 1 ? nullptr : reinterpret_cast<unsigned int*>(flags[k]);
        p.chunk_shift = CHUNK_SHIFT;
        if (k == 0 && src->last_ev && src->last_stream != d.stream) cudaStreamWaitEvent(d.stream, src->last_ev, 0);
        chain_digest_kernel<<<(unsigned)ntiles, 256, 0, d.stream>>>(p);  // one 256 KiB span per CTA, scheduled in order
        g.st_launch++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            rc = fail(LZYB_E_LOCAL, "chain kernel launch failed: %s", cudaGetErrorString(e));
            break;
        }
        g.st_pull++;
        if (up.dev == dd.dev && !up.remote)
            g.st_bpull_l += total;
        else
            g.st_bpull_p += total;
        dd.delivered = total;
        dd.state = LZYB_ST_FILLING;
        dd.digest_pending = true;
        note(dd, d.stream);
    }
    if (rc == LZYB_OK) {
        for (int k = 0; k < n && rc == LZYB_OK; ++k) rc = lzyb_input_finish(dsts[k], src_h, nullptr);
    }
    for (int k = 0; k < n; ++k) {
        if (!flags[k]) continue;
        DeviceCtx& d = *g.devs[chain[k]->dev];
        DeviceGuard dg(d.ordinal);
        cudaStreamSynchronize(d.stream);
        free_frame(d, flags[k], flag_bytes);
    }
    return rc;
}

int lzyb_broadcast(lzyb_handle_t src_h, const lzyb_handle_t* dsts, int n, uint64_t chunk_bytes) {
    if (n <= 0 || !dsts) return fail(LZYB_E_ARG, "no consumers");
    if (chunk_bytes == 0) {
        // the device-chained kernel needs every stage on its own GPU (a stage that spins must not be able to
        // starve the stage it waits for); otherwise fall back to the host-driven chain
        std::vector<int> devs;
        bool distinct = true;
        for (int k = 0; k < n; ++k) {
            auto s = find(dsts[k]);
            if (!s) return fail(LZYB_E_ARG, "unknown consumer handle #%d", k);
            for (int d : devs) distinct = distinct && d != s->dev;
            devs.push_back(s->dev);
        }
        if (distinct) return broadcast_device_chained(src_h, dsts, n);
        chunk_bytes = 64ull << 20;
    }
    return broadcast_host_driven(src_h, dsts, n, chunk_bytes);
}

// ---- pinned host ring --------------------------------------------------------------------------

static int ring_alloc(uint64_t bytes, uint64_t* off) {
    std::lock_guard<std::mutex> lk(g.ring_mu);
    while (!g.ring_frames.empty() && !g.ring_frames.front().live) {
        g.ring_used -= g.ring_frames.front().bytes;
        g.ring_frames.pop_front();
    }
    if (g.ring_frames.empty()) g.ring_head = 0;
    if (bytes > g.ring_bytes) return fail(LZYB_E_LOCAL, "object of %llu bytes exceeds the host ring (%llu)", (unsigned long long)bytes,
                                          (unsigned long long)g.ring_bytes);
    uint64_t tail = g.ring_frames.empty() ? 0 : g.ring_frames.front().off;
    uint64_t head = g.ring_head;
    if (!g.ring_frames.empty() && head <= tail && g.ring_used > 0) {
        // wrapped: free space is [head, tail)
        if (head + bytes > tail) return fail(LZYB_E_LOCAL, "host ring full");
    } else {
        if (head + bytes > g.ring_bytes) {  // wrap: pad the end
            uint64_t pad = g.ring_bytes - head;
            if (bytes > tail) return fail(LZYB_E_LOCAL, "host ring full");
            if (pad) {
                g.ring_frames.push_back({head, pad, false});
                g.ring_used += pad;
            }
            head = 0;
        }
    }
    *off = head;
    g.ring_frames.push_back({head, bytes, true});
    g.ring_used += bytes;
    g.ring_head = head + bytes;
    return LZYB_OK;
}

static SlotCtl* borrow_ctl(DeviceCtx& d, int* idx) {
    std::lock_guard<std::mutex> lk(d.mu);
    if (d.ctl_free.empty()) return nullptr;
    *idx = d.ctl_free.back();
    d.ctl_free.pop_back();
    return d.ctl_slab + *idx;
}
static void return_ctl(DeviceCtx& d, int idx) {
    std::lock_guard<std::mutex> lk(d.mu);
    d.ctl_free.push_back(idx);
}

// run one blocking scan (+ publish or verify) over host-mapped or device memory
static int scan_blocking(DeviceCtx& d, const uint8_t* src, uint64_t nbytes, int fin, const uint64_t expect[2],
                         lzyb_frame_header_t* dst_hdr, const lzyb_frame_header_t* tmpl, uint64_t out_digest[2],
                         cudaStream_t st) {
    int cidx = -1, midx = -1;
    SlotCtl* ctl = borrow_ctl(d, &cidx);
    if (!ctl) return fail(LZYB_E_LOCAL, "no free accumulator block");
    Mailbox* mb;
    {
        std::lock_guard<std::mutex> lk(g.mbox_mu);
        if (g.mbox_free.empty()) {
            return_ctl(d, cidx);
            return fail(LZYB_E_LOCAL, "out of mailboxes");
        }
        midx = g.mbox_free.back();
        g.mbox_free.pop_back();
        mb = g.mbox_slab + midx;
        memset((void*)mb, 0, sizeof(Mailbox));
    }
    CopyParams p;
    memset(&p, 0, sizeof(p));
    p.src = src;
    p.nbytes = nbytes;
    p.ctl = ctl;
    p.mbox = mb;
    p.seq = 1;
    p.total_nbytes = nbytes;
    p.fin = fin;
    if (expect) {
        p.expect[0] = expect[0];
        p.expect[1] = expect[1];
    }
    p.dst_hdr = dst_hdr;
    if (tmpl) p.hdr_template = *tmpl;
    int rc = launch_stream(d, p, false, st);
    cudaError_t e = cudaStreamSynchronize(st);
    int status = (int)mb->status;
    if (out_digest) {
        out_digest[0] = mb->digest[0];
        out_digest[1] = mb->digest[1];
    }
    {
        std::lock_guard<std::mutex> lk(g.mbox_mu);
        g.mbox_free.push_back(midx);
    }
    return_ctl(d, cidx);
    if (rc != LZYB_OK) return rc;
    if (e != cudaSuccess) return fail(LZYB_E_LOCAL, "scan failed: %s", cudaGetErrorString(e));
    if (status == 0) return fail(LZYB_E_LOCAL, "scan did not report");
    return status;
}

int lzyb_frame_host_object(const void* bytes, uint64_t len, uint64_t* ring_off) {
    if (!g.inited) return fail(LZYB_E_STATE, "lzyb_init has not been called (or found no GPU)");
    if (!ring_off || (len && !bytes)) return fail(LZYB_E_ARG, "null argument");
    uint64_t frame = HEADER_BYTES + round_up(len ? len : 16, 128);
    uint64_t off = 0;
    int rc = ring_alloc(frame, &off);
    if (rc != LZYB_OK) return rc;
    uint8_t* base = g.ring + off;
    memset(base, 0, HEADER_BYTES);
    if (len) memcpy(base + HEADER_BYTES, bytes, len);
    DeviceCtx& d = *g.devs[0];
    DeviceGuard dg(d.ordinal);
    lzyb_frame_header_t tmpl;
    memset(&tmpl, 0, sizeof(tmpl));
    memcpy(tmpl.magic, "LZYB200", 8);
    tmpl.version = LZYB_ABI_VERSION;
    tmpl.dtype = LZYB_DT_PICKLE;
    tmpl.nbytes = len;
    tmpl.ndim = 1;
    tmpl.shape[0] = len;
    // the frame kernel reads the payload through the device mapping of the pinned ring and writes
    // the header (digest included) back into the ring
    int st = scan_blocking(d, base + HEADER_BYTES, len, FIN_PUBLISH, nullptr, reinterpret_cast<lzyb_frame_header_t*>(base),
                           &tmpl, nullptr, d.stream);
    if (st < 0) {
        lzyb_ring_release(off);
        return st;
    }
    g.st_frame++;
    g.st_bframe += len;
    *ring_off = off;
    return LZYB_OK;
}

static int ring_check(uint64_t off, lzyb_frame_header_t** hdr) {
    if (!g.inited) return fail(LZYB_E_STATE, "lzyb_init has not been called");
    std::lock_guard<std::mutex> lk(g.ring_mu);
    for (auto& f : g.ring_frames) {
        if (f.off == off && f.live) {
            *hdr = reinterpret_cast<lzyb_frame_header_t*>(g.ring + off);
            return LZYB_OK;
        }
    }
    return fail(LZYB_E_NOT_FOUND, "no live frame at ring offset %llu", (unsigned long long)off);
}

int lzyb_ring_frame_info(uint64_t ring_off, lzyb_frame_header_t* out) {
    lzyb_frame_header_t* h = nullptr;
    int rc = ring_check(ring_off, &h);
    if (rc != LZYB_OK) return rc;
    if (!out) return fail(LZYB_E_ARG, "null argument");
    *out = *h;
    return LZYB_OK;
}

int64_t lzyb_ring_read(uint64_t ring_off, void* dst, uint64_t cap) {
    lzyb_frame_header_t* h = nullptr;
    int rc = ring_check(ring_off, &h);
    if (rc != LZYB_OK) return rc;
    if (memcmp(h->magic, "LZYB200", 8) != 0 || !(h->flags & 1u)) return fail(LZYB_E_PEER, "ring frame %llu has no valid header", (unsigned long long)ring_off);
    if (h->nbytes > cap) return fail(LZYB_E_ARG, "frame holds %llu bytes, buffer %llu", (unsigned long long)h->nbytes, (unsigned long long)cap);
    DeviceCtx& d = *g.devs[0];
    DeviceGuard dg(d.ordinal);
    uint64_t expect[2] = {h->digest[0], h->digest[1]};
    int st = scan_blocking(d, g.ring + ring_off + HEADER_BYTES, h->nbytes, FIN_VERIFY, expect, nullptr, nullptr, nullptr, d.stream);
    if (st < 0) return st;
    g.st_frame++;
    if (st == 2) return fail(LZYB_E_PEER, "digest mismatch on ring frame %llu", (unsigned long long)ring_off);
    if (h->nbytes) memcpy(dst, g.ring + ring_off + HEADER_BYTES, h->nbytes);
    return (int64_t)h->nbytes;
}

int lzyb_ring_release(uint64_t ring_off) {
    std::lock_guard<std::mutex> lk(g.ring_mu);
    for (auto& f : g.ring_frames) {
        if (f.off == ring_off && f.live) {
            f.live = false;
            return LZYB_OK;
        }
    }
    return LZYB_OK;
}

// ---- utilities ---------------------------------------------------------------------------------

int lzyb_content_key(const uint64_t digest[2], uint32_t dtype_flags, uint32_t ndim, const uint64_t* shape, uint64_t key_out[2]) {
    if (!digest || !key_out || (ndim && !shape)) return fail(LZYB_E_ARG, "null argument");
    if (ndim > LZYB_MAX_DIMS) return fail(LZYB_E_ARG, "ndim %u > %d", ndim, LZYB_MAX_DIMS);
    lzyb_content_key_fold(digest, dtype_flags, ndim, shape, key_out);
    return LZYB_OK;
}

int lzyb_digest_device(int gpu, const void* dptr, uint64_t nbytes, uint64_t digest_out[2], void* stream) {
    if (!g.inited) return fail(LZYB_E_STATE, "lzyb_init has not been called (or found no GPU)");
    int di = gpu - g.first;
    if (di < 0 || di >= (int)g.devs.size()) return fail(LZYB_E_ARG, "GPU %d is not managed by this process", gpu);
    if (!digest_out || (nbytes && !dptr)) return fail(LZYB_E_ARG, "null argument");
    DeviceCtx& d = *g.devs[di];
    DeviceGuard dg(d.ordinal);
    int st = scan_blocking(d, static_cast<const uint8_t*>(dptr), nbytes, FIN_PUBLISH, nullptr, nullptr, nullptr, digest_out,
                           pick_stream(d, stream));
    if (st < 0) return st;
    g.st_scan++;
    return LZYB_OK;
}

int lzyb_fill_pattern(int gpu, void* dptr, uint64_t nbytes, uint64_t seed, void* stream) {
    if (!g.inited) return fail(LZYB_E_STATE, "lzyb_init has not been called (or found no GPU)");
    int di = gpu - g.first;
    if (di < 0 || di >= (int)g.devs.size()) return fail(LZYB_E_ARG, "GPU %d is not managed by this process", gpu);
    if (reinterpret_cast<uintptr_t>(dptr) & 15) return fail(LZYB_E_ARG, "fill target must be 16-byte aligned");
    DeviceCtx& d = *g.devs[di];
    DeviceGuard dg(d.ordinal);
    uint64_t nvec = nbytes >> 4;
    uint64_t want = (nvec + 255) / 256;
    int grid = d.sm_count * 8;
    if ((uint64_t)grid > want) grid = (int)(want ? want : 1);
    fill_pattern_kernel<<<grid, 256, 0, pick_stream(d, stream)>>>(static_cast<uint8_t*>(dptr), nbytes, seed);
    g.st_launch++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LZYB_E_LOCAL, "fill launch failed: %s", cudaGetErrorString(e));
    return LZYB_OK;
}

int lzyb_stats(lzyb_stats_t* out) {
    if (!out) return fail(LZYB_E_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    out->kernel_launches = g.st_launch;
    out->pack_launches = g.st_pack;
    out->pull_launches = g.st_pull;
    out->scan_launches = g.st_scan;
    out->frame_launches = g.st_frame;
    out->bytes_packed = g.st_bpack;
    out->bytes_pulled_local = g.st_bpull_l;
    out->bytes_pulled_peer = g.st_bpull_p;
    out->bytes_framed_host = g.st_bframe;
    out->h2d_bytes = g.st_h2d;
    out->d2h_bytes = g.st_d2h;
    out->slots_live = g.st_live;
    out->hbm_bytes_live = g.st_hbm_live;
    uint64_t pooled = 0;
    for (auto& d : g.devs) pooled += d->pooled_bytes;
    out->hbm_bytes_pooled = pooled;
    out->dma_copies = g.st_dma;
    out->push_launches = g.st_push;
    out->frames_moved = g.st_moves;
    return LZYB_OK;
}

int lzyb_stats_reset(void) {
    g.st_launch = g.st_pack = g.st_pull = g.st_scan = g.st_frame = 0;
    g.st_dma = 0;
    g.st_push = g.st_moves = 0;
    g.st_bpack = g.st_bpull_l = g.st_bpull_p = g.st_bframe = g.st_h2d = g.st_d2h = 0;
    return LZYB_OK;
}

int lzyb_set_copy_engine(int engine) {
    if (engine < 0 || engine > 4)
        return fail(LZYB_E_ARG, "engine must be 0 (auto), 1 (LDG/STG), 2 (TMA bulk), 3 (copy engine + trailing scan, peer pulls) or "
                                "4 (push kernel + landed scan, in-process peer transfers)");
    g.engine = engine;
    return LZYB_OK;
}

int lzyb_set_push_min_bytes(uint64_t nbytes) {
    g.push_min = nbytes ? nbytes : PUSH_MIN_DEFAULT;
    return LZYB_OK;
}

int lzyb_set_dma_schedule(uint64_t last_bytes, int growth) {
    if (last_bytes != 0 && (last_bytes < (1ull << 20) || (last_bytes & 15ull)))
        return fail(LZYB_E_ARG, "last sub-chunk must be 0 (auto) or a multiple of 16 bytes, at least 1 MiB");
    if (growth < 1 || growth > 64) return fail(LZYB_E_ARG, "growth must be in [1, 64]");
    g.dma_last_bytes = last_bytes;
    g.dma_growth = growth;
    return LZYB_OK;
}

int lzyb_dma_schedule(uint64_t nbytes, uint64_t* sizes_out, int cap) {
    const std::vector<uint64_t> sizes = dma_schedule(nbytes);
    for (size_t i = 0; i < sizes.size() && (int)i < cap; ++i) sizes_out[i] = sizes[i];
    return (int)sizes.size();
}

int lzyb_set_tma_config(int index) {
    if (index < 0 || index >= kNumTmaCfgs) return fail(LZYB_E_ARG, "tma config index must be in [0, %d)", kNumTmaCfgs);
    g.tma_cfg = index;
    return LZYB_OK;
}

int lzyb_set_debug_flags(int flags) {
    g_debug_flags = flags;
    return LZYB_OK;
}

int lzyb_set_grid(int ctas_per_sm) {
    if (ctas_per_sm == GRID_AUTO) {
        g.ctas_per_sm = GRID_AUTO;
        return LZYB_OK;
    }
    if (ctas_per_sm < -64 || ctas_per_sm > 16) return fail(LZYB_E_ARG, "ctas_per_sm must be in [-64, 16] (0 = one tile per CTA, -T = T tiles per CTA) or -1000 (auto)");
    g.ctas_per_sm = ctas_per_sm;
    return LZYB_OK;
}

int lzyb_fault_inject(lzyb_handle_t h, int kind, uint64_t arg) {
    if (h == 0) {
        if (kind == LZYB_FAULT_SLAB_RESERVE) {
            g.ctl_reserve = (int)std::min<uint64_t>(arg, CTL_SLAB);
            return LZYB_OK;
        }
        g.global_fault = kind;
        return LZYB_OK;
    }
    auto s = find(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slot handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (kind == LZYB_FAULT_CORRUPT_ONCE) {
        if (arg >= s->nbytes) return fail(LZYB_E_ARG, "corrupt offset beyond payload");
        DeviceCtx& d = *g.devs[s->dev];
        DeviceGuard dg(d.ordinal);
        if (s->last_ev) cudaEventSynchronize(s->last_ev);
        corrupt_byte_kernel<<<1, 1, 0, d.stream>>>(s->payload() + arg);
        CU(cudaStreamSynchronize(d.stream));
        return LZYB_OK;
    }
    s->fault_kind = kind;
    s->fault_arg = arg;
    return LZYB_OK;
}

}  // extern "C"
