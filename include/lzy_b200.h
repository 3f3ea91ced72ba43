/*
 * lzy_b200.h — C ABI of liblzy_b200.so, the B200-native slot/channel data plane.
 *
 * This is the drop-in boundary for ONE path of lambdazy/lzy: the slot/channel data plane that
 * carries @op inputs/outputs between workers.  Every entry point names the reference interface
 * it replaces (paths relative to the reference tree).  Plain pointers and sizes only; no
 * exceptions cross the ABI; the caller owns strings, the library owns device memory.
 *
 * Binding sketches (JNI for the Java worker, ctypes for pylzy) are in INTEGRATION.md.
 *
 * Threading: all functions are thread-safe; operations on one handle serialise on that handle.
 * A handle may be closed from another thread (InputSlot.close / OutputSlot.close semantics,
 * lzy/slots/src/main/java/ai/lzy/slots/InputSlot.java:110-117, OutputSlot.java:90-133).
 */
#ifndef LZY_B200_H
#define LZY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZYB_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------------ */
/* Byte-count returning calls use: >= 0 bytes moved, LZYB_EOF (-1) end of stream — the contract
 * of InputTransfer.transferChunkTo (lzy/slots/.../transfers/InputTransfer.java:8-11) —
 * and < -1 for errors.  LZYB_E_PEER is the ABI image of InputTransfer.ReadException (peer
 * fault: report TransferFailed, ask the channel manager for another peer, resume at offset);
 * every other error is the image of IOException (local fault: the slot fails). */
#define LZYB_OK 0
#define LZYB_EOF (-1)
#define LZYB_E_PEER (-2)      /* peer fault: digest mismatch, injected peer failure, dead peer      */
#define LZYB_E_LOCAL (-3)     /* local fault: CUDA error, out of memory, sink closed                */
#define LZYB_E_ARG (-4)       /* bad argument / misuse (unknown handle, misaligned offset, ...)     */
#define LZYB_E_STATE (-5)     /* wrong state (read before commit, transfer on closed slot, ...)     */
#define LZYB_E_NOGPU (-6)     /* no usable CUDA device: the product path never falls back to CPU    */
#define LZYB_E_NOT_FOUND (-7) /* gRPC NOT_FOUND image (unknown slot / channel / transfer)           */
#define LZYB_E_PRECONDITION (-8) /* gRPC FAILED_PRECONDITION image (transfer is not active)         */
#define LZYB_E_UNAVAILABLE (-9)  /* gRPC UNAVAILABLE image (unbind while transfers pending)         */
#define LZYB_E_INTERNAL (-10)    /* gRPC INTERNAL image                                              */
#define LZYB_E_UNIMPLEMENTED (-11) /* gRPC UNIMPLEMENTED image (Read on an input slot)              */
#define LZYB_E_INVALID (-12)     /* gRPC INVALID_ARGUMENT image                                      */
#define LZYB_E_EXISTS (-13)      /* peer id already associated with channel                          */

typedef uint64_t lzyb_handle_t;

/* element type codes carried in the frame header (LMD.DataScheme metadata "lzy.dtype",
 * model/src/main/proto/ai/lzy/v1/common/data-scheme.proto:8-14) */
enum lzyb_dtype {
    LZYB_DT_BYTES = 0, LZYB_DT_U8 = 1, LZYB_DT_I8 = 2, LZYB_DT_I16 = 3, LZYB_DT_I32 = 4, LZYB_DT_I64 = 5,
    LZYB_DT_F16 = 6, LZYB_DT_BF16 = 7, LZYB_DT_F32 = 8, LZYB_DT_F64 = 9, LZYB_DT_BOOL = 10,
    LZYB_DT_U16 = 11, LZYB_DT_U32 = 12, LZYB_DT_U64 = 13, LZYB_DT_PICKLE = 32
};

enum lzyb_slot_state {
    LZYB_ST_EMPTY = 0,     /* created, payload not complete                                       */
    LZYB_ST_FILLING = 1,   /* input slot: download in progress (InputSlot.State.DOWNLOADING)       */
    LZYB_ST_READY = 2,     /* committed / fully downloaded and verified                            */
    LZYB_ST_FAILED = 3,
    LZYB_ST_CLOSED = 4,
    LZYB_ST_MOVED = 5      /* output slot whose frame went to its sole consumer (lzyb_slot_set_movable)      */
};

#define LZYB_MAX_DIMS 8

/* 128-byte frame header that precedes every slot payload in HBM (and every object in the host
 * ring).  A slot buffer is therefore self-describing: a consumer that maps it over CUDA IPC
 * needs nothing else to size, type and verify the payload. */
typedef struct lzyb_frame_header {
    uint8_t magic[8];          /* "LZYB200\0"                                                      */
    uint32_t version;          /* LZYB_ABI_VERSION                                                 */
    uint32_t dtype;            /* enum lzyb_dtype                                                  */
    uint64_t nbytes;           /* payload bytes                                                    */
    uint32_t ndim;
    uint32_t flags;            /* bit0: committed                                                  */
    uint64_t shape[LZYB_MAX_DIMS];
    uint64_t digest[2];        /* lzyb_digest_finalize output                                      */
    uint64_t sum;              /* raw partials, kept so digests of sub-ranges stay composable      */
    uint64_t xr;
} lzyb_frame_header_t;         /* sizeof == 128 */

typedef struct lzyb_slot_info_s {
    uint64_t nbytes;
    int32_t gpu;
    int32_t state;             /* enum lzyb_slot_state                                             */
    int32_t is_input;
    int32_t is_remote;         /* imported over CUDA IPC                                           */
    uint64_t delivered;        /* input slots: bytes delivered so far (the resume offset)          */
    uint64_t digest[2];
    uint32_t dtype;
    uint32_t ndim;
    uint64_t shape[LZYB_MAX_DIMS];
} lzyb_slot_info_t;

/* opaque descriptor that lets another worker process map a committed slot (CUDA IPC);
 * this is what rides in LC.PeerDescription.SlotPeer.peer_url = "cuda+ipc://…"
 * (model/src/main/proto/ai/lzy/v1/common/channel.proto:20-40) */
typedef struct lzyb_ipc_desc {
    uint8_t ipc_handle[64];    /* cudaIpcMemHandle_t                                               */
    uint64_t buffer_bytes;     /* header + padded payload                                          */
    uint64_t nbytes;
    int32_t gpu;               /* CUDA ordinal of the owning device                                */
    int32_t pid;
    uint64_t digest[2];
    uint64_t sum;
    uint64_t xr;
    uint32_t dtype;
    uint32_t ndim;
    uint64_t shape[LZYB_MAX_DIMS];
    uint8_t reserved[64];      /* [0..8): arrival word of a push sink (lzyb_sink_export), else zero        */
} lzyb_ipc_desc_t;             /* sizeof == 256 */

typedef struct lzyb_stats_s {
    uint64_t kernel_launches;  /* launches of this library's kernels since init / last reset       */
    uint64_t pack_launches;
    uint64_t pull_launches;
    uint64_t scan_launches;
    uint64_t frame_launches;
    uint64_t bytes_packed;
    uint64_t bytes_pulled_local;
    uint64_t bytes_pulled_peer;
    uint64_t bytes_framed_host;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t slots_live;
    uint64_t hbm_bytes_live;
    uint64_t hbm_bytes_pooled;
    uint64_t dma_copies;       /* copy-engine sub-chunk copies issued by peer pulls (engine 3 / auto)  */
    uint64_t push_launches;    /* streaming-kernel launches on the SOURCE GPU of a GPU→GPU transfer (pushes)  */
    uint64_t frames_moved;     /* same-GPU deliveries done by ownership hand-over instead of a copy    */
} lzyb_stats_t;

/* ---- lifecycle ---------------------------------------------------------------------------- */

/* Manage CUDA devices [first_gpu, first_gpu + n_gpus) from this process (n_gpus <= 0: all
 * visible).  One worker per GPU = first_gpu = LOCAL_RANK, n_gpus = 1; a single-process runtime
 * that places ops on several GPUs passes the whole range and gets peer access enabled between
 * them.  host_ring_bytes sizes the pinned host ring for non-tensor objects (0 = 64 MiB).
 * Replaces Slots(...) construction in WorkerApiImpl.init (lzy/worker/.../WorkerApiImpl.java:255-264).
 * Returns LZYB_E_NOGPU when no device is usable; there is no CPU fallback. */
int lzyb_init(int first_gpu, int n_gpus, uint64_t host_ring_bytes);
int lzyb_shutdown(void);
int lzyb_is_initialized(void);
int lzyb_abi_version(void);
int lzyb_device_count(void);                 /* devices managed by this process                   */
const char* lzyb_last_error(void);           /* thread-local, valid until the next call           */
const char* lzyb_strerror(int code);
/* lets a pluggable backend (lzy_b200_host.h: lzyb_backend_ops_t) report through the same channel */
int lzyb_set_last_error(const char* message);

/* ---- output slots: OutputSlotBackend (lzy/slots/.../backends/OutputSlotBackend.java:6-16) ----- */

/* OutputSlotBackend constructor (OutputPipeBackend.java:27-41 creates a fifo + temp file; here
 * an HBM frame of header + nbytes on `gpu`). */
int lzyb_slot_create(const char* slot_id, int gpu, uint64_t nbytes, lzyb_handle_t* out);
/* Device pointer of the payload (16-byte aligned) so a producer can write its result in place —
 * the analogue of the op writing into the fifo at <mount>/<slot> (pylzy/lzy/api/v1/startup.py:75-77). */
int lzyb_slot_device_ptr(lzyb_handle_t h, void** dptr);
int lzyb_slot_set_meta(lzyb_handle_t h, uint32_t dtype, uint32_t ndim, const uint64_t* shape);
/* Fused pack + frame + digest: copy `nbytes` from a device pointer on the slot's GPU (any
 * alignment) into the slot payload, computing the digest on the way (one pass, 2 B of HBM
 * traffic per payload byte).  Replaces fifo→tmpfile IOUtils.copyLarge
 * (OutputPipeBackend.java:43-53) + HashingIO (hashing.py:36-38).  Asynchronous on `stream`
 * (cudaStream_t; NULL = the library's stream of that GPU). */
int lzyb_slot_pack(lzyb_handle_t h, const void* src_dptr, uint64_t nbytes, void* stream);
/* Strided variant: gather `rows` rows of `row_bytes` bytes, `src_pitch` bytes apart. */
int lzyb_slot_pack_2d(lzyb_handle_t h, const void* src_dptr, uint64_t rows, uint64_t row_bytes,
                      uint64_t src_pitch, void* stream);
/* Host source (ndarray / bytes): staged H2D copy into the payload, then digest scan. */
int lzyb_slot_pack_host(lzyb_handle_t h, const void* src_host, uint64_t nbytes, void* stream);
/* Multi-part host source: a payload assembled from several host buffers (a pickle skeleton followed by
 * its out-of-band buffers).  Parts are written back to back at 16-byte aligned offsets (the library
 * zero-fills the ≤ 15 padding bytes after a part); `offset` must equal the bytes already written;
 * the part with last != 0 must end exactly at the slot's nbytes and publishes the digest. */
int lzyb_slot_pack_host_part(lzyb_handle_t h, uint64_t offset, const void* src_host, uint64_t len, int last, void* stream);
/* = OutputSlotBackend.waitCompleted(): make the payload final.  If the payload was written in
 * place the digest is computed by a read-only scan (1 B/B); if it was packed the digest is
 * already there.  Blocks until the frame header is valid; writes the 128-bit digest. */
int lzyb_slot_commit(lzyb_handle_t h, uint64_t digest_out[2], void* stream);
/* = OutputSlotBackend.readFromOffset(offset) for legacy (gRPC `Read`) consumers: copies up to
 * `len` payload bytes starting at `offset` into host memory.  Returns bytes copied, LZYB_EOF
 * at/after the end (OutputSlot.ReadThread, OutputSlot.java:238-265). */
int64_t lzyb_slot_read(lzyb_handle_t h, uint64_t offset, void* dst_host, uint64_t len);
int lzyb_slot_info(lzyb_handle_t h, lzyb_slot_info_t* out);
/* the same without waiting for anything: sizes, state, dtype, shape; the digest fields are filled only if the digest is
 * already known (see lzyb_set_lazy_commit) */
int lzyb_slot_meta(lzyb_handle_t h, lzyb_slot_info_t* out);
/* order `stream` (NULL: the calling thread) after the last launch that wrote this slot's frame — what a reader on another
 * stream needs before it touches lzyb_slot_device_ptr() */
int lzyb_slot_wait(lzyb_handle_t h, void* stream);
int lzyb_slot_header(lzyb_handle_t h, lzyb_frame_header_t* out);   /* D2H copy of the frame header */
/* = backend.close(): idempotent, returns HBM to the pool, cancels in-flight work.  Frames are
 * reference counted: every lzyb_slot_retain() needs one more close before the HBM is released. */
int lzyb_slot_close(lzyb_handle_t h);
int lzyb_slot_retain(lzyb_handle_t h);
/* The owner of a committed output slot declares that the channel has exactly ONE consumer and no other reader (what
 * the channel manager knows from its peer table, channel-manager/.../services/SlotsService.java:440-519).  A whole-
 * payload lzyb_transfer_chunk from this slot by a consumer on the same GPU then MOVES the frame instead of copying
 * it — InputSlotBackend.toOutput()'s ownership hand-over (FileInputBackend.java:38-45) applied to the producer →
 * consumer edge — provided nobody else holds the frame (no lzyb_slot_retain outstanding).  The producer slot reads
 * LZYB_ST_MOVED afterwards (further reads fail with LZYB_E_PEER).  Off by default: a slot may have late consumers. */
int lzyb_slot_set_movable(lzyb_handle_t h, int movable);
/* look a slot up by id (SlotsService.slots map, lzy/slots/.../SlotsService.java:21-30) */
int lzyb_slot_find(const char* slot_id, lzyb_handle_t* out);

/* ---- input slots: InputSlotBackend + InputTransfer ------------------------------------------- */

/* InputSlotBackend constructor (FileInputBackend.java:19-31: creates the sink file). */
int lzyb_input_open(const char* slot_id, int gpu, uint64_t nbytes, lzyb_handle_t* out);
/* = InputTransfer.transferChunkTo(sink) (InputTransfer.java:8-11; SlotInputTransfer.java:43-99):
 * consumer-pull up to max_len bytes of `src_peer`'s payload, starting at `offset`, into `dst`
 * at the same offset, accumulating the digest of what arrived.  `offset` must equal the bytes
 * already delivered (the reference repositions the sink to currentOffset, :47-49) and be a
 * multiple of 16.  Same-GPU peers are an HBM copy; other GPUs are read over NVLink through the
 * peer mapping.  Returns bytes moved, LZYB_EOF when offset == nbytes, LZYB_E_PEER / LZYB_E_LOCAL.
 * stream == NULL: runs on the library's copy stream of dst's GPU and returns after completion
 * (Java-like blocking); otherwise asynchronous on `stream`.  Asynchronous callers keep `src_peer` open until
 * `stream` has passed the transfer (lzyb_input_finish, which waits for the verdict, is such a point): closing the
 * last reference of a frame waits for the writes into it, not for readers the library was not told about. */
int64_t lzyb_transfer_chunk(lzyb_handle_t dst, lzyb_handle_t src_peer, uint64_t offset,
                            uint64_t max_len, void* stream);
/* End of stream (the EOS control message, slots-api.proto:33-47): compare the accumulated
 * digest with the producer's frame header, publish dst's own header.  LZYB_OK, or LZYB_E_PEER
 * on mismatch (dst is reset to offset 0 so another peer can be tried). Blocks. */
int lzyb_input_finish(lzyb_handle_t dst, lzyb_handle_t src_peer, void* stream);
/* Host source for an input slot (StorageInputTransfer / a legacy gRPC chunk,
 * transfers/StorageInputTransfer.java:32-41): write `len` host bytes at `offset`. */
int64_t lzyb_input_write_host(lzyb_handle_t dst, uint64_t offset, const void* src_host, uint64_t len,
                              void* stream);
/* expected digest supplied by the control plane (DataScheme.metadata "lzy.checksum") for
 * sources that have no frame header (host/storage peers); NULL = no expectation, the digest
 * computed while the bytes landed is published */
int lzyb_input_finish_expect(lzyb_handle_t dst, const uint64_t digest[2], void* stream);
/* The payload of an input slot was filled in place by an external writer that this library does not
 * drive — an `ncclBroadcast` into lzyb_slot_device_ptr(dst) when one output fans out to many
 * consumers.  Scans the payload, checks it against the producer's digest (NULL: publish what is
 * there) and makes the slot READY; LZYB_E_PEER on mismatch.  Blocks. */
int lzyb_input_adopt_external(lzyb_handle_t dst, const uint64_t digest[2], uint32_t dtype, uint32_t ndim,
                              const uint64_t* shape, void* stream);
/* = InputSlotBackend.toOutput() (FileInputBackend.java:38-45): the delivered bytes become an
 * output slot ("<slotId>-out" re-publication, InputSlot.java:160-168) with no copy. */
int lzyb_input_to_output(lzyb_handle_t h);
/* restart a failed download at offset 0 (InputSlot.download reopening the sink, InputSlot.java:232-238) */
int lzyb_input_reset(lzyb_handle_t h);

/* ---- cross-process peers -------------------------------------------------------------------- */
int lzyb_slot_export(lzyb_handle_t h, lzyb_ipc_desc_t* out);
int lzyb_peer_import(const lzyb_ipc_desc_t* desc, int onto_gpu, lzyb_handle_t* out);
/* Producer-issued PUSH between worker processes (inside one process lzyb_transfer_chunk does it by itself).
 * SM-issued NVLink traffic crosses the link in 128-byte packets: a pull costs 1.31 wire bytes per payload byte once
 * both directions of a GPU are busy (its requests share the wire with the payload the same GPU serves), a push 1.19
 * (profiles/r02_peer_variants_ncu.csv).  So between processes the consumer exports its EMPTY input frame and the
 * PRODUCER's process launches the streaming kernel (TMA ring: cp.async.bulk global → shared → peer global, digest folded
 * from the shared-memory stages) with that frame as its destination.  The verdict is the digest of what was pushed
 * against the digest the producer committed; the pushing kernel writes the sink's frame header, the verdict and —
 * last, with release semantics at system scope — the arrival word the consumer chose into the frame's trailer.
 * The consumer's GPU runs a one-thread arrival_wait_kernel on that word and fills the slot's mailbox: no copy
 * engine, no cross-process event, the consumer's stream continues with a complete frame.
 *   consumer:  lzyb_input_open → lzyb_sink_export(desc) → [desc travels to the producer] →
 *              lzyb_input_expect_push(dst, producer's slot descriptor, 0, 0, stream) → lzyb_input_finish_expect
 *   producer:  lzyb_slot_push(src, sink desc, 0, 0, stream)
 * A push carries the whole payload (offset 0, len 0 or nbytes); offset-resumed partial deliveries take the pull path
 * (lzyb_peer_import + lzyb_transfer_chunk).  A push that never comes is LZYB_E_PEER after ≈ 4 s, not a hang.
 * Ordering inside ONE process that both pushes and receives on the same GPU (a ring): issue lzyb_slot_push before
 * lzyb_input_expect_push — the first push into a frame maps it (cudaIpcOpenMemHandle), which synchronises the device
 * and would wait for a spinning waiter's timeout. */
int lzyb_sink_export(lzyb_handle_t input, lzyb_ipc_desc_t* out);
int lzyb_slot_push(lzyb_handle_t src, const lzyb_ipc_desc_t* sink, uint64_t offset, uint64_t len, void* stream);
int64_t lzyb_input_expect_push(lzyb_handle_t dst, const lzyb_ipc_desc_t* producer, uint64_t offset, uint64_t len, void* stream);

/* ---- fan-out ------------------------------------------------------------------------------ */
/* One producer, n consumers on distinct GPUs of this process (SlotsService.bind creates one
 * transfer per consumer, channel-manager/.../services/SlotsService.java:457-466; OutputSlot.read
 * one ReadThread per consumer, OutputSlot.java:66-72).  A pull chain over NVLink: consumer k pulls
 * from consumer k-1 while k-1 is still receiving, so the producer's egress is paid once.
 * chunk_bytes == 0 (default): ONE kernel launch per consumer GPU, all running concurrently and
 * handing 1 MiB chunks to each other through progress counters in peer memory (acquire/release at
 * system scope) — needs every consumer on its own GPU.  chunk_bytes > 0: host-driven chain, one
 * launch per (consumer, chunk), chained by CUDA events.  All dsts end READY and verified. Blocks. */
int lzyb_broadcast(lzyb_handle_t src, const lzyb_handle_t* dsts, int n, uint64_t chunk_bytes);

/* ---- pinned host ring for non-tensor objects -------------------------------------------------- */
/* Frame `len` pickled bytes into the pinned host ring: a CUDA kernel computes the digest and
 * writes the 128-byte frame header in front of the payload.  ring_off identifies the frame. */
int lzyb_frame_host_object(const void* bytes, uint64_t len, uint64_t* ring_off);
/* Verify (CUDA digest scan over the mapped ring) and copy a frame's payload out. */
int64_t lzyb_ring_read(uint64_t ring_off, void* dst, uint64_t cap);
int lzyb_ring_frame_info(uint64_t ring_off, lzyb_frame_header_t* out);
int lzyb_ring_release(uint64_t ring_off);

/* ---- utilities ----------------------------------------------------------------------------- */
/* Content address of a framed value = payload digest ⊕ (dtype|flags, ndim, shape): what storage URIs and the
 * snapshot's data_hash are built from (the reference's md5 covers the serialized stream, dtype and shape
 * included: pylzy/lzy/api/v1/snapshot.py:147-155).  Pure host arithmetic (include/lzyb_digest.h), needs no GPU. */
int lzyb_content_key(const uint64_t digest[2], uint32_t dtype_flags, uint32_t ndim, const uint64_t* shape,
                     uint64_t key_out[2]);
int lzyb_digest_device(int gpu, const void* dptr, uint64_t nbytes, uint64_t digest_out[2], void* stream);
/* deterministic 64-bit LCG fill (BASELINE config C3's payload) */
int lzyb_fill_pattern(int gpu, void* dptr, uint64_t nbytes, uint64_t seed, void* stream);
int lzyb_copy_to_host(lzyb_handle_t h, uint64_t offset, void* dst_host, uint64_t len, void* stream);
int lzyb_stats(lzyb_stats_t* out);
int lzyb_stats_reset(void);
/* how large contiguous transfers are carried: 0 = auto, 1 = LDG/STG 16-B vectors, 2 = TMA bulk
 * (cp.async.bulk + mbarrier ring), 3 = peer pulls ride the copy engine in sub-chunks while a scan kernel on
 * the sink GPU digests each landed sub-chunk in the shadow of the next copy (same-GPU work as in auto),
 * 4 = GPU→GPU transfers inside this process are PUSHED: the streaming kernel (TMA ring, digest folded from shared
 * memory) runs on the SOURCE GPU and stores into the sink frame through the peer mapping (same-GPU work as in auto).
 * auto, between two GPUs of this process: a PULL (2, or 1 for short / unaligned chunks) while the sink GPU is not itself
 * serving a transfer — read responses are the cheapest SM traffic, 770–780 GB/s — and a PUSH (4) for transfers of at
 * least lzyb_set_push_min_bytes (default 4 MiB) into a GPU that is sourcing one right now: in a ring or a bidirectional
 * exchange a pull's requests share the wire with the payload the same GPU serves (657–671 GB/s per direction), posted
 * writes keep 684–700.  A source mapped from another process is pulled (2) unless its owner pushes (lzyb_slot_push).
 * auto never uses the copy engine; 3 stays available by name.  1 and 2 take the digest from the bytes as they LAND,
 * a push verifies what it hands to the wire (push_kernels.cuh). */
int lzyb_set_copy_engine(int engine);
int lzyb_set_push_min_bytes(uint64_t nbytes);
/* Commit in stream order.  Off (default): lzyb_slot_commit blocks until the pack / scan has finished and the digest is known
 * — OutputSlotBackend.waitCompleted() as the reference means it.  On: lzyb_slot_commit(h, NULL, …) only QUEUES the
 * completion — the slot is READY for everything that is ordered behind the pack on the device (a same-GPU hand-over, an op
 * body on the worker's stream, lzyb_copy_to_host / lzyb_slot_read / transfers, which wait for the frame's last writer on
 * their own stream) — and the digest is harvested by the first call that needs it (lzyb_slot_info, _export, a copying
 * transfer, lzyb_slot_commit with digest_out).  lzyb_input_finish on a slot that received its frame by hand-over returns
 * without waiting.  A runtime that keeps every reader stream-ordered (B200Runtime) turns it on so that an op's successor is
 * not held up by the op's pack kernel; Java-style blocking callers leave it off. */
int lzyb_set_lazy_commit(int on);
/* Grid of a cross-GPU TMA-ring transfer (pull or push): `waves` × the resident CTA count, each CTA carrying 1/waves of the
 * tiles.  1 (default): a persistent grid — nothing else runs on that GPU until the transfer has drained (a 0.35 ms pack
 * launched into a 1 GiB pull ends 1.2 ms later, tools/overlap_probe.py).  > 1: CTAs retire all along the transfer and a kernel
 * on a higher-priority stream gets their SMs. */
int lzyb_set_transfer_waves(int waves);
/* Pinned host blocks for the host side of small reads and writes (the consumer of a host-origin payload — an ndarray, a
 * pickled object — reads it back out of HBM; into pageable memory that costs a staging copy and a page fault per 4 KiB).
 * Blocks are pooled by power-of-two size (64 KiB … 16 MiB) and never shrink below what is handed out; lzyb_host_alloc
 * returns LZYB_E_LOCAL once `limit` bytes (default 512 MiB, lzyb_set_host_pool_limit) are outstanding + pooled, or for a
 * size outside the classes — the caller then uses ordinary memory.  lzyb_host_free of a pointer that is not an outstanding
 * block is LZYB_E_ARG.  Blocks still outstanding at lzyb_shutdown stay valid (they are released when freed). */
int lzyb_host_alloc(uint64_t nbytes, void** out);
int lzyb_host_free(void* block);
int lzyb_set_host_pool_limit(uint64_t bytes);
/* sub-chunk schedule of a copy-engine pull: the last sub-chunk is last_bytes (0 = payload/32 clamped to
 * 4..32 MiB; else a multiple of 16, at least 1 MiB), each earlier one `growth` times larger (default 6; 1 =
 * uniform), the first takes the remainder */
int lzyb_set_dma_schedule(uint64_t last_bytes, int growth);
/* the sub-chunk sizes a copy-engine pull of nbytes would use under the current schedule, first to last (pure
 * arithmetic, needs no GPU): writes at most cap entries, returns how many there are */
int lzyb_dma_schedule(uint64_t nbytes, uint64_t* sizes_out, int cap);
/* LDG/STG grid shape: n > 0 = persistent, n CTAs per SM; 0 = one 32 KiB tile per CTA; -T = T tiles per CTA;
 * -1000 = auto (default) */
int lzyb_set_grid(int ctas_per_sm);
/* shape of the TMA shared-memory ring, for measurement: 0 = 6×16 KiB ×2 CTAs/SM, 1 = 3×32 KiB ×2 (default),
 * 2 = 6×32 KiB ×1, 3 = 12×16 KiB ×1, 4 = 8×8 KiB ×3, 5 = 4×16 KiB ×3, 6 = 2×8 KiB ×1 (stress shape: the shallowest ring) */
int lzyb_set_tma_config(int index);
/* measurement / stress only.  bit 0: run the TMA ring without the digest arithmetic (digests are then wrong);
 * bit 1: launch the TMA ring with 4 CTAs, so every CTA recycles its stages thousands of times */
int lzyb_set_debug_flags(int flags);

/* fault injection at the backend seam (the reference's InMemBackend failOpen/failRead/failAfter,
 * lzy/slots/src/test/java/ai/lzy/slots/SlotsTest.java:866-943) */
enum lzyb_fault {
    LZYB_FAULT_NONE = 0,
    LZYB_FAULT_PEER_AFTER = 1,   /* transfer_chunk returns LZYB_E_PEER once `arg` bytes were delivered */
    LZYB_FAULT_CORRUPT_ONCE = 2, /* flip one payload byte at offset `arg` after the next chunk lands   */
    LZYB_FAULT_LOCAL_OPEN = 3,   /* next input_open fails with LZYB_E_LOCAL                            */
    LZYB_FAULT_SLAB_RESERVE = 4  /* (handle 0) treat `arg` per-device slot control entries as taken, so slab
                                    exhaustion can be provoked without 8192 live slots; 0 restores          */
};
int lzyb_fault_inject(lzyb_handle_t slot_or_zero, int kind, uint64_t arg);

#ifdef __cplusplus
}
#endif
#endif /* LZY_B200_H */
