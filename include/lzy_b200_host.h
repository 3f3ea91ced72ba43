/*
 * lzy_b200_host.h — C ABI of the host-side control logic that sits either side of the data plane:
 * the channel manager's peer/transfer matcher and the graph dispatcher.  Same library
 * (liblzy_b200.so), same conventions as lzy_b200.h (plain structs, negative LZYB_E_* codes that are
 * the images of the reference's gRPC status codes).  Nothing here touches CUDA.
 */
#ifndef LZY_B200_HOST_H
#define LZY_B200_HOST_H

#include "lzy_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- LC.PeerDescription (model/src/main/proto/ai/lzy/v1/common/channel.proto:20-40) ---------- */
#define LZYB_PEER_NONE 0
#define LZYB_PEER_SLOT 1     /* oneof slot_peer { peer_url }                                       */
#define LZYB_PEER_STORAGE 2  /* oneof storage_peer { storage_uri, credentials }                    */

typedef struct lzyb_peer_desc {
    char peer_id[128];
    int32_t kind;            /* LZYB_PEER_SLOT | LZYB_PEER_STORAGE                                */
    int32_t reserved;
    char url[512];           /* SlotPeer.peer_url  or  StoragePeer.storage_uri                    */
} lzyb_peer_desc_t;

/* LCMS.BindRequest.Role (lzy/channel-manager-api/.../channel-manager.proto) */
#define LZYB_ROLE_UNSPECIFIED 0
#define LZYB_ROLE_PRODUCER 1
#define LZYB_ROLE_CONSUMER 2

/* BindResponse / TransferFailedResponse: optional peer + transfer id */
typedef struct lzyb_bind_result {
    int32_t has_peer;
    int32_t reserved;
    lzyb_peer_desc_t peer;
    char transfer_id[64];
} lzyb_bind_result_t;

/* transfer states (channel-manager/.../db/TransferDao.java:47-52) */
enum lzyb_transfer_state { LZYB_TR_PENDING = 0, LZYB_TR_ACTIVE = 1, LZYB_TR_COMPLETED = 2, LZYB_TR_FAILED = 3 };

/* peer priorities (channel-manager/.../db/PeerDao.java:44-53) */
#define LZYB_PRIO_PRIMARY 7
#define LZYB_PRIO_BACKUP 5
#define LZYB_PRIO_DONT_USE (-1)

/* What the channel manager asks the outside world to do after a reply has been sent:
 *  START_TRANSFER = ActionScheduler.startTransfer → LzySlotsApi.StartTransfer(slot_id = consumer,
 *  peer = producer) (channel-manager/.../ActionScheduler.java:61-91);
 *  ABORT_WORKFLOW = LzyServiceClient.destroyChannelAndWorkflow (LzyServiceClient.java:37-55). */
#define LZYB_ACT_NONE 0
#define LZYB_ACT_START_TRANSFER 1
#define LZYB_ACT_ABORT_WORKFLOW 2

typedef struct lzyb_cm_action {
    int32_t kind;
    int32_t reserved;
    char channel_id[64];
    char transfer_id[64];
    char consumer_id[128];        /* StartTransferRequest.slot_id                                  */
    char consumer_url[512];
    lzyb_peer_desc_t producer;    /* StartTransferRequest.peer                                     */
    char dial_url[512];           /* the URL the reference dials: the PRODUCER's (ActionScheduler.java:64) */
    char execution_id[128];
    char workflow_name[128];
    char reason[256];
    char idempotency_key[128];
} lzyb_cm_action_t;

/* ---- channel manager ------------------------------------------------------------------------ */
int lzyb_cm_create(lzyb_handle_t* cm);
int lzyb_cm_free(lzyb_handle_t cm);
/* deterministic tie-break for tests: FIND_PRODUCER orders by priority DESC, RANDOM() (PeerDaoImpl.java:61-65) */
int lzyb_cm_seed(lzyb_handle_t cm, uint64_t seed);

/* LzyChannelManagerPrivate.GetOrCreate (channel-manager/.../services/ChannelService.java:39-116).
 * storage_role: LZYB_ROLE_PRODUCER (pre-existing data) or LZYB_ROLE_CONSUMER (op output). */
int lzyb_cm_get_or_create(lzyb_handle_t cm, const char* user_id, const char* workflow_name, const char* execution_id,
                          const char* data_format, const char* scheme_format, const char* scheme_content,
                          int storage_role, const char* storage_uri, char channel_id_out[64]);
/* Destroy / DestroyAll (ChannelService.java:119-155) */
int lzyb_cm_destroy(lzyb_handle_t cm, const char* channel_id, const char* reason);
int lzyb_cm_destroy_all(lzyb_handle_t cm, const char* execution_id, const char* reason);

/* LzyChannelManager.Bind (channel-manager/.../services/SlotsService.java:67-145) */
int lzyb_cm_bind(lzyb_handle_t cm, const char* idempotency_key, const char* channel_id, const char* peer_id,
                 const char* execution_id, const char* peer_url, int role, lzyb_bind_result_t* out);
/* Unbind (:148-188): LZYB_E_UNAVAILABLE while the peer has PENDING/ACTIVE transfers */
int lzyb_cm_unbind(lzyb_handle_t cm, const char* channel_id, const char* peer_id);
/* TransferCompleted (:294-358) */
int lzyb_cm_transfer_completed(lzyb_handle_t cm, const char* idempotency_key, const char* channel_id,
                               const char* transfer_id);
/* TransferFailed (:191-291): on success `out` carries the new peer and transfer id */
int lzyb_cm_transfer_failed(lzyb_handle_t cm, const char* idempotency_key, const char* channel_id,
                            const char* transfer_id, const char* description, lzyb_bind_result_t* out);
/* GetChannelsStatus (:361-407): writes a JSON document
 *   {"channels":[{"channel_id":..,"execution_id":..,"scheme":{..},"producers":[{"peer_id":..,"kind":..,"url":..}],"consumers":[..]}]}
 * channel_ids may be NULL/0 = every channel of the execution.  Returns bytes needed (may exceed cap). */
int64_t lzyb_cm_channels_status(lzyb_handle_t cm, const char* execution_id, const char* const* channel_ids,
                                int n_channel_ids, char* json_out, uint64_t cap);

/* deferred actions: returns 1 and fills `out` when one is available within timeout_ms, else 0 */
int lzyb_cm_poll_action(lzyb_handle_t cm, lzyb_cm_action_t* out, int timeout_ms);
/* outcome of a START_TRANSFER action: ok → PENDING→ACTIVE (ActionScheduler.java:79-91);
 * !ok → channel dropped + ABORT_WORKFLOW queued (:92-104) */
int lzyb_cm_action_result(lzyb_handle_t cm, const char* channel_id, const char* transfer_id, int ok, const char* reason);
/* ActionScheduler.restoreActions (:47-59): re-queue START_TRANSFER for every PENDING transfer */
int lzyb_cm_restore_actions(lzyb_handle_t cm);

/* introspection used by the conformance tests (DbTest.java) */
int lzyb_cm_transfer_state(lzyb_handle_t cm, const char* channel_id, const char* transfer_id, int* state_out);
int lzyb_cm_peer_priority(lzyb_handle_t cm, const char* channel_id, const char* peer_id, int* priority_out);
int lzyb_cm_is_aborted(lzyb_handle_t cm, const char* execution_id, char* reason_out, uint64_t cap);

/* ---- slots runtime: the worker-side state machines of lzy/slots -------------------------------- */
/* Replaces, behind the same seams (paths relative to lzy/slots/src/main/java/ai/lzy/slots/):
 *   Slots.java:44-106                     one per worker: slots server + execution contexts
 *   SlotsService.java:21-69               slot registry, Read / StartTransfer routing
 *   SlotsExecutionContext.java:45-107     beforeExecution / afterExecution / close / add
 *   InputSlot.java:127-295                BINDING → WAITING_FOR_PEER → DOWNLOADING → READY (+ "-out")
 *   OutputSlot.java:135-273               waitCompleted → register → Bind(PRODUCER) → storage upload
 *   transfers/TransferFactory.java:19-33  peer-kind dispatch (slot peer → NVLink pull, storage → host blob)
 * and the worker staging order of lzy/worker/.../WorkerApiImpl.java:191-225. */

/* The backend seam (InputSlotBackend / OutputSlotBackend / InputTransfer).  The default table is the
 * HBM data plane of lzy_b200.h; the reference's tests swap in an in-memory backend with failure
 * knobs at exactly this seam (SlotsTest.InMemBackend, SlotsTest.java:866-912) and so do ours. */
typedef struct lzyb_backend_ops {
    int (*slot_create)(const char* slot_id, int gpu, uint64_t nbytes, lzyb_handle_t* out);
    int (*input_open)(const char* slot_id, int gpu, uint64_t nbytes, lzyb_handle_t* out);
    int (*slot_commit)(lzyb_handle_t h, uint64_t digest_out[2], void* stream);
    int64_t (*slot_read)(lzyb_handle_t h, uint64_t offset, void* dst_host, uint64_t len);
    int (*slot_info)(lzyb_handle_t h, lzyb_slot_info_t* out);
    int (*slot_find)(const char* slot_id, lzyb_handle_t* out);
    int (*slot_close)(lzyb_handle_t h);
    int64_t (*transfer_chunk)(lzyb_handle_t dst, lzyb_handle_t src_peer, uint64_t offset, uint64_t max_len, void* stream);
    int (*input_finish)(lzyb_handle_t dst, lzyb_handle_t src_peer, void* stream);
    int64_t (*input_write_host)(lzyb_handle_t dst, uint64_t offset, const void* src_host, uint64_t len, void* stream);
    int (*input_finish_expect)(lzyb_handle_t dst, const uint64_t digest[2], void* stream);
    int (*input_to_output)(lzyb_handle_t h);
    int (*input_reset)(lzyb_handle_t h);
    int (*slot_retain)(lzyb_handle_t h);
    /* optional (may be NULL): dtype / shape of a frame that is being filled from storage */
    int (*slot_set_meta)(lzyb_handle_t h, uint32_t dtype, uint32_t ndim, const uint64_t* shape);
    /* optional (may be NULL): a committed frame as a descriptor another worker process can map, and the mapping
     * of such a descriptor as a local read-only source (HBM backend: CUDA IPC, lzyb_slot_export / lzyb_peer_import) */
    int (*slot_export)(lzyb_handle_t h, lzyb_ipc_desc_t* out);
    int (*peer_import)(const lzyb_ipc_desc_t* desc, int onto_gpu, lzyb_handle_t* out);
    /* optional (may be NULL): declare that a committed output frame has exactly one consumer, which may then take the
     * frame over instead of copying it (HBM backend: lzyb_slot_set_movable) */
    int (*slot_set_movable)(lzyb_handle_t h, int movable);
    /* optional (may be NULL → slot_info): sizes / state / dtype / shape without waiting for a digest that is still being
     * computed (HBM backend: lzyb_slot_meta) */
    int (*slot_meta)(lzyb_handle_t h, lzyb_slot_info_t* out);
} lzyb_backend_ops_t;
int lzyb_slots_set_backend(const lzyb_backend_ops_t* ops); /* NULL = the HBM backend */

/* How the slot state machines reach the channel manager (LzyChannelManager.Bind / Unbind / TransferCompleted /
 * TransferFailed, channel-manager.proto:13-58).  Default: the in-process matcher (lzyb_cm_* above, the `cm` handle is
 * a lzyb_cm_create handle).  A worker whose channel manager is another process installs a table that performs the
 * RPCs; `cm` is then an opaque cookie chosen by the installer and passed through unchanged.  Same signatures, same
 * result codes (a gRPC status maps to its LZYB_E_* image) and lzyb_set_last_error for the description. */
typedef struct lzyb_cm_client_ops {
    int (*bind)(lzyb_handle_t cm, const char* idempotency_key, const char* channel_id, const char* peer_id,
                const char* execution_id, const char* peer_url, int role, lzyb_bind_result_t* out);
    int (*unbind)(lzyb_handle_t cm, const char* channel_id, const char* peer_id);
    int (*transfer_completed)(lzyb_handle_t cm, const char* idempotency_key, const char* channel_id, const char* transfer_id);
    int (*transfer_failed)(lzyb_handle_t cm, const char* idempotency_key, const char* channel_id, const char* transfer_id,
                           const char* description, lzyb_bind_result_t* out);
} lzyb_cm_client_ops_t;
int lzyb_slots_set_cm_client(const lzyb_cm_client_ops_t* ops); /* NULL = the in-process matcher */

/* Slot peers that live in another worker process.  A consumer whose peer is not in this process's registry asks the
 * resolver for the peer's descriptor — given the peer's URL (what that worker put into BindRequest.peer_url, e.g.
 * "cuda+ipc://host:port?gpu=3") and its id — and maps it with the backend's peer_import; the transfer then pulls from
 * the mapping exactly as from a local peer.  The other side of that question is lzyb_slots_describe: the descriptor of
 * a producer slot registered on this worker (NOT_FOUND otherwise, as LzySlotsApi.Read answers).  On the wire the pair is
 * one additive RPC next to LzySlotsApi (lzy_b200/grpc_api: LzybPeerApi.Describe); nothing in the reference's protos changes. */
typedef int (*lzyb_peer_resolver_t)(const char* peer_url, const char* peer_id, lzyb_ipc_desc_t* out);
int lzyb_slots_set_peer_resolver(lzyb_peer_resolver_t resolver); /* NULL = peers must be local (default) */
int lzyb_slots_describe(lzyb_handle_t slots, const char* peer_id, lzyb_ipc_desc_t* out);

/* The same two seams over a unix socket, for this repository's own processes on ONE box (csrc/host/local_rpc.cpp).  gRPC stays
 * the wire contract towards stock peers; between a worker process per GPU and the channel manager of one box a task's six
 * control calls (Bind, Describe, TransferCompleted for an input; Bind, upload, TransferCompleted for an output) cost
 * 0.5–0.9 ms each through two interpreters — here they are length-prefixed frames served by threads of this library, with the
 * same arguments, result codes and lzyb_last_error texts.
 *   lzyb_cm_serve_local(cm, path)          serve Bind / Unbind / TransferCompleted / TransferFailed of matcher `cm` on `path`
 *   lzyb_slots_connect_cm_local(path, &c)  route this process's slot state machines to that socket (installs the client
 *                                          table of lzyb_slots_set_cm_client); `c` goes where a matcher handle would
 *                                          (lzyb_slots_create).  LZYB_E_UNAVAILABLE if nobody listens there
 *   lzyb_slots_serve_local(slots, path)    serve Describe(peer_id) of worker `slots`; a worker that does puts
 *                                          "uds=<path>" into its URL's query ("cuda+ipc://host:port?gpu=3&uds=/run/w3.sock")
 *                                          and consumers of this box resolve its peers over the socket instead of
 *                                          through the resolver callback
 * The matcher's deferred StartTransfer keeps travelling as LzySlotsApi.StartTransfer. */
int lzyb_cm_serve_local(lzyb_handle_t cm, const char* socket_path);
int lzyb_cm_stop_local(lzyb_handle_t cm);
int lzyb_slots_connect_cm_local(const char* socket_path, lzyb_handle_t* cm_cookie);
int lzyb_slots_disconnect_cm_local(void);
int lzyb_slots_serve_local(lzyb_handle_t slots, const char* socket_path);
int lzyb_slots_stop_local(lzyb_handle_t slots);

#define LZYB_DIR_INPUT 1
#define LZYB_DIR_OUTPUT 2

/* Slots(...): one per worker.  `api_url` is what the worker puts in BindRequest.peer_url. */
int lzyb_slots_create(lzyb_handle_t cm, int gpu, const char* api_url, lzyb_handle_t* out);
int lzyb_slots_free(lzyb_handle_t slots);
/* bytes per transferChunkTo call (0 = the whole remaining payload in one launch; the reference's
 * gRPC chunk is 2 MiB, OutputSlot.java:241) */
int lzyb_slots_set_chunk_bytes(lzyb_handle_t slots, uint64_t chunk_bytes);
/* the same for a transfer whose source frame is on ANOTHER GPU (or in another process), when chunk_bytes is 0.  A transfer
 * kernel is a persistent grid: a kernel launched next to it on either GPU — an op's pack — waits for it to drain (measured,
 * tools/overlap_probe.py: a 0.35 ms pack launched into a 1 GiB pull ends 1.2 ms later).  In slices the other kernel gets the
 * GPU at the next slice boundary (0.5 ms with 128 MiB slices) for ≈ 8 % on the transfer itself.  0 (default) = one launch. */
int lzyb_slots_set_peer_chunk_bytes(lzyb_handle_t slots, uint64_t chunk_bytes);
/* Slots.context(requestId, executionId, taskId, slots, slotToChannelMapping) */
int lzyb_ctx_create(lzyb_handle_t slots, const char* execution_id, const char* task_id, lzyb_handle_t* out);
/* one LMS.Slot + its channel assignment (model/.../slot.proto:10-30, operation.proto:31-37) */
int lzyb_ctx_add_slot(lzyb_handle_t ctx, const char* slot_name, const char* channel_id, int direction);
/* the same with hints.  LZYB_SLOTF_SOLE_CONSUMER (output slots): the graph has exactly ONE consumer of this output, on
 * this worker's GPU, and nobody else (no client proxy, no cached result, no fan-out) can ask for it — what the
 * reference's channel manager could read off its peer table (channel-manager/.../services/SlotsService.java:440-519).
 * The slot's frame is then not referenced by the `hbm://` store ("uploading" to a store nobody reads would only pin
 * the frame) and is marked movable, so the consumer's InputSlot takes the frame over instead of copying it
 * (FileInputBackend.toOutput's ownership hand-over, FileInputBackend.java:38-45). */
#define LZYB_SLOTF_SOLE_CONSUMER 1u
int lzyb_ctx_add_slot_flags(lzyb_handle_t ctx, const char* slot_name, const char* channel_id, int direction, uint32_t flags);
/* beforeExecution(): build the slots; block until every input is READY (or a slot failed) */
int lzyb_ctx_before_execution(lzyb_handle_t ctx, int timeout_ms);
/* non-blocking half of beforeExecution(): create the slots and let the inputs bind/download in the
 * background (downstream prefetch); a later before_execution() only waits */
int lzyb_ctx_prefetch(lzyb_handle_t ctx);
/* the delivered frame of an input slot (after before_execution) */
int lzyb_ctx_input_handle(lzyb_handle_t ctx, const char* slot_name, lzyb_handle_t* slot_out);
/* the op opens its output: allocate the frame to pack / write into (the fifo of OutputPipeBackend) */
int lzyb_ctx_output_open(lzyb_handle_t ctx, const char* slot_name, uint64_t nbytes, lzyb_handle_t* slot_out);
/* the op closed its output: data is complete (waitCompleted returns) → register → Bind(PRODUCER) */
int lzyb_ctx_output_done(lzyb_handle_t ctx, const char* slot_name);
/* afterExecution(): block until every output is bound (and uploaded when the channel has a storage consumer) */
int lzyb_ctx_after_execution(lzyb_handle_t ctx, int timeout_ms);
int lzyb_ctx_close(lzyb_handle_t ctx);
/* input slot state for tests/metrics: enum below */
enum lzyb_input_state { LZYB_IS_BINDING = 0, LZYB_IS_WAITING_FOR_PEER = 1, LZYB_IS_DOWNLOADING = 2, LZYB_IS_READY = 3, LZYB_IS_CLOSED = 4 };
int lzyb_ctx_input_state(lzyb_handle_t ctx, const char* slot_name);

/* LzySlotsApi server side (lzy/slots-api/src/main/proto/ai/lzy/v1/slots/slots-api.proto:13-47) */
int lzyb_slots_start_transfer(lzyb_handle_t slots, const char* slot_id, const lzyb_peer_desc_t* peer, const char* transfer_id);
/* Read(peer_id, offset): one chunk per call (≤ len bytes); LZYB_EOF = the EOS control message;
 * LZYB_E_NOT_FOUND unknown slot; LZYB_E_UNIMPLEMENTED input slot */
int64_t lzyb_slots_read(lzyb_handle_t slots, const char* peer_id, uint64_t offset, void* dst_host, uint64_t len);

/* the channel manager's deferred actions need somebody to execute them (ActionScheduler's executor):
 * start a thread that polls `cm` and delivers StartTransfer to the registered Slots instances */
int lzyb_cm_pump_start(lzyb_handle_t cm);
int lzyb_cm_pump_stop(lzyb_handle_t cm);

/* storage peers: on one box the reference's S3 rendezvous/backup store becomes a host-memory blob
 * store keyed by the storage URI (StorageOutputTransfer / StorageInputTransfer,
 * lzy/slots/.../transfers/StorageOutputTransfer.java:30-40, StorageInputTransfer.java:32-41) */
int lzyb_blob_put(const char* uri, const void* data, uint64_t len);
int64_t lzyb_blob_size(const char* uri);                       /* LZYB_E_NOT_FOUND if absent */
int64_t lzyb_blob_get(const char* uri, void* dst, uint64_t cap);
int lzyb_blob_delete(const char* uri);
/* HBM frame store: storage URIs with the scheme "hbm://" name frames that stay resident in HBM.
 * A producer bound to a channel whose storage consumer is hbm://… "uploads" by handing a reference
 * of its frame to the store (no copy); a consumer whose peer is such a storage producer pulls from
 * that frame with the same NVLink/HBM kernel as from a slot peer.  put retains, drop releases. */
int lzyb_frame_store_put(const char* uri, lzyb_handle_t slot);
int lzyb_frame_store_get(const char* uri, lzyb_handle_t* slot_out);   /* LZYB_E_NOT_FOUND if absent */
int lzyb_frame_store_drop(const char* uri);
/* drop every frame whose URI starts with `prefix` (NULL/"" = all); returns how many were dropped */
int lzyb_frame_store_clear(const char* prefix);
int64_t lzyb_frame_store_bytes(void);
/* Spill tier below the frame store — the part of the reference's storage peers that outlives HBM residency
 * (BACKUP producers, PeerDao.java:44-53; cached op results under stable URIs, core/workflow.py:247-281).
 * set_spill: directory for spilled frames (NULL / "" switches the tier off) and the GPU a frame is loaded
 * onto when nobody says otherwise.  spill: write the resident frame at `uri` to <dir>/<escaped uri> as
 * [128-byte lzyb_frame_header_t][payload] (temp file + rename; keeps the frame resident).  get falls back to
 * the tier: a URI that is not resident but has a spill file is read into a fresh frame — digest verified
 * against the header while it lands, LZYB_E_PEER if the file is corrupt — registered and returned; load does
 * the same onto a chosen GPU.  spilled: 1 / 0.  unspill: delete the file.  A file without the frame magic is a
 * foreign object: its bytes become a DT_BYTES payload whose digest is computed on the way in. */
int lzyb_frame_store_set_spill(const char* dir, int default_gpu);
int lzyb_frame_store_spill(const char* uri);
int lzyb_frame_store_load(const char* uri, int gpu, lzyb_handle_t* slot_out);
int lzyb_frame_store_spilled(const char* uri);
int lzyb_frame_store_resident(const char* uri);   /* 1 / 0; never loads */
int lzyb_frame_store_unspill(const char* uri);
/* Per-transfer records of this worker's input slots (most recent 4096), as a JSON array of
 *   {"slot_id","channel_id","peer_id","peer_kind":"slot"|"storage","bytes","ns","failovers","ok","error"}
 * — bytes delivered, wall time from the first peer to READY (or failure), the peer the bytes finally came from and
 * how many producers were tried.  The reference keeps no data-plane metric at all.  Returns the bytes
 * needed including the terminator (may exceed cap); clear != 0 empties the log. */
int64_t lzyb_slots_transfer_log(lzyb_handle_t slots, char* json_out, uint64_t cap, int clear);
/* republish delivered inputs as "<slot>-out" producers (InputSlot.java:160-168); default on.  The
 * Python runtime turns it off so that an op owns its input frame and may mutate it in place. */
int lzyb_slots_set_republish(lzyb_handle_t slots, int on);
/* observability: how many OS threads the slot state machines have created so far in this process (they are recycled:
 * a finished slot's thread serves the next slot, so this stays near the peak number of slots blocked at once) */
uint64_t lzyb_slots_threads_spawned(void);

/* ---- graph dispatch -------------------------------------------------------------------------- */
/* Replaces BfsGraphProcessor + Algorithms (lzy/graph-executor/.../exec/BfsGraphProcessor.java:44-258,
 * algo/Algorithms.java:7-148) and the dependency build of graph-executor-2
 * (lzy/graph-executor-2/.../algo/Algorithms.java:10-73).  Tasks are 0..n-1; an edge (a, b) says
 * "a channel produced by task a is consumed by task b" (DirectChannelChecker.java:24-32). */
enum lzyb_graph_status { LZYB_G_WAITING = 0, LZYB_G_EXECUTING = 1, LZYB_G_COMPLETED = 2, LZYB_G_FAILED = 3 };
enum lzyb_task_status { LZYB_T_WAITING = 0, LZYB_T_EXECUTING = 1, LZYB_T_SUCCESS = 2, LZYB_T_ERROR = 3 };

int lzyb_graph_create(int n_tasks, const int32_t* edges_from, const int32_t* edges_to, int n_edges, int allow_cycles,
                      lzyb_handle_t* out);
int lzyb_graph_free(lzyb_handle_t g);
/* One dispatch step (BfsGraphProcessor.exec): returns how many tasks became runnable and writes
 * their ids; they are marked EXECUTING.  A task is runnable when every producer of each of its
 * input channels has finished successfully; members of one strongly connected component start
 * together.  No sleep between steps (the reference sleeps 1 s, BfsGraphProcessor.java:148). */
int lzyb_graph_next(lzyb_handle_t g, int32_t* runnable_out, int cap);
int lzyb_graph_task_done(lzyb_handle_t g, int32_t task, int ok, const char* error_description);
/* tasks EXECUTING now (GraphExecutionState.currentExecutionGroup) */
int lzyb_graph_executing(lzyb_handle_t g, int32_t* out, int cap);
int lzyb_graph_status(lzyb_handle_t g);
int lzyb_graph_task_status(lzyb_handle_t g, int32_t task);
/* stop(graph, description): FAILED; returns the tasks that must be killed (were EXECUTING) */
int lzyb_graph_stop(lzyb_handle_t g, const char* description, int32_t* kill_out, int cap);
int lzyb_graph_error(lzyb_handle_t g, char* description_out, uint64_t cap, int32_t* failed_task_out);
/* LocalRuntime.exec order (pylzy/lzy/api/v1/local/runtime.py:48-85) is available for comparison */
int lzyb_graph_topo_order(lzyb_handle_t g, int32_t* order_out, int cap);
/* graph-executor-2 Algorithms.buildTaskDependents cycle check: 1 if the task graph has a cycle */
int lzyb_graph_has_cycle(lzyb_handle_t g);

#ifdef __cplusplus
}
#endif
#endif /* LZY_B200_HOST_H */
