// slots_runtime.cpp — the worker-side slot state machines, re-built over HBM frames.
//
// What each piece replaces (paths relative to lzy/slots/src/main/java/ai/lzy/slots/):
//   SlotsInstance     Slots.java:44-106              (the gRPC server becomes an in-process registry; the RPC
//                                                      entry points lzyb_slots_start_transfer / lzyb_slots_read stay)
//   SlotsServiceImpl  SlotsService.java:21-69        register / unregister / read / startTransfer routing + error codes
//   ExecCtx           SlotsExecutionContext.java:45-107
//   InputSlotImpl     InputSlot.java:127-295         one thread per input slot, same state machine, same failover loop
//   OutputSlotImpl    OutputSlot.java:135-273        prepare thread: waitCompleted → register → bind → storage upload
//   make_input_transfer / storage upload              transfers/TransferFactory.java:19-33, SlotInputTransfer.java:43-99,
//                                                      StorageInputTransfer.java:32-41, StorageOutputTransfer.java:30-40
// The byte movers are gone: a "transfer" is a CUDA pull kernel over NVLink/HBM issued through the
// backend table (default: the lzyb_* data plane), verified against the producer's digest.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../../include/lzy_b200_host.h"
#include "host_common.h"
#include "local_rpc.h"

namespace {

using lzyb_host::fail;
using lzyb_host::scopy;

// ---- backend seam ------------------------------------------------------------------------------
const lzyb_backend_ops_t kHbmBackend = {
    lzyb_slot_create, lzyb_input_open,      lzyb_slot_commit,         lzyb_slot_read,       lzyb_slot_info,
    lzyb_slot_find,   lzyb_slot_close,      lzyb_transfer_chunk,      lzyb_input_finish,    lzyb_input_write_host,
    lzyb_input_finish_expect, lzyb_input_to_output, lzyb_input_reset, lzyb_slot_retain, lzyb_slot_set_meta,
    lzyb_slot_export, lzyb_peer_import, lzyb_slot_set_movable, lzyb_slot_meta,
};
std::atomic<const lzyb_backend_ops_t*> g_ops{&kHbmBackend};
lzyb_backend_ops_t g_custom_ops;
const lzyb_backend_ops_t& ops() { return *g_ops.load(); }
// sizes, state, dtype, shape of a slot — never waits for a digest that is still being computed
int slot_sizes(lzyb_handle_t h, lzyb_slot_info_t* out) { return ops().slot_meta ? ops().slot_meta(h, out) : ops().slot_info(h, out); }

// ---- the matcher as the slots see it: in-process calls by default, a client of a remote channel manager otherwise ----
const lzyb_cm_client_ops_t kLocalMatcher = {lzyb_cm_bind, lzyb_cm_unbind, lzyb_cm_transfer_completed, lzyb_cm_transfer_failed};
std::atomic<const lzyb_cm_client_ops_t*> g_cmc{&kLocalMatcher};
lzyb_cm_client_ops_t g_custom_cmc;
const lzyb_cm_client_ops_t& cmc() { return *g_cmc.load(); }

// ---- peers in other worker processes ---------------------------------------------------------------------------------
std::atomic<lzyb_peer_resolver_t> g_peer_resolver{nullptr};

// ---- host blob store (storage peers) -----------------------------------------------------------------
struct Blob {
    std::vector<uint8_t> data;
    bool has_digest = false;   // known when the object was uploaded from a committed slot
    uint64_t digest[2] = {0, 0};
};
std::mutex g_blob_mu;
std::map<std::string, std::shared_ptr<Blob>> g_blobs;

std::shared_ptr<Blob> blob_find(const std::string& uri) {
    std::lock_guard<std::mutex> lk(g_blob_mu);
    auto it = g_blobs.find(uri);
    return it == g_blobs.end() ? nullptr : it->second;
}

// ---- HBM frame store ("hbm://" storage peers) -----------------------------------------------------------
std::mutex g_frames_mu;
std::map<std::string, lzyb_handle_t> g_frames;
bool is_hbm_uri(const char* uri) { return uri && strncmp(uri, "hbm://", 6) == 0; }
lzyb_handle_t frame_resident(const std::string& uri) {
    std::lock_guard<std::mutex> lk(g_frames_mu);
    auto it = g_frames.find(uri);
    return it == g_frames.end() ? 0 : it->second;
}

// ---- spill tier: frames that left HBM (or another process' HBM) stay fetchable from files ------------------
std::mutex g_spill_mu;       // guards the configuration
std::string g_spill_dir;
int g_spill_gpu = 0;
std::mutex g_spill_load_mu;  // one load at a time: two consumers of one spilled URI must share the frame

std::string spill_dir() {
    std::lock_guard<std::mutex> lk(g_spill_mu);
    return g_spill_dir;
}

// file name of a URI: readable, collision-free, bounded ('%XX' escapes; long names end in a 128-bit FNV-1a tag)
std::string spill_path(const std::string& dir, const std::string& uri) {
    static const char* hex = "0123456789abcdef";
    std::string name;
    for (unsigned char c : uri) {
        if (isalnum(c) || c == '.' || c == '_' || c == '-')
            name.push_back((char)c);
        else {
            name.push_back('%');
            name.push_back(hex[c >> 4]);
            name.push_back(hex[c & 15]);
        }
    }
    if (name.size() > 200) {
        uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x84222325cbf29ce4ull;
        for (unsigned char c : uri) {
            h1 = (h1 ^ c) * 0x100000001b3ull;
            h2 = (h2 ^ (c + 0x9eull)) * 0x100000001b3ull;
        }
        char tag[40];
        snprintf(tag, sizeof(tag), "-%016llx%016llx", (unsigned long long)h1, (unsigned long long)h2);
        name = name.substr(0, 160) + tag;
    }
    return dir + "/" + name + ".lzyb";
}

bool read_fully(int fd, void* dst, uint64_t n) {
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) {
        ssize_t r = ::read(fd, p, n > (1ull << 30) ? (1ull << 30) : n);
        if (r <= 0) return false;
        p += r;
        n -= (uint64_t)r;
    }
    return true;
}

bool write_fully(int fd, const void* src, uint64_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(src);
    while (n) {
        ssize_t w = ::write(fd, p, n > (1ull << 30) ? (1ull << 30) : n);
        if (w <= 0) return false;
        p += w;
        n -= (uint64_t)w;
    }
    return true;
}

constexpr uint64_t SPILL_STEP = 64ull << 20;

int spill_write(const std::string& path, lzyb_handle_t slot) {
    lzyb_slot_info_t info;
    int rc = ops().slot_info(slot, &info);
    if (rc != LZYB_OK) return rc;
    if (info.state != LZYB_ST_READY) return fail(LZYB_E_STATE, "only a complete frame can be spilled");
    lzyb_frame_header_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    memcpy(hdr.magic, "LZYB200", 8);
    hdr.version = LZYB_ABI_VERSION;
    hdr.dtype = info.dtype;
    hdr.nbytes = info.nbytes;
    hdr.ndim = info.ndim;
    hdr.flags = 1;
    memcpy(hdr.shape, info.shape, sizeof(hdr.shape));
    hdr.digest[0] = info.digest[0];
    hdr.digest[1] = info.digest[1];
    static std::atomic<uint64_t> tmp_seq{0};  // two writers of one URI (threads or processes) never share a temp file
    const std::string tmp = path + ".tmp." + std::to_string((long long)getpid()) + "." + std::to_string(++tmp_seq);
    int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(LZYB_E_LOCAL, "cannot create %s: %s", tmp.c_str(), strerror(errno));
    bool ok = write_fully(fd, &hdr, sizeof(hdr));
    std::vector<uint8_t> buf(std::min<uint64_t>(std::max<uint64_t>(info.nbytes, 1), SPILL_STEP));
    for (uint64_t off = 0; ok && off < info.nbytes;) {
        int64_t n = ops().slot_read(slot, off, buf.data(), std::min<uint64_t>(info.nbytes - off, buf.size()));
        if (n <= 0) {
            ok = false;
            break;
        }
        ok = write_fully(fd, buf.data(), (uint64_t)n);
        off += (uint64_t)n;
    }
    ok = ok && ::fsync(fd) == 0;
    ::close(fd);
    if (!ok || ::rename(tmp.c_str(), path.c_str()) != 0) {
        ::unlink(tmp.c_str());
        return fail(LZYB_E_LOCAL, "cannot write %s", path.c_str());
    }
    return LZYB_OK;
}

// read a spill file into a fresh frame on `gpu`; the digest computed while the bytes land is checked against
// the header's (a foreign file has no header: its digest becomes the published one)
int spill_read(const std::string& path, const std::string& uri, int gpu, lzyb_handle_t* out) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return fail(LZYB_E_NOT_FOUND, "no frame at %s", uri.c_str());
    struct stat st;
    if (::fstat(fd, &st) != 0) {
        ::close(fd);
        return fail(LZYB_E_LOCAL, "cannot stat %s", path.c_str());
    }
    lzyb_frame_header_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    const bool magic = (uint64_t)st.st_size >= sizeof(hdr) && read_fully(fd, &hdr, sizeof(hdr)) && memcmp(hdr.magic, "LZYB200", 8) == 0;
    const bool framed = magic && hdr.nbytes + sizeof(hdr) == (uint64_t)st.st_size && hdr.ndim <= LZYB_MAX_DIMS;
    if (magic && !framed) {
        // a frame file that contradicts itself (truncated payload, impossible rank) is rotten, not foreign: serving
        // header + partial payload as bytes would hand consumers wrong data under a valid-looking digest
        ::close(fd);
        return fail(LZYB_E_PEER, "corrupt frame file for %s: header says %llu payload bytes, ndim %u; file has %llu bytes", uri.c_str(),
                    (unsigned long long)hdr.nbytes, hdr.ndim, (unsigned long long)st.st_size);
    }
    uint64_t nbytes = framed ? hdr.nbytes : (uint64_t)st.st_size;
    if (!framed) ::lseek(fd, 0, SEEK_SET);
    static std::atomic<uint64_t> seq{0};
    const std::string slot_id = "spill/" + std::to_string(++seq) + "/" + uri;
    lzyb_handle_t h = 0;
    int rc = ops().input_open(slot_id.c_str(), gpu, nbytes, &h);
    if (rc != LZYB_OK) {
        ::close(fd);
        return rc;
    }
    if (framed && ops().slot_set_meta) rc = ops().slot_set_meta(h, hdr.dtype, hdr.ndim, hdr.shape);
    std::vector<uint8_t> buf(std::min<uint64_t>(std::max<uint64_t>(nbytes, 1), SPILL_STEP));
    for (uint64_t off = 0; rc == LZYB_OK && off < nbytes;) {
        const uint64_t step = std::min<uint64_t>(nbytes - off, buf.size());
        if (!read_fully(fd, buf.data(), step)) {
            rc = fail(LZYB_E_LOCAL, "short read from %s", path.c_str());
            break;
        }
        int64_t n = ops().input_write_host(h, off, buf.data(), step, nullptr);
        if (n < 0) {
            rc = (int)n;
            break;
        }
        off += (uint64_t)n;
    }
    ::close(fd);
    if (rc == LZYB_OK) rc = ops().input_finish_expect(h, framed ? hdr.digest : nullptr, nullptr);
    if (rc != LZYB_OK) {
        const std::string why = lzyb_last_error();
        ops().slot_close(h);
        return fail(rc, "spilled frame %s is unusable: %s", uri.c_str(), why.c_str());
    }
    *out = h;
    return LZYB_OK;
}

// ---- "file://" storage peers: the box-local durable storage (what FsStorageClient is to the reference's local
// runtime, pylzy/lzy/storage/async_/fs.py) — same file format as the spill tier, reachable from every process ----------
bool is_file_uri(const char* uri) { return uri && strncmp(uri, "file://", 7) == 0; }
std::string file_path_of(const std::string& uri) { return uri.substr(7); }

void make_parent_dirs(const std::string& path) {
    for (size_t i = 1; i < path.size(); ++i)
        if (path[i] == '/') ::mkdir(path.substr(0, i).c_str(), 0755);
}

// what a stored file says about itself without reading the payload
struct StoredFile {
    bool framed = false;
    bool corrupt = false;  // carries the frame magic but its header contradicts the file
    lzyb_frame_header_t hdr{};
    uint64_t nbytes = 0;
};

bool stat_stored_file(const std::string& path, StoredFile* out) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    bool ok = ::fstat(fd, &st) == 0;
    if (ok) {
        const bool magic = (uint64_t)st.st_size >= sizeof(out->hdr) && read_fully(fd, &out->hdr, sizeof(out->hdr)) &&
                           memcmp(out->hdr.magic, "LZYB200", 8) == 0;
        out->framed = magic && out->hdr.nbytes + sizeof(out->hdr) == (uint64_t)st.st_size && out->hdr.ndim <= LZYB_MAX_DIMS;
        out->corrupt = magic && !out->framed;
        out->nbytes = out->framed ? out->hdr.nbytes : (uint64_t)st.st_size;
    }
    ::close(fd);
    return ok;
}

// resident frame, or — when the spill tier holds the URI — a frame loaded from it (registered, so that every
// later consumer shares it).  LZYB_E_NOT_FOUND if neither; the loader's error if the file exists but is unusable.
int frame_lookup(const std::string& uri, int gpu, lzyb_handle_t* out) {
    if ((*out = frame_resident(uri))) return LZYB_OK;
    const std::string dir = spill_dir();
    if (dir.empty()) return LZYB_E_NOT_FOUND;
    const std::string path = spill_path(dir, uri);
    if (::access(path.c_str(), R_OK) != 0) return LZYB_E_NOT_FOUND;
    std::lock_guard<std::mutex> lk(g_spill_load_mu);
    if ((*out = frame_resident(uri))) return LZYB_OK;
    if (gpu < 0) {
        std::lock_guard<std::mutex> cfg(g_spill_mu);
        gpu = g_spill_gpu;
    }
    int rc = spill_read(path, uri, gpu, out);
    if (rc != LZYB_OK) return rc;
    std::lock_guard<std::mutex> fl(g_frames_mu);
    g_frames[uri] = *out;  // the store owns the loader's reference
    return LZYB_OK;
}

lzyb_handle_t frame_find(const std::string& uri) {
    lzyb_handle_t h = 0;
    return frame_lookup(uri, -1, &h) == LZYB_OK ? h : 0;
}

// ---- a tiny future -------------------------------------------------------------------------------------
struct Completion {
    std::mutex mu;
    std::condition_variable cv;
    int state = 0;  // 0 pending, 1 ok, 2 failed
    std::string error;
    void complete() {
        std::lock_guard<std::mutex> lk(mu);
        if (state == 0) state = 1;
        cv.notify_all();
    }
    void fail_with(const std::string& e) {
        std::lock_guard<std::mutex> lk(mu);
        if (state == 0) {
            state = 2;
            error = e;
        }
        cv.notify_all();
    }
    bool done() {
        std::lock_guard<std::mutex> lk(mu);
        return state != 0;
    }
    // 1 ok, 2 failed, 0 timeout
    int wait(int timeout_ms) {
        std::unique_lock<std::mutex> lk(mu);
        if (timeout_ms < 0)
            cv.wait(lk, [&] { return state != 0; });
        else
            cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return state != 0; });
        return state;
    }
};

// ---- slot threads ----------------------------------------------------------------------------------------
// Every slot runs its state machine on a thread of its own, as in the reference (InputSlot and OutputSlot.PrepareThread
// are Threads): it may block for as long as it likes — waiting for a peer, for the op to finish its output.  The
// threads are recycled: one that has finished a slot waits a little for the next slot instead of being created anew
// (two creations + joins per op otherwise).  The pool grows on demand and is never destroyed: its idle threads may
// outlive main().
class SlotThreads {
public:
    struct Ticket {
        std::mutex mu;
        std::condition_variable cv;
        bool done = false;
        void wait() {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return done; });
        }
    };
    using TicketPtr = std::shared_ptr<Ticket>;

    TicketPtr run(std::function<void()> body) {
        auto ticket = std::make_shared<Ticket>();
        bool spawn;
        {
            std::lock_guard<std::mutex> lk(mu_);
            queue_.emplace_back(std::move(body), ticket);
            spawn = one_shot_ || queue_.size() > (size_t)idle_;
            if (!spawn) cv_.notify_one();
        }
        if (spawn) {
            ++spawned_;
            std::thread([this] { serve(); }).detach();
        }
        return ticket;
    }
    uint64_t spawned() const { return spawned_.load(); }

private:
    void serve() {
        for (;;) {
            std::pair<std::function<void()>, TicketPtr> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (queue_.empty()) {
                    if (one_shot_) return;
                    ++idle_;
                    bool got = cv_.wait_for(lk, std::chrono::seconds(2), [&] { return !queue_.empty(); });
                    --idle_;
                    if (!got) return;
                }
                job = std::move(queue_.front());
                queue_.pop_front();
            }
            job.first();
            job.first = nullptr;  // drop what the body captured (the slot) before announcing the end
            {
                std::lock_guard<std::mutex> lk(job.second->mu);
                job.second->done = true;
            }
            job.second->cv.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::pair<std::function<void()>, TicketPtr>> queue_;
    int idle_ = 0;
    std::atomic<uint64_t> spawned_{0};
    // LZYB_SLOT_THREADS=fresh: a new thread per slot, as before the pool (for A/B measurements)
    const bool one_shot_ = [] {
        const char* v = getenv("LZYB_SLOT_THREADS");
        return v && strcmp(v, "fresh") == 0;
    }();
};

SlotThreads& slot_threads() {
    static SlotThreads* pool = new SlotThreads;
    return *pool;
}

struct SlotsInstance;
struct ExecCtx;

// ---- Slot / SlotInternal ---------------------------------------------------------------------------
struct SlotBase {
    std::string id, channel_id;
    ExecCtx* ctx = nullptr;
    virtual ~SlotBase() = default;
    virtual bool is_input() const = 0;
    virtual int start_transfer(const lzyb_peer_desc_t& peer, const std::string& transfer_id) = 0;
    virtual int64_t read(uint64_t offset, void* dst, uint64_t len) = 0;
    virtual Completion& before() = 0;
    virtual Completion& after() = 0;
    virtual void close() = 0;
    virtual void join() = 0;
};

// ---- SlotsService --------------------------------------------------------------------------------------
struct SlotsServiceImpl {
    std::shared_mutex mu;
    std::unordered_map<std::string, std::shared_ptr<SlotBase>> slots;
    void reg(const std::shared_ptr<SlotBase>& s) {
        std::unique_lock<std::shared_mutex> lk(mu);
        slots[s->id] = s;
    }
    void unreg(const std::string& id) {
        std::unique_lock<std::shared_mutex> lk(mu);
        slots.erase(id);
    }
    std::shared_ptr<SlotBase> find(const std::string& id) {
        std::shared_lock<std::shared_mutex> lk(mu);
        auto it = slots.find(id);
        return it == slots.end() ? nullptr : it->second;
    }
};

// One finished (or failed) input transfer: what the reference never records ("no data-plane byte /
// throughput metric exists") — bytes, wall time, where they finally came from, how many producers were tried.
struct TransferRecord {
    std::string slot_id, channel_id, peer_id, error;
    int peer_kind = 0;
    uint64_t bytes = 0, ns = 0;
    uint64_t ns_setup = 0, ns_issue = 0, ns_wait = 0, ns_complete = 0;  // of the last attempt: open, launch, verdict, TransferCompleted
    int failovers = 0;
    bool ok = false;
};

struct SlotsInstance {
    std::mutex log_mu;
    std::deque<TransferRecord> transfer_log;  // the most recent 4096
    void record(TransferRecord r) {
        std::lock_guard<std::mutex> lk(log_mu);
        if (transfer_log.size() >= 4096) transfer_log.pop_front();
        transfer_log.push_back(std::move(r));
    }
    lzyb_handle_t handle = 0;
    lzyb_handle_t cm = 0;
    int gpu = 0;
    std::string api_url;
    std::atomic<uint64_t> chunk_bytes{0};
    std::atomic<uint64_t> peer_chunk_bytes{0};
    std::atomic<bool> republish{true};
    SlotsServiceImpl service;
    std::mutex mu;
    std::vector<std::shared_ptr<ExecCtx>> contexts;
    std::atomic<uint64_t> idk{0};
    std::string next_idk(const std::string& what) { return api_url + "/" + what + "/" + std::to_string(++idk); }
};

struct OutputSlotImpl;
struct InputSlotImpl;

struct ExecCtx : std::enable_shared_from_this<ExecCtx> {
    lzyb_handle_t handle = 0;
    SlotsInstance* inst = nullptr;
    std::string execution_id, task_id;
    struct Desc {
        std::string name, channel;
        int direction;
        uint32_t flags = 0;
    };
    std::vector<Desc> descs;
    std::mutex mu;
    std::vector<std::shared_ptr<SlotBase>> slots;
    std::map<std::string, std::shared_ptr<InputSlotImpl>> inputs;
    std::map<std::string, std::shared_ptr<OutputSlotImpl>> outputs;
    bool built = false, closed = false;
    void add(const std::shared_ptr<SlotBase>& s) {
        std::lock_guard<std::mutex> lk(mu);
        slots.push_back(s);
    }
};

// ---- OutputSlot ------------------------------------------------------------------------------------------
struct OutputSlotImpl : SlotBase, std::enable_shared_from_this<OutputSlotImpl> {
    SlotsInstance* inst;
    lzyb_handle_t backend = 0;  // HBM frame (OutputSlotBackend)
    bool own_backend = true;
    Completion before_, complete_;
    std::mutex mu;
    std::condition_variable cv;
    bool data_done = false, closed = false;
    bool sole_consumer = false;      // LZYB_SLOTF_SOLE_CONSUMER: the one consumer takes the frame over
    SlotThreads::TicketPtr prepare;  // PrepareThread
    bool is_input() const override { return false; }
    Completion& before() override { return before_; }
    Completion& after() override { return complete_; }

    int start_transfer(const lzyb_peer_desc_t&, const std::string&) override {
        return fail(LZYB_E_INTERNAL, "Cannot start transfer in output slot");  // NotImplementedException → INTERNAL
    }

    int64_t read(uint64_t offset, void* dst, uint64_t len) override {
        // OutputSlot.ReadThread: bytes [offset, end) in ≤ len chunks, then EOS (OutputSlot.java:238-265)
        int64_t n = ops().slot_read(backend, offset, dst, len);
        if (n < LZYB_EOF) return fail(LZYB_E_INTERNAL, "Error while reading from backend: %s", lzyb_last_error());
        return n;
    }

    void start() {
        before_.complete();
        auto self = shared_from_this();
        prepare = slot_threads().run([self] { self->run(); });
    }

    void signal_data_done() {
        std::lock_guard<std::mutex> lk(mu);
        data_done = true;
        cv.notify_all();
    }

    void run() {
        // PrepareThread.runImpl (OutputSlot.java:158-192)
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return data_done || closed; });  // backend.waitCompleted(): the op closed its output
            if (closed) {
                complete_.fail_with("Slot closed before ready");
                return;
            }
        }
        // (no digest asked for: under lzyb_set_lazy_commit the backend may then complete the slot in stream order)
        if (ops().slot_commit(backend, nullptr, nullptr) != LZYB_OK) {
            complete_.fail_with(std::string("Error while binding output slot: ") + lzyb_last_error());
            return;
        }
        if (sole_consumer && ops().slot_set_movable) ops().slot_set_movable(backend, 1);
        inst->service.reg(shared_from_this());  // enable requests from outside
        lzyb_bind_result_t res;
        int rc = cmc().bind(inst->cm, inst->next_idk("bind").c_str(), channel_id.c_str(), id.c_str(),
                              ctx ? ctx->execution_id.c_str() : "", inst->api_url.c_str(), LZYB_ROLE_PRODUCER, &res);
        if (rc != LZYB_OK) {
            complete_.fail_with(std::string("Error while binding output slot: ") + lzyb_last_error());
            return;
        }
        if (res.has_peer) {
            // bind answered with the channel's storage consumer: upload, then report
            bool ok = res.peer.kind == LZYB_PEER_STORAGE && upload_to_storage(res.peer.url);
            if (!ok) {
                std::string why = res.peer.kind == LZYB_PEER_STORAGE ? lzyb_last_error() : "Cannot create transfer for peer";
                lzyb_bind_result_t dummy;
                cmc().transfer_failed(inst->cm, inst->next_idk("fail").c_str(), channel_id.c_str(), res.transfer_id,
                                        why.c_str(), &dummy);
                complete_.fail_with("Error while transferring data to peer " + std::string(res.peer.peer_id) + ": " + why);
                return;
            }
            if (cmc().transfer_completed(inst->cm, inst->next_idk("done").c_str(), channel_id.c_str(), res.transfer_id) != LZYB_OK) {
                complete_.fail_with(std::string("Error while binding output slot: ") + lzyb_last_error());
                return;
            }
        }
        complete_.complete();  // "Slot is ready"
    }

    // StorageOutputTransfer.readFrom(backend.readFromOffset(0))
    bool upload_to_storage(const char* uri) {
        // a frame promised to its one consumer is not parked in a store nobody reads: the reference would hold the
        // store's second reference and force the copy this hint exists to avoid
        if (is_hbm_uri(uri) && sole_consumer && ops().slot_set_movable) return true;
        if (is_hbm_uri(uri)) return lzyb_frame_store_put(uri, backend) == LZYB_OK;  // stays in HBM: no copy
        if (is_file_uri(uri)) {  // durable and visible to every process of the box
            const std::string path = file_path_of(uri);
            make_parent_dirs(path);
            return spill_write(path, backend) == LZYB_OK;
        }
        lzyb_slot_info_t info;
        if (ops().slot_info(backend, &info) != LZYB_OK) return false;
        auto blob = std::make_shared<Blob>();
        blob->data.resize(info.nbytes);
        uint64_t off = 0;
        while (off < info.nbytes) {
            uint64_t want = std::min<uint64_t>(info.nbytes - off, 64ull << 20);
            int64_t n = ops().slot_read(backend, off, blob->data.data() + off, want);
            if (n < 0) return false;
            off += (uint64_t)n;
        }
        blob->has_digest = true;
        blob->digest[0] = info.digest[0];
        blob->digest[1] = info.digest[1];
        std::lock_guard<std::mutex> lk(g_blob_mu);
        g_blobs[uri] = blob;
        return true;
    }

    void close() override {
        // OutputSlot.close (OutputSlot.java:90-133): unregister, fail the future, unbind (errors ignored), close backend
        {
            std::lock_guard<std::mutex> lk(mu);
            if (closed) return;
            closed = true;
            cv.notify_all();
        }
        inst->service.unreg(id);
        complete_.fail_with("Slot closed before ready");
        cmc().unbind(inst->cm, channel_id.c_str(), id.c_str());
        if (backend && own_backend) ops().slot_close(backend);
    }

    void join() override {
        if (prepare) prepare->wait();
    }
};

// ---- InputTransfer variants ------------------------------------------------------------------------------
struct InputTransfer {
    virtual ~InputTransfer() = default;
    // bytes moved, LZYB_EOF, LZYB_E_PEER (ReadException) or another negative code (IOException)
    virtual int64_t transfer_chunk_to(lzyb_handle_t sink, uint64_t offset, uint64_t max_len) = 0;
    virtual int finish(lzyb_handle_t sink) = 0;
    virtual uint64_t nbytes() const = 0;
    virtual bool crosses_gpus(int /*sink_gpu*/) const { return false; }  // the source frame is on another GPU / in another process
};

// SlotInputTransfer: consumer-pull from the peer's frame (same GPU: HBM; other GPU: NVLink)
struct SlotPullTransfer : InputTransfer {
    lzyb_handle_t peer = 0;
    uint64_t size = 0;
    bool retained = false;  // a frame-store frame may be evicted to the spill tier while we read it: hold it
    int peer_gpu = -1;
    bool peer_remote = false;
    bool crosses_gpus(int sink_gpu) const override { return peer_remote || (peer_gpu >= 0 && peer_gpu != sink_gpu); }
    ~SlotPullTransfer() override {
        if (retained) ops().slot_close(peer);
    }
    int64_t transfer_chunk_to(lzyb_handle_t sink, uint64_t offset, uint64_t max_len) override {
        return ops().transfer_chunk(sink, peer, offset, max_len, nullptr);
    }
    int finish(lzyb_handle_t sink) override { return ops().input_finish(sink, peer, nullptr); }
    uint64_t nbytes() const override { return size; }
};

// StorageInputTransfer over a "file://" object: header (if any) → dtype / shape / digest, payload streamed into the sink
struct FileTransfer : InputTransfer {
    std::string path;
    StoredFile info;
    bool sent = false;
    int64_t transfer_chunk_to(lzyb_handle_t sink, uint64_t offset, uint64_t) override {
        if (sent || (offset >= info.nbytes && info.nbytes != 0)) return LZYB_EOF;
        if (offset != 0 && ops().input_reset(sink) != LZYB_OK) return LZYB_E_LOCAL;  // restart: the digest covers the whole object
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return fail(LZYB_E_PEER, "Error while reading from storage: cannot open %s", path.c_str());
        if (info.framed) {
            ::lseek(fd, sizeof(lzyb_frame_header_t), SEEK_SET);
            if (ops().slot_set_meta) ops().slot_set_meta(sink, info.hdr.dtype, info.hdr.ndim, info.hdr.shape);
        }
        std::vector<uint8_t> buf(std::min<uint64_t>(std::max<uint64_t>(info.nbytes, 1), SPILL_STEP));
        int64_t rc = (int64_t)info.nbytes;
        for (uint64_t off = 0; off < info.nbytes;) {
            const uint64_t step = std::min<uint64_t>(info.nbytes - off, buf.size());
            if (!read_fully(fd, buf.data(), step)) {
                rc = fail(LZYB_E_PEER, "Error while reading from storage: short read from %s", path.c_str());
                break;
            }
            int64_t n = ops().input_write_host(sink, off, buf.data(), step, nullptr);
            if (n < 0) {
                rc = n;
                break;
            }
            off += (uint64_t)n;
        }
        ::close(fd);
        if (rc >= 0) sent = true;
        return rc;
    }
    int finish(lzyb_handle_t sink) override { return ops().input_finish_expect(sink, info.framed ? info.hdr.digest : nullptr, nullptr); }
    uint64_t nbytes() const override { return info.nbytes; }
};

// StorageInputTransfer: the whole object in one go, then EOF (StorageInputTransfer.java:32-41)
struct BlobTransfer : InputTransfer {
    std::shared_ptr<Blob> blob;
    bool sent = false;
    int64_t transfer_chunk_to(lzyb_handle_t sink, uint64_t offset, uint64_t) override {
        if (sent || offset >= blob->data.size()) return LZYB_EOF;
        // a storage object cannot resume mid-way with a digest of its own: restart the sink at 0
        if (offset != 0) {
            if (ops().input_reset(sink) != LZYB_OK) return LZYB_E_LOCAL;
        }
        uint64_t off = 0;
        const uint64_t total = blob->data.size();
        while (off < total) {
            uint64_t step = std::min<uint64_t>(total - off, 64ull << 20);
            int64_t n = ops().input_write_host(sink, off, blob->data.data() + off, step, nullptr);
            if (n < 0) return n == LZYB_E_PEER ? LZYB_E_PEER : n;
            off += (uint64_t)n;
        }
        sent = true;
        return (int64_t)total;
    }
    int finish(lzyb_handle_t sink) override {
        // an object uploaded from a slot carries that slot's digest; a pre-existing object has none
        // and the digest computed while it landed becomes the published one
        return ops().input_finish_expect(sink, blob->has_digest ? blob->digest : nullptr, nullptr);
    }
    uint64_t nbytes() const override { return blob->data.size(); }
};

// ---- InputSlot ---------------------------------------------------------------------------------------------
struct InputSlotImpl : SlotBase, std::enable_shared_from_this<InputSlotImpl> {
    SlotsInstance* inst;
    lzyb_handle_t backend = 0;  // HBM frame (InputSlotBackend), opened once the peer (hence the size) is known
    std::atomic<int> state{LZYB_IS_BINDING};
    Completion ready_, after_;
    std::mutex mu;
    std::condition_variable cv;
    bool have_peer = false, closed = false, handed_over = false;
    lzyb_peer_desc_t peer{};
    std::string transfer_id;
    SlotThreads::TicketPtr thr;
    bool is_input() const override { return true; }
    Completion& before() override { return ready_; }
    Completion& after() override { return after_; }

    int64_t read(uint64_t, void*, uint64_t) override {
        return fail(LZYB_E_UNIMPLEMENTED, "unexpected `read` call on input slot %s", id.c_str());
    }

    int start_transfer(const lzyb_peer_desc_t& p, const std::string& tid) override {
        std::lock_guard<std::mutex> lk(mu);
        if (have_peer) return fail(LZYB_E_INTERNAL, "Transfer is already started in this slot");  // IllegalStateException → INTERNAL
        peer = p;
        transfer_id = tid;
        have_peer = true;
        cv.notify_all();
        return LZYB_OK;
    }

    void start() {
        after_.complete();
        auto self = shared_from_this();
        thr = slot_threads().run([self] {
            std::string err;
            if (!self->run(err)) self->ready_.fail_with(err.empty() ? "Closed before ready" : err);
            self->clear();
        });
    }

    std::unique_ptr<InputTransfer> make_transfer(const lzyb_peer_desc_t& p, std::string& err) {
        // TransferFactory.input(peer, offset)
        // a storage peer is served from the frame store whenever the frame is there (resident, or spilled: then it is
        // loaded straight onto this worker's GPU) — for "hbm://" URIs that is the only place, for the others a cache
        lzyb_handle_t cached = 0;
        const int lookup = p.kind == LZYB_PEER_STORAGE ? frame_lookup(p.url, inst->gpu, &cached) : LZYB_E_NOT_FOUND;
        if (p.kind == LZYB_PEER_STORAGE && (is_hbm_uri(p.url) || lookup == LZYB_OK)) {
            lzyb_handle_t h = cached;
            lzyb_slot_info_t info;
            int rc = lookup;
            if (rc != LZYB_OK || ops().slot_retain(h) != LZYB_OK) {
                err = std::string("Error while reading from storage: ") +
                      (rc == LZYB_OK || rc == LZYB_E_NOT_FOUND ? std::string("no frame at ") + p.url : std::string(lzyb_last_error()));
                return nullptr;
            }
            auto t = std::make_unique<SlotPullTransfer>();
            t->peer = h;
            t->retained = true;
            if (slot_sizes(h, &info) != LZYB_OK) {
                err = std::string("Error while reading from storage: ") + lzyb_last_error();
                return nullptr;
            }
            t->size = info.nbytes;
            return t;
        }
        if (p.kind == LZYB_PEER_STORAGE && is_file_uri(p.url)) {
            auto t = std::make_unique<FileTransfer>();
            t->path = file_path_of(p.url);
            if (!stat_stored_file(t->path, &t->info)) {
                err = std::string("Error while reading from storage: no object at ") + p.url;
                return nullptr;
            }
            if (t->info.corrupt) {  // a rotten frame file is a read fault of this peer, never foreign bytes
                err = std::string("Error while reading from storage: corrupt frame file at ") + p.url;
                return nullptr;
            }
            return t;
        }
        if (p.kind == LZYB_PEER_STORAGE) {
            auto blob = blob_find(p.url);
            if (!blob) {
                err = std::string("Error while reading from storage: no object at ") + p.url;
                return nullptr;
            }
            auto t = std::make_unique<BlobTransfer>();
            t->blob = blob;
            return t;
        }
        lzyb_handle_t h = 0;
        bool imported = false;
        if (ops().slot_find(p.peer_id, &h) != LZYB_OK) {
            // not in this process: another worker holds it.  Ask that worker for the frame's descriptor and map it
            // (TransferFactory.input would open a gRPC stream to peer_url here, SlotInputTransfer.java:43-60).
            const std::string local_err = lzyb_last_error();
            lzyb_peer_resolver_t resolve = g_peer_resolver.load();
            lzyb_ipc_desc_t desc;
            std::string peer_socket;  // a worker of this box that serves Describe on a unix socket says so in its URL
            const bool by_socket = lzyb_host::local_peer_socket(p.url, &peer_socket);
            if ((!resolve && !by_socket) || !ops().peer_import || p.url[0] == 0 || inst->api_url == p.url) {
                err = std::string("Cannot read from slot peer ") + p.peer_id + ": " + local_err;
                return nullptr;
            }
            memset(&desc, 0, sizeof(desc));
            const int found = by_socket ? lzyb_host::local_describe(peer_socket, p.peer_id, &desc) : resolve(p.url, p.peer_id, &desc);
            if (found != LZYB_OK || ops().peer_import(&desc, inst->gpu, &h) != LZYB_OK) {
                err = std::string("Cannot read from slot peer ") + p.peer_id + " at " + p.url + ": " + lzyb_last_error();
                return nullptr;
            }
            imported = true;
        }
        lzyb_slot_info_t info;
        if (slot_sizes(h, &info) != LZYB_OK) {
            err = std::string("Cannot read from slot peer ") + p.peer_id + ": " + lzyb_last_error();
            if (imported) ops().slot_close(h);
            return nullptr;
        }
        auto t = std::make_unique<SlotPullTransfer>();
        t->peer = h;
        t->retained = imported;  // the mapping belongs to this transfer: closed with it
        t->size = info.nbytes;
        t->peer_gpu = info.gpu;
        t->peer_remote = info.is_remote != 0;
        return t;
    }

    bool run(std::string& err) {
        // InputSlot.runImpl (InputSlot.java:127-169)
        inst->service.reg(shared_from_this());  // registering before bind
        lzyb_bind_result_t resp;
        if (is_closed()) return false;
        int rc = cmc().bind(inst->cm, inst->next_idk("bind").c_str(), channel_id.c_str(), id.c_str(),
                              ctx ? ctx->execution_id.c_str() : "", inst->api_url.c_str(), LZYB_ROLE_CONSUMER, &resp);
        if (rc != LZYB_OK) {
            err = std::string("Bind failed: ") + lzyb_last_error();
            return false;
        }
        lzyb_peer_desc_t cur;
        std::string tid;
        if (!resp.has_peer) {
            state = LZYB_IS_WAITING_FOR_PEER;
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return have_peer || closed; });
            if (closed) return false;
            cur = peer;
            tid = transfer_id;
        } else {
            std::lock_guard<std::mutex> lk(mu);
            if (closed) return false;
            have_peer = true;  // a late StartTransfer must now fail
            cur = resp.peer;
            tid = resp.transfer_id;
        }
        state = LZYB_IS_DOWNLOADING;
        inst->service.unreg(id);  // got peer, unregistering
        if (!download(cur, tid, err)) return false;
        state = LZYB_IS_READY;
        {
            std::lock_guard<std::mutex> lk(mu);
            handed_over = true;  // from here on the frame lives until the context is closed
        }
        // re-publish the delivered bytes as producer "<slotId>-out" on the same channel
        // (InputSlot.java:160-168).  The companion joins the context before `ready` completes so that
        // an afterExecution() issued right after beforeExecution() already sees it.
        if (inst->republish.load() && ops().input_to_output(backend) == LZYB_OK) {
            auto out = std::make_shared<OutputSlotImpl>();
            out->inst = inst;
            out->id = id + "-out";
            out->channel_id = channel_id;
            out->ctx = ctx;
            out->backend = backend;
            out->own_backend = false;  // the frame stays owned by this slot's context entry
            {
                std::lock_guard<std::mutex> lk(mu);
                handed_over = true;
            }
            if (ctx) ctx->add(out);
            out->start();
            out->signal_data_done();
        }
        ready_.complete();
        return true;
    }

    bool is_closed() {
        std::lock_guard<std::mutex> lk(mu);
        return closed;
    }

    bool download(lzyb_peer_desc_t cur, std::string tid, std::string& err) {
        TransferRecord rec;
        rec.slot_id = id;
        rec.channel_id = channel_id;
        const auto t0 = std::chrono::steady_clock::now();
        const bool ok = download_loop(cur, tid, err, rec);
        rec.ok = ok;
        rec.ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        if (!rec.ok) rec.error = err;
        if (backend) {
            lzyb_slot_info_t info;
            if (slot_sizes(backend, &info) == LZYB_OK) rec.bytes = rec.ok ? info.nbytes : info.delivered;
        }
        inst->record(std::move(rec));
        return ok;
    }

    bool download_loop(lzyb_peer_desc_t cur, std::string tid, std::string& err, TransferRecord& rec) {
        // InputSlot.download (InputSlot.java:223-295)
        uint64_t offset = 0;
        std::string terr;
        rec.peer_id = cur.peer_id;
        rec.peer_kind = cur.kind;
        using clk = std::chrono::steady_clock;
        auto since = [](clk::time_point t) { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now() - t).count(); };
        auto t_mark = clk::now();
        auto transfer = make_transfer(cur, terr);
        std::string peer_fail_reason = transfer ? "" : terr;
        bool need_new_peer = !transfer;
        while (true) {
            if (is_closed()) {
                err = "Closed before ready";
                return false;
            }
            if (need_new_peer) {
                lzyb_bind_result_t np;
                int rc = cmc().transfer_failed(inst->cm, inst->next_idk("failed").c_str(), channel_id.c_str(), tid.c_str(),
                                                 peer_fail_reason.c_str(), &np);
                if (rc != LZYB_OK || !np.has_peer) {
                    err = std::string("Cannot get data from any peer: ") + lzyb_last_error();
                    return false;
                }
                cur = np.peer;
                tid = np.transfer_id;
                rec.peer_id = cur.peer_id;
                rec.peer_kind = cur.kind;
                rec.failovers++;
                need_new_peer = false;
                peer_fail_reason.clear();
                transfer = make_transfer(cur, terr);
                if (!transfer) {
                    peer_fail_reason = terr;
                    need_new_peer = true;
                    continue;
                }
            }
            if (!backend) {  // FileInputBackend ctor: the sink exists once the size is known
                if (ops().input_open(id.c_str(), inst->gpu, transfer->nbytes(), &backend) != LZYB_OK) {
                    err = std::string("Error while reading into backend: ") + lzyb_last_error();
                    return false;
                }
            }
            lzyb_slot_info_t info;
            if (slot_sizes(backend, &info) == LZYB_OK) offset = info.delivered;  // reopening the channel: resume here
            if (offset == 0) {
                rec.ns_setup = since(t_mark);
                t_mark = clk::now();
            }
            uint64_t chunk = inst->chunk_bytes.load();
            if (!chunk && transfer->crosses_gpus(inst->gpu)) chunk = inst->peer_chunk_bytes.load();
            int64_t moved = transfer->transfer_chunk_to(backend, offset, chunk ? chunk : ~0ull);
            if (moved == LZYB_E_PEER) {  // ReadException: mark the peer bad, ask for another one
                peer_fail_reason = lzyb_last_error();
                need_new_peer = true;
                continue;
            }
            if (moved < LZYB_EOF) {  // error on the backend side: this slot fails
                err = std::string("Error while reading into backend: ") + lzyb_last_error();
                return false;
            }
            if (moved == LZYB_EOF) {
                rec.ns_issue = since(t_mark);
                t_mark = clk::now();
                int frc = transfer->finish(backend);
                rec.ns_wait = since(t_mark);
                t_mark = clk::now();
                if (frc == LZYB_E_PEER) {  // delivered bytes are not what the producer committed
                    peer_fail_reason = lzyb_last_error();
                    need_new_peer = true;
                    continue;
                }
                if (frc != LZYB_OK) {
                    err = std::string("Error while reading into backend: ") + lzyb_last_error();
                    return false;
                }
                if (cmc().transfer_completed(inst->cm, inst->next_idk("completed").c_str(), channel_id.c_str(), tid.c_str()) != LZYB_OK) {
                    err = std::string("TransferCompleted failed: ") + lzyb_last_error();
                    return false;
                }
                rec.ns_complete = since(t_mark);
                return true;
            }
            offset += (uint64_t)moved;
        }
    }

    void clear() {
        // InputSlot.clear (InputSlot.java:195-221): runs after success as well as after failure
        bool keep;
        {
            std::lock_guard<std::mutex> lk(mu);
            keep = handed_over;
            closed = true;
        }
        if (!keep && backend) {
            ops().slot_close(backend);
            backend = 0;
        }
        state = LZYB_IS_CLOSED;
        inst->service.unreg(id);
        ready_.fail_with("Closed before ready");
        cmc().unbind(inst->cm, channel_id.c_str(), id.c_str());  // error ignored
    }

    void close() override {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (closed && !handed_over) return;
            closed = true;
            cv.notify_all();
        }
    }

    void join() override {
        if (thr) thr->wait();
    }

    void release_frame() {
        if (backend) {
            ops().slot_close(backend);
            backend = 0;
        }
    }
};

// ---- registries ------------------------------------------------------------------------------------------
std::shared_mutex g_mu;
std::unordered_map<lzyb_handle_t, std::shared_ptr<SlotsInstance>> g_instances;
std::unordered_map<lzyb_handle_t, std::shared_ptr<ExecCtx>> g_ctxs;
lzyb_handle_t g_next = 0x6000000000000001ull;

std::shared_ptr<SlotsInstance> get_inst(lzyb_handle_t h) {
    std::shared_lock<std::shared_mutex> lk(g_mu);
    auto it = g_instances.find(h);
    return it == g_instances.end() ? nullptr : it->second;
}
std::shared_ptr<ExecCtx> get_ctx(lzyb_handle_t h) {
    std::shared_lock<std::shared_mutex> lk(g_mu);
    auto it = g_ctxs.find(h);
    return it == g_ctxs.end() ? nullptr : it->second;
}

int build_slots(ExecCtx& c) {
    // SlotsExecutionContext.beforeExecution, first half (SlotsExecutionContext.java:45-63)
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.closed) return fail(LZYB_E_STATE, "context is closed");
    if (c.built) return LZYB_OK;
    c.built = true;
    for (auto& d : c.descs) {
        if (d.direction == LZYB_DIR_INPUT) {
            auto s = std::make_shared<InputSlotImpl>();
            s->inst = c.inst;
            s->id = d.name;
            s->channel_id = d.channel;
            s->ctx = &c;
            c.slots.push_back(s);
            c.inputs[d.name] = s;
            s->start();
        } else {
            auto s = std::make_shared<OutputSlotImpl>();
            s->inst = c.inst;
            s->id = d.name;
            s->channel_id = d.channel;
            s->ctx = &c;
            s->sole_consumer = (d.flags & LZYB_SLOTF_SOLE_CONSUMER) != 0;
            c.slots.push_back(s);
            c.outputs[d.name] = s;
            s->start();
        }
    }
    return LZYB_OK;
}

int wait_all(ExecCtx& c, bool before, int timeout_ms) {
    auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms < 0 ? 0 : timeout_ms);
    size_t i = 0;
    while (true) {
        std::shared_ptr<SlotBase> s;
        {
            std::lock_guard<std::mutex> lk(c.mu);
            if (i >= c.slots.size()) break;  // slots added meanwhile ("-out" companions) are picked up too
            s = c.slots[i++];
        }
        Completion& f = before ? s->before() : s->after();
        int left = -1;
        if (timeout_ms >= 0) {
            auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(deadline - std::chrono::steady_clock::now()).count();
            left = ms < 0 ? 0 : (int)ms;
        }
        int st = f.wait(left);
        if (st == 0) return fail(LZYB_E_UNAVAILABLE, "timed out waiting for slot %s", s->id.c_str());
        if (st == 2) return fail(LZYB_E_LOCAL, "slot %s failed: %s", s->id.c_str(), f.error.c_str());
    }
    return LZYB_OK;
}

void close_ctx(ExecCtx& c) {
    std::vector<std::shared_ptr<SlotBase>> slots;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.closed) return;
        c.closed = true;
        slots = c.slots;
    }
    for (auto& s : slots) s->close();
    for (auto& s : slots) s->join();
    // companions created while closing
    {
        std::lock_guard<std::mutex> lk(c.mu);
        slots = c.slots;
    }
    for (auto& s : slots) {
        s->close();
        s->join();
    }
    for (auto& kv : c.inputs) kv.second->release_frame();
}

// ---- channel-manager action pump -----------------------------------------------------------------------
struct Pump {
    std::thread thr;
    std::atomic<bool> stop{false};
};
std::mutex g_pump_mu;
std::map<lzyb_handle_t, std::shared_ptr<Pump>> g_pumps;

void pump_loop(lzyb_handle_t cm, std::shared_ptr<Pump> p) {
    while (!p->stop.load()) {
        lzyb_cm_action_t a;
        int got = lzyb_cm_poll_action(cm, &a, 20);
        if (got <= 0) continue;
        if (a.kind == LZYB_ACT_START_TRANSFER) {
            // ActionScheduler.startTransfer: deliver StartTransfer(slot_id = consumer, peer = producer).
            // The consumer is looked up by slot id across the workers of this process.
            std::vector<std::shared_ptr<SlotsInstance>> insts;
            {
                std::shared_lock<std::shared_mutex> lk(g_mu);
                for (auto& kv : g_instances)
                    if (kv.second->cm == cm || lzyb_host::local_matcher_behind(kv.second->cm) == cm) insts.push_back(kv.second);
            }
            // The reference issues StartTransfer first and marks the transfer ACTIVE afterwards
            // (ActionScheduler.java:74-91); a consumer that finishes its download in between would
            // get FAILED_PRECONDITION from TransferCompleted.  With GPU-speed transfers that window
            // is real, so the transfer is activated first and the channel is failed if delivery fails.
            lzyb_cm_action_result(cm, a.channel_id, a.transfer_id, 1, "");
            int rc = LZYB_E_NOT_FOUND;
            for (auto& inst : insts) {
                auto s = inst->service.find(a.consumer_id);
                if (!s) continue;
                rc = s->start_transfer(a.producer, a.transfer_id);
                break;
            }
            if (rc != LZYB_OK) lzyb_cm_action_result(cm, a.channel_id, a.transfer_id, 0, "Cannot connect.");
        } else if (a.kind == LZYB_ACT_ABORT_WORKFLOW) {
            // lzy-service aborts the workflow: every context of that execution is closed
            std::vector<std::shared_ptr<ExecCtx>> ctxs;
            {
                std::shared_lock<std::shared_mutex> lk(g_mu);
                for (auto& kv : g_ctxs)
                    if ((kv.second->inst->cm == cm || lzyb_host::local_matcher_behind(kv.second->inst->cm) == cm) &&
                        kv.second->execution_id == a.execution_id)
                        ctxs.push_back(kv.second);
            }
            for (auto& c : ctxs) {
                std::vector<std::shared_ptr<SlotBase>> slots;
                {
                    std::lock_guard<std::mutex> lk(c->mu);
                    slots = c->slots;
                }
                for (auto& s : slots) s->close();
            }
        }
    }
}

}  // namespace

extern "C" {

int lzyb_slots_set_backend(const lzyb_backend_ops_t* o) {
    if (!o) {
        g_ops = &kHbmBackend;
        return LZYB_OK;
    }
    g_custom_ops = *o;
    g_ops = &g_custom_ops;
    return LZYB_OK;
}

uint64_t lzyb_slots_threads_spawned(void) { return slot_threads().spawned(); }

int lzyb_slots_set_peer_resolver(lzyb_peer_resolver_t resolver) {
    g_peer_resolver = resolver;
    return LZYB_OK;
}

int lzyb_slots_describe(lzyb_handle_t h, const char* peer_id, lzyb_ipc_desc_t* out) {
    auto inst = get_inst(h);
    if (!inst) return fail(LZYB_E_ARG, "unknown slots handle");
    if (!peer_id || !out) return fail(LZYB_E_ARG, "null argument");
    auto s = inst->service.find(peer_id);
    if (!s) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", peer_id);
    if (s->is_input()) return fail(LZYB_E_UNIMPLEMENTED, "Slot %s is an input slot", peer_id);
    if (!ops().slot_export) return fail(LZYB_E_UNIMPLEMENTED, "this backend cannot export frames");
    lzyb_handle_t frame = 0;
    int rc = ops().slot_find(peer_id, &frame);
    if (rc != LZYB_OK) return fail(LZYB_E_NOT_FOUND, "Slot %s has no frame: %s", peer_id, lzyb_last_error());
    return ops().slot_export(frame, out);
}

int lzyb_slots_set_cm_client(const lzyb_cm_client_ops_t* o) {
    if (!o) {
        g_cmc = &kLocalMatcher;
        return LZYB_OK;
    }
    if (!o->bind || !o->unbind || !o->transfer_completed || !o->transfer_failed) return fail(LZYB_E_ARG, "incomplete client table");
    g_custom_cmc = *o;
    g_cmc = &g_custom_cmc;
    return LZYB_OK;
}

int lzyb_slots_create(lzyb_handle_t cm, int gpu, const char* api_url, lzyb_handle_t* out) {
    if (!out || !api_url) return fail(LZYB_E_ARG, "null argument");
    auto s = std::make_shared<SlotsInstance>();
    s->cm = cm;
    s->gpu = gpu;
    s->api_url = api_url;
    std::unique_lock<std::shared_mutex> lk(g_mu);
    s->handle = g_next++;
    g_instances[s->handle] = s;
    *out = s->handle;
    return LZYB_OK;
}

int lzyb_slots_free(lzyb_handle_t h) {
    std::shared_ptr<SlotsInstance> s;
    std::vector<std::shared_ptr<ExecCtx>> ctxs;
    {
        std::unique_lock<std::shared_mutex> lk(g_mu);
        auto it = g_instances.find(h);
        if (it == g_instances.end()) return LZYB_OK;
        s = it->second;
        g_instances.erase(it);
        for (auto it2 = g_ctxs.begin(); it2 != g_ctxs.end();) {
            if (it2->second->inst == s.get()) {
                ctxs.push_back(it2->second);
                it2 = g_ctxs.erase(it2);
            } else
                ++it2;
        }
    }
    for (auto& c : ctxs) close_ctx(*c);  // Slots.close(): every context is closed
    return LZYB_OK;
}

int lzyb_slots_set_chunk_bytes(lzyb_handle_t h, uint64_t chunk) {
    auto s = get_inst(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slots handle");
    if (chunk & 15) return fail(LZYB_E_ARG, "chunk size must be a multiple of 16");
    s->chunk_bytes = chunk;
    return LZYB_OK;
}

int lzyb_slots_set_peer_chunk_bytes(lzyb_handle_t h, uint64_t chunk) {
    auto s = get_inst(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slots handle");
    if (chunk & 15) return fail(LZYB_E_ARG, "chunk size must be a multiple of 16");
    s->peer_chunk_bytes = chunk;
    return LZYB_OK;
}

int lzyb_ctx_create(lzyb_handle_t slots, const char* execution_id, const char* task_id, lzyb_handle_t* out) {
    auto inst = get_inst(slots);
    if (!inst) return fail(LZYB_E_ARG, "unknown slots handle");
    if (!out) return fail(LZYB_E_ARG, "null argument");
    auto c = std::make_shared<ExecCtx>();
    c->inst = inst.get();
    c->execution_id = execution_id ? execution_id : "";
    c->task_id = task_id ? task_id : "";
    std::unique_lock<std::shared_mutex> lk(g_mu);
    c->handle = g_next++;
    g_ctxs[c->handle] = c;
    *out = c->handle;
    return LZYB_OK;
}

int lzyb_ctx_add_slot(lzyb_handle_t h, const char* slot_name, const char* channel_id, int direction) {
    return lzyb_ctx_add_slot_flags(h, slot_name, channel_id, direction, 0);
}

int lzyb_ctx_add_slot_flags(lzyb_handle_t h, const char* slot_name, const char* channel_id, int direction, uint32_t flags) {
    auto c = get_ctx(h);
    if (!c) return fail(LZYB_E_ARG, "unknown context handle");
    if (!slot_name || !channel_id) return fail(LZYB_E_ARG, "null argument");
    if (direction != LZYB_DIR_INPUT && direction != LZYB_DIR_OUTPUT) return fail(LZYB_E_INVALID, "slot direction must be INPUT or OUTPUT");
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->built) return fail(LZYB_E_STATE, "slots are already built");
    c->descs.push_back({slot_name, channel_id, direction, flags});
    return LZYB_OK;
}

int lzyb_ctx_prefetch(lzyb_handle_t h) {
    auto c = get_ctx(h);
    if (!c) return fail(LZYB_E_ARG, "unknown context handle");
    return build_slots(*c);
}

int lzyb_ctx_before_execution(lzyb_handle_t h, int timeout_ms) {
    auto c = get_ctx(h);
    if (!c) return fail(LZYB_E_ARG, "unknown context handle");
    int rc = build_slots(*c);
    if (rc != LZYB_OK) return rc;
    rc = wait_all(*c, true, timeout_ms);
    if (rc != LZYB_OK && rc != LZYB_E_UNAVAILABLE) {
        std::string keep = lzyb_last_error();
        close_ctx(*c);  // "Failed to initialize slots" → close()
        return fail(rc, "%s", keep.c_str());
    }
    return rc;
}

int lzyb_ctx_input_handle(lzyb_handle_t h, const char* slot_name, lzyb_handle_t* out) {
    auto c = get_ctx(h);
    if (!c || !slot_name || !out) return fail(LZYB_E_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->inputs.find(slot_name);
    if (it == c->inputs.end()) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", slot_name);
    if (it->second->state.load() != LZYB_IS_READY && !it->second->ready_.done()) return fail(LZYB_E_STATE, "slot %s is not ready", slot_name);
    if (!it->second->backend) return fail(LZYB_E_STATE, "slot %s has no data", slot_name);
    *out = it->second->backend;
    return LZYB_OK;
}

int lzyb_ctx_input_state(lzyb_handle_t h, const char* slot_name) {
    auto c = get_ctx(h);
    if (!c || !slot_name) return fail(LZYB_E_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->inputs.find(slot_name);
    if (it == c->inputs.end()) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", slot_name);
    if (it->second->ready_.done() && it->second->ready_.state == 1) return LZYB_IS_READY;
    return it->second->state.load();
}

int lzyb_ctx_output_open(lzyb_handle_t h, const char* slot_name, uint64_t nbytes, lzyb_handle_t* out) {
    auto c = get_ctx(h);
    if (!c || !slot_name || !out) return fail(LZYB_E_ARG, "bad argument");
    int rc = build_slots(*c);
    if (rc != LZYB_OK) return rc;
    std::shared_ptr<OutputSlotImpl> s;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        auto it = c->outputs.find(slot_name);
        if (it == c->outputs.end()) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", slot_name);
        s = it->second;
    }
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->closed) return fail(LZYB_E_STATE, "slot %s is closed", slot_name);
    if (s->backend) return fail(LZYB_E_STATE, "output %s is already open", slot_name);
    rc = ops().slot_create(slot_name, c->inst->gpu, nbytes, &s->backend);
    if (rc != LZYB_OK) return rc;
    *out = s->backend;
    return LZYB_OK;
}

int lzyb_ctx_output_done(lzyb_handle_t h, const char* slot_name) {
    auto c = get_ctx(h);
    if (!c || !slot_name) return fail(LZYB_E_ARG, "bad argument");
    std::shared_ptr<OutputSlotImpl> s;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        auto it = c->outputs.find(slot_name);
        if (it == c->outputs.end()) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", slot_name);
        s = it->second;
    }
    if (!s->backend) return fail(LZYB_E_STATE, "output %s was never opened", slot_name);
    s->signal_data_done();
    return LZYB_OK;
}

int lzyb_ctx_after_execution(lzyb_handle_t h, int timeout_ms) {
    auto c = get_ctx(h);
    if (!c) return fail(LZYB_E_ARG, "unknown context handle");
    int rc = wait_all(*c, false, timeout_ms);
    if (rc != LZYB_OK && rc != LZYB_E_UNAVAILABLE) {
        std::string keep = lzyb_last_error();
        close_ctx(*c);  // "Failed to finalize slots" → close()
        return fail(rc, "%s", keep.c_str());
    }
    return rc;
}

int lzyb_ctx_close(lzyb_handle_t h) {
    std::shared_ptr<ExecCtx> c;
    {
        std::unique_lock<std::shared_mutex> lk(g_mu);
        auto it = g_ctxs.find(h);
        if (it == g_ctxs.end()) return LZYB_OK;
        c = it->second;
        g_ctxs.erase(it);
    }
    close_ctx(*c);
    return LZYB_OK;
}

int lzyb_slots_start_transfer(lzyb_handle_t h, const char* slot_id, const lzyb_peer_desc_t* peer, const char* transfer_id) {
    auto inst = get_inst(h);
    if (!inst) return fail(LZYB_E_ARG, "unknown slots handle");
    if (!slot_id || !peer) return fail(LZYB_E_ARG, "null argument");
    auto s = inst->service.find(slot_id);
    if (!s) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", slot_id);
    return s->start_transfer(*peer, transfer_id ? transfer_id : "");
}

int64_t lzyb_slots_read(lzyb_handle_t h, const char* peer_id, uint64_t offset, void* dst, uint64_t len) {
    auto inst = get_inst(h);
    if (!inst) return fail(LZYB_E_ARG, "unknown slots handle");
    if (!peer_id) return fail(LZYB_E_ARG, "null argument");
    auto s = inst->service.find(peer_id);
    if (!s) return fail(LZYB_E_NOT_FOUND, "Slot not found: %s", peer_id);
    return s->read(offset, dst, len);
}

int lzyb_cm_pump_start(lzyb_handle_t cm) {
    std::lock_guard<std::mutex> lk(g_pump_mu);
    if (g_pumps.count(cm)) return LZYB_OK;
    auto p = std::make_shared<Pump>();
    p->thr = std::thread(pump_loop, cm, p);
    g_pumps[cm] = p;
    return LZYB_OK;
}

int lzyb_cm_pump_stop(lzyb_handle_t cm) {
    std::shared_ptr<Pump> p;
    {
        std::lock_guard<std::mutex> lk(g_pump_mu);
        auto it = g_pumps.find(cm);
        if (it == g_pumps.end()) return LZYB_OK;
        p = it->second;
        g_pumps.erase(it);
    }
    p->stop = true;
    if (p->thr.joinable()) p->thr.join();
    return LZYB_OK;
}

int lzyb_blob_put(const char* uri, const void* data, uint64_t len) {
    if (!uri || (len && !data)) return fail(LZYB_E_ARG, "null argument");
    auto b = std::make_shared<Blob>();
    b->data.assign(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + len);
    std::lock_guard<std::mutex> lk(g_blob_mu);
    g_blobs[uri] = b;
    return LZYB_OK;
}

int64_t lzyb_blob_size(const char* uri) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    auto b = blob_find(uri);
    if (!b) return fail(LZYB_E_NOT_FOUND, "no object at %s", uri);
    return (int64_t)b->data.size();
}

int64_t lzyb_blob_get(const char* uri, void* dst, uint64_t cap) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    auto b = blob_find(uri);
    if (!b) return fail(LZYB_E_NOT_FOUND, "no object at %s", uri);
    if (b->data.size() > cap) return fail(LZYB_E_ARG, "object holds %zu bytes, buffer %llu", b->data.size(), (unsigned long long)cap);
    if (!b->data.empty()) memcpy(dst, b->data.data(), b->data.size());
    return (int64_t)b->data.size();
}

int lzyb_frame_store_put(const char* uri, lzyb_handle_t slot) {
    if (!uri || !slot) return fail(LZYB_E_ARG, "null argument");
    int rc = ops().slot_retain(slot);
    if (rc != LZYB_OK) return rc;
    lzyb_handle_t old = 0;
    {
        std::lock_guard<std::mutex> lk(g_frames_mu);
        auto it = g_frames.find(uri);
        if (it != g_frames.end()) old = it->second;
        g_frames[uri] = slot;
    }
    if (old) ops().slot_close(old);
    return LZYB_OK;
}

int lzyb_frame_store_get(const char* uri, lzyb_handle_t* out) {
    if (!uri || !out) return fail(LZYB_E_ARG, "null argument");
    int rc = frame_lookup(uri, -1, out);
    if (rc == LZYB_E_NOT_FOUND) return fail(LZYB_E_NOT_FOUND, "no frame at %s", uri);
    return rc;  // LZYB_OK, or why an existing spill file could not be used
}

int lzyb_frame_store_set_spill(const char* dir, int default_gpu) {
    std::lock_guard<std::mutex> lk(g_spill_mu);
    if (!dir || !*dir) {
        g_spill_dir.clear();
        return LZYB_OK;
    }
    if (::mkdir(dir, 0755) != 0 && errno != EEXIST) return fail(LZYB_E_LOCAL, "cannot create %s: %s", dir, strerror(errno));
    g_spill_dir = dir;
    g_spill_gpu = default_gpu;
    return LZYB_OK;
}

int lzyb_frame_store_spill(const char* uri) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    const std::string dir = spill_dir();
    if (dir.empty()) return fail(LZYB_E_STATE, "no spill directory configured");
    lzyb_handle_t h = frame_resident(uri);
    if (!h) return fail(LZYB_E_NOT_FOUND, "no resident frame at %s", uri);
    return spill_write(spill_path(dir, uri), h);
}

int lzyb_frame_store_load(const char* uri, int gpu, lzyb_handle_t* out) {
    if (!uri || !out) return fail(LZYB_E_ARG, "null argument");
    int rc = frame_lookup(uri, gpu, out);
    if (rc == LZYB_E_NOT_FOUND) return fail(LZYB_E_NOT_FOUND, "no frame at %s", uri);
    return rc;
}

int lzyb_frame_store_resident(const char* uri) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    return frame_resident(uri) ? 1 : 0;
}

int lzyb_frame_store_spilled(const char* uri) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    const std::string dir = spill_dir();
    if (dir.empty()) return 0;
    return ::access(spill_path(dir, uri).c_str(), R_OK) == 0 ? 1 : 0;
}

int lzyb_frame_store_unspill(const char* uri) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    const std::string dir = spill_dir();
    if (dir.empty()) return LZYB_OK;
    ::unlink(spill_path(dir, uri).c_str());
    return LZYB_OK;
}

int lzyb_frame_store_drop(const char* uri) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    lzyb_handle_t h = 0;
    {
        std::lock_guard<std::mutex> lk(g_frames_mu);
        auto it = g_frames.find(uri);
        if (it == g_frames.end()) return LZYB_OK;
        h = it->second;
        g_frames.erase(it);
    }
    ops().slot_close(h);
    return LZYB_OK;
}

int lzyb_frame_store_clear(const char* prefix) {
    std::vector<lzyb_handle_t> victims;
    {
        std::lock_guard<std::mutex> lk(g_frames_mu);
        const std::string p = prefix ? prefix : "";
        for (auto it = g_frames.begin(); it != g_frames.end();) {
            if (it->first.compare(0, p.size(), p) == 0) {
                victims.push_back(it->second);
                it = g_frames.erase(it);
            } else
                ++it;
        }
    }
    for (auto h : victims) ops().slot_close(h);
    return (int)victims.size();
}

int64_t lzyb_frame_store_bytes(void) {
    std::vector<lzyb_handle_t> hs;
    {
        std::lock_guard<std::mutex> lk(g_frames_mu);
        for (auto& kv : g_frames) hs.push_back(kv.second);
    }
    std::sort(hs.begin(), hs.end());
    hs.erase(std::unique(hs.begin(), hs.end()), hs.end());
    int64_t total = 0;
    for (auto h : hs) {
        lzyb_slot_info_t info;
        if (slot_sizes(h, &info) == LZYB_OK) total += (int64_t)info.nbytes;
    }
    return total;
}

int64_t lzyb_slots_transfer_log(lzyb_handle_t h, char* json_out, uint64_t cap, int clear) {
    auto s = get_inst(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slots handle");
    std::ostringstream os;
    auto quote = [&os](const std::string& v) {
        os << '"';
        for (char c : v) {
            if (c == '"' || c == '\\')
                os << '\\' << c;
            else if ((unsigned char)c < 0x20)
                os << ' ';
            else
                os << c;
        }
        os << '"';
    };
    {
        std::lock_guard<std::mutex> lk(s->log_mu);
        os << '[';
        bool first = true;
        for (const auto& r : s->transfer_log) {
            if (!first) os << ',';
            first = false;
            os << "{\"slot_id\":";
            quote(r.slot_id);
            os << ",\"channel_id\":";
            quote(r.channel_id);
            os << ",\"peer_id\":";
            quote(r.peer_id);
            os << ",\"peer_kind\":\"" << (r.peer_kind == LZYB_PEER_STORAGE ? "storage" : "slot") << "\",\"bytes\":" << r.bytes
               << ",\"ns\":" << r.ns << ",\"ns_setup\":" << r.ns_setup << ",\"ns_issue\":" << r.ns_issue << ",\"ns_wait\":" << r.ns_wait
               << ",\"ns_complete\":" << r.ns_complete << ",\"failovers\":" << r.failovers << ",\"ok\":" << (r.ok ? "true" : "false") << ",\"error\":";
            quote(r.error);
            os << '}';
        }
        os << ']';
        if (clear) s->transfer_log.clear();
    }
    const std::string out = os.str();
    if (json_out && cap) {
        const uint64_t n = std::min<uint64_t>(cap - 1, out.size());
        memcpy(json_out, out.data(), n);
        json_out[n] = 0;
    }
    return (int64_t)out.size() + 1;
}

int lzyb_slots_set_republish(lzyb_handle_t h, int on) {
    auto s = get_inst(h);
    if (!s) return fail(LZYB_E_ARG, "unknown slots handle");
    s->republish = on != 0;
    return LZYB_OK;
}

int lzyb_blob_delete(const char* uri) {
    if (!uri) return fail(LZYB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(g_blob_mu);
    g_blobs.erase(uri);
    return LZYB_OK;
}

}  // extern "C"
