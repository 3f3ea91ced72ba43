// host_stress.cpp — a native, multi-threaded workout of the C++ host logic (channel manager, slot state
// machines, storage / spill tiers) over the in-memory backend, built with -fsanitize=thread and with
// -fsanitize=address,undefined by tests/test_host_sanitizers.py.  The reference has no race detector at all
// (SURVEY §5); its slot code relies on `synchronized` and DB transactions — here every slot is a thread, the
// action pump is a thread, and the tests above them are Python, so this is where the C++ gets its sanitizer time.
//
// Workload per worker thread (deterministic from its seed): producer → consumer pairs with early and late
// consumers, 1 → 3 fan-outs, producers that die half-way after their storage backup exists (failover with offset
// resume), spill + drop + reload of stored frames, contexts closed while their inputs are still waiting.
// Exit code 0 = every delivered byte matched; sanitizer findings make the run fail on their own.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lzy_b200_host.h"

extern "C" {
extern const lzyb_backend_ops_t hostmem_backend_ops;
int hostmem_write(lzyb_handle_t h, uint64_t off, const void* src, uint64_t len);
int64_t hostmem_peek(lzyb_handle_t h, uint64_t off, void* dst, uint64_t len);
int hostmem_set_fail_after(lzyb_handle_t h, long long n);
int hostmem_live_slots(void);
}

namespace {

std::atomic<int> g_failures{0};

#define REQUIRE(cond, ...)                                  \
    do {                                                    \
        if (!(cond)) {                                      \
            fprintf(stderr, "FAILED %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);                   \
            fprintf(stderr, " [%s]\n", lzyb_last_error()); \
            g_failures++;                                   \
            return;                                         \
        }                                                   \
    } while (0)

std::vector<uint8_t> payload(std::mt19937_64& rng, size_t n) {
    std::vector<uint8_t> v(n);
    for (auto& b : v) b = static_cast<uint8_t>(rng());
    return v;
}

struct Ctx {
    lzyb_handle_t h = 0;
    Ctx(lzyb_handle_t slots, const std::string& exec, const std::string& task) {
        if (lzyb_ctx_create(slots, exec.c_str(), task.c_str(), &h) != LZYB_OK) h = 0;
    }
    ~Ctx() {
        if (h) lzyb_ctx_close(h);
    }
};

bool delivered_equals(lzyb_handle_t ctx, const std::string& slot, const std::vector<uint8_t>& want) {
    lzyb_handle_t h = 0;
    if (lzyb_ctx_input_handle(ctx, slot.c_str(), &h) != LZYB_OK) return false;
    std::vector<uint8_t> got(want.size() + 1);
    int64_t n = hostmem_peek(h, 0, got.data(), got.size());
    return n == (int64_t)want.size() && memcmp(got.data(), want.data(), want.size()) == 0;
}

void worker(int id, lzyb_handle_t cm, const std::vector<lzyb_handle_t>& slots, int jobs) {
    std::mt19937_64 rng(1234 + id);
    const std::string exec = "exec-" + std::to_string(id);
    static const size_t kSizes[] = {5, 4096, 70001, 1 << 20};
    for (int j = 0; j < jobs; ++j) {
        const std::string tag = std::to_string(id) + "-" + std::to_string(j);
        const auto data = payload(rng, kSizes[rng() % 4]);
        const int kind = (int)(rng() % 5);
        const bool flaky = kind == 2 && data.size() > 64;
        // slot names are unique per task, as the reference's "/{call}/{entry}" paths are; a flaky producer needs a
        // backup that is a copy (host blob), not a reference to its own frame (hbm://)
        const std::string in = "/in/" + tag, out = "/out/" + tag;
        const std::string uri = (flaky ? "mem://stress/" : "hbm://stress/") + tag;
        char chan[64];
        REQUIRE(lzyb_cm_get_or_create(cm, "u", "wf", exec.c_str(), "", "", "", LZYB_ROLE_CONSUMER, uri.c_str(), chan) == LZYB_OK,
                "get_or_create %s", tag.c_str());
        const lzyb_handle_t wp = slots[rng() % slots.size()], wc = slots[rng() % slots.size()];
        const bool early = kind == 0 || kind == 3;
        const int consumers = kind == 1 ? 3 : 1;

        std::vector<std::unique_ptr<Ctx>> cons;
        for (int c = 0; c < consumers; ++c) {
            cons.emplace_back(new Ctx(c == 0 ? wc : slots[rng() % slots.size()], exec, "c" + tag + "." + std::to_string(c)));
            REQUIRE(cons.back()->h, "consumer ctx");
            REQUIRE(lzyb_ctx_add_slot(cons.back()->h, (in + "." + std::to_string(c)).c_str(), chan, LZYB_DIR_INPUT) == LZYB_OK, "add in");
            if (early) REQUIRE(lzyb_ctx_prefetch(cons.back()->h) == LZYB_OK, "prefetch");
        }
        if (kind == 3) {  // a context that is closed while its input still waits for a peer
            Ctx doomed(wc, exec, "doomed" + tag);
            REQUIRE(doomed.h && lzyb_ctx_add_slot(doomed.h, (in + ".doomed").c_str(), chan, LZYB_DIR_INPUT) == LZYB_OK, "doomed");
            REQUIRE(lzyb_ctx_prefetch(doomed.h) == LZYB_OK, "doomed prefetch");
        }

        Ctx prod(wp, exec, "p" + tag);
        REQUIRE(prod.h && lzyb_ctx_add_slot(prod.h, out.c_str(), chan, LZYB_DIR_OUTPUT) == LZYB_OK, "producer ctx");
        REQUIRE(lzyb_ctx_before_execution(prod.h, 30000) == LZYB_OK, "producer before");
        lzyb_handle_t oh = 0;
        REQUIRE(lzyb_ctx_output_open(prod.h, out.c_str(), data.size(), &oh) == LZYB_OK, "output_open");
        REQUIRE(hostmem_write(oh, 0, data.data(), data.size()) == 0, "write");
        REQUIRE(lzyb_ctx_output_done(prod.h, out.c_str()) == LZYB_OK, "output_done");
        if (flaky || !early) REQUIRE(lzyb_ctx_after_execution(prod.h, 30000) == LZYB_OK, "producer after");
        if (flaky) hostmem_set_fail_after(oh, (long long)((data.size() / 2) & ~15ull));  // backup exists: fail over to it

        for (size_t c = 0; c < cons.size(); ++c) {
            REQUIRE(lzyb_ctx_before_execution(cons[c]->h, 30000) == LZYB_OK, "consumer before (%s, kind %d)", tag.c_str(), kind);
            REQUIRE(delivered_equals(cons[c]->h, in + "." + std::to_string(c), data), "bytes differ (%s, kind %d)", tag.c_str(), kind);
        }
        if (!flaky && early) REQUIRE(lzyb_ctx_after_execution(prod.h, 30000) == LZYB_OK, "producer after (late)");

        if (kind == 4) {  // the stored frame leaves "HBM" and a later consumer brings it back from the tier
            REQUIRE(lzyb_frame_store_spill(uri.c_str()) == LZYB_OK, "spill");
            cons.clear();
            lzyb_ctx_close(prod.h);
            prod.h = 0;
            REQUIRE(lzyb_frame_store_drop(uri.c_str()) == LZYB_OK, "drop");
            Ctx late(wc, exec, "late" + tag);
            REQUIRE(late.h && lzyb_ctx_add_slot(late.h, (in + ".late").c_str(), chan, LZYB_DIR_INPUT) == LZYB_OK, "late ctx");
            REQUIRE(lzyb_ctx_before_execution(late.h, 30000) == LZYB_OK, "late before");
            REQUIRE(delivered_equals(late.h, in + ".late", data), "reloaded bytes differ");
            lzyb_frame_store_unspill(uri.c_str());
        }
        cons.clear();
        if (prod.h) {
            lzyb_ctx_close(prod.h);
            prod.h = 0;
        }
        if (flaky)
            lzyb_blob_delete(uri.c_str());
        else
            lzyb_frame_store_drop(uri.c_str());
    }
}

}  // namespace

int main(int argc, char** argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 4, jobs = argc > 2 ? atoi(argv[2]) : 20;
    const char* spill_dir = argc > 3 ? argv[3] : "/tmp/lzyb_host_stress_spill";
    lzyb_slots_set_backend(&hostmem_backend_ops);
    if (lzyb_frame_store_set_spill(spill_dir, 0) != LZYB_OK) return 2;
    // argv[4]: a unix socket path — the slot state machines then reach the matcher through the box-local transport
    // (csrc/host/local_rpc.cpp: serving threads, per-thread client connections) instead of calling it in process
    const char* socket_path = argc > 4 ? argv[4] : nullptr;
    lzyb_handle_t cm = 0;
    if (lzyb_cm_create(&cm) != LZYB_OK || lzyb_cm_seed(cm, 7) != LZYB_OK || lzyb_cm_pump_start(cm) != LZYB_OK) return 2;
    lzyb_handle_t via = cm;  // what the slots instances are created with: the matcher, or the connection to its socket
    if (socket_path && (lzyb_cm_serve_local(cm, socket_path) != LZYB_OK || lzyb_slots_connect_cm_local(socket_path, &via) != LZYB_OK)) {
        fprintf(stderr, "local socket: %s\n", lzyb_last_error());
        return 2;
    }
    std::vector<lzyb_handle_t> slots(4);
    for (size_t i = 0; i < slots.size(); ++i)
        if (lzyb_slots_create(via, 0, ("lzyb://worker-" + std::to_string(i)).c_str(), &slots[i]) != LZYB_OK) return 2;

    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t) ts.emplace_back(worker, t, cm, std::cref(slots), jobs);
    for (auto& t : ts) t.join();

    for (auto s : slots) lzyb_slots_free(s);
    if (socket_path) {
        lzyb_slots_disconnect_cm_local();
        lzyb_cm_stop_local(cm);
    }
    lzyb_cm_pump_stop(cm);
    lzyb_cm_free(cm);
    lzyb_frame_store_clear("");
    lzyb_frame_store_set_spill(nullptr, 0);
    const int live = hostmem_live_slots();
    printf("{\"threads\": %d, \"jobs_per_thread\": %d, \"failures\": %d, \"live_slots\": %d}\n", threads, jobs, g_failures.load(), live);
    return g_failures.load() == 0 && live == 0 ? 0 : 1;
}
