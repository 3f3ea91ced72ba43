// host_chain_bench.cpp — what one op costs in the C++ host logic alone (contexts, slot state machines, channel
// manager, action pump), measured without a GPU and without Python: an 8-op chain per step, as bench.py's workload,
// over the in-memory backend with payloads small enough that no byte-moving shows up.  Results stay in HBM
// ("hbm://" storage consumer: a reference, no copy), as in the shipped runtime.
//
//   g++ -std=c++17 -O2 -Iinclude tools/host_chain_bench.cpp tests/native/dataplane_stubs.cpp \
//       lzy_b200/csrc/host/*.cpp tests/support/hostmem_backend.c(obj) -lpthread
//   (tools/host_chain_bench.sh does it)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/lzy_b200_host.h"

extern "C" {
extern const lzyb_backend_ops_t hostmem_backend_ops;
int hostmem_write(lzyb_handle_t h, uint64_t off, const void* src, uint64_t len);
int hostmem_live_slots(void);
}

#define CHECK(expr)                                                                   \
    do {                                                                              \
        if ((expr) != LZYB_OK) {                                                      \
            fprintf(stderr, "%s failed: %s (line %d)\n", #expr, lzyb_last_error(), __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

static double g_phase[8];
static const char* kPhase[8] = {"get_or_create", "ctx+add_slot", "prefetch", "before", "open+write+done", "after", "close", "destroy+clear"};
struct Phase {
    int i;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    explicit Phase(int i) : i(i) {}
    ~Phase() { g_phase[i] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t).count(); }
};

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 2000, n_ops = argc > 2 ? atoi(argv[2]) : 8;
    const bool prefetch = argc > 3 ? atoi(argv[3]) != 0 : true;
    const bool republish = argc > 4 ? atoi(argv[4]) != 0 : false;  // "<slot>-out" companions: off in B200Runtime's default
    lzyb_slots_set_backend(&hostmem_backend_ops);
    lzyb_handle_t cm = 0, slots = 0;
    CHECK(lzyb_cm_create(&cm));
    CHECK(lzyb_cm_seed(cm, 7));
    CHECK(lzyb_cm_pump_start(cm));
    // argv[5]: a unix socket path — Bind / TransferCompleted / Unbind of every slot then cross the box-local transport
    // (csrc/host/local_rpc.cpp) instead of being function calls: the transport's cost per op, without Python
    const char* socket_path = argc > 5 ? argv[5] : nullptr;
    lzyb_handle_t via = cm;
    if (socket_path) {
        CHECK(lzyb_cm_serve_local(cm, socket_path));
        CHECK(lzyb_slots_connect_cm_local(socket_path, &via));
    }
    CHECK(lzyb_slots_create(via, 0, "lzyb://worker-0", &slots));
    CHECK(lzyb_slots_set_republish(slots, republish ? 1 : 0));
    const char payload[256] = "x";
    using clk = std::chrono::steady_clock;
    clk::time_point t0;
    const int warm = steps / 10 + 1;
    for (int k = 0; k < steps + warm; ++k) {
        if (k == warm) t0 = clk::now();
        if (getenv("LZYB_BENCH_TRACE") && k % 100 == 0) {
            static clk::time_point last = clk::now();
            fprintf(stderr, "step %d: %.1f us/op over the last 100 steps |", k, std::chrono::duration<double, std::micro>(clk::now() - last).count() / 100 / n_ops);
            for (int p = 0; p < 8; ++p) {
                fprintf(stderr, " %s %.0f", kPhase[p], g_phase[p] / 100 / n_ops);
                g_phase[p] = 0;
            }
            if (FILE* f = fopen("/proc/self/status", "r")) {
                char line[256];
                while (fgets(line, sizeof line, f))
                    if (!strncmp(line, "Threads:", 8) || !strncmp(line, "VmRSS:", 6)) {
                        line[strcspn(line, "\n")] = 0;
                        fprintf(stderr, " | %s", line);
                    }
                fclose(f);
            }
            fprintf(stderr, "\n");
            last = clk::now();
        }
        std::vector<std::string> chans(n_ops);
        std::vector<lzyb_handle_t> ctxs(n_ops);
        for (int i = 0; i < n_ops; ++i) {
            Phase ph(0);
            char chan[64];
            const std::string uri = "hbm://bench/" + std::to_string(k) + "/" + std::to_string(i);
            CHECK(lzyb_cm_get_or_create(cm, "u", "wf", "exec", "", "", "", LZYB_ROLE_CONSUMER, uri.c_str(), chan));
            chans[i] = chan;
        }
        auto name = [&](int i, const char* what) { return "/t" + std::to_string(k) + "-" + std::to_string(i) + "/" + what; };
        for (int i = 0; i < n_ops; ++i) {
            Phase ph(1);
            CHECK(lzyb_ctx_create(slots, "exec", ("t" + std::to_string(k) + "-" + std::to_string(i)).c_str(), &ctxs[i]));
            if (i > 0) CHECK(lzyb_ctx_add_slot(ctxs[i], name(i, "in").c_str(), chans[i - 1].c_str(), LZYB_DIR_INPUT));
            CHECK(lzyb_ctx_add_slot(ctxs[i], name(i, "out").c_str(), chans[i].c_str(), LZYB_DIR_OUTPUT));
        }
        if (prefetch) {
            Phase ph(2);
            for (int i = 0; i < n_ops; ++i) CHECK(lzyb_ctx_prefetch(ctxs[i]));
        }
        for (int i = 0; i < n_ops; ++i) {
            {
                Phase ph(3);
                CHECK(lzyb_ctx_before_execution(ctxs[i], 30000));
            }
            {
                Phase ph(4);
                lzyb_handle_t h = 0;
                CHECK(lzyb_ctx_output_open(ctxs[i], name(i, "out").c_str(), sizeof(payload), &h));
                hostmem_write(h, 0, payload, sizeof(payload));
                CHECK(lzyb_ctx_output_done(ctxs[i], name(i, "out").c_str()));
            }
            Phase ph(5);
            CHECK(lzyb_ctx_after_execution(ctxs[i], 30000));
        }
        {
            Phase ph(6);
            for (int i = 0; i < n_ops; ++i) CHECK(lzyb_ctx_close(ctxs[i]));
        }
        Phase ph(7);
        CHECK(lzyb_cm_destroy_all(cm, "exec", "done"));
        lzyb_frame_store_clear("hbm://bench/");
    }
    const double us = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
    lzyb_slots_free(slots);
    if (socket_path) {
        lzyb_slots_disconnect_cm_local();
        lzyb_cm_stop_local(cm);
    }
    lzyb_cm_pump_stop(cm);
    lzyb_cm_free(cm);
    printf("{\"matcher_calls\": \"%s\", \"steps\": %d, \"ops_per_step\": %d, \"prefetch\": %s, \"republish\": %s, \"us_per_op\": %.1f, \"live_slots\": %d}\n",
           socket_path ? "unix socket (local_rpc.cpp)" : "in process", steps,
           n_ops, prefetch ? "true" : "false", republish ? "true" : "false", us / steps / n_ops, hostmem_live_slots());
    return 0;
}
