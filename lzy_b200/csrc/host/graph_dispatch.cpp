// graph_dispatch.cpp — which @op tasks may run now.
//
// Replaces the graph-executor's BFS dispatch loop:
//   lzy/graph-executor/src/main/java/ai/lzy/graph/exec/BfsGraphProcessor.java
//       exec :44-87 (status machine), stop :94-122, nextStep :124-167 (Thread.sleep(1000) per step, :148),
//       getNextExecutionGroup :176-258
//   lzy/graph-executor/src/main/java/ai/lzy/graph/exec/impl/DirectChannelChecker.java:24-32
//       an edge is ready when its producer task reported Success
//   lzy/graph-executor/src/main/java/ai/lzy/graph/algo/Algorithms.java
//       condenseGraph :56-117 (strongly connected components run as one group), findRoots :26-54,
//       getNextBfsGroup :7-24
// and uses the dependency-count idea of the unfinished rewrite
//   lzy/graph-executor-2/src/main/java/ai/lzy/graph/algo/Algorithms.java:10-73 (buildTaskDependents + cycle check).
//
// Same observable rule: a task becomes runnable when every producer task of each of
// its input channels has finished successfully; independent runnable tasks start together; members
// of one strongly connected component start together; the first failure stops the graph and the
// still-running tasks are reported for killing.  What is gone is the polling: readiness is
// maintained incrementally by per-component pending-parent counters, so a step costs O(out-degree)
// instead of a scheduler RPC per task plus a one-second sleep.
#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/lzy_b200_host.h"
#include "host_common.h"

namespace {

using lzyb_host::fail;

struct Graph {
    std::mutex mu;
    int n = 0;
    std::vector<std::vector<int>> children, parents;  // task level, deduplicated
    std::vector<int> comp_of;                         // task -> component
    std::vector<std::vector<int>> comp_tasks;
    std::vector<std::set<int>> task_child_comps;      // task -> other components it feeds
    std::vector<int> comp_pending_parents;            // distinct outside producer tasks not yet successful
    std::vector<int> comp_remaining;                  // member tasks not yet successful
    std::vector<char> comp_started;
    std::vector<int> task_status;
    std::vector<int> ready_comps;                     // runnable, not yet handed out
    int status = LZYB_G_WAITING;
    int done = 0;
    bool cyclic = false;
    std::string error;
    int failed_task = -1;
};

std::shared_mutex g_mu;
std::unordered_map<lzyb_handle_t, std::shared_ptr<Graph>> g_graphs;
lzyb_handle_t g_next = 0x5000000000000001ull;

std::shared_ptr<Graph> get(lzyb_handle_t h) {
    std::shared_lock<std::shared_mutex> lk(g_mu);
    auto it = g_graphs.find(h);
    return it == g_graphs.end() ? nullptr : it->second;
}

// iterative Tarjan: components come out in reverse topological order
void condense(Graph& g) {
    const int n = g.n;
    std::vector<int> index(n, -1), low(n, 0), stack, it(n, 0);
    std::vector<char> on(n, 0);
    g.comp_of.assign(n, -1);
    int idx = 0;
    for (int s = 0; s < n; ++s) {
        if (index[s] != -1) continue;
        std::vector<int> call{s};
        index[s] = low[s] = idx++;
        stack.push_back(s);
        on[s] = 1;
        while (!call.empty()) {
            int v = call.back();
            if (it[v] < (int)g.children[v].size()) {
                int w = g.children[v][it[v]++];
                if (index[w] == -1) {
                    index[w] = low[w] = idx++;
                    stack.push_back(w);
                    on[w] = 1;
                    call.push_back(w);
                } else if (on[w]) {
                    low[v] = std::min(low[v], index[w]);
                }
            } else {
                if (low[v] == index[v]) {
                    std::vector<int> comp;
                    int w;
                    do {
                        w = stack.back();
                        stack.pop_back();
                        on[w] = 0;
                        g.comp_of[w] = (int)g.comp_tasks.size();
                        comp.push_back(w);
                    } while (w != v);
                    std::sort(comp.begin(), comp.end());
                    if (comp.size() > 1) g.cyclic = true;
                    g.comp_tasks.push_back(std::move(comp));
                }
                call.pop_back();
                if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[v]);
            }
        }
    }
    const int nc = (int)g.comp_tasks.size();
    g.task_child_comps.assign(n, {});
    g.comp_pending_parents.assign(nc, 0);
    g.comp_remaining.assign(nc, 0);
    g.comp_started.assign(nc, 0);
    for (int c = 0; c < nc; ++c) g.comp_remaining[c] = (int)g.comp_tasks[c].size();
    for (int v = 0; v < n; ++v)
        for (int w : g.children[v]) {
            if (v == w) g.cyclic = true;
            // readiness is decided per channel edge on the *producer task* (DirectChannelChecker.ready,
            // BfsGraphProcessor.java:236-246), not on the producer's whole component
            int a = g.comp_of[v], b = g.comp_of[w];
            if (a != b && g.task_child_comps[v].insert(b).second) g.comp_pending_parents[b]++;
        }
    for (int c = 0; c < nc; ++c)
        if (g.comp_pending_parents[c] == 0) g.ready_comps.push_back(c);  // Algorithms.findRoots
    std::sort(g.ready_comps.begin(), g.ready_comps.end(),
              [&](int a, int b) { return g.comp_tasks[a][0] < g.comp_tasks[b][0]; });
}

// GraphExecutionState.currentExecutionGroup is kept per condensed component: a component stays in
// the group until *all* of its members succeeded (BfsGraphProcessor.java:214-227;
// GraphExecutorTest.testCycles expects {1, 2} to be "executing now" after 2 alone completed)
int collect_executing(Graph& g, int32_t* out, int cap) {
    int k = 0;
    for (int v = 0; v < g.n; ++v) {
        const int c = g.comp_of[v];
        const bool in_group = g.comp_started[c] && g.comp_remaining[c] > 0 && g.task_status[v] != LZYB_T_ERROR;
        if (in_group) {
            if (out && k < cap) out[k] = v;
            ++k;
        }
    }
    return k;
}

int collect_running(Graph& g, int32_t* out, int cap) {
    int k = 0;
    for (int v = 0; v < g.n; ++v)
        if (g.task_status[v] == LZYB_T_EXECUTING) {
            if (out && k < cap) out[k] = v;
            ++k;
        }
    return k;
}

}  // namespace

extern "C" {

int lzyb_graph_create(int n_tasks, const int32_t* from, const int32_t* to, int n_edges, int allow_cycles, lzyb_handle_t* out) {
    if (n_tasks < 0 || n_edges < 0 || !out || (n_edges && (!from || !to))) return fail(LZYB_E_ARG, "bad graph arguments");
    auto g = std::make_shared<Graph>();
    g->n = n_tasks;
    g->children.assign(n_tasks, {});
    g->parents.assign(n_tasks, {});
    g->task_status.assign(n_tasks, LZYB_T_WAITING);
    for (int e = 0; e < n_edges; ++e) {
        if (from[e] < 0 || from[e] >= n_tasks || to[e] < 0 || to[e] >= n_tasks)
            return fail(LZYB_E_INVALID, "Validation error: edge %d references an unknown task", e);
        g->children[from[e]].push_back(to[e]);
        g->parents[to[e]].push_back(from[e]);
    }
    for (auto* adj : {&g->children, &g->parents})
        for (auto& v : *adj) {
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
        }
    condense(*g);
    if (g->cyclic && !allow_cycles)  // graph-executor-2 rejects cyclic task graphs (Algorithms.java:75-108)
        return fail(LZYB_E_INVALID, "Validation error: task graph has a cycle");
    if (n_tasks == 0) g->status = LZYB_G_COMPLETED;
    std::unique_lock<std::shared_mutex> lk(g_mu);
    lzyb_handle_t h = g_next++;
    g_graphs[h] = g;
    *out = h;
    return LZYB_OK;
}

int lzyb_graph_free(lzyb_handle_t h) {
    std::unique_lock<std::shared_mutex> lk(g_mu);
    g_graphs.erase(h);
    return LZYB_OK;
}

int lzyb_graph_next(lzyb_handle_t h, int32_t* runnable_out, int cap) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->status == LZYB_G_FAILED || g->status == LZYB_G_COMPLETED) return 0;
    int k = 0;
    size_t used = 0;
    for (; used < g->ready_comps.size(); ++used) {
        int c = g->ready_comps[used];
        if (k + (int)g->comp_tasks[c].size() > cap) break;  // hand out whole components only
        g->comp_started[c] = 1;
        for (int v : g->comp_tasks[c]) {
            g->task_status[v] = LZYB_T_EXECUTING;
            runnable_out[k++] = v;
        }
    }
    g->ready_comps.erase(g->ready_comps.begin(), g->ready_comps.begin() + used);
    if (k > 0) g->status = LZYB_G_EXECUTING;
    return k;
}

int lzyb_graph_task_done(lzyb_handle_t h, int32_t task, int ok, const char* error_description) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    if (task < 0 || task >= g->n) return fail(LZYB_E_ARG, "unknown task %d", task);
    if (g->status == LZYB_G_FAILED) {
        if (g->task_status[task] == LZYB_T_EXECUTING) g->task_status[task] = ok ? LZYB_T_SUCCESS : LZYB_T_ERROR;
        return LZYB_OK;
    }
    if (!ok) {
        // "TaskVertex <id> is in error state, stopping graph execution": FAILED, everything still running is killed
        g->task_status[task] = LZYB_T_ERROR;
        g->status = LZYB_G_FAILED;
        g->error = error_description ? error_description : "";
        g->failed_task = task;
        g->ready_comps.clear();
        return LZYB_OK;
    }
    if (g->task_status[task] != LZYB_T_EXECUTING)
        return fail(LZYB_E_STATE, "task %d is not executing (status %d)", task, g->task_status[task]);
    g->task_status[task] = LZYB_T_SUCCESS;
    g->done++;
    int c = g->comp_of[task];
    --g->comp_remaining[c];
    std::vector<int> fresh;
    for (int ch : g->task_child_comps[task])
        if (--g->comp_pending_parents[ch] == 0) fresh.push_back(ch);
    std::sort(fresh.begin(), fresh.end(), [&](int a, int b) { return g->comp_tasks[a][0] < g->comp_tasks[b][0]; });
    g->ready_comps.insert(g->ready_comps.end(), fresh.begin(), fresh.end());
    if (g->done == g->n) g->status = LZYB_G_COMPLETED;
    return LZYB_OK;
}

int lzyb_graph_executing(lzyb_handle_t h, int32_t* out, int cap) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    return collect_executing(*g, out, cap);
}

int lzyb_graph_status(lzyb_handle_t h) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    return g->status;
}

int lzyb_graph_task_status(lzyb_handle_t h, int32_t task) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    if (task < 0 || task >= g->n) return fail(LZYB_E_ARG, "unknown task %d", task);
    return g->task_status[task];
}

int lzyb_graph_stop(lzyb_handle_t h, const char* description, int32_t* kill_out, int cap) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->status == LZYB_G_COMPLETED || g->status == LZYB_G_FAILED) return 0;
    int k = collect_running(*g, kill_out, cap);
    g->status = LZYB_G_FAILED;
    g->error = description ? description : "";
    g->ready_comps.clear();
    // api.kill(...) for every execution: the scheduler mock then reports ERROR for them
    // (GraphExecutorTest.testStop waits for ERROR on 1, 2, 3)
    for (int v = 0; v < g->n; ++v)
        if (g->task_status[v] == LZYB_T_EXECUTING) g->task_status[v] = LZYB_T_ERROR;
    return k;
}

int lzyb_graph_error(lzyb_handle_t h, char* description_out, uint64_t cap, int32_t* failed_task_out) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    if (description_out && cap) lzyb_host::scopy(description_out, cap, g->error);
    if (failed_task_out) *failed_task_out = g->failed_task;
    return g->status == LZYB_G_FAILED ? 1 : 0;
}

int lzyb_graph_has_cycle(lzyb_handle_t h) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    return g->cyclic ? 1 : 0;
}

int lzyb_graph_topo_order(lzyb_handle_t h, int32_t* order_out, int cap) {
    auto g = get(h);
    if (!g) return fail(LZYB_E_ARG, "unknown graph handle");
    std::lock_guard<std::mutex> lk(g->mu);
    if (cap < g->n) return fail(LZYB_E_ARG, "order buffer too small");
    // Kahn over tasks with smallest-id-first tie break: one valid sequential order (the local
    // runtime runs ops one at a time in *a* topological order, local/runtime.py:62-85)
    std::vector<int> indeg(g->n, 0);
    for (int v = 0; v < g->n; ++v) indeg[v] = (int)g->parents[v].size();
    std::set<int> ready;
    for (int v = 0; v < g->n; ++v)
        if (!indeg[v]) ready.insert(v);
    int k = 0;
    while (!ready.empty()) {
        int v = *ready.begin();
        ready.erase(ready.begin());
        order_out[k++] = v;
        for (int w : g->children[v])
            if (--indeg[w] == 0) ready.insert(w);
    }
    return k;  // < n when the graph is cyclic
}

}  // extern "C"
