// local_rpc.cpp — the matcher's four calls and LzybPeerApi.Describe over a unix socket, for the processes of ONE box.
//
// The kept wire contract is gRPC: a stock worker or channel manager reaches this repository's services through
// lzy_b200/grpc_api, and that stays so.  But between this repository's OWN processes on one box — a worker process per
// GPU and the channel manager (lzy/worker/.../Worker.java:36-56 is that shape) — every input slot costs Bind + Describe +
// TransferCompleted and every output slot Bind + TransferCompleted, and through two Python interpreters a unary call is
// 0.5–0.9 ms (tools/worker_chain_cpu.py: six sequential calls were most of a task's 9 ms).  The same calls, with the same
// arguments, result codes and error texts, travel here as length-prefixed frames over a SOCK_STREAM unix socket, served by
// threads of this library: no interpreter on either side.
//
//   server side   lzyb_cm_serve_local(cm, path)        Bind / Unbind / TransferCompleted / TransferFailed of that matcher
//                 lzyb_slots_serve_local(slots, path)  Describe(peer_id) → lzyb_ipc_desc_t of a producer slot of that worker
//   client side   lzyb_slots_connect_cm_local(path, &cookie)  installs the client table of the slots runtime's seam
//                 (lzyb_slots_set_cm_client); the cookie stands where a matcher handle would (lzyb_slots_create)
//                 a peer URL that carries "uds=<path>" is resolved over that socket (slots_runtime.cpp: make_transfer)
//
// Deferred actions (the matcher's StartTransfer towards an early consumer) keep travelling as LzySlotsApi.StartTransfer.
//
// Frame:   u32 length of what follows | u32 op | fields: (u32 length, bytes)*
// Answer:  u32 length of what follows | i32 rc | u32 length, error text | u32 length, payload (the raw result struct)
// Both ends are the same build on the same machine: structs travel as they lie in memory.
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../../include/lzy_b200.h"
#include "../../../include/lzy_b200_host.h"
#include "host_common.h"
#include "local_rpc.h"

namespace {

using lzyb_host::fail;

enum Op : uint32_t { OP_BIND = 1, OP_UNBIND = 2, OP_COMPLETED = 3, OP_FAILED = 4, OP_DESCRIBE = 5 };
constexpr uint32_t MAX_FRAME = 1u << 20;
constexpr long CALL_TIMEOUT_S = 30;

// ---- plumbing --------------------------------------------------------------------------------------------------------
bool read_all(int fd, void* buf, size_t n) {
    auto* p = static_cast<uint8_t*>(buf);
    while (n) {
        ssize_t r = ::recv(fd, p, n, 0);
        if (r == 0) return false;
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += r;
        n -= static_cast<size_t>(r);
    }
    return true;
}

bool write_all(int fd, const void* buf, size_t n) {
    auto* p = static_cast<const uint8_t*>(buf);
    while (n) {
        ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += r;
        n -= static_cast<size_t>(r);
    }
    return true;
}

void put_u32(std::vector<uint8_t>& b, uint32_t v) {
    const auto* p = reinterpret_cast<const uint8_t*>(&v);
    b.insert(b.end(), p, p + 4);
}

void put_field(std::vector<uint8_t>& b, const void* data, size_t n) {
    put_u32(b, static_cast<uint32_t>(n));
    const auto* p = static_cast<const uint8_t*>(data);
    b.insert(b.end(), p, p + n);
}

void put_str(std::vector<uint8_t>& b, const char* s) { put_field(b, s ? s : "", s ? strlen(s) : 0); }

struct Reader {
    const uint8_t* p;
    size_t left;
    bool ok = true;
    uint32_t u32() {
        if (left < 4) {
            ok = false;
            return 0;
        }
        uint32_t v;
        memcpy(&v, p, 4);
        p += 4;
        left -= 4;
        return v;
    }
    std::string str() {
        uint32_t n = u32();
        if (!ok || left < n) {
            ok = false;
            return {};
        }
        std::string s(reinterpret_cast<const char*>(p), n);
        p += n;
        left -= n;
        return s;
    }
};

int make_addr(const char* path, sockaddr_un* addr) {
    if (!path || !*path) return fail(LZYB_E_ARG, "empty socket path");
    if (strlen(path) >= sizeof(addr->sun_path)) return fail(LZYB_E_ARG, "socket path longer than %zu bytes: %s", sizeof(addr->sun_path) - 1, path);
    memset(addr, 0, sizeof(*addr));
    addr->sun_family = AF_UNIX;
    strcpy(addr->sun_path, path);
    return LZYB_OK;
}

// ---- server ----------------------------------------------------------------------------------------------------------
struct Server {
    lzyb_handle_t target = 0;  // the matcher, or the slots instance
    bool is_cm = true;
    std::string path;
    int listen_fd = -1;
    std::atomic<bool> stopping{false};
    std::thread acceptor;
    struct Conn {
        int fd = -1;  // −1 once its thread has closed it (guarded by `mu`: stop() must not shut down a recycled descriptor)
        std::thread th;
    };
    std::mutex mu;
    std::vector<std::shared_ptr<Conn>> conns;

    void answer(int fd, int rc, const void* payload, size_t payload_len) {
        std::vector<uint8_t> out;
        out.reserve(64 + payload_len);
        put_u32(out, 0);  // length of what follows, patched below
        put_u32(out, static_cast<uint32_t>(static_cast<int32_t>(rc)));
        put_str(out, rc == LZYB_OK ? "" : lzyb_last_error());
        put_field(out, payload, rc == LZYB_OK ? payload_len : 0);
        const uint32_t len = static_cast<uint32_t>(out.size() - 4);
        memcpy(out.data(), &len, 4);
        write_all(fd, out.data(), out.size());
    }

    void serve(const std::shared_ptr<Conn>& conn) {
        const int fd = conn->fd;
        std::vector<uint8_t> buf;
        for (;;) {
            uint32_t len = 0;
            if (!read_all(fd, &len, 4) || len < 4 || len > MAX_FRAME) break;
            buf.resize(len);
            if (!read_all(fd, buf.data(), len)) break;
            Reader r{buf.data(), buf.size()};
            const uint32_t op = r.u32();
            if (is_cm && (op == OP_BIND || op == OP_FAILED)) {
                lzyb_bind_result_t res;
                memset(&res, 0, sizeof(res));
                int rc;
                if (op == OP_BIND) {
                    std::string idk = r.str(), channel = r.str(), peer = r.str(), exec = r.str(), url = r.str(), role = r.str();
                    rc = r.ok ? lzyb_cm_bind(target, idk.empty() ? nullptr : idk.c_str(), channel.c_str(), peer.c_str(), exec.c_str(), url.c_str(),
                                             atoi(role.c_str()), &res)
                              : fail(LZYB_E_ARG, "malformed Bind frame");
                } else {
                    std::string idk = r.str(), channel = r.str(), tid = r.str(), desc = r.str();
                    rc = r.ok ? lzyb_cm_transfer_failed(target, idk.empty() ? nullptr : idk.c_str(), channel.c_str(), tid.c_str(), desc.c_str(), &res)
                              : fail(LZYB_E_ARG, "malformed TransferFailed frame");
                }
                answer(fd, rc, &res, sizeof(res));
            } else if (is_cm && op == OP_UNBIND) {
                std::string channel = r.str(), peer = r.str();
                answer(fd, r.ok ? lzyb_cm_unbind(target, channel.c_str(), peer.c_str()) : fail(LZYB_E_ARG, "malformed Unbind frame"), nullptr, 0);
            } else if (is_cm && op == OP_COMPLETED) {
                std::string idk = r.str(), channel = r.str(), tid = r.str();
                answer(fd,
                       r.ok ? lzyb_cm_transfer_completed(target, idk.empty() ? nullptr : idk.c_str(), channel.c_str(), tid.c_str())
                            : fail(LZYB_E_ARG, "malformed TransferCompleted frame"),
                       nullptr, 0);
            } else if (!is_cm && op == OP_DESCRIBE) {
                std::string peer = r.str();
                lzyb_ipc_desc_t desc;
                memset(&desc, 0, sizeof(desc));
                answer(fd, r.ok ? lzyb_slots_describe(target, peer.c_str(), &desc) : fail(LZYB_E_ARG, "malformed Describe frame"), &desc, sizeof(desc));
            } else {
                answer(fd, fail(LZYB_E_UNIMPLEMENTED, "this socket does not serve call %u", op), nullptr, 0);
            }
        }
        std::lock_guard<std::mutex> lk(mu);
        ::close(fd);
        conn->fd = -1;
    }

    void accept_loop() {
        for (;;) {
            int fd = ::accept(listen_fd, nullptr, nullptr);
            if (fd < 0) {
                if (errno == EINTR) continue;
                break;  // the listening socket was shut down
            }
            std::lock_guard<std::mutex> lk(mu);
            if (stopping.load()) {
                ::close(fd);
                break;
            }
            // connections whose thread has finished (a client thread that went away) are reaped here: their last locked act was
            // to mark the descriptor closed, so the join returns at once
            for (size_t i = 0; i < conns.size();) {
                if (conns[i]->fd < 0) {
                    if (conns[i]->th.joinable()) conns[i]->th.join();
                    conns.erase(conns.begin() + static_cast<long>(i));
                } else {
                    ++i;
                }
            }
            auto conn = std::make_shared<Conn>();
            conn->fd = fd;
            conns.push_back(conn);
            conn->th = std::thread([this, conn] { serve(conn); });
        }
    }

    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stopping = true;
        }
        if (listen_fd >= 0) {
            ::shutdown(listen_fd, SHUT_RDWR);
            ::close(listen_fd);
        }
        if (acceptor.joinable()) acceptor.join();
        std::vector<std::shared_ptr<Conn>> mine;
        {
            std::lock_guard<std::mutex> lk(mu);
            mine = conns;
            for (auto& c : mine)
                if (c->fd >= 0) ::shutdown(c->fd, SHUT_RDWR);  // wakes a serve() blocked in recv; it closes its own descriptor
        }
        for (auto& c : mine)
            if (c->th.joinable()) c->th.join();
        ::unlink(path.c_str());
    }
};

std::mutex g_servers_mu;
std::map<std::pair<bool, lzyb_handle_t>, std::shared_ptr<Server>> g_servers;  // (is_cm, handle) → its server

int serve_local(lzyb_handle_t target, bool is_cm, const char* path) {
    sockaddr_un addr;
    if (int rc = make_addr(path, &addr); rc != LZYB_OK) return rc;
    std::lock_guard<std::mutex> lk(g_servers_mu);
    if (g_servers.count({is_cm, target})) return fail(LZYB_E_EXISTS, "already serving on a local socket");
    int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return fail(LZYB_E_LOCAL, "socket: %s", strerror(errno));
    ::unlink(path);  // a stale file of a previous incarnation
    if (::bind(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || ::listen(fd, 128) != 0) {
        int e = errno;
        ::close(fd);
        return fail(LZYB_E_LOCAL, "cannot listen on %s: %s", path, strerror(e));
    }
    auto s = std::make_shared<Server>();
    s->target = target;
    s->is_cm = is_cm;
    s->path = path;
    s->listen_fd = fd;
    s->acceptor = std::thread([s] { s->accept_loop(); });
    g_servers[{is_cm, target}] = s;
    return LZYB_OK;
}

int stop_local(lzyb_handle_t target, bool is_cm) {
    std::shared_ptr<Server> s;
    {
        std::lock_guard<std::mutex> lk(g_servers_mu);
        auto it = g_servers.find({is_cm, target});
        if (it == g_servers.end()) return LZYB_OK;
        s = it->second;
        g_servers.erase(it);
    }
    s->stop();
    return LZYB_OK;
}

// ---- client ----------------------------------------------------------------------------------------------------------
// One connection per (calling thread, socket path): the slot state machines run on recycled threads of their own, so a
// connection lives as long as its thread and nobody shares it.
struct Conns {
    std::map<std::string, int> fds;
    ~Conns() {
        for (auto& kv : fds) ::close(kv.second);
    }
};
thread_local Conns t_conns;

int dial(const std::string& path) {
    sockaddr_un addr;
    if (make_addr(path.c_str(), &addr) != LZYB_OK) return -1;
    int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return -1;
    if (::connect(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
        ::close(fd);
        return -1;
    }
    // the matcher answers in microseconds; a peer that has not answered for this long is gone (the gRPC clients give a call 30 s too)
    timeval tv{CALL_TIMEOUT_S, 0};
    ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    ::setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    return fd;
}

// Sends `frame` (already length-prefixed), receives the answer into `payload` (exactly payload_len bytes when rc == OK and a
// payload is expected).  A connection that turns out dead before any byte of the answer arrived is re-dialled once: every
// call is idempotent at the matcher (idempotency keys; Unbind of an unbound peer is NOT_FOUND either way).
int call(const std::string& path, const std::vector<uint8_t>& frame, void* payload, size_t payload_len) {
    for (int attempt = 0; attempt < 2; ++attempt) {
        auto it = t_conns.fds.find(path);
        int fd = it == t_conns.fds.end() ? -1 : it->second;
        if (fd < 0) {
            fd = dial(path);
            if (fd < 0) return fail(LZYB_E_UNAVAILABLE, "cannot reach %s: %s", path.c_str(), strerror(errno));
            t_conns.fds[path] = fd;
        }
        uint32_t len = 0;
        if (!write_all(fd, frame.data(), frame.size()) || !read_all(fd, &len, 4)) {
            ::close(fd);
            t_conns.fds.erase(path);
            continue;
        }
        std::vector<uint8_t> buf(len);
        if (len < 12 || len > MAX_FRAME || !read_all(fd, buf.data(), len)) {
            ::close(fd);
            t_conns.fds.erase(path);
            return fail(LZYB_E_UNAVAILABLE, "connection to %s broke inside an answer", path.c_str());
        }
        Reader r{buf.data(), buf.size()};
        const int32_t rc = static_cast<int32_t>(r.u32());
        std::string err = r.str();
        std::string body = r.str();
        if (!r.ok) return fail(LZYB_E_INTERNAL, "malformed answer from %s", path.c_str());
        if (rc != LZYB_OK) return fail(rc, "%s", err.c_str());
        if (payload && payload_len) {
            if (body.size() != payload_len) return fail(LZYB_E_INTERNAL, "answer from %s carries %zu bytes, expected %zu", path.c_str(), body.size(), payload_len);
            memcpy(payload, body.data(), payload_len);
        }
        return LZYB_OK;
    }
    return fail(LZYB_E_UNAVAILABLE, "cannot reach %s", path.c_str());
}

std::vector<uint8_t> begin(uint32_t op) {
    std::vector<uint8_t> f;
    f.reserve(1024);
    put_u32(f, 0);
    put_u32(f, op);
    return f;
}

const std::vector<uint8_t>& seal(std::vector<uint8_t>& f) {
    uint32_t len = static_cast<uint32_t>(f.size() - 4);
    memcpy(f.data(), &len, 4);
    return f;
}

constexpr lzyb_handle_t COOKIE_BASE = 0x10CA10000ull;
std::mutex g_paths_mu;
std::vector<std::string> g_paths;  // cookie − COOKIE_BASE → socket path of that channel manager

bool path_of(lzyb_handle_t cookie, std::string* out) {
    std::lock_guard<std::mutex> lk(g_paths_mu);
    if (cookie < COOKIE_BASE || cookie - COOKIE_BASE >= g_paths.size()) return false;
    *out = g_paths[cookie - COOKIE_BASE];
    return true;
}

int c_bind(lzyb_handle_t cm, const char* idk, const char* channel_id, const char* peer_id, const char* execution_id, const char* peer_url,
           int role, lzyb_bind_result_t* out) {
    std::string path;
    if (!path_of(cm, &path)) return fail(LZYB_E_ARG, "not a local channel-manager connection");
    auto f = begin(OP_BIND);
    put_str(f, idk);
    put_str(f, channel_id);
    put_str(f, peer_id);
    put_str(f, execution_id);
    put_str(f, peer_url);
    put_str(f, std::to_string(role).c_str());
    return call(path, seal(f), out, sizeof(*out));
}

int c_unbind(lzyb_handle_t cm, const char* channel_id, const char* peer_id) {
    std::string path;
    if (!path_of(cm, &path)) return fail(LZYB_E_ARG, "not a local channel-manager connection");
    auto f = begin(OP_UNBIND);
    put_str(f, channel_id);
    put_str(f, peer_id);
    return call(path, seal(f), nullptr, 0);
}

int c_completed(lzyb_handle_t cm, const char* idk, const char* channel_id, const char* transfer_id) {
    std::string path;
    if (!path_of(cm, &path)) return fail(LZYB_E_ARG, "not a local channel-manager connection");
    auto f = begin(OP_COMPLETED);
    put_str(f, idk);
    put_str(f, channel_id);
    put_str(f, transfer_id);
    return call(path, seal(f), nullptr, 0);
}

int c_failed(lzyb_handle_t cm, const char* idk, const char* channel_id, const char* transfer_id, const char* description,
             lzyb_bind_result_t* out) {
    std::string path;
    if (!path_of(cm, &path)) return fail(LZYB_E_ARG, "not a local channel-manager connection");
    auto f = begin(OP_FAILED);
    put_str(f, idk);
    put_str(f, channel_id);
    put_str(f, transfer_id);
    put_str(f, description);
    return call(path, seal(f), out, sizeof(*out));
}

const lzyb_cm_client_ops_t kLocalSocketClient = {c_bind, c_unbind, c_completed, c_failed};

}  // namespace

namespace lzyb_host {

bool local_peer_socket(const char* peer_url, std::string* path) {
    // "...?gpu=3&uds=/tmp/w3.sock" — the value runs to the next '&' or the end
    if (!peer_url) return false;
    const char* q = strchr(peer_url, '?');
    while (q) {
        ++q;
        if (strncmp(q, "uds=", 4) == 0) {
            const char* e = strchr(q, '&');
            path->assign(q + 4, e ? static_cast<size_t>(e - q - 4) : strlen(q + 4));
            return !path->empty();
        }
        q = strchr(q, '&');
    }
    return false;
}

lzyb_handle_t local_matcher_behind(lzyb_handle_t cookie) {
    std::string path;
    if (!path_of(cookie, &path)) return 0;
    std::lock_guard<std::mutex> lk(g_servers_mu);
    for (auto& kv : g_servers)
        if (kv.first.first && kv.second->path == path) return kv.second->target;
    return 0;
}

int local_describe(const std::string& socket_path, const char* peer_id, lzyb_ipc_desc_t* out) {
    auto f = begin(OP_DESCRIBE);
    put_str(f, peer_id);
    return call(socket_path, seal(f), out, sizeof(*out));
}

}  // namespace lzyb_host

extern "C" {

int lzyb_cm_serve_local(lzyb_handle_t cm, const char* socket_path) { return serve_local(cm, true, socket_path); }
int lzyb_cm_stop_local(lzyb_handle_t cm) { return stop_local(cm, true); }
int lzyb_slots_serve_local(lzyb_handle_t slots, const char* socket_path) { return serve_local(slots, false, socket_path); }
int lzyb_slots_stop_local(lzyb_handle_t slots) { return stop_local(slots, false); }

int lzyb_slots_connect_cm_local(const char* socket_path, lzyb_handle_t* cm_cookie) {
    if (!cm_cookie) return fail(LZYB_E_ARG, "null argument");
    sockaddr_un addr;
    if (int rc = make_addr(socket_path, &addr); rc != LZYB_OK) return rc;
    int fd = dial(socket_path);  // fail now, not inside the first slot's Bind
    if (fd < 0) return fail(LZYB_E_UNAVAILABLE, "cannot reach the channel manager at %s: %s", socket_path, strerror(errno));
    ::close(fd);
    {
        std::lock_guard<std::mutex> lk(g_paths_mu);
        g_paths.emplace_back(socket_path);
        *cm_cookie = COOKIE_BASE + g_paths.size() - 1;
    }
    return lzyb_slots_set_cm_client(&kLocalSocketClient);
}

int lzyb_slots_disconnect_cm_local(void) { return lzyb_slots_set_cm_client(nullptr); }

}  // extern "C"
