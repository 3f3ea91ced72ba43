// channel_manager.cpp — in-memory peer/transfer matcher with the RPC semantics of the reference's
// channel manager.  One box, no multi-tenant database: the Postgres tables
// (lzy/channel-manager/src/main/resources/db/channel-manager/migrations/V1_0__Init_tables.sql)
// become three maps guarded by one mutex, and every "transaction" is a critical section.
//
// Behaviour follows (paths relative to lzy/channel-manager/src/main/java/ai/lzy/channelmanager/):
//   services/SlotsService.java    bind :67-145, unbind :148-188, transferFailed :191-291,
//                                 transferCompleted :294-358, getChannelsStatus :361-407,
//                                 createProducer :440-478, createConsumer :489-519
//   services/ChannelService.java  getOrCreate :39-116, destroy :119-136, destroyAll :139-155
//   ActionScheduler.java          runStartTransferAction/startTransfer :43-106, restoreActions :47-59
//   db/PeerDao.java               Priority PRIMARY=7 / BACKUP=5 / DONT_USE=-1 :44-53
//   db/PeerDaoImpl.java           FIND_PRODUCER: priority >= 0, ORDER BY priority DESC, RANDOM() :61-65
//   LzyServiceClient.java         destroyChannelAndWorkflow :37-55
// Error codes are the images of the gRPC statuses the reference returns (include/lzy_b200.h).
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <shared_mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/lzy_b200_host.h"
#include "host_common.h"

namespace {

using lzyb_host::fail;
using lzyb_host::scopy;

struct PeerRec {
    std::string id, channel_id;
    int role = 0;
    int kind = 0;
    std::string url;
    int priority = LZYB_PRIO_PRIMARY;
    bool connected = false;
    std::string idk, request_hash;
    std::string conn_idk, conn_request_hash;
    uint64_t order = 0;  // insertion order (stable listing)
};

struct TransferRec {
    std::string id, channel_id, from_id, to_id;
    int state = LZYB_TR_PENDING;
    std::string error, idk, request_hash, state_change_idk;
};

struct ChannelRec {
    std::string id, user_id, execution_id, workflow_name;
    std::string data_format, scheme_format, scheme_content;
    std::string storage_producer_uri, storage_consumer_uri;
    std::map<std::string, PeerRec> peers;          // by peer id
    std::map<std::string, TransferRec> transfers;  // by transfer id
};

struct ChannelManager {
    std::mutex mu;
    std::condition_variable cv;
    std::map<std::string, ChannelRec> channels;
    std::deque<lzyb_cm_action_t> actions;
    std::map<std::string, std::string> aborted;  // execution id -> reason
    std::mt19937_64 rng{0x1a2b3c4d5e6f7081ull};
    uint64_t seq = 0;

    std::string gen(const char* prefix) {
        char buf[96];
        snprintf(buf, sizeof(buf), "%s%016llx-%llu", prefix, (unsigned long long)rng(), (unsigned long long)++seq);
        return buf;
    }
};

std::shared_mutex g_mu;
std::unordered_map<lzyb_handle_t, std::shared_ptr<ChannelManager>> g_cms;
lzyb_handle_t g_next = 0x4000000000000001ull;

std::shared_ptr<ChannelManager> get(lzyb_handle_t h) {
    std::shared_lock<std::shared_mutex> lk(g_mu);
    auto it = g_cms.find(h);
    return it == g_cms.end() ? nullptr : it->second;
}

void fill_peer(const PeerRec& p, lzyb_peer_desc_t* out) {
    memset(out, 0, sizeof(*out));
    scopy(out->peer_id, p.id);
    out->kind = p.kind;
    scopy(out->url, p.url);
}

// request hash = IdempotencyUtils.md5(request): any stable function of the request fields will do
std::string req_hash(std::initializer_list<std::string> parts) {
    uint64_t h = 1469598103934665603ull;
    for (auto& s : parts) {
        for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
        h = (h ^ 0xff) * 1099511628211ull;
    }
    char buf[32];
    snprintf(buf, sizeof(buf), "%016llx", (unsigned long long)h);
    return buf;
}

// PeerDao.create — idempotent on (id, channel) for the same (idempotency key, request hash)
int peer_create(ChannelRec& ch, const std::string& id, int role, int kind, const std::string& url, int priority,
                bool connected, const std::string& idk, const std::string& rh, ChannelManager& cm) {
    auto it = ch.peers.find(id);
    if (it != ch.peers.end()) {
        if (it->second.idk != idk || it->second.request_hash != rh)
            return fail(LZYB_E_INTERNAL, "Cannot save peer description in db: Peer with id='%s' already associated with channel with id='%s'",
                        id.c_str(), ch.id.c_str());
        return LZYB_OK;
    }
    PeerRec p;
    p.id = id;
    p.channel_id = ch.id;
    p.role = role;
    p.kind = kind;
    p.url = url;
    p.priority = priority;
    p.connected = connected;
    p.idk = idk;
    p.request_hash = rh;
    p.order = ++cm.seq;
    ch.peers.emplace(id, std::move(p));
    return LZYB_OK;
}

// PeerDao.drop with ON DELETE CASCADE on transfers(from_id|to_id)
void peer_drop(ChannelRec& ch, const std::string& id) {
    ch.peers.erase(id);
    for (auto it = ch.transfers.begin(); it != ch.transfers.end();) {
        if (it->second.from_id == id || it->second.to_id == id)
            it = ch.transfers.erase(it);
        else
            ++it;
    }
}

// FIND_PRODUCER (PeerDaoImpl.java:61-65)
PeerRec* find_producer(ChannelRec& ch, ChannelManager& cm) {
    int best = -1;
    std::vector<PeerRec*> cands;
    for (auto& kv : ch.peers) {
        PeerRec& p = kv.second;
        if (p.role != LZYB_ROLE_PRODUCER || p.priority < 0) continue;
        if (p.priority > best) {
            best = p.priority;
            cands.clear();
        }
        if (p.priority == best) cands.push_back(&p);
    }
    if (cands.empty()) return nullptr;
    return cands[cm.rng() % cands.size()];
}

// TransferDao.create — idempotent per (idempotency key): a replay returns the same id
std::string transfer_create(ChannelRec& ch, const std::string& from, const std::string& to, int state,
                            const std::string& idk, const std::string& rh, ChannelManager& cm) {
    for (auto& kv : ch.transfers)
        if (kv.second.idk == idk && kv.second.request_hash == rh && kv.second.from_id == from && kv.second.to_id == to)
            return kv.first;
    TransferRec t;
    t.id = cm.gen("transfer-");
    t.channel_id = ch.id;
    t.from_id = from;
    t.to_id = to;
    t.state = state;
    t.idk = idk;
    t.request_hash = rh;
    std::string id = t.id;
    ch.transfers.emplace(id, std::move(t));
    return id;
}

void queue_abort(ChannelManager& cm, const ChannelRec& ch, const std::string& reason, const std::string& idk) {
    lzyb_cm_action_t a;
    memset(&a, 0, sizeof(a));
    a.kind = LZYB_ACT_ABORT_WORKFLOW;
    scopy(a.channel_id, ch.id);
    scopy(a.execution_id, ch.execution_id);
    scopy(a.workflow_name, ch.workflow_name);
    scopy(a.reason, reason);
    scopy(a.idempotency_key, idk);
    cm.actions.push_back(a);
    cm.aborted[ch.execution_id] = reason;
    cm.cv.notify_all();
}

// LzyServiceClient.destroyChannelAndWorkflow: drop the channel, abort the workflow
void destroy_channel_and_workflow(ChannelManager& cm, const std::string& channel_id, const std::string& reason,
                                  const std::string& idk) {
    auto it = cm.channels.find(channel_id);
    if (it == cm.channels.end()) return;
    ChannelRec ch = it->second;
    cm.channels.erase(it);
    queue_abort(cm, ch, reason, idk);
}

void queue_start_transfer(ChannelManager& cm, const ChannelRec& ch, const std::string& transfer_id, const PeerRec& from,
                          const PeerRec& to, const std::string& idk) {
    lzyb_cm_action_t a;
    memset(&a, 0, sizeof(a));
    a.kind = LZYB_ACT_START_TRANSFER;
    scopy(a.channel_id, ch.id);
    scopy(a.transfer_id, transfer_id);
    scopy(a.consumer_id, to.id);
    scopy(a.consumer_url, to.url);
    fill_peer(from, &a.producer);
    // the reference dials the producer's URL with slot_id = consumer id (ActionScheduler.java:64,74-77);
    // kept visible here, the in-process dispatcher resolves the consumer by slot id
    scopy(a.dial_url, from.url);
    scopy(a.execution_id, ch.execution_id);
    scopy(a.workflow_name, ch.workflow_name);
    scopy(a.idempotency_key, idk);
    cm.actions.push_back(a);
    cm.cv.notify_all();
}

void json_escape(std::ostringstream& os, const std::string& s) {
    os << '"';
    for (char c : s) {
        switch (c) {
            case '"': os << "\\\""; break;
            case '\\': os << "\\\\"; break;
            case '\n': os << "\\n"; break;
            case '\t': os << "\\t"; break;
            default:
                if ((unsigned char)c < 0x20) {
                    char b[8];
                    snprintf(b, sizeof(b), "\\u%04x", c);
                    os << b;
                } else
                    os << c;
        }
    }
    os << '"';
}

}  // namespace

extern "C" {

int lzyb_cm_create(lzyb_handle_t* out) {
    if (!out) return fail(LZYB_E_ARG, "null argument");
    std::unique_lock<std::shared_mutex> lk(g_mu);
    lzyb_handle_t h = g_next++;
    g_cms[h] = std::make_shared<ChannelManager>();
    *out = h;
    return LZYB_OK;
}

int lzyb_cm_free(lzyb_handle_t h) {
    std::unique_lock<std::shared_mutex> lk(g_mu);
    g_cms.erase(h);
    return LZYB_OK;
}

int lzyb_cm_seed(lzyb_handle_t h, uint64_t seed) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    std::lock_guard<std::mutex> lk(cm->mu);
    cm->rng.seed(seed);
    return LZYB_OK;
}

int lzyb_cm_get_or_create(lzyb_handle_t h, const char* user_id, const char* workflow_name, const char* execution_id,
                          const char* data_format, const char* scheme_format, const char* scheme_content,
                          int storage_role, const char* storage_uri, char channel_id_out[64]) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    if (!user_id || !workflow_name || !execution_id || !channel_id_out) return fail(LZYB_E_ARG, "null argument");
    if ((storage_role != LZYB_ROLE_PRODUCER && storage_role != LZYB_ROLE_CONSUMER) || !storage_uri)
        return fail(LZYB_E_INVALID, "Consumer or producer must be set");
    std::lock_guard<std::mutex> lk(cm->mu);
    const std::string prod = storage_role == LZYB_ROLE_PRODUCER ? storage_uri : "";
    const std::string cons = storage_role == LZYB_ROLE_CONSUMER ? storage_uri : "";
    // channelDao.find(user, execution, producerUri, consumerUri)
    for (auto& kv : cm->channels) {
        ChannelRec& c = kv.second;
        if (c.user_id == user_id && c.execution_id == execution_id && c.storage_producer_uri == prod &&
            c.storage_consumer_uri == cons) {
            scopy(channel_id_out, 64, c.id);
            return LZYB_OK;
        }
    }
    ChannelRec c;
    c.id = cm->gen("channel-");
    c.user_id = user_id;
    c.execution_id = execution_id;
    c.workflow_name = workflow_name;
    c.data_format = data_format ? data_format : "";
    c.scheme_format = scheme_format ? scheme_format : "";
    c.scheme_content = scheme_content ? scheme_content : "";
    c.storage_producer_uri = prod;
    c.storage_consumer_uri = cons;
    std::string peer_id = cm->gen("storage-peer-");
    int rc = peer_create(c, peer_id, storage_role, LZYB_PEER_STORAGE, storage_uri, LZYB_PRIO_BACKUP, false, cm->gen("idk-"),
                         req_hash({user_id, execution_id, prod, cons}), *cm);
    if (rc != LZYB_OK) return rc;
    scopy(channel_id_out, 64, c.id);
    std::string id = c.id;
    cm->channels.emplace(id, std::move(c));
    return LZYB_OK;
}

int lzyb_cm_destroy(lzyb_handle_t h, const char* channel_id, const char* /*reason*/) {
    auto cm = get(h);
    if (!cm || !channel_id) return fail(LZYB_E_ARG, "unknown channel manager handle");
    std::lock_guard<std::mutex> lk(cm->mu);
    cm->channels.erase(channel_id);  // "Just dropping channel from db here" (ChannelService.java:124-127)
    return LZYB_OK;
}

int lzyb_cm_destroy_all(lzyb_handle_t h, const char* execution_id, const char* /*reason*/) {
    auto cm = get(h);
    if (!cm || !execution_id) return fail(LZYB_E_ARG, "unknown channel manager handle");
    std::lock_guard<std::mutex> lk(cm->mu);
    for (auto it = cm->channels.begin(); it != cm->channels.end();) {
        if (it->second.execution_id == execution_id)
            it = cm->channels.erase(it);
        else
            ++it;
    }
    return LZYB_OK;
}

int lzyb_cm_bind(lzyb_handle_t h, const char* idempotency_key, const char* channel_id, const char* peer_id,
                 const char* execution_id, const char* peer_url, int role, lzyb_bind_result_t* out) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    if (!channel_id || !peer_id || !out) return fail(LZYB_E_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    std::lock_guard<std::mutex> lk(cm->mu);
    auto cit = cm->channels.find(channel_id);
    if (cit == cm->channels.end()) return fail(LZYB_E_INTERNAL, "Cannot get channel %s", channel_id);  // getChannelAndCheckAccess
    ChannelRec& ch = cit->second;
    if (role != LZYB_ROLE_PRODUCER && role != LZYB_ROLE_CONSUMER)
        return fail(LZYB_E_INVALID, "Role of peer not set or unsupported");
    const std::string idk = idempotency_key ? idempotency_key : cm->gen("idk-");
    const std::string url = peer_url ? peer_url : "";
    const std::string rh = req_hash({channel_id, peer_id, execution_id ? execution_id : "", url, role == LZYB_ROLE_PRODUCER ? "P" : "C"});

    if (role == LZYB_ROLE_CONSUMER) {
        // createConsumer: one critical section {findProducer, insert peer, create ACTIVE transfer}
        auto existing = ch.peers.find(peer_id);
        if (existing != ch.peers.end() && existing->second.idk == idk && existing->second.request_hash == rh) {
            // replay of the same request: answer from the stored transfer
            for (auto& kv : ch.transfers) {
                if (kv.second.to_id == peer_id && kv.second.idk == idk) {
                    auto pit = ch.peers.find(kv.second.from_id);
                    if (pit != ch.peers.end()) {
                        out->has_peer = 1;
                        fill_peer(pit->second, &out->peer);
                        scopy(out->transfer_id, kv.first);
                    }
                    return LZYB_OK;
                }
            }
            return LZYB_OK;
        }
        PeerRec* producer = find_producer(ch, *cm);
        int rc = peer_create(ch, peer_id, LZYB_ROLE_CONSUMER, LZYB_PEER_SLOT, url, LZYB_PRIO_PRIMARY, producer != nullptr, idk, rh, *cm);
        if (rc != LZYB_OK) return rc;
        if (producer) {
            std::string tid = transfer_create(ch, producer->id, peer_id, LZYB_TR_ACTIVE, idk, rh, *cm);
            out->has_peer = 1;
            fill_peer(*producer, &out->peer);
            scopy(out->transfer_id, tid);
        }
        return LZYB_OK;
    }

    // createProducer: insert PRIMARY producer, claim every not-connected consumer, one transfer each
    int rc = peer_create(ch, peer_id, LZYB_ROLE_PRODUCER, LZYB_PEER_SLOT, url, LZYB_PRIO_PRIMARY, false, idk, rh, *cm);
    if (rc != LZYB_OK) return rc;
    std::vector<PeerRec*> consumers;
    for (auto& kv : ch.peers) {  // listConnectedConsumersByRequest (replay)
        PeerRec& p = kv.second;
        if (p.role == LZYB_ROLE_CONSUMER && p.connected && p.conn_idk == idk) {
            if (p.conn_request_hash != rh) return fail(LZYB_E_INTERNAL, "Cannot save peer description in db: Idempotency key conflict");
            consumers.push_back(&p);
        }
    }
    if (consumers.empty()) {  // markConsumersAsConnected
        for (auto& kv : ch.peers) {
            PeerRec& p = kv.second;
            if (p.role == LZYB_ROLE_CONSUMER && !p.connected) {
                p.connected = true;
                p.conn_idk = idk;
                p.conn_request_hash = rh;
                consumers.push_back(&p);
            }
        }
    }
    std::sort(consumers.begin(), consumers.end(), [](PeerRec* a, PeerRec* b) { return a->order < b->order; });
    const PeerRec& prod = ch.peers.at(peer_id);
    struct PT {
        PeerRec* peer;
        std::string tid;
    };
    std::vector<PT> transfers;
    for (size_t i = 0; i < consumers.size(); ++i) {
        PeerRec* c = consumers[i];
        int state = c->kind == LZYB_PEER_STORAGE ? LZYB_TR_ACTIVE : LZYB_TR_PENDING;
        std::string tid = transfer_create(ch, peer_id, c->id, state, idk + "_" + std::to_string(i), rh, *cm);
        transfers.push_back({c, tid});
    }
    for (auto& pt : transfers) {  // the storage consumer (if any) goes back in the reply
        if (pt.peer->kind == LZYB_PEER_STORAGE) {
            out->has_peer = 1;
            fill_peer(*pt.peer, &out->peer);
            scopy(out->transfer_id, pt.tid);
            break;
        }
    }
    // "Start actions after response is sent"
    for (size_t i = 0; i < transfers.size(); ++i) {
        if (transfers[i].peer->kind != LZYB_PEER_STORAGE)
            queue_start_transfer(*cm, ch, transfers[i].tid, prod, *transfers[i].peer, idk + "_" + std::to_string(i));
    }
    return LZYB_OK;
}

int lzyb_cm_unbind(lzyb_handle_t h, const char* channel_id, const char* peer_id) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    if (!channel_id || !peer_id) return fail(LZYB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(cm->mu);
    auto cit = cm->channels.find(channel_id);
    if (cit == cm->channels.end()) return fail(LZYB_E_INTERNAL, "Cannot get channel %s", channel_id);
    ChannelRec& ch = cit->second;
    for (auto& kv : ch.transfers) {
        const TransferRec& t = kv.second;
        if ((t.from_id == peer_id || t.to_id == peer_id) && (t.state == LZYB_TR_ACTIVE || t.state == LZYB_TR_PENDING))
            return fail(LZYB_E_UNAVAILABLE, "Cannot unbind this slot now, there are some not started transfers");
    }
    peer_drop(ch, peer_id);
    return LZYB_OK;
}

int lzyb_cm_transfer_completed(lzyb_handle_t h, const char* idempotency_key, const char* channel_id, const char* transfer_id) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    if (!channel_id || !transfer_id) return fail(LZYB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(cm->mu);
    auto cit = cm->channels.find(channel_id);
    if (cit == cm->channels.end()) return fail(LZYB_E_INTERNAL, "Cannot get channel %s", channel_id);
    ChannelRec& ch = cit->second;
    const std::string idk = idempotency_key ? idempotency_key : cm->gen("idk-");
    auto tit = ch.transfers.find(transfer_id);
    if (tit == ch.transfers.end()) return fail(LZYB_E_NOT_FOUND, "Cannot find transfer in db");
    TransferRec& t = tit->second;
    if ((t.state != LZYB_TR_COMPLETED && t.state != LZYB_TR_ACTIVE) || (t.state == LZYB_TR_COMPLETED && t.state_change_idk != idk))
        return fail(LZYB_E_PRECONDITION, "Transfer is not active");
    t.state = LZYB_TR_COMPLETED;
    t.state_change_idk = idk;
    auto to = ch.peers.find(t.to_id);
    if (to != ch.peers.end() && to->second.kind == LZYB_PEER_STORAGE) {
        // data uploaded to storage: drop it as consumer, re-create it as BACKUP producer
        PeerRec old = to->second;
        peer_drop(ch, old.id);  // cascades: the completed transfer row goes with the peer
        int rc = peer_create(ch, old.id, LZYB_ROLE_PRODUCER, LZYB_PEER_STORAGE, old.url, LZYB_PRIO_BACKUP, false, idk,
                             req_hash({channel_id, transfer_id}), *cm);
        if (rc != LZYB_OK) return rc;
    }
    return LZYB_OK;
}

int lzyb_cm_transfer_failed(lzyb_handle_t h, const char* idempotency_key, const char* channel_id, const char* transfer_id,
                            const char* description, lzyb_bind_result_t* out) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    if (!channel_id || !transfer_id || !out) return fail(LZYB_E_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    std::lock_guard<std::mutex> lk(cm->mu);
    auto cit = cm->channels.find(channel_id);
    if (cit == cm->channels.end()) return fail(LZYB_E_INTERNAL, "Cannot get channel %s", channel_id);
    ChannelRec& ch = cit->second;
    const std::string idk = idempotency_key ? idempotency_key : cm->gen("idk-");
    const std::string rh = req_hash({channel_id, transfer_id, description ? description : ""});
    auto tit = ch.transfers.find(transfer_id);
    if (tit == ch.transfers.end()) return fail(LZYB_E_NOT_FOUND, "Cannot find transfer in db");
    TransferRec& t = tit->second;
    if ((t.state != LZYB_TR_FAILED && t.state != LZYB_TR_ACTIVE) || (t.state == LZYB_TR_FAILED && t.state_change_idk != idk))
        return fail(LZYB_E_PRECONDITION, "Transfer is not active");
    const std::string consumer_id = t.to_id, failed_from = t.from_id;
    if (t.state == LZYB_TR_ACTIVE) {
        t.state = LZYB_TR_FAILED;
        t.error = description ? description : "";
        t.state_change_idk = idk;
        auto to = ch.peers.find(t.to_id);
        if (to != ch.peers.end() && to->second.kind == LZYB_PEER_STORAGE) {
            destroy_channel_and_workflow(*cm, ch.id, "Uploading data to storage failed", idk);
            return fail(LZYB_E_INTERNAL, "Uploading data to storage failed");
        }
        auto from = ch.peers.find(failed_from);
        if (from != ch.peers.end()) from->second.priority -= 1;  // decrementPriority
    }
    PeerRec* np = find_producer(ch, *cm);
    if (!np) {
        destroy_channel_and_workflow(*cm, ch.id, "No more producers in channel, while there are consumers", idk);
        return fail(LZYB_E_INTERNAL, "No more producers in channel, while there are consumers");
    }
    // NB: the reference passes transfer.from().id() as the *to* of the new transfer
    // (SlotsService.java:271-272); the consumer is what it means, and what is used here
    std::string tid = transfer_create(ch, np->id, consumer_id, LZYB_TR_ACTIVE, idk, rh, *cm);
    out->has_peer = 1;
    fill_peer(*np, &out->peer);
    scopy(out->transfer_id, tid);
    return LZYB_OK;
}

int64_t lzyb_cm_channels_status(lzyb_handle_t h, const char* execution_id, const char* const* channel_ids, int n_ids,
                                char* json_out, uint64_t cap) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    std::lock_guard<std::mutex> lk(cm->mu);
    std::set<std::string> want;
    for (int i = 0; i < n_ids; ++i)
        if (channel_ids && channel_ids[i]) want.insert(channel_ids[i]);
    std::ostringstream os;
    os << "{\"channels\":[";
    bool first = true;
    for (auto& kv : cm->channels) {
        const ChannelRec& c = kv.second;
        if (execution_id && *execution_id && c.execution_id != execution_id) continue;
        if (!want.empty() && !want.count(c.id)) continue;
        if (!first) os << ',';
        first = false;
        os << "{\"channel_id\":";
        json_escape(os, c.id);
        os << ",\"execution_id\":";
        json_escape(os, c.execution_id);
        os << ",\"scheme\":{\"data_format\":";
        json_escape(os, c.data_format);
        os << ",\"scheme_format\":";
        json_escape(os, c.scheme_format);
        os << ",\"scheme_content\":";
        json_escape(os, c.scheme_content);
        os << "}";
        for (int role : {LZYB_ROLE_PRODUCER, LZYB_ROLE_CONSUMER}) {
            os << (role == LZYB_ROLE_PRODUCER ? ",\"producers\":[" : ",\"consumers\":[");
            std::vector<const PeerRec*> ps;
            for (auto& pk : c.peers)
                if (pk.second.role == role) ps.push_back(&pk.second);
            std::sort(ps.begin(), ps.end(), [](const PeerRec* a, const PeerRec* b) { return a->order < b->order; });
            for (size_t i = 0; i < ps.size(); ++i) {
                if (i) os << ',';
                os << "{\"peer_id\":";
                json_escape(os, ps[i]->id);
                os << ",\"kind\":" << (ps[i]->kind == LZYB_PEER_STORAGE ? "\"storage\"" : "\"slot\"") << ",\"url\":";
                json_escape(os, ps[i]->url);
                os << ",\"priority\":" << ps[i]->priority << ",\"connected\":" << (ps[i]->connected ? "true" : "false") << "}";
            }
            os << "]";
        }
        os << "}";
    }
    os << "]}";
    std::string s = os.str();
    if (json_out && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(json_out, s.data(), n);
        json_out[n] = 0;
    }
    return (int64_t)s.size() + 1;
}

int lzyb_cm_poll_action(lzyb_handle_t h, lzyb_cm_action_t* out, int timeout_ms) {
    auto cm = get(h);
    if (!cm || !out) return fail(LZYB_E_ARG, "unknown channel manager handle");
    std::unique_lock<std::mutex> lk(cm->mu);
    if (cm->actions.empty() && timeout_ms > 0)
        cm->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !cm->actions.empty(); });
    if (cm->actions.empty()) return 0;
    *out = cm->actions.front();
    cm->actions.pop_front();
    return 1;
}

int lzyb_cm_action_result(lzyb_handle_t h, const char* channel_id, const char* transfer_id, int ok, const char* reason) {
    auto cm = get(h);
    if (!cm || !channel_id || !transfer_id) return fail(LZYB_E_ARG, "unknown channel manager handle");
    std::lock_guard<std::mutex> lk(cm->mu);
    auto cit = cm->channels.find(channel_id);
    if (cit == cm->channels.end()) return fail(LZYB_E_NOT_FOUND, "channel %s is gone", channel_id);
    ChannelRec& ch = cit->second;
    if (!ok) {
        destroy_channel_and_workflow(*cm, ch.id, reason ? reason : "Cannot connect.", "");
        return LZYB_OK;
    }
    auto tit = ch.transfers.find(transfer_id);
    if (tit == ch.transfers.end() || tit->second.state != LZYB_TR_PENDING) {
        // ConcurrentModificationException("Unexpected state of transfer") → the action fails → abort
        destroy_channel_and_workflow(*cm, ch.id, "Unexpected state of transfer", "");
        return fail(LZYB_E_PRECONDITION, "Unexpected state of transfer");
    }
    tit->second.state = LZYB_TR_ACTIVE;
    return LZYB_OK;
}

int lzyb_cm_restore_actions(lzyb_handle_t h) {
    auto cm = get(h);
    if (!cm) return fail(LZYB_E_ARG, "unknown channel manager handle");
    std::lock_guard<std::mutex> lk(cm->mu);
    int n = 0;
    for (auto& ck : cm->channels) {
        ChannelRec& ch = ck.second;
        for (auto& tk : ch.transfers) {
            if (tk.second.state != LZYB_TR_PENDING) continue;
            auto f = ch.peers.find(tk.second.from_id), t = ch.peers.find(tk.second.to_id);
            if (f == ch.peers.end() || t == ch.peers.end()) continue;
            queue_start_transfer(*cm, ch, tk.first, f->second, t->second, tk.second.idk);
            ++n;
        }
    }
    return n;
}

int lzyb_cm_transfer_state(lzyb_handle_t h, const char* channel_id, const char* transfer_id, int* state_out) {
    auto cm = get(h);
    if (!cm || !channel_id || !transfer_id || !state_out) return fail(LZYB_E_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(cm->mu);
    auto cit = cm->channels.find(channel_id);
    if (cit == cm->channels.end()) return fail(LZYB_E_NOT_FOUND, "no channel %s", channel_id);
    auto tit = cit->second.transfers.find(transfer_id);
    if (tit == cit->second.transfers.end()) return fail(LZYB_E_NOT_FOUND, "Cannot find transfer in db");
    *state_out = tit->second.state;
    return LZYB_OK;
}

int lzyb_cm_peer_priority(lzyb_handle_t h, const char* channel_id, const char* peer_id, int* priority_out) {
    auto cm = get(h);
    if (!cm || !channel_id || !peer_id || !priority_out) return fail(LZYB_E_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(cm->mu);
    auto cit = cm->channels.find(channel_id);
    if (cit == cm->channels.end()) return fail(LZYB_E_NOT_FOUND, "no channel %s", channel_id);
    auto pit = cit->second.peers.find(peer_id);
    if (pit == cit->second.peers.end()) return fail(LZYB_E_NOT_FOUND, "no peer %s", peer_id);
    *priority_out = pit->second.priority;
    return LZYB_OK;
}

int lzyb_cm_is_aborted(lzyb_handle_t h, const char* execution_id, char* reason_out, uint64_t cap) {
    auto cm = get(h);
    if (!cm || !execution_id) return fail(LZYB_E_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(cm->mu);
    auto it = cm->aborted.find(execution_id);
    if (it == cm->aborted.end()) return 0;
    if (reason_out && cap) scopy(reason_out, cap, it->second);
    return 1;
}

}  // extern "C"
