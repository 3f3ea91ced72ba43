// local_rpc.h — what slots_runtime.cpp needs of the box-local transport (local_rpc.cpp)
#pragma once
#include <string>

#include "../../../include/lzy_b200.h"
#include "../../../include/lzy_b200_host.h"

namespace lzyb_host {

// does this peer URL name a unix socket of its worker ("...&uds=<path>")?
bool local_peer_socket(const char* peer_url, std::string* path);
// LzybPeerApi.Describe over that socket: the descriptor of producer slot `peer_id` of the worker behind it
int local_describe(const std::string& socket_path, const char* peer_id, lzyb_ipc_desc_t* out);

// a matcher connection made with lzyb_slots_connect_cm_local whose socket is served by THIS process: that matcher's handle
// (0 otherwise).  The in-process action pump uses it to find the slots instances that reach its matcher through the socket
// (one-process topologies: tests, the sanitizer workout).
lzyb_handle_t local_matcher_behind(lzyb_handle_t cookie);

}  // namespace lzyb_host
