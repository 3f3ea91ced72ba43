// Test-only: lets the host control logic (lzy_b200/csrc/host/*.cpp) link into a sanitizer-instrumented driver
// without the CUDA half of liblzy_b200.so.  The default backend table of slots_runtime.cpp names the lzyb_* data
// plane; the driver always installs the in-memory backend, so these are never called — they only have to exist.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/lzy_b200_host.h"
#include "../../lzy_b200/csrc/host/host_common.h"

namespace {
[[noreturn]] void unreachable(const char* what) {
    fprintf(stderr, "data-plane stub %s reached: the in-memory backend is not installed\n", what);
    abort();
}
}  // namespace

extern "C" {
const char* lzyb_last_error(void) { return lzyb_host::err_buf(); }
int lzyb_set_last_error(const char* m) {
    char* b = lzyb_host::err_buf();
    strncpy(b, m ? m : "", lzyb_host::ERR_CAP - 1);
    b[lzyb_host::ERR_CAP - 1] = 0;
    return LZYB_OK;
}
int lzyb_slot_create(const char*, int, uint64_t, lzyb_handle_t*) { unreachable("slot_create"); }
int lzyb_input_open(const char*, int, uint64_t, lzyb_handle_t*) { unreachable("input_open"); }
int lzyb_slot_commit(lzyb_handle_t, uint64_t*, void*) { unreachable("slot_commit"); }
int64_t lzyb_slot_read(lzyb_handle_t, uint64_t, void*, uint64_t) { unreachable("slot_read"); }
int lzyb_slot_info(lzyb_handle_t, lzyb_slot_info_t*) { unreachable("slot_info"); }
int lzyb_slot_find(const char*, lzyb_handle_t*) { unreachable("slot_find"); }
int lzyb_slot_close(lzyb_handle_t) { unreachable("slot_close"); }
int64_t lzyb_transfer_chunk(lzyb_handle_t, lzyb_handle_t, uint64_t, uint64_t, void*) { unreachable("transfer_chunk"); }
int lzyb_input_finish(lzyb_handle_t, lzyb_handle_t, void*) { unreachable("input_finish"); }
int64_t lzyb_input_write_host(lzyb_handle_t, uint64_t, const void*, uint64_t, void*) { unreachable("input_write_host"); }
int lzyb_input_finish_expect(lzyb_handle_t, const uint64_t*, void*) { unreachable("input_finish_expect"); }
int lzyb_input_to_output(lzyb_handle_t) { unreachable("input_to_output"); }
int lzyb_input_reset(lzyb_handle_t) { unreachable("input_reset"); }
int lzyb_slot_retain(lzyb_handle_t) { unreachable("slot_retain"); }
int lzyb_slot_set_meta(lzyb_handle_t, uint32_t, uint32_t, const uint64_t*) { unreachable("slot_set_meta"); }
int lzyb_slot_export(lzyb_handle_t, lzyb_ipc_desc_t*) { unreachable("slot_export"); }
int lzyb_peer_import(const lzyb_ipc_desc_t*, int, lzyb_handle_t*) { unreachable("peer_import"); }
int lzyb_slot_set_movable(lzyb_handle_t, int) { unreachable("slot_set_movable"); }
int lzyb_slot_meta(lzyb_handle_t, lzyb_slot_info_t*) { unreachable("slot_meta"); }
}
