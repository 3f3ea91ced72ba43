#include "host_common.h"

#include <cstdio>

namespace lzyb_host {

char* err_buf() {
    static thread_local char buf[ERR_CAP] = "";
    return buf;
}

int vfail(int code, const char* fmt, va_list ap) {
    vsnprintf(err_buf(), ERR_CAP, fmt, ap);
    return code;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfail(code, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace lzyb_host
