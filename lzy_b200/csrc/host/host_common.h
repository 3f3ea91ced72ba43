// host_common.h — shared by the CUDA side (capi.cu) and the host-only control logic.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstring>
#include <string>

namespace lzyb_host {

// thread-local message behind lzyb_last_error()
char* err_buf();
constexpr size_t ERR_CAP = 512;
int vfail(int code, const char* fmt, va_list ap);
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

inline void scopy(char* dst, size_t cap, const std::string& s) {
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(dst, s.data(), n);
    dst[n] = 0;
}
template <size_t N>
inline void scopy(char (&dst)[N], const std::string& s) {
    scopy(dst, N, s);
}

}  // namespace lzyb_host
