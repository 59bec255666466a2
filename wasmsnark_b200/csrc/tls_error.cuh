// tls_error.cuh -- per-thread copy of the last failure text (host code), shared by ws_ctx and ws_multi.
#pragma once
#include <string>

namespace ws {

// The text of the calling THREAD's last failure and the handle it happened on.  A handle may be shared by several host
// threads (calls are serialised, see ws_ctx::mu): thread A's failed call must not have its message replaced -- or the
// string reallocated under A's eyes -- by thread B's failure between A's call and A's ws_last_error().
struct TlsError {
    const void* owner = nullptr;
    std::string text;
};
inline TlsError& tls_error() {
    static thread_local TlsError e;
    return e;
}
inline void tls_error_set(const void* owner, const std::string& text) {
    TlsError& e = tls_error();
    e.owner = owner;
    e.text = text;
}

inline void tls_error_forget(const void* owner) {      // the handle is being destroyed: its address may be reused
    TlsError& e = tls_error();
    if (e.owner == owner) e.owner = nullptr;
}

}  // namespace ws
