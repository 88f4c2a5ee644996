// Error string, version and launch counter of libcodes_b200.so.
#include "cb_common.cuh"
#include <atomic>

namespace cb {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};  // process-wide (autograd runs backward on its own thread)

char* error_buffer() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
}  // namespace cb

extern "C" {
const char* cb_last_error(void) { return cb::error_buffer(); }
int32_t cb_version(void) { return 100; }
int64_t cb_launch_count(void) { return cb::g_launches.load(); }
void cb_reset_launch_count(void) { cb::g_launches.store(0); }
}
