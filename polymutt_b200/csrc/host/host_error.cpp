#include "host_error.h"

#include <cstdarg>
#include <cstdio>

namespace pmh {
static thread_local char g_msg[1024] = "";
int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_msg, sizeof g_msg, fmt, ap);
  va_end(ap);
  return code;
}
const char *last_error() { return g_msg; }
void clear_error() { g_msg[0] = 0; }
}  // namespace pmh
