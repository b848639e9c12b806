// Thread-local last-error string shared by the host helpers and the C-ABI (pm_last_error()).
#pragma once
namespace pmh {
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
const char *last_error();
void clear_error();
}  // namespace pmh
