/* Shim: map the absent pcreposix.h onto POSIX regex.h (base/Regex.h is not on the hot path). Ours. */
#ifndef PM_SHIM_PCREPOSIX_H
#define PM_SHIM_PCREPOSIX_H
#include <regex.h>
#endif
