// api_common.cpp — error reporting shared by every entry point of the C ABI
// (reference analogue: PbrtError, src/core/error/error.rs:4-30; failures return non-zero + message).
#include "../../../include/pbrt_r3_b200.h"

#include <string>

static thread_local std::string g_last_error;

void pt_set_error(const std::string& s) { g_last_error = s; }

extern "C" {
const char* pt_last_error(void) { return g_last_error.c_str(); }
int pt_abi_version(void) { return PT_ABI_VERSION; }
}
