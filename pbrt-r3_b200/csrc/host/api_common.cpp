// api_common.cpp — error reporting shared by every entry point of the C ABI
// (reference analogue: PbrtError, src/core/error/error.rs:4-30; failures return non-zero + message).
#include "../../../include/pbrt_r3_b200.h"

#include <string>

static thread_local std::string g_last_error;

void pt_set_error(const std::string& s) { g_last_error = s; }

// PbrtOptions singleton (core/options/pbrt_options.rs)
static pt_options g_options = {0, 0};
const pt_options& pt_current_options() { return g_options; }

extern "C" {
const char* pt_last_error(void) { return g_last_error.c_str(); }
int pt_abi_version(void) { return PT_ABI_VERSION; }
void pt_set_options(const pt_options* o) {
  if (o) g_options = *o;
  else g_options = pt_options{0, 0};
}
void pt_get_options(pt_options* out) { if (out) *out = g_options; }
}
