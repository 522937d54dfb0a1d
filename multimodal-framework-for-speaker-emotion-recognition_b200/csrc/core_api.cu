// Shared host-side helpers of liblsthm_b200.so: ABI version and the thread-local error string every entry point reports through.
#include <string>

#include "../../include/lsthm_b200.h"
#include "common.cuh"

namespace lsthm {

thread_local std::string g_err;
static int fail(const std::string &m) {
    g_err = m;
    return 1;
}
int set_error(const char *what, cudaError_t e) { return fail(std::string(what) + ": " + cudaGetErrorString(e)); }
int fail_msg(const char *msg) { return fail(msg); }

}  // namespace lsthm

extern "C" {

int lsthm_abi_version(void) { return LSTHM_ABI_VERSION; }
const char *lsthm_last_error(void) { return lsthm::g_err.c_str(); }

}  // extern "C"
