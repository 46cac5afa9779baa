#include "error.h"
#include "../../include/pbrt_b200.h"
namespace pbrt_b200 {
static thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
const char* last_error_cstr() { return g_err.c_str(); }
}  // namespace pbrt_b200
extern "C" const char* pbrt_b200_last_error(void) { return pbrt_b200::last_error_cstr(); }
extern "C" int pbrt_b200_abi_version(void) { return PBRT_B200_ABI_VERSION; }
