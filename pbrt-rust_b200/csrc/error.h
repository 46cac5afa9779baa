// Thread-local error reporting behind pbrt_b200_last_error().
#pragma once
#include <string>
namespace pbrt_b200 {
int fail(int code, const std::string& msg);
const char* last_error_cstr();
}  // namespace pbrt_b200
