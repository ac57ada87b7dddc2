// host_util.hpp — error plumbing shared by the host-side translation units.
//
// The reference reports failures by throwing csrt::MyException
// (include/csrt/utils/misc.hpp:54-63); behind a C ABI that becomes
// "negative return code + message retrievable with b200pt_last_error()".
#pragma once
#include <string>

namespace b200pt {

// Records `msg` as the process-wide last error (for failures that happen
// before a handle exists) and returns `code` so callers can `return SetGlobalError(...)`.
int SetGlobalError(int code, const std::string &msg);
const char *GlobalError();

} // namespace b200pt
