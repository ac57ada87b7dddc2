#include "host_util.hpp"

#include <mutex>

namespace b200pt {

namespace {
std::mutex g_error_mutex;
std::string g_error;
} // namespace

int SetGlobalError(int code, const std::string &msg) {
    std::lock_guard<std::mutex> lock(g_error_mutex);
    g_error = msg;
    return code;
}

const char *GlobalError() {
    std::lock_guard<std::mutex> lock(g_error_mutex);
    return g_error.c_str();
}

} // namespace b200pt
