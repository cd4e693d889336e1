// api.cu — error text, version and launch accounting of the C-ABI.
#include "common.cuh"

namespace nncf {
static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
std::atomic<int64_t> g_launches{0};
}  // namespace nncf

extern "C" const char* nncf_last_error(void) { return nncf::t_error.c_str(); }
extern "C" int nncf_version(void) { return 100; }
extern "C" int64_t nncf_launch_count(void) { return nncf::g_launches.load(); }
