// Library-wide state of libmaddpg_b200: error message, launch counter, version string.
#include "mdp_common.cuh"

namespace mdp {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace mdp

extern "C" const char* mdp_last_error(void) { return mdp::g_err; }
extern "C" const char* mdp_version(void) { return "maddpg_b200 0.1 (sm_100a)"; }
extern "C" int64_t mdp_launch_count(void) { return (int64_t)mdp::g_launches.load(); }
