// Library-wide state of libmaddpg_b200: error message, launch counter, version string.
#include "mdp_common.cuh"

namespace mdp {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace mdp

extern "C" const char* mdp_last_error(void) { return mdp::g_err; }
extern "C" const char* mdp_version(void) { return "maddpg_b200 0.1 (sm_100a)"; }
extern "C" int64_t mdp_launch_count(void) { return (int64_t)mdp::g_launches.load(); }

// Blocks until everything enqueued on `stream` has completed (cudaStreamSynchronize): the host side of a call with host
// result buffers (mdp_host_step) needs exactly this and nothing else from the runtime.
extern "C" int mdp_stream_synchronize(void* stream) {
  MDP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return MDP_OK;
}

// The uniform draws behind the kernels' Gumbel noise, computed on the HOST with the same Philox4x32-10 keying as
// philox_u (mdp_mlp.cuh): element (r, a) = u(seed, counter, tag = agent, row = row0 + r, col = a).  Parity tests feed these
// to the CPU oracle so that a free-running device rollout can be replayed on the CPU (TF / numpy streams cannot
// be reproduced, so the tests share THIS stream instead).
extern "C" int mdp_philox_uniform(uint64_t seed, uint64_t counter, uint32_t tag, int64_t row0, int32_t nrows, int32_t ncols,
                                  float* h_out) {
  MDP_REQUIRE(h_out && nrows >= 0 && ncols >= 0, "mdp_philox_uniform: bad argument");
  for (int r = 0; r < nrows; ++r)
    for (int a = 0; a < ncols; ++a) {
      const long long row = row0 + r;
      const uint4 v = mdp::Philox::gen(seed, (uint32_t)row, (uint32_t)(row >> 32) ^ (tag << 8) ^ (uint32_t)a, (uint32_t)counter,
                                       (uint32_t)(counter >> 32));
      h_out[(size_t)r * ncols + a] = mdp::Philox::u01(v.x);
    }
  return MDP_OK;
}
