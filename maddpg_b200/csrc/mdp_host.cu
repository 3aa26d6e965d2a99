// Host-buffer entry points: the reference loop body (experiments/train.py:112-120) for E lockstep env
// instances in ONE call whose inputs and outputs are HOST arrays.  Everything between the two copies runs
// in the kernels of mdp_train.cu / mdp_env.cu / mdp_replay.cu; nothing here computes on the CPU.
#include "mdp_env_dev.cuh"

using namespace mdp;

static inline int64_t up256(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int mdp_host_step_layout(const mdp_env* env, int32_t E, int64_t* offs4, int64_t* total_bytes) {
  MDP_REQUIRE(env && offs4 && total_bytes && E > 0, "mdp_host_step_layout: bad argument");
  mdp_env_dims d;
  int rc = mdp_env_get_dims(env, &d);
  if (rc) return rc;
  int64_t o = 0;
  offs4[0] = o; o = up256(o + 4ll * E * d.obs_stride);   // next observations (E, obs_stride) f32
  offs4[1] = o; o = up256(o + 4ll * E * d.n_agents);     // rewards (E, n_agents) f32
  offs4[2] = o; o = up256(o + 4ll * E * d.act_stride);   // sampled actions (E, act_stride) f32
  offs4[3] = o; o = up256(o + 1ll * E * d.n_agents);     // done (E, n_agents) u8
  *total_bytes = o;
  return MDP_OK;
}

// Copies between page-locked host memory and device memory issued as a KERNEL (the SMs read / write the host buffer
// through the unified address space).  A copy-engine transfer of ~1 MB costs ~10 us of fixed latency on top of its
// ~20 us on the wire (measured: H2D 0.9 MB + sync = 28.5 us, D2H 1.2 MB + sync = 33.9 us on a B200 / PCIe Gen5 box);
// a kernel starts within a launch latency and needs no engine hand-over between dependent operations.
__global__ void __launch_bounds__(256) k_copy16(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}

// up to four regions in one launch (a range's next observations, rewards, actions and done flags)
struct CopyRegions {
  uint4* dst[4];
  const uint4* src[4];
  unsigned n16[4];
};
__global__ void __launch_bounds__(256) k_copy16x4(CopyRegions R) {
  const unsigned stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    for (unsigned i = t; i < R.n16[r]; i += stride) R.dst[r][i] = R.src[r][i];
}

static int copy_kernel(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  MDP_REQUIRE(bytes % 16 == 0 && ((uintptr_t)dst & 15) == 0 && ((uintptr_t)src & 15) == 0, "host copy kernel: 16-byte alignment");
  const size_t n16 = bytes / 16;
  int grid = (int)((n16 + 255) / 256);
  if (grid > 148 * 4) grid = 148 * 4;
  k_copy16<<<grid, 256, 0, st>>>((uint4*)dst, (const uint4*)src, n16);
  return check_launch("k_copy16");
}

// is `p` (a host pointer) directly usable by kernels?  page-locked + mapped under the unified address space
static bool device_accessible(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}

static int ensure_pipeline(mdp_env* env) {
  if (env->pipeline_ready) return MDP_OK;
  for (int i = 0; i < mdp_env::kMaxChunks; ++i) {
    MDP_CUDA(cudaStreamCreateWithFlags(&env->chunk_stream[i], cudaStreamNonBlocking));
    MDP_CUDA(cudaEventCreateWithFlags(&env->chunk_done[i], cudaEventDisableTiming));
  }
  MDP_CUDA(cudaEventCreateWithFlags(&env->fork_ev, cudaEventDisableTiming));
  env->pipeline_ready = 1;
  return MDP_OK;
}

// The loop body for E env instances split into n_chunks ranges, each on its own stream: the H2D copy of range c+1 and the
// D2H copy of range c-1 run on the two copy engines while range c computes.  All streams fork from / join `stream` through
// events, so the call is one unit of work on `stream` (and is capturable into a CUDA graph as such).
static int host_step_impl(mdp_env* env, mdp_core* core, int32_t E, int32_t n_chunks, void* state, const float* h_obs_in,
                          float* d_obs_in, void* d_out, void* h_out, float* ring, int64_t ring_capacity, int32_t ring_row_stride,
                          int64_t ring_cursor, uint64_t seed, uint64_t counter, void* stream) {
  MDP_REQUIRE(env && core && state && h_obs_in && d_obs_in && d_out && h_out && E > 0, "mdp_host_step: bad argument");
  MDP_REQUIRE(n_chunks >= 1 && n_chunks <= mdp_env::kMaxChunks && E % n_chunks == 0,
              "mdp_host_step: %d env instances do not split into %d chunks", E, n_chunks);
  mdp_env_dims d;
  int rc = mdp_env_get_dims(env, &d);
  if (rc) return rc;
  int64_t off[4], total;
  rc = mdp_host_step_layout(env, E, off, &total);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* dout = static_cast<char*>(d_out);
  char* hout = static_cast<char*>(h_out);
  float* d_obs = reinterpret_cast<float*>(dout + off[0]);
  float* d_rew = reinterpret_cast<float*>(dout + off[1]);
  float* d_act = reinterpret_cast<float*>(dout + off[2]);
  uint8_t* d_done = reinterpret_cast<uint8_t*>(dout + off[3]);
  const bool sm_copies = env->host_copy_mode == 1 && device_accessible(h_obs_in) && device_accessible(h_out);
  if (n_chunks > 1) {
    rc = ensure_pipeline(env);
    if (rc) return rc;
    MDP_CUDA(cudaEventRecord(env->fork_ev, st));
  }
  const int n = E / n_chunks;
  // copy kernels move 16-byte words: a range's slices of the small arrays qualify when n * n_agents is a multiple of 16
  const bool tail_in_chunks = sm_copies && n_chunks > 1 && (n * d.n_agents) % 16 == 0 && (n * d.act_stride) % 4 == 0;
  for (int c = 0; c < n_chunks; ++c) {
    cudaStream_t cs = n_chunks > 1 ? env->chunk_stream[c] : st;
    const int e0 = c * n;
    const size_t obs_off = (size_t)e0 * d.obs_stride, obs_bytes = 4ull * n * d.obs_stride;
    if (n_chunks > 1) MDP_CUDA(cudaStreamWaitEvent(cs, env->fork_ev, 0));
    // obs_n (host) -> device: the argument of agent.action(obs), train.py:112
    if (sm_copies) {
      rc = copy_kernel(d_obs_in + obs_off, h_obs_in + obs_off, obs_bytes, cs);
      if (rc) return rc;
    } else {
      MDP_CUDA(cudaMemcpyAsync(d_obs_in + obs_off, h_obs_in + obs_off, obs_bytes, cudaMemcpyHostToDevice, cs));
    }
    // action_n = [agent.action(obs) ...]: grouped actor inference + Gumbel-softmax sampling
    rc = mdp::actor_act_range(core, 0, d.n_agents, 0, n, d_obs_in + obs_off, d.obs_stride, d_act + (size_t)e0 * d.act_stride,
                              d.act_stride, nullptr, seed, counter, nullptr, e0, cs);
    if (rc) return rc;
    // new_obs_n, rew_n, done_n = env.step(action_n) (train.py:114) + agent.experience(...) for every agent (train.py:119-120)
    rc = mdp::env_step_range(env, E, e0, n, state, d_act, d_obs, d_rew, d_done, ring ? d_obs_in : nullptr, ring, ring_capacity,
                             ring_row_stride, ring_cursor, cs);
    if (rc) return rc;
    if (n_chunks > 1) {
      // the range's next observations go home as soon as they exist; the small arrays follow in one copy after the join
      if (sm_copies && tail_in_chunks) {
        // everything the range produced goes home in one launch: next observations, rewards, actions, done flags
        CopyRegions R;
        const size_t o4[4] = {4 * obs_off, 4ull * e0 * d.n_agents, 4ull * e0 * d.act_stride, 1ull * e0 * d.n_agents};
        const size_t b4[4] = {obs_bytes, 4ull * n * d.n_agents, 4ull * n * d.act_stride, 1ull * n * d.n_agents};
        for (int r = 0; r < 4; ++r) {
          R.dst[r] = reinterpret_cast<uint4*>(hout + off[r] + o4[r]);
          R.src[r] = reinterpret_cast<const uint4*>(dout + off[r] + o4[r]);
          R.n16[r] = (unsigned)(b4[r] / 16);
        }
        int grid = (int)((obs_bytes / 16 + 255) / 256);
        if (grid > 148 * 4) grid = 148 * 4;
        k_copy16x4<<<grid, 256, 0, cs>>>(R);
        rc = check_launch("k_copy16x4");
        if (rc) return rc;
      } else if (sm_copies) {
        rc = copy_kernel(hout + off[0] + 4 * obs_off, dout + off[0] + 4 * obs_off, obs_bytes, cs);
        if (rc) return rc;
      } else {
        MDP_CUDA(cudaMemcpyAsync(hout + off[0] + 4 * obs_off, dout + off[0] + 4 * obs_off, obs_bytes, cudaMemcpyDeviceToHost, cs));
      }
      MDP_CUDA(cudaEventRecord(env->chunk_done[c], cs));
    }
  }
  if (n_chunks > 1) {
    for (int c = 0; c < n_chunks; ++c) MDP_CUDA(cudaStreamWaitEvent(st, env->chunk_done[c], 0));
    if (tail_in_chunks) return MDP_OK;
    if (sm_copies) return copy_kernel(hout + off[1], dout + off[1], (size_t)(total - off[1]), st);
    MDP_CUDA(cudaMemcpyAsync(hout + off[1], dout + off[1], (size_t)(total - off[1]), cudaMemcpyDeviceToHost, st));
  } else {
    // one packed device -> host copy of everything the loop reads back
    if (sm_copies) return copy_kernel(h_out, d_out, (size_t)total, st);
    MDP_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)total, cudaMemcpyDeviceToHost, st));
  }
  return MDP_OK;
}

extern "C" int mdp_host_copy_mode(mdp_env* env, int32_t mode) {
  MDP_REQUIRE(env && (mode == 0 || mode == 1), "mdp_host_copy_mode: mode must be 0 (copy engines) or 1 (copy kernels)");
  env->host_copy_mode = mode;
  return MDP_OK;
}

extern "C" int mdp_host_step(mdp_env* env, mdp_core* core, int32_t E, void* state, const float* h_obs_in, float* d_obs_in,
                             void* d_out, void* h_out, float* ring, int64_t ring_capacity, int32_t ring_row_stride,
                             int64_t ring_cursor, uint64_t seed, uint64_t counter, void* stream) {
  return host_step_impl(env, core, E, 1, state, h_obs_in, d_obs_in, d_out, h_out, ring, ring_capacity, ring_row_stride, ring_cursor,
                        seed, counter, stream);
}

extern "C" int mdp_host_step_pipelined(mdp_env* env, mdp_core* core, int32_t E, int32_t n_chunks, void* state,
                                       const float* h_obs_in, float* d_obs_in, void* d_out, void* h_out, float* ring,
                                       int64_t ring_capacity, int32_t ring_row_stride, int64_t ring_cursor, uint64_t seed,
                                       uint64_t counter, void* stream) {
  return host_step_impl(env, core, E, n_chunks, state, h_obs_in, d_obs_in, d_out, h_out, ring, ring_capacity, ring_row_stride,
                        ring_cursor, seed, counter, stream);
}
