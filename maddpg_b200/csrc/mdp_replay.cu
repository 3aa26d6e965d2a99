// Device-resident replay ring: vectorised insert and gather-based sample (sm_100a).
//
// Replaces maddpg/trainer/replay_buffer.py (reference): add :25-32, _encode_sample :34-44,
// sample_index :55-56.  The reference keeps one python list of tuples per agent; every agent
// inserts every step and every agent update gathers ALL agents' buffers at the same indices
// (maddpg/trainer/maddpg.py:173-178), so the device ring stores one JOINT row per transition:
//   [obs_0..obs_{n-1} | act_0..act_{n-1} | pad][next_obs_0.. | pad][rew_0.. | done_0.. | pad]
// A sampled index is then ONE contiguous, 16-byte aligned row whose first x_dim floats are
// exactly the centralized critic input -- the gather is a pure row copy (TMA bulk-copy engine
// in mode 1), and per-agent sample_index() results are column views of it.
// Both kernels are HBM bound: insert writes 4*(2D_i+K_i+2) bytes per agent-transition, gather
// reads and writes B*row bytes.
#include "mdp_common.cuh"

namespace mdp {

// One warp per transition row; lanes sweep the joint row so global stores are coalesced.
__global__ void k_replay_insert(mdp_ring_layout L, float* __restrict__ ring, long long capacity, long long cursor,
                                int E, int agent, const float* __restrict__ obs, int obs_stride,
                                const float* __restrict__ act, int act_stride, const float* __restrict__ rew,
                                int rew_stride, const float* __restrict__ next_obs, int next_obs_stride,
                                const uint8_t* __restrict__ done, int done_stride,
                                const unsigned long long* __restrict__ ctl) {
  if (ctl) cursor = (cursor + (long long)ctl[1]) % capacity;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long long e = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < E;
       e += (long long)gridDim.x * warps_per_block) {
    float* row = ring + ((cursor + e) % capacity) * (long long)L.row_stride;
    if (agent < 0) {
      const float* o = obs + e * obs_stride;
      const float* a = act + e * act_stride;
      const float* n = next_obs + e * next_obs_stride;
      for (int c = lane; c < L.obs_sum; c += 32) row[c] = o[c];
      for (int c = lane; c < L.act_sum; c += 32) row[L.obs_sum + c] = a[c];
      for (int c = lane; c < L.obs_sum; c += 32) row[L.nx_off + c] = n[c];
      for (int c = lane; c < L.n_agents; c += 32) {
        row[L.rw_off + c] = rew[e * rew_stride + c];
        row[L.dn_off + c] = done[e * done_stride + c] ? 1.0f : 0.0f;
      }
    } else {
      // per-agent experience(): obs/act/next_obs point at that agent's own (E, dim) arrays
      const int D = L.obs_dim[agent], K = L.act_dim[agent];
      const float* o = obs + e * obs_stride;
      const float* a = act + e * act_stride;
      const float* n = next_obs + e * next_obs_stride;
      for (int c = lane; c < D; c += 32) {
        row[L.obs_off[agent] + c] = o[c];
        row[L.nx_off + L.obs_off[agent] + c] = n[c];
      }
      for (int c = lane; c < K; c += 32) row[L.obs_sum + L.act_off[agent] + c] = a[c];
      if (lane == 0) {
        row[L.rw_off + agent] = rew[e * rew_stride];
        row[L.dn_off + agent] = done[e * done_stride] ? 1.0f : 0.0f;
      }
    }
  }
}

// mode 0: one warp per sampled row, 128-bit loads/stores (rows are 16-byte aligned by layout)
__global__ void k_replay_gather_vec(const float4* __restrict__ ring, long long capacity, int row_vec4,
                                    const long long* __restrict__ idx, int B, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int b = blockIdx.x * warps_per_block + (threadIdx.x >> 5); b < B; b += gridDim.x * warps_per_block) {
    long long r = idx[b];
    const float4* src = ring + r * row_vec4;
    float4* dst = out + (long long)b * row_vec4;
    int c = lane;
    // 4 independent 16-byte requests in flight per lane
    for (; c + 96 < row_vec4; c += 128) {
      float4 v0 = __ldg(src + c), v1 = __ldg(src + c + 32), v2 = __ldg(src + c + 64), v3 = __ldg(src + c + 96);
      dst[c] = v0; dst[c + 32] = v1; dst[c + 64] = v2; dst[c + 96] = v3;
    }
    for (; c < row_vec4; c += 32) dst[c] = __ldg(src + c);
  }
}

// mode 1: TMA bulk-copy engine.  One elected thread per CTA moves ROWS_PER_CTA whole rows
// global -> shared with cp.async.bulk (completion on an mbarrier), then shared -> global with
// cp.async.bulk.global.shared::cta (bulk-group completion).  No register staging at all.

template <int ROWS>
__global__ void k_replay_gather_bulk(const float* __restrict__ ring, long long capacity, int row_stride,
                                     const long long* __restrict__ idx, int B, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  float* buf = reinterpret_cast<float*>(smem_raw);
  const uint32_t row_bytes = (uint32_t)row_stride * 4u;
  const int b0 = blockIdx.x * ROWS;
  const int nrows = min(ROWS, B - b0);
  if (nrows <= 0) return;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                 "r"(row_bytes * (uint32_t)nrows)
                 : "memory");
    for (int r = 0; r < nrows; ++r) {
      const float* src = ring + idx[b0 + r] * (long long)row_stride;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(buf + (size_t)r * row_stride)),
          "l"(src), "r"(row_bytes), "r"(smem_u32(&bar))
          : "memory");
    }
    // wait for all rows (phase 0)
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0u)
          : "memory");
    }
    // rows b0..b0+nrows of `out` are contiguous: one bulk store
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (long long)b0 * row_stride),
                 "r"(smem_u32(buf)), "r"(row_bytes * (uint32_t)nrows)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// ctl = {philox_counter, ring_cursor, episode, ring_length}
__global__ void k_ctl_advance(unsigned long long* ctl, unsigned long long d_counter, long long d_rows, long long capacity,
                              unsigned long long d_episode) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ctl[0] += d_counter;
    ctl[1] = (unsigned long long)(((long long)ctl[1] + d_rows) % capacity);
    ctl[2] += d_episode;
    long long len = (long long)ctl[3] + d_rows;
    ctl[3] = (unsigned long long)(len > capacity ? capacity : len);
  }
}

__global__ void k_replay_make_index(long long* __restrict__ idx_out, int B, long long length, uint64_t seed,
                                    uint64_t counter, const unsigned long long* __restrict__ ctl) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (ctl) counter += ctl[0];
  if (length <= 0) length = (long long)ctl[3];
  uint4 r = Philox::gen(seed, (uint32_t)b, 0x1D3Au, (uint32_t)counter, (uint32_t)(counter >> 32));
  long long i = (long long)(Philox::u01d(r.x, r.y) * (double)length);
  i = i < length ? i : length - 1;
  idx_out[b] = i > 0 ? i : 0;  // an empty ring (length 0) yields row 0, never a negative row (callers gate on the ring length)
}

}  // namespace mdp

using namespace mdp;

extern "C" int mdp_ctl_advance(uint64_t* ctl, uint64_t d_counter, int64_t d_rows, int64_t capacity, uint64_t d_episode,
                               void* stream) {
  MDP_REQUIRE(ctl && capacity > 0 && d_rows >= 0, "mdp_ctl_advance: bad argument");
  k_ctl_advance<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long*)ctl, d_counter, d_rows, capacity, d_episode);
  return check_launch("k_ctl_advance");
}

extern "C" int mdp_replay_make_index(int64_t* idx_out, int32_t B, int64_t length, uint64_t seed, uint64_t counter,
                                     const uint64_t* ctl, void* stream) {
  MDP_REQUIRE(idx_out && B > 0 && (length > 0 || ctl), "mdp_replay_make_index: bad argument");
  k_replay_make_index<<<cdiv(B, 256), 256, 0, (cudaStream_t)stream>>>((long long*)idx_out, B, length, seed, counter,
                                                                    (const unsigned long long*)ctl);
  return check_launch("k_replay_make_index");
}

extern "C" int mdp_ring_make_layout(int32_t n_agents, const int32_t* obs_dim, const int32_t* act_dim,
                                    mdp_ring_layout* out) {
  MDP_REQUIRE(out && obs_dim && act_dim && n_agents > 0 && n_agents <= MDP_MAX_AGENTS, "mdp_ring_make_layout: bad argument");
  memset(out, 0, sizeof(*out));
  out->n_agents = n_agents;
  int od = 0, ad = 0;
  for (int i = 0; i < n_agents; ++i) {
    MDP_REQUIRE(obs_dim[i] > 0 && act_dim[i] > 0, "mdp_ring_make_layout: agent %d has empty obs/act", i);
    out->obs_dim[i] = obs_dim[i];
    out->act_dim[i] = act_dim[i];
    out->obs_off[i] = od;
    out->act_off[i] = ad;
    od += obs_dim[i];
    ad += act_dim[i];
  }
  out->obs_sum = od;
  out->act_sum = ad;
  out->x_dim = od + ad;
  out->nx_off = round_up(out->x_dim, 4);
  out->rw_off = out->nx_off + round_up(od, 4);
  out->dn_off = out->rw_off + n_agents;
  out->row_stride = round_up(out->dn_off + n_agents, 4);
  return MDP_OK;
}

extern "C" int mdp_replay_insert(const mdp_ring_layout* lay, float* ring, int64_t capacity, int64_t cursor, int32_t E,
                                 int32_t agent, const float* obs, int32_t obs_stride, const float* act,
                                 int32_t act_stride, const float* rew, int32_t rew_stride, const float* next_obs,
                                 int32_t next_obs_stride, const uint8_t* done, int32_t done_stride, void* stream) {
  return mdp::replay_insert_ctl(lay, ring, capacity, cursor, E, agent, obs, obs_stride, act, act_stride, rew, rew_stride,
                                next_obs, next_obs_stride, done, done_stride, nullptr, stream);
}

int mdp::replay_insert_ctl(const mdp_ring_layout* lay, float* ring, int64_t capacity, int64_t cursor, int32_t E,
                           int32_t agent, const float* obs, int32_t obs_stride, const float* act, int32_t act_stride,
                           const float* rew, int32_t rew_stride, const float* next_obs, int32_t next_obs_stride,
                           const uint8_t* done, int32_t done_stride, const unsigned long long* ctl, void* stream) {
  MDP_REQUIRE(lay && ring && obs && act && rew && next_obs && done, "mdp_replay_insert: null argument");
  MDP_REQUIRE(capacity > 0 && E > 0 && E <= capacity && cursor >= 0 && cursor < capacity,
              "mdp_replay_insert: bad sizes (capacity %lld, cursor %lld, E %d)", (long long)capacity, (long long)cursor, E);
  MDP_REQUIRE(agent >= -1 && agent < lay->n_agents, "mdp_replay_insert: bad agent %d", agent);
  const int wpb = 8;
  int grid = cdiv(E, wpb);
  if (grid > 148 * 16) grid = 148 * 16;
  k_replay_insert<<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(*lay, ring, capacity, cursor, E, agent, obs, obs_stride,
                                                              act, act_stride, rew, rew_stride, next_obs,
                                                              next_obs_stride, done, done_stride, ctl);
  return check_launch("k_replay_insert");
}

extern "C" int mdp_replay_gather(const float* ring, int64_t capacity, int32_t row_stride, const int64_t* idx, int32_t B,
                                 float* out, int32_t mode, void* stream) {
  MDP_REQUIRE(ring && idx && out && B > 0 && capacity > 0, "mdp_replay_gather: bad argument");
  MDP_REQUIRE(row_stride > 0 && row_stride % 4 == 0, "mdp_replay_gather: row_stride %d not a multiple of 4 floats", row_stride);
  MDP_REQUIRE((((uintptr_t)ring) & 15) == 0 && (((uintptr_t)out) & 15) == 0, "mdp_replay_gather: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    const int wpb = 8;
    int grid = cdiv(B, wpb);
    if (grid > 148 * 16) grid = 148 * 16;
    k_replay_gather_vec<<<grid, wpb * 32, 0, st>>>((const float4*)ring, capacity, row_stride / 4,
                                                   (const long long*)idx, B, (float4*)out);
    return check_launch("k_replay_gather_vec");
  }
  MDP_REQUIRE(mode == 1, "mdp_replay_gather: unknown mode %d", mode);
  // rows per CTA: fill up to ~64 KB of shared memory, at most 16 rows
  const size_t row_bytes = (size_t)row_stride * 4;
  MDP_REQUIRE(row_bytes <= 200 * 1024, "mdp_replay_gather: row of %zu bytes exceeds the bulk path's shared memory", row_bytes);
  int rows = (int)(65536 / row_bytes);
  if (rows < 1) rows = 1;
  if (rows > 16) rows = 16;
  // instantiate a few fixed row counts
  int R = rows >= 16 ? 16 : rows >= 8 ? 8 : rows >= 4 ? 4 : rows >= 2 ? 2 : 1;
  size_t smem = row_bytes * R;
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(B, R), 32, smem, st>>>(ring, capacity, row_stride, (const long long*)idx, B, out);
    return check_launch("k_replay_gather_bulk");
  };
  switch (R) {
    case 16: return launch(k_replay_gather_bulk<16>);
    case 8: return launch(k_replay_gather_bulk<8>);
    case 4: return launch(k_replay_gather_bulk<4>);
    case 2: return launch(k_replay_gather_bulk<2>);
    default: return launch(k_replay_gather_bulk<1>);
  }
}
