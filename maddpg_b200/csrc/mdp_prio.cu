// Device-resident prioritized replay: the reference's SumTree / PrioritizedReplayMemory
// (/root/reference/maddpg/trainer/prioritized_replay_buffer.py:19-201) as float64 array kernels.
//
// The tree is the reference's array (2^(k+1) - 1 float64 nodes, k = ceil(log2(capacity)), leaf of data slot d at index
// d + 2^k - 2) and every kernel reproduces the reference's float64 roundings in the reference's order, so that the sampled
// indices are bit-exact given the same uniforms -- including the structure's quirks, which oracle/prioritized.py lists and
// tests/golden/prioritized_ref.npz pins to the real class:
//   * slot 0 lives in the last INTERNAL node q = 2^k - 2 (its chain to the root is one level ahead of the true leaves');
//   * add() only marks slots dirty; the tree changes at the next get_leaf (k_sumtree_flush = SumTree.update_all);
//   * update_all adds the SUM of both children's deltas to a parent (one rounding); update() adds one delta per ancestor, in
//     batch order (k_sumtree_update_levels folds each node's deltas sequentially, in batch order, one thread per node);
//   * for odd k, update_all with slot 0 pending hands the root the wrong delta (k_sumtree_flush reproduces it).
#include <math.h>

#include "mdp_common.cuh"

namespace mdp {

struct Ival { long long lo, hi; };  // inclusive interval of node indices, empty when lo > hi
struct Dirty { Ival a, b; };        // the dirty nodes of one tree level: at most two intervals, a below b

__host__ __device__ inline bool ival_empty(const Ival& v) { return v.lo > v.hi; }
__host__ __device__ inline long long ival_count(const Ival& v) { return v.lo > v.hi ? 0 : v.hi - v.lo + 1; }
__host__ __device__ inline bool dirty_has(const Dirty& d, long long n) {
  return (n >= d.a.lo && n <= d.a.hi) || (n >= d.b.lo && n <= d.b.hi);
}
// true leaves (depth k) of the dirty data slots [start, start + count) (circular), slot 0 excluded (it is not a true leaf)
__host__ __device__ inline Dirty leaf_intervals(int k, long long cap, long long start, long long count) {
  const long long base = (1LL << k) - 2;
  Dirty d;
  d.a.lo = d.b.lo = 1; d.a.hi = d.b.hi = 0;
  const long long end = start + count;
  if (end <= cap) {
    const long long lo = start > 1 ? start : 1;
    if (lo <= end - 1) { d.a.lo = lo + base; d.a.hi = end - 1 + base; }
  } else {
    const long long lo = start > 1 ? start : 1;
    d.b.lo = lo + base; d.b.hi = cap - 1 + base;
    if (end - cap - 1 >= 1) { d.a.lo = 1 + base; d.a.hi = end - cap - 1 + base; }
    if (ival_empty(d.a)) { d.a = d.b; d.b.lo = 1; d.b.hi = 0; }
  }
  return d;
}
__host__ __device__ inline Dirty parent_intervals(Dirty d) {
  if (!ival_empty(d.a)) { d.a.lo = (d.a.lo - 1) >> 1; d.a.hi = (d.a.hi - 1) >> 1; }
  if (!ival_empty(d.b)) { d.b.lo = (d.b.lo - 1) >> 1; d.b.hi = (d.b.hi - 1) >> 1; }
  if (!ival_empty(d.a) && !ival_empty(d.b) && d.a.hi >= d.b.lo) { d.a.hi = d.b.hi; d.b.lo = 1; d.b.hi = 0; }
  return d;
}
__host__ __device__ inline bool slot0_dirty(long long cap, long long start, long long count) {
  return count > 0 && (start == 0 || start + count > cap);
}

// SumTree.update_all (:58-100) for the pending adds: iterations [t_begin, t_end) of the level-synchronous form (oracle/prioritized.py
// SumTreeOracle.update_all).  Iteration t touches the dirty nodes at depth k - t and slot 0's ancestor at depth k - 1 - t.
// One launch per wide level (any grid), or several levels in ONE CTA (gridDim.x == 1, barrier between levels).
// hdr[0] carries slot 0's delta between launches.
__global__ void __launch_bounds__(1024) k_sumtree_flush(double* __restrict__ tree, double* __restrict__ delta, int k, long long cap,
                                                         long long start, long long count, double value, int t_begin, int t_end,
                                                         double* __restrict__ hdr) {
  Dirty cur = leaf_intervals(k, cap, start, count), prev = cur;
  for (int t = 0; t < t_begin; ++t) { prev = cur; cur = parent_intervals(cur); }
  const bool s0 = slot0_dirty(cap, start, count);
  const long long q = (1LL << k) - 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (int t = t_begin; t < t_end; ++t) {
    const long long na = ival_count(cur.a), nb = ival_count(cur.b);
    for (long long i = tid; i < na + nb; i += nth) {
      const long long n = i < na ? cur.a.lo + i : cur.b.lo + (i - na);
      if (t == 0) {
        const double d = __dsub_rn(value, tree[n]);
        tree[n] = value;
        delta[n] = d;
      } else {
        const long long l = 2 * n + 1, r = l + 1;
        const bool hl = dirty_has(prev, l), hr = dirty_has(prev, r);
        const double dl = hl ? delta[l] : 0.0, dr = hr ? delta[r] : 0.0;
        double s = __dadd_rn(dl, dr);
        // The reference's parallel (node, delta) lists lose their alignment once slot 0's chain has reached the root (:95-99):
        // for odd k the root then receives slot 0's delta again plus node 1's, and the right subtree's delta is dropped
        // (oracle/prioritized.py, SumTreeOracle.update_all).  Reproduced: "results identical" includes the tree array.
        if (n == 0 && s0 && (k & 1)) s = (hl && hr) ? __dadd_rn(hdr[0], dl) : hdr[0];
        tree[n] = __dadd_rn(tree[n], s);
        delta[n] = s;
      }
    }
    if (s0 && tid == 0) {
      if (t == 0) {
        const double c0 = __dsub_rn(value, tree[q]);
        tree[q] = value;
        hdr[0] = c0;
      } else if (t <= k - 1) {
        const long long m = (1LL << (k - t)) - 2;
        tree[m] = __dadd_rn(tree[m], hdr[0]);
      }
    }
    if (t + 1 < t_end) __syncthreads();  // multi-level launches are single-CTA (host contract)
    prev = cur;
    cur = parent_intervals(cur);
  }
}

// PrioritizedReplayMemory.sample, the reads that precede the first get_leaf (:175-182): hdr[1] = total_p, hdr[2] = min over the
// LAST `cap` entries of the tree array (bit pattern: non-negative doubles order like their unsigned bits)
__global__ void k_sumtree_prep(const double* __restrict__ tree, double* __restrict__ hdr, int* __restrict__ flag) {
  atomicAnd(flag, ~1);
  hdr[1] = tree[0];
  reinterpret_cast<unsigned long long*>(hdr)[2] = 0x7FF0000000000000ull;  // +inf
}
__global__ void __launch_bounds__(256) k_sumtree_min(const double* __restrict__ tree, long long size, long long cap,
                                                      double* __restrict__ hdr) {
  const unsigned long long* bits = reinterpret_cast<const unsigned long long*>(tree) + (size - cap);
  unsigned long long m = ~0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long b = bits[i] & 0x7FFFFFFFFFFFFFFFull;  // -0.0 == 0.0
    m = b < m ? b : m;
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o);
    m = x < m ? x : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMin(reinterpret_cast<unsigned long long*>(hdr) + 2, m);
}

// the n stratified descents of sample() (:183-191): v = a + (b - a) * u like numpy's legacy uniform(a, b), get_leaf (:111-142),
// prob = p / total_p (after the flush), weight = (prob / min_prob) ^ -beta.  flag |= 1 where the reference raises IndexError
// (data index >= capacity: a descent through slot 0's node).
__global__ void __launch_bounds__(256) k_sumtree_sample(const double* __restrict__ tree, int k, long long cap, long long size, int n,
                                                         const double* __restrict__ u, const double* __restrict__ hdr, double beta,
                                                         long long* __restrict__ leaf_out, long long* __restrict__ data_out,
                                                         double* __restrict__ isw_out, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double total0 = hdr[1], min_prob = __ddiv_rn(hdr[2], total0);
  const double seg = __ddiv_rn(total0, (double)n);
  const double a = __dmul_rn(seg, (double)i), b = __dmul_rn(seg, (double)(i + 1));
  double v = __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), u[i]));
  long long parent = 0;
  while (true) {
    const long long cl = 2 * parent + 1;
    if (cl >= size) break;
    const double tl = tree[cl];
    if (v <= tl) {
      parent = cl;
    } else {
      v = __dsub_rn(v, tl);
      parent = cl + 1;
    }
  }
  const double p = tree[parent];
  const long long data = parent - (1LL << k) + 2;
  leaf_out[i] = parent;
  data_out[i] = data;
  isw_out[i] = pow(__ddiv_rn(__ddiv_rn(p, tree[0]), min_prob), -beta);
  if (data >= cap) atomicOr(flag, 1);
}

// ---------------------------------------------------------------------------------------------
// batch_update (:196-201): SumTree.update(ti, p) for every element in batch order.
// ---------------------------------------------------------------------------------------------
constexpr int PRIO_MAX_B = 4096;

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* keys, int P) {
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int idx = threadIdx.x; idx < P; idx += blockDim.x) {
        const int ixj = idx ^ j;
        if (ixj > idx) {
          const unsigned long long x = keys[idx], y = keys[ixj];
          if ((x > y) == ((idx & k2) == 0)) { keys[idx] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// Stage 1 (one CTA): priorities (injected, or (min(|err| + eps, upper)) ^ alpha), per-element change = p - (value the leaf holds
// when its turn comes: the previous duplicate's p, else the tree's), leaves set to the last duplicate's p.
// flag |= 2: some index is not a true leaf (or out of range) -> nothing is written here and k_sumtree_update_serial does it all.
__global__ void __launch_bounds__(1024) k_sumtree_update_leaves(double* __restrict__ tree, int k, long long size,
                                                                 const long long* __restrict__ tidx, int B,
                                                                 const double* __restrict__ abs_err, const double* __restrict__ prio,
                                                                 double eps, double upper, double alpha, double* __restrict__ pval,
                                                                 double* __restrict__ change, int* __restrict__ flag) {
  extern __shared__ unsigned long long s_keys[];
  double* s_p = reinterpret_cast<double*>(s_keys + PRIO_MAX_B);
  __shared__ int s_bad;
  if (threadIdx.x == 0) {
    s_bad = 0;
    atomicAnd(flag, ~2);
  }
  int P = 1;
  while (P < B) P <<= 1;
  __syncthreads();
  const long long first_leaf = (1LL << k) - 1;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    if (i < B) {
      const long long ti = tidx[i];
      const double p = prio ? prio[i] : pow(fmin(__dadd_rn(abs_err[i], eps), upper), alpha);
      s_p[i] = p;
      pval[i] = p;
      if (ti < first_leaf || ti >= size) s_bad = 1;
      s_keys[i] = ((unsigned long long)ti << 16) | (unsigned long long)i;
    } else {
      s_keys[i] = ~0ull;
    }
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0) atomicOr(flag, 2);
    return;
  }
  bitonic_sort_u64(s_keys, P);
  for (int s = threadIdx.x; s < B; s += blockDim.x) {  // reads of the old leaf values
    const unsigned long long key = s_keys[s];
    const int i = (int)(key & 0xffff);
    const long long ti = (long long)(key >> 16);
    const bool dup = s > 0 && (s_keys[s - 1] >> 16) == (key >> 16);
    const double prev = dup ? s_p[s_keys[s - 1] & 0xffff] : tree[ti];
    change[i] = __dsub_rn(s_p[i], prev);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < B; s += blockDim.x) {  // then the writes
    const unsigned long long key = s_keys[s];
    if (s == B - 1 || (s_keys[s + 1] >> 16) != (key >> 16)) tree[key >> 16] = s_p[key & 0xffff];
  }
}

// Stage 2: CTA d owns tree depth d in [0, k): a node's deltas are folded in batch order by one thread (the reference's
// `tree[idx] += change` sequence), distinct nodes in parallel.
//   * depth <= 5 (at most 32 nodes): one warp, lane = node; every lane walks the batch in order and adds the elements that
//     belong to its node (a skipped element leaves the chain untouched, so the roundings are the reference's) -- no sort, the
//     element loads are lane-uniform broadcasts and run ahead of the dependent adds;
//   * deeper levels: the elements are grouped per node by a bitonic sort on (node, batch position), segments are short.
__global__ void __launch_bounds__(1024) k_sumtree_update_levels(double* __restrict__ tree, int k, const long long* __restrict__ tidx,
                                                                 int B, const double* __restrict__ change,
                                                                 const int* __restrict__ flag) {
  extern __shared__ unsigned long long s_keys[];
  double* s_c = reinterpret_cast<double*>(s_keys + PRIO_MAX_B);
  if (*flag & 2) return;
  const int d = blockIdx.x;
  if (d <= 5) {
    unsigned* s_node = reinterpret_cast<unsigned*>(s_keys);  // node of element i, relative to the first node of the level
    const long long first = (1LL << d) - 1;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      s_node[i] = (unsigned)((((tidx[i] + 1) >> (k - d)) - 1) - first);
      s_c[i] = change[i];
    }
    __syncthreads();
    if (threadIdx.x < (1u << d)) {
      const unsigned me = threadIdx.x;
      double acc = tree[first + me];
      bool any = false;
      int i = 0;
      for (; i + 8 <= B; i += 8) {
        unsigned nd[8];
        double c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { nd[j] = s_node[i + j]; c[j] = s_c[i + j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (nd[j] == me) { acc = __dadd_rn(acc, c[j]); any = true; }
      }
      for (; i < B; ++i)
        if (s_node[i] == me) { acc = __dadd_rn(acc, s_c[i]); any = true; }
      if (any) tree[first + me] = acc;
    }
    return;
  }
  int P = 1;
  while (P < B) P <<= 1;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    if (i < B) {
      const unsigned long long node = (unsigned long long)(((tidx[i] + 1) >> (k - d)) - 1);
      s_keys[i] = (node << 16) | (unsigned long long)i;
      s_c[i] = change[i];
    } else {
      s_keys[i] = ~0ull;
    }
  }
  __syncthreads();
  bitonic_sort_u64(s_keys, P);
  for (int s = threadIdx.x; s < B; s += blockDim.x) {
    const unsigned long long node = s_keys[s] >> 16;
    if (s > 0 && (s_keys[s - 1] >> 16) == node) continue;  // not the head of its segment
    double acc = tree[node];
    int e = s;
    do {
      acc = __dadd_rn(acc, s_c[s_keys[e] & 0xffff]);
      ++e;
    } while (e < B && (s_keys[e] >> 16) == node);
    tree[node] = acc;
  }
}

// The reference loop itself on one thread: batches that name internal nodes (slot 0's node among them), where an element's
// change depends on earlier elements' propagation.  Negative indices wrap like python's; indices >= size are skipped.
__global__ void k_sumtree_update_serial(double* __restrict__ tree, long long size, const long long* __restrict__ tidx, int B,
                                        const double* __restrict__ pval, const int* __restrict__ flag) {
  if (!(*flag & 2)) return;
  for (int i = 0; i < B; ++i) {
    long long ti = tidx[i];
    if (ti < 0) ti += size;
    if (ti < 0 || ti >= size) continue;
    const double change = __dsub_rn(pval[i], tree[ti]);
    tree[ti] = pval[i];
    while (ti != 0) {
      ti = (ti - 1) >> 1;
      tree[ti] = __dadd_rn(tree[ti], change);
    }
  }
}

static int tree_k(int64_t capacity) {
  int k = 0;
  while ((1LL << k) < capacity) ++k;
  return k;
}

}  // namespace mdp

using namespace mdp;

extern "C" int mdp_sumtree_layout(int64_t capacity, int64_t* tree_size, int32_t* k_out, int64_t* scratch_doubles) {
  MDP_REQUIRE(capacity >= 3, "mdp_sumtree_layout: capacity %lld < 3", (long long)capacity);
  const int k = tree_k(capacity);
  MDP_REQUIRE(k <= 30, "mdp_sumtree_layout: capacity too large");
  if (tree_size) *tree_size = (2LL << k) - 1;
  if (k_out) *k_out = k;
  // hdr (8 doubles: slot-0 delta, total before the flush, min bits, flag word, ...) + per-node deltas + per-element p / change
  if (scratch_doubles) *scratch_doubles = 8 + ((2LL << k) - 1) + 2 * PRIO_MAX_B;
  return MDP_OK;
}

extern "C" int mdp_sumtree_flush(double* tree, int64_t capacity, int64_t start, int64_t count, double value, double* scratch,
                                 void* stream) {
  MDP_REQUIRE(tree && scratch && capacity >= 3, "mdp_sumtree_flush: bad argument");
  MDP_REQUIRE(start >= 0 && start < capacity && count >= 0 && count <= capacity, "mdp_sumtree_flush: bad dirty range");
  if (count == 0) return MDP_OK;
  if (count == capacity) start = 0;
  const int k = tree_k(capacity);
  cudaStream_t st = (cudaStream_t)stream;
  double* hdr = scratch;
  double* delta = scratch + 8;
  Dirty cur = leaf_intervals(k, capacity, start, count);
  int t = 0;
  for (; t <= k; ++t) {  // wide levels: one launch each
    const long long nodes = ival_count(cur.a) + ival_count(cur.b);
    if (nodes <= 4096) break;
    const long long grid = (nodes + 255) / 256;
    k_sumtree_flush<<<(unsigned)(grid < 1184 ? grid : 1184), 256, 0, st>>>(tree, delta, k, capacity, start, count, value, t, t + 1, hdr);
    int rc = check_launch("k_sumtree_flush");
    if (rc) return rc;
    cur = parent_intervals(cur);
  }
  if (t <= k) {  // the narrow top of the tree: all remaining levels in one CTA
    k_sumtree_flush<<<1, 1024, 0, st>>>(tree, delta, k, capacity, start, count, value, t, k + 1, hdr);
    return check_launch("k_sumtree_flush");
  }
  return MDP_OK;
}

extern "C" int mdp_sumtree_sample(double* tree, int64_t capacity, int64_t dirty_start, int64_t dirty_count, double dirty_value,
                                  int32_t n, const double* uniforms, double beta, int64_t* tree_idx_out, int64_t* data_idx_out,
                                  double* weights_out, int32_t* flag, double* scratch, void* stream) {
  MDP_REQUIRE(tree && uniforms && tree_idx_out && data_idx_out && weights_out && flag && scratch && n > 0 && capacity >= 3,
              "mdp_sumtree_sample: bad argument");
  const int k = tree_k(capacity);
  const long long size = (2LL << k) - 1;
  cudaStream_t st = (cudaStream_t)stream;
  k_sumtree_prep<<<1, 1, 0, st>>>(tree, scratch, flag);
  int rc = check_launch("k_sumtree_prep");
  if (rc) return rc;
  const long long grid = (capacity + 256 * 8 - 1) / (256 * 8);
  k_sumtree_min<<<(unsigned)(grid < 1184 ? grid : 1184), 256, 0, st>>>(tree, size, capacity, scratch);
  rc = check_launch("k_sumtree_min");
  if (rc) return rc;
  rc = mdp_sumtree_flush(tree, capacity, dirty_start, dirty_count, dirty_value, scratch, stream);
  if (rc) return rc;
  k_sumtree_sample<<<cdiv(n, 256), 256, 0, st>>>(tree, k, capacity, size, n, uniforms, scratch, beta, (long long*)tree_idx_out,
                                                 (long long*)data_idx_out, weights_out, flag);
  return check_launch("k_sumtree_sample");
}

extern "C" int mdp_sumtree_update(double* tree, int64_t capacity, const int64_t* tree_idx, int32_t B, const double* abs_errors,
                                  const double* priorities, double epsilon, double abs_err_upper, double alpha, int32_t* flag,
                                  double* scratch, void* stream) {
  MDP_REQUIRE(tree && tree_idx && flag && scratch && (abs_errors || priorities) && capacity >= 3, "mdp_sumtree_update: bad argument");
  MDP_REQUIRE(B > 0 && B <= PRIO_MAX_B, "mdp_sumtree_update: batch %d outside [1, %d]", B, PRIO_MAX_B);
  const int k = tree_k(capacity);
  const long long size = (2LL << k) - 1;
  cudaStream_t st = (cudaStream_t)stream;
  double* pval = scratch + 8 + size;
  double* change = pval + PRIO_MAX_B;
  const size_t smem = PRIO_MAX_B * (sizeof(unsigned long long) + sizeof(double));
  static bool attr_set = false;
  if (!attr_set) {
    MDP_CUDA(cudaFuncSetAttribute(k_sumtree_update_leaves, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MDP_CUDA(cudaFuncSetAttribute(k_sumtree_update_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k_sumtree_update_leaves<<<1, 1024, smem, st>>>(tree, k, size, (const long long*)tree_idx, B, abs_errors, priorities, epsilon,
                                                 abs_err_upper, alpha, pval, change, flag);
  int rc = check_launch("k_sumtree_update_leaves");
  if (rc) return rc;
  k_sumtree_update_levels<<<k, 1024, smem, st>>>(tree, k, (const long long*)tree_idx, B, change, flag);
  rc = check_launch("k_sumtree_update_levels");
  if (rc) return rc;
  k_sumtree_update_serial<<<1, 1, 0, st>>>(tree, size, (const long long*)tree_idx, B, pval, flag);
  return check_launch("k_sumtree_update_serial");
}
