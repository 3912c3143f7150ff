// kernels.cuh — the sm_100a kernels of the gradient-reduction path.
//
//   K1  allreduce_inplace_kernel   fused bucket allreduce (two-shot P2P / NVLS) + pre-divide + stats
//   K1b allreduce_oneshot_kernel   out-of-place one-shot allreduce for small tensors
//   K2  pull_kernel                allgather / broadcast by peer loads
//   K4  pack_kernel / pack_bulk_kernel   multi-tensor bucket pack/unpack (LDG/STG and cp.async.bulk)
//   K5  multi_stats_kernel / multi_scale_kernel   multi-tensor norm statistics / scaling
//   K6  finalize_norm_kernel / scale_flat_kernel / clamp_flat_kernel
//   K7  pad_copy_kernel
//   K8  exchange_small_kernel
//
// All of these are bandwidth-bound (NVLink or HBM); none is a contraction, so no tensor cores.
#pragma once
#include "device_utils.cuh"

namespace b200dp {

constexpr int kThreads = 512;        // comm kernels: 1 CTA / SM, 16 warps of 16-byte accesses
constexpr int kStreamThreads = 256;  // local HBM streaming kernels
constexpr int kMTMax = 640;          // tensors per multi-tensor launch (table lives in kernel params)
constexpr int kMTChunk = 8192;       // elements per chunk in multi_stats / multi_scale
constexpr int kPackChunk = 32768;    // bytes per chunk in the LDG/STG pack kernel

// ------------------------------------------------------------------------------------------------
// statistics hand-off: every CTA leaves its partial, the last CTA to arrive folds them in a fixed
// tree and publishes one Record to slot [slot][rank] of every peer's control segment.
// ------------------------------------------------------------------------------------------------
struct StatsSink {
  Record* partial;    // local scratch [gridDim.x]
  uint32_t* ticket;   // local arrival counter
  Peers control;      // control-segment bases (records live at kRecordOff)
  int slot;           // < 0: no stats
  int rank, world;    // world == 1 -> publish locally only
};

__device__ __forceinline__ void publish_stats(const StatsSink& sink, const Stats& mine) {
  __shared__ uint32_t s_last;
  double bs;
  float bm;
  uint32_t bnf;
  block_reduce_stats(mine, &bs, &bm, &bnf);
  if (threadIdx.x == 0) {
    Record r;
    r.sumsq = bs;
    r.maxabs = bm;
    r.nonfinite = bnf;
    sink.partial[blockIdx.x] = r;
    __threadfence();
    uint32_t t = atomicAdd(sink.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    double v = 0.0;
    float m = 0.f;
    uint32_t nf = 0;
    for (int i = threadIdx.x; i < int(gridDim.x); i += 32) {  // fixed assignment -> deterministic
      const Record* p = sink.partial + i;
      v += __ldcg(&p->sumsq);
      m = fmaxf(m, __ldcg(&p->maxabs));
      nf |= __ldcg(&p->nonfinite);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v += __shfl_xor_sync(0xffffffffu, v, o);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      nf |= __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if (int(threadIdx.x) < sink.world) {
      Record* dst = reinterpret_cast<Record*>(reinterpret_cast<char*>(sink.control.p[threadIdx.x]) + kRecordOff) +
                    (size_t(sink.slot) * B200DP_MAX_WORLD + sink.rank);
      Record r;
      r.sumsq = v;
      r.maxabs = m;
      r.nonfinite = nf;
      *dst = r;
    }
    if (threadIdx.x == 0) *sink.ticket = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K1: in-place fused allreduce of one gradient bucket.
//   x_i <- sum_r (x_i^(r) * prescale), r ascending, fp32 accumulate, no FMA contraction, so the
//   result equals the oracle's sequential definition bit for bit (TWOSHOT).  Rank r owns the r-th
//   1/N of the 16-byte vectors: it loads that slice from all N copies, reduces, and stores the
//   result into all N copies.  Nobody else touches those addresses, so in place is race-free.
//   NVLS: the owner issues multimem.ld_reduce (switch adds the N copies) and multimem.st.
// ------------------------------------------------------------------------------------------------
struct AllreduceArgs {
  BarrierCtx bar;
  Peers bufs;        // segment bases on every rank
  char* mc;          // multicast VA of the segment (NVLS) or nullptr
  size_t byte_off;   // bucket start inside the segment (16-byte aligned)
  size_t nvec;       // 16-byte vectors in the bucket
  float prescale;
  int flags;         // B200DP_F_*
  StatsSink sink;
};

// Loads one 16-byte vector of the bucket from all W ranks (issued back to back: W requests in flight).
template <int W>
__device__ __forceinline__ void load_all_ranks(const AllreduceArgs& a, size_t byte, uint4* v) {
#pragma unroll
  for (int r = 0; r < W; ++r) v[r] = ld_stream16(reinterpret_cast<const char*>(a.bufs.p[r]) + byte);
}

// acc = fl(x_0 * s); acc = fl(acc + fl(x_r * s)), r ascending; returns the packed result, feeds the stats.
template <typename T, int W>
__device__ __forceinline__ uint4 reduce_loaded(const uint4* v, float prescale, Stats& st, bool want_stats) {
  constexpr int N = Vec16<T>::N;
  float acc[N], x[N];
  Vec16<T>::unpack(v[0], acc);
#pragma unroll
  for (int k = 0; k < N; ++k) acc[k] = __fmul_rn(acc[k], prescale);
#pragma unroll
  for (int r = 1; r < W; ++r) {
    Vec16<T>::unpack(v[r], x);
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = __fadd_rn(acc[k], __fmul_rn(x[k], prescale));
  }
  uint4 out = Vec16<T>::pack(acc);
  if (want_stats) {
    Vec16<T>::unpack(out, x);  // stats of the values as stored (after rounding to T)
#pragma unroll
    for (int k = 0; k < N; ++k) st.add(x[k]);
  }
  return out;
}

// any world size: one rank at a time (less memory-level parallelism, but no W-sized register array)
template <typename T>
__device__ __forceinline__ uint4 reduce_vec_generic(const AllreduceArgs& a, size_t byte, float prescale, Stats& st,
                                                    bool want_stats) {
  constexpr int N = Vec16<T>::N;
  float acc[N], x[N];
  uint4 v0 = ld_stream16(reinterpret_cast<const char*>(a.bufs.p[0]) + byte);
  Vec16<T>::unpack(v0, acc);
#pragma unroll
  for (int k = 0; k < N; ++k) acc[k] = __fmul_rn(acc[k], prescale);
  for (int r = 1; r < a.bar.world; ++r) {
    uint4 v = ld_stream16(reinterpret_cast<const char*>(a.bufs.p[r]) + byte);
    Vec16<T>::unpack(v, x);
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = __fadd_rn(acc[k], __fmul_rn(x[k], prescale));
  }
  uint4 out = Vec16<T>::pack(acc);
  if (want_stats) {
    Vec16<T>::unpack(out, x);
#pragma unroll
    for (int k = 0; k < N; ++k) st.add(x[k]);
  }
  return out;
}

// vectors a thread keeps in flight per loop trip: W x kUnroll 16-byte loads (two-shot), kUnroll (NVLS)
template <int W>
struct UnrollFor {
  static constexpr int value = W >= 5 ? 2 : 4;  // keeps W x U x 4 data registers <= 64
};

// W > 0: two-shot specialised for W ranks; W == 0: two-shot, any world size; NVLS: W is the number of
// switch-side reductions a thread keeps in flight (4 or 8).
template <typename T, int W, bool NVLS>
__global__ void __launch_bounds__(kThreads, 1) allreduce_inplace_kernel(const AllreduceArgs a) {
  const BarrierCtx& b = a.bar;
  const uint32_t base = barrier_seq_begin(b);
  const bool want_stats = a.sink.slot >= 0;
  const int world = b.world;

  // peers' producers (their backward kernels) are stream-ordered before their copy of this kernel
  cross_rank_barrier(b, base + 1);

  const size_t per_rank = (a.nvec + world - 1) / world;
  const size_t lo = size_t(b.rank) * per_rank;
  const size_t hi = (lo + per_rank < a.nvec) ? lo + per_rank : a.nvec;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t first = lo + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  Stats st;
  constexpr int N = Vec16<T>::N;

  if constexpr (NVLS) {
    constexpr int U = W;
    char* mcb = a.mc + a.byte_off;
    for (size_t i = first; i < hi; i += U * stride) {
      uint4 s[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // all switch-side reductions of this trip are in flight together
        const size_t idx = i + u * stride;
        ok[u] = idx < hi;
        if (ok[u]) s[u] = MultimemOps<T>::ld_reduce(mcb + idx * 16);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        float f[N];
        Vec16<T>::unpack(s[u], f);
#pragma unroll
        for (int k = 0; k < N; ++k) f[k] = __fmul_rn(f[k], a.prescale);
        const uint4 o = Vec16<T>::pack(f);
        if (want_stats) {
          Vec16<T>::unpack(o, f);
#pragma unroll
          for (int k = 0; k < N; ++k) st.add(f[k]);
        }
        multimem_st16(mcb + (i + u * stride) * 16, o);
      }
    }
  } else if constexpr (W > 0) {
    constexpr int U = UnrollFor<W>::value;
    for (size_t i = first; i < hi; i += U * stride) {
      uint4 v[U][W];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t idx = i + u * stride;
        ok[u] = idx < hi;
        if (ok[u]) load_all_ranks<W>(a, a.byte_off + idx * 16, v[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const uint4 o = reduce_loaded<T, W>(v[u], a.prescale, st, want_stats);
        const size_t byte = a.byte_off + (i + u * stride) * 16;
#pragma unroll
        for (int r = 0; r < W; ++r) st_stream16(reinterpret_cast<char*>(a.bufs.p[r]) + byte, o);
      }
    }
  } else {
    for (size_t i = first; i < hi; i += stride) {
      const size_t byte = a.byte_off + i * 16;
      const uint4 o = reduce_vec_generic<T>(a, byte, a.prescale, st, want_stats);
      for (int r = 0; r < world; ++r) st_stream16(reinterpret_cast<char*>(a.bufs.p[r]) + byte, o);
    }
  }

  if (want_stats) publish_stats(a.sink, st);

  if (a.flags & B200DP_F_NO_END_BARRIER) {
    // a later kernel on this channel (or b200dp_fence) provides the "everything has landed" guarantee
    __syncthreads();
    barrier_seq_end(b, base, 1);
    return;
  }
  // all reduced slices (and stats records) have landed everywhere before anyone's kernel retires
  cross_rank_barrier(b, base + 2);
  barrier_seq_end(b, base, 2);
}

// barrier-only kernel (one CTA): closes a sequence of NO_END_BARRIER launches
__global__ void __launch_bounds__(32) fence_kernel(const BarrierCtx b) {
  const uint32_t base = barrier_seq_begin(b);
  cross_rank_barrier(b, base + 1);
  barrier_seq_end(b, base, 1);
}

// ------------------------------------------------------------------------------------------------
// K1b: out-of-place one-shot allreduce: dst_i = (sum_r staging_r[i]) / divisor.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct ElemTraits {
  using Acc = T;
  static __device__ __forceinline__ Acc up(T v) { return v; }
  static __device__ __forceinline__ T down(Acc v) { return v; }
  static constexpr bool kFloat = false;
};
template <>
struct ElemTraits<float> {
  using Acc = float;
  static __device__ __forceinline__ Acc up(float v) { return v; }
  static __device__ __forceinline__ float down(Acc v) { return v; }
  static constexpr bool kFloat = true;
};
template <>
struct ElemTraits<double> {
  using Acc = double;
  static __device__ __forceinline__ Acc up(double v) { return v; }
  static __device__ __forceinline__ double down(Acc v) { return v; }
  static constexpr bool kFloat = true;
};
template <>
struct ElemTraits<__nv_bfloat16> {
  using Acc = float;
  static __device__ __forceinline__ Acc up(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 down(Acc v) { return __float2bfloat16_rn(v); }
  static constexpr bool kFloat = true;
};
template <>
struct ElemTraits<__half> {
  using Acc = float;
  static __device__ __forceinline__ Acc up(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half down(Acc v) { return __float2half_rn(v); }
  static constexpr bool kFloat = true;
};

struct OneshotArgs {
  BarrierCtx bar;
  Peers bufs;
  size_t byte_off;  // staging offset, 16-byte aligned
  size_t count;     // elements
  double divisor;   // 1 = plain sum
  void* dst;        // local output, 16-byte aligned
};

template <typename T>
__device__ __forceinline__ T oneshot_finish(typename ElemTraits<T>::Acc acc, double divisor) {
  using Tr = ElemTraits<T>;
  T s = Tr::down(acc);
  if constexpr (Tr::kFloat) {
    if (divisor != 1.0) {  // reference: all_reduce(SUM) then `t /= N` in the tensor's own dtype
      using A = typename Tr::Acc;
      s = Tr::down(Tr::up(s) / A(divisor));
    }
  }
  return s;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) allreduce_oneshot_kernel(const OneshotArgs a) {
  using Tr = ElemTraits<T>;
  using Acc = typename Tr::Acc;
  constexpr int N = 16 / sizeof(T);
  const BarrierCtx& b = a.bar;
  const uint32_t base = barrier_seq_begin(b);
  cross_rank_barrier(b, base + 1);

  const size_t nvec = a.count / N;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Acc acc[N];
    union {
      uint4 v;
      T e[N];
    } u;
    u.v = ld_stream16(reinterpret_cast<const char*>(a.bufs.p[0]) + a.byte_off + i * 16);
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = Tr::up(u.e[k]);
    for (int r = 1; r < b.world; ++r) {
      u.v = ld_stream16(reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off + i * 16);
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] = acc[k] + Tr::up(u.e[k]);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) u.e[k] = oneshot_finish<T>(acc[k], a.divisor);
    reinterpret_cast<uint4*>(a.dst)[i] = u.v;
  }
  if (blockIdx.x == 0) {  // ragged tail, element-wise
    for (size_t e = nvec * N + threadIdx.x; e < a.count; e += blockDim.x) {
      Acc acc = Tr::up(reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.bufs.p[0]) + a.byte_off)[e]);
      for (int r = 1; r < b.world; ++r)
        acc = acc + Tr::up(reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off)[e]);
      reinterpret_cast<T*>(a.dst)[e] = oneshot_finish<T>(acc, a.divisor);
    }
  }
  // nobody may overwrite its staging before every peer has read it
  cross_rank_barrier(b, base + 2);
  barrier_seq_end(b, base, 2);
}

// ------------------------------------------------------------------------------------------------
// K2/K3: pull `nbytes` from peers' staging into local memory (allgather, broadcast).  Bit-exact.
// ------------------------------------------------------------------------------------------------
struct PullArgs {
  BarrierCtx bar;
  Peers bufs;
  size_t byte_off;
  size_t nbytes;
  char* dst;
  size_t dst_stride;
  int first, nranks;
};

__device__ __forceinline__ void copy_span(char* dst, const char* src, size_t nbytes, size_t tid, size_t nthreads) {
  // src is always 16-byte aligned (staging); pick the widest unit dst allows
  const uintptr_t d = reinterpret_cast<uintptr_t>(dst);
  if ((d & 15) == 0) {
    const size_t nvec = nbytes >> 4;
    constexpr int U = 4;  // peer loads in flight per thread (NVLink latency is ~2-3 us)
    for (size_t i = tid; i < nvec; i += U * nthreads) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + u * nthreads;
        if (j < nvec) v[u] = ld_stream16(src + j * 16);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + u * nthreads;
        if (j < nvec) reinterpret_cast<uint4*>(dst)[j] = v[u];
      }
    }
    for (size_t i = (nvec << 4) + tid; i < nbytes; i += nthreads) dst[i] = src[i];
  } else if ((d & 3) == 0) {
    const size_t nw = nbytes >> 2;
    for (size_t i = tid; i < nw; i += nthreads)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    for (size_t i = (nw << 2) + tid; i < nbytes; i += nthreads) dst[i] = src[i];
  } else {
    for (size_t i = tid; i < nbytes; i += nthreads) dst[i] = src[i];
  }
}

__global__ void __launch_bounds__(kThreads, 1) pull_kernel(const PullArgs a) {
  const BarrierCtx& b = a.bar;
  const uint32_t base = barrier_seq_begin(b);
  cross_rank_barrier(b, base + 1);
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  // start with the next rank so that the N readers fan out over N different sources
  for (int k = 1; k <= b.world; ++k) {
    const int r = (b.rank + k) % b.world;
    if (r < a.first || r >= a.first + a.nranks) continue;
    copy_span(a.dst + size_t(r - a.first) * a.dst_stride, reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off,
              a.nbytes, tid, nthreads);
  }
  cross_rank_barrier(b, base + 2);
  barrier_seq_end(b, base, 2);
}

// ------------------------------------------------------------------------------------------------
// K2c / K1c: fused large-message collectives for tensors that live in ordinary (not peer-mapped) torch memory.
// ONE launch each: CTA b stages ITS portion of the local input into the symmetric staging segment, meets CTA b of
// every other rank at a flag barrier, and then moves exactly that portion over NVLink.  Because the hand-off is per
// CTA, portions pipeline against each other (some CTAs still stage while others already pull) and there is no
// separate pack launch, no grid-wide barrier and no second pass over the output.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_any(char* dst, const char* src, size_t nbytes, size_t tid, size_t nthreads) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src);
  if ((al & 15) == 0) {
    const size_t nvec = nbytes >> 4;
    constexpr int U = 4;
    for (size_t i = tid; i < nvec; i += U * nthreads) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + u * nthreads;
        if (j < nvec) v[u] = ld_stream16(src + j * 16);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + u * nthreads;
        if (j < nvec) st_stream16(dst + j * 16, v[u]);
      }
    }
    for (size_t i = (nvec << 4) + tid; i < nbytes; i += nthreads) dst[i] = src[i];
  } else if ((al & 3) == 0) {
    const size_t nw = nbytes >> 2;
    for (size_t i = tid; i < nw; i += nthreads)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    for (size_t i = (nw << 2) + tid; i < nbytes; i += nthreads) dst[i] = src[i];
  } else {
    for (size_t i = tid; i < nbytes; i += nthreads) dst[i] = src[i];
  }
}

struct GatherArgs {
  BarrierCtx bar;
  Peers bufs;          // staging segment bases
  size_t byte_off;     // staging window (16-byte aligned)
  size_t nbytes;       // bytes per rank in this call
  const char* src;     // local input
  char* dst;           // local output: rank r's block at dst + r * dst_stride
  size_t dst_stride;
};

__global__ void __launch_bounds__(kThreads, 1) gather_fused_kernel(const GatherArgs a) {
  const BarrierCtx& b = a.bar;
  const uint32_t base = barrier_seq_begin(b);
  const size_t nvec = (a.nbytes + 15) >> 4;
  const size_t per = (nvec + gridDim.x - 1) / gridDim.x;
  size_t lo = size_t(blockIdx.x) * per * 16, hi = lo + per * 16;
  if (lo > a.nbytes) lo = a.nbytes;
  if (hi > a.nbytes) hi = a.nbytes;
  const size_t n = hi - lo;
  char* mine = reinterpret_cast<char*>(a.bufs.p[b.rank]) + a.byte_off + lo;
  copy_any(mine, a.src + lo, n, threadIdx.x, blockDim.x);
  cross_rank_barrier(b, base + 1);  // CTA b of every rank has staged portion b
  // All N sources at once: the 16 warps split into (up to) N groups, group g copies from rank (rank + g) % N, so the
  // local block (HBM) and every peer's block (NVLink) stream concurrently instead of one after the other, and each
  // GPU serves its N - 1 readers evenly.  With more ranks than warps the groups take several sources in turn.
  const int nwarps = blockDim.x >> 5;
  const int groups = b.world < nwarps ? b.world : nwarps;
  const int per_group = nwarps / groups;               // warps per group (the last group takes the remainder)
  int gidx = (threadIdx.x >> 5) / per_group;
  if (gidx >= groups) gidx = groups - 1;
  const int gfirst = gidx * per_group;
  const int gwarps = (gidx == groups - 1) ? nwarps - gfirst : per_group;
  const size_t gtid = size_t(threadIdx.x) - size_t(gfirst) * 32;
  const size_t gthreads = size_t(gwarps) * 32;
  for (int k = gidx; k < b.world; k += groups) {
    const int r = (b.rank + k) % b.world;
    copy_span(a.dst + size_t(r) * a.dst_stride + lo, reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off + lo, n, gtid,
              gthreads);
  }
  // nobody may restage (next call) while a peer still reads this window
  cross_rank_barrier(b, base + 2);
  barrier_seq_end(b, base, 2);
}

struct ReduceArgs {
  BarrierCtx bar;
  Peers bufs;        // staging segment bases
  char* mc;          // multicast VA of the staging segment (NVLS) or nullptr
  size_t byte_off;   // staging window (16-byte aligned)
  size_t count;      // elements
  const void* src;   // local input
  void* dst;         // local output
  float divisor;     // 1 = plain sum; else out = T(float(T(sum)) / divisor)  (the reference's `t /= N` in t's dtype)
};

// T in {float, bf16, half}.  W > 0: ranks known at compile time; W == 0: any world size.
template <typename T, int W, bool NVLS>
__global__ void __launch_bounds__(kThreads, 1) reduce_fused_kernel(const ReduceArgs a) {
  constexpr int N = Vec16<T>::N;
  const BarrierCtx& b = a.bar;
  const uint32_t base = barrier_seq_begin(b);
  const int world = b.world;
  const size_t nbytes = a.count * sizeof(T);
  const size_t nvec = (nbytes + 15) >> 4;
  const size_t per = (nvec + gridDim.x - 1) / gridDim.x;
  size_t vlo = size_t(blockIdx.x) * per, vhi = vlo + per;
  if (vlo > nvec) vlo = nvec;
  if (vhi > nvec) vhi = nvec;
  size_t blo = vlo * 16, bhi = vhi * 16;
  if (bhi > nbytes) bhi = nbytes;
  if (blo > nbytes) blo = nbytes;
  char* mine = reinterpret_cast<char*>(a.bufs.p[b.rank]) + a.byte_off;
  // A: stage this CTA's portion (the bytes past `count` in the last vector keep whatever they held: never copied out)
  copy_any(mine + blo, reinterpret_cast<const char*>(a.src) + blo, bhi - blo, threadIdx.x, blockDim.x);
  cross_rank_barrier(b, base + 1);
  // B: rank r owns the r-th 1/N of the portion: reduce it over all ranks' staged copies (rank order, fp32 accumulate,
  // one rounding to T), store the result into every rank's staging
  const size_t sub = (vhi - vlo + world - 1) / world;
  size_t slo = vlo + size_t(b.rank) * sub, shi = slo + sub;
  if (slo > vhi) slo = vhi;
  if (shi > vhi) shi = vhi;
  if constexpr (NVLS) {
    char* mcb = a.mc + a.byte_off;
    constexpr int U = 4;
    for (size_t i = slo + threadIdx.x; i < shi; i += size_t(U) * blockDim.x) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j < shi) v[u] = MultimemOps<T>::ld_reduce(mcb + j * 16);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j < shi) multimem_st16(mcb + j * 16, v[u]);
      }
    }
  } else {
    constexpr int U = (W == 0 || W >= 5) ? 2 : 4;
    for (size_t i = slo + threadIdx.x; i < shi; i += size_t(U) * blockDim.x) {
      float acc[U][N];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + size_t(u) * blockDim.x;
        ok[u] = j < shi;
        if (!ok[u]) continue;
        if constexpr (W > 0) {
          uint4 v[W];
#pragma unroll
          for (int r = 0; r < W; ++r) v[r] = ld_stream16(reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off + j * 16);
          Vec16<T>::unpack(v[0], acc[u]);
#pragma unroll
          for (int r = 1; r < W; ++r) {
            float x[N];
            Vec16<T>::unpack(v[r], x);
#pragma unroll
            for (int k = 0; k < N; ++k) acc[u][k] = __fadd_rn(acc[u][k], x[k]);
          }
        } else {
          Vec16<T>::unpack(ld_stream16(reinterpret_cast<const char*>(a.bufs.p[0]) + a.byte_off + j * 16), acc[u]);
          for (int r = 1; r < world; ++r) {
            float x[N];
            Vec16<T>::unpack(ld_stream16(reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off + j * 16), x);
#pragma unroll
            for (int k = 0; k < N; ++k) acc[u][k] = __fadd_rn(acc[u][k], x[k]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const size_t j = i + size_t(u) * blockDim.x;
        const uint4 o = Vec16<T>::pack(acc[u]);
        for (int r = 0; r < world; ++r) st_stream16(reinterpret_cast<char*>(a.bufs.p[r]) + a.byte_off + j * 16, o);
      }
    }
  }
  cross_rank_barrier(b, base + 2);  // every owner's results for portion b have landed in my staging
  // C: staging -> output, dividing on the way if asked (peers only ever read my staging before barrier 2 and write it
  // between the barriers, so the next call may restage at once: no closing barrier)
  const char* st = mine;
  char* out = reinterpret_cast<char*>(a.dst);
  if (a.divisor == 1.0f) {
    copy_span(out + blo, st + blo, bhi - blo, threadIdx.x, blockDim.x);
  } else {
    const size_t elo = blo / sizeof(T), ehi = bhi / sizeof(T);
    const T* sp = reinterpret_cast<const T*>(st);
    T* op = reinterpret_cast<T*>(out);
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      for (size_t j = vlo + threadIdx.x; j < vhi; j += blockDim.x) {
        if ((j + 1) * N <= a.count) {
          float f[N];
          Vec16<T>::unpack(ld_stream16(st + j * 16), f);
#pragma unroll
          for (int k = 0; k < N; ++k) f[k] = __fdiv_rn(f[k], a.divisor);
          *reinterpret_cast<uint4*>(out + j * 16) = Vec16<T>::pack(f);
        } else {
          for (size_t e = j * N; e < a.count; ++e) op[e] = from_float<T>(__fdiv_rn(to_float<T>(sp[e]), a.divisor));
        }
      }
    } else {
      for (size_t e = elo + threadIdx.x; e < ehi; e += blockDim.x) op[e] = from_float<T>(__fdiv_rn(to_float<T>(sp[e]), a.divisor));
    }
  }
  barrier_seq_end(b, base, 2);
}

// ------------------------------------------------------------------------------------------------
// K1d: pipelined out-of-place allreduce for large tensors in ordinary torch memory.  One launch, two kinds of CTA:
//   * COMM CTAs (the first `ncomm`, the same indices on every rank) own the NVLink side: per chunk they wait until the
//     local copy CTAs staged it, meet their peers at a flag barrier, reduce this rank's share of the chunk over all
//     ranks' staged copies (NVLS multimem, or two-shot peer loads / stores) and, after a second barrier, hand the
//     chunk back;
//   * COPY CTAs (all the others) own the HBM side: stage chunk c (input -> symmetric staging), then copy a finished
//     chunk (staging -> output, dividing on the way for "mean"), a few chunks behind.
// Chunk c+1 is being staged and chunk c-1 copied out while chunk c crosses the switch, so the time is
// max(NVLink, HBM) instead of their sum — what the three-launch sequence (pack, in-place allreduce, pack) cannot do.
// NVLS streams best with few CTAs, HBM copies want many: the split gives each side what it needs.
// Local hand-off = two counters per chunk in device memory (zeroed by the host before the launch).
// ------------------------------------------------------------------------------------------------
struct PipeReduceArgs {
  BarrierCtx bar;
  Peers bufs;          // staging segment bases
  char* mc;            // multicast VA of the staging segment (NVLS) or nullptr
  size_t byte_off;     // staging window (16-byte aligned)
  size_t count;        // elements
  const void* src;
  void* dst;
  float divisor;
  uint32_t* staged;    // [nchunks] arrivals of copy CTAs (local)
  uint32_t* reduced;   // [nchunks] arrivals of comm CTAs (local)
  size_t chunk_vecs;   // 16-byte vectors per chunk
  int nchunks;
  int ncomm;           // comm CTAs; gridDim.x - ncomm copy CTAs
  int lag;             // copy CTAs copy out chunk c - lag after staging chunk c
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all threads of the CTA call it; returns when *p >= want (or the context timed out)
__device__ __forceinline__ void wait_counter(const BarrierCtx& b, const uint32_t* p, uint32_t want) {
  if (threadIdx.x == 0) {
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while (ld_acquire_gpu(p) < want) {
      if ((++spins & 0x3ff) == 0) {
        unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > b.timeout_ns || *reinterpret_cast<volatile int*>(b.status) != 0) {
          *reinterpret_cast<volatile int*>(b.status) = B200DP_ETIMEOUT;
          __threadfence_system();
          break;
        }
      }
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void signal_counter(uint32_t* p) {
  __syncthreads();  // every thread's stores of this phase are ordered before the release below
  if (threadIdx.x == 0) red_release_gpu_add(p, 1u);
}

template <typename T, int W, bool NVLS>
__global__ void __launch_bounds__(kThreads, 1) reduce_pipelined_kernel(const PipeReduceArgs a) {
  constexpr int N = Vec16<T>::N;
  const BarrierCtx& b = a.bar;
  const int world = b.world;
  const size_t nbytes = a.count * sizeof(T);
  const size_t nvec = (nbytes + 15) >> 4;
  char* mine = reinterpret_cast<char*>(a.bufs.p[b.rank]) + a.byte_off;

  if (int(blockIdx.x) < a.ncomm) {
    // ------------------------------------------------------------------ COMM role
    // two groups (even / odd CTA index) take the even / odd chunks: while one group sits in its flag barriers the
    // other one streams, so the links never idle for a barrier round trip
    const uint32_t base = barrier_seq_begin(b);
    const uint32_t ncopy = gridDim.x - a.ncomm;
    const int ngroups = a.ncomm >= 2 ? 2 : 1;
    const int group = int(blockIdx.x) % ngroups, gidx = int(blockIdx.x) / ngroups;
    const int gsize = (a.ncomm - group + ngroups - 1) / ngroups;
    uint32_t k = 0;  // chunks this CTA has handled
    int prev = -1;
    for (int c = group; c < a.nchunks; c += ngroups, ++k) {
      const size_t clo = size_t(c) * a.chunk_vecs;
      const size_t chi = (clo + a.chunk_vecs < nvec) ? clo + a.chunk_vecs : nvec;
      wait_counter(b, a.staged + c, ncopy);  // my copy CTAs staged chunk c
      // ONE barrier per chunk.  A peer CTA that arrives here has (a) staged chunk c and (b) finished storing its share of
      // this group's previous chunk into my staging (its stores precede its flag release in program order): the
      // "results have landed" barrier of chunk `prev` and the "inputs are staged" barrier of chunk c are the same one.
      cross_rank_barrier(b, base + 1 + k);
      if (prev >= 0 && threadIdx.x == 0) red_release_gpu_add(a.reduced + prev, 1u);
      prev = c;
      // this rank's share of the chunk, then this comm CTA's share of that
      const size_t per_rank = (chi - clo + world - 1) / world;
      size_t rlo = clo + size_t(b.rank) * per_rank, rhi = rlo + per_rank;
      if (rlo > chi) rlo = chi;
      if (rhi > chi) rhi = chi;
      const size_t per_cta = (rhi - rlo + gsize - 1) / gsize;
      size_t slo = rlo + size_t(gidx) * per_cta, shi = slo + per_cta;
      if (slo > rhi) slo = rhi;
      if (shi > rhi) shi = rhi;
      if constexpr (NVLS) {
        char* mcb = a.mc + a.byte_off;
        constexpr int U = 4;
        for (size_t i = slo + threadIdx.x; i < shi; i += size_t(U) * blockDim.x) {
          uint4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t j = i + size_t(u) * blockDim.x;
            if (j < shi) v[u] = MultimemOps<T>::ld_reduce(mcb + j * 16);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t j = i + size_t(u) * blockDim.x;
            if (j < shi) multimem_st16(mcb + j * 16, v[u]);
          }
        }
      } else {
        constexpr int U = (W == 0 || W >= 5) ? 2 : 4;
        for (size_t i = slo + threadIdx.x; i < shi; i += size_t(U) * blockDim.x) {
          float acc[U][N];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t j = i + size_t(u) * blockDim.x;
            ok[u] = j < shi;
            if (!ok[u]) continue;
            if constexpr (W > 0) {
              uint4 v[W];
#pragma unroll
              for (int r = 0; r < W; ++r) v[r] = ld_stream16(reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off + j * 16);
              Vec16<T>::unpack(v[0], acc[u]);
#pragma unroll
              for (int r = 1; r < W; ++r) {
                float x[N];
                Vec16<T>::unpack(v[r], x);
#pragma unroll
                for (int k = 0; k < N; ++k) acc[u][k] = __fadd_rn(acc[u][k], x[k]);
              }
            } else {
              Vec16<T>::unpack(ld_stream16(reinterpret_cast<const char*>(a.bufs.p[0]) + a.byte_off + j * 16), acc[u]);
              for (int r = 1; r < world; ++r) {
                float x[N];
                Vec16<T>::unpack(ld_stream16(reinterpret_cast<const char*>(a.bufs.p[r]) + a.byte_off + j * 16), x);
#pragma unroll
                for (int k = 0; k < N; ++k) acc[u][k] = __fadd_rn(acc[u][k], x[k]);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const size_t j = i + size_t(u) * blockDim.x;
            const uint4 o = Vec16<T>::pack(acc[u]);
            for (int r = 0; r < world; ++r) st_stream16(reinterpret_cast<char*>(a.bufs.p[r]) + a.byte_off + j * 16, o);
          }
        }
      }
    }
    if (prev >= 0) {
      cross_rank_barrier(b, base + 1 + k);  // closing: every owner's share of the last chunk has landed in my staging
      if (threadIdx.x == 0) red_release_gpu_add(a.reduced + prev, 1u);
      ++k;
    }
    barrier_seq_end(b, base, k);  // the same k on every rank: it depends only on (blockIdx, nchunks, ncomm)
    return;
  }

  // -------------------------------------------------------------------- COPY role
  const int me = int(blockIdx.x) - a.ncomm, ncopy = int(gridDim.x) - a.ncomm;
  const size_t tid = threadIdx.x, nthreads = blockDim.x;
  auto my_span = [&](int c, size_t& blo, size_t& bhi) {
    const size_t clo = size_t(c) * a.chunk_vecs;
    const size_t chi = (clo + a.chunk_vecs < nvec) ? clo + a.chunk_vecs : nvec;
    const size_t per = (chi - clo + ncopy - 1) / ncopy;
    size_t vlo = clo + size_t(me) * per, vhi = vlo + per;
    if (vlo > chi) vlo = chi;
    if (vhi > chi) vhi = chi;
    blo = vlo * 16;
    bhi = vhi * 16;
    if (blo > nbytes) blo = nbytes;
    if (bhi > nbytes) bhi = nbytes;
  };
  const int ngroups_c = a.ncomm >= 2 ? 2 : 1;
  auto copy_out = [&](int c) {
    const int g = c % ngroups_c;
    wait_counter(b, a.reduced + c, uint32_t((a.ncomm - g + ngroups_c - 1) / ngroups_c));
    size_t blo, bhi;
    my_span(c, blo, bhi);
    char* out = reinterpret_cast<char*>(a.dst);
    if (a.divisor == 1.0f) {
      copy_span(out + blo, mine + blo, bhi - blo, tid, nthreads);
    } else if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      const T* sp = reinterpret_cast<const T*>(mine);
      T* op = reinterpret_cast<T*>(out);
      for (size_t j = blo / 16 + tid; j < (bhi + 15) / 16; j += nthreads) {
        if ((j + 1) * N <= a.count) {
          float f[N];
          Vec16<T>::unpack(ld_stream16(mine + j * 16), f);
#pragma unroll
          for (int q = 0; q < N; ++q) f[q] = __fdiv_rn(f[q], a.divisor);
          *reinterpret_cast<uint4*>(out + j * 16) = Vec16<T>::pack(f);
        } else {
          for (size_t e = j * N; e < a.count; ++e) op[e] = from_float<T>(__fdiv_rn(to_float<T>(sp[e]), a.divisor));
        }
      }
    } else {
      const T* sp = reinterpret_cast<const T*>(mine);
      T* op = reinterpret_cast<T*>(out);
      for (size_t e = blo / sizeof(T) + tid; e < bhi / sizeof(T); e += nthreads)
        op[e] = from_float<T>(__fdiv_rn(to_float<T>(sp[e]), a.divisor));
    }
  };
  for (int c = 0; c < a.nchunks; ++c) {
    size_t blo, bhi;
    my_span(c, blo, bhi);
    copy_any(mine + blo, reinterpret_cast<const char*>(a.src) + blo, bhi - blo, tid, nthreads);
    signal_counter(a.staged + c);
    if (c >= a.lag) copy_out(c - a.lag);
  }
  for (int c = (a.nchunks > a.lag ? a.nchunks - a.lag : 0); c < a.nchunks; ++c) copy_out(c);
}

// ------------------------------------------------------------------------------------------------
// K1e: two-rank out-of-place allreduce by PUSH.  With two ranks a one-shot exchange moves exactly as many bytes over
// NVLink as the two-shot (S per direction) and needs no staging copy and no copy-out: every CTA pushes its portion of the
// local input straight into the PEER's window (posted NVLink stores), round by round, and — one flag barrier per round,
// the next round's push already in flight — adds what the peer pushed to its own input in rank order (x_0 + x_1, fp32
// accumulate, one rounding) and writes the output.  HBM: read S (push) + read 2S + write S (+ S incoming).
// Preconditions (checked by the host): world == 2, src / dst 16-byte aligned, byte count a multiple of 16.
// ------------------------------------------------------------------------------------------------
struct Push2Args {
  BarrierCtx bar;
  Peers bufs;          // staging segment bases
  size_t byte_off;     // staging window (16-byte aligned)
  size_t nvec;         // 16-byte vectors
  const char* src;
  char* dst;
  float divisor;
  size_t round_vecs;   // vectors per CTA per round
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) reduce_push2_kernel(const Push2Args a) {
  constexpr int N = Vec16<T>::N;
  const BarrierCtx& b = a.bar;
  const uint32_t base = barrier_seq_begin(b);
  const int peer = 1 - b.rank;
  const size_t per = (a.nvec + gridDim.x - 1) / gridDim.x;
  size_t vlo = size_t(blockIdx.x) * per, vhi = vlo + per;
  if (vlo > a.nvec) vlo = a.nvec;
  if (vhi > a.nvec) vhi = a.nvec;
  // the number of rounds must be the same for CTA b on both ranks: it depends only on (nvec, grid, round size)
  const uint32_t rounds = uint32_t((per + a.round_vecs - 1) / a.round_vecs);
  char* peer_win = reinterpret_cast<char*>(a.bufs.p[peer]) + a.byte_off;
  const char* my_win = reinterpret_cast<const char*>(a.bufs.p[b.rank]) + a.byte_off;
  constexpr int U = 4;
  auto push = [&](uint32_t r) {
    size_t lo = vlo + size_t(r) * a.round_vecs, hi = lo + a.round_vecs;
    if (lo > vhi) lo = vhi;
    if (hi > vhi) hi = vhi;
    for (size_t i = lo + threadIdx.x; i < hi; i += size_t(U) * blockDim.x) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j < hi) v[u] = ld_stream16(a.src + j * 16);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j < hi) st_stream16(peer_win + j * 16, v[u]);
      }
    }
  };
  if (rounds > 0) push(0);
  for (uint32_t r = 0; r < rounds; ++r) {
    if (r + 1 < rounds) push(r + 1);     // keep the link busy while waiting for the peer's round r
    cross_rank_barrier(b, base + 1 + r);  // the peer's rounds <= r have landed in my window
    size_t lo = vlo + size_t(r) * a.round_vecs, hi = lo + a.round_vecs;
    if (lo > vhi) lo = vhi;
    if (hi > vhi) hi = vhi;
    for (size_t i = lo + threadIdx.x; i < hi; i += size_t(U) * blockDim.x) {
      uint4 x[U], y[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j < hi) {
          x[u] = ld_stream16(a.src + j * 16);
          y[u] = ld_stream16(my_win + j * 16);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t j = i + size_t(u) * blockDim.x;
        if (j >= hi) continue;
        float f0[N], f1[N];
        Vec16<T>::unpack(b.rank == 0 ? x[u] : y[u], f0);  // rank order: x_0 first
        Vec16<T>::unpack(b.rank == 0 ? y[u] : x[u], f1);
#pragma unroll
        for (int k = 0; k < N; ++k) f0[k] = __fadd_rn(f0[k], f1[k]);
        uint4 o = Vec16<T>::pack(f0);
        if (a.divisor != 1.0f) {
          Vec16<T>::unpack(o, f0);
#pragma unroll
          for (int k = 0; k < N; ++k) f0[k] = __fdiv_rn(f0[k], a.divisor);
          o = Vec16<T>::pack(f0);
        }
        *reinterpret_cast<uint4*>(a.dst + j * 16) = o;
      }
    }
  }
  // the peer must be done reading its window before my next call pushes into it again
  cross_rank_barrier(b, base + 1 + rounds);
  barrier_seq_end(b, base, rounds + 1);
}

// ------------------------------------------------------------------------------------------------
// K2b: latency path for small messages: ONE launch, ONE barrier.  Every rank pushes its data into a
// fixed window of every peer's staging segment (posted NVLink stores), then reads all N windows
// locally and either concatenates (gather) or sums in rank order (reduce).  Windows are private to a
// CTA index and double-buffered by the parity of that CTA's barrier counter, so no closing barrier is
// needed: a window is rewritten two calls later, after every peer passed the next call's barrier.
// ------------------------------------------------------------------------------------------------
constexpr int kPushCtas = 32;             // CTAs (= windows) per call
constexpr size_t kPushWindow = 16384;     // bytes per (CTA, parity, source rank)
constexpr size_t kPushAreaBytes = size_t(kPushCtas) * 2 * B200DP_MAX_WORLD * kPushWindow;  // 16 MiB

struct PushArgs {
  BarrierCtx bar;
  Peers bufs;        // staging segment bases
  size_t area_off;   // start of the push area inside the staging segment
  const char* src;   // local input, nbytes
  char* dst;         // local output: gather: world*nbytes, reduce: nbytes
  size_t nbytes;     // per rank
  double divisor;    // reduce: 1 = plain sum
};

__device__ __forceinline__ char* push_window(void* base, size_t area_off, int cta, uint32_t parity, int rank) {
  return reinterpret_cast<char*>(base) + area_off +
         ((size_t(cta) * 2 + parity) * B200DP_MAX_WORLD + rank) * kPushWindow;
}

// MODE 0: gather (bytes).  MODE 1: reduce over elements of type T.
template <int MODE, typename T>
__global__ void __launch_bounds__(kThreads, 1) push_kernel(const PushArgs a) {
  const BarrierCtx& b = a.bar;
  const uint32_t base = barrier_seq_begin(b);
  const uint32_t parity = base & 1u;  // this kernel is the only one that advances the counter by an odd amount
  const size_t lo = size_t(blockIdx.x) * kPushWindow;
  const size_t n = a.nbytes > lo ? (a.nbytes - lo < kPushWindow ? a.nbytes - lo : kPushWindow) : 0;
  const char* src = a.src + lo;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(a.dst) | a.nbytes) & 15) == 0;

  // phase 1: my bytes -> window [cta][parity][my rank] of every rank (including mine)
  if (vec) {
    for (size_t i = threadIdx.x; i < (n >> 4); i += blockDim.x) {
      const uint4 v = ld_stream16(src + i * 16);
      for (int r = 0; r < b.world; ++r) st_stream16(push_window(a.bufs.p[r], a.area_off, blockIdx.x, parity, b.rank) + i * 16, v);
    }
  } else {
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
      const char v = src[i];
      for (int r = 0; r < b.world; ++r) push_window(a.bufs.p[r], a.area_off, blockIdx.x, parity, b.rank)[i] = v;
    }
  }
  cross_rank_barrier(b, base + 1);

  // phase 2: everything is local now
  char* mine = reinterpret_cast<char*>(a.bufs.p[b.rank]);
  if constexpr (MODE == 0) {
    for (int r = 0; r < b.world; ++r) {
      const char* w = push_window(mine, a.area_off, blockIdx.x, parity, r);
      char* d = a.dst + size_t(r) * a.nbytes + lo;
      if (vec) {
        for (size_t i = threadIdx.x; i < (n >> 4); i += blockDim.x) reinterpret_cast<uint4*>(d)[i] = *reinterpret_cast<const uint4*>(w + i * 16);
      } else {
        for (size_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = w[i];
      }
    }
  } else {
    using Tr = ElemTraits<T>;
    using Acc = typename Tr::Acc;
    const size_t ne = n / sizeof(T);
    T* d = reinterpret_cast<T*>(a.dst + lo);
    for (size_t i = threadIdx.x; i < ne; i += blockDim.x) {
      Acc acc = Tr::up(reinterpret_cast<const T*>(push_window(mine, a.area_off, blockIdx.x, parity, 0))[i]);
      for (int r = 1; r < b.world; ++r)
        acc = acc + Tr::up(reinterpret_cast<const T*>(push_window(mine, a.area_off, blockIdx.x, parity, r))[i]);
      d[i] = oneshot_finish<T>(acc, a.divisor);
    }
  }
  barrier_seq_end(b, base, 1);
}

// ------------------------------------------------------------------------------------------------
// K4: multi-tensor copy (bucket pack / unpack / staging).  One launch for a whole bucket.
// ------------------------------------------------------------------------------------------------
// CAP = table capacity.  The table travels as a kernel parameter; a gradient bucket holds ~10 tensors, and launching with
// a 900-byte table instead of an 18 KB one is several microseconds cheaper on the (host-bound) autograd thread.
template <int CAP>
struct PackTableT {
  const char* src[CAP];
  char* dst[CAP];
  unsigned long long bytes[CAP];
  uint32_t chunk_start[CAP + 1];  // prefix sum of ceil(bytes / chunk_bytes)
  uint32_t chunk_bytes;           // work unit: kPackChunk (LDG/STG) or one TMA stage (bulk)
  int n;
};
constexpr int kPackSmall = 32;
using PackTable = PackTableT<kMTMax>;
using PackTableSmall = PackTableT<kPackSmall>;

__device__ __forceinline__ int find_tensor(const uint32_t* chunk_start, int n, uint32_t chunk) {
  int lo = 0, hi = n;  // last t with chunk_start[t] <= chunk
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (chunk_start[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

// LDG/STG variant: any alignment.  A CTA takes 32 KB chunks; every thread keeps kPackILP 16-byte loads in flight
// before the first store of a trip (256 threads x 4 x 16 B = 16 KB per CTA, x up to 8 CTAs per SM), which is what
// covers HBM latency; no shared memory and few registers, so it co-resides with the backward GEMMs it overlaps.
constexpr int kPackILP = 4;
template <int CAP>
__global__ void __launch_bounds__(kStreamThreads) pack_kernel(const __grid_constant__ PackTableT<CAP> t) {
  const uint32_t total = t.chunk_start[t.n];
  const size_t chunk_bytes = t.chunk_bytes;
  for (uint32_t c = blockIdx.x; c < total; c += gridDim.x) {
    const int ti = find_tensor(t.chunk_start, t.n, c);
    const size_t off = size_t(c - t.chunk_start[ti]) * chunk_bytes;
    const size_t rem = t.bytes[ti] - off;
    const size_t n = rem < chunk_bytes ? rem : chunk_bytes;
    const char* s = t.src[ti] + off;
    char* d = t.dst[ti] + off;
    const uintptr_t al = reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d);
    if ((al & 15) == 0) {
      const size_t nvec = n >> 4;
      for (size_t i = threadIdx.x; i < nvec; i += size_t(kPackILP) * blockDim.x) {
        uint4 v[kPackILP];
#pragma unroll
        for (int u = 0; u < kPackILP; ++u) {
          const size_t j = i + size_t(u) * blockDim.x;
          if (j < nvec) v[u] = ld_stream16(s + j * 16);
        }
#pragma unroll
        for (int u = 0; u < kPackILP; ++u) {
          const size_t j = i + size_t(u) * blockDim.x;
          if (j < nvec) st_stream16(d + j * 16, v[u]);
        }
      }
      for (size_t i = (nvec << 4) + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    } else if ((al & 3) == 0) {
      const size_t nw = n >> 2;
      for (size_t i = threadIdx.x; i < nw; i += blockDim.x)
        reinterpret_cast<uint32_t*>(d)[i] = reinterpret_cast<const uint32_t*>(s)[i];
      for (size_t i = (nw << 2) + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    } else {
      for (size_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
  }
}

// --- TMA-engine variant: cp.async.bulk global->smem->global.  Every warp of the CTA elects one thread that drives its
// own ring of `stages` stages (stage size = t.chunk_bytes) over the chunks g, g+G, ... (g = global warp index), so a
// CTA has warps x stages bulk transfers in flight; the LSU and the register file stay idle, the copy engine streams.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int kBulkMaxWarps = 8;
constexpr int kBulkMaxStages = 8;

// All spans must be 16-byte aligned with a size that is a multiple of 16 (the host splits off the rest).
template <int CAP>
__global__ void __launch_bounds__(kBulkMaxWarps * 32) pack_bulk_kernel(const __grid_constant__ PackTableT<CAP> t, int* status,
                                                                        int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_all[kBulkMaxWarps][kBulkMaxStages];
  if ((threadIdx.x & 31) != 0) return;  // one elected thread per warp; the TMA engine does the work
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint64_t* full = full_all[warp];
  const uint32_t stage_bytes = t.chunk_bytes;
  unsigned char* ring = smem + size_t(warp) * stages * stage_bytes;
  const uint32_t total = t.chunk_start[t.n];
  for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  const uint32_t g = blockIdx.x * nwarps + warp, G = gridDim.x * nwarps;
  auto chunk_of = [&](uint32_t j, const char*& s, char*& d, uint32_t& n) {
    const uint32_t c = g + j * G;
    const int ti = find_tensor(t.chunk_start, t.n, c);
    const size_t off = size_t(c - t.chunk_start[ti]) * stage_bytes;
    const size_t rem = t.bytes[ti] - off;
    n = uint32_t(rem < size_t(stage_bytes) ? rem : size_t(stage_bytes));
    s = t.src[ti] + off;
    d = t.dst[ti] + off;
  };
  const uint32_t mine = g < total ? (total - g + G - 1) / G : 0;
  auto issue_load = [&](uint32_t j) {
    const char* s;
    char* d;
    uint32_t n;
    chunk_of(j, s, d, n);
    const int st = j % stages;
    mbar_expect_tx(&full[st], n);
    bulk_g2s(ring + size_t(st) * stage_bytes, s, n, &full[st]);
  };
  const uint32_t pro = mine < uint32_t(stages) ? mine : uint32_t(stages);
  for (uint32_t j = 0; j < pro; ++j) issue_load(j);
  for (uint32_t j = 0; j < mine; ++j) {
    const int st = j % stages;
    const uint32_t parity = (j / stages) & 1;
    uint32_t spins = 0;
    while (!mbar_try_wait(&full[st], parity)) {
      if (++spins > (1u << 26)) {  // ~seconds: never hang the GPU on a lost transaction
        *reinterpret_cast<volatile int*>(status) = B200DP_ETIMEOUT;
        return;
      }
    }
    const char* s;
    char* d;
    uint32_t n;
    chunk_of(j, s, d, n);
    bulk_s2g(d, ring + size_t(st) * stage_bytes, n);
    bulk_commit();
    // refill the stage used by chunk j-1 once its store has finished reading shared memory
    if (j >= 1 && (j - 1) + stages < mine) {
      bulk_wait_read<1>();
      issue_load(j - 1 + stages);
    }
  }
  bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------
// K5: multi-tensor statistics / scaling for gradients outside a segment.
// ------------------------------------------------------------------------------------------------
struct MTTable {
  void* ptr[kMTMax];
  unsigned long long numel[kMTMax];
  uint32_t chunk_start[kMTMax + 1];  // prefix sum of ceil(numel / kMTChunk)
  int n;
};

template <typename T>
__global__ void __launch_bounds__(kStreamThreads) multi_stats_kernel(const __grid_constant__ MTTable t,
                                                                      const StatsSink sink) {
  constexpr int N = Vec16<T>::N;
  const uint32_t total = t.chunk_start[t.n];
  Stats st;
  for (uint32_t c = blockIdx.x; c < total; c += gridDim.x) {
    const int ti = find_tensor(t.chunk_start, t.n, c);
    const size_t off = size_t(c - t.chunk_start[ti]) * kMTChunk;
    const size_t rem = t.numel[ti] - off;
    const size_t n = rem < size_t(kMTChunk) ? rem : size_t(kMTChunk);
    const T* p = reinterpret_cast<const T*>(t.ptr[ti]) + off;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const size_t nvec = n / N;
      for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[N];
        Vec16<T>::unpack(ld_stream16(reinterpret_cast<const char*>(p) + i * 16), f);
#pragma unroll
        for (int k = 0; k < N; ++k) st.add(f[k]);
      }
      for (size_t i = nvec * N + threadIdx.x; i < n; i += blockDim.x) st.add(to_float<T>(p[i]));
    } else {
      for (size_t i = threadIdx.x; i < n; i += blockDim.x) st.add(to_float<T>(p[i]));
    }
  }
  publish_stats(sink, st);
}

template <typename T>
__global__ void __launch_bounds__(kStreamThreads) multi_scale_kernel(const __grid_constant__ MTTable t,
                                                                      const float* __restrict__ mult_ptr) {
  const float m = *mult_ptr;
  if (m == 1.0f) return;
  constexpr int N = Vec16<T>::N;
  const uint32_t total = t.chunk_start[t.n];
  for (uint32_t c = blockIdx.x; c < total; c += gridDim.x) {
    const int ti = find_tensor(t.chunk_start, t.n, c);
    const size_t off = size_t(c - t.chunk_start[ti]) * kMTChunk;
    const size_t rem = t.numel[ti] - off;
    const size_t n = rem < size_t(kMTChunk) ? rem : size_t(kMTChunk);
    T* p = reinterpret_cast<T*>(t.ptr[ti]) + off;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const size_t nvec = n / N;
      for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[N];
        char* q = reinterpret_cast<char*>(p) + i * 16;
        Vec16<T>::unpack(ld_stream16(q), f);
#pragma unroll
        for (int k = 0; k < N; ++k) f[k] = __fmul_rn(f[k], m);
        st_stream16(q, Vec16<T>::pack(f));
      }
      for (size_t i = nvec * N + threadIdx.x; i < n; i += blockDim.x) p[i] = from_float<T>(to_float<T>(p[i]) * m);
    } else {
      for (size_t i = threadIdx.x; i < n; i += blockDim.x) p[i] = from_float<T>(to_float<T>(p[i]) * m);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K9: fused optimizer consumption: one pass Adam/AdamW over fp32 parameters whose
// moments live in flat buffers (gradients: fusion-buffer views or loose tensors).  Arithmetic follows torch's
// foreach implementation (torch/optim/adam.py `_multi_tensor_adam`, non-capturable branch) operation by
// operation in float32: decoupled decay `p *= 1 - lr*wd` (AdamW) or `g += wd*p` (Adam), `m.lerp_(g, 1-b1)`,
// `v = v*b2 + (1-b2)*g*g`, `denom = sqrt(v)/sqrt(1-b2^t) + eps`, `p += (-lr/(1-b1^t)) * (m/denom)`.
// The clip / unscale multiplier and the found-inf flag are read from device memory, so the separate
// multiply pass over the gradients disappears and an overflowed step is skipped without a host sync.
// ------------------------------------------------------------------------------------------------
struct AdamTable {
  float* param[kMTMax];
  const float* grad[kMTMax];         // gradient of this parameter (a fusion-buffer view or a loose tensor)
  unsigned long long numel[kMTMax];
  unsigned long long off[kMTMax];    // element offset of this parameter in the flat moment buffers
  uint32_t chunk_start[kMTMax + 1];  // prefix sum of ceil(numel / kMTChunk)
  int n;
};
struct AdamHyper {
  float decay_mul;      // AdamW: 1 - lr*wd (1 = no decay)
  float l2;             // Adam: wd added to the gradient (0 = none)
  float lerp_w;         // 1 - beta1
  float beta2, one_minus_beta2;
  float neg_step_size;  // -lr / (1 - beta1^t)
  float bc2_sqrt;       // sqrt(1 - beta2^t)
  float eps;
};

// Roundings follow torch's CUDA foreach functors as nvcc compiles them (a + alpha*b contracts into one fma):
//   _foreach_add(g, p, alpha=wd)            -> fma(wd, p, g)
//   _foreach_lerp_(m, g, w)                 -> fma(w, g - m, m)             (|w| < 0.5, else g - (g-m)*(1-w))
//   _foreach_mul_(v, b2); _foreach_addcmul_ -> v = v*b2; v = fma(1-b2, g*g, v)
//   _foreach_sqrt / _foreach_div_ / _foreach_add_ -> sqrt.rn, div.rn, add
//   _foreach_addcdiv_(p, m, denom, step)    -> fma(step, m / denom, p)
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h, float mult) {
  g = __fmul_rn(g, mult);
  if (h.decay_mul != 1.0f) p = __fmul_rn(p, h.decay_mul);
  if (h.l2 != 0.0f) g = __fmaf_rn(h.l2, p, g);
  const float diff = __fsub_rn(g, m);
  m = (fabsf(h.lerp_w) < 0.5f) ? __fmaf_rn(h.lerp_w, diff, m) : __fsub_rn(g, __fmul_rn(diff, 1.0f - h.lerp_w));
  v = __fmul_rn(v, h.beta2);
  v = __fmaf_rn(h.one_minus_beta2, __fmul_rn(g, g), v);
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), h.bc2_sqrt), h.eps);
  p = __fmaf_rn(h.neg_step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(kStreamThreads) multi_adam_kernel(const __grid_constant__ AdamTable t,
                                                                    float* __restrict__ exp_avg,
                                                                    float* __restrict__ exp_avg_sq, const AdamHyper h,
                                                                    const float* __restrict__ mult_ptr,
                                                                    const float* __restrict__ found_inf_ptr) {
  if (found_inf_ptr != nullptr && *found_inf_ptr != 0.0f) return;  // overflow: the whole step is skipped
  const float mult = mult_ptr != nullptr ? *mult_ptr : 1.0f;
  const uint32_t total = t.chunk_start[t.n];
  for (uint32_t c = blockIdx.x; c < total; c += gridDim.x) {
    const int ti = find_tensor(t.chunk_start, t.n, c);
    const size_t lo = size_t(c - t.chunk_start[ti]) * kMTChunk;
    const size_t rem = t.numel[ti] - lo;
    const size_t n = rem < size_t(kMTChunk) ? rem : size_t(kMTChunk);
    float* p = t.param[ti] + lo;
    const float* g = t.grad[ti] + lo;
    float* m = exp_avg + t.off[ti] + lo;
    float* v = exp_avg_sq + t.off[ti] + lo;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v);
    size_t done = 0;
    if ((al & 15) == 0) {
      const size_t nvec = n >> 2;
      for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        float fp[4], fg[4], fm[4], fv[4];
        Vec16<float>::unpack(ld_stream16(p + i * 4), fp);
        Vec16<float>::unpack(ld_stream16(g + i * 4), fg);
        Vec16<float>::unpack(ld_stream16(m + i * 4), fm);
        Vec16<float>::unpack(ld_stream16(v + i * 4), fv);
#pragma unroll
        for (int k = 0; k < 4; ++k) adam_update(fp[k], fg[k], fm[k], fv[k], h, mult);
        st_stream16(p + i * 4, Vec16<float>::pack(fp));
        st_stream16(m + i * 4, Vec16<float>::pack(fm));
        st_stream16(v + i * 4, Vec16<float>::pack(fv));
      }
      done = nvec << 2;
    }
    for (size_t i = done + threadIdx.x; i < n; i += blockDim.x) adam_update(p[i], g[i], m[i], v[i], h, mult);
  }
}

// ------------------------------------------------------------------------------------------------
// K6: norm finalisation + flat scale / clamp.
// ------------------------------------------------------------------------------------------------
// total over records [slot0, slot0+nslots) x ranks; every rank sums the same records in the same
// order, so total_norm and the multiplier are bit-identical on all ranks.
__global__ void __launch_bounds__(256) finalize_norm_kernel(const char* control_local, int slot0, int nslots, int r0,
                                                            int nr, int is_inf_norm, float max_norm,
                                                            const float* scale_ptr, float* out) {
  const Record* recs = reinterpret_cast<const Record*>(control_local + kRecordOff);
  double v = 0.0;
  float m = 0.f;
  uint32_t nf = 0;
  const int total = nslots * nr;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const Record* p = recs + (size_t(slot0 + i / nr) * B200DP_MAX_WORLD + (r0 + i % nr));
    v += __ldcg(&p->sumsq);
    m = fmaxf(m, __ldcg(&p->maxabs));
    nf |= __ldcg(&p->nonfinite);
  }
  __shared__ double sh_v[8];
  __shared__ float sh_m[8];
  __shared__ uint32_t sh_nf[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    nf |= __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh_v[threadIdx.x >> 5] = v;
    sh_m[threadIdx.x >> 5] = m;
    sh_nf[threadIdx.x >> 5] = nf;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    v = 0.0;
    m = 0.f;
    nf = 0;
    for (int w = 0; w < 8; ++w) {
      v += sh_v[w];
      m = fmaxf(m, sh_m[w]);
      nf |= sh_nf[w];
    }
    float inv_scale = 1.0f;
    if (scale_ptr != nullptr) inv_scale = float(1.0 / double(*scale_ptr));  // torch: scale.double().reciprocal().float()
    // 2-norm: inf/NaN inputs propagate through the sum of squares by themselves
    float total_norm = is_inf_norm ? (nf ? __int_as_float(0x7f800000) : m) : float(sqrt(v));
    total_norm = total_norm * inv_scale;
    float mult = inv_scale;
    if (max_norm >= 0.f) {
      float coef = max_norm / (total_norm + 1e-6f);  // torch clip_grad_norm_: clamp(max_norm/(total+1e-6), max=1)
      coef = coef > 1.0f ? 1.0f : coef;              // NaN stays NaN, like torch.clamp
      mult = inv_scale * coef;
    }
    out[0] = total_norm;
    out[1] = mult;
    out[2] = nf ? 1.0f : 0.0f;
  }
}

template <typename T>
__global__ void __launch_bounds__(kStreamThreads) scale_flat_kernel(T* __restrict__ p, size_t count,
                                                                     const float* __restrict__ mult_ptr) {
  const float m = *mult_ptr;
  if (m == 1.0f) return;  // clip inactive: zero HBM traffic
  constexpr int N = Vec16<T>::N;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  // head elements up to 16-byte alignment
  size_t head = 0;
  const uintptr_t mis = reinterpret_cast<uintptr_t>(p) & 15;
  if (mis) head = (16 - mis) / sizeof(T);
  if (head > count) head = count;
  for (size_t i = tid; i < head; i += stride) p[i] = from_float<T>(to_float<T>(p[i]) * m);
  char* q = reinterpret_cast<char*>(p + head);
  const size_t nvec = (count - head) / N;
  for (size_t i = tid; i < nvec; i += 2 * stride) {
    const size_t i1 = i + stride;
    const bool two = i1 < nvec;
    uint4 a = ld_stream16(q + i * 16);
    uint4 b = two ? ld_stream16(q + i1 * 16) : make_uint4(0, 0, 0, 0);
    float f[N];
    Vec16<T>::unpack(a, f);
#pragma unroll
    for (int k = 0; k < N; ++k) f[k] = __fmul_rn(f[k], m);
    st_stream16(q + i * 16, Vec16<T>::pack(f));
    if (two) {
      Vec16<T>::unpack(b, f);
#pragma unroll
      for (int k = 0; k < N; ++k) f[k] = __fmul_rn(f[k], m);
      st_stream16(q + i1 * 16, Vec16<T>::pack(f));
    }
  }
  for (size_t i = head + nvec * N + tid; i < count; i += stride) p[i] = from_float<T>(to_float<T>(p[i]) * m);
}

template <typename T>
__global__ void __launch_bounds__(kStreamThreads) clamp_flat_kernel(T* __restrict__ p, size_t count, float v) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    float x = to_float<T>(p[i]);
    x = x < -v ? -v : (x > v ? v : x);  // NaN passes through, like torch.clamp_
    p[i] = from_float<T>(x);
  }
}

// ------------------------------------------------------------------------------------------------
// K6b: dtype conversion with divide, for the half-precision *wire* format (the reference's fp16/bf16
// compress hooks: buffer.to(dtype).div_(N) before the allreduce, copy back to fp32 after;
// src/accelerate/utils/dataclasses.py:206-216 -> torch default_hooks._compress_hook).
// dst = Tout( float(Tout(src)) / divisor ): cast first, then divide in the narrow type, like torch.
// ------------------------------------------------------------------------------------------------
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kStreamThreads) cast_kernel(const Tin* __restrict__ src, Tout* __restrict__ dst,
                                                               size_t count, float divisor) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  constexpr int G = 8;  // elements per thread-trip: 32 B of f32 <-> 16 B of bf16/f16
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 31) == 0;
  const size_t groups = aligned ? count / G : 0;
  for (size_t gidx = tid; gidx < groups; gidx += stride) {
    float f[G];
    if constexpr (sizeof(Tin) == 4) {
      Vec16<float>::unpack(ld_stream16(reinterpret_cast<const char*>(src) + gidx * 32), f);
      Vec16<float>::unpack(ld_stream16(reinterpret_cast<const char*>(src) + gidx * 32 + 16), f + 4);
    } else {
      Vec16<Tin>::unpack(ld_stream16(reinterpret_cast<const char*>(src) + gidx * 16), f);
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
      float v = to_float<Tout>(from_float<Tout>(f[k]));
      f[k] = divisor == 1.0f ? v : v / divisor;
    }
    if constexpr (sizeof(Tout) == 4) {
      st_stream16(reinterpret_cast<char*>(dst) + gidx * 32, Vec16<float>::pack(f));
      st_stream16(reinterpret_cast<char*>(dst) + gidx * 32 + 16, Vec16<float>::pack(f + 4));
    } else {
      st_stream16(reinterpret_cast<char*>(dst) + gidx * 16, Vec16<Tout>::pack(f));
    }
  }
  for (size_t i = groups * G + tid; i < count; i += stride) {
    float v = to_float<Tout>(from_float<Tout>(to_float<Tin>(src[i])));
    dst[i] = from_float<Tout>(divisor == 1.0f ? v : v / divisor);
  }
}

// ------------------------------------------------------------------------------------------------
// K7: fused pad copy.  Unit = U bytes (1,2,4,8,16); dst [outer, new_dim, inner_u], src [outer, old_dim, inner_u].
// ------------------------------------------------------------------------------------------------
template <typename U>
__global__ void __launch_bounds__(kStreamThreads) pad_copy_kernel(const U* __restrict__ src, U* __restrict__ dst,
                                                                   size_t outer, size_t old_dim, size_t new_dim,
                                                                   size_t inner_u, uint4 pat16, size_t shift) {
  const size_t total = outer * new_dim * inner_u;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const unsigned char* pat = reinterpret_cast<const unsigned char*>(&pat16);
  for (size_t o = size_t(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += stride) {
    const size_t i = o % inner_u;
    const size_t rest = o / inner_u;
    const size_t d = rest % new_dim;
    const size_t ou = rest / new_dim;
    U val;
    if (d >= shift && d < shift + old_dim) {
      val = src[(ou * old_dim + (d - shift)) * inner_u + i];
    } else {
      val = *reinterpret_cast<const U*>(pat + ((i * sizeof(U)) & 15));
    }
    dst[o] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// K8: tiny all-to-all of <=128 bytes per rank, result lands in mapped host memory.
// ------------------------------------------------------------------------------------------------
struct SmallPayload {
  uint32_t w[B200DP_SMALL_MAX / 4];
};
__global__ void __launch_bounds__(kThreads, 1) exchange_small_kernel(const BarrierCtx b, const SmallPayload payload,
                                                                      int nwords, uint32_t* host_out,
                                                                      volatile int* host_flag, int flag_val) {
  const uint32_t base = barrier_seq_begin(b);
  const uint32_t parity = base & 1u;  // exactly one barrier per launch on this channel -> alternates
  const int t = threadIdx.x;
  const int n = b.world * nwords;
  if (t < n) {
    const int peer = t / nwords, w = t % nwords;
    uint32_t* slot = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(b.pads.p[peer]) + kSmallOff) +
                     ((size_t(parity) * B200DP_MAX_WORLD + b.rank) * (B200DP_SMALL_MAX / 4) + w);
    *slot = payload.w[w];
  }
  cross_rank_barrier(b, base + 1);
  if (t < n) {
    const int src = t / nwords, w = t % nwords;
    const uint32_t* slot = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(b.pads.p[b.rank]) + kSmallOff) +
                           ((size_t(parity) * B200DP_MAX_WORLD + src) * (B200DP_SMALL_MAX / 4) + w);
    host_out[src * nwords + w] = __ldcg(slot);
  }
  __threadfence_system();
  __syncthreads();
  if (t == 0) {
    __threadfence_system();
    *host_flag = flag_val;
  }
  barrier_seq_end(b, base, 1);
}

}  // namespace b200dp
