// device_utils.cuh — device-side building blocks shared by the b200dp kernels (sm_100a only).
//
//  * system-scope release/acquire flag barrier between the same-index CTAs of all ranks
//  * 16-byte vector loads/stores with L1 bypass for streaming data
//  * multimem (NVLS) load-reduce / store wrappers
//  * block reductions by warp shuffle
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200dp.h"

namespace b200dp {

struct Peers {
  void* p[B200DP_MAX_WORLD];
};

// Everything a kernel needs to run cross-rank barriers on one channel.  Passed by value.
struct BarrierCtx {
  Peers pads;                    // control-segment base of every rank (pads.p[rank] is local)
  uint32_t* ctr;                 // local per-CTA sequence counters of this channel [MAX_BLOCKS]
  int* status;                   // mapped host status word (0 = ok)
  unsigned long long timeout_ns; // flag-wait timeout
  int rank, world, channel;
};

// ---- control segment layout (identical on every rank) ----
constexpr size_t kPadBytes = size_t(B200DP_NUM_CHANNELS) * B200DP_MAX_BLOCKS * B200DP_MAX_WORLD * 4;
struct __align__(16) Record {  // stats of one bucket slice as computed by one rank
  double sumsq;
  float maxabs;
  uint32_t nonfinite;
};
constexpr size_t kRecordOff = (kPadBytes + 255) / 256 * 256;
constexpr size_t kRecordBytes = size_t(B200DP_MAX_SLOTS) * B200DP_MAX_WORLD * sizeof(Record);
constexpr size_t kSmallOff = kRecordOff + kRecordBytes;
constexpr size_t kSmallBytes = size_t(2) * B200DP_MAX_WORLD * B200DP_SMALL_MAX;
constexpr size_t kControlBytes = (kSmallOff + kSmallBytes + 4095) / 4096 * 4096;

__device__ __forceinline__ size_t pad_slot(int channel, int block, int src) {
  return (size_t(channel) * B200DP_MAX_BLOCKS + block) * B200DP_MAX_WORLD + src;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Barrier between CTA `blockIdx.x` of every rank.  `val` is this barrier's sequence value.
// All threads of the CTA must call it.  Writes made by any thread of this CTA before the call are
// visible to every thread of the peer CTAs after they return (release/acquire at system scope,
// made CTA-wide by the two bar.sync).
__device__ __forceinline__ void cross_rank_barrier(const BarrierCtx& b, uint32_t val) {
  __syncthreads();
  if (threadIdx.x < b.world) {
    uint32_t* remote = reinterpret_cast<uint32_t*>(b.pads.p[threadIdx.x]) + pad_slot(b.channel, blockIdx.x, b.rank);
    st_release_sys(remote, val);  // release at system scope: cumulative over the CTA's writes ordered by the bar.sync above
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(b.pads.p[b.rank]) + pad_slot(b.channel, blockIdx.x, threadIdx.x);
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while (int32_t(ld_acquire_sys(mine) - val) < 0) {
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

// per-CTA sequence base: read once at kernel start, advanced by `used` at kernel end.
__device__ __forceinline__ uint32_t barrier_seq_begin(const BarrierCtx& b) {
  __shared__ uint32_t s_base;
  if (threadIdx.x == 0) s_base = b.ctr[blockIdx.x];
  __syncthreads();
  return s_base;
}
__device__ __forceinline__ void barrier_seq_end(const BarrierCtx& b, uint32_t base, uint32_t used) {
  if (threadIdx.x == 0) b.ctr[blockIdx.x] = base + used;
}

// ---- 16-byte streaming accesses ----
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ---- multimem (NVLS): the switch sums the N copies / replicates the store ----
template <typename T>
struct MultimemOps;
template <>
struct MultimemOps<float> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
  }
};
template <>
struct MultimemOps<__nv_bfloat16> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
  }
};
template <>
struct MultimemOps<__half> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
  }
};
__device__ __forceinline__ void multimem_st16(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// ---- element <-> float conversion over one 16-byte vector ----
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Vec16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// ---- per-thread running statistics of reduced values ----
struct Stats {
  float sumsq = 0.f;
  float maxabs = 0.f;
  uint32_t nonfinite = 0;
  __device__ __forceinline__ void add(float x) {
    sumsq = fmaf(x, x, sumsq);
    float a = fabsf(x);
    maxabs = fmaxf(maxabs, a);  // fmaxf drops NaN; nonfinite catches it
    nonfinite |= (a <= 3.402823466e+38f) ? 0u : 1u;  // false for inf and NaN
  }
};

// CTA reduction of Stats -> thread 0 gets the result (sumsq in double from per-warp floats).
__device__ __forceinline__ void block_reduce_stats(const Stats& s, double* out_sumsq, float* out_max, uint32_t* out_nf) {
  __shared__ double sh_sum[32];
  __shared__ float sh_max[32];
  __shared__ uint32_t sh_nf[32];
  double v = double(s.sumsq);
  float m = s.maxabs;
  uint32_t nf = s.nonfinite;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    nf |= __shfl_xor_sync(0xffffffffu, nf, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh_sum[warp] = v;
    sh_max[warp] = m;
    sh_nf[warp] = nf;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? sh_sum[lane] : 0.0;
    m = lane < nw ? sh_max[lane] : 0.f;
    nf = lane < nw ? sh_nf[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v += __shfl_xor_sync(0xffffffffu, v, o);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      nf |= __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if (lane == 0) {
      *out_sumsq = v;
      *out_max = m;
      *out_nf = nf;
    }
  }
}

}  // namespace b200dp
