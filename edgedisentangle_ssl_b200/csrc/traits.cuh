// Lane <-> element mappings shared by the layer and pair kernels (see disga.cu).
#pragma once
#include "edis_common.cuh"

namespace edis {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------ lane <-> element traits
// VecT: D == 4*LPC.  A warp covers CPW = KV*32/LPC channels; register block k of a lane holds
// the float4 at float offset (k*32 + lane)*4 from the group's first column.
template <int KV_, int LPC_>
struct VecT {
  static constexpr int KV = KV_, LPC = LPC_, LPCV = LPC_;
  static constexpr int NCH = KV;      // channel slots per lane
  static constexpr int RPC = 4;       // registers per channel slot
  static constexpr int R = 4 * KV;
  static constexpr int CPK = 32 / LPC;
  static constexpr int CPW = KV * CPK;
  static constexpr bool kVec = true;
  static constexpr bool kRing = LPC_ == 16;   // D == 64 layouts: the bulk-copy (ring) kernels exist for these
  __device__ static __forceinline__ int ch(int k, int lane) { return k * CPK + lane / LPC; }
  __device__ static __forceinline__ bool writer(int lane) { return (lane % LPC) == 0; }
  // value of channel `cc` (0..CPW) held by that channel's lanes -> every lane of the warp
  __device__ static __forceinline__ float bcast(const float (&v)[NCH], int cc) {
    float r = 0.0f;
#pragma unroll
    for (int k = 0; k < KV; ++k)
      if (k == cc / CPK) r = v[k];
    return __shfl_sync(0xffffffffu, r, (cc % CPK) * LPC);
  }
  __device__ static __forceinline__ float reduce(float v) {
#pragma unroll
    for (int o = LPC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }
  // "Owned slot" scheme: after reduce_own every lane holds the channel sum of ONE of its KV
  // slots (slot own(lane)), so the per-channel scalar math (sigmoid, exp, dropout hash, softmax
  // backward) is issued once per warp instead of KV times.  Transposed butterfly: KV-1 + log2
  // (LPC/KV) ... shuffles instead of KV*log2(LPC).
  static constexpr int SPAN = LPC / KV;   // lanes that end up owning the same slot
  __device__ static __forceinline__ int own(int lane) { return (lane % LPC) / SPAN; }
  __device__ static __forceinline__ int own_ch(int lane) { return own(lane) * CPK + lane / LPC; }
  __device__ static __forceinline__ bool own_writer(int lane) { return (lane % SPAN) == 0; }
  __device__ static __forceinline__ float reduce_own(float (&v)[NCH], int lane) {
    int off = LPC / 2;
#pragma unroll
    for (int n = KV; n > 1; n >>= 1) {
      const bool hi = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < n / 2; ++i) {
        const float keep = hi ? v[i + n / 2] : v[i];
        const float send = hi ? v[i] : v[i + n / 2];
        v[i] = keep + __shfl_xor_sync(FULL, send, off);
      }
      off >>= 1;
    }
    float r = v[0];
#pragma unroll
    for (int o = SPAN / 2; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    return r;
  }
  // value owned for slot k of this lane's channel group -> this lane
  __device__ static __forceinline__ float from_owner(float mine, int k, int lane) {
    return __shfl_sync(FULL, mine, (lane / LPC) * LPC + k * SPAN);
  }
  // value owned for warp channel cc (0..CPW) -> every lane
  __device__ static __forceinline__ float from_channel(float mine, int cc) {
    return __shfl_sync(FULL, mine, (cc % CPK) * LPC + (cc / CPK) * SPAN);
  }
  __device__ static __forceinline__ void load(float (&r)[R], const float* base, int lane, int) {
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const float4 v = ldg4(base + (k * 32 + lane) * 4);
      r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
    }
  }
  template <bool HOT>
  __device__ static __forceinline__ void load_pol(float (&r)[R], const float* base, int lane, int) {
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const float4 v = ldg4_pol<HOT>(base + (k * 32 + lane) * 4);
      r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
    }
  }
  __device__ static __forceinline__ void store(float* base, const float (&r)[R], int lane, int) {
#pragma unroll
    for (int k = 0; k < KV; ++k)
      *reinterpret_cast<float4*>(base + (k * 32 + lane) * 4) =
          make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
  }
  __device__ static __forceinline__ void atomic_add(float* base, const float (&r)[R], int lane, int) {
#pragma unroll
    for (int k = 0; k < KV; ++k)
      red_add4(base + (k * 32 + lane) * 4, r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
  }
};

// ScaT: any D <= 32*ND, one channel per warp, lane-strided scalar accesses.
template <int ND_>
struct ScaT {
  static constexpr int NCH = 1, RPC = ND_, R = ND_, CPW = 1, KV = 0, LPCV = 1;
  static constexpr bool kVec = false;
  static constexpr bool kRing = false;
  __device__ static __forceinline__ int ch(int, int) { return 0; }
  __device__ static __forceinline__ bool writer(int lane) { return lane == 0; }
  __device__ static __forceinline__ float bcast(const float (&v)[NCH], int) { return v[0]; }
  __device__ static __forceinline__ int own(int) { return 0; }
  __device__ static __forceinline__ int own_ch(int) { return 0; }
  __device__ static __forceinline__ bool own_writer(int lane) { return lane == 0; }
  __device__ static __forceinline__ float reduce_own(float (&v)[NCH], int) { return reduce(v[0]); }
  __device__ static __forceinline__ float from_owner(float mine, int, int) { return mine; }
  __device__ static __forceinline__ float from_channel(float mine, int) { return mine; }
  __device__ static __forceinline__ float reduce(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }
  __device__ static __forceinline__ void load(float (&r)[R], const float* base, int lane, int D) {
#pragma unroll
    for (int k = 0; k < ND_; ++k) r[k] = (k * 32 + lane < D) ? __ldg(base + k * 32 + lane) : 0.0f;
  }
  template <bool HOT>
  __device__ static __forceinline__ void load_pol(float (&r)[R], const float* base, int lane, int D) {
    load(r, base, lane, D);
  }
  __device__ static __forceinline__ void store(float* base, const float (&r)[R], int lane, int D) {
#pragma unroll
    for (int k = 0; k < ND_; ++k)
      if (k * 32 + lane < D) base[k * 32 + lane] = r[k];
  }
  __device__ static __forceinline__ void atomic_add(float* base, const float (&r)[R], int lane, int D) {
#pragma unroll
    for (int k = 0; k < ND_; ++k)
      if (k * 32 + lane < D) atomicAdd(base + k * 32 + lane, r[k]);
  }
};

template <class T>
__device__ __forceinline__ void zero(float (&r)[T::R]) {
#pragma unroll
  for (int i = 0; i < T::R; ++i) r[i] = 0.0f;
}


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

int env_int(const char* name, int dflt);

}  // namespace edis
