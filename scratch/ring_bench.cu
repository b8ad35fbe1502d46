// Microbenchmark: gather 2 KB rows by index and reduce them per 26-edge "row" --
//   (A) per-lane LDG.128 with U edges in flight (what the layer kernels do today)
//   (B) per-warp shared-memory ring filled by cp.async.bulk (1-D TMA) + mbarrier, LDS.128 consumers.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo scratch/ring_bench.cu -o scratch/ring_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// work emulation: per edge 16 FMA dot + 5-shuffle reduce + EXTRA dependent ALU ops
template <int EXTRA>
__device__ __forceinline__ float edge_math(const float (&h)[16], const float (&dh)[16], float (&acc)[16], int lane) {
  float g = 0.f;
#pragma unroll
  for (int r = 0; r < 16; ++r) g = fmaf(dh[r], h[r], g);
#pragma unroll
  for (int o = 16; o > 1; o >>= 1) g += __shfl_xor_sync(FULL, g, o);
  float s = g;
#pragma unroll
  for (int k = 0; k < EXTRA; ++k) s = fmaf(s, 0.999f, 0.001f);
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = fmaf(s, h[r], acc[r]);
  return s;
}

template <int U, int EXTRA>
__global__ void __launch_bounds__(256, 2) k_ldg(const float* __restrict__ V, const int* __restrict__ nbr, int64_t e_total,
                                                int row_len, float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * 8, w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t rows = e_total / row_len;
  float dh[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) dh[r] = 0.01f * (r + lane);
  for (int64_t row = w; row < rows; row += nw) {
    float acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.f;
    const int64_t eb = row * row_len;
    const int myj = lane < row_len ? __ldg(nbr + eb + lane) : 0;
    for (int t = 0; t < row_len; t += U) {
      float h[U][16];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = __shfl_sync(FULL, myj, min(t + u, row_len - 1));
        const float* vp = V + j * 512;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(vp + (k * 32 + lane) * 4));
          h[u][4 * k] = v.x; h[u][4 * k + 1] = v.y; h[u][4 * k + 2] = v.z; h[u][4 * k + 3] = v.w;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) edge_math<EXTRA>(h[u], dh, acc, lane);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(out + row * 512 + (k * 32 + lane) * 4) = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
  }
}

// (B) each warp owns a contiguous edge range and a private ring of NS x 2 KB slots
template <int NS, int EXTRA>
__global__ void __launch_bounds__(256, 2) k_ring(const float* __restrict__ V, const int* __restrict__ nbr, int64_t e_total,
                                                 int row_len, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint8_t* ring = smem + (size_t)wid * NS * 2048;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)8 * NS * 2048) + wid * NS;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(bars + s), 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t nw = (int64_t)gridDim.x * 8, w = (int64_t)blockIdx.x * 8 + wid;
  const int64_t rows = e_total / row_len;
  const int64_t r0 = rows * w / nw, r1 = rows * (w + 1) / nw;
  const int64_t e0 = r0 * row_len, e1 = r1 * row_len;
  float dh[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) dh[r] = 0.01f * (r + lane);
  // producer cursor
  int64_t pe = e0;
  auto issue = [&](int64_t e) {   // all lanes call; lane 0 issues
    if (lane == 0) {
      const int64_t j = __ldg(nbr + e);
      const int s = (int)((e - e0) % NS);
      const uint32_t bar = smem_u32(bars + s);
      mbar_expect_tx(bar, 2048);
      bulk_g2s(smem_u32(ring + s * 2048), V + j * 512, 2048, bar);
    }
  };
  for (; pe < e1 && pe < e0 + NS; ++pe) issue(pe);
  int64_t e = e0;
  for (int64_t row = r0; row < r1; ++row) {
    float acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.f;
    for (int t = 0; t < row_len; ++t, ++e) {
      const int64_t k = e - e0;
      const int s = (int)(k % NS);
      const uint32_t parity = (uint32_t)((k / NS) & 1);
      mbar_wait(smem_u32(bars + s), parity);
      float h[16];
      const float* sp = reinterpret_cast<const float*>(ring + s * 2048);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 v = *reinterpret_cast<const float4*>(sp + (kk * 32 + lane) * 4);
        h[4 * kk] = v.x; h[4 * kk + 1] = v.y; h[4 * kk + 2] = v.z; h[4 * kk + 3] = v.w;
      }
      __syncwarp();
      if (pe < e1) { issue(pe); ++pe; }
      edge_math<EXTRA>(h, dh, acc, lane);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(out + row * 512 + (k * 32 + lane) * 4) = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
  }
}

template <class F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const int64_t n = argc > 2 ? atoll(argv[2]) : 2400000;
  const int64_t e = (argc > 3 ? atoll(argv[3]) : 62400000) / 26 * 26;
  const int row_len = 26;
  float* V; int* nbr; float* out;
  CK(cudaMalloc(&V, n * 512 * 4));
  CK(cudaMemset(V, 0, n * 512 * 4));
  CK(cudaMalloc(&nbr, e * 4));
  CK(cudaMalloc(&out, (e / row_len) * 512 * 4));
  std::vector<int> h(e);
  std::mt19937_64 rng(1);
  const bool skew = argc > 1 && atoi(argv[1]) == 1;
  for (int64_t i = 0; i < e; ++i) {
    double u = (rng() >> 11) * (1.0 / 9007199254740992.0);
    h[i] = skew ? (int)(n * u * u * u) : (int)(n * u);     // skew: cubic -> hot head
  }
  CK(cudaMemcpy(nbr, h.data(), e * 4, cudaMemcpyHostToDevice));
  const double gb = (double)e * 2048 / 1e9 + (double)(e / row_len) * 2048 / 1e9;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("gather %.1f GB per launch (%s indices), %d SMs\n", gb, skew ? "skewed" : "uniform", sms);
#define RUN_LDG(U, X) { float ms = time_ms([&] { k_ldg<U, X><<<sms * 2, 256>>>(V, nbr, e, row_len, out); }); \
    printf("ldg  U=%d extra=%3d : %7.2f ms  %7.0f GB/s\n", U, X, ms, gb / ms * 1e3); }
#define RUN_RING(NS, X, OCC) { size_t sm = (size_t)8 * NS * 2048 + 8 * NS * 8; \
    CK(cudaFuncSetAttribute(k_ring<NS, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    float ms = time_ms([&] { k_ring<NS, X><<<sms * OCC, 256, sm>>>(V, nbr, e, row_len, out); }); \
    printf("ring NS=%d extra=%3d occ=%d : %7.2f ms  %7.0f GB/s\n", NS, X, OCC, ms, gb / ms * 1e3); }
  RUN_LDG(1, 0) RUN_LDG(2, 0) RUN_LDG(4, 0)
  RUN_LDG(2, 60) RUN_LDG(2, 150)
  RUN_RING(2, 0, 2) RUN_RING(4, 0, 2) RUN_RING(6, 0, 2) RUN_RING(4, 0, 1) RUN_RING(8, 0, 1) RUN_RING(12, 0, 1)
  RUN_RING(4, 60, 2) RUN_RING(6, 60, 2) RUN_RING(4, 150, 2) RUN_RING(6, 150, 2) RUN_RING(12, 150, 1)
  CK(cudaDeviceSynchronize());
  return 0;
}
