// 3xTF32 (fp32-accurate) GEMMs on the tensor cores via CUTLASS' OpMultiplyAddFastF32 (mma.sync TF32, operands split
// big/small in registers): the three shapes of the node projection and its backward.  Timing + accuracy vs fp64.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
#include "cutlass/cutlass.h"
#include "cutlass/gemm/device/gemm_universal.h"
#include "cutlass/epilogue/thread/linear_combination.h"

#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)

template <class LA, class LB, class TB, class WP, int ST>
using Gemm3x = cutlass::gemm::device::GemmUniversal<
    float, LA, float, LB, float, cutlass::layout::RowMajor, float, cutlass::arch::OpClassTensorOp, cutlass::arch::Sm80,
    TB, WP, cutlass::gemm::GemmShape<16, 8, 8>,
    cutlass::epilogue::thread::LinearCombination<float, 4, float, float>,
    cutlass::gemm::threadblock::GemmIdentityThreadblockSwizzle<>, ST, 4, 4, cutlass::arch::OpMultiplyAddFastF32>;

using RM = cutlass::layout::RowMajor;
using CM = cutlass::layout::ColumnMajor;
using GemmNN = Gemm3x<RM, RM, cutlass::gemm::GemmShape<128, 128, 16>, cutlass::gemm::GemmShape<64, 64, 16>, 3>;   // x W
using GemmNT = Gemm3x<RM, CM, cutlass::gemm::GemmShape<128, 64, 16>, cutlass::gemm::GemmShape<64, 32, 16>, 4>;    // g W^T
using GemmTN = Gemm3x<CM, RM, cutlass::gemm::GemmShape<128, 128, 16>, cutlass::gemm::GemmShape<64, 64, 16>, 3>;   // x^T g

__global__ void fill(float* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u ^ seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    p[i] = ((x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.0f;
  }
}

template <class G>
float run(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int slices, void* ws, size_t wsb) {
  typename G::Arguments args(slices > 1 ? cutlass::gemm::GemmUniversalMode::kGemm : cutlass::gemm::GemmUniversalMode::kGemm,
                             {M, N, K}, slices, {1.0f, 0.0f}, A, B, C, C, 0, 0, 0, 0, lda, ldb, N, N);
  G op;
  if (op.can_implement(args) != cutlass::Status::kSuccess) { printf("cannot implement\n"); return -1; }
  if (G::get_workspace_size(args) > wsb) { printf("workspace %zu > %zu\n", G::get_workspace_size(args), wsb); return -1; }
  if (op.initialize(args, ws) != cutlass::Status::kSuccess) { printf("init failed\n"); return -1; }
  op();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) { op.initialize(args, ws); op(); }
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main(int argc, char** argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 2400000;
  size_t wsb = (size_t)1 << 30; void* ws; CK(cudaMalloc(&ws, wsb));
  float *X, *W, *G, *O;
  CK(cudaMalloc(&X, (size_t)S * 128 * 4)); CK(cudaMalloc(&W, (size_t)128 * 1536 * 4));
  CK(cudaMalloc(&G, (size_t)S * 1536 * 4)); CK(cudaMalloc(&O, (size_t)S * 1536 * 4));
  fill<<<2048, 256>>>(X, (size_t)S * 128, 1); fill<<<64, 256>>>(W, (size_t)128 * 1536, 2); fill<<<2048, 256>>>(G, (size_t)S * 1536, 3);
  CK(cudaDeviceSynchronize());
  std::vector<float> hx(128 * 64), hw(128 * 1536), hg(64 * 1536), ho(64 * 1536);
  // (A) O[S,1024] = X[S,100] W[100,1024]
  for (int K : {100, 64}) {
    float ms = run<GemmNN>(S, 1024, K, X, K, W, 1024, O, 1, ws, wsb);
    CK(cudaMemcpy(hx.data(), X, (size_t)64 * K * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hw.data(), W, (size_t)K * 1024 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ho.data(), O, (size_t)64 * 1024 * 4, cudaMemcpyDeviceToHost));
    double err = 0, sc = 0;
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 1024; j += 37) { double a = 0, s = 0; for (int k = 0; k < K; ++k) { a += (double)hx[i * K + k] * hw[k * 1024 + j]; s += fabs((double)hx[i * K + k] * hw[k * 1024 + j]); } err = fmax(err, fabs(a - ho[i * 1024 + j])); sc = fmax(sc, s); }
    printf("NN  x[S,%d] W[%d,1024]       : %6.2f ms  %6.1f TFLOP/s   err/sum|terms| %.2e\n", K, K, ms, 2.0 * S * 1024 * K / ms / 1e9, err / sc);
  }
  // (B) O[S,F] = G[S,1024] W^T   (W stored [F,1024])
  for (int F : {64, 100}) {
    float ms = run<GemmNT>(S, F, 1024, G, 1024, W, 1024, O, 1, ws, wsb);
    CK(cudaMemcpy(hg.data(), G, (size_t)64 * 1024 * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hw.data(), W, (size_t)F * 1024 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ho.data(), O, (size_t)64 * F * 4, cudaMemcpyDeviceToHost));
    double err = 0, sc = 0;
    for (int i = 0; i < 64; ++i) for (int j = 0; j < F; ++j) { double a = 0, s = 0; for (int k = 0; k < 1024; ++k) { a += (double)hg[i * 1024 + k] * hw[j * 1024 + k]; s += fabs((double)hg[i * 1024 + k] * hw[j * 1024 + k]); } err = fmax(err, fabs(a - ho[i * F + j])); sc = fmax(sc, s); }
    printf("NT  g[S,1024] W^T[1024,%d]    : %6.2f ms  %6.1f TFLOP/s   err/sum|terms| %.2e\n", F, ms, 2.0 * S * 1024 * F / ms / 1e9, err / sc);
  }
  // (C) O[F,1024] = X^T[F,S] G[S,1024], split-K
  for (int F : {100, 64}) for (int slices : {32, 64, 128}) {
    float ms = run<GemmTN>(F, 1024, S, X, F, G, 1024, O, slices, ws, wsb);
    if (ms < 0) continue;
    // float64 check of 8 entries of row 0
    std::vector<float> col((size_t)S), gcol((size_t)S * 8);
    CK(cudaMemcpy2D(col.data(), 4, X, (size_t)F * 4, 4, S, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy2D(gcol.data(), 8 * 4, G, (size_t)1024 * 4, 8 * 4, S, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ho.data(), O, 8 * 4, cudaMemcpyDeviceToHost));
    double err = 0, sc = 0;
    for (int j = 0; j < 8; ++j) { double a = 0, s = 0; for (int k = 0; k < S; ++k) { a += (double)col[k] * gcol[(size_t)k * 8 + j]; s += fabs((double)col[k] * gcol[(size_t)k * 8 + j]); } err = fmax(err, fabs(a - ho[j])); sc = fmax(sc, s); }
    printf("TN  x^T[%d,S] g[S,1024] k=%3d : %6.2f ms  %6.1f TFLOP/s   err/sum|terms| %.2e\n", F, slices, ms, 2.0 * S * 1024 * F / ms / 1e9, err / sc);
  }
  return 0;
}
