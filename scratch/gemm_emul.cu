// FP32 GEMM on B200: native SIMT fp32 vs cuBLAS 12.9 BF16x9 emulation (tensor cores), shapes of the node projection
// and its backward.  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scratch/gemm_emul.cu -lcublas -o gemm_emul
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)
#define CB(x) do { cublasStatus_t st__ = (x); if (st__ != CUBLAS_STATUS_SUCCESS) { printf("%s: %s\n", #x, cublasGetStatusName(st__)); } } while (0)

__global__ void fill(float* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u ^ seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    p[i] = ((x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.0f;
  }
}
// row-major C[M,N] = op(A) op(B): computed as column-major C^T = op(B)^T op(A)^T
static cublasStatus_t gemm_rm(cublasHandle_t h, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                              float* C, int ldc, cublasComputeType_t ct) {
  const float one = 1.f, zero = 0.f;
  return cublasGemmEx(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &one, B, CUDA_R_32F, ldb, A, CUDA_R_32F, lda,
                      &zero, C, CUDA_R_32F, ldc, ct, CUBLAS_GEMM_DEFAULT);
}
int main() {
  cublasHandle_t h;
  CB(cublasCreate(&h));
  int ver; cublasGetVersion(h, &ver); printf("cublas %d\n", ver);
  size_t wsb = (size_t)4 << 30;
  void* ws; CK(cudaMalloc(&ws, wsb));
  CB(cublasSetWorkspace(h, ws, wsb));
  struct Shape { const char* name; bool ta, tb; int M, N, K; };
  // A[M,K] (or A^T stored [K,M] when ta), B[K,N]
  const Shape shapes[] = {{"proj fwd   x[S,100] W[100,1024]", false, false, 2400000, 1024, 100},
                          {"proj fwd L2 x[S,64] W[64,1024]", false, false, 2400000, 1024, 64},
                          {"dX   g[S,1024] W^T[1024,64]", false, true, 2400000, 64, 1024},
                          {"dW   x^T[100,S] g[S,1024]", true, false, 100, 1024, 2400000},
                          {"fuser h[S,512] W^T[512,64]", false, true, 2400000, 64, 512},
                          {"fuser dW g^T[64,S] h[S,512]", true, false, 64, 512, 2400000}};
  for (const Shape& s : shapes) {
    const size_t na = (size_t)s.M * s.K, nb = (size_t)s.K * s.N, nc = (size_t)s.M * s.N;
    float *A, *B, *C, *C2;
    CK(cudaMalloc(&A, na * 4)); CK(cudaMalloc(&B, nb * 4)); CK(cudaMalloc(&C, nc * 4)); CK(cudaMalloc(&C2, nc * 4));
    fill<<<1024, 256>>>(A, na, 1); fill<<<1024, 256>>>(B, nb, 2);
    const int lda = s.ta ? s.M : s.K, ldb = s.tb ? s.K : s.N;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const cublasComputeType_t cts[] = {CUBLAS_COMPUTE_32F, CUBLAS_COMPUTE_32F_EMULATED_16BFX9, CUBLAS_COMPUTE_32F_FAST_TF32};
    const char* names[] = {"fp32 native", "bf16x9 emu ", "tf32 1-pass "};
    for (int v = 0; v < 3; ++v) {
      float* out = v == 0 ? C : C2;
      cublasStatus_t st = gemm_rm(h, s.ta, s.tb, s.M, s.N, s.K, A, lda, B, ldb, out, s.N, cts[v]);
      CK(cudaDeviceSynchronize());
      if (st != CUBLAS_STATUS_SUCCESS) { printf("  %-34s %s: %s\n", s.name, names[v], cublasGetStatusName(st)); continue; }
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) gemm_rm(h, s.ta, s.tb, s.M, s.N, s.K, A, lda, B, ldb, out, s.N, cts[v]);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      double err = 0, mx = 0;
      if (v > 0) {
        const size_t cnt = nc < 4000000 ? nc : 4000000;
        std::vector<float> r(cnt), g(cnt);
        CK(cudaMemcpy(r.data(), C, cnt * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(g.data(), C2, cnt * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt; ++i) { err = fmax(err, fabs((double)r[i] - g[i])); mx = fmax(mx, fabs((double)r[i])); }
      }
      printf("  %-34s %s: %7.2f ms  %6.1f TFLOP/s  max|diff vs native|/max = %.2e\n", s.name, names[v], ms,
             2.0 * s.M * s.N * s.K / ms / 1e9, mx > 0 ? err / mx : 0.0);
      if (s.ta && v < 2) {
        // long-K weight gradient: 16 output entries against a float64 evaluation on the host
        std::vector<float> ha(na), hb((size_t)s.K * 16);
        CK(cudaMemcpy(ha.data(), A, na * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy2D(hb.data(), 16 * 4, B, (size_t)s.N * 4, 16 * 4, s.K, cudaMemcpyDeviceToHost));
        std::vector<float> hc(16);
        CK(cudaMemcpy(hc.data(), out, 16 * 4, cudaMemcpyDeviceToHost));      // row 0 of C, first 16 columns
        double worst = 0, scale = 0;
        for (int j = 0; j < 16; ++j) {
          double acc = 0, sa = 0;
          for (int k = 0; k < s.K; ++k) { acc += (double)ha[(size_t)k * s.M] * hb[(size_t)k * 16 + j]; sa += fabs((double)ha[(size_t)k * s.M] * hb[(size_t)k * 16 + j]); }
          worst = fmax(worst, fabs(acc - hc[j])); scale = fmax(scale, sa);
        }
        printf("      vs float64 (16 entries of row 0): max|err| / sum|terms| = %.2e\n", worst / scale);
      }
    }
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(C2);
  }
  return 0;
}
