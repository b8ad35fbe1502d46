// SSL-side kernels for sm_100a: pair scoring on arbitrary (i, j) lists, the fused
// channel-sum -> sigmoid -> class-balanced MSE reduction, the constant-label NLL tail of
// DifHead, and the stand-alone COO sp_softmax / sp_matmul drop-ins.
// Replaces /root/reference/layers.py:355-389 (edge_auxs), pretrainer.py:613-627 / 730-737,
// utils.py:287-298, pretrainer.py:825-832 and utils.py:192-207.
#include "edis_common.cuh"
#include "traits.cuh"
#include <algorithm>

namespace edis {

constexpr int kPairBlock = 32;  // consecutive pairs handled by one warp (lists are sorted by i)

struct PairArgs {
  int64_t m, n_units;
  const int64_t *pi, *pj;
  int c_lo, Cs, G, C, D;
  const float *P, *Q, *a;
  int64_t ldp, ldq;
  float* out;
  const float* g_out;
  float *gP, *gQ, *ga;
  // att 3 sign record (see edis_common.cuh): written by the forward, read by k_pair_bwd_sign
  unsigned char* sign;
  // k_pair_bwd_sign: one pass per side.  key = pi (row pass, perm == NULL) or pj (column pass,
  // perm = pair ids sorted by j); X / ldx = the side's operand rows (P or Q), gX its gradient
  const int64_t* key;
  const int32_t *perm, *col_perm;
  const float* X;
  int64_t ldx;
  float* gX;
};

template <class T, int ATT>
__global__ void __launch_bounds__(256) k_pair_fwd(const PairArgs A) {
  constexpr int R = T::R, NCH = T::NCH, RPC = T::RPC;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  int64_t unit = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (; unit < A.n_units; unit += nwarps) {
    const int64_t blk = unit / A.G;
    const int grp = static_cast<int>(unit - blk * A.G);
    const int c0 = A.c_lo + grp * T::CPW;
    const int off = c0 * A.D;
    int cidx[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) cidx[k] = c0 + T::ch(k, lane);
    constexpr int SBPL = (R + 7) / 8;
    float ar[R];
    if (ATT == 3) T::load(ar, A.a + off, lane, A.D);
    const int64_t base = blk * kPairBlock;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPairBlock), A.m - base));
    const int64_t myi = lane < cnt ? A.pi[base + lane] : 0;
    const int64_t myj = lane < cnt ? A.pj[base + lane] : 0;
    int64_t cur_i = -1;
    float pr[R], sd[NCH];
    for (int t = 0; t < cnt; ++t) {
      const int64_t i = __shfl_sync(FULL, myi, t);
      const int64_t j = __shfl_sync(FULL, myj, t);
      float e[NCH];
      unsigned mask = 0u;
      if (ATT == 1) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) e[k] = __ldg(A.P + i * A.ldp + cidx[k]) + __ldg(A.Q + j * A.ldq + cidx[k]);
        (void)sd;
      } else {
        if (i != cur_i) {
          T::load(pr, A.P + i * A.ldp + off, lane, A.D);
          cur_i = i;
        }
        float q[R], part[NCH];
        T::load(q, A.Q + j * A.ldq + off, lane, A.D);
#pragma unroll
        for (int k = 0; k < NCH; ++k) part[k] = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (ATT == 3) {
            const float w = -pr[r] - q[r];                 // -(P_i + Q_j): its sign bit <=> z > 0
            mask = sign_push(mask, w);
            part[r / RPC] = fmaf(ar[r], lrelu01(-w), part[r / RPC]);
          } else {
            part[r / RPC] = fmaf(pr[r], q[r], part[r / RPC]);
          }
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) e[k] = T::reduce(part[k]);
        if (ATT == 3 && A.sign) {
          const int64_t so = (((base + t) * A.G + grp) * 32 + lane) * SBPL;
          if (SBPL == 1) st_stream(A.sign + so, static_cast<unsigned char>(mask));
          else st_stream(reinterpret_cast<unsigned short*>(A.sign + so), static_cast<unsigned short>(mask));
        }
      }
      if (T::writer(lane)) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) A.out[(base + t) * A.Cs + (cidx[k] - A.c_lo)] = e[k];
      }
    }
  }
}

// (A ring / bulk-copy variant of this kernel -- contiguous pair runs per warp, one cp.async.bulk per pair for the
// channel group's 1 KB slice of Q_j -- was built and measured in round 2: parity-green, but the SupEdge step on
// config A went from 707 to 724 ms.  With two channel groups per pair the copies are only 1 KB each and the
// per-copy issue cost is paid twice; removed again.  profiles/README.md.)
// Backward: gP_i and ga accumulate in registers over runs of equal i (lists are row-sorted);
// gQ_j (and gP_i at run ends) go out as 128-bit vector reductions.
template <class T, int ATT>
__global__ void __launch_bounds__(256) k_pair_bwd(const PairArgs A) {
  constexpr int R = T::R, NCH = T::NCH, RPC = T::RPC;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  int64_t unit = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int CD = A.C * A.D;
  float da[R];
  zero<T>(da);
  int da_grp = -1;
  for (; unit < A.n_units; unit += nwarps) {
    const int64_t blk = unit / A.G;
    const int grp = static_cast<int>(unit - blk * A.G);
    const int c0 = A.c_lo + grp * T::CPW;
    const int off = c0 * A.D;
    if (ATT == 3 && grp != da_grp) {
      if (da_grp >= 0) T::atomic_add(A.ga + (A.c_lo + da_grp * T::CPW) * A.D, da, lane, A.D);
      zero<T>(da);
      da_grp = grp;
    }
    int cidx[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) cidx[k] = c0 + T::ch(k, lane);
    float ar[R];
    if (ATT == 3) T::load(ar, A.a + off, lane, A.D);
    const int64_t base = blk * kPairBlock;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPairBlock), A.m - base));
    const int64_t myi = lane < cnt ? A.pi[base + lane] : 0;
    const int64_t myj = lane < cnt ? A.pj[base + lane] : 0;
    int64_t cur_i = -1;
    float pr[R], dP[R];
    zero<T>(dP);
    for (int t = 0; t < cnt; ++t) {
      const int64_t i = __shfl_sync(FULL, myi, t);
      const int64_t j = __shfl_sync(FULL, myj, t);
      float g[NCH];
#pragma unroll
      for (int k = 0; k < NCH; ++k) g[k] = __ldg(A.g_out + (base + t) * A.Cs + (cidx[k] - A.c_lo));
      if (ATT == 1) {
        if (T::writer(lane)) {
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            atomicAdd(A.gP + i * A.C + cidx[k], g[k]);
            atomicAdd(A.gQ + j * A.C + cidx[k], g[k]);
          }
        }
        continue;
      }
      if (i != cur_i) {
        if (cur_i >= 0) T::atomic_add(A.gP + cur_i * CD + off, dP, lane, A.D);
        zero<T>(dP);
        T::load(pr, A.P + i * A.ldp + off, lane, A.D);
        cur_i = i;
      }
      float q[R], dq[R];
      T::load(q, A.Q + j * A.ldq + off, lane, A.D);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (ATT == 3) {
          const float z = pr[r] + q[r];
          const float dz = g[r / RPC] * ar[r] * (z > 0.0f ? 1.0f : 0.01f);
          dP[r] += dz;
          dq[r] = dz;
          da[r] = fmaf(g[r / RPC], lrelu01(z), da[r]);
        } else {
          dP[r] = fmaf(g[r / RPC], q[r], dP[r]);
          dq[r] = g[r / RPC] * pr[r];
        }
      }
      T::atomic_add(A.gQ + j * CD + off, dq, lane, A.D);
    }
    if (ATT >= 2 && cur_i >= 0) T::atomic_add(A.gP + cur_i * CD + off, dP, lane, A.D);
  }
  if (ATT == 3 && da_grp >= 0) T::atomic_add(A.ga + (A.c_lo + da_grp * T::CPW) * A.D, da, lane, A.D);
}

// att 3 backward from the forward's sign record, one launch per side (no 2 KB atomic per pair, no
// re-gather of the other side's rows): over runs of equal key (i: the list is row-sorted; j: through
// the caller's column-sorted permutation) accumulate U = sum g * [z > 0] and S = sum g per channel;
// at the run end  u = 0.99 U + 0.01 S,  da += X_key (.) u  (Euler, as in the layer kernels),
// gX_key += a (.) u  (one vector reduction per run; runs may straddle 32-pair blocks).
template <class T>
__global__ void __launch_bounds__(256) k_pair_bwd_sign(const PairArgs A) {
  constexpr int R = T::R, NCH = T::NCH, RPC = T::RPC;
  constexpr int SBPL = (R + 7) / 8;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  int64_t unit = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int CD = A.C * A.D;
  float da[R];
  zero<T>(da);
  int da_grp = -1;
  for (; unit < A.n_units; unit += nwarps) {
    const int64_t blk = unit / A.G;
    const int grp = static_cast<int>(unit - blk * A.G);
    const int c0 = A.c_lo + grp * T::CPW;
    const int off = c0 * A.D;
    if (grp != da_grp) {
      if (da_grp >= 0) T::atomic_add(A.ga + (A.c_lo + da_grp * T::CPW) * A.D, da, lane, A.D);
      zero<T>(da);
      da_grp = grp;
    }
    int cidx[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) cidx[k] = c0 + T::ch(k, lane) - A.c_lo;
    float ar[R];
    T::load(ar, A.a + off, lane, A.D);
    const int64_t base = blk * kPairBlock;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPairBlock), A.m - base));
    int mypid = 0, mykey = 0;       // m < 2^31 (checked by the entry point), node ids < 2^31
    if (lane < cnt) {
      mypid = A.perm ? A.perm[base + lane] : static_cast<int>(base + lane);
      mykey = static_cast<int>(A.key[mypid]);
    }
    int cur = -1;
    float U[R], S[NCH];
    auto flush = [&]() {
      float x[R], o[R];
      T::load(x, A.X + static_cast<int64_t>(cur) * A.ldx + off, lane, A.D);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float u = fmaf(0.99f, U[r], 0.01f * S[r / RPC]);
        da[r] = fmaf(x[r], u, da[r]);
        o[r] = ar[r] * u;
      }
      T::atomic_add(A.gX + static_cast<int64_t>(cur) * CD + off, o, lane, A.D);
    };
    // PU pairs in flight per warp: the record / gradient loads of a group are issued together
    // (the column pass reads them at random addresses -- one pair at a time is latency-bound)
    constexpr int PU = 4;
    for (int t0 = 0; t0 < cnt; t0 += PU) {
      int keys[PU];
      unsigned sg[PU];
      float g[PU][NCH];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int t = min(t0 + u, cnt - 1);
        const int64_t pid = __shfl_sync(FULL, mypid, t);
        keys[u] = __shfl_sync(FULL, mykey, t);
        const int64_t so = ((pid * A.G + grp) * 32 + lane) * SBPL;
        sg[u] = SBPL == 1 ? static_cast<unsigned>(ld_stream(A.sign + so))
                          : static_cast<unsigned>(ld_stream(reinterpret_cast<const unsigned short*>(A.sign + so)));
#pragma unroll
        for (int k = 0; k < NCH; ++k) g[u][k] = __ldg(A.g_out + pid * A.Cs + cidx[k]);
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        if (t0 + u < cnt) {
          if (keys[u] != cur) {
            if (cur >= 0) flush();
            cur = keys[u];
            zero<T>(U);
#pragma unroll
            for (int k = 0; k < NCH; ++k) S[k] = 0.0f;
          }
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            S[k] += g[u][k];
#pragma unroll
            for (int r = k * RPC; r < (k + 1) * RPC; ++r) U[r] = fmaf(sign_pos_f<R>(sg[u], r), g[u][k], U[r]);
          }
        }
      }
    }
    if (cur >= 0) flush();
  }
  if (da_grp >= 0) T::atomic_add(A.ga + (A.c_lo + da_grp * T::CPW) * A.D, da, lane, A.D);
}

template <class T>
static int launch_pair_sign(PairArgs A, cudaStream_t st) {
  A.G = A.Cs / T::CPW;
  const int64_t blocks_of_pairs = (A.m + kPairBlock - 1) / kPairBlock;
  A.n_units = blocks_of_pairs * A.G;
  if (A.n_units == 0) return EDIS_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = launch_grid(reinterpret_cast<const void*>(&k_pair_bwd_sign<T>), 256, 0, sms);
  const int64_t need = (A.n_units + 7) / 8;
  if (need < blocks) blocks = static_cast<int>(need);
  // row side (the list is sorted by i), then column side through the permutation
  A.key = A.pi; A.perm = nullptr; A.X = A.P; A.ldx = A.ldp; A.gX = A.gP;
  k_pair_bwd_sign<T><<<blocks, 256, 0, st>>>(A);
  A.key = A.pj; A.perm = A.col_perm; A.X = A.Q; A.ldx = A.ldq; A.gX = A.gQ;
  k_pair_bwd_sign<T><<<blocks, 256, 0, st>>>(A);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

template <class T, int ATT>
static int launch_pair(bool bwd, PairArgs A, cudaStream_t st) {
  if (bwd && ATT == 3 && A.sign && A.col_perm) return launch_pair_sign<T>(A, st);
  A.G = A.Cs / T::CPW;
  const int64_t blocks_of_pairs = (A.m + kPairBlock - 1) / kPairBlock;
  A.n_units = blocks_of_pairs * A.G;
  if (A.n_units == 0) return EDIS_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const void* fn = bwd ? reinterpret_cast<const void*>(&k_pair_bwd<T, ATT>)
                       : reinterpret_cast<const void*>(&k_pair_fwd<T, ATT>);
  int blocks = launch_grid(fn, 256, 0, sms);
  const int64_t need = (A.n_units + 7) / 8;
  if (need < blocks) blocks = static_cast<int>(need);
  if (bwd) k_pair_bwd<T, ATT><<<blocks, 256, 0, st>>>(A);
  else k_pair_fwd<T, ATT><<<blocks, 256, 0, st>>>(A);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

template <class T>
static int launch_pair_att(bool bwd, const PairArgs& A, int att, cudaStream_t st) {
  switch (att) {
    case 1: return launch_pair<T, 1>(bwd, A, st);
    case 2: return launch_pair<T, 2>(bwd, A, st);
    case 3: return launch_pair<T, 3>(bwd, A, st);
  }
  set_error("att must be 1, 2 or 3");
  return EDIS_ERR_ARG;
}

static int launch_pair_any(bool bwd, const PairArgs& A, int att, cudaStream_t st) {
  const int D = A.D, Cs = A.Cs;
  const bool aligned = att == 1 || ((reinterpret_cast<uintptr_t>(A.P) % 16 == 0) && (reinterpret_cast<uintptr_t>(A.Q) % 16 == 0) &&
                                    A.ldp % 4 == 0 && A.ldq % 4 == 0);
  if (!aligned && A.sign) {
    set_error("pair scoring with a sign record needs 16-byte aligned P / Q rows (ld %% 4 == 0)");
    return EDIS_ERR_ARG;
  }
  // (whole-row warps, VecT<4,16>, measured slower here: pair fwd 94 vs 86 ms, bwd 158 vs 138 ms at M = 210M)
  if (aligned && D == 64 && Cs % 4 == 0) return launch_pair_att<VecT<2, 16>>(bwd, A, att, st);
  if (aligned && D == 64 && Cs % 2 == 0) return launch_pair_att<VecT<1, 16>>(bwd, A, att, st);
  if (aligned && D == 128) return launch_pair_att<VecT<1, 32>>(bwd, A, att, st);
  if (D <= 32) return launch_pair_att<ScaT<1>>(bwd, A, att, st);
  if (D <= 64) return launch_pair_att<ScaT<2>>(bwd, A, att, st);
  if (D <= 128) return launch_pair_att<ScaT<4>>(bwd, A, att, st);
  if (D <= 256) return launch_pair_att<ScaT<8>>(bwd, A, att, st);
  set_error("unsupported channel width D=%d (max 256)", D);
  return EDIS_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  }
  return v;  // valid in thread 0
}

__global__ void k_wmse_fwd(int64_t m, int cs, const float* scores, const float* target, float w_neg,
                           double* acc) {
  double local = 0.0;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < m;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float s = 0.0f;
    for (int c = 0; c < cs; ++c) s += scores[k * cs + c];
    const float pr = 1.0f / (1.0f + expf(-s));
    const float t = target[k];
    const float d = pr - t;
    local += static_cast<double>((t != 0.0f ? 1.0f : w_neg) * d * d);
  }
  const double tot = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(acc, tot);
}
__global__ void k_finalize_mean(const double* acc, double inv_count, float* loss) {
  loss[0] = static_cast<float>(acc[0] * inv_count);
}
__global__ void k_wmse_bwd(int64_t m, int cs, const float* scores, const float* target, float w_neg,
                           float inv_m, const float* g_loss, float* g_scores) {
  const float gl = g_loss[0];
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < m;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float s = 0.0f;
    for (int c = 0; c < cs; ++c) s += scores[k * cs + c];
    const float pr = 1.0f / (1.0f + expf(-s));
    const float t = target[k];
    const float g = gl * inv_m * 2.0f * (t != 0.0f ? 1.0f : w_neg) * (pr - t) * pr * (1.0f - pr);
    for (int c = 0; c < cs; ++c) g_scores[k * cs + c] = g;
  }
}

__global__ void k_nll_fwd(int64_t n, int kk, const float* logits, int label, double* acc) {
  double local = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* row = logits + i * kk;
    float mx = row[0];
    for (int c = 1; c < kk; ++c) mx = fmaxf(mx, row[c]);
    float se = 0.0f;
    for (int c = 0; c < kk; ++c) se += expf(row[c] - mx);
    local += static_cast<double>(mx + logf(se) - row[label]);
  }
  const double tot = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(acc, tot);
}
__global__ void k_nll_bwd(int64_t n, int kk, const float* logits, int label, float inv_n,
                          const float* g_loss, float* g_logits) {
  const float gl = g_loss[0] * inv_n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* row = logits + i * kk;
    float mx = row[0];
    for (int c = 1; c < kk; ++c) mx = fmaxf(mx, row[c]);
    float se = 0.0f;
    for (int c = 0; c < kk; ++c) se += expf(row[c] - mx);
    const float inv = 1.0f / se;
    for (int c = 0; c < kk; ++c)
      g_logits[i * kk + c] = gl * (expf(row[c] - mx) * inv - (c == label ? 1.0f : 0.0f));
  }
}

// ------------------------------------------------------------------ COO drop-ins
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // monotone int mapping of IEEE floats
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__global__ void k_fill(float* p, float v, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_coo_max(int64_t e, const float* values, float* vmax) {
  float m = -INFINITY;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < e;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, values[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) atomic_max_float(vmax, m);
}
__global__ void k_coo_exp_sum(int64_t e, const int64_t* row, const float* values, const float* vmax,
                              float* out, float* denom) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= e) return;
  const float ex = expf(values[k] - vmax[0]);
  out[k] = ex;
  atomicAdd(denom + row[k], ex);
}
__global__ void k_coo_div(int64_t e, const int64_t* row, float* out, const float* denom) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < e) out[k] = out[k] / (denom[row[k]] + 1e-10f);
}
__global__ void k_coo_rowdot(int64_t e, const int64_t* row, const float* out, const float* g, float* rowdot) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < e) atomicAdd(rowdot + row[k], out[k] * g[k]);
}
__global__ void k_coo_softmax_bwd(int64_t e, const int64_t* row, const float* out, const float* g,
                                  const float* rowdot, float* gv) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < e) gv[k] = out[k] * (g[k] - rowdot[row[k]]);
}
// one warp per entry, lanes stride the feature axis
__global__ void k_coo_spmm_fwd(int64_t e, int64_t f, const int64_t* row, const int64_t* col,
                               const float* values, const float* mat, float* out) {
  const int64_t k = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (k >= e) return;
  const int lane = threadIdx.x & 31;
  const float v = values[k];
  const float* src = mat + col[k] * f;
  float* dst = out + row[k] * f;
  for (int64_t x = lane; x < f; x += 32) atomicAdd(dst + x, v * src[x]);
}
__global__ void k_coo_spmm_bwd(int64_t e, int64_t f, const int64_t* row, const int64_t* col,
                               const float* values, const float* mat, const float* g_out,
                               float* g_values, float* g_mat) {
  const int64_t k = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (k >= e) return;
  const int lane = threadIdx.x & 31;
  const float v = values[k];
  const float* src = mat + col[k] * f;
  const float* go = g_out + row[k] * f;
  float* gm = g_mat + col[k] * f;
  float dot = 0.0f;
  for (int64_t x = lane; x < f; x += 32) {
    const float g = go[x];
    dot = fmaf(g, src[x], dot);
    atomicAdd(gm + x, v * g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
  if (lane == 0) g_values[k] = dot;
}

static inline unsigned nblk(int64_t n, int per = 256) { return static_cast<unsigned>((n + per - 1) / per); }
static inline unsigned ngrid(int64_t n) { return static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 148 * 8)); }

}  // namespace edis

using namespace edis;

static int pair_common(const char* who, const edis_layer_desc* d, int64_t n, int64_t m,
                       const int64_t* pi, const int64_t* pj, int32_t c_lo, int32_t c_hi,
                       const float* P, const float* Q, const float* a, PairArgs* A) {
  EDIS_CHECK_ARG(d && d->att >= 1 && d->att <= 3, "%s: bad descriptor", who);
  EDIS_CHECK_ARG(n > 0 && m >= 0 && (m == 0 || (pi && pj)) && P && Q, "%s: null pointer", who);
  EDIS_CHECK_ARG(0 <= c_lo && c_lo < c_hi && c_hi <= d->C, "%s: channel range [%d,%d) of %d", who, c_lo, c_hi, d->C);
  EDIS_CHECK_ARG(d->att != 3 || a, "%s: att=3 needs a", who);
  A->m = m; A->pi = pi; A->pj = pj; A->c_lo = c_lo; A->Cs = c_hi - c_lo; A->C = d->C; A->D = d->D;
  A->P = P; A->Q = Q; A->a = a;
  return EDIS_OK;
}

// bytes of the att-3 sign record of a pair list (layout private to the pair kernels)
extern "C" int64_t edis_pair_sign_bytes(const edis_layer_desc* d, int64_t m, int32_t c_lo, int32_t c_hi) {
  if (!d || m < 0 || c_hi <= c_lo) return EDIS_ERR_ARG;
  if (d->att != 3) return 0;
  const int cs = c_hi - c_lo;
  // 32 lanes x SBPL bytes per (pair, channel group); channels per group as in launch_pair_any
  const int cpw = d->D != 64 ? 1 : (cs % 4 == 0 ? 4 : (cs % 2 == 0 ? 2 : 1));
  return m * static_cast<int64_t>(cs / cpw) * 32 + 64;
}

extern "C" int edis_pair_score_fwd(const edis_layer_desc* d, int64_t n, int64_t m, const int64_t* pi,
                                   const int64_t* pj, int32_t c_lo, int32_t c_hi, const float* P,
                                   int64_t ldp, const float* Q, int64_t ldq, const float* a,
                                   float* out, uint8_t* psign, void* stream) {
  PairArgs A = {};
  int rc = pair_common("edis_pair_score_fwd", d, n, m, pi, pj, c_lo, c_hi, P, Q, a, &A);
  if (rc) return rc;
  EDIS_CHECK_ARG(out || m == 0, "edis_pair_score_fwd: null out");
  A.ldp = ldp; A.ldq = ldq; A.out = out; A.sign = psign;
  return launch_pair_any(false, A, d->att, static_cast<cudaStream_t>(stream));
}

extern "C" int edis_pair_score_bwd(const edis_layer_desc* d, int64_t n, int64_t m, const int64_t* pi,
                                   const int64_t* pj, int32_t c_lo, int32_t c_hi, const float* P,
                                   int64_t ldp, const float* Q, int64_t ldq, const float* a,
                                   const float* g_out, const uint8_t* psign, const int32_t* col_perm,
                                   float* gP, float* gQ, float* ga, void* stream) {
  PairArgs A = {};
  int rc = pair_common("edis_pair_score_bwd", d, n, m, pi, pj, c_lo, c_hi, P, Q, a, &A);
  if (rc) return rc;
  EDIS_CHECK_ARG((g_out && gP && gQ) || m == 0, "edis_pair_score_bwd: null pointer");
  EDIS_CHECK_ARG(d->att != 3 || ga, "edis_pair_score_bwd: att=3 needs ga");
  EDIS_CHECK_ARG((m < (int64_t(1) << 31) && n < (int64_t(1) << 31)) || !col_perm,
                 "edis_pair_score_bwd: the sign-record path needs m, n < 2^31");
  A.ldp = ldp; A.ldq = ldq; A.g_out = g_out; A.gP = gP; A.gQ = gQ; A.ga = ga;
  A.sign = const_cast<uint8_t*>(psign); A.col_perm = col_perm;
  return launch_pair_any(true, A, d->att, static_cast<cudaStream_t>(stream));
}

static float neg_weight(int64_t m, int64_t n_pos) {
  // utils.py:288-291: edge_num / (shape[0]**2 - edge_num), python double -> float32 fill
  const double total = static_cast<double>(m) * static_cast<double>(m);
  return static_cast<float>(static_cast<double>(n_pos) / (total - static_cast<double>(n_pos)));
}

extern "C" int edis_ssl_wmse_fwd(int64_t m, int32_t cs, const float* scores, const float* target,
                                 int64_t n_pos, int64_t m_total, float* loss, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  EDIS_CHECK_ARG(m > 0 && cs > 0 && scores && target && loss && m_total >= m, "edis_ssl_wmse_fwd: bad arguments");
  if (!workspace || workspace_bytes < 8) {
    set_error("edis_ssl_wmse_fwd: workspace must hold 8 bytes");
    return EDIS_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(workspace);
  EDIS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  k_wmse_fwd<<<ngrid(m), 256, 0, st>>>(m, cs, scores, target, neg_weight(m_total, n_pos), acc);
  k_finalize_mean<<<1, 1, 0, st>>>(acc, 1.0 / static_cast<double>(m_total), loss);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

extern "C" int edis_ssl_wmse_bwd(int64_t m, int32_t cs, const float* scores, const float* target,
                                 int64_t n_pos, int64_t m_total, const float* g_loss, float* g_scores,
                                 void* stream) {
  EDIS_CHECK_ARG(m > 0 && cs > 0 && scores && target && g_loss && g_scores && m_total >= m,
                 "edis_ssl_wmse_bwd: bad arguments");
  k_wmse_bwd<<<ngrid(m), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      m, cs, scores, target, neg_weight(m_total, n_pos), static_cast<float>(1.0 / static_cast<double>(m_total)),
      g_loss, g_scores);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

extern "C" int edis_nll_const_label_fwd(int64_t n, int32_t k, const float* logits, int32_t label,
                                        float* loss, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  EDIS_CHECK_ARG(n > 0 && k > 0 && label >= 0 && label < k && logits && loss, "edis_nll_const_label_fwd: bad arguments");
  if (!workspace || workspace_bytes < 8) {
    set_error("edis_nll_const_label_fwd: workspace must hold 8 bytes");
    return EDIS_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(workspace);
  EDIS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  k_nll_fwd<<<ngrid(n), 256, 0, st>>>(n, k, logits, label, acc);
  k_finalize_mean<<<1, 1, 0, st>>>(acc, 1.0 / static_cast<double>(n), loss);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

extern "C" int edis_nll_const_label_bwd(int64_t n, int32_t k, const float* logits, int32_t label,
                                        const float* g_loss, float* g_logits, void* stream) {
  EDIS_CHECK_ARG(n > 0 && k > 0 && label >= 0 && label < k && logits && g_loss && g_logits,
                 "edis_nll_const_label_bwd: bad arguments");
  k_nll_bwd<<<ngrid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, k, logits, label, static_cast<float>(1.0 / static_cast<double>(n)), g_loss, g_logits);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

extern "C" int edis_sp_softmax_fwd(int64_t n, int64_t e, const int64_t* row, const float* values,
                                   float* out, float* denom, float* vmax, void* stream) {
  EDIS_CHECK_ARG(n > 0 && e >= 0 && (e == 0 || (row && values && out)) && denom && vmax, "edis_sp_softmax_fwd: bad arguments");
  if (e == 0) return EDIS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_fill<<<1, 32, 0, st>>>(vmax, -INFINITY, 1);
  EDIS_CUDA(cudaMemsetAsync(denom, 0, n * sizeof(float), st));
  k_coo_max<<<ngrid(e), 256, 0, st>>>(e, values, vmax);
  k_coo_exp_sum<<<nblk(e), 256, 0, st>>>(e, row, values, vmax, out, denom);
  k_coo_div<<<nblk(e), 256, 0, st>>>(e, row, out, denom);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

extern "C" int edis_sp_softmax_bwd(int64_t n, int64_t e, const int64_t* row, const float* out,
                                   const float* g_out, float* g_values, float* rowdot, void* stream) {
  EDIS_CHECK_ARG(n > 0 && e >= 0 && (e == 0 || (row && out && g_out && g_values)) && rowdot, "edis_sp_softmax_bwd: bad arguments");
  if (e == 0) return EDIS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EDIS_CUDA(cudaMemsetAsync(rowdot, 0, n * sizeof(float), st));
  k_coo_rowdot<<<nblk(e), 256, 0, st>>>(e, row, out, g_out, rowdot);
  k_coo_softmax_bwd<<<nblk(e), 256, 0, st>>>(e, row, out, g_out, rowdot, g_values);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

extern "C" int edis_sp_matmul_fwd(int64_t n, int64_t e, int64_t f, const int64_t* row,
                                  const int64_t* col, const float* values, const float* mat,
                                  float* out, void* stream) {
  EDIS_CHECK_ARG(n > 0 && e >= 0 && f > 0 && mat && out && (e == 0 || (row && col && values)), "edis_sp_matmul_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EDIS_CUDA(cudaMemsetAsync(out, 0, n * f * sizeof(float), st));
  if (e > 0) k_coo_spmm_fwd<<<nblk(e * 32), 256, 0, st>>>(e, f, row, col, values, mat, out);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}

extern "C" int edis_sp_matmul_bwd(int64_t n, int64_t e, int64_t f, const int64_t* row,
                                  const int64_t* col, const float* values, const float* mat,
                                  const float* g_out, float* g_values, float* g_mat, void* stream) {
  EDIS_CHECK_ARG(n > 0 && e >= 0 && f > 0 && mat && g_out && g_mat && (e == 0 || (row && col && values && g_values)),
                 "edis_sp_matmul_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EDIS_CUDA(cudaMemsetAsync(g_mat, 0, n * f * sizeof(float), st));
  if (e > 0) k_coo_spmm_bwd<<<nblk(e * 32), 256, 0, st>>>(e, f, row, col, values, mat, g_out, g_values, g_mat);
  EDIS_CUDA(cudaGetLastError());
  return EDIS_OK;
}
