// Brute-force k-nearest-neighbour classification of embeddings on the GPU (eval-time consumer of the encoder).
//
// Replaces solvers/vince_solver.py:651-693: sklearn KDTree(all_features).query(all_features, k=11), drop the first
// (self) match, scipy.stats.mode over the 10 neighbour labels, mean(pred == label).  n is ~10^4 and D = 128, so the
// exact O(n^2 D) search is 13 GFLOP: KNN_Q queries per block share every candidate row they stream (L2-resident), each
// thread keeps a sorted top-k list in registers, a shared-memory merge picks the block's k best per query.
#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int KNN_Q = 4;          // queries per block (4 x (k+1) x 2 registers of sorted lists per thread)
constexpr int KNN_T = 128;        // threads per block
constexpr int KNN_KMAX = 16;      // k + 1 <= KNN_KMAX

struct KnnItem {
  float d;
  int idx;
};
// strict weak order: by distance, ties by index (KDTree returns equal-distance neighbours in index order for exact
// duplicates; any fixed rule keeps the result deterministic)
__device__ __forceinline__ bool knn_less(const KnnItem& a, const KnnItem& b) {
  return a.d < b.d || (a.d == b.d && a.idx < b.idx);
}

template <int KK>
__global__ void __launch_bounds__(KNN_T) knn_kernel(const float* __restrict__ feats, const int64_t* __restrict__ labels,
                                                    int n, int D, int64_t* __restrict__ nbr_idx,
                                                    float* __restrict__ nbr_dist, int64_t* __restrict__ pred,
                                                    int n_classes_hint) {
  extern __shared__ float sm[];                      // [KNN_Q][D] queries, then the merge area
  float* qs = sm;
  KnnItem* merge = reinterpret_cast<KnnItem*>(sm + KNN_Q * D);     // [KNN_Q][KNN_T][KK]
  const int q0 = blockIdx.x * KNN_Q;
  for (int i = threadIdx.x; i < KNN_Q * D; i += KNN_T) {
    const int q = q0 + i / D;
    qs[i] = q < n ? feats[(int64_t)q * D + (i % D)] : 0.f;
  }
  __syncthreads();
  KnnItem best[KNN_Q][KK];
#pragma unroll
  for (int q = 0; q < KNN_Q; ++q)
#pragma unroll
    for (int i = 0; i < KK; ++i) best[q][i].d = INFINITY, best[q][i].idx = 0x7fffffff;
  for (int j = threadIdx.x; j < n; j += KNN_T) {
    const float* c = feats + (int64_t)j * D;
    float acc[KNN_Q];
#pragma unroll
    for (int q = 0; q < KNN_Q; ++q) acc[q] = 0.f;
    for (int e = 0; e < D; e += 4) {
      const float4 cv = __ldg(reinterpret_cast<const float4*>(c + e));
#pragma unroll
      for (int q = 0; q < KNN_Q; ++q) {
        const float4 qv = *reinterpret_cast<const float4*>(qs + q * D + e);
        const float dx = qv.x - cv.x, dy = qv.y - cv.y, dz = qv.z - cv.z, dw = qv.w - cv.w;
        acc[q] = fmaf(dx, dx, acc[q]);
        acc[q] = fmaf(dy, dy, acc[q]);
        acc[q] = fmaf(dz, dz, acc[q]);
        acc[q] = fmaf(dw, dw, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < KNN_Q; ++q) {
      KnnItem it;
      it.d = acc[q], it.idx = j;
      if (knn_less(it, best[q][KK - 1])) {
        best[q][KK - 1] = it;
#pragma unroll
        for (int i = KK - 1; i > 0; --i) {
          if (knn_less(best[q][i], best[q][i - 1])) {
            const KnnItem t = best[q][i];
            best[q][i] = best[q][i - 1];
            best[q][i - 1] = t;
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < KNN_Q; ++q)
#pragma unroll
    for (int i = 0; i < KK; ++i) merge[(q * KNN_T + threadIdx.x) * KK + i] = best[q][i];
  __syncthreads();
  // one warp per query merges the KNN_T sorted per-thread lists: KK rounds of (each lane scans the heads of its 4
  // lists, warp argmin, the winning lane advances that list)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < KNN_Q; q += KNN_T / 32) {
    if (q0 + q >= n) continue;
    int head[KNN_T / 32];                       // this lane owns lists lane, lane+32, ...
#pragma unroll
    for (int t = 0; t < KNN_T / 32; ++t) head[t] = 0;
    int64_t my_lab[KK];
    for (int r = 0; r < KK; ++r) {
      KnnItem loc;
      loc.d = INFINITY, loc.idx = 0x7fffffff;
      int loc_t = 0;
#pragma unroll
      for (int t = 0; t < KNN_T / 32; ++t) {
        if (head[t] < KK) {
          const KnnItem it = merge[(q * KNN_T + lane + 32 * t) * KK + head[t]];
          if (knn_less(it, loc)) loc = it, loc_t = t;
        }
      }
      // warp argmin
      KnnItem w = loc;
      int wl = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        KnnItem o_it;
        o_it.d = __shfl_xor_sync(0xffffffffu, w.d, o);
        o_it.idx = __shfl_xor_sync(0xffffffffu, w.idx, o);
        const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
        if (knn_less(o_it, w)) w = o_it, wl = ol;
      }
      if (wl == lane) {
#pragma unroll
        for (int t = 0; t < KNN_T / 32; ++t)
          if (t == loc_t) head[t]++;
      }
      // rank 0 is the self match the reference drops (neighbors[:, 1:], vince_solver.py:677)
      if (r > 0) {
        if (lane == 0) {
          nbr_idx[(int64_t)(q0 + q) * (KK - 1) + r - 1] = w.idx;
          if (nbr_dist) nbr_dist[(int64_t)(q0 + q) * (KK - 1) + r - 1] = sqrtf(w.d);
        }
        my_lab[r] = (w.idx >= 0 && w.idx < n && labels) ? labels[w.idx] : -1;
      }
    }
    if (lane == 0 && pred != nullptr) {
      // scipy.stats.mode: most frequent label, the SMALLEST one on ties
      int64_t best_lab = -1;
      int best_cnt = 0;
      for (int a = 1; a < KK; ++a) {
        int cnt = 0;
        for (int b = 1; b < KK; ++b) cnt += (my_lab[b] == my_lab[a]);
        if (cnt > best_cnt || (cnt == best_cnt && my_lab[a] < best_lab)) best_cnt = cnt, best_lab = my_lab[a];
      }
      pred[q0 + q] = best_lab;
    }
  }
}

int knn_launch(const float* feats, const int64_t* labels, int n, int D, int k, int64_t* nbr_idx, float* nbr_dist,
               int64_t* pred, cudaStream_t stream) {
  VB_REQUIRE(n > 0 && D > 0 && D % 4 == 0, "knn: n=%d D=%d (D must be a positive multiple of 4)", n, D);
  VB_REQUIRE(k >= 1 && k + 1 <= KNN_KMAX && k + 1 <= n, "knn: k=%d unsupported (1 <= k <= %d, k < n)", k, KNN_KMAX - 1);
  VB_REQUIRE((reinterpret_cast<uintptr_t>(feats) & 15) == 0, "knn: features must be 16-byte aligned");
  const int blocks = (n + KNN_Q - 1) / KNN_Q;
  const int kk = k + 1;
  const size_t smem = (size_t)KNN_Q * D * sizeof(float) + (size_t)KNN_Q * KNN_T * kk * sizeof(KnnItem);
  VB_REQUIRE(smem <= 200 * 1024, "knn: D=%d too large", D);
#define VB_KNN(KK_)                                                                                                \
  if (kk == KK_) {                                                                                                 \
    VB_CHECK_CUDA(cudaFuncSetAttribute(knn_kernel<KK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    knn_kernel<KK_><<<blocks, KNN_T, smem, stream>>>(feats, labels, n, D, nbr_idx, nbr_dist, pred, 0);            \
    VB_CHECK_CUDA(cudaGetLastError());                                                                             \
    return VB_OK;                                                                                                  \
  }
  VB_KNN(2) VB_KNN(4) VB_KNN(6) VB_KNN(11) VB_KNN(16)
#undef VB_KNN
  VB_REQUIRE(false, "knn: k=%d is not instantiated (supported k: 1, 3, 5, 10, 15)", k);
}

}  // namespace vb
