// Fused InfoNCE backward w.r.t. the queries for sm_100a (reproduces what autograd computes
// through vince_model.py:213-233 + loss_util.py:7-62 when `dist.backward()` is called in vince_solver.py:465).
//
//   L = grad_dist * mean_{i,p} l_ip,   l_ip = -(z_ip - m_i) + log(exp(z_ip - m_i) + Zneg_i),   z = sigma / T
//   dL/dz_ij (j negative) = g * R_i * exp(z_ij - m_i),      R_i = sum_p 1 / (exp(z_ip - m_i) + Zneg_i)
//   dL/dz_ip (p positive) = -g * (1 - w_ip),                w_ip = exp(z_ip - m_i) / (exp(z_ip - m_i) + Zneg_i)
//   dL/dq_i = (1/T) * sum_j dL/dz_ij * n_j                  g = grad_dist / (B * nP)
// Keys and queue carry no gradient (no_grad / detach, vince_model.py:598,610; storage_queue.py:53).
//
// Like the forward this is ONE streaming pass over [keys || queue] that never writes a [B, B+K] matrix:
//   main kernel : each CTA keeps a 128-row block of queries in shared memory and streams 128-column tiles through a
//                 TMA ring; the similarities are recomputed on the tensor cores (tcgen05 kind::tf32 into
//                 double-buffered TMEM) exactly as in the forward; 256 threads (two per query row, half of the
//                 feature dimension each) turn every tile into probabilities e_ij = exp(z_ij - m_i) (0 on positives)
//                 straight from TMEM and accumulate  A_i += e_ij * n_j  in registers, reading n_j from the very
//                 shared-memory tile the MMA consumed (warp-wide broadcast loads).  Per-CTA partial sums go to a
//                 workspace [row block][slice][128][D].
//   finalize    : one warp per row sums the partials in a fixed order, applies g * R_i / T and adds the positives'
//                 term with exact fp32 keys.
//   symmetric   : for the self-batch loss (vince_model.py:213-222) the columns are the queries themselves and also
//                 carry gradient; a small B x B kernel adds  (1/T) * sum_i dL/dz_ij * q_i  to row j.
// The second product runs on the CUDA cores (B*(B+K)*D FMAs, ~40 us at B=256, K=65536, D=128); moving it to the
// tensor cores needs an MN-major B operand and is left for a later round.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int NB_BM = 128;
constexpr int NB_BN = 128;
constexpr int NB_THREADS = 64 + 256;     // warp 0 TMA, warp 1 MMA, warps 2..9 probability + accumulate
constexpr int NB_MAX_STAGES = 4;

struct NceBwdParams {
  CUtensorMap q_map, keys_map, queue_map;
  int B, Bk, K, D;
  int nf;
  int nmb, nkt, nqt, slices;
  int num_stages;
  float scale_log2;          // log2(e) / T
  const float* row_lse;      // [B][2] (row max of z, Zneg) from the forward
  float* partial;            // [nmb][slices][128][D]
};

__device__ __forceinline__ float bwd_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(NB_THREADS, 1) infonce_bwd_kernel(const __grid_constant__ NceBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int atoms = D / 32;                        // 128-byte swizzle atoms along D
  constexpr uint32_t q_bytes = atoms * NB_BM * 128;
  constexpr uint32_t stage_bytes = atoms * NB_BN * 128;
  constexpr int DH = D / 2;                            // features per accumulating thread

  uint8_t* q_smem = smem;
  uint8_t* stages = smem + q_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stages + (size_t)p.num_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + NB_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + NB_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* q_full = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(q_full + 1);

  const int m_blk = blockIdx.x % p.nmb;
  const int sidx = blockIdx.x / p.nmb;
  const int T = p.nkt + p.nqt;
  const int t_begin = (int)(((int64_t)T * sidx) / p.slices);
  const int t_end = (int)(((int64_t)T * (sidx + 1)) / p.slices);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    if (p.nkt) tma_prefetch_desc(&p.keys_map);
    if (p.nqt) tma_prefetch_desc(&p.queue_map);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 256);                   // the accumulating threads release a tile, not the MMA
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 256);
    }
    mbar_init(q_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * NB_BN);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (t_end > t_begin) {
      if (elect_one()) {
        mbar_expect_tx(q_full, q_bytes);
        for (int a = 0; a < atoms; ++a) tma_load_2d(q_smem + a * NB_BM * 128, &p.q_map, q_full, a * 32, m_blk * NB_BM);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = stages + (size_t)stage * stage_bytes;
        const bool is_key = t < p.nkt;
        const CUtensorMap* map = is_key ? &p.keys_map : &p.queue_map;
        const int row0 = (is_key ? t : t - p.nkt) * NB_BN;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          for (int a = 0; a < atoms; ++a) tma_load_2d(st + a * NB_BN * 128, map, &full_bar[stage], a * 32, row0);
        }
        __syncwarp();
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (t_end > t_begin) {
      constexpr uint32_t idesc = make_idesc(UMMA_FMT_TF32, NB_BM, NB_BN);
      mbar_wait(q_full, 0);
      tc_fence_after_sync();
      const uint32_t q_addr = smem_u32(q_smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin, lt = 0; t < t_end; ++t, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint32_t st = smem_u32(stages + (size_t)stage * stage_bytes);
        const uint32_t d_tmem = tmem_base + acc * NB_BN;
        if (elect_one()) {
          for (int a = 0; a < atoms; ++a) {
            const uint64_t da0 = make_smem_desc(q_addr + a * NB_BM * 128, 16, 1024, UMMA_LAYOUT_SW128);
            const uint64_t db0 = make_smem_desc(st + a * NB_BN * 128, 16, 1024, UMMA_LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, (a | k) != 0);
          }
          umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ---- probabilities from TMEM, A_i += e_ij * n_j; thread = (query row, half of the feature dimension) ----
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                  // warps 2..5 -> features [0, D/2), warps 6..9 -> [D/2, D)
    const int r = quarter * 32 + lane;
    const int i = m_blk * NB_BM + r;                   // global query row (may be >= B in the last block)
    const int pos_lo = p.nf > 0 ? (i / p.nf) * p.nf : -1;
    const int pos_hi = p.nf > 0 ? pos_lo + p.nf : -1;
    const float c = p.scale_log2;
    const float off = (i < p.B) ? -p.row_lse[2 * i] * 1.4426950408889634f : 0.f;      // -m_i * log2(e)
    float acc_d[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) acc_d[d] = 0.f;
    int stage = 0;
    for (int t = t_begin, lt = 0; t < t_end; ++t, ++lt) {
      const int acc = lt & 1;
      const uint32_t acc_phase = (lt >> 1) & 1;
      const bool is_key = t < p.nkt;
      const int j0 = (is_key ? t : t - p.nkt) * NB_BN;
      const int limit = is_key ? p.Bk : p.K;
      const bool needs_mask = is_key || (j0 + NB_BN > limit);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t st = smem_u32(stages + (size_t)stage * stage_bytes);
#pragma unroll 1
      for (int chunk = 0; chunk < NB_BN / 32; ++chunk) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * NB_BN + chunk * 32, raw);
        tmem_ld_wait();
        if (chunk == NB_BN / 32 - 1) {
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty[acc]);
        }
        float e[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          e[x] = bwd_exp2(fmaf(__uint_as_float(raw[x]), c, off));
          if (needs_mask) {
            const int j = j0 + chunk * 32 + x;
            const bool is_pos = is_key && j >= pos_lo && j < pos_hi;
            if (j >= limit || is_pos) e[x] = 0.f;
          }
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const int j = chunk * 32 + x;                // row of the tile == column of the similarity matrix
          const float pj = e[x];
#pragma unroll
          for (int d4 = 0; d4 < DH / 4; ++d4) {
            const int d = half * DH + d4 * 4;          // feature index; atom d / 32, 16-byte chunk (d % 32) / 4
            const float4 n = lds_v4(st + (uint32_t)((d >> 5) * (NB_BN * 128) + j * 128 + ((((d & 31) >> 2) ^ (j & 7)) << 4)));
            acc_d[4 * d4 + 0] = fmaf(pj, n.x, acc_d[4 * d4 + 0]);
            acc_d[4 * d4 + 1] = fmaf(pj, n.y, acc_d[4 * d4 + 1]);
            acc_d[4 * d4 + 2] = fmaf(pj, n.z, acc_d[4 * d4 + 2]);
            acc_d[4 * d4 + 3] = fmaf(pj, n.w, acc_d[4 * d4 + 3]);
          }
        }
      }
      mbar_arrive(&empty_bar[stage]);                  // this thread is done with the shared-memory tile
      if (++stage == p.num_stages) stage = 0;
    }
    float* dst = p.partial + (((size_t)m_blk * p.slices + sidx) * NB_BM + r) * D + half * DH;
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4)
      *reinterpret_cast<float4*>(dst + 4 * d4) =
          make_float4(acc_d[4 * d4], acc_d[4 * d4 + 1], acc_d[4 * d4 + 2], acc_d[4 * d4 + 3]);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * NB_BN);
  }
}

struct NceBwdFinalizeParams {
  const float* keys;
  const float* partial;
  const float* pos_sim;      // [B][nP] raw positive similarities (forward)
  const float* row_lse;      // [B][2]
  int B, D, nf, nP, nmb, slices;
  float temperature, g;      // g = grad_dist / (B * nP)
  int accumulate;            // add to dq instead of overwriting
  float* dq;
};

// one warp per query row: fixed-order reduction over the slices, scaling, positives with exact fp32 keys
__global__ void __launch_bounds__(256) infonce_bwd_finalize_kernel(const NceBwdFinalizeParams p) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= p.B) return;
  const int m_blk = i / NB_BM, r = i % NB_BM;
  const float m = p.row_lse[2 * i], Zn = p.row_lse[2 * i + 1];
  const int pos0 = p.nf > 0 ? (i / p.nf) * p.nf : i;
  float R = 0.f;
  float coef[8];
  for (int pp = 0; pp < p.nP; ++pp) {
    const float es = expf(p.pos_sim[(size_t)i * p.nP + pp] / p.temperature - m);
    const float denom = es + Zn;
    R += 1.f / denom;
    coef[pp] = -p.g * (1.f - es / denom);               // dL/dz_ip
  }
  const float gneg = p.g * R;
  for (int d = lane; d < p.D; d += 32) {
    float a = 0.f;
    for (int s = 0; s < p.slices; ++s) a += p.partial[(((size_t)m_blk * p.slices + s) * NB_BM + r) * p.D + d];
    float v = gneg * a;
    for (int pp = 0; pp < p.nP; ++pp) v = fmaf(coef[pp], p.keys[(size_t)(pos0 + pp) * p.D + d], v);
    v /= p.temperature;
    float* o = p.dq + (size_t)i * p.D + d;
    *o = p.accumulate ? *o + v : v;
  }
}

// self-batch loss: column role of the queries.  dq_j += (1/T) * sum_i dL/dz_ij * q_i, one block per column j.
__global__ void __launch_bounds__(128) infonce_bwd_symmetric_kernel(const float* __restrict__ q,
                                                                    const float* __restrict__ pos_sim,
                                                                    const float* __restrict__ row_lse, int B, int D,
                                                                    int nf, float temperature, float g,
                                                                    float* __restrict__ dq) {
  extern __shared__ float sh[];                         // [D] q_j | [blockDim] coefficients
  float* qj = sh;
  float* w = sh + D;
  const int j = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) qj[d] = q[(size_t)j * D + d];
  float accd[4] = {0.f, 0.f, 0.f, 0.f};                 // this thread owns features threadIdx.x + 128*k (D <= 512)
  for (int i0 = 0; i0 < B; i0 += blockDim.x) {
    __syncthreads();
    const int i = i0 + threadIdx.x;
    float gij = 0.f;
    if (i < B) {
      const float m = row_lse[2 * i], Zn = row_lse[2 * i + 1];
      float R = 0.f;
      for (int pp = 0; pp < nf; ++pp) R += 1.f / (expf(pos_sim[(size_t)i * nf + pp] / temperature - m) + Zn);
      float dot = 0.f;
      for (int d = 0; d < D; ++d) dot = fmaf(q[(size_t)i * D + d], qj[d], dot);
      const float es = expf(dot / temperature - m);
      const bool is_pos = (i / nf) == (j / nf);
      gij = is_pos ? -g * (1.f - es / (es + Zn)) : g * R * es;
    }
    w[threadIdx.x] = gij;
    __syncthreads();
    const int n = min((int)blockDim.x, B - i0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = threadIdx.x + k * blockDim.x;
      if (d < D) {
        float a = accd[k];
        for (int ii = 0; ii < n; ++ii) a = fmaf(w[ii], q[(size_t)(i0 + ii) * D + d], a);
        accd[k] = a;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = threadIdx.x + k * blockDim.x;
    if (d < D) dq[(size_t)j * D + d] += accd[k] / temperature;
  }
}

size_t infonce_bwd_workspace_bytes(int B, int D) {
  const size_t nmb = (B + NB_BM - 1) / NB_BM;
  const size_t rounded = 2 * nmb * NB_BM * (size_t)D * sizeof(float);        // q_tf32, keys_tf32 (padded rows)
  const size_t partial = nmb * 148 * NB_BM * (size_t)D * sizeof(float);
  return rounded + partial + 1024;
}

template <int D>
static int launch_bwd_main(const NceBwdParams& kp, int grid, cudaStream_t stream) {
  const size_t q_bytes = (size_t)(D / 32) * NB_BM * 128;
  const size_t stage_bytes = (size_t)(D / 32) * NB_BN * 128;
  const size_t smem = 1024 + q_bytes + (2 * NB_MAX_STAGES + 5) * 8 + 16 + kp.num_stages * stage_bytes;
  VB_CHECK_CUDA(cudaFuncSetAttribute(infonce_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  infonce_bwd_kernel<D><<<grid, NB_THREADS, smem, stream>>>(kp);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

int infonce_bwd_launch(const InfoNceDesc& d, float grad_dist, int symmetric, int accumulate, float* dq,
                       cudaStream_t stream) {
  VB_REQUIRE(d.B > 0 && d.D > 0, "infonce_bwd: empty batch");
  VB_REQUIRE(d.D % 32 == 0 && d.D <= 128, "infonce_bwd: embedding size %d unsupported (multiple of 32, <= 128)", d.D);
  VB_REQUIRE(d.q && d.keys && dq, "infonce_bwd: q / keys / dq null");
  VB_REQUIRE(d.K == 0 || d.queue_tf32, "infonce_bwd: queue pointer null");
  VB_REQUIRE(d.Bk == d.B, "infonce_bwd: keys must have one row per query (Bk=%d, B=%d)", d.Bk, d.B);
  VB_REQUIRE(d.num_frames >= 0 && d.num_frames <= 8, "infonce_bwd: num_frames %d unsupported", d.num_frames);
  VB_REQUIRE(d.num_frames == 0 || d.B % d.num_frames == 0, "infonce_bwd: batch %d not a multiple of num_frames %d", d.B,
             d.num_frames);
  VB_REQUIRE(!symmetric || (d.num_frames > 0 && d.K == 0 && d.keys == d.q),
             "infonce_bwd: symmetric mode is the self-batch loss (keys == q, no queue, num_frames > 0)");
  VB_REQUIRE(d.temperature > 0.f, "infonce_bwd: temperature must be positive");
  VB_REQUIRE(d.pos_sim && d.row_lse, "infonce_bwd: the forward's pos_sim / row_lse are required");
  VB_REQUIRE(d.workspace && (reinterpret_cast<uintptr_t>(d.workspace) & 255) == 0,
             "infonce_bwd: workspace must be non-null and 256-byte aligned");

  const int nmb = (d.B + NB_BM - 1) / NB_BM;
  const bool ibc = d.num_frames > 0;
  uint8_t* ws = reinterpret_cast<uint8_t*>(d.workspace);
  float* q_r = reinterpret_cast<float*>(ws);
  float* k_r = q_r + (size_t)nmb * NB_BM * d.D;
  float* partial = k_r + (size_t)nmb * NB_BM * d.D;

  int rc = round_tf32_launch(d.q, q_r, (int64_t)d.B * d.D, stream);
  if (rc) return rc;
  if (ibc) {
    rc = round_tf32_launch(d.keys, k_r, (int64_t)d.Bk * d.D, stream);
    if (rc) return rc;
  }
  NceBwdParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.B = d.B, kp.Bk = d.Bk, kp.K = d.K, kp.D = d.D, kp.nf = d.num_frames;
  kp.nmb = nmb;
  kp.nkt = ibc ? (d.Bk + NB_BN - 1) / NB_BN : 0;
  kp.nqt = (d.K + NB_BN - 1) / NB_BN;
  const int T = kp.nkt + kp.nqt;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms > 148) sms = 148;
  int slices = sms / nmb;
  if (slices < 1) slices = 1;
  if (slices > T) slices = T;
  kp.slices = slices;
  kp.scale_log2 = (float)(1.4426950408889634 / (double)d.temperature);
  kp.row_lse = d.row_lse;
  kp.partial = partial;
  if (T > 0) {
    rc = encode_tma_2d(&kp.q_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, q_r, d.D, d.B, (uint64_t)d.D * 4, 32, NB_BM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (kp.nkt) {
      rc = encode_tma_2d(&kp.keys_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, k_r, d.D, d.Bk, (uint64_t)d.D * 4, 32, NB_BN,
                         CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    if (kp.nqt) {
      rc = encode_tma_2d(&kp.queue_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.queue_tf32, d.D, d.K, (uint64_t)d.D * 4, 32,
                         NB_BN, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    const size_t q_bytes = (size_t)(d.D / 32) * NB_BM * 128;
    const size_t stage_bytes = (size_t)(d.D / 32) * NB_BN * 128;
    const size_t fixed = 1024 + q_bytes + (2 * NB_MAX_STAGES + 5) * 8 + 16;
    int stages = (int)((224 * 1024 - fixed) / stage_bytes);
    if (stages > NB_MAX_STAGES) stages = NB_MAX_STAGES;
    VB_REQUIRE(stages >= 2, "infonce_bwd: not enough shared memory");
    kp.num_stages = stages;
    const int grid = nmb * slices;
    if (d.D == 32) rc = launch_bwd_main<32>(kp, grid, stream);
    else if (d.D == 64) rc = launch_bwd_main<64>(kp, grid, stream);
    else if (d.D == 96) rc = launch_bwd_main<96>(kp, grid, stream);
    else rc = launch_bwd_main<128>(kp, grid, stream);
    if (rc) return rc;
  }
  NceBwdFinalizeParams fp;
  fp.keys = d.keys, fp.partial = partial, fp.pos_sim = d.pos_sim, fp.row_lse = d.row_lse;
  fp.B = d.B, fp.D = d.D, fp.nf = d.num_frames, fp.nP = ibc ? d.num_frames : 1;
  fp.nmb = nmb, fp.slices = T > 0 ? slices : 0;
  fp.temperature = d.temperature;
  fp.g = grad_dist / ((float)d.B * (float)fp.nP);
  fp.accumulate = accumulate;
  fp.dq = dq;
  infonce_bwd_finalize_kernel<<<(d.B + 7) / 8, 256, 0, stream>>>(fp);
  VB_CHECK_CUDA(cudaGetLastError());
  if (symmetric) {
    const size_t smem = (size_t)(d.D + 128) * sizeof(float);
    infonce_bwd_symmetric_kernel<<<d.B, 128, smem, stream>>>(d.q, d.pos_sim, d.row_lse, d.B, d.D, d.num_frames,
                                                            d.temperature, fp.g, dq);
    VB_CHECK_CUDA(cudaGetLastError());
  }
  return VB_OK;
}

}  // namespace vb
