// Fused InfoNCE forward for sm_100a (replaces vince_model.py:198-250 cat+mm, loss_util.py:7-62 and the
// metric passes of vince_model.py:314-342): the [B, Bk+K] similarity matrix is never written.
//
//   ONE launch (B > 0 and at least one column tile):
//   main loop     : persistent, warp-specialised.  Each CTA owns one 128-row block of queries (resident in
//                   shared memory) and a contiguous range of 128-column tiles of [keys || queue]; tiles are
//                   streamed by TMA, multiplied on the tensor cores (tcgen05 kind::tf32, fp32 accumulate in
//                   double-buffered TMEM) and consumed straight from TMEM by 128 epilogue threads (one query
//                   row each) that keep an online (max, sum-exp) over the NEGATIVE columns.
//   tail          : the LAST CTA of a query block to finish (ticket counter) merges that block's per-CTA partials
//                   (coalesced: one thread per row), evaluates the positives with exact fp32 dot products and emits the
//                   per-positive losses / weights and the saved row statistics; the last CTA overall then reduces the
//                   five scalars the reference computes every step (loss, softmax weight, accuracy, cosine_sim,
//                   cosine_sim_neg_max) in a fixed (deterministic) order.  No finalize / scalars launches.
//
// Operands are fp32 values rounded (to nearest) to TF32 so that the tensor core's truncation of the low 13 mantissa
// bits is exact: the queue by its owner at enqueue time (`queue_tf32`, see StorageQueue); queries and keys by the
// epilogue threads, in place in shared memory, right after their TMA tiles land (no pre-pass, no rounded copies in
// HBM); positives use the un-rounded data.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int NCE_BM = 128;
constexpr int NCE_BN = 128;
constexpr int NCE_THREADS = 192;
constexpr int NCE_MAX_STAGES = 4;

struct NceParams {
  CUtensorMap q_map, keys_map, queue_map;
  int B, Bk, K, D;
  int nf;
  int nmb, nkt, nqt, slices;
  int num_stages;
  float scale_log2;
  float2* partials;
  // fused tail
  unsigned int* counters;          // [nmb + 1], zero at launch: per-query-block tickets, then the block-level ticket
  const float* q;                  // exact fp32 operands for the positives
  const float* keys;
  int nP;
  float temperature;
  float* dists;
  float* weights;
  float* pos_sim;
  float* neg_max;
  float* row_lse;
  float* scalars;
  int debug;                       // timing experiments only (VINCE_B200_NCE_DEBUG bits): results are garbage
};

__device__ __forceinline__ uint32_t rna_tf32(uint32_t v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(__uint_as_float(v)));
  return u;
}
// round a shared-memory tile of fp32 to TF32 in place (128 epilogue threads; element-wise, so the swizzle is irrelevant)
__device__ __forceinline__ void round_tile_tf32(uint8_t* tile, uint32_t bytes, int etid) {
  uint4* v = reinterpret_cast<uint4*>(tile);
  for (uint32_t i = etid; i < bytes / 16; i += 128) {
    uint4 x = v[i];
    x.x = rna_tf32(x.x), x.y = rna_tf32(x.y), x.z = rna_tf32(x.z), x.w = rna_tf32(x.w);
    v[i] = x;
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(NCE_THREADS, 1) infonce_main_kernel(const __grid_constant__ NceParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int atoms = p.D / 32;                          // 128-byte swizzle atoms along D
  const uint32_t q_bytes = atoms * NCE_BM * 128;
  const uint32_t stage_bytes = atoms * NCE_BN * 128;

  uint8_t* q_smem = smem;
  uint8_t* stages = smem + q_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stages + (size_t)p.num_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + NCE_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + NCE_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* q_full = tmem_empty + 2;
  uint64_t* q_ready = q_full + 1;                      // queries rounded to TF32 in place (128 arrivals)
  uint64_t* key_ready = q_ready + 1;                   // one phase per key tile: rounded in place (128 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(key_ready + 1);
  uint32_t* flags = tmem_ptr + 1;                      // [2]: last CTA of the query block / last CTA overall

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
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    mbar_init(q_full, 1);
    mbar_init(q_ready, 128);
    mbar_init(key_ready, 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * NCE_BN);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  // producer / MMA warps: warp-uniform control flow, one elected lane issues (keeps operands in uniform registers)
  if (warp == 0) {
    if (t_end > t_begin) {
      if (elect_one()) {
        mbar_expect_tx(q_full, q_bytes);
        for (int a = 0; a < atoms; ++a) tma_load_2d(q_smem + a * NCE_BM * 128, &p.q_map, q_full, a * 32, m_blk * NCE_BM);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = stages + (size_t)stage * stage_bytes;
        const bool is_key = t < p.nkt;
        const CUtensorMap* map = is_key ? &p.keys_map : &p.queue_map;
        const int row0 = (is_key ? t : t - p.nkt) * NCE_BN;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          for (int a = 0; a < atoms; ++a) tma_load_2d(st + a * NCE_BN * 128, map, &full_bar[stage], a * 32, row0);
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
      constexpr uint32_t idesc = make_idesc(UMMA_FMT_TF32, NCE_BM, NCE_BN);
      mbar_wait(q_ready, 0);
      tc_fence_after_sync();
      const uint32_t q_addr = smem_u32(q_smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin, lt = 0; t < t_end; ++t, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        if (t < p.nkt) mbar_wait(key_ready, (uint32_t)(t - t_begin) & 1u);     // key tiles come first: lt == t - t_begin
        else mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint32_t st = smem_u32(stages + (size_t)stage * stage_bytes);
        const uint32_t d_tmem = tmem_base + acc * NCE_BN;
        if (elect_one()) {
          for (int a = 0; a < atoms; ++a) {
            const uint64_t da0 = make_smem_desc(q_addr + a * NCE_BM * 128, 16, 1024, UMMA_LAYOUT_SW128);
            const uint64_t db0 = make_smem_desc(st + a * NCE_BN * 128, 16, 1024, UMMA_LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < 4; ++k)            // 8 tf32 (32 bytes) per MMA: start address advances by 2 units
              umma_tf32(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, (a | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
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
    // ---- online softmax over negatives, one query row per thread ----
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int i = m_blk * NCE_BM + r;                 // global query row
    const int pos_lo = p.nf > 0 ? (i / p.nf) * p.nf : -1;
    const int pos_hi = p.nf > 0 ? pos_lo + p.nf : -1;
    float nmax = -INFINITY;                           // running max of raw similarities over negatives
    float Z = 0.f;                                    // sum exp2((sigma - nmax) * scale_log2) over negatives
    const float c = p.scale_log2;
    const int etid = threadIdx.x - 64;
    if (t_end > t_begin) {
      // queries: exact fp32 from HBM -> TF32 (round to nearest) in place, then visible to the tensor core (async proxy)
      mbar_wait(q_full, 0);
      round_tile_tf32(q_smem, q_bytes, etid);
      fence_proxy_async_smem();
      mbar_arrive(q_ready);
    }
    // positives: q_i . k_j in exact fp32 (vince_model.py:213-233 computes them in fp32).  The rows of this query block are
    // dealt out over its `slices` CTAs (one or two rows each), and a row is one warp's work: the 32 lanes split the D
    // elements (coalesced 128-byte reads), lane 0 stores the result straight into the pos_sim output, which the block's
    // last CTA reads back in the tail.  Runs under the TMA prologue; doing all 128 rows in the tail instead cost the last
    // CTA ~60 us of serial L2 round trips (ncu: SMs active 27k of 157k elapsed cycles).
    {
      int idx = 0;
      for (int rr = (p.debug & 4) ? NCE_BM : sidx; rr < NCE_BM; rr += p.slices, ++idx) {
        const int ri = m_blk * NCE_BM + rr;
        if ((idx & 3) != quarter || ri >= p.B) continue;          // warp-uniform
        const int pos0 = p.nf > 0 ? (ri / p.nf) * p.nf : ri;
        const float* qr = p.q + (size_t)ri * p.D;
        float qv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) qv[t] = (lane + 32 * t < p.D) ? __ldg(qr + lane + 32 * t) : 0.f;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          if (pp < p.nP) {
            const float* kr = p.keys + (size_t)(pos0 + pp) * p.D;
            float d = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) d = fmaf(qv[t], (lane + 32 * t < p.D) ? __ldg(kr + lane + 32 * t) : 0.f, d);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            if (lane == 0) p.pos_sim[(size_t)ri * p.nP + pp] = d;
          }
        }
      }
    }
    for (int t = t_begin, lt = 0; t < t_end; ++t, ++lt) {
      const int acc = lt & 1;
      const uint32_t acc_phase = (lt >> 1) & 1;
      const bool is_key = t < p.nkt;
      if (is_key) {
        // key tiles (the first nkt tiles) arrive un-rounded too: same in-place treatment before the MMA may read them
        const int stage = lt % p.num_stages;
        mbar_wait(&full_bar[stage], (uint32_t)(lt / p.num_stages) & 1u);
        round_tile_tf32(stages + (size_t)stage * stage_bytes, stage_bytes, etid);
        fence_proxy_async_smem();
        mbar_arrive(key_ready);
      }
      const int j0 = (is_key ? t : t - p.nkt) * NCE_BN;
      const int limit = is_key ? p.Bk : p.K;
      const bool needs_mask = is_key || (j0 + NCE_BN > limit);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int chunk = 0; chunk < NCE_BN / 32; ++chunk) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * NCE_BN + chunk * 32, raw);
        tmem_ld_wait();
        if (chunk == NCE_BN / 32 - 1) {
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty[acc]);
        }
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(raw[e]);
        if (needs_mask) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int j = j0 + chunk * 32 + e;
            const bool is_pos = is_key && j >= pos_lo && j < pos_hi;
            if (j >= limit || is_pos) v[e] = -INFINITY;
          }
        }
        float cm = v[0];
#pragma unroll
        for (int e = 1; e < 32; ++e) cm = fmaxf(cm, v[e]);
        if (cm > nmax) {
          Z *= fast_exp2((nmax - cm) * c);            // nmax = -inf on first use: Z is 0, exp2(-inf) = 0
          nmax = cm;
        }
        const float off = (nmax == -INFINITY) ? 0.f : -nmax * c;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          s0 += fast_exp2(fmaf(v[e], c, off));
          s1 += fast_exp2(fmaf(v[e + 1], c, off));
        }
        Z += s0 + s1;
      }
    }
    p.partials[((size_t)m_blk * p.slices + sidx) * NCE_BM + r] = make_float2(nmax, Z);

    // ---------------- fused tail: last CTA of this query block finalizes its 128 rows ----------------
    if (p.debug & 8) goto tail_done;
    __threadfence();
    named_bar_sync(1, 128);
    if (etid == 0) flags[0] = (atomicAdd(&p.counters[m_blk], 1u) == (unsigned)p.slices - 1u) ? 1u : 0u;
    named_bar_sync(1, 128);
    if (flags[0]) {
      __threadfence();
      if (i < p.B && !(p.debug & 1)) {
        // merge the per-CTA partials of this row (coalesced across the block's threads; unrolled so that several of the
        // L2 round trips are in flight at once)
        const float2* part = p.partials + (size_t)m_blk * p.slices * NCE_BM + r;
        // one pass, 16 partials at a time held in registers so that 16 L2 round trips are in flight together (a plain
        // loop, even unrolled, issued them one after another: 31 us for the 2 x 74 loads of a row)
        float gmax = -INFINITY, Zs = 0.f;
        for (int s0 = 0; s0 < p.slices; s0 += 16) {
          float2 v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u)
            v[u] = (s0 + u < p.slices) ? __ldcg(&part[(size_t)(s0 + u) * NCE_BM]) : make_float2(-INFINITY, 0.f);
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            if (v[u].x != -INFINITY) {
              const float mn = fmaxf(gmax, v[u].x);
              Zs = Zs * exp2f((gmax - mn) * c) + v[u].y * exp2f((v[u].x - mn) * c);     // gmax = -inf: Zs is 0, exp2(-inf) = 0
              gmax = mn;
            }
          }
        }
        // positives: computed (exact fp32) at the start of the kernel by the CTAs of this block, see above
        float sig[8];
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) sig[pp] = (pp < p.nP) ? __ldcg(&p.pos_sim[(size_t)i * p.nP + pp]) : 0.f;
        float zmax = (gmax == -INFINITY) ? -INFINITY : gmax / p.temperature;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
          if (pp < p.nP) zmax = fmaxf(zmax, sig[pp] / p.temperature);
        // Zneg relative to the row max over ALL columns (loss_util.py:24)
        const float Zn = (gmax == -INFINITY) ? 0.f : Zs * expf(gmax / p.temperature - zmax);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          if (pp < p.nP) {
            const float sp = sig[pp] / p.temperature - zmax;
            const float logsm = sp - logf(expf(sp) + Zn);
            p.dists[(size_t)i * p.nP + pp] = -logsm;
            p.weights[(size_t)i * p.nP + pp] = expf(logsm);
          }
        }
        p.neg_max[i] = gmax;
        p.row_lse[2 * i] = zmax;
        p.row_lse[2 * i + 1] = Zn;
      }
      // ---------------- last CTA overall: the five scalars, fixed reduction order ----------------
      __threadfence();
      named_bar_sync(1, 128);
      if (etid == 0) flags[1] = (atomicAdd(&p.counters[p.nmb], 1u) == (unsigned)p.nmb - 1u) ? 1u : 0u;
      named_bar_sync(1, 128);
      if (flags[1] && !(p.debug & 2)) {
        __threadfence();
        float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int row = etid; row < p.B; row += 128) {
          const float nm = __ldcg(&p.neg_max[row]);
          for (int pp = 0; pp < p.nP; ++pp) {
            a[0] += __ldcg(&p.dists[(size_t)row * p.nP + pp]);
            a[1] += __ldcg(&p.weights[(size_t)row * p.nP + pp]);
            const float ps = __ldcg(&p.pos_sim[(size_t)row * p.nP + pp]);
            a[2] += ps > nm ? 1.f : 0.f;
            a[3] += ps;
          }
          a[4] += nm;
        }
        float* red = reinterpret_cast<float*>(stages);      // the operand ring is idle by now: [5][4] warp partials
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
          if (lane == 0) red[k * 4 + quarter] = a[k];
        }
        named_bar_sync(1, 128);
        if (etid < 5) {
          const float tot = (red[etid * 4 + 0] + red[etid * 4 + 1]) + (red[etid * 4 + 2] + red[etid * 4 + 3]);
          p.scalars[etid] = tot / (etid == 4 ? (float)p.B : (float)p.B * (float)p.nP);
        }
      }
    }
  tail_done:;
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * NCE_BN);
  }
}

// round-to-nearest fp32 -> tf32 (kept in an fp32 container)
__global__ void round_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x[i]));
    out[i] = __uint_as_float(u);
  }
}

int round_tf32_launch(const float* x, float* out, int64_t n, cudaStream_t stream) {
  if (n == 0) return VB_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  round_tf32_kernel<<<(int)blocks, 256, 0, stream>>>(x, out, n);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

struct NceFinalizeParams {
  const float* q;
  const float* keys;
  const float2* partials;
  int B, D, nf, nP, nmb, slices;
  float temperature, scale_log2;
  float* dists;
  float* weights;
  float* pos_sim;
  float* neg_max;
  float* row_lse;
  float* scalars;
};

// one warp per query row; the five scalar means are produced afterwards by nce_scalars_kernel in a fixed order
__global__ void __launch_bounds__(256) infonce_finalize_kernel(const NceFinalizeParams p) {
  const int lane = threadIdx.x & 31;
  {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= p.B) return;
    const int m_blk = i / NCE_BM, r = i % NCE_BM;
    // merge partials over slices (lanes stride over slices)
    float nmax = -INFINITY;
    for (int s = lane; s < p.slices; s += 32)
      nmax = fmaxf(nmax, p.partials[((size_t)m_blk * p.slices + s) * NCE_BM + r].x);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
    float Z = 0.f;
    for (int s = lane; s < p.slices; s += 32) {
      const float2 pr = p.partials[((size_t)m_blk * p.slices + s) * NCE_BM + r];
      if (pr.x != -INFINITY) Z += pr.y * exp2f((pr.x - nmax) * p.scale_log2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(0xffffffffu, Z, o);
    // positives: exact fp32 dot products
    const int pos0 = p.nf > 0 ? (i / p.nf) * p.nf : i;
    float zmax = (nmax == -INFINITY) ? -INFINITY : nmax / p.temperature;
    float sig[8];
    for (int pp = 0; pp < p.nP; ++pp) {
      const float* kr = p.keys + (size_t)(pos0 + pp) * p.D;
      const float* qr = p.q + (size_t)i * p.D;
      float d = 0.f;
      for (int e = lane; e < p.D; e += 32) d = fmaf(qr[e], kr[e], d);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      sig[pp] = d;
      zmax = fmaxf(zmax, d / p.temperature);
    }
    // Zneg relative to the row max over ALL columns (loss_util.py:24)
    const float Zn = (nmax == -INFINITY) ? 0.f : Z * expf(nmax / p.temperature - zmax);
    for (int pp = 0; pp < p.nP; ++pp) {
      const float s = sig[pp] / p.temperature - zmax;
      const float logsm = s - logf(expf(s) + Zn);
      if (lane == 0) {
        p.dists[(size_t)i * p.nP + pp] = -logsm;
        p.weights[(size_t)i * p.nP + pp] = expf(logsm);
        p.pos_sim[(size_t)i * p.nP + pp] = sig[pp];
      }
    }
    if (lane == 0) {
      p.neg_max[i] = nmax;
      p.row_lse[2 * i] = zmax;
      p.row_lse[2 * i + 1] = Zn;
    }
  }
}

__global__ void nce_scalars_kernel(const float* __restrict__ dists, const float* __restrict__ weights,
                                   const float* __restrict__ pos_sim, const float* __restrict__ neg_max, int R, int nP,
                                   float* __restrict__ scalars);

size_t infonce_workspace_bytes(int B, int D) {
  const size_t nmb = (B + NCE_BM - 1) / NCE_BM;
  // [partials: nmb x 148 slices x 128 rows float2][ticket counters: nmb + 1]; sized as in ABI v1-v3, which also kept
  // TF32-rounded copies of q / keys here (now rounded in shared memory by the kernel itself)
  const size_t rounded = 2 * nmb * NCE_BM * (size_t)D * sizeof(float);
  const size_t partials = nmb * 148 * NCE_BM * sizeof(float2);
  return rounded + partials + 1024;
}

int infonce_fwd_launch(const InfoNceDesc& d, cudaStream_t stream) {
  VB_REQUIRE(d.B > 0 && d.D > 0, "infonce: empty batch");
  VB_REQUIRE(d.D % 32 == 0 && d.D <= 128, "infonce: embedding size %d unsupported (multiple of 32, <= 128)", d.D);
  VB_REQUIRE(d.q && d.keys, "infonce: q / keys null");
  VB_REQUIRE(d.K == 0 || d.queue_tf32, "infonce: queue pointer null");
  VB_REQUIRE(d.Bk == d.B, "infonce: keys must have one row per query (Bk=%d, B=%d)", d.Bk, d.B);
  VB_REQUIRE(d.num_frames >= 0 && d.num_frames <= 8, "infonce: num_frames %d unsupported", d.num_frames);
  VB_REQUIRE(d.num_frames == 0 || d.B % d.num_frames == 0, "infonce: batch %d not a multiple of num_frames %d", d.B,
             d.num_frames);
  VB_REQUIRE(d.temperature > 0.f, "infonce: temperature must be positive");
  VB_REQUIRE(d.workspace, "infonce: workspace null");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(d.workspace) & 255) == 0, "infonce: workspace must be 256-byte aligned");

  const int nmb = (d.B + NCE_BM - 1) / NCE_BM;
  const bool ibc = d.num_frames > 0;
  VB_REQUIRE((reinterpret_cast<uintptr_t>(d.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.keys) & 15) == 0,
             "infonce: q / keys must be 16-byte aligned (TMA)");
  uint8_t* ws = reinterpret_cast<uint8_t*>(d.workspace);
  float2* partials = reinterpret_cast<float2*>(ws);
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws + (size_t)nmb * 148 * NCE_BM * sizeof(float2));

  int rc = VB_OK;
  NceParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.B = d.B, kp.Bk = d.Bk, kp.K = d.K, kp.D = d.D, kp.nf = d.num_frames;
  kp.nmb = nmb;
  kp.nkt = ibc ? (d.Bk + NCE_BN - 1) / NCE_BN : 0;
  kp.nqt = (d.K + NCE_BN - 1) / NCE_BN;
  const int T = kp.nkt + kp.nqt;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms > 148) sms = 148;                            // partials workspace is sized for <= 148 slices
  int slices = sms / nmb;
  if (slices < 1) slices = 1;
  if (slices > T) slices = T;
  kp.slices = slices;
  kp.scale_log2 = (float)(1.4426950408889634 / (double)d.temperature);
  kp.partials = partials;
  kp.counters = counters;
  kp.q = d.q, kp.keys = d.keys;
  kp.nP = ibc ? d.num_frames : 1;
  kp.temperature = d.temperature;
  kp.dists = d.dists, kp.weights = d.weights, kp.pos_sim = d.pos_sim, kp.neg_max = d.neg_max, kp.row_lse = d.row_lse;
  kp.scalars = d.scalars;
  kp.debug = getenv("VINCE_B200_NCE_DEBUG") ? atoi(getenv("VINCE_B200_NCE_DEBUG")) : 0;

  if (T > 0) {
    rc = encode_tma_2d(&kp.q_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.q, d.D, d.B, (uint64_t)d.D * 4, 32, NCE_BM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (kp.nkt) {
      rc = encode_tma_2d(&kp.keys_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.keys, d.D, d.Bk, (uint64_t)d.D * 4, 32, NCE_BN,
                         CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    if (kp.nqt) {
      rc = encode_tma_2d(&kp.queue_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.queue_tf32, d.D, d.K, (uint64_t)d.D * 4, 32,
                         NCE_BN, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    const size_t q_bytes = (size_t)(d.D / 32) * NCE_BM * 128;
    const size_t stage_bytes = (size_t)(d.D / 32) * NCE_BN * 128;
    const size_t fixed = 1024 + q_bytes + (2 * NCE_MAX_STAGES + 7) * 8 + 16;
    int stages = (int)((227 * 1024 - fixed) / stage_bytes);
    if (stages > NCE_MAX_STAGES) stages = NCE_MAX_STAGES;
    VB_REQUIRE(stages >= 2, "infonce: not enough shared memory");
    kp.num_stages = stages;
    const size_t smem = fixed + stages * stage_bytes;
    static bool smem_set[64] = {false};                 // the attribute is per device / context
    if (dev < 0 || dev >= 64 || !smem_set[dev]) {
      VB_CHECK_CUDA(cudaFuncSetAttribute(infonce_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      if (dev >= 0 && dev < 64) smem_set[dev] = true;
    }
    VB_CHECK_CUDA(cudaMemsetAsync(counters, 0, (size_t)(nmb + 1) * sizeof(unsigned int), stream));
    infonce_main_kernel<<<nmb * slices, NCE_THREADS, smem, stream>>>(kp);
    VB_CHECK_CUDA(cudaGetLastError());
    return VB_OK;                                       // finalize + scalars happen in the kernel's tail
  }
  // no column tile at all (K == 0 without inter-batch comparison): positives only

  NceFinalizeParams fp;
  fp.q = d.q, fp.keys = d.keys, fp.partials = partials;
  fp.B = d.B, fp.D = d.D, fp.nf = d.num_frames, fp.nP = ibc ? d.num_frames : 1;
  fp.nmb = nmb, fp.slices = 0;
  fp.temperature = d.temperature, fp.scale_log2 = kp.scale_log2;
  fp.dists = d.dists, fp.weights = d.weights, fp.pos_sim = d.pos_sim, fp.neg_max = d.neg_max, fp.row_lse = d.row_lse;
  fp.scalars = d.scalars;
  infonce_finalize_kernel<<<(d.B + 7) / 8, 256, 0, stream>>>(fp);
  VB_CHECK_CUDA(cudaGetLastError());
  nce_scalars_kernel<<<1, 256, 0, stream>>>(d.dists, d.weights, d.pos_sim, d.neg_max, d.B, fp.nP, d.scalars);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb

// ------------------------------------------------------------------------------------------------
// Generic masked cross entropy over an EXPLICIT similarity matrix (the literal signature of
// loss_util.similarity_cross_entropy, utils/loss_util.py:7-62, equal-count branch :35-38, plus the metric
// quantities of vince_model.py:327-333).  Not on the solver's path (the fused kernel above is); one block per
// row, two coalesced passes.  Positives are emitted in column order, like the reference's boolean gather.
// ------------------------------------------------------------------------------------------------
namespace vb {

constexpr int MCE_MAX_POS = 64;

__global__ void __launch_bounds__(256) masked_ce_kernel(const float* __restrict__ sims,
                                                        const uint8_t* __restrict__ mask, int R, int C, int nP,
                                                        float temperature, float* __restrict__ dists,
                                                        float* __restrict__ weights, float* __restrict__ pos_sim,
                                                        float* __restrict__ neg_max, float* __restrict__ row_lse,
                                                        int* __restrict__ error_flag) {
  __shared__ float red_a[8], red_b[8];
  __shared__ int pos_col[MCE_MAX_POS];
  __shared__ float pos_val[MCE_MAX_POS];
  __shared__ int pos_count;
  const int row = blockIdx.x;
  const float* s = sims + (size_t)row * C;
  const uint8_t* m = mask + (size_t)row * C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) pos_count = 0;
  __syncthreads();
  float zmax = -INFINITY, nmax = -INFINITY;
  for (int j = tid; j < C; j += blockDim.x) {
    const float v = s[j];
    zmax = fmaxf(zmax, v / temperature);
    if (m[j]) {
      const int slot = atomicAdd(&pos_count, 1);
      if (slot < MCE_MAX_POS) pos_col[slot] = j, pos_val[slot] = v;
    } else {
      nmax = fmaxf(nmax, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  }
  if (lane == 0) red_a[warp] = zmax, red_b[warp] = nmax;
  __syncthreads();
  zmax = red_a[0], nmax = red_b[0];
  for (int w = 1; w < 8; ++w) zmax = fmaxf(zmax, red_a[w]), nmax = fmaxf(nmax, red_b[w]);
  __syncthreads();
  float z = 0.f;
  for (int j = tid; j < C; j += blockDim.x)
    if (!m[j]) z += expf(s[j] / temperature - zmax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  if (lane == 0) red_a[warp] = z;
  __syncthreads();
  if (tid == 0) {
    float Zn = 0.f;
    for (int w = 0; w < 8; ++w) Zn += red_a[w];
    if (pos_count != nP) {
      atomicExch(error_flag, 1);      // rows with different positive counts: the reference's USE_FLOAT branch
    } else {
      // insertion sort by column (nP is tiny)
      for (int a = 1; a < nP; ++a) {
        const int c = pos_col[a];
        const float v = pos_val[a];
        int b = a - 1;
        while (b >= 0 && pos_col[b] > c) {
          pos_col[b + 1] = pos_col[b], pos_val[b + 1] = pos_val[b];
          --b;
        }
        pos_col[b + 1] = c, pos_val[b + 1] = v;
      }
      for (int pp = 0; pp < nP; ++pp) {
        const float sp = pos_val[pp] / temperature - zmax;
        const float logsm = sp - logf(expf(sp) + Zn);
        dists[(size_t)row * nP + pp] = -logsm;
        weights[(size_t)row * nP + pp] = expf(logsm);
        pos_sim[(size_t)row * nP + pp] = pos_val[pp];
      }
    }
    neg_max[row] = nmax;
    row_lse[2 * row] = zmax;
    row_lse[2 * row + 1] = Zn;
  }
}

// deterministic means of the per-row results -> scalars[0..4] (same slots as the fused kernel)
__global__ void __launch_bounds__(256) nce_scalars_kernel(const float* __restrict__ dists,
                                                          const float* __restrict__ weights,
                                                          const float* __restrict__ pos_sim,
                                                          const float* __restrict__ neg_max, int R, int nP,
                                                          float* __restrict__ scalars) {
  __shared__ float red[5][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = tid; i < R; i += blockDim.x) {
    const float nm = neg_max[i];
    for (int pp = 0; pp < nP; ++pp) {
      a[0] += dists[(size_t)i * nP + pp];
      a[1] += weights[(size_t)i * nP + pp];
      const float ps = pos_sim[(size_t)i * nP + pp];
      a[2] += ps > nm ? 1.f : 0.f;
      a[3] += ps;
    }
    a[4] += nm;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) red[k][warp] = a[k];
  }
  __syncthreads();
  if (tid < 5) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[tid][w];
    scalars[tid] = s / (tid == 4 ? (float)R : (float)R * (float)nP);
  }
}

int masked_ce_launch(const float* sims, const uint8_t* mask, int R, int C, int nP, float temperature, float* dists,
                     float* weights, float* pos_sim, float* neg_max, float* row_lse, float* scalars, int* error_flag,
                     cudaStream_t stream) {
  VB_REQUIRE(R > 0 && C > 0, "masked_ce: empty matrix");
  VB_REQUIRE(nP >= 1 && nP <= MCE_MAX_POS, "masked_ce: positives per row must be in [1, %d], got %d", MCE_MAX_POS, nP);
  VB_REQUIRE(temperature > 0.f, "masked_ce: temperature must be positive");
  masked_ce_kernel<<<R, 256, 0, stream>>>(sims, mask, R, C, nP, temperature, dists, weights, pos_sim, neg_max, row_lse,
                                          error_flag);
  VB_CHECK_CUDA(cudaGetLastError());
  nce_scalars_kernel<<<1, 256, 0, stream>>>(dists, weights, pos_sim, neg_max, R, nP, scalars);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb
