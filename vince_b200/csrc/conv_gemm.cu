// Persistent warp-specialised implicit-GEMM convolution / linear kernel for sm_100a.
//
//   out[M, N] = epilogue( A[M, K] * W[N, K]^T )         fp32 accumulate in TMEM
//
// * A is either a plain row-major [M, K] matrix (1x1 stride-1 convs, Linear layers; 2-D tiled TMA) or an
//   NHWC activation tensor gathered on the fly by TMA *im2col* loads (KxK / strided convs): k-block kb
//   maps to filter tap (r, s) = (kb / cin_blocks) and input-channel block (kb % cin_blocks).
// * Operands are bf16.  To reproduce the reference's fp32 arithmetic (torch conv2d / mm, SURVEY.md 7
//   "hard parts") every fp32 value is carried as a bf16 (hi, lo) pair and each k-step issues three MMAs
//   hi*hi + lo*hi + hi*lo into the same accumulator ("bf16x3", passes = 3).  passes = 1 is plain bf16.
// * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
//   warps 2..5 = epilogue (TMEM -> registers -> swizzled smem -> TMA store; per-channel sum / sum-of-squares
//   for train-mode BatchNorm accumulated in fp64).  Accumulators are double-buffered in TMEM so the epilogue
//   of tile t overlaps the main loop of tile t+1.  One CTA per SM; tiles are handed out round-robin with the
//   M index fastest so that concurrently running CTAs share the weight tile through L2.
#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int BM = 128;          // rows (output pixels) per tile  == UMMA M == TMEM lanes
constexpr int BK = 64;           // bf16 elements per k-block       == one 128-byte swizzle row
constexpr int UMMA_K = 16;       // bf16
constexpr int GEMM_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int STAGING_BYTES = BM * 128;   // 128 rows x 32 fp32

struct GemmKernelParams {
  CUtensorMap a_hi, a_lo, b_hi, b_lo, out;
  int M, N;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int a_mode;                    // 0 tiled [M,K]; 1 im2col NHWC
  int PQ, Q, stride, pad_h, pad_w, S, cin_blocks;
  int passes;                    // 1 or 3
  int num_stages;
  const float* scale;            // optional per-channel multiplier
  const float* bias;             // optional per-channel addend
  int relu;
  double* stats;                 // optional [2][N]: sum, sum of squares over rows (raw accumulators)
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_gemm_kernel(const __grid_constant__ GemmKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int planes = (p.passes == 3) ? 2 : 1;
  constexpr uint32_t A_TILE = BM * 128;            // bytes per plane
  constexpr uint32_t B_TILE = BN * 128;
  const uint32_t stage_bytes = (A_TILE + B_TILE) * planes;

  uint8_t* stages = smem;
  uint8_t* staging = smem + (size_t)p.num_stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  double* smem_stats = reinterpret_cast<double*>(tmem_ptr + 2);     // [2][BN]

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_hi);
    tma_prefetch_desc(&p.b_hi);
    tma_prefetch_desc(&p.out);
    if (planes == 2) {
      tma_prefetch_desc(&p.a_lo);
      tma_prefetch_desc(&p.b_lo);
    }
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * BN);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < 2 * BN; i += 128) smem_stats[i] = 0.0;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.num_m_blocks;
        const int n_blk = tile / p.num_m_blocks;
        const int m0 = m_blk * BM;
        int img = 0, ph = 0, qw = 0;
        if (p.a_mode == 1) {
          img = m0 / p.PQ;
          const int rem = m0 - img * p.PQ;
          const int p0 = rem / p.Q;
          const int q0 = rem - p0 * p.Q;
          ph = p0 * p.stride - p.pad_h;
          qw = q0 * p.stride - p.pad_w;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stages + (size_t)stage * stage_bytes;
          uint8_t* a_hi = st;
          uint8_t* a_lo = st + A_TILE;
          uint8_t* b_hi = st + A_TILE * planes;
          uint8_t* b_lo = b_hi + B_TILE;
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          if (p.a_mode == 0) {
            tma_load_2d(a_hi, &p.a_hi, &full_bar[stage], kb * BK, m0);
            if (planes == 2) tma_load_2d(a_lo, &p.a_lo, &full_bar[stage], kb * BK, m0);
          } else {
            const int tap = kb / p.cin_blocks;
            const int cb = kb - tap * p.cin_blocks;
            const int r = tap / p.S;
            const int s = tap - r * p.S;
            tma_load_im2col_4d(a_hi, &p.a_hi, &full_bar[stage], cb * BK, qw, ph, img, (uint16_t)s, (uint16_t)r);
            if (planes == 2)
              tma_load_im2col_4d(a_lo, &p.a_lo, &full_bar[stage], cb * BK, qw, ph, img, (uint16_t)s, (uint16_t)r);
          }
          tma_load_2d(b_hi, &p.b_hi, &full_bar[stage], kb * BK, n_blk * BN);
          if (planes == 2) tma_load_2d(b_lo, &p.b_lo, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(UMMA_FMT_BF16, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local_t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_t) {
        const int acc = local_t & 1;
        const uint32_t acc_phase = (local_t >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t st = smem_u32(stages + (size_t)stage * stage_bytes);
          const uint32_t a_hi = st, a_lo = st + A_TILE;
          const uint32_t b_hi = st + A_TILE * planes, b_lo = b_hi + B_TILE;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;   // bytes inside the 128-byte swizzle row
            const uint64_t da_hi = make_smem_desc(a_hi + koff, 16, 1024, UMMA_LAYOUT_SW128);
            const uint64_t db_hi = make_smem_desc(b_hi + koff, 16, 1024, UMMA_LAYOUT_SW128);
            if (planes == 2) {
              const uint64_t da_lo = make_smem_desc(a_lo + koff, 16, 1024, UMMA_LAYOUT_SW128);
              const uint64_t db_lo = make_smem_desc(b_lo + koff, 16, 1024, UMMA_LAYOUT_SW128);
              // small cross terms first, dominant term last
              umma_bf16(d_tmem, da_lo, db_hi, idesc, (kb | k) != 0);
              umma_bf16(d_tmem, da_hi, db_lo, idesc, 1);
              umma_bf16(d_tmem, da_hi, db_hi, idesc, 1);
            } else {
              umma_bf16(d_tmem, da_hi, db_hi, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);              // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue (warps 2..5) =======================
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - 64;             // 0..127
    const bool store_leader = (etid == 0);
    int local_t = 0;
    int cur_n_blk = -1;
    bool store_pending = false;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_t) {
      const int m_blk = tile % p.num_m_blocks;
      const int n_blk = tile / p.num_m_blocks;
      const int acc = local_t & 1;
      const uint32_t acc_phase = (local_t >> 1) & 1;
      if (p.stats != nullptr && n_blk != cur_n_blk) {
        if (cur_n_blk >= 0) {
          named_bar_sync(1, 128);
          for (int i = etid; i < BN; i += 128) {
            if (cur_n_blk * BN + i < p.N) {
              atomicAdd(&p.stats[cur_n_blk * BN + i], smem_stats[i]);
              atomicAdd(&p.stats[p.N + cur_n_blk * BN + i], smem_stats[BN + i]);
            }
            smem_stats[i] = 0.0;
            smem_stats[BN + i] = 0.0;
          }
          named_bar_sync(1, 128);
        }
        cur_n_blk = n_blk;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + chunk * 32, raw);
        tmem_ld_wait();
        if (chunk == BN / 32 - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty[acc]);
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const int c0 = n_blk * BN + chunk * 32;
        if (p.stats != nullptr) {
          // butterfly transpose-reduce: afterwards lane L holds the sum over the warp's 32 rows of channel c0+L
          float s1[32], s2[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) s1[i] = v[i], s2[i] = v[i] * v[i];
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send1 = upper ? s1[i] : s1[i + off];
              const float keep1 = upper ? s1[i + off] : s1[i];
              s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
              const float send2 = upper ? s2[i] : s2[i + off];
              const float keep2 = upper ? s2[i + off] : s2[i];
              s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
            }
          }
          atomicAdd(&smem_stats[chunk * 32 + lane], (double)s1[0]);
          atomicAdd(&smem_stats[BN + chunk * 32 + lane], (double)s2[0]);
        }
        if (p.scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= (c0 + i < p.N) ? __ldg(p.scale + c0 + i) : 0.f;
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += (c0 + i < p.N) ? __ldg(p.bias + c0 + i) : 0.f;
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        // staging buffer must have been drained by the previous TMA store
        if (store_leader && store_pending) tma_store_wait_read0();
        named_bar_sync(2, 128);
        {
          // 128-byte row `row`, 16-byte chunk j stored at (j ^ (row & 7)) : SWIZZLE_128B, conflict-free
          uint8_t* rowp = staging + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 f = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            *reinterpret_cast<float4*>(rowp + ((j ^ (row & 7)) << 4)) = f;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (store_leader) {
          tma_store_2d(&p.out, staging, c0, m_blk * BM);
          tma_store_commit();
        }
        store_pending = true;
      }
    }
    if (p.stats != nullptr && cur_n_blk >= 0) {
      named_bar_sync(1, 128);
      for (int i = etid; i < BN; i += 128) {
        if (cur_n_blk * BN + i < p.N) {
          atomicAdd(&p.stats[cur_n_blk * BN + i], smem_stats[i]);
          atomicAdd(&p.stats[p.N + cur_n_blk * BN + i], smem_stats[BN + i]);
        }
      }
    }
    if (store_leader) tma_store_wait0();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static size_t gemm_smem_bytes(int bn, int passes, int* stages_out) {
  const size_t planes = passes == 3 ? 2 : 1;
  const size_t stage = (size_t)(BM * 128 + bn * 128) * planes;
  const size_t fixed = 1024 /*align slack*/ + STAGING_BYTES + (2 * MAX_STAGES + 4) * 8 + 16 + 2 * bn * 8;
  const size_t budget = 227 * 1024;
  int stages = (int)((budget - fixed) / stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  *stages_out = stages;
  return fixed + stage * stages;
}

template <int BN>
static int launch_gemm(const GemmKernelParams& kp, int stages, size_t smem, int grid, cudaStream_t stream) {
  VB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_gemm_kernel<BN><<<grid, GEMM_THREADS, smem, stream>>>(kp);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

int conv_gemm_launch(const ConvGemmDesc& d, cudaStream_t stream) {
  VB_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "conv_gemm: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
  VB_REQUIRE(d.K % BK == 0, "conv_gemm: K=%d must be a multiple of %d", d.K, BK);
  VB_REQUIRE(d.N % 32 == 0, "conv_gemm: N=%d must be a multiple of 32", d.N);
  VB_REQUIRE(d.passes == 1 || d.passes == 3, "conv_gemm: passes must be 1 or 3");
  VB_REQUIRE(d.a_hi && d.b_hi && d.out, "conv_gemm: null operand");
  VB_REQUIRE(d.passes == 1 || (d.a_lo && d.b_lo), "conv_gemm: bf16x3 needs lo planes");
  VB_REQUIRE(!(d.stats && (d.bias || d.scale || d.relu)), "conv_gemm: stats are defined on raw accumulators only");
  int bn = d.block_n;
  if (bn == 0) bn = (d.N % 128 == 0) ? 128 : 64;
  VB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "conv_gemm: block_n must be 64/128/256");

  GemmKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = d.M;
  kp.N = d.N;
  kp.num_m_blocks = (d.M + BM - 1) / BM;
  kp.num_n_blocks = (d.N + bn - 1) / bn;
  kp.num_k_blocks = d.K / BK;
  kp.passes = d.passes;
  kp.scale = d.scale;
  kp.bias = d.bias;
  kp.relu = d.relu;
  kp.stats = d.stats;
  int rc;
  if (d.im2col) {
    VB_REQUIRE(d.Cin % BK == 0, "conv_gemm: Cin=%d must be a multiple of %d", d.Cin, BK);
    VB_REQUIRE(d.K == d.R * d.S * d.Cin, "conv_gemm: K != R*S*Cin");
    const int P = (d.H + d.pad_lo_h + d.pad_hi_h - d.R) / d.stride + 1;
    const int Q = (d.W + d.pad_lo_w + d.pad_hi_w - d.S) / d.stride + 1;
    VB_REQUIRE(d.M == d.batch * P * Q, "conv_gemm: M=%d != batch*P*Q=%d", d.M, d.batch * P * Q);
    kp.a_mode = 1;
    kp.PQ = P * Q;
    kp.Q = Q;
    kp.stride = d.stride;
    kp.pad_h = d.pad_lo_h;
    kp.pad_w = d.pad_lo_w;
    kp.S = d.S;
    kp.cin_blocks = d.Cin / BK;
    rc = encode_tma_im2col_nhwc(&kp.a_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.a_hi, d.batch, d.H, d.W, d.Cin, d.pad_lo_h,
                                d.pad_lo_w, d.pad_hi_h, d.pad_hi_w, d.R, d.S, d.stride, BK, BM,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (d.passes == 3) {
      rc = encode_tma_im2col_nhwc(&kp.a_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.a_lo, d.batch, d.H, d.W, d.Cin,
                                  d.pad_lo_h, d.pad_lo_w, d.pad_hi_h, d.pad_hi_w, d.R, d.S, d.stride, BK, BM,
                                  CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  } else {
    kp.a_mode = 0;
    rc = encode_tma_2d(&kp.a_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.a_hi, d.K, d.M, (uint64_t)d.K * 2, BK, BM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (d.passes == 3) {
      rc = encode_tma_2d(&kp.a_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.a_lo, d.K, d.M, (uint64_t)d.K * 2, BK, BM,
                         CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  rc = encode_tma_2d(&kp.b_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.b_hi, d.K, d.N, (uint64_t)d.K * 2, BK, bn,
                     CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (d.passes == 3) {
    rc = encode_tma_2d(&kp.b_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.b_lo, d.K, d.N, (uint64_t)d.K * 2, BK, bn,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  rc = encode_tma_2d(&kp.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.out, d.N, d.M, (uint64_t)d.N * 4, 32, BM,
                     CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;

  int stages = 0;
  const size_t smem = gemm_smem_bytes(bn, d.passes, &stages);
  VB_REQUIRE(stages >= 2, "conv_gemm: not enough shared memory for a 2-stage pipeline (bn=%d)", bn);
  kp.num_stages = stages;
  const int tiles = kp.num_m_blocks * kp.num_n_blocks;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (bn == 64) return launch_gemm<64>(kp, stages, smem, grid, stream);
  if (bn == 128) return launch_gemm<128>(kp, stages, smem, grid, stream);
  return launch_gemm<256>(kp, stages, smem, grid, stream);
}

}  // namespace vb
