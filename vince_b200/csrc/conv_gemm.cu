// Persistent warp-specialised implicit-GEMM convolution / linear kernel for sm_100a.
//
//   out[M, N] = epilogue( A[M, K] * W[N, K]^T )         fp32 accumulate in TMEM
//
// A-operand modes
//   0 TILED   A is a plain row-major [M, K] matrix (1x1 stride-1 convs, Linear layers): 2-D tiled TMA.
//   1 IM2COL  NHWC activation gathered on the fly by TMA *im2col* loads (strided / 7x7-stem / small-map convs):
//             k-block kb = filter tap (r, s) x 64-channel block.
//   2 HALO    3x3 stride-1 pad-1 convs on larger maps: a tile is TH full image rows; ONE tiled 4-D TMA load brings the
//             (TH+2) x (W+2) x 64ch input halo (zero-filled outside the image) into shared memory and all nine
//             filter taps read it through row-shifted UMMA descriptors (start address + (r*(W+2)+s)*128 B), cutting
//             the L2->SM activation traffic ~4.7x versus im2col.  Positions x >= W of each padded row are computed
//             and discarded.
// Operands are fp16.  To reproduce the reference's fp32 arithmetic (torch conv2d / mm, SURVEY.md 7 "hard parts") every
// fp32 value is carried as a fp16 (hi, lo) pair and each k-step issues three MMAs lo*hi + hi*lo + hi*hi into the same
// accumulator ("fp16x3", passes = 3).  passes = 1 is plain fp16.
//
// Warp roles: warps 0..3 = TMA producers (activation hi / lo planes, weight hi / lo planes), warp 4 = TMEM allocator + MMA
// issuer, then one or two SETS of four epilogue warps (one warp per TMEM lane quarter; TMEM -> registers -> swizzled smem
// -> TMA store; per-channel sum / sum-of-squares for train-mode BatchNorm accumulated in double-float / fp64; the LAST CTA
// to finish turns the sums into per-channel scale/shift and updates the running statistics, so no separate BN-finalize
// launch exists), and - apply epilogue with a residual on the small-K GEMMs only - two residual fetcher warps.
// 288 / 416 / 480 threads.  Accumulators are double-buffered in TMEM so the epilogue of tile t overlaps the main loop of
// tile t+1.  One CTA per SM; tiles are handed out round-robin with the M index fastest so that concurrently running CTAs
// share the weight tile through L2.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vb {

constexpr int BM = 128;          // rows (output positions) per tile == UMMA M == TMEM lanes
constexpr int BK = 64;           // fp16 elements per k-block        == one 128-byte swizzle row
constexpr int UMMA_K = 16;       // fp16
constexpr int GEMM_THREADS = 288;       // 4 TMA producer warps + 1 MMA warp + 4 epilogue warps
constexpr int EPI_WARP0 = 5;            // first epilogue warp
constexpr int EPI_TID0 = EPI_WARP0 * 32;
constexpr int MAX_RING = 8;
constexpr int STAGING_BYTES = BM * 128;   // 128 rows x 32 fp32
constexpr size_t SMEM_BUDGET = 226 * 1024;    // of 227 KB: leaves the 1 KB system reservation of one co-resident block

struct GemmKernelParams {
  CUtensorMap a_hi, a_lo, b_hi, b_lo, out;
  CUtensorMap out_hi, out_lo;    // EPI == 1: fp16 (hi, lo) output planes (box 32 channels x 128 rows, SWIZZLE_64B)
  int res_fetch;                 // EPI == 1: residual tiles are fetched by two extra warps (cp.async) into a ring of
  int res_slots;                 // res_slots 16 KB tiles behind the output staging tiles (0: row owners load registers)
  int staging_bufs;              // output staging tiles in shared memory (1, 2 or 4)
  uint32_t staging_total;        // bytes of the staging area: staging_bufs output tiles + res_slots residual tiles
  // EPI == 1 ("apply" epilogue): out = relu?( acc*alpha*coef[c] + coef[N+c] + residual ) split into fp16 planes
  const float* ep_coef;          // [2][N] per-channel (scale, shift): BatchNorm coefficients
  int res_kind;                  // 0 none, 1 fp16 planes [M,N], 2 bn(res_raw) with res_coef (downsample branch)
  const __half* res_hi;
  const __half* res_lo;
  const float* res_raw;
  const float* res_coef;
  int stats_only;                // EPI == 0: accumulate BatchNorm sums / finalize but store nothing
  int stat_channels;             // BatchNorm channels: N, or (EPI == 2, transposed statistics pass) the kernel's M extent
  // batched split-K GEMM (weight gradients; tiled A mode only): tile -> (batch, m_blk, n_blk), batch -> (tap, split);
  // A k-offset = split * kchunk; B k-offset = split * kchunk + (tap/3 - 1) * shift_w and B rows (tap%3) * N + n for
  // taps == 9 (B = three column-shifted copies stacked along the rows); output rows batch * out_batch_rows + m
  int num_tiles;
  int tiles_per_batch;           // num_m_blocks * num_n_blocks
  int splits;                    // k-splits per tap
  int kchunk;                    // elements of K per split (multiple of 64)
  int shift_w;                   // padded row pitch of the pixel axis (0: no taps)
  int out_batch_rows;
  const float* alpha_dev;        // optional device scalar multiplied into alpha (dynamic power-of-two gradient scale)
  int bn_save;                   // finalize also writes bn_coef[2N..3N) = mean, [3N..4N) = 1/sqrt(var+eps)
  int M, N;
  int num_m_blocks, num_n_blocks;
  int a_mode;                    // 0 tiled, 1 im2col, 2 halo
  int a_chunks;                  // A loads per tile (modes 0/1: k-blocks; halo: 64-channel blocks)
  int b_per_a;                   // B loads (k-blocks) per A load (modes 0/1: 1; halo: 9 taps)
  int PQ, Q, stride, pad_h, pad_w, S, cin_blocks;        // im2col geometry
  int Wp, TH, tiles_per_img, H, W;                        // halo geometry
  uint32_t a_plane_bytes;        // smem bytes reserved per A plane per stage (multiple of 1024)
  uint32_t a_tx_bytes;           // bytes one A-plane TMA load delivers
  int passes;
  int a_stages, b_stages;
  const float* scale;
  const float* bias;
  int relu;
  float alpha;                   // accumulators are multiplied by alpha first (exact power-of-two weight descale)
  double* stats;                 // optional [2][N]
  // train-mode BatchNorm finalize by the last CTA (all optional; need stats)
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_running_mean;
  float* bn_running_var;
  long long* bn_nbt;
  float* bn_coef;                // [2][N]: scale, shift
  unsigned int* bn_counter;
  float bn_momentum, bn_eps;
  double bn_count;
  int debug_skip_mma;            // debug only: issue no MMAs (pure TMA-pipeline timing), results are garbage
  unsigned long long* trace;     // debug timeline (VB_TRACE builds only): [4 tags][512] clock64 stamps of CTA 0
};

#ifdef VB_TRACE
#define VB_TRACE_EVENT(tag, idx)                                                              \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x < 2 && (idx) < 512) {                                \
      unsigned long long _t;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                  \
      p.trace[(blockIdx.x * 8 + (tag)) * 512 + (idx)] = _t;                                   \
    }                                                                                         \
  } while (0)
#else
#define VB_TRACE_EVENT(tag, idx) \
  do {                           \
  } while (0)
#endif

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// CG = 1: one CTA per 128-row tile.  CG = 2: a CTA pair (cluster of 2 on one TPC) computes a 256-row tile with
// tcgen05.mma.cta_group::2 - the leader's single issuing thread drives both SMs' tensor cores, each CTA loads its own
// 128 A rows and HALF of the weight tile, which halves both the MMA-dispatch load and the weight bytes per SM.
// HALO selects the A-operand mode 2 code paths at compile time (the MMA-issuing thread is the critical resource:
// its loop must carry no mode checks, divisions or 64-bit descriptor arithmetic).
// RES ("resident weights"): the layer has a single n-block and this CTA's share of the whole [N, K] weight matrix
// fits in shared memory next to the activation ring, so it is loaded ONCE per CTA; afterwards only activations
// stream and the issuing thread waits on / commits to one barrier per activation stage instead of one per k-block.
// EPI = 0: raw fp32 output (+ scale / bias / ReLU, + BatchNorm sums and finalize).  EPI = 1: "apply" epilogue - per-channel
// scale/shift (BatchNorm with known coefficients) + residual + ReLU, written as the fp16 (hi, lo) planes the next
// convolution consumes: eval-mode BatchNorm folded into the producing convolution (resnet.py:76-92,117-137), and the
// second pass of the train-mode "statistics pass + recompute pass" scheme for wide 1x1 convolutions.
// EPI = 2: TRANSPOSED statistics pass, the first pass of that scheme: the launcher swaps the operand roles (kernel rows =
// output channels from the weight matrix, kernel columns = pixels), so an accumulator ROW holds one channel over BN
// pixels and a thread's tcgen05.ld delivers 32 values of ITS channel - BatchNorm sum / sum of squares are then plain
// in-register adds (no staging tile, no barrier, no cross-thread reduction) and nothing is stored at all.
// EW = number of epilogue warp SETS (4 warps each, one per TMEM lane quarter).  The epilogue of the small-K layers is
// latency-bound with one warp per scheduler (issue slots ~25 % busy, ~1500-2900 cycles per 32-column chunk); with EW = 2
// the sets work on alternate chunks of a tile, each with its own staging tile, store leader and named barriers.
template <int BN, int CG, bool HALO, bool RES, int EPI, int EW>
// Register budget: 96 per thread unless the apply epilogue (which keeps a chunk of residual in registers) needs more -
// the small footprint leaves room for two or three streaming (BatchNorm-apply) blocks of the OTHER encoder's stream on
// the same SM, which is where the two-stream schedule gets its overlap.  (One set: 288 threads, two blocks' worth of
// registers; two sets: 416 threads launched, bounds declared for 640 so that ptxas caps at 65536 / 640 -> 96.)
__global__ void __launch_bounds__(EPI == 1 ? GEMM_THREADS + 128 * (EW - 1) + 64 : (EW == 1 ? GEMM_THREADS : 640), (EW == 1 && EPI != 1) ? 2 : 1)
conv_gemm_kernel(const __grid_constant__ GemmKernelParams p) {
  constexpr int ETHREADS = 128 * EW;               // epilogue threads
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int planes = (p.passes == 3) ? 2 : 1;
  constexpr uint32_t B_TILE = (BN / CG) * 128;     // weight rows held by THIS CTA
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int tile_start = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t a_stage_bytes = p.a_plane_bytes * planes;
  const uint32_t b_stage_bytes = B_TILE * planes;
  constexpr bool merged = !HALO && !RES;           // A and B rings advance in lockstep and share barriers

  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + (size_t)p.a_stages * a_stage_bytes;
  uint8_t* staging = b_ring + (size_t)p.b_stages * b_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(staging + p.staging_total);
  uint64_t* a_empty = a_full + MAX_RING;
  uint64_t* b_full = a_empty + MAX_RING;
  uint64_t* b_empty = b_full + MAX_RING;
  uint64_t* tmem_full = b_empty + MAX_RING;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;             // [4] residual ring (EPI == 1 with fetcher warps)
  uint64_t* res_empty = res_full + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_empty + 4);
  uint32_t* last_flag = tmem_ptr + 1;
  double* smem_stats = reinterpret_cast<double*>(tmem_ptr + 2);     // [4 epilogue warps][2][BN], warp-private

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_hi);
    tma_prefetch_desc(&p.b_hi);
    if (EPI == 1) {
      tma_prefetch_desc(&p.out_hi);
      tma_prefetch_desc(&p.out_lo);
    } else if (EPI == 0) {
      tma_prefetch_desc(&p.out);
    }
    if (planes == 2) {
      tma_prefetch_desc(&p.a_lo);
      tma_prefetch_desc(&p.b_lo);
    }
    // "full" barriers: every CTA's TMA loads signal its OWN barrier (local complete_tx); in a pair the leader's
    // barrier takes one more arrival, forwarded by the peer's otherwise idle warp 1 once the peer's data has landed
    // one arrival (with its expect_tx) per producer warp = per fp16 plane.  When every A load pairs with exactly
    // one B load (tiled / im2col modes) both operands share the B ring's barriers: one wait per k-block.
    const uint32_t fwd = (CG == 2 && cta_rank == 0) ? 1u : 0u;
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], (uint32_t)planes + fwd);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < (RES ? 1 : p.b_stages); ++s) {       // RES: one barrier for the one-time weight load
      mbar_init(&b_full[s], (uint32_t)planes * (merged ? 2u : 1u) + fwd);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&res_full[s], 64);                 // every fetcher thread arrives once its copies have landed
      mbar_init(&res_empty[s], 128);               // every thread of the consuming epilogue set
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CG == 2 ? 8 * EW : ETHREADS);    // pair: one arrive per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 4) {
    if (CG == 2) {
      tmem_alloc2(tmem_ptr, 2 * BN);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_ptr, 2 * BN);
      tmem_relinquish();
    }
  }
  if (EPI == 0 && threadIdx.x >= EPI_TID0) {       // (EPI 1 keeps 4 x BN coefficients there, EPI 2 nothing)
    for (int i = threadIdx.x - EPI_TID0; i < 8 * BN + 2; i += ETHREADS) smem_stats[i] = 0.0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                 // peer barriers initialised before any remote arrive / multicast
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.num_tiles;               // tiles_per_batch (num_m_blocks counts 128*CG-row tiles) x batches

  // The producer and MMA warps keep their control flow WARP-UNIFORM (all 32 lanes walk the loops and wait on the
  // barriers) and elect one lane only around the TMA / tcgen05 instructions themselves: tile indices, stage
  // counters, smem addresses and descriptors then live in uniform registers, which is what UTMALDG / UTCHMMA
  // consume.  Running the whole loop under `if (lane == 0)` makes ptxas treat every operand as divergent and
  // emit a waterfall of R2UR moves per instruction (~90 cycles per MMA issue, measured).
  if (warp < 2) {
    // ======================= TMA producers, activations: warp 0 = hi plane, warp 1 = lo plane =======================
    // (a single thread can only issue a TMA load every ~250 cycles, measured; one warp per operand plane keeps the
    //  four loads of a k-block in flight concurrently)
    const int plane = warp;
    if (plane < planes) {
      const CUtensorMap* amap = plane == 0 ? &p.a_hi : &p.a_lo;
      int as = 0;
      uint32_t aphase = 0;
      int tr_a = 0;
      for (int tile = tile_start; tile < num_tiles; tile += tile_step) {
        const int tb = tile % p.tiles_per_batch;
        const int m_blk = (tb % p.num_m_blocks) * CG + (int)cta_rank;
        const int m0 = m_blk * BM;
        const int a_koff = ((tile / p.tiles_per_batch) % p.splits) * p.kchunk;      // 0 unless batched split-K
        int img = 0, ph = 0, qw = 0;
        if (p.a_mode == 1) {
          img = m0 / p.PQ;
          const int rem = m0 - img * p.PQ;
          const int p0 = rem / p.Q;
          const int q0 = rem - p0 * p.Q;
          ph = p0 * p.stride - p.pad_h;
          qw = q0 * p.stride - p.pad_w;
        } else if (HALO) {
          img = m_blk / p.tiles_per_img;
          ph = (m_blk - img * p.tiles_per_img) * p.TH - 1;        // first halo row
        }
        for (int ac = 0; ac < p.a_chunks; ++ac) {
          uint64_t* a_full_bar = merged ? &b_full[as] : &a_full[as];
          mbar_wait(merged ? &b_empty[as] : &a_empty[as], aphase ^ 1);
          if (lane == 0 && plane == 0) VB_TRACE_EVENT(4, tr_a);
          ++tr_a;
          uint8_t* dst = a_ring + (size_t)as * a_stage_bytes + (size_t)plane * p.a_plane_bytes;
          int tap = 0, cb = 0, r = 0, sx = 0;
          if (p.a_mode == 1) {
            tap = ac / p.cin_blocks;
            cb = ac - tap * p.cin_blocks;
            r = tap / p.S;
            sx = tap - r * p.S;
          }
          if (elect_one()) {
            if (p.debug_skip_mma & 2) mbar_arrive(a_full_bar);
            else {
            mbar_expect_tx(a_full_bar, p.a_tx_bytes);
            if (p.a_mode == 0) tma_load_2d(dst, amap, a_full_bar, ac * BK + a_koff, m0);
            else if (p.a_mode == 1) tma_load_im2col_4d(dst, amap, a_full_bar, cb * BK, qw, ph, img, (uint16_t)sx, (uint16_t)r);
            else tma_load_4d(dst, amap, a_full_bar, ac * BK, -1, ph, img);
            }
          }
          __syncwarp();
          if (++as == p.a_stages) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
    }
  } else if (warp < 4) {
    // ======================= TMA producers, weights: warp 2 = hi plane, warp 3 = lo plane =======================
    const int plane = warp - 2;
    if (plane < planes) {
      const CUtensorMap* bmap = plane == 0 ? &p.b_hi : &p.b_lo;
      int bs = 0;
      uint32_t bphase = 0;
      int tr_p = 0;
      if (RES) {
        // whole weight share of this CTA, once: slot (ac * taps + bi) holds the k-block the MMA loop consumes then
        if (tile_start < num_tiles) {
          const int nrow = (int)cta_rank * (BN / CG);
          const int total_kb = p.a_chunks * p.b_per_a;
          if (elect_one()) {
            mbar_expect_tx(&b_full[0], (uint32_t)total_kb * B_TILE);
            for (int ac = 0; ac < p.a_chunks; ++ac)
              for (int bi = 0; bi < p.b_per_a; ++bi) {
                const int kb = HALO ? (bi * p.a_chunks + ac) : ac;
                uint8_t* dst = b_ring + (size_t)(ac * p.b_per_a + bi) * b_stage_bytes + (size_t)plane * B_TILE;
                tma_load_2d(dst, bmap, &b_full[0], kb * BK, nrow);
              }
          }
          __syncwarp();
        }
      } else {
      for (int tile = tile_start; tile < num_tiles; tile += tile_step) {
        const int n_blk = (tile % p.tiles_per_batch) / p.num_m_blocks;
        int b_koff = 0, b_row0 = 0;
        if (p.splits > 1 || p.shift_w != 0) {
          const int batch = tile / p.tiles_per_batch;
          const int tap = batch / p.splits;
          b_koff = (batch - tap * p.splits) * p.kchunk;
          // TMA needs 16-byte aligned global addresses, so only the ROW part of a tap's shift (shift_w, a multiple
          // of 8 elements) is applied as a coordinate; the column part selects one of three pre-shifted copies of the
          // operand stacked along its row axis (vince_transpose_pad, copies = 3)
          if (p.shift_w != 0) b_koff += (tap / 3 - 1) * p.shift_w, b_row0 = (tap % 3) * p.N;
        }
        // this CTA's share of the weight tile (all of it, or its half when paired)
        const int nrow = b_row0 + n_blk * BN + (int)cta_rank * (BN / CG);
        for (int ac = 0; ac < p.a_chunks; ++ac) {
          for (int bi = 0; bi < p.b_per_a; ++bi) {
            mbar_wait(&b_empty[bs], bphase ^ 1);
            if (lane == 0 && plane == 0) VB_TRACE_EVENT(0, tr_p);
            ++tr_p;
            uint8_t* dst = b_ring + (size_t)bs * b_stage_bytes + (size_t)plane * B_TILE;
            // weight k-block: modes 0/1 -> ac; halo -> tap bi, channel block ac
            const int kb = HALO ? (bi * p.a_chunks + ac) : ac;
            if (elect_one()) {
              if (p.debug_skip_mma & 4) mbar_arrive(&b_full[bs]);
              else {
                mbar_expect_tx(&b_full[bs], B_TILE);
                tma_load_2d(dst, bmap, &b_full[bs], kb * BK + b_koff, nrow);
              }
            }
            __syncwarp();
            if (++bs == p.b_stages) {
              bs = 0;
              bphase ^= 1;
            }
          }
        }
      }
      }
    }
  } else if (warp == 4) {
    // ======================= MMA issuer (leader CTA only when CG == 2) =======================
    constexpr uint32_t idesc = make_idesc(UMMA_FMT_F16, BM * CG, BN);
    int as = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    int local_t = 0;
    int tr_m = 0;
    if (CG == 2 && cta_rank == 1) {
      // peer CTA: forward "my operands have landed" to the leader's barriers, one remote arrive per stage use
      if (RES && tile_start < num_tiles) {
        mbar_wait(&b_full[0], 0);
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&b_full[0]), 0));
      }
      for (int tile = tile_start; tile < num_tiles; tile += tile_step) {
        for (int ac = 0; ac < p.a_chunks; ++ac) {
          if (!merged) {
            mbar_wait(&a_full[as], aphase);
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&a_full[as]), 0));
          }
          for (int bi = 0; bi < (RES ? 0 : p.b_per_a); ++bi) {
            mbar_wait(&b_full[bs], bphase);
            if (lane == 0) VB_TRACE_EVENT(5, tr_m);
            ++tr_m;
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&b_full[bs]), 0));
            if (++bs == p.b_stages) {
              bs = 0;
              bphase ^= 1;
            }
          }
          if (++as == p.a_stages) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
    }
    // ---- leader: the one thread per CTA (pair) that feeds the tensor cores ----
    // Shared-memory matrix descriptors are handled as two 32-bit words: the high word (SBO = 1024 B, version 1,
    // SWIZZLE_128B) is a constant, the low word is (address >> 4) | LBO(16 B) << 16; stepping to the next 16-element
    // k-slice adds 2 to the low word.  Everything below is warp-uniform so it lives in uniform registers.
    constexpr uint32_t DESC_HI = 0x40004040u;
    const uint32_t a_ring_w = (smem_u32(a_ring) >> 4) | 0x10000u;
    const uint32_t b_ring_w = (smem_u32(b_ring) >> 4) | 0x10000u;
    const uint32_t a_stage_w = a_stage_bytes >> 4, a_plane_w = p.a_plane_bytes >> 4;
    const uint32_t b_stage_w = b_stage_bytes >> 4;
    constexpr uint32_t b_plane_w = B_TILE >> 4;
    const uint32_t b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
    const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);
    const bool skip_mma = (p.debug_skip_mma & 1) != 0;
    const int a_chunks = p.a_chunks, a_stages = p.a_stages, b_stages = p.b_stages;
    const uint32_t wp8 = (uint32_t)p.Wp * 8u;            // halo: one padded image row in descriptor units
    if (RES && cta_rank == 0 && tile_start < num_tiles) {
      mbar_wait(&b_full[0], 0);                          // resident weights (both CTAs' shares) have landed
      tc_fence_after_sync();
    }
    for (int tile = (cta_rank == 0 ? tile_start : num_tiles); tile < num_tiles; tile += tile_step, ++local_t) {
      const int acc = local_t & 1;
      const uint32_t acc_phase = (local_t >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      uint32_t accumulate = 0;
      for (int ac = 0; ac < a_chunks; ++ac) {
        uint32_t a_w = 0;
        if (!merged) {
          mbar_wait_addr(a_full0 + 8u * as, aphase);
          if (RES) tc_fence_after_sync();
          a_w = a_ring_w + (uint32_t)as * a_stage_w;
        }
        const bool last_chunk = (ac == a_chunks - 1);
        constexpr int TAPS = HALO ? 9 : 1;
#pragma unroll
        for (int bi = 0; bi < TAPS; ++bi) {
          if (!RES) {
            mbar_wait_addr(b_full0 + 8u * bs, bphase);
            tc_fence_after_sync();
          }
          if (lane == 0) VB_TRACE_EVENT(1, tr_m);
          // halo mode: filter tap (r, s) = rows shifted by r*(W+2)+s inside the halo tile
          const uint32_t aw = HALO ? a_w + (uint32_t)(bi / 3) * wp8 + (uint32_t)(bi % 3) * 8u
                                   : (merged ? a_ring_w + (uint32_t)bs * a_stage_w : a_w);
          const uint32_t bw = b_ring_w + (uint32_t)(RES ? ac * TAPS + bi : bs) * b_stage_w;
          if (elect_one()) {
            if (!skip_mma) {
              if (planes == 2) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  // small cross terms first, dominant term last
                  umma_f16_w<CG>(d_tmem, aw + a_plane_w + 2 * k, bw + 2 * k, DESC_HI, idesc, accumulate);
                  umma_f16_w<CG>(d_tmem, aw + 2 * k, bw + b_plane_w + 2 * k, DESC_HI, idesc, 1);
                  umma_f16_w<CG>(d_tmem, aw + 2 * k, bw + 2 * k, DESC_HI, idesc, 1);
                  accumulate = 1;
                }
              } else {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  umma_f16_w<CG>(d_tmem, aw + 2 * k, bw + 2 * k, DESC_HI, idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            const bool last_b = (bi == TAPS - 1);
            if (!RES) umma_commit_addr<CG>(b_empty0 + 8u * bs);      // frees the weight slot (in BOTH CTAs of a pair)
            if (!merged && last_b) umma_commit_addr<CG>(a_empty0 + 8u * as);   // ... the activation slot after its last tap
            if (last_b && last_chunk) umma_commit_addr<CG>(smem_u32(&tmem_full[acc]));   // accumulators ready
          }
          accumulate = 1;
          if (lane == 0) VB_TRACE_EVENT(2, tr_m);
          ++tr_m;
          if (!RES) {
            if (++bs == b_stages) {
              bs = 0;
              bphase ^= 1;
            }
          }
        }
        if (!merged) {
          if (++as == a_stages) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
    }
  } else if (EPI == 1 && warp >= EPI_WARP0 + 4 * EW) {
    // ======================= residual fetchers (two warps; launched only when p.res_fetch) =======================
    // The residual tile of chunk k of this CTA (32 channels x 128 rows: 8 KB hi + 8 KB lo planes, or 16 KB of the raw fp32
    // tensor) is copied with cp.async (LSU path: it does not queue behind the TMA output stores) into slot k % res_slots of
    // a ring behind the staging tiles, in the swizzled layout the row owners read conflict-free; the copies of a slot are
    // tracked by its mbarrier (cp.async.mbarrier.arrive), so the fetchers run res_slots chunks ahead of the epilogue and
    // neither the row owners nor their proxy fence ever wait on a global load.
    const int f = threadIdx.x - (EPI_WARP0 + 4 * EW) * 32;          // 0..63
    const int R = p.res_slots;
    const uint32_t ring_s = smem_u32(staging) + (uint32_t)p.staging_bufs * STAGING_BYTES;
    const int res_kind = p.res_kind;
    const long long Mrows = p.M;
    const int N = p.N;
    uint32_t k = 0;
    for (int tile = tile_start; tile < num_tiles; tile += tile_step) {
      const int tb = tile % p.tiles_per_batch;
      const int m_blk = (tb % p.num_m_blocks) * CG + (int)cta_rank;
      const int n_blk = tb / p.num_m_blocks;
      const long long m0 = (long long)m_blk * BM;
      int nvalid = BM;
      if (m0 + nvalid > Mrows) nvalid = (int)(Mrows - m0 > 0 ? Mrows - m0 : 0);
      for (int chunk = 0; chunk < BN / 32; ++chunk, ++k) {
        const uint32_t slot = k % (uint32_t)R;
        mbar_wait(&res_empty[slot], ((k / (uint32_t)R) & 1u) ^ 1u);
        const uint32_t dst = ring_s + slot * STAGING_BYTES;
        const int c0 = n_blk * BN + chunk * 32;
        if (res_kind == 1) {
#pragma unroll 4
          for (int it = 0; it < 16; ++it) {
            const int idx = it * 64 + f;                 // 512 hi pieces then 512 lo pieces of 16 bytes
            const int plane = idx >> 9, rem = idx & 511;
            const int r = rem >> 2, j = rem & 3;
            const __half* src = plane == 0 ? p.res_hi : p.res_lo;
            if (r < nvalid && src != nullptr)
              cp_async_16_l2_256(dst + (uint32_t)plane * 8192u + (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)),
                                 src + (m0 + r) * N + c0 + j * 8);
          }
        } else {
#pragma unroll 4
          for (int it = 0; it < 16; ++it) {
            const int idx = it * 64 + f;                 // 1024 pieces of 16 bytes: 8 per 128-byte row
            const int r = idx >> 3, j = idx & 7;
            if (r < nvalid)
              cp_async_16_l2_256(dst + (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)), p.res_raw + (m0 + r) * N + c0 + j * 4);
          }
        }
        cp_async_mbar_arrive_noinc(&res_full[slot]);
      }
    }
    cp_async_wait_all();
  } else {
    // ======================= epilogue (warps 5..8) =======================
    // Per 32-column chunk: TMEM -> registers (one accumulator row per thread) -> epilogue math -> swizzled staging
    // tile in shared memory -> one named barrier -> TMA store (asynchronous: the copy-out costs the SM's LSU nothing;
    // coalesced st.global from the staged tile was measured 1.6x SLOWER on the wide layers) plus, for train-mode
    // BatchNorm, column sums read back from the staged tile.  With two staging tiles (small-K, epilogue-bound layers)
    // the store of chunk j drains while chunk j+1 is produced and a chunk costs ONE barrier; with one tile the
    // round-1 scheme (wait for the store, barrier, write, barrier) is kept.
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - EPI_TID0;       // 0..ETHREADS-1
    const int eset = etid >> 7;                    // epilogue warp set of this thread
    const int ltid = etid & 127;                   // thread index inside the set
    const uint32_t bar_a = 2u + 2u * (uint32_t)eset, bar_b = 3u + 2u * (uint32_t)eset;     // the set's named barriers
    const uint32_t staging_s = smem_u32(staging);
    const int nbuf = p.staging_bufs;               // output staging tiles in total, nbuf / EW per set
    const bool one_tile = nbuf == EW;              // this set cycles through a single staging tile
    // kernel parameters used per chunk, read once (the asm barriers would otherwise force constant-bank re-reads)
    const bool has_stats = (EPI != 1) && p.stats != nullptr;
    const bool do_store = p.stats_only == 0;
    const float alpha = p.alpha * (p.alpha_dev != nullptr ? __ldg(p.alpha_dev) : 1.f);
    const int N = p.N;
    const long long M = p.M;
    const int relu = p.relu;
    const int res_kind = (EPI == 1) ? p.res_kind : 0;
    const bool store_leader = (ltid == 0);
    bool store_pending = false;
    const float* const ep_scale = p.scale;
    const float* const ep_bias = p.bias;
    const uint32_t stats_s = smem_u32(smem_stats + 1);        // 16-byte aligned: [4 warps][BN] x {sum, err, sq, err}
    // halo mode: which output pixel (if any) this accumulator row is
    int hy = 0, hx = 0;
    if (HALO) {
      hy = row / p.Wp;
      hx = row - hy * p.Wp;
    }
    int local_t = 0;
    int cur_n_blk = -1;
    uint32_t tmem_empty_leader[2] = {0u, 0u};
    if (CG == 2) {
      tmem_empty_leader[0] = mapa_shared(smem_u32(&tmem_empty[0]), 0);
      tmem_empty_leader[1] = mapa_shared(smem_u32(&tmem_empty[1]), 0);
    }
    // first global row and number of valid staging rows of a tile
    auto tile_rows = [&](int tile, long long& m0, int& nvalid, int& n_blk) {
      const int tb = tile % p.tiles_per_batch;
      const int m_blk = (tb % p.num_m_blocks) * CG + (int)cta_rank;
      n_blk = tb / p.num_m_blocks;
      if (HALO) {
        const int img = m_blk / p.tiles_per_img;
        const int y0 = (m_blk - img * p.tiles_per_img) * p.TH;
        m0 = ((long long)img * p.H + y0) * p.W;             // TH full image rows are contiguous in NHWC
        nvalid = min(p.TH, p.H - y0) * p.W;
      } else {
        m0 = (long long)m_blk * BM;
        nvalid = BM;
      }
      if (m0 + nvalid > M) nvalid = (int)(M - m0 > 0 ? M - m0 : 0);
    };
    // EPI == 1: the residual of a chunk goes straight from global memory into the row owner's registers - its 32 channels
    // are 64 contiguous bytes per fp16 plane (128 bytes of the raw fp32 tensor), fetched with 256-bit loads that bypass L1
    // and ask L2 for the whole 256-byte neighbourhood the following chunks of the tile will want.  The loads for the
    // set's NEXT chunk are issued as soon as the current ones are consumed.  (Two earlier schemes were slower: cp.async
    // tiles one chunk ahead; and TMA tiles through an mbarrier ring, which queue in the SM's TMA unit behind the output
    // stores - in a write-bound kernel every residual load then waited for the store ahead of it, ncu: half of the
    // epilogue's samples on that barrier, reads and writes serialised at ~480 us for the 56x56 64->256 layer.)
    constexpr int CHUNKS = BN / 32;
    uint32_t rres[32];
    const bool row_ok0 = HALO ? ((hy < p.TH) && (hx < p.W)) : true;
    const int srow0 = HALO ? hy * p.W + hx : row;
    auto load_residual = [&](int tile, int chunk) {
#pragma unroll
      for (int i = 0; i < 32; ++i) rres[i] = 0u;
      if (res_kind == 0 || tile >= num_tiles) return;
      long long m0;
      int nvalid, n_blk;
      tile_rows(tile, m0, nvalid, n_blk);
      if (!(row_ok0 && srow0 < nvalid)) return;
      const long long e = (m0 + srow0) * N + n_blk * BN + chunk * 32;
      if (res_kind == 1) {
        ldg256_stream(p.res_hi + e, &rres[0]);
        ldg256_stream(p.res_hi + e + 16, &rres[8]);
        if (p.res_lo != nullptr) {
          ldg256_stream(p.res_lo + e, &rres[16]);
          ldg256_stream(p.res_lo + e + 16, &rres[24]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) ldg256_stream(p.res_raw + e + 8 * j, &rres[8 * j]);
      }
    };
    const bool res_fetch = (EPI == 1) && p.res_fetch != 0 && res_kind != 0;
    const uint32_t ring_s = staging_s + (uint32_t)nbuf * STAGING_BYTES;
    if (EPI == 1 && !res_fetch) load_residual(tile_start, eset);
    if (EPI == 2) {
      // ---- transposed statistics pass: this thread owns channel (m_blk * 128 + row) of every tile it sees ----
      const int NC = p.stat_channels;
      float s_hi = 0.f, s_er = 0.f, q_hi = 0.f, q_er = 0.f;      // double-float (value, error) running sums
      int cur_ch = -1;
      auto flush = [&]() {
        if (cur_ch >= 0 && cur_ch < NC && p.stats != nullptr) {
          const double a = (double)alpha;
          atomicAdd(&p.stats[cur_ch], ((double)s_hi + (double)s_er) * a);
          atomicAdd(&p.stats[NC + cur_ch], ((double)q_hi + (double)q_er) * a * a);
        }
        s_hi = s_er = q_hi = q_er = 0.f;
      };
      for (int tile = tile_start; tile < num_tiles; tile += tile_step, ++local_t) {
        const int acc = local_t & 1;
        const uint32_t acc_phase = (local_t >> 1) & 1;
        const int m_blk = ((tile % p.tiles_per_batch) % p.num_m_blocks) * CG + (int)cta_rank;
        const int ch = m_blk * BM + row;
        if (ch != cur_ch) {
          flush();
          cur_ch = ch;
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after_sync();
#pragma unroll 1
        for (int chunk = eset; chunk < BN / 32; chunk += EW) {
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + chunk * 32, raw);
          tmem_ld_wait();
          if (chunk + EW >= BN / 32) {
            tc_fence_before_sync();
            if (CG == 2) {
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(tmem_empty_leader[acc]);
            } else {
              mbar_arrive(&tmem_empty[acc]);
            }
          }
          // pixels past the end of the activation matrix were zero-filled by TMA: they add nothing
          float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f, qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float x0 = __uint_as_float(raw[i]), x1 = __uint_as_float(raw[i + 1]);
            const float x2 = __uint_as_float(raw[i + 2]), x3 = __uint_as_float(raw[i + 3]);
            sa += x0, sb += x1, sc += x2, sd += x3;
            qa = fmaf(x0, x0, qa), qb = fmaf(x1, x1, qb), qc = fmaf(x2, x2, qc), qd = fmaf(x3, x3, qd);
          }
          two_sum_acc(s_hi, s_er, (sa + sb) + (sc + sd));
          two_sum_acc(q_hi, q_er, (qa + qb) + (qc + qd));
        }
      }
      flush();
    }
    for (int tile = (EPI == 2 ? num_tiles : tile_start); tile < num_tiles; tile += tile_step, ++local_t) {
      const int acc = local_t & 1;
      const uint32_t acc_phase = (local_t >> 1) & 1;
      long long m0;
      int nvalid, n_blk;
      tile_rows(tile, m0, nvalid, n_blk);
      const int m_blk = ((tile % p.tiles_per_batch) % p.num_m_blocks) * CG + (int)cta_rank;
      const int out_row0 = (tile / p.tiles_per_batch) * p.out_batch_rows + m_blk * BM;
      int img = 0, y0 = 0;                         // halo mode: TMA store coordinates of the tile
      if (HALO) {
        img = m_blk / p.tiles_per_img;
        y0 = (m_blk - img * p.tiles_per_img) * p.TH;
      }
      bool valid = true;
      int srow = row;                              // row inside the staging tile
      if (HALO) {
        valid = (hy < p.TH) && (hx < p.W);
        srow = hy * p.W + hx;
      }
      valid = valid && srow < nvalid;
      if (EPI == 1 && res_kind != 0 && !res_fetch && tile + tile_step < num_tiles) {
        // the residual rows of this CTA's NEXT tile: ask L2 for them now, a whole tile time before the 256-bit loads
        // above want them (those are issued only one chunk ahead, less than a DRAM round trip under load)
        long long m0n;
        int nvn, nbn;
        tile_rows(tile + tile_step, m0n, nvn, nbn);
        if (row_ok0 && srow0 < nvn) {
          const long long e = (m0n + srow0) * N + nbn * BN;
          if (res_kind == 1) {
            for (int l = eset; l < BN / 64; l += EW) {
              prefetch_l2(p.res_hi + e + 64 * l);
              if (p.res_lo != nullptr) prefetch_l2(p.res_lo + e + 64 * l);
            }
          } else {
            for (int l = eset; l < BN / 32; l += EW) prefetch_l2(p.res_raw + e + 32 * l);
          }
        }
      }
      if (EPI == 1 && n_blk != cur_n_blk) {
        // per-channel coefficients of this n-block, cached in the (otherwise unused) statistics area:
        // [0,BN) scale  [BN,2BN) shift  [2BN,3BN) residual scale  [3BN,4BN) residual shift
        float* cf = reinterpret_cast<float*>(smem_stats + 1);      // +8 bytes: 16-byte aligned for the vector reads
        named_bar_sync(1, ETHREADS);                    // readers of the previous n-block's coefficients are done
        for (int i = etid; i < BN; i += ETHREADS) {
          const int c = n_blk * BN + i;
          const bool ok = c < N;
          cf[i] = ok ? __ldg(p.ep_coef + c) : 0.f;
          cf[BN + i] = ok ? __ldg(p.ep_coef + N + c) : 0.f;
          if (res_kind == 2) {
            cf[2 * BN + i] = ok ? __ldg(p.res_coef + c) : 0.f;
            cf[3 * BN + i] = ok ? __ldg(p.res_coef + N + c) : 0.f;
          }
        }
        named_bar_sync(1, ETHREADS);
        cur_n_blk = n_blk;
      }
      if (has_stats && n_blk != cur_n_blk) {
        if (cur_n_blk >= 0) {
          named_bar_sync(1, ETHREADS);
          for (int i = etid; i < BN; i += ETHREADS) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint32_t slot = stats_s + (uint32_t)(w * BN + i) * 16u;
              const float4 t = lds_v4(slot);
              a += (double)t.x + (double)t.y;
              b += (double)t.z + (double)t.w;
              sts_v4(slot, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            if (cur_n_blk * BN + i < N) {
              atomicAdd(&p.stats[cur_n_blk * BN + i], a);
              atomicAdd(&p.stats[N + cur_n_blk * BN + i], b);
            }
          }
          named_bar_sync(1, ETHREADS);
        }
        cur_n_blk = n_blk;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      if (etid == 0) VB_TRACE_EVENT(3, local_t);
#pragma unroll 1
      for (int chunk = eset; chunk < BN / 32; chunk += EW) {
        const uint32_t chunk_ctr = (uint32_t)local_t * (uint32_t)(BN / 32) + (uint32_t)chunk;   // chunk index in this CTA
        const int c0 = n_blk * BN + chunk * 32;
        // the set's tile(s): tiles [eset * nbuf / EW, ...), alternating with the set's own chunk count when it has two
        const uint32_t buf = staging_s + ((uint32_t)eset * (uint32_t)(nbuf / EW) +
                                          (one_tile ? 0u : ((chunk_ctr / (uint32_t)EW) & 1u))) * STAGING_BYTES;
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + chunk * 32, raw);
        tmem_ld_wait();
        if (chunk + EW >= BN / 32) {
          // all TMEM reads of this accumulator (by this set) are done: hand it back to the MMA warp
          tc_fence_before_sync();
          if (CG == 2) {
            // the leader's MMA warp owns both accumulators: one remote arrive per epilogue warp
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader[acc]);
          } else {
            mbar_arrive(&tmem_empty[acc]);
          }
        }
        if (EPI == 1) {
          // fetcher mode: the residual tile of this chunk sits in ring slot chunk_ctr % res_slots once its barrier flips;
          // this row's 32 channels are copied to registers (same conflict-free swizzle as the staging tiles)
          if (res_fetch) {
            const uint32_t slot = chunk_ctr % (uint32_t)p.res_slots;
            mbar_wait(&res_full[slot], (chunk_ctr / (uint32_t)p.res_slots) & 1u);
            const uint32_t rt = ring_s + slot * STAGING_BYTES;
            if (res_kind == 1) {
              const uint32_t rsw = (uint32_t)((srow >> 1) & 3);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 fh = lds_v4(rt + (uint32_t)(srow * 64 + ((j ^ rsw) << 4)));
                const float4 fl = lds_v4(rt + 8192u + (uint32_t)(srow * 64 + ((j ^ rsw) << 4)));
                rres[4 * j + 0] = __float_as_uint(fh.x), rres[4 * j + 1] = __float_as_uint(fh.y);
                rres[4 * j + 2] = __float_as_uint(fh.z), rres[4 * j + 3] = __float_as_uint(fh.w);
                rres[16 + 4 * j + 0] = __float_as_uint(fl.x), rres[16 + 4 * j + 1] = __float_as_uint(fl.y);
                rres[16 + 4 * j + 2] = __float_as_uint(fl.z), rres[16 + 4 * j + 3] = __float_as_uint(fl.w);
              }
              if (p.res_lo == nullptr) {
#pragma unroll
                for (int j = 16; j < 32; ++j) rres[j] = 0u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 fr = lds_v4(rt + (uint32_t)(srow * 128 + ((j ^ (srow & 7)) << 4)));
                rres[4 * j + 0] = __float_as_uint(fr.x), rres[4 * j + 1] = __float_as_uint(fr.y);
                rres[4 * j + 2] = __float_as_uint(fr.z), rres[4 * j + 3] = __float_as_uint(fr.w);
              }
            }
            if (!valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) rres[j] = 0u;
            }
            mbar_arrive(&res_empty[slot]);              // the slot may be refilled (this thread holds its row in registers)
          }
          // with one output staging tile everybody must be done with the previous chunk's tile before it is overwritten
          if (one_tile) {
            if (store_leader && store_pending) tma_store_wait_read0();
            named_bar_sync(bar_a, 128);
          }
          const uint32_t cfs = stats_s + (uint32_t)(chunk * 32) * 4u;
          // two 8 KB tiles (hi | lo) of 64-byte rows; 16-byte piece j of row `srow` stored at j ^ ((srow >> 1) & 3),
          // conflict-free.  One 8-channel group at a time keeps the live registers at raw[32] + a handful.
          const uint32_t sw = (uint32_t)((srow >> 1) & 3);
          const uint32_t rowp = buf + (uint32_t)srow * 64u;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float4 sc = lds_v4(cfs + (2 * g8 + q) * 16), sh = lds_v4(cfs + BN * 4 + (2 * g8 + q) * 16);
              v[4 * q + 0] = fmaf(__uint_as_float(raw[8 * g8 + 4 * q + 0]) * alpha, sc.x, sh.x);
              v[4 * q + 1] = fmaf(__uint_as_float(raw[8 * g8 + 4 * q + 1]) * alpha, sc.y, sh.y);
              v[4 * q + 2] = fmaf(__uint_as_float(raw[8 * g8 + 4 * q + 2]) * alpha, sc.z, sh.z);
              v[4 * q + 3] = fmaf(__uint_as_float(raw[8 * g8 + 4 * q + 3]) * alpha, sc.w, sh.w);
            }
            if (res_kind == 1) {
              const uint32_t hw[4] = {rres[4 * g8], rres[4 * g8 + 1], rres[4 * g8 + 2], rres[4 * g8 + 3]};
              const uint32_t lw[4] = {rres[16 + 4 * g8], rres[17 + 4 * g8], rres[18 + 4 * g8], rres[19 + 4 * g8]};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __half2 ha = *reinterpret_cast<const __half2*>(&hw[i]), hb = *reinterpret_cast<const __half2*>(&lw[i]);
                v[2 * i] += __low2float(ha) + __low2float(hb);
                v[2 * i + 1] += __high2float(ha) + __high2float(hb);
              }
            } else if (res_kind == 2) {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const float4 rsc = lds_v4(cfs + 2 * BN * 4 + (2 * g8 + q) * 16);
                const float4 rsh = lds_v4(cfs + 3 * BN * 4 + (2 * g8 + q) * 16);
                // (rows that are not output pixels hold zeros and are never stored)
                v[4 * q + 0] += fmaf(__uint_as_float(rres[8 * g8 + 4 * q + 0]), rsc.x, rsh.x);
                v[4 * q + 1] += fmaf(__uint_as_float(rres[8 * g8 + 4 * q + 1]), rsc.y, rsh.y);
                v[4 * q + 2] += fmaf(__uint_as_float(rres[8 * g8 + 4 * q + 2]), rsc.z, rsh.z);
                v[4 * q + 3] += fmaf(__uint_as_float(rres[8 * g8 + 4 * q + 3]), rsc.w, rsh.w);
              }
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            uint32_t hq[4], lq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __half h0, l0, h1, l1;
              split_f16(v[2 * i], h0, l0);
              split_f16(v[2 * i + 1], h1, l1);
              const __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
              hq[i] = *reinterpret_cast<const uint32_t*>(&hh);
              lq[i] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            if (valid) {
              const uint32_t off = (uint32_t)((g8 ^ sw) << 4);
              sts_v4(rowp + off, make_float4(__uint_as_float(hq[0]), __uint_as_float(hq[1]), __uint_as_float(hq[2]),
                                             __uint_as_float(hq[3])));
              sts_v4(rowp + 8192u + off, make_float4(__uint_as_float(lq[0]), __uint_as_float(lq[1]),
                                                     __uint_as_float(lq[2]), __uint_as_float(lq[3])));
            }
          }
          // this chunk's residual is consumed: fetch the one of this set's next chunk (same tile, or the next tile's first).
          // (Issued here, before the proxy fence, the fence absorbs part of the load latency - ncu: 25 % of the epilogue's
          //  samples on fence + barrier - but issuing after the barrier was measured slower still: 497 vs 434 us.)
          if (res_kind != 0 && !res_fetch) {
            if (chunk + EW < CHUNKS) load_residual(tile, chunk + EW);
            else load_residual(tile + tile_step, eset);
          }
          fence_proxy_async_smem();
          // two tiles: the PREVIOUS chunk's store (other tile) must have drained before the barrier lets anybody start
          // the next chunk, which overwrites that tile
          if (!one_tile && store_leader && store_pending) tma_store_wait_read0();
          named_bar_sync(bar_b, 128);
          if (store_leader) {
            const void* src = staging + (buf - staging_s);
            if (HALO) {
              tma_store_4d(&p.out_hi, src, c0, 0, y0, img);
              if (planes == 2) tma_store_4d(&p.out_lo, reinterpret_cast<const uint8_t*>(src) + 8192, c0, 0, y0, img);
            } else {
              tma_store_2d(&p.out_hi, src, c0, out_row0);
              if (planes == 2) tma_store_2d(&p.out_lo, reinterpret_cast<const uint8_t*>(src) + 8192, c0, out_row0);
            }
            tma_store_commit();
          }
          store_pending = true;
          continue;
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(raw[i]) * alpha : 0.f;
        if (ep_scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= (c0 + i < N) ? __ldg(ep_scale + c0 + i) : 0.f;
        }
        if (ep_bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += (c0 + i < N) ? __ldg(ep_bias + c0 + i) : 0.f;
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        // one staging tile: everybody must be done reading the previous chunk before it is overwritten; two tiles:
        // the barrier below (of the previous chunk) already guarantees that for the tile of two chunks ago
        if (one_tile) {
          if (store_leader && store_pending) tma_store_wait_read0();
          named_bar_sync(bar_a, 128);
        }
        if (valid) {
          // 128-byte row `srow`, 16-byte piece j stored at (j ^ (srow & 7)): SWIZZLE_128B, conflict-free
          const uint32_t rowp = buf + (uint32_t)srow * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_v4(rowp + (uint32_t)((j ^ (srow & 7)) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
        if (do_store) {
          fence_proxy_async_smem();
          if (!one_tile && store_leader && store_pending) tma_store_wait_read0();
        }
        named_bar_sync(bar_b, 128);
        if (do_store) {
          if (store_leader) {
            const void* src = staging + (buf - staging_s);
            if (HALO) tma_store_4d(&p.out, src, c0, 0, y0, img);
            else tma_store_2d(&p.out, src, c0, out_row0);
            tma_store_commit();
          }
          store_pending = true;
        }
        if (has_stats) {
          // BatchNorm sums from the staged tile: warp w adds up its 32 staging rows of column `lane` (conflict-free
          // 128-byte row reads, four independent chains) and folds the two partial sums into warp-private
          // double-float (sum, error) accumulators with an error-free TwoSum - fp64-grade accumulation at the cost of
          // a dozen FADDs (fp64 adds cost 25 % of the epilogue's stall samples on this part, profiles/r02_summary.md)
          const int r_end = min(32, nvalid - quarter * 32);
          float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f, qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
          const uint32_t base = buf + (uint32_t)((quarter * 32) * 128 + ((lane & 3) << 2));
          const int cj = lane >> 2;
          if (r_end == 32) {
#pragma unroll
            for (int r = 0; r < 32; r += 4) {
              const float x0 = lds_f32(base + (uint32_t)((r + 0) * 128 + ((cj ^ ((r + 0) & 7)) << 4)));
              const float x1 = lds_f32(base + (uint32_t)((r + 1) * 128 + ((cj ^ ((r + 1) & 7)) << 4)));
              const float x2 = lds_f32(base + (uint32_t)((r + 2) * 128 + ((cj ^ ((r + 2) & 7)) << 4)));
              const float x3 = lds_f32(base + (uint32_t)((r + 3) * 128 + ((cj ^ ((r + 3) & 7)) << 4)));
              sa += x0, sb += x1, sc += x2, sd += x3;
              qa = fmaf(x0, x0, qa), qb = fmaf(x1, x1, qb), qc = fmaf(x2, x2, qc), qd = fmaf(x3, x3, qd);
            }
          } else {
            for (int r = 0; r < r_end; ++r) {
              const float x0 = lds_f32(base + (uint32_t)(r * 128 + ((cj ^ (r & 7)) << 4)));
              sa += x0;
              qa = fmaf(x0, x0, qa);
            }
          }
          const uint32_t my = stats_s + (uint32_t)(quarter * BN + chunk * 32 + lane) * 16u;
          float4 t = lds_v4(my);
          two_sum_acc(t.x, t.y, (sa + sb) + (sc + sd));
          two_sum_acc(t.z, t.w, (qa + qb) + (qc + qd));
          sts_v4(my, t);
        }
      }
    }
    if (store_leader) tma_store_wait0();
    if (has_stats) {
      if (cur_n_blk >= 0) {
        named_bar_sync(1, ETHREADS);
        for (int i = etid; i < BN; i += ETHREADS) {
          double a = 0.0, b = 0.0;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float4 t = lds_v4(stats_s + (uint32_t)(w * BN + i) * 16u);
            a += (double)t.x + (double)t.y;
            b += (double)t.z + (double)t.w;
          }
          if (cur_n_blk * BN + i < N) {
            atomicAdd(&p.stats[cur_n_blk * BN + i], a);
            atomicAdd(&p.stats[N + cur_n_blk * BN + i], b);
          }
        }
      }
      if (p.bn_coef != nullptr) {
        // The last CTA to arrive owns the BatchNorm finalize: batch mean / biased variance -> per-channel
        // scale = gamma / sqrt(var + eps), shift = beta - mean * scale, running stats with the unbiased variance.
        __threadfence();
        named_bar_sync(1, ETHREADS);
        if (etid == 0) {
          const unsigned int ticket = atomicAdd(p.bn_counter, 1u);
          *last_flag = (ticket == gridDim.x - 1) ? 1u : 0u;
        }
        named_bar_sync(1, ETHREADS);
        if (*last_flag) {
          __threadfence();
          const int N = p.stat_channels;            // (shadows the kernel's column extent: they differ when EPI == 2)
          for (int c = etid; c < N; c += ETHREADS) {
            const double mean = __ldcg(&p.stats[c]) / p.bn_count;
            double var = __ldcg(&p.stats[N + c]) / p.bn_count - mean * mean;
            if (var < 0.0) var = 0.0;
            const double unbiased = p.bn_count > 1.0 ? var * (p.bn_count / (p.bn_count - 1.0)) : var;
            p.bn_running_mean[c] =
                (float)((1.0 - (double)p.bn_momentum) * (double)p.bn_running_mean[c] + (double)p.bn_momentum * mean);
            p.bn_running_var[c] =
                (float)((1.0 - (double)p.bn_momentum) * (double)p.bn_running_var[c] + (double)p.bn_momentum * unbiased);
            const float inv = (float)(1.0 / sqrt(var + (double)p.bn_eps));
            const float sc = p.bn_gamma[c] * inv;
            p.bn_coef[c] = sc;
            p.bn_coef[N + c] = p.bn_beta[c] - (float)mean * sc;
            if (p.bn_save) {                        // saved for the BatchNorm backward
              p.bn_coef[2 * N + c] = (float)mean;
              p.bn_coef[3 * N + c] = inv;
            }
          }
          if (etid == 0 && p.bn_nbt != nullptr) *p.bn_nbt += 1;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                 // no CTA of the pair may exit while the other can still signal it
  if (warp == 4) {
    tc_fence_after_sync();
    if (CG == 2) tmem_dealloc2(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// eval-mode BatchNorm: per-channel scale/shift from the running statistics
__global__ void bn_eval_coef_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                    float* __restrict__ coef, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv = (float)(1.0 / sqrt((double)rv[c] + (double)eps));
  const float sc = gamma[c] * inv;
  coef[c] = sc;
  coef[C + c] = beta[c] - rm[c] * sc;
}

int bn_eval_coef_launch(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, float* coef,
                        int C, cudaStream_t stream) {
  if (C == 0) return VB_OK;
  bn_eval_coef_kernel<<<(C + 127) / 128, 128, 0, stream>>>(gamma, beta, rm, rv, eps, coef, C);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
// per-device caches: one process may drive several GPUs (feature_extractor_gpu_ids[0] != model device, 2-GPU tests)
constexpr int MAX_DEVICES = 64;
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}
static int num_sms() {
  static int sms[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (sms[dev] == 0) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

// Output staging tiles: two (one per epilogue set) when a tile has at most 8 k-blocks, i.e. when the layer is bound by
// its epilogue and shared memory is not needed for a deep operand ring; four (two per set: one barrier per chunk instead
// of two) for plain GEMMs of one or two k-blocks; behind them the residual ring of the fetcher warps, when used.
// Apply epilogue with a residual on the small-K plain GEMMs (the two-pass route's recompute pass): the residual tiles are
// fetched by two extra warps into a shared-memory ring (see the kernel); needs the two epilogue sets' configuration
static bool use_res_fetch(const ConvGemmDesc& d) {
  if (!(d.out_hi != nullptr && d.res_kind != 0 && !d.im2col && d.kchunk == 0 && d.K / BK <= 8)) return false;
  const char* e = getenv("VINCE_B200_RES_FETCH");    // debug / A-B comparison: 0 = row owners load the residual themselves
  if (e && atoi(e) == 0) return false;
  e = getenv("VINCE_B200_EPI_SETS");
  if (e && atoi(e) == 1) return false;
  return true;
}
static int staging_tiles(const ConvGemmDesc& d) {
  // one or two k-blocks per tile: two tiles for each of the two epilogue warp sets (one barrier per chunk)
  int nbuf = (d.K / BK <= 2 && !d.im2col) ? 4 : ((d.K / BK <= 8) ? 2 : 1);
  // the stem (multi-tap im2col over the packed window layout, 4 k-blocks) is bound by its operand ring, not by the epilogue:
  // one tile and one epilogue set leave room for one more im2col stage (B200: 440 -> 399 us at B = 256)
  if (d.im2col && d.R * d.S > 1 && d.K / BK <= 8) nbuf = 1;
  if (d.res_slots > 0) nbuf = 2;                     // the residual ring takes the room of the second pair of tiles
  const char* e = getenv("VINCE_B200_STAGING");      // debug / A-B comparison: 1, 2 or 4
  if (e && (atoi(e) == 1 || atoi(e) == 2 || (atoi(e) == 4 && nbuf == 4))) nbuf = atoi(e);
  return nbuf;
}
static size_t staging_bytes(const ConvGemmDesc& d) {
  if (d.stats_only == 2) return 0;                   // transposed statistics pass: nothing is staged
  return (size_t)STAGING_BYTES * (staging_tiles(d) + d.res_slots);
}
static size_t fixed_smem(int bn, const ConvGemmDesc& d) {
  // alignment slack + staging + barriers + flags + per-channel area (EPI 0: [4 warps][2][bn] double-float sums,
  // EPI 1: 4 x bn coefficients, EPI 2: unused)
  const size_t chan = d.out_hi != nullptr ? (size_t)4 * bn * 4 + 16 : (d.stats_only == 2 ? 16 : (size_t)8 * bn * 8);
  return 1024 + staging_bytes(d) + (4 * MAX_RING + 12) * 8 + 32 + chan;
}

template <int BN, int CG, bool HALO, bool RES, int EPI, int EW = 1>
static int launch_gemm(const GemmKernelParams& kp, size_t smem, int grid, cudaStream_t stream) {
  const int fetch_threads = (EPI == 1 && kp.res_fetch) ? 64 : 0;       // two residual fetcher warps
  static bool smem_set[MAX_DEVICES] = {false};       // per instantiation AND device (the attribute is per context)
  const int dev = current_device();
  constexpr int THREADS = GEMM_THREADS + 128 * (EW - 1);
  if (!smem_set[dev]) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, CG, HALO, RES, EPI, EW>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET));
    smem_set[dev] = true;
  }
  if (CG == 1) {
    conv_gemm_kernel<BN, CG, HALO, RES, EPI, EW><<<grid, THREADS + fetch_threads, smem, stream>>>(kp);
  } else {
    grid &= ~1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS + fetch_threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, CG, HALO, RES, EPI, EW>, kp));
  }
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// Tile width policy: minimise waves x max(main loop, epilogue) per tile with measured per-MMA / per-k-block / per-chunk
// costs (profiles/r01_summary.md); wider tiles amortise better unless they leave SMs idle in the last wave.
static int auto_block_n(const ConvGemmDesc& d, int cg) {
  const int cands[3] = {256, 128, 64};
  const long m_blocks = ((d.M + BM - 1) / BM + cg - 1) / cg;
  const long units = num_sms() / cg;                 // CTAs (cg = 1) or CTA pairs (cg = 2) working concurrently
  long best_cost = -1;
  int best = 64;
  for (int bn : cands) {
    if (bn > 64 && d.N % bn != 0) continue;
    {
      // a 2-stage ring of (A tile + this CTA's weight share) must fit next to the fixed buffers
      const size_t planes = d.passes == 3 ? 2 : 1;
      const size_t stage = ((size_t)BM * 128 + (size_t)(bn / cg) * 128) * planes;
      // ... or, for a single n-block, the whole weight share resident next to two activation stages (RES mode)
      const size_t a_st = (size_t)BM * 128 * planes, b_st = (size_t)(bn / cg) * 128 * planes;
      const bool res_fits = !d.im2col && d.kchunk == 0 && bn >= d.N && m_blocks >= 2 * units &&
                            fixed_smem(bn, d) + (size_t)(d.K / BK) * b_st + 2 * a_st <= SMEM_BUDGET;
      if (bn > 64 && !res_fits && fixed_smem(bn, d) + 2 * stage > SMEM_BUDGET) continue;
    }
    long tiles = m_blocks * ((d.N + bn - 1) / bn);
    if (d.kchunk > 0) tiles *= (long)d.taps * ((d.K + d.kchunk - 1) / d.kchunk);     // batched split-K: every batch
    const long waves = (tiles + units - 1) / units;
    const long passes = d.passes == 3 ? 3 : 1;
    // per MMA: the tensor pipe needs bn/2 cycles, the issuing thread ~50 (measured with the 32-bit-descriptor issue
    // loop: N = 64 MMAs retire every ~50 cycles, N >= 128 at the tensor rate); per k-block ~120 cycles of barrier /
    // commit work.  With these constants the model reproduces the measured preference of 128-wide tiles for layer4
    // (98 tiles of 256 leave the second wave of 74 CTA pairs two-thirds idle: 158 -> 128 us) and of 256 for layer3.
    const long per_mma = bn / 2 > 50 ? bn / 2 : 50;
    const long mainloop = (long)(d.K / BK) * (4 * passes * per_mma + 120);
    const long epilogue = 1100L * (bn / 32);         // overlaps the next tile's main loop (double-buffered TMEM)
    const long cost = waves * (mainloop > epilogue ? mainloop : epilogue);
    // at most 4 k-blocks: the layer is bound by its output stream, and every halving of the tile width re-reads the
    // activations once more - take the widest tile that fits (measured: R50 128->512 161 us at 256 vs 202 us at 128)
    if (d.K / BK <= 4) return bn;
    if (best_cost < 0 || cost < best_cost) best_cost = cost, best = bn;
  }
  return best;
}

// Can this convolution run in HALO mode, and with which tile height?
static int halo_tile_rows(const ConvGemmDesc& d) {
  if (!d.im2col || d.halo_mode == 0) return 0;
  if (!(d.R == 3 && d.S == 3 && d.stride == 1 && d.pad_lo_h == 1 && d.pad_lo_w == 1 && d.pad_hi_h == 1 && d.pad_hi_w == 1))
    return 0;
  const int Wp = d.W + 2;
  if (Wp > BM) return 0;
  int th = BM / Wp;
  if (th > d.H) th = d.H;
  const int tiles_per_img = (d.H + th - 1) / th;
  // useful fraction of the MMA rows; below ~70% the plain im2col path wins
  const double eff = (double)(d.H * d.W) / ((double)tiles_per_img * BM);
  if (d.halo_mode < 0 && eff < 0.7) return 0;
  // two halo stages + at least two weight stages must fit in shared memory, else fall back to im2col
  int bn = d.block_n ? d.block_n : auto_block_n(d, ((d.M + BM - 1) / BM >= 8) ? 2 : 1);
  const size_t planes = d.passes == 3 ? 2 : 1;
  const size_t a_stage = (size_t)(((th + 2) * Wp + 7) / 8) * 1024 * planes;
  const size_t b_stage = (size_t)bn * 128 * planes;
  if (fixed_smem(bn, d) + 2 * a_stage + 2 * b_stage > SMEM_BUDGET) return 0;
  return th;
}

int conv_gemm_launch(const ConvGemmDesc& d_in, cudaStream_t stream) {
  ConvGemmDesc d = d_in;
  // Transposed statistics pass (stats_only == 2; 1x1 stride-1 convolutions / plain GEMMs only): the operand roles are
  // swapped - kernel rows = the N output channels (weights as the A operand), kernel columns = the M pixels
  // (activations as the B operand, 256 per tile) - so that every accumulator row is one channel (EPI = 2 above).
  const bool tstats = d.stats_only == 2;
  const int stat_channels = d.N;
  const double stat_count = (double)d.M;
  if (tstats) {
    VB_REQUIRE(!d.im2col && d.kchunk == 0, "conv_gemm: the transposed statistics pass handles plain [M,K] x [N,K] GEMMs");
    VB_REQUIRE(d.stats && !d.out && !d.out_hi && !d.scale && !d.bias && !d.relu,
               "conv_gemm: the transposed statistics pass produces BatchNorm sums only");
    d.a_hi = d_in.b_hi, d.a_lo = d_in.b_lo, d.b_hi = d_in.a_hi, d.b_lo = d_in.a_lo;
    d.M = d_in.N, d.N = d_in.M;
    d.block_n = 256;
  }
  VB_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "conv_gemm: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
  d.res_slots = use_res_fetch(d) ? 3 : 0;          // (tile-width decisions below account for a 3-slot ring)
  const bool batched = d.kchunk > 0;               // batched split-K GEMM (weight gradients)
  VB_REQUIRE(batched || d.K % BK == 0, "conv_gemm: K=%d must be a multiple of %d", d.K, BK);
  if (batched) {
    VB_REQUIRE(!d.im2col && !d.stats && !d.out_hi && !d.scale && !d.bias && d.out, "conv_gemm: batched split-K is a plain GEMM");
    VB_REQUIRE(d.kchunk % BK == 0 && (d.taps == 1 || d.taps == 9), "conv_gemm: bad split-K geometry");
    VB_REQUIRE(d.taps == 1 || (d.shift_w > 0 && d.shift_w % 8 == 0), "conv_gemm: shift_w must be a positive multiple of 8");
    VB_REQUIRE(d.K % 8 == 0, "conv_gemm: split-K needs a 16-byte aligned row pitch (K %% 8 == 0)");
  }
  VB_REQUIRE(tstats || d.N % 32 == 0, "conv_gemm: N=%d must be a multiple of 32", d.N);
  VB_REQUIRE(d.passes == 1 || d.passes == 3, "conv_gemm: passes must be 1 or 3");
  const bool planes_out = d.out_hi != nullptr;
  VB_REQUIRE(d.a_hi && d.b_hi && (d.out || planes_out || d.stats_only), "conv_gemm: null operand");
  VB_REQUIRE(!(planes_out && d.out), "conv_gemm: give either the fp32 output or the fp16 output planes");
  if (planes_out) {
    VB_REQUIRE(d.ep_coef, "conv_gemm: the apply epilogue needs per-channel (scale, shift) coefficients");
    VB_REQUIRE(d.passes == 1 || d.out_lo, "conv_gemm: fp16x3 needs the lo output plane");
    VB_REQUIRE(!d.stats && !d.scale && !d.bias && !d.stats_only, "conv_gemm: apply epilogue excludes stats / scale / bias");
    VB_REQUIRE(d.res_kind >= 0 && d.res_kind <= 2, "conv_gemm: res_kind %d", d.res_kind);
    VB_REQUIRE(d.res_kind != 1 || d.res_hi, "conv_gemm: residual planes null");
    VB_REQUIRE(d.res_kind != 2 || (d.res_raw && d.res_coef), "conv_gemm: residual raw / coefficients null");
  }
  VB_REQUIRE(!d.stats_only || d.stats, "conv_gemm: stats_only without a stats buffer");
  VB_REQUIRE(d.stats_only >= 0 && d.stats_only <= 2, "conv_gemm: stats_only must be 0, 1 or 2");
  VB_REQUIRE(d.passes == 1 || (d.a_lo && d.b_lo), "conv_gemm: fp16x3 needs lo planes");
  VB_REQUIRE(!(d.stats && (d.bias || d.scale || d.relu)), "conv_gemm: stats are defined on raw accumulators only");
  VB_REQUIRE(!d.bn_coef || (d.stats && d.bn_gamma && d.bn_beta && d.bn_running_mean && d.bn_running_var && d.bn_counter),
             "conv_gemm: BatchNorm finalize needs stats, gamma, beta, running stats and a counter");
  const size_t planes = d.passes == 3 ? 2 : 1;
  // CTA pairs (cta_group::2): worthwhile whenever there are at least a few 256-row tiles per SM pair
  int cg = ((d.M + BM - 1) / BM >= 8) ? 2 : 1;
  if (batched) cg = 1;
  // at most 4 k-blocks per tile: the layer is bound by its epilogue / output stream, where the pair's cross-CTA
  // hand-shakes only cost (stem: 430 -> 399 us measured with single CTAs); with at least 128 output channels and two k-blocks the pair's
  // halved weight traffic wins again (ResNet-50 256->1024: 114 -> 99 us, 128->512: 158 -> 144 us, r02 conv_variants)
  if (d.K <= 4 * BK && (d.N <= 64 || d.K <= BK) &&
      !(getenv("VINCE_B200_SMALLK_PAIR") && atoi(getenv("VINCE_B200_SMALLK_PAIR")) == 1))
    cg = 1;
  if (tstats) {
    // a pair covers 256 channels x 256 pixels per tile: each CTA loads its 128 weight rows and half of the pixel tile
    cg = (d.M % 256 == 0) ? 2 : 1;
    const char* e = getenv("VINCE_B200_TSTATS_PAIR");  // debug / A-B comparison: 0 single CTAs, 1 pairs when possible
    if (e && atoi(e) == 0) cg = 1;
  }
  {
    const char* e = getenv("VINCE_B200_CTA_PAIR");     // 0 forces single-CTA tiles (debug / A-B comparison)
    if (e && atoi(e) == 0) cg = 1;
  }
  int bn = d.block_n;
  if (bn == 0) bn = auto_block_n(d, cg);
  VB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "conv_gemm: block_n must be 64/128/256");

  GemmKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = d.M;
  kp.N = d.N;
  kp.num_m_blocks = (d.M + BM - 1) / BM;
  kp.num_n_blocks = (d.N + bn - 1) / bn;
  kp.passes = d.passes;
  kp.scale = d.scale;
  kp.bias = d.bias;
  kp.relu = d.relu;
  kp.alpha = d.alpha != 0.f ? d.alpha : 1.f;
  kp.stats = d.stats;
  kp.bn_gamma = d.bn_gamma, kp.bn_beta = d.bn_beta, kp.bn_running_mean = d.bn_running_mean;
  kp.bn_running_var = d.bn_running_var, kp.bn_nbt = reinterpret_cast<long long*>(d.bn_num_batches_tracked);
  kp.bn_coef = d.bn_coef, kp.bn_counter = d.bn_counter, kp.bn_momentum = d.bn_momentum, kp.bn_eps = d.bn_eps;
  kp.bn_count = stat_count;
  kp.stat_channels = stat_channels;
  kp.trace = reinterpret_cast<unsigned long long*>(d.trace);
  kp.stats_only = d.stats_only;
  kp.ep_coef = d.ep_coef;
  kp.res_kind = planes_out ? d.res_kind : 0;
  kp.res_hi = reinterpret_cast<const __half*>(d.res_hi), kp.res_lo = reinterpret_cast<const __half*>(d.res_lo);
  kp.res_raw = d.res_raw, kp.res_coef = d.res_coef;
  kp.debug_skip_mma = getenv("VINCE_B200_DEBUG_SKIP_MMA") ? atoi(getenv("VINCE_B200_DEBUG_SKIP_MMA")) : 0;
  kp.a_plane_bytes = BM * 128;
  kp.a_tx_bytes = BM * 128;
  kp.a_chunks = batched ? d.kchunk / BK : d.K / BK;
  kp.b_per_a = 1;
  kp.splits = 1, kp.kchunk = 0, kp.shift_w = 0, kp.out_batch_rows = 0;
  kp.alpha_dev = d.alpha_dev;
  kp.bn_save = d.bn_save;
  int nbatch = 1;
  if (batched) {
    kp.splits = (d.K + d.kchunk - 1) / d.kchunk;
    kp.kchunk = d.kchunk;
    kp.shift_w = d.taps == 9 ? d.shift_w : 0;
    kp.out_batch_rows = ((d.M + BM - 1) / BM) * BM;
    nbatch = d.taps * kp.splits;
  }
  int rc;
  const int th = halo_tile_rows(d);
  if (d.im2col) {
    VB_REQUIRE(d.Cin % BK == 0, "conv_gemm: Cin=%d must be a multiple of %d", d.Cin, BK);
    VB_REQUIRE(d.K == d.R * d.S * d.Cin, "conv_gemm: K != R*S*Cin");
    const int P = (d.H + d.pad_lo_h + d.pad_hi_h - d.R) / d.stride + 1;
    const int Q = (d.W + d.pad_lo_w + d.pad_hi_w - d.S) / d.stride + 1;
    VB_REQUIRE(d.M == d.batch * P * Q, "conv_gemm: M=%d != batch*P*Q=%d", d.M, d.batch * P * Q);
    kp.cin_blocks = d.Cin / BK;
  }
  if (th > 0) {
    // ---- HALO mode ----
    const int Wp = d.W + 2;
    kp.a_mode = 2;
    kp.Wp = Wp, kp.TH = th, kp.H = d.H, kp.W = d.W;
    kp.tiles_per_img = (d.H + th - 1) / th;
    kp.num_m_blocks = d.batch * kp.tiles_per_img;
    kp.a_chunks = kp.cin_blocks;
    kp.b_per_a = 9;
    const uint32_t halo_rows = (uint32_t)(th + 2) * Wp;
    kp.a_tx_bytes = halo_rows * 128;
    kp.a_plane_bytes = ((halo_rows + 7) / 8) * 1024;
    rc = encode_tma_4d_nhwc(&kp.a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.a_hi, d.batch, d.H, d.W, d.Cin, BK, Wp, th + 2,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (d.passes == 3) {
      rc = encode_tma_4d_nhwc(&kp.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.a_lo, d.batch, d.H, d.W, d.Cin, BK, Wp,
                              th + 2, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  } else if (d.im2col) {
    const int P = (d.H + d.pad_lo_h + d.pad_hi_h - d.R) / d.stride + 1;
    const int Q = (d.W + d.pad_lo_w + d.pad_hi_w - d.S) / d.stride + 1;
    kp.a_mode = 1;
    kp.PQ = P * Q;
    kp.Q = Q;
    kp.stride = d.stride;
    kp.pad_h = d.pad_lo_h;
    kp.pad_w = d.pad_lo_w;
    kp.S = d.S;
    rc = encode_tma_im2col_nhwc(&kp.a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.a_hi, d.batch, d.H, d.W, d.Cin, d.pad_lo_h,
                                d.pad_lo_w, d.pad_hi_h, d.pad_hi_w, d.R, d.S, d.stride, BK, BM,
                                CU_TENSOR_MAP_SWIZZLE_128B, d.a_pixel_stride, d.a_row_stride, d.a_img_stride);
    if (rc) return rc;
    if (d.passes == 3) {
      rc = encode_tma_im2col_nhwc(&kp.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.a_lo, d.batch, d.H, d.W, d.Cin,
                                  d.pad_lo_h, d.pad_lo_w, d.pad_hi_h, d.pad_hi_w, d.R, d.S, d.stride, BK, BM,
                                  CU_TENSOR_MAP_SWIZZLE_128B, d.a_pixel_stride, d.a_row_stride, d.a_img_stride);
      if (rc) return rc;
    }
  } else {
    kp.a_mode = 0;
    rc = encode_tma_2d(&kp.a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.a_hi, d.K, d.M, (uint64_t)d.K * 2, BK, BM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (d.passes == 3) {
      rc = encode_tma_2d(&kp.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.a_lo, d.K, d.M, (uint64_t)d.K * 2, BK, BM,
                         CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  // each CTA of a pair loads its half of the weight tile
  const uint64_t b_rows = (uint64_t)d.N * ((batched && d.taps == 9) ? 3 : 1);
  rc = encode_tma_2d(&kp.b_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.b_hi, d.K, b_rows, (uint64_t)d.K * 2, BK, bn / cg,
                     CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (d.passes == 3) {
    rc = encode_tma_2d(&kp.b_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.b_lo, d.K, b_rows, (uint64_t)d.K * 2, BK, bn / cg,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (kp.a_mode == 2) {
    const int th_ = kp.TH;
    if (planes_out) {
      rc = encode_tma_4d_nhwc(&kp.out_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.out_hi, d.batch, d.H, d.W, d.N, 32, d.W, th_,
                              CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      if (d.out_lo) {
        rc = encode_tma_4d_nhwc(&kp.out_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.out_lo, d.batch, d.H, d.W, d.N, 32, d.W,
                                th_, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
      }
    } else if (d.out) {
      rc = encode_tma_4d_nhwc(&kp.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.out, d.batch, d.H, d.W, d.N, 32, d.W, th_,
                              CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  } else {
    if (planes_out) {
      rc = encode_tma_2d(&kp.out_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.out_hi, d.N, d.M, (uint64_t)d.N * 2, 32, BM,
                         CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      if (d.out_lo) {
        rc = encode_tma_2d(&kp.out_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, d.out_lo, d.N, d.M, (uint64_t)d.N * 2, 32, BM,
                           CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
      }
    } else if (d.out) {
      const uint64_t out_rows = batched ? (uint64_t)nbatch * kp.out_batch_rows : (uint64_t)d.M;
      rc = encode_tma_2d(&kp.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.out, d.N, out_rows, (uint64_t)d.N * 4, 32, BM,
                         CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  // two epilogue warp sets where the epilogue is the bottleneck: few k-blocks per tile (the layers that already get two
  // staging tiles), never the 3x3 halo path; the transposed statistics pass has no staging at all
  int ew = ((staging_tiles(d) >= 2 || tstats) && kp.a_mode != 2) ? 2 : 1;
  {
    const char* e = getenv("VINCE_B200_EPI_SETS");     // debug / A-B comparison: 1 forces a single set
    if (e && atoi(e) == 1) ew = 1;
  }
  // ---- shared-memory budget: A ring + B ring ----
  const size_t a_stage = (size_t)kp.a_plane_bytes * planes;
  const size_t b_stage = (size_t)(bn / cg) * 128 * planes;
  if (d.res_slots > 0) {
    if (ew != 2) {
      d.res_slots = 0;                               // (e.g. forced single set): row owners load the residual themselves
    } else {
      // a fourth slot when it fits next to the minimum operand configuration, else make sure three do
      const size_t need = (kp.num_n_blocks == 1 && ((kp.num_m_blocks + cg - 1) / cg >= 2 * (num_sms() / cg)))
                              ? (size_t)kp.a_chunks * kp.b_per_a * b_stage + 2 * a_stage : 2 * (a_stage + b_stage);
      d.res_slots = 4;
      while (d.res_slots >= 2 && fixed_smem(bn, d) + need > SMEM_BUDGET) --d.res_slots;
      if (d.res_slots < 2) d.res_slots = 0;
    }
  }
  kp.res_fetch = d.res_slots > 0 ? 1 : 0;
  kp.res_slots = d.res_slots;
  kp.staging_bufs = staging_tiles(d);
  kp.staging_total = (uint32_t)staging_bytes(d);
  const size_t avail = SMEM_BUDGET - fixed_smem(bn, d);
  // resident weights: single n-block, this CTA's whole weight share + two activation stages fit, and every CTA
  // works through several tiles (otherwise the ring version overlaps the weight load just as well)
  const int total_kb = kp.a_chunks * kp.b_per_a;
  const bool res_ok = kp.num_n_blocks == 1 && (size_t)total_kb * b_stage + 2 * a_stage <= avail;
  bool res = !batched && res_ok && (kp.num_m_blocks + cg - 1) / cg >= 2 * (num_sms() / cg);
  {
    const char* e = getenv("VINCE_B200_RESIDENT");     // debug / A-B comparison: 0 disables, 2 forces when feasible
    if (e && atoi(e) == 0) res = false;
    if (e && atoi(e) == 2) res = res_ok && !batched;
  }
  if (tstats) res = false;
  if (res) {
    kp.b_stages = total_kb;
    int st = (int)((avail - (size_t)total_kb * b_stage) / a_stage);
    kp.a_stages = st > MAX_RING ? MAX_RING : st;
  } else if (kp.a_mode == 2) {
    kp.a_stages = 2;
    VB_REQUIRE(avail > 2 * a_stage + 2 * b_stage, "conv_gemm: halo tile does not fit in shared memory");
    int bs = (int)((avail - 2 * a_stage) / b_stage);
    kp.b_stages = bs > MAX_RING ? MAX_RING : bs;
  } else {
    int st = (int)(avail / (a_stage + b_stage));
    if (st > MAX_RING) st = MAX_RING;
    VB_REQUIRE(st >= 2, "conv_gemm: not enough shared memory for a 2-stage pipeline (bn=%d)", bn);
    kp.a_stages = kp.b_stages = st;
  }
  const size_t smem = fixed_smem(bn, d) + kp.a_stages * a_stage + kp.b_stages * b_stage;
  // 128-row blocks -> (128*cg)-row tiles; a pair whose second half lies past the end loads zeros and stores nothing
  kp.num_m_blocks = (kp.num_m_blocks + cg - 1) / cg;
  kp.tiles_per_batch = kp.num_m_blocks * kp.num_n_blocks;
  kp.num_tiles = kp.tiles_per_batch * nbatch;
  const int tiles = kp.num_tiles;
  int grid = tiles * cg < num_sms() ? tiles * cg : num_sms();
  if (getenv("VINCE_B200_DEBUG_GRID") && atoi(getenv("VINCE_B200_DEBUG_GRID")) < grid) grid = atoi(getenv("VINCE_B200_DEBUG_GRID"));
  const bool halo = kp.a_mode == 2;
  if (cg == 2) grid &= ~1;
  if (tstats) {
    if (ew == 2)
      return cg == 2 ? launch_gemm<256, 2, false, false, 2, 2>(kp, smem, grid, stream)
                     : launch_gemm<256, 1, false, false, 2, 2>(kp, smem, grid, stream);
    return cg == 2 ? launch_gemm<256, 2, false, false, 2>(kp, smem, grid, stream)
                   : launch_gemm<256, 1, false, false, 2>(kp, smem, grid, stream);
  }
#define VB_LAUNCH_E(BN_, CG_, EPI_)                                                             \
  {                                                                                             \
    if (halo) return res ? launch_gemm<BN_, CG_, true, true, EPI_>(kp, smem, grid, stream)      \
                         : launch_gemm<BN_, CG_, true, false, EPI_>(kp, smem, grid, stream);    \
    if (ew == 2) return res ? launch_gemm<BN_, CG_, false, true, EPI_, 2>(kp, smem, grid, stream)    \
                            : launch_gemm<BN_, CG_, false, false, EPI_, 2>(kp, smem, grid, stream);  \
    return res ? launch_gemm<BN_, CG_, false, true, EPI_>(kp, smem, grid, stream)               \
               : launch_gemm<BN_, CG_, false, false, EPI_>(kp, smem, grid, stream);             \
  }
#define VB_LAUNCH(BN_, CG_)                  \
  if (bn == BN_ && cg == CG_) {              \
    if (planes_out) VB_LAUNCH_E(BN_, CG_, 1) \
    VB_LAUNCH_E(BN_, CG_, 0)                 \
  }
  VB_LAUNCH(64, 1) VB_LAUNCH(128, 1) VB_LAUNCH(256, 1) VB_LAUNCH(64, 2) VB_LAUNCH(128, 2) VB_LAUNCH(256, 2)
#undef VB_LAUNCH
#undef VB_LAUNCH_E
  VB_REQUIRE(false, "conv_gemm: no kernel for bn=%d cg=%d", bn, cg);
}

}  // namespace vb
