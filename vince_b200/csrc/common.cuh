// Blackwell (sm_100a) primitives shared by the vince_b200 kernels: mbarrier, TMA, tcgen05/TMEM wrappers,
// UMMA descriptor builders, error plumbing.  Hand-written inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vb {

// ------------------------------------------------------------------------------------------------
// error plumbing (C ABI: every entry point returns 0 or a negative code; message via vince_last_error)
// ------------------------------------------------------------------------------------------------
enum : int {
  VB_OK = 0,
  VB_ERR_INVALID = -1,    // bad argument / unsupported shape
  VB_ERR_CUDA = -2,       // CUDA runtime / driver error
  VB_ERR_NCCL = -3,       // NCCL error or libnccl not loadable
  VB_ERR_UNSUPPORTED = -4
};

void set_error(const char* fmt, ...);
const char* get_error();

#define VB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::vb::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ::vb::VB_ERR_CUDA;                                                               \
    }                                                                                         \
  } while (0)

#define VB_REQUIRE(cond, ...)                \
  do {                                       \
    if (!(cond)) {                           \
      ::vb::set_error(__VA_ARGS__);          \
      return ::vb::VB_ERR_INVALID;           \
    }                                        \
  } while (0)

// cuTensorMapEncode* are driver-API entry points; resolved at run time through cudart so the
// library has no link-time dependency on libcuda (it must dlopen on a GPU-less box).
int encode_tma_2d(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, uint64_t inner, uint64_t outer,
                  uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz);
int encode_tma_im2col_nhwc(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int N, int H, int W, int C,
                           int pad_lo_h, int pad_lo_w, int pad_hi_h, int pad_hi_w, int R, int S, int stride,
                           uint32_t channels_per_pixel, uint32_t pixels_per_column, CUtensorMapSwizzle swz,
                           long long pixel_stride = 0, long long row_stride = 0, long long img_stride = 0);

// NHWC tensor [N,H,W,C] as a 4-D tiled map with box {box_c, box_w, box_h, 1} (halo loads / NHWC tile stores)
int encode_tma_4d_nhwc(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int N, int H, int W, int C,
                       uint32_t box_c, uint32_t box_w, uint32_t box_h, CUtensorMapSwizzle swz);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred;
}

// ---- explicit shared-memory accesses ----
// The dynamic shared-memory base is re-aligned through an integer round trip, after which the compiler no longer
// knows the address space and emits GENERIC loads / stores (LD.E / ST.E) for ordinary pointer dereferences.  These
// wrappers take 32-bit shared addresses and emit LDS / STS.  Loads are `asm volatile` without a memory clobber: they
// keep their order relative to the (volatile) barrier instructions but do not serialise surrounding arithmetic.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// ---- cp.async (LDGSTS): 16-byte asynchronous global -> shared copies, no register staging ----
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
// same, asking L2 to fetch the surrounding 256 bytes (the neighbouring channel chunks of the row are wanted next)
__device__ __forceinline__ void cp_async_16_l2_256(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
// this thread's arrival on `bar` is deferred until all of its prior cp.async copies have landed (the barrier's arrival
// count must include it: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Error-free accumulation (Knuth TwoSum): (s, e) += x with s + e carrying the exact running sum to ~2^-48 relative.
__device__ __forceinline__ void two_sum_acc(float& s, float& e, float x) {
  const float t = __fadd_rn(s, x);
  const float bp = __fadd_rn(t, -s);
  const float err = __fadd_rn(__fadd_rn(s, -__fadd_rn(t, -bp)), __fadd_rn(x, -bp));
  e = __fadd_rn(e, err);
  s = t;
}

// 32 bytes from global memory in one request (sm_100 256-bit load), read-only path, no L1 allocation, L2 asked to fetch the
// surrounding 256 bytes: for streams whose neighbours are wanted a little later by the same SM (`p` 32-byte aligned)
__device__ __forceinline__ void ldg256_stream(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}

// fire-and-forget request to bring the 128-byte line holding `p` into L2
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug (wrong expect_tx byte count, missing arrive) would otherwise hang the GPU until
// the host watchdog fires.  ~4e9 cycles (~2 s) is far beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("vince_b200: mbarrier wait timed out (block %d thread %d smem 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// same, on a shared-memory address (keeps barrier indexing in 32-bit uniform arithmetic)
__device__ __forceinline__ uint32_t mbar_try_wait_addr(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok;
}
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait_addr(bar_addr, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("vince_b200: mbarrier wait timed out (block %d thread %d smem 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar_addr, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  if (mbar_try_wait_addr(bar_addr, parity)) return;
  if (mbar_try_wait_addr(bar_addr, parity)) return;
  mbar_wait_slow(bar_addr, parity);
}

// ---- proxies / fences ----
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// NHWC im2col load: coords {c, w, h, n} of the base pixel (in input space, may be negative), filter offsets {s, r}
__device__ __forceinline__ void tma_load_im2col_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c,
                                                   int32_t w, int32_t h, int32_t n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- UMMA (tcgen05.mma), single-CTA, A and B from shared memory (64-bit descriptor form, InfoNCE kernels) ----
// D[tmem] (+)= A[smem] * B[smem]^T ; `accumulate`==0 overwrites D.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05 ops of this thread complete
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on one TPC drive one 256-row MMA; only the leader (cluster
// rank 0) issues tcgen05.mma, each CTA loads its own A rows and HALF of the B tile, all "full" barriers live in the
// leader, "empty"/"tmem_full" barriers are signalled in both CTAs by a multicast commit. ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `cta_rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
// Remote arrive WITHOUT release semantics: a cluster-scope release makes ptxas emit a fence + L1 invalidate
// (~0.6 us, measured - it serialised the whole CTA-pair pipeline).  The data this signals was written by TMA and
// is consumed by the tensor core (both async proxy), ordered by the mbarrier phases themselves.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// tcgen05.mma with the shared-memory descriptors given as 32-bit words: low word per operand (address >> 4 | LBO),
// one common high word (SBO | version | swizzle mode).  CG = 1: single CTA; CG = 2: CTA pair (issued by the leader).
template <int CG>
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  if (CG == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi)
        : "memory");
  }
}
// commit on a barrier given by its shared-memory address (CG = 2: same offset in both CTAs of the pair)
template <int CG>
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
  if (CG == 2) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar_addr), "h"(mask)
                 : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
  }
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t <- lane base+t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ----
// Shared-memory matrix descriptor for a K-major operand tile whose rows are one swizzle span wide
// (128 B rows for SWIZZLE_128B, written by TMA with the matching swizzle).  8-row core-matrix groups are
// `sbo_bytes` apart (8 * row bytes for a dense tile).  Bit layout: start>>4 [0,14) | LBO>>4 [16,30) |
// SBO>>4 [32,46) | version=1 [46,48) | layout [61,64) (2 = SWIZZLE_128B, 4 = 64B, 6 = 32B, 0 = none).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
constexpr uint32_t UMMA_LAYOUT_SW128 = 2;

// Instruction descriptor (kind::f16 / kind::tf32), fp32 accumulate, A and B K-major.
// c_format=F32 bit[4,6)=1 | a_format [7,10) | b_format [10,13) | N>>3 [17,23) | M>>4 [24,29)
constexpr uint32_t UMMA_FMT_F16 = 0, UMMA_FMT_TF32 = 2;
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// fp16 hi/lo split of an fp32 value: hi = fp16(x) (11-bit significand), lo = fp16(x - hi), so x ~= hi + lo with
// |x - hi - lo| <= 2^-24 |x| while lo stays a normal fp16 number (|x| >~ 0.25) and <= 3e-8 absolute below that
// (fp16 subnormal spacing) - fp32-grade for the O(1) activations / O(1e-2) weights of a BatchNorm ResNet.  The three
// products hi*hi + hi*lo + lo*hi are exact in the fp32 accumulator; the dropped lo*lo term is <= 2^-24 relative.
// Conversions saturate at +-65504 instead of overflowing to inf.
__device__ __forceinline__ __half f16_sat(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return __ushort_as_half(h);
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = f16_sat(x);
  lo = f16_sat(x - __half2float(hi));
}
#endif  // __CUDACC__

}  // namespace vb
