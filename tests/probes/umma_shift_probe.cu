// Hardware probe (test infrastructure, not part of the library): does a K-major SWIZZLE_128B UMMA shared-memory
// descriptor whose start address is shifted by s rows (s * 128 bytes, s not a multiple of 8) read rows s..s+127 of a
// tile that TMA wrote with the matching swizzle?  And does it need the descriptor's base_offset field?
// Answer decides whether a 3x3 stride-1 convolution can reuse ONE halo tile in smem for all 9 filter taps.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I vince_b200/csrc tests/probes/umma_shift_probe.cu \
//        vince_b200/csrc/host.cu -o gpurun_out/umma_shift_probe -lcuda   &&   ./gpurun_out/umma_shift_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace vb;

constexpr int ROWS = 256, K = 64, N = 64;

struct Params {
  CUtensorMap a_map, b_map;
  float* out;   // [variants][shifts][128][64]
  int shifts[16];
  int n_shifts;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;                       // 256 rows * 128 B
  uint8_t* b_s = smem + ROWS * 128;          // 64 rows * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + N * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * 128 + N * 128);
    tma_load_2d(a_s, &p.a_map, bar, 0, 0);
    tma_load_2d(b_s, &p.b_map, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int si = 0; si < p.n_shifts; ++si) {
      const int s = p.shifts[si];
      if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc(UMMA_FMT_F16, 128, N);
        for (int k = 0; k < K / 16; ++k) {
          uint64_t da = make_smem_desc(smem_u32(a_s) + s * 128 + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
          if (variant == 1) da |= (uint64_t)(s & 7) << 49;          // base_offset
          const uint64_t db = make_smem_desc(smem_u32(b_s) + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
          umma_f16(tmem, da, db, idesc, k != 0);
        }
        umma_commit(mma_bar);
      }
      mbar_wait(mma_bar, phase);
      phase ^= 1;
      tc_fence_after_sync();
      float* dst = p.out + ((size_t)(variant * p.n_shifts + si) * 128 + warp * 32 + lane) * N;
      for (int c = 0; c < N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) dst[c * 32 + i] = __uint_as_float(r[i]);
      }
      tc_fence_before_sync();
      __syncthreads();
      tc_fence_after_sync();
    }
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> a(ROWS * K), b(N * K);
  std::vector<float> af(ROWS * K), bf(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) { af[i] = (float)(rand() % 17 - 8); a[i] = __float2half(af[i]); }
  for (int i = 0; i < N * K; ++i) { bf[i] = (float)(rand() % 9 - 4); b[i] = __float2half(bf[i]); }
  __half *da, *db;
  cudaMalloc(&da, a.size() * 2);
  cudaMalloc(&db, b.size() * 2);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  Params p;
  int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 58, 59, 60, 116, 117, 118};
  p.n_shifts = sizeof(shifts) / sizeof(int);
  for (int i = 0; i < p.n_shifts; ++i) p.shifts[i] = shifts[i];
  const size_t out_elems = (size_t)2 * p.n_shifts * 128 * N;
  cudaMalloc(&p.out, out_elems * 4);
  cudaMemset(p.out, 0xff, out_elems * 4);
  if (encode_tma_2d(&p.a_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, da, K, ROWS, K * 2, 64, ROWS, CU_TENSOR_MAP_SWIZZLE_128B) ||
      encode_tma_2d(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, db, K, N, K * 2, 64, N, CU_TENSOR_MAP_SWIZZLE_128B)) {
    printf("tma encode failed: %s\n", get_error());
    return 1;
  }
  const int smem = 1024 + ROWS * 128 + N * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(out_elems);
  cudaMemcpy(out.data(), p.out, out_elems * 4, cudaMemcpyDeviceToHost);
  for (int variant = 0; variant < 2; ++variant)
    for (int si = 0; si < p.n_shifts; ++si) {
      const int s = shifts[si];
      int bad = 0, checked = 0;
      for (int i = 0; i < 128 && i + s < ROWS; ++i)
        for (int j = 0; j < N; ++j) {
          float ref = 0.f;
          for (int k = 0; k < K; ++k) ref += af[(i + s) * K + k] * bf[j * K + k];
          const float got = out[((size_t)(variant * p.n_shifts + si) * 128 + i) * N + j];
          ++checked;
          if (got != ref) ++bad;
        }
      printf("shift %3d base_offset=%s : %s (%d / %d mismatches)\n", s, variant ? "s&7" : "0  ", bad ? "WRONG" : "ok", bad,
             checked);
    }
  return 0;
}
