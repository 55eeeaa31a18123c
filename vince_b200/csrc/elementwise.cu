// HBM-bound kernels of the encoder + queue path: layout packing, weight preparation, train/eval BatchNorm apply
// (+ReLU, +residual, +max/avg pooling), bf16 hi/lo splitting, L2 normalisation, jigsaw gathers, and the fused
// multi-tensor momentum-EMA + ring-buffer enqueue.  All are coalesced, vectorised (16-byte) streaming kernels.
#include "common.cuh"
#include "kernels.h"

namespace vb {

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

struct alignas(8) bf16x4 {
  __nv_bfloat16 v[4];
};
struct alignas(16) bf16x8 {
  __nv_bfloat16 v[8];
};

__device__ __forceinline__ void split4(const float4& f, bf16x4& hi, bf16x4& lo) {
  split_bf16(f.x, hi.v[0], lo.v[0]);
  split_bf16(f.y, hi.v[1], lo.v[1]);
  split_bf16(f.z, hi.v[2], lo.v[2]);
  split_bf16(f.w, hi.v[3], lo.v[3]);
}
__device__ __forceinline__ float4 join4(const bf16x4& hi, const bf16x4& lo) {
  return make_float4(__bfloat162float(hi.v[0]) + __bfloat162float(lo.v[0]),
                     __bfloat162float(hi.v[1]) + __bfloat162float(lo.v[1]),
                     __bfloat162float(hi.v[2]) + __bfloat162float(lo.v[2]),
                     __bfloat162float(hi.v[3]) + __bfloat162float(lo.v[3]));
}

// ------------------------------------------------------------------------------------------------
// stem packing: one block per (image n, row pair j).  The six input rows (2 image rows x 3 channels) are staged
// in shared memory with coalesced loads (every input row belongs to exactly one j, so the image is read once),
// then each thread emits 16-byte groups of the 64-element packed "pixel" for consecutive q.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x,
                                                        const int64_t* __restrict__ gather_idx,
                                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                        int H, int W, int Hj, int Q) {
  extern __shared__ float rows[];                 // [6][W + 6], 3 zero columns of left padding
  __shared__ int lut[64];                         // element e -> offset of (r2, c, s) inside `rows`, or -1
  const int j = blockIdx.x;
  const int n = blockIdx.y;
  const int Wp = W + 6;
  const int tid = threadIdx.x;
  if (tid < 64) {
    int off = -1;
    if (tid < 42) {
      const int r2 = tid / 21, rem = tid - r2 * 21, s = rem / 3, c = rem - s * 3;
      off = (r2 * 3 + c) * Wp + s;
    }
    lut[tid] = off;
  }
  const int64_t src_n = gather_idx ? gather_idx[n] : n;
  const float* xn = x + src_n * 3 * (int64_t)H * W;
  for (int i = tid; i < 6 * Wp; i += blockDim.x) {
    const int rc = i / Wp, col = i - rc * Wp - 3;
    const int r2 = rc / 3, c = rc - r2 * 3;
    const int row = 2 * j - 1 + r2;
    float v = 0.f;
    if (row >= 0 && row < H && col >= 0 && col < W) v = __ldg(xn + ((int64_t)c * H + row) * W + col);
    rows[i] = v;
  }
  __syncthreads();
  const int64_t base = (((int64_t)n * Hj + j) * Q) * 64;
  for (int g = tid; g < Q * 8; g += blockDim.x) {
    const int q = g >> 3, grp = g & 7;
    bf16x8 oh, ol;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int off = lut[grp * 8 + i];
      const float val = off >= 0 ? rows[off + 2 * q] : 0.f;
      split_bf16(val, oh.v[i], ol.v[i]);
    }
    *reinterpret_cast<bf16x8*>(hi + base + (int64_t)g * 8) = oh;
    if (lo) *reinterpret_cast<bf16x8*>(lo + base + (int64_t)g * 8) = ol;
  }
}

int stem_pack_launch(const float* x, const int64_t* gather_idx, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int H,
                     int W, int Hj, int Q, cudaStream_t stream) {
  if (N == 0) return VB_OK;
  VB_REQUIRE(N <= 65535, "stem_pack: batch %d too large", N);
  const size_t smem = (size_t)6 * (W + 6) * sizeof(float);
  VB_REQUIRE(smem <= 48 * 1024, "stem_pack: image width %d too large", W);
  dim3 grid(Hj, N);
  stem_pack_kernel<<<grid, 256, smem, stream>>>(x, gather_idx, hi, lo, H, W, Hj, Q);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// weight preparation: OIHW fp32 -> K-major [Cout][R][S][Cin] bf16 hi/lo (or the packed stem layout)
// ------------------------------------------------------------------------------------------------
__global__ void weight_prep_kernel(const WeightPrepEntry* __restrict__ table, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo) {
  const WeightPrepEntry e = table[blockIdx.y];
  const int64_t K = e.kind == 1 ? 256 : (int64_t)e.R * e.S * e.Cin;
  const int64_t total = (int64_t)e.Cout * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / K);
    const int k = (int)(i - (int64_t)co * K);
    float w = 0.f;
    if (e.kind == 0) {
      const int rs = k / e.Cin;
      const int ci = k - rs * e.Cin;
      w = __ldg(e.src + ((int64_t)co * e.Cin + ci) * (e.R * e.S) + rs);
    } else {
      // stem: k = t*64 + r2*21 + s*3 + c  <-  w[co, c, r = 2t + r2, s]   (7x7, Cin = 3)
      const int t = k >> 6;
      const int el = k & 63;
      if (el < 42) {
        const int r2 = el / 21;
        const int rem = el - r2 * 21;
        const int s = rem / 3;
        const int c = rem - s * 3;
        const int r = 2 * t + r2;
        if (r < 7) w = __ldg(e.src + (((int64_t)co * 3 + c) * 7 + r) * 7 + s);
      }
    }
    __nv_bfloat16 h, l;
    split_bf16(w, h, l);
    hi[e.dst_off + i] = h;
    if (lo) lo[e.dst_off + i] = l;
  }
}

int weight_prep_launch(const WeightPrepEntry* table_dev, int n_entries, int64_t max_elems, __nv_bfloat16* hi,
                       __nv_bfloat16* lo, cudaStream_t stream) {
  if (n_entries == 0) return VB_OK;
  int bx = div_up(max_elems, 256 * 8);
  if (bx > 1024) bx = 1024;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_entries);
  weight_prep_kernel<<<grid, 256, 0, stream>>>(table_dev, hi, lo);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm is applied from per-channel (scale, shift) pairs `coef[2][C]` produced by the convolution kernel's fused
// finalize (train mode) or by bn_eval_coef (eval mode): these kernels are pure fp32 streaming passes.
// ------------------------------------------------------------------------------------------------
struct BnSideDev {
  const float* raw;
  const float* coef;
};
static BnSideDev to_dev(const BnSide& s) {
  BnSideDev d;
  d.raw = s.raw, d.coef = s.coef;
  return d;
}
__device__ __forceinline__ void bn_coeffs4(const BnSideDev& s, int c, int C, float4& sc, float4& sh) {
  sc = __ldg(reinterpret_cast<const float4*>(s.coef + c));
  sh = __ldg(reinterpret_cast<const float4*>(s.coef + C + c));
}

__device__ __forceinline__ float4 fma4(const float4& x, const float4& a, const float4& b) {
  return make_float4(fmaf(x.x, a.x, b.x), fmaf(x.y, a.y, b.y), fmaf(x.z, a.z, b.z), fmaf(x.w, a.w, b.w));
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 relu4(const float4& a) {
  return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// ------------------------------------------------------------------------------------------------
// bn_apply: out = relu?( bn(main) + residual )  ->  bf16 hi/lo planes (and/or fp32)
// thread t owns channel group (t % C4) and strides over rows; total threads is a multiple of C4.
// ------------------------------------------------------------------------------------------------
__global__ void bn_apply_kernel(BnSideDev main, int res_kind, const __nv_bfloat16* __restrict__ res_hi,
                                const __nv_bfloat16* __restrict__ res_lo, BnSideDev res_bn, int relu,
                                __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                float* __restrict__ out_f32, int64_t M, int C) {
  const int C4 = C >> 2;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
  const int cg = (int)(gtid % C4);
  const int64_t row0 = gtid / C4;
  const int64_t row_stride = gthreads / C4;
  const int c = cg * 4;
  float4 sc, sh, rsc, rsh;
  bn_coeffs4(main, c, C, sc, sh);
  if (res_kind == 2) bn_coeffs4(res_bn, c, C, rsc, rsh);
  for (int64_t r = row0; r < M; r += row_stride) {
    const int64_t off = r * C + c;
    float4 v = fma4(__ldcs(reinterpret_cast<const float4*>(main.raw + off)), sc, sh);
    if (res_kind == 1) {
      const bf16x4 h = *reinterpret_cast<const bf16x4*>(res_hi + off);
      bf16x4 l;
      if (res_lo) l = *reinterpret_cast<const bf16x4*>(res_lo + off);
      else l.v[0] = l.v[1] = l.v[2] = l.v[3] = __float2bfloat16_rn(0.f);
      v = add4(v, join4(h, l));
    } else if (res_kind == 2) {
      v = add4(v, fma4(__ldcs(reinterpret_cast<const float4*>(res_bn.raw + off)), rsc, rsh));
    }
    if (relu) v = relu4(v);
    if (out_hi) {
      bf16x4 h, l;
      split4(v, h, l);
      *reinterpret_cast<bf16x4*>(out_hi + off) = h;
      if (out_lo) *reinterpret_cast<bf16x4*>(out_lo + off) = l;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = v;
  }
}

static int elementwise_grid(int64_t work_threads, int C4, int threads) {
  // a grid whose total thread count is a multiple of C4 (lcm handling: threads=256, C4 is a power of two <= 512
  // for every ResNet width; otherwise fall back to one row per C4 threads with padding handled by the caller)
  int64_t blocks = (work_threads + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // make blocks*threads a multiple of C4
  int64_t unit = C4 / threads;       // blocks must be a multiple of this when C4 > threads
  if (unit > 1) blocks = ((blocks + unit - 1) / unit) * unit;
  return (int)blocks;
}

int bn_apply_launch(const BnSide& main, int res_kind, const __nv_bfloat16* res_hi, const __nv_bfloat16* res_lo,
                    const BnSide& res_bn, int relu, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, float* out_f32,
                    int64_t M, int C, cudaStream_t stream) {
  VB_REQUIRE(C % 4 == 0, "bn_apply: C=%d must be a multiple of 4", C);
  const int C4 = C / 4;
  const int threads = 256;
  VB_REQUIRE((C4 <= threads && threads % C4 == 0) || (C4 > threads && C4 % threads == 0),
             "bn_apply: unsupported channel count %d", C);
  if (M == 0) return VB_OK;
  const int blocks = elementwise_grid(M * C4, C4, threads);
  bn_apply_kernel<<<blocks, threads, 0, stream>>>(to_dev(main), res_kind, res_hi, res_lo, to_dev(res_bn), relu, out_hi,
                                                  out_lo, out_f32, M, C);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// stem: bn + relu + maxpool 3x3 stride 2 pad 1 (NHWC); padding behaves as -inf (torch max_pool2d).
// One block per (image, output row, tile of MP_TQ output columns): the 3 x (2*MP_TQ+1) input pixels are normalised
// once while being staged in shared memory (coalesced 16-byte loads), then reduced.
// ------------------------------------------------------------------------------------------------
constexpr int MP_TQ = 14;
__global__ void __launch_bounds__(256) bn_relu_maxpool_kernel(BnSideDev bn, __nv_bfloat16* __restrict__ out_hi,
                                                              __nv_bfloat16* __restrict__ out_lo, int N, int P, int Q,
                                                              int C, int P2, int Q2) {
  extern __shared__ float4 tile4[];                   // [3][2*MP_TQ+1][C/4]
  const int C4 = C >> 2;
  const int q2_0 = blockIdx.x * MP_TQ;
  const int p2 = blockIdx.y;
  const int n = blockIdx.z;
  const int tid = threadIdx.x;
  const int cols = 2 * MP_TQ + 1;
  const int cg = tid % C4;
  float4 sc, sh;
  bn_coeffs4(bn, cg * 4, C, sc, sh);
  const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  // blockDim (256) is a multiple of C4, so a thread always serves the same channel group
  for (int i = tid; i < 3 * cols * C4; i += blockDim.x) {
    const int pix = i / C4;
    const int dy = pix / cols, dx = pix - dy * cols;
    const int y = 2 * p2 - 1 + dy, xx = 2 * q2_0 - 1 + dx;
    float4 v = ninf;
    if (y >= 0 && y < P && xx >= 0 && xx < Q)
      v = relu4(fma4(__ldcs(reinterpret_cast<const float4*>(bn.raw + (((int64_t)n * P + y) * Q + xx) * C) + cg), sc, sh));
    tile4[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < MP_TQ * C4; i += blockDim.x) {
    const int t = i / C4;
    const int q2 = q2_0 + t;
    if (q2 >= Q2) break;
    float4 m = ninf;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) m = max4(m, tile4[(dy * cols + 2 * t + dx) * C4 + cg]);
    bf16x4 h, l;
    split4(m, h, l);
    const int64_t off = ((((int64_t)n * P2 + p2) * Q2) + q2) * C + cg * 4;
    *reinterpret_cast<bf16x4*>(out_hi + off) = h;
    if (out_lo) *reinterpret_cast<bf16x4*>(out_lo + off) = l;
  }
}

int bn_relu_maxpool_launch(const BnSide& bn, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int N, int P, int Q, int C,
                           int P2, int Q2, cudaStream_t stream) {
  VB_REQUIRE(C % 4 == 0 && C <= 256 && 256 % (C / 4) == 0, "bn_relu_maxpool: unsupported channel count %d", C);
  VB_REQUIRE(N <= 65535 && P2 <= 65535, "bn_relu_maxpool: shape too large");
  if ((int64_t)N * P2 * Q2 == 0) return VB_OK;
  const size_t smem = (size_t)3 * (2 * MP_TQ + 1) * (C / 4) * sizeof(float4);
  VB_REQUIRE(smem <= 48 * 1024, "bn_relu_maxpool: channel count %d too large", C);
  dim3 grid((Q2 + MP_TQ - 1) / MP_TQ, P2, N);
  bn_relu_maxpool_kernel<<<grid, 256, smem, stream>>>(to_dev(bn), out_hi, out_lo, N, P, Q, C, P2, Q2);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// last residual block: relu(bn(main)+residual) -> NCHW fp32 spatial features + global average pool
// block = (image n, 32-channel slab); smem transpose so both the NHWC read and the NCHW write are coalesced
// ------------------------------------------------------------------------------------------------
constexpr int FP_CH = 32;
__global__ void bn_final_pool_kernel(BnSideDev main, int res_kind, const __nv_bfloat16* __restrict__ res_hi,
                                     const __nv_bfloat16* __restrict__ res_lo, BnSideDev res_bn,
                                     const int64_t* __restrict__ scatter_idx, float* __restrict__ spatial,
                                     float* __restrict__ pooled, int N, int HW, int C) {
  extern __shared__ float tile[];                 // [FP_CH][HW + 1]
  const int n = blockIdx.y;
  const int c_base = blockIdx.x * FP_CH;
  const int tid = threadIdx.x;
  const int cg = tid & 7;                         // 8 channel groups of 4
  const int c = c_base + cg * 4;
  float4 sc, sh, rsc, rsh;
  bn_coeffs4(main, c, C, sc, sh);
  if (res_kind == 2) bn_coeffs4(res_bn, c, C, rsc, rsh);
  const int ld = HW + 1;
  for (int pix = tid >> 3; pix < HW; pix += blockDim.x >> 3) {
    const int64_t off = ((int64_t)n * HW + pix) * C + c;
    float4 v = fma4(*reinterpret_cast<const float4*>(main.raw + off), sc, sh);
    if (res_kind == 1) {
      const bf16x4 h = *reinterpret_cast<const bf16x4*>(res_hi + off);
      bf16x4 l;
      if (res_lo) l = *reinterpret_cast<const bf16x4*>(res_lo + off);
      else l.v[0] = l.v[1] = l.v[2] = l.v[3] = __float2bfloat16_rn(0.f);
      v = add4(v, join4(h, l));
    } else if (res_kind == 2) {
      v = add4(v, fma4(*reinterpret_cast<const float4*>(res_bn.raw + off), rsc, rsh));
    }
    v = relu4(v);
    tile[(cg * 4 + 0) * ld + pix] = v.x;
    tile[(cg * 4 + 1) * ld + pix] = v.y;
    tile[(cg * 4 + 2) * ld + pix] = v.z;
    tile[(cg * 4 + 3) * ld + pix] = v.w;
  }
  __syncthreads();
  const int64_t out_n = scatter_idx ? scatter_idx[n] : n;
  if (spatial) {
    float* dst = spatial + ((int64_t)out_n * C + c_base) * HW;      // FP_CH*HW contiguous floats
    for (int i = tid; i < FP_CH * HW; i += blockDim.x) dst[i] = tile[(i / HW) * ld + (i % HW)];
  }
  // mean over HW: one warp per 4 channels
  const int warp = tid >> 5, lane = tid & 31;
  for (int ch = warp; ch < FP_CH; ch += blockDim.x >> 5) {
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) s += tile[ch * ld + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) pooled[(int64_t)out_n * C + c_base + ch] = s / (float)HW;
  }
}

int bn_final_pool_launch(const BnSide& main, int res_kind, const __nv_bfloat16* res_hi, const __nv_bfloat16* res_lo,
                         const BnSide& res_bn, const int64_t* scatter_idx, float* spatial_nchw, float* pooled, int N,
                         int HW, int C, cudaStream_t stream) {
  VB_REQUIRE(C % FP_CH == 0, "bn_final_pool: C=%d must be a multiple of %d", C, FP_CH);
  const size_t smem = (size_t)FP_CH * (HW + 1) * sizeof(float);
  VB_REQUIRE(smem <= 200 * 1024, "bn_final_pool: spatial size %d too large", HW);
  if (N == 0) return VB_OK;
  if (smem > 48 * 1024)
    VB_CHECK_CUDA(cudaFuncSetAttribute(bn_final_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(C / FP_CH, N);
  bn_final_pool_kernel<<<grid, 256, smem, stream>>>(to_dev(main), res_kind, res_hi, res_lo, to_dev(res_bn), scatter_idx,
                                                    spatial_nchw, pooled, N, HW, C);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 h, l;
    split_bf16(x[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}
int split_bf16_launch(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, cudaStream_t stream) {
  if (n == 0) return VB_OK;
  int blocks = div_up(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_bf16_kernel<<<blocks, 256, 0, stream>>>(x, hi, lo, n);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// F.normalize(x, dim=1): x / max(||x||_2, eps); one warp per row
__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int D, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * D;
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) ss = fmaf(xr[i], xr[i], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(sqrtf(ss), eps);
  for (int i = lane; i < D; i += 32) out[(int64_t)row * D + i] = xr[i] / denom;
}
int l2_normalize_launch(const float* x, float* out, int rows, int D, float eps, cudaStream_t stream) {
  if (rows == 0) return VB_OK;
  l2_normalize_kernel<<<div_up(rows, 8), 256, 0, stream>>>(x, out, rows, D, eps);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// jigsaw: [N,C,H,W] (rows gathered by idx) -> [9N, C, H3, W3], patch order row-major (vince_model.py:144-155)
__global__ void jigsaw_patchify_kernel(const float* __restrict__ x, const int64_t* __restrict__ gather_idx,
                                       float* __restrict__ out, int N, int C, int H, int W, int H3, int W3,
                                       int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int xw = (int)(t % W3); t /= W3;
    const int yh = (int)(t % H3); t /= H3;
    const int c = (int)(t % C); t /= C;
    const int patch = (int)(t % 9);
    const int n = (int)(t / 9);
    const int row = (patch / 3) * H3 + yh;
    const int col = (patch % 3) * W3 + xw;
    const int64_t sn = gather_idx ? gather_idx[n] : n;
    float v = 0.f;
    if (row < H && col < W) v = x[((sn * C + c) * H + row) * W + col];
    out[i] = v;
  }
}
int jigsaw_patchify_launch(const float* x, const int64_t* gather_idx, float* out, int N, int C, int H, int W, int H3,
                           int W3, cudaStream_t stream) {
  const int64_t total = (int64_t)N * 9 * C * H3 * W3;
  if (total == 0) return VB_OK;
  int blocks = div_up(total, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  jigsaw_patchify_kernel<<<blocks, 256, 0, stream>>>(x, gather_idx, out, N, C, H, W, H3, W3, total);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

__global__ void jigsaw_gather_kernel(const float* __restrict__ in, const int64_t* __restrict__ order,
                                     float* __restrict__ out, int N, int C, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t t = i / C;
    const int j = (int)(t % 9);
    const int n = (int)(t / 9);
    out[i] = in[((int64_t)n * 9 + order[n * 9 + j]) * C + c];
  }
}
int jigsaw_gather_launch(const float* in, const int64_t* order, float* out, int N, int C, cudaStream_t stream) {
  const int64_t total = (int64_t)N * 9 * C;
  if (total == 0) return VB_OK;
  int blocks = div_up(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  jigsaw_gather_kernel<<<blocks, 256, 0, stream>>>(in, order, out, N, C, total);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// fused multi-tensor momentum EMA + ring-buffer enqueue (vince_model.py:587-592, storage_queue.py:31-49)
// blocks [0, n_chunks): theta_k <- m*theta_k + (1-m)*theta_q for one <=8192-element chunk of one tensor
// blocks [n_chunks, ...): copy the new keys into the queue slice(s)
// ------------------------------------------------------------------------------------------------
__global__ void ema_enqueue_kernel(const EmaChunk* __restrict__ table, int n_chunks, float momentum, float one_minus,
                                   float* __restrict__ queue, float* __restrict__ queue_tf32,
                                   const float* __restrict__ keys, int64_t n0, int64_t dst0, int64_t n1, int64_t dst1,
                                   int64_t src1) {
  const int b = blockIdx.x;
  if (b < n_chunks) {
    const EmaChunk ch = table[b];
    const bool vec = ((reinterpret_cast<uintptr_t>(ch.dst) | reinterpret_cast<uintptr_t>(ch.src)) & 15) == 0;
    if (vec) {
      const int64_t n4 = ch.count >> 2;
      float4* d4 = reinterpret_cast<float4*>(ch.dst);
      const float4* s4 = reinterpret_cast<const float4*>(ch.src);
      for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 k = d4[i];
        const float4 q = __ldcs(s4 + i);
        // same operation order as queue_param.mul_(m).add_(1-m, param): round(m*k) then + (1-m)*q
        k.x = __fmaf_rn(one_minus, q.x, __fmul_rn(k.x, momentum));
        k.y = __fmaf_rn(one_minus, q.y, __fmul_rn(k.y, momentum));
        k.z = __fmaf_rn(one_minus, q.z, __fmul_rn(k.z, momentum));
        k.w = __fmaf_rn(one_minus, q.w, __fmul_rn(k.w, momentum));
        d4[i] = k;
      }
      for (int64_t i = (n4 << 2) + threadIdx.x; i < ch.count; i += blockDim.x)
        ch.dst[i] = __fmaf_rn(one_minus, ch.src[i], __fmul_rn(ch.dst[i], momentum));
    } else {
      for (int64_t i = threadIdx.x; i < ch.count; i += blockDim.x)
        ch.dst[i] = __fmaf_rn(one_minus, ch.src[i], __fmul_rn(ch.dst[i], momentum));
    }
  } else {
    const int64_t eb = b - n_chunks;
    const int64_t nblk = gridDim.x - n_chunks;
    const int64_t total = n0 + n1;                 // all multiples of 4 (D % 4 == 0, checked on the host)
    for (int64_t i = (eb * blockDim.x + threadIdx.x) * 4; i < total; i += nblk * blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(keys + (i < n0 ? i : src1 + (i - n0)));
      const int64_t doff = i < n0 ? dst0 + i : dst1 + (i - n0);
      *reinterpret_cast<float4*>(queue + doff) = v;
      if (queue_tf32) {
        uint32_t a, b2, c2, d2;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(v.x));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b2) : "f"(v.y));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c2) : "f"(v.z));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d2) : "f"(v.w));
        *reinterpret_cast<float4*>(queue_tf32 + doff) =
            make_float4(__uint_as_float(a), __uint_as_float(b2), __uint_as_float(c2), __uint_as_float(d2));
      }
    }
  }
}

int ema_enqueue_launch(const EmaChunk* table_dev, int n_chunks, float momentum, float one_minus, float* queue,
                       float* queue_tf32, const float* keys,
                       int64_t n0_elems, int64_t dst0_off, int64_t n1_elems, int64_t dst1_off, int64_t src1_off,
                       cudaStream_t stream) {
  VB_REQUIRE(n_chunks >= 0, "ema_enqueue: negative chunk count");
  VB_REQUIRE(((n0_elems | n1_elems | dst0_off | dst1_off | src1_off) & 3) == 0,
             "ema_enqueue: feature size must be a multiple of 4");
  const int64_t total = n0_elems + n1_elems;
  int enq_blocks = total > 0 ? div_up(total, 256 * 4) : 0;
  if (enq_blocks > 148) enq_blocks = 148;
  const int blocks = n_chunks + enq_blocks;
  if (blocks == 0) return VB_OK;
  ema_enqueue_kernel<<<blocks, 256, 0, stream>>>(table_dev, n_chunks, momentum, one_minus, queue, queue_tf32, keys, n0_elems,
                                                 dst0_off,
                                                 n1_elems, dst1_off, src1_off);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb
