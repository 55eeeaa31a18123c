// HBM-bound kernels of the encoder + queue path: layout packing, weight preparation, train/eval BatchNorm apply
// (+ReLU, +residual, +max/avg pooling), fp16 hi/lo splitting, L2 normalisation, jigsaw gathers, and the fused
// multi-tensor momentum-EMA + ring-buffer enqueue.  All are coalesced, vectorised (16-byte) streaming kernels.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vb {

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

struct alignas(8) h16x4 {
  __half v[4];
};
struct alignas(16) h16x8 {
  __half v[8];
};

__device__ __forceinline__ void split4(const float4& f, h16x4& hi, h16x4& lo) {
  split_f16(f.x, hi.v[0], lo.v[0]);
  split_f16(f.y, hi.v[1], lo.v[1]);
  split_f16(f.z, hi.v[2], lo.v[2]);
  split_f16(f.w, hi.v[3], lo.v[3]);
}
__device__ __forceinline__ float4 join4(const h16x4& hi, const h16x4& lo) {
  return make_float4(__half2float(hi.v[0]) + __half2float(lo.v[0]),
                     __half2float(hi.v[1]) + __half2float(lo.v[1]),
                     __half2float(hi.v[2]) + __half2float(lo.v[2]),
                     __half2float(hi.v[3]) + __half2float(lo.v[3]));
}

// ------------------------------------------------------------------------------------------------
// stem packing (overlapping-window layout): X[n, a, b, 16] holds the 2x2x3 input patch of row pair a-2 and column
// pair b-2:  X[n,a,b,(dr*2+dc)*3+c] = x[idx[n], c, 2(a-2)+dr, 2(b-2)+dc]  (0 outside the image, 0 for e >= 12).
// The 64 contiguous elements starting at (a, b = q) are then exactly the 2 x 8 x 3 input window (rows 2a-4, 2a-3,
// columns 2q-4 .. 2q+3) that output column q needs from that row pair, so the conv kernel's TMA descriptor reads
// 64-"channel" pixels at a pixel stride of 16 elements: the packed tensor is 1.4x the image instead of 5.4x.
// One thread per packed pixel (n, a, b): six 8-byte loads (2 rows x 3 channels x 2 adjacent columns; consecutive
// threads read consecutive column pairs, so every image element is fetched exactly once, coalesced) and four
// 16-byte stores.  Odd image widths (jigsaw patches, 225^2) take the scalar-load path (rows are then not 8-byte aligned).
// ------------------------------------------------------------------------------------------------
// Jigsaw (vince_model.py:144-155): with `grid` = 3 the N source images are read as 9N patches of PH x PW (patch order
// row-major, image n' = 9 n + py*3 + px starts at source pixel (py*PH, px*PW)); source pixels beyond H x W are the
// reference's zero padding to a multiple of 3.  Patchify is thereby folded into the stem's loads (the 9N patch tensor
// is never written).  grid = 1: PH = H, PW = W, plain frames.
struct StemSrc {
  int H, W;              // source image size
  int PH, PW;            // logical (patch) image size
  int grid;              // 1 or 3
};
template <bool EVEN_W>
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x,
                                                        const int64_t* __restrict__ gather_idx,
                                                        __half* __restrict__ hi, __half* __restrict__ lo,
                                                        StemSrc g, int Ha, int Wb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ha * Wb) return;
  const int n = blockIdx.y;
  const int a = idx / Wb, b = idx - a * Wb;
  const int g2 = g.grid * g.grid;
  const int img = n / g2, patch = n - img * g2;
  const int roff = (patch / g.grid) * g.PH, coff = (patch % g.grid) * g.PW;
  const int64_t src_n = gather_idx ? gather_idx[img] : img;
  const float* xn = x + src_n * 3 * (int64_t)g.H * g.W;
  const int row0 = 2 * (a - 2), col0 = 2 * (b - 2);
  float v[12];                                       // element (dr*2 + dc)*3 + c
#pragma unroll
  for (int dr = 0; dr < 2; ++dr) {
    const int row = row0 + dr;
    const bool row_ok = row >= 0 && row < g.PH && row + roff < g.H;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v0 = 0.f, v1 = 0.f;
      if (row_ok) {
        const float* rp = xn + ((int64_t)c * g.H + row + roff) * g.W + coff;
        if (EVEN_W) {
          if (col0 >= 0 && col0 < g.PW) {            // W even, col0 even: both columns are inside or both outside
            const float2 t = __ldg(reinterpret_cast<const float2*>(rp + col0));
            v0 = t.x, v1 = t.y;
          }
        } else {
          if (col0 >= 0 && col0 < g.PW && col0 + coff < g.W) v0 = __ldg(rp + col0);
          if (col0 + 1 >= 0 && col0 + 1 < g.PW && col0 + 1 + coff < g.W) v1 = __ldg(rp + col0 + 1);
        }
      }
      v[(dr * 2 + 0) * 3 + c] = v0;
      v[(dr * 2 + 1) * 3 + c] = v1;
    }
  }
  h16x8 oh[2], ol[2];
#pragma unroll
  for (int e = 0; e < 16; ++e) split_f16(e < 12 ? v[e] : 0.f, oh[e >> 3].v[e & 7], ol[e >> 3].v[e & 7]);
  const int64_t base = (((int64_t)n * Ha + a) * Wb + b) * 16;
  h16x8* dh = reinterpret_cast<h16x8*>(hi + base);
  dh[0] = oh[0], dh[1] = oh[1];
  if (lo) {
    h16x8* dl = reinterpret_cast<h16x8*>(lo + base);
    dl[0] = ol[0], dl[1] = ol[1];
  }
}

// uint8 HWC frames (what the dataset workers hold before ToTensor + Normalize, utils/transforms.py:89-101): the
// normalisation ((x / 255) - mean[c]) / std[c] - the same fp32 operations in the same order - is fused into the
// packing, so a step moves 1 byte per input element over PCIe and HBM instead of 4.
struct StemNorm {
  float mean[3], std[3];
};
__global__ void __launch_bounds__(256) stem_pack_u8_kernel(const uint8_t* __restrict__ x,
                                                           const int64_t* __restrict__ gather_idx, StemNorm nrm,
                                                           __half* __restrict__ hi, __half* __restrict__ lo,
                                                           StemSrc g, int Ha, int Wb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ha * Wb) return;
  const int n = blockIdx.y;
  const int a = idx / Wb, b = idx - a * Wb;
  const int g2 = g.grid * g.grid;
  const int img = n / g2, patch = n - img * g2;
  const int roff = (patch / g.grid) * g.PH, coff = (patch % g.grid) * g.PW;
  const int64_t src_n = gather_idx ? gather_idx[img] : img;
  const uint8_t* xn = x + src_n * 3 * (int64_t)g.H * g.W;
  const int row0 = 2 * (a - 2), col0 = 2 * (b - 2);
  float v[12];                                       // element (dr*2 + dc)*3 + c; 0 outside the image (zero padding
#pragma unroll                                       // applies to the NORMALISED image, as in the reference)
  for (int dr = 0; dr < 2; ++dr) {
    const int row = row0 + dr;
#pragma unroll
    for (int dc = 0; dc < 2; ++dc) {
      const int col = col0 + dc;
      const bool ok = row >= 0 && row < g.PH && col >= 0 && col < g.PW && row + roff < g.H && col + coff < g.W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float val = 0.f;
        if (ok) {
          const float px = (float)__ldg(xn + ((int64_t)(row + roff) * g.W + col + coff) * 3 + c);
          val = __fdiv_rn(__fsub_rn(__fdiv_rn(px, 255.f), nrm.mean[c]), nrm.std[c]);
        }
        v[(dr * 2 + dc) * 3 + c] = val;
      }
    }
  }
  h16x8 oh[2], ol[2];
#pragma unroll
  for (int e = 0; e < 16; ++e) split_f16(e < 12 ? v[e] : 0.f, oh[e >> 3].v[e & 7], ol[e >> 3].v[e & 7]);
  const int64_t base = (((int64_t)n * Ha + a) * Wb + b) * 16;
  h16x8* dh = reinterpret_cast<h16x8*>(hi + base);
  dh[0] = oh[0], dh[1] = oh[1];
  if (lo) {
    h16x8* dl = reinterpret_cast<h16x8*>(lo + base);
    dl[0] = ol[0], dl[1] = ol[1];
  }
}

static int stem_src(StemSrc& g, int H, int W, int grid, const char* what) {
  VB_REQUIRE(grid == 1 || grid == 3, "%s: grid must be 1 (frames) or 3 (jigsaw patches), got %d", what, grid);
  g.H = H, g.W = W, g.grid = grid, g.PH = H, g.PW = W;
  if (grid == 3) {
    // vince_model.py:145-146: pad BOTH axes by 3 - dim % 3 when EITHER is not a multiple of 3
    const bool pad = (H % 3) != 0 || (W % 3) != 0;
    g.PH = (pad ? H + 3 - H % 3 : H) / 3;
    g.PW = (pad ? W + 3 - W % 3 : W) / 3;
  }
  return VB_OK;
}

int stem_pack_u8_launch(const uint8_t* x, const int64_t* gather_idx, const float* mean3, const float* std3, __half* hi,
                        __half* lo, int N, int H, int W, int grid_p, cudaStream_t stream) {
  if (N == 0) return VB_OK;
  StemSrc g;
  int rc = stem_src(g, H, W, grid_p, "stem_pack_u8");
  if (rc) return rc;
  const int NP = N * grid_p * grid_p;
  VB_REQUIRE(NP <= 65535, "stem_pack_u8: batch %d too large", NP);
  const int Ha = (g.PH - 1) / 2 + 4, Wb = (g.PW - 1) / 2 + 4;
  StemNorm nrm;
  for (int c = 0; c < 3; ++c) {
    VB_REQUIRE(std3[c] != 0.f, "stem_pack_u8: std[%d] is zero", c);
    nrm.mean[c] = mean3[c], nrm.std[c] = std3[c];
  }
  dim3 grid((Ha * Wb + 255) / 256, NP);
  stem_pack_u8_kernel<<<grid, 256, 0, stream>>>(x, gather_idx, nrm, hi, lo, g, Ha, Wb);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

int stem_pack_launch(const float* x, const int64_t* gather_idx, __half* hi, __half* lo, int N, int H,
                     int W, int grid_p, cudaStream_t stream) {
  if (N == 0) return VB_OK;
  StemSrc g;
  int rc = stem_src(g, H, W, grid_p, "stem_pack");
  if (rc) return rc;
  const int NP = N * grid_p * grid_p;
  VB_REQUIRE(NP <= 65535, "stem_pack: batch %d too large", NP);
  const int Ha = (g.PH - 1) / 2 + 4, Wb = (g.PW - 1) / 2 + 4;
  dim3 grid((Ha * Wb + 255) / 256, NP);
  const bool even = grid_p == 1 && (W % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
  if (even) stem_pack_kernel<true><<<grid, 256, 0, stream>>>(x, gather_idx, hi, lo, g, Ha, Wb);
  else stem_pack_kernel<false><<<grid, 256, 0, stream>>>(x, gather_idx, hi, lo, g, Ha, Wb);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// weight preparation: OIHW fp32 -> K-major [Cout][R][S][Cin] fp16 hi/lo (or the packed stem layout)
// ------------------------------------------------------------------------------------------------
// One block per (output channel, tensor): the channel's [Cin][R*S] filter is staged in shared memory with
// coalesced loads and written back K-major ([R*S][Cin]) with coalesced 2-byte stores; 1x1 / Linear weights need no
// transpose and are streamed directly.
constexpr int WP_SMEM_FLOATS = 9216;               // 36 KB: up to 1024 input channels of a 3x3 filter
__device__ __forceinline__ float weight_prep_stem(const float* w147, int k) {
  // stem: k = t*64 + j*16 + (dr*2+dc)*3 + c  <-  w[c, r = 2t+dr-1, s = 2j+dc-1]   (7x7, Cin = 3):
  // tap t is the row pair p-2+t, j the column pair q-2+j of output pixel (p, q) (see stem_pack_kernel)
  const int t = k >> 6, j = (k >> 4) & 3, el = k & 15;
  if (el >= 12) return 0.f;
  const int dr = el / 6, dc = (el % 6) / 3, c = el % 3;
  const int r = 2 * t + dr - 1, sx = 2 * j + dc - 1;
  return (r >= 0 && r < 7 && sx >= 0 && sx < 7) ? w147[(c * 7 + r) * 7 + sx] : 0.f;
}

__global__ void __launch_bounds__(256) weight_prep_kernel(const WeightPrepEntry* __restrict__ table,
                                                          __half* __restrict__ hi, __half* __restrict__ lo) {
  extern __shared__ float wbuf[];
  const WeightPrepEntry e = table[blockIdx.y];
  const int RS = e.R * e.S;
  if (e.kind == 2) {
    // data-gradient weights: row ci of [Cin][R][S][Cout], k = (r', s', co) <- W[co][ci][R-1-r'][S-1-s'] (the flipped,
    // channel-transposed filter: dX = conv_stride1(dilate(dY), W'))
    const int ci = blockIdx.x;
    if (ci >= e.Cin) return;
    const int K = RS * e.Cout;
    const int64_t dst = e.dst_off + (int64_t)ci * K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const int rs = k / e.Cout, co = k - rs * e.Cout;
      const float w = __ldg(e.src + ((int64_t)co * e.Cin + ci) * RS + (RS - 1 - rs));
      __half h, l;
      split_f16(ldexpf(w, e.scale_log2), h, l);
      hi[dst + k] = h;
      if (lo) lo[dst + k] = l;
    }
    return;
  }
  const int co = blockIdx.x;
  if (co >= e.Cout) return;
  const int K = e.kind == 1 ? 256 : RS * e.Cin;
  const int n_src = e.kind == 1 ? 147 : RS * e.Cin;
  const float* src = e.src + (int64_t)co * n_src;
  const int64_t dst = e.dst_off + (int64_t)co * K;
  const bool staged = (e.kind == 1 || RS > 1) && n_src <= WP_SMEM_FLOATS;
  if (staged) {
    for (int i = threadIdx.x; i < n_src; i += blockDim.x) wbuf[i] = __ldg(src + i);
    __syncthreads();
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float w;
    if (e.kind == 1) {
      w = weight_prep_stem(wbuf, k);
    } else if (RS == 1) {
      w = __ldg(src + k);
    } else {
      const int rs = k / e.Cin, ci = k - rs * e.Cin;
      w = staged ? wbuf[ci * RS + rs] : __ldg(src + ci * RS + rs);
    }
    __half h, l;
    split_f16(ldexpf(w, e.scale_log2), h, l);
    hi[dst + k] = h;
    if (lo) lo[dst + k] = l;
  }
}

int weight_prep_launch(const WeightPrepEntry* table_dev, int n_entries, int max_cout, __half* hi,
                       __half* lo, cudaStream_t stream) {
  if (n_entries == 0 || max_cout == 0) return VB_OK;
  VB_REQUIRE(max_cout > 0 && n_entries <= 65535, "weight_prep: bad table size");
  dim3 grid(max_cout, n_entries);
  weight_prep_kernel<<<grid, 256, WP_SMEM_FLOATS * sizeof(float), stream>>>(table_dev, hi, lo);
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
// bn_apply: out = relu?( bn(main) + residual )  ->  fp16 hi/lo planes (and/or fp32)
// A thread owns V (4 or 8) consecutive channels of a row.  A block works on contiguous tiles of
// (blockDim / (C/V)) * U rows: all U rows' loads are issued before any arithmetic (memory-level parallelism), and a
// tile is one contiguous span of every tensor (DRAM page locality).  Tiles are handed out grid-strided.
// ------------------------------------------------------------------------------------------------
template <int V>
struct VecF {
  float4 q[V / 4];
};
template <int V>
struct alignas(V * 2) VecH {
  __half v[V];
};

template <int V, int U, int RES>
__global__ void __launch_bounds__(256, 4) bn_apply_kernel(BnSideDev main, const __half* __restrict__ res_hi,
                                                       const __half* __restrict__ res_lo, BnSideDev res_bn, int relu,
                                                       __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                       float* __restrict__ out_f32, int64_t M, int C) {
  const int CV = C / V;                               // threads per row
  const int rows_per_pass = blockDim.x / CV;          // rows covered by the block at once (>= 1: checked on the host)
  const int cg = threadIdx.x % CV;
  const int lrow = threadIdx.x / CV;
  const int c = cg * V;
  float sc[V], sh[V], rsc[V], rsh[V];
#pragma unroll
  for (int i = 0; i < V; i += 4) {
    float4 a, b;
    bn_coeffs4(main, c + i, C, a, b);
    sc[i] = a.x, sc[i + 1] = a.y, sc[i + 2] = a.z, sc[i + 3] = a.w;
    sh[i] = b.x, sh[i + 1] = b.y, sh[i + 2] = b.z, sh[i + 3] = b.w;
    if (RES == 2) {
      bn_coeffs4(res_bn, c + i, C, a, b);
      rsc[i] = a.x, rsc[i + 1] = a.y, rsc[i + 2] = a.z, rsc[i + 3] = a.w;
      rsh[i] = b.x, rsh[i + 1] = b.y, rsh[i + 2] = b.z, rsh[i + 3] = b.w;
    }
  }
  const int64_t tile_rows = (int64_t)rows_per_pass * U;
  for (int64_t t0 = (int64_t)blockIdx.x * tile_rows; t0 < M; t0 += (int64_t)gridDim.x * tile_rows) {
    VecF<V> raw[U], rres[U];
    VecH<V> rh[U], rl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = t0 + (int64_t)u * rows_per_pass + lrow;
      if (rr < M) {
        const int64_t off = rr * C + c;
#pragma unroll
        for (int i = 0; i < V / 4; ++i) raw[u].q[i] = __ldcs(reinterpret_cast<const float4*>(main.raw + off) + i);
        if (RES == 1) {
          rh[u] = *reinterpret_cast<const VecH<V>*>(res_hi + off);
          if (res_lo) rl[u] = *reinterpret_cast<const VecH<V>*>(res_lo + off);
        } else if (RES == 2) {
#pragma unroll
          for (int i = 0; i < V / 4; ++i) rres[u].q[i] = __ldcs(reinterpret_cast<const float4*>(res_bn.raw + off) + i);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = t0 + (int64_t)u * rows_per_pass + lrow;
      if (rr < M) {
        const int64_t off = rr * C + c;
        float v[V];
        const float* rp = reinterpret_cast<const float*>(&raw[u]);
        const float* rq = reinterpret_cast<const float*>(&rres[u]);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          v[i] = fmaf(rp[i], sc[i], sh[i]);
          if (RES == 1) v[i] += __half2float(rh[u].v[i]) + (res_lo ? __half2float(rl[u].v[i]) : 0.f);
          if (RES == 2) v[i] += fmaf(rq[i], rsc[i], rsh[i]);
          if (relu) v[i] = fmaxf(v[i], 0.f);
        }
        if (out_hi) {
          VecH<V> h, l;
#pragma unroll
          for (int i = 0; i < V; ++i) split_f16(v[i], h.v[i], l.v[i]);
          *reinterpret_cast<VecH<V>*>(out_hi + off) = h;
          if (out_lo) *reinterpret_cast<VecH<V>*>(out_lo + off) = l;
        }
        if (out_f32) {
#pragma unroll
          for (int i = 0; i < V / 4; ++i)
            reinterpret_cast<float4*>(out_f32 + off)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
}

template <int V, int U>
static void bn_apply_dispatch(int res_kind, int blocks, int threads, cudaStream_t stream, BnSideDev main,
                              const __half* res_hi, const __half* res_lo, BnSideDev res_bn, int relu, __half* out_hi,
                              __half* out_lo, float* out_f32, int64_t M, int C) {
  if (res_kind == 0)
    bn_apply_kernel<V, U, 0><<<blocks, threads, 0, stream>>>(main, res_hi, res_lo, res_bn, relu, out_hi, out_lo, out_f32, M, C);
  else if (res_kind == 1)
    bn_apply_kernel<V, U, 1><<<blocks, threads, 0, stream>>>(main, res_hi, res_lo, res_bn, relu, out_hi, out_lo, out_f32, M, C);
  else
    bn_apply_kernel<V, U, 2><<<blocks, threads, 0, stream>>>(main, res_hi, res_lo, res_bn, relu, out_hi, out_lo, out_f32, M, C);
}

int bn_apply_launch(const BnSide& main, int res_kind, const __half* res_hi, const __half* res_lo,
                    const BnSide& res_bn, int relu, __half* out_hi, __half* out_lo, float* out_f32,
                    int64_t M, int C, cudaStream_t stream) {
  VB_REQUIRE(C % 4 == 0, "bn_apply: C=%d must be a multiple of 4", C);
  if (M == 0) return VB_OK;
  // 8 channels per thread when the width allows it (16-byte fp16 stores), one row in flight per thread: measured on
  // B200 as fast as deeper unrolling (5.1-5.9 TB/s at the layer1 shape) at 52-64 registers, which lets two blocks
  // share an SM with a resident convolution CTA of the other encoder's stream
  const int V = (C % 8 == 0 && C / 8 <= 256) ? 8 : 4;
  const int threads = 256;
  const int CV = C / V;
  VB_REQUIRE(CV <= threads && threads % CV == 0, "bn_apply: unsupported channel count %d", C);
  const int64_t tile_rows = threads / CV;
  int64_t blocks = (M + tile_rows - 1) / tile_rows;
  const int64_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  const BnSideDev m = to_dev(main), r = to_dev(res_bn);
  if (V == 8)
    bn_apply_dispatch<8, 1>(res_kind, (int)blocks, threads, stream, m, res_hi, res_lo, r, relu, out_hi, out_lo, out_f32, M, C);
  else
    bn_apply_dispatch<4, 1>(res_kind, (int)blocks, threads, stream, m, res_hi, res_lo, r, relu, out_hi, out_lo, out_f32, M, C);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// stem: bn + relu + maxpool 3x3 stride 2 pad 1 (NHWC); padding behaves as -inf (torch max_pool2d).
// A thread owns (4 channels, one output column) and walks down a segment of output rows keeping the horizontal
// max of the last input row in registers, so every input row is fetched once per segment (+1 row of overlap) and
// the six 16-byte loads of an output row are issued back to back.  Neighbouring columns share one input column
// through L1.  Block = C/4 channel groups x MP_TQ output columns; grid = (column tiles, row segments, images).
// ------------------------------------------------------------------------------------------------
constexpr int MP_ROWS = 14;                           // output rows per segment
__device__ __forceinline__ float4 mp_hmax(const float* __restrict__ rowp, int xc, int Q, int C, const float4& sc,
                                          const float4& sh) {
  // max over input columns xc-1, xc, xc+1 of relu(bn(.)); rowp points at (row, col 0, this thread's channels)
  const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  float4 l = ninf, m, r = ninf;
  const float4* c4 = reinterpret_cast<const float4*>(rowp + (int64_t)xc * C);
  m = __ldg(c4);
  const bool hl = xc - 1 >= 0, hr = xc + 1 < Q;
  if (hl) l = __ldg(reinterpret_cast<const float4*>(rowp + (int64_t)(xc - 1) * C));
  if (hr) r = __ldg(reinterpret_cast<const float4*>(rowp + (int64_t)(xc + 1) * C));
  m = relu4(fma4(m, sc, sh));
  if (hl) m = max4(m, relu4(fma4(l, sc, sh)));
  if (hr) m = max4(m, relu4(fma4(r, sc, sh)));
  return m;
}

__global__ void __launch_bounds__(256) bn_relu_maxpool_kernel(BnSideDev bn, __half* __restrict__ out_hi,
                                                              __half* __restrict__ out_lo, int N, int P, int Q,
                                                              int C, int P2, int Q2, int tq) {
  const int C4 = C >> 2;
  const int cg = threadIdx.x % C4;
  const int q2 = blockIdx.x * tq + threadIdx.x / C4;
  const int n = blockIdx.z;
  if (q2 >= Q2) return;
  const int p2_0 = blockIdx.y * MP_ROWS;
  const int p2_1 = min(p2_0 + MP_ROWS, P2);
  float4 sc, sh;
  bn_coeffs4(bn, cg * 4, C, sc, sh);
  const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const int xc = 2 * q2;
  const float* img = bn.raw + (int64_t)n * P * Q * C + cg * 4;
  float4 prev = ninf;                                 // horizontal max of input row 2*p2 - 1
  if (2 * p2_0 - 1 >= 0) prev = mp_hmax(img + (int64_t)(2 * p2_0 - 1) * Q * C, xc, Q, C, sc, sh);
  for (int p2 = p2_0; p2 < p2_1; ++p2) {
    const int y0 = 2 * p2, y1 = 2 * p2 + 1;
    const float4 h0 = mp_hmax(img + (int64_t)y0 * Q * C, xc, Q, C, sc, sh);      // y0 <= P-1 always
    float4 h1 = ninf;
    if (y1 < P) h1 = mp_hmax(img + (int64_t)y1 * Q * C, xc, Q, C, sc, sh);
    const float4 m = max4(max4(prev, h0), h1);
    prev = h1;
    h16x4 h, l;
    split4(m, h, l);
    const int64_t off = ((((int64_t)n * P2 + p2) * Q2) + q2) * C + cg * 4;
    *reinterpret_cast<h16x4*>(out_hi + off) = h;
    if (out_lo) *reinterpret_cast<h16x4*>(out_lo + off) = l;
  }
}

int bn_relu_maxpool_launch(const BnSide& bn, __half* out_hi, __half* out_lo, int N, int P, int Q, int C,
                           int P2, int Q2, cudaStream_t stream) {
  VB_REQUIRE(C % 4 == 0 && C <= 1024, "bn_relu_maxpool: unsupported channel count %d", C);
  VB_REQUIRE(N <= 65535 && P2 <= 65535 * MP_ROWS, "bn_relu_maxpool: shape too large");
  if ((int64_t)N * P2 * Q2 == 0) return VB_OK;
  const int C4 = C / 4;
  // output columns per block: the divisor-friendly choice with the fewest idle threads
  int max_tq = 256 / C4;
  if (max_tq < 1) max_tq = 1;
  int tq = max_tq, best_waste = 1 << 30;
  for (int t = max_tq; t >= (max_tq + 1) / 2; --t) {
    const int tiles = (Q2 + t - 1) / t;
    const int waste = tiles * t - Q2;
    if (waste < best_waste) best_waste = waste, tq = t;
  }
  dim3 grid((Q2 + tq - 1) / tq, (P2 + MP_ROWS - 1) / MP_ROWS, N);
  bn_relu_maxpool_kernel<<<grid, tq * C4, 0, stream>>>(to_dev(bn), out_hi, out_lo, N, P, Q, C, P2, Q2, tq);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// last residual block: relu(bn(main)+residual) -> NCHW fp32 spatial features + global average pool
// block = (image n, 32-channel slab); smem transpose so both the NHWC read and the NCHW write are coalesced
// ------------------------------------------------------------------------------------------------
constexpr int FP_CH = 32;
__global__ void bn_final_pool_kernel(BnSideDev main, int res_kind, const __half* __restrict__ res_hi,
                                     const __half* __restrict__ res_lo, BnSideDev res_bn,
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
      const h16x4 h = *reinterpret_cast<const h16x4*>(res_hi + off);
      h16x4 l;
      if (res_lo) l = *reinterpret_cast<const h16x4*>(res_lo + off);
      else l.v[0] = l.v[1] = l.v[2] = l.v[3] = __float2half_rn(0.f);
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

int bn_final_pool_launch(const BnSide& main, int res_kind, const __half* res_hi, const __half* res_lo,
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
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                  __half* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    __half h, l;
    split_f16(x[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}
int split_f16_launch(const float* x, __half* hi, __half* lo, int64_t n, cudaStream_t stream) {
  if (n == 0) return VB_OK;
  int blocks = div_up(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_f16_kernel<<<blocks, 256, 0, stream>>>(x, hi, lo, n);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// Debug aid: how many values of an fp16 plane sit at the saturation bound +-65504 (the conversions of common.cuh clamp
// instead of overflowing to inf, so a network whose activations outgrow fp16's range would otherwise clip silently)
__global__ void count_saturated_kernel(const __half* __restrict__ x, int64_t n, unsigned long long* __restrict__ count) {
  unsigned int local = 0;
  const unsigned short* u = reinterpret_cast<const unsigned short*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    local += ((u[i] & 0x7fffu) >= 0x7bffu) ? 1u : 0u;        // |x| == 65504 (or inf / nan, which must never appear)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local != 0) atomicAdd(count, (unsigned long long)local);
}
int count_saturated_launch(const __half* x, int64_t n, unsigned long long* count, cudaStream_t stream) {
  if (n == 0) return VB_OK;
  int blocks = div_up(n, 256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_saturated_kernel<<<blocks, 256, 0, stream>>>(x, n, count);
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
