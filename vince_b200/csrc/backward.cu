// Query-encoder backward (what autograd does for vince_solver.py:463-469): the HBM-bound kernels
// around the tensor-core GEMMs.  The contractions themselves - data gradients (dgrad) and weight gradients (wgrad) -
// run on conv_gemm.cu's tcgen05 kernel: dgrad is a stride-1 convolution of the (zero-dilated) output gradient with the
// flipped / transposed filter (weight_prep kind 2), wgrad is a batched split-K GEMM over the pixel axis between the
// TRANSPOSED output gradient and the transposed, zero-padded input activation, one batch per filter tap.
//
//   bn_bwd_reduce    per-channel sum(dZ), sum(dZ * xhat) and max|dZ| of a conv+BN(+residual)(+ReLU) unit
//   bn_bwd_finalize  -> d gamma, d beta, the per-channel means the apply pass needs, and a power-of-two scale that puts
//                    the unit's dRaw into fp16's normal range (undone exactly by the consumers)
//   bn_bwd_apply     dRaw = gamma*invstd * (dZ - mean(dZ) - xhat * mean(dZ*xhat)) as fp16 (hi, lo) planes (optionally
//                    zero-dilated for stride-2 convolutions) and/or fp32; optionally the masked dZ for the skip path
//   transpose_pad    [pixels, C] planes -> [C][padded pixel axis] planes (operands of the wgrad GEMM)
//   wgrad_reduce     split-K partials -> OIHW fp32 gradient (accumulating)
//   maxpool_bwd, stem_wgrad, sgemm (projection head), normalize_bwd, colsum, sgd_step
#include "common.cuh"
#include "kernels.h"

namespace vb {

static inline int div_up64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// dZ of one element: (dA + dB) * [bcast scale] masked by the unit's ReLU
// ------------------------------------------------------------------------------------------------
struct BnBwdArgs {
  const float* dA;          // [M, C] fp32 gradient wrt the unit's output (or [M / bcast_hw, C] when bcast_hw > 0)
  const float* dB;          // optional second addend [M, C]
  int bcast_hw;             // > 0: dA is per image ([N, C]); row m reads dA[m / bcast_hw] * (1 / bcast_hw) (avg-pool backward)
  int mask_kind;            // 0 none, 1 relu(raw*sc + sh) > 0, 2 saved output planes > 0
  const __half* out_hi;     // mask_kind 2
  const __half* out_lo;
  const float* raw;         // [M, C] raw conv output (true scale)
  const float* coef;        // [4][C]: scale, shift, mean, invstd
  int64_t M;
  int C;
};

__device__ __forceinline__ float4 bwd_dz4(const BnBwdArgs& a, int64_t m, int c, const float4& raw, const float4& sc,
                                          const float4& sh) {
  float4 d;
  if (a.bcast_hw > 0) {
    const float inv = 1.f / (float)a.bcast_hw;
    const float4 t = __ldg(reinterpret_cast<const float4*>(a.dA + (m / a.bcast_hw) * a.C + c));
    d = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
  } else {
    d = __ldg(reinterpret_cast<const float4*>(a.dA + m * a.C + c));
  }
  if (a.dB != nullptr) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(a.dB + m * a.C + c));
    d.x += t.x, d.y += t.y, d.z += t.z, d.w += t.w;
  }
  if (a.mask_kind == 1) {
    if (!(fmaf(raw.x, sc.x, sh.x) > 0.f)) d.x = 0.f;
    if (!(fmaf(raw.y, sc.y, sh.y) > 0.f)) d.y = 0.f;
    if (!(fmaf(raw.z, sc.z, sh.z) > 0.f)) d.z = 0.f;
    if (!(fmaf(raw.w, sc.w, sh.w) > 0.f)) d.w = 0.f;
  } else if (a.mask_kind == 2) {
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(a.out_hi + m * a.C + c));
    uint2 l = make_uint2(0u, 0u);
    if (a.out_lo != nullptr) l = __ldg(reinterpret_cast<const uint2*>(a.out_lo + m * a.C + c));
    const __half2 h0 = *reinterpret_cast<const __half2*>(&h.x), h1 = *reinterpret_cast<const __half2*>(&h.y);
    const __half2 l0 = *reinterpret_cast<const __half2*>(&l.x), l1 = *reinterpret_cast<const __half2*>(&l.y);
    if (!(__low2float(h0) + __low2float(l0) > 0.f)) d.x = 0.f;
    if (!(__high2float(h0) + __high2float(l0) > 0.f)) d.y = 0.f;
    if (!(__low2float(h1) + __low2float(l1) > 0.f)) d.z = 0.f;
    if (!(__high2float(h1) + __high2float(l1) > 0.f)) d.w = 0.f;
  }
  return d;
}

// block = CG channel groups (4 channels each) x RP rows; grid.y tiles the channel groups, grid.x strides over row tiles
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(BnBwdArgs a, double* __restrict__ sums /*[2][C]*/,
                                                            unsigned int* __restrict__ maxbits, int cgpb) {
  __shared__ float red[2][256][4];
  const int rp = blockDim.x / cgpb;                  // rows per pass
  const int cgi = threadIdx.x % cgpb, lrow = threadIdx.x / cgpb;
  const int c = (blockIdx.y * cgpb + cgi) * 4;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.coef + c));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(a.coef + a.C + c));
  const float4 mu = __ldg(reinterpret_cast<const float4*>(a.coef + 2 * a.C + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(a.coef + 3 * a.C + c));
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float mx = 0.f;
  // two rows per trip: both rows' loads (raw, dA, dB / mask planes) are issued before either is consumed - twice the bytes
  // in flight per thread (this pass ran at half the HBM rate of the apply pass) - and accumulated in row order, so the sums
  // are bit-identical to the one-row loop
  const int64_t stride = (int64_t)gridDim.x * rp;
  auto acc = [&](const float4& raw, const float4& d) {
    s1[0] += d.x, s1[1] += d.y, s1[2] += d.z, s1[3] += d.w;
    s2[0] = fmaf(d.x, (raw.x - mu.x) * is.x, s2[0]);
    s2[1] = fmaf(d.y, (raw.y - mu.y) * is.y, s2[1]);
    s2[2] = fmaf(d.z, (raw.z - mu.z) * is.z, s2[2]);
    s2[3] = fmaf(d.w, (raw.w - mu.w) * is.w, s2[3]);
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
  };
  int64_t m = (int64_t)blockIdx.x * rp + lrow;
  for (; m + stride < a.M; m += 2 * stride) {
    const float4 raw0 = __ldg(reinterpret_cast<const float4*>(a.raw + m * a.C + c));
    const float4 raw1 = __ldg(reinterpret_cast<const float4*>(a.raw + (m + stride) * a.C + c));
    const float4 d0 = bwd_dz4(a, m, c, raw0, sc, sh);
    const float4 d1 = bwd_dz4(a, m + stride, c, raw1, sc, sh);
    acc(raw0, d0);
    acc(raw1, d1);
  }
  if (m < a.M) {
    const float4 raw = __ldg(reinterpret_cast<const float4*>(a.raw + m * a.C + c));
    acc(raw, bwd_dz4(a, m, c, raw, sc, sh));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[0][threadIdx.x][i] = s1[i], red[1][threadIdx.x][i] = s2[i];
  __syncthreads();
  if (lrow == 0) {
    double t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
    for (int r = 0; r < rp; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) t1[i] += (double)red[0][r * cgpb + cgi][i], t2[i] += (double)red[1][r * cgpb + cgi][i];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&sums[c + i], t1[i]);
      atomicAdd(&sums[a.C + c + i], t2[i]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(maxbits, __float_as_uint(mx));      // non-negative floats order as uints
}

// one block: d gamma / d beta (+=), per-channel means for the apply pass, and the power-of-two plane scale
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, const unsigned int* __restrict__ maxbits,
                                       const float* __restrict__ coef, double count, float* __restrict__ k12 /*[2][C]*/,
                                       float* __restrict__ scale2 /*[2]: 2^e, 2^-e*/, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate, int C) {
  __shared__ float smax[256];
  float mxsc = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = sums[c], s2 = sums[C + c];
    k12[c] = (float)(s1 / count);
    k12[C + c] = (float)(s2 / count);
    if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s2;
    if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s1;
    mxsc = fmaxf(mxsc, fabsf(coef[c]));
  }
  smax[threadIdx.x] = mxsc;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // |dRaw| <~ max|scale| * 3 * max|dZ| (the two mean terms are bounded by max|dZ| each for |xhat| ~ 1): put that
    // bound at 2^12 so that the bulk of the values has a NORMAL fp16 lo plane and nothing saturates
    const float bound = 3.f * smax[0] * __uint_as_float(*maxbits);
    int e = 0;
    if (bound > 0.f && isfinite(bound)) {
      int eb;
      frexpf(bound, &eb);                          // bound = f * 2^eb, f in [0.5, 1)
      e = 12 - eb;
      if (e > 100) e = 100;
      if (e < -100) e = -100;
    }
    scale2[0] = ldexpf(1.f, e);
    scale2[1] = ldexpf(1.f, -e);
  }
}

// dRaw -> planes (row m, or the zero-dilated position of a stride-`dil` convolution) and/or fp32; optional masked dZ out
struct BnBwdOut {
  __half* d_hi;             // optional [Mout, C] planes of 2^e * dRaw
  __half* d_lo;
  float* d_f32;             // optional [M, C] fp32 dRaw (unscaled)
  float* dz_out;            // optional [M, C] masked dZ (skip path)
  int dil;                  // 1, or 2: row (n, p, q) is written at (n, dil*p, dil*q) of an [N, Hd, Wd, C] tensor
  int P, Q, Hd, Wd;         // geometry for dil > 1
};
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(BnBwdArgs a, const float* __restrict__ k12,
                                                           const float* __restrict__ scale2, BnBwdOut o, int cgpb) {
  const int rp = blockDim.x / cgpb;
  const int cgi = threadIdx.x % cgpb, lrow = threadIdx.x / cgpb;
  const int c = (blockIdx.y * cgpb + cgi) * 4;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.coef + c));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(a.coef + a.C + c));
  const float4 mu = __ldg(reinterpret_cast<const float4*>(a.coef + 2 * a.C + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(a.coef + 3 * a.C + c));
  const float4 k1 = __ldg(reinterpret_cast<const float4*>(k12 + c));
  const float4 k2 = __ldg(reinterpret_cast<const float4*>(k12 + a.C + c));
  const float s = __ldg(scale2);
  for (int64_t m = (int64_t)blockIdx.x * rp + lrow; m < a.M; m += (int64_t)gridDim.x * rp) {
    const float4 raw = __ldg(reinterpret_cast<const float4*>(a.raw + m * a.C + c));
    const float4 d = bwd_dz4(a, m, c, raw, sc, sh);
    float4 g;
    g.x = sc.x * (d.x - k1.x - (raw.x - mu.x) * is.x * k2.x);
    g.y = sc.y * (d.y - k1.y - (raw.y - mu.y) * is.y * k2.y);
    g.z = sc.z * (d.z - k1.z - (raw.z - mu.z) * is.z * k2.z);
    g.w = sc.w * (d.w - k1.w - (raw.w - mu.w) * is.w * k2.w);
    if (o.dz_out) *reinterpret_cast<float4*>(o.dz_out + m * a.C + c) = d;
    if (o.d_f32) *reinterpret_cast<float4*>(o.d_f32 + m * a.C + c) = g;
    if (o.d_hi) {
      int64_t mo = m;
      if (o.dil > 1) {
        const int64_t pq = (int64_t)o.P * o.Q;
        const int64_t n = m / pq;
        const int rem = (int)(m - n * pq);
        const int p = rem / o.Q, q = rem - p * o.Q;
        mo = (n * o.Hd + (int64_t)p * o.dil) * o.Wd + (int64_t)q * o.dil;
      }
      __half h[4], l[4];
      split_f16(g.x * s, h[0], l[0]);
      split_f16(g.y * s, h[1], l[1]);
      split_f16(g.z * s, h[2], l[2]);
      split_f16(g.w * s, h[3], l[3]);
      const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
      const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
      uint2 hv, lv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h01), hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      lv.x = *reinterpret_cast<const uint32_t*>(&l01), lv.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(o.d_hi + mo * a.C + c) = hv;
      if (o.d_lo) *reinterpret_cast<uint2*>(o.d_lo + mo * a.C + c) = lv;
    }
  }
}

static int bwd_grid(int64_t M, int C, int& cgpb, dim3& grid) {
  VB_REQUIRE(C % 4 == 0, "bn_bwd: C=%d must be a multiple of 4", C);
  const int cg = C / 4;
  cgpb = cg < 256 ? cg : 256;
  VB_REQUIRE(256 % cgpb == 0 && cg % cgpb == 0, "bn_bwd: unsupported channel count %d", C);
  const int rp = 256 / cgpb;
  int64_t bx = (M + rp - 1) / rp;
  const int64_t cap = 148 * 8 / (cg / cgpb);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  grid = dim3((unsigned)bx, (unsigned)(cg / cgpb));
  return VB_OK;
}

static BnBwdArgs to_args(const BnBwdDesc& d) {
  BnBwdArgs a;
  a.dA = d.dA, a.dB = d.dB, a.bcast_hw = d.bcast_hw, a.mask_kind = d.mask_kind;
  a.out_hi = reinterpret_cast<const __half*>(d.out_hi), a.out_lo = reinterpret_cast<const __half*>(d.out_lo);
  a.raw = d.raw, a.coef = d.coef, a.M = d.M, a.C = d.C;
  return a;
}

int bn_bwd_launch(const BnBwdDesc& d, cudaStream_t stream) {
  VB_REQUIRE(d.dA && d.raw && d.coef && d.work && d.M > 0, "bn_bwd: null argument");
  VB_REQUIRE(d.mask_kind >= 0 && d.mask_kind <= 2 && (d.mask_kind != 2 || d.out_hi), "bn_bwd: bad mask");
  VB_REQUIRE(d.dil == 1 || d.dil == 2, "bn_bwd: dil must be 1 or 2");
  int cgpb;
  dim3 grid;
  int rc = bwd_grid(d.M, d.C, cgpb, grid);
  if (rc) return rc;
  // work buffer (doubles): [2C] sums | [C] = 2C floats k12 | [1] = 2 floats scale | [1] max bits
  double* sums = d.work;
  float* k12 = reinterpret_cast<float*>(d.work + 2 * d.C);
  float* scale2 = reinterpret_cast<float*>(d.work + 3 * d.C);
  unsigned int* maxbits = reinterpret_cast<unsigned int*>(d.work + 3 * d.C + 1);
  VB_CHECK_CUDA(cudaMemsetAsync(d.work, 0, (size_t)(3 * d.C + 2) * sizeof(double), stream));
  const BnBwdArgs a = to_args(d);
  bn_bwd_reduce_kernel<<<grid, 256, 0, stream>>>(a, sums, maxbits, cgpb);
  bn_bwd_finalize_kernel<<<1, 256, 0, stream>>>(sums, maxbits, d.coef, (double)d.M, k12, scale2, d.dgamma, d.dbeta,
                                                d.accumulate, d.C);
  BnBwdOut o;
  o.d_hi = reinterpret_cast<__half*>(d.d_hi), o.d_lo = reinterpret_cast<__half*>(d.d_lo), o.d_f32 = d.d_f32;
  o.dz_out = d.dz_out, o.dil = d.dil, o.P = d.P, o.Q = d.Q, o.Hd = d.Hd, o.Wd = d.Wd;
  bn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(a, k12, scale2, o, cgpb);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// planes [N*P*Q, C] -> transposed, zero-padded planes [C][ld]: element (n, p, q, c) lands at column
// n*L + (p*st + off)*Wp + (q*st + off), L = Hp*Wp.  dst must be zeroed by the caller when padding / dilation leaves holes.
// 32 x 32 shared-memory tile: coalesced on both sides.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_pad_kernel(const __half* __restrict__ s_hi, const __half* __restrict__ s_lo,
                                                            __half* __restrict__ d_hi, __half* __restrict__ d_lo, int64_t M,
                                                            int C, int P, int Q, int st, int off, int Hp, int Wp,
                                                            int64_t ld, int copies) {
  // tile = 64 pixels x 64 channels: 16-byte loads along the channels, a shared-memory transpose, then stores with the
  // lanes along the pixel axis (32 consecutive pixels of an image row are 64 contiguous bytes of a destination row)
  __shared__ __half th[64][66], tl[64][66];
  __shared__ int64_t colv[64];
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  if (t < 64) {
    const int64_t m = m0 + t;
    int64_t col = -1;
    if (m < M) {
      const int64_t pq = (int64_t)P * Q;
      const int64_t n = m / pq;
      const int rem = (int)(m - n * pq);
      const int p = rem / Q, q = rem - p * Q;
      col = n * (int64_t)Hp * Wp + (int64_t)(p * st + off) * Wp + (q * st + off);
    }
    colv[t] = col;
  }
  {
    const int chunk = t & 7;                         // 8 channels = 16 bytes
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int px = (t >> 3) + 32 * j;
      const int64_t m = m0 + px;
      uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
      if (m < M && c0 + chunk * 8 < C) {
        vh = __ldg(reinterpret_cast<const uint4*>(s_hi + m * C + c0 + chunk * 8));
        if (s_lo) vl = __ldg(reinterpret_cast<const uint4*>(s_lo + m * C + c0 + chunk * 8));
      }
      const __half* ph = reinterpret_cast<const __half*>(&vh);
      const __half* pl = reinterpret_cast<const __half*>(&vl);
#pragma unroll
      for (int e = 0; e < 8; ++e) th[chunk * 8 + e][px] = ph[e], tl[chunk * 8 + e][px] = pl[e];
    }
  }
  __syncthreads();
  // copies == 3: rows [j*C, (j+1)*C) hold the tensor shifted by j - 1 columns (copy_j[k] = x[k + j - 1]), so that a
  // consumer can realise +-1 column shifts with 16-byte aligned TMA coordinates by picking a copy
  const int px = t & 63;
  const int64_t col = colv[px];
  if (col < 0) return;
  for (int cc = (t >> 6); cc < 64; cc += 4) {
    const int c = c0 + cc;
    if (c >= C) break;
    const __half h = th[cc][px], l = tl[cc][px];
    for (int j = 0; j < copies; ++j) {
      const int64_t k = copies == 3 ? col - (j - 1) : col;
      if (k < 0 || k >= ld) continue;
      d_hi[((int64_t)j * C + c) * ld + k] = h;
      if (d_lo) d_lo[((int64_t)j * C + c) * ld + k] = l;
    }
  }
}

int transpose_pad_launch(const __half* s_hi, const __half* s_lo, __half* d_hi, __half* d_lo, int64_t M, int C, int P,
                         int Q, int st, int off, int Hp, int Wp, int64_t ld, int copies, cudaStream_t stream) {
  VB_REQUIRE(s_hi && d_hi && M > 0 && C > 0, "transpose_pad: null argument");
  VB_REQUIRE(copies == 1 || copies == 3, "transpose_pad: copies must be 1 or 3");
  VB_REQUIRE(C % 8 == 0, "transpose_pad: C=%d must be a multiple of 8", C);
  dim3 grid((unsigned)div_up64(M, 64), (unsigned)((C + 63) / 64));
  transpose_pad_kernel<<<grid, 256, 0, stream>>>(s_hi, s_lo, d_hi, d_lo, M, C, P, Q, st, off, Hp, Wp, ld, copies);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// split-K partials [taps*splits][Mpad][Cin] -> grad OIHW [Cout][Cin][R][S] (+=), times scale (device scalar pointers
// multiplied together: the 2^-e of the gradient planes)
// ------------------------------------------------------------------------------------------------
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int taps, int splits, int Mpad, int Cout, int Cin,
                                    const float* __restrict__ scale_dev, float scale, float* __restrict__ grad,
                                    int accumulate) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  const float s = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates (co, tap, ci) with ci fastest: coalesced reads of the partials
    const int ci = (int)(i % Cin);
    const int64_t t = i / Cin;
    const int tap = (int)(t % taps);
    const int co = (int)(t / taps);
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += part[((int64_t)(tap * splits + sp) * Mpad + co) * Cin + ci];
    const int64_t g = ((int64_t)co * Cin + ci) * taps + tap;
    grad[g] = (accumulate ? grad[g] : 0.f) + acc * s;
  }
}
int wgrad_reduce_launch(const float* part, int taps, int splits, int Mpad, int Cout, int Cin, const float* scale_dev,
                        float scale, float* grad, int accumulate, cudaStream_t stream) {
  VB_REQUIRE(part && grad, "wgrad_reduce: null pointer");
  const int64_t total = (int64_t)Cout * Cin * taps;
  int blocks = div_up64(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(part, taps, splits, Mpad, Cout, Cin, scale_dev, scale, grad, accumulate);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// max-pool 3x3/2 pad 1 backward through relu(bn(raw)): every pooled gradient goes to the FIRST maximum of its window
// (row-major scan, strict >, as torch's max_pool2d).  dst [N,P,Q,C] fp32 must be zeroed; contributions are atomics
// (<= 4 per element).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dA, const float* __restrict__ dB,
                                                          const float* __restrict__ raw, const float* __restrict__ coef,
                                                          float* __restrict__ dst, int N, int P, int Q, int C, int P2,
                                                          int Q2) {
  const int64_t total = (int64_t)N * P2 * Q2 * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int q2 = (int)(t % Q2);
    t /= Q2;
    const int p2 = (int)(t % P2);
    const int n = (int)(t / P2);
    float g = dA[i];
    if (dB) g += dB[i];
    const float sc = __ldg(coef + c), sh = __ldg(coef + C + c);
    float best = -INFINITY;
    int by = -1, bx = -1;
    for (int dy = 0; dy < 3; ++dy) {
      const int y = 2 * p2 - 1 + dy;
      if (y < 0 || y >= P) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int x = 2 * q2 - 1 + dx;
        if (x < 0 || x >= Q) continue;
        const float v = fmaxf(fmaf(__ldg(raw + (((int64_t)n * P + y) * Q + x) * C + c), sc, sh), 0.f);
        if (v > best) best = v, by = y, bx = x;
      }
    }
    if (by >= 0 && g != 0.f) atomicAdd(dst + (((int64_t)n * P + by) * Q + bx) * C + c, g);
  }
}
int maxpool_bwd_launch(const float* dA, const float* dB, const float* raw, const float* coef, float* dst, int N, int P,
                       int Q, int C, cudaStream_t stream) {
  VB_REQUIRE(dA && raw && coef && dst, "maxpool_bwd: null pointer");
  const int P2 = (P - 1) / 2 + 1, Q2 = (Q - 1) / 2 + 1;
  VB_CHECK_CUDA(cudaMemsetAsync(dst, 0, (size_t)N * P * Q * C * sizeof(float), stream));
  const int64_t total = (int64_t)N * P2 * Q2 * C;
  int blocks = div_up64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  maxpool_bwd_kernel<<<blocks, 256, 0, stream>>>(dA, dB, raw, coef, dst, N, P, Q, C, P2, Q2);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// stem weight gradient: dW[co][c][r][s] = sum_{n,p,q} dRaw[n,p,q,co] * x[idx[n], c, 2p-3+r, 2q-3+s]  (7x7/2, pad 3,
// Cin = 3).  CUDA cores (K = 147 does not fill a tensor-core tile): a block walks output-pixel tiles of 8 x 8, stages the
// 21 x 21 x 3 input patch and the [64 px][64 co] gradient tile in shared memory; thread (co, tap group) accumulates in
// registers over all its tiles and flushes once with atomics.
// ------------------------------------------------------------------------------------------------
constexpr int SW_TP = 8;                             // output tile 8 x 8 pixels
constexpr int SW_IN = 2 * SW_TP + 5;                 // 21 input rows / columns
constexpr int SW_TAPS = 10;                          // taps per thread: 16 tap groups x 10 >= 147
// Register tiling: thread = (4 output channels) x (10 filter taps); per pixel one 16-byte shared-memory read of the
// four gradient values and one (warp-broadcast) read per tap of the input patch feed 40 FMAs - 0.28 shared-memory
// reads per FMA instead of the 1.0 of the first version, which was bound by the LDS pipe at 1/6 of the FMA rate.
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const uint8_t* __restrict__ x8,
                                                         const int64_t* __restrict__ gather_idx, float m0, float m1,
                                                         float m2, float s0, float s1, float s2,
                                                         const float* __restrict__ draw, float* __restrict__ grad,
                                                         int N, int H, int W, int P, int Q) {
  __shared__ float xin[3][SW_IN][SW_IN + 1];
  __shared__ __align__(16) float dr[SW_TP * SW_TP][68];
  const int cog = threadIdx.x & 15;                  // output channels 4*cog .. 4*cog+3
  const int tg = threadIdx.x >> 4;                   // taps tg, tg+16, ...
  float acc[SW_TAPS][4];
  int toff[SW_TAPS];                                 // offset of tap (c, r, s) inside xin, -1 if beyond the 147 taps
#pragma unroll
  for (int i = 0; i < SW_TAPS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int tap = tg + 16 * i;
    const int c = tap / 49, r = (tap % 49) / 7, sx = tap % 7;
    toff[i] = tap < 147 ? (c * SW_IN + r) * (SW_IN + 1) + sx : -1;
  }
  const float* xf = &xin[0][0][0];
  const int tiles_y = (P + SW_TP - 1) / SW_TP, tiles_x = (Q + SW_TP - 1) / SW_TP;
  const int64_t ntiles = (int64_t)N * tiles_y * tiles_x;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = (int)(tile / (tiles_y * tiles_x));
    const int tr = (int)(tile % (tiles_y * tiles_x));
    const int p0 = (tr / tiles_x) * SW_TP, q0 = (tr % tiles_x) * SW_TP;
    const int64_t src_n = gather_idx ? gather_idx[n] : n;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * SW_IN * SW_IN; i += 256) {
      const int c = i / (SW_IN * SW_IN), r2 = (i / SW_IN) % SW_IN, c2 = i % SW_IN;
      const int y = 2 * p0 - 3 + r2, xx = 2 * q0 - 3 + c2;
      float v = 0.f;
      if (y >= 0 && y < H && xx >= 0 && xx < W) {
        if (x8 != nullptr) {
          const float px = (float)__ldg(x8 + ((src_n * H + y) * (int64_t)W + xx) * 3 + c);
          v = __fdiv_rn(__fsub_rn(__fdiv_rn(px, 255.f), mean[c]), stdv[c]);
        } else {
          v = __ldg(x + ((src_n * 3 + c) * (int64_t)H + y) * W + xx);
        }
      }
      xin[c][r2][c2] = v;
    }
    for (int i = threadIdx.x; i < SW_TP * SW_TP * 16; i += 256) {
      const int px = i >> 4, c4 = i & 15;
      const int p = p0 + px / SW_TP, q = q0 + px % SW_TP;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < P && q < Q) v = __ldg(reinterpret_cast<const float4*>(draw + (((int64_t)n * P + p) * Q + q) * 64 + c4 * 4));
      *reinterpret_cast<float4*>(&dr[px][c4 * 4]) = v;
    }
    __syncthreads();
    for (int py = 0; py < SW_TP; ++py) {
#pragma unroll
      for (int pxx = 0; pxx < SW_TP; ++pxx) {
        const float4 dv = *reinterpret_cast<const float4*>(&dr[py * SW_TP + pxx][cog * 4]);
        const int base = (2 * py) * (SW_IN + 1) + 2 * pxx;
#pragma unroll
        for (int i = 0; i < SW_TAPS; ++i) {
          if (toff[i] >= 0) {
            const float xv = xf[toff[i] + base];
            acc[i][0] = fmaf(dv.x, xv, acc[i][0]);
            acc[i][1] = fmaf(dv.y, xv, acc[i][1]);
            acc[i][2] = fmaf(dv.z, xv, acc[i][2]);
            acc[i][3] = fmaf(dv.w, xv, acc[i][3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < SW_TAPS; ++i) {
    const int tap = tg + 16 * i;
    if (tap < 147) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(grad + (cog * 4 + j) * 147 + tap, acc[i][j]);
    }
  }
}
int stem_wgrad_launch(const float* x, const uint8_t* x8, const int64_t* gather_idx, const float* mean3,
                      const float* std3, const float* draw, float* grad, int N, int H, int W, int accumulate,
                      cudaStream_t stream) {
  VB_REQUIRE((x || x8) && draw && grad, "stem_wgrad: null pointer");
  const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
  if (!accumulate) VB_CHECK_CUDA(cudaMemsetAsync(grad, 0, 64 * 147 * sizeof(float), stream));
  float m[3] = {0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
  if (x8) {
    VB_REQUIRE(mean3 && std3, "stem_wgrad: uint8 input needs mean / std");
    for (int c = 0; c < 3; ++c) m[c] = mean3[c], s[c] = std3[c];
  }
  stem_wgrad_kernel<<<148 * 2, 256, 0, stream>>>(x, x8, gather_idx, m[0], m[1], m[2], s[0], s[1], s[2], draw, grad, N, H, W,
                                                 P, Q);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// projection head: plain fp32 GEMM on CUDA cores (B = 256 rows: 1 GMAC per layer, not worth a tensor-core path)
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N], row-major with leading dimensions; ta / tb = operand stored transposed
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ Cm, int M, int N, int K, int lda, int ldb,
                                                    int ldc, int ta, int tb, int accumulate,
                                                    const float* __restrict__ relu_mask_src) {
  __shared__ float As[16][64 + 1], Bs[16][64 + 1];
  const int bm = blockIdx.y * 64, bn = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i / 64, mm = i % 64;
      const int gm = bm + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? (ta ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk]) : 0.f;
      const int gn = bn + mm;
      Bs[kk][mm] = (gn < N && gk < K) ? (tb ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i], b[i] = Bs[kk][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = bm + ty * 4 + i, gn = bn + tx * 4 + j;
      if (gm < M && gn < N) {
        float v = acc[i][j] + (accumulate ? Cm[(int64_t)gm * ldc + gn] : 0.f);
        if (relu_mask_src && !(relu_mask_src[(int64_t)gm * ldc + gn] > 0.f)) v = 0.f;
        Cm[(int64_t)gm * ldc + gn] = v;
      }
    }
}
int sgemm_launch(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int ta, int tb,
                 int accumulate, const float* relu_mask_src, cudaStream_t stream) {
  VB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "sgemm: bad argument");
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  sgemm_kernel<<<grid, 256, 0, stream>>>(A, B, C, M, N, K, lda, ldb, ldc, ta, tb, accumulate, relu_mask_src);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// column sums: out[c] (+)= sum_r x[r, c]   (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int R, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += x[(int64_t)r * C + c];
  out[c] = (accumulate ? out[c] : 0.f) + s;
}
int colsum_launch(const float* x, float* out, int R, int C, int accumulate, cudaStream_t stream) {
  VB_REQUIRE(x && out, "colsum: null pointer");
  colsum_kernel<<<(C + 127) / 128, 128, 0, stream>>>(x, out, R, C, accumulate);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// F.normalize backward: y = x / max(||x||, eps);  dx = (dy - y * (y . dy)) / max(||x||, eps); one warp per row
__global__ void normalize_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                     int rows, int D, float eps, float gscale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * D;
  const float* gr = dy + (int64_t)row * D;
  float ss = 0.f, dot = 0.f;
  for (int i = lane; i < D; i += 32) ss = fmaf(xr[i], xr[i], ss), dot = fmaf(xr[i], gr[i], dot);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o), dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float nrm = sqrtf(ss);
  const float den = fmaxf(nrm, eps);
  // y = x / den; y . dy = dot / den.  For ||x|| < eps the clamp makes y linear in x: dx = dy / eps.
  for (int i = lane; i < D; i += 32) {
    const float g = gr[i] * gscale;
    float v;
    if (nrm >= eps) v = (g - xr[i] * (dot * gscale) / (den * den)) / den;
    else v = g / den;
    dx[(int64_t)row * D + i] = v;
  }
}
int normalize_bwd_launch(const float* x, const float* dy, float* dx, int rows, int D, float eps, float gscale,
                         cudaStream_t stream) {
  VB_REQUIRE(x && dy && dx, "normalize_bwd: null pointer");
  if (rows == 0) return VB_OK;
  normalize_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, dy, dx, rows, D, eps, gscale);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// fused multi-tensor SGD with momentum and weight decay (torch.optim.SGD semantics, vince_solver.py:252-256):
//   d = g * gscale + wd * p;  buf = first ? d : mu * buf + d;  p -= lr * buf
// ------------------------------------------------------------------------------------------------
__global__ void sgd_kernel(const SgdChunk* __restrict__ table, float lr, float mu, float wd, float gscale, int first) {
  const SgdChunk ch = table[blockIdx.x];
  for (int64_t i = threadIdx.x; i < ch.count; i += blockDim.x) {
    const float p = ch.param[i];
    const float d = fmaf(wd, p, ch.grad[i] * gscale);
    float b = d;
    if (ch.buf != nullptr) {
      b = first ? d : fmaf(mu, ch.buf[i], d);
      ch.buf[i] = b;
    }
    ch.param[i] = p - lr * b;
  }
}
int sgd_launch(const SgdChunk* table_dev, int n_chunks, float lr, float momentum, float weight_decay, float grad_scale,
               int first_step, cudaStream_t stream) {
  if (n_chunks == 0) return VB_OK;
  VB_REQUIRE(table_dev, "sgd: null table");
  sgd_kernel<<<n_chunks, 256, 0, stream>>>(table_dev, lr, momentum, weight_decay, grad_scale, first_step);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb
