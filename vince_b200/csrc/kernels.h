// Internal C++ launch API of libvince_b200 (the exported C ABI in include/vince_b200.h wraps these 1:1).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace vb {

// ---- conv_gemm.cu -------------------------------------------------------------------------------
struct ConvGemmDesc {
  const void* a_hi;      // fp16 plane(s) of A: [M,K] row-major, or NHWC [batch,H,W,Cin] when im2col
  const void* a_lo;
  const void* b_hi;      // fp16 plane(s) of the weights, [N, K] row-major, K ordered (r, s, cin)
  const void* b_lo;
  float* out;            // fp32 [M, N]
  int M, N, K;
  int im2col;
  int batch, H, W, Cin, R, S, stride, pad_lo_h, pad_lo_w, pad_hi_h, pad_hi_w;
  int passes;            // 3 = fp16x3 (fp32-grade), 1 = plain fp16
  int block_n;           // 0 = auto
  const float* scale;    // optional [N]
  const float* bias;     // optional [N]
  int relu;
  float alpha;           // out = epilogue(alpha * A*W^T); 0 means 1 (used to undo the power-of-two weight pre-scale)
  double* stats;         // optional [2][N] (+=): per-channel sum and sum of squares of the raw outputs
  int halo_mode;         // 3x3 stride-1 convs: -1 auto (use the shared-memory halo path when efficient), 0 off, 1 force
  // optional train-mode BatchNorm finalize fused into the kernel tail (needs stats): the last CTA writes
  // bn_coef[2][N] = (scale, shift) and updates the running statistics / num_batches_tracked
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_running_mean;
  float* bn_running_var;
  int64_t* bn_num_batches_tracked;
  float* bn_coef;
  unsigned int* bn_counter;   // must be 0 at launch
  float bn_momentum, bn_eps;
  long long a_pixel_stride, a_row_stride, a_img_stride;   // im2col: element strides (0 = dense NHWC)
  void* trace;                // debug builds (-DVB_TRACE) only: [4][512] uint64 timeline of CTA 0
  // ---- "apply" epilogue (out_hi != null, out == null): the output is written as the fp16 (hi, lo) planes of
  //      relu?( alpha*acc*ep_coef[c] + ep_coef[N+c] + residual ), i.e. BatchNorm with known coefficients (+ the residual
  //      add and ReLU of resnet.py:76-92,117-137) folded into the producing convolution
  void* out_hi;
  void* out_lo;
  const float* ep_coef;       // [2][N] (scale, shift)
  int res_kind;               // 0 none, 1 fp16 planes (res_hi, res_lo) [M,N], 2 bn(res_raw [M,N] fp32) with res_coef [2][N]
  const void* res_hi;
  const void* res_lo;
  const float* res_raw;
  const float* res_coef;
  // ---- statistics-only pass (needs stats): BatchNorm sums (+ finalize) are produced, nothing is stored
  int stats_only;
  int bn_save;                // fused finalize also stores mean / inv-std: bn_coef is then [4][N]
  const float* alpha_dev;     // optional device scalar multiplied into alpha
  // ---- batched split-K GEMM (kchunk > 0; weight gradients): out[(tap*splits + split)*Mpad + m, n] =
  //      sum_{k in split} A[m, k] * B[n, k + shift(tap)], K = full contraction extent, Mpad = M rounded up to 128,
  //      shift(tap) = (tap/3 - 1)*shift_w + tap%3 - 1 for taps == 9, else 0; out-of-range k reads as zero
  int kchunk;                 // K elements per split (multiple of 64)
  int taps;                   // 1 or 9
  int shift_w;
  int res_slots;              // internal (set by conv_gemm_launch): residual tiles of the apply epilogue's cp.async ring
};
int conv_gemm_launch(const ConvGemmDesc& d, cudaStream_t stream);
// eval-mode BatchNorm: coef[2][C] = (gamma / sqrt(running_var + eps), beta - running_mean * scale)
int bn_eval_coef_launch(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, float* coef,
                        int C, cudaStream_t stream);

// ---- elementwise.cu -----------------------------------------------------------------------------
// Stem input packing: NCHW fp32 image -> X[n, a, b, 16] fp16 hi/lo (overlapping-window layout) with
//   X[n,a,b,(dr*2+dc)*3 + c] = x[idx[n], c, 2(a-2)+dr, 2(b-2)+dc]   (0 outside the image, 0 for e >= 12)
// so that the 7x7/2 pad-3 stem conv becomes a 4x1 stride-1 im2col conv over row pairs with 64-element pixels read
// at a pixel stride of 16 elements.
// `grid` = 1: N frames of H x W.  `grid` = 3: the jigsaw branch - the N frames are read as 9N patches (patchify of
// vince_model.py:144-155 folded into the stem's loads, zero padding to a multiple of 3 included).
int stem_pack_launch(const float* x, const int64_t* gather_idx, __half* hi, __half* lo, int N, int H,
                     int W, int grid, cudaStream_t stream);

// same packing from uint8 HWC frames [N,H,W,3] with ((x/255) - mean[c]) / std[c] fused in (mean3 / std3: host arrays)
int stem_pack_u8_launch(const uint8_t* x, const int64_t* gather_idx, const float* mean3, const float* std3, __half* hi,
                        __half* lo, int N, int H, int W, int grid, cudaStream_t stream);

struct WeightPrepEntry {   // one per weight tensor; lives in device memory
  const float* src;        // OIHW fp32 (Linear: [Cout, Cin] with R=S=1)
  int64_t dst_off;         // element offset into the hi / lo planes
  int32_t Cout, Cin, R, S;
  int32_t kind;            // 0: [Cout][R][S][Cin]; 1: stem packing [64][4][4][16]
  int32_t scale_log2;      // weights are multiplied by 2^scale_log2 before the fp16 split (keeps lo planes normal)
};
int weight_prep_launch(const WeightPrepEntry* table_dev, int n_entries, int max_cout, __half* hi,
                       __half* lo, cudaStream_t stream);

struct BnSide {
  const float* raw;        // [M, C] raw conv output
  const float* coef;       // [2][C] per-channel (scale, shift) written by conv_gemm's fused finalize / bn_eval_coef
};
// out = relu?( bn(main) + residual ), residual = none | hi+lo planes | bn(second raw tensor)
int bn_apply_launch(const BnSide& main, int res_kind, const __half* res_hi, const __half* res_lo,
                    const BnSide& res_bn, int relu, __half* out_hi, __half* out_lo, float* out_f32,
                    int64_t M, int C, cudaStream_t stream);
// stem: bn + relu + 3x3/2 pad-1 max pool, NHWC
int bn_relu_maxpool_launch(const BnSide& bn, __half* out_hi, __half* out_lo, int N, int P, int Q, int C,
                           int P2, int Q2, cudaStream_t stream);
// last block: relu(bn(main)+residual) -> NCHW fp32 spatial features (rows scattered by scatter_idx) + global mean
int bn_final_pool_launch(const BnSide& main, int res_kind, const __half* res_hi, const __half* res_lo,
                         const BnSide& res_bn, const int64_t* scatter_idx, float* spatial_nchw, float* pooled, int N,
                         int HW, int C, cudaStream_t stream);
int split_f16_launch(const float* x, __half* hi, __half* lo, int64_t n, cudaStream_t stream);
// *count += number of plane values at the fp16 saturation bound (debug aid, see vince_count_saturated)
int count_saturated_launch(const __half* x, int64_t n, unsigned long long* count, cudaStream_t stream);
int l2_normalize_launch(const float* x, float* out, int rows, int D, float eps, cudaStream_t stream);
// NCHW fp32 [N,C,H,W] -> jigsaw patches NCHW [9N,C,H3,W3] (pad bottom/right with zeros to a multiple of 3)
int jigsaw_patchify_launch(const float* x, const int64_t* gather_idx, float* out, int N, int C, int H, int W, int H3,
                           int W3, cudaStream_t stream);
// out[n, j*C + c] = in[(n*9 + order[n, j]), c]
int jigsaw_gather_launch(const float* in, const int64_t* order, float* out, int N, int C, cudaStream_t stream);

struct EmaChunk {        // device-resident table, one entry per <=EMA_CHUNK contiguous elements
  float* dst;            // key-encoder parameter
  const float* src;      // query-encoder parameter
  int64_t count;
};
constexpr int EMA_CHUNK = 8192;
// theta_k <- m*theta_k + (1-m)*theta_q over the table, and (same launch) ring-buffer enqueue of up to two row slices
// `one_minus` = float(1 - momentum) evaluated in double by the caller, as the reference's Python does
int ema_enqueue_launch(const EmaChunk* table_dev, int n_chunks, float momentum, float one_minus, float* queue,
                       float* queue_tf32, const float* keys,
                       int64_t n0_elems, int64_t dst0_off, int64_t n1_elems, int64_t dst1_off, int64_t src1_off,
                       cudaStream_t stream);

// ---- infonce.cu ---------------------------------------------------------------------------------
struct InfoNceDesc {
  const float* q;          // [B, D] queries (exact fp32)
  const float* keys;       // [Bk, D] current keys (exact fp32); Bk == B
  const float* queue_tf32; // [K, D] queue, values pre-rounded (RN) to TF32; may be null if K == 0
  int B, Bk, K, D;
  int num_frames;          // > 0 (ibc): keys are the first Bk columns, positives of row i = {j : j/nf == i/nf};
                           // 0 (MoCo): keys are NOT columns, the single positive of row i is keys[i]
  float temperature;
  // outputs
  float* dists;            // [B, nP]   nP = num_frames (ibc) or 1 (MoCo)
  float* weights;          // [B, nP]
  float* pos_sim;          // [B, nP]   raw positive similarities
  float* neg_max;          // [B]       max raw similarity over the negatives
  float* row_lse;          // [B, 2]    (row max of z over all columns, Zneg relative to it) - saved for backward
  float* scalars;          // [8]       dist, softmax_weight, nce_accuracy, cosine_sim, cosine_sim_neg_max
  void* workspace;         // >= infonce_workspace_bytes(B, D), 256-byte aligned
};
size_t infonce_workspace_bytes(int B, int D);
int infonce_fwd_launch(const InfoNceDesc& d, cudaStream_t stream);
int round_tf32_launch(const float* x, float* out, int64_t n, cudaStream_t stream);
// backward w.r.t. q (infonce_bwd.cu): d.pos_sim / d.row_lse are the forward's outputs (inputs here); symmetric = the
// self-batch loss where the columns are the queries themselves; accumulate adds to dq instead of overwriting it
size_t infonce_bwd_workspace_bytes(int B, int D);
int infonce_bwd_launch(const InfoNceDesc& d, float grad_dist, int symmetric, int accumulate, float* dq,
                       cudaStream_t stream);
// explicit-matrix masked cross entropy (loss_util.similarity_cross_entropy's literal signature); error_flag is set
// to 1 when some row does not have exactly nP positives
int masked_ce_launch(const float* sims, const uint8_t* mask, int R, int C, int nP, float temperature, float* dists,
                     float* weights, float* pos_sim, float* neg_max, float* row_lse, float* scalars, int* error_flag,
                     cudaStream_t stream);

// ---- backward.cu (query-encoder backward) ------------------------------------------------
// One conv+BN(+residual)(+ReLU) unit: dZ = (dA + dB) masked by the unit's ReLU; per-channel sums -> d gamma / d beta;
// dRaw = gamma*invstd*(dZ - mean(dZ) - xhat*mean(dZ*xhat)) written as 2^e-scaled fp16 planes (optionally zero-dilated)
// and / or fp32.  `work` = (3C + 2) doubles of scratch; afterwards work + 3C holds the floats (2^e, 2^-e).
struct BnBwdDesc {
  const float* dA;
  const float* dB;
  int bcast_hw;
  int mask_kind;            // 0 none, 1 recompute relu(raw*scale + shift) > 0, 2 saved output planes > 0
  const void* out_hi;
  const void* out_lo;
  const float* raw;
  const float* coef;        // [4][C] scale, shift, mean, invstd (vince_conv_fwd with bn_save)
  int64_t M;
  int C;
  double* work;
  float* dgamma;            // optional [C]
  float* dbeta;
  int accumulate;           // add to dgamma / dbeta instead of overwriting
  void* d_hi;               // optional planes of 2^e * dRaw
  void* d_lo;
  float* d_f32;             // optional fp32 dRaw
  float* dz_out;            // optional masked dZ [M, C]
  int dil, P, Q, Hd, Wd;
};
int bn_bwd_launch(const BnBwdDesc& d, cudaStream_t stream);
int transpose_pad_launch(const __half* s_hi, const __half* s_lo, __half* d_hi, __half* d_lo, int64_t M, int C, int P,
                         int Q, int st, int off, int Hp, int Wp, int64_t ld, int copies, cudaStream_t stream);
int wgrad_reduce_launch(const float* part, int taps, int splits, int Mpad, int Cout, int Cin, const float* scale_dev,
                        float scale, float* grad, int accumulate, cudaStream_t stream);
int maxpool_bwd_launch(const float* dA, const float* dB, const float* raw, const float* coef, float* dst, int N, int P,
                       int Q, int C, cudaStream_t stream);
int stem_wgrad_launch(const float* x, const uint8_t* x8, const int64_t* gather_idx, const float* mean3,
                      const float* std3, const float* draw, float* grad, int N, int H, int W, int accumulate,
                      cudaStream_t stream);
int sgemm_launch(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int ta, int tb,
                 int accumulate, const float* relu_mask_src, cudaStream_t stream);
int colsum_launch(const float* x, float* out, int R, int C, int accumulate, cudaStream_t stream);
int normalize_bwd_launch(const float* x, const float* dy, float* dx, int rows, int D, float eps, float gscale,
                         cudaStream_t stream);
struct SgdChunk {
  float* param;
  const float* grad;
  float* buf;               // momentum buffer (null: no momentum)
  int64_t count;
};
int sgd_launch(const SgdChunk* table_dev, int n_chunks, float lr, float momentum, float weight_decay, float grad_scale,
               int first_step, cudaStream_t stream);

// ---- knn.cu ------------------------------------------------------------------------------------------------
// exact k-NN of every row of feats [n, D] among all rows (Euclidean), self match dropped; nbr_idx [n, k],
// nbr_dist [n, k] (optional), pred [n] = most frequent neighbour label (smallest on ties; optional with labels)
int knn_launch(const float* feats, const int64_t* labels, int n, int D, int k, int64_t* nbr_idx, float* nbr_dist,
               int64_t* pred, cudaStream_t stream);

}  // namespace vb
