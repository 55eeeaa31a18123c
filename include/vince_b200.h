/* libvince_b200 - C ABI of the B200-native VINCE hot path (encoder forward, fused InfoNCE, EMA + enqueue).
 *
 * The reference (danielgordon10/vince) is pure Python on PyTorch and has NO FFI of its own; every entry point
 * below replaces a PyTorch library call site on its hot path (cited per function, paths relative to the
 * reference root).  The Python classes in vince_b200/ (VinceModel, VinceQueueModel, StorageQueue,
 * loss_util.similarity_cross_entropy) bind these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success or a negative vince_status; it never throws and never syncs the device;
 *     vince_last_error() returns a thread-local message for the last failure on the calling thread;
 *   - all tensor arguments are raw DEVICE pointers owned by the caller, with explicit shapes; nothing is
 *     allocated inside (scratch comes in through `workspace` arguments sized by *_workspace_bytes());
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - activations between convolutions are NHWC, carried as a pair of fp16 planes (hi, lo) with
 *     hi = fp16(x), lo = fp16(x - hi), x ~= hi + lo (|err| <= 2^-24 |x| while lo is a normal fp16 number;
 *     conversions saturate at +-65504); with `passes` = 3 each MMA k-step computes hi*hi + lo*hi + hi*lo into an
 *     fp32 accumulator, so the result matches the reference's fp32 arithmetic to ~1e-6 per layer; `passes` = 1
 *     uses the hi plane only (plain fp16, lo pointers may be NULL).  Weights are pre-scaled by an exact power of
 *     two (vince_weight_entry.scale_log2) that vince_conv_desc.alpha undoes.
 */
#ifndef VINCE_B200_H_
#define VINCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  VINCE_OK = 0,
  VINCE_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  VINCE_ERR_CUDA = -2,        /* CUDA runtime or driver error */
  VINCE_ERR_NCCL = -3,        /* NCCL error, or libnccl could not be loaded */
  VINCE_ERR_UNSUPPORTED = -4
} vince_status;

const char* vince_last_error(void);
int vince_abi_version(void);

/* ---- convolution / linear layer as a tcgen05 implicit GEMM ------------------------------------------------
 * replaces: torchvision resnet conv2d calls reached from models/building_blocks/backbone_models.py:50-53
 *           (spec: models/building_blocks/resnet.py:76-92, 117-137, 231-247), the nn.Linear layers of
 *           models/vince_model.py:38-49,163,171,177, and the statistics half of train-mode nn.BatchNorm2d.
 * out[M,N] (fp32, row-major == NHWC) = epilogue(A * W^T); epilogue = optional per-channel scale, bias, ReLU.
 * `stats` (optional, [2][N] fp64, accumulated with +=) receives per-channel sum and sum of squares of the raw
 * outputs.  When `bn_coef` is also given the kernel finishes train-mode BatchNorm itself (resnet.py:79,83 in
 * training mode): the last CTA converts the sums into bn_coef[2][N] = (gamma/sqrt(var+eps), beta - mean*scale),
 * updates running_mean / running_var (momentum, unbiased variance) and increments num_batches_tracked.
 * 3x3 stride-1 pad-1 convolutions use a shared-memory halo path (one TMA load per tile serves all nine taps) when
 * halo_mode = -1 (auto) finds it efficient, or always when halo_mode = 1; 0 forces the TMA-im2col path. */
typedef struct {
  const void* a_hi;       /* fp16: [M,K] row-major, or NHWC [batch,H,W,Cin] when im2col != 0 */
  const void* a_lo;
  const void* w_hi;       /* fp16 [N,K], K ordered (r, s, cin): see vince_weight_prep */
  const void* w_lo;
  float* out;
  int32_t M, N, K;
  int32_t im2col;
  int32_t batch, H, W, Cin, R, S, stride, pad_lo_h, pad_lo_w, pad_hi_h, pad_hi_w;
  int32_t passes;         /* 3 or 1 */
  int32_t block_n;        /* 0 = auto; 64 / 128 / 256 */
  const float* scale;
  const float* bias;
  int32_t relu;
  int32_t halo_mode;      /* -1 auto, 0 off, 1 force */
  double* stats;
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_running_mean;
  float* bn_running_var;
  int64_t* bn_num_batches_tracked;   /* may be NULL */
  float* bn_coef;
  uint32_t* bn_counter;   /* one zero-initialised word per launch */
  float bn_momentum, bn_eps;
  /* im2col only: element strides of the activation tensor between pixels / rows / images; 0 = dense NHWC
   * (Cin, W*Cin, H*W*Cin).  A pixel stride smaller than Cin describes overlapping pixel windows (the packed
   * stem input of vince_stem_pack). */
  int64_t a_pixel_stride, a_row_stride, a_img_stride;
  float alpha;            /* accumulators are multiplied by alpha before the epilogue (0 = 1): undoes the exact
                           * power-of-two pre-scale of vince_weight_prep */
  int32_t stats_only;     /* 1: statistics pass - `stats` (+ the fused finalize) are produced, nothing is stored
                           * (`out` may be NULL).  First half of the train-mode two-pass scheme for wide 1x1 convolutions:
                           * pass 1 computes the batch statistics, pass 2 recomputes the GEMM with the apply epilogue,
                           * so the raw [M,N] fp32 tensor and the separate vince_bn_apply pass never touch HBM.
                           * 2 (ABI v4; im2col == 0 only): the same pass with the operand roles swapped inside the kernel
                           * (accumulator rows = output channels, columns = pixels), which turns the per-channel sums
                           * into in-register adds: 2.8x faster on the 56x56 64->256 expansion (74 vs 207 us) */
  /* "apply" epilogue (ABI v2), selected by out_hi != NULL (then out must be NULL and stats / scale / bias unused):
   *   planes(out) = relu?( alpha*acc*ep_coef[c] + ep_coef[N+c] + residual )
   * replaces: nn.BatchNorm2d (eval mode, or train mode with coefficients from a statistics pass) + the residual add
   * + ReLU of resnet.py:76-92,117-137 folded into the producing convolution; same fp32 operation order as
   * vince_bn_apply, so both routes give identical bits. */
  void* out_hi;           /* fp16 [M,N] */
  void* out_lo;
  const float* ep_coef;   /* [2][N] (scale, shift) */
  int32_t res_kind;       /* 0 none, 1 planes (res_hi,res_lo) [M,N], 2 bn(res_raw [M,N] fp32) with res_coef [2][N] */
  int32_t reserved;
  const void* res_hi;
  const void* res_lo;
  const float* res_raw;
  const float* res_coef;
  /* training support (ABI v3) */
  int32_t bn_save;        /* 1: the fused finalize also stores the batch mean and 1/sqrt(var+eps): bn_coef is [4][N]
                           * (scale, shift, mean, inv-std) - what the BatchNorm backward needs */
  int32_t kchunk;         /* > 0: batched split-K GEMM for weight gradients (replaces autograd's conv2d / linear weight
                           * gradient at vince_solver.py:465): out[(tap*splits+split)*Mpad + m, n] = sum over the split's
                           * k of A[m,k] * B[n, k + shift(tap)]; K = whole contraction extent (pixels), kchunk = k per
                           * split (multiple of 64), Mpad = M rounded up to 128, splits = ceil(K / kchunk) */
  int32_t taps;           /* 1, or 9 with shift(tap) = (tap/3 - 1)*shift_w + tap%3 - 1 (3x3 filter taps as shifts along
                           * a zero-padded pixel axis of row pitch shift_w) */
  int32_t shift_w;
  const float* alpha_dev; /* optional DEVICE scalar multiplied into alpha (dynamic power-of-two gradient scaling) */
} vince_conv_desc;
int vince_conv_fwd(const vince_conv_desc* desc, void* stream);
/* eval-mode BatchNorm coefficients from the running statistics: coef[2][C] */
int vince_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, float* coef, int32_t C, void* stream);

/* ---- stem input packing: NCHW fp32 -> X[n, a, b, 16] fp16 (hi, lo), overlapping-window layout ---------------
 * replaces: the input side of conv1 (7x7/2, pad 3; resnet.py:170,233) and the shuffle gather data[shuffle_order]
 *           (vince_model.py:142) when gather_idx != NULL.
 * X[n,a,b,(dr*2+dc)*3+c] = x[gather_idx[n], c, 2(a-2)+dr, 2(b-2)+dc] (0 outside the image and for elements 12..15);
 * Ha = P+3, Wb = Q+3 with P = (H-1)/2+1, Q = (W-1)/2+1.  The 64 elements starting at (a, b=q) are the 2x8x3 input
 * window of output column q in row pair a, so the stem runs as vince_conv_fwd with batch,H,W,Cin = N,Ha,Q,64,
 * R=4, S=1, stride 1, no padding, a_pixel_stride = 16, a_row_stride = 16*Wb, a_img_stride = 16*Wb*Ha and weights
 * prepared with kind = 1. */
int vince_stem_pack(const float* x, const int64_t* gather_idx, void* x_hi, void* x_lo, int32_t N, int32_t H, int32_t W,
                    void* stream);

/* Same packing straight from uint8 HWC frames [N,H,W,3] (what the dataset workers hold before
 * utils/transforms.py:89-101 applies ToTensor(scale=255) + Normalize): ((x / 255) - mean3[c]) / std3[c] in fp32, same
 * operation order, fused into the packing (1 byte per element over PCIe / HBM instead of 4).
 * mean3 / std3 are HOST arrays of 3 floats (constants.py:28-29 divided by 255 for the reference's pipelines). */
int vince_stem_pack_u8(const uint8_t* x_nhwc, const int64_t* gather_idx, const float* mean3, const float* std3,
                       void* x_hi, void* x_lo, int32_t N, int32_t H, int32_t W, void* stream);

/* Jigsaw variants (grid = 3): replace the patchify of vince_model.py:144-155 - pad bottom/right with zeros to a
 * multiple of 3 (BOTH axes by 3 - dim % 3 when either is not divisible, :145-146), cut 3x3 patches, row-major patch
 * order, [N,3,H,W] -> [9N,3,PH,PW] - by folding it into the stem's loads: packed image 9n + 3*py + px is patch
 * (py, px) of frame gather_idx[n]; the 9N patch tensor is never written.  x_hi / x_lo hold 9N packed images of the
 * PH x PW geometry.  grid = 1 is identical to the plain entry points. */
int vince_stem_pack_grid(const float* x, const int64_t* gather_idx, void* x_hi, void* x_lo, int32_t N, int32_t H,
                         int32_t W, int32_t grid, void* stream);
int vince_stem_pack_u8_grid(const uint8_t* x_nhwc, const int64_t* gather_idx, const float* mean3, const float* std3,
                            void* x_hi, void* x_lo, int32_t N, int32_t H, int32_t W, int32_t grid, void* stream);

/* ---- weight preparation (multi-tensor): OIHW fp32 -> K-major fp16 (hi, lo) ----------------------------------
 * replaces: nothing in the reference (cuDNN consumes OIHW directly); run after every weight update.
 * `table_dev` is a DEVICE array of n_entries descriptors; one block per (output channel, tensor), so `max_cout` is the
 * largest Cout in the table. */
typedef struct {
  const float* src;       /* [Cout,Cin,R,S] fp32 */
  int64_t dst_off;        /* element offset into w_hi / w_lo */
  int32_t Cout, Cin, R, S;
  int32_t kind;           /* 0: [Cout][R][S][Cin];  1: packed 7x7 stem -> [64][4 row pairs][4 col pairs][16] */
  int32_t scale_log2;     /* weights are multiplied by 2^scale_log2 before the split, so that the lo plane of O(1e-2)
                           * weights stays a normal fp16 number; pass alpha = 2^-scale_log2 to vince_conv_fwd */
} vince_weight_entry;
int vince_weight_prep(const vince_weight_entry* table_dev, int32_t n_entries, int64_t max_cout, void* w_hi, void* w_lo,
                      void* stream);

/* ---- BatchNorm apply (+residual, +ReLU) and pooling ----------------------------------------------------------
 * replaces: the normalisation half of nn.BatchNorm2d, the residual add and ReLU of resnet.py:76-92 / 117-137,
 *           MaxPool2d (resnet.py:173,236) and AdaptiveAvgPool2d (vince_model.py:33,128).
 * `coef` = [2][C] (scale, shift) from vince_conv_fwd's fused finalize (train) or vince_bn_eval_coef (eval). */
typedef struct {
  const float* raw;       /* [M,C] raw conv output (NHWC) */
  const float* coef;      /* [2][C] */
} vince_bn_side;
/* out = relu?( bn(main) + residual ); res_kind 0 none, 1 (res_hi,res_lo) planes, 2 bn(res_bn) (downsample branch) */
int vince_bn_apply(const vince_bn_side* main, int32_t res_kind, const void* res_hi, const void* res_lo,
                   const vince_bn_side* res_bn, int32_t relu, void* out_hi, void* out_lo, float* out_f32, int64_t M,
                   int32_t C, void* stream);
int vince_bn_relu_maxpool(const vince_bn_side* bn, void* out_hi, void* out_lo, int32_t N, int32_t P, int32_t Q, int32_t C,
                          void* stream);
/* last block: relu(bn(main)+residual) -> spatial_features NCHW fp32 [N,C,h,w] + extracted_features [N,C]; output
 * row n is written at scatter_idx[n] (the un-shuffle of vince_model.py:184-192) */
int vince_bn_final_pool(const vince_bn_side* main, int32_t res_kind, const void* res_hi, const void* res_lo,
                        const vince_bn_side* res_bn, const int64_t* scatter_idx, float* spatial_nchw, float* pooled,
                        int32_t N, int32_t HW, int32_t C, void* stream);

/* ---- small helpers --------------------------------------------------------------------------------------------
 * vince_l2_normalize replaces F.normalize(dim=1) (vince_model.py:180); the jigsaw pair replaces
 * vince_model.py:144-155 and :164-170. */
int vince_split_f16(const float* x, void* hi, void* lo, int64_t n, void* stream);
int vince_round_tf32(const float* x, float* out, int64_t n, void* stream);
/* debug aid (no reference counterpart): *count += how many of the n fp16 values of an activation / weight plane sit at
 * the saturation bound +-65504.  The fp32 -> fp16 (hi, lo) split clamps instead of overflowing, so a network whose
 * activations outgrow fp16's range (large BatchNorm gammas, unnormalised residual sums) clips silently unless checked:
 * EncoderRunner runs this over every plane it produces when VINCE_B200_CHECK_SATURATION=1. */
int vince_count_saturated(const void* plane_f16, int64_t n, uint64_t* count, void* stream);
int vince_l2_normalize(const float* x, float* out, int32_t rows, int32_t D, float eps, void* stream);
int vince_jigsaw_patchify(const float* x, const int64_t* gather_idx, float* out, int32_t N, int32_t C, int32_t H,
                          int32_t W, void* stream);
int vince_jigsaw_gather(const float* in, const int64_t* order, float* out, int32_t N, int32_t C, void* stream);

/* ---- fused InfoNCE forward ----------------------------------------------------------------------------------
 * replaces: VinceModel.forward similarity matmuls (vince_model.py:213-233), loss_util.similarity_cross_entropy
 *           (utils/loss_util.py:7-62) and the metric passes of VinceModel.get_metrics (vince_model.py:314-342).
 * ONE kernel launch (ABI v4): q / keys are exact fp32 (16-byte aligned, D a multiple of 32, <= 128) and are rounded to
 * TF32 inside the kernel; the last CTA of each 128-row query block merges the per-CTA (max, sum-exp) partials, adds the
 * exact-fp32 positives and writes the per-row outputs, the last CTA overall reduces the five scalars. */
typedef struct {
  const float* q;
  const float* keys;
  const float* queue_tf32;
  int32_t B, Bk, K, D;
  int32_t num_frames;     /* >0: inter-batch comparison, block-diagonal positives; 0: MoCo (l_pos = q_i . k_i) */
  float temperature;
  float* dists;
  float* weights;
  float* pos_sim;
  float* neg_max;
  float* row_lse;
  float* scalars;         /* [8]: 0 dist, 1 softmax_weight, 2 nce_accuracy, 3 cosine_sim, 4 cosine_sim_neg_max */
  void* workspace;
} vince_infonce_desc;
size_t vince_infonce_workspace_bytes(int32_t B, int32_t D);
/* ---- fused InfoNCE backward w.r.t. the queries ---------------------------------------------------------------
 * replaces: what autograd computes for `embeddings` through vince_model.py:213-233 + utils/loss_util.py:7-62 when
 *           vince_solver.py:465 calls loss.backward() (keys and queue are detached: vince_model.py:598,610,
 *           storage_queue.py:53).  dq[B,D] = d(grad_dist * dist)/dq.  `desc` is the forward's descriptor: q, keys,
 *           queue_tf32, shapes and temperature as in the forward, pos_sim and row_lse as the forward WROTE them
 *           (inputs here); dists / weights / neg_max / scalars are ignored; workspace >=
 *           vince_infonce_bwd_workspace_bytes.  symmetric = 1: the self-batch loss (vince_model.py:213-222; keys ==
 *           q, K == 0), whose columns carry gradient too.  accumulate = 1 adds to dq. */
size_t vince_infonce_bwd_workspace_bytes(int32_t B, int32_t D);
int vince_infonce_bwd(const vince_infonce_desc* desc, float grad_dist, int32_t symmetric, int32_t accumulate, float* dq,
                      void* stream);
int vince_infonce_fwd(const vince_infonce_desc* desc, void* stream);

/* explicit-matrix form with the literal signature of loss_util.similarity_cross_entropy (utils/loss_util.py:7-62,
 * equal-count branch) plus the metric quantities of vince_model.py:327-333: sims [rows, cols] fp32, mask [rows, cols]
 * bool (1 byte), exactly n_pos positives per row (else *error_flag is set to 1).  Outputs as vince_infonce_fwd. */
int vince_masked_ce_fwd(const float* sims, const uint8_t* mask, int32_t rows, int32_t cols, int32_t n_pos,
                        float temperature, float* dists, float* weights, float* pos_sim, float* neg_max, float* row_lse,
                        float* scalars, int32_t* error_flag, void* stream);

/* ---- fused momentum EMA (multi-tensor) + ring-buffer enqueue ----------------------------------------------------
 * replaces: VinceQueueModel.param_update (vince_model.py:587-592; ~2 launches per parameter tensor) and the
 *           copy_ calls of StorageQueue.enqueue (utils/storage_queue.py:38,46).
 * The table holds one entry per <= 8192 contiguous floats of one parameter tensor.  The enqueue part copies
 * keys[0:n0] to queue[dst0:...] and keys[src1:src1+n1] to queue[dst1:...] (element offsets; n1 > 0 on wrap-around);
 * when queue_tf32 != NULL the TF32-rounded shadow used by vince_infonce_fwd is written in the same pass. */
typedef struct {
  float* dst;
  const float* src;
  int64_t count;
} vince_ema_chunk;
int vince_ema_enqueue(const vince_ema_chunk* table_dev, int32_t n_chunks, float momentum, float one_minus_momentum,
                      float* queue, float* queue_tf32, const float* keys, int64_t n0, int64_t dst0, int64_t n1,
                      int64_t dst1, int64_t src1, void* stream);

/* ---- query-encoder backward + optimiser ---------------------------------------------------
 * replaces: what torch autograd + torch.optim.SGD do for the query encoder at solvers/vince_solver.py:252-256,463-469
 *           (loss.backward(); optimizer.step()).  The contractions run on vince_conv_fwd:
 *   data gradient    dX = conv_stride1(dilate(dRaw), W')   with W' from vince_weight_prep kind = 2 (flipped, channel-
 *                    transposed filter), alpha_dev = the 2^-e of the dRaw planes;
 *   weight gradient  batched split-K GEMM (kchunk / taps / shift_w) between vince_transpose_pad'ed dRaw and input planes,
 *                    reduced by vince_wgrad_reduce into the OIHW fp32 gradient.
 * vince_bn_bwd handles one conv + BatchNorm (+ residual) (+ ReLU) unit (train-mode batch statistics, resnet.py:76-92,
 * 117-137): dZ = (dA + dB) masked by the ReLU; d gamma = sum dZ*xhat, d beta = sum dZ;
 * dRaw = gamma*invstd*(dZ - mean dZ - xhat * mean(dZ*xhat)) written as fp16 planes scaled by a power of two 2^e chosen
 * per unit (kept at work + 3C as two floats 2^e, 2^-e), optionally zero-dilated for stride-2 convolutions. */
typedef struct {
  const float* dA;          /* [M,C] gradient wrt the unit's output ([M/bcast_hw, C] if bcast_hw > 0: avg-pool backward) */
  const float* dB;          /* optional second addend */
  int32_t bcast_hw;
  int32_t mask_kind;        /* 0 none, 1 relu(raw*scale+shift) > 0 recomputed, 2 saved output planes > 0 */
  const void* out_hi;
  const void* out_lo;
  const float* raw;         /* [M,C] raw conv output */
  const float* coef;        /* [4][C] scale, shift, mean, inv-std (vince_conv_fwd with bn_save = 1) */
  int64_t M;
  int32_t C;
  double* work;             /* (3C + 2) doubles */
  float* dgamma;            /* optional [C] */
  float* dbeta;
  int32_t accumulate;
  void* d_hi;               /* optional: planes of 2^e * dRaw, [M,C] or dilated [N,Hd,Wd,C] (pre-zeroed) */
  void* d_lo;
  float* d_f32;             /* optional: fp32 dRaw */
  float* dz_out;            /* optional: masked dZ (gradient of the skip path) */
  int32_t dil, P, Q, Hd, Wd;
} vince_bn_bwd_desc;
int vince_bn_bwd(const vince_bn_bwd_desc* desc, void* stream);
/* planes [N*P*Q, C] -> [copies*C][ld] planes, (n,p,q,c) at column n*Hp*Wp + (p*stride+offset)*Wp + q*stride+offset;
 * copies = 3 additionally stores the tensor shifted by -1 / 0 / +1 columns in rows [0,C) / [C,2C) / [2C,3C) (TMA needs
 * 16-byte aligned addresses: the 3x3 taps' column shifts are realised by picking a copy) */
int vince_transpose_pad(const void* src_hi, const void* src_lo, void* dst_hi, void* dst_lo, int64_t M, int32_t C, int32_t P,
                        int32_t Q, int32_t stride, int32_t offset, int32_t Hp, int32_t Wp, int64_t ld, int32_t copies,
                        void* stream);
/* partials [taps*splits][Mpad][Cin] -> grad [Cout][Cin][taps] (OIHW), times scale * (*scale_dev) */
int vince_wgrad_reduce(const float* partials, int32_t taps, int32_t splits, int32_t Mpad, int32_t Cout, int32_t Cin,
                       const float* scale_dev, float scale, float* grad, int32_t accumulate, void* stream);
/* max-pool 3x3/2 backward through relu(bn(raw)) of the stem (resnet.py:173,236): dst [N,P,Q,C] fp32 */
int vince_maxpool_bwd(const float* dA, const float* dB, const float* raw, const float* coef, float* dst, int32_t N,
                      int32_t P, int32_t Q, int32_t C, void* stream);
/* conv1 (7x7/2, Cin = 3) weight gradient from the fp32 dRaw [N,P,Q,64] and the input frames (fp32 NCHW or uint8 HWC) */
int vince_stem_wgrad(const float* x, const uint8_t* x_u8, const int64_t* gather_idx, const float* mean3,
                     const float* std3, const float* draw, float* grad, int32_t N, int32_t H, int32_t W,
                     int32_t accumulate, void* stream);
/* projection head: fp32 GEMM C (+)= op(A) op(B), optional ReLU mask (zero where relu_mask_src <= 0); column sums;
 * F.normalize backward (vince_model.py:177-180) */
int vince_sgemm(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t lda, int32_t ldb,
                int32_t ldc, int32_t trans_a, int32_t trans_b, int32_t accumulate, const float* relu_mask_src,
                void* stream);
int vince_colsum(const float* x, float* out, int32_t R, int32_t C, int32_t accumulate, void* stream);
int vince_normalize_bwd(const float* x, const float* dy, float* dx, int32_t rows, int32_t D, float eps, float gscale,
                        void* stream);
/* fused multi-tensor SGD, torch.optim.SGD semantics (vince_solver.py:252-256): d = g*grad_scale + wd*p;
 * buf = first_step ? d : momentum*buf + d; p -= lr*buf */
typedef struct {
  float* param;
  const float* grad;
  float* buf;
  int64_t count;
} vince_sgd_chunk;
int vince_sgd_step(const vince_sgd_chunk* table_dev, int32_t n_chunks, float lr, float momentum, float weight_decay,
                   float grad_scale, int32_t first_step, void* stream);

/* ---- eval: exact k-nearest-neighbour label voting over embeddings ------------------------------------------------
 * replaces: the kNN-CIFAR evaluation of solvers/vince_solver.py:651-693 (sklearn KDTree(all_features).query(k=11) on
 * the host, first (self) match dropped, scipy.stats.mode over the neighbour labels).  feats [n,D] fp32 (D % 4 == 0),
 * labels [n] int64; outputs nbr_idx [n,k], nbr_dist [n,k] (may be NULL), pred [n] (may be NULL; most frequent label,
 * the smallest one on ties).  k in {1, 3, 5, 10, 15}. */
int vince_knn_classify(const float* feats, const int64_t* labels, int32_t n, int32_t D, int32_t k, int64_t* nbr_idx,
                       float* nbr_dist, int64_t* pred, void* stream);

/* ---- multi-GPU: NCCL all-gather of the keys straight into the ring buffer -------------------------------------
 * replaces: the implicit nn.DataParallel gather of vince_model.py:35,125 + StorageQueue.enqueue; one process per
 * GPU.  `unique_id` is the 128-byte ncclUniqueId produced by vince_comm_unique_id on rank 0 and broadcast by the
 * host program.  vince_allgather_enqueue gathers keys[n_local, D] of every rank, in rank order, into
 * queue rows [tail, tail + world*n_local) mod K (and the TF32 shadow), using `scratch` (world*n_local*D floats). */
int vince_comm_unique_id(void* unique_id_128);
int vince_comm_init(void** comm_out, const void* unique_id_128, int32_t world, int32_t rank);
int vince_comm_destroy(void* comm);
int vince_allgather_enqueue(void* comm, const float* keys, int64_t n_local, int32_t D, float* queue, float* queue_tf32,
                            int64_t K, int64_t tail, float* scratch, void* stream);
/* same, with the momentum EMA of vince_ema_enqueue applied by the scatter kernel's launch (vince_solver.py:497-499 in
 * one call: all-gather + enqueue + vince_update) */
int vince_allgather_enqueue_ema(void* comm, const float* keys, int64_t n_local, int32_t D, float* queue,
                                float* queue_tf32, int64_t K, int64_t tail, float* scratch,
                                const vince_ema_chunk* table_dev, int32_t n_chunks, float momentum,
                                float one_minus_momentum, void* stream);

/* in-place sum all-reduce of the flat gradient buffer over the communicator of vince_comm_init (data-parallel training:
 * replaces the gradient reduction nn.DataParallel performs on GPU 0) */
int vince_allreduce_sum(void* comm, float* buf, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VINCE_B200_H_ */
