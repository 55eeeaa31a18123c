// extern "C" surface of libvince_b200 (see include/vince_b200.h).  Thin: validates, converts the POD structs to
// the internal descriptors, launches.  NCCL is bound at run time with dlopen so the library loads (and exports
// every symbol) on machines without NCCL or a GPU.
#include <dlfcn.h>
#include <stdlib.h>

#include "../../include/vince_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace vb;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline __half* HF(void* p) { return reinterpret_cast<__half*>(p); }
static inline const __half* HF(const void* p) { return reinterpret_cast<const __half*>(p); }

static BnSide to_side(const vince_bn_side* s) {
  BnSide o;
  memset(&o, 0, sizeof(o));
  if (s) o.raw = s->raw, o.coef = s->coef;
  return o;
}

static int check_side(const vince_bn_side* s, const char* what) {
  VB_REQUIRE(s != nullptr, "%s: null bn side", what);
  VB_REQUIRE(s->raw && s->coef, "%s: null pointer in bn side", what);
  return VB_OK;
}

extern "C" {

const char* vince_last_error(void) { return get_error(); }
int vince_abi_version(void) { return 4; }

int vince_conv_fwd(const vince_conv_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr, "vince_conv_fwd: null descriptor");
  ConvGemmDesc g;
  memset(&g, 0, sizeof(g));
  g.a_hi = d->a_hi, g.a_lo = d->a_lo, g.b_hi = d->w_hi, g.b_lo = d->w_lo, g.out = d->out;
  g.M = d->M, g.N = d->N, g.K = d->K, g.im2col = d->im2col;
  g.batch = d->batch, g.H = d->H, g.W = d->W, g.Cin = d->Cin, g.R = d->R, g.S = d->S, g.stride = d->stride;
  g.pad_lo_h = d->pad_lo_h, g.pad_lo_w = d->pad_lo_w, g.pad_hi_h = d->pad_hi_h, g.pad_hi_w = d->pad_hi_w;
  g.passes = d->passes, g.block_n = d->block_n, g.scale = d->scale, g.bias = d->bias, g.relu = d->relu;
  g.stats = d->stats;
  g.halo_mode = d->halo_mode;
  g.bn_gamma = d->bn_gamma, g.bn_beta = d->bn_beta, g.bn_running_mean = d->bn_running_mean;
  g.bn_running_var = d->bn_running_var, g.bn_num_batches_tracked = d->bn_num_batches_tracked;
  g.bn_coef = d->bn_coef, g.bn_counter = d->bn_counter, g.bn_momentum = d->bn_momentum, g.bn_eps = d->bn_eps;
  g.a_pixel_stride = d->a_pixel_stride, g.a_row_stride = d->a_row_stride, g.a_img_stride = d->a_img_stride;
  g.alpha = d->alpha;
  g.out_hi = d->out_hi, g.out_lo = d->out_lo, g.ep_coef = d->ep_coef, g.res_kind = d->res_kind;
  g.res_hi = d->res_hi, g.res_lo = d->res_lo, g.res_raw = d->res_raw, g.res_coef = d->res_coef;
  g.stats_only = d->stats_only;
  g.bn_save = d->bn_save, g.alpha_dev = d->alpha_dev;
  g.kchunk = d->kchunk, g.taps = d->taps, g.shift_w = d->shift_w;
  g.trace = getenv("VINCE_B200_TRACE_PTR") ? reinterpret_cast<void*>(strtoull(getenv("VINCE_B200_TRACE_PTR"), nullptr, 0)) : nullptr;
  return conv_gemm_launch(g, S(stream));
}

int vince_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, float* coef, int32_t C, void* stream) {
  VB_REQUIRE(C == 0 || (gamma && beta && running_mean && running_var && coef), "vince_bn_eval_coef: null pointer");
  return bn_eval_coef_launch(gamma, beta, running_mean, running_var, eps, coef, C, S(stream));
}

int vince_stem_pack(const float* x, const int64_t* gather_idx, void* x_hi, void* x_lo, int32_t N, int32_t H, int32_t W,
                    void* stream) {
  VB_REQUIRE(x && x_hi, "vince_stem_pack: null pointer");
  VB_REQUIRE(N >= 0 && H > 0 && W > 0, "vince_stem_pack: bad shape N=%d H=%d W=%d", N, H, W);
  return stem_pack_launch(x, gather_idx, HF(x_hi), HF(x_lo), N, H, W, 1, S(stream));
}

int vince_stem_pack_grid(const float* x, const int64_t* gather_idx, void* x_hi, void* x_lo, int32_t N, int32_t H,
                         int32_t W, int32_t grid, void* stream) {
  VB_REQUIRE(x && x_hi, "vince_stem_pack_grid: null pointer");
  VB_REQUIRE(N >= 0 && H > 0 && W > 0, "vince_stem_pack_grid: bad shape N=%d H=%d W=%d", N, H, W);
  return stem_pack_launch(x, gather_idx, HF(x_hi), HF(x_lo), N, H, W, grid, S(stream));
}

int vince_stem_pack_u8(const uint8_t* x_nhwc, const int64_t* gather_idx, const float* mean3, const float* std3,
                       void* x_hi, void* x_lo, int32_t N, int32_t H, int32_t W, void* stream) {
  VB_REQUIRE(x_nhwc && x_hi && mean3 && std3, "vince_stem_pack_u8: null pointer");
  VB_REQUIRE(N >= 0 && H > 0 && W > 0, "vince_stem_pack_u8: bad shape N=%d H=%d W=%d", N, H, W);
  return stem_pack_u8_launch(x_nhwc, gather_idx, mean3, std3, HF(x_hi), HF(x_lo), N, H, W, 1, S(stream));
}

int vince_stem_pack_u8_grid(const uint8_t* x_nhwc, const int64_t* gather_idx, const float* mean3, const float* std3,
                            void* x_hi, void* x_lo, int32_t N, int32_t H, int32_t W, int32_t grid, void* stream) {
  VB_REQUIRE(x_nhwc && x_hi && mean3 && std3, "vince_stem_pack_u8_grid: null pointer");
  VB_REQUIRE(N >= 0 && H > 0 && W > 0, "vince_stem_pack_u8_grid: bad shape N=%d H=%d W=%d", N, H, W);
  return stem_pack_u8_launch(x_nhwc, gather_idx, mean3, std3, HF(x_hi), HF(x_lo), N, H, W, grid, S(stream));
}

int vince_weight_prep(const vince_weight_entry* table_dev, int32_t n_entries, int64_t max_cout, void* w_hi, void* w_lo,
                      void* stream) {
  static_assert(sizeof(vince_weight_entry) == sizeof(WeightPrepEntry), "ABI struct mismatch");
  VB_REQUIRE(n_entries == 0 || (table_dev && w_hi), "vince_weight_prep: null pointer");
  VB_REQUIRE(max_cout >= 0 && max_cout <= 0x7fffffff, "vince_weight_prep: bad max_cout");
  return weight_prep_launch(reinterpret_cast<const WeightPrepEntry*>(table_dev), n_entries, (int)max_cout, HF(w_hi),
                            HF(w_lo), S(stream));
}

int vince_bn_apply(const vince_bn_side* main, int32_t res_kind, const void* res_hi, const void* res_lo,
                   const vince_bn_side* res_bn, int32_t relu, void* out_hi, void* out_lo, float* out_f32, int64_t M,
                   int32_t C, void* stream) {
  int rc = check_side(main, "vince_bn_apply");
  if (rc) return rc;
  VB_REQUIRE(res_kind >= 0 && res_kind <= 2, "vince_bn_apply: res_kind %d", res_kind);
  VB_REQUIRE(res_kind != 1 || res_hi, "vince_bn_apply: residual planes null");
  if (res_kind == 2 && (rc = check_side(res_bn, "vince_bn_apply(residual)"))) return rc;
  VB_REQUIRE(out_hi || out_f32, "vince_bn_apply: no output");
  return bn_apply_launch(to_side(main), res_kind, HF(res_hi), HF(res_lo), to_side(res_bn), relu, HF(out_hi), HF(out_lo),
                         out_f32, M, C, S(stream));
}

int vince_bn_relu_maxpool(const vince_bn_side* bn, void* out_hi, void* out_lo, int32_t N, int32_t P, int32_t Q, int32_t C,
                          void* stream) {
  int rc = check_side(bn, "vince_bn_relu_maxpool");
  if (rc) return rc;
  VB_REQUIRE(out_hi, "vince_bn_relu_maxpool: null output");
  const int P2 = (P + 2 - 3) / 2 + 1, Q2 = (Q + 2 - 3) / 2 + 1;
  return bn_relu_maxpool_launch(to_side(bn), HF(out_hi), HF(out_lo), N, P, Q, C, P2, Q2, S(stream));
}

int vince_bn_final_pool(const vince_bn_side* main, int32_t res_kind, const void* res_hi, const void* res_lo,
                        const vince_bn_side* res_bn, const int64_t* scatter_idx, float* spatial_nchw, float* pooled,
                        int32_t N, int32_t HW, int32_t C, void* stream) {
  int rc = check_side(main, "vince_bn_final_pool");
  if (rc) return rc;
  VB_REQUIRE(res_kind >= 0 && res_kind <= 2, "vince_bn_final_pool: res_kind %d", res_kind);
  VB_REQUIRE(res_kind != 1 || res_hi, "vince_bn_final_pool: residual planes null");
  if (res_kind == 2 && (rc = check_side(res_bn, "vince_bn_final_pool(residual)"))) return rc;
  VB_REQUIRE(pooled, "vince_bn_final_pool: pooled output null");
  return bn_final_pool_launch(to_side(main), res_kind, HF(res_hi), HF(res_lo), to_side(res_bn), scatter_idx,
                              spatial_nchw, pooled, N, HW, C, S(stream));
}

int vince_count_saturated(const void* plane, int64_t n, uint64_t* count, void* stream) {
  VB_REQUIRE(n == 0 || (plane && count), "vince_count_saturated: null pointer");
  return count_saturated_launch(reinterpret_cast<const __half*>(plane), n, reinterpret_cast<unsigned long long*>(count),
                                S(stream));
}

int vince_split_f16(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  VB_REQUIRE(n == 0 || (x && hi), "vince_split_f16: null pointer");
  return split_f16_launch(x, HF(hi), HF(lo), n, S(stream));
}
int vince_round_tf32(const float* x, float* out, int64_t n, void* stream) {
  VB_REQUIRE(n == 0 || (x && out), "vince_round_tf32: null pointer");
  return round_tf32_launch(x, out, n, S(stream));
}
int vince_l2_normalize(const float* x, float* out, int32_t rows, int32_t D, float eps, void* stream) {
  VB_REQUIRE(rows == 0 || (x && out), "vince_l2_normalize: null pointer");
  return l2_normalize_launch(x, out, rows, D, eps, S(stream));
}
int vince_jigsaw_patchify(const float* x, const int64_t* gather_idx, float* out, int32_t N, int32_t C, int32_t H,
                          int32_t W, void* stream) {
  VB_REQUIRE(N == 0 || (x && out), "vince_jigsaw_patchify: null pointer");
  // vince_model.py:145-146: BOTH axes grow by 3 - dim % 3 when EITHER is not a multiple of 3
  const bool pad = (H % 3) != 0 || (W % 3) != 0;
  const int Hp = pad ? H + 3 - H % 3 : H, Wp = pad ? W + 3 - W % 3 : W;
  return jigsaw_patchify_launch(x, gather_idx, out, N, C, H, W, Hp / 3, Wp / 3, S(stream));
}
int vince_jigsaw_gather(const float* in, const int64_t* order, float* out, int32_t N, int32_t C, void* stream) {
  VB_REQUIRE(N == 0 || (in && order && out), "vince_jigsaw_gather: null pointer");
  return jigsaw_gather_launch(in, order, out, N, C, S(stream));
}

size_t vince_infonce_workspace_bytes(int32_t B, int32_t D) { return infonce_workspace_bytes(B, D); }

int vince_infonce_fwd(const vince_infonce_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr, "vince_infonce_fwd: null descriptor");
  VB_REQUIRE(d->dists && d->weights && d->pos_sim && d->neg_max && d->row_lse && d->scalars,
             "vince_infonce_fwd: null output pointer");
  InfoNceDesc n;
  memset(&n, 0, sizeof(n));
  n.q = d->q, n.keys = d->keys, n.queue_tf32 = d->queue_tf32;
  n.B = d->B, n.Bk = d->Bk, n.K = d->K, n.D = d->D, n.num_frames = d->num_frames, n.temperature = d->temperature;
  n.dists = d->dists, n.weights = d->weights, n.pos_sim = d->pos_sim, n.neg_max = d->neg_max, n.row_lse = d->row_lse;
  n.scalars = d->scalars, n.workspace = d->workspace;
  return infonce_fwd_launch(n, S(stream));
}

size_t vince_infonce_bwd_workspace_bytes(int32_t B, int32_t D) { return infonce_bwd_workspace_bytes(B, D); }

int vince_infonce_bwd(const vince_infonce_desc* d, float grad_dist, int32_t symmetric, int32_t accumulate, float* dq,
                      void* stream) {
  VB_REQUIRE(d != nullptr, "vince_infonce_bwd: null descriptor");
  InfoNceDesc n;
  memset(&n, 0, sizeof(n));
  n.q = d->q, n.keys = d->keys, n.queue_tf32 = d->queue_tf32;
  n.B = d->B, n.Bk = d->Bk, n.K = d->K, n.D = d->D, n.num_frames = d->num_frames, n.temperature = d->temperature;
  n.pos_sim = d->pos_sim, n.row_lse = d->row_lse, n.workspace = d->workspace;
  return infonce_bwd_launch(n, grad_dist, symmetric, accumulate, dq, S(stream));
}

int vince_masked_ce_fwd(const float* sims, const uint8_t* mask, int32_t rows, int32_t cols, int32_t n_pos,
                        float temperature, float* dists, float* weights, float* pos_sim, float* neg_max, float* row_lse,
                        float* scalars, int32_t* error_flag, void* stream) {
  VB_REQUIRE(sims && mask && dists && weights && pos_sim && neg_max && row_lse && scalars && error_flag,
             "vince_masked_ce_fwd: null pointer");
  return masked_ce_launch(sims, mask, rows, cols, n_pos, temperature, dists, weights, pos_sim, neg_max, row_lse, scalars,
                          error_flag, S(stream));
}

int vince_ema_enqueue(const vince_ema_chunk* table_dev, int32_t n_chunks, float momentum, float one_minus_momentum,
                      float* queue, float* queue_tf32, const float* keys, int64_t n0, int64_t dst0, int64_t n1,
                      int64_t dst1, int64_t src1, void* stream) {
  static_assert(sizeof(vince_ema_chunk) == sizeof(EmaChunk), "ABI struct mismatch");
  VB_REQUIRE(n_chunks == 0 || table_dev, "vince_ema_enqueue: null table");
  VB_REQUIRE((n0 + n1 == 0) || (queue && keys), "vince_ema_enqueue: null queue / keys");
  VB_REQUIRE(n0 >= 0 && n1 >= 0, "vince_ema_enqueue: negative count");
  return ema_enqueue_launch(reinterpret_cast<const EmaChunk*>(table_dev), n_chunks, momentum, one_minus_momentum, queue,
                            queue_tf32, keys, n0, dst0, n1, dst1, src1, S(stream));
}

// ---- query-encoder backward -----------------------------------------------------------------
int vince_bn_bwd(const vince_bn_bwd_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr, "vince_bn_bwd: null descriptor");
  static_assert(sizeof(vince_bn_bwd_desc) == sizeof(BnBwdDesc), "ABI struct mismatch");
  return bn_bwd_launch(*reinterpret_cast<const BnBwdDesc*>(d), S(stream));
}
int vince_transpose_pad(const void* src_hi, const void* src_lo, void* dst_hi, void* dst_lo, int64_t M, int32_t C, int32_t P,
                        int32_t Q, int32_t stride, int32_t offset, int32_t Hp, int32_t Wp, int64_t ld, int32_t copies,
                        void* stream) {
  return transpose_pad_launch(HF(src_hi), HF(src_lo), HF(dst_hi), HF(dst_lo), M, C, P, Q, stride, offset, Hp, Wp, ld,
                              copies, S(stream));
}
int vince_wgrad_reduce(const float* partials, int32_t taps, int32_t splits, int32_t Mpad, int32_t Cout, int32_t Cin,
                       const float* scale_dev, float scale, float* grad, int32_t accumulate, void* stream) {
  return wgrad_reduce_launch(partials, taps, splits, Mpad, Cout, Cin, scale_dev, scale, grad, accumulate, S(stream));
}
int vince_maxpool_bwd(const float* dA, const float* dB, const float* raw, const float* coef, float* dst, int32_t N,
                      int32_t P, int32_t Q, int32_t C, void* stream) {
  return maxpool_bwd_launch(dA, dB, raw, coef, dst, N, P, Q, C, S(stream));
}
int vince_stem_wgrad(const float* x, const uint8_t* x_u8, const int64_t* gather_idx, const float* mean3,
                     const float* std3, const float* draw, float* grad, int32_t N, int32_t H, int32_t W,
                     int32_t accumulate, void* stream) {
  return stem_wgrad_launch(x, x_u8, gather_idx, mean3, std3, draw, grad, N, H, W, accumulate, S(stream));
}
int vince_sgemm(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t lda, int32_t ldb,
                int32_t ldc, int32_t trans_a, int32_t trans_b, int32_t accumulate, const float* relu_mask_src,
                void* stream) {
  return sgemm_launch(A, B, C, M, N, K, lda, ldb, ldc, trans_a, trans_b, accumulate, relu_mask_src, S(stream));
}
int vince_colsum(const float* x, float* out, int32_t R, int32_t C, int32_t accumulate, void* stream) {
  return colsum_launch(x, out, R, C, accumulate, S(stream));
}
int vince_normalize_bwd(const float* x, const float* dy, float* dx, int32_t rows, int32_t D, float eps, float gscale,
                        void* stream) {
  return normalize_bwd_launch(x, dy, dx, rows, D, eps, gscale, S(stream));
}
int vince_sgd_step(const vince_sgd_chunk* table_dev, int32_t n_chunks, float lr, float momentum, float weight_decay,
                   float grad_scale, int32_t first_step, void* stream) {
  static_assert(sizeof(vince_sgd_chunk) == sizeof(SgdChunk), "ABI struct mismatch");
  return sgd_launch(reinterpret_cast<const SgdChunk*>(table_dev), n_chunks, lr, momentum, weight_decay, grad_scale,
                    first_step, S(stream));
}

int vince_knn_classify(const float* feats, const int64_t* labels, int32_t n, int32_t D, int32_t k, int64_t* nbr_idx,
                       float* nbr_dist, int64_t* pred, void* stream) {
  VB_REQUIRE(feats && nbr_idx, "vince_knn_classify: null pointer");
  VB_REQUIRE(!pred || labels, "vince_knn_classify: predictions need labels");
  return knn_launch(feats, labels, n, D, k, nbr_idx, nbr_dist, pred, S(stream));
}

// ---------------------------------------------------------------------------------------------------------------
// NCCL (bound lazily)
// ---------------------------------------------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
static struct {
  void* handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int /*dtype*/, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
} g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return VB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);   // prefer the instance torch already loaded
    if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("libnccl.so.2 could not be loaded: %s", dlerror());
    return VB_ERR_NCCL;
  }
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(h, "ncclAllGather"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) {
    set_error("libnccl is missing required symbols");
    return VB_ERR_NCCL;
  }
  g_nccl.handle = h;
  return VB_OK;
}

#define VB_CHECK_NCCL(expr)                                                                              \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != 0) {                                                                                       \
      set_error("NCCL error %d (%s) at %s", (int)_r, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", #expr); \
      return VB_ERR_NCCL;                                                                                \
    }                                                                                                    \
  } while (0)

int vince_comm_unique_id(void* unique_id_128) {
  VB_REQUIRE(unique_id_128, "vince_comm_unique_id: null buffer");
  int rc = load_nccl();
  if (rc) return rc;
  VB_CHECK_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(unique_id_128)));
  return VB_OK;
}

int vince_comm_init(void** comm_out, const void* unique_id_128, int32_t world, int32_t rank) {
  VB_REQUIRE(comm_out && unique_id_128, "vince_comm_init: null pointer");
  VB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "vince_comm_init: bad rank %d / world %d", rank, world);
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, unique_id_128, sizeof(id));
  ncclComm_t comm = nullptr;
  VB_CHECK_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  *comm_out = comm;
  return VB_OK;
}

int vince_comm_destroy(void* comm) {
  if (!comm) return VB_OK;
  int rc = load_nccl();
  if (rc) return rc;
  VB_CHECK_NCCL(g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(comm)));
  return VB_OK;
}

static int allgather_enqueue_impl(void* comm, const float* keys, int64_t n_local, int32_t D, float* queue,
                                  float* queue_tf32, int64_t K, int64_t tail, float* scratch,
                                  const vince_ema_chunk* table_dev, int32_t n_chunks, float momentum, float one_minus,
                                  void* stream) {
  VB_REQUIRE(comm && keys && queue && scratch, "vince_allgather_enqueue: null pointer");
  VB_REQUIRE(D > 0 && D % 4 == 0, "vince_allgather_enqueue: D=%d must be a positive multiple of 4", D);
  VB_REQUIRE(K > 0 && tail >= 0 && tail <= K, "vince_allgather_enqueue: bad tail %lld for K=%lld", (long long)tail,
             (long long)K);
  VB_REQUIRE(n_chunks == 0 || table_dev, "vince_allgather_enqueue: null EMA table");
  int rc = load_nccl();
  if (rc) return rc;
  int world = 0;
  {
    typedef ncclResult_t (*CountFn)(ncclComm_t, int*);
    CountFn count = reinterpret_cast<CountFn>(dlsym(g_nccl.handle, "ncclCommCount"));
    VB_REQUIRE(count, "libnccl lacks ncclCommCount");
    VB_CHECK_NCCL(count(reinterpret_cast<ncclComm_t>(comm), &world));
  }
  const int64_t total_rows = n_local * world;
  VB_REQUIRE(total_rows <= K, "vince_allgather_enqueue: gathered batch (%lld rows) exceeds the queue (%lld)",
             (long long)total_rows, (long long)K);
  // rank-ordered gather (ncclFloat32 == 7), then ONE kernel scatters into the ring (two slices on wrap-around) and,
  // when a table is given, applies the momentum EMA in the same launch
  VB_CHECK_NCCL(g_nccl.AllGather(keys, scratch, (size_t)(n_local * D), 7, reinterpret_cast<ncclComm_t>(comm), S(stream)));
  int64_t t = tail;
  if (t + total_rows > K && t == K) t = 0;   // storage_queue.py:35-43 with an empty head slice
  const int64_t first = (t + total_rows > K) ? (K - t) : total_rows;
  const int64_t second = total_rows - first;
  return ema_enqueue_launch(reinterpret_cast<const EmaChunk*>(table_dev), n_chunks, momentum, one_minus, queue,
                            queue_tf32, scratch, first * D, t * D, second * D, 0, first * D, S(stream));
}

int vince_allgather_enqueue(void* comm, const float* keys, int64_t n_local, int32_t D, float* queue, float* queue_tf32,
                            int64_t K, int64_t tail, float* scratch, void* stream) {
  return allgather_enqueue_impl(comm, keys, n_local, D, queue, queue_tf32, K, tail, scratch, nullptr, 0, 0.f, 1.f,
                                stream);
}

int vince_allgather_enqueue_ema(void* comm, const float* keys, int64_t n_local, int32_t D, float* queue,
                                float* queue_tf32, int64_t K, int64_t tail, float* scratch,
                                const vince_ema_chunk* table_dev, int32_t n_chunks, float momentum,
                                float one_minus_momentum, void* stream) {
  return allgather_enqueue_impl(comm, keys, n_local, D, queue, queue_tf32, K, tail, scratch, table_dev, n_chunks,
                                momentum, one_minus_momentum, stream);
}

int vince_allreduce_sum(void* comm, float* buf, int64_t n, void* stream) {
  VB_REQUIRE(comm && buf && n >= 0, "vince_allreduce_sum: bad argument");
  int rc = load_nccl();
  if (rc) return rc;
  VB_REQUIRE(g_nccl.AllReduce, "libnccl lacks ncclAllReduce");
  if (n == 0) return VB_OK;
  // in place, ncclFloat32 == 7, ncclSum == 0
  VB_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, 7, 0, reinterpret_cast<ncclComm_t>(comm), S(stream)));
  return VB_OK;
}

}  // extern "C"
