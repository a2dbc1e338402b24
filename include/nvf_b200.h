/*
 * nvf_b200.h - C ABI of the B200-native NVF leaf-block decoder hot path.
 *
 * This is the drop-in boundary for the ONE data-parallel path of huzi96/NVFPCC
 * that this repository accelerates: the per-leaf-block Neural Volumetric Field
 * decoder (forward for encode/decode; forward + rate-distortion loss + backward
 * for train).  The reference has no FFI layer of its own - its boundary is the
 * Python nn.Module pair
 *     CompDecoder.forward(x, q)        utils/network.py:4758-4779
 *     Net.forward / Net.reconstruct    NVFPCC.py:41-49
 * plus the loss functions utils/loss.py:61-121 and the threshold idiom
 * NVFPCC.py:631-637.  Each entry point below names the reference lines it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain C, no torch types; all pointers are DEVICE pointers unless the
 *     name ends in _host; all tensors fp32, contiguous, native PyTorch layout
 *     (NCDHW; convT kernels (Cin,Cout,5,5,5); conv kernels (Cout,Cin,k,k,k)).
 *   - "effective" weights: the caller applies the tiny parameter transforms
 *     (kernel quantisation/noise + kernel_init, b + b_init, GDN reparam,
 *     utils/network.py:606-620, gdn_3d.py:143-150) and passes the results.
 *   - the library never allocates, frees or retains device memory: the caller
 *     provides inputs, outputs and a workspace of nvf_workspace_bytes().
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*)
 *     and performs no host synchronisation.
 *   - return value: NVF_OK or a negative error code; nvf_strerror() decodes it.
 *     Nothing throws or exits across this boundary.
 */
#ifndef NVF_B200_H_
#define NVF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVF_ABI_VERSION 1

enum {
  NVF_OK = 0,
  NVF_ERR_INVALID_ARG = -1,  /* NULL pointer, negative size, ...            */
  NVF_ERR_UNSUPPORTED = -2,  /* channel configuration not compiled in       */
  NVF_ERR_WORKSPACE = -3,    /* workspace too small                         */
  NVF_ERR_CUDA = -4,         /* a CUDA runtime call failed (see nvf_last_cuda_error) */
  NVF_ERR_NO_DEVICE = -5,    /* no sm_100 device / kernel image not loadable */
  NVF_ERR_BITSTREAM = -6     /* corrupt or truncated entropy-coded stream (nvf_prep_b200.h) */
};

/* Decoder geometry: latent channels `ch`, chanstr c0,c1,c2,c3
 * (NVFPCC.py:34-39 Net.__init__, --ch / --chanstr). */
typedef struct NvfDesc {
  int32_t ch, c0, c1, c2, c3;
} NvfDesc;

/* Effective decoder tensors (device pointers).  Order follows
 * CompDecoder.__init__, utils/network.py:4658-4751. */
typedef struct NvfWeights {
  const float* up0_w;   const float* up0_b;    /* (ch,c0,5,5,5), (c0)  convT s2 p2 op1 */
  const float* igdn_beta; const float* igdn_gamma; /* (c0), (c0,c0) effective, gdn_3d.py:143-150 */
  const float* conv0_w; const float* conv0_b;  /* (c0,c1,5,5,5), (c1)  convT s2 p2 op1 */
  const float* up1_w;   const float* up1_b;    /* (c1,c2,5,5,5), (c2)  convT s2 p0     */
  const float* conv1_w; const float* conv1_b;  /* (c2,c2,4,4,4), (c2)  conv  k4 p0     */
  const float* up2_w;   const float* up2_b;    /* (c2,c3,5,5,5), (c3)  convT s2 p0     */
  const float* conv2_w; const float* conv2_b;  /* (c3,c3,4,4,4), (c3)  conv  k4 p0     */
  const float* cls2_w;  const float* cls2_b;   /* (1,c3,3,3,3), (1)    conv2_cls       */
  const float* cls1_w;  const float* cls1_b;   /* (1,c2,3,3,3), (1)    conv1_cls; NULL ok for decode */
  const float* cls0_w;  const float* cls0_b;   /* (1,c1,3,3,3), (1)    conv0_cls; NULL ok for decode */
} NvfWeights;

/* Gradients w.r.t. the effective tensors, same shapes as NvfWeights. */
typedef struct NvfWeightGrads {
  float* up0_w;   float* up0_b;
  float* igdn_beta; float* igdn_gamma;
  float* conv0_w; float* conv0_b;
  float* up1_w;   float* up1_b;
  float* conv1_w; float* conv1_b;
  float* up2_w;   float* up2_b;
  float* conv2_w; float* conv2_b;
  float* cls2_w;  float* cls2_b;
  float* cls1_w;  float* cls1_b;
  float* cls0_w;  float* cls0_b;
} NvfWeightGrads;

enum { NVF_MODE_DECODE = 0, NVF_MODE_TRAIN = 1 };

int nvf_abi_version(void);
const char* nvf_strerror(int code);
/* cudaError_t of the last failing CUDA call on this thread (0 if none). */
int nvf_last_cuda_error(void);
/* number of CUDA kernels this library has launched so far in this process. */
long long nvf_launch_count(void);
/* 1 if a fused (single-kernel, activations on chip) decode kernel exists for desc. */
int nvf_has_fused_decode(const NvfDesc* desc);

/* Workspace size for a call with `n_blocks` leaf blocks in `mode`. */
int nvf_workspace_bytes(const NvfDesc* desc, int64_t n_blocks, int mode, size_t* bytes_out);

/*
 * Decode: latents -> occupancy -> points.  Replaces the per-block loop of
 * decode(), NVFPCC.py:625-638 (net.reconstruct -> to_sparse -> `F > thh`
 * pruning -> coords + origin) and the same idiom in encode(), :516-539, by ONE
 * batched call.
 *   latent     [n_blocks, ch, 2,2,2]   rounded latents (NVFPCC.py:604-607)
 *   origins    [n_blocks, 3] int32 leaf origins or NULL (then coords are in-block)
 *   thh        strict threshold on the sigmoid probability (NVFPCC.py:632)
 *   prob_out   [n_blocks,1,32,32,32] or NULL  - dense probabilities (`out_dense`)
 *   mask_out   [n_blocks,1024] uint32 or NULL - bit k of word (i*32+j) = p[i,j,k] > thh
 *   counts_out [n_blocks] int32               - points per block
 *   coords_out [cap,3] int32 or NULL          - origin + (i,j,k), block order then
 *                                               row-major (i,j,k) order
 *   total_out  [1] int64                      - total number of points (may exceed cap;
 *                                               only the first cap are written)
 */
int nvf_decode(const NvfDesc* desc, const NvfWeights* w, const float* latent, const int32_t* origins,
               int64_t n_blocks, float thh, float* prob_out, uint32_t* mask_out, int32_t* counts_out,
               int32_t* coords_out, int64_t cap, int64_t* total_out, void* workspace, size_t workspace_bytes,
               void* stream);

/*
 * Second half of nvf_decode on its own: occupancy masks + counts -> ordered
 * points (used when the caller's first `cap` was too small).  Workspace:
 * (n_blocks+1) * 8 bytes.  Replaces NVFPCC.py:631-637 like nvf_decode.
 */
int nvf_emit_points(const uint32_t* mask, const int32_t* counts, const int32_t* origins, int64_t n_blocks,
                    int32_t* coords_out, int64_t cap, int64_t* total_out, void* workspace, size_t workspace_bytes,
                    void* stream);

/*
 * Training forward (CompDecoder.forward in train mode incl. aux heads,
 * utils/network.py:4758-4768): writes probabilities of the three heads and
 * keeps every activation needed by nvf_train_backward in `workspace`.
 *   out  [n,1,32,32,32], cls1 [n,1,16,16,16], cls0 [n,1,8,8,8]
 */
int nvf_train_forward(const NvfDesc* desc, const NvfWeights* w, const float* latent, int64_t n_blocks, float* out,
                      float* cls1, float* cls0, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Distortion losses + metrics + gradient seeds.  Replaces MultiscaleProcessor
 * (NVFPCC.py:76-88), get_surf_focal_dense (utils/loss.py:94-111, beta=1,
 * alpha=alpha_main), get_focal_dense x2 (utils/loss.py:61-72, alpha=alpha_aux),
 * get_sse1 / get_acc_dense (utils/loss.py:74-84,113-121).
 *   gt [n,1,32,32,32] fp32 0/1, dist [n,1,32,32,32] fp32
 *   sums_out [16] float64: 0 bce, 1 ms0 (8^3 head), 2 ms1 (16^3 head),
 *     3 sse(thh_metric), 4 denom, 5..8 tp,ap,tn,an main head (thh 0.5),
 *     9..12 tp,ap,tn,an 8^3 head, 13..16 -> see NVF_LOSS_SUMS
 *   g_out/g_cls1/g_cls0: dLoss/dprobability * 1 (same shapes as the heads) or
 *     NULL; when given they are the seeds nvf_train_backward consumes.
 */
#define NVF_LOSS_SUMS 20
#define NVF_LOSS_CHUNKS 16 /* workspace: sizeof(double) * NVF_LOSS_SUMS * (NVF_LOSS_CHUNKS * n_blocks + 1) bytes */
int nvf_loss_seeds(const float* out, const float* cls1, const float* cls0, const float* gt, const float* dist,
                   int64_t n_blocks, float alpha_main, float alpha_aux, float thh_metric, double* sums_out,
                   float* g_out, float* g_cls1, float* g_cls0, void* workspace, size_t workspace_bytes,
                   void* stream);

/*
 * Training backward: given dLoss/dprobability of the three heads, accumulates
 * dLoss/d(effective weights) and writes dLoss/d(latent).  Replaces autograd
 * through CompDecoder.forward.
 *   g_out/g_cls1/g_cls0  gradients w.r.t. the sigmoid outputs (NULL = zero)
 *   flags: NVF_BWD_WGRAD compute weight grads (else skipped, NVFPCC.py:225-251
 *          discards them), NVF_BWD_DLATENT compute d_latent (NVFPCC.py:149-223
 *          discards it).  Weight-gradient buffers are OVERWRITTEN.
 *          NVF_BWD_DLOGIT: g_out/g_cls1/g_cls0 are gradients w.r.t. the LOGITS of the
 *          three heads (the sigmoid derivative is already folded in; all three required).
 */
enum { NVF_BWD_WGRAD = 1, NVF_BWD_DLATENT = 2, NVF_BWD_DLOGIT = 4 };
int nvf_train_backward(const NvfDesc* desc, const NvfWeights* w, const float* latent, int64_t n_blocks,
                       const float* g_out, const float* g_cls1, const float* g_cls0, int flags,
                       const NvfWeightGrads* gw, float* g_latent, void* workspace, size_t workspace_bytes,
                       void* stream);

/*
 * Parameter-side transforms of the training step, fused (forward + backward).
 * Replaces, per call, the ~250 small framework launches behind
 *   Q{ConvTranspose3d,Conv3d}.forward / IConv3d.forward weight preparation
 *     (W_eff = Q_q(kernel) + kernel_init, b_eff = b + b_init; utils/network.py:606-620, 669-686, 735-740),
 *   the IGDN reparametrisation (gdn_3d.py:143-150) and
 *   net_bits = [likelihood_model(bypass_round16(p)) for p in get_q_params()] (utils/network.py:4777-4778, 287-305).
 * Layer order of the arrays: up0 conv0 up1 conv1 up2 conv2 conv2_cls conv1_cls conv0_cls
 * (the first NVF_NUM_QUANT are the quantised ones).
 */
#define NVF_NUM_CONV 9
#define NVF_NUM_QUANT 7
#define NVF_PARAM_WS_BYTES (64 * 1024)
typedef struct NvfParamSet {
  const float* kernel[NVF_NUM_CONV];       /* trainable deltas, native PyTorch layouts   */
  const float* kernel_init[NVF_NUM_CONV];  /* seed-derived buffers                       */
  const float* b[NVF_NUM_CONV];
  const float* b_init[NVF_NUM_CONV];
  const float* igdn_beta;                  /* raw (c0)                                   */
  const float* igdn_gamma;                 /* raw (c0,c0)                                */
  const float* lik_sigma;                  /* likelihood_model.sigma (1)                 */
  const float* lik_mu;                     /* likelihood_model.mu (1)                    */
} NvfParamSet;
typedef struct NvfParamGrads {
  float* kernel[NVF_NUM_CONV];
  float* b[NVF_NUM_CONV];
  float* igdn_beta; float* igdn_gamma; float* lik_sigma; float* lik_mu;
} NvfParamGrads;
/*   q      0 none, 1 additive U(-1/32,1/32) noise, 2 straight-through round16
 *   noise  U(0,1) samples, one per element of the 7 quantised kernels in layer order (q == 1), else NULL
 *   eff    OUT: the 20 effective tensors (same struct as the weight gradients: plain float*)
 *   net_bits OUT [NVF_NUM_QUANT] */
int nvf_param_prep(const NvfDesc* desc, const NvfParamSet* params, int q, const float* noise, float beta_bound,
                   float gamma_bound, float pedestal, const NvfWeightGrads* eff, float* net_bits, void* workspace,
                   size_t workspace_bytes, void* stream);
/*   g_eff      gradients w.r.t. the 20 effective tensors, g_net_bits [NVF_NUM_QUANT]
 *   out        OUT: gradients w.r.t. kernel / b / raw IGDN beta,gamma / likelihood sigma,mu */
int nvf_param_prep_backward(const NvfDesc* desc, const NvfParamSet* params, float beta_bound, float gamma_bound,
                            const NvfWeights* g_eff, const float* g_net_bits, const NvfParamGrads* out,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * Latent head, fused (forward + backward).  Replaces SingleLayerLatentGen.forward
 * (utils/network.py:4610-4612: 1x1x1 IConv3d + GDN3d, gdn_3d.py:78-92) followed by
 * QuantGaussianLikelihood.forward (utils/network.py:4514-4539: straight-through round,
 * rate of the noisy (train) / rounded (eval) latent under N(mu_c, |sigma_c|), floor 1e-8).
 *   ch        latent channels, 1..4
 *   emb       [n, ch, 2,2,2]
 *   noise     [n, ch, 2,2,2] U(0,1) samples or NULL; the rate argument in train mode is
 *             y + (noise - 0.5) * noise_scale   (:4516-4525)
 *   latent_out [n, ch, 2,2,2] rounded latent (decoder input); bits_out [1] summed rate in bits
 *   workspace NVF_LATENT_WS_BYTES, zero-filled once by the caller; calls leave it zeroed where it matters
 * Backward: g_latent = dLoss/d(latent_out) or NULL, g_bits [1] = dLoss/d(bits_out) (device scalar);
 *   `grads` (all six pointers) or NULL, g_emb [n,ch,2,2,2] or NULL.
 */
#define NVF_LATENT_WS_BYTES (128 * 1024)
typedef struct NvfLatentParams {
  const float* kernel; const float* kernel_init;  /* latent_gen.h_analysis_2 (ch,ch,1,1,1) */
  const float* b; const float* b_init;            /* (ch) */
  const float* gdn_beta; const float* gdn_gamma;  /* latent_gen.gdn_2 raw (ch), (ch,ch) */
  const float* sigma; const float* mu;            /* entropy_coder (1,ch,1,1,1) */
} NvfLatentParams;
typedef struct NvfLatentGrads {
  float* kernel; float* b; float* gdn_beta; float* gdn_gamma; float* sigma; float* mu;
} NvfLatentGrads;
int nvf_latent_forward(int ch, const NvfLatentParams* params, const float* emb, const float* noise, float noise_scale,
                       int train, int64_t n_blocks, float beta_bound, float gamma_bound, float pedestal,
                       float* latent_out, float* bits_out, void* workspace, size_t workspace_bytes, void* stream);
int nvf_latent_backward(int ch, const NvfLatentParams* params, const float* emb, const float* noise,
                        float noise_scale, int train, int64_t n_blocks, float beta_bound, float gamma_bound,
                        float pedestal, const float* g_latent, const float* g_bits, const NvfLatentGrads* grads,
                        float* g_emb, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Total rate-distortion loss and the scalars train() logs (NVFPCC.py:161-164,190-221,196):
 *   loss = bce + ms0 + ms1 + lmbda * (w1 * latent_bits / n_pts + w2 * sum(net_bits) / n_total)
 *   sums        the NVF_LOSS_SUMS doubles of nvf_loss_seeds ([0] bce, [1] ms0, [2] ms1)
 *   n_pts       [1] device scalar: occupied voxels of the (global) batch (NVFPCC.py:154)
 *   loss_out [1]; stats_out [7] or NULL: loss bce ms0 ms1 b_latent b_net n_pts
 * Backward: g_loss [1] -> g_dist [3] (bce, ms0, ms1), g_latent_bits [1], g_net_bits [NVF_NUM_QUANT].
 */
int nvf_rd_total(const double* sums, const float* latent_bits, const float* net_bits, const float* n_pts,
                 float n_total, float lmbda, float w1, float w2, float* loss_out, float* stats_out, void* stream);
int nvf_rd_total_backward(const float* g_loss, const float* n_pts, float n_total, float lmbda, float w1, float w2,
                          float* g_dist, float* g_latent_bits, float* g_net_bits, void* stream);

/*
 * One Adam update of a flat parameter buffer (torch.optim.Adam defaults as used by train(),
 * NVFPCC.py:116,124,222,250): no weight decay, no amsgrad.  `step` [1] float device counter
 * (incremented by the call), `lr` [1] device scalar (so captured graphs follow LR schedules).
 */
int nvf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* step,
                  const float* lr, float beta1, float beta2, float eps, void* stream);

/*
 * Data-parallel weight loop (NVFPCC.py:154,161,222 with one process per GPU): the all-reduce of the flat
 * shared-weight gradient FUSED with nvf_adam_step, over NVLink / NVSwitch peer memory instead of an NCCL call plus two
 * launches.  Every rank allocates one symmetric buffer with nvf_symm_alloc (cudaMalloc + a 64-byte CUDA IPC handle
 * the host code exchanges, e.g. with torch.distributed.all_gather_object) and maps its peers' buffers with
 * nvf_symm_open; bytes >= nvf_symm_bytes(n).  nvf_adam_allreduce_step then does, in ONE kernel per rank: copy the
 * rank's gradient into its buffer, publish a sequence flag to every peer (st.release.sys), wait for all peers'
 * flags, read the `world` gradients over NVLink, add them in rank order (bit-identical weights on every rank) and
 * apply the Adam update of nvf_adam_step (same arithmetic; `step` is advanced by the call).
 *   peers [world] host array: peers[r] = rank r's buffer as mapped in this process (peers[rank] = own buffer)
 *   ctl   16 zero-initialised device bytes owned by the caller (sequence number + tickets), one per optimizer
 * Every rank must make the same sequence of calls (as with any collective).  world <= 16, one node.
 */
size_t nvf_symm_bytes(int64_t n);
int nvf_symm_alloc(size_t bytes, void** ptr_out, void* handle64_out);
int nvf_symm_open(const void* handle64, void** ptr_out);
int nvf_symm_close(void* ptr);
int nvf_symm_free(void* ptr);
int nvf_adam_allreduce_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* step,
                            const float* lr, float beta1, float beta2, float eps, void* const* peers, int rank,
                            int world, void* ctl, void* stream);

/*
 * One weight-loop step of train() as ONE call (NVFPCC.py:149-197 up to `loss.backward()`): latent head ->
 * parameter transforms -> decoder forward -> rate-distortion loss -> backward of all of it, with the gradients
 * of the RAW trainable tensors written to caller-chosen destinations (slices of one flat gradient buffer, so the
 * all-reduce and nvf_adam_step follow without a gather).  It composes the entry points above behind one boundary
 * and replaces what only exists between them in a framework: the minibatch gather (blocks read their dataset rows
 * through `idx`), the noise tensors (drawn in-kernel: Philox4x32-10 keyed by `seed`, indexed by the device-resident
 * `rng_counter`, which the call advances), the loss finalisation / total loss / cotangent kernels (last-CTA
 * finalisation inside the loss kernel) and the packing of the effective kernels (written by the parameter kernel).
 *   emb [n_rows, ch,2,2,2], gt / dist [n_rows,1,32,32,32]: the resident dataset when idx != NULL (idx [n] int64,
 *   device; out-of-range -> row 0 + bit 0 of *status), else the batch itself (n_rows ignored).
 *   flags: NVF_BWD_WGRAD -> g_latent / g_params are written (every pointer required);
 *          NVF_BWD_DLATENT -> g_emb [n, ch,2,2,2] (batch order) is written.
 *   noise_latent [n,ch,2,2,2] / noise_kernel (layout of nvf_param_prep): explicit U(0,1) draws (tests); NULL = in-kernel.
 *   n_pts [1] device scalar (NVFPCC.py:154, batch-global); w2_grad: weight of the network-rate term in the backward
 *   pass (w2 on one rank, 0 on the others under data parallelism).
 *   stats [7] (loss bce ms0 ms1 b_latent b_net n_pts), sums [NVF_LOSS_SUMS] double.
 */
typedef struct NvfStepArgs {
  NvfDesc desc;
  int32_t n, q, flags, train_mode;
  const float* emb; const float* gt; const float* dist; const int64_t* idx; int64_t n_rows; int32_t* status;
  NvfLatentParams latent; NvfParamSet params;
  NvfLatentGrads g_latent; NvfParamGrads g_params; float* g_emb;
  const float* noise_latent; const float* noise_kernel;
  uint64_t seed; uint64_t* rng_counter;
  const float* n_pts;
  float n_total, lmbda, w1, w2, w2_grad, alpha_main, alpha_aux, thh_metric, noise_scale;
  float latent_beta_bound, latent_gamma_bound, latent_pedestal, igdn_beta_bound, igdn_gamma_bound, igdn_pedestal;
  float* stats; double* sums;
} NvfStepArgs;
int nvf_train_step_workspace_bytes(const NvfDesc* desc, int64_t n_blocks, size_t* bytes_out);
/* workspace: nvf_train_step_workspace_bytes(), zero-filled ONCE by the caller before its first use (ticket words). */
int nvf_train_step(const NvfStepArgs* args, void* workspace, size_t workspace_bytes, void* stream);
/* U[0,1) values number first..first+n-1 of noise stream `stream_id` (1 latent, 2 kernel) at step `step`: the exact
 * values the in-kernel generators use (tests: distribution, determinism, replay of a fused step). */
int nvf_rng_uniform(uint64_t seed, uint64_t step, int stream_id, int64_t first, int64_t n, float* out, void* stream);

/* FP32 FFMA throughput micro-benchmark (roofline denominator):
 * runs `iters` dependent-chain FFMA batches on every SM; returns the number of
 * FLOPs executed in flops_out.  The caller times it with CUDA events. */
int nvf_ffma_microbench(int variant, int64_t iters, float* sink /*[>= 148*1024]*/, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NVF_B200_H_ */
