/* ddpm_b200.h -- C-ABI of the B200-native DDPM hot path (libddpm_b200.so).
 *
 * Drop-in boundary for the compute that nereaqing/Polyp-Image-Generator reaches through
 * diffusers/torch on its DDPM path.  The reference has NO native interface of its own (SURVEY.md §2.3:
 * zero .cu/.cpp files); every entry point below therefore cites the *Python call site* whose
 * third-party arithmetic it replaces.  All paths are relative to /root/reference/.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - the caller owns all buffers (incl. workspaces), kernels never allocate and never synchronise
 *   - `stream` is a cudaStream_t passed as void*
 *   - return 0 on success, <0 on error; ddpm_last_error() returns a thread-local message
 *   - activations are NHWC bf16 with an explicit pixel stride `ld` (elements), so channel slices work
 *   - conv/linear weights are bf16 [Cout][tap][Cin] (reduction index contiguous)
 */
#ifndef DDPM_B200_H_
#define DDPM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DDPM_MAX_TAPS 9
#define DDPM_ABI_VERSION 3

const char* ddpm_last_error(void);
int ddpm_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * Scheduler / loss elementwise kernels (fp32, single pass, HBM-bound)
 * ---------------------------------------------------------------------------------------------- */

/* DDPMScheduler.add_noise -- generator_model/train_from_scratch.py:93
 * out[b,i] = sqrt_ac[t[b]] * x0[b,i] + sqrt_1mac[t[b]] * noise[b,i]; each product rounded separately
 * (no FMA contraction) so results are bit-identical to the torch expression.  t: int64[batch]. */
int ddpm_add_noise(const float* x0, const float* noise, const long long* t, const float* sqrt_ac,
                   const float* sqrt_1mac, float* out, int batch, long long per_sample, int num_train_timesteps,
                   void* stream);

/* F.mse_loss(pred, target) forward + backward -- generator_model/train_from_scratch.py:101,103
 * loss_sum += sum((pred-target)^2) (caller zeroes it and divides by n), dpred = (2/n) * (pred-target). */
int ddpm_mse_fwd_bwd(const float* pred, const float* target, float* loss_sum, float* dpred, long long n,
                     void* stream);

/* x *= *scale (device scalar): folds autograd's grad_output (e.g. GradScaler's scale) into dpred. */
int ddpm_scale_by_device_scalar(float* x, const float* scale, long long n, void* stream);

/* clip_grad_norm_(params, max_norm) + AdamW.step() over a FLAT fp32 parameter / gradient / moment arena --
 * generator_model/train_from_scratch.py:106-108, :273.
 *   ddpm_sumsq_f32: out += sum x^2 (caller zeroes out).
 *   ddpm_adamw_flat: scal[0] (step) += 1; clip = min(1, max_norm / (sqrt(*gnorm_sq) + 1e-6)) when gnorm_sq != NULL;
 *     g' = clip*g; p *= 1 - lr*wd; m += (g'-m)(1-b1); v = b2 v + (1-b2) g'^2;
 *     p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps)          (torch.optim.AdamW's op order)
 *   scal: 4 device floats (step, clip, bias corrections), zero-initialised once; lr_dev (device scalar) overrides lr. */
int ddpm_sumsq_f32(const float* x, long long n, float* out, void* stream);
int ddpm_adamw_flat(float* p, const float* g, float* m, float* v, long long n, float* scal, const float* gnorm_sq,
                    float max_norm, const float* lr_dev, float lr, float beta1, float beta2, float eps,
                    float weight_decay, void* stream);

/* DDPMScheduler.step (fixed_small, epsilon, clip_sample) -- inside DDPMPipeline.__call__,
 * generator_model/train_from_scratch.py:51-54.
 *   x0 = clamp((x - sqrt_beta_prod*eps) / sqrt_alpha_prod, -clip, clip)   (clip<=0: no clamp)
 *   prev = c0*x0 + ct*x (+ sigma*z when z != NULL)
 * op order and rounding follow the torch expression.  pred_x0 may be NULL. */
int ddpm_scheduler_step(const float* eps, const float* x, const float* z, float* prev, float* pred_x0, long long n,
                        float sqrt_alpha_prod, float sqrt_beta_prod, float c0, float ct, float sigma, float clip,
                        void* stream);

/* Same step with z ~ N(0,1) drawn in-kernel (Philox4x32-10 + Box-Muller): 12 B/elem instead of 16. */
int ddpm_scheduler_step_philox(const float* eps, const float* x, float* prev, long long n, float sqrt_alpha_prod,
                               float sqrt_beta_prod, float c0, float ct, float sigma, float clip,
                               unsigned long long seed, unsigned long long offset, void* stream);
/* DDIMScheduler.step (epsilon prediction) -- strided sampler on the same UNet:
 *   x0 = clamp((x - sqrt_beta_prod*eps)/sqrt_alpha_prod);  eps' = use_clipped ? (x - sqrt_alpha_prod*x0)/sqrt_beta_prod : eps
 *   prev = sqrt_alpha_prod_prev*x0 + dir_coef*eps' (+ sigma*z when z != NULL);  pred_x0 optional. */
int ddpm_ddim_step(const float* eps, const float* x, const float* z, float* prev, float* pred_x0, long long n,
                   float sqrt_alpha_prod, float sqrt_beta_prod, float sqrt_alpha_prod_prev, float dir_coef, float sigma,
                   float clip, int use_clipped_model_output, void* stream);

/* UniPCMultistepScheduler.step (the sampler the LoRA scripts build at train_with_lora_all_classes.py:314 /
 * train_with_lora_per_class.py:308; diffusers defaults: solver_order 2, epsilon, predict_x0, "bh2", lower_order_final).
 * Scalars are computed on the host in fp32 in diffusers' op order; the two kernels keep torch's elementwise op order:
 *   ddpm_unipc_x0:     out = (x - sigma_t * eps) / alpha_t                           (convert_model_output)
 *   ddpm_unipc_update: out = (cx * x - cm * m0) - cb * res,
 *                      res = rho0 * ((m1 - m0) / rk)  [m1 != NULL]   (+)   rho_t * (mt - m0)  [mt != NULL]
 *                      (predictor: m1 = previous x0 prediction at order 2; corrector: mt = this step's x0 prediction) */
int ddpm_unipc_x0(const float* eps, const float* x, float* out, long long n, float sigma_t, float alpha_t, void* stream);
int ddpm_unipc_update(const float* x, const float* m0, const float* m1, const float* mt, float* out, long long n,
                      float cx, float cm, float cb, float rk, float rho0, float rho_t, void* stream);

/* DDPMPipeline post-processing: (x/2+0.5).clamp(0,1) -> NHWC uint8 via round(x*255). x: NCHW fp32. */
int ddpm_to_uint8_nhwc(const float* x, unsigned char* out, int n, int c, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------------------
 * UNet2DModel forward / backward building blocks -- generator_model/train_from_scratch.py:100,103
 * (model built at generator_model/PolypGeneratorModel.py:25-48)
 * ---------------------------------------------------------------------------------------------- */

/* Implicit-GEMM conv / linear on tcgen05 (fprop and dgrad; nn.Conv2d / nn.Linear forward and input grad).
 *   out[pix, co] = bias[co] + temb[n, co] + res[pix, co] + sum_{tap, ci} X[pix + (dn,dh,dw)[tap], ci] * W[co][wk[tap] + ci]
 * X may be the channel concat of two tensors (x0 | x1).  Out-of-range pixels read as zero (= padding). */
typedef struct ddpm_conv_args {
  const void* x0; int c0; long long ld0;     /* source 0: bf16 NHWC, c0 channels (multiple of 64) */
  const void* x1; int c1; long long ld1;     /* optional source 1 (NULL / 0) */
  int n, h, w;                               /* output pixel grid */
  int src_n;                                 /* batch extent of the sources (0 -> n; 4n for space-to-depth) */
  int ntaps;
  int tap_dn[DDPM_MAX_TAPS], tap_dh[DDPM_MAX_TAPS], tap_dw[DDPM_MAX_TAPS], tap_wk[DDPM_MAX_TAPS];
  const void* wgt; int cout; long long ldw; long long k_total; /* bf16 [cout][ldw]; k_total 0 -> ldw */
  void* out; float* out_f32; long long ldo;  /* bf16 output, or fp32 output when out_f32 != NULL */
  const float* bias;                         /* fp32 [cout] or NULL */
  const float* temb; int ld_temb;            /* fp32 [n][ld_temb] or NULL: ResnetBlock2D time-embedding add */
  const void* res; long long ldr;            /* bf16 NHWC residual or NULL */
  /* Optional GroupNorm-backward fusion (gn_sums != NULL), for the dgrad GEMM whose result is the gradient w.r.t.
   * y = act(GroupNorm(x)):  out = dz = result * act'(x*ka + kb) and gn_sums[n][co][0..1] += (sum dz, sum dz*x) over
   * the sample's pixels.  x = (gn_x0 | gn_x1) bf16 NHWC with gn_c0 + (cout - gn_c0) channels; gn_coef is the
   * per-(sample, channel) affine table written by ddpm_gn_fwd.  Requires h*w >= 128 or h*w % 32 == 0; gn_sums must
   * be zero-filled. */
  const void* gn_x0; long long gn_ld0; int gn_c0;
  const void* gn_x1; long long gn_ld1;
  const float* gn_coef;
  int gn_silu;
  float* gn_sums;
  /* Optional statistics of the GroupNorm that CONSUMES this output: out_csum[n][co / 4][0..1] += (sum out, sum out^2)
   * over the sample's pixels and the 4 channels of a granule, of the bf16 values actually stored (zero-filled by the
   * caller; same h*w rule as gn_sums; excludes gn_sums / out_f32; GroupNorm groups are multiples of 4 channels).
   * ddpm_gn_stats_from_csum + ddpm_gn_apply then replace the two-phase ddpm_gn_fwd. */
  float* out_csum;
  /* Optional split-K workspace (fp32, caller-owned): low-resolution layers whose tile grid covers less than half the
   * SMs split the reduction over blockIdx.z, accumulate fp32 partial sums here and finish in a second pass.
   * ddpm_conv_gemm_workspace_elems() says how many elements this problem would use (0 = it does not split). */
  float* splitk_ws; long long splitk_ws_elems;
  /* fp32-faithful inference (split-bf16, see "fp32-faithful mode" below): 1 = `out` and `res` are SPLIT tensors --
   * per pixel cout "hi" channels followed by cout "lo" channels (ldo / ldr >= 2*cout), value = hi + lo; the fp32
   * accumulator is stored as hi = bf16(v), lo = bf16(v - hi).  Excludes gn_sums / out_csum.  The A operand of such a
   * conv is a split tensor passed as x0 = [hi | lo] (c0 = 2*cin) and x1 = its hi half (c1 = cin), with the weight
   * operand [W_hi | W_hi | W_lo] per tap: (A_hi + A_lo) W_hi + A_hi W_lo, fp32 accumulation -- three bf16 MMAs per
   * product instead of one, ~2^-17 relative error instead of 2^-9. */
  int split_io;
} ddpm_conv_args;
int ddpm_conv_gemm(const ddpm_conv_args* args, void* stream);
/* Number of column strips per image row the halo-resident 3x3 kernel (conv_halo.cu) uses at image width w, 0 when that
 * width runs on the generic implicit-GEMM kernel.  The host uses it to decide where the GroupNorm statistics / backward
 * fusions (out_csum, gn_sums) are free (their reductions hide behind the next tile's mainloop only in that kernel). */
int ddpm_conv_halo_strips(int w);
long long ddpm_conv_gemm_workspace_elems(const ddpm_conv_args* args);

/* Conv / linear weight gradient on tcgen05:
 *   dw[co][wk[tap] + ci] (+)= sum_pix dY[pix, co] * X[pix + (dn,dh,dw)[tap], ci]      (fp32, split-K) */
typedef struct ddpm_wgrad_args {
  const void* dy; long long ldy; int cout;   /* bf16 NHWC grad of the conv output (cout multiple of 64) */
  const void* x0; int c0; long long ld0;
  const void* x1; int c1; long long ld1;
  int n, h, w; int src_n;
  int ntaps;
  int tap_dn[DDPM_MAX_TAPS], tap_dh[DDPM_MAX_TAPS], tap_dw[DDPM_MAX_TAPS], tap_wk[DDPM_MAX_TAPS];
  float* dw; long long ldw;                  /* fp32 [cout][ldw] */
  int accumulate;                            /* 1: add into dw, 0: dw must be zero-filled or splits==1 */
  int splits;                                /* 0 = auto */
  float* dbias;                              /* optional fp32 [cout]: += sum_pix dY[pix, co] (bias gradient) */
} ddpm_wgrad_args;
int ddpm_conv_wgrad(const ddpm_wgrad_args* args, void* stream);

/* fp32 master weight [cout][taps][cin] -> bf16 fprop copy (same layout) and, if wd != NULL, the dgrad copy
 * wd[ci][taps-1-tap][co] (taps flipped, in/out transposed). */
int ddpm_prep_weight(const float* w, void* wf, long long ldwf, void* wd, long long ldwd, int cout, int taps, int cin,
                     void* stream);

/* The same for ALL layers of the model in one launch: a device-resident table of descriptors (pointers into the
 * fp32 parameter arena and the bf16 operand arena), tile_begin = running count of DDPM_PREP_TILE x DDPM_PREP_TILE
 * tiles (ascending). */
#define DDPM_PREP_TILE 64
typedef struct ddpm_prep_desc {
  const float* w; void* wf; long long ldwf; void* wd; long long ldwd;
  int cout, taps, cin;
  int tile_begin, tiles_x, tiles_y;   /* tiles_x = ceil(cin/TILE), tiles_y = ceil(cout/TILE); taps tiles in z */
} ddpm_prep_desc;
int ddpm_prep_weights_batched(const ddpm_prep_desc* table_dev, int n_entries, int total_tiles, int with_d,
                              void* stream);

/* The 3-channel boundary convs on tensor cores (conv_in forward / conv_out dgrad: K = 27 padded to one 64-wide
 * k-block; conv_out forward: N = 3 padded to 32 output columns):
 *   ddpm_im2col3: patches[pix][tap*cin + k] = src[n, k, h+tap/3-1, w+tap%3-1] (bf16 [n*h*w][64], zero padded);
 *                 chan_sum[k] += sum of src[:, k] when non-NULL (conv_out bias gradient).
 *   ddpm_nhwc_to_nchw_f32: out[n][k][h][w] = src[pix][k], k < cout <= 4 (src fp32 NHWC, pixel stride ld). */
int ddpm_im2col3(const float* src, void* patches, int n, int h, int w, int cin, float* chan_sum, void* stream);
int ddpm_nhwc_to_nchw_f32(const float* src, long long ld, float* out, int n, int h, int w, int cout, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32-faithful mode (reference: the sampling loop of train_from_scratch.py:39-66,121-125 and the LoRA trainers run
 * WITHOUT autocast, i.e. in fp32).  Activations travel as "split-bf16" tensors: per pixel C hi channels then C lo
 * channels, value = hi + lo (16 mantissa bits).  The tensor cores stay the compute engine (ddpm_conv_args.split_io);
 * the *_split entry points below are the same ops as their bf16 namesakes on split tensors, with exact (non-approx)
 * transcendental arithmetic.  c0 / c1 / ld* count LOGICAL channels / elements of the split row (ld >= 2*c).
 * ---------------------------------------------------------------------------------------------- */
int ddpm_gn_stats_split(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                        int groups, float* stats, void* stream);
int ddpm_gn_apply_split(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                        int groups, const float* stats, float eps, const float* gamma, const float* beta, int silu,
                        void* y, long long ldy, void* stream);
/* qkv: split rows [q k v hi (3*heads*d) | q k v lo]; o: split rows [hi (heads*d) | lo]; narrow heads (d = 8..64). */
int ddpm_attn_fwd_split(const void* qkv, long long ldqkv, void* o, long long ldo, int b, int t, int heads, int d,
                        float scale, void* stream);
/* patches[pix][128] bf16: columns [0, 9cin) hi, [32, 32 + 9cin) lo, [64, 64 + 9cin) hi again (the three MMA terms of
 * conv_in as ONE two-k-block GEMM against [W_hi | W_hi | W_lo] laid out the same way), zero elsewhere; cin <= 3. */
int ddpm_im2col3_split(const float* src, void* patches, int n, int h, int w, int cin, void* stream);

/* GroupNorm statistics: stats[n][g] = (sum, sumsq) over the (possibly concatenated) channels of group g. */
int ddpm_gn_stats(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                  int groups, float* stats, void* stream);
/* y = act(GroupNorm(x)) with act = SiLU (silu=1) or identity; y bf16 NHWC contiguous over c0+c1 channels.
 * coef (may be NULL): per-(sample, channel) affine table [n][(c0+c1)/2][4] for the gn_sums fusion of ddpm_conv_gemm. */
int ddpm_gn_apply(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                  int groups, const float* stats, float eps, const float* gamma, const float* beta, int silu,
                  void* y, long long ldy, float* coef, void* stream);
/* stats (as ddpm_gn_stats) from the per-(sample, 4-channel granule) moments accumulated by the producing conv epilogues
 * (ddpm_conv_args.out_csum), for an input that may be the concat of two tensors: with it the GroupNorm forward is ONE
 * streaming pass (ddpm_gn_apply, 4 B/elem) instead of the two-phase ddpm_gn_fwd. */
int ddpm_gn_stats_from_csum(const float* csum0, int c0, const float* csum1, int c1, int n, int groups, float* stats,
                            void* stream);
/* Fused GroupNorm forward: stats (as ddpm_gn_stats) AND y = act(GroupNorm(x)) in one persistent, cooperatively
 * launched kernel whose second phase re-reads x from L2.  ws: n ints (team barrier counters). */
int ddpm_gn_fwd(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                int groups, float eps, const float* gamma, const float* beta, int silu, float* stats, void* y,
                long long ldy, float* coef, int* ws, void* stream);
/* coef (optional, fp32 [n][(c0+c1)/2][4]): (ka0, ka1, kb0, kb1) per channel pair with GN(x) = x*ka + kb; consumed by
 * the GroupNorm-backward fusion of ddpm_conv_gemm. */
/* GroupNorm(+SiLU) backward.  ws: workspace of n*(c0+c1)*2 floats followed by n ints (team barrier counters).
 *   dgamma/dbeta are accumulated (+=) when non-NULL.  dx = GN'(dy) + add0 + add1, written split over dx0|dx1. */
int ddpm_gn_bwd(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                int groups, const float* stats, float eps, const float* gamma, const float* beta, int silu,
                const void* dy, long long lddy, const void* add0, long long ldadd0, const void* add1,
                long long ldadd1, void* dx0, long long lddx0, void* dx1, long long lddx1, float* dgamma,
                float* dbeta, float* ws, void* stream);

/* Second half of the GroupNorm backward when the first half ran in a conv epilogue (gn_sums of ddpm_conv_gemm):
 *   dx = dz*rstd*gamma - xhat*rstd*mean_g(gamma*dz*xhat) - rstd*mean_g(gamma*dz) + add0 + add1,  split over dx0|dx1;
 * dgamma[c] += sum_n (rstd*(S2 - mean*S1)), dbeta[c] += sum_n S1 when non-NULL.  One streaming pass, no
 * transcendentals.  sums: [n][c0+c1][2] as accumulated by the conv epilogue. */
int ddpm_gn_bwd_apply(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                      int groups, const float* stats, float eps, const float* gamma, const void* dz, long long lddz,
                      const float* sums, const void* add0, long long ldadd0, const void* add1, long long ldadd1,
                      void* dx0, long long lddx0, void* dx1, long long lddx1, float* dgamma, float* dbeta,
                      float* out_nc, long long ld_nc, float* out_c, void* stream);
/* dgamma / dbeta reductions alone (call the two functions above with dgamma = dbeta = NULL first): lets the caller
 * issue them on another stream.  stats == NULL: sums = the centred sums ddpm_gn_bwd left at the start of its ws;
 * otherwise sums = raw moments from the conv-epilogue fusion (as passed to ddpm_gn_bwd_apply). */
int ddpm_gn_bwd_dparams(const float* sums, const float* stats, int n, int c, int groups, int hw, float eps,
                        float* dgamma, float* dbeta, void* stream);
/* out_nc[n][c] += sum_pix dx, out_c[c] += sum_{n,pix} dx when non-NULL (time-embedding / conv-bias gradients of the
 * layer that produced x, fused here instead of a separate pass over dx). */

/* Self-attention core on fused qkv [b*t][3*heads*d] bf16 (AttnProcessor2_0's scaled_dot_product_attention). */
int ddpm_attn_fwd(const void* qkv, long long ldqkv, void* o, long long ldo, float* lse, int b, int t, int heads,
                  int d, float scale, void* stream);
int ddpm_attn_bwd(const void* qkv, long long ldqkv, const void* o, long long ldo, const void* d_o, long long lddo,
                  const float* lse, void* dqkv, long long lddqkv, int b, int t, int heads, int d, float scale,
                  void* stream);

/* Wide-head attention (head_dim > 64, e.g. the single 512-wide head of the google/ddpm-celebahq-256 architecture that
 * train_with_lora_all_classes.py:316-330 fine-tunes): the core is composed from a batched tcgen05 GEMM and row softmax
 * kernels (bgemm.cu) over the same fused qkv buffer.
 *   ddpm_bgemm: C[z] = alpha * A[z] (m x k) * B[z] (k x n) for z = (batch, head); bf16 operands, fp32 accumulation,
 *     C bf16 or fp32 (c_f32).  x_mn = 0: reduction index contiguous (A(i,kk) = a[i*lda+kk], B(kk,j) = b[j*ldb+kk]);
 *     x_mn = 1: output index contiguous (A(i,kk) = a[kk*lda+i], B(kk,j) = b[kk*ldb+j]).  Strides in elements,
 *     multiples of 8; *_head / *_batch are the element offsets between heads / samples.
 *   ddpm_softmax_rows: P = softmax(S) per row (S fp32 with the d^-1/2 scale already applied, P bf16).
 *   ddpm_softmax_rows_bwd: dS = scale * P o (dP - rowsum(P o dP)) (dS bf16, leading dimension ldp). */
int ddpm_bgemm(const void* a, long long lda, long long a_head, long long a_batch, int a_mn, const void* b,
               long long ldb, long long b_head, long long b_batch, int b_mn, void* c, long long ldc, long long c_head,
               long long c_batch, int c_f32, int m, int n, int k, int heads, int batch, float alpha, void* stream);
int ddpm_softmax_rows(const float* s, long long lds, void* p, long long ldp, long long rows, int t, void* stream);
int ddpm_softmax_rows_bwd(const void* p, long long ldp, const float* dp, long long lddp, void* ds, long long rows,
                          int t, float scale, void* stream);

/* The same attention core as ONE fused tcgen05 kernel (attn_wide.cu) where it fits: t <= 256 tokens and head_dim a
 * multiple of 128 (the 16x16 / 8x8 blocks of the celebahq architecture at 256^2: t = 256 / 64, one head of 512).
 * S = Q K^T accumulates in TMEM, the softmax runs from TMEM into shared memory, O = P V follows from there; S and P
 * never travel through L2 / HBM.  qkv as in ddpm_attn_fwd; o bf16 [b*t][heads*d] (row stride ldo);
 * probs: NULL (inference) or bf16 [b][heads][t][ldp] receiving softmax(scale * Q K^T) for the backward pass
 * (ddpm_bgemm / ddpm_softmax_rows_bwd).  ddpm_attn_wide_supported -> 1 when (t, heads, d) fits the fused kernel. */
int ddpm_attn_wide_supported(int t, int heads, int d);
int ddpm_attn_wide_fwd(const void* qkv, long long ldqkv, void* o, long long ldo, void* probs, long long ldp, int b,
                       int t, int heads, int d, float scale, void* stream);

/* Timesteps(128) sinusoid -> fp32 [b][dim]; t int64[b] (device); freqs fp32[dim/2] (device) =
 * exp(-ln(10000) * j / (dim/2 - freq_shift)) tabulated by the host. */
int ddpm_timestep_embedding(const long long* t, const float* freqs, float* out, int b, int dim,
                            int flip_sin_to_cos, void* stream);
/* Small fp32 linears of the time-embedding path: y[m][n] = bias[n] + sum_k act(x[m][k]) * w[n][k]. */
int ddpm_linear_f32(const float* x, const float* w, const float* bias, float* y, int m, int n, int k, int silu_in,
                    void* stream);
/* dw[n][k] += sum_m dy[m][n]*act(x[m][k]); db[n] += sum_m dy[m][n]  (db may be NULL) */
int ddpm_linear_f32_wgrad(const float* x, const float* dy, float* dw, float* db, int m, int n, int k, int silu_in,
                          void* stream);
/* dx[m][k] (+)= act'(x[m][k]) * sum_n dy[m][n]*w[n][k] */
int ddpm_linear_f32_dgrad(const float* dy, const float* w, const float* x, float* dx, int m, int n, int k,
                          int silu_in, int accumulate, void* stream);

/* out_nc[n][c] (=) sum_hw x[n,hw,c] and/or out_c[c] += sum_{n,hw} x[n,hw,c]; x bf16 NHWC. Either may be NULL. */
int ddpm_reduce_hw(const void* x, long long ld, int n, int hw, int c, float* out_nc, long long ld_nc, float* out_c,
                   void* stream);

/* LoRA adapter-branch dropout (peft lora_dropout; generator_model/train_with_lora_all_classes.py:316-322):
 * out = (add ? add : 0) + x * keep / (1-p), keep regenerated from (seed, offset) by Philox -- no mask is stored.
 * Forward: add = NULL.  Backward: x = grad of the dropped tensor, add = the other gradient contribution. bf16.
 * tick (may be NULL): device-resident step counter added to the upper counter word, so CUDA-graph replays of a
 * training step draw fresh masks. */
int ddpm_dropout(const void* x, const void* add, void* out, long long n, float p, unsigned long long seed,
                 unsigned long long offset, const unsigned long long* tick, void* stream);

/* Input transform of generator_model/PolypDiffusionDataset.py:52-59 on the device:
 * Resize((S,S)) [Pillow's two-pass antialiased bilinear resampler, 8-bit fixed point] -> hflip -> ToTensor ->
 * Normalize([0.5],[0.5]).  bounds int32 [out][2] = (first source index, tap count), coeffs int32 [out][ksize] =
 * Pillow's normalize_coeffs_8bpc tables (22 fractional bits), built by the host.  c = 1 or 3 interleaved channels.
 *   ddpm_resize_h_u8:        dst u8 [rows][out_w][c] from src u8 [rows][w][c]                (horizontal pass)
 *   ddpm_resize_v_normalize: out f32 [b][c][out_h][w] from src u8 [b][h][w][c]; flip u8 [b] or NULL (vertical pass,
 *                            flip, /255, (x-0.5)/0.5).  Pass identity tables to skip a resampling pass. */
int ddpm_resize_h_u8(const unsigned char* src, unsigned char* dst, long long rows, int w, int c, int out_w,
                     const int* bounds, const int* coeffs, int ksize, void* stream);
int ddpm_resize_v_normalize(const unsigned char* src, float* out, int b, int h, int w, int c, int out_h,
                            const int* bounds, const int* coeffs, int ksize, const unsigned char* flip, void* stream);

/* Layout helpers (bf16 NHWC, contiguous outputs). */
int ddpm_space_to_depth(const void* x, long long ldx, void* out, int n, int h, int w, int c, int pad_lo,
                        void* stream);                      /* out[(ph*2+pw)*n + b][h/2][w/2][c] */
int ddpm_zero_insert2x(const void* dy, long long ldy, void* out, int n, int ho, int wo, int c, int h, int w,
                       void* stream);                       /* out[b][2i][2j] = dy[b][i][j], else 0 */
int ddpm_upsample2x(const void* x, long long ldx, void* out, int n, int h, int w, int c, void* stream);
int ddpm_sumpool2x(const void* dy, long long ldy, const void* add, long long ldadd, void* out, int n, int h, int w,
                   int c, void* stream);                    /* out[b][i][j] = sum 2x2 dy + add */

#ifdef __cplusplus
}
#endif
#endif /* DDPM_B200_H_ */
