/*
 * semb200.h -- C ABI of libsemb200.so: hand-written sm_100a kernels for the conv-stack
 * hot path of BAMresearch/automatic-sem-image-segmentation (Release 1.2.0).
 *
 * The reference has no FFI of its own: every FLOP of its hot path is a Keras layer call
 * that lands in torch.nn.functional on the torch backend.  The
 * entry points below are therefore cut at exactly those library-call sites; each one
 * names the reference call site(s) (file:line under Releases/Version 1.2.0/) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch types.  All tensor pointers are DEVICE
 *     pointers owned by the caller; the library never allocates or frees tensor memory.
 *   - activations are NHWC with a channel pitch (elements) and a channel offset, so a
 *     tensor can be a channel slice of a wider buffer (skip-concat without a copy).
 *   - `dtype` selects the activation storage type: SEMB_F32 (parity mode) or SEMB_BF16
 *     (throughput mode).  Weights, statistics, gradients of weights are always fp32.
 *   - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous.
 *   - every function returns SEMB_OK (0) or a negative SEMB_E* code; the message is
 *     available from semb_last_error() (thread-local).
 */
#ifndef SEMB200_H
#define SEMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEMB_VERSION 100

enum { SEMB_OK = 0, SEMB_ESHAPE = -1, SEMB_EALIGN = -2, SEMB_EARCH = -3, SEMB_EWORKSPACE = -4, SEMB_ECUDA = -5 };
enum { SEMB_F32 = 0, SEMB_BF16 = 1 };
enum { SEMB_PAD_ZERO = 0, SEMB_PAD_REFLECT = 1 };
enum { SEMB_ACT_NONE = 0, SEMB_ACT_RELU = 1, SEMB_ACT_LEAKY = 2, SEMB_ACT_SIGMOID = 3, SEMB_ACT_TANH = 4 };
/* how an operand of semb_affine_act_* is transformed before the add */
enum { SEMB_AFF_NONE = 0,      /* identity (plain residual operand)                                */
       SEMB_AFF_PLAIN = 1,     /* x*scale+shift, scale/shift are constants (inference BN)          */
       SEMB_AFF_BATCH = 2 };   /* x*scale+shift where scale/shift come from batch statistics of x  */

/* A view of an NHWC activation tensor (possibly a channel slice of a wider buffer). */
typedef struct {
    void*   ptr;     /* base of the underlying buffer (element type = dtype)            */
    int32_t C;       /* channels of this view (PHYSICAL, i.e. padded: C % 8 == 0)        */
    int32_t pitch;   /* elements between consecutive pixels of the underlying buffer    */
    int32_t coff;    /* first channel of the view inside the buffer (coff % 8 == 0)     */
} semb_tensor;
/* Channel padding: every tensor the kernels see has its channel count padded to a multiple of 8
 * (16 bytes of bf16).  The host keeps the map logical -> physical channel (multi_res_block's
 * concat of 8|17|26 channels is stored as 8|24|32); padded lanes hold exact zeros, the matching
 * weight rows/columns and BN parameters are zero, so they stay zero through forward, backward
 * and Adam.  All C / Cin / Cout below are physical. */

/* Geometry of one Conv2D (forward direction).  Keras call sites:
 * UNet_Segmentation.py:421 (conv2d_bn), CycleGAN.py:327,333,340,372,393,429,448. */
typedef struct {
    int32_t N, H, W;          /* input  batch / height / width                          */
    int32_t OH, OW;           /* output height / width                                  */
    int32_t Cin, Cout;
    int32_t R, S;             /* kernel height / width                                  */
    int32_t stride;           /* 1 or 2                                                 */
    int32_t pad_t, pad_l;     /* leading padding; trailing is implied by OH/OW          */
    int32_t pad_mode;         /* SEMB_PAD_ZERO | SEMB_PAD_REFLECT                       */
    int32_t dtype;            /* activation storage                                     */
} semb_conv_geom;

int         semb_version(void);
const char* semb_last_error(void);
/* number of kernels launched by this library in this process since load (bench.py's gpu_launches) */
int64_t     semb_launch_count(void);
/* 1 when the current device is compute capability 10.x */
int         semb_device_ok(void);

/* ---- convolutions ------------------------------------------------------------------------- */

/* y[n,oy,ox,co] (+)= bias[co] + sum_{r,s,ci} x[n, oy*stride-pad_t+r, ox*stride-pad_l+s, ci] * w[r,s,ci,co]
 * w: fp32 HWIO (Keras Conv2D kernel layout).  bias may be NULL.
 * stats (may be NULL): FP64 accumulators (strides in doubles), sum at stats[n*stats_nstride + c], sum of
 * squares at stats[n*stats_nstride + stats_cstride + c]; nstride = 0 gives BatchNorm (per-channel)
 * moments, nstride = 2*cstride gives GroupNormalization(groups=-1) (per-sample) moments.  Partials are
 * reduced in a fixed order inside a CTA and combined across CTAs with fp64 atomics, so the forward pass is
 * reproducible run to run (fp32 atomics were observed to flip ReLU/max-pool decisions in backward).
 * Replaces F.conv2d under keras.layers.Conv2D + the keras.ops.moments pass of the following
 * BatchNormalization / GroupNormalization (UNet_Segmentation.py:421-422, CycleGAN.py:327-329). */
int semb_conv2d_fwd(const semb_conv_geom* g, const semb_tensor* x, const float* w, const float* bias,
                    const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                    int32_t accumulate, void* stream);

/* dx[n,iy,ix,ci] (+)= bias[ci] + sum_{r,s,co : iy = oy*stride-pad_t+r} dy[n,oy,ox,co] * w[r,s,ci,co]
 * Zero padding only (reflect-padded convs are differentiated on the padded domain and folded
 * with semb_pad_crop).  This is (a) the data gradient of semb_conv2d_fwd (autograd of F.conv2d) and
 * (b) the FORWARD of keras.layers.Conv2DTranspose whose Keras kernel (kh,kw,Cout,Cin) is read as
 * HWIO with I=Cout_T, O=Cin_T (UNet_Segmentation.py:542-551, CycleGAN.py:353): g describes the
 * equivalent strided conv that maps the transposed-conv OUTPUT back to its INPUT.  bias/stats as above
 * (indexed by ci). */
int semb_conv2d_dgrad(const semb_conv_geom* g, const semb_tensor* dy, const float* w, const float* bias,
                      const semb_tensor* dx, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                      int32_t accumulate, void* stream);

/* dw[r,s,ci,co] += sum_{n,oy,ox} x[n, oy*stride-pad_t+r, ox*stride-pad_l+s, ci] * dy[n,oy,ox,co]   (fp32 HWIO)
 * dbias[co] += sum dy  (dbias may be NULL).  Accumulates: the caller zeroes the flat gradient buffer.
 * Weight gradient of F.conv2d / F.conv_transpose2d (loss.backward()). */
int semb_conv2d_wgrad(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy,
                      float* dw, float* dbias, void* stream);

/* ---- tensor-core (tcgen05 / TMEM) convolutions, bf16 storage ------------------------------ */

/* Packs fp32 HWIO weights (8-padded channels) into the bf16 UMMA shared-memory image read by
 * semb_conv2d_fwd_tc: [n-chunk][k-chunk][tap][k/8][n][8].  flip=1 packs the spatially mirrored,
 * channel-transposed kernel, so that the same implicit-GEMM kernel computes the stride-1 data gradient
 * (call semb_conv2d_fwd_tc with x:=dy, y:=dx, Cin/Cout swapped, pad := k-1-pad).
 * Returns the packed size in bytes (also when dst==NULL, for sizing) or a negative SEMB_E* code. */
int64_t semb_pack_weights_tc(const float* w, int32_t R, int32_t S, int32_t Cin, int32_t Cout,
                             int32_t flip, void* dst, void* stream);

/* One launch for all weight images of a network.  The caller keeps a HOST array of njobs opaque job records
 * (semb_pack_batch_job_size() bytes each), fills record `index` with semb_pack_batch_prepare (which returns the first
 * block of the next job, i.e. the running block count, or a negative SEMB_E* code), copies the array to the device once
 * and then calls semb_pack_weights_tc_batch(device_jobs, njobs, total_blocks) after every weight update. */
int64_t semb_pack_batch_job_size(void);
int semb_pack_batch_prepare(int32_t index, const float* w, void* dst, int32_t R, int32_t S, int32_t Cin, int32_t Cout,
                            int32_t flip, int32_t block0, void* host_jobs);
int semb_pack_weights_tc_batch(const void* device_jobs, int32_t njobs, int32_t total_blocks, void* stream);

/* Same contract as semb_conv2d_fwd for stride 1, R=S in {1,3}, bf16 storage: im2col-free implicit GEMM on
 * tcgen05.mma (M = 128 output pixels per CTA, N = Cout, K = taps x Cin) with the fp32 accumulator in TMEM.
 * The A operand is the NHWC halo tile staged once in shared memory by the TMA unit; the nine taps are shifted UMMA
 * descriptors over that tile.  Zero padding only (the TMA's out-of-bounds fill): the reflect-padded residual convs of
 * CycleGAN.py:326-333 run as pad 0 ('valid') over an input materialised with semb_pad_crop.
 * UNet_Segmentation.py:421,465-468,490-499; CycleGAN.py:327,333. */
int semb_conv2d_fwd_tc(const semb_conv_geom* g, const semb_tensor* x, const void* w_packed, const float* bias,
                       const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                       int32_t accumulate, void* stream);

/* Conv2DTranspose(2x2, stride 2, bias) forward (UNet_Segmentation.py:542-551) in ONE launch: the 1x1 tensor-core conv with
 * g->Cout = 4*C virtual outputs (channel (2r+s)*C + c, packed like any 1x1 kernel) whose epilogue stores pixel (y, x),
 * quadrant (r, s) at pixel (2y+r, 2x+s) of `up` = the (N, UH, UW, C) destination view (a channel slice of the skip-concat
 * buffer; UH in {2H-1, 2H}) and adds the C-long bias -- the depth-to-space pass semb_pixel_shuffle2 and the 4C-channel
 * intermediate are gone from the forward pass.  No moments, no accumulate. */
int semb_conv2d_fwd_tc_d2s(const semb_conv_geom* g, const semb_tensor* x, const void* w_packed, const float* bias,
                           const semb_tensor* up, int32_t UH, int32_t UW, void* stream);

/* Weight gradient of the same layers on tcgen05: D[co][ci] per tap, K = output pixels (split over CTAs), both
 * operands read as MN-major UMMA operands from the NHWC channel-group planes; accumulates (+=) into the fp32
 * HWIO gradient with coalesced reductions.  Same contract as semb_conv2d_wgrad without dbias. */
int semb_conv2d_wgrad_tc(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw, void* stream);

/* Same weight gradient with caller-owned scratch: for 3x3 layers with >= 128 input and output channels (the CycleGAN
 * residual / sampling convs, CycleGAN.py:323-358) x and dy are first re-laid out as planar [N][C/8][H][W][8] copies in
 * `workspace`, whose TMA boxes have 128-byte rows instead of one 16-byte row per pixel and channel group.
 * semb_conv2d_wgrad_tc_workspace returns the bytes needed (0: no planar path for this geometry, call the plain entry). */
int64_t semb_conv2d_wgrad_tc_workspace(const semb_conv_geom* g);
int semb_conv2d_wgrad_tc_ws(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw, void* workspace,
                            int64_t workspace_bytes, void* stream);

/* ---- stride-2 convolutions on the stride-1 tensor-core kernels (space-to-depth) ------------------------------
 * A k x k (k = 3, 4; also the 5x5 'same' convs of the WGAN critic, WassersteinGAN.py:571-614, with pad_t = pad_l = 1, which fill the
 * whole 3x3) stride-2 Conv2D / Conv2DTranspose (CycleGAN.py:339-358, 425-451) is run as a 3x3-embedded 2x2
 * stride-1 conv over the space-to-depth image (H/2, W/2, 4C): semb_pixel_shuffle2x moves activations between the two
 * domains (DH x DW is the true size of the full-resolution tensor, 2H-1 <= DH <= 2H: a missing last row / column reads
 * as zero and is not written; acc = 1 accumulates in dir 0), semb_s2d_weights builds the virtual fp32 HWIO kernel
 * w3 (3,3,4*Cin,Cout) from the Keras kernel (dir 0) or adds the gradient of w3 into the Keras gradient (dir 1), and
 * semb_fold_stats4 turns the moments of a 4C-channel space-to-depth tensor into those of its C-channel image. */
int semb_pixel_shuffle2x(const semb_tensor* src, const semb_tensor* dst, int32_t N, int32_t H, int32_t W, int32_t DH,
                         int32_t DW, const float* bias, int32_t dir, int32_t acc, int32_t dtype, void* stream);
int semb_s2d_weights(float* w, int32_t k, int32_t pad_t, int32_t pad_l, int32_t Cin, int32_t Cout, float* w3, int32_t dir,
                     void* stream);
/* res_path unit (UNet_Segmentation.py:490-499): Conv2D 3x3 (wa: 3,3,Cin,Ca) and the 1x1 shortcut Conv2D (ws: 1,1,Cin,Cs) of the
 * same input as ONE 3x3 conv with Ca+Cs outputs.  dir 0: w3 (3,3,Cin,Ca+Cs) <- [wa | ws on the centre tap, 0 elsewhere];
 * dir 1: dwa += dw3[..., :Ca], dws += dw3[1,1,:,Ca:] (gradient of the virtual kernel folded into the Keras-layout gradients). */
int semb_merge_weights(float* wa, float* ws, int32_t Cin, int32_t Ca, int32_t Cs, float* w3, int32_t dir, void* stream);
int semb_fold_stats4(const void* temp, void* stats, int32_t groups, int32_t C, int32_t stats_nstride, int32_t stats_cstride,
                     void* stream);

/* ---- 7x7 single-channel convolutions as 1x1 tensor-core convolutions (CycleGAN.py:372, 393: generator stem / head) ----
 * "big" is the reflect-padded domain (N, H+k-1, W+k-1, .), "small" the output domain (N, H, W, .), T8 = pad8(k*k):
 *   mode 0  small = patches (T8 ch):  patches[p][t] = big[p + (r,s)][0]                (stem forward)
 *   mode 1  dbig[u][0] = sum_t dpatches[u - (r,s)][t]                                   (stem data gradient)
 *   mode 2  small[p][0] = bias[0] + sum_t big[p + (r,s)][t],  big = per-tap 1x1 conv    (head forward)
 *   mode 3  dbig[u][t] = dsmall[u - (r,s)][0]                                           (head backward)
 * semb_tapfold_weights maps the Keras kernel to the virtual 1x1 kernel (kind 0: w1[t][co] = w[t][0][co]; kind 1:
 * w1[ci][t] = w[t][ci][0]; dir 0 writes w1, dir 1 adds the gradient of w1 into the Keras-layout gradient). */
int semb_tap_patch(const semb_tensor* small, const semb_tensor* big, int32_t N, int32_t H, int32_t W, int32_t k,
                   const float* bias, int32_t mode, int32_t dtype, void* stream);
int semb_tapfold_weights(float* w, int32_t k, int32_t Cin, int32_t Cout, float* w1, int32_t kind, int32_t dir, void* stream);

/* ---- normalisation + activation (fused elementwise) --------------------------------------- */

/* From moments to the affine that BatchNormalization / GroupNormalization applies.
 * stats: the FP64 moment accumulators written by the conv / affine kernels.
 * For i in [0, groups*C): mean = sum/count, var = sumsq/count - mean^2 (Keras ops.moments),
 * invstd = rsqrt(var+eps), scale = gamma[c]*invstd (gamma NULL -> 1), shift = beta[c]-mean*scale.
 * Writes scale, shift, mean, invstd (each groups x cstride, fp32).  If moving_mean != NULL
 * (BatchNorm training): moving = moving*momentum + batch*(1-momentum) with the biased variance.
 * UNet_Segmentation.py:422,470,473,494,502; CycleGAN.py:329,335,342,355,374. */
int semb_norm_finalize(const void* stats, int32_t groups, int32_t C, int32_t cstride, int32_t stats_nstride,
                       float count, float eps, const float* gamma, const float* beta,
                       float* scale, float* shift, float* mean, float* invstd,
                       float* moving_mean, float* moving_var, float momentum, void* stream);

/* Inference-mode BatchNorm affine from moving statistics (no reduction). */
int semb_norm_from_moving(int32_t C, float eps, const float* gamma, const float* beta,
                          const float* moving_mean, const float* moving_var,
                          float* scale, float* shift, void* stream);

/* y = act( A(a) + actb(B(b)) )   with  A(a) = a*scale_a[g,c]+shift_a[g,c]  (mode_a)  and the same for b
 * (b optional: b==NULL).  g = n when *_nstride != 0 (per-sample / InstanceNorm), else 0.
 * Optionally accumulates moments of y into stats (same layout as semb_conv2d_fwd).
 * Covers BatchNormalization+Activation, add+Activation+BatchNormalization of multi_res_block /
 * res_path (UNet_Segmentation.py:425,469-473,492-494), GroupNormalization+ReLU/LeakyReLU and the
 * residual add of CycleGAN.py:329-336, and the final tanh/sigmoid. */
typedef struct {
    int32_t N, HW, C;
    int32_t dtype;
    int32_t act, actb;
    int32_t mode_a, mode_b;           /* SEMB_AFF_* */
    int32_t aff_nstride;              /* 0 = per-channel, else per-sample stride of scale/shift arrays */
    /* Optional: the b operand as a CONCATENATION of up to three tensors along the channel axis (`concatenate([a, b, c])`
     * of multi_res_block, UNet_Segmentation.py:469) without a concat buffer: segment s holds the channels
     * [seg_c0[s], seg_c0[s] + seg_b[s].C) of b.  nseg_b = 0: b is the plain tensor argument.  When segments are given
     * the `b` (and, in the backward kernels, `db`) arguments only need valid C; seg_db[s] receive the gradient slices
     * (ptr NULL = not needed).  An 8-channel slice of a 32-channel concat buffer costs the full 64-byte DRAM sector per
     * pixel; compact per-branch tensors do not. */
    int32_t nseg_b;
    int32_t seg_c0[3];
    semb_tensor seg_b[3];
    semb_tensor seg_db[3];
} semb_affine_desc;

int semb_affine_act_fwd(const semb_affine_desc* d,
                        const semb_tensor* a, const float* scale_a, const float* shift_a,
                        const semb_tensor* b, const float* scale_b, const float* shift_b,
                        const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                        void* stream);

/* Backward of semb_affine_act_fwd, pass 1 of 2.  The activation derivative is recomputed from the
 * pre-activation u = A(a)+actb(B(b)) (same fp32 arithmetic as the forward), so the saved output is not re-read.
 * With g = dy*act'(u), gb = g*actb'(B(b)) and xhat = (x-mean)*invstd accumulates sums[0]=sum g,
 * sums[1]=sum g*xhat_a, sums[2]=sum gb, sums[3]=sum gb*xhat_b (each at k*cstride + c, plus n*nstride for
 * per-sample mode); only operands in SEMB_AFF_BATCH mode are reduced. */
int semb_affine_act_bwd_reduce(const semb_affine_desc* d, const semb_tensor* dy,
                               const semb_tensor* a, const semb_tensor* b,
                               const float* scale_a, const float* shift_a, const float* mean_a, const float* invstd_a,
                               const float* scale_b, const float* shift_b, const float* mean_b, const float* invstd_b,
                               float* sums, int32_t sums_nstride, int32_t sums_cstride, void* stream);

/* C-length finalize of the BN/IN backward: for operand `which` (0=a, 1=b) turns the sums into
 * c1 = sum(g)/count, c2 = sum(g*xhat)/count, and accumulates dgamma += sum(g*xhat), dbeta += sum(g)
 * (either may be NULL; per-sample mode accumulates over samples). */
int semb_norm_bwd_finalize(const float* sums, int32_t which, int32_t groups, int32_t C, int32_t cstride,
                           int32_t sums_nstride, float count, const float* mean, const float* invstd,
                           float* c1, float* c2, float* dgamma, float* dbeta, void* stream);

/* pass 2 of 2: da (+)= scale_a*(g - c1_a - xhat_a*c2_a)   [SEMB_AFF_BATCH]
 *                     = scale_a*g [PLAIN] = g [NONE];  same for db with gb.  da/db may be NULL. */
int semb_affine_act_bwd_apply(const semb_affine_desc* d, const semb_tensor* dy,
                              const semb_tensor* a, const semb_tensor* b,
                              const float* scale_a, const float* shift_a, const float* mean_a, const float* invstd_a,
                              const float* c1_a, const float* c2_a,
                              const float* scale_b, const float* shift_b, const float* mean_b, const float* invstd_b,
                              const float* c1_b, const float* c2_b,
                              const semb_tensor* da, int32_t acc_a, const semb_tensor* db, int32_t acc_b,
                              void* stream);

/* semb_norm_finalize folded into its consumer: one record per normalised operand.  stats != NULL: the kernel computes
 * scale/shift from the fp64 moments (same arithmetic as semb_norm_finalize) and one block per group also WRITES
 * scale / shift / mean / invstd (mean / invstd may be NULL) and, for per-channel statistics, updates the moving
 * statistics (moving_mean may be NULL).  stats == NULL: scale / shift are inputs (constant or already finalized). */
typedef struct {
    const void* stats;                 /* fp64 moment accumulators, or NULL                          */
    int32_t stats_nstride, cstride;    /* strides in doubles / elements, as in semb_norm_finalize    */
    float count, eps;
    const float *gamma, *beta;         /* gamma may be NULL (scale=False)                            */
    float *scale, *shift, *mean, *invstd;
    float *moving_mean, *moving_var;
    float momentum;
} semb_norm_fin;

/* semb_affine_act_fwd with semb_norm_finalize folded in (85 C-length launches fewer per UNet step). */
int semb_affine_act_fwd_fin(const semb_affine_desc* d, const semb_tensor* a, const semb_norm_fin* fin_a,
                            const semb_tensor* b, const semb_norm_fin* fin_b, const semb_tensor* y, void* stats,
                            int32_t stats_nstride, int32_t stats_cstride, void* stream);

/* semb_affine_act_bwd_apply reading c1 = sum(g)/count, c2 = sum(g*xhat)/count straight from the sums written by
 * semb_affine_act_bwd_reduce, and accumulating dgamma / dbeta (each may be NULL): semb_norm_bwd_finalize folded in. */
int semb_affine_act_bwd_apply_sums(const semb_affine_desc* d, const semb_tensor* dy, const semb_tensor* a, const semb_tensor* b,
                                   const float* scale_a, const float* shift_a, const float* mean_a, const float* invstd_a,
                                   float count_a, float* dgamma_a, float* dbeta_a,
                                   const float* scale_b, const float* shift_b, const float* mean_b, const float* invstd_b,
                                   float count_b, float* dgamma_b, float* dbeta_b,
                                   const float* sums, int32_t sums_nstride, int32_t sums_cstride,
                                   const semb_tensor* da, int32_t acc_a, const semb_tensor* db, int32_t acc_b, void* stream);

/* Both backward passes and the C-length finalize in ONE cooperative launch: pass 1 (the sums of
 * semb_affine_act_bwd_reduce), a grid-wide barrier, dgamma += sum(g*xhat) / dbeta += sum(g) (each may be NULL), then
 * pass 2 (semb_affine_act_bwd_apply with c1 = sum(g)/count, c2 = sum(g*xhat)/count) walking every block's pixel range
 * backwards so that it re-reads from L2 what pass 1 read last.  At least one operand must be in SEMB_AFF_BATCH mode.
 * `sums` (4 x cstride floats per group) and `barrier` (two 32-bit words) must be ZERO on entry; the barrier words are
 * zero again on exit.  Returns SEMB_EWORKSPACE when the grid cannot be co-resident (fall back to the two-pass calls). */
int semb_affine_act_bwd_fused(const semb_affine_desc* d, const semb_tensor* dy, const semb_tensor* a, const semb_tensor* b,
                              const float* scale_a, const float* shift_a, const float* mean_a, const float* invstd_a,
                              float count_a, float* dgamma_a, float* dbeta_a,
                              const float* scale_b, const float* shift_b, const float* mean_b, const float* invstd_b,
                              float count_b, float* dgamma_b, float* dbeta_b,
                              float* sums, int32_t sums_nstride, int32_t sums_cstride, void* barrier,
                              const semb_tensor* da, int32_t acc_a, const semb_tensor* db, int32_t acc_b, void* stream);

/* out[c] += sum over pixels of x[.,c]   (bias gradient of Conv2DTranspose / biased Conv2D) */
int semb_channel_sum(const semb_tensor* x, int32_t N, int32_t HW, float* out, int32_t dtype, void* stream);

/* ---- pooling / padding ------------------------------------------------------------------- */

/* keras.layers.MaxPooling2D((2,2)) (UNet_Segmentation.py:525-537); bwd routes to the first maximum
 * in window scan order like ATen's max_pool2d backward. */
int semb_maxpool2x2_fwd(const semb_tensor* x, const semb_tensor* y, int32_t N, int32_t H, int32_t W,
                        int32_t dtype, void* stream);
int semb_maxpool2x2_bwd(const semb_tensor* x, const semb_tensor* dy, const semb_tensor* dx, int32_t N,
                        int32_t H, int32_t W, int32_t dtype, int32_t accumulate, void* stream);

/* mode 0: y = reflect_pad(x) (ReflectionPadding2D, UNet_Segmentation.py:578-589)
 * mode 1: y = crop(x)        (Cropping2D, :554)            top/left = offsets into x
 * mode 2: y (+)= zero_pad(x) (gradient of crop)
 * mode 3: y (+)= reflect_fold(x) (gradient of reflect_pad: mirrored borders are added back)
 * (H,W) is the size of x, (OH,OW) of y. */
int semb_pad_crop(const semb_tensor* x, const semb_tensor* y, int32_t N, int32_t H, int32_t W,
                  int32_t OH, int32_t OW, int32_t top, int32_t left, int32_t mode, int32_t dtype,
                  int32_t accumulate, void* stream);

/* Conv2DTranspose(2x2, stride 2) (UNet_Segmentation.py:542-551) is computed as a 1x1 conv with 4*C outputs
 * (one block of C per kernel position) followed by this depth-to-space pass:
 * dir 0: dst[n,2y+r,2x+s,c] = src[n,y,x,(2r+s)*C+c] + bias[c]  (bias may be NULL);  dir 1: the inverse gather (its gradient).
 * (H,W) is the size of src; dst is (2H,2W) and may be a channel slice of the skip-concat buffer. */
int semb_pixel_shuffle2(const semb_tensor* src, const semb_tensor* dst, int32_t N, int32_t H, int32_t W,
                        const float* bias, int32_t dir, int32_t dtype, void* stream);

/* UpSampling2D(size=(2,2)) nearest neighbour (CycleGAN.py:349, the use_resize_convolution branch of `upsample`).
 * dir 0: big[n,2y+r,2x+s,c] = small[n,y,x,c];  dir 1: small (+)= sum over the 2x2 block of big (its gradient).
 * (H,W) is the size of `small`; `big` is (2H,2W). */
int semb_upsample2x(const semb_tensor* small, const semb_tensor* big, int32_t N, int32_t H, int32_t W, int32_t dir,
                    int32_t accumulate, int32_t dtype, void* stream);

/* ---- WGAN-GP (WassersteinGAN.py) ------------------------------------------------------------------------
 * semb_mask_mul: y (+)= x * slope(z) * m over n_pixels pixels of C channels; slope(z) = 1 where z > 0 else negative_slope
 *   (z == NULL: 1), m = dropout keep mask already scaled by 1/(1-rate) (m == NULL: 1).  x = z: LeakyReLU(0.2) followed by
 *   Dropout (conv_block :547-567) in one pass; x = dy: its gradient; x = u: the critic linearised at z, which is what the
 *   gradient penalty's second-order term needs (dP/dW = backprop of <grad_x D(x_hat), dP/dgrad> through the SAME masks, since
 *   LeakyReLU and Dropout are piecewise linear -- the double backward of gradient_penalty :88-121 without a second-order graph).
 * semb_gp_direction: g = grad_x D(x_hat) (N samples x HW pixels x C channels).  Per sample norm = ||g_n||_2;
 *   sums[0] += (norm-1)^2, sums[1] += norm;  u_n = scale * (norm-1)/norm * g_n  (scale = 2*gp_weight/N gives
 *   u = d(gp_weight * mean_n (norm_n-1)^2)/dg). */
int semb_mask_mul(const semb_tensor* x, const semb_tensor* z, const semb_tensor* m, const semb_tensor* y, int64_t n_pixels,
                  float negative_slope, int32_t accumulate, int32_t dtype, void* stream);
int semb_gp_direction(const semb_tensor* g, const semb_tensor* u, int32_t N, int64_t HW, float scale, float* sums, int32_t dtype,
                      void* stream);

/* ---- parity mode on the tensor cores (fp32 storage, split bf16 operands) ----------------------------------------------
 * north_star: "within 1e-3 relative fp32 (bit-exact for the argmax mask)" for an implicit GEMM on tcgen05.  An fp32 value
 * is x = xh + xm + xl (three bf16 terms, exact to 2^-24); x*w ~ xh*wh + xh*wm + xm*wh + xh*wl + xl*wh + xm*wm keeps every
 * product above 2^-24 (6 terms).  The 3-term variant (x = xh + xl', xh*wh + xl'*wh + xh*wl') is 2^-16-accurate: measured
 * 1e-3 on the UNet's sigmoid map after 61 layers, i.e. NOT enough for a bit-exact mask; the engine uses 6 terms.
 * semb_split_bf16: fp32 tensor (C channels) -> bf16 tensor with 6C channels [xh|xh|xm|xh|xl|xm] (or 3C: [xh|xl'|xh]).
 * semb_split_weights: fp32 kernel (R,S,Cin,Cout) -> fp32 stacked kernel whose bf16 rounding (semb_pack_weights_tc) is
 *   [wh;wm;wh;wl;wh;wm] (3 terms: [wh;wh;wl']) along the input channels (axis 0, forward) or the output channels (axis 1:
 *   the flipped pack contracts over them, data gradient).
 * semb_conv2d_fwd_tc_f32: semb_conv2d_fwd_tc on the split operand (g->Cin = terms*C, g->dtype = SEMB_BF16) with an fp32
 *   result tensor y (pitch / coff in floats; accumulate adds to the fp32 content); moments come from the fp32 accumulators.
 * The weight gradient of such a layer is `terms` calls of semb_conv2d_wgrad_tc on channel slices of the split x and dy. */
int semb_split_bf16(const semb_tensor* src_f32, const semb_tensor* dst_bf16, int64_t n_pixels, void* stream);
int semb_split_weights(const float* w, int32_t R, int32_t S, int32_t Cin, int32_t Cout, float* ws, int32_t axis, int32_t terms,
                       void* stream);
int semb_conv2d_fwd_tc_f32(const semb_conv_geom* g, const semb_tensor* x3, const void* w_packed, const float* bias,
                           const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride, int32_t accumulate,
                           void* stream);

/* ---- tiled inference: HelperFunctions.tile_image / stitch_image (:17-141) as index arithmetic on the device ----------
 * The reference cuts an image into overlapping tiles on the host, calls the model tile by tile and stitches on the host
 * (UNet_Segmentation.py:335-343).  Here the image is uploaded once; tiles are numbered x-major (k = ix*ny + iy), xs[nx] /
 * ys[ny] are the per-axis tile offsets (device int32 arrays computed by the host from the reference's grid rule).
 * gather: tiles[k-k0] (fp32, th x tw, one channel) for k in [k0, k0+count), zero beyond the image edge.
 * stitch: all nx*ny predicted tiles -> out (H x W fp32); mode 0 maximum, 1 average, 2 centre crop (manage_overlap_mode). */
int semb_tile_gather(const float* img, int32_t H, int32_t W, float* tiles, int32_t th, int32_t tw, const int32_t* xs, int32_t nx,
                     const int32_t* ys, int32_t ny, int32_t k0, int32_t count, void* stream);
int semb_tile_stitch(const float* tiles, int32_t th, int32_t tw, float* out, int32_t H, int32_t W, const int32_t* xs, int32_t nx,
                     const int32_t* ys, int32_t ny, int32_t mode, void* stream);

/* ---- losses ------------------------------------------------------------------------------ */

/* weighted_bce (UNet_Segmentation.py:379-384) + the 'mae' and 'acc' metrics (:395) + d(loss)/d(p).
 * out[0] += sum w*bce, out[1] += sum |y-p|, out[2] += #((p>0.5)==y); caller divides by count.
 * dp = w*(-(y/p)+(1-y)/(1-p))/count inside the Keras clip range [1e-7,1-1e-7], else 0 (dp may be NULL). */
int semb_loss_wbce(const semb_tensor* p, const float* y_true, const semb_tensor* dp, int64_t count,
                   float weighting, float* out, int32_t dtype, void* stream);

/* Same loss / metrics / gradient, with p = sigmoid(z*scale[0] + shift[0]) recomputed in fp32 from the head's stored
 * pre-activation z (conv2d_bn(..., activation='sigmoid'), UNet_Segmentation.py:556-557; scale/shift = its BatchNorm affine).
 * The product path uses this one: a probability stored in bf16 rounds to exactly 1 above ~0.998, where Keras' fp32 clip
 * (1e-7) is still far away.  dp is still d(loss)/d(p), consumed by the sigmoid backward. */
int semb_loss_wbce_logits(const semb_tensor* z, const float* scale, const float* shift, const float* y_true,
                          const semb_tensor* dp, int64_t count, float weighting, float* out, int32_t dtype, void* stream);

/* sum |a-b| (kind 0, MeanAbsoluteError) or sum (a-b)^2 (kind 1, MeanSquaredError) into out[0] over the
 * first c_logical channels; b==NULL compares against the constant `target` (LSGAN labels,
 * CycleGAN.py:301-308).  da (+)= gscale * d/da, gscale already containing lambda/count. */
int semb_loss_l1_l2(const semb_tensor* a, const semb_tensor* b, float target, int32_t kind, int64_t n_pixels,
                    int32_t c_logical, float gscale, const semb_tensor* da, int32_t accumulate, float* out,
                    int32_t dtype, void* stream);

/* ---- optimizer / utilities --------------------------------------------------------------- */

/* keras.optimizers.Adam over one flat fp32 buffer (UNet_Segmentation.py:390-393, CycleGAN.py:168-171):
 * m += (g-m)(1-b1); v += (g*g-v)(1-b2); w -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)  (epsilon OUTSIDE
 * the bias correction, unlike torch.optim.Adam).  gscale multiplies g first (1/world_size after the NCCL
 * sum).  state: 16 device bytes {int64 t; float alpha; float pad}, zero-initialised by the caller; t is
 * incremented on the device so the call replays inside a CUDA graph.  lr_ptr: device float. */
int semb_adam_step(float* w, const float* g, float* m, float* v, int64_t n, const float* lr_ptr,
                   float beta1, float beta2, float eps, float gscale, void* state, void* stream);

int semb_fill_f32(float* p, int64_t n, float value, void* stream);
/* The NHWC float32 boundary of the Keras model call: dst view (dtype, 8-padded) <- src fp32 dense
 * [n_pixels, src_C] (extra channels zero-filled), and back (first dst_C channels). */
int semb_cast_in(const float* src, int32_t src_C, const semb_tensor* dst, int64_t n_pixels, int32_t dtype, void* stream);
int semb_cast_out(const semb_tensor* src, float* dst, int32_t dst_C, int64_t n_pixels, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMB200_H */
