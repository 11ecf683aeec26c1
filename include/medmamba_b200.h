/* medmamba_b200 -- C ABI of the B200-native SS2D hot path.
 *
 * The reference (leeminsun1205/MedMamba) is pure Python; the only native boundary on its hot
 * path is the third-party CUDA extension behind
 *     mamba_ssm.ops.selective_scan_interface.selective_scan_fn      (MedMamba.py:12, 273-279)
 * Every entry point below replaces that call, or the chain of ATen kernels the reference
 * launches around it in SS2D.forward / SS_Conv_SSM.forward, and cites the lines it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer in the current CUDA
 *     context unless stated otherwise; the caller owns and allocates every buffer;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it, nothing
 *     synchronises; the functions are stateless and re-entrant;
 *   - return value: MMB_OK (0), or a negative status: MMB_ERR_INVALID_ARG, MMB_ERR_UNSUPPORTED,
 *     or MMB_ERR_CUDA_BASE - cudaError_t for a launch failure.  Nothing throws.
 *   - element types are named by mmb_dtype; state, A, D, delta_bias and all parameter
 *     gradients are always fp32.
 */
#ifndef MEDMAMBA_B200_H
#define MEDMAMBA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMB_OK 0
#define MMB_ERR_INVALID_ARG (-1)
#define MMB_ERR_UNSUPPORTED (-2)
#define MMB_ERR_CUDA_BASE (-1000)

typedef enum { MMB_F32 = 0, MMB_BF16 = 1, MMB_F16 = 2 } mmb_dtype;

#define MMB_ABI_VERSION 4

/* ABI version of the loaded library (host only, no CUDA call). */
int mmb_abi_version(void);

/* SHA-256 (hex) of the sources and compiler flags this library was built from, or "unknown" (host only).
 * The Python loader compares it with the digest of the sources next to it and refuses a stale library: ctypes
 * checks neither arity nor types, so an old binary behind new call sites would corrupt memory silently. */
const char* mmb_source_digest(void);

/* Static description of a status code (host only). */
const char* mmb_status_string(int status);

/* Spacing (in steps) of the state checkpoints mmb_scan_fwd writes and mmb_scan_bwd recomputes from
 * (host only).  The `chunk_state` buffer holds ceil(seqlen / chunk) * dim * dstate floats per batch
 * element. */
int mmb_scan_chunk_len(void);

/* selective_scan_fn forward -- replaces selective_scan_cuda.fwd reached from MedMamba.py:273-279
 * (semantics: temp.py:57-139).
 *   u, delta, z, out : (batch, dim, seqlen), innermost stride 1, dtype io_dtype
 *                      (strides in ELEMENTS: *_bs per batch, *_ds per channel row)
 *   A                : (dim, dstate) fp32 contiguous          Dv, delta_bias: (dim) fp32 or NULL
 *   Bm, Cm           : (batch, ngroups, dstate, seqlen) dtype bc_dtype, arbitrary element strides
 *                      (b, g, n, l); dim % ngroups == 0, group g serves rows g*dim/ngroups ...
 *   z                : NULL, or the gate: out *= silu(z)
 *   last_state       : NULL or (batch, dim, dstate) fp32 contiguous
 *   chunk_state      : NULL or (batch, dim, nchunks, dstate) fp32: state at the END of each chunk
 *                      of mmb_scan_chunk_len() steps (what the backward recomputes from)
 * dstate <= 16 per launch: states are independent, so a caller with more runs groups of 16 over strided A / Bm / Cm views
 * and adds the outputs (medmamba_b200/selective_scan_interface.py:_wide_state_scan). */
int mmb_scan_fwd(const void* u, const void* delta, const float* A, const void* Bm, const void* Cm,
                 const float* Dv, const void* z, const float* delta_bias, void* out,
                 float* last_state, float* chunk_state,
                 int batch, int dim, int seqlen, int dstate, int ngroups,
                 int64_t u_bs, int64_t u_ds, int64_t delta_bs, int64_t delta_ds,
                 int64_t z_bs, int64_t z_ds, int64_t out_bs, int64_t out_ds,
                 int64_t B_bs, int64_t B_gs, int64_t B_ns, int64_t B_ls,
                 int64_t C_bs, int64_t C_gs, int64_t C_ns, int64_t C_ls,
                 int delta_softplus, int io_dtype, int bc_dtype, void* stream);

/* Row tiles mmb_scan_bwd splits a B/C group into (host only): the leading extent of dB_part / dC_part. */
int mmb_scan_bwd_row_tiles(int dim, int ngroups);

/* selective_scan_fn backward -- replaces selective_scan_cuda.bwd behind loss.backward() (train.py:284);
 * formulas: SURVEY.md Appendix B.  Inputs as mmb_scan_fwd plus
 *   dout        : (batch, dim, seqlen) gradient of out, dtype io_dtype, strides dout_bs / dout_ds
 *   chunk_state : the checkpoints the forward wrote (may be NULL when seqlen <= mmb_scan_chunk_len())
 * Outputs (all overwritten, no accumulation, no atomics):
 *   du, ddelta, dz : (batch, dim, seqlen) dense, dtype io_dtype (ddelta is w.r.t. the RAW delta input;
 *                    dz only when z != NULL)
 *   dB_part, dC_part : (row_tiles, batch, ngroups, dstate, seqlen) fp32 -- sum over axis 0 = dB, dC
 *   dA_part : (batch, dim, dstate) fp32;  dD_part, dbias_part : (batch, dim) fp32 -- sum over axis 0
 * dstate <= 16. */
int mmb_scan_bwd(const void* u, const void* delta, const float* A, const void* Bm, const void* Cm,
                 const float* Dv, const void* z, const float* delta_bias, const void* dout,
                 const float* chunk_state, void* du, void* ddelta, void* dz,
                 float* dB_part, float* dC_part, float* dA_part, float* dD_part, float* dbias_part,
                 int batch, int dim, int seqlen, int dstate, int ngroups,
                 int64_t u_bs, int64_t u_ds, int64_t delta_bs, int64_t delta_ds,
                 int64_t z_bs, int64_t z_ds, int64_t dout_bs, int64_t dout_ds,
                 int64_t B_bs, int64_t B_gs, int64_t B_ns, int64_t B_ls,
                 int64_t C_bs, int64_t C_gs, int64_t C_ns, int64_t C_ls,
                 int delta_softplus, int io_dtype, int bc_dtype, void* stream);

/* ---- fused SS2D path (channels-last; every tensor is indexed by token position p = h*W + w) ---- */

/* Depthwise 3x3 conv (padding 1) + bias + SiLU on a channels-last view.  Replaces the NHWC->NCHW
 * copy, cuDNN depthwise conv and SiLU of MedMamba.py:294-295 (conv defined at :153-161).
 *   x      : (batch, H, W, D) view, channel stride 1, pixel stride x_pixel_stride, batch stride
 *            x_batch_stride (elements) -- the first half of the in_proj output has pitch 2*D
 *   weight : (D, 1, 3, 3) fp32 contiguous;  bias: (D) fp32 or NULL
 *   out    : (batch, H, W, D) dense, dtype out_dtype (MMB_F32 or MMB_BF16)
 * D % 4 == 0. */
int mmb_dwconv3x3_silu_fwd(const void* x, const float* weight, const float* bias, void* out,
                           int batch, int H, int W, int D, int64_t x_pixel_stride, int64_t x_batch_stride,
                           int in_dtype, int out_dtype, void* stream);

/* dt_rank padded to the widths the core kernel is instantiated for (host only);
 * MMB_ERR_UNSUPPORTED when dt_rank > 32. */
int mmb_ss2d_core_dt_pad(int dt_rank);

/* Blocks per direction of the training geometry (host only): the third extent of `hsave`. */
int mmb_ss2d_core_train_blocks(int H, int W);

/* Launch geometry mmb_ss2d_core_fwd would use for this problem (host only, nothing is launched): lanes per channel
 * (1, 2 or 4), channels per CTA (a multiple of 32 / lanes, <= 256) and the number of channel tiles.  The choice comes
 * from a cost model fitted to measurements; exposed for tests and for capacity planning. */
int mmb_ss2d_core_plan(int batch, int H, int W, int D, int* lanes_per_channel, int* channels_per_cta,
                       int* channel_tiles);

/* Workspace of mmb_ss2d_core_fwd for this problem, in bytes (host only, nothing is launched; < 0: status).  The
 * caller allocates it (any contents; 16-byte aligned) and passes it to every forward call of that shape.  It holds the
 * segment summaries of the L-parallel passes.  Optional outputs describe the plan: segments per sequence (1 = the
 * sequences run whole, no workspace is touched) and resident CTAs per SM.  save_states != 0 plans the training forward. */
int64_t mmb_ss2d_core_fwd_workspace_bytes(int batch, int H, int W, int D, int dstate, int dt_rank, int xc_dtype,
                                          int save_states, int* segments, int* ctas_per_sm);

/* Four-direction selective scan of SS2D.forward_corev0 in one call.  Replaces the cross-scan
 * (MedMamba.py:256-257), the dt_proj einsum and its copy (:262, :266), selective_scan_fn (:273-279,
 * delta_softplus=True, delta_bias=dt_projs_bias, z=None) and the flips / transposes of the
 * cross-merge (:282-286).
 *   xc    : (batch, H, W, D) dense, xc_dtype MMB_F32 or MMB_BF16 -- u of all four directions
 *   proj  : (batch, H, W, 4, 32 + dt_pad) fp32 -- per direction k the x_proj of the token:
 *           [0,16) = B_n, [16,32) = C_n (rows n >= dstate zero), [32, 32+dt_rank) = dt_r, rest zero
 *   Wdt   : (4, D, dt_rank)   dt_bias: (4, D)   A: (4*D, dstate) (= -exp(A_logs))   Ds: (4*D)
 *   ydir  : (batch, H, W, 4, D) in ydir_dtype (MMB_F32, or MMB_BF16 with bf16 xc) -- direction k's scan output
 *           stored at the token it belongs to.
 *           MMB_F32: y_k = <C, h> + Ds_k * u.   MMB_BF16: the state term <C, h> alone (rounded to bf16);
 *           mmb_outnorm_gate_fwd adds u * sum_k Ds_k in fp32, so the large skip term is never rounded to bf16.
 *   hsave : NULL (inference), or (batch, 4, mmb_ss2d_core_train_blocks(H, W), D, 16) fp32: the state after
 *           every block of 8 steps, in each direction's own time order -- what mmb_ss2d_core_bwd recomputes from
 *   workspace : mmb_ss2d_core_fwd_workspace_bytes(...) bytes of scratch (see there)
 * Launches with fewer sequences than resident CTA slots split every sequence into segments: a first pass computes
 * the segment summaries (end state from a zero start, sum of delta), a second pass scans all segments in parallel
 * from the carried prefixes (the chunked scan north_star asks for).  A, Wdt 16-byte aligned.
 * Direction order and index maps: SURVEY.md Appendix A.  D % 4 == 0 (fp32 xc) or D % 8 == 0 (bf16 xc),
 * dstate <= 16, dt_rank <= 32. */
int mmb_ss2d_core_fwd(const void* xc, const float* proj, const float* Wdt, const float* dt_bias,
                      const float* A, const float* Ds, void* ydir, float* hsave, void* workspace,
                      int64_t workspace_bytes, int batch, int H, int W, int D, int dstate, int dt_rank, int dt_pad,
                      int xc_dtype, int ydir_dtype, void* stream);

/* y = ((y0 + y2) + y1) + y3 over ydir's direction slices (the operand order of MedMamba.py:298),
 * LayerNorm over D (MedMamba.py:300, eps as given) and * SiLU(z) (MedMamba.py:301).
 *   ydir : (tokens, 4, D), ydir_dtype MMB_F32 (full y_k) or MMB_BF16 (state terms only: then u = xc (tokens, D)
 *          bf16 dense and Dsum (D) fp32 = sum_k Ds_k must be given and y += u * Dsum);
 *   z    : (tokens, D) view with pixel stride z_pixel_stride, dtype z_dtype
 *   out  : (tokens, D) dense, dtype out_dtype (== z_dtype);  ymerged: NULL or (tokens, D) fp32, the
 *          pre-norm sum (kept for the backward).  D % 4 == 0, D <= 1024. */
int mmb_outnorm_gate_fwd(const void* ydir, const void* z, const float* gamma, const float* beta,
                         void* out, float* ymerged, const void* xc, const float* Dsum, int64_t tokens, int D,
                         int64_t z_pixel_stride, float eps, int ydir_dtype, int z_dtype, int out_dtype, void* stream);

/* out[..., 2j] = left[..., j] + inp[..., 2j];  out[..., 2j+1] = ssm[..., j] + inp[..., 2j+1]
 * -- torch.cat + channel_shuffle(groups=2) + residual of MedMamba.py:355-357 (:308-320).
 *   left, ssm : (tokens, c) views, channel stride 1, dtype branch_dtype
 *   inp       : (tokens, 2c) view, dtype res_dtype;  out: (tokens, 2c) dense, dtype res_dtype
 * c % 4 == 0.  branch_dtype == res_dtype, or 16-bit branches onto an fp32 residual stream (autocast). */
int mmb_shuffle_cat_residual_fwd(const void* left, const void* ssm, const void* inp, void* out,
                                 int64_t tokens, int c, int64_t left_pixel_stride,
                                 int64_t ssm_pixel_stride, int64_t inp_pixel_stride, int branch_dtype,
                                 int res_dtype, void* stream);

/* ---- callers either side of the path ---- */

/* LayerNorm over the channels of a channels-last token matrix: ln_1 on the strided right half of the residual
 * stream (MedMamba.py:351), the patch-embed norm (:75), the patch-merging norm (:116).  Inference path.
 *   x : (tokens, D) view, channel stride 1, pixel stride x_pixel_stride;  out: (tokens, D) dense
 * D % 4 == 0, D <= 2048. */
int mmb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* out, int64_t tokens, int D,
                      int64_t x_pixel_stride, float eps, int in_dtype, int out_dtype, void* stream);

/* out[tok, c] = x[tok, c] * scale[c] + shift[c]: the eval-mode BatchNorm that opens the CNN branch
 * (MedMamba.py:338, :352-353) fused with the gather of the strided left half and the cast to the conv dtype.
 *   x : (tokens, C) view, channel stride 1, pixel stride x_pixel_stride;  out: (tokens, C) dense.  C % 4 == 0. */
int mmb_affine_cast_fwd(const void* x, const float* scale, const float* shift, void* out, int64_t tokens, int C,
                        int64_t x_pixel_stride, int in_dtype, int out_dtype, void* stream);

/* Patch embedding (MedMamba.py:54-76): Conv2d(3 -> embed_dim, kernel 4, stride 4) of an NCHW image batch, the
 * NCHW -> NHWC permute and LayerNorm(embed_dim) in one pass.  Inference path; fp32 accumulation and statistics.
 *   x      : (batch, 3, Hin, Win) dense, dtype in_dtype (fp32 or bf16), 16-byte aligned rows (Win % 4 == 0)
 *   weight : (embed_dim, 3, 4, 4) fp32;  conv_bias: (embed_dim) fp32 or NULL;  gamma, beta: (embed_dim) fp32
 *   out    : (batch, Hin/4, Win/4, embed_dim) fp32 dense
 *   math_mode : 0 = fp32 FMA (bit-faithful to an fp32 convolution); 1 = the bf16 autocast arithmetic of the reference's
 *               convolution on tensor cores: operands rounded to bf16, fp32 accumulation (mma.sync m16n8k16), bias and
 *               LayerNorm in fp32.  Mode 1 needs fp32 images with Win / 4 <= 128 and falls back to mode 0 otherwise.
 * Hin % 4 == 0, Win % 4 == 0, embed_dim % 32 == 0, embed_dim <= 128; other shapes return MMB_ERR_UNSUPPORTED. */
int mmb_patch_embed_ln_fwd(const void* x, const float* weight, const float* conv_bias, const float* gamma,
                           const float* beta, float* out, int batch, int Hin, int Win, int embed_dim, float eps,
                           int in_dtype, int math_mode, void* stream);

/* Patch merging up to the norm (MedMamba.py:93-117): the 2x2 neighbourhood gather (x0, x1, x2, x3 = pixels
 * (2i,2j), (2i+1,2j), (2i,2j+1), (2i+1,2j+1)), their concatenation and LayerNorm(4C) in one pass; odd trailing
 * rows / columns are dropped like the reference's SHAPE_FIX.  The Linear(4C -> 2C) that follows stays a GEMM.
 *   x : (batch, H, W, C) dense channels-last;  gamma, beta: (4C) fp32;  out: (batch, H/2, W/2, 4C) dense
 * C % 4 == 0, C <= 512. */
int mmb_patch_merge_ln_fwd(const void* x, const float* gamma, const float* beta, void* out, int batch, int H, int W,
                           int C, float eps, int in_dtype, int out_dtype, void* stream);

/* ---- backward of the fused path (training; loss.backward(), train.py:284).  Parameter gradients come back
 * as per-CTA / per-batch partials that the caller sums over the leading axis: no float atomics, results are
 * bit-reproducible. ---- */

/* Rows of the partial buffers of mmb_outnorm_gate_bwd / mmb_layernorm_bwd, and of mmb_dwconv3x3_silu_bwd_ds (host only). */
int mmb_partial_blocks(void);
int mmb_dwconv_partial_blocks(void);

/* Channel groups (one per warp: 32 or 16 channels, by launch size) whose dproj partials mmb_ss2d_core_bwd writes for this
 * problem (host only): the leading extent of dproj_part.  Every partial row is written; nothing needs a memset. */
int mmb_ss2d_core_bwd_tiles(int batch, int D);

/* Gradient of mmb_ss2d_core_fwd given dY (batch, H, W, D) fp32 -- the gradient of the merged sum, identical for
 * the four direction slices of ydir.  xc, proj, Wdt, dt_bias, A, Ds as in the forward; hsave as written by it.
 *   dudir      : (batch, H, W, 4, D) fp32 -- d xc through direction k's `u` (sum over k = that part of d xc)
 *   dproj_part : (tiles, batch, H, W, 4, 32 + dt_pad) fp32 -- gradient of proj in proj's own layout
 *                [dB_n | dC_n | d dt_r]; sum over axis 0
 *   dA_part    : (batch, 4*D, dstate);  dW_part: (batch, 4*D, dt_pad) (= dWdt, first dt_rank columns);
 *   dD_part, db_part : (batch, 4*D) (= dDs, d dt_bias) -- all fp32, sum over axis 0 */
int mmb_ss2d_core_bwd(const void* xc, const float* proj, const float* dY, const float* Wdt,
                      const float* dt_bias, const float* A, const float* Ds, const float* hsave,
                      float* dudir, float* dproj_part, float* dA_part, float* dW_part, float* dD_part,
                      float* db_part, int batch, int H, int W, int D, int dstate, int dt_rank, int dt_pad, int xc_dtype,
                      void* stream);

/* Backward of mmb_layernorm_fwd (ln_1 and the patch norms in training).  The statistics are recomputed from x.
 *   x : the forward's input view (tokens, D), pixel stride x_pixel_stride, dtype x_dtype;  dy: (tokens, D) dense
 *   dx: (tokens, D) dense in x_dtype;  dgb_part: (mmb_partial_blocks(), 2, D) fp32 partials of dgamma / dbeta.
 * D % 4 == 0, D <= 512 (wider rows: MMB_ERR_UNSUPPORTED, the caller keeps torch's LayerNorm). */
int mmb_layernorm_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgb_part, int64_t tokens,
                      int D, int64_t x_pixel_stride, float eps, int x_dtype, int dy_dtype, void* stream);

/* Backward of mmb_outnorm_gate_fwd: dout (tokens, D) dense in z_dtype, ymerged from the forward ->
 *   dy (tokens, D) fp32, dz (tokens, D) in z_dtype with pixel stride dz_pixel_stride (D for a dense tensor; 2*D writes
 *   it straight into the z half of the gradient of the in_proj output),
 *   dgb_part (mmb_partial_blocks(), 2, D) fp32: [.,0,:] dgamma, [.,1,:] dbeta partials. */
int mmb_outnorm_gate_bwd(const void* dout, const float* ymerged, const void* z, const float* gamma,
                         const float* beta, float* dy, void* dz, float* dgb_part, int64_t tokens, int D,
                         int64_t z_pixel_stride, int64_t dz_pixel_stride, float eps, int z_dtype, void* stream);

/* Backward of mmb_dwconv3x3_silu_fwd, step 1: ds = g * silu'(dwconv(x) + bias) (pre-activation recomputed), where the
 * upstream gradient g of xc is the sum of the addends given (each may be NULL, not all): dxc (batch, H, W, D) fp32;
 * dudir (batch, H, W, 4, D) fp32, the per-direction gradients of mmb_ss2d_core_bwd; dxc_extra (batch, H, W, D) dense in
 * in_dtype, the x_proj GEMM's input gradient.  Folding the sum in here replaces a reduction, an add and a cast kernel.
 *   ds (batch, H, W, D) fp32, dwb_part (mmb_dwconv_partial_blocks(), D, 10) fp32: [., c, 0..8] dweight taps, [., c, 9] dbias. */
int mmb_dwconv3x3_silu_bwd_ds(const void* x, const float* weight, const float* bias, const float* dxc,
                              const float* dudir, const void* dxc_extra, float* ds, float* dwb_part, int batch, int H,
                              int W, int D, int64_t x_pixel_stride, int64_t x_batch_stride, int in_dtype, void* stream);

/* ... step 2: dx = ds correlated with the flipped 3x3 kernel; dx (batch, H, W, D) in out_dtype with pixel stride
 * dx_pixel_stride (D: dense; 2*D: the x half of the gradient of the in_proj output). */
int mmb_dwconv3x3_bwd_dx(const float* ds, const float* weight, void* dx, int batch, int H, int W, int D,
                         int64_t dx_pixel_stride, int out_dtype, void* stream);

/* Backward of mmb_shuffle_cat_residual_fwd w.r.t. the two branches (d inp = dout):
 *   dleft[..., j] = dout[..., 2j], dssm[..., j] = dout[..., 2j+1]; dense (tokens, c) in branch_dtype. */
int mmb_shuffle_cat_residual_bwd(const void* dout, void* dleft, void* dssm, int64_t tokens, int c,
                                 int res_dtype, int branch_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEDMAMBA_B200_H */
