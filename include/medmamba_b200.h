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

#define MMB_ABI_VERSION 1

/* ABI version of the loaded library (host only, no CUDA call). */
int mmb_abi_version(void);

/* Static description of a status code (host only). */
const char* mmb_status_string(int status);

/* Length of the L-chunks mmb_scan_fwd checkpoints the state at (host only).  The optional
 * `chunk_state` buffer of mmb_scan_fwd holds ceil(seqlen / chunk) * dim * dstate floats per
 * batch element. */
int mmb_scan_chunk_len(int batch, int dim, int seqlen);

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
 * dstate <= 16. */
int mmb_scan_fwd(const void* u, const void* delta, const float* A, const void* Bm, const void* Cm,
                 const float* Dv, const void* z, const float* delta_bias, void* out,
                 float* last_state, float* chunk_state,
                 int batch, int dim, int seqlen, int dstate, int ngroups,
                 int64_t u_bs, int64_t u_ds, int64_t delta_bs, int64_t delta_ds,
                 int64_t z_bs, int64_t z_ds, int64_t out_bs, int64_t out_ds,
                 int64_t B_bs, int64_t B_gs, int64_t B_ns, int64_t B_ls,
                 int64_t C_bs, int64_t C_gs, int64_t C_ns, int64_t C_ls,
                 int delta_softplus, int io_dtype, int bc_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEDMAMBA_B200_H */
