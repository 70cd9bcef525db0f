/*
 * ss2d_b200.h — C ABI of libss2d_b200.so: the B200-native (sm_100a) SS2D hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / strides and a cudaStream_t passed
 * as void*, launches asynchronously on that stream and returns 0 on success, a negative SS2D_E* code
 * for a rejected argument (nothing launched) or a positive cudaError_t.  No torch types, no global
 * state; all buffers are DEVICE pointers owned by the caller (the torch shim allocates them).
 *
 * Each entry point replaces one reference interface (paths relative to the c95yang/FocalNet tree):
 *
 *   ss2d_selective_scan_fwd   selective_scan_fwd  kernels/selective_scan/csrc/selective_scan/cusoflex/
 *                             selective_scan_oflex.cpp:157-243 (pybind `fwd`, :361) and the "core" twin
 *                             cus/selective_scan.cpp (out dtype == input dtype)
 *   ss2d_selective_scan_bwd   selective_scan_bwd  selective_scan_oflex.cpp:245-358 (pybind `bwd`, :362)
 *   ss2d_cross_scan           CrossScanTriton.forward / CrossMergeTriton.backward  ITS/models/csm_triton.py:163-175,202-210
 *   ss2d_cross_merge          CrossMergeTriton.forward / CrossScanTriton.backward  ITS/models/csm_triton.py:188-200,177-185
 *   ss2d_cross_scan_fwd/_bwd  (fused) CrossScan -> selective scan -> CrossMerge, i.e. the body of
 *                             cross_selective_scan  ITS/models/vmamba_layers.py:261-291 without the 4x copies
 *   ss2d_dwconv_silu_fwd/_bwd permute + depthwise 3x3 conv + bias + SiLU  ITS/models/vmamba_layers.py:460-469,591-594
 *   ss2d_merge_norm_gate_fwd/_bwd  transpose + LayerNorm (+ z gate)       ITS/models/vmamba_layers.py:296-297,599
 *
 * Struct fields mirror SSMParamsBase / SSMParamsBwd (csrc/selective_scan/selective_scan.h:26-90) with
 * 64-bit strides (the reference's uint32 strides overflow past 4.29 G elements).  Strides are in ELEMENTS.
 */
#ifndef SS2D_B200_H_
#define SS2D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS2D_ABI_VERSION 3

/* element types of u / delta / B / C / z (in_dtype) and of out / dout (out_dtype) */
enum { SS2D_F32 = 0, SS2D_F16 = 1, SS2D_BF16 = 2 };

/* Kernel family that serves a scan (field `family` of the parameter structs).  The two families keep their checkpoints
 * in different layouts, so the forward and the backward of one problem MUST run on the same family: callers pass
 * SS2D_FAMILY_AUTO to the forward, read the choice back with ss2d_scan_family() and hand that value to the backward.
 * A pinned family that does not cover the shape (state-lanes needs dstate == 16) falls through to the other one. */
enum { SS2D_FAMILY_AUTO = 0, SS2D_FAMILY_STATELANES = 1, SS2D_FAMILY_WARPSCAN = 2 };

/* negative return codes (argument validation, nothing was launched) */
enum {
    SS2D_OK = 0,
    SS2D_EINVAL = -22,      /* null pointer / bad size / dim % ngroups != 0 / dstate > 256          */
    SS2D_EDTYPE = -2,       /* unsupported dtype combination (out must be F32 or == in)              */
    SS2D_ESTRIDE = -3,      /* a last-dimension stride != 1 (selective_scan_oflex.cpp:181-182,198), or x / ckpt /
                               work (and, for the fused seam with `work`, y / dy / dx / ddelta / dB / dC) not
                               16-byte aligned                                                        */
    SS2D_EDEVICE = -4       /* no sm_100 device / kernel image missing                               */
};

#define SS2D_MAX_DSTATE 256 /* selective_scan_oflex.cpp:192 */
#define SS2D_REF_CHUNK 2048 /* selective_scan_oflex.cpp:218: x has ceil(L/2048) checkpoints          */
#define SS2D_CKPT_STEPS 256 /* checkpoint spacing of the general (warp-scan) kernels, any dstate      */
#define SS2D_SL_BLOCK 16    /* checkpoint spacing of the state-lanes kernels (dstate == 16)          */

/* ---------------------------------------------------------------------------------------------
 * selective scan, scan-order operands (seam S1)
 *   u, delta : (batch, dim, seqlen)   in_dtype, last stride 1
 *   A        : (dim, dstate)          f32 contiguous
 *   B, C     : (batch, ngroups, dstate, seqlen) in_dtype, last stride 1
 *   D, delta_bias : (dim) f32 or NULL ; z : (batch, dim, seqlen) in_dtype or NULL
 *   out      : (batch, dim, seqlen)   out_dtype, contiguous rows (stride given)
 *   x        : (batch, dim, ceil(seqlen/2048), 2*dstate) f32 contiguous — (running prod a, h) at the
 *              end of every 2048-step chunk, exactly the reference's checkpoint tensor; may be NULL
 *   ckpt     : f32 workspace of ss2d_scan_ckpt_floats() elements, or NULL — the states h the backward restarts
 *              its recomputation from.  The layout is private to the library (it depends on which kernel
 *              family serves the shape): dstate == 16 -> (batch, ceil(seqlen/16), dim, 16), h at the end of
 *              every 16-step block; otherwise (batch, dim, ceil(seqlen/256), dstate), h every 256 steps
 *   out_z    : (batch, dim, seqlen) out_dtype, only with z: out_z = out * silu(z) (out stays un-gated)
 * ------------------------------------------------------------------------------------------- */
typedef struct ss2d_scan_fwd_params {
    int64_t batch, dim, seqlen, dstate, ngroups;
    int32_t in_dtype, out_dtype, delta_softplus, family;
    const void *u, *delta;
    const float *A;
    const void *B, *C;
    const float *D, *delta_bias;
    const void *z;
    int64_t u_bstride, u_dstride, delta_bstride, delta_dstride;
    int64_t B_bstride, B_gstride, B_nstride, C_bstride, C_gstride, C_nstride;
    int64_t z_bstride, z_dstride;
    void *out;
    int64_t out_bstride, out_dstride;
    void *out_z;
    float *x;
    float *ckpt;
} ss2d_scan_fwd_params;

/*   dout     : (batch, dim, seqlen) out_dtype, last stride 1
 *   ckpt     : the workspace written by the forward, or NULL together with x == the reference
 *              tensor (then the library rebuilds the checkpoints in `ckpt_scratch` first)
 *   du, ddelta : (batch, dim, seqlen) in_dtype contiguous ; dz likewise (only with z; needs out)
 *   dA (dim,dstate), dD, ddelta_bias (dim) : f32, MUST BE ZEROED by the caller (accumulated over batch)
 *   dB, dC   : (batch, ngroups, dstate, seqlen) f32 contiguous, MUST BE ZEROED (accumulated over the
 *              channels of a group); the shim casts them to in_dtype afterwards (oflex.cpp:356)     */
typedef struct ss2d_scan_bwd_params {
    ss2d_scan_fwd_params f; /* the forward operands (out/out_z unused unless z != NULL: then f.out = un-gated out) */
    const void *dout;
    int64_t dout_bstride, dout_dstride;
    float *ckpt_scratch; /* ss2d_scan_ckpt_floats() f32 elements, required when f.ckpt == NULL and seqlen > SS2D_SL_BLOCK */
    void *du, *ddelta, *dz;
    float *dA, *dB, *dC, *dD, *ddelta_bias;
} ss2d_scan_bwd_params;

int ss2d_abi_version(void);
/* "sm_100a" build tag + kernel variant list; static string */
const char *ss2d_build_info(void);
/* human-readable text for a return code of this library (negative) or of CUDA (positive) */
const char *ss2d_error_string(int code);

/* number of f32 elements of the `ckpt` workspace for a scan of this shape (0 on invalid sizes); one size serves
 * whichever family takes the shape */
int64_t ss2d_scan_ckpt_floats(int64_t batch, int64_t dim, int64_t seqlen, int64_t dstate);
/* the family (SS2D_FAMILY_STATELANES / _WARPSCAN) that serves a scan with these sizes / dtypes and this `family` request;
 * only batch, dim, seqlen, dstate, ngroups and family are read */
int ss2d_scan_family(const ss2d_scan_fwd_params *p);
/* TEST HOOK: what SS2D_FAMILY_AUTO resolves to for the whole process — 0 (default): by problem size; 1 / 2: that family.
 * Returns the previous value.  The product path never calls it (there is no environment switch). */
int ss2d_set_default_family(int family);

int ss2d_selective_scan_fwd(const ss2d_scan_fwd_params *p, void *stream);
int ss2d_selective_scan_bwd(const ss2d_scan_bwd_params *p, void *stream);

/* x:(B,C,H,W) -> xs:(B,4,C,H*W); dtype as enum above; both contiguous */
int ss2d_cross_scan(const void *x, void *xs, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype, void *stream);
/* ys:(B,4,C,H*W) -> y:(B,C,H*W) (spatial order), sum of the four un-permuted directions */
int ss2d_cross_merge(const void *ys, void *y, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype, void *stream);

/* src:(B,4,C,H*W), plane (b,k,c) in SPATIAL order -> dst: same planes, each in its direction k's SCAN order
 * (inverse != 0: scan order -> spatial).  The per-direction ("1b1") permutation of triton_cross_scan_1b1
 * (csm_triton.py:83-120); the fused path applies it to the 38-row x_dbl and to its gradient only. */
int ss2d_cross_permute(const void *src, void *dst, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype, int32_t inverse,
                       void *stream);

/* ---------------------------------------------------------------------------------------------
 * fused SS2D core (seam S3): CrossScan and CrossMerge are applied in the scan kernel's load / store
 * addressing — no (B,4,D,L) copy of x (100.7 MB at the microbench) and no (B,4,D,L) ys is ever materialised.
 *   x      : (batch, D, H, W)         in_dtype, contiguous, SPATIAL order: the u of all four directions; the
 *            kernel of direction k reads pixel perm_k(l) for scan step l (vmamba_layers.py:35-37)
 *   delta  : (batch, 4*D, H*W)        in_dtype, SCAN order of each direction (dt_proj output, pre-softplus)
 *   B, C   : (batch, 4, dstate, H*W)  in_dtype, SCAN order of each direction.  delta/B/C are projections of the
 *            38-row x_dbl, which the host permutes per direction before dt_proj — 5x fewer bytes than permuting x
 *   A (4*D,dstate), Dskip (4*D), delta_bias (4*D) : f32, channel index k*D+d (vmamba_layers.py:273-279)
 *   y      : (batch, D, H*W) f32, SPATIAL order, = CrossMerge of the four scans; MUST BE ZEROED by the caller
 *            (each direction accumulates with red.global.add.f32, so the sum order is not deterministic)
 *            y == NULL with ckpt != NULL: states-only sweep (rebuilds ckpt for a backward that did not keep it)
 *   ckpt   : ss2d_scan_ckpt_floats(batch, 4*D, H*W, dstate) f32 elements, consumed by the backward; NULL: not written
 *            (inference)
 * ------------------------------------------------------------------------------------------- */
typedef struct ss2d_cross_fwd_params {
    int64_t batch, D, H, W, dstate;
    int32_t in_dtype, delta_softplus;
    int32_t family;        /* SS2D_FAMILY_*: as in ss2d_scan_fwd_params (ss2d_cross_family() reports the choice)          */
    int32_t deterministic; /* != 0: y / dx are bit-reproducible run to run, like triton_cross_merge (csm_triton.py:45-80).
                              State-lanes kernels: always (y and y^T each take two commutative adds onto zero, then one
                              transposing add: y = (y0 + y2) + (y1 + y3)).  Warp-scan kernels: one launch per direction in
                              the order k = 0, 1, 2, 3 instead of one launch whose red.global.add arrive in any order.   */
    const void *x, *delta, *B, *C;
    const float *A, *Dskip, *delta_bias;
    float *y;
    float *ckpt;
    int64_t bc_bstride, bc_gstride; /* element strides of B and C over batch / direction (rows of dstate are L apart);
                                       0 = contiguous.  Lets B, C be views into the permuted x_dbl (no .contiguous() copy) */
    float *work;                    /* NULL, or scratch of ss2d_cross_work_floats() f32 elements (16-byte aligned).  With it,
                                       fp32 / dstate 16 / H*W % 16 == 0 problems run on the state-lanes kernels: directions
                                       1 and 3 walk ONE transposed copy of x (and accumulate into a transposed y that is
                                       folded back at the end) so that every direction moves contiguous runs.  Pass it to
                                       the forward AND the backward of a problem, or to neither (the checkpoint layouts of
                                       the two kernel families differ). */
} ss2d_cross_fwd_params;

/*   dy : (batch, D, H*W) f32 spatial (gathered per direction = CrossMerge.backward).
 *   dx : (batch, D, H*W) f32 spatial, ZEROED (du of the 4 directions accumulated = CrossScan.backward);
 *   ddelta (batch,4*D,L) in_dtype scan order; dB, dC (batch,4,dstate,L) f32 scan order ZEROED;
 *   dA (4*D,dstate), dDskip, ddelta_bias (4*D) f32 ZEROED.  f.ckpt (from the forward) is required when L > 256
 *   (L > 16 with f.work); ckpt_scratch is reserved (must be NULL); f.work: 3*batch*D*H*W f32 elements here. */
typedef struct ss2d_cross_bwd_params {
    ss2d_cross_fwd_params f;
    const float *dy;
    float *ckpt_scratch;
    float *dx;
    void *ddelta;
    float *dA, *dB, *dC, *dDskip, *ddelta_bias;
} ss2d_cross_bwd_params;

/* f32 elements of `work`: 2*batch*D*H*W for the forward, 3*batch*D*H*W for the backward; 0 when the state-lanes
 * kernels do not take the problem (not fp32, dstate != 16, H*W % 16 != 0, fewer than ~4.6 k channel sequences) */
int64_t ss2d_cross_work_floats(int64_t batch, int64_t D, int64_t H, int64_t W, int64_t dstate, int32_t in_dtype, int32_t backward,
                               int32_t family);
/* family serving the fused problem (given `work` is supplied as ss2d_cross_work_floats asks) */
int ss2d_cross_family(int64_t batch, int64_t D, int64_t H, int64_t W, int64_t dstate, int32_t in_dtype, int32_t family);
/* src: `planes` images of (H, W) -> dst: images of (W, H); dst = src^T, or dst += src^T when accumulate != 0 */
int ss2d_plane_transpose(const float *src, float *dst, int64_t planes, int64_t H, int64_t W, int32_t accumulate, void *stream);

int ss2d_cross_scan_fwd(const ss2d_cross_fwd_params *p, void *stream);
int ss2d_cross_scan_bwd(const ss2d_cross_bwd_params *p, void *stream);

/* ---------------------------------------------------------------------------------------------
 * depthwise 3x3 conv (pad 1) + bias + SiLU reading channels-last, writing channels-first.
 *   xin  : (batch, H, W, cstride) f32 — first C channels of every pixel are convolved (the in_proj
 *          output keeps x | z interleaved per pixel, vmamba_layers.py:585-587)
 *   weight (C,3,3), bias (C) or NULL : f32 ;  out : (batch, C, H, W) f32
 * bwd: dout (batch,C,H,W) -> dxin (batch,H,W,dx_cstride) first C channels written; dweight (C,9), dbias (C)
 *      f32 ZEROED (accumulated); dpre_scratch (batch,H,W,C) f32 work buffer
 * ------------------------------------------------------------------------------------------- */
int ss2d_dwconv_silu_fwd(const float *xin, int64_t cstride, const float *weight, const float *bias, float *out,
                         int64_t batch, int64_t C, int64_t H, int64_t W, void *stream);
int ss2d_dwconv_silu_bwd(const float *xin, int64_t cstride, const float *weight, const float *bias, const float *dout,
                         float *dpre_scratch, float *dxin, int64_t dx_cstride, float *dweight, float *dbias,
                         int64_t batch, int64_t C, int64_t H, int64_t W, void *stream);

/* ---------------------------------------------------------------------------------------------
 * epilogue of the SS2D core: transpose + LayerNorm over the D channels (+ SiLU gate) in one pass.
 *   y    : (batch, D, L) f32, the merged scan output in spatial order (ss2d_cross_fwd_params.y)
 *   weight, bias : (D) f32 LayerNorm affine ; eps as nn.LayerNorm
 *   z    : NULL, or the raw z-half of the in_proj output, channels-last: pixel (b,l) starts at z + (b*L+l)*z_pstride
 *          (vmamba_layers.py:587-589,599: z = act(z); y = y * z) ; out : (batch, L, D) f32 channels-last
 * replaces y.transpose(1,2).contiguous() + out_norm(y) (+ y * z)  ITS/models/vmamba_layers.py:296-297,599.  D <= 512.
 * bwd: dout (batch,L,D) -> dy (batch,D,L); dz (first D channels of every pixel, dz_pstride) or NULL;
 *      dweight, dbias (D) f32 ZEROED (accumulated).
 * ------------------------------------------------------------------------------------------- */
int ss2d_merge_norm_gate_fwd(const float *y, const float *weight, const float *bias, float eps, const float *z,
                             int64_t z_pstride, float *out, int64_t batch, int64_t D, int64_t L, void *stream);
int ss2d_merge_norm_gate_bwd(const float *y, const float *weight, const float *bias, float eps, const float *z,
                             int64_t z_pstride, const float *dout, float *dy, float *dz, int64_t dz_pstride,
                             float *dweight, float *dbias, int64_t batch, int64_t D, int64_t L, void *stream);

/* ---------------------------------------------------------------------------------------------
 * low-rank dt projection of the SS2D core (first half of SURVEY §8f row N2): replaces
 *     dts = F.conv1d(dts.contiguous().view(B, -1, L), dt_projs_weight.view(K * D, -1, 1), groups=K)
 * of cross_selective_scan (ITS/models/vmamba_layers.py:264; einsum twin :270) and its backward, as plain HBM streams.
 *   dtlr : f32, element (b, k, r, l) at dtlr + b*sb + k*sk + r*sr + l  (the dt rows of the permuted x_dbl; no copy)
 *   W    : (K, D, R) f32 = dt_projs_weight ; out / dout : (B, K*D, L) f32 contiguous ; R <= 8, D <= 512
 *   bwd  : d_dtlr (B, K, R, L) f32 contiguous, written ; dW (K, D, R) f32 ZEROED (accumulated)
 * L % 4 == 0 and 16-byte aligned rows are required (SS2D_ESTRIDE otherwise: the caller keeps the library conv).
 * ------------------------------------------------------------------------------------------- */
int ss2d_dt_proj_fwd(const float *dtlr, int64_t sb, int64_t sk, int64_t sr, const float *W, float *out, int64_t B, int64_t K,
                     int64_t D, int64_t R, int64_t L, void *stream);
int ss2d_dt_proj_bwd(const float *dout, const float *dtlr, int64_t sb, int64_t sk, int64_t sr, const float *W, float *d_dtlr,
                     float *dW, int64_t B, int64_t K, int64_t D, int64_t R, int64_t L, void *stream);

/* ---------------------------------------------------------------------------------------------
 * optimizer side of the data-parallel training step: global-norm clip + Adam + zero-grad over ONE
 * flat fp32 bucket, two launches.  Replaces clip_grad_norm_(params, 0.001); optimizer.step(); optimizer.zero_grad()
 * of ITS/train.py:61,89-91 (Adam(lr, betas=(0.9, 0.999), eps=1e-8), train.py:16).
 *   param, grad, exp_avg, exp_avg_sq : n f32 each, 16-byte aligned; grad holds the all-reduce SUM over ranks and is ZEROED
 *   on return; grad_scale = 1 / world size; partials: SS2D_OPTIM_PARTIALS f32 scratch; norm_out: NULL or 1 f32 that
 *   receives the total norm of the averaged gradient (what clip_grad_norm_ returns); step >= 1 (bias correction);
 *   max_norm <= 0 disables clipping.
 * ------------------------------------------------------------------------------------------- */
#define SS2D_OPTIM_PARTIALS 592
int64_t ss2d_optim_partials(void);
int ss2d_optim_clip_adam(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float *partials,
                         float *norm_out, float lr, float beta1, float beta2, float eps, int64_t step, float max_norm,
                         float grad_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SS2D_B200_H_ */
