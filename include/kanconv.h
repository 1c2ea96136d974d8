/* kanconv.h - C ABI of libkanconv.so: the KAN-convolution hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (GadGadGad/Convolutional-KAN-for-Image-Classification) is pure eager PyTorch and has NO native
 * interface of its own.  This header therefore defines the boundary a native replacement
 * has to offer underneath the reference's Python plugin API (layers/kan_conv.py:726-745, CONV_KAN_FACTORY):
 * each entry point below replaces the arithmetic of the reference lines cited next to it.  All pointers are raw
 * device pointers owned by the caller (PyTorch's caching allocator in the Python binding); the library allocates
 * nothing persistent, is re-entrant, takes the CUDA stream explicitly and never throws across the ABI.
 *
 * Conventions
 *   - activations are fp32 NCHW; one call handles ONE group of a grouped layer: `cin`/`cout` are per-group
 *     channel counts and `x_batch_stride` / `z_batch_stride` (in elements) let the caller point into the
 *     channel slice of the full tensor (the reference loops over groups in Python, kan_layers.py:249-258).
 *   - weights are fp32 in the reference's own parameter layout, so state_dicts are
 *     interchangeable.  `w_basis` inner index: c*nb+j (B-spline, Chebyshev, RBF) or j*cin+c (Gram).
 *   - return value: KC_OK, or a negative kc_status; kc_last_error() gives the message of the calling thread.
 *     KC_ERR_UNSUPPORTED means "this entry point cannot run this shape" - the binding raises NotImplementedError;
 *     there is no CPU or library fallback behind any entry point.
 */
#ifndef KANCONV_H_
#define KANCONV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KC_ABI_VERSION 2
#define KC_MAX_BASIS 16   /* max basis width nb  (G+K, D+1 or G)      */
#define KC_MAX_PARAMS 40  /* knots / rbf grid etc. carried in kc_desc */

typedef enum kc_status {
  KC_OK = 0,
  KC_ERR_INVALID = -1,      /* bad argument (binding raises ValueError)           */
  KC_ERR_UNSUPPORTED = -2,  /* shape not covered by this entry point               */
  KC_ERR_CUDA = -3          /* CUDA runtime error, message in kc_last_error()      */
} kc_status;

typedef enum kc_basis_kind {
  KC_BASIS_BSPLINE = 0, /* KANConvNDLayer      layers/kan_layers.py:203-233  (Cox-de Boor)               */
  KC_BASIS_CHEBY = 1,   /* ChebyKANConvNDLayer layers/cheby_kan_layers.py:93-96                          */
  KC_BASIS_GRAM = 2,    /* GRAMKANConvNDLayer  layers/gram_kan_layers.py:150-181 (SiLU of Gram polys)    */
  KC_BASIS_RBF = 3,     /* FastKANConvNDLayer  utils/utils.py:19-33 (Gaussian RBF)                       */
  /* Polynomials by a three-term recurrence with constant coefficients on t = tanh(x) (or on x itself, params[0] != 0):
   *   p_0 = c0,  p_1 = a1 t + b1,  p_i = (A_i t + B_i) p_{i-1} + C_i p_{i-2}     (i = 2 .. nb-1)
   * params = { pre, c0, a1, b1, A_2, B_2, C_2, A_3, B_3, C_3, ... }, nparams = 4 + 3 * max(nb - 2, 0).
   * One functor for the reference's Hermite / Gegenbauer / Laguerre / Lucas / Fibonacci / Bessel / Taylor layers
   * (layers/hermite_kan_layers.py:127-150 and the same method of the sibling files: expanded channel c*nb + j) and, with
   * the degree-major channel order j*cin + c of `torch.concatenate(polys, dim=1)`, for the Legendre / Jacobi layers
   * (layers/legendre_kan_layers.py:108-121, layers/jacobi_kan_layers.py:119-137). */
  KC_BASIS_RECUR = 4,   /* channel-major  (c*nb + j)  */
  KC_BASIS_RECUR_DM = 5 /* degree-major   (j*cin + c) */
} kc_basis_kind;

typedef enum kc_act_kind {
  KC_ACT_NONE = -1,    /* layer has no base branch (Chebyshev)                              */
  KC_ACT_IDENTITY = 0, /* base_activation=None -> nn.Identity  (kan_layers.py:132)          */
  KC_ACT_GELU = 1,     /* exact erf GELU (layer default, kan_layers.py:276)                 */
  KC_ACT_SILU = 2      /* model default (models/kan_vgg.py:47)                              */
} kc_act_kind;

typedef enum kc_norm_kind { KC_NORM_NONE = 0, KC_NORM_INSTANCE = 1, KC_NORM_BATCH = 2 } kc_norm_kind;
typedef enum kc_out_act_kind { KC_OUT_NONE = 0, KC_OUT_PRELU = 1, KC_OUT_SILU = 2 } kc_out_act_kind;

/* Geometry + basis description of ONE group of one KAN convolution layer (plain data, no pointers). */
typedef struct kc_desc {
  int32_t basis;              /* kc_basis_kind                                                     */
  int32_t act;                /* kc_act_kind of the base branch                                    */
  int32_t n, cin, h, w;       /* input  [n, cin, h, w]   (cin per group)                           */
  int32_t cout, ho, wo;       /* output [n, cout, ho, wo] (cout per group)                         */
  int32_t kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int32_t nb;                 /* basis width: G+K | D+1 | G                                        */
  int32_t order;              /* spline order K | polynomial degree D | unused                     */
  int32_t nparams;            /* B-spline: G+2K+1 knots; RBF: G grid points then the denominator;
                                 RECUR: 4 + 3*(nb-2) recurrence coefficients (see kc_basis_kind)     */
  int64_t x_batch_stride;     /* elements between images of x   (>= cin*h*w)                       */
  int64_t z_batch_stride;     /* elements between images of z/dz (>= cout*ho*wo)                   */
  float params[KC_MAX_PARAMS];
} kc_desc;

/* Row-wise normalisation + output activation over [n, c, hw] planes (kan_layers.py:241-243). */
typedef struct kc_norm_desc {
  int32_t norm;        /* kc_norm_kind     */
  int32_t out_act;     /* kc_out_act_kind  */
  int32_t n, c, hw;
  int32_t affine;      /* gamma/beta present */
  int64_t batch_stride;/* elements between images (>= c*hw) for z, y, dy, dz alike */
  float eps;
} kc_norm_desc;

int kc_version(void);
const char* kc_last_error(void);
/* Number of SMs / compute capability of the current device, for the binding's sanity check (returns KC_OK). */
int kc_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Number of CUDA kernels this library has launched in the calling process so far (monotonic; for bench accounting). */
long long kc_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * FP32 path (CUDA-core FFMA, fp32 accumulate): any kernel size / stride / dilation / padding, nb <= 16.
 * Parity target: <= 1e-5 relative to the reference fp32 modules.
 * ------------------------------------------------------------------------------------------------------- */

/* z = conv(act(x_base), w_base) + conv(basis(x_basis), w_basis)       [kan_layers.py:199-241; cheby:93-97;
 * gram:173-186; fast:103-109].  x_basis == x_base except for FastKAN (basis of the normalised input).
 * w_base may be NULL iff desc->act == KC_ACT_NONE.  `beta` = GRAM beta_weights[D+1] (device), else NULL. */
int kc_conv_fwd_f32(const kc_desc* d, const float* x_base, const float* x_basis, const float* w_base,
                    const float* w_basis, const float* beta, float* z, void* stream);

/* Input gradient.  dx_base/dx_basis receive the base-branch and basis-branch parts; if they are the same pointer
 * the sum is written once.  `dbeta` (GRAM only, may be NULL): a buffer of kc_dbeta_floats(d, tc) floats; on return its
 * first D+1 floats hold d/d beta_weights of this call (gram_kan_layers.py:150-170).  The kernels write one partial row
 * per thread block behind them and a fixed-order reduction adds the rows: the result is deterministic, nothing has to be
 * zeroed, and there are no atomics.
 * Replaces autograd's backward of kan_layers.py:199-239 (the reference has no hand-written backward). */
size_t kc_dbeta_floats(const kc_desc* d, int tc);      /* tc = 0: kc_conv_dgrad_f32, 1: kc_conv_dgrad_tc; 0 if not GRAM */
int kc_conv_dgrad_f32(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                      const float* w_base, const float* w_basis, const float* beta, float* dx_base,
                      float* dx_basis, float* dbeta, void* stream);

/* Weight gradients in the reference layouts.  `workspace` must hold kc_wgrad_workspace_bytes(d) bytes. */
size_t kc_wgrad_workspace_bytes(const kc_desc* d);
int kc_conv_wgrad_f32(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                      const float* beta, float* dw_base, float* dw_basis, void* workspace, void* stream);

/* Depthwise KAN convolution: `channels` single-channel groups (desc->cin == desc->cout == 1, batch strides covering all
 * channels) in one launch each for forward, dX and dW.  Replaces the per-group Python loop of
 * KANConvNDLayer.forward (layers/kan_layers.py:249-258) for the `replace_depthwise=True` stage of MobileNetV2
 * (models/kan_mobilenetv2.py:112-124: groups == channels, up to 960 iterations per layer upstream).
 * w_base [channels][kh*kw] and w_basis [channels][nb][kh*kw] are the per-group filters stacked in group order.  FP32; every
 * family but GRAM; nb + 1 <= 9, kh*kw <= 9.  kc_dwconv_supported() != 0 iff the three entry points accept the shape. */
int kc_dwconv_supported(const kc_desc* d, int channels);
int kc_dwconv_fwd_f32(const kc_desc* d, int channels, const float* x_base, const float* x_basis, const float* w_base,
                      const float* w_basis, float* z, void* stream);
int kc_dwconv_dgrad_f32(const kc_desc* d, int channels, const float* dz, const float* x_base, const float* x_basis,
                        const float* w_base, const float* w_basis, float* dx_base, float* dx_basis, void* stream);
size_t kc_dwconv_wgrad_workspace_bytes(const kc_desc* d, int channels);
int kc_dwconv_wgrad_f32(const kc_desc* d, int channels, const float* dz, const float* x_base, const float* x_basis,
                        float* dw_base, float* dw_basis, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Max pooling between convolution stages (nn.MaxPool2d(k, s), no padding / dilation; models/kan_vgg.py "M" entries).
 * x is [planes][h][w] fp32 (planes = n*c), y / idx are [planes][ho][wo] with ho = (h-k)/s + 1; idx holds the
 * row-major position of the maximum inside its window (one byte, k <= 11).  Ties keep the first maximum and NaN
 * propagates, as in ATen.  bwd is a deterministic gather and writes every element of dx.
 * ------------------------------------------------------------------------------------------------------- */
int kc_maxpool2d_fwd(const float* x, float* y, unsigned char* idx, long long planes, int h, int w, int k, int s,
                     int ho, int wo, void* stream);
int kc_maxpool2d_bwd(const float* dy, const unsigned char* idx, float* dx, long long planes, int h, int w, int k, int s,
                     int ho, int wo, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Normalisation + output activation (memory-bound, vectorised):  y = out_act(gamma * (z-mean)*rstd + beta)
 * [kan_layers.py:242-243, gram:187, cheby:98; fast:106 uses it on the INPUT with out_act = NONE].
 * mean/rstd are [n*c] (instance) or [c] (batch) and are outputs of fwd / inputs of bwd.
 * alpha = PReLU weight (1 element, device).  bwd writes dz and per-plane partials that it then reduces into
 * dgamma[c], dbeta[c], dalpha[1] (any of which may be NULL).  `partials` needs 3*n*c + 2*c floats.
 * fwd `scratch` (2*n*c floats) is only used for KC_NORM_BATCH with batch statistics (per-plane sums) and may be NULL
 * otherwise.  KC_NORM_BATCH with scratch == NULL means "normalise with the GIVEN statistics": mean / rstd are inputs
 * (nn.BatchNorm in eval mode: running_mean and rsqrt(running_var + eps)) and are not modified.
 * bwd with `given_stats` != 0 (eval-mode batch norm) treats mean / rstd as constants: dz = rstd * gamma * dy * act'.
 * ------------------------------------------------------------------------------------------------------- */
int kc_norm_act_fwd(const kc_norm_desc* d, const float* z, const float* gamma, const float* beta,
                    const float* alpha, float* y, float* mean, float* rstd, float* scratch, void* stream);
int kc_norm_act_bwd(const kc_norm_desc* d, const float* dy, const float* z, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, const float* alpha, float* dz, float* dgamma,
                    float* dbeta, float* dalpha, float* partials, int given_stats, void* stream);

/* LayerNorm over the features of each row + output activation: the tail of the fully-connected KANLayer
 * (layers/kan_layers.py:110-112, nn.LayerNorm(out_features) -> nn.PReLU; models/kans.py:300-327 stacks them).  z, y, dy, dz are
 * [rows, features] fp32; gamma / beta are per FEATURE (elementwise affine); mean / rstd are [rows]; `partials` = rows floats. */
typedef struct kc_rownorm_desc {
  int32_t rows, features;
  int32_t out_act;     /* kc_out_act_kind */
  int32_t affine;      /* gamma/beta present */
  float eps;
} kc_rownorm_desc;
int kc_layernorm_act_fwd(const kc_rownorm_desc* d, const float* z, const float* gamma, const float* beta,
                         const float* alpha, float* y, float* mean, float* rstd, void* stream);
int kc_layernorm_act_bwd(const kc_rownorm_desc* d, const float* dy, const float* z, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, const float* alpha, float* dz, float* dgamma,
                         float* dbeta, float* dalpha, float* partials, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * BF16 tensor-core path (tcgen05.mma, fp32 accumulate in TMEM).  Dilation 1, stride 1..4 (a strided layer runs on the
 * stride-1 position grid and stores the sampled outputs), kernel size <= 8, padding < kernel size, basis width
 * nb <= 8 (zero-padded to 4 or 8 inside), any channel counts.  kc_tc_supported() tells the binding which path a
 * shape takes.
 * Parity target: <= 2e-2 relative / 1e-3 absolute to the reference fp32 modules.
 * ------------------------------------------------------------------------------------------------------- */
int kc_tc_supported(const kc_desc* d);
/* 1 when kc_conv_fwd_tc needs a phi_out buffer even for inference: 1x1 layers run as a basis pre-pass into phi followed by
 * a GEMM that streams phi back (fusing the basis into the GEMM leaves the tensor cores idle when there is one tap). */
int kc_tc_fwd_needs_phi(const kc_desc* d);
/* Bytes of the packed bf16 weight images of the forward (which=0) and dgrad (which=1) kernels, of the flat bf16
 * dz buffer that kc_tc_dz_flat fills (which=2), of the wgrad workspace when the forward did not save its basis rows
 * (which=3: split-K partial sums + a transient basis buffer), of the saved basis rows `phi` that kc_conv_fwd_tc can
 * emit for kc_conv_wgrad_tc (which=4), and of the wgrad workspace when `phi` is supplied (which=5). */
size_t kc_tc_bytes(const kc_desc* d, int which);
/* dz (fp32 NCHW) -> bf16 "flat" layout [n*(h+pad_h)*(w+pad_w)][round_up(cout,16)], zeros at padding positions: the
 * operand format of the tensor-core dgrad / wgrad kernels (pass it as their `workspace` / `dz_flat`). */
int kc_tc_dz_flat(const kc_desc* d, const float* dz, void* dz_flat, void* stream);
/* fp32 reference-layout weights -> bf16 UMMA-canonical K-block images (once per optimizer step).  Either output
 * pointer may be NULL to skip that image. */
int kc_tc_pack_weights(const kc_desc* d, const float* w_base, const float* w_basis, void* packed_fwd,
                       void* packed_dgrad, void* stream);
/* `phi_out` (NULL, or kc_tc_bytes(d, 4) bytes): the bf16 basis / base-activation rows the kernel evaluates anyway are
 * also written in the plane-major layout the weight-gradient kernel consumes, so the backward pass does not have to
 * re-evaluate them (the reference keeps the expanded basis tensor alive for autograd in the same way). */
int kc_conv_fwd_tc(const kc_desc* d, const float* x_base, const float* x_basis, const void* packed_fwd,
                   const float* beta, float* z, void* phi_out, void* stream);
int kc_conv_dgrad_tc(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                     const void* packed_dgrad, const float* beta, float* dx_base, float* dx_basis, float* dbeta,
                     void* workspace, void* stream);
/* `dz_flat` = buffer filled by kc_tc_dz_flat; `phi` = rows saved by kc_conv_fwd_tc, or NULL to re-evaluate them from x;
 * `workspace` = kc_tc_bytes(d, phi ? 5 : 3) bytes. */
int kc_conv_wgrad_tc(const kc_desc* d, const void* dz_flat, const float* x_base, const float* x_basis,
                     const float* beta, const void* phi, float* dw_base, float* dw_basis, void* workspace,
                     void* stream);

/* Tile geometry the tensor-core kernels choose for a shape (host-side, needs no GPU; for integrators and tests).
 * which 0 = forward, 1 = dgrad; out[8] = {M sub-tiles per CTA, N tile, number of N tiles, A ring depth, taps per weight
 * stage, weight ring depth, M tiles, dynamic shared memory in bytes}. */
int kc_tc_geometry(const kc_desc* d, int which, long long* out);

/* Backward of (InstanceNorm -> output activation) fused with the conversion of dz to the bf16 flat operand layout:
 * equivalent to kc_norm_act_bwd followed by kc_tc_dz_flat, without the fp32 dz round trip through HBM (reads dy and z once,
 * writes 2 bytes per element).  `conv` = descriptor of the convolution that produced z (its geometry fixes the flat layout),
 * `dz_flat` = kc_tc_bytes(conv, 2) bytes, `partials` = 3*n*c floats, `dalpha` (1 float) may be NULL.  Covers instance norm
 * without affine parameters on stride-1 layers; kc_norm_bwd_dz_flat_supported() == 0 means "use the two-step path".
 * [backward of kan_layers.py:241-243, gram:187, cheby:98] */
int kc_norm_bwd_dz_flat_supported(const kc_desc* conv, const kc_norm_desc* d);
int kc_norm_bwd_dz_flat(const kc_desc* conv, const kc_norm_desc* d, const float* dy, const float* z, const float* mean,
                        const float* rstd, const float* alpha, void* dz_flat, float* dalpha, float* partials, void* stream);

/* Self-test of the tcgen05 shared-memory descriptor conventions this library relies on: runs a 128xNx64 bf16 GEMM
 * through the same UMMA helpers as the convolution kernels and returns the max abs error against a CUDA-core
 * evaluation of the same product (expected < 1e-2).  mode 0 = K-major A/B, 1 = MN-major A/B. */
int kc_tc_selftest(int mode, float* max_abs_err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KANCONV_H_ */
