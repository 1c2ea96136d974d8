// kc_common.cuh - shared device code of libkanconv: basis functions (value + analytic derivative), base
// activations, error plumbing.  Math spec: SURVEY.md Appendix A; reference lines cited per function.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "kanconv.h"

#define KC_MAX_ORDER 6     // max B-spline order K handled by the local de Boor scheme
#define KC_MAX_DEGREE 15   // max polynomial degree (nb = D+1 <= KC_MAX_BASIS)

// ---------------------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------------------
void kc_set_error(const char* fmt, ...);
#define KC_FAIL(code, ...)        \
  do {                            \
    kc_set_error(__VA_ARGS__);    \
    return (code);                \
  } while (0)
#define KC_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
void kc_count_launch();   // per-process counter of kernels launched by this library (kc_launch_count in the ABI)
#define KC_LAUNCH_CHECK(name)                                                            \
  do {                                                                                   \
    kc_count_launch();                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

int kc_validate_desc(const kc_desc* d);   // common argument checks (kc_api.cu)
bool kc_knots_uniform_cubic(const kc_desc* d, float* t0, float* inv_h);   // kc_api.cu
int kc_sm_count();                         // SMs of the current device (kc_api.cu)
// kc_norm.cu: per-plane partial sums -> dgamma / dbeta / dalpha (any may be NULL), fixed summation order
int kc_norm_partials_to_params(const kc_norm_desc* d, const float* partials, float* dgamma, float* dbeta, float* dalpha, void* stream);
// kc_norm_cluster.cu: shared-memory-resident InstanceNorm forward for large planes; KC_ERR_UNSUPPORTED if the shape is not covered
int kc_instnorm_fwd_cluster(const kc_norm_desc* d, const float* z, const float* gamma, const float* beta, const float* alpha,
                            float* y, float* mean, float* rstd, void* stream);
// GRAM d/d beta_weights: fixed-order sum of the `nrows` partial rows behind dbeta[0 .. KC_MAX_BASIS) into it (kc_norm.cu)
int kc_dbeta_reduce(float* dbeta, long long nrows, void* stream);

// ---------------------------------------------------------------------------------------------------------
// Basis context: copied from kc_desc into shared memory at kernel start (dynamic indexing of knots).
// ---------------------------------------------------------------------------------------------------------
struct KcBasisCtx {
  int kind, nb, order, nparams;
  float p[KC_MAX_PARAMS];
  float inv_h;                       // B-spline: 1/(mean knot spacing) for the interval guess
  float gbeta[KC_MAX_BASIS];         // GRAM: beta(n, n+1) = coef_n * beta_weights[n]   (gram_kan_layers.py:150-153)
};

__host__ __device__ inline float kc_gram_coef(int n) {
  // ((m+n)(m-n) n^2) / (m^2 / (4 n^2 - 1)) with m = n+1
  float m = (float)(n + 1), fn = (float)n;
  return ((m + fn) * (m - fn) * fn * fn) / (m * m / (4.0f * fn * fn - 1.0f));
}

// Every thread of the block calls this; ends with __syncthreads().
__device__ inline void kc_load_basis_ctx(KcBasisCtx* B, const kc_desc& d, const float* __restrict__ beta_w) {
  int t = threadIdx.x + threadIdx.y * blockDim.x;
  int nt = blockDim.x * blockDim.y;
  if (t == 0) {
    B->kind = d.basis; B->nb = d.nb; B->order = d.order; B->nparams = d.nparams;
    float h = 1.0f;
    if (d.basis == KC_BASIS_BSPLINE && d.nparams > 1) h = (d.params[d.nparams - 1] - d.params[0]) / (float)(d.nparams - 1);
    B->inv_h = 1.0f / h;
  }
  for (int i = t; i < KC_MAX_PARAMS; i += nt) B->p[i] = (i < d.nparams) ? d.params[i] : 0.0f;
  for (int i = t; i < KC_MAX_BASIS; i += nt) {
    float v = 0.0f;
    if (d.basis == KC_BASIS_GRAM && beta_w != nullptr && i >= 1 && i <= d.order - 1) v = kc_gram_coef(i) * beta_w[i];
    B->gbeta[i] = v;
  }
  __syncthreads();
}

// Chebyshev: where clamp(tanh x, -1+1e-7, 1-1e-7) is active autograd MASKS the gradient (clamp backward is a select, not a
// product), so dx is exactly 0 there even when the incoming gradient is NaN / Inf (cheby_kan_layers.py:93-96).
__device__ __forceinline__ bool kc_cheby_clamped(float t) { return (t < -1.0f + 1e-7f) || (t > 1.0f - 1e-7f); }

// GRAM with dropout: the binding applies tanh and Dropout to the input itself and sets params[0] = 1 (nparams = 1)
__device__ __forceinline__ bool kc_gram_presquashed(const KcBasisCtx& B) { return B.nparams > 0 && B.p[0] != 0.0f; }

// ---------------------------------------------------------------------------------------------------------
// base activations (kan_layers.py:199; gram:173; fast:103)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float kc_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float kc_silu(float x) { return x * kc_sigmoid(x); }
__device__ __forceinline__ float kc_silu_grad(float x) {
  float s = kc_sigmoid(x);
  return s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float kc_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float kc_gelu_grad(float x) {
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  float pdf = 0.39894228040143268f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float kc_act(int kind, float x) {
  if (kind == KC_ACT_GELU) return kc_gelu(x);
  if (kind == KC_ACT_SILU) return kc_silu(x);
  return x;
}
__device__ __forceinline__ float kc_act_grad(int kind, float x) {
  if (kind == KC_ACT_GELU) return kc_gelu_grad(x);
  if (kind == KC_ACT_SILU) return kc_silu_grad(x);
  return 1.0f;
}

// ---------------------------------------------------------------------------------------------------------
// B-spline: local de Boor scheme on the K+1 basis functions that are non-zero at x.
// Equivalent to the reference's full Cox-de Boor recursion (kan_layers.py:209-233): the reference evaluates all
// G+2K order-0 indicators and recurses; all but K+1 entries are exact zeros, and the surviving terms are the
// products formed here.  Indicator semantics are half-open [t_i, t_{i+1}) on the actual fp32 knots.
// Returns the index j0 of the first non-zero basis function (may be negative; callers clip to [0, nb)),
// or INT_MIN/2 when x is outside the knot span (all-zero basis).  NaN input yields NaN values on all K+1 taps.
//   N[r]  = B^K_{j0+r}(x),  dN[r] = d/dx B^K_{j0+r}(x),  r = 0..K
// ---------------------------------------------------------------------------------------------------------
#define KC_BSPLINE_OUTSIDE (-1000000)

__device__ __forceinline__ int kc_bspline_local(const KcBasisCtx& B, float x, float* N, float* dN) {
  const int K = B.order;
  const int nk = B.nparams;          // G + 2K + 1 knots
  const float* t = B.p;
#pragma unroll
  for (int r = 0; r <= KC_MAX_ORDER; ++r) { N[r] = 0.0f; dN[r] = 0.0f; }
  if (!(fabsf(x) <= 3.402823466e38f)) x = __int_as_float(0x7fc00000);   // +-Inf: the reference's recursion gives 0 * Inf = NaN
  if (x != x) {                      // NaN propagates like in the reference (0 * NaN)
#pragma unroll
    for (int r = 0; r <= KC_MAX_ORDER; ++r) { N[r] = x; dN[r] = x; }
    return 0;
  }
  if (!(x >= t[0]) || !(x < t[nk - 1])) return KC_BSPLINE_OUTSIDE;
  int i0 = (int)floorf((x - t[0]) * B.inv_h);
  i0 = max(0, min(i0, nk - 2));
  while (i0 > 0 && x < t[i0]) --i0;
  while (i0 < nk - 2 && x >= t[i0 + 1]) ++i0;
  // triangular scheme; Nm holds order-(K-1) values for the derivative
  float left[KC_MAX_ORDER + 1], right[KC_MAX_ORDER + 1], Nm[KC_MAX_ORDER + 1];
#pragma unroll
  for (int r = 0; r <= KC_MAX_ORDER; ++r) { left[r] = 0.f; right[r] = 0.f; Nm[r] = 0.f; }
  N[0] = 1.0f;
#pragma unroll
  for (int j = 1; j <= KC_MAX_ORDER; ++j) {
    if (j <= K) {
      if (j == K) {
#pragma unroll
        for (int r = 0; r <= KC_MAX_ORDER; ++r) Nm[r] = N[r];
      }
      // knots outside the vector (only reachable for basis functions that get clipped) are extrapolated uniformly
      int il = i0 + 1 - j, ir = i0 + j;
      float tl = (il >= 0) ? t[il] : t[0] + (float)il / B.inv_h;
      float tr = (ir <= nk - 1) ? t[ir] : t[nk - 1] + (float)(ir - (nk - 1)) / B.inv_h;
      left[j] = x - tl;
      right[j] = tr - x;
      float saved = 0.0f;
#pragma unroll
      for (int r = 0; r < KC_MAX_ORDER; ++r) {
        if (r < j) {
          float den = right[r + 1] + left[j - r];
          float temp = N[r] / den;
          N[r] = saved + right[r + 1] * temp;
          saved = left[j - r] * temp;
        }
      }
      N[j] = saved;
    }
  }
  // derivative: dB^K_i = K * ( B^{K-1}_i/(t_{i+K}-t_i) - B^{K-1}_{i+1}/(t_{i+K+1}-t_{i+1}) ), i = i0-K+r.
  // Order-(K-1) functions non-zero at x are indices i0-K+1 .. i0, stored in Nm[0..K-1] (Nm[s] = B^{K-1}_{i0-K+1+s}).
  if (K >= 1) {
#pragma unroll
    for (int r = 0; r <= KC_MAX_ORDER; ++r) {
      if (r <= K) {
        float a = 0.0f, b = 0.0f;
        if (r >= 1) {      // B^{K-1}_{i} with i = i0-K+r  -> Nm[r-1]; denominator t_{i+K}-t_i = right[r] + left[K-r+1]
          a = Nm[r - 1] / (right[r] + left[K - r + 1]);
        }
        if (r <= K - 1) {  // B^{K-1}_{i+1} -> Nm[r]; denominator t_{i+K+1}-t_{i+1} = right[r+1] + left[K-r]
          b = Nm[r] / (right[r + 1] + left[K - r]);
        }
        dN[r] = (float)K * (a - b);
      }
    }
  }
  return i0 - K;
}

// Three-term recurrence polynomials p_j(t) and d p_j / dx = p_j'(t) * dt for j < nb (KC_BASIS_RECUR*; coefficient layout in
// kanconv.h).  Host + device so that the arithmetic can be checked on the CPU (tests/test_recur_host_cpu.py).
__host__ __device__ inline void kc_recur_eval(const float* p, int nb, float t, float dt, float* phi, float* dphi, int stride) {
  float p0 = p[1], d0 = 0.0f;                       // p_0 = c0
  float p1 = p[2] * t + p[3], d1 = p[2];            // p_1 = a1 t + b1
  phi[0] = p0;
  if (dphi) dphi[0] = 0.0f;
  if (nb > 1) {
    phi[stride] = p1;
    if (dphi) dphi[stride] = d1 * dt;
  }
  for (int i = 2; i < nb; ++i) {
    const float A = p[4 + 3 * (i - 2)], Bc = p[5 + 3 * (i - 2)], C = p[6 + 3 * (i - 2)];
    const float m = A * t + Bc;
    const float p2 = m * p1 + C * p0;
    const float d2 = A * p1 + m * d1 + C * d0;
    phi[i * stride] = p2;
    if (dphi) dphi[i * stride] = d2 * dt;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Generic evaluation into memory (shared or local): phi[j*stride], dphi[j*stride] for j < nb.
// dphi may be nullptr.  For GRAM, dphi is d/dx (through tanh); see kc_gram_dbeta for d/d beta.
// ---------------------------------------------------------------------------------------------------------
__device__ inline void kc_eval_basis(const KcBasisCtx& B, float x, float* phi, float* dphi, int stride) {
  const int nb = B.nb;
  if (B.kind == KC_BASIS_BSPLINE) {
    float N[KC_MAX_ORDER + 1], dN[KC_MAX_ORDER + 1];
    int j0 = kc_bspline_local(B, x, N, dN);
    for (int j = 0; j < nb; ++j) { phi[j * stride] = 0.0f; if (dphi) dphi[j * stride] = 0.0f; }
    if (j0 != KC_BSPLINE_OUTSIDE) {
      if (x != x) {
        for (int j = 0; j < nb; ++j) { phi[j * stride] = x; if (dphi) dphi[j * stride] = x; }
      } else {
#pragma unroll
        for (int r = 0; r <= KC_MAX_ORDER; ++r) {
          int j = j0 + r;
          if (r <= B.order && j >= 0 && j < nb) { phi[j * stride] = N[r]; if (dphi) dphi[j * stride] = dN[r]; }
        }
      }
    }
  } else if (B.kind == KC_BASIS_CHEBY) {
    // cheby_kan_layers.py:93-96: cos(d * acos(clamp(tanh x, -1+1e-7, 1-1e-7)))
    const float lo = -1.0f + 1e-7f, hi = 1.0f - 1e-7f;
    float t = tanhf(x);
    float c = fminf(fmaxf(t, lo), hi);
    bool clamped = (t < lo) || (t > hi);
    if (t != t) c = t;
    float th = acosf(c);
    float dth = clamped ? 0.0f : -(1.0f - t * t) / sqrtf(1.0f - c * c);   // d theta / dx
    for (int j = 0; j < nb; ++j) {
      float a = (float)j * th;
      phi[j * stride] = cosf(a);
      if (dphi) dphi[j * stride] = -(float)j * sinf(a) * dth;
    }
  } else if (B.kind == KC_BASIS_GRAM) {
    // gram_kan_layers.py:155-181: p0=1, p1=t, p_i = t p_{i-1} - beta(i-1,i) p_{i-2};  phi = SiLU(p)
    // params[0] != 0 ("pre-squashed"): the input already is t = dropout(tanh(x)) (gram_kan_layers.py:176-179), no tanh here
    const bool pre = kc_gram_presquashed(B);
    float t = pre ? x : tanhf(x);
    float dt = pre ? 1.0f : 1.0f - t * t;
    float p0 = 1.0f, p1 = t, d0 = 0.0f, d1 = 1.0f;
    phi[0] = kc_silu(1.0f);
    if (dphi) dphi[0] = 0.0f;
    if (nb > 1) {
      phi[stride] = kc_silu(t);
      if (dphi) dphi[stride] = kc_silu_grad(t) * dt;
    }
    for (int i = 2; i < nb; ++i) {
      float b = B.gbeta[i - 1];
      float p2 = t * p1 - b * p0;
      float d2 = p1 + t * d1 - b * d0;
      phi[i * stride] = kc_silu(p2);
      if (dphi) dphi[i * stride] = kc_silu_grad(p2) * d2 * dt;
      p0 = p1; p1 = p2; d0 = d1; d1 = d2;
    }
  } else if (B.kind == KC_BASIS_RECUR || B.kind == KC_BASIS_RECUR_DM) {
    // three-term recurrence families (kanconv.h, kc_basis_kind): hermite_kan_layers.py:127-150 and siblings, on t = tanh(x);
    // Legendre's min-max normalised input arrives "pre-squashed" (params[0] != 0), like GRAM's dropout path
    const bool pre = B.p[0] != 0.0f;
    const float t = pre ? x : tanhf(x);
    const float dt = pre ? 1.0f : 1.0f - t * t;
    kc_recur_eval(B.p, nb, t, dt, phi, dphi, stride);
  } else {
    // utils/utils.py:32-33: exp(-((u - g_j)/den)^2); params = grid[0..G-1], den
    float den = B.p[nb];
    for (int j = 0; j < nb; ++j) {
      float q = (x - B.p[j]) / den;
      float e = expf(-(q * q));
      phi[j * stride] = e;
      if (dphi) dphi[j * stride] = e * (-2.0f * q / den);
    }
  }
}

// GRAM only: sum_d dphi_d * SiLU'(p_d) * (d p_d / d beta_n) for n = 1..D-1, accumulated into acc[n].
// dphi_d = gradient arriving at Phi_{c,d}.  (autograd path through gram_kan_layers.py:150-170.)
__device__ inline void kc_gram_dbeta(const KcBasisCtx& B, float x, const float* g, int gstride, float* acc) {
  const int nb = B.nb;
  float t = kc_gram_presquashed(B) ? x : tanhf(x);
  for (int n = 1; n <= nb - 2; ++n) {
    // q_i = d p_i / d beta_n : q_i = t q_{i-1} - [i-1 == n] p_{i-2} - beta_{i-1} q_{i-2}
    float p0 = 1.0f, p1 = t, q0 = 0.0f, q1 = 0.0f, s = 0.0f;
    for (int i = 2; i < nb; ++i) {
      float b = B.gbeta[i - 1];
      float p2 = t * p1 - b * p0;
      float q2 = t * q1 - ((i - 1 == n) ? p0 : 0.0f) - b * q0;
      s += g[i * gstride] * kc_silu_grad(p2) * q2;
      p0 = p1; p1 = p2; q0 = q1; q1 = q2;
    }
    acc[n] += s * kc_gram_coef(n);
  }
}

// index of (channel c, basis j) inside the reference's w_basis inner dimension
__host__ __device__ __forceinline__ bool kc_degree_major(int kind) { return kind == KC_BASIS_GRAM || kind == KC_BASIS_RECUR_DM; }
__device__ __forceinline__ int kc_wbasis_index(int kind, int c, int j, int cin, int nb) {
  return kc_degree_major(kind) ? (j * cin + c) : (c * nb + j);
}

__device__ __forceinline__ float kc_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
