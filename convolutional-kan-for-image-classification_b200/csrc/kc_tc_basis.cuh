// kc_tc_basis.cuh - basis evaluation shared by the tensor-core kernels (forward producers, dgrad epilogues, phi pre-pass).
#pragma once
#include "kc_common.cuh"
#include "kc_umma.cuh"

namespace kc {

// ---- optional timeline trace: compiled in only with -DKANCONV_DEBUG (python -m kanconv_b200.build with KANCONV_DEBUG=1).
// kc_debug_trace*() point the per-TU device pointer at a buffer of 4 x 1024 clock stamps; one CTA records one stamp per
// pipeline event of representative threads.  In the product build stamp() is an empty inline function.
#ifdef KANCONV_DEBUG
#define KC_TRACE_DECL(sym) __device__ long long* sym = nullptr;
struct Tracer {
  long long* p; int n;
  __device__ Tracer(long long* t, int role, bool on) {
    p = (on && t != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0) ? t + role * 1024 : nullptr; n = 0;
  }
  __device__ __forceinline__ void stamp() { if (p != nullptr && n < 1024) p[n++] = clock64(); }
};
#define KC_TRACER(name, sym, role, on) Tracer name(sym, role, on)
#else
#define KC_TRACE_DECL(sym)
struct Tracer {
  __device__ __forceinline__ void stamp() {}
};
#define KC_TRACER(name, sym, role, on) Tracer name
#endif

// Base activation of the tensor-core paths: the result is rounded to bf16 (2^-9), so SiLU uses the fast exponential and
// reciprocal (5 instructions instead of ~30 for expf + IEEE division; the forward producers evaluate one per input element
// per N tile).  -Inf gives NaN and NaN propagates, like x * sigmoid(x) in the reference.
__device__ __forceinline__ float tc_act(int kind, float x) {
  if (kind == KC_ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));
  return kc_act(kind, x);
}

// ---- uniform cubic B-spline, closed form ---------------------------------------------------------
// The 4 non-zero weights land at j = i0-3 .. i0.  Branch-free: the 4 weights are packed into 64 bits and moved to their
// slots with clamped PTX shifts (shift amounts >= 64, including "negative" ones, yield 0), so independent evaluations
// can be interleaved by the compiler.
__device__ __forceinline__ unsigned long long shl64(unsigned long long v, int s) {
  unsigned long long r;
  asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(s));
  return r;
}
__device__ __forceinline__ unsigned long long shr64(unsigned long long v, int s) {
  unsigned long long r;
  asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(s));
  return r;
}
// true for NaN and +-Inf: the reference's Cox-de Boor recursion turns both into NaN on EVERY basis function
// ((x - t_i) / d * B with B = 0 gives 0 * Inf; kan_layers.py:212-233)
__device__ __forceinline__ bool tc_nonfinite(float x) { return !(fabsf(x) <= 3.402823466e38f); }

// 8 basis values of one input as packed bf16.  valid = false (padding position / channel beyond cin): all-zero row.
__device__ __forceinline__ uint4 cubic8(float x, float t0, float inv_h, int nintervals, bool valid) {
  const float u = (x - t0) * inv_h;
  const bool ok = valid && (u >= 0.0f) && (u < (float)nintervals);    // outside the knot span: all-zero row
  const float fi = floorf(u);
  const float f = u - fi;
  const int i0 = min(max((int)fi, 0), 15);
  const float s6 = 1.0f / 6.0f;
  const float w0 = fmaf(fmaf(fmaf(-s6, f, 0.5f), f, -0.5f), f, s6);
  const float w1 = fmaf(fmaf(0.5f, f, -1.0f) * f, f, 4.0f * s6);
  const float w2 = fmaf(fmaf(fmaf(-0.5f, f, 0.5f), f, 0.5f), f, s6);
  const float w3 = f * f * f * s6;
  unsigned long long v = (unsigned long long)pack_bf16(w0, w1) | ((unsigned long long)pack_bf16(w2, w3) << 32);
  v = ok ? v : 0ull;
  const int sh = 16 * (i0 - 3);
  unsigned long long lo = shl64(v, sh) | shr64(v, -sh);
  unsigned long long hi = shr64(v, 64 - sh) | shl64(v, sh - 64);
  if (valid && tc_nonfinite(x)) lo = hi = 0x7fc07fc07fc07fc0ull;      // NaN / Inf input: NaN on all eight (see above)
  return make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
}

// Basis value (and optionally derivative) for the tensor-core path: same formulas as kc_eval_basis, evaluated with fast
// intrinsics (ex2-based exp / tanh / sigmoid, Chebyshev polynomials by recurrence instead of cos(j acos c)); the results
// are rounded to bf16 anyway.  B-splines that are not the uniform cubic case go through the exact evaluator.
__device__ __forceinline__ float tc_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ float tc_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ inline void tc_eval_basis(const KcBasisCtx& B, float x, float* phi, float* dphi) {
  const int nb = B.nb;
  if (B.kind == KC_BASIS_RBF) {
    const float inv_den = __fdividef(1.0f, B.p[nb]);
    for (int j = 0; j < nb; ++j) {
      const float q = (x - B.p[j]) * inv_den;
      const float e = __expf(-(q * q));
      phi[j] = e;
      if (dphi) dphi[j] = e * (-2.0f * q * inv_den);
    }
  } else if (B.kind == KC_BASIS_CHEBY) {
    const float lo = -1.0f + 1e-7f, hi = 1.0f - 1e-7f;
    const float t = tc_tanh(x);
    float c = fminf(fmaxf(t, lo), hi);
    if (t != t) c = t;
    const float dc = (t < lo || t > hi) ? 0.0f : 1.0f - t * t;       // d c / d x
    float T0 = 1.0f, T1 = c, U0 = 0.0f, U1 = 1.0f;                   // T_j, U_{j-1}
    for (int j = 0; j < nb; ++j) {
      phi[j] = T0;
      if (dphi) dphi[j] = (float)j * U0 * dc;
      const float T2 = 2.0f * c * T1 - T0, U2 = 2.0f * c * U1 - U0;
      T0 = T1; T1 = T2; U0 = U1; U1 = U2;
    }
  } else if (B.kind == KC_BASIS_GRAM) {
    const bool pre = kc_gram_presquashed(B);
    const float t = pre ? x : tc_tanh(x), dt = pre ? 1.0f : 1.0f - t * t;
    float p0 = 1.0f, p1 = t, d0 = 0.0f, d1 = 1.0f;
    for (int i = 0; i < nb; ++i) {
      const float sg = tc_sigmoid(p0);
      phi[i] = p0 * sg;
      if (dphi) dphi[i] = sg * fmaf(p0, 1.0f - sg, 1.0f) * d0 * dt;
      const float b = (i + 1 < KC_MAX_BASIS) ? B.gbeta[i + 1] : 0.0f;   // p_{i+2} = t p_{i+1} - beta(i+1, i+2) p_i
      const float p2 = t * p1 - b * p0, d2 = p1 + t * d1 - b * d0;
      p0 = p1; p1 = p2; d0 = d1; d1 = d2;
    }
  } else if (B.kind == KC_BASIS_RECUR || B.kind == KC_BASIS_RECUR_DM) {
    const bool pre = B.p[0] != 0.0f;
    const float t = pre ? x : tc_tanh(x), dt = pre ? 1.0f : 1.0f - t * t;
    kc_recur_eval(B.p, nb, t, dt, phi, dphi, 1);
  } else {
    kc_eval_basis(B, x, phi, dphi, 1);
  }
}


}  // namespace kc
