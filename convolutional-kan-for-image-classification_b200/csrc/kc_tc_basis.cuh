// kc_tc_basis.cuh - basis evaluation shared by the tensor-core kernels (forward producers, dgrad epilogues, phi pre-pass).
#pragma once
#include "kc_common.cuh"

namespace kc {

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
    const float t = tc_tanh(x), dt = 1.0f - t * t;
    float p0 = 1.0f, p1 = t, d0 = 0.0f, d1 = 1.0f;
    for (int i = 0; i < nb; ++i) {
      const float sg = tc_sigmoid(p0);
      phi[i] = p0 * sg;
      if (dphi) dphi[i] = sg * fmaf(p0, 1.0f - sg, 1.0f) * d0 * dt;
      const float b = (i + 1 < KC_MAX_BASIS) ? B.gbeta[i + 1] : 0.0f;   // p_{i+2} = t p_{i+1} - beta(i+1, i+2) p_i
      const float p2 = t * p1 - b * p0, d2 = p1 + t * d1 - b * d0;
      p0 = p1; p1 = p2; d0 = d1; d1 = d2;
    }
  } else {
    kc_eval_basis(B, x, phi, dphi, 1);
  }
}


}  // namespace kc
