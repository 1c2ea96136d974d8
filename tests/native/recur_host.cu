// Test helper (NOT part of libkanconv): kc_recur_eval of csrc/kc_common.cuh - the very function the CUDA kernels call for the
// KC_BASIS_RECUR* families - compiled for the HOST, so that tests/test_layers_host_cpu.py can hold its arithmetic (values and
// derivatives, coefficient layout of kc_desc.params) to the oracle without a GPU.
#include <math.h>

#include "kc_common.cuh"

extern "C" int recur_eval_host(const float* params, int nb, const float* x, int n, float* phi, float* dphi) {
  for (int i = 0; i < n; ++i) {
    const bool pre = params[0] != 0.0f;
    const float t = pre ? x[i] : tanhf(x[i]);
    const float dt = pre ? 1.0f : 1.0f - t * t;
    kc_recur_eval(params, nb, t, dt, phi + (long long)i * nb, dphi + (long long)i * nb, 1);
  }
  return 0;
}
