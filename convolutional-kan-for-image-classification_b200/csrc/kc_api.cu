// kc_api.cu - C-ABI plumbing of libkanconv: version, per-thread error string, descriptor validation.
#include <stdarg.h>
#include <string.h>

#include "kc_common.cuh"

#include <atomic>

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}

void kc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long kc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void kc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* kc_last_error(void) { return g_err; }
extern "C" int kc_version(void) { return KC_ABI_VERSION; }

extern "C" int kc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  KC_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  KC_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return KC_OK;
}

// Uniform cubic B-spline on G + 2K + 1 = 12 knots (the reference's default grid_size 5 / spline_order 3): the tensor-core
// kernels then use the closed form of SURVEY Appendix A.2 with t0 and 1/h taken from the (fp32-rounded) knot vector.
bool kc_knots_uniform_cubic(const kc_desc* d, float* t0, float* inv_h) {
  if (d->basis != KC_BASIS_BSPLINE || d->order != 3 || d->nb != 8 || d->nparams != 12) return false;
  double h = ((double)d->params[11] - (double)d->params[0]) / 11.0;
  if (!(h > 0)) return false;
  for (int i = 0; i < 12; ++i) {
    double e = (double)d->params[0] + h * i - (double)d->params[i];
    if (e < 0) e = -e;
    if (e > 1e-5 * h) return false;
  }
  *t0 = d->params[0];
  *inv_h = (float)(1.0 / h);
  return true;
}

// Number of SMs of the current device (cached per device index); the persistent kernels launch one CTA per SM.
int kc_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    cached[dev] = sms;
  }
  return cached[dev];
}

int kc_validate_desc(const kc_desc* d) {
  if (!d) KC_FAIL(KC_ERR_INVALID, "null kc_desc");
  if (d->basis < KC_BASIS_BSPLINE || d->basis > KC_BASIS_RECUR_DM) KC_FAIL(KC_ERR_INVALID, "kc_desc: unknown basis kind %d", d->basis);
  if (d->act < KC_ACT_NONE || d->act > KC_ACT_SILU) KC_FAIL(KC_ERR_INVALID, "kc_desc: unknown activation kind %d", d->act);
  if (d->n <= 0 || d->cin <= 0 || d->h <= 0 || d->w <= 0 || d->cout <= 0)
    KC_FAIL(KC_ERR_INVALID, "kc_desc: n, cin, h, w, cout must be positive");
  if (d->kh <= 0 || d->kw <= 0 || d->stride_h <= 0 || d->stride_w <= 0 || d->dil_h <= 0 || d->dil_w <= 0 || d->pad_h < 0 || d->pad_w < 0)
    KC_FAIL(KC_ERR_INVALID, "kc_desc: bad kernel/stride/dilation/padding");
  int ho = (d->h + 2 * d->pad_h - d->dil_h * (d->kh - 1) - 1) / d->stride_h + 1;
  int wo = (d->w + 2 * d->pad_w - d->dil_w * (d->kw - 1) - 1) / d->stride_w + 1;
  if (ho != d->ho || wo != d->wo || ho <= 0 || wo <= 0)
    KC_FAIL(KC_ERR_INVALID, "kc_desc: ho/wo (%d,%d) inconsistent with geometry (%d,%d)", d->ho, d->wo, ho, wo);
  if (d->nb <= 0 || d->nb > KC_MAX_BASIS) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_desc: basis width %d outside [1,%d]", d->nb, KC_MAX_BASIS);
  if (d->nparams < 0 || d->nparams > KC_MAX_PARAMS) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_desc: nparams %d > %d", d->nparams, KC_MAX_PARAMS);
  if (d->basis == KC_BASIS_BSPLINE) {
    if (d->order < 0 || d->order > KC_MAX_ORDER) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_desc: spline order %d outside [0,%d]", d->order, KC_MAX_ORDER);
    if (d->nparams != d->nb + d->order + 1) KC_FAIL(KC_ERR_INVALID, "kc_desc: B-spline needs nb+order+1 = %d knots, got %d", d->nb + d->order + 1, d->nparams);
  }
  if ((d->basis == KC_BASIS_CHEBY || d->basis == KC_BASIS_GRAM) && d->nb != d->order + 1)
    KC_FAIL(KC_ERR_INVALID, "kc_desc: polynomial basis needs nb == degree+1");
  if (d->basis == KC_BASIS_CHEBY && d->act != KC_ACT_NONE) KC_FAIL(KC_ERR_INVALID, "kc_desc: Chebyshev layer has no base branch");
  if ((d->basis == KC_BASIS_RECUR || d->basis == KC_BASIS_RECUR_DM) && d->nparams != 4 + 3 * (d->nb > 2 ? d->nb - 2 : 0))
    KC_FAIL(KC_ERR_INVALID, "kc_desc: recurrence basis of width %d needs %d params, got %d", d->nb, 4 + 3 * (d->nb > 2 ? d->nb - 2 : 0), d->nparams);
  if (d->basis == KC_BASIS_RBF && d->nparams != d->nb + 1) KC_FAIL(KC_ERR_INVALID, "kc_desc: RBF needs nb grid points + denominator");
  if (d->x_batch_stride < (long long)d->cin * d->h * d->w) KC_FAIL(KC_ERR_INVALID, "kc_desc: x_batch_stride too small");
  if (d->z_batch_stride < (long long)d->cout * d->ho * d->wo) KC_FAIL(KC_ERR_INVALID, "kc_desc: z_batch_stride too small");
  return KC_OK;
}
