// kc_pool.cu - max pooling between KAN convolution stages (the nn.MaxPool2d(2, 2) of the reference's VGG stacks,
// models/kan_vgg.py "M" entries).  HBM-bound byte shuffling: the forward reads every input once and writes the maximum
// plus a one-byte window index; the backward is a deterministic gather (no atomics): one thread per INPUT element looks
// up the (at most ceil(k/s)^2) windows that cover it.  Comparison semantics follow ATen's max_pool2d: the first maximum
// in row-major window order wins, NaN propagates.
#include "kc_common.cuh"

namespace {

__global__ void __launch_bounds__(256) kc_maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             unsigned char* __restrict__ idx, long long planes, int h, int w, int k,
                                                             int s, int ho, int wo) {
  const long long total = planes * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo);
    const long long t = i / wo;
    const int oy = (int)(t % ho);
    const long long pl = t / ho;
    const float* xp = x + pl * h * w + (long long)(oy * s) * w + ox * s;
    float best = -INFINITY;
    int bi = 0;
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) {
        const float v = __ldg(xp + a * w + b);
        if (v > best || v != v) { best = v; bi = a * k + b; }
      }
    y[i] = best;
    idx[i] = (unsigned char)bi;
  }
}

// 2x2 / stride 2 with even width: one thread produces two adjacent outputs from two 16-byte loads.
__global__ void __launch_bounds__(256) kc_maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              unsigned char* __restrict__ idx, long long planes, int h, int w, int ho,
                                                              int wo) {
  const int wo2 = wo >> 1;
  const long long total = planes * ho * wo2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox2 = (int)(i % wo2);
    const long long t = i / wo2;
    const int oy = (int)(t % ho);
    const long long pl = t / ho;
    const float* xp = x + pl * h * w + (long long)(oy * 2) * w + ox2 * 4;
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(xp));
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(xp + w));
    float o[2];
    unsigned char ix[2];
    const float v[2][4] = {{r0.x, r0.y, r1.x, r1.y}, {r0.z, r0.w, r1.z, r1.w}};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float best = -INFINITY;
      int bi = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (v[j][q] > best || v[j][q] != v[j][q]) { best = v[j][q]; bi = q; }
      o[j] = best;
      ix[j] = (unsigned char)bi;
    }
    const long long ob = (pl * ho + oy) * wo + ox2 * 2;
    *reinterpret_cast<float2*>(y + ob) = make_float2(o[0], o[1]);
    *reinterpret_cast<uchar2*>(idx + ob) = make_uchar2(ix[0], ix[1]);
  }
}

__global__ void __launch_bounds__(256) kc_maxpool_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                             float* __restrict__ dx, long long planes, int h, int w, int k, int s,
                                                             int ho, int wo) {
  const long long total = planes * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const long long t = i / w;
    const int yy = (int)(t % h);
    const long long pl = t / h;
    const int oy1 = min(yy / s, ho - 1), ox1 = min(xx / s, wo - 1);
    const int oy0 = max(0, (yy - k + s) / s), ox0 = max(0, (xx - k + s) / s);     // ceil((yy - k + 1) / s) for yy-k+1 >= 0
    float acc = 0.0f;
    for (int oy = oy0; oy <= oy1; ++oy)
      for (int ox = ox0; ox <= ox1; ++ox) {
        const int a = yy - oy * s, b = xx - ox * s;
        if (a < 0 || a >= k || b < 0 || b >= k) continue;
        const long long o = (pl * ho + oy) * wo + ox;
        if ((int)__ldg(idx + o) == a * k + b) acc += __ldg(dy + o);
      }
    dx[i] = acc;
  }
}

// 2x2 / stride 2, even width and height: one thread scatters two adjacent outputs into two 16-byte stores.
__global__ void __launch_bounds__(256) kc_maxpool2_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                              float* __restrict__ dx, long long planes, int h, int w, int ho, int wo) {
  const int wo2 = wo >> 1;
  const long long total = planes * ho * wo2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox2 = (int)(i % wo2);
    const long long t = i / wo2;
    const int oy = (int)(t % ho);
    const long long pl = t / ho;
    const long long ob = (pl * ho + oy) * wo + ox2 * 2;
    const float2 g = __ldg(reinterpret_cast<const float2*>(dy + ob));
    const uchar2 ix = *reinterpret_cast<const uchar2*>(idx + ob);
    float* xp = dx + pl * h * w + (long long)(oy * 2) * w + ox2 * 4;
    *reinterpret_cast<float4*>(xp) = make_float4(ix.x == 0 ? g.x : 0.f, ix.x == 1 ? g.x : 0.f, ix.y == 0 ? g.y : 0.f, ix.y == 1 ? g.y : 0.f);
    *reinterpret_cast<float4*>(xp + w) = make_float4(ix.x == 2 ? g.x : 0.f, ix.x == 3 ? g.x : 0.f, ix.y == 2 ? g.y : 0.f, ix.y == 3 ? g.y : 0.f);
  }
}

int pool_check(long long planes, int h, int w, int k, int s, int ho, int wo) {
  if (planes < 0 || h < 1 || w < 1 || k < 1 || s < 1) KC_FAIL(KC_ERR_INVALID, "kc_maxpool2d: bad shape");
  if (k > 11) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_maxpool2d: window > 11 (index is one byte)");
  if (h < k || w < k || ho != (h - k) / s + 1 || wo != (w - k) / s + 1) KC_FAIL(KC_ERR_INVALID, "kc_maxpool2d: output size does not match floor((in - k) / s) + 1");
  return KC_OK;
}

int pool_blocks(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)kc_sm_count() * 32;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

bool fast2(const void* p0, const void* p1, const void* p2, int h, int w, int k, int s) {
  return k == 2 && s == 2 && (w % 4) == 0 && (h % 2) == 0 && (((uintptr_t)p0 | (uintptr_t)p1) % 16) == 0 && ((uintptr_t)p2 % 2) == 0;
}

}  // namespace

extern "C" int kc_maxpool2d_fwd(const float* x, float* y, unsigned char* idx, long long planes, int h, int w, int k, int s,
                                int ho, int wo, void* stream) {
  int rc = pool_check(planes, h, w, k, s, ho, wo);
  if (rc != KC_OK) return rc;
  if (!x || !y || !idx) KC_FAIL(KC_ERR_INVALID, "kc_maxpool2d_fwd: null pointer");
  if (planes == 0) return KC_OK;
  if (fast2(x, x, idx, h, w, k, s) && ((uintptr_t)y % 8) == 0) {
    kc_maxpool2_fwd_kernel<<<pool_blocks(planes * ho * (wo / 2)), 256, 0, (cudaStream_t)stream>>>(x, y, idx, planes, h, w, ho, wo);
    KC_LAUNCH_CHECK("kc_maxpool2_fwd_kernel");
  } else {
    kc_maxpool_fwd_kernel<<<pool_blocks(planes * ho * wo), 256, 0, (cudaStream_t)stream>>>(x, y, idx, planes, h, w, k, s, ho, wo);
    KC_LAUNCH_CHECK("kc_maxpool_fwd_kernel");
  }
  return KC_OK;
}

extern "C" int kc_maxpool2d_bwd(const float* dy, const unsigned char* idx, float* dx, long long planes, int h, int w, int k, int s,
                                int ho, int wo, void* stream) {
  int rc = pool_check(planes, h, w, k, s, ho, wo);
  if (rc != KC_OK) return rc;
  if (!dy || !idx || !dx) KC_FAIL(KC_ERR_INVALID, "kc_maxpool2d_bwd: null pointer");
  if (planes == 0) return KC_OK;
  if (fast2(dx, dx, idx, h, w, k, s) && ((uintptr_t)dy % 8) == 0) {
    kc_maxpool2_bwd_kernel<<<pool_blocks(planes * ho * (wo / 2)), 256, 0, (cudaStream_t)stream>>>(dy, idx, dx, planes, h, w, ho, wo);
    KC_LAUNCH_CHECK("kc_maxpool2_bwd_kernel");
  } else {
    kc_maxpool_bwd_kernel<<<pool_blocks(planes * (long long)h * w), 256, 0, (cudaStream_t)stream>>>(dy, idx, dx, planes, h, w, k, s, ho, wo);
    KC_LAUNCH_CHECK("kc_maxpool_bwd_kernel");
  }
  return KC_OK;
}
