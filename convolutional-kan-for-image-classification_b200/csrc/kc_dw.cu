// kc_dw.cu - depthwise KAN convolution (HBM bound): `groups` single-channel layers in ONE launch.
//
// The reference builds the depthwise stage of MobileNetV2 as a grouped KAN convolution with groups == channels when
// `replace_depthwise=True` (models/kan_mobilenetv2.py:112-124) and then runs up to 960 Python iterations of
// `forward_kan` per layer (layers/kan_layers.py:249-258).  Per group the contraction is one channel x (nb + 1) basis rows
// x kh*kw taps - no GEMM to speak of, so this is CUDA-core work bounded by the activation traffic:
//   forward   read x once (+ halo), write z once                        4 * (C*H*W + C*Ho*Wo) bytes per image
//   dgrad     read dz (taps hit L1 / L2), read x, write dx              4 * (C*Ho*Wo + 2*C*H*W)
//   wgrad     read x, read dz (taps hit L1 / L2)                        4 * (C*H*W + C*Ho*Wo)
// The basis is evaluated once per input element (into shared memory in the forward, in registers in the backward kernels),
// not once per tap.  FP32 throughout; the weight gradient is reduced in a fixed order (per-block partial rows, then one
// pass over them): no atomics, bit-identical reruns.
#include "kc_common.cuh"

namespace {

constexpr int kDwThreads = 256;
constexpr int kDwMaxNb1 = 9;     // nb + 1 (basis rows + base branch) held in registers by the weight-gradient kernel
constexpr int kDwMaxT = 9;       // taps held in registers by the weight-gradient kernel

struct DwGeom {
  int C;                // channels = groups
  int th, tw;           // output tile
  int ih, iw;           // input tile incl. halo: (th - 1) * stride + (k - 1) * dil + 1
  int tiles_y, tiles_x;
  int ipb;              // forward: images per block (small planes)
  int nb1;              // nb + (has_base ? 1 : 0)
};

__device__ __forceinline__ KcBasisCtx* dw_ctx(unsigned char* smem) { return reinterpret_cast<KcBasisCtx*>(smem); }
__host__ __device__ inline size_t dw_ctx_bytes() { return ((sizeof(KcBasisCtx) + 15) / 16) * 16; }

// Output index of an input coordinate under one filter tap: nn = i + pad - tap * dil must be a non-negative multiple of the
// stride below the output extent.  S = 1 / 2: compile-time strides (integer division by a run-time stride costs ~25
// instructions, four of them per tap dominated the first version of the backward kernels); S = 0: any stride.
template <int S>
__device__ __forceinline__ bool dw_out_index(int nn, int stride, int limit, int& o) {
  if (nn < 0) return false;
  if (S == 1) {
    o = nn;
  } else if (S == 2) {
    if (nn & 1) return false;
    o = nn >> 1;
  } else {
    if (nn % stride != 0) return false;
    o = nn / stride;
  }
  return o < limit;
}

// ---- forward: one block = one output tile of `ipb` consecutive images of one channel ---------------------------------------------
// Large planes: 32 x 32 outputs per block (four per thread; the halo costs 1.13 basis evaluations per output at stride 1);
// small planes: the whole plane of several images, so that a block always has ~1000 outputs to amortise its prologue
// (basis context, this channel's filters) and its two barriers.
// NB1 = basis rows + base branch as a compile-time bound of the accumulation loop (0: run-time)
template <int NB1>
__global__ void __launch_bounds__(kDwThreads)
kc_dw_fwd_kernel(const kc_desc d, const DwGeom g, const float* __restrict__ x_base, const float* __restrict__ x_basis,
                 const float* __restrict__ w_base, const float* __restrict__ w_basis, float* __restrict__ z) {
  extern __shared__ __align__(16) unsigned char smem[];
  KcBasisCtx* B = dw_ctx(smem);
  const int T = d.kh * d.kw, nb = d.nb, nb1 = NB1 > 0 ? NB1 : g.nb1;
  const bool has_base = d.act != KC_ACT_NONE;
  float* wsm = reinterpret_cast<float*>(smem + dw_ctx_bytes());         // [nb1][T]
  float* phi = wsm + ((nb1 * T + 3) / 4) * 4;                           // [nb1][ipb * ih * iw]
  const int tile_elems = g.ih * g.iw, tile_outs = g.th * g.tw;
  const int c = blockIdx.y, n0 = blockIdx.z * g.ipb;
  const int nimg = (d.n - n0 < g.ipb) ? d.n - n0 : g.ipb;
  const int all_elems = g.ipb * tile_elems;                             // row pitch of phi
  const int ty = blockIdx.x / g.tiles_x, tx = blockIdx.x - ty * g.tiles_x;
  const int oy0 = ty * g.th, ox0 = tx * g.tw;
  const int iy0 = oy0 * d.stride_h - d.pad_h, ix0 = ox0 * d.stride_w - d.pad_w;
  kc_load_basis_ctx(B, d, nullptr);
  for (int i = threadIdx.x; i < nb1 * T; i += kDwThreads) {
    const int j = i / T, t = i - j * T;
    wsm[i] = (j < nb) ? w_basis[((long long)c * nb + j) * T + t] : w_base[(long long)c * T + t];
  }
  // basis rows of the input tiles (zero outside the image: the reference pads the EXPANDED tensor)
  for (int e = threadIdx.x; e < nimg * tile_elems; e += kDwThreads) {
    const int img = e / tile_elems, le = e - img * tile_elems;
    const int ly = le / g.iw, lx = le - ly * g.iw;
    const int iy = iy0 + ly, ix = ix0 + lx;
    if (iy >= 0 && iy < d.h && ix >= 0 && ix < d.w) {
      const long long off = (long long)(n0 + img) * d.x_batch_stride + (long long)c * d.h * d.w + (long long)iy * d.w + ix;
      kc_eval_basis(*B, x_basis[off], phi + e, nullptr, all_elems);
      if (has_base) phi[nb * all_elems + e] = kc_act(d.act, x_base[off]);
    } else {
      for (int j = 0; j < nb1; ++j) phi[j * all_elems + e] = 0.0f;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < nimg * tile_outs; o += kDwThreads) {
    const int img = o / tile_outs, lo = o - img * tile_outs;
    const int ly = lo / g.tw, lx = lo - ly * g.tw;
    const int oy = oy0 + ly, ox = ox0 + lx;
    if (oy >= d.ho || ox >= d.wo) continue;
    const float* ph = phi + img * tile_elems + (ly * d.stride_h) * g.iw + lx * d.stride_w;
    float acc = 0.0f;
    for (int r = 0; r < d.kh; ++r)
      for (int s = 0; s < d.kw; ++s) {
        const int e = r * d.dil_h * g.iw + s * d.dil_w;
        const int t = r * d.kw + s;
        if (NB1 > 0) {
#pragma unroll
          for (int j = 0; j < NB1; ++j) acc = fmaf(wsm[j * T + t], ph[j * all_elems + e], acc);
        } else {
          for (int j = 0; j < nb1; ++j) acc = fmaf(wsm[j * T + t], ph[j * all_elems + e], acc);
        }
      }
    z[(long long)(n0 + img) * d.z_batch_stride + (long long)c * d.ho * d.wo + (long long)oy * d.wo + ox] = acc;
  }
}

// ---- dgrad: one thread = one input element; block = a 256-element chunk of one (image, channel) plane --------------------------
template <int S>
__global__ void __launch_bounds__(kDwThreads)
kc_dw_dgrad_kernel(const kc_desc d, const DwGeom g, const float* __restrict__ dz, const float* __restrict__ x_base,
                   const float* __restrict__ x_basis, const float* __restrict__ w_base, const float* __restrict__ w_basis,
                   float* dx_base, float* dx_basis) {
  extern __shared__ __align__(16) unsigned char smem[];
  KcBasisCtx* B = dw_ctx(smem);
  float* wsm = reinterpret_cast<float*>(smem + dw_ctx_bytes());         // [T][nb1]: this channel's filters, tap-major
  const int T = d.kh * d.kw, nb = d.nb, nb1 = g.nb1;
  const bool has_base = d.act != KC_ACT_NONE;
  const int plane_id = blockIdx.x;                                      // n * C + c
  const int n = plane_id / g.C, c = plane_id - n * g.C;
  kc_load_basis_ctx(B, d, nullptr);
  for (int i = threadIdx.x; i < nb1 * T; i += blockDim.x) {
    const int t = i / nb1, j = i - t * nb1;
    wsm[i] = (j < nb) ? w_basis[((long long)c * nb + j) * T + t] : w_base[(long long)c * T + t];
  }
  __syncthreads();
  const int HW = d.h * d.w;
  const int p = blockIdx.y * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int iy = p / d.w, ix = p - iy * d.w;
  float acc[kDwMaxNb1];
#pragma unroll
  for (int j = 0; j < kDwMaxNb1; ++j) acc[j] = 0.0f;
  const float* dzp = dz + (long long)n * d.z_batch_stride + (long long)c * d.ho * d.wo;
  for (int r = 0; r < d.kh; ++r) {
    int oy;
    if (!dw_out_index<S>(iy + d.pad_h - r * d.dil_h, d.stride_h, d.ho, oy)) continue;
    for (int s = 0; s < d.kw; ++s) {
      int ox;
      if (!dw_out_index<S>(ix + d.pad_w - s * d.dil_w, d.stride_w, d.wo, ox)) continue;
      const float gz = dzp[oy * d.wo + ox];
      const float* wt = wsm + (r * d.kw + s) * nb1;
#pragma unroll
      for (int j = 0; j < kDwMaxNb1; ++j)
        if (j < nb1) acc[j] = fmaf(gz, wt[j], acc[j]);
    }
  }
  const long long off = (long long)n * d.x_batch_stride + (long long)c * HW + p;
  float phi[KC_MAX_BASIS], dphi[KC_MAX_BASIS];
  const float xs = x_basis[off];
  kc_eval_basis(*B, xs, phi, dphi, 1);
  float gs = 0.0f, ga = 0.0f;
#pragma unroll
  for (int j = 0; j < kDwMaxNb1; ++j) {
    if (j < nb) gs = fmaf(acc[j], dphi[j], gs);
    if (j == nb) ga = acc[j];
  }
  if (d.basis == KC_BASIS_CHEBY && kc_cheby_clamped(tanhf(xs))) gs = 0.0f;
  const float gb = has_base ? ga * kc_act_grad(d.act, x_base[off]) : 0.0f;
  if (dx_base == dx_basis) {
    dx_basis[off] = gs + gb;
  } else {
    dx_basis[off] = gs;
    if (has_base && dx_base != nullptr) dx_base[off] = gb;
  }
}

// ---- wgrad: block (split, channel) sums phi_j(x) * dz over its share of the channel's input positions ---------------------------
// partial[(split * C + c) * 81 + j * 9 + t]; j = nb is the base branch.
template <int NB1, int S>
__global__ void __launch_bounds__(kDwThreads)
kc_dw_wgrad_kernel(const kc_desc d, const DwGeom g, const float* __restrict__ dz, const float* __restrict__ x_base,
                   const float* __restrict__ x_basis, float* __restrict__ partial, int nsplit) {
  extern __shared__ __align__(16) unsigned char smem[];
  KcBasisCtx* B = dw_ctx(smem);
  float* red = reinterpret_cast<float*>(smem + dw_ctx_bytes());          // [warps][81]
  kc_load_basis_ctx(B, d, nullptr);
  const int nb = d.nb, nb1 = g.nb1;
  const bool has_base = d.act != KC_ACT_NONE;
  const int c = blockIdx.y, split = blockIdx.x;
  const int HW = d.h * d.w, total = d.n * HW;                            // < 2^31 (dw_geometry)
  const int per = (total + nsplit - 1) / nsplit;
  const int lo = split * per, hi = (lo + per < total) ? lo + per : total;
  const int T = d.kh * d.kw;
  float acc[NB1][kDwMaxT];
#pragma unroll
  for (int j = 0; j < NB1; ++j)
#pragma unroll
    for (int t = 0; t < kDwMaxT; ++t) acc[j][t] = 0.0f;
  for (int i = lo + threadIdx.x; i < hi; i += kDwThreads) {
    const int n = i / HW;
    const int p = i - n * HW;
    const int iy = p / d.w, ix = p - iy * d.w;
    const long long off = (long long)n * d.x_batch_stride + (long long)c * HW + p;
    float phi[KC_MAX_BASIS];
    kc_eval_basis(*B, x_basis[off], phi, nullptr, 1);
    float row[NB1];
#pragma unroll
    for (int j = 0; j < NB1; ++j) row[j] = (j < nb) ? phi[j] : 0.0f;
    if (has_base) {
      const float a = kc_act(d.act, x_base[off]);
#pragma unroll
      for (int j = 0; j < NB1; ++j)
        if (j == nb) row[j] = a;
    }
    const float* dzp = dz + (long long)n * d.z_batch_stride + (long long)c * d.ho * d.wo;
    int r = 0, s = 0;                                                    // tap t = r * kw + s without a division
#pragma unroll
    for (int t = 0; t < kDwMaxT; ++t) {
      float gz = 0.0f;
      if (t < T) {
        int oy, ox;
        if (dw_out_index<S>(iy + d.pad_h - r * d.dil_h, d.stride_h, d.ho, oy) &&
            dw_out_index<S>(ix + d.pad_w - s * d.dil_w, d.stride_w, d.wo, ox))
          gz = dzp[oy * d.wo + ox];
        if (++s == d.kw) { s = 0; ++r; }
      }
#pragma unroll
      for (int j = 0; j < NB1; ++j) acc[j][t] = fmaf(row[j], gz, acc[j][t]);
    }
  }
  // fixed-order block reduction: warp shuffles, then the warps' slots in order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NB1; ++j)
#pragma unroll
    for (int t = 0; t < kDwMaxT; ++t) {
      const float v = kc_warp_sum(acc[j][t]);
      if (lane == 0) red[warp * (kDwMaxNb1 * kDwMaxT) + j * kDwMaxT + t] = v;
    }
  __syncthreads();
  for (int e = threadIdx.x; e < NB1 * kDwMaxT; e += kDwThreads) {
    float sum = 0.0f;
    for (int w = 0; w < kDwThreads / 32; ++w) sum += red[w * (kDwMaxNb1 * kDwMaxT) + e];
    partial[((long long)split * g.C + c) * (kDwMaxNb1 * kDwMaxT) + e] = sum;
  }
  (void)nb1;
}

// partial rows -> dw_basis [C][nb][T], dw_base [C][T]; one thread per output element, splits added in order
__global__ void __launch_bounds__(kDwThreads)
kc_dw_wgrad_reduce_kernel(const kc_desc d, const DwGeom g, const float* __restrict__ partial, float* __restrict__ dw_base,
                          float* __restrict__ dw_basis, int nsplit) {
  const int T = d.kh * d.kw, nb = d.nb, nb1 = g.nb1;
  const long long total = (long long)g.C * nb1 * T;
  for (long long i = (long long)blockIdx.x * kDwThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kDwThreads) {
    const int c = (int)(i / (nb1 * T));
    const int rem = (int)(i - (long long)c * nb1 * T);
    const int j = rem / T, t = rem - j * T;
    float s = 0.0f;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[((long long)sp * g.C + c) * (kDwMaxNb1 * kDwMaxT) + j * kDwMaxT + t];
    if (j < nb) dw_basis[((long long)c * nb + j) * T + t] = s;
    else if (dw_base != nullptr) dw_base[(long long)c * T + t] = s;
  }
}

int dw_geometry(const kc_desc* d, int channels, DwGeom* g, size_t* fwd_smem) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  if (channels <= 0 || channels > 65535) KC_FAIL(KC_ERR_UNSUPPORTED, "depthwise path: channels %d outside [1, 65535]", channels);
  if (d->cin != 1 || d->cout != 1) KC_FAIL(KC_ERR_UNSUPPORTED, "depthwise path needs one input and one output channel per group");
  if (d->basis == KC_BASIS_GRAM) KC_FAIL(KC_ERR_UNSUPPORTED, "depthwise path: the GRAM family (shared beta_weights) is not covered");
  if (d->n > 65535) KC_FAIL(KC_ERR_UNSUPPORTED, "depthwise path: batch > 65535");
  if ((long long)d->n * d->h * d->w >= (1LL << 31) || (long long)d->n * channels >= (1LL << 31) ||
      (long long)d->h * d->w > 65535LL * 32)
    KC_FAIL(KC_ERR_UNSUPPORTED, "depthwise path: plane or batch too large for 32-bit indexing");
  const bool has_base = d->act != KC_ACT_NONE;
  g->C = channels;
  g->nb1 = d->nb + (has_base ? 1 : 0);
  if (g->nb1 > kDwMaxNb1 || d->kh * d->kw > kDwMaxT)
    KC_FAIL(KC_ERR_UNSUPPORTED, "depthwise path needs nb + 1 <= %d and kh * kw <= %d", kDwMaxNb1, kDwMaxT);
  if (d->x_batch_stride < (long long)channels * d->h * d->w || d->z_batch_stride < (long long)channels * d->ho * d->wo)
    KC_FAIL(KC_ERR_INVALID, "depthwise path: batch strides must cover all %d channels", channels);
  const int T = d->kh * d->kw;
  const size_t budget = 64 * 1024;           // three blocks per SM
  for (int tile = 32; tile >= 4; tile /= 2) {
    g->th = d->ho < tile ? d->ho : tile;
    g->tw = d->wo < tile ? d->wo : tile;
    g->ih = (g->th - 1) * d->stride_h + (d->kh - 1) * d->dil_h + 1;
    g->iw = (g->tw - 1) * d->stride_w + (d->kw - 1) * d->dil_w + 1;
    const size_t fixed = dw_ctx_bytes() + (size_t)(((g->nb1 * T + 3) / 4) * 4) * sizeof(float);
    const size_t per_img = (size_t)g->nb1 * g->ih * g->iw * sizeof(float);
    if (fixed + per_img <= budget) {
      g->tiles_y = (d->ho + g->th - 1) / g->th;
      g->tiles_x = (d->wo + g->tw - 1) / g->tw;
      int ipb = 1024 / (g->th * g->tw);                  // ~1000 outputs per block
      const int fit = (int)((budget - fixed) / per_img);
      if (ipb > fit) ipb = fit;
      if (ipb > d->n) ipb = d->n;
      if (ipb < 1) ipb = 1;
      g->ipb = ipb;
      *fwd_smem = fixed + (size_t)ipb * per_img;
      return KC_OK;
    }
  }
  KC_FAIL(KC_ERR_UNSUPPORTED, "depthwise path: input tile does not fit shared memory (dilation / stride too large)");
}

int dw_wgrad_splits(const kc_desc* d, int channels) {
  // enough blocks for ~4 waves, at least ~2048 input positions per block
  const long long total = (long long)d->n * d->h * d->w;
  long long want = ((long long)kc_sm_count() * 4 + channels - 1) / channels;
  long long cap = (total + 2047) / 2048;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

}  // namespace

extern "C" int kc_dwconv_supported(const kc_desc* d, int channels) {
  DwGeom g;
  size_t smem = 0;
  return dw_geometry(d, channels, &g, &smem) == KC_OK ? 1 : 0;
}

extern "C" int kc_dwconv_fwd_f32(const kc_desc* d, int channels, const float* x_base, const float* x_basis,
                                 const float* w_base, const float* w_basis, float* z, void* stream) {
  DwGeom g;
  size_t smem = 0;
  int rc = dw_geometry(d, channels, &g, &smem);
  if (rc != KC_OK) return rc;
  if (!x_basis || !w_basis || !z) KC_FAIL(KC_ERR_INVALID, "kc_dwconv_fwd_f32: null pointer");
  if (d->act != KC_ACT_NONE && (!x_base || !w_base)) KC_FAIL(KC_ERR_INVALID, "kc_dwconv_fwd_f32: base branch needs x_base and w_base");
  dim3 grid((unsigned)(g.tiles_y * g.tiles_x), (unsigned)channels, (unsigned)((d->n + g.ipb - 1) / g.ipb));
#define KC_DW_FWD(NB1)                                                                                                       \
  do {                                                                                                                       \
    KC_CUDA_CHECK(cudaFuncSetAttribute(kc_dw_fwd_kernel<NB1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    kc_dw_fwd_kernel<NB1><<<grid, kDwThreads, smem, (cudaStream_t)stream>>>(*d, g, x_base, x_basis, w_base, w_basis, z);     \
  } while (0)
  switch (g.nb1) {            // the widths of the reference's models and defaults; anything else takes the run-time loop
    case 4: KC_DW_FWD(4); break;      // Chebyshev degree 3
    case 5: KC_DW_FWD(5); break;      // degree-3 polynomial families with a base branch
    case 6: KC_DW_FWD(6); break;      // FastKAN grid 5 (the models) + base
    case 9: KC_DW_FWD(9); break;      // B-spline G + K = 8, FastKAN grid 8 + base
    default: KC_DW_FWD(0); break;
  }
#undef KC_DW_FWD
  KC_LAUNCH_CHECK("kc_dw_fwd_kernel");
  return KC_OK;
}

extern "C" int kc_dwconv_dgrad_f32(const kc_desc* d, int channels, const float* dz, const float* x_base, const float* x_basis,
                                   const float* w_base, const float* w_basis, float* dx_base, float* dx_basis, void* stream) {
  DwGeom g;
  size_t smem = 0;
  int rc = dw_geometry(d, channels, &g, &smem);
  if (rc != KC_OK) return rc;
  if (!dz || !x_basis || !w_basis || !dx_basis) KC_FAIL(KC_ERR_INVALID, "kc_dwconv_dgrad_f32: null pointer");
  if (d->act != KC_ACT_NONE && (!x_base || !w_base)) KC_FAIL(KC_ERR_INVALID, "kc_dwconv_dgrad_f32: base branch needs x_base and w_base");
  const int HW = d->h * d->w;
  const int threads = HW >= kDwThreads ? kDwThreads : ((HW + 31) / 32) * 32;      // small planes: one short block per plane
  dim3 grid((unsigned)(d->n * channels), (unsigned)((HW + threads - 1) / threads));
  const size_t dsmem = dw_ctx_bytes() + (size_t)g.nb1 * d->kh * d->kw * sizeof(float);
  const int S = (d->stride_h == d->stride_w && d->stride_h <= 2) ? d->stride_h : 0;
  if (S == 1) kc_dw_dgrad_kernel<1><<<grid, threads, dsmem, (cudaStream_t)stream>>>(*d, g, dz, x_base, x_basis, w_base, w_basis, dx_base, dx_basis);
  else if (S == 2) kc_dw_dgrad_kernel<2><<<grid, threads, dsmem, (cudaStream_t)stream>>>(*d, g, dz, x_base, x_basis, w_base, w_basis, dx_base, dx_basis);
  else kc_dw_dgrad_kernel<0><<<grid, threads, dsmem, (cudaStream_t)stream>>>(*d, g, dz, x_base, x_basis, w_base, w_basis, dx_base, dx_basis);
  KC_LAUNCH_CHECK("kc_dw_dgrad_kernel");
  return KC_OK;
}

extern "C" size_t kc_dwconv_wgrad_workspace_bytes(const kc_desc* d, int channels) {
  DwGeom g;
  size_t smem = 0;
  if (dw_geometry(d, channels, &g, &smem) != KC_OK) return 0;
  return (size_t)dw_wgrad_splits(d, channels) * channels * kDwMaxNb1 * kDwMaxT * sizeof(float);
}

extern "C" int kc_dwconv_wgrad_f32(const kc_desc* d, int channels, const float* dz, const float* x_base, const float* x_basis,
                                   float* dw_base, float* dw_basis, void* workspace, void* stream) {
  DwGeom g;
  size_t smem = 0;
  int rc = dw_geometry(d, channels, &g, &smem);
  if (rc != KC_OK) return rc;
  if (!dz || !x_basis || !dw_basis || !workspace) KC_FAIL(KC_ERR_INVALID, "kc_dwconv_wgrad_f32: null pointer");
  if (d->act != KC_ACT_NONE && (!x_base || !dw_base)) KC_FAIL(KC_ERR_INVALID, "kc_dwconv_wgrad_f32: base branch needs x_base and dw_base");
  const int nsplit = dw_wgrad_splits(d, channels);
  const size_t wsmem = dw_ctx_bytes() + (size_t)(kDwThreads / 32) * kDwMaxNb1 * kDwMaxT * sizeof(float);
  dim3 grid((unsigned)nsplit, (unsigned)channels);
  const int S = (d->stride_h == d->stride_w && d->stride_h <= 2) ? d->stride_h : 0;
#define KC_DW_WGRAD(NB1)                                                                                                     \
  do {                                                                                                                       \
    if (S == 1) kc_dw_wgrad_kernel<NB1, 1><<<grid, kDwThreads, wsmem, (cudaStream_t)stream>>>(*d, g, dz, x_base, x_basis, (float*)workspace, nsplit); \
    else if (S == 2) kc_dw_wgrad_kernel<NB1, 2><<<grid, kDwThreads, wsmem, (cudaStream_t)stream>>>(*d, g, dz, x_base, x_basis, (float*)workspace, nsplit); \
    else kc_dw_wgrad_kernel<NB1, 0><<<grid, kDwThreads, wsmem, (cudaStream_t)stream>>>(*d, g, dz, x_base, x_basis, (float*)workspace, nsplit); \
  } while (0)
  if (g.nb1 <= 4) KC_DW_WGRAD(4);             // accumulator rows held in registers: the next size up
  else if (g.nb1 <= 6) KC_DW_WGRAD(6);
  else KC_DW_WGRAD(9);
#undef KC_DW_WGRAD
  KC_LAUNCH_CHECK("kc_dw_wgrad_kernel");
  const long long total = (long long)channels * g.nb1 * d->kh * d->kw;
  long long blocks = (total + kDwThreads - 1) / kDwThreads;
  if (blocks > (long long)kc_sm_count() * 8) blocks = (long long)kc_sm_count() * 8;
  kc_dw_wgrad_reduce_kernel<<<(unsigned)blocks, kDwThreads, 0, (cudaStream_t)stream>>>(*d, g, (const float*)workspace, dw_base,
                                                                                      dw_basis, nsplit);
  KC_LAUNCH_CHECK("kc_dw_wgrad_reduce_kernel");
  return KC_OK;
}
