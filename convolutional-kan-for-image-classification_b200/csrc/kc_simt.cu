// kc_simt.cu - FP32 CUDA-core (FFMA) kernels of the KAN convolution: forward, input gradient, weight gradient.
// General in kernel size / stride / dilation / padding and in the basis family; the basis expansion is evaluated
// on the fly into shared memory, the 8x-expanded tensor of the reference (kan_layers.py:236-239) never exists.
// This is the <=1e-5 parity path and the path for shapes the tensor-core kernels do not cover.
#include "kc_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTilePix = 64;    // output (fwd/wgrad) or input (dgrad) pixels per block
constexpr int kTileCout = 64;   // couts per block (fwd, wgrad)
constexpr int kMaxWB = KC_MAX_BASIS + 1;

__device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------------
// forward: z[p, co] = sum_{c, tap} ( sum_j Phi_j(x_basis[p@tap, c]) * Wb[co, c, j, tap] + act(x_base[p@tap, c]) * Wa[co, c, tap] )
// block = 64 output pixels x 64 couts, thread = 4 pixels x 4 couts, K-loop over (channel, tap-chunk).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
kc_fwd_simt_kernel(const __grid_constant__ kc_desc d, const float* __restrict__ x_base,
                   const float* __restrict__ x_basis, const float* __restrict__ w_base,
                   const float* __restrict__ w_basis, const float* __restrict__ beta_w, float* __restrict__ z,
                   int tch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KcBasisCtx* B = reinterpret_cast<KcBasisCtx*>(smem_raw);
  int* pix_n = reinterpret_cast<int*>(smem_raw + ((sizeof(KcBasisCtx) + 15) / 16) * 16);
  int* pix_oy = pix_n + kTilePix;
  int* pix_ox = pix_oy + kTilePix;
  float* A_s = reinterpret_cast<float*>(pix_ox + kTilePix);
  const bool has_base = d.act != KC_ACT_NONE;
  const int nb = d.nb, WB = nb + (has_base ? 1 : 0);
  float* W_s = A_s + tch * WB * kTilePix;

  kc_load_basis_ctx(B, d, beta_w);

  const int tid = threadIdx.x;
  const int T = d.kh * d.kw;
  const int HoWo = d.ho * d.wo;
  const long long P = (long long)d.n * HoWo;
  const long long p0 = (long long)blockIdx.x * kTilePix;
  const int co0 = blockIdx.y * kTileCout;
  if (tid < kTilePix) {
    long long p = p0 + tid;
    if (p < P) {
      int n = (int)(p / HoWo);
      int rem = (int)(p - (long long)n * HoWo);
      pix_n[tid] = n; pix_oy[tid] = rem / d.wo; pix_ox[tid] = rem % d.wo;
    } else {
      pix_n[tid] = -1; pix_oy[tid] = 0; pix_ox[tid] = 0;
    }
  }
  __syncthreads();

  const int tp = tid & 15, tc = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int HW = d.h * d.w;
  const int nbw = d.cin * nb;      // inner dim of w_basis
  for (int c = 0; c < d.cin; ++c) {
    // two-level accumulation: the (nb + 1) * T products of one input channel are summed in `part`, then added to `acc`.  With a
    // single fp32 accumulator over K = cin * (nb + 1) * T terms (41 472 for the 512-channel layers) the rounding error grows
    // like sqrt(K) and the layer output was 8e-6 off (the reference's oneDNN convolution: 1e-6); blocked, it stays ~2e-6.
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.0f;
    for (int t0 = 0; t0 < T; t0 += tch) {
      const int ntc = min(tch, T - t0);
      // ---- stage A: basis + base activation of the tap-shifted input pixel -------------------------------
      for (int it = tid; it < ntc * kTilePix; it += kThreads) {
        int tl = it / kTilePix, p = it % kTilePix;
        int tap = t0 + tl, r = tap / d.kw, s = tap % d.kw;
        int n = pix_n[p];
        int iy = pix_oy[p] * d.stride_h - d.pad_h + r * d.dil_h;
        int ix = pix_ox[p] * d.stride_w - d.pad_w + s * d.dil_w;
        float* dst = A_s + (tl * WB) * kTilePix + p;
        if (n >= 0 && iy >= 0 && iy < d.h && ix >= 0 && ix < d.w) {
          long long off = (long long)n * d.x_batch_stride + (long long)c * HW + iy * d.w + ix;
          kc_eval_basis(*B, x_basis[off], dst, nullptr, kTilePix);
          if (has_base) dst[nb * kTilePix] = kc_act(d.act, x_base[off]);
        } else {   // zero padding lives in basis space (the reference pads the EXPANDED tensor)
          for (int j = 0; j < WB; ++j) dst[j * kTilePix] = 0.0f;
        }
      }
      // ---- stage W ---------------------------------------------------------------------------------------
      for (int it = tid; it < ntc * WB * kTileCout; it += kThreads) {
        int co = it % kTileCout, rest = it / kTileCout;
        int j = rest % WB, tl = rest / WB;
        int tap = t0 + tl, cog = co0 + co;
        float v = 0.0f;
        if (cog < d.cout) {
          if (j < nb) v = w_basis[((long long)cog * nbw + kc_wbasis_index(d.basis, c, j, d.cin, nb)) * T + tap];
          else v = w_base[((long long)cog * d.cin + c) * T + tap];
        }
        W_s[(tl * WB + j) * kTileCout + co] = v;
      }
      __syncthreads();
      // ---- FMA -------------------------------------------------------------------------------------------
      const int rows = ntc * WB;
      for (int rj = 0; rj < rows; ++rj) {
        float4 a = *reinterpret_cast<const float4*>(A_s + rj * kTilePix + tp * 4);
        float4 w = *reinterpret_cast<const float4*>(W_s + rj * kTileCout + tc * 4);
        float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], wv[j], part[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int p = tp * 4 + i;
    int n = pix_n[p];
    if (n < 0) continue;
    long long base = (long long)n * d.z_batch_stride + pix_oy[p] * d.wo + pix_ox[p];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = co0 + tc * 4 + j;
      if (co < d.cout) z[base + (long long)co * HoWo] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// dgrad: dPhi[p_in, c, j] = sum_{co, tap} dz[p_out(p_in, tap), co] * W[co, c, j, tap];
//        dx = sum_j dPhi_j * Phi'_j(x_basis) + dA * act'(x_base)      (dPhi never leaves registers)
// block = 64 input pixels x 4 channels; thread = (pixel, channel) with nb+1 accumulators.
// ---------------------------------------------------------------------------------------------------------
constexpr int kDgCh = 4;       // channels per block
constexpr int kDgCo = 16;      // couts per K step
constexpr int kDgWBP = 20;     // padded WB (float4-aligned rows)

__global__ void __launch_bounds__(kThreads)
kc_dgrad_simt_kernel(const __grid_constant__ kc_desc d, const float* __restrict__ dz,
                     const float* __restrict__ x_base, const float* __restrict__ x_basis,
                     const float* __restrict__ w_base, const float* __restrict__ w_basis,
                     const float* __restrict__ beta_w, float* __restrict__ dx_base, float* __restrict__ dx_basis,
                     float* __restrict__ dbeta, int tch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KcBasisCtx* B = reinterpret_cast<KcBasisCtx*>(smem_raw);
  int* pix_n = reinterpret_cast<int*>(smem_raw + ((sizeof(KcBasisCtx) + 15) / 16) * 16);
  int* pix_iy = pix_n + kTilePix;
  int* pix_ix = pix_iy + kTilePix;
  float* red = reinterpret_cast<float*>(pix_ix + kTilePix);          // [warps][KC_MAX_BASIS] floats for the dbeta reduction
  float* dz_s = red + (kThreads / 32) * KC_MAX_BASIS;                 // [tch][kDgCo][64]
  float* W_s = dz_s + tch * kDgCo * kTilePix;                         // [tch][kDgCo][kDgCh][kDgWBP]
  const bool has_base = d.act != KC_ACT_NONE;
  const int nb = d.nb, WB = nb + (has_base ? 1 : 0);

  kc_load_basis_ctx(B, d, beta_w);
  const int tid = threadIdx.x;
  const int T = d.kh * d.kw, HW = d.h * d.w, HoWo = d.ho * d.wo;
  const long long P = (long long)d.n * HW;
  const long long p0 = (long long)blockIdx.x * kTilePix;
  const int c0 = blockIdx.y * kDgCh;
  if (tid < kTilePix) {
    long long p = p0 + tid;
    if (p < P) {
      int n = (int)(p / HW);
      int rem = (int)(p - (long long)n * HW);
      pix_n[tid] = n; pix_iy[tid] = rem / d.w; pix_ix[tid] = rem % d.w;
    } else {
      pix_n[tid] = -1; pix_iy[tid] = 0; pix_ix[tid] = 0;
    }
  }
  if (tid < KC_MAX_BASIS) red[tid] = 0.0f;
  __syncthreads();

  const int p = tid & 63, cq = tid >> 6;
  float acc[kDgWBP];
#pragma unroll
  for (int j = 0; j < kDgWBP; ++j) acc[j] = 0.0f;
  const int nbw = d.cin * nb;

  for (int cob = 0; cob < d.cout; cob += kDgCo) {
    for (int t0 = 0; t0 < T; t0 += tch) {
      const int ntc = min(tch, T - t0);
      for (int it = tid; it < ntc * kDgCo * kTilePix; it += kThreads) {
        int pp = it % kTilePix, rest = it / kTilePix;
        int co = rest % kDgCo, tl = rest / kDgCo;
        int tap = t0 + tl, r = tap / d.kw, s = tap % d.kw;
        int n = pix_n[pp], cog = cob + co;
        float v = 0.0f;
        if (n >= 0 && cog < d.cout) {
          int ny = pix_iy[pp] + d.pad_h - r * d.dil_h, nx = pix_ix[pp] + d.pad_w - s * d.dil_w;
          if (ny >= 0 && nx >= 0 && ny % d.stride_h == 0 && nx % d.stride_w == 0) {
            int oy = ny / d.stride_h, ox = nx / d.stride_w;
            if (oy < d.ho && ox < d.wo) v = dz[(long long)n * d.z_batch_stride + (long long)cog * HoWo + oy * d.wo + ox];
          }
        }
        dz_s[(tl * kDgCo + co) * kTilePix + pp] = v;
      }
      for (int it = tid; it < ntc * kDgCo * kDgCh * kDgWBP; it += kThreads) {
        int j = it % kDgWBP, rest = it / kDgWBP;
        int ch = rest % kDgCh; rest /= kDgCh;
        int co = rest % kDgCo, tl = rest / kDgCo;
        int tap = t0 + tl, cog = cob + co, c = c0 + ch;
        float v = 0.0f;
        if (cog < d.cout && c < d.cin && j < WB) {
          if (j < nb) v = w_basis[((long long)cog * nbw + kc_wbasis_index(d.basis, c, j, d.cin, nb)) * T + tap];
          else v = w_base[((long long)cog * d.cin + c) * T + tap];
        }
        W_s[it] = v;
      }
      __syncthreads();
      for (int k = 0; k < ntc * kDgCo; ++k) {
        float g = dz_s[k * kTilePix + p];
        const float4* wr = reinterpret_cast<const float4*>(W_s + (k * kDgCh + cq) * kDgWBP);
#pragma unroll
        for (int q = 0; q < kDgWBP / 4; ++q) {
          float4 w = wr[q];
          acc[q * 4 + 0] = fmaf(g, w.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(g, w.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(g, w.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(g, w.w, acc[q * 4 + 3]);
        }
      }
      __syncthreads();
    }
  }
  // ---- epilogue: multiply by the analytic basis derivative ----------------------------------------------
  const int c = c0 + cq, n = pix_n[p];
  float dbl[KC_MAX_BASIS];
#pragma unroll
  for (int j = 0; j < KC_MAX_BASIS; ++j) dbl[j] = 0.0f;
  if (n >= 0 && c < d.cin) {
    long long off = (long long)n * d.x_batch_stride + (long long)c * HW + pix_iy[p] * d.w + pix_ix[p];
    float phi[KC_MAX_BASIS], dphi[KC_MAX_BASIS];
    float xb = x_basis[off];
    kc_eval_basis(*B, xb, phi, dphi, 1);
    float gs = 0.0f;
#pragma unroll
    for (int j = 0; j < KC_MAX_BASIS; ++j)
      if (j < nb) gs = fmaf(acc[j], dphi[j], gs);
    if (d.basis == KC_BASIS_CHEBY && kc_cheby_clamped(tanhf(xb))) gs = 0.0f;
    float gb = 0.0f;
    if (has_base) {
      float ga = 0.0f;
#pragma unroll
      for (int j = 0; j < kDgWBP; ++j)
        if (j == nb) ga = acc[j];
      gb = ga * kc_act_grad(d.act, x_base[off]);
    }
    if (dx_base == dx_basis) {
      dx_basis[off] = gs + gb;
    } else {
      dx_basis[off] = gs;
      if (has_base && dx_base != nullptr) dx_base[off] = gb;
    }
    if (d.basis == KC_BASIS_GRAM && dbeta != nullptr) {
      float g[KC_MAX_BASIS];
#pragma unroll
      for (int j = 0; j < KC_MAX_BASIS; ++j) g[j] = acc[j];
      kc_gram_dbeta(*B, xb, g, 1, dbl);
    }
  }
  if (d.basis == KC_BASIS_GRAM && dbeta != nullptr) {
    // deterministic: warp sums -> per-warp slots -> fixed-order block sum -> this block's row of the partials buffer
    // (kc_dbeta_reduce_kernel adds the rows in fixed order; no atomics anywhere)
#pragma unroll
    for (int nn = 0; nn < KC_MAX_BASIS; ++nn) {
      const float v = kc_warp_sum(dbl[nn]);
      if ((tid & 31) == 0) red[(tid >> 5) * KC_MAX_BASIS + nn] = v;
    }
    __syncthreads();
    if (tid < KC_MAX_BASIS) {
      float v = 0.0f;
      for (int w = 0; w < kThreads / 32; ++w) v += red[w * KC_MAX_BASIS + tid];
      const long long row = (long long)blockIdx.y * gridDim.x + blockIdx.x;
      dbeta[(1 + row) * KC_MAX_BASIS + tid] = (tid >= 1 && tid <= nb - 2) ? v : 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: dW[co, c, j, tap] = sum_p dz[p, co] * Phi_j(x_basis[p@tap, c])   (+ base row j = nb)
// block = 64 couts x one channel x one tap-chunk x one pixel split; partial sums go to the workspace
// [split][cout][cin][T][WB] and are reduced in fixed order by kc_wgrad_reduce_kernel (deterministic).
// ---------------------------------------------------------------------------------------------------------
constexpr int kWgPix = 32;       // pixels per K step
constexpr int kWgMaxAcc = 40;    // accumulators per thread (pairs per quarter, multiple of 4)
constexpr int kWgDzPitch = kTileCout + 1;

__global__ void __launch_bounds__(kThreads)
kc_wgrad_simt_kernel(const __grid_constant__ kc_desc d, const float* __restrict__ dz,
                     const float* __restrict__ x_base, const float* __restrict__ x_basis,
                     const float* __restrict__ beta_w, float* __restrict__ ws, int tch, int nsplit,
                     int pix_per_split) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KcBasisCtx* B = reinterpret_cast<KcBasisCtx*>(smem_raw);
  float* dz_s = reinterpret_cast<float*>(smem_raw + ((sizeof(KcBasisCtx) + 15) / 16) * 16);   // [32][65]
  const bool has_base = d.act != KC_ACT_NONE;
  const int nb = d.nb, WB = nb + (has_base ? 1 : 0);
  const int T = d.kh * d.kw, HW = d.h * d.w, HoWo = d.ho * d.wo;
  const int ntchunks = ceil_div(T, tch);
  const int c = blockIdx.y / ntchunks, t0 = (blockIdx.y % ntchunks) * tch;
  const int ntc = min(tch, T - t0);
  const int npairs = ntc * WB;
  const int ppq = ((ceil_div(npairs, 4) + 3) / 4) * 4;     // pairs per quarter, multiple of 4, <= kWgMaxAcc
  const int pitch = ppq * 4;
  float* Phi_s = dz_s + kWgPix * kWgDzPitch;                 // [32][pitch]
  const int co0 = blockIdx.z * kTileCout;
  const int split = blockIdx.x;

  kc_load_basis_ctx(B, d, beta_w);
  const int tid = threadIdx.x;
  const int co = tid & 63, q = tid >> 6;
  float acc[kWgMaxAcc];
#pragma unroll
  for (int i = 0; i < kWgMaxAcc; ++i) acc[i] = 0.0f;

  const long long P = (long long)d.n * HoWo;
  const long long pbeg = (long long)split * pix_per_split;
  const long long pend = min(P, pbeg + pix_per_split);
  for (long long pb = pbeg; pb < pend; pb += kWgPix) {
    const int np = (int)min((long long)kWgPix, pend - pb);
    for (int it = tid; it < kWgPix * kTileCout; it += kThreads) {
      int pl = it % kWgPix, cc = it / kWgPix;
      float v = 0.0f;
      if (pl < np && co0 + cc < d.cout) {
        long long p = pb + pl;
        int n = (int)(p / HoWo);
        int rem = (int)(p - (long long)n * HoWo);
        v = dz[(long long)n * d.z_batch_stride + (long long)(co0 + cc) * HoWo + rem];
      }
      dz_s[pl * kWgDzPitch + cc] = v;
    }
    for (int it = tid; it < kWgPix * ntc; it += kThreads) {
      int pl = it % kWgPix, tl = it / kWgPix;
      float* dst = Phi_s + pl * pitch + tl * WB;
      bool ok = false;
      if (pl < np) {
        long long p = pb + pl;
        int n = (int)(p / HoWo);
        int rem = (int)(p - (long long)n * HoWo);
        int oy = rem / d.wo, ox = rem % d.wo;
        int tap = t0 + tl, r = tap / d.kw, s = tap % d.kw;
        int iy = oy * d.stride_h - d.pad_h + r * d.dil_h, ix = ox * d.stride_w - d.pad_w + s * d.dil_w;
        if (iy >= 0 && iy < d.h && ix >= 0 && ix < d.w) {
          long long off = (long long)n * d.x_batch_stride + (long long)c * HW + iy * d.w + ix;
          kc_eval_basis(*B, x_basis[off], dst, nullptr, 1);
          if (has_base) dst[nb] = kc_act(d.act, x_base[off]);
          ok = true;
        }
      }
      if (!ok)
        for (int j = 0; j < WB; ++j) dst[j] = 0.0f;
    }
    // zero the padding columns once per step (cheap; keeps the float4 reads defined)
    for (int it = tid; it < kWgPix * (pitch - npairs); it += kThreads) {
      int pl = it / (pitch - npairs), k = it % (pitch - npairs);
      Phi_s[pl * pitch + npairs + k] = 0.0f;
    }
    __syncthreads();
    for (int pl = 0; pl < kWgPix; ++pl) {
      float g = dz_s[pl * kWgDzPitch + co];
      const float4* pr = reinterpret_cast<const float4*>(Phi_s + pl * pitch + q * ppq);
#pragma unroll
      for (int i = 0; i < kWgMaxAcc / 4; ++i) {
        if (i * 4 < ppq) {
          float4 f = pr[i];
          acc[i * 4 + 0] = fmaf(g, f.x, acc[i * 4 + 0]);
          acc[i * 4 + 1] = fmaf(g, f.y, acc[i * 4 + 1]);
          acc[i * 4 + 2] = fmaf(g, f.z, acc[i * 4 + 2]);
          acc[i * 4 + 3] = fmaf(g, f.w, acc[i * 4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  if (co0 + co < d.cout) {
    // ws[split][co][c][tap][WB]
    float* dst = ws + ((((long long)split * d.cout + (co0 + co)) * d.cin + c) * T + t0) * WB;
#pragma unroll
    for (int i = 0; i < kWgMaxAcc; ++i) {
      int pair = q * ppq + i;
      if (i < ppq && pair < npairs) dst[pair] = acc[i];
    }
  }
}

__global__ void __launch_bounds__(kThreads)
kc_wgrad_reduce_kernel(const __grid_constant__ kc_desc d, const float* __restrict__ ws, float* __restrict__ dw_base,
                       float* __restrict__ dw_basis, int nsplit) {
  const bool has_base = d.act != KC_ACT_NONE;
  const int nb = d.nb, WB = nb + (has_base ? 1 : 0), T = d.kh * d.kw;
  const long long total = (long long)d.cout * d.cin * T * WB;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < nsplit; ++k) s += ws[(long long)k * total + i];
    int j = (int)(i % WB);
    long long rest = i / WB;
    int tap = (int)(rest % T); rest /= T;
    int c = (int)(rest % d.cin);
    int co = (int)(rest / d.cin);
    if (j < nb) dw_basis[((long long)co * d.cin * nb + kc_wbasis_index(d.basis, c, j, d.cin, nb)) * T + tap] = s;
    else dw_base[((long long)co * d.cin + c) * T + tap] = s;
  }
}

size_t basis_ctx_bytes() { return ((sizeof(KcBasisCtx) + 15) / 16) * 16; }

int pick_tch(int T, int per_tap_floats, size_t fixed_bytes, size_t budget) {
  int tch = T;
  while (tch > 1 && fixed_bytes + (size_t)tch * per_tap_floats * 4 > budget) --tch;
  return tch;
}

struct WgradPlan { int tch, ntchunks, nsplit, pix_per_split; size_t smem; };

WgradPlan plan_wgrad(const kc_desc* d) {
  WgradPlan pl;
  const int WB = d->nb + (d->act != KC_ACT_NONE ? 1 : 0), T = d->kh * d->kw;
  pl.tch = T;
  while (pl.tch > 1 && pl.tch * WB > kWgMaxAcc * 4) --pl.tch;
  pl.ntchunks = (T + pl.tch - 1) / pl.tch;
  const long long P = (long long)d->n * d->ho * d->wo;
  const long long base_blocks = (long long)d->cin * pl.ntchunks * ((d->cout + kTileCout - 1) / kTileCout);
  long long want = (4LL * kc_sm_count() + base_blocks - 1) / base_blocks;
  long long max_split = (P + 4 * kWgPix - 1) / (4 * kWgPix);     // at least 128 pixels per split
  long long ns = want < 1 ? 1 : want;
  if (ns > max_split) ns = max_split;
  if (ns < 1) ns = 1;
  if (ns > 1024) ns = 1024;
  long long pps = (P + ns - 1) / ns;
  pps = ((pps + kWgPix - 1) / kWgPix) * kWgPix;
  ns = (P + pps - 1) / pps;
  pl.nsplit = (int)ns;
  pl.pix_per_split = (int)pps;
  int npairs = pl.tch * WB;
  int ppq = ((((npairs + 3) / 4) + 3) / 4) * 4;
  pl.smem = basis_ctx_bytes() + (size_t)kWgPix * kWgDzPitch * 4 + (size_t)kWgPix * ppq * 4 * 4;
  return pl;
}

}  // namespace

extern "C" int kc_conv_fwd_f32(const kc_desc* d, const float* x_base, const float* x_basis, const float* w_base,
                               const float* w_basis, const float* beta, float* z, void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  if (!x_basis || !w_basis || !z) KC_FAIL(KC_ERR_INVALID, "kc_conv_fwd_f32: null pointer");
  const bool has_base = d->act != KC_ACT_NONE;
  if (has_base && (!x_base || !w_base)) KC_FAIL(KC_ERR_INVALID, "kc_conv_fwd_f32: base branch needs x_base and w_base");
  if (d->basis == KC_BASIS_GRAM && !beta) KC_FAIL(KC_ERR_INVALID, "kc_conv_fwd_f32: GRAM basis needs beta_weights");
  const int WB = d->nb + (has_base ? 1 : 0), T = d->kh * d->kw;
  size_t fixed = basis_ctx_bytes() + 3 * kTilePix * sizeof(int);
  int tch = pick_tch(T, WB * (kTilePix + kTileCout), fixed, 96 * 1024);
  size_t smem = fixed + (size_t)tch * WB * (kTilePix + kTileCout) * 4;
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_fwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long P = (long long)d->n * d->ho * d->wo;
  dim3 grid((unsigned)((P + kTilePix - 1) / kTilePix), (unsigned)((d->cout + kTileCout - 1) / kTileCout));
  kc_fwd_simt_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(*d, x_base, x_basis, w_base, w_basis, beta, z, tch);
  KC_LAUNCH_CHECK("kc_fwd_simt_kernel");
  return KC_OK;
}

long long kc_simt_dgrad_blocks(const kc_desc* d) {      // thread blocks of kc_dgrad_simt_kernel = rows of GRAM dbeta partials
  const long long P = (long long)d->n * d->h * d->w;
  return ((P + kTilePix - 1) / kTilePix) * ((d->cin + kDgCh - 1) / kDgCh);
}

extern "C" int kc_conv_dgrad_f32(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                                 const float* w_base, const float* w_basis, const float* beta, float* dx_base,
                                 float* dx_basis, float* dbeta, void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  if (!dz || !x_basis || !w_basis || !dx_basis) KC_FAIL(KC_ERR_INVALID, "kc_conv_dgrad_f32: null pointer");
  const bool has_base = d->act != KC_ACT_NONE;
  if (has_base && (!x_base || !w_base)) KC_FAIL(KC_ERR_INVALID, "kc_conv_dgrad_f32: base branch needs x_base and w_base");
  if (d->basis == KC_BASIS_GRAM && !beta) KC_FAIL(KC_ERR_INVALID, "kc_conv_dgrad_f32: GRAM basis needs beta_weights");
  const int T = d->kh * d->kw;
  size_t fixed = basis_ctx_bytes() + 3 * kTilePix * sizeof(int) + (kThreads / 32) * KC_MAX_BASIS * sizeof(float);
  int per_tap = kDgCo * kTilePix + kDgCo * kDgCh * kDgWBP;
  int tch = pick_tch(T, per_tap, fixed, 96 * 1024);
  size_t smem = fixed + (size_t)tch * per_tap * 4;
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_dgrad_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long P = (long long)d->n * d->h * d->w;
  dim3 grid((unsigned)((P + kTilePix - 1) / kTilePix), (unsigned)((d->cin + kDgCh - 1) / kDgCh));
  kc_dgrad_simt_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(*d, dz, x_base, x_basis, w_base, w_basis, beta,
                                                                      dx_base, dx_basis, dbeta, tch);
  KC_LAUNCH_CHECK("kc_dgrad_simt_kernel");
  if (d->basis == KC_BASIS_GRAM && dbeta != nullptr) return kc_dbeta_reduce(dbeta, (long long)grid.x * grid.y, stream);
  return KC_OK;
}

extern "C" size_t kc_wgrad_workspace_bytes(const kc_desc* d) {
  if (kc_validate_desc(d) != KC_OK) return 0;
  WgradPlan pl = plan_wgrad(d);
  const int WB = d->nb + (d->act != KC_ACT_NONE ? 1 : 0), T = d->kh * d->kw;
  return (size_t)pl.nsplit * d->cout * d->cin * T * WB * sizeof(float);
}

extern "C" int kc_conv_wgrad_f32(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                                 const float* beta, float* dw_base, float* dw_basis, void* workspace, void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  if (!dz || !x_basis || !dw_basis || !workspace) KC_FAIL(KC_ERR_INVALID, "kc_conv_wgrad_f32: null pointer");
  const bool has_base = d->act != KC_ACT_NONE;
  if (has_base && (!x_base || !dw_base)) KC_FAIL(KC_ERR_INVALID, "kc_conv_wgrad_f32: base branch needs x_base and dw_base");
  if (d->basis == KC_BASIS_GRAM && !beta) KC_FAIL(KC_ERR_INVALID, "kc_conv_wgrad_f32: GRAM basis needs beta_weights");
  WgradPlan pl = plan_wgrad(d);
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_wgrad_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  dim3 grid((unsigned)pl.nsplit, (unsigned)(d->cin * pl.ntchunks), (unsigned)((d->cout + kTileCout - 1) / kTileCout));
  kc_wgrad_simt_kernel<<<grid, kThreads, pl.smem, (cudaStream_t)stream>>>(*d, dz, x_base, x_basis, beta, (float*)workspace,
                                                                         pl.tch, pl.nsplit, pl.pix_per_split);
  KC_LAUNCH_CHECK("kc_wgrad_simt_kernel");
  const int WB = d->nb + (has_base ? 1 : 0), T = d->kh * d->kw;
  long long total = (long long)d->cout * d->cin * T * WB;
  int blocks = (int)((total + kThreads - 1) / kThreads);
  if (blocks > kc_sm_count() * 8) blocks = kc_sm_count() * 8;
  kc_wgrad_reduce_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(*d, (const float*)workspace, dw_base, dw_basis, pl.nsplit);
  KC_LAUNCH_CHECK("kc_wgrad_reduce_kernel");
  return KC_OK;
}
