// kc_tc.cu - BF16 tensor-core (tcgen05 / TMEM) kernels of the KAN convolution for sm_100a.
//
// "Flat-shift" implicit GEMM.  With stride 1 / dilation 1 the padded input of the whole batch is viewed as ONE flat
// sequence with row pitch P = W + pad_w and (H + pad_h) rows per image; the gap positions are zero.  Output position m
// and filter tap (r, s) then read input position  m + (r - pad_h) * P + (s - pad_w)  - a pure shift.  A CTA owns 128
// consecutive flat positions (the MMA M dimension).  Producer warps evaluate the basis functions (or the base activation)
// ONCE per needed input position and channel into shared memory as bf16 "planes" [k-core][row][8 x bf16]; because the
// no-swizzle UMMA layout accepts any 16-byte aligned start address, the A operand of tap (r, s) is just the same buffer
// viewed from row r*SS + s.  The expanded tensor of the reference (kan_layers.py:236-239) never reaches HBM, and the
// basis is evaluated (kh*SEG)/(128) ~ 3x per input element instead of 9x.  B (packed bf16 weights) streams in through
// cp.async.bulk (TMA engine) behind an mbarrier ring; tcgen05.mma accumulates fp32 in TMEM; 4 epilogue warps read TMEM
// with tcgen05.ld and write z (fp32 NCHW) coalesced.
//
// Warp roles (448 threads): 0-3 epilogue | 4 MMA issuer + TMEM allocator | 5 weight loader | 6-13 basis producers.
#include <string.h>

#include "kc_common.cuh"
#include "kc_umma.cuh"

namespace {

using namespace kc;

constexpr int kTcThreads = 448;
constexpr int kProdThreads = 256;
constexpr int kProdWarp0 = 6;
constexpr int kTileM = 128;
constexpr int kMaxBStages = 4;
constexpr size_t kSmemLimit = 227 * 1024;

struct TcGeom {
  int Cp, cps, nsc, ngroups, nbc, last_base_cols;
  int P, IMG, ph, pw;          // flat pitch, flat size of one image, effective padding of THIS gemm
  long long L;                 // flat length of the batch
  int seglen, SS, nrows, plane_bytes;
  int ntile, n_ntiles, tmem_cols, bstages;
  long long mtiles;
  long long wimg_bytes_per_ntile;
  size_t smem_bytes;
  int fast_cubic;
  float t0, inv_h;
};

struct TcFwdArgs {
  kc_desc d;
  TcGeom g;
  const float* x_base;
  const float* x_basis;
  const unsigned char* wp;
  const float* beta;
  float* z;
};

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------------------
// basis -> 8 packed bf16
// ---------------------------------------------------------------------------------------------------------
// Uniform cubic B-spline, closed form: the 4 non-zero weights land at j = i0-3 .. i0.
__device__ __forceinline__ uint4 cubic8(float x, float t0, float inv_h, int nintervals) {
  float u = (x - t0) * inv_h;
  uint4 out = make_uint4(0u, 0u, 0u, 0u);
  if (!(u >= 0.0f) || !(u < (float)nintervals)) return out;     // outside the knot span (or NaN): all-zero row
  float fi = floorf(u);
  float f = u - fi, omf = 1.0f - f;
  int i0 = (int)fi;
  float f2 = f * f, f3 = f2 * f;
  const float s = 1.0f / 6.0f;
  float w0 = omf * omf * omf * s;
  float w1 = (3.0f * f3 - 6.0f * f2 + 4.0f) * s;
  float w2 = (-3.0f * f3 + 3.0f * f2 + 3.0f * f + 1.0f) * s;
  float w3 = f3 * s;
  unsigned long long v = (unsigned long long)pack_bf16(w0, w1) | ((unsigned long long)pack_bf16(w2, w3) << 32);
  int sh = 16 * (i0 - 3);
  unsigned long long lo, hi;
  if (sh < 0) { lo = v >> (-sh); hi = 0ull; }
  else if (sh == 0) { lo = v; hi = 0ull; }
  else if (sh < 64) { lo = v << sh; hi = v >> (64 - sh); }
  else { lo = 0ull; hi = (sh < 128) ? (v << (sh - 64)) : 0ull; }
  out.x = (unsigned)lo; out.y = (unsigned)(lo >> 32); out.z = (unsigned)hi; out.w = (unsigned)(hi >> 32);
  return out;
}

__device__ __forceinline__ uint4 basis8(const KcBasisCtx& B, const TcGeom& g, float x) {
  if (g.fast_cubic) return cubic8(x, g.t0, g.inv_h, B.nparams - 1);
  float phi[KC_MAX_BASIS];
  kc_eval_basis(B, x, phi, nullptr, 1);
  return make_uint4(pack_bf16(phi[0], phi[1]), pack_bf16(phi[2], phi[3]), pack_bf16(phi[4], phi[5]), pack_bf16(phi[6], phi[7]));
}
__device__ __forceinline__ uint2 basis4(const KcBasisCtx& B, float x) {
  float phi[KC_MAX_BASIS];
  kc_eval_basis(B, x, phi, nullptr, 1);
  return make_uint2(pack_bf16(phi[0], phi[1]), pack_bf16(phi[2], phi[3]));
}

// ---------------------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) kc_fwd_tc_kernel(const __grid_constant__ TcFwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const kc_desc& d = a.d;
  const TcGeom& g = a.g;
  // ---- carve shared memory ---------------------------------------------------------------------------
  const int abuf_bytes = 8 * g.plane_bytes;
  const int bstage_bytes = 8 * g.ntile * 16;
  unsigned char* abuf0 = smem;
  unsigned char* bst0 = abuf0 + 2 * abuf_bytes;
  int* rowoff = reinterpret_cast<int*>(bst0 + g.bstages * bstage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(rowoff) + round_up(g.nrows * 4, 16));
  uint64_t* a_full = bars;                 // [2]
  uint64_t* a_empty = bars + 2;            // [2]
  uint64_t* b_full = bars + 4;             // [kMaxBStages]
  uint64_t* b_empty = bars + 4 + kMaxBStages;
  uint64_t* acc_full = bars + 4 + 2 * kMaxBStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5 + 2 * kMaxBStages);
  KcBasisCtx* B = reinterpret_cast<KcBasisCtx*>(reinterpret_cast<unsigned char*>(tmem_ptr) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * kTileM;
  const int nt = blockIdx.y;
  const int T = d.kh * d.kw, HW = d.h * d.w;
  const bool has_base = d.act != KC_ACT_NONE;
  const int nchunks = g.nsc + (has_base ? g.nbc : 0);

  if (threadIdx.x == 0) {
    mbar_init(&a_full[0], kProdThreads); mbar_init(&a_full[1], kProdThreads);
    mbar_init(&a_empty[0], 1); mbar_init(&a_empty[1], 1);
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_ptr, (uint32_t)g.tmem_cols);
  kc_load_basis_ctx(B, d, a.beta);        // ends with __syncthreads()
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= kProdWarp0) {
    // ================================ basis / activation producers ====================================
    const int tp = threadIdx.x - kProdWarp0 * 32;
    const long long qbase = m0 - (long long)g.ph * g.P - g.pw;
    for (int b = tp; b < g.nrows; b += kProdThreads) {
      int r = min(b / g.SS, d.kh - 1);
      long long q = qbase + (long long)r * g.P + (b - r * g.SS);
      int off = -1;
      if (q >= 0 && q < g.L) {
        int n = (int)(q / g.IMG);
        int rem = (int)(q - (long long)n * g.IMG);
        int y = rem / g.P, x = rem - y * g.P;
        if (y < d.h && x < d.w) off = (int)((long long)n * d.x_batch_stride + y * d.w + x);
      }
      rowoff[b] = off;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads) : "memory");
    for (int q = 0; q < nchunks; ++q) {
      const int buf = q & 1;
      unsigned char* ab = abuf0 + buf * abuf_bytes;
      mbar_wait(&a_empty[buf], ((q >> 1) & 1) ^ 1);
      if (q < g.nsc) {
        if (d.nb == 8) {
          for (int cl = 0; cl < 8; ++cl) {
            const int c = q * 8 + cl;
            const float* xc = a.x_basis + (long long)c * HW;
            uint4* plane = reinterpret_cast<uint4*>(ab + cl * g.plane_bytes);
            for (int b = tp; b < g.nrows; b += kProdThreads) {
              int off = rowoff[b];
              uint4 v = make_uint4(0u, 0u, 0u, 0u);
              if (off >= 0 && c < d.cin) v = basis8(*B, g, __ldg(xc + off));
              plane[b] = v;
            }
          }
        } else {   // nb == 4: two channels share one 16-byte k-core
          for (int pl = 0; pl < 8; ++pl) {
            const int c = q * 16 + pl * 2;
            const float* xc = a.x_basis + (long long)c * HW;
            uint4* plane = reinterpret_cast<uint4*>(ab + pl * g.plane_bytes);
            for (int b = tp; b < g.nrows; b += kProdThreads) {
              int off = rowoff[b];
              uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
              if (off >= 0) {
                if (c < d.cin) lo = basis4(*B, __ldg(xc + off));
                if (c + 1 < d.cin) hi = basis4(*B, __ldg(xc + HW + off));
              }
              plane[b] = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
          }
        }
      } else {
        const int bq = q - g.nsc;
        const int ncols = (bq == g.nbc - 1) ? g.last_base_cols : 8;
        for (int pl = 0; pl < ncols; ++pl) {
          const int grp = bq * 8 + pl;
          const float* xc = a.x_base + (long long)grp * 8 * HW;
          uint4* plane = reinterpret_cast<uint4*>(ab + pl * g.plane_bytes);
          for (int b = tp; b < g.nrows; b += kProdThreads) {
            int off = rowoff[b];
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (off >= 0 && grp < g.ngroups) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = (grp * 8 + i < d.cin) ? kc_act(d.act, __ldg(xc + (long long)i * HW + off)) : 0.0f;
              v = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
            plane[b] = v;
          }
        }
      }
      fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      mbar_arrive(&a_full[buf]);
    }
  } else if (warp == 4) {
    // ================================ MMA issuer ======================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kTileM, g.ntile, 0, 0);
      int stage = 0;
      uint32_t bphase = 0, accumulate = 0;
      for (int q = 0; q < nchunks; ++q) {
        const int buf = q & 1;
        const int ncols = (q < g.nsc) ? 8 : ((q - g.nsc == g.nbc - 1) ? g.last_base_cols : 8);
        mbar_wait(&a_full[buf], (q >> 1) & 1);
        tc_fence_after();
        const uint32_t abase = smem_u32(abuf0 + buf * abuf_bytes);
        for (int t = 0; t < T; ++t) {
          mbar_wait(&b_full[stage], bphase);
          tc_fence_after();
          const int r = t / d.kw, s = t - r * d.kw;
          const uint32_t aaddr = abase + (uint32_t)(r * g.SS + s) * 16u;
          const uint32_t baddr = smem_u32(bst0 + stage * bstage_bytes);
          for (int i = 0; i < ncols / 2; ++i) {
            uint64_t ad = make_smem_desc(aaddr + (uint32_t)(2 * i) * g.plane_bytes, (uint32_t)g.plane_bytes, 128u);
            uint64_t bd = make_smem_desc(baddr + (uint32_t)(2 * i) * g.ntile * 16u, (uint32_t)g.ntile * 16u, 128u);
            tc_mma_bf16(tmem_base, ad, bd, idesc, accumulate);
            accumulate = 1;
          }
          tc_commit(&b_empty[stage]);
          if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
        }
        tc_commit(&a_empty[buf]);
      }
      tc_commit(acc_full);
    }
    __syncwarp();
  } else if (warp == 5) {
    // ================================ weight loader (TMA engine bulk copies) ==========================
    if (lane == 0) {
      int stage = 0;
      uint32_t bphase = 0;
      const unsigned char* wsrc = a.wp + (long long)nt * g.wimg_bytes_per_ntile;
      for (int q = 0; q < nchunks; ++q) {
        const int ncols = (q < g.nsc) ? 8 : ((q - g.nsc == g.nbc - 1) ? g.last_base_cols : 8);
        const uint32_t bytes = (uint32_t)ncols * g.ntile * 16u;
        for (int t = 0; t < T; ++t) {
          mbar_wait(&b_empty[stage], bphase ^ 1);
          mbar_arrive_expect_tx(&b_full[stage], bytes);
          bulk_g2s(bst0 + stage * bstage_bytes, wsrc, bytes, &b_full[stage]);
          wsrc += bytes;
          if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue: TMEM -> registers -> z (fp32 NCHW) =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const long long q = m0 + warp * 32 + lane;
    bool valid = false;
    long long zoff = 0;
    const int HoWo = d.ho * d.wo;
    if (q < g.L) {
      int n = (int)(q / g.IMG);
      int rem = (int)(q - (long long)n * g.IMG);
      int y = rem / g.P, x = rem - y * g.P;
      if (y < d.ho && x < d.wo) { valid = true; zoff = (long long)n * d.z_batch_stride + y * d.wo + x; }
    }
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int n0 = nt * g.ntile;
    for (int c0 = 0; c0 < g.ntile; c0 += 32) {
      uint32_t r[32];
      if (g.ntile - c0 >= 32) {
        tmem_ld32(trow + (uint32_t)c0, r);
      } else {
        tmem_ld16(trow + (uint32_t)c0, r);
#pragma unroll
        for (int i = 16; i < 32; ++i) r[i] = 0u;
      }
      tmem_ld_wait();
      const int lim = min(32, g.ntile - c0);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        int co = n0 + c0 + i;
        if (valid && i < lim && co < d.cout) a.z[zoff + (long long)co * HoWo] = __uint_as_float(r[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// weight packing: fp32 reference layout -> bf16 K-major no-swizzle images [ntile][chunk][tap][k-core][cout][8]
// ---------------------------------------------------------------------------------------------------------
struct TcPackArgs { kc_desc d; TcGeom g; const float* w_base; const float* w_basis; uint4* out; };

__global__ void __launch_bounds__(256) kc_pack_fwd_kernel(const __grid_constant__ TcPackArgs a) {
  const kc_desc& d = a.d;
  const TcGeom& g = a.g;
  const int T = d.kh * d.kw, nb = d.nb;
  const bool has_base = d.act != KC_ACT_NONE;
  const long long vec_per_ntile = g.wimg_bytes_per_ntile / 16;
  const long long total = vec_per_ntile * g.n_ntiles;
  const long long spline_vecs = (long long)g.nsc * T * 8 * g.ntile;
  const long long full_chunk = (long long)T * 8 * g.ntile;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const int nt = (int)(v / vec_per_ntile);
    long long vl = v - (long long)nt * vec_per_ntile;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.0f;
    if (vl < spline_vecs) {
      int q = (int)(vl / full_chunk);
      long long rem = vl - (long long)q * full_chunk;
      int t = (int)(rem / (8 * g.ntile));
      int rem2 = (int)(rem - (long long)t * 8 * g.ntile);
      int kc = rem2 / g.ntile, nl = rem2 - kc * g.ntile;
      int co = nt * g.ntile + nl;
      if (co < d.cout) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int c, j;
          if (nb == 8) { c = q * 8 + kc; j = e; } else { c = q * 16 + kc * 2 + (e >> 2); j = e & 3; }
          if (c < d.cin) f[e] = a.w_basis[((long long)co * d.cin * nb + kc_wbasis_index(d.basis, c, j, d.cin, nb)) * T + t];
        }
      }
    } else if (has_base) {
      long long vb = vl - spline_vecs;
      int bq = (int)min((long long)(g.nbc - 1), vb / full_chunk);
      long long rem = vb - (long long)bq * full_chunk;
      int ncols = (bq == g.nbc - 1) ? g.last_base_cols : 8;
      int t = (int)(rem / (ncols * g.ntile));
      int rem2 = (int)(rem - (long long)t * ncols * g.ntile);
      int kc = rem2 / g.ntile, nl = rem2 - kc * g.ntile;
      int co = nt * g.ntile + nl, grp = bq * 8 + kc;
      if (co < d.cout && grp < g.ngroups) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int c = grp * 8 + e;
          if (c < d.cin) f[e] = a.w_base[((long long)co * d.cin + c) * T + t];
        }
      }
    }
    a.out[v] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// ---------------------------------------------------------------------------------------------------------
// host-side geometry
// ---------------------------------------------------------------------------------------------------------
bool knots_uniform_cubic(const kc_desc* d, float* t0, float* inv_h) {
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

int tc_forward_geometry(const kc_desc* d, TcGeom* g) {
  if (d->stride_h != 1 || d->stride_w != 1 || d->dil_h != 1 || d->dil_w != 1) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs stride 1 and dilation 1");
  if (d->nb != 8 && d->nb != 4) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs basis width 4 or 8 (got %d)", d->nb);
  if (d->pad_h > d->kh - 1 || d->pad_w > d->kw - 1) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs padding < kernel size");
  if (d->kw > 8 || d->kh > 8) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs kernel size <= 8");
  if ((long long)d->n * d->x_batch_stride >= (1LL << 31)) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs < 2^31 input elements");
  memset(g, 0, sizeof(*g));
  const bool has_base = d->act != KC_ACT_NONE;
  const int T = d->kh * d->kw;
  g->cps = (d->nb == 8) ? 8 : 16;
  g->Cp = round_up(d->cin, g->cps);
  g->nsc = g->Cp / g->cps;
  g->ngroups = g->Cp / 8;
  g->nbc = has_base ? (g->ngroups + 7) / 8 : 0;
  g->last_base_cols = has_base ? round_up(g->ngroups - (g->nbc - 1) * 8, 2) : 0;
  g->ph = d->pad_h; g->pw = d->pad_w;
  g->P = d->w + d->pad_w;
  g->IMG = (d->h + d->pad_h) * g->P;
  g->L = (long long)d->n * g->IMG;
  g->seglen = round_up(kTileM + d->kw - 1, 8);
  g->SS = g->P < g->seglen ? g->P : g->seglen;
  g->nrows = (d->kh - 1) * g->SS + g->seglen;
  g->plane_bytes = g->nrows * 16 + 16;       // +16 B: consecutive planes start 4 banks apart
  int want_tiles = (d->cout + 255) / 256;
  g->ntile = round_up((d->cout + want_tiles - 1) / want_tiles, 16);
  g->n_ntiles = (d->cout + g->ntile - 1) / g->ntile;
  g->tmem_cols = 32;
  while (g->tmem_cols < g->ntile) g->tmem_cols *= 2;
  g->mtiles = (g->L + kTileM - 1) / kTileM;
  int base_cols = has_base ? (g->nbc - 1) * 8 + g->last_base_cols : 0;
  g->wimg_bytes_per_ntile = (long long)T * g->ntile * 16 * (g->nsc * 8 + base_cols);
  size_t fixed = 2 * 8 * (size_t)g->plane_bytes + round_up(g->nrows * 4, 16) + (5 + 2 * kMaxBStages) * 8 + 16 + sizeof(KcBasisCtx) + 128;
  size_t bstage = 8 * (size_t)g->ntile * 16;
  g->bstages = kMaxBStages;
  while (g->bstages > 2 && fixed + g->bstages * bstage > kSmemLimit) --g->bstages;
  g->smem_bytes = fixed + g->bstages * bstage;
  if (g->smem_bytes > kSmemLimit) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path: tile does not fit shared memory");
  g->fast_cubic = knots_uniform_cubic(d, &g->t0, &g->inv_h) ? 1 : 0;
  return KC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// UMMA descriptor self-test kernel (one CTA, 128 threads): D[128 x 64] = A[128 x 64] * B[64 x 64]^T
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float st_val(int i, int k, int salt) { return (float)(((i * 7 + k * 13 + salt) % 17) - 8) * 0.125f; }

__global__ void __launch_bounds__(128, 1) kc_umma_selftest_kernel(int mode, float* max_err) {
  __shared__ __align__(128) unsigned char sA[8 * 2176 + 2048 * 8];
  __shared__ __align__(128) unsigned char sB[8 * 1024];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int mn_major = mode & 1, swap = (mode >> 1) & 1;
  // A: 128 x 64, B: 64 x 64 (row = n), both bf16
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_start, a_step, b_step;
  for (int i = tid; i < (int)sizeof(sA) / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0u;
  __syncthreads();
  __nv_bfloat16* A16 = reinterpret_cast<__nv_bfloat16*>(sA);
  __nv_bfloat16* B16 = reinterpret_cast<__nv_bfloat16*>(sB);
  if (!mn_major) {
    // K-major planes [k-core][row][8]; A planes have 136 rows and the matrix starts at row 3 (unaligned-start view test)
    for (int i = tid; i < 128 * 64; i += 128) {
      int m = i / 64, k = i % 64;
      A16[((k / 8) * 136 + (m + 3)) * 8 + (k % 8)] = __float2bfloat16(st_val(m, k, 1));
    }
    for (int i = tid; i < 64 * 64; i += 128) {
      int n = i / 64, k = i % 64;
      B16[((k / 8) * 64 + n) * 8 + (k % 8)] = __float2bfloat16(st_val(n, k, 5));
    }
    a_lbo = 136 * 16; a_sbo = 128; b_lbo = 64 * 16; b_sbo = 128;
    a_start = 3 * 16; a_step = 2 * a_lbo; b_step = 2 * b_lbo;
  } else {
    // MN-major planes [mn-group][k row][8 mn]
    for (int i = tid; i < 128 * 64; i += 128) {
      int m = i / 64, k = i % 64;
      A16[((m / 8) * 64 + k) * 8 + (m % 8)] = __float2bfloat16(st_val(m, k, 1));
    }
    for (int i = tid; i < 64 * 64; i += 128) {
      int n = i / 64, k = i % 64;
      B16[((n / 8) * 64 + k) * 8 + (n % 8)] = __float2bfloat16(st_val(n, k, 5));
    }
    a_lbo = 128; a_sbo = 1024; b_lbo = 128; b_sbo = 1024;
    a_start = 0; a_step = 256; b_step = 256;
  }
  if (swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_ptr, 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64, mn_major, mn_major);
    for (int i = 0; i < 4; ++i) {
      uint64_t ad = make_smem_desc(smem_u32(sA) + a_start + i * a_step, a_lbo, a_sbo);
      uint64_t bd = make_smem_desc(smem_u32(sB) + i * b_step, b_lbo, b_sbo);
      tc_mma_bf16(tb, ad, bd, idesc, i > 0 ? 1u : 0u);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float err = 0.0f;
  const int m = tid;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) {
      int n = c0 + i;
      float ref = 0.0f;
      for (int k = 0; k < 64; ++k) ref = fmaf(st_val(m, k, 1), st_val(n, k, 5), ref);
      float e = fabsf(__uint_as_float(r[i]) - ref);
      if (!(e <= 1e30f)) e = 1e30f;
      err = fmaxf(err, e);
    }
  }
  atomicMax(reinterpret_cast<int*>(max_err), __float_as_int(err));
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64);
}

}  // namespace

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" int kc_tc_supported(const kc_desc* d) {
  if (kc_validate_desc(d) != KC_OK) return 0;
  TcGeom g;
  return tc_forward_geometry(d, &g) == KC_OK ? 1 : 0;
}

extern "C" size_t kc_tc_bytes(const kc_desc* d, int which) {
  if (kc_validate_desc(d) != KC_OK) return 0;
  TcGeom g;
  if (tc_forward_geometry(d, &g) != KC_OK) return 0;
  if (which == 0) return (size_t)g.wimg_bytes_per_ntile * g.n_ntiles;
  return 0;
}

extern "C" int kc_tc_pack_weights(const kc_desc* d, const float* w_base, const float* w_basis, void* packed_fwd,
                                  void* packed_dgrad, void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  TcGeom g;
  rc = tc_forward_geometry(d, &g);
  if (rc != KC_OK) return rc;
  if (!w_basis || !packed_fwd) KC_FAIL(KC_ERR_INVALID, "kc_tc_pack_weights: null pointer");
  if (d->act != KC_ACT_NONE && !w_base) KC_FAIL(KC_ERR_INVALID, "kc_tc_pack_weights: base branch needs w_base");
  (void)packed_dgrad;
  TcPackArgs a;
  a.d = *d; a.g = g; a.w_base = w_base; a.w_basis = w_basis; a.out = (uint4*)packed_fwd;
  long long total = g.wimg_bytes_per_ntile / 16 * g.n_ntiles;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  kc_pack_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  KC_LAUNCH_CHECK("kc_pack_fwd_kernel");
  return KC_OK;
}

extern "C" int kc_conv_fwd_tc(const kc_desc* d, const float* x_base, const float* x_basis, const void* packed_fwd,
                              const float* beta, float* z, void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  TcGeom g;
  rc = tc_forward_geometry(d, &g);
  if (rc != KC_OK) return rc;
  if (!x_basis || !packed_fwd || !z) KC_FAIL(KC_ERR_INVALID, "kc_conv_fwd_tc: null pointer");
  if (d->act != KC_ACT_NONE && !x_base) KC_FAIL(KC_ERR_INVALID, "kc_conv_fwd_tc: base branch needs x_base");
  if (d->basis == KC_BASIS_GRAM && !beta) KC_FAIL(KC_ERR_INVALID, "kc_conv_fwd_tc: GRAM basis needs beta_weights");
  if (g.mtiles > 0x7fffffffLL) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_conv_fwd_tc: too many tiles");
  TcFwdArgs a;
  a.d = *d; a.g = g; a.x_base = x_base; a.x_basis = x_basis; a.wp = (const unsigned char*)packed_fwd; a.beta = beta; a.z = z;
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  dim3 grid((unsigned)g.mtiles, (unsigned)g.n_ntiles);
  kc_fwd_tc_kernel<<<grid, kTcThreads, g.smem_bytes, (cudaStream_t)stream>>>(a);
  KC_LAUNCH_CHECK("kc_fwd_tc_kernel");
  return KC_OK;
}

extern "C" int kc_conv_dgrad_tc(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                                const void* packed_dgrad, const float* beta, float* dx_base, float* dx_basis,
                                float* dbeta, void* workspace, void* stream) {
  (void)d; (void)dz; (void)x_base; (void)x_basis; (void)packed_dgrad; (void)beta; (void)dx_base; (void)dx_basis;
  (void)dbeta; (void)workspace; (void)stream;
  KC_FAIL(KC_ERR_UNSUPPORTED, "kc_conv_dgrad_tc: not built yet");
}

extern "C" int kc_conv_wgrad_tc(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                                const float* beta, float* dw_base, float* dw_basis, void* workspace, void* stream) {
  (void)d; (void)dz; (void)x_base; (void)x_basis; (void)beta; (void)dw_base; (void)dw_basis; (void)workspace; (void)stream;
  KC_FAIL(KC_ERR_UNSUPPORTED, "kc_conv_wgrad_tc: not built yet");
}

extern "C" int kc_tc_selftest(int mode, float* max_abs_err, void* stream) {
  if (!max_abs_err) KC_FAIL(KC_ERR_INVALID, "kc_tc_selftest: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  float* dev = nullptr;
  KC_CUDA_CHECK(cudaMalloc(&dev, sizeof(float)));
  cudaError_t e = cudaMemsetAsync(dev, 0, sizeof(float), st);
  if (e == cudaSuccess) {
    kc_umma_selftest_kernel<<<1, 128, 0, st>>>(mode, dev);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(max_abs_err, dev, sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(dev);
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "kc_tc_selftest: %s", cudaGetErrorString(e));
  return KC_OK;
}
