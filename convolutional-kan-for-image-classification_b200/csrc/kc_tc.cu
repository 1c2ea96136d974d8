// kc_tc.cu - BF16 tensor-core (tcgen05 / TMEM) kernels of the KAN convolution for sm_100a.
//
// "Flat-shift" implicit GEMM.  With stride 1 / dilation 1 the padded input of the whole batch is viewed as ONE flat
// sequence with row pitch P = W + pad_w and (H + pad_h) rows per image; the gap positions are zero.  Output position m
// and filter tap (r, s) then read input position  m + (r - pad_h) * P + (s - pad_w)  - a pure shift.  A CTA owns
// nsub * 128 consecutive flat positions (nsub MMA M-tiles, nsub * N <= 512 TMEM columns).  Producer warps evaluate the
// basis functions (or the base activation) ONCE per needed input position and channel into shared memory as bf16
// "planes" [k-core][row][8 x bf16]; because the no-swizzle UMMA layout accepts any 16-byte aligned start address, the
// A operand of (sub-tile i, tap (r, s)) is just the same buffer viewed from row r*SS + s + 128*i.  The expanded tensor
// of the reference (kan_layers.py:236-239) never reaches HBM and the basis is evaluated ~(M + 2P)/M times per input
// element instead of 9 times.  B (packed bf16 weights, K = 32 per ring step) streams in through cp.async.bulk (TMA
// engine) behind a deep mbarrier ring - measured: the MMA-completion -> commit -> refill -> MMA round trip is ~3.6k
// cycles, so every B stage feeds nsub*2 MMAs and the ring holds up to 16 stages; tcgen05.mma accumulates fp32 in TMEM;
// the producer warps read TMEM with tcgen05.ld after the mainloop and write z (fp32 NCHW) coalesced.  The rows a CTA
// evaluates for its own positions are also saved as bf16 planes (phi) for the weight gradient.  Strided layers run on
// the stride-1 position grid (sampled epilogue); 1x1 layers use the persistent kernel below on a pre-computed phi.
//
// This file also holds the persistent dgrad kernel (TMA-fed operands, double-buffered TMEM accumulators, analytic
// basis-derivative epilogue), the weight packing kernels, the dz -> flat bf16 conversion and the host-side geometry.
//
// Warp roles of kc_tc_kernel (608 threads): 0-15 producers / epilogue | 16 weight loader | 17-18 MMA issuers (17 also
// allocates TMEM).  Two issuers, each owning half of the sub-tile accumulators: the tensor-core queue is shallow, so the
// ~250 cycles of per-ring-step bookkeeping of one issuer would otherwise leave the pipe idle at small N.
// The MMA issuer is the highest warp id on purpose: the scheduler arbitrates highest-warp-id-first, and the single issuing
// thread must never wait behind the 16 ALU-heavy producer warps (measured: 245 -> ~50 cycles per tcgen05.mma).
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "kc_common.cuh"
#include "kc_umma.cuh"
#include "kc_tc_basis.cuh"

size_t kc_tc_wgrad_ws_bytes(const kc_desc* d, int which);                                              // kc_tc_wgrad.cu
int kc_tc_wgrad_phi_layout(const kc_desc* d, long long* L, int* spline_planes, int* base_planes);    // kc_tc_wgrad.cu
int kc_tc_phi_prepass(const kc_desc* d, const float* x_base, const float* x_basis, const float* beta, void* phi, void* stream);
long long kc_simt_dgrad_blocks(const kc_desc* d);                                                        // kc_simt.cu

namespace {

using namespace kc;

constexpr int kTcThreads = 608;        // 19 warps
constexpr int kProdThreads = 512;      // warps 0-15
constexpr int kLoaderWarp = 16, kMmaWarp = 17;   // MMA issuers: warps 17 and 18 (sub-tiles split between them)
constexpr int kRowThreads = 256;       // producer thread t owns rows (t & 255) + 256*k and plane half (t >> 8)
constexpr int kRB = 4;                 // rows per producer thread  (nrows <= 1024)
constexpr int kPL = 4;                 // k-cores ("planes") per chunk: K = 32 per ring step
constexpr int kTileM = 128;
constexpr int kMaxA = 3;
constexpr int kMaxBStages = 16;
constexpr int kNumBars = 3 * kMaxA + 3 * kMaxBStages + 1;
constexpr size_t kSmemLimit = 227 * 1024;

struct TcGeom {
  int Cp, cps, nsc, ngroups, nbc, last_base_cols;
  int P, IMG, ph, pw;          // flat pitch, flat size of one image, effective padding of THIS gemm
  long long L;                 // flat length of the batch
  int nsub, mcta;              // M sub-tiles per CTA, rows per CTA
  int SS, nrows, plane_bytes;
  int ntile, n_ntiles, tmem_cols, bstages, na;
  int tps;                     // filter taps per B ring stage (1, or kw = a whole filter row)
  int cpt;                     // dgrad: input channels per N tile (16, or 14 for the persistent kernel with a base column)
  int persistent;              // dgrad: 1 = kc_dgrad_persistent_kernel (double-buffered TMEM, epilogue overlaps the MMAs)
  int from_phi;                // forward of a 1x1 convolution: basis rows come from the phi buffer (pre-pass + persistent GEMM)
  int pair;                    // persistent dgrad: 1 = CTA pairs (tcgen05 cta_group::2), each CTA holds half of the N rows of a weight stage
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
  unsigned char* phi_out;       // optional: basis rows saved for the weight gradient, plane-major [plane][L][8] bf16
  int phi_base_plane0;          // index of the first base-activation plane in phi_out
  // dgrad mode only
  const unsigned char* dzf;     // bf16 plane-major [cq/8][L][8], zero at padding / invalid output positions
  int cq;                       // channels per flat row (cout rounded up to 16)
  float* dx_base;
  float* dx_basis;
  float* dbeta;                 // GRAM only: d/d beta_weights accumulator (atomics), else nullptr
};

constexpr int kModeFwd = 0, kModeDgrad = 1;

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

KC_TRACE_DECL(g_trace)

// generic (any family) evaluators are kept out of line so that their local arrays do not inflate the register
// allocation of the closed-form cubic fast path
// RBF / Chebyshev values j < 8 packed to bf16, registers only (compile-time indices); out of line so that the closed-form
// cubic path of the producers keeps its register budget.
__device__ __noinline__ uint4 basis8_lean(const KcBasisCtx& B, float x) {
  const int nb = B.nb;
  float phi[8];
  if (B.kind == KC_BASIS_RBF) {
    const float inv_den = __fdividef(1.0f, B.p[nb]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float q = (x - B.p[j]) * inv_den;
      phi[j] = j < nb ? __expf(-(q * q)) : 0.0f;
    }
  } else {
    const float lo = -1.0f + 1e-7f, hi = 1.0f - 1e-7f;
    const float t = tc_tanh(x);
    float c = fminf(fmaxf(t, lo), hi);
    if (t != t) c = t;
    float T0 = 1.0f, T1 = c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      phi[j] = j < nb ? T0 : 0.0f;
      const float T2 = 2.0f * c * T1 - T0;
      T0 = T1; T1 = T2;
    }
  }
  return make_uint4(pack_bf16(phi[0], phi[1]), pack_bf16(phi[2], phi[3]), pack_bf16(phi[4], phi[5]), pack_bf16(phi[6], phi[7]));
}
__device__ __noinline__ uint4 basis8_generic(const KcBasisCtx& B, float x) {
  float phi[KC_MAX_BASIS];
#pragma unroll
  for (int j = 0; j < 8; ++j) phi[j] = 0.0f;      // widths 5..7 are zero-padded to 8 (their packed weights are zero too)
  tc_eval_basis(B, x, phi, nullptr);
  return make_uint4(pack_bf16(phi[0], phi[1]), pack_bf16(phi[2], phi[3]), pack_bf16(phi[4], phi[5]), pack_bf16(phi[6], phi[7]));
}
__device__ __forceinline__ uint4 basis8(const KcBasisCtx& B, const TcGeom& g, float x, bool valid) {
  if (g.fast_cubic) return cubic8(x, g.t0, g.inv_h, B.nparams - 1, valid);
  if (!valid) return make_uint4(0u, 0u, 0u, 0u);
  return (B.kind == KC_BASIS_RBF || B.kind == KC_BASIS_CHEBY) ? basis8_lean(B, x) : basis8_generic(B, x);
}
__device__ __noinline__ uint2 basis4(const KcBasisCtx& B, float x) {
  if (B.kind == KC_BASIS_RBF || B.kind == KC_BASIS_CHEBY) {
    const uint4 v = basis8_lean(B, x);
    return make_uint2(v.x, v.y);
  }
  float phi[KC_MAX_BASIS];
#pragma unroll
  for (int j = 0; j < 4; ++j) phi[j] = 0.0f;      // widths 1..3 are zero-padded to 4
  tc_eval_basis(B, x, phi, nullptr);
  return make_uint2(pack_bf16(phi[0], phi[1]), pack_bf16(phi[2], phi[3]));
}

// d/dx of the uniform cubic B-spline weights dotted with the 8 incoming gradients r[0..7] (closed form, Appendix A.2).
// Same pack-and-shift trick as the forward producer: the 4 non-zero derivative weights are packed as bf16 into 64 bits,
// moved to their slots j = i0-3..i0 with clamped shifts and unpacked with one shift/mask each (~40 instructions instead
// of a select chain per j; the bf16 rounding of the weights is far inside the BF16 tolerance of this path).
__device__ __forceinline__ float cubic8_dot_grad(float x, float t0, float inv_h, int nintervals, const uint32_t* r) {
  const float u = (x - t0) * inv_h;
  const bool ok = (u >= 0.0f) && (u < (float)nintervals);
  const float fi = floorf(u);
  const float f = u - fi, omf = 1.0f - f;
  const int i0 = min(max((int)fi, 0), 15);
  const float d0 = -0.5f * omf * omf, d1 = fmaf(1.5f * f, f, -2.0f * f), d2 = fmaf(fmaf(-1.5f, f, 1.0f), f, 0.5f), d3 = 0.5f * f * f;
  unsigned long long v = (unsigned long long)pack_bf16(d0, d1) | ((unsigned long long)pack_bf16(d2, d3) << 32);
  v = ok ? v : 0ull;
  const int sh = 16 * (i0 - 3);
  const unsigned long long lo = shl64(v, sh) | shr64(v, -sh);
  const unsigned long long hi = shr64(v, 64 - sh) | shl64(v, sh - 64);
  const uint32_t w[4] = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
  float acc = 0.0f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    acc = fmaf(__uint_as_float(r[2 * p]), __uint_as_float(w[p] << 16), acc);
    acc = fmaf(__uint_as_float(r[2 * p + 1]), __uint_as_float(w[p] & 0xffff0000u), acc);
  }
  if (tc_nonfinite(x)) acc = __int_as_float(0x7fc00000);               // the reference's autograd yields NaN here
  return acc * inv_h;
}

// base-activation derivative with fast intrinsics (tensor-core path only; the FP32 path uses kc_act_grad)
__device__ __forceinline__ float act_grad_fast(int kind, float x) {
  if (kind == KC_ACT_SILU) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-x));
    return s * fmaf(x, 1.0f - s, 1.0f);
  }
  if (kind == KC_ACT_GELU) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    return fmaf(x, 0.39894228040143268f * __expf(-0.5f * x * x), cdf);
  }
  return 1.0f;
}

// Generic (any basis family) contraction of the incoming gradients r[0..nb) with d(basis)/dx; optionally accumulates the
// GRAM d/d(beta_weights) partials.  Out of line: rare path, keeps the dgrad epilogue loop compact.
__device__ __noinline__ float dgrad_generic(const KcBasisCtx& B, float x, const uint32_t* r, float* dbl) {
  float phi[KC_MAX_BASIS], dphi[KC_MAX_BASIS], gg[KC_MAX_BASIS];
  tc_eval_basis(B, x, phi, dphi);
  float gs = 0.0f;
  for (int j = 0; j < KC_MAX_BASIS; ++j) {
    gg[j] = j < B.nb ? __uint_as_float(r[j]) : 0.0f;
    if (j < B.nb) gs = fmaf(gg[j], dphi[j], gs);
  }
  if (dbl != nullptr) kc_gram_dbeta(B, x, gg, 1, dbl);
  if (B.kind == KC_BASIS_CHEBY && kc_cheby_clamped(tc_tanh(x))) gs = 0.0f;
  return gs;
}

__device__ __forceinline__ int chunk_cols(const TcGeom& g, int q) {
  return (q < g.nsc || q - g.nsc != g.nbc - 1) ? kPL : g.last_base_cols;
}

// Order in which kc_tc_kernel walks the K chunks of a tile.  A base-activation chunk costs the producers 2-3x a spline chunk
// (32 channels per chunk instead of 4), so the base chunks are spread between the spline chunks instead of following them all:
// base chunk b comes right after the spline chunks of its own 32 channels (their x values are still in L2), the producers
// bank their slack of the cheap chunks in the A ring and spend it on the expensive one, and the MMA warps rarely wait for
// rows.  Producers, weight loader and MMA issuers run the same sequence.
struct ChunkSeq {
  int nsc, nbc, s, b;
  __device__ __forceinline__ ChunkSeq(int nsc_, int nbc_) : nsc(nsc_), nbc(nbc_), s(0), b(0) {}
  __device__ __forceinline__ int next() {           // chunk id: < nsc spline chunk, else base chunk (id - nsc)
    if (b < nbc && (s >= nsc || nbc * s >= (b + 1) * nsc)) return nsc + b++;
    return s++;
  }
};

// Straight-line issue of the NSUB x (K/16) MMAs of one ring step.  a_lo / b_lo are complete low descriptor words
// (LBO field | start address >> 4) of sub-tile 0, k-core pair 0; all other descriptors differ by small constants, so
// the issuing thread executes ~2 integer adds per tcgen05.mma and no branches.
template <int NSUB, bool PAIR = false>
__device__ __forceinline__ void issue_step(uint32_t tmem_base, uint32_t ntile, uint32_t a_lo, uint32_t b_lo, uint32_t a_k2,
                                           uint32_t b_k2, uint32_t desc_hi, uint32_t idesc, uint32_t first, bool two, int i0,
                                           int istep) {
#pragma unroll
  for (int ii = 0; ii < NSUB; ++ii) {
    const int i = i0 + ii * istep;
    const uint32_t al = a_lo + (uint32_t)(i * kTileM);
    const uint32_t td = tmem_base + (uint32_t)i * ntile;
    if (PAIR) {
      tc_mma_bf16_pair(td, ((uint64_t)desc_hi << 32) | al, ((uint64_t)desc_hi << 32) | b_lo, idesc, first);
      if (two) tc_mma_bf16_pair(td, ((uint64_t)desc_hi << 32) | (al + a_k2), ((uint64_t)desc_hi << 32) | (b_lo + b_k2), idesc, 1u);
    } else {
      tc_mma_bf16(td, ((uint64_t)desc_hi << 32) | al, ((uint64_t)desc_hi << 32) | b_lo, idesc, first);
      if (two) tc_mma_bf16(td, ((uint64_t)desc_hi << 32) | (al + a_k2), ((uint64_t)desc_hi << 32) | (b_lo + b_k2), idesc, 1u);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------------------
// PAIR (forward only): the two CTAs of a cluster own two neighbouring position tiles and the same N tile and run
// tcgen05 cta_group::2 MMAs (M = 256): each CTA evaluates its own basis rows and loads HALF of every weight stage (ntile / 2 of
// the N rows), the leader CTA issues for both.  Narrow tiles stop being bound by the operand fetch (N = 64: 43 instead of 66
// cycles, N = 128: 64 instead of 74), the weight stream from L2 halves, and a whole chunk of taps fits one ring stage.
template <int MODE, bool PAIR>
__global__ void __launch_bounds__(kTcThreads, 1) kc_tc_kernel(const __grid_constant__ TcFwdArgs a) {
  static_assert(!(PAIR && MODE != kModeFwd), "pairs: forward only");
  extern __shared__ __align__(128) unsigned char smem[];
  const kc_desc& d = a.d;
  const TcGeom& g = a.g;
  // ---- carve shared memory ---------------------------------------------------------------------------
  const int abuf_bytes = kPL * g.plane_bytes;
  const int nw = PAIR ? g.ntile / 2 : g.ntile;   // weight rows (of the MMA's N) this CTA holds
  const int btap_bytes = kPL * nw * 16;
  const int bstage_bytes = g.tps * btap_bytes;
  unsigned char* abuf0 = smem;
  unsigned char* bst0 = abuf0 + g.na * abuf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bst0 + g.bstages * bstage_bytes);
  uint64_t* a_full = bars;                       // [kMaxA]
  uint64_t* a_empty = bars + kMaxA;              // [kMaxA]
  uint64_t* b_full = bars + 2 * kMaxA;           // [kMaxBStages]
  uint64_t* b_empty = b_full + kMaxBStages;      // [kMaxBStages]
  uint64_t* acc_full = b_empty + kMaxBStages;
  uint64_t* pa_full = acc_full + 1;              // [kMaxA]        pair, leader: the peer's basis rows are in place
  uint64_t* pb_full = pa_full + kMaxA;           // [kMaxBStages]  pair, leader: the peer's half of a weight stage has landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  KcBasisCtx* B = reinterpret_cast<KcBasisCtx*>(reinterpret_cast<unsigned char*>(tmem_ptr) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  KC_TRACER(tre, g_trace, 3, threadIdx.x == 0);      // CTA life cycle: entry, producer loop done, accumulators ready, results written
  tre.stamp();
  const long long m0 = (long long)blockIdx.x * g.mcta;
  const int nt = blockIdx.y;
  const int T = d.kh * d.kw, HW = d.h * d.w;
  const bool has_base = d.act != KC_ACT_NONE;
  const int nchunks = (MODE == kModeDgrad) ? g.nbc : g.nsc + (has_base ? g.nbc : 0);

  if (threadIdx.x == 0) {
    const uint32_t nmw = g.nsub >= 2 ? 2u : 1u;      // issuing warps
    for (int i = 0; i < kMaxA; ++i) { mbar_init(&a_full[i], kProdThreads); mbar_init(&a_empty[i], nmw); mbar_init(&pa_full[i], 1); }
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], nmw); mbar_init(&pb_full[s], 1); }
    mbar_init(acc_full, nmw);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { if (PAIR) tmem_alloc_pair(tmem_ptr, (uint32_t)g.tmem_cols); else tmem_alloc(tmem_ptr, (uint32_t)g.tmem_cols); }
  kc_load_basis_ctx(B, d, a.beta);        // ends with __syncthreads()
  tc_fence_before();
  __syncthreads();
  if (PAIR) { cluster_arrive(); cluster_wait(); }        // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kLoaderWarp) {
    // ================================ basis / activation producers ====================================
    const int tp = threadIdx.x;                                      // 0 .. 511
    const int r0 = tp & (kRowThreads - 1), half = tp >> 8;
    const long long qbase = m0 - (long long)g.ph * g.P - g.pw;
    int offs[kRB];                                                   // >=0 input offset | -1 zero (padding) row | -2 no row
#pragma unroll
    for (int k = 0; k < kRB; ++k) {
      const int b = r0 + k * kRowThreads;
      int off = -2;
      if (b < g.nrows) {
        off = -1;
        int r = min(b / g.SS, d.kh - 1);
        long long q = qbase + (long long)r * g.P + (b - r * g.SS);
        if (q >= 0 && q < g.L) {
          if (MODE == kModeDgrad) {
            off = (int)q;                 // the flat dz buffer is indexed by the flat position itself
          } else {
            int n = (int)(q / g.IMG);
            int rem = (int)(q - (long long)n * g.IMG);
            int y = rem / g.P, x = rem - y * g.P;
            if (y < d.h && x < d.w) off = (int)((long long)n * d.x_batch_stride + y * d.w + x);
          }
        }
      }
      offs[k] = off;
    }
    // Rows of the centre strip are this tile's own flat positions m0 .. m0+mcta-1: when phi_out is given, their basis rows
    // are also written to global memory in the plane-major layout of the weight-gradient kernel (first cout tile only).
    unsigned ownmask = 0u;
    unsigned char* phi_row = nullptr;
    const long long phi_ps = g.L * 16;                                // bytes between planes
    if (MODE == kModeFwd && a.phi_out != nullptr && nt == 0) {
      const int bc0 = r0 - (g.ph * g.SS + g.pw);
#pragma unroll
      for (int k = 0; k < kRB; ++k) {
        const int bc = bc0 + k * kRowThreads;
        if (offs[k] != -2 && bc >= 0 && bc < g.mcta && m0 + bc < g.L) ownmask |= 1u << k;
      }
      phi_row = a.phi_out + (m0 + bc0) * 16;
    }
    const int plane_bytes = g.plane_bytes, cin = d.cin, nb = d.nb > 4 ? 8 : 4;   // padded basis width
    // x values of the NEXT spline chunk are fetched while the current one is evaluated (nb == 8 path):
    // this thread owns planes half*2 + {0,1} = channels q*4 + half*2 + {0,1}
    float xnext[kRB][2];
    auto fetch8 = [&](int q, float (&xv)[kRB][2]) {
      const int c0 = q * 4 + half * 2;
      const float* xc = a.x_basis + (long long)c0 * HW;
#pragma unroll
      for (int k = 0; k < kRB; ++k)
#pragma unroll
        for (int cl = 0; cl < 2; ++cl)
          xv[k][cl] = (offs[k] >= 0 && c0 + cl < cin) ? __ldg(xc + (long long)cl * HW + offs[k]) : 0.0f;
    };
    KC_TRACER(trp, g_trace, 0, tp == 0);
    if (nb == 8 && g.nsc > 0) fetch8(0, xnext);
    if (MODE == kModeDgrad) {
      // copy chunks: k-core = 8 consecutive output channels of one flat position = one 16-byte vector of the plane-major
      // dz buffer.  cp.async straight into the A ring, na-1 chunks in flight per thread (no registers, no exposed latency).
      const int depth = g.na - 1;
      for (int it = 0; it < nchunks + depth; ++it) {
        if (it < nchunks) {
          const int bufc = it % g.na;
          const uint32_t phc = (uint32_t)(it / g.na) & 1u;
          const int ncols = chunk_cols(g, it);
          mbar_wait(&a_empty[bufc], phc ^ 1);
          unsigned char* ab = abuf0 + bufc * abuf_bytes;
#pragma unroll
          for (int k = 0; k < kRB; ++k) {
            if (offs[k] == -2) continue;
            const int b = r0 + k * kRowThreads;
#pragma unroll
            for (int cl = 0; cl < 2; ++cl) {
              const int pl = half * 2 + cl, grp = it * kPL + pl;
              if (pl >= ncols) continue;
              const bool ok = offs[k] >= 0 && grp * 8 < a.cq;
              cp_async16(ab + pl * plane_bytes + b * 16, ok ? a.dzf + ((long long)grp * g.L + offs[k]) * 16 : a.dzf, ok ? 16u : 0u);
            }
          }
        }
        cp_async_commit();
        if (it >= depth) {
          if (depth == 2) cp_async_wait<2>(); else cp_async_wait<1>();
          fence_proxy_async_smem();
          mbar_arrive(&a_full[(it - depth) % g.na]);
        }
      }
    }
    int buf = 0;
    uint32_t aphase = 0;
    ChunkSeq pseq(g.nsc, has_base ? g.nbc : 0);
    for (int pq = 0; MODE != kModeDgrad && pq < nchunks; ++pq) {
      const int q = pseq.next();
      unsigned char* ab = abuf0 + buf * abuf_bytes;
      trp.stamp();                                   // chunk start
      if (q < g.nsc && nb == 8) {
        float xv[kRB][2];
#pragma unroll
        for (int k = 0; k < kRB; ++k) { xv[k][0] = xnext[k][0]; xv[k][1] = xnext[k][1]; }
        if (q + 1 < g.nsc) fetch8(q + 1, xnext);
        mbar_wait(&a_empty[buf], aphase ^ 1);
        trp.stamp();                                 // buffer free
#pragma unroll
        for (int k = 0; k < kRB; ++k) {
          const int b = r0 + k * kRowThreads;
#pragma unroll
          for (int cl = 0; cl < 2; ++cl) {
            const int pl = half * 2 + cl;
            uint4 v = basis8(*B, g, xv[k][cl], offs[k] >= 0 && q * 4 + pl < cin);
            if (offs[k] != -2) reinterpret_cast<uint4*>(ab + pl * plane_bytes)[b] = v;
            if (MODE == kModeFwd && (ownmask >> k) & 1u)
              *reinterpret_cast<uint4*>(phi_row + (long long)(q * kPL + pl) * phi_ps + k * (kRowThreads * 16)) = v;
          }
        }
      } else if (q < g.nsc) {   // nb == 4: two channels share one 16-byte k-core, 8 channels per chunk
        float xv[kRB][4];
        const int c0 = q * 8 + half * 4;
        const float* xc = a.x_basis + (long long)c0 * HW;
#pragma unroll
        for (int k = 0; k < kRB; ++k)
#pragma unroll
          for (int cl = 0; cl < 4; ++cl)
            xv[k][cl] = (offs[k] >= 0 && c0 + cl < cin) ? __ldg(xc + (long long)cl * HW + offs[k]) : 0.0f;
        mbar_wait(&a_empty[buf], aphase ^ 1);
        trp.stamp();
#pragma unroll
        for (int k = 0; k < kRB; ++k) {
          if (offs[k] == -2) continue;
          const int b = r0 + k * kRowThreads;
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int pl = half * 2 + pp, c = q * 8 + pl * 2;
            uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
            if (offs[k] >= 0) {
              if (c < cin) lo = basis4(*B, xv[k][2 * pp]);
              if (c + 1 < cin) hi = basis4(*B, xv[k][2 * pp + 1]);
            }
            const uint4 v = make_uint4(lo.x, lo.y, hi.x, hi.y);
            reinterpret_cast<uint4*>(ab + pl * plane_bytes)[b] = v;
            if (MODE == kModeFwd && (ownmask >> k) & 1u)
              *reinterpret_cast<uint4*>(phi_row + (long long)(q * kPL + pl) * phi_ps + k * (kRowThreads * 16)) = v;
          }
        }
      } else {                  // base-activation chunk: k-core = 8 consecutive channels of one position
        // Four batches (2 planes x 2 row pairs) of 16 loads per thread, software-pipelined: the loads of batch b + 1 are in
        // flight while batch b is evaluated, so only the first batch's latency is exposed (the chunk used to expose four).
        const int bq = q - g.nsc;
        const int ncols = chunk_cols(g, q);
        auto loadb = [&](int b, float (&xv)[2][8]) {
          const int pl = half + 2 * (b >> 1), kk = (b & 1) * 2, grp = bq * kPL + pl;
          const float* xc = a.x_base + (long long)grp * 8 * HW;
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              xv[k][i] = (pl < ncols && offs[kk + k] >= 0 && grp * 8 + i < cin) ? __ldg(xc + (long long)i * HW + offs[kk + k]) : 0.0f;
        };
        auto emitb = [&](int b, const float (&xv)[2][8]) {
          const int pl = half + 2 * (b >> 1), kk = (b & 1) * 2, grp = bq * kPL + pl;
          if (pl >= ncols) return;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (offs[kk + k] == -2) continue;
            const int b_row = r0 + (kk + k) * kRowThreads;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = (offs[kk + k] >= 0 && grp * 8 + i < cin) ? tc_act(d.act, xv[k][i]) : 0.0f;
            const uint4 v = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            reinterpret_cast<uint4*>(ab + pl * plane_bytes)[b_row] = v;
            if (MODE == kModeFwd && (ownmask >> (kk + k)) & 1u)
              *reinterpret_cast<uint4*>(phi_row + (long long)(a.phi_base_plane0 + grp) * phi_ps + (kk + k) * (kRowThreads * 16)) = v;
          }
        };
        float xa[2][8], xb[2][8];
        loadb(0, xa);
        loadb(1, xb);
        mbar_wait(&a_empty[buf], aphase ^ 1);
        trp.stamp();
        emitb(0, xa);
        loadb(2, xa);
        emitb(1, xb);
        loadb(3, xb);
        emitb(2, xa);
        emitb(3, xb);
      }
      trp.stamp();                                   // stores done
      fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      mbar_arrive(&a_full[buf]);
      trp.stamp();                                   // arrived
      if (++buf == g.na) { buf = 0; aphase ^= 1; }
    }
  }
  if (PAIR && rank != 0 && warp >= kMmaWarp) {
    // ================================ peer CTA of a pair: relay "in place" to the leader, which issues the MMAs ========
    // warp 17: basis-row buffers, warp 18: weight stages
    if (warp == kMmaWarp) {
      int buf = 0;
      uint32_t aphase = 0;
      for (int pq = 0; pq < nchunks; ++pq) {
        mbar_wait(&a_full[buf], aphase);
        if (lane == 0) mbar_arrive_cluster(&pa_full[buf], 0);
        __syncwarp();
        if (++buf == g.na) { buf = 0; aphase ^= 1; }
      }
    } else if (warp == kMmaWarp + 1) {
      int stage = 0;
      uint32_t bphase = 0;
      for (int pq = 0; pq < nchunks; ++pq)
        for (int t = 0; t < T; t += g.tps) {
          mbar_wait(&b_full[stage], bphase);
          if (lane == 0) mbar_arrive_cluster(&pb_full[stage], 0);
          __syncwarp();
          if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
        }
    }
  } else if (warp >= kMmaWarp && warp - kMmaWarp < (g.nsub >= 2 ? 2 : 1)) {
    // ================================ MMA issuers =====================================================
    // The whole warp walks the pipeline with warp-uniform state; one elected lane issues tcgen05.mma / commit.
    // Per MMA only the 14-bit start-address fields of the two descriptors change (a few uniform integer adds).
    {
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kTileM : kTileM, g.ntile, 0, 0);
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);                                  // SBO = 128 B, version = 1
      const uint32_t a_lo_c = ((uint32_t)(g.plane_bytes >> 4) & 0x3FFFu) << 16;         // LBO = plane pitch
      const uint32_t b_lo_c = ((uint32_t)(nw) & 0x3FFFu) << 16;                         // LBO = (weight rows held) * 16 B
      const uint32_t a_k2 = (uint32_t)(2 * g.plane_bytes) >> 4, b_k2 = (uint32_t)(2 * nw);
      const uint32_t abuf_u = smem_u32(abuf0) >> 4, bst_u = smem_u32(bst0) >> 4;
      const uint32_t abuf_sz = (uint32_t)abuf_bytes >> 4, bst_sz = (uint32_t)bstage_bytes >> 4;
      const int ntile = g.ntile, kw = d.kw, SS = g.SS, tps = g.tps;
      const int mw = warp - kMmaWarp, nmw = g.nsub >= 2 ? 2 : 1;
      const int nsub = (g.nsub - mw + nmw - 1) / nmw;       // sub-tiles mw, mw + nmw, ... of this issuer
      int stage = 0, buf = 0;
      uint32_t bphase = 0, aphase = 0;
      KC_TRACER(trm, g_trace, 1, lane == 0 && mw == 0);
      ChunkSeq mseq(MODE == kModeDgrad ? 0 : g.nsc, MODE == kModeDgrad ? g.nbc : (has_base ? g.nbc : 0));
      for (int pq = 0; pq < nchunks; ++pq) {
        const int q = mseq.next();
        const int nk2 = chunk_cols(g, q) >> 1;
        const uint32_t btap_u = (uint32_t)(chunk_cols(g, q) * nw);         // 16-byte units per tap image
        trm.stamp();                                 // before a_full wait
        mbar_wait(&a_full[buf], aphase);
        if (PAIR) mbar_wait_cluster(&pa_full[buf], aphase);
        tc_fence_after();
        trm.stamp();                                 // a_full acquired
        uint32_t arow = abuf_u + (uint32_t)buf * abuf_sz;      // (address >> 4) of tap (0,0), sub-tile 0, k2 = 0
        int s = 0;
        for (int t = 0; t < T; t += tps) {
          mbar_wait(&b_full[stage], bphase);
          if (PAIR) mbar_wait_cluster(&pb_full[stage], bphase);
          tc_fence_after();
          trm.stamp();                               // b_full acquired
          uint32_t b_lo = b_lo_c + bst_u + (uint32_t)stage * bst_sz;
          const bool two = nk2 == 2;
          const bool leader = elect_one_sync();
          for (int tt = 0; tt < tps; ++tt) {
            const uint32_t first = (pq | (t + tt)) != 0 ? 1u : 0u;
            if (leader) {
              if (nsub == 2) issue_step<2, PAIR>(tmem_base, (uint32_t)ntile, a_lo_c + arow, b_lo, a_k2, b_k2, desc_hi, idesc, first, two, mw, nmw);
              else issue_step<1, PAIR>(tmem_base, (uint32_t)ntile, a_lo_c + arow, b_lo, a_k2, b_k2, desc_hi, idesc, first, two, mw, nmw);
            }
            b_lo += btap_u;
            // next tap: one position right, or first position of the next filter row (SS rows down)
            if (++s == kw) { s = 0; arow += (uint32_t)(SS - (kw - 1)); } else { arow += 1u; }
          }
          if (leader) {
            if (PAIR) {
              tc_commit_pair(&b_empty[stage]);
              if (t + tps >= T) tc_commit_pair(&a_empty[buf]);
            } else {
              tc_commit(&b_empty[stage]);
              if (t + tps >= T) tc_commit(&a_empty[buf]);
            }
          }
          __syncwarp();
          if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
        }
        if (++buf == g.na) { buf = 0; aphase ^= 1; }
      }
      if (elect_one_sync()) { if (PAIR) tc_commit_pair(acc_full); else tc_commit(acc_full); }
    }
    __syncwarp();
  } else if (warp == kLoaderWarp) {
    // ================================ weight loader (TMA engine bulk copies) ==========================
    if (PAIR) {
      // the image rows are [chunk][tap][k-core][ntile][8]: this CTA's half of a (tap, k-core) slab is nw * 16 contiguous
      // bytes; lane i copies the slabs i, i + 32, ... of a stage
      int stage = 0;
      uint32_t bphase = 0;
      const unsigned char* wimg = a.wp + (long long)nt * g.wimg_bytes_per_ntile + (long long)rank * nw * 16;
      const long long full_chunk_bytes = (long long)kPL * g.ntile * 16 * T;
      ChunkSeq lseq(g.nsc, has_base ? g.nbc : 0);
      for (int pq = 0; pq < nchunks; ++pq) {
        const int q = lseq.next();
        const int ncols = chunk_cols(g, q);
        const int nslab = g.tps * ncols;                                   // slabs per stage
        const unsigned char* wsrc = wimg + (long long)q * full_chunk_bytes;
        for (int t = 0; t < T; t += g.tps) {
          mbar_wait(&b_empty[stage], bphase ^ 1);
          if (lane == 0) mbar_arrive_expect_tx(&b_full[stage], (uint32_t)nslab * (uint32_t)nw * 16u);
          __syncwarp();
          unsigned char* dstb = bst0 + stage * bstage_bytes;
          for (int sl = lane; sl < nslab; sl += 32)
            bulk_g2s(dstb + (size_t)sl * nw * 16, wsrc + (size_t)sl * g.ntile * 16, (uint32_t)nw * 16u, &b_full[stage]);
          wsrc += (size_t)nslab * g.ntile * 16;
          if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t bphase = 0;
      const unsigned char* wimg = a.wp + (long long)nt * g.wimg_bytes_per_ntile;
      KC_TRACER(trl, g_trace, 2, true);
      // the packed image holds the chunks in id order (spline chunks, then base chunks; only the last one can be narrower)
      const long long full_chunk_bytes = (long long)kPL * g.ntile * 16 * T;
      ChunkSeq lseq(MODE == kModeDgrad ? 0 : g.nsc, MODE == kModeDgrad ? g.nbc : (has_base ? g.nbc : 0));
      for (int pq = 0; pq < nchunks; ++pq) {
        const int q = lseq.next();
        const unsigned char* wsrc = wimg + (long long)q * full_chunk_bytes;
        const uint32_t bytes = (uint32_t)chunk_cols(g, q) * g.ntile * 16u * (uint32_t)g.tps;
        for (int t = 0; t < T; t += g.tps) {
          mbar_wait(&b_empty[stage], bphase ^ 1);
          trl.stamp();                               // b_empty acquired
          mbar_arrive_expect_tx(&b_full[stage], bytes);
          bulk_g2s(bst0 + stage * bstage_bytes, wsrc, bytes, &b_full[stage]);
          wsrc += bytes;
          if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
        }
      }
    }
    __syncwarp();
  }
  if (MODE == kModeFwd && warp < 16) {
    // ================================ epilogue: TMEM -> registers -> z (fp32 NCHW) =====================
    // All 16 producer warps: warp w reads the TMEM lanes of its hardware quarter (w % 4) and every fourth group of 16
    // columns (w / 4); a store instruction writes one output channel of 32 consecutive positions (128 B).
    tre.stamp();                                     // producer work done
    mbar_wait(acc_full, 0);
    tc_fence_after();
    tre.stamp();                                     // accumulator ready
    const int HoWo = d.ho * d.wo;
    const int n0 = nt * g.ntile;
    const int quarter = warp & 3, cgrp = warp >> 2;
    for (int i = 0; i < g.nsub; ++i) {
      const long long q = m0 + i * kTileM + quarter * 32 + lane;
      bool valid = false;
      long long zoff = 0;
      if (q < g.L) {
        const unsigned uq = (unsigned)q, n = uq / (unsigned)g.IMG, rem = uq - n * (unsigned)g.IMG;
        const unsigned y = rem / (unsigned)g.P, x = rem - y * (unsigned)g.P;
        unsigned yo = y, xo = x;
        bool on_grid = true;
        if (d.stride_h != 1 || d.stride_w != 1) {
          yo = y / (unsigned)d.stride_h; xo = x / (unsigned)d.stride_w;
          on_grid = yo * (unsigned)d.stride_h == y && xo * (unsigned)d.stride_w == x;
        }
        if (on_grid && yo < (unsigned)d.ho && xo < (unsigned)d.wo) { valid = true; zoff = (long long)n * d.z_batch_stride + yo * d.wo + xo; }
      }
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * g.ntile);
      for (int c0 = cgrp * 16; c0 < g.ntile; c0 += 64) {
        uint32_t r[16];
        tmem_ld16(trow + (uint32_t)c0, r);          // ntile is a multiple of 16
        tmem_ld_wait();
        float* zp = a.z + zoff + (long long)(n0 + c0) * HoWo;
        const int lim = min(16, d.cout - (n0 + c0));
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < lim) zp[(long long)j * HoWo] = __uint_as_float(r[j]);
        }
      }
    }
    tre.stamp();                                     // z written
  }
  if (MODE == kModeDgrad && warp < 16) {
    // ================================ dgrad epilogue: dPhi (TMEM) x analytic basis derivative -> dx =======
    // Columns of this N tile are (channel cl, j) with j < nb the basis gradients and j == nb the base-branch gradient;
    // dPhi never leaves the SM (the reference's autograd materialises it, ~50 elementwise backward launches).
    // All 16 producer warps take part: warp w reads TMEM lanes 32*(w%4).. (its hardware quarter) and handles the four
    // channels 4*(w/4) .. +3.  The loop over sub-tiles is a real loop with a compact body (the rare generic-basis and
    // GRAM paths are out of line): a fully unrolled version was instruction-cache bound (40k cycles per CTA).
    const int nb = d.nb, wb = nb + (has_base ? 1 : 0);
    const bool alias = a.dx_base == a.dx_basis, same_x = a.x_base == a.x_basis;
    const bool gram = d.basis == KC_BASIS_GRAM && a.dbeta != nullptr;
    const int quarter = warp & 3, cgrp = warp >> 2;
    const int c0 = nt * 16 + cgrp * 4;
    tre.stamp();                                     // epilogue entered (producer loop done)
    if (g.fast_cubic && !gram && same_x) {
      // ---- closed-form cubic path (nb == 8): everything stays in registers ----------------------------------------
      // x of sub-tile i+1 is fetched while sub-tile i is evaluated, and the TMEM read of channel c+1 is issued before
      // channel c is evaluated, so neither latency is exposed.
      const int nint = B->nparams - 1, act = d.act;
      const float t0 = g.t0, inv_h = g.inv_h;
      const int nch = min(4, d.cin - c0);                             // warp-uniform, may be <= 0
      const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cgrp * 4 * wb);
      const float* xch = a.x_basis + (long long)c0 * HW;
      float* dxo = a.dx_basis + (long long)c0 * HW;
      float* dxb = (!alias && has_base && a.dx_base != nullptr) ? a.dx_base + (long long)c0 * HW : nullptr;
      auto locate = [&](int i) -> int {                               // element offset of the flat position in x, or -1
        const long long q = m0 + i * kTileM + quarter * 32 + lane;
        if (q >= g.L) return -1;
        const unsigned uq = (unsigned)q, n = uq / (unsigned)g.IMG, rem = uq - n * (unsigned)g.IMG;
        const unsigned y = rem / (unsigned)g.P, x = rem - y * (unsigned)g.P;
        return (y < (unsigned)d.h && x < (unsigned)d.w) ? (int)((long long)n * d.x_batch_stride + y * d.w + x) : -1;
      };
      auto fetch = [&](int off, float (&xv)[4]) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) xv[c4] = (off >= 0 && c4 < nch) ? ldg_early(xch + off + (long long)c4 * HW) : 0.0f;
      };
      auto ld9 = [&](uint32_t taddr, uint32_t (&r)[9]) {
        tmem_ld8(taddr, r);
        if (has_base) tmem_ld1(taddr + 8u, r + 8); else r[8] = 0u;
      };
      float xn[4];
      int offn = locate(0);
      fetch(offn, xn);
      tre.stamp();
      mbar_wait(acc_full, 0);
      tc_fence_after();
      tre.stamp();                                   // accumulator ready
#pragma unroll 1
      for (int i = 0; i < g.nsub; ++i) {
        float xc[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) xc[c4] = xn[c4];
        const int off = offn;
        if (i + 1 < g.nsub) { offn = locate(i + 1); fetch(offn, xn); }
        const uint32_t trow = tq + (uint32_t)(i * g.ntile);
        auto emit = [&](int c4, const uint32_t (&r)[9]) {
          if (off < 0) return;
          const float gs = cubic8_dot_grad(xc[c4], t0, inv_h, nint, r);
          const float gb = has_base ? __uint_as_float(r[8]) * act_grad_fast(act, xc[c4]) : 0.0f;
          const long long o = off + (long long)c4 * HW;
          if (dxb != nullptr) { dxo[o] = gs; dxb[o] = gb; } else { dxo[o] = alias ? gs + gb : gs; }
        };
        uint32_t ra[9], rb[9];
        if (nch > 0) ld9(trow, ra);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; c4 += 2) {
          if (c4 + 1 < nch) ld9(trow + (uint32_t)((c4 + 1) * wb), rb);
          if (c4 < nch) emit(c4, ra);
          tmem_ld_wait();
          if (c4 + 2 < nch) ld9(trow + (uint32_t)((c4 + 2) * wb), ra);
          if (c4 + 1 < nch) emit(c4 + 1, rb);
          tmem_ld_wait();
        }
      }
    } else {
      float dbl[KC_MAX_BASIS];
#pragma unroll
      for (int j = 0; j < KC_MAX_BASIS; ++j) dbl[j] = 0.0f;
      bool waited = false;
  #pragma unroll 1
      for (int i = 0; i < g.nsub; ++i) {
        const long long q = m0 + i * kTileM + quarter * 32 + lane;
        long long off = -1;
        if (q < g.L) {
          int n = (int)(q / g.IMG);
          int rem = (int)(q - (long long)n * g.IMG);
          int y = rem / g.P, x = rem - y * g.P;
          if (y < d.h && x < d.w) off = (long long)n * d.x_batch_stride + y * d.w + x;
        }
        float xs[4], xb[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const bool ok = off >= 0 && c0 + c4 < d.cin;
          xs[c4] = ok ? ldg_early(a.x_basis + off + (long long)(c0 + c4) * HW) : 0.0f;
          xb[c4] = (ok && has_base && !same_x) ? ldg_early(a.x_base + off + (long long)(c0 + c4) * HW) : xs[c4];
        }
        if (!waited) {                                 // the x loads of the first sub-tile fly while the MMAs drain
          tre.stamp();
          mbar_wait(acc_full, 0);
          tc_fence_after();
          tre.stamp();                                 // accumulator ready
          waited = true;
        }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * g.ntile);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int c = c0 + c4;
          if (c >= d.cin) break;                       // warp-uniform
          uint32_t r[16];
          tmem_ld16(trow + (uint32_t)((cgrp * 4 + c4) * wb), r);
          tmem_ld_wait();
          if (off >= 0) {
            const long long o = off + (long long)c * HW;
            float gs;
            if (g.fast_cubic) gs = cubic8_dot_grad(xs[c4], g.t0, g.inv_h, B->nparams - 1, r);
            else gs = dgrad_generic(*B, xs[c4], r, gram ? dbl : nullptr);
            float gb = 0.0f;
            if (has_base) {
              float ga = __uint_as_float(r[8]);          // nb == 8 (the common case); other widths: pick column nb
              if (nb != 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (j == nb) ga = __uint_as_float(r[j]);
              }
              gb = ga * act_grad_fast(d.act, xb[c4]);
            }
            if (alias) {
              a.dx_basis[o] = gs + gb;
            } else {
              a.dx_basis[o] = gs;
              if (has_base && a.dx_base != nullptr) a.dx_base[o] = gb;
            }
          }
        }
      }
      if (gram) {
        // deterministic: every epilogue warp writes its own row of the partials buffer (zeros where it had no work);
        // kc_dbeta_reduce_kernel adds the rows in fixed order
        const long long row = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16 + warp;
#pragma unroll
        for (int nn = 0; nn < KC_MAX_BASIS; ++nn) {
          const float v = kc_warp_sum(dbl[nn]);
          if (lane == nn) a.dbeta[(1 + row) * KC_MAX_BASIS + nn] = (nn >= 1 && nn <= nb - 2) ? v : 0.0f;
        }
      }
    }
    tre.stamp();                                     // dx written
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) { cluster_arrive(); cluster_wait(); }        // both CTAs are done with the pair's tensor memory and shared memory
  if (warp == kMmaWarp) { if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)g.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols); }
}

// d(basis_j)/dx for the RBF and Chebyshev families, j < 8, registers only (every index is a compile-time constant):
// the epilogue of the persistent dgrad for the layers that have no closed-form cubic basis.
// Returns true when the gradient is masked to exactly zero (Chebyshev clamp active).
__device__ __forceinline__ bool tc_basis_grad8(const KcBasisCtx& B, float x, float (&dphi)[8]) {
  const int nb = B.nb;
  if (B.kind == KC_BASIS_RBF) {
    const float inv_den = __fdividef(1.0f, B.p[nb]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float q = (x - B.p[j]) * inv_den;
      dphi[j] = j < nb ? __expf(-(q * q)) * (-2.0f * q * inv_den) : 0.0f;
    }
    return false;
  } else {      // KC_BASIS_CHEBY: d T_j(c) / dx = j U_{j-1}(c) (1 - t^2), c = clamp(tanh x)
    const float lo = -1.0f + 1e-7f, hi = 1.0f - 1e-7f;
    const float t = tc_tanh(x);
    float c = fminf(fmaxf(t, lo), hi);
    if (t != t) c = t;
    const float dc = (t < lo || t > hi) ? 0.0f : 1.0f - t * t;
    float U0 = 0.0f, U1 = 1.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dphi[j] = j < nb ? (float)j * U0 * dc : 0.0f;
      const float U2 = 2.0f * c * U1 - U0;
      U0 = U1; U1 = U2;
    }
    return kc_cheby_clamped(t);
  }
}

// GRAM (gram_kan_layers.py:150-181), registers only: d SiLU(p_j(t)) / dx for j < 8 and, for the d/d beta_weights reduction, the
// sums over j of g_j * SiLU'(p_j) * dp_j/d beta(n, n+1) for n = 1 .. nb-2 (accumulated into dbl[n - 1] with the coefficient
// of kc_gram_coef applied).  p_0 = 1, p_1 = t, p_i = t p_{i-1} - beta_{i-1} p_{i-2};  q^n_i = dp_i/d beta_n obeys
// q_i = t q_{i-1} - [i-1 == n] p_{i-2} - beta_{i-1} q_{i-2}.  Every index is a compile-time constant after unrolling.
__device__ __forceinline__ void tc_gram_grad8(const KcBasisCtx& B, float x, const uint32_t* r, float (&dphi)[8], float (&dbl)[6],
                                              bool want_dbeta) {
  const int nb = B.nb;
  const bool pre = kc_gram_presquashed(B);
  const float t = pre ? x : tc_tanh(x), dt = pre ? 1.0f : 1.0f - t * t;
  float p[8], sgp[8];                      // p_j and SiLU'(p_j)
  float p0 = 1.0f, p1 = t, d0 = 0.0f, d1 = 1.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float sg = tc_sigmoid(p0);
    p[j] = p0;
    sgp[j] = sg * fmaf(p0, 1.0f - sg, 1.0f);
    dphi[j] = j < nb ? sgp[j] * d0 * dt : 0.0f;
    const float b = B.gbeta[j + 1 < KC_MAX_BASIS ? j + 1 : 0];
    const float p2 = t * p1 - b * p0, d2 = p1 + t * d1 - b * d0;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
  if (!want_dbeta) return;
#pragma unroll
  for (int n = 1; n <= 6; ++n) {
    if (n <= nb - 2) {
      float q0 = 0.0f, q1 = 0.0f, s = 0.0f;        // q_0 = q_1 = 0
#pragma unroll
      for (int i = 2; i < 8; ++i) {
        if (i < nb) {
          const float q2 = t * q1 - ((i - 1 == n) ? p[i - 2] : 0.0f) - B.gbeta[i - 1] * q0;
          s = fmaf(__uint_as_float(r[i]) * sgp[i], q2, s);
          q0 = q1; q1 = q2;
        }
      }
      dbl[n - 1] = fmaf(s, kc_gram_coef(n), dbl[n - 1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent GEMM kernel with both operands fed by the TMA engine: one CTA per SM walks a list of (256-position tile, N tile)
// pairs.  The TMEM holds TWO accumulator sets (2 sub-tiles x <= 128 columns each), so the 16 epilogue warps drain tile t
// while the MMA warps already accumulate tile t+1.
//   warps 0-15 epilogue | 16 A loader (strips of a plane-major flat buffer) | 17 weight loader | 18-19 MMA issuers
// FAM 1: dgrad, closed-form cubic basis   (A = dz_flat, N tile = cpt channels x (nb + base) columns padded to 128)
// FAM 0: dgrad, RBF / Chebyshev basis
// FAM 3: dgrad, Gram basis (adds the deterministic d/d beta_weights partial rows: one row per CTA and epilogue warp)
// FAM 2: forward of a 1x1 convolution from the saved basis rows (A = phi, N tile = output channels): pointwise layers have one
//        tap of MMA work per K chunk, so evaluating the basis inside the GEMM kernel leaves the tensor cores idle; the basis is
//        written once by the pre-pass (it is needed for the weight gradient anyway) and this kernel streams it back.
// ---------------------------------------------------------------------------------------------------------
constexpr int kDgThreads = 640, kDgEpiWarps = 16, kDgProdWarp0 = 16, kDgLoaderWarp = 17, kDgMmaWarp0 = 18;
constexpr int kDgBars = 3 * kMaxA + 3 * kMaxBStages + 4;

// PAIR: two CTAs of a cluster work on two neighbouring position tiles and the same N tile with tcgen05 cta_group::2 MMAs
// (M = 256): each CTA loads its own dz rows and HALF of every weight stage (64 of the 128 N rows), the leader CTA issues the
// MMAs for both.  An M = 128, N = 128 MMA is bound by the operand fetch from shared memory (74 cycles instead of 64); the
// pair halves the weight fetch per CTA (64 cycles, tools/mma_rate_2cta.py) and the weight traffic from L2.
template <int FAM, bool PAIR>
__global__ void __launch_bounds__(kDgThreads, 1) kc_dgrad_persistent_kernel(const __grid_constant__ TcFwdArgs a) {
  constexpr bool CUBIC = FAM == 1, FWD = FAM == 2, GRAMF = FAM == 3;
  static_assert(!(FWD && PAIR), "the pointwise forward runs on single CTAs");
  extern __shared__ __align__(128) unsigned char smem[];
  const kc_desc& d = a.d;
  const TcGeom& g = a.g;
  const int abuf_bytes = kPL * g.plane_bytes;
  const int nw = PAIR ? g.ntile / 2 : g.ntile;     // weight rows (of the MMA's N) this CTA holds
  const int btap_bytes = kPL * nw * 16;
  const int bstage_bytes = g.tps * btap_bytes;
  unsigned char* abuf0 = smem;
  unsigned char* bst0 = abuf0 + g.na * abuf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bst0 + g.bstages * bstage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kMaxA;
  uint64_t* b_full = bars + 2 * kMaxA;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* acc_full = b_empty + kMaxBStages;      // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint64_t* pa_full = acc_empty + 2;               // [kMaxA]        pair, leader: the peer's dz rows have landed
  uint64_t* pb_full = pa_full + kMaxA;             // [kMaxBStages]  pair, leader: the peer's half of a weight stage has landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + kDgBars);
  KcBasisCtx* B = reinterpret_cast<KcBasisCtx*>(reinterpret_cast<unsigned char*>(tmem_ptr) + 16);      // !CUBIC only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = d.kh * d.kw, HW = d.h * d.w;
  const bool has_base = d.act != KC_ACT_NONE;
  const int wb = d.nb + (has_base ? 1 : 0);
  const int nchunks = FWD ? g.nsc + (has_base ? g.nbc : 0) : g.nbc;
  // tile walk: a (pair of) CTA(s) takes every tstep-th (position tile [pair], N tile); in a pair CTA `rank` owns position
  // tile 2 * m + rank (an odd tile count leaves the last peer with positions >= L: nothing is loaded or stored for them)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const long long ntiles = (PAIR ? (g.mtiles + 1) / 2 : g.mtiles) * g.n_ntiles;
  const long long tile0 = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long tstep = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  auto mt_of = [&](long long tile) -> long long { const long long m = tile / g.n_ntiles; return PAIR ? 2 * m + rank : m; };
  if (FAM == 0 || GRAMF) kc_load_basis_ctx(B, d, a.beta);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2); mbar_init(&pa_full[i], 1); }
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 2); mbar_init(&pb_full[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 2); mbar_init(&acc_empty[i], PAIR ? 2 * kDgEpiWarps : kDgEpiWarps); }
    fence_barrier_init();
  }
  if (warp == kDgMmaWarp0) { if (PAIR) tmem_alloc_pair(tmem_ptr, 512u); else tmem_alloc(tmem_ptr, 512u); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) { cluster_arrive(); cluster_wait(); }        // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kDgProdWarp0) {
    // ================================ dz loader: bulk copies (TMA engine) of the flat dz planes ====================
    // The rows of a chunk are kh strips of consecutive flat positions per plane: lane (r, pl) copies strip r of plane pl
    // with one cp.async.bulk.  Rows outside the flat sequence (first / last position tiles only) are zero-filled.
    const int r = lane / kPL, pl = lane % kPL;
    const int len = r < d.kh - 1 ? g.SS : g.nrows - (d.kh - 1) * g.SS;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    KC_TRACER(trp, g_trace, 0, lane == 0);
    int buf = 0;
    uint32_t ph = 1;
    for (long long tile = tile0; tile < ntiles; tile += tstep) {
      const long long qs = mt_of(tile) * g.mcta - (long long)g.ph * g.P - g.pw + (long long)r * g.P;
      long long lo = qs < 0 ? -qs : 0, hi = qs + len > g.L ? g.L - qs : len;
      if (lo > len) lo = len;
      if (hi < lo) hi = lo;
      for (int q = 0; q < nchunks; ++q) {
        const int ncols = chunk_cols(g, q);
        const bool active = r < d.kh && pl < ncols;
        trp.stamp();
        mbar_wait(&a_empty[buf], ph);
        trp.stamp();
        unsigned char* dst = abuf0 + buf * abuf_bytes + pl * g.plane_bytes + (r * g.SS) * 16;
        const bool edge = active && (lo > 0 || hi < len);
        if (__any_sync(0xffffffffu, edge)) {
          if (edge) {
            for (int i = 0; i < (int)lo; ++i) *reinterpret_cast<uint4*>(dst + i * 16) = zero4;
            for (int i = (int)hi; i < len; ++i) *reinterpret_cast<uint4*>(dst + i * 16) = zero4;
          }
          fence_proxy_async_smem();
        }
        const uint32_t mybytes = active ? (uint32_t)(hi - lo) * 16u : 0u;
        const uint32_t total = __reduce_add_sync(0xffffffffu, mybytes);
        if (lane == 0) mbar_arrive_expect_tx(&a_full[buf], total);
        __syncwarp();
        // plane of the flat buffer: dz_flat planes are consecutive; in phi the base planes follow the (padded) spline planes
        const int plane = (!FWD || q < g.nsc) ? q * kPL + pl : a.phi_base_plane0 + (q - g.nsc) * kPL + pl;
        if (mybytes != 0u)
          bulk_g2s(dst + lo * 16, a.dzf + ((long long)plane * g.L + qs + lo) * 16, mybytes, &a_full[buf]);
        if (++buf == g.na) { buf = 0; ph ^= 1; }
      }
    }
  } else if (warp == kDgLoaderWarp) {
    // ================================ weight loader (TMA engine bulk copies) ==========================
    if (lane == 0) {
      int stage = 0;
      uint32_t bphase = 1;
      KC_TRACER(trl, g_trace, 2, true);
      for (long long tile = tile0; tile < ntiles; tile += tstep) {
        const int nt = (int)(tile % g.n_ntiles);
        // pair: the image of an N tile is stored as two halves of nw rows each, CTA `rank` streams its half
        const unsigned char* wsrc = a.wp + (long long)nt * g.wimg_bytes_per_ntile + (PAIR ? (long long)rank * (g.wimg_bytes_per_ntile / 2) : 0);
        for (int q = 0; q < nchunks; ++q) {
          const uint32_t bytes = (uint32_t)chunk_cols(g, q) * (uint32_t)nw * 16u * (uint32_t)g.tps;
          for (int t = 0; t < T; t += g.tps) {
            mbar_wait(&b_empty[stage], bphase);
            trl.stamp();
            mbar_arrive_expect_tx(&b_full[stage], bytes);
            bulk_g2s(bst0 + stage * bstage_bytes, wsrc, bytes, &b_full[stage]);
            wsrc += bytes;
            if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= kDgMmaWarp0 && rank != 0) {
    // ================================ peer CTA of a pair: relay "landed" to the leader, which issues the MMAs ==========
    // warp 18: dz buffers, warp 19: weight stages
    if (warp == kDgMmaWarp0) {
      int buf = 0;
      uint32_t aphase = 0;
      for (long long tile = tile0; tile < ntiles; tile += tstep)
        for (int q = 0; q < nchunks; ++q) {
          mbar_wait(&a_full[buf], aphase);
          if (lane == 0) mbar_arrive_cluster(&pa_full[buf], 0);
          __syncwarp();
          if (++buf == g.na) { buf = 0; aphase ^= 1; }
        }
    } else {
      int stage = 0;
      uint32_t bphase = 0;
      for (long long tile = tile0; tile < ntiles; tile += tstep)
        for (int q = 0; q < nchunks; ++q)
          for (int t = 0; t < T; t += g.tps) {
            mbar_wait(&b_full[stage], bphase);
            if (lane == 0) mbar_arrive_cluster(&pb_full[stage], 0);
            __syncwarp();
            if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
          }
    }
  } else if (warp >= kDgMmaWarp0) {
    // ================================ MMA issuers: warp mw accumulates sub-tile mw ====================
    const int mw = warp - kDgMmaWarp0;
    const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kTileM : kTileM, g.ntile, 0, 0);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_c = ((uint32_t)(g.plane_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t b_lo_c = ((uint32_t)(nw) & 0x3FFFu) << 16;
    const uint32_t a_k2 = (uint32_t)(2 * g.plane_bytes) >> 4, b_k2 = (uint32_t)(2 * nw);
    const uint32_t abuf_u = smem_u32(abuf0) >> 4, bst_u = smem_u32(bst0) >> 4;
    const uint32_t abuf_sz = (uint32_t)abuf_bytes >> 4, bst_sz = (uint32_t)bstage_bytes >> 4;
    const int ntile = g.ntile, kw = d.kw, SS = g.SS, tps = g.tps;
    int stage = 0, buf = 0;
    uint32_t bphase = 0, aphase = 0;
    uint32_t it = 0;
    KC_TRACER(trm, g_trace, 1, lane == 0 && mw == 0);
    for (long long tile = tile0; tile < ntiles; tile += tstep, ++it) {
      const uint32_t acc = it & 1u;
      trm.stamp();                                           // tile start
      // the epilogue (of both CTAs in a pair) has drained this accumulator set
      if (PAIR) mbar_wait_cluster(&acc_empty[acc], ((it >> 1) & 1u) ^ 1u); else mbar_wait(&acc_empty[acc], ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      trm.stamp();                                           // accumulator set free
      const uint32_t tacc = tmem_base + acc * (uint32_t)(2 * ntile);
      for (int q = 0; q < nchunks; ++q) {
        const int nk2 = chunk_cols(g, q) >> 1;
        const uint32_t btap_u = (uint32_t)(chunk_cols(g, q) * nw);
        mbar_wait(&a_full[buf], aphase);
        if (PAIR) mbar_wait_cluster(&pa_full[buf], aphase);
        tc_fence_after();
        trm.stamp();                                         // a_full acquired
        uint32_t arow = abuf_u + (uint32_t)buf * abuf_sz;
        int s = 0;
        for (int t = 0; t < T; t += tps) {
          mbar_wait(&b_full[stage], bphase);
          if (PAIR) mbar_wait_cluster(&pb_full[stage], bphase);
          tc_fence_after();
          trm.stamp();                                       // b_full acquired
          uint32_t b_lo = b_lo_c + bst_u + (uint32_t)stage * bst_sz;
          const bool two = nk2 == 2;
          const bool leader = elect_one_sync();
          for (int tt = 0; tt < tps; ++tt) {
            const uint32_t first = (q | (t + tt)) != 0 ? 1u : 0u;
            if (leader) issue_step<1, PAIR>(tacc, (uint32_t)ntile, a_lo_c + arow, b_lo, a_k2, b_k2, desc_hi, idesc, first, two, mw, 1);
            b_lo += btap_u;
            if (++s == kw) { s = 0; arow += (uint32_t)(SS - (kw - 1)); } else { arow += 1u; }
          }
          if (leader) {
            if (PAIR) {
              tc_commit_pair(&b_empty[stage]);
              if (t + tps >= T) tc_commit_pair(&a_empty[buf]);
            } else {
              tc_commit(&b_empty[stage]);
              if (t + tps >= T) tc_commit(&a_empty[buf]);
            }
          }
          __syncwarp();
          if (++stage == g.bstages) { stage = 0; bphase ^= 1; }
        }
        if (++buf == g.na) { buf = 0; aphase ^= 1; }
      }
      if (elect_one_sync()) { if (PAIR) tc_commit_pair(&acc_full[acc]); else tc_commit(&acc_full[acc]); }
      __syncwarp();
    }
  } else {
    if (FWD) {
      // ============================== epilogue (forward from phi): TMEM -> z (fp32 NCHW) ===========================
      const int quarter = warp & 3, cgrp = warp >> 2;
      const int HoWo = d.ho * d.wo;
      const long long nsteps = tile0 < ntiles ? ((ntiles - tile0 + tstep - 1) / tstep) * 2 : 0;
#pragma unroll 1
      for (long long step = 0; step < nsteps; ++step) {
        const uint32_t it = (uint32_t)(step >> 1), acc = it & 1u, sub = (uint32_t)(step & 1);
        const long long tile = tile0 + (step >> 1) * tstep;
        const long long mt = tile / g.n_ntiles;
        const int n0 = (int)(tile - mt * g.n_ntiles) * g.ntile;
        const long long q = mt * g.mcta + sub * kTileM + quarter * 32 + lane;
        bool valid = false;
        long long zoff = 0;
        if (q < g.L) {
          const unsigned uq = (unsigned)q, n = uq / (unsigned)g.IMG, rem = uq - n * (unsigned)g.IMG;
          const unsigned y = rem / (unsigned)g.P, x = rem - y * (unsigned)g.P;
          unsigned yo = y, xo = x;
          bool on_grid = true;
          if (d.stride_h != 1 || d.stride_w != 1) {
            yo = y / (unsigned)d.stride_h; xo = x / (unsigned)d.stride_w;
            on_grid = yo * (unsigned)d.stride_h == y && xo * (unsigned)d.stride_w == x;
          }
          if (on_grid && yo < (unsigned)d.ho && xo < (unsigned)d.wo) { valid = true; zoff = (long long)n * d.z_batch_stride + yo * d.wo + xo; }
        }
        if (sub == 0) {
          mbar_wait(&acc_full[acc], (it >> 1) & 1u);
          tc_fence_after();
        }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (uint32_t)(2 * g.ntile) + sub * (uint32_t)g.ntile;
        for (int c0 = cgrp * 16; c0 < g.ntile; c0 += 64) {
          uint32_t r[16];
          tmem_ld16(trow + (uint32_t)c0, r);
          tmem_ld_wait();
          float* zp = a.z + zoff + (long long)(n0 + c0) * HoWo;
          const int lim = min(16, d.cout - (n0 + c0));
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < lim) zp[(long long)j * HoWo] = __uint_as_float(r[j]);
          }
        }
        if (sub == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
      }
    } else {
    // ================================ epilogue: dPhi (TMEM) x analytic basis derivative -> dx =========
    // warp w: TMEM lanes of quarter w % 4, channel group w / 4 (the cpt channels of the tile are split four ways)
    constexpr int kCh = CUBIC ? 4 : 8;                             // channels per warp (cpt <= 4 * kCh)
    const bool alias = a.dx_base == a.dx_basis, same_x = a.x_base == a.x_basis;
    const int quarter = warp & 3, cgrp = warp >> 2;
    const int chs = (cgrp * g.cpt) / 4, chn = ((cgrp + 1) * g.cpt) / 4 - chs;
    const int nint = d.nparams - 1, act = d.act, nb = d.nb;
    const float t0 = g.t0, inv_h = g.inv_h;
    const long long nsteps = tile0 < ntiles ? ((ntiles - tile0 + tstep - 1) / tstep) * 2 : 0;     // (tile, sub-tile) pairs of this CTA
    auto locate = [&](long long step, int& c0) -> int {           // x offset of this lane's position in step, or -1
      const long long tile = tile0 + (step >> 1) * tstep;
      c0 = (int)(tile % g.n_ntiles) * g.cpt + chs;
      const long long q = mt_of(tile) * g.mcta + (step & 1) * kTileM + quarter * 32 + lane;
      if (q >= g.L) return -1;
      const unsigned uq = (unsigned)q, n = uq / (unsigned)g.IMG, rem = uq - n * (unsigned)g.IMG;
      const unsigned y = rem / (unsigned)g.P, x = rem - y * (unsigned)g.P;
      return (y < (unsigned)d.h && x < (unsigned)d.w) ? (int)((long long)n * d.x_batch_stride + y * d.w + x) : -1;
    };
    auto fetch = [&](int off, int c0, float (&xv)[kCh]) {
#pragma unroll
      for (int c4 = 0; c4 < kCh; ++c4)
        xv[c4] = (off >= 0 && c4 < chn && c0 + c4 < d.cin) ? ldg_early(a.x_basis + off + (long long)(c0 + c4) * HW) : 0.0f;
    };
    auto ld9 = [&](uint32_t taddr, uint32_t (&r)[9]) {            // the wb <= 9 gradient columns of one channel
      tmem_ld8(taddr, r);
      if (wb > 8) tmem_ld1(taddr + 8u, r + 8); else r[8] = 0u;
    };
    KC_TRACER(tre, g_trace, 3, threadIdx.x == 0);
    float dbl[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};          // GRAM: this thread's d/d beta(n, n+1) sums, n = 1 .. 6
    float xn[kCh];
    int c0n = 0;
    int offn = nsteps > 0 ? locate(0, c0n) : -1;
    fetch(offn, c0n, xn);
#pragma unroll 1
    for (long long step = 0; step < nsteps; ++step) {
      const uint32_t it = (uint32_t)(step >> 1), acc = it & 1u, sub = (uint32_t)(step & 1);
      float xc[kCh];
#pragma unroll
      for (int c4 = 0; c4 < kCh; ++c4) xc[c4] = xn[c4];
      const int off = offn, c0 = c0n;
      if (step + 1 < nsteps) { offn = locate(step + 1, c0n); fetch(offn, c0n, xn); }
      if (sub == 0) {
        tre.stamp();
        mbar_wait(&acc_full[acc], (it >> 1) & 1u);
        tc_fence_after();
        tre.stamp();                                           // accumulator ready
      }
      const int nch = min(chn, d.cin - c0);                      // warp-uniform, may be <= 0
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (uint32_t)(2 * g.ntile) + sub * (uint32_t)g.ntile +
                            (uint32_t)(chs * wb);
      auto emit = [&](int c4, const uint32_t (&r)[9]) {
        if (off < 0) return;
        const long long o = off + (long long)(c0 + c4) * HW;
        float gs, ga;                                           // spline-branch gradient, incoming base-branch gradient
        if (CUBIC) {
          gs = cubic8_dot_grad(xc[c4], t0, inv_h, nint, r);
          ga = __uint_as_float(r[8]);
        } else {
          float dphi[8];
          bool masked = false;
          if (GRAMF) tc_gram_grad8(*B, xc[c4], r, dphi, dbl, a.dbeta != nullptr);
          else masked = tc_basis_grad8(*B, xc[c4], dphi);
          gs = 0.0f;
          ga = __uint_as_float(r[8]);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            if (jj < nb) gs = fmaf(__uint_as_float(r[jj]), dphi[jj], gs);      // columns >= nb belong to the base branch / next channel
            if (jj == nb) ga = __uint_as_float(r[jj]);
          }
          if (masked) gs = 0.0f;
        }
        float gb = 0.0f;
        if (has_base) gb = ga * act_grad_fast(act, same_x ? xc[c4] : __ldg(a.x_base + o));
        if (alias) {
          a.dx_basis[o] = gs + gb;
        } else {
          a.dx_basis[o] = gs;
          if (has_base && a.dx_base != nullptr) a.dx_base[o] = gb;
        }
      };
      uint32_t ra[9], rb[9];
      if (nch > 0) ld9(trow, ra);
      tmem_ld_wait();
#pragma unroll
      for (int c4 = 0; c4 < kCh; c4 += 2) {
        if (c4 + 1 < nch) ld9(trow + (uint32_t)((c4 + 1) * wb), rb);
        if (c4 < nch) emit(c4, ra);
        tmem_ld_wait();
        if (c4 + 2 < nch) ld9(trow + (uint32_t)((c4 + 2) * wb), ra);
        if (c4 + 1 < nch) emit(c4 + 1, rb);
        tmem_ld_wait();
      }
      if (sub == 1) {                                            // both sub-tiles of this accumulator set are in registers / stored
        tre.stamp();                                             // tile written
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(&acc_empty[acc], 0); else mbar_arrive(&acc_empty[acc]); }
      }
    }
    if (GRAMF && a.dbeta != nullptr) {
      // deterministic d/d beta_weights: one partial row per (CTA, epilogue warp); kc_dbeta_reduce_kernel adds them in order
      const long long row = (long long)blockIdx.x * kDgEpiWarps + warp;
#pragma unroll
      for (int nn = 0; nn < KC_MAX_BASIS; ++nn) {
        const float v = (nn >= 1 && nn <= 6) ? kc_warp_sum(dbl[nn >= 1 && nn <= 6 ? nn - 1 : 0]) : 0.0f;
        if (lane == nn) a.dbeta[(1 + row) * KC_MAX_BASIS + nn] = (nn >= 1 && nn <= nb - 2) ? v : 0.0f;
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) { cluster_arrive(); cluster_wait(); }        // both CTAs are done with the pair's tensor memory and shared memory
  if (warp == kDgMmaWarp0) { if (PAIR) tmem_dealloc_pair(tmem_base, 512u); else tmem_dealloc(tmem_base, 512u); }
}

// ---------------------------------------------------------------------------------------------------------
// weight packing: fp32 reference layout -> bf16 K-major no-swizzle images [ntile][chunk][tap][k-core][cout][8]
// ---------------------------------------------------------------------------------------------------------
struct TcPackArgs { kc_desc d; TcGeom g; const float* w_base; const float* w_basis; uint4* out; };

// One thread = one (N tile, chunk, k-core, column) and ALL taps: for the B-spline / RBF / Chebyshev layouts the 8 x T
// source floats of a spline k-core (8 basis functions of one channel, T taps) and of a base k-core (8 channels, T taps)
// are one contiguous run, so every 32-byte sector that is fetched is fully used; the T output vectors go to the T tap
// images of the chunk (consecutive threads = consecutive columns -> contiguous 16-byte stores).
constexpr int kPackMaxT = 9;          // taps held in registers per pass (larger filters take several passes)

__global__ void __launch_bounds__(256) kc_pack_fwd_kernel(const __grid_constant__ TcPackArgs a) {
  const kc_desc& d = a.d;
  const TcGeom& g = a.g;
  const int T = d.kh * d.kw, nb = d.nb;
  const bool has_base = d.act != KC_ACT_NONE;
  const int nchunks = g.nsc + (has_base ? g.nbc : 0);
  const long long vec_per_ntile = g.wimg_bytes_per_ntile / 16;
  const long long full_chunk = (long long)T * kPL * g.ntile;
  const long long per_chunk = (long long)kPL * g.ntile;                        // threads per (N tile, chunk)
  const long long total = (long long)g.n_ntiles * nchunks * per_chunk;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
    const int nt = (int)(u / (nchunks * per_chunk));
    const long long ul = u - (long long)nt * nchunks * per_chunk;
    const int q = (int)(ul / per_chunk);
    const int rem = (int)(ul - (long long)q * per_chunk);
    const int kc = rem / g.ntile, nl = rem - kc * g.ntile;
    const int co = nt * g.ntile + nl;
    const bool spline = q < g.nsc;
    const int bq = q - g.nsc;
    const int ncols = (spline || bq != g.nbc - 1) ? kPL : g.last_base_cols;
    if (kc >= ncols) continue;
    // element e of the vector reads src[e * estride + t]; estride < 0 marks "zero"
    const float* src[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      src[e] = nullptr;
      if (co >= d.cout) continue;
      if (spline) {
        int c, jj;
        if (nb > 4) { c = q * 4 + kc; jj = e; } else { c = q * 8 + kc * 2 + (e >> 2); jj = e & 3; }
        if (c < d.cin && jj < nb) src[e] = a.w_basis + ((long long)co * d.cin * nb + kc_wbasis_index(d.basis, c, jj, d.cin, nb)) * T;
      } else {
        const int c = (bq * kPL + kc) * 8 + e;
        if (c < d.cin) src[e] = a.w_base + ((long long)co * d.cin + c) * T;
      }
    }
    uint4* dst = a.out + (long long)nt * vec_per_ntile + (spline ? (long long)q * full_chunk : (long long)g.nsc * full_chunk + (long long)bq * full_chunk) +
                 (long long)kc * g.ntile + nl;
    const long long tstride = (long long)ncols * g.ntile;
    for (int t0 = 0; t0 < T; t0 += kPackMaxT) {
      float f[kPackMaxT][8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int tt = 0; tt < kPackMaxT; ++tt) f[tt][e] = (src[e] != nullptr && t0 + tt < T) ? __ldg(src[e] + t0 + tt) : 0.0f;
#pragma unroll
      for (int tt = 0; tt < kPackMaxT; ++tt)
        if (t0 + tt < T)
          dst[(long long)(t0 + tt) * tstride] = make_uint4(pack_bf16(f[tt][0], f[tt][1]), pack_bf16(f[tt][2], f[tt][3]),
                                                            pack_bf16(f[tt][4], f[tt][5]), pack_bf16(f[tt][6], f[tt][7]));
    }
  }
}

// Forward image through a shared-memory transpose (channel-major weight layouts, T <= 9): the source of one (chunk, cout)
// is ONE contiguous run of nk x 8 x T floats (nk k-cores), read coalesced by a warp; the tile [32 couts][run] is then
// re-read column-wise (pitch run+1: conflict-free) to emit the 16-byte vectors with consecutive threads = consecutive
// couts.  grid = (cout tiles of 32, chunks, N tiles).
constexpr int kPackTileCo = 32;
constexpr int kPackRunMax = kPL * 8 * kPackMaxT;

__global__ void __launch_bounds__(256) kc_pack_fwd_tile_kernel(const __grid_constant__ TcPackArgs a) {
  __shared__ float tile[kPackTileCo][kPackRunMax + 1];
  const kc_desc& d = a.d;
  const TcGeom& g = a.g;
  const int T = d.kh * d.kw, nb = d.nb;
  const int nt = blockIdx.z, q = blockIdx.y, nl0 = blockIdx.x * kPackTileCo;
  const bool spline = q < g.nsc;
  const int bq = q - g.nsc;
  const int ncols = (spline || bq != g.nbc - 1) ? kPL : g.last_base_cols;
  const long long vec_per_ntile = g.wimg_bytes_per_ntile / 16;
  const long long full_chunk = (long long)T * kPL * g.ntile;
  // source run of one cout: channels [ch0, ch0 + nch) x (nb | 1) x T floats, zero beyond cin
  const int ch_per_core = spline ? (nb > 4 ? 1 : 2) : 8, per_ch = spline ? nb : 1;
  const int ch0 = spline ? q * kPL * ch_per_core : bq * kPL * 8;
  const int nch_live = max(0, min(ncols * ch_per_core, d.cin - ch0));
  const int run = nch_live * per_ch * T;                                  // floats actually present
  const float* base = spline ? a.w_basis : a.w_base;
  const long long co_stride = (long long)d.cin * per_ch * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cl = warp; cl < kPackTileCo; cl += 8) {
    const int co = nt * g.ntile + nl0 + cl;
    const bool ok = nl0 + cl < g.ntile && co < d.cout;
    const float* src = base + (long long)co * co_stride + (long long)ch0 * per_ch * T;
    for (int i = lane; i < run; i += 32) tile[cl][i] = ok ? __ldg(src + i) : 0.0f;
  }
  __syncthreads();
  uint4* dst = a.out + (long long)nt * vec_per_ntile + (spline ? (long long)q * full_chunk : (long long)g.nsc * full_chunk + (long long)bq * full_chunk);
  const int nvec = T * ncols * kPackTileCo;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    const int cl = v % kPackTileCo, rest = v / kPackTileCo, kc = rest % ncols, t = rest / ncols;
    if (nl0 + cl >= g.ntile) continue;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int idx;                                                            // float index inside the run, or -1
      if (!spline) idx = (kc * 8 + e < nch_live) ? (kc * 8 + e) * T + t : -1;
      else if (nb > 4) idx = (kc < nch_live && e < nb) ? (kc * nb + e) * T + t : -1;
      else idx = (kc * 2 + (e >> 2) < nch_live && (e & 3) < nb) ? ((kc * 2 + (e >> 2)) * nb + (e & 3)) * T + t : -1;
      f[e] = idx >= 0 ? tile[cl][idx] : 0.0f;
    }
    dst[((long long)t * ncols + kc) * g.ntile + nl0 + cl] =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// dgrad weights: [ntile = 16 input channels][chunk of 32 couts][flipped tap][k-core = 8 couts][n = (cl, j)][8]
// Same thread mapping (all taps per thread): the T floats of one (cout, channel, j) are contiguous and consecutive threads
// (consecutive j, then channel) read adjacent runs.
__global__ void __launch_bounds__(256) kc_pack_dgrad_kernel(const __grid_constant__ TcPackArgs a) {
  const kc_desc& d = a.d;
  const TcGeom& g = a.g;
  const int T = d.kh * d.kw, nb = d.nb;
  const bool has_base = d.act != KC_ACT_NONE;
  const int wb = nb + (has_base ? 1 : 0);
  const long long vec_per_ntile = g.wimg_bytes_per_ntile / 16;
  const long long per_chunk = (long long)kPL * g.ntile;
  const long long total = (long long)g.n_ntiles * g.nbc * per_chunk;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
    const int nt = (int)(u / (g.nbc * per_chunk));
    const long long ul = u - (long long)nt * g.nbc * per_chunk;
    const int bq = (int)(ul / per_chunk);
    const int rem = (int)(ul - (long long)bq * per_chunk);
    const int kc = rem / g.ntile, nl = rem - kc * g.ntile;
    const int ncols = (bq == g.nbc - 1) ? g.last_base_cols : kPL;
    if (kc >= ncols) continue;
    const int cl = nl / wb, c = nt * g.cpt + cl, jj = nl % wb;
    const float* src[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int co = (bq * kPL + kc) * 8 + e;
      src[e] = nullptr;
      if (co < d.cout && c < d.cin && cl < g.cpt)
        src[e] = jj < nb ? a.w_basis + ((long long)co * d.cin * nb + kc_wbasis_index(d.basis, c, jj, d.cin, nb)) * T
                         : a.w_base + ((long long)co * d.cin + c) * T;
    }
    // CTA pairs: the image of an N tile is two halves of nw = ntile / 2 rows, each laid out like a whole image of nw rows
    const int nw = g.pair ? g.ntile / 2 : g.ntile, half = nl / nw, nlh = nl - half * nw;
    uint4* dst = a.out + (long long)nt * vec_per_ntile + (long long)half * (vec_per_ntile / 2) + (long long)bq * ((long long)T * kPL * nw) +
                 (long long)kc * nw + nlh;
    const long long tstride = (long long)ncols * nw;
    for (int t0 = 0; t0 < T; t0 += kPackMaxT) {
      float f[kPackMaxT][8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int tt = 0; tt < kPackMaxT; ++tt) f[tt][e] = (src[e] != nullptr && t0 + tt < T) ? __ldg(src[e] + t0 + tt) : 0.0f;
#pragma unroll
      for (int tt = 0; tt < kPackMaxT; ++tt)
        if (t0 + tt < T)     // image tap index is the flipped filter tap
          dst[(long long)(T - 1 - (t0 + tt)) * tstride] = make_uint4(pack_bf16(f[tt][0], f[tt][1]), pack_bf16(f[tt][2], f[tt][3]),
                                                                      pack_bf16(f[tt][4], f[tt][5]), pack_bf16(f[tt][6], f[tt][7]));
    }
  }
}

// dz (fp32 NCHW) -> flat bf16, plane-major [cq/8][L][8]: plane p holds output channels 8p..8p+7 of flat position
// q = n*IMG + y*P + x as one 16-byte vector; zero where (y, x) is not an output pixel.  One thread per (q, plane): its
// eight channel reads are each coalesced across the warp (consecutive q), the 16-byte stores are contiguous.
__global__ void __launch_bounds__(256) kc_dz_flat_kernel(const __grid_constant__ kc_desc d, int P, int IMG, long long L, int cq,
                                                         const float* __restrict__ dz, unsigned char* __restrict__ out) {
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  const int pl = blockIdx.y;
  if (q >= L) return;
  const int HoWo = d.ho * d.wo;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = 0.0f;
  const int n = (int)(q / IMG);
  const int rem = (int)(q - (long long)n * IMG);
  const int y = rem / P, x = rem - y * P;
  int yo = y, xo = x;
  bool on_grid = true;
  if (d.stride_h != 1 || d.stride_w != 1) {       // strided layer: dz lives on every stride-th position of the stride-1 grid
    yo = y / d.stride_h; xo = x / d.stride_w;
    on_grid = yo * d.stride_h == y && xo * d.stride_w == x;
  }
  if (on_grid && yo < d.ho && xo < d.wo) {
    const float* src = dz + (long long)n * d.z_batch_stride + yo * d.wo + xo;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (pl * 8 + e < d.cout) f[e] = __ldg(src + (long long)(pl * 8 + e) * HoWo);
  }
  *reinterpret_cast<uint4*>(out + ((long long)pl * L + q) * 16) =
      make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------------------
// host-side geometry
// ---------------------------------------------------------------------------------------------------------
size_t tc_fixed_smem() { return (size_t)kNumBars * 8 + 16 + sizeof(KcBasisCtx) + 128; }

// Common tail of the forward / dgrad geometry: choose nsub, ring depths and shared-memory carve-up.
// kcores = total number of 16-byte k-cores per tap in the weight image of one N tile.
int tc_fit(const kc_desc* d, TcGeom* g, int T, int kcores, int pair) {
  const size_t btap = (size_t)kPL * (pair ? g->ntile / 2 : g->ntile) * 16;        // per CTA
  // nsub: 128-row sub-tiles per CTA, bounded by TMEM (512 columns), the producer row mapping (1024 rows) and shared memory.
  // Among the nsub that fit, take the one with the smallest estimated kernel time:
  //   waves(nsub) * [ max(nsub * MMA cycles of the K loop, weight-image bytes / ~32 B per cycle from L2) + fixed + nsub * epilogue ]
  // (a CTA streams the whole weight image of its N tile whatever nsub is, so nsub = 1 is L2-bandwidth bound; large nsub
  // leaves SMs idle when the batch is small).
  bool found = false;
  double best_cost = 0.0;
  const int nchunks = (kcores + kPL - 1) / kPL;
  const int sms = kc_sm_count();
  for (int nsub = 4; nsub >= 1; --nsub) {
    if (nsub * g->ntile > 512) continue;
    const int mcta = nsub * kTileM;
    const int seglen = round_up(mcta + d->kw - 1, 8);
    const int SS = g->P < seglen ? g->P : seglen;
    const int nrows = (d->kh - 1) * SS + seglen;
    if (nrows > kRB * kRowThreads) continue;
    const long long mtiles = (g->L + mcta - 1) / mcta;
    const long long ctas = (pair ? (mtiles + 1) / 2 * 2 : mtiles) * g->n_ntiles;
    // CTAs do not run in lockstep, so wave quantisation only matters while the grid is a wave or two
    const double waves = ctas < 2 * sms ? (double)((ctas + sms - 1) / sms) : (double)ctas / (double)sms;
    const double mma_cyc = (double)nchunks * T * 2.0 * (g->ntile / 2 > 48 ? g->ntile / 2 : 48);     // per sub-tile
    const double b_cyc = (double)nchunks * T * (double)btap / 32.0;
    const double cta_cyc = (nsub * mma_cyc > b_cyc ? nsub * mma_cyc : b_cyc) + 8000.0 + nsub * 3000.0;
    const double cost = waves * cta_cyc;
    if (found && cost >= best_cost) continue;
    const int plane_bytes = nrows * 16 + 16;       // +16 B: consecutive planes start 4 banks apart
    // preference: whole-filter-row stages with a 3-deep A ring, then with a 2-deep A ring, then single-tap stages
    // (a ring stage that carries kw taps pays the issuing warp's per-stage cost - barrier wait, fence, commits, several
    // hundred cycles - once per filter row)
    // coarsest first: all taps of a chunk per stage (>= 2 stages), a filter row (>= 3), a single tap (>= 6)
    // (a filter row with a 4-deep ring is preferred over a deeper A ring with only 3 stages)
    const int cand_na[8] = {3, 2, 3, 2, 3, 2, 3, 2}, cand_tps[8] = {T, T, d->kw, d->kw, d->kw, d->kw, 1, 1};
    const int cand_min[8] = {2, 2, 4, 4, 3, 3, 6, 6};
    for (int ci = 0; ci < 8; ++ci) {
      const int na = cand_na[ci], tps = cand_tps[ci];
      if (ci >= 2 && ci < 4 && tps == T) continue;                        // kw == T: covered by the first pair
      if (ci >= 6 && d->kw == 1) continue;                                // kw == 1: covered by the filter-row candidates
      const size_t fixed = tc_fixed_smem() + (size_t)na * kPL * plane_bytes;
      const size_t bstage = btap * tps;
      if (fixed + cand_min[ci] * bstage > kSmemLimit) continue;
      int bst = (int)((kSmemLimit - fixed) / bstage);
      if (bst > kMaxBStages) bst = kMaxBStages;
      g->nsub = nsub; g->mcta = mcta; g->SS = SS; g->nrows = nrows; g->plane_bytes = plane_bytes; g->tps = tps;
      g->na = na; g->bstages = bst; g->mtiles = mtiles; g->smem_bytes = fixed + bst * bstage; g->pair = pair;
      found = true;
      best_cost = cost;
      break;
    }
  }
  if (!found) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path: tile does not fit shared memory");
  g->tmem_cols = 32;
  while (g->tmem_cols < g->nsub * g->ntile) g->tmem_cols *= 2;
  g->wimg_bytes_per_ntile = (long long)T * g->ntile * 16 * kcores;
  g->fast_cubic = kc_knots_uniform_cubic(d, &g->t0, &g->inv_h) ? 1 : 0;
  return KC_OK;
}

int tc_common_checks(const kc_desc* d) {
  // A strided convolution is the stride-1 convolution sampled at every stride-th position: the kernels run on the stride-1
  // position grid, the forward epilogue stores only the sampled outputs and dz_flat scatters dz onto that grid (zeros in
  // between).  Costs stride_h * stride_w times the MMA work - meant for the strided stem / downsampling layers of a model.
  if (d->dil_h != 1 || d->dil_w != 1) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs dilation 1");
  if (d->stride_h < 1 || d->stride_w < 1 || d->stride_h > 4 || d->stride_w > 4) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs stride <= 4");
  if (d->nb < 1 || d->nb > 8) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs basis width <= 8 (got %d)", d->nb);
  if (d->pad_h > d->kh - 1 || d->pad_w > d->kw - 1) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs padding < kernel size");
  if (d->kw > 8 || d->kh > 8) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs kernel size <= 8");
  if ((long long)d->n * d->x_batch_stride >= (1LL << 31)) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core path needs < 2^31 input elements");
  return KC_OK;
}

// dgrad GEMM: rows = input positions, K = (tap, cout), N = (channel, basis j | base) in tiles of 16 channels.
int tc_dgrad_geometry(const kc_desc* d, TcGeom* g) {
  int rc = tc_common_checks(d);
  if (rc != KC_OK) return rc;
  memset(g, 0, sizeof(*g));
  const bool has_base = d->act != KC_ACT_NONE;
  const int T = d->kh * d->kw, wb = d->nb + (has_base ? 1 : 0);
  const int cq = round_up(d->cout, 16), planes = cq / 8;
  g->Cp = cq; g->cps = 0; g->nsc = 0; g->ngroups = planes;
  g->nbc = (planes + kPL - 1) / kPL;
  g->last_base_cols = planes - (g->nbc - 1) * kPL;
  g->ph = d->kh - 1 - d->pad_h; g->pw = d->kw - 1 - d->pad_w;      // transposed convolution: flipped taps
  g->P = d->w + d->pad_w;
  g->IMG = (d->h + d->pad_h) * g->P;
  g->L = (long long)d->n * g->IMG;
  if (g->L >= (1LL << 31)) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core dgrad needs < 2^31 flat positions");
  // Persistent kernel (closed-form cubic basis): N tile = cpt channels x wb columns padded to 128, two sub-tiles per
  // accumulator set, two sets in TMEM (4 x 128 = 512 columns).
  float t0 = 0.0f, inv_h = 0.0f;
  const bool cubic = kc_knots_uniform_cubic(d, &t0, &inv_h);
  if ((cubic || d->basis == KC_BASIS_RBF || d->basis == KC_BASIS_CHEBY || d->basis == KC_BASIS_GRAM) && d->kh * d->kw <= 64) {
    // channels per N tile: cpt * wb <= 128 columns, <= 4 (8) channels per epilogue warp, and the 8-column TMEM read of the
    // last channel must stay inside the tile ((cpt - 1) * wb + 8 <= 128)
    int cpt = 128 / wb;
    if (cpt > 120 / wb + 1) cpt = 120 / wb + 1;
    if (cpt > (cubic ? 16 : 32)) cpt = cubic ? 16 : 32;
    const int ntile = 128;
    const int mcta = 2 * kTileM;
    const int seglen = round_up(mcta + d->kw - 1, 8);
    const int SS = g->P < seglen ? g->P : seglen;
    const int nrows = (d->kh - 1) * SS + seglen;
    const int plane_bytes = nrows * 16 + 16;
    static const int pair_enabled = []() { const char* e = getenv("KANCONV_DGRAD_PAIR"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
    // pairs pay off where the kernel is bound by the MMA rate (>= 4 K chunks per tile, i.e. cout >= 128); with fewer chunks
    // the epilogue sets the pace and coupling two CTAs only adds hand-over latency (64 -> 64 @224: 7 % slower)
    const int pair = pair_enabled && (long long)((g->L + mcta - 1) / mcta) >= 2 && planes >= 4 * kPL ? 1 : 0;
    const size_t btap = (size_t)kPL * (pair ? ntile / 2 : ntile) * 16;      // per CTA
    const size_t fixed0 = (size_t)kDgBars * 8 + 16 + sizeof(KcBasisCtx) + 128;
    if (d->kh * kPL <= 32) {
      // weight ring granularity, coarsest first: a whole 32-cout chunk (all taps) per stage, a filter row, a single tap -
      // the barrier round of a stage costs the issuing warps a few hundred cycles, so fewer and larger stages win as long
      // as the ring stays >= 2-3 deep
      const int cand_na[6] = {3, 2, 3, 2, 3, 2}, cand_tps[6] = {T, T, d->kw, d->kw, 1, 1}, cand_min[6] = {2, 2, 3, 3, 6, 6};
      for (int ci = 0; ci < 6; ++ci) {
        const int na = cand_na[ci], tps = cand_tps[ci];
        if (ci >= 2 && tps == cand_tps[ci - 2]) continue;                 // kw == T or kw == 1: same as a coarser candidate
        const size_t fixed = fixed0 + (size_t)na * kPL * plane_bytes;
        const size_t bstage = btap * tps;
        if (fixed + cand_min[ci] * bstage > kSmemLimit) continue;
        int bst = (int)((kSmemLimit - fixed) / bstage);
        if (bst > kMaxBStages) bst = kMaxBStages;
        g->persistent = 1; g->pair = pair; g->cpt = cpt; g->ntile = ntile; g->n_ntiles = (d->cin + cpt - 1) / cpt;
        g->nsub = 2; g->mcta = mcta; g->SS = SS; g->nrows = nrows; g->plane_bytes = plane_bytes; g->tps = tps; g->na = na;
        g->bstages = bst; g->mtiles = (g->L + mcta - 1) / mcta; g->smem_bytes = fixed + bst * bstage; g->tmem_cols = 512;
        g->wimg_bytes_per_ntile = (long long)T * ntile * 16 * planes;
        g->fast_cubic = cubic ? 1 : 0; g->t0 = t0; g->inv_h = inv_h;
        return KC_OK;
      }
    }
  }
  g->cpt = 16;
  g->ntile = 16 * wb;
  g->n_ntiles = (d->cin + 15) / 16;
  return tc_fit(d, g, T, planes, 0);
}

int tc_forward_geometry(const kc_desc* d, TcGeom* g) {
  int rc0 = tc_common_checks(d);
  if (rc0 != KC_OK) return rc0;
  memset(g, 0, sizeof(*g));
  const bool has_base = d->act != KC_ACT_NONE;
  const int T = d->kh * d->kw;
  g->cps = (d->nb > 4) ? 4 : 8;            // basis width is zero-padded to 8 or 4
  g->Cp = round_up(d->cin, 8);
  g->nsc = g->Cp / g->cps;
  g->ngroups = g->Cp / 8;
  g->nbc = has_base ? (g->ngroups + kPL - 1) / kPL : 0;
  g->last_base_cols = has_base ? round_up(g->ngroups - (g->nbc - 1) * kPL, 2) : 0;
  g->ph = d->pad_h; g->pw = d->pad_w;
  g->P = d->w + d->pad_w;
  g->IMG = (d->h + d->pad_h) * g->P;
  g->L = (long long)d->n * g->IMG;
  int want_tiles = (d->cout + 255) / 256;
  g->ntile = round_up((d->cout + want_tiles - 1) / want_tiles, 16);
  g->n_ntiles = (d->cout + g->ntile - 1) / g->ntile;
  const int kcores = (has_base ? (g->nbc - 1) * kPL + g->last_base_cols : 0) + g->nsc * kPL;
  // Pointwise layers, and stem layers with <= 8 input channels (an RGB image): persistent GEMM over the basis rows of a pre-pass
  // (see kc_dgrad_persistent_kernel, FAM 2).  Both have so little MMA work per position that the fused kernel's CTA (evaluate,
  // multiply, store, one after the other) is latency-bound; the persistent kernel overlaps the epilogue of a tile with the
  // MMAs of the next one, and the rows (18 B per position and channel) are what the weight gradient reads anyway.
  long long Lw = 0;
  int splanes = 0, bplanes = 0;
  static const int stem_enabled = []() { const char* e = getenv("KANCONV_STEM_FROM_PHI"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
  const bool stem = stem_enabled && T > 1 && d->cin <= 8 && d->kh * kPL <= 32 && d->stride_h == 1 && d->stride_w == 1;
  if ((T == 1 || stem) && kc_tc_wgrad_phi_layout(d, &Lw, &splanes, &bplanes) == KC_OK && Lw == g->L) {
    const int ntile = d->cout >= 128 ? 128 : round_up(d->cout, 16);
    const int mcta = 2 * kTileM;
    const int seglen = T == 1 ? mcta : round_up(mcta + d->kw - 1, 8), SS = g->P < seglen ? g->P : seglen;
    const int nrows = (d->kh - 1) * SS + seglen;
    const int plane_bytes = nrows * 16 + 16;
    const size_t btap = (size_t)kPL * ntile * 16 * (T == 1 ? 1 : T);      // stem: all taps of a chunk per weight stage
    const size_t fixed = (size_t)kDgBars * 8 + 16 + sizeof(KcBasisCtx) + 128 + (size_t)3 * kPL * plane_bytes;
    int bst = fixed < kSmemLimit ? (int)((kSmemLimit - fixed) / btap) : 0;
    if (bst > kMaxBStages) bst = kMaxBStages;
    if (bst >= (T == 1 ? 4 : 2)) {
      g->from_phi = 1; g->persistent = 1; g->ntile = ntile; g->n_ntiles = (d->cout + ntile - 1) / ntile;
      g->nsub = 2; g->mcta = mcta; g->SS = SS; g->nrows = nrows; g->plane_bytes = plane_bytes; g->tps = T == 1 ? 1 : T; g->na = 3;
      g->bstages = bst; g->mtiles = (g->L + mcta - 1) / mcta; g->smem_bytes = fixed + bst * btap; g->tmem_cols = 512;
      g->wimg_bytes_per_ntile = (long long)T * ntile * 16 * kcores;
      g->fast_cubic = kc_knots_uniform_cubic(d, &g->t0, &g->inv_h) ? 1 : 0;
      return KC_OK;
    }
  }
  static const int fwd_pair_enabled = []() { const char* e = getenv("KANCONV_FWD_PAIR"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
  // (not for N tiles below 128: those layers are bound by the basis producers, and coupling two CTAs' producers costs 10 %)
  return tc_fit(d, g, T, kcores, fwd_pair_enabled && g->ntile % 16 == 0 && g->ntile >= 128 && g->L > 2 * kTileM ? 1 : 0);
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
template <int FAM, bool PAIR>
cudaError_t launch_dgrad(const TcFwdArgs& a, unsigned nctas, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kc_dgrad_persistent_kernel<FAM, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.g.smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nctas, 1, 1);
  cfg.blockDim = dim3(kDgThreads, 1, 1);
  cfg.dynamicSmemBytes = a.g.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2u : 1u;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kc_dgrad_persistent_kernel<FAM, PAIR>, a);
}

extern "C" int kc_tc_supported(const kc_desc* d) {
  if (kc_validate_desc(d) != KC_OK) return 0;
  TcGeom g;
  return tc_forward_geometry(d, &g) == KC_OK ? 1 : 0;
}

extern "C" int kc_tc_fwd_needs_phi(const kc_desc* d) {
  if (kc_validate_desc(d) != KC_OK) return 0;
  TcGeom g;
  return (tc_forward_geometry(d, &g) == KC_OK && g.from_phi) ? 1 : 0;
}

extern "C" size_t kc_tc_bytes(const kc_desc* d, int which) {
  if (kc_validate_desc(d) != KC_OK) return 0;
  TcGeom g;
  if (tc_forward_geometry(d, &g) != KC_OK) return 0;
  if (which == 0) return (size_t)g.wimg_bytes_per_ntile * g.n_ntiles;
  TcGeom gd;
  if (tc_dgrad_geometry(d, &gd) != KC_OK) return 0;
  if (which == 1) return (size_t)gd.wimg_bytes_per_ntile * gd.n_ntiles;
  if (which == 2) return (size_t)gd.L * gd.Cp * 2;
  if (which == 3) return kc_tc_wgrad_ws_bytes(d, 0);
  if (which == 4) return kc_tc_wgrad_ws_bytes(d, 1);
  if (which == 5) return kc_tc_wgrad_ws_bytes(d, 2);
  return 0;
}

// Layout of the bf16 flat dz buffer (kc_tc_dz_flat / kc_norm_bwd_dz_flat): row pitch, positions per image, total positions,
// channels per position (cout rounded up to 16).
int kc_tc_flat_layout(const kc_desc* d, int* P, int* IMG, long long* L, int* cq) {
  TcGeom g;
  int rc = tc_dgrad_geometry(d, &g);
  if (rc != KC_OK) return rc;
  *P = g.P; *IMG = g.IMG; *L = g.L; *cq = g.Cp;
  return KC_OK;
}

// CTAs of the persistent dgrad launch: one per SM, or one per tile when there are fewer tiles (pairs: an even number, one
// pair per two neighbouring position tiles).
unsigned tc_dgrad_grid(const TcGeom& g) {
  const long long sms = kc_sm_count();
  if (g.pair) {
    const long long npairs = ((g.mtiles + 1) / 2) * g.n_ntiles;
    return (unsigned)(2 * (npairs < sms / 2 ? npairs : sms / 2));
  }
  const long long ntiles = g.mtiles * g.n_ntiles;
  return (unsigned)(ntiles < sms ? ntiles : sms);
}

// GRAM only: floats of the `dbeta` buffer (KC_MAX_BASIS results + one partial row per thread block / epilogue warp).
extern "C" size_t kc_dbeta_floats(const kc_desc* d, int tc) {
  if (kc_validate_desc(d) != KC_OK || d->basis != KC_BASIS_GRAM) return 0;
  long long rows;
  if (tc) {
    TcGeom g;
    if (tc_dgrad_geometry(d, &g) != KC_OK) return 0;
    if (g.persistent) {           // one row per (CTA, epilogue warp) of the persistent grid
      rows = (long long)tc_dgrad_grid(g) * kDgEpiWarps;
    } else {
      rows = g.mtiles * g.n_ntiles * 16;
    }
  } else {
    rows = kc_simt_dgrad_blocks(d);
  }
  return (size_t)(1 + rows) * KC_MAX_BASIS;
}

extern "C" int kc_tc_pack_weights(const kc_desc* d, const float* w_base, const float* w_basis, void* packed_fwd,
                                  void* packed_dgrad, void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  TcGeom g;
  rc = tc_forward_geometry(d, &g);
  if (rc != KC_OK) return rc;
  if (!w_basis || (!packed_fwd && !packed_dgrad)) KC_FAIL(KC_ERR_INVALID, "kc_tc_pack_weights: null pointer");
  if (d->act != KC_ACT_NONE && !w_base) KC_FAIL(KC_ERR_INVALID, "kc_tc_pack_weights: base branch needs w_base");
  TcPackArgs a;
  a.d = *d; a.g = g; a.w_base = w_base; a.w_basis = w_basis; a.out = (uint4*)packed_fwd;
  const int T = d->kh * d->kw;
  long long total = g.wimg_bytes_per_ntile / 16 * g.n_ntiles / T + (long long)kPL * g.ntile * g.n_ntiles;   // one thread per (vector, all taps)
  int blocks = (int)((total + 255) / 256);
  const int max_blocks = kc_sm_count() * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  if (packed_fwd != nullptr) {
    const int nchunks = g.nsc + (d->act != KC_ACT_NONE ? g.nbc : 0);
    if (!kc_degree_major(d->basis) && T <= kPackMaxT && nchunks <= 65535 && g.n_ntiles <= 65535) {
      dim3 pgrid((unsigned)((g.ntile + kPackTileCo - 1) / kPackTileCo), (unsigned)nchunks, (unsigned)g.n_ntiles);
      kc_pack_fwd_tile_kernel<<<pgrid, 256, 0, (cudaStream_t)stream>>>(a);
    } else {
      kc_pack_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    }
    KC_LAUNCH_CHECK("kc_pack_fwd_kernel");
  }
  if (packed_dgrad != nullptr) {
    TcGeom gd;
    rc = tc_dgrad_geometry(d, &gd);
    if (rc != KC_OK) return rc;
    a.g = gd; a.out = (uint4*)packed_dgrad;
    total = gd.wimg_bytes_per_ntile / 16 * gd.n_ntiles / T + (long long)kPL * gd.ntile * gd.n_ntiles;
    blocks = (int)((total + 255) / 256);
    if (blocks > max_blocks) blocks = max_blocks;
    kc_pack_dgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    KC_LAUNCH_CHECK("kc_pack_dgrad_kernel");
  }
  return KC_OK;
}

extern "C" int kc_tc_dz_flat(const kc_desc* d, const float* dz, void* dz_flat, void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  TcGeom g;
  rc = tc_dgrad_geometry(d, &g);
  if (rc != KC_OK) return rc;
  if (!dz || !dz_flat) KC_FAIL(KC_ERR_INVALID, "kc_tc_dz_flat: null pointer");
  dim3 grid((unsigned)((g.L + 255) / 256), (unsigned)(g.Cp / 8));
  kc_dz_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*d, g.P, g.IMG, g.L, g.Cp, dz, (unsigned char*)dz_flat);
  KC_LAUNCH_CHECK("kc_dz_flat_kernel");
  return KC_OK;
}

extern "C" int kc_conv_fwd_tc(const kc_desc* d, const float* x_base, const float* x_basis, const void* packed_fwd,
                              const float* beta, float* z, void* phi_out, void* stream) {
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
  memset(&a, 0, sizeof(a));
  a.d = *d; a.g = g; a.x_base = x_base; a.x_basis = x_basis; a.wp = (const unsigned char*)packed_fwd; a.beta = beta; a.z = z;
  if (g.from_phi) {
    // pointwise layer: basis rows by the pre-pass, then the persistent GEMM streams them back through the TMA engine
    if (phi_out == nullptr) KC_FAIL(KC_ERR_INVALID, "kc_conv_fwd_tc: 1x1 layers need the phi buffer (kc_tc_fwd_needs_phi, kc_tc_bytes(d, 4))");
    long long Lw = 0;
    int splanes = 0, bplanes = 0;
    rc = kc_tc_wgrad_phi_layout(d, &Lw, &splanes, &bplanes);
    if (rc != KC_OK) return rc;
    rc = kc_tc_phi_prepass(d, x_base, x_basis, beta, phi_out, stream);
    if (rc != KC_OK) return rc;
    a.dzf = (const unsigned char*)phi_out;
    a.phi_base_plane0 = splanes;
    const int sms = kc_sm_count();
    const long long ntiles = g.mtiles * g.n_ntiles;
    KC_CUDA_CHECK(cudaFuncSetAttribute(kc_dgrad_persistent_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    kc_dgrad_persistent_kernel<2, false><<<(unsigned)(ntiles < sms ? ntiles : sms), kDgThreads, g.smem_bytes, (cudaStream_t)stream>>>(a);
    KC_LAUNCH_CHECK("kc_tc_kernel<fwd>");
    return KC_OK;
  }
  if (phi_out != nullptr) {
    long long Lw = 0;
    int splanes = 0, bplanes = 0;
    rc = kc_tc_wgrad_phi_layout(d, &Lw, &splanes, &bplanes);
    if (rc != KC_OK) return rc;
    if (Lw != g.L) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_conv_fwd_tc: forward / wgrad flat lengths differ");
    a.phi_out = (unsigned char*)phi_out;
    a.phi_base_plane0 = splanes;
    // the planes behind the channels this kernel expands (padding of the weight gradient's M = 128 tiles) stay unwritten: the
    // weight-gradient kernel zero-fills them in shared memory instead of reading them (WgGeom::vs_planes / vb_planes)
  }
  if (g.pair) {
    KC_CUDA_CHECK(cudaFuncSetAttribute(kc_tc_kernel<kModeFwd, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((g.mtiles + 1) / 2 * 2), (unsigned)g.n_ntiles, 1);      // an odd tile count gets an idle partner
    cfg.blockDim = dim3(kTcThreads, 1, 1);
    cfg.dynamicSmemBytes = g.smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    KC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kc_tc_kernel<kModeFwd, true>, a));
  } else {
    KC_CUDA_CHECK(cudaFuncSetAttribute(kc_tc_kernel<kModeFwd, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    dim3 grid((unsigned)g.mtiles, (unsigned)g.n_ntiles);
    kc_tc_kernel<kModeFwd, false><<<grid, kTcThreads, g.smem_bytes, (cudaStream_t)stream>>>(a);
  }
  KC_LAUNCH_CHECK("kc_tc_kernel<fwd>");
  return KC_OK;
}

extern "C" int kc_conv_dgrad_tc(const kc_desc* d, const float* dz, const float* x_base, const float* x_basis,
                                const void* packed_dgrad, const float* beta, float* dx_base, float* dx_basis,
                                float* dbeta, void* workspace, void* stream) {
  (void)dz;
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  TcGeom g;
  rc = tc_dgrad_geometry(d, &g);
  if (rc != KC_OK) return rc;
  if (!x_basis || !packed_dgrad || !dx_basis || !workspace) KC_FAIL(KC_ERR_INVALID, "kc_conv_dgrad_tc: null pointer");
  if (d->act != KC_ACT_NONE && !x_base) KC_FAIL(KC_ERR_INVALID, "kc_conv_dgrad_tc: base branch needs x_base");
  if (g.mtiles > 0x7fffffffLL) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_conv_dgrad_tc: too many tiles");
  TcFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.d = *d; a.g = g; a.x_base = x_base; a.x_basis = x_basis; a.wp = (const unsigned char*)packed_dgrad; a.beta = beta;
  a.dzf = (const unsigned char*)workspace; a.cq = g.Cp; a.dx_base = dx_base; a.dx_basis = dx_basis;
  a.dbeta = (d->basis == KC_BASIS_GRAM) ? dbeta : nullptr;
  if (g.persistent) {
    const unsigned nctas = tc_dgrad_grid(g);
    const int fam = g.fast_cubic ? 1 : d->basis == KC_BASIS_GRAM ? 3 : 0;
    cudaError_t e;
    if (g.pair) e = fam == 1 ? launch_dgrad<1, true>(a, nctas, (cudaStream_t)stream) : fam == 3 ? launch_dgrad<3, true>(a, nctas, (cudaStream_t)stream)
                                                                                         : launch_dgrad<0, true>(a, nctas, (cudaStream_t)stream);
    else e = fam == 1 ? launch_dgrad<1, false>(a, nctas, (cudaStream_t)stream) : fam == 3 ? launch_dgrad<3, false>(a, nctas, (cudaStream_t)stream)
                                                                                : launch_dgrad<0, false>(a, nctas, (cudaStream_t)stream);
    KC_CUDA_CHECK(e);
    KC_LAUNCH_CHECK("kc_dgrad_persistent_kernel");
    if (fam == 3 && a.dbeta != nullptr) return kc_dbeta_reduce(a.dbeta, (long long)nctas * kDgEpiWarps, stream);
    return KC_OK;
  }
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_tc_kernel<kModeDgrad, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  dim3 grid((unsigned)g.mtiles, (unsigned)g.n_ntiles);
  kc_tc_kernel<kModeDgrad, false><<<grid, kTcThreads, g.smem_bytes, (cudaStream_t)stream>>>(a);
  KC_LAUNCH_CHECK("kc_tc_kernel<dgrad>");
  if (a.dbeta != nullptr) return kc_dbeta_reduce(a.dbeta, (long long)grid.x * grid.y * 16, stream);
  return KC_OK;
}

// Tile geometry the kernels choose for a shape (introspection for integrators and the CPU tests; no GPU needed).
// which 0 = forward, 1 = dgrad; out = {nsub, ntile, n_ntiles, na, tps, bstages, mtiles, smem bytes}
extern "C" int kc_tc_geometry(const kc_desc* d, int which, long long* out) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  if (!out) KC_FAIL(KC_ERR_INVALID, "kc_tc_geometry: null pointer");
  TcGeom g;
  rc = which == 0 ? tc_forward_geometry(d, &g) : tc_dgrad_geometry(d, &g);
  if (rc != KC_OK) return rc;
  out[0] = g.nsub; out[1] = g.ntile; out[2] = g.n_ntiles; out[3] = g.na; out[4] = g.tps; out[5] = g.bstages; out[6] = g.mtiles;
  out[7] = (long long)g.smem_bytes;
  return KC_OK;
}

#ifdef KANCONV_DEBUG
// Debug build only (not part of include/kanconv.h): enable / disable the timeline trace of the kernels in this file.
extern "C" int kc_debug_trace(void* device_buffer) {
  long long* p = (long long*)device_buffer;
  KC_CUDA_CHECK(cudaMemcpyToSymbol(g_trace, &p, sizeof(p)));
  return KC_OK;
}
#endif

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
