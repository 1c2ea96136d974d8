// kc_tc_wgrad.cu - BF16 tensor-core weight gradient of the KAN convolution (tcgen05 / TMEM), sm_100a.
//
//   dW[cout][tap][c, j] = sum over flat output positions m of  dz[m][cout] * Phi_j(x[m + off(tap)][c])      (j = nb: base act)
//
// GEMM view per CTA: D[M = 128 expanded input rows (16 channels x 8 basis, or 128 base channels)][N = cout tile] for the
// kw taps of ONE filter row r, reduced over a range of 128-position blocks (split-K).  Both operands are "MN-major"
// no-swizzle UMMA layouts whose K dimension (positions) runs along 16-byte rows:
//   A = Phi  planes [channel (8 j) | 8-channel group][row = position + tap shift][8 x bf16]   - 16-byte copies of phi, the
//                                                                 rows the forward kernel saved (or kc_phi_flat_kernel)
//   B = dz   planes [8 couts][row = position][8 x bf16]                                        - 16-byte copies of the
//                                                                                               flat bf16 dz buffer
// so the three taps of the filter row are again shifted VIEWS (start row + s) of the same Phi stage.  Each tap has its own TMEM accumulator (kw * N <= 512 columns).  Partial sums go to a
// workspace [split][unit][s][128][N] and kc_wgrad_tc_reduce_kernel adds the splits in fixed order (deterministic) while
// scattering into the reference's parameter layouts.
#include <stdlib.h>
#include <string.h>

#include "kc_common.cuh"
#include "kc_umma.cuh"
#include "kc_tc_basis.cuh"

namespace {

using namespace kc;

constexpr int kThreadsW = 576;       // warps 0-15 producers / epilogue, 16-17 MMA issuers (highest warp ids)
constexpr int kProdW = 512;
constexpr int kMmaWarpW = 16;
constexpr int kMaxStagesW = 6;
constexpr int kNumBarsW = 3 * kMaxStagesW + 1;
constexpr size_t kSmemLimitW = 227 * 1024;

struct WgGeom {
  int P, IMG;
  long long L;
  int cq;                       // channels per row of the flat dz buffer
  int cps;                      // input channels per spline chunk (16 for nb = 8, 32 for nb = 4)
  int nsc, nbc, nchunks;        // spline / base chunks (M = 128 rows each)
  int ntile, n_ct;              // cout tile (MMA N), number of cout tiles
  int units;                    // n_ct * nchunks * kh
  int pair;                     // 1 = CTA pairs (tcgen05 cta_group::2): two units of a cout tile share the dz stage, each CTA loads half
  int units_ct, upc;            // units per cout tile (nchunks * kh), the same rounded up to an even number (grid.x = n_ct * upc)
  int bplanes_cta;              // dz planes a CTA loads per stage (bplanes, or bplanes / 2 in a pair)
  int merged;                   // stem layers (kh * (cin + base planes) <= 16): ONE unit per cout tile whose 16 Phi planes are
                                // (filter row r, channel | base plane) - all filter rows and both branches in one M = 128 tile
  int mnv;                      // merged: planes per filter row = cin + base planes
  int vs_planes, vb_planes;     // spline / base planes of phi that hold data (channels < cin rounded up to 8); the planes behind
                                // them exist only as padding of the M = 128 tiles: never written, never read (zero-filled in smem)
  int arows, aplane_bytes;      // Phi rows per stage (kKS + kw - 1, padded), plane pitch
  int bplanes, bplane_bytes;
  int stages, stage_bytes, a_bytes;
  int ks;                       // positions per pipeline stage (64 or 128)
  int prefetch;                 // cp.async stages in flight per producer thread
  int tmem_cols;
  long long nblk;               // 64-position blocks in total
  int nsplit;
  long long blk_per_split;
  size_t smem_bytes;
  size_t ws_bytes, phi_bytes;
  int fast_cubic;
  float t0, inv_h;
};

struct WgArgs {
  kc_desc d;
  WgGeom g;
  const float* x_base;
  const float* x_basis;
  const unsigned char* dzf;     // bf16 plane-major [cq/8][L][8]
  const unsigned char* phi;     // bf16 plane-major [nchunks*16][L][8]  (kc_phi_flat_kernel)
  const float* beta;
  float* ws;
};

__host__ __device__ inline int round_up_w(int a, int b) { return (a + b - 1) / b * b; }

KC_TRACE_DECL(g_trace_w)

__device__ __noinline__ uint4 basis8w_generic(const KcBasisCtx& B, float x) {
  float phi[KC_MAX_BASIS];
#pragma unroll
  for (int j = 0; j < 8; ++j) phi[j] = 0.0f;
  tc_eval_basis(B, x, phi, nullptr);          // same evaluator as the forward producers: the rows must match what they store
  return make_uint4(pack_bf16(phi[0], phi[1]), pack_bf16(phi[2], phi[3]), pack_bf16(phi[4], phi[5]), pack_bf16(phi[6], phi[7]));
}
__device__ __noinline__ uint2 basis4w(const KcBasisCtx& B, float x) {
  float phi[KC_MAX_BASIS];
#pragma unroll
  for (int j = 0; j < 4; ++j) phi[j] = 0.0f;
  tc_eval_basis(B, x, phi, nullptr);
  return make_uint2(pack_bf16(phi[0], phi[1]), pack_bf16(phi[2], phi[3]));
}

// Pre-pass: evaluate the A operand ONCE per (position, channel) into a transient bf16 plane-major buffer
//   phi[chunk][plane 0..15][L][8]     spline chunk: plane = channel (nb = 8) or channel pair (nb = 4), 8 basis values
//                                     base chunk  : plane = 8 consecutive channels, their base activations
// (zero at padding positions).  Reads of x are coalesced along the flat position, writes are contiguous 16-byte vectors.
// The wgrad kernel would otherwise re-evaluate every basis value kh * n_ct (6-12) times; see DESIGN.md.
__global__ void __launch_bounds__(256) kc_phi_flat_kernel(const __grid_constant__ WgArgs a, unsigned char* __restrict__ out) {
  __shared__ KcBasisCtx Bs;
  const kc_desc& d = a.d;
  const WgGeom& g = a.g;
  kc_load_basis_ctx(&Bs, d, a.beta);
  // one thread = one flat position of one chunk: the position is decoded once, then the chunk's 16 planes are produced
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  const int chunk = blockIdx.y;
  if (q >= g.L) return;
  const int HW = d.h * d.w, nb = d.nb > 4 ? 8 : 4;      // padded basis width
  const int n = (int)(q / g.IMG);
  const int rem = (int)(q - (long long)n * g.IMG);
  const int y = rem / g.P, x = rem - y * g.P;
  const bool inside = y < d.h && x < d.w;
  const long long off = inside ? (long long)n * d.x_batch_stride + y * d.w + x : 0;
  unsigned char* dst = out + ((long long)chunk * 16 * g.L + q) * 16;
  const long long pstride = g.L * 16;
  if (chunk >= g.nsc) {
    const int cb = (chunk - g.nsc) * 128;
#pragma unroll 4
    for (int pl = 0; pl < 16; ++pl) {
      if ((chunk - g.nsc) * 16 + pl >= g.vb_planes) break;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cb + pl * 8 + i;
        f[i] = (inside && c < d.cin) ? tc_act(d.act, __ldg(a.x_base + off + (long long)c * HW)) : 0.0f;
      }
      *reinterpret_cast<uint4*>(dst + pl * pstride) =
          make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    }
  } else if (nb == 8) {
    float xv[16];
#pragma unroll
    for (int pl = 0; pl < 16; ++pl) {
      const int c = chunk * g.cps + pl;
      xv[pl] = (inside && c < d.cin) ? __ldg(a.x_basis + off + (long long)c * HW) : 0.0f;
    }
#pragma unroll
    for (int pl = 0; pl < 16; ++pl) {
      if (chunk * 16 + pl >= g.vs_planes) break;
      const bool ok = inside && chunk * g.cps + pl < d.cin;
      uint4 v;
      if (g.fast_cubic) v = cubic8(xv[pl], g.t0, g.inv_h, Bs.nparams - 1, ok);
      else v = ok ? basis8w_generic(Bs, xv[pl]) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(dst + pl * pstride) = v;
    }
  } else {
    for (int pl = 0; pl < 16; ++pl) {
      if (chunk * 16 + pl >= g.vs_planes) break;
      const int c = chunk * g.cps + pl * 2;
      uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
      if (inside && c < d.cin) lo = basis4w(Bs, __ldg(a.x_basis + off + (long long)c * HW));
      if (inside && c + 1 < d.cin) hi = basis4w(Bs, __ldg(a.x_basis + off + (long long)(c + 1) * HW));
      *reinterpret_cast<uint4*>(dst + pl * pstride) = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
  }
}

// Producer role of the wgrad kernel (512 threads): pure 16-byte cp.async copies of the Phi planes (A, with the filter-row
// shift) and the dz planes (B) straight into the stage; rows are the fastest index so global reads are contiguous.
// stages - 1 stages are kept in flight per thread (commit_group / wait_group), then the landed stage is fenced for the
// tensor-core proxy and published on its `full` barrier.

template <int KS, bool PAIR>
__device__ __forceinline__ void wg_produce(const WgArgs& a, unsigned char* smem, uint64_t* full, uint64_t* empty, int r, int chunk,
                                           int ct, long long blk0, int nblocks, uint32_t rank, bool dummy) {
  constexpr int kItems = KS == 64 ? 3 : KS == 96 ? 4 : 5;       // 16-byte vectors per producer thread and operand per stage
  const kc_desc& d = a.d;
  const WgGeom& g = a.g;
  const int tid = threadIdx.x;
  const int nAitems = g.arows * 16, nBitems = KS * g.bplanes_cta;
  const int bpl0 = ct * g.bplanes + (PAIR ? (int)rank * g.bplanes_cta : 0);        // first dz plane of this CTA
  const long long qoff = blk0 * KS + (long long)(r - d.pad_h) * g.P - d.pad_w;
  int arow[kItems], brow[kItems];
  int ashift[kItems];                                       // merged units: extra row offset of the item's filter row
  uint32_t adst[kItems], bdst[kItems];                      // byte offsets inside a stage, 0xffffffff = no item
  const unsigned char* aplane[kItems];
  const unsigned char* bplane[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int it = tid + kProdW * k;
    arow[k] = it % g.arows;
    const int apl = it / g.arows;
    // planes without data (channel padding of the tile, or the dummy partner of a pair) are zeroed once by the kernel prologue
    // and never copied
    bool aok = !dummy && (chunk < g.nsc ? chunk * 16 + apl < g.vs_planes : (chunk - g.nsc) * 16 + apl < g.vb_planes);
    int src_plane = chunk * 16 + (it < nAitems ? apl : 0);
    ashift[k] = 0;
    if (g.merged) {                                         // plane apl = (filter row, channel c < cin | base plane)
      const int mr = apl / g.mnv, v = apl - mr * g.mnv;
      aok = mr < d.kh;
      src_plane = v < d.cin ? v : g.nsc * 16 + (v - d.cin);
      ashift[k] = aok ? mr * g.P : 0;
    }
    adst[k] = (it < nAitems && aok) ? (uint32_t)(apl * g.aplane_bytes + arow[k] * 16) : 0xffffffffu;
    aplane[k] = a.phi + ((long long)src_plane * g.L) * 16;
    brow[k] = it % KS;
    const int bpl = it / KS;
    const bool bok = it < nBitems && (bpl0 + bpl) * 8 < g.cq;
    bdst[k] = (it < nBitems) ? (uint32_t)(g.a_bytes + bpl * g.bplane_bytes + brow[k] * 16) : 0xffffffffu;
    bplane[k] = bok ? a.dzf + ((long long)(bpl0 + bpl) * g.L) * 16 : nullptr;
  }
  KC_TRACER(trp, g_trace_w, 0, tid == 0);
  // The arrival on a stage's `full` barrier is attached to the completion of this thread's copies
  // (cp.async.mbarrier.arrive.noinc), so the thread never waits for its own data and every free stage is in flight.  (The
  // former commit_group / wait_group / fence.proxy.async sequence stalled each thread until ALL its outstanding copies had
  // landed: one stage in flight, a full memory latency per 128 positions.)
  int st = 0;
  uint32_t ph = 1;
  for (int it = 0; it < nblocks; ++it) {
    trp.stamp();
    mbar_wait(&empty[st], ph);
    trp.stamp();
    unsigned char* stage = smem + (size_t)st * g.stage_bytes;
    const long long q0 = qoff + (long long)it * KS, m0 = (blk0 + it) * KS;
    if (q0 >= 0 && q0 + (g.merged ? (long long)(d.kh - 1) * g.P : 0) + g.arows <= g.L && m0 + KS <= g.L) {     // interior block: no per-row range checks
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        if (adst[k] != 0xffffffffu) cp_async16(stage + adst[k], aplane[k] + (q0 + ashift[k] + arow[k]) * 16, 16u);
        if (bdst[k] != 0xffffffffu) {
          const bool ok = bplane[k] != nullptr;
          cp_async16(stage + bdst[k], ok ? bplane[k] + (m0 + brow[k]) * 16 : a.dzf, ok ? 16u : 0u);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        if (adst[k] != 0xffffffffu) {
          const long long q = q0 + ashift[k] + arow[k];
          const bool ok = q >= 0 && q < g.L;
          cp_async16(stage + adst[k], ok ? aplane[k] + q * 16 : a.phi, ok ? 16u : 0u);
        }
        if (bdst[k] != 0xffffffffu) {
          const long long m = m0 + brow[k];
          const bool ok = bplane[k] != nullptr && m < g.L;
          cp_async16(stage + bdst[k], ok ? bplane[k] + m * 16 : a.dzf, ok ? 16u : 0u);
        }
      }
    }
    cp_async_mbar_arrive_noinc(&full[st]);
    if (++st == g.stages) { st = 0; ph ^= 1; }
  }
}

// Straight-line issue of the KW x (kKS/16) MMAs of one stage (tap s reads the Phi planes from row s; k-step ks advances both
// operands by 16 rows).  Descriptor low words differ by small constants only.
template <bool PAIR>
__device__ __forceinline__ void wg_mma(uint32_t td, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (PAIR) tc_mma_bf16_pair(td, ad, bd, idesc, acc); else tc_mma_bf16(td, ad, bd, idesc, acc);
}
template <int S0, int S1, int KS, bool PAIR>          // taps [S0, S1)
__device__ __forceinline__ void wg_issue(uint32_t tmem_base, uint32_t ntile, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi,
                                         uint32_t b_hi, uint32_t idesc, uint32_t first) {
#pragma unroll
  for (int s = S0; s < S1; ++s) {
    const uint32_t td = tmem_base + (uint32_t)s * ntile;
#pragma unroll
    for (int ks = 0; ks < KS / 16; ++ks)
      wg_mma<PAIR>(td, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(s + 16 * ks)),
                   ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(16 * ks)), idesc, ks == 0 ? first : 1u);
  }
}

template <int KS, bool PAIR>
__global__ void __launch_bounds__(kThreadsW, 1) kc_wgrad_tc_kernel(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const kc_desc& d = a.d;
  const WgGeom& g = a.g;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * g.stage_bytes);
  uint64_t* full = bars;                         // [kMaxStagesW]
  uint64_t* empty = bars + kMaxStagesW;          // [kMaxStagesW]
  uint64_t* acc_full = bars + 2 * kMaxStagesW;
  uint64_t* pfull = bars + 2 * kMaxStagesW + 1;  // [kMaxStagesW] pair only, in the leader: the peer CTA's stage has landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + kNumBarsW);
  KcBasisCtx* B = reinterpret_cast<KcBasisCtx*>(reinterpret_cast<unsigned char*>(tmem_ptr) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  KC_TRACER(trl, g_trace_w, 2, threadIdx.x == 0);          // CTA life cycle: entry, set-up done, mainloop done, accumulators ready, end
  trl.stamp();
  // unit = ((cout tile * nchunks + chunk) * kh + r); in a pair the two CTAs of a cluster are two units of the SAME cout tile
  // (they share the dz stage) and a cout tile with an odd number of units gets a dummy partner that only loads its dz half
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  int unit, ct, chunk, r;
  bool dummy = false;
  if (g.merged) {
    unit = ct = blockIdx.x; chunk = 0; r = 0;
  } else if (PAIR) {
    ct = blockIdx.x / g.upc;
    const int u = blockIdx.x - ct * g.upc;
    dummy = u >= g.units_ct;
    r = u % d.kh;
    chunk = dummy ? 0 : u / d.kh;
    unit = ct * g.units_ct + u;
  } else {
    unit = blockIdx.x;
    r = unit % d.kh;
    chunk = (unit / d.kh) % g.nchunks;
    ct = unit / (d.kh * g.nchunks);
  }
  const int split = blockIdx.y;
  const long long blk0 = (long long)split * g.blk_per_split;
  const long long blk1 = min(g.nblk, blk0 + g.blk_per_split);
  const int nblocks = (int)(blk1 - blk0);

  if (threadIdx.x == 0) {
    const uint32_t nmw = d.kw == 3 ? 2u : 1u;          // MMA-issuing warps: with three taps, warp 16 takes taps 0-1, warp 17 tap 2
    for (int i = 0; i < kMaxStagesW; ++i) { mbar_init(&full[i], (uint32_t)kProdW); mbar_init(&empty[i], nmw); mbar_init(&pfull[i], 1u); }
    mbar_init(acc_full, nmw);
    fence_barrier_init();
  }
  if (warp == kMmaWarpW) { if (PAIR) tmem_alloc_pair(tmem_ptr, (uint32_t)g.tmem_cols); else tmem_alloc(tmem_ptr, (uint32_t)g.tmem_cols); }
  kc_load_basis_ctx(B, d, a.beta);
  {
    // Phi planes of this unit that hold no data (channel padding / dummy partner): zero them in every stage, once
    const int chunk_planes = dummy      ? 0
                             : g.merged ? d.kh * g.mnv
                             : chunk < g.nsc ? g.vs_planes - chunk * 16 : g.vb_planes - (chunk - g.nsc) * 16;
    const int nvalid = chunk_planes < 0 ? 0 : chunk_planes > 16 ? 16 : chunk_planes;
    if (nvalid < 16) {
      const int vec_per_plane = g.aplane_bytes / 16, nvec = (16 - nvalid) * vec_per_plane;
      for (int st = 0; st < g.stages; ++st) {
        uint4* base = reinterpret_cast<uint4*>(smem + (size_t)st * g.stage_bytes + (size_t)nvalid * g.aplane_bytes);
        for (int i = threadIdx.x; i < nvec; i += kThreadsW) base[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) { cluster_arrive(); cluster_wait(); }        // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  trl.stamp();

  if (warp < 16) {
    // ============================ producers: cp.async copies of the Phi (A) and dz (B) planes ===========================
    wg_produce<KS, PAIR>(a, smem, full, empty, r, chunk, ct, blk0, nblocks, rank, dummy);
    trl.stamp();
  }
  if (warp >= kMmaWarpW && (warp == kMmaWarpW || d.kw == 3) && rank == 0) {
    // ============================ MMA issuers (whole warp uniform, one elected lane issues) =========================
    // The taps have separate accumulators, so two warps can issue independently: the barrier wait / commit overhead of one
    // overlaps the MMAs of the other.  In a pair only the leader CTA issues (M = 256: its own 128 Phi rows and the peer's).
    const int mw = warp - kMmaWarpW;
    const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, g.ntile, 1, 1);   // both operands MN-major
    const uint32_t lo_c = (128u >> 4) << 16;                            // LBO = 128 B between 8-row K groups
    const uint32_t a_hi = ((uint32_t)g.aplane_bytes >> 4) | (1u << 14); // SBO = plane pitch (8 M rows), version 1
    const uint32_t b_hi = ((uint32_t)g.bplane_bytes >> 4) | (1u << 14);
    const uint32_t smem_u = smem_u32(smem) >> 4, stage_u = (uint32_t)g.stage_bytes >> 4, a_u = (uint32_t)g.a_bytes >> 4;
    const int kw = d.kw, ntile = g.ntile;
    int st = 0;
    uint32_t ph = 0;
    KC_TRACER(trm, g_trace_w, 1, lane == 0 && mw == 0);
    for (int bi = 0; bi < nblocks; ++bi) {
      trm.stamp();
      mbar_wait(&full[st], ph);
      if (PAIR) mbar_wait_cluster(&pfull[st], ph);
      tc_fence_after();
      trm.stamp();
      const uint32_t au = smem_u + (uint32_t)st * stage_u, bu = au + a_u;
      const uint32_t first = bi != 0 ? 1u : 0u;
      if (elect_one_sync()) {
        const uint32_t a_lo = lo_c | au, b_lo = lo_c | bu;
        if (kw == 3 && mw == 0) wg_issue<0, 2, KS, PAIR>(tmem_base, (uint32_t)ntile, a_lo, b_lo, a_hi, b_hi, idesc, first);
        else if (kw == 3) wg_issue<2, 3, KS, PAIR>(tmem_base, (uint32_t)ntile, a_lo, b_lo, a_hi, b_hi, idesc, first);
        else if (kw == 1) wg_issue<0, 1, KS, PAIR>(tmem_base, (uint32_t)ntile, a_lo, b_lo, a_hi, b_hi, idesc, first);
        else {
          for (int s = 0; s < kw; ++s) {
            const uint32_t td = tmem_base + (uint32_t)(s * ntile);
#pragma unroll
            for (int ks = 0; ks < KS / 16; ++ks)
              wg_mma<PAIR>(td, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(s + 16 * ks)),
                           ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(16 * ks)), idesc, ks == 0 ? first : 1u);
          }
        }
        if (PAIR) tc_commit_pair(&empty[st]); else tc_commit(&empty[st]);
      }
      __syncwarp();
      if (++st == g.stages) { st = 0; ph ^= 1; }
    }
    if (elect_one_sync()) { if (PAIR) tc_commit_pair(acc_full); else tc_commit(acc_full); }
    __syncwarp();
  }
  if (PAIR && rank == 1 && warp == kMmaWarpW) {
    // ============================ peer CTA: tell the leader that this CTA's half of a stage has landed ===============
    int st = 0;
    uint32_t ph = 0;
    for (int bi = 0; bi < nblocks; ++bi) {
      mbar_wait(&full[st], ph);
      if (lane == 0) mbar_arrive_cluster(&pfull[st], 0);
      __syncwarp();
      if (++st == g.stages) { st = 0; ph ^= 1; }
    }
  }
  if (warp < 16) {
    // ============================ epilogue: TMEM -> fp32 partial sums in the split workspace ========================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    trl.stamp();
    const int quarter = warp & 3, cgrp = warp >> 2;
    const int m = quarter * 32 + lane;
    const int ncol16 = dummy ? 0 : d.kw * g.ntile / 16;
    float* wsu = a.ws + ((long long)split * g.units + unit) * ((long long)d.kw * 128 * g.ntile);
    for (int c16 = cgrp; c16 < ncol16; c16 += 4) {
      uint32_t rr[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c16 * 16), rr);
      tmem_ld_wait();
      const int col = c16 * 16, s = col / g.ntile, n = col - s * g.ntile;
      float4* dst = reinterpret_cast<float4*>(wsu + ((long long)s * 128 + m) * g.ntile + n);
      const bool live = nblocks > 0;
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4)
        dst[v4] = live ? make_float4(__uint_as_float(rr[4 * v4]), __uint_as_float(rr[4 * v4 + 1]), __uint_as_float(rr[4 * v4 + 2]),
                                     __uint_as_float(rr[4 * v4 + 3]))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  trl.stamp();
  if (PAIR) { cluster_arrive(); cluster_wait(); }        // both CTAs are done with the pair's tensor memory and shared memory
  if (warp == kMmaWarpW) { if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)g.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols); }
}

// Sum the splits (fixed order) and scatter into the reference layouts dw_basis [cout][cin*nb][kh][kw] (inner index
// c*nb+j, or j*cin+c for GRAM) and dw_base [cout][cin][kh][kw].  Threads walk the workspace layout, so the nsplit reads
// per element are coalesced; only the single write per weight is scattered.
__global__ void __launch_bounds__(256) kc_wgrad_tc_reduce_kernel(const __grid_constant__ kc_desc d, const __grid_constant__ WgGeom g,
                                                                 const float* __restrict__ ws, float* __restrict__ dw_base,
                                                                 float* __restrict__ dw_basis) {
  const int nb = d.nb, T = d.kh * d.kw;
  const long long unit_sz = (long long)d.kw * 128 * g.ntile;
  const long long total = (long long)g.units * unit_sz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % g.ntile);
    long long rest = i / g.ntile;
    const int m = (int)(rest % 128); rest /= 128;
    const int s = (int)(rest % d.kw);
    const int unit = (int)(rest / d.kw);
    int r = unit % d.kh;
    int chunk = (unit / d.kh) % g.nchunks;
    int ct = unit / (d.kh * g.nchunks);
    int mm = m;                                  // row inside the (chunk, r) tile
    if (g.merged) {                              // row m = 8 x plane (r, v) + j: v < cin a spline plane, else a base plane
      ct = unit;
      const int apl = m / 8, mr = apl / g.mnv, v = apl - mr * g.mnv;
      if (mr >= d.kh) continue;
      r = mr;
      if (v < d.cin) { chunk = 0; mm = v * 8 + (m & 7); } else { chunk = g.nsc; mm = (v - d.cin) * 8 + (m & 7); }
    }
    const int co = ct * g.ntile + n;
    if (co >= d.cout) continue;
    float* dst;
    if (chunk < g.nsc) {
      const int nbp = nb > 4 ? 8 : 4;
      const int c = chunk * g.cps + mm / nbp, j = mm % nbp;
      if (c >= d.cin || j >= nb) continue;
      dst = dw_basis + ((long long)co * d.cin * nb + kc_wbasis_index(d.basis, c, j, d.cin, nb)) * T + r * d.kw + s;
    } else {
      const int c = (chunk - g.nsc) * 128 + mm;
      if (c >= d.cin) continue;
      dst = dw_base + ((long long)co * d.cin + c) * T + r * d.kw + s;
    }
    float acc = 0.0f;
    for (int k = 0; k < g.nsplit; ++k) acc += ws[(long long)k * total + i];
    *dst = acc;
  }
}

// Same reduction through a shared-memory transpose (filters with kh*kw <= 9): a block owns 32 expanded rows x 32 couts of one
// (cout tile, chunk) for ALL taps, sums the splits with coalesced 128-byte reads, and writes runs that are contiguous in
// the reference layout ([cout][(c, j)][kh][kw]: 32 rows x T taps = 288 consecutive floats per cout for nb = 8).
constexpr int kRedT = 9;
// KH, KW > 0: compile-time filter size - the kh * kw * 4 * nsplit workspace reads of a thread are independent, and with the
// loops unrolled they are all in flight at once (the run-time loops issued one 128-byte row per round trip: 124 us for the
// 100 MB of a 512 -> 512 layer, 0.8 TB/s).  KH = 0: any filter with kh * kw <= 9.
template <int KH, int KW>
__global__ void __launch_bounds__(256) kc_wgrad_tc_reduce_tile_kernel(const __grid_constant__ kc_desc d, const __grid_constant__ WgGeom g,
                                                                      const float* __restrict__ ws, float* __restrict__ dw_base,
                                                                      float* __restrict__ dw_basis) {
  __shared__ float tile[32][32 * kRedT + 1];
  const int nb = d.nb, T = d.kh * d.kw;
  const int nsub_n = (g.ntile + 31) / 32;
  const int n0 = (blockIdx.x % nsub_n) * 32, m0 = (blockIdx.x / nsub_n) * 32;
  const int chunk = blockIdx.y % g.nchunks, ct = blockIdx.y / g.nchunks;
  const long long unit_sz = (long long)d.kw * 128 * g.ntile;
  const long long total = (long long)g.units * unit_sz;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool ncol_ok = n0 + lane < g.ntile;
  if (KH > 0) {
    float acc[KH * KW * 4 + 1];
#pragma unroll
    for (int e = 0; e < KH * KW * 4; ++e) acc[e] = 0.0f;
    if (ncol_ok) {
      const float* base = ws + (long long)((ct * g.nchunks + chunk) * KH) * unit_sz + (long long)(m0 + warp) * g.ntile + n0 + lane;
      for (int k = 0; k < g.nsplit; ++k) {                      // splits in fixed order (deterministic)
        const float* bk = base + (long long)k * total;
#pragma unroll
        for (int r = 0; r < KH; ++r)
#pragma unroll
          for (int s = 0; s < KW; ++s)
#pragma unroll
            for (int mm = 0; mm < 4; ++mm)
              acc[(r * KW + s) * 4 + mm] += __ldg(bk + (long long)r * unit_sz + ((long long)s * 128 + 8 * mm) * g.ntile);
      }
    }
#pragma unroll
    for (int r = 0; r < KH; ++r)
#pragma unroll
      for (int s = 0; s < KW; ++s)
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) tile[lane][(warp + 8 * mm) * (KH * KW) + r * KW + s] = acc[(r * KW + s) * 4 + mm];
  } else {
    for (int r = 0; r < d.kh; ++r) {
      const int unit = (ct * g.nchunks + chunk) * d.kh + r;
      for (int s = 0; s < d.kw; ++s)
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
          const int ml = warp + 8 * mm;
          const long long i = (long long)unit * unit_sz + ((long long)s * 128 + m0 + ml) * g.ntile + n0 + lane;
          float acc = 0.0f;
          if (ncol_ok)
            for (int k = 0; k < g.nsplit; ++k) acc += ws[(long long)k * total + i];
          tile[lane][ml * T + r * d.kw + s] = acc;
        }
    }
  }
  __syncthreads();
  const int run = 32 * T;
  for (int idx = threadIdx.x; idx < 32 * run; idx += 256) {
    const int col = idx / run, i = idx - col * run;
    const int ml = i / T, t = i - ml * T;
    const int co = ct * g.ntile + n0 + col, m = m0 + ml;
    if (n0 + col >= g.ntile || co >= d.cout) continue;
    float* dst;
    if (chunk < g.nsc) {
      const int nbp = nb > 4 ? 8 : 4;
      const int c = chunk * g.cps + m / nbp, j = m % nbp;
      if (c >= d.cin || j >= nb) continue;
      dst = dw_basis + ((long long)co * d.cin * nb + kc_wbasis_index(d.basis, c, j, d.cin, nb)) * T + t;
    } else {
      const int c = (chunk - g.nsc) * 128 + m;
      if (c >= d.cin) continue;
      dst = dw_base + ((long long)co * d.cin + c) * T + t;
    }
    *dst = tile[col][i];
  }
}

int wgrad_geometry(const kc_desc* d, WgGeom* g) {
  if (d->dil_h != 1 || d->dil_w != 1) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad needs dilation 1");
  if (d->stride_h < 1 || d->stride_w < 1 || d->stride_h > 4 || d->stride_w > 4) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad needs stride <= 4");
  if (d->nb < 1 || d->nb > 8) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad needs basis width <= 8 (got %d)", d->nb);
  if (d->pad_h > d->kh - 1 || d->pad_w > d->kw - 1) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad needs padding < kernel size");
  if (d->kw > 8 || d->kh > 8) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad needs kernel size <= 8");
  if ((long long)d->n * d->x_batch_stride >= (1LL << 31)) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad needs < 2^31 input elements");
  memset(g, 0, sizeof(*g));
  const bool has_base = d->act != KC_ACT_NONE;
  g->P = d->w + d->pad_w;
  g->IMG = (d->h + d->pad_h) * g->P;
  g->L = (long long)d->n * g->IMG;
  if (g->L >= (1LL << 31)) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad needs < 2^31 flat positions");
  g->cq = round_up_w(d->cout, 16);
  g->cps = 128 / (d->nb > 4 ? 8 : 4);
  g->nsc = (d->cin + g->cps - 1) / g->cps;
  g->nbc = has_base ? (d->cin + 127) / 128 : 0;
  g->nchunks = g->nsc + g->nbc;
  {
    const int Cp = round_up_w(d->cin, 8);            // channels the forward kernel expands (its k-cores hold 8 channels)
    g->vs_planes = d->nb > 4 ? Cp : Cp / 2;
    g->vb_planes = has_base ? round_up_w(Cp / 8, 2) : 0;
    if (g->vs_planes > g->nsc * 16) g->vs_planes = g->nsc * 16;
    if (g->vb_planes > g->nbc * 16) g->vb_planes = g->nbc * 16;
  }
  int nmax = (512 / d->kw) / 16 * 16;
  if (nmax > 192) nmax = 192;            // 64 positions x (ntile / 8) dz planes must fit the producer mapping (3 vectors per thread)
  int want = (g->cq + nmax - 1) / nmax;
  g->ntile = round_up_w((g->cq + want - 1) / want, 16);
  g->n_ct = (g->cq + g->ntile - 1) / g->ntile;
  g->units = g->n_ct * g->nchunks * d->kh;
  {
    // stem layers: every filter row and both branches fit ONE M = 128 tile -> one unit per cout tile instead of nchunks * kh
    static const int merge_enabled = []() { const char* e = getenv("KANCONV_WGRAD_MERGE"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
    const int nbase = has_base ? (d->cin + 7) / 8 : 0;
    if (merge_enabled && d->nb > 4 && d->nb <= 8 && d->kh * (d->cin + nbase) <= 16) {
      g->merged = 1;
      g->mnv = d->cin + nbase;
      g->units = g->n_ct;
    }
  }
  // CTA pairs (tcgen05 cta_group::2, M = 256): two units of a cout tile share every dz stage and each CTA loads half of it.
  // An M = 128 MMA with N <= 128 is bound by the operand fetch from shared memory (66 / 74 cycles at N = 64 / 128 instead of
  // 32 / 64); the pair halves the B fetch per CTA (43 / 64 cycles, tools/mma_rate_2cta.py).  KANCONV_WGRAD_PAIR=0: single CTAs.
  g->units_ct = g->merged ? 1 : g->nchunks * d->kh;
  {
    static const int pair_enabled = []() { const char* e = getenv("KANCONV_WGRAD_PAIR"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
    g->pair = (pair_enabled && g->units_ct >= 2) ? 1 : 0;
  }
  g->upc = g->pair ? round_up_w(g->units_ct, 2) : g->units_ct;
  // positions per ring stage: as many MMAs per barrier round as the producer mapping and a >= 3-deep ring allow
  // (128 -> 24 MMAs per tap row for cout tiles <= 64, else 96 -> 18, else 64 -> 12)
  g->bplanes = g->ntile / 8;
  g->bplanes_cta = g->pair ? g->bplanes / 2 : g->bplanes;
  const int cand[3] = {128, 96, 64};
  size_t fixed = kNumBarsW * 8 + 16 + sizeof(KcBasisCtx) + 128;
  bool ok = false;
  for (int ci = 0; ci < 3 && !ok; ++ci) {
    g->ks = cand[ci];
    g->arows = round_up_w(g->ks + d->kw - 1, 8);
    g->aplane_bytes = g->arows * 16 + 16;
    g->a_bytes = 16 * g->aplane_bytes;
    g->bplane_bytes = g->ks * 16 + 16;
    g->stage_bytes = round_up_w(g->a_bytes + g->bplanes_cta * g->bplane_bytes, 128);
    const int items = g->ks == 64 ? 3 : g->ks == 96 ? 4 : 5;
    if (g->arows * 16 > items * kProdW || g->ks * g->bplanes_cta > items * kProdW) continue;
    g->stages = (int)((kSmemLimitW - fixed) / g->stage_bytes);
    if (g->stages > kMaxStagesW) g->stages = kMaxStagesW;
    if (g->stages < (g->ks == 64 ? 2 : 3)) continue;
    ok = true;
  }
  if (!ok) KC_FAIL(KC_ERR_UNSUPPORTED, "tensor-core wgrad: stage does not fit shared memory / the producer mapping");
  g->prefetch = g->stages - 1;                     // informational: stages in flight while one is being consumed
  g->smem_bytes = fixed + (size_t)g->stages * g->stage_bytes;
  g->tmem_cols = 32;
  while (g->tmem_cols < d->kw * g->ntile) g->tmem_cols *= 2;
  g->nblk = (g->L + g->ks - 1) / g->ks;
  // Split-K count.  Every CTA of a launch does the same work and one CTA fits an SM, so the launch runs in whole waves of
  // kc_sm_count() CTAs: among the counts that give about 2-4 waves take the one whose last wave is fullest (units * splits
  // just below a multiple of the SM count; e.g. 15 units: 29 splits = 2.94 waves instead of 30 = 3.04), ties to the fewest
  // splits (less workspace traffic, fewer prologues / epilogues).
  const long long sms = kc_sm_count();
  const long long minblk = 1024 / g->ks;                            // at least 1024 positions per CTA
  const long long max_split = g->nblk / minblk > 0 ? g->nblk / minblk : 1;
  const long long grid_x = g->pair ? (long long)g->n_ct * g->upc : g->units;      // CTAs per split (incl. dummy partners)
  long long lo = (2 * sms + grid_x - 1) / grid_x, hi = (4 * sms) / grid_x;
  if (lo < 1) lo = 1;
  if (hi < lo) hi = lo;
  if (hi > max_split) hi = max_split;
  if (lo > hi) lo = hi;
  long long ns = lo;
  double best = -1.0;
  for (long long c = lo; c <= hi; ++c) {
    const long long bps = (g->nblk + c - 1) / c, real = (g->nblk + bps - 1) / bps;   // splits that actually get blocks
    const long long ctas = real * grid_x, waves = (ctas + sms - 1) / sms;
    // time ~ waves * blocks per CTA (+ a fixed prologue / epilogue of ~10 block-times per CTA)
    const double t = (double)waves * (double)(bps + 10);
    const double score = (double)g->nblk * grid_x / (double)sms / t;
    if (score > best + 1e-9) { best = score; ns = c; }
  }
  g->blk_per_split = (g->nblk + ns - 1) / ns;
  g->nsplit = (int)((g->nblk + g->blk_per_split - 1) / g->blk_per_split);
  g->ws_bytes = (((size_t)g->nsplit * g->units * d->kw * 128 * g->ntile * sizeof(float)) + 255) / 256 * 256;
  g->phi_bytes = (size_t)g->nchunks * 16 * (size_t)g->L * 16;
  g->fast_cubic = kc_knots_uniform_cubic(d, &g->t0, &g->inv_h) ? 1 : 0;
  return KC_OK;
}

template <int KS, bool PAIR>
cudaError_t launch_wgrad(const WgArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kc_wgrad_tc_kernel<KS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreadsW, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2u : 1u;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kc_wgrad_tc_kernel<KS, PAIR>, a);
}

}  // namespace

#ifdef KANCONV_DEBUG
extern "C" int kc_debug_trace_wgrad(void* device_buffer) {       // debug build only, not part of include/kanconv.h
  long long* p = (long long*)device_buffer;
  KC_CUDA_CHECK(cudaMemcpyToSymbol(g_trace_w, &p, sizeof(p)));
  return KC_OK;
}
#endif

// which: 0 = split-K workspace + transient basis buffer, 1 = saved basis rows (phi), 2 = split-K workspace only
size_t kc_tc_wgrad_ws_bytes(const kc_desc* d, int which) {
  WgGeom g;
  if (wgrad_geometry(d, &g) != KC_OK) return 0;
  return which == 0 ? g.ws_bytes + g.phi_bytes : which == 1 ? g.phi_bytes : g.ws_bytes;
}

// Plane counts of the phi layout (for the forward kernel that fills it): flat length, spline planes, base planes.
int kc_tc_wgrad_phi_layout(const kc_desc* d, long long* L, int* spline_planes, int* base_planes) {
  WgGeom g;
  int rc = wgrad_geometry(d, &g);
  if (rc != KC_OK) return rc;
  *L = g.L; *spline_planes = g.nsc * 16; *base_planes = g.nbc * 16;
  return KC_OK;
}

// Basis rows of a layer input in the phi layout (the pre-pass of the weight gradient, also used by the pointwise forward).
int kc_tc_phi_prepass(const kc_desc* d, const float* x_base, const float* x_basis, const float* beta, void* phi, void* stream) {
  WgGeom g;
  int rc = wgrad_geometry(d, &g);
  if (rc != KC_OK) return rc;
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.d = *d; a.g = g; a.x_base = x_base; a.x_basis = x_basis; a.beta = beta;
  dim3 pgrid((unsigned)((g.L + 255) / 256), (unsigned)g.nchunks);
  kc_phi_flat_kernel<<<pgrid, 256, 0, (cudaStream_t)stream>>>(a, (unsigned char*)phi);
  KC_LAUNCH_CHECK("kc_phi_flat_kernel");
  return KC_OK;
}

extern "C" int kc_conv_wgrad_tc(const kc_desc* d, const void* dz_flat, const float* x_base, const float* x_basis,
                                const float* beta, const void* phi_saved, float* dw_base, float* dw_basis, void* workspace,
                                void* stream) {
  int rc = kc_validate_desc(d);
  if (rc != KC_OK) return rc;
  WgGeom g;
  rc = wgrad_geometry(d, &g);
  if (rc != KC_OK) return rc;
  if (!dz_flat || !x_basis || !dw_basis || !workspace) KC_FAIL(KC_ERR_INVALID, "kc_conv_wgrad_tc: null pointer");
  if (d->act != KC_ACT_NONE && (!x_base || !dw_base)) KC_FAIL(KC_ERR_INVALID, "kc_conv_wgrad_tc: base branch needs x_base and dw_base");
  if (d->basis == KC_BASIS_GRAM && !beta) KC_FAIL(KC_ERR_INVALID, "kc_conv_wgrad_tc: GRAM basis needs beta_weights");
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.d = *d; a.g = g; a.x_base = x_base; a.x_basis = x_basis; a.dzf = (const unsigned char*)dz_flat; a.beta = beta;
  a.ws = (float*)workspace;
  if (phi_saved != nullptr) {
    a.phi = (const unsigned char*)phi_saved;          // rows written by kc_conv_fwd_tc
  } else {
    unsigned char* phi = (unsigned char*)workspace + g.ws_bytes;
    a.phi = phi;
    dim3 pgrid((unsigned)((g.L + 255) / 256), (unsigned)g.nchunks);
    kc_phi_flat_kernel<<<pgrid, 256, 0, (cudaStream_t)stream>>>(a, phi);
    KC_LAUNCH_CHECK("kc_phi_flat_kernel");
  }
  if (g.pair) {
    dim3 grid((unsigned)(g.n_ct * g.upc), (unsigned)g.nsplit);
    cudaError_t e = g.ks == 96    ? launch_wgrad<96, true>(a, grid, g.smem_bytes, (cudaStream_t)stream)
                    : g.ks == 128 ? launch_wgrad<128, true>(a, grid, g.smem_bytes, (cudaStream_t)stream)
                                  : launch_wgrad<64, true>(a, grid, g.smem_bytes, (cudaStream_t)stream);
    KC_CUDA_CHECK(e);
  } else {
    dim3 grid((unsigned)g.units, (unsigned)g.nsplit);
    cudaError_t e = g.ks == 96    ? launch_wgrad<96, false>(a, grid, g.smem_bytes, (cudaStream_t)stream)
                    : g.ks == 128 ? launch_wgrad<128, false>(a, grid, g.smem_bytes, (cudaStream_t)stream)
                                  : launch_wgrad<64, false>(a, grid, g.smem_bytes, (cudaStream_t)stream);
    KC_CUDA_CHECK(e);
  }
  KC_LAUNCH_CHECK("kc_wgrad_tc_kernel");
  long long total = (long long)g.units * d->kw * 128 * g.ntile;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kc_sm_count() * 16) blocks = kc_sm_count() * 16;
  // the tile kernel needs enough (cout tile, chunk) pairs to fill the machine; few-channel layers (many splits, few
  // units) keep the element-parallel kernel
  const long long tile_blocks = (long long)((g.ntile + 31) / 32) * 4 * g.n_ct * g.nchunks;
  if (!g.merged && d->kh * d->kw <= kRedT && (long long)g.n_ct * g.nchunks <= 65535 && tile_blocks >= 2 * kc_sm_count()) {
    dim3 rgrid((unsigned)(((g.ntile + 31) / 32) * 4), (unsigned)(g.n_ct * g.nchunks));
    if (d->kh == 3 && d->kw == 3) kc_wgrad_tc_reduce_tile_kernel<3, 3><<<rgrid, 256, 0, (cudaStream_t)stream>>>(*d, g, (const float*)workspace, dw_base, dw_basis);
    else kc_wgrad_tc_reduce_tile_kernel<0, 0><<<rgrid, 256, 0, (cudaStream_t)stream>>>(*d, g, (const float*)workspace, dw_base, dw_basis);
  } else {
    kc_wgrad_tc_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*d, g, (const float*)workspace, dw_base, dw_basis);
  }
  KC_LAUNCH_CHECK("kc_wgrad_tc_reduce_kernel");
  return KC_OK;
}
