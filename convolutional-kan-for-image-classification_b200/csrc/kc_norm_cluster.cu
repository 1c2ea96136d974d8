// kc_norm_cluster.cu - HBM-roofline versions of the InstanceNorm kernels for large planes (sm_100a):
// the part of a plane a CTA works on stays RESIDENT IN SHARED MEMORY between the statistics pass and the apply pass, a plane
// that does not fit one CTA is split over a thread-block cluster, and the per-CTA partial statistics are exchanged through
// distributed shared memory.  Every activation byte is read from HBM once (the two-pass kernels of kc_norm.cu re-read the
// plane from L2, which misses for the 200 KB planes of the 224x224 layers: measured 1.4-1.7x the algorithmic traffic).
//
//   kc_instnorm_fwd_cluster_kernel   y = out_act(gamma * (z - mean) * rstd + beta)          8 B per element (read z, write y)
//   kc_norm_bwd_flat_cluster_kernel  backward of (InstanceNorm -> out_act) fused with the conversion of dz to the bf16 "flat"
//                                    operand layout of the tensor-core dgrad / wgrad kernels: 8 B read (dy, z) + 2 B written
//                                    per element instead of 12 B (norm backward) + 6 B (kc_dz_flat_kernel).
// The bulk loads are TMA-engine copies (cp.async.bulk, SASS UBLKCP) of whole contiguous plane chunks: no registers, the
// whole chunk is in flight at once.  Replaces native_batch_norm(_backward) + prelu / silu (+ backward) of
// kan_layers.py:241-243, gram_kan_layers.py:187, cheby_kan_layers.py:98.
#include <stdlib.h>

#include "kc_common.cuh"
#include "kc_norm_common.cuh"
#include "kc_umma.cuh"

int kc_tc_flat_layout(const kc_desc* d, int* P, int* IMG, long long* L, int* cq);      // kc_tc.cu

namespace {

using namespace kc;

constexpr int kFwdThreads = 256;
constexpr int kBwdThreadsBig = 1024, kBwdThreadsSmall = 256;  // one CTA per SM (chunk up to 200 KB) / four CTAs per SM (<= 52 KB)
constexpr int kBwdThreadsMid = 512;                           // two CTAs per SM (chunk <= 100 KB, clusters of 16)
constexpr int kBwdThreadsTiny = 64;                           // chunks of <= 64 float4 columns per channel (14x14 planes)
constexpr uint32_t kBulkPiece = 32 * 1024;       // bytes per cp.async.bulk request
constexpr size_t kSmemMax = 200 * 1024 + 1024;   // largest resident chunk (8 channels x 28 rows x 224 floats = 200 704 B)

// issue the bulk copies of `bytes` (multiple of 16) from src to dst in pieces; all complete on `bar`
__device__ __forceinline__ void bulk_load_chunk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  for (uint32_t o = 0; o < bytes; o += kBulkPiece)
    bulk_g2s((unsigned char*)dst + o, (const unsigned char*)src + o, min(kBulkPiece, bytes - o), bar);
}

// ---------------------------------------------------------------------------------------------------------
// forward: one cluster per plane (n, c); CTA `rank` owns elements [rank * chunk, rank * chunk + cnt)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads)
kc_instnorm_fwd_cluster_kernel(const __grid_constant__ kc_norm_desc d, int chunk, const float* __restrict__ z,
                               const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ alpha_p,
                               float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* buf = reinterpret_cast<float*>(smem_raw);
  __shared__ float sh[32];
  __shared__ float part[4];          // this CTA's {count, mean, M2}: read by the peers through DSMEM
  __shared__ float stat[2];          // combined {mean, rstd}
  __shared__ __align__(8) uint64_t bar;
  const uint32_t cs = cluster_nctarank(), rank = cluster_ctarank();
  const int plane = blockIdx.x / cs, n = plane / d.c, ch = plane % d.c;
  const long long off = (long long)n * d.batch_stride + (long long)ch * d.hw + (long long)rank * chunk;
  const int cnt = max(0, min(chunk, d.hw - (int)rank * chunk));
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0 && cnt > 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)cnt * 4u);
    bulk_load_chunk(buf, z + off, (uint32_t)cnt * 4u, &bar);
  }
  if (cnt > 0) mbar_wait(&bar, 0);
  const float4* b4 = reinterpret_cast<const float4*>(buf);
  const int n4 = cnt >> 2;
  float s = 0.0f;
  for (int i = threadIdx.x; i < n4; i += kFwdThreads) { const float4 v = b4[i]; s += (v.x + v.y) + (v.z + v.w); }
  const float lmean = cnt > 0 ? block_sum(s, sh) / (float)cnt : block_sum(s, sh);
  float m2 = 0.0f;
  for (int i = threadIdx.x; i < n4; i += kFwdThreads) {
    const float4 v = b4[i];
    const float a = v.x - lmean, b = v.y - lmean, c = v.z - lmean, e = v.w - lmean;
    m2 = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(e, e, m2))));
  }
  m2 = block_sum(m2, sh);
  if (threadIdx.x == 0) { part[0] = (float)cnt; part[1] = lmean; part[2] = m2; }
  __syncthreads();
  cluster_arrive();                  // release: part[] is visible to the cluster
  cluster_wait();
  if (threadIdx.x == 0) {
    // Chan et al. combination of the per-CTA (count, mean, M2) in rank order: deterministic
    float gm = 0.0f;
    for (uint32_t r = 0; r < cs; ++r) gm += dsmem_ld(&part[0], r) * dsmem_ld(&part[1], r);
    gm /= (float)d.hw;
    float M2 = 0.0f;
    for (uint32_t r = 0; r < cs; ++r) {
      const float dm = dsmem_ld(&part[1], r) - gm;
      M2 += dsmem_ld(&part[2], r) + dsmem_ld(&part[0], r) * dm * dm;
    }
    const float rs = rsqrtf(M2 / (float)d.hw + d.eps);          // biased variance, like F.instance_norm
    stat[0] = gm; stat[1] = rs;
    if (rank == 0) { mean_out[plane] = gm; rstd_out[plane] = rs; }
  }
  __syncthreads();
  cluster_arrive();                  // this CTA no longer reads its peers' shared memory
  const float mean = stat[0], rstd = stat[1];
  const float g = (d.affine && gamma) ? gamma[ch] : 1.0f;
  const float b = (d.affine && beta) ? beta[ch] : 0.0f;
  const float alpha = (d.out_act == KC_OUT_PRELU) ? alpha_p[0] : 0.0f;
  const float sc = rstd * g, sh0 = b - mean * rstd * g;
  float4* y4 = reinterpret_cast<float4*>(y + off);
  const int kind = d.out_act;
  for (int i = threadIdx.x; i < n4; i += kFwdThreads) {
    const float4 v = b4[i];
    float4 o;
    o.x = out_act(kind, fmaf(v.x, sc, sh0), alpha);
    o.y = out_act(kind, fmaf(v.y, sc, sh0), alpha);
    o.z = out_act(kind, fmaf(v.z, sc, sh0), alpha);
    o.w = out_act(kind, fmaf(v.w, sc, sh0), alpha);
    y4[i] = o;
  }
  cluster_wait();                    // peers may still be reading part[]: do not exit before they are done
}

// ---------------------------------------------------------------------------------------------------------
// backward + bf16 flat dz: one cluster per (image n, group of 8 output channels); CTA `rank` owns `rows` rows of all
// eight planes.  zhat of the chunk stays in shared memory; dy is read twice (HBM, then L2: 200 KB per CTA in flight).
//   dv = dy * act'(zhat);  dz = rstd * (dv - mean_hw(dv) - zhat * mean_hw(dv * zhat))             (no affine)
// dz_flat layout (kc_tc.cu): plane pl = 8 channels, [pl][q = n*IMG + y*P + x][8] bf16, zero at the gap column(s) x >= wo
// and gap rows y >= ho.  partials[0][plane] = sum dv*zhat, [1][plane] = sum dv, [2][plane] = sum dy*zhat*[zhat<=0].
// ---------------------------------------------------------------------------------------------------------
struct NbfArgs {
  kc_norm_desc d;
  int ho, wo, rows;            // output map, rows per CTA
  int P, IMG;                  // flat row pitch, flat positions per image
  long long L;
  int groups8;                 // 8-channel planes of the flat buffer (cq / 8)
  const float* dy;
  const float* z;
  const float* mean;
  const float* rstd;
  const float* alpha;
  unsigned char* dzf;
  float* partials;
};

// Both phases are written for <= 64 registers per thread so that 32 warps are resident per SM (1024 threads with the 200 KB
// chunk of the 224x224 planes, 4 x 256 threads otherwise): an elementwise kernel at this bandwidth is ISSUE-bound unless the
// instruction count per element is small (a 512-thread / 128-register version ran at 2.7 TB/s; a version with the activation
// kind as a run-time switch and scalar dy loads executed 63 instructions per element - profiles/r2_ncu_summary.md).  Hence:
// activation kind and "all eight channels present" are template parameters, every global / shared access is 16 bytes wide.
template <int KIND>
__device__ __forceinline__ float act_grad_t(float v, float alpha) {
  if (KIND == KC_OUT_PRELU) return v > 0.0f ? 1.0f : alpha;
  if (KIND == KC_OUT_SILU) return kc_silu_grad(v);
  return 1.0f;
}

// CH = channels per CTA: 8 (one CTA writes whole 16-byte vectors of the flat buffer) or 4 (two CTAs each write one 8-byte half;
// used for the 224x224 planes so that the resident chunk is 100 KB and TWO CTAs share an SM - with one CTA per SM the bulk
// load, the two phases and the cluster barriers of a chunk run back to back and the SM idles in between).
template <int kBwdThreads, int CH, int KIND, bool FULL>
__global__ void __launch_bounds__(kBwdThreads, kBwdThreads >= 1024 ? 1 : kBwdThreads >= 512 ? 2 : kBwdThreads >= 256 ? 4 : 16)
kc_norm_bwd_flat_cluster_kernel(const __grid_constant__ NbfArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* zbuf = reinterpret_cast<float*>(smem_raw);           // [CH][cap] zhat of this CTA's rows
  __shared__ float red[kBwdThreads / 32][3 * CH];
  __shared__ float part[3 * CH];     // this CTA's sums, read by the peers through DSMEM
  __shared__ float4 coef[CH];        // phase 2: {rstd, rstd * mean(dv), rstd * mean(dv * zhat), 0} per channel
  __shared__ float stat[2 * CH];     // mean[CH], rstd[CH]
  __shared__ __align__(8) uint64_t bar[CH];     // one per channel: phase 1 starts on channel 0 while the others are in flight
  constexpr int kParts = 8 / CH;     // CTAs (clusters) that share one 8-channel group of the flat buffer
  const kc_norm_desc& d = a.d;
  const uint32_t cs = cluster_nctarank(), rank = cluster_ctarank();
  const int cl = blockIdx.x / cs, half = cl % kParts, n = (cl / kParts) / a.groups8, g8 = (cl / kParts) % a.groups8;
  const int c0 = g8 * 8 + half * CH, nch = FULL ? CH : max(0, min(CH, d.c - c0));
  const int r0 = (int)rank * a.rows, r1 = min(a.ho, r0 + a.rows);
  const int cnt = max(0, r1 - r0) * a.wo, cap = a.rows * a.wo, hw = a.ho * a.wo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)n * d.batch_stride + (long long)c0 * hw + (long long)r0 * a.wo;      // channel c: + c * hw
  if (tid < CH) mbar_init(&bar[tid], 1);
  if (tid == 0) fence_barrier_init();
  if (tid < 2 * CH) {
    const int c = tid % CH;
    stat[tid] = c < nch ? (tid < CH ? a.mean : a.rstd)[n * d.c + c0 + c] : 0.0f;
  }
  __syncthreads();
  if (tid == 0 && cnt > 0)
    for (int c = 0; c < nch; ++c) {
      mbar_arrive_expect_tx(&bar[c], (uint32_t)cnt * 4u);
      bulk_load_chunk(zbuf + (size_t)c * cap, a.z + base + (long long)c * hw, (uint32_t)cnt * 4u, &bar[c]);
    }
  const float alpha = (KIND == KC_OUT_PRELU) ? a.alpha[0] : 0.0f;
  // ---- phase 1: per-channel sums over this CTA's rows; zhat replaces z in shared memory ---------------------------
  // A thread owns at most two float4 columns of a channel (host: n4 <= 2 * threads).  The dy vectors of channel c + 1 are
  // requested before channel c is processed, and those of channel 0 before the wait for the bulk copy of z.  (A generic
  // prefetch queue over any number of columns executed 30 % more instructions and was slower.)
  const int n4 = cnt >> 2, hw4 = hw >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(a.dy + base);
  const int i0 = tid, i1 = tid + kBwdThreads;
  const bool has0 = i0 < n4, has1 = i1 < n4;
  const float4 zero_f4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ga = (has0 && nch > 0) ? ldg4_early(g4 + i0) : zero_f4;
  float4 gb = (has1 && nch > 0) ? ldg4_early(g4 + i1) : zero_f4;
#pragma unroll 1
  for (int c = 0; c < CH; ++c) {
    float s_dvz = 0.0f, s_dv = 0.0f, s_da = 0.0f;
    if (c < nch) {                                   // warp-uniform
      const float4 ca = ga, cb = gb;
      if (c + 1 < nch) {
        const float4* gn = g4 + (c + 1) * hw4;
        if (has0) ga = ldg4_early(gn + i0);
        if (has1) gb = ldg4_early(gn + i1);
      }
      const float mean = stat[c], rstd = stat[CH + c];
      float4* zb = reinterpret_cast<float4*>(zbuf + (size_t)c * cap);
      if (cnt > 0) mbar_wait(&bar[c], 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 0 ? has0 : has1) {
          const float4 g = h == 0 ? ca : cb;
          float4 zv = zb[h == 0 ? i0 : i1];
          zv.x = (zv.x - mean) * rstd; zv.y = (zv.y - mean) * rstd; zv.z = (zv.z - mean) * rstd; zv.w = (zv.w - mean) * rstd;
          zb[h == 0 ? i0 : i1] = zv;
          const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float dv = gg[e] * act_grad_t<KIND>(zz[e], alpha);
            s_dvz = fmaf(dv, zz[e], s_dvz);
            s_dv += dv;
            if (KIND == KC_OUT_PRELU && !(zz[e] > 0.0f)) s_da = fmaf(gg[e], zz[e], s_da);
          }
        }
      }
    }
    s_dvz = kc_warp_sum(s_dvz); s_dv = kc_warp_sum(s_dv);
    if (KIND == KC_OUT_PRELU) s_da = kc_warp_sum(s_da);
    if (lane == 0) { red[warp][c] = s_dvz; red[warp][CH + c] = s_dv; red[warp][2 * CH + c] = s_da; }
  }
  __syncthreads();
  if (tid < 3 * CH) {
    float v = 0.0f;
    for (int w = 0; w < kBwdThreads / 32; ++w) v += red[w][tid];
    part[tid] = v;
  }
  __syncthreads();
  cluster_arrive();
  cluster_wait();
  if (tid < 3 * CH) {
    float v = 0.0f;
    for (uint32_t r = 0; r < cs; ++r) v += dsmem_ld(&part[tid], r);       // rank order: deterministic
    red[0][tid] = v;
    const int c = tid % CH, which = tid / CH;
    if (rank == 0 && c < nch) a.partials[(long long)which * d.n * d.c + n * d.c + c0 + c] = v;
  }
  __syncthreads();
  if (tid < CH) {
    const float inv_hw = 1.0f / (float)hw, rs = stat[CH + tid];
    coef[tid] = make_float4(rs, rs * red[0][CH + tid] * inv_hw, rs * red[0][tid] * inv_hw, 0.0f);
  }
  __syncthreads();
  cluster_arrive();                  // done with the peers' shared memory
  // ---- phase 2: four consecutive positions x CH channels per thread -> four (half) vectors of the flat buffer --------
  //   dz = rstd * (dv - mean(dv) - zhat * mean(dv * zhat)) = coef.x * dv - (coef.y + zhat * coef.z)
  // dy is read again (L2 hits: this CTA streamed the same bytes in phase 1).  Channels are taken two at a time and packed
  // straight into the output words (bf16x2 = channels 2k, 2k+1); the dy vectors of the next channel pair are requested
  // before the current pair is evaluated.
  unsigned char* outp = a.dzf + ((long long)g8 * a.L + (long long)n * a.IMG) * 16 + half * (2 * CH);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < n4; i += kBwdThreads) {
    uint32_t o[4][CH / 2];
    float4 ga = (FULL || 0 < nch) ? ldg4_early(g4 + i) : zero_f4;
    float4 gb = (FULL || 1 < nch) ? ldg4_early(g4 + hw4 + i) : zero_f4;
#pragma unroll
    for (int k = 0; k < CH / 2; ++k) {
      const float4 gpair[2] = {ga, gb};
      if (k + 1 < CH / 2) {
        ga = (FULL || 2 * k + 2 < nch) ? ldg4_early(g4 + (2 * k + 2) * hw4 + i) : zero_f4;
        gb = (FULL || 2 * k + 3 < nch) ? ldg4_early(g4 + (2 * k + 3) * hw4 + i) : zero_f4;
      }
      float f[2][4];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = 2 * k + q;
        if (FULL || c < nch) {
          const float4 g = gpair[q];
          const float4 zv = *(reinterpret_cast<const float4*>(zbuf + (size_t)c * cap) + i);
          const float4 kf = coef[c];
          const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) f[q][e] = fmaf(kf.x, gg[e] * act_grad_t<KIND>(zz[e], alpha), -fmaf(zz[e], kf.z, kf.y));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) f[q][e] = 0.0f;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e][k] = pack_bf16(f[0][e], f[1][e]);
    }
    const int p = i * 4;
    int yl = p / a.wo, x = p - yl * a.wo;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      unsigned char* dst = outp + ((long long)(r0 + yl) * a.P + x) * 16;
      if (CH == 8) *reinterpret_cast<uint4*>(dst) = make_uint4(o[e][0], o[e][1], o[e][2 % (CH / 2)], o[e][3 % (CH / 2)]);
      else *reinterpret_cast<uint2*>(dst) = make_uint2(o[e][0], o[e][1]);
      if (++x == a.wo) {
        if (half == 0)
          for (int gx = 1; gx <= a.P - a.wo; ++gx) *reinterpret_cast<uint4*>(dst + gx * 16 - half * (2 * CH)) = zero4;   // gap column(s)
        x = 0; ++yl;
      }
    }
  }
  if (half == 0 && r1 == a.ho && r0 < a.ho) {     // the CTA that owns the last row also zero-fills the gap rows of the image
    const int q0r = a.ho * a.P, q1r = a.IMG;
    uint4* dst = reinterpret_cast<uint4*>(outp);
    for (int q = q0r + tid; q < q1r; q += kBwdThreads) dst[q] = zero4;
  }
  cluster_wait();
}

// ---------------------------------------------------------------------------------------------------------
// Small planes (hw <= 1024: the 32x32 .. 2x2 maps of KAN-VGG on CIFAR-sized inputs and the 28x28 / 14x14 tail of VGG16 @224):
// ONE WARP PER PLANE, the plane lives in registers (<= 8 float4 per lane), statistics by warp shuffles - no shared-memory
// passes, no block barriers.  A block-per-plane kernel spends its time in three latency-bound passes and two block reductions
// for a few hundred bytes (measured 0.6-2 TB/s).
// ---------------------------------------------------------------------------------------------------------
constexpr int kWarpPlaneMax4 = 8;        // float4 per lane (instantiated for 2 and 8: the 14x14 / 16x16 planes need two)

template <int VPL>
__global__ void __launch_bounds__(256, VPL <= 2 ? 8 : 4)
kc_instnorm_fwd_warp_kernel(const __grid_constant__ kc_norm_desc d, const float* __restrict__ z, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const float* __restrict__ alpha_p, float* __restrict__ y,
                            float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const long long plane = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (plane >= (long long)d.n * d.c) return;
  const int n = (int)(plane / d.c), ch = (int)(plane % d.c), n4 = d.hw >> 2;
  const long long off = (long long)n * d.batch_stride + (long long)ch * d.hw;
  const float4* z4 = reinterpret_cast<const float4*>(z + off);
  float4 v[VPL];
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < n4 ? __ldg(z4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mean = kc_warp_sum(s) / (float)d.hw;
  float m2 = 0.0f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (lane + 32 * k < n4) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
      m2 = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(e, e, m2))));
    }
  }
  const float rstd = rsqrtf(kc_warp_sum(m2) / (float)d.hw + d.eps);
  if (lane == 0) { mean_out[plane] = mean; rstd_out[plane] = rstd; }
  const float g = (d.affine && gamma) ? gamma[ch] : 1.0f;
  const float b = (d.affine && beta) ? beta[ch] : 0.0f;
  const float alpha = (d.out_act == KC_OUT_PRELU) ? alpha_p[0] : 0.0f;
  const float sc = rstd * g, sh0 = b - mean * rstd * g;
  const int kind = d.out_act;
  float4* y4 = reinterpret_cast<float4*>(y + off);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    if (i < n4) {
      float4 o;
      o.x = out_act(kind, fmaf(v[k].x, sc, sh0), alpha); o.y = out_act(kind, fmaf(v[k].y, sc, sh0), alpha);
      o.z = out_act(kind, fmaf(v[k].z, sc, sh0), alpha); o.w = out_act(kind, fmaf(v[k].w, sc, sh0), alpha);
      y4[i] = o;
    }
  }
}

// Backward + bf16 flat dz for small planes: one CTA per (image, 8-channel group), warp w = channel w.  dz of the eight
// channels meets in a shared-memory tile [position][8 x bf16] and leaves as coalesced 16-byte vectors (gap columns / rows
// zero-filled).  Same sums and the same partials layout as the cluster kernel.
template <int KIND, int VPL>
__global__ void __launch_bounds__(256, VPL <= 2 ? 4 : 2)
kc_norm_bwd_flat_warp_kernel(const __grid_constant__ NbfArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* tile = reinterpret_cast<uint16_t*>(smem_raw);        // [hw][8] bf16
  const kc_norm_desc& d = a.d;
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int n = blockIdx.x / a.groups8, g8 = blockIdx.x % a.groups8;
  const int ch = g8 * 8 + c, hw = a.ho * a.wo, n4 = hw >> 2;
  const bool live = ch < d.c;
  const float alpha = (KIND == KC_OUT_PRELU) ? a.alpha[0] : 0.0f;
  float4 zv[VPL], gv[VPL];
  float s_dvz = 0.0f, s_dv = 0.0f, s_da = 0.0f, rstd = 0.0f;
  if (live) {
    const long long off = (long long)n * d.batch_stride + (long long)ch * hw;
    const float4* z4 = reinterpret_cast<const float4*>(a.z + off);
    const float4* g4 = reinterpret_cast<const float4*>(a.dy + off);
    const float mean = a.mean[n * d.c + ch];
    rstd = a.rstd[n * d.c + ch];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + 32 * k;
      zv[k] = i < n4 ? __ldg(z4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      gv[k] = i < n4 ? __ldg(g4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (lane + 32 * k < n4) {
        float zz[4] = {(zv[k].x - mean) * rstd, (zv[k].y - mean) * rstd, (zv[k].z - mean) * rstd, (zv[k].w - mean) * rstd};
        float gg[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dv = gg[e] * act_grad_t<KIND>(zz[e], alpha);
          s_dvz = fmaf(dv, zz[e], s_dvz);
          s_dv += dv;
          if (KIND == KC_OUT_PRELU && !(zz[e] > 0.0f)) s_da = fmaf(gg[e], zz[e], s_da);
          gg[e] = dv;
        }
        zv[k] = make_float4(zz[0], zz[1], zz[2], zz[3]);       // zhat
        gv[k] = make_float4(gg[0], gg[1], gg[2], gg[3]);       // dv
      }
    }
    s_dvz = kc_warp_sum(s_dvz); s_dv = kc_warp_sum(s_dv);
    if (KIND == KC_OUT_PRELU) s_da = kc_warp_sum(s_da);
    if (lane == 0) {
      const long long NP = (long long)d.n * d.c, pl = (long long)n * d.c + ch;
      a.partials[pl] = s_dvz; a.partials[NP + pl] = s_dv; a.partials[2 * NP + pl] = s_da;
    }
  }
  const float m1 = s_dv / (float)hw, m2 = s_dvz / (float)hw;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    if (i < n4) {
      const float zz[4] = {zv[k].x, zv[k].y, zv[k].z, zv[k].w}, dv[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float dz = live ? rstd * (dv[e] - m1 - zz[e] * m2) : 0.0f;
        tile[(size_t)(4 * i + e) * 8 + c] = __bfloat16_as_ushort(__float2bfloat16_rn(dz));
      }
    }
  }
  __syncthreads();
  unsigned char* outp = a.dzf + ((long long)g8 * a.L + (long long)n * a.IMG) * 16;
  const uint4* t4 = reinterpret_cast<const uint4*>(tile);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int q = threadIdx.x; q < a.IMG; q += 256) {               // every flat position of the image, gaps included
    const int yy = q / a.P, x = q - yy * a.P;
    reinterpret_cast<uint4*>(outp)[q] = (yy < a.ho && x < a.wo) ? t4[yy * a.wo + x] : zero4;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Micro planes (hw <= 128: the 8x8 .. 2x2 maps at the tail of KAN-VGG on 32x32 inputs): a plane needs only LPP = 1 .. 32 lanes
// (one float4 each), so a warp carries 32 / LPP planes (consecutive images of one channel) and the reductions are segmented
// shuffles.  With a warp (or a block) per 2x2 plane the step of BASELINE config 3 spent 3 of 11 ms in launch-bound norm kernels.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float seg_sum(float v, int lpp) {
  for (int o = lpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
kc_instnorm_fwd_micro_kernel(const __grid_constant__ kc_norm_desc d, int lpp, const float* __restrict__ z, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const float* __restrict__ alpha_p, float* __restrict__ y,
                             float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  // plane index p = n * c + ch; a warp takes 32 / lpp consecutive planes
  const int lane = threadIdx.x & 31, per_warp = 32 / lpp, i = lane % lpp, n4 = d.hw >> 2;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long plane = warp * per_warp + lane / lpp;
  const bool live = plane < (long long)d.n * d.c && i < n4;
  const int n = live ? (int)(plane / d.c) : 0, ch = live ? (int)(plane % d.c) : 0;
  const long long off = (long long)n * d.batch_stride + (long long)ch * d.hw;
  const float4 v = live ? __ldg(reinterpret_cast<const float4*>(z + off) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float mean = seg_sum((v.x + v.y) + (v.z + v.w), lpp) / (float)d.hw;
  const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
  const float rstd = rsqrtf(seg_sum(live ? fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, a3 * a3))) : 0.0f, lpp) / (float)d.hw + d.eps);
  if (!live) return;
  if (i == 0) { mean_out[plane] = mean; rstd_out[plane] = rstd; }
  const float g = (d.affine && gamma) ? gamma[ch] : 1.0f;
  const float b = (d.affine && beta) ? beta[ch] : 0.0f;
  const float alpha = (d.out_act == KC_OUT_PRELU) ? alpha_p[0] : 0.0f;
  const float sc = rstd * g, sh0 = b - mean * rstd * g;
  const int kind = d.out_act;
  float4 o;
  o.x = out_act(kind, fmaf(v.x, sc, sh0), alpha); o.y = out_act(kind, fmaf(v.y, sc, sh0), alpha);
  o.z = out_act(kind, fmaf(v.z, sc, sh0), alpha); o.w = out_act(kind, fmaf(v.w, sc, sh0), alpha);
  reinterpret_cast<float4*>(y + off)[i] = o;
}

// Backward + flat dz for micro planes: CTA = 8 warps = the 8 channels of a group; warp c carries channel c of 32 / lpp
// consecutive images.  The dz tile [image][position][8 x bf16] is assembled in shared memory and written out as whole images
// of the flat buffer (gap columns / rows zero-filled).
template <int KIND>
__global__ void __launch_bounds__(256)
kc_norm_bwd_flat_micro_kernel(const __grid_constant__ NbfArgs a, int lpp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* tile = reinterpret_cast<uint16_t*>(smem_raw);        // [images per CTA][hw][8] bf16
  const kc_norm_desc& d = a.d;
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5, ipc = 32 / lpp, i = lane % lpp, il = lane / lpp;
  const int nblk = (d.n + ipc - 1) / ipc;                        // image blocks per channel group
  const int g8 = blockIdx.x / nblk, n0 = (blockIdx.x % nblk) * ipc, n = n0 + il;
  const int ch = g8 * 8 + c, hw = a.ho * a.wo, n4 = hw >> 2;
  const bool live = ch < d.c && n < d.n && i < n4;
  const float alpha = (KIND == KC_OUT_PRELU) ? a.alpha[0] : 0.0f;
  float zz[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
  float s_dvz = 0.0f, s_dv = 0.0f, s_da = 0.0f, rstd = 0.0f;
  if (live) {
    const long long off = (long long)n * d.batch_stride + (long long)ch * hw;
    const float4 zv = __ldg(reinterpret_cast<const float4*>(a.z + off) + i);
    const float4 gv = __ldg(reinterpret_cast<const float4*>(a.dy + off) + i);
    const float mean = a.mean[n * d.c + ch];
    rstd = a.rstd[n * d.c + ch];
    zz[0] = (zv.x - mean) * rstd; zz[1] = (zv.y - mean) * rstd; zz[2] = (zv.z - mean) * rstd; zz[3] = (zv.w - mean) * rstd;
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dv[e] = gg[e] * act_grad_t<KIND>(zz[e], alpha);
      s_dvz = fmaf(dv[e], zz[e], s_dvz);
      s_dv += dv[e];
      if (KIND == KC_OUT_PRELU && !(zz[e] > 0.0f)) s_da = fmaf(gg[e], zz[e], s_da);
    }
  }
  s_dvz = seg_sum(s_dvz, lpp); s_dv = seg_sum(s_dv, lpp); s_da = seg_sum(s_da, lpp);
  if (live && i == 0) {
    const long long NP = (long long)d.n * d.c, pl = (long long)n * d.c + ch;
    a.partials[pl] = s_dvz; a.partials[NP + pl] = s_dv; a.partials[2 * NP + pl] = s_da;
  }
  const float m1 = s_dv / (float)hw, m2 = s_dvz / (float)hw;
  if (i < n4 && n < d.n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float dz = live ? rstd * (dv[e] - m1 - zz[e] * m2) : 0.0f;
      tile[((size_t)il * hw + 4 * i + e) * 8 + c] = __bfloat16_as_ushort(__float2bfloat16_rn(dz));
    }
  }
  __syncthreads();
  const uint4* t4 = reinterpret_cast<const uint4*>(tile);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  const int nimg = min(ipc, d.n - n0);
  uint4* outp = reinterpret_cast<uint4*>(a.dzf + ((long long)g8 * a.L + (long long)n0 * a.IMG) * 16);
  for (int q = threadIdx.x; q < nimg * a.IMG; q += 256) {        // consecutive images are consecutive in the flat buffer
    const int im = q / a.IMG, r = q - im * a.IMG, yy = r / a.P, x = r - yy * a.P;
    outp[q] = (yy < a.ho && x < a.wo) ? t4[(size_t)im * hw + yy * a.wo + x] : zero4;
  }
}

int pick_cluster(int units, size_t bytes_per_unit, size_t want, size_t limit, int* per_cta) {
  // smallest cluster size cs in {1, 2, 4, 8} whose per-CTA share of `units` needs <= want bytes; else the smallest that fits limit
  for (int pass = 0; pass < 2; ++pass) {
    const size_t cap = pass == 0 ? want : limit;
    for (int cs = 1; cs <= 8; cs *= 2) {
      const int per = (units + cs - 1) / cs;
      if ((size_t)per * bytes_per_unit <= cap) { *per_cta = per; return cs; }
    }
  }
  return 0;
}

template <typename Kernel, typename... Args>
cudaError_t launch_cluster(Kernel kernel, unsigned grid, unsigned threads, size_t smem, int cs, cudaStream_t st, Args... args) {
  if (cs > 8) {                      // clusters of 16 are a non-portable size: opt in per kernel
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int T, int CH, int KIND>
cudaError_t launch_nbf(const NbfArgs& a, unsigned grid, size_t smem, int cs, bool full, cudaStream_t st) {
  cudaError_t e;
  if (full) {
    e = cudaFuncSetAttribute(kc_norm_bwd_flat_cluster_kernel<T, CH, KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_cluster(kc_norm_bwd_flat_cluster_kernel<T, CH, KIND, true>, grid, T, smem, cs, st, a);
  }
  e = cudaFuncSetAttribute(kc_norm_bwd_flat_cluster_kernel<T, CH, KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return launch_cluster(kc_norm_bwd_flat_cluster_kernel<T, CH, KIND, false>, grid, T, smem, cs, st, a);
}

template <int T, int CH>
cudaError_t launch_nbf_kind(const NbfArgs& a, unsigned grid, size_t smem, int cs, bool full, cudaStream_t st) {
  if (a.d.out_act == KC_OUT_PRELU) return launch_nbf<T, CH, KC_OUT_PRELU>(a, grid, smem, cs, full, st);
  if (a.d.out_act == KC_OUT_SILU) return launch_nbf<T, CH, KC_OUT_SILU>(a, grid, smem, cs, full, st);
  return launch_nbf<T, CH, KC_OUT_NONE>(a, grid, smem, cs, full, st);
}

// Launch shape of the fused backward: cluster size, rows per CTA, threads.  The resident chunk (8 channels x rows x wo floats) is
// kept <= 52 KB when a cluster of <= 8 CTAs allows it (four 256-thread CTAs per SM; 64-thread CTAs for the 14x14 planes), else
// <= 200 KB with one 1024-thread CTA per SM (the 224x224 planes).  A thread owns at most two float4 columns per channel.
// (Splitting the 8 channels over two CTAs - 100 KB chunks, two CTAs per SM, 8-byte half-vector stores - measured 6 % slower.)
// Planes that do not fit 52 KB chunks in a cluster of 8 (224x224): a cluster of 16 (non-portable size) with 100 KB chunks and
// 512-thread CTAs puts TWO CTAs on an SM, so the phases of one (bulk load, sums, cluster exchange, apply) overlap the other's
// instead of running back to back on an otherwise idle SM; KANCONV_NORM_BWD_CS16=0 keeps the 8 x 200 KB / 1024-thread shape.
struct NbfPlan { int ch, cs, rows, threads; size_t smem; };
bool plan_nbf(int ho, int wo, NbfPlan* pl) {
  static const int cs16 = []() { const char* e = getenv("KANCONV_NORM_BWD_CS16"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
  for (int pass = 0; pass < 3; ++pass) {
    if (pass == 1 && !cs16) continue;
    for (int cs = pass == 1 ? 16 : 1; cs <= (pass == 1 ? 16 : 8); cs *= 2) {
      const int rows = (ho + cs - 1) / cs;
      const size_t smem = (size_t)rows * wo * 32;
      if (smem > (pass == 0 ? (size_t)52 * 1024 : pass == 1 ? (size_t)101 * 1024 : kSmemMax)) continue;
      const int threads = pass == 2 ? kBwdThreadsBig : pass == 1 ? kBwdThreadsMid : rows * wo <= 4 * kBwdThreadsTiny ? kBwdThreadsTiny : kBwdThreadsSmall;
      if (rows * wo > 8 * threads) continue;
      if (pass == 1 && (rows * cs != ho || ((rows * wo) & 3))) continue;      // equal, 16-byte aligned chunks only
      *pl = {8, cs, rows, threads, smem};
      return true;
    }
  }
  return false;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Forward through the shared-memory-resident kernel when the shape allows it.  Returns KC_ERR_UNSUPPORTED (without touching
// the error string semantics of the caller) when it does not: kc_norm_act_fwd then runs the generic two-pass kernel.
int kc_instnorm_fwd_cluster(const kc_norm_desc* d, const float* z, const float* gamma, const float* beta, const float* alpha,
                            float* y, float* mean, float* rstd, void* stream) {
  const char* env = getenv("KANCONV_NORM_CLUSTER");          // debug switch: 0 = always the generic two-pass kernel
  if (env != nullptr && env[0] == '0') return KC_ERR_UNSUPPORTED;
  if (d->norm == KC_NORM_INSTANCE && d->hw <= 128 * kWarpPlaneMax4 && (d->hw & 3) == 0 && (d->batch_stride & 3) == 0 && aligned16(z) &&
      aligned16(y)) {              // small planes: one warp per plane
    const long long planes = (long long)d->n * d->c;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->hw <= 128) {            // micro planes: several planes per warp
      int lpp = 1;
      while (lpp * 4 < d->hw) lpp *= 2;
      const long long warps = (planes + 32 / lpp - 1) / (32 / lpp);
      kc_instnorm_fwd_micro_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(*d, lpp, z, gamma, beta, alpha, y, mean, rstd);
    } else if (d->hw <= 256) {
      kc_instnorm_fwd_warp_kernel<2><<<(unsigned)((planes + 7) / 8), 256, 0, st>>>(*d, z, gamma, beta, alpha, y, mean, rstd);
    } else {
      kc_instnorm_fwd_warp_kernel<8><<<(unsigned)((planes + 7) / 8), 256, 0, st>>>(*d, z, gamma, beta, alpha, y, mean, rstd);
    }
    KC_LAUNCH_CHECK("kc_instnorm_fwd_warp_kernel");
    return KC_OK;
  }
  if (d->norm != KC_NORM_INSTANCE || d->hw < 3136 || (d->hw & 3) || (d->batch_stride & 3) || !aligned16(z) || !aligned16(y))
    return KC_ERR_UNSUPPORTED;
  int chunk4 = 0;
  const int cs = pick_cluster(d->hw / 4, 16, 52 * 1024, kSmemMax, &chunk4);
  if (cs == 0) return KC_ERR_UNSUPPORTED;
  const int chunk = chunk4 * 4;
  const size_t smem = (size_t)chunk * 4;
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_instnorm_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = (long long)d->n * d->c * cs;
  if (grid > 0x7fffffffLL) return KC_ERR_UNSUPPORTED;
  cudaError_t e = launch_cluster(kc_instnorm_fwd_cluster_kernel, (unsigned)grid, kFwdThreads, smem, cs, (cudaStream_t)stream, *d, chunk,
                                 z, gamma, beta, alpha, y, mean, rstd);
  kc_count_launch();
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "launch of kc_instnorm_fwd_cluster_kernel failed: %s", cudaGetErrorString(e));
  return KC_OK;
}

extern "C" int kc_norm_bwd_dz_flat_supported(const kc_desc* conv, const kc_norm_desc* d) {
  if (!conv || !d) return 0;
  if (d->norm != KC_NORM_INSTANCE || d->affine) return 0;
  if (conv->stride_h != 1 || conv->stride_w != 1) return 0;
  if (d->c != conv->cout || d->n != conv->n || d->hw != conv->ho * conv->wo) return 0;
  int P = 0, IMG = 0, cq = 0;
  long long L = 0;
  if (kc_tc_flat_layout(conv, &P, &IMG, &L, &cq) != KC_OK) return 0;
  if (conv->wo > P || conv->ho * P > IMG) return 0;
  const int hw = d->hw;
  if ((hw & 3) || (d->batch_stride & 3)) return 0;
  NbfPlan pl;
  if (!plan_nbf(conv->ho, conv->wo, &pl)) return 0;
  if (pl.cs > 1 && ((pl.rows * conv->wo) & 3)) return 0;    // every CTA's chunk must start 16-byte aligned
  return 1;
}

extern "C" int kc_norm_bwd_dz_flat(const kc_desc* conv, const kc_norm_desc* d, const float* dy, const float* z, const float* mean,
                                   const float* rstd, const float* alpha, void* dz_flat, float* dalpha, float* partials,
                                   void* stream) {
  if (!kc_norm_bwd_dz_flat_supported(conv, d)) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_norm_bwd_dz_flat: shape / norm kind not covered by the fused kernel");
  if (!dy || !z || !mean || !rstd || !dz_flat || !partials) KC_FAIL(KC_ERR_INVALID, "kc_norm_bwd_dz_flat: null pointer");
  if (d->out_act == KC_OUT_PRELU && !alpha) KC_FAIL(KC_ERR_INVALID, "kc_norm_bwd_dz_flat: PReLU needs alpha");
  if (!aligned16(dy) || !aligned16(z) || !aligned16(dz_flat)) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_norm_bwd_dz_flat: pointers must be 16-byte aligned");
  NbfArgs a;
  a.d = *d; a.ho = conv->ho; a.wo = conv->wo;
  int cq = 0;
  int rc = kc_tc_flat_layout(conv, &a.P, &a.IMG, &a.L, &cq);
  if (rc != KC_OK) return rc;
  a.groups8 = cq / 8;
  NbfPlan pl;
  plan_nbf(conv->ho, conv->wo, &pl);
  a.rows = pl.rows;
  a.dy = dy; a.z = z; a.mean = mean; a.rstd = rstd; a.alpha = alpha; a.dzf = (unsigned char*)dz_flat; a.partials = partials;
  if (d->hw <= 128 * kWarpPlaneMax4) {       // small planes: one warp per channel plane, eight channels per CTA
    cudaStream_t st = (cudaStream_t)stream;
    if (d->hw <= 128) {            // micro planes: 32 / lpp images per warp
      int lpp = 1;
      while (lpp * 4 < d->hw) lpp *= 2;
      const int ipc = 32 / lpp;
      const size_t tile = (size_t)ipc * d->hw * 16;
      const unsigned blocks = (unsigned)((long long)((d->n + ipc - 1) / ipc) * a.groups8);
      if (d->out_act == KC_OUT_PRELU) kc_norm_bwd_flat_micro_kernel<KC_OUT_PRELU><<<blocks, 256, tile, st>>>(a, lpp);
      else if (d->out_act == KC_OUT_SILU) kc_norm_bwd_flat_micro_kernel<KC_OUT_SILU><<<blocks, 256, tile, st>>>(a, lpp);
      else kc_norm_bwd_flat_micro_kernel<KC_OUT_NONE><<<blocks, 256, tile, st>>>(a, lpp);
    } else {
      const size_t tile = (size_t)d->hw * 16;
      const unsigned blocks = (unsigned)((long long)d->n * a.groups8);
      const bool v2 = d->hw <= 256;
      if (d->out_act == KC_OUT_PRELU) { if (v2) kc_norm_bwd_flat_warp_kernel<KC_OUT_PRELU, 2><<<blocks, 256, tile, st>>>(a); else kc_norm_bwd_flat_warp_kernel<KC_OUT_PRELU, 8><<<blocks, 256, tile, st>>>(a); }
      else if (d->out_act == KC_OUT_SILU) { if (v2) kc_norm_bwd_flat_warp_kernel<KC_OUT_SILU, 2><<<blocks, 256, tile, st>>>(a); else kc_norm_bwd_flat_warp_kernel<KC_OUT_SILU, 8><<<blocks, 256, tile, st>>>(a); }
      else { if (v2) kc_norm_bwd_flat_warp_kernel<KC_OUT_NONE, 2><<<blocks, 256, tile, st>>>(a); else kc_norm_bwd_flat_warp_kernel<KC_OUT_NONE, 8><<<blocks, 256, tile, st>>>(a); }
    }
    KC_LAUNCH_CHECK("kc_norm_bwd_flat_warp_kernel");
    if (dalpha != nullptr) return kc_norm_partials_to_params(d, partials, nullptr, nullptr, dalpha, stream);
    return KC_OK;
  }
  const long long grid = (long long)d->n * a.groups8 * (8 / pl.ch) * pl.cs;
  if (grid > 0x7fffffffLL) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_norm_bwd_dz_flat: grid too large");
  const bool full = (d->c % 8) == 0 && a.groups8 * 8 == d->c;          // every 8-channel group of the flat buffer is complete
  cudaError_t e = pl.threads == kBwdThreadsBig ? launch_nbf_kind<kBwdThreadsBig, 8>(a, (unsigned)grid, pl.smem, pl.cs, full, (cudaStream_t)stream)
                  : pl.threads == kBwdThreadsMid ? launch_nbf_kind<kBwdThreadsMid, 8>(a, (unsigned)grid, pl.smem, pl.cs, full, (cudaStream_t)stream)
                  : pl.threads == kBwdThreadsTiny ? launch_nbf_kind<kBwdThreadsTiny, 8>(a, (unsigned)grid, pl.smem, pl.cs, full, (cudaStream_t)stream)
                                                  : launch_nbf_kind<kBwdThreadsSmall, 8>(a, (unsigned)grid, pl.smem, pl.cs, full, (cudaStream_t)stream);
  kc_count_launch();
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "launch of kc_norm_bwd_flat_cluster_kernel failed: %s", cudaGetErrorString(e));
  if (dalpha != nullptr) return kc_norm_partials_to_params(d, partials, nullptr, nullptr, dalpha, stream);
  return KC_OK;
}
