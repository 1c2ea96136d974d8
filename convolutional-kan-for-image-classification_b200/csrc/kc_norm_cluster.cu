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
#include "kc_common.cuh"
#include "kc_norm_common.cuh"
#include "kc_umma.cuh"

int kc_tc_flat_layout(const kc_desc* d, int* P, int* IMG, long long* L, int* cq);      // kc_tc.cu

namespace {

using namespace kc;

constexpr int kFwdThreads = 256;
constexpr int kBwdThreads = 512;
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

__global__ void __launch_bounds__(kBwdThreads)
kc_norm_bwd_flat_cluster_kernel(const __grid_constant__ NbfArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* zbuf = reinterpret_cast<float*>(smem_raw);           // [8][cap] zhat of this CTA's rows
  __shared__ float red[kBwdThreads / 32][24];
  __shared__ float part[24];         // this CTA's sums, read by the peers through DSMEM
  __shared__ float tot[24];          // cluster totals
  __shared__ __align__(8) uint64_t bar;
  const kc_norm_desc& d = a.d;
  const uint32_t cs = cluster_nctarank(), rank = cluster_ctarank();
  const int cl = blockIdx.x / cs, n = cl / a.groups8, g8 = cl % a.groups8;
  const int c0 = g8 * 8, nch = max(0, min(8, d.c - c0));
  const int r0 = (int)rank * a.rows, r1 = min(a.ho, r0 + a.rows);
  const int cnt = max(0, r1 - r0) * a.wo, cap = a.rows * a.wo, hw = a.ho * a.wo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)n * d.batch_stride + (long long)c0 * hw + (long long)r0 * a.wo;      // channel c: + c * hw
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (tid == 0 && cnt > 0 && nch > 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)(cnt * 4 * nch));
    for (int c = 0; c < nch; ++c) bulk_load_chunk(zbuf + (size_t)c * cap, a.z + base + (long long)c * hw, (uint32_t)cnt * 4u, &bar);
  }
  const int kind = d.out_act;
  const float alpha = (kind == KC_OUT_PRELU) ? a.alpha[0] : 0.0f;
  float mean[8], rstd[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const bool ok = c < nch;
    mean[c] = ok ? a.mean[n * d.c + c0 + c] : 0.0f;
    rstd[c] = ok ? a.rstd[n * d.c + c0 + c] : 0.0f;
  }
  if (cnt > 0 && nch > 0) mbar_wait(&bar, 0);
  // ---- phase 1: sums over this CTA's rows; zhat replaces z in shared memory -------------------------------------
  float acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = 0.0f;
  const int n4 = cnt >> 2;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c < nch) {
      float4* zb = reinterpret_cast<float4*>(zbuf + (size_t)c * cap);
      const float4* g4 = reinterpret_cast<const float4*>(a.dy + base + (long long)c * hw);
      for (int i = tid; i < n4; i += kBwdThreads) {
        float4 zv = zb[i];
        const float4 gv = __ldg(g4 + i);
        zv.x = (zv.x - mean[c]) * rstd[c]; zv.y = (zv.y - mean[c]) * rstd[c];
        zv.z = (zv.z - mean[c]) * rstd[c]; zv.w = (zv.w - mean[c]) * rstd[c];
        zb[i] = zv;
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dv = gg[e] * out_act_grad(kind, zz[e], alpha);
          acc[c] = fmaf(dv, zz[e], acc[c]);
          acc[8 + c] += dv;
          if (kind == KC_OUT_PRELU && !(zz[e] > 0.0f)) acc[16 + c] = fmaf(gg[e], zz[e], acc[16 + c]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const float v = kc_warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (tid < 24) {
    float v = 0.0f;
    for (int w = 0; w < kBwdThreads / 32; ++w) v += red[w][tid];
    part[tid] = v;
  }
  __syncthreads();
  cluster_arrive();
  cluster_wait();
  if (tid < 24) {
    float v = 0.0f;
    for (uint32_t r = 0; r < cs; ++r) v += dsmem_ld(&part[tid], r);       // rank order: deterministic
    tot[tid] = v;
    const int c = tid & 7, which = tid >> 3;
    if (rank == 0 && c < nch) a.partials[(long long)which * d.n * d.c + n * d.c + c0 + c] = v;
  }
  __syncthreads();
  cluster_arrive();                  // done with the peers' shared memory
  // ---- phase 2: dz of 8 channels per position -> one 16-byte bf16 vector of the flat buffer -----------------------
  float m1[8], m2[8];
  const float inv_hw = 1.0f / (float)hw;
#pragma unroll
  for (int c = 0; c < 8; ++c) { m2[c] = tot[c] * inv_hw; m1[c] = tot[8 + c] * inv_hw; }
  unsigned char* outp = a.dzf + ((long long)g8 * a.L + (long long)n * a.IMG) * 16;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  const float* gbase = a.dy + base;
  for (int p = tid; p < cnt; p += kBwdThreads) {
    const int yl = p / a.wo, x = p - yl * a.wo;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v = 0.0f;
      if (c < nch) {
        const float zh = zbuf[(size_t)c * cap + p];
        const float dv = __ldg(gbase + (long long)c * hw + p) * out_act_grad(kind, zh, alpha);
        v = rstd[c] * (dv - m1[c] - zh * m2[c]);
      }
      f[c] = v;
    }
    uint4* dst = reinterpret_cast<uint4*>(outp + ((long long)(r0 + yl) * a.P + x) * 16);
    *dst = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    if (x == a.wo - 1)
      for (int gx = 1; gx <= a.P - a.wo; ++gx) dst[gx] = zero4;             // gap column(s) of this row
  }
  if (r1 == a.ho && r0 < a.ho) {     // the CTA that owns the last row also zero-fills the gap rows of the image
    const int q0 = a.ho * a.P, q1 = a.IMG;
    uint4* dst = reinterpret_cast<uint4*>(outp);
    for (int q = q0 + tid; q < q1; q += kBwdThreads) dst[q] = zero4;
  }
  cluster_wait();
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

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Forward through the shared-memory-resident kernel when the shape allows it.  Returns KC_ERR_UNSUPPORTED (without touching
// the error string semantics of the caller) when it does not: kc_norm_act_fwd then runs the generic two-pass kernel.
int kc_instnorm_fwd_cluster(const kc_norm_desc* d, const float* z, const float* gamma, const float* beta, const float* alpha,
                            float* y, float* mean, float* rstd, void* stream) {
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
  int rows = 0;
  const int cs = pick_cluster(conv->ho, (size_t)conv->wo * 32, 52 * 1024, kSmemMax, &rows);
  if (cs == 0) return 0;
  if (cs > 1 && ((rows * conv->wo) & 3)) return 0;          // every CTA's chunk must start 16-byte aligned
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
  const int cs = pick_cluster(conv->ho, (size_t)conv->wo * 32, 52 * 1024, kSmemMax, &a.rows);
  a.dy = dy; a.z = z; a.mean = mean; a.rstd = rstd; a.alpha = alpha; a.dzf = (unsigned char*)dz_flat; a.partials = partials;
  const size_t smem = (size_t)a.rows * a.wo * 32;
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_norm_bwd_flat_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = (long long)d->n * a.groups8 * cs;
  if (grid > 0x7fffffffLL) KC_FAIL(KC_ERR_UNSUPPORTED, "kc_norm_bwd_dz_flat: grid too large");
  cudaError_t e = launch_cluster(kc_norm_bwd_flat_cluster_kernel, (unsigned)grid, kBwdThreads, smem, cs, (cudaStream_t)stream, a);
  kc_count_launch();
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "launch of kc_norm_bwd_flat_cluster_kernel failed: %s", cudaGetErrorString(e));
  if (dalpha != nullptr) return kc_norm_partials_to_params(d, partials, nullptr, nullptr, dalpha, stream);
  return KC_OK;
}
