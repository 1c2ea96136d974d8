// kc_debug.cu - micro-benchmarks behind the measurements quoted in DESIGN.md section 4 (tcgen05.mma issue rate for the
// no-swizzle layouts and tile widths used by the convolution kernels, cp.async.bulk latency / request rate).
// NOT part of the product library: this file is compiled and its kc_debug_* symbols exist only when the library is built
// with KANCONV_DEBUG=1 (python -m kanconv_b200.build); include/kanconv.h does not declare them.  tools/mma_rate.py and
// tools/bulk_bench.py are the callers.
#ifdef KANCONV_DEBUG
#include "kc_common.cuh"
#include "kc_umma.cuh"

using namespace kc;

// Debug only: raw tcgen05.mma rate from resident smem operands, with knobs that mimic the convolution main loop:
// nsub accumulators used round-robin, a commit every `commit_every` MMAs (0 = only at the end), and `writers` extra
// warps streaming 16-byte st.shared into an unrelated smem region while the MMAs run.
__global__ void __launch_bounds__(576, 1) kc_mma_rate_kernel(int N, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int iters,
                                                              int a_rowshift, int nsub, int commit_every, int writers,
                                                              int mn_major, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar, dummy[8];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1000000); done = 0; fence_barrier_init(); }
  if (warp == 17) tmem_alloc(&tmem_ptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (tid == 17 * 32) {
    const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
    const uint32_t abase = smem_u32(sm) + a_rowshift * 16, bbase = smem_u32(sm) + 96 * 1024;
    long long t0 = clock64();
    int cnt = 0;
    for (int it = 0; it < iters; ++it) {
      for (int s = 0; s < nsub; ++s) {
        for (int i = 0; i < 2; ++i) {
          uint64_t ad = make_smem_desc(abase + (mn_major ? s * 16 + i * 256 : s * 2048 + i * 2 * a_lbo), a_lbo, a_sbo);
          uint64_t bd = make_smem_desc(bbase + (mn_major ? i * 256 : i * 2 * b_lbo), b_lbo, b_sbo);
          tc_mma_bf16(tb + s * N, ad, bd, idesc, 1u);
          if (commit_every > 0 && (++cnt % commit_every) == 0) tc_commit(&dummy[(cnt / commit_every) & 7]);
        }
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = (float)(t1 - t0) / (float)(iters * nsub * 2);
    done = 1;
  } else if (warp < writers) {
    uint4* dst = reinterpret_cast<uint4*>(sm + 128 * 1024) + tid;
    uint4 v = make_uint4(tid, 1, 2, 3);
    while (!done) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k * 512 % 1024] = v;
      v.x += 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tb, 512);
}

extern "C" int kc_debug_mma_rate(int N, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int iters, int a_rowshift, int nsub,
                                 int commit_every, int writers, int mn_major, float* cycles) {
  float* dev = nullptr;
  KC_CUDA_CHECK(cudaMalloc(&dev, 32 * sizeof(float)));
  KC_CUDA_CHECK(cudaMemset(dev, 0, 32 * sizeof(float)));
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  kc_mma_rate_kernel<<<1, 576, 160 * 1024>>>(N, a_lbo, a_sbo, b_lbo, b_sbo, iters, a_rowshift, nsub, commit_every, writers, mn_major, dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles, dev, 32 * sizeof(float), cudaMemcpyDeviceToHost);     // caller passes float[32]
  cudaFree(dev);
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "kc_debug_mma_rate: %s", cudaGetErrorString(e));
  return KC_OK;
}

// Debug only: tcgen05.mma execution rate with the warp-uniform elect issue path, mimicking one ring step of the conv
// kernels: nsub accumulators x 2 k-steps per iteration, optional commit per iteration, optional writer warps hammering
// shared memory with 16-byte stores, optional unaligned A view.
__global__ void __launch_bounds__(576, 1) kc_mma_rate2_kernel(int N, int mn_major, int iters, int nsub, int commit_each, int writers,
                                                               int a_shift_rows, float* out, const unsigned char* bulk_src, int bulk_streams) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar, dummy[8], bbar[4];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1000000); for (int i = 0; i < 4; ++i) mbar_init(&bbar[i], 1); done = 0; fence_barrier_init(); }
  if (warp == 17) tmem_alloc(&tmem_ptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 17) {
    const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
    const uint32_t au = (smem_u32(sm) >> 4) + (uint32_t)a_shift_rows, bu = (smem_u32(sm) + 100 * 1024) >> 4;
    const uint32_t a_pitch = mn_major ? 1168u : 15456u, b_pitch = mn_major ? 1040u : (uint32_t)N * 16u;
    const uint32_t a_lo_c = (mn_major ? 8u : (a_pitch >> 4)) << 16, b_lo_c = (mn_major ? 8u : (b_pitch >> 4)) << 16;
    const uint32_t a_hi = (mn_major ? (a_pitch >> 4) : 8u) | (1u << 14), b_hi = (mn_major ? (b_pitch >> 4) : 8u) | (1u << 14);
    const uint32_t a_step = mn_major ? 16u : (2u * a_pitch) >> 4, b_step = mn_major ? 16u : (2u * b_pitch) >> 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        for (int s = 0; s < nsub; ++s) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_c | (au + s * 128 + ks * a_step));
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo_c | (bu + ks * b_step));
            tc_mma_bf16(tb + s * N, ad, bd, idesc, 1u);
          }
        }
        if (commit_each) tc_commit(&dummy[it & 7]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 17 * 32) { out[0] = (float)(t1 - t0) / (float)(iters * nsub * 2); done = 1; }
  } else if (warp == 16 && bulk_streams > 0 && bulk_src != nullptr) {
    // bulk_streams (<= 4) cp.async.bulk copies of 4 KB each kept in flight into an unrelated shared-memory region while the
    // MMAs run (the weight loader of the convolution kernels); out[1] = bytes per clock achieved
    if (threadIdx.x == 16 * 32) {
      uint32_t ph[4] = {0, 0, 0, 0};
      long long n = 0, t0 = clock64();
      for (int i = 0; i < bulk_streams; ++i) { mbar_arrive_expect_tx(&bbar[i], 4096u); bulk_g2s(sm + 144 * 1024 + i * 4096, bulk_src + i * 4096, 4096u, &bbar[i]); }
      while (!done) {
        for (int i = 0; i < bulk_streams; ++i) {
          mbar_wait(&bbar[i], ph[i]); ph[i] ^= 1u; ++n;
          mbar_arrive_expect_tx(&bbar[i], 4096u);
          bulk_g2s(sm + 144 * 1024 + i * 4096, bulk_src + ((n * 4096) & 0xffffff), 4096u, &bbar[i]);
        }
      }
      for (int i = 0; i < bulk_streams; ++i) mbar_wait(&bbar[i], ph[i]);
      out[1] = (float)(n * 4096) / (float)(clock64() - t0);
    }
  } else if (warp < writers) {
    uint4* dst = reinterpret_cast<uint4*>(sm + 128 * 1024) + threadIdx.x;
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    while (!done) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[(k * 512) & 1023] = v;
      v.x += 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tb, 512);
}

// Debug only: the same measurement for a CTA PAIR (tcgen05.mma.cta_group::2, M = 256: 128 rows of A per CTA, the N rows of B
// split between the two CTAs).  The leader issues; out[0] = cycles per MMA, out[1] / out[2] = accumulator element (lane 0,
// column 0) of the leader / the peer (all-ones operands: must equal the accumulated K), out[3] / out[4] = TMEM base addresses.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    kc_mma_rate_2cta_kernel(int N, int mn_major, int iters, int nsub, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3f803f80u;   // bf16 1.0
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_arrive();
  cluster_wait();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(256, N, mn_major, mn_major);
    const uint32_t au = smem_u32(sm) >> 4, bu = (smem_u32(sm) + 100 * 1024) >> 4;
    const uint32_t nh = (uint32_t)N / 2;                       // B rows held by each CTA
    const uint32_t a_pitch = mn_major ? 1168u : 15456u, b_pitch = mn_major ? 1040u : nh * 16u;
    const uint32_t a_lo_c = (mn_major ? 8u : (a_pitch >> 4)) << 16, b_lo_c = (mn_major ? 8u : (b_pitch >> 4)) << 16;
    const uint32_t a_hi = (mn_major ? (a_pitch >> 4) : 8u) | (1u << 14), b_hi = (mn_major ? (b_pitch >> 4) : 8u) | (1u << 14);
    const uint32_t a_step = mn_major ? 16u : (2u * a_pitch) >> 4, b_step = mn_major ? 16u : (2u * b_pitch) >> 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        for (int s = 0; s < nsub; ++s) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_c | (au + s * 128 + ks * a_step));
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo_c | (bu + ks * b_step));
            const uint32_t acc = it > 0 || ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb + s * N), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                         : "memory");
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync())
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                   "h"((uint16_t)3)
                   : "memory");
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (float)(t1 - t0) / (float)(iters * nsub * 2);
  }
  if (warp == 0) {
    if (rank != 0) mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[8];
    tmem_ld8(tb, r);
    tmem_ld_wait();
    if (threadIdx.x == 0) { out[1 + rank] = __uint_as_float(r[0]); out[3 + rank] = (float)tb; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();
  cluster_wait();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

extern "C" int kc_debug_mma_rate_2cta(int N, int mn_major, int iters, int nsub, float* host_out5) {
  float* dev = nullptr;
  KC_CUDA_CHECK(cudaMalloc(&dev, 5 * sizeof(float)));
  KC_CUDA_CHECK(cudaMemset(dev, 0, 5 * sizeof(float)));
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_mma_rate_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  kc_mma_rate_2cta_kernel<<<2, 128, 160 * 1024>>>(N, mn_major, iters, nsub, dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(host_out5, dev, 5 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "kc_debug_mma_rate_2cta: %s", cudaGetErrorString(e));
  return KC_OK;
}

extern "C" int kc_debug_mma_rate2_bulk(int N, int mn_major, int iters, int nsub, int commit_each, int writers, int a_shift_rows,
                                       int bulk_streams, float* cycles2) {
  float* dev = nullptr;
  unsigned char* src = nullptr;
  KC_CUDA_CHECK(cudaMalloc(&dev, 2 * sizeof(float)));
  KC_CUDA_CHECK(cudaMalloc(&src, (1 << 24) + 65536));
  KC_CUDA_CHECK(cudaMemset(src, 0x3c, (1 << 24) + 65536));
  KC_CUDA_CHECK(cudaMemset(dev, 0, 2 * sizeof(float)));
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  kc_mma_rate2_kernel<<<1, 576, 160 * 1024>>>(N, mn_major, iters, nsub, commit_each, writers, a_shift_rows, dev, src, bulk_streams);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles2, dev, 2 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(dev); cudaFree(src);
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "kc_debug_mma_rate2_bulk: %s", cudaGetErrorString(e));
  return KC_OK;
}

extern "C" int kc_debug_mma_rate2(int N, int mn_major, int iters, int nsub, int commit_each, int writers, int a_shift_rows, float* cycles) {
  float* dev = nullptr;
  KC_CUDA_CHECK(cudaMalloc(&dev, sizeof(float)));
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  kc_mma_rate2_kernel<<<1, 576, 160 * 1024>>>(N, mn_major, iters, nsub, commit_each, writers, a_shift_rows, dev, nullptr, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles, dev, sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "kc_debug_mma_rate2: %s", cudaGetErrorString(e));
  return KC_OK;
}

// Debug only: does the A-operand collector hint relieve the shared-memory bound of narrow MMAs?  Pattern of the weight-
// gradient kernel: per k-step one A tile feeds three MMAs (three taps = three accumulators, B read from shifted rows).
__global__ void __launch_bounds__(576, 1) kc_mma_rate3_kernel(int N, int mn_major, int iters, int reuse, int writers, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); done = 0; fence_barrier_init(); }
  if (warp == 17) tmem_alloc(&tmem_ptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 17) {
    const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
    const uint32_t au = smem_u32(sm) >> 4, bu = (smem_u32(sm) + 100 * 1024) >> 4;
    const uint32_t a_pitch = mn_major ? 1168u : 15456u, b_pitch = mn_major ? 1168u : (uint32_t)N * 16u + 64u;
    const uint32_t a_lo_c = (mn_major ? 8u : (a_pitch >> 4)) << 16, b_lo_c = (mn_major ? 8u : (b_pitch >> 4)) << 16;
    const uint32_t a_hi = (mn_major ? (a_pitch >> 4) : 8u) | (1u << 14), b_hi = (mn_major ? (b_pitch >> 4) : 8u) | (1u << 14);
    const uint32_t a_step = mn_major ? 16u : (2u * a_pitch) >> 4, b_step = mn_major ? 16u : (2u * b_pitch) >> 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_c | (au + (ks & 1) * a_step + (ks >> 1) * 3u));
          const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo_c | (bu + (ks & 1) * b_step));
          if (reuse) {
            tc_mma_bf16_keep<1>(tb, ad, bd, idesc, 1u);
            tc_mma_bf16_keep<2>(tb + N, ad, bd + 1u, idesc, 1u);
            tc_mma_bf16_keep<3>(tb + 2 * N, ad, bd + 2u, idesc, 1u);
          } else {
            tc_mma_bf16(tb, ad, bd, idesc, 1u);
            tc_mma_bf16(tb + N, ad, bd + 1u, idesc, 1u);
            tc_mma_bf16(tb + 2 * N, ad, bd + 2u, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 17 * 32) { out[0] = (float)(t1 - t0) / (float)(iters * 12); done = 1; }
  } else if (warp < writers) {
    uint4* dst = reinterpret_cast<uint4*>(sm + 128 * 1024) + threadIdx.x;
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    while (!done) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[(k * 512) & 1023] = v;
      v.x += 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tb, 512);
}

extern "C" int kc_debug_mma_rate3(int N, int mn_major, int iters, int reuse, int writers, float* cycles) {
  float* dev = nullptr;
  KC_CUDA_CHECK(cudaMalloc(&dev, sizeof(float)));
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_mma_rate3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  kc_mma_rate3_kernel<<<1, 576, 160 * 1024>>>(N, mn_major, iters, reuse, writers, dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles, dev, sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "kc_debug_mma_rate3: %s", cudaGetErrorString(e));
  return KC_OK;
}

// Debug only: latency / throughput of cp.async.bulk global->shared.  Each CTA issues `depth` copies of `bytes` back to back
// (ring of `depth` buffers), `iters` rounds; same_addr != 0 makes every CTA read the same global range.
__global__ void __launch_bounds__(32, 1) kc_bulk_bench_kernel(const unsigned char* src, int bytes, int depth, int iters,
                                                               int same_addr, long long span, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned char* base = src + (same_addr ? 0 : ((long long)blockIdx.x * (long long)bytes * depth) % span);
    long long off = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < depth; ++k) {
        mbar_arrive_expect_tx(&bars[k], (uint32_t)bytes);
        bulk_g2s(sm + (size_t)k * bytes, base + off, (uint32_t)bytes, &bars[k]);
        off = (off + bytes) % (span / 2);
      }
      for (int k = 0; k < depth; ++k) mbar_wait(&bars[k], it & 1);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (float)(t1 - t0) / (float)iters;
  }
}

extern "C" int kc_debug_bulk_bench(const void* src, long long span, int bytes, int depth, int iters, int nctas, int same_addr,
                                   float* cycles_per_round) {
  float* dev = nullptr;
  KC_CUDA_CHECK(cudaMalloc(&dev, sizeof(float)));
  KC_CUDA_CHECK(cudaFuncSetAttribute(kc_bulk_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes * depth));
  kc_bulk_bench_kernel<<<nctas, 32, (size_t)bytes * depth>>>((const unsigned char*)src, bytes, depth, iters, same_addr, span, dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles_per_round, dev, sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "kc_debug_bulk_bench: %s", cudaGetErrorString(e));
  return KC_OK;
}
#endif  // KANCONV_DEBUG
