// kc_norm_common.cuh - helpers shared by the normalisation kernels (kc_norm.cu, kc_norm_cluster.cu).
#pragma once
#include "kc_common.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  // all threads of the block must call; returns the total to every thread
  v = kc_warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0f;
  if (wid == 0) {
    t = kc_warp_sum(t);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}

__device__ __forceinline__ float out_act(int kind, float v, float alpha) {
  if (kind == KC_OUT_PRELU) return v > 0.0f ? v : alpha * v;
  if (kind == KC_OUT_SILU) return kc_silu(v);
  return v;
}
__device__ __forceinline__ float out_act_grad(int kind, float v, float alpha) {
  if (kind == KC_OUT_PRELU) return v > 0.0f ? 1.0f : alpha;
  if (kind == KC_OUT_SILU) return kc_silu_grad(v);
  return 1.0f;
}

}  // namespace
