// bw_probe.cu -- read-bandwidth probe: which load path streams HBM fastest on B200?
//   (a) TMA 1-D bulk copies into a shared-memory ring (one producer warp, trivial consumers)
//   (b) per-lane cp.async (LDGSTS) rings, no block-level sync
//   (c) direct 128-bit LDG, unrolled
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu && ./bw_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vectordb-from-scratch_b200/csrc/common.cuh"
using namespace gfi;

__global__ void __launch_bounds__(288, 1) k_tma(const float* x, size_t nbytes, int stage_bytes, int ns, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = (uint64_t*)(smem + (size_t)ns * stage_bytes);
  uint64_t* empty = full + 16;
  int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int s = 0; s < ns; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); } fence_mbar_init(); }
  __syncthreads();
  size_t nblk = nbytes / stage_bytes;
  uint32_t it = 0; float acc = 0.f;
  if (warp == 8) {
    for (size_t b = blockIdx.x; b < nblk; b += gridDim.x, ++it) {
      int s = it % ns; uint32_t ph = (it / ns) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) { mbar_arrive_expect_tx(&full[s], stage_bytes); bulk_g2s(smem + (size_t)s * stage_bytes, (const char*)x + b * stage_bytes, stage_bytes, &full[s]); }
      __syncwarp();
    }
  } else {
    for (size_t b = blockIdx.x; b < nblk; b += gridDim.x, ++it) {
      int s = it % ns; uint32_t ph = (it / ns) & 1;
      mbar_wait(&full[s], ph);
      const float4* p = (const float4*)(smem + (size_t)s * stage_bytes);
      for (int i = tid; i < stage_bytes / 16; i += 256) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int CH, int NSW>
__global__ void __launch_bounds__(256, 1) k_cpasync(const float* x, size_t nbytes, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float4* ring = (float4*)smem + (size_t)warp * NSW * CH * 32;
  const size_t unit = (size_t)CH * 512;  // bytes per warp-unit
  size_t nunits = nbytes / unit;
  size_t gw = (size_t)blockIdx.x * 8 + warp, W = (size_t)gridDim.x * 8;
  float acc = 0.f;
  size_t iu = gw;
  int istage = 0;
  auto issue = [&]() {
    if (iu < nunits) {
      const char* src = (const char*)x + iu * unit;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        uint32_t dst = smem_u32(ring + ((size_t)istage * CH + c) * 32 + lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + ((size_t)c * 32 + lane) * 16) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    iu += W; istage = (istage + 1 == NSW) ? 0 : istage + 1;
  };
  for (int s = 0; s < NSW - 1; ++s) issue();
  int cstage = 0;
  for (size_t cu = gw; cu < nunits; cu += W) {
    asm volatile("cp.async.wait_group %0;" ::"n"(NSW - 2) : "memory");
#pragma unroll
    for (int c = 0; c < CH; ++c) { float4 v = ring[((size_t)cstage * CH + c) * 32 + lane]; acc += v.x + v.y + v.z + v.w; }
    cstage = (cstage + 1 == NSW) ? 0 : cstage + 1;
    issue();
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int CH>
__global__ void __launch_bounds__(512) k_ldg(const float* x, size_t nbytes, float* out) {
  const size_t unit = (size_t)CH * 512;
  size_t nunits = nbytes / unit;
  int lane = threadIdx.x & 31;
  size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, W = ((size_t)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (size_t u = gw; u < nunits; u += W) {
    const float4* p = (const float4*)((const char*)x + u * unit) + lane;
    float4 v[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[c].x), "=f"(v[c].y), "=f"(v[c].z), "=f"(v[c].w) : "l"(p + c * 32));
#pragma unroll
    for (int c = 0; c < CH; ++c) acc += v[c].x + v[c].y + v[c].z + v[c].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <class F> float time_it(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  return ms / reps;
}

int main() {
  size_t nbytes = (size_t)12 << 30;
  float *x, *out; cudaMalloc(&x, nbytes); cudaMalloc(&out, 16); cudaMemset(x, 0, nbytes);
  int sms = 148;
  for (int stage_kb : {8, 16, 24, 32}) for (int ns : {4, 6}) {
    int sb = stage_kb * 1024; size_t sm = (size_t)ns * sb + 512;
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    float ms = time_it([&] { k_tma<<<sms, 288, sm>>>(x, nbytes, sb, ns, out); }, 3);
    printf("tma  stage=%2dKB ns=%d : %.3f ms  %.0f GB/s\n", stage_kb, ns, ms, nbytes / ms / 1e6);
  }
  {
    size_t sm = (size_t)8 * 6 * 8 * 512;
    cudaFuncSetAttribute(k_cpasync<8, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    float ms = time_it([&] { k_cpasync<8, 6><<<sms, 256, sm>>>(x, nbytes, out); }, 3);
    printf("cpasync CH=8 NSW=6 (192KB ring): %.3f ms  %.0f GB/s\n", ms, nbytes / ms / 1e6);
    sm = (size_t)8 * 12 * 4 * 512;
    cudaFuncSetAttribute(k_cpasync<4, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    ms = time_it([&] { k_cpasync<4, 12><<<sms, 256, sm>>>(x, nbytes, out); }, 3);
    printf("cpasync CH=4 NSW=12 (192KB ring): %.3f ms  %.0f GB/s\n", ms, nbytes / ms / 1e6);
    sm = (size_t)8 * 3 * 8 * 512;
    cudaFuncSetAttribute(k_cpasync<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    ms = time_it([&] { k_cpasync<8, 3><<<sms * 2, 256, sm>>>(x, nbytes, out); }, 3);
    printf("cpasync CH=8 NSW=3 x2 CTAs/SM   : %.3f ms  %.0f GB/s\n", ms, nbytes / ms / 1e6);
  }
  for (int mult : {2, 4}) {
    float ms = time_it([&] { k_ldg<8><<<sms * mult, 512>>>(x, nbytes, out); }, 3);
    printf("ldg CH=8 grid=%dx148 x512thr: %.3f ms  %.0f GB/s\n", mult, ms, nbytes / ms / 1e6);
    ms = time_it([&] { k_ldg<4><<<sms * mult, 512>>>(x, nbytes, out); }, 3);
    printf("ldg CH=4 grid=%dx148 x512thr: %.3f ms  %.0f GB/s\n", mult, ms, nbytes / ms / 1e6);
  }
  return 0;
}
