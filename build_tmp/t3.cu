__global__ void k(const float* a, float* o) {
  float x = a[threadIdx.x], y = a[threadIdx.x + 32], z = a[threadIdx.x + 64], r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x), "f"(y), "f"(z));
  o[threadIdx.x] = r;
}
