// Pipe-throughput microbenchmarks for the instruction mix of the fused raw->sRGB kernel (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define ITERS 4096
__device__ __forceinline__ void fma2(float &dx, float &dy, float ax, float ay, float bx, float by, float cx, float cy) {
  asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};"
      " fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}" : "=f"(dx), "=f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(cx), "f"(cy));
}
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float *out, float a, float b, const float2 *tab, const uint32_t *idx) {
  __shared__ float2 lut[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) lut[i] = tab[i];
  __syncthreads();
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = a + threadIdx.x * 1e-6f + j;
  uint32_t key = idx[threadIdx.x];
  double d0 = a, d1 = b;
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {  // scalar FFMA 3-reg
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = fmaf(x[j], a, b);
    } else if (MODE == 1) {  // FFMA2
#pragma unroll
      for (int j = 0; j < 8; j += 2) fma2(x[j], x[j + 1], x[j], x[j + 1], a, a, b, b);
#pragma unroll
      for (int j = 0; j < 8; j += 2) fma2(x[j], x[j + 1], x[j], x[j + 1], a, a, b, b);
    } else if (MODE == 2) {  // FMUL + FADD unfused
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = __fadd_rn(__fmul_rn(x[j], a), b);
    } else if (MODE == 3) {  // FFMA2 + FMNMX co-issue: 8 fma-lanes + 8 alu ops
#pragma unroll
      for (int j = 0; j < 8; j += 2) fma2(x[j], x[j + 1], x[j], x[j + 1], a, a, b, b);
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = fminf(x[j], b);
    } else if (MODE == 4) {  // scalar FFMA + FMNMX
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = fminf(fmaf(x[j], a, b), b);
    } else if (MODE == 5) {  // LDS.64 random gather (8 per iter)
#pragma unroll
      for (int j = 0; j < 8; j++) { float2 e = lut[key & 4095]; key = key * 1664525u + 1013904223u + __float_as_uint(e.x); x[j] += e.y; }
    } else if (MODE == 6) {  // LDS.32 random gather
      const float *l1 = reinterpret_cast<const float *>(lut);
#pragma unroll
      for (int j = 0; j < 8; j++) { float e = l1[key & 4095]; key = key * 1664525u + 1013904223u + __float_as_uint(e); x[j] += e; }
    } else if (MODE == 7) {  // DFMA
#pragma unroll
      for (int j = 0; j < 4; j++) { d0 = fma(d0, d1, d1); d1 = fma(d1, d0, d0); }
    } else if (MODE == 8) {  // F2F f32->f64->f32
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = (float)((double)x[j] + d0);
    } else if (MODE == 9) {  // I2F.U16
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = x[j] + (float)(uint16_t)(__float_as_uint(x[j]) >> 3);
    } else if (MODE == 10) {  // FSEL/FSETP pairs
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = x[j] > a ? x[j] : b + x[j];
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)d0 + (float)d1 + key;
}
template <int MODE> void run(const char *name, double ops_per_iter, float *out, const float2 *tab, const uint32_t *idx) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 1024>>>(out, 1.0001f, 0.5f, tab, idx);
  cudaEventRecord(e0);
  k<MODE><<<148, 1024>>>(out, 1.0001f, 0.5f, tab, idx);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double total = ops_per_iter * ITERS * 148.0 * 1024.0;
  printf("%-28s %8.3f ms  %8.1f Gop/s  %6.2f thread-ops/clk/SM @1.965GHz\n", name, ms, total / ms / 1e6, total / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  float *out; float2 *tab; uint32_t *idx;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&tab, 8192 * 8); cudaMalloc(&idx, 1024 * 4);
  cudaMemset(tab, 0, 8192 * 8);
  uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = i * 2654435761u;
  cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("FFMA scalar (fma/iter=8)", 8, out, tab, idx);
  run<1>("FFMA2 (f32 fma/iter=16)", 16, out, tab, idx);
  run<2>("FMUL+FADD (instr/iter=16)", 16, out, tab, idx);
  run<3>("FFMA2x4+FMNMXx8 (instr=12)", 12, out, tab, idx);
  run<4>("FFMA+FMNMX x8 (instr=16)", 16, out, tab, idx);
  run<5>("LDS.64 random (8/iter)", 8, out, tab, idx);
  run<6>("LDS.32 random (8/iter)", 8, out, tab, idx);
  run<7>("DFMA (8/iter)", 8, out, tab, idx);
  run<8>("F2F pair (16 cvt/iter)", 16, out, tab, idx);
  run<9>("I2F.U16 (8/iter)", 8, out, tab, idx);
  run<10>("FSETP+FSEL+FADD (24)", 24, out, tab, idx);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
