// Probe which 2-D u16 TMA box/tensor configurations the hardware accepts:  ./tma_probe W H BOXW BOXH X Y
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int n, uint16_t *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint16_t *tile = reinterpret_cast<uint16_t *>(smem);
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + ((n * 2 + 127) / 128) * 128);
  uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(tile);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(d), "l"(&tmap), "r"(x), "r"(y), "r"(b) : "memory");
  }
  __syncthreads();
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(b), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char **argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), BW = atoi(argv[3]), BH = atoi(argv[4]), X = atoi(argv[5]), Y = atoi(argv[6]);
  std::vector<uint16_t> h(W * H);
  for (int i = 0; i < W * H; i++) h[i] = (uint16_t)(1 + (i % 60000));
  uint16_t *d, *o; cudaMalloc(&d, W * H * 2); cudaMalloc(&o, BW * BH * 2);
  cudaMemcpy(d, h.data(), W * H * 2, cudaMemcpyHostToDevice);
  void *f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W * 2};
  cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)f)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("W=%d H=%d box=%dx%d at (%d,%d): encode=%d ", W, H, BW, BH, X, Y, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
  int n = BW * BH;
  size_t smem = ((n * 2 + 127) / 128) * 128 + 16;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<1, 256, smem>>>(map, X, Y, n, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel=%s ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<uint16_t> g(n); cudaMemcpy(g.data(), o, n * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BH; r2++) for (int c = 0; c < BW; c++) {
      int yy = Y + r2, xx = X + c; uint16_t want = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? h[yy * W + xx] : 0;
      bad += g[r2 * BW + c] != want;
    }
    printf("mismatches=%d", bad);
  }
  printf("\n");
  return 0;
}
