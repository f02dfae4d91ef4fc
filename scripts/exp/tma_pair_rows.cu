// Probe: TMA tensor store / load of boxes that start 8 bytes off 16-byte alignment inside a row PAIR of a
// (B, V, 3) fp32 tensor (row pitch 82 680 B = 8 mod 16; pair pitch 165 360 B = 0 mod 16).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void store_kernel(const __grid_constant__ CUtensorMap map, int V3, int col0, int mode) {
  __shared__ __align__(128) float tile[2][16][24];
  const int lane = threadIdx.x;
  for (int i = lane; i < 2 * 16 * 24; i += 32) (&tile[0][0][0])[i] = 1000.f * (i / 384) + 10.f * ((i % 384) / 24) + 0.01f * (i % 24);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    for (int par = 0; par < 2; ++par) {
      if (mode == 1 && par == 1) continue;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(&map), "r"(smem_addr(&tile[par][0][0])), "r"(par * V3 + col0), "r"(0) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int B = 32, V = 6890, V3 = V * 3;
  float* d;
  cudaMalloc(&d, (size_t)B * V3 * 4);
  cudaMemset(d, 0, (size_t)B * V3 * 4);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)2 * V3, (cuuint64_t)B / 2};
  cuuint64_t strides[1] = {(cuuint64_t)2 * V3 * 4};
  cuuint32_t box[2] = {24, 16}, estr[2] = {1, 1};
  CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  const int col0 = 48;
  store_kernel<<<1, 32>>>(map, V3, col0, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> h((size_t)B * V3);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < V3; ++c) {
      float want = 0.f;
      if (c >= col0 && c < col0 + 24 && !(mode == 1 && (b & 1))) want = 1000.f * (b & 1) + 10.f * (b >> 1) + 0.01f * (c - col0);
      if (h[(size_t)b * V3 + c] != want && bad++ < 5) printf("mismatch b=%d c=%d got %f want %f\n", b, c, h[(size_t)b * V3 + c], want);
    }
  printf("mode %d: %s (%d mismatches)\n", mode, bad ? "WRONG" : "ok", bad);
  return bad != 0;
}
