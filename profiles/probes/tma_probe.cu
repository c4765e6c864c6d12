// Stand-alone probe for the TMA-staged lookup: runs one variant (argv[1]) of the issue/wait sequence
// and checks a box against a CPU gather.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define BOX 12
__device__ __forceinline__ uint32_t sm(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VARIANT>
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int bx, int by, int plane) {
  __shared__ __align__(128) float patch[160];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x;
  const uint32_t ba = sm(&bar);
  bool leader = lane == 0;
  if (VARIANT == 2) { uint32_t p; asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(p)); leader = p != 0; }
  if (leader) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (VARIANT == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (leader) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(BOX * BOX * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(sm(patch)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bx), "r"(by), "r"(plane), "r"(ba) : "memory");
  }
  __syncwarp();
  uint32_t done = 0; long spins = 0;
  for (; spins < (1 << 24) && !done; ++spins)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(ba) : "memory");
  if (lane == 0) out[BOX * BOX] = done ? (float)spins : -1.0f;
  for (int i = lane; i < BOX * BOX; i += 32) out[i] = done ? patch[i] : -777.0f;
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  int ld = argc > 2 ? atoi(argv[2]) : 56, h = argc > 3 ? atoi(argv[3]) : 30, planes = 100;
  int bx = argc > 4 ? atoi(argv[4]) : -3, by = argc > 5 ? atoi(argv[5]) : 25, plane = 7;
  std::vector<float> host((size_t)planes * h * ld);
  for (size_t i = 0; i < host.size(); ++i) host[i] = (float)(i % 100003) * 0.001f;
  float *d, *o; cudaMalloc(&d, host.size() * 4); cudaMalloc(&o, (BOX * BOX + 1) * 4);
  cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*ENC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)h, (cuuint64_t)planes}; cuuint64_t str[2] = {(cuuint64_t)ld * 4, (cuuint64_t)h * ld * 4};
  cuuint32_t box[3] = {BOX, BOX, 1}, es[3] = {1, 1, 1};
  CUresult r = ((ENC)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d ld %d h %d box@(%d,%d): encode rc=%d q=%d\n", variant, ld, h, bx, by, (int)r, (int)q);
  if (variant == 0) k<0><<<1, 32>>>(tm, o, bx, by, plane); else if (variant == 1) k<1><<<1, 32>>>(tm, o, bx, by, plane); else k<2><<<1, 32>>>(tm, o, bx, by, plane);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  sync: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> out(BOX * BOX + 1); cudaMemcpy(out.data(), o, out.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < BOX; ++y) for (int x = 0; x < BOX; ++x) {
    int gx = bx + x, gy = by + y; float ref = (gx >= 0 && gx < ld && gy >= 0 && gy < h) ? host[((size_t)plane * h + gy) * ld + gx] : 0.f;
    if (out[y * BOX + x] != ref) ++bad;
  }
  printf("  spins %.0f mismatches %d (first vals %.3f %.3f %.3f %.3f)\n", out[BOX * BOX], bad, out[0], out[3], out[4], out[BOX * 4 + 5]);
  return 0;
}
