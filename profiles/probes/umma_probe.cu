// Stand-alone probe of the tcgen05 building blocks the sparse-attention kernel needs (sm_100a):
//   D[128 x N] (TMEM, fp32) = A[128 x K] (tf32) * B[N x K]^T,   K = 128, N = 128
// variant 0: A, B both K-major SWIZZLE_128B in shared memory (Q.K^T shape)
// variant 1/2: B MN-major SWIZZLE_128B (V of P.V: rows = k, n contiguous), two LBO/SBO conventions
// variant 3/4: like 0 / 1 but A comes from TMEM (written with tcgen05.st) -- P of P.V
// variant 5/6: B MN-major in the SWIZZLE_128B_BASE32B layout (layout type 1) with A from smem / TMEM.  CUTLASS
//              (cutlass/gemm/collective/builders/sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
//              available smem layout") explains why variants 1/2/4 return zeros: 32-bit MN-major operands need the
//              32-byte-base swizzle  Swizzle<2,5,2>: atom = 4 k-rows x 128 B, 32-byte granule g of row r stored at g ^ (r & 3);
//              canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp:238-268).
//              Verified on B200: variants 5 and 6 reproduce variant 3's result (profiles/r1_attn_vmn_probe.log).
// Each run prints max |D - ref| (fp32 host reference; TF32 inputs => ~1e-2 abs at these magnitudes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <stdint.h>

#define M_ 128
#define N_ 128
#define K_ 128

__device__ __forceinline__ uint32_t sm(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B: block kb (32 floats of K) = [rows][128 B]; 8-row groups of 1024 B; 16-B chunk c of row r stored at c ^ (r & 7)
__device__ __forceinline__ uint32_t off_kmajor(int row, int k, int rows) {
  int kb = k >> 5, kk = k & 31, c = kk >> 2, w = kk & 3, r = row & 7, grp = row >> 3;
  return (uint32_t)(kb * rows * 128 + grp * 1024 + r * 128 + ((c ^ r) << 4) + w * 4);
}
// MN-major SWIZZLE_128B for B[k][n]: atom = 8 k-rows x 128 B (32 n); atom (kg, nb) at (kg*NB + nb)*1024
__device__ __forceinline__ uint32_t off_mnmajor(int k, int n, int nblocks) {
  int kg = k >> 3, kr = k & 7, nb = n >> 5, nn = n & 31, c = nn >> 2, w = nn & 3;
  return (uint32_t)((kg * nblocks + nb) * 1024 + kr * 128 + ((c ^ kr) << 4) + w * 4);
}
// MN-major SWIZZLE_128B_BASE32B for B[k][n]: n-block (32 n = 128 B) major, then k rows of 128 B; 4-row swizzle atoms
__device__ __forceinline__ uint32_t off_mn32(int k, int n, int krows) {
  int nb = n >> 5, nn = n & 31, g = nn >> 3, w = nn & 7;
  return (uint32_t)(nb * krows * 128 + k * 128 + ((g ^ (k & 3)) << 5) + w * 4);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
  d |= (uint64_t)layout << 61;     // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // 128 x 128 floats = 64 KB
  uint8_t* sB = smem + 65536;               // 64 KB
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool b_mn32 = (variant == 5 || variant == 6);
  const bool b_mn = (variant == 1 || variant == 2 || variant == 4) || b_mn32;
  const bool a_tmem = (variant == 3 || variant == 4 || variant == 6);
  for (int i = tid; i < M_ * K_; i += 128) { int m = i / K_, k = i % K_; *(float*)(sA + off_kmajor(m, k, M_)) = A[i]; }
  for (int i = tid; i < N_ * K_; i += 128) {
    int n = i / K_, k = i % K_;
    uint32_t o = b_mn32 ? off_mn32(k, n, K_) : b_mn ? off_mnmajor(k, n, N_ / 32) : off_kmajor(n, k, N_);
    *(float*)(sB + o) = B[i];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t tD = tbase, tA = tbase + 256;                       // D: columns [0,128), A operand copy: [256,384)
  if (a_tmem) {                                                        // A[m][k] -> TMEM lane m, column tA + k
    const int m = warp * 32 + lane;
    for (int k0 = 0; k0 < K_; k0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[m * K_ + k0 + j]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tA + ((uint32_t)(warp * 32) << 16) + k0),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) {
    // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), b_major bit 16, N>>3 at 17, M>>4 at 24
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N_ >> 3) << 17) | ((uint32_t)(M_ >> 4) << 24);
    if (b_mn) idesc |= (1u << 16);
    for (int ks = 0; ks < K_ / 8; ++ks) {
      const int kb = ks >> 2, k8 = ks & 3;
      const uint64_t adesc = make_desc(sm(sA) + kb * (M_ * 128) + k8 * 32, 16, 1024);
      uint64_t bdesc;
      if (!b_mn) bdesc = make_desc(sm(sB) + kb * (N_ * 128) + k8 * 32, 16, 1024);
      else if (b_mn32) {                                             // 8 k-rows per MMA = 1024 B; LBO = n-block stride, SBO = 4-row group
        bdesc = make_desc(sm(sB) + ks * 1024, K_ * 128, 512, 1);
      } else {
        const uint32_t kstep = (N_ / 32) * 1024;                     // one 8-row k-group of atoms
        bdesc = (variant == 2) ? make_desc(sm(sB) + ks * kstep, kstep, 1024) : make_desc(sm(sB) + ks * kstep, 1024, kstep);
      }
      const uint32_t acc = ks > 0;
      if (!a_tmem)
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(tD),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }" ::"r"(tD),
                     "r"(tA + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sm(&bar)) : "memory");
  }
  uint32_t done = 0;
  for (long spin = 0; spin < (1 << 24) && !done; ++spin)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(sm(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (done) {
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < N_; c0 += 32) {
      uint32_t v[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                     "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                     "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(tD + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[m * N_ + c0 + j] = __uint_as_float(v[j]);
    }
  } else if (tid == 0) D[0] = -12345.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> A(M_ * K_), B(N_ * K_), D(M_ * N_), R(M_ * N_);
  srand(1);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.0f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.0f;
  for (int m = 0; m < M_; ++m) for (int n = 0; n < N_; ++n) { double s = 0; for (int k = 0; k < K_; ++k) s += (double)A[m * K_ + k] * B[n * K_ + k]; R[m * N_ + n] = (float)s; }
  float *dA, *dB, *dD; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  probe<<<1, 128, 131072 + 1024>>>(dA, dB, dD, variant);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d: %s\n", variant, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0; int worst = 0;
  for (int i = 0; i < M_ * N_; ++i) { double d = fabs(D[i] - R[i]); if (d > maxerr) { maxerr = d; worst = i; } maxref = fmax(maxref, fabs(R[i])); }
  printf("  max|D-ref| = %.5f (max|ref| %.3f) worst at (%d,%d): got %.4f want %.4f; D[0]=%.4f D[1]=%.4f D[128]=%.4f\n", maxerr, maxref, worst / N_,
         worst % N_, D[worst], R[worst], D[0], D[1], D[N_]);
  return 0;
}
