// Issue-rate probe for tcgen05.mma (sm_100a): how many cycles does one M=128 MMA cost as a function of N, of where A
// comes from (shared memory descriptor vs TMEM) and of the shared-memory layout (K-major SWIZZLE_128B / SWIZZLE_32B /
// no swizzle), for kind::tf32 (K = 8 per instruction) and kind::f16 (bf16, K = 16)?  One CTA, one elected lane issues
// `iters` back-to-back MMAs into one accumulator (or alternating between two), commits, and the warp waits on the mbarrier;
// cycles = clock64 around issue + completion.  Operands are zeros (only the rate matters here; layouts are validated by
// umma_probe.cu and the parity tests).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_probe umma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>

__device__ __forceinline__ uint32_t sm(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect() {
  uint32_t pred = 0;
  asm volatile("{ .reg .b32 r; .reg .pred p; elect.sync r|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

// mode 0: SS SW128 tf32   1: TS (A in TMEM) SW128 B tf32   2: SS SW32 tf32   3: SS no-swizzle tf32   4: SS SW128 bf16   5: TS bf16
__global__ void __launch_bounds__(128) rate(int mode, int N, int iters, int two_acc, int advance, long long* out, int M = 128) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    const bool bf16 = mode >= 4, ts = (mode == 1 || mode == 5);
    const uint32_t fmt = bf16 ? 1u : 2u;                           // a/b format: 1 = bf16 (kind::f16), 2 = tf32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t layout = 2, sbo = 1024, lbo = 16, kstep = 32;         // SW128: 8-row groups 1 KB apart, k-step = 32 B inside the row
    if (mode == 2) { layout = 6; sbo = 256; kstep = 4096; }        // SW32: rows of 32 B, one k-step = whole [128 x 32 B] slab
    if (mode == 3) { layout = 0; sbo = 128; lbo = 2048; kstep = 4096; }   // interleaved 8x16B core matrices
    const uint32_t a_addr = sm(smem), b_addr = sm(smem + 64 * 1024);
    const uint64_t ad0 = make_desc(a_addr, lbo, sbo, layout), bd0 = make_desc(b_addr, lbo, sbo, layout);
    const uint32_t tD0 = tb, tD1 = tb + 256, tA = tb + 384;
    long long t0 = clock64();
    if (elect()) {
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint64_t ad = ad0 + (advance ? (uint64_t)(((u * kstep) % 16384) >> 4) : 0);
          const uint64_t bd = bd0 + (advance ? (uint64_t)(((u * kstep) % 16384) >> 4) : 0);
          const uint32_t td = (two_acc && (u & 1)) ? tD1 : tD0;
          if (!ts) {
            if (!bf16) asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                                    ::"r"(td), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            else asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                              ::"r"(td), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
          } else {
            if (!bf16) asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }"
                                    ::"r"(td), "r"(tA + u * 8), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            else asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }"
                              ::"r"(td), "r"(tA + u * 8), "l"(bd), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sm(&bar)) : "memory");
    }
    __syncwarp();
    long long t1 = clock64();
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 26) && !done; ++spin)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(sm(&bar)) : "memory");
    long long t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

// Where do the rows of an M = 64 accumulator live in TMEM?  A[m][0] = m + 1 (other k zero), B[n][0] = 1: D[m][n] = m + 1.
// Every one of the 128 lanes reads column 0 of the accumulator; printed as lane -> value.
__global__ void __launch_bounds__(128) m64_layout(float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // K-major SW128: row r at r*128 (+ 1 KB per 8-row group), 16-byte chunk c stored at c ^ (r & 7); element k = 0 is in chunk 0
  if (tid < 64) *reinterpret_cast<float*>(smem + (tid >> 3) * 1024 + (tid & 7) * 128 + (((0) ^ (tid & 7)) << 4)) = (float)(tid + 1);
  if (tid < 32) *reinterpret_cast<float*>(smem + 32 * 1024 + (tid >> 3) * 1024 + (tid & 7) * 128 + (((0) ^ (tid & 7)) << 4)) = 1.0f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm(&tmem_base_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  // pre-fill the accumulator columns with a sentinel so untouched lanes are visible
  {
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(-7.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(tb + ((uint32_t)(warp * 32) << 16)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
                 "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    if (elect()) {
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                   ::"r"(tb), "l"(make_desc(sm(smem), 16, 1024, 2)), "l"(make_desc(sm(smem + 32 * 1024), 16, 1024, 2)), "r"(idesc), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sm(&bar)) : "memory");
    }
    __syncwarp();
  }
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 24) && !done; ++spin)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(sm(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(tb + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[tid] = __uint_as_float(r0);
  out[128 + tid] = __uint_as_float(r1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64));
}

// Epilogue timing: after one committed MMA, how long do (a) tcgen05.ld 32x32b.x32 + wait, (b) 8 STS.128 + syncwarp,
// (c) 8 x (LDS.128 + STG.128) take for 4 warps?  Second round repeats (a)-(c) to separate one-time from steady costs.
__global__ void __launch_bounds__(192) epi_probe(float* gout, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * 1024 / 4; i += 192) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  if (warp == 5) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (elect()) {
      for (int i = 0; i < 64; ++i)
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                     ::"r"(tb), "l"(make_desc(sm(smem), 16, 1024, 2)), "l"(make_desc(sm(smem + 32 * 1024), 16, 1024, 2)), "r"(idesc), "r"(1u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sm(&bar)) : "memory");
    }
    __syncwarp();
  } else if (warp < 4) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(sm(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    long long t[8];
    t[0] = clock64();
    for (int round = 0; round < 2; ++round) {
      uint32_t v[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                     "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                     "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(tb + round * 32 + ((uint32_t)(warp * 32) << 16)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      t[1 + 3 * round] = clock64();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      __syncwarp();
      t[2 + 3 * round] = clock64();
      const int rsub = lane >> 3, c4 = lane & 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * 36 + 4 * c4);
        *reinterpret_cast<float4*>(gout + ((long)(warp * 32 + 4 * i + rsub) * 128 + round * 32 + 4 * c4)) = a;
      }
      __syncwarp();
      t[3 + 3 * round] = clock64();
    }
    if (tid == 0) for (int i = 0; i < 7; ++i) out[i] = t[i] - t[0];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  {
    float* go; long long h[8];
    cudaMalloc(&go, 128 * 128 * 4);
    cudaFuncSetAttribute(epi_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 2048);
    for (int rep = 0; rep < 3; ++rep) {
      epi_probe<<<1, 192, 64 * 1024 + 2048>>>(go, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("epi_probe: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
      printf("epilogue probe rep %d (cycles after acc_full): round0 tmem-ld %lld  staged %lld  stored %lld | round1 tmem-ld %lld staged %lld stored %lld\n",
             rep, h[1], h[2], h[3], h[4], h[5], h[6]);
    }
    if (getenv("EPI_ONLY")) return 0;
  }
  {
    float* o; float h[256];
    cudaMalloc(&o, 1024);
    cudaFuncSetAttribute(m64_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 2048);
    m64_layout<<<1, 128, 64 * 1024 + 2048>>>(o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=64 accumulator layout probe: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
      cudaMemcpy(h, o, 1024, cudaMemcpyDeviceToHost);
      for (int l = 0; l < 128; ++l) printf("%s lane %3d: col0 %5.1f col1 %5.1f", l % 4 == 0 ? "\n" : " |", l, h[l], h[128 + l]);
      printf("\n");
    } else return 1;
  }
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[6] = {"SS sw128 tf32", "TS tf32 (A in TMEM)", "SS sw32 tf32", "SS interleave tf32", "SS sw128 bf16", "TS bf16"};
  const int iters = 512;
  for (int N = 32; N <= 256; N *= 2) {                             // M = 64 instruction shape, SS sw128 tf32
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
      rate<<<1, 128, 200 * 1024>>>(0, N, iters, 0, 1, d, 64);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("M=64 N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    printf("%-22s N=%3d M=64 advance=1: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA\n", "SS sw128 tf32", N, (double)h[0] / iters, (double)h[1] / iters);
  }
  for (int mode = 0; mode < 6; ++mode)
    for (int N = 32; N <= 256; N *= 2)
      for (int two = 0; two < 2; ++two) {
        if (two && N > 128) continue;
        for (int adv = 0; adv < 2; ++adv) {
          long long h[2] = {0, 0};
          for (int rep = 0; rep < 2; ++rep) {
            rate<<<1, 128, 200 * 1024>>>(mode, N, iters, two, adv, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          }
          printf("%-22s N=%3d two_acc=%d advance=%d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal %5.1f)\n", names[mode], N, two, adv,
                 (double)h[0] / iters, (double)h[1] / iters, mode >= 4 ? N / 2.0 : N / 2.0);
        }
      }
  return 0;
}
