// Mask-guided sparse window attention on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Replaces SparseWindowAttention.forward (model/modules/sparse_transformer.py:177-275) for *masked*
// windows: 128 query rows x 128 head dims per CTA, keys streamed in tiles of 64 (own + rolled tokens by
// table lookup, pooled tokens), flash-style online softmax.  Unmasked windows (45x45 per frame) stay on
// the small mma.sync kernel in mma_kernels.cu.
//
//   S = Q K^T   tcgen05.mma kind::tf32, A = Q (smem, K-major SWIZZLE_128B), B = K tile (smem, K-major SW128),
//               D = S in TMEM (two 64-column buffers)
//   P = exp2(S - m)   softmax warps: tcgen05.ld S -> registers -> tcgen05.st P (TMEM)
//   O += P V    tcgen05.mma with A = P straight from TMEM, B = V^T tile (smem, K-major SW128: the producer
//               warps transpose V while staging it; bank-conflict free), D = O in TMEM (128 columns)
// O is rescaled in TMEM only when a row maximum grows by more than 2^8 (lazy rescaling); the final 1/l is
// applied in the epilogue.  Roles: warps 0-3 softmax / correction / epilogue (one query row per thread =
// one TMEM lane), warps 4-7 K/V producers (cp.async + transposing stores), warp 8 TMEM allocator + MMA issuer.
// All waits are bounded and trap instead of hanging.  Descriptor encodings were validated in isolation
// with profiles/probes/umma_probe.cu (variants 0 and 3).
#include "pp_elem.cuh"
#include "pp_mma.cuh"
#include "../../include/propainter_b200.h"

#define UA_BM 128
#define UA_BN 64
#define UA_THREADS 288
#define UA_Q_BYTES (UA_BM * 128 * 4)               // 64 KB: 4 k-blocks of [128 rows x 128 B]
#define UA_K_BYTES (UA_BN * 128 * 4)               // 32 KB: 4 k-blocks of [64 rows x 128 B]
#define UA_V_BYTES (128 * UA_BN * 4)               // 32 KB: 2 k-blocks of [128 rows x 128 B]  (V^T: rows = head dim)
#define UA_STAGE_BYTES (UA_K_BYTES + UA_V_BYTES)
#define UA_SMEM_BYTES (UA_Q_BYTES + 2 * UA_STAGE_BYTES + 1024)

__device__ __forceinline__ uint32_t ua_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ua_bar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ua_bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ua_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 26) && !done; ++spin)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (!done) __trap();
}
__device__ __forceinline__ void ua_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (version 1, LBO unused = 16 B, SBO = 1024 B between 8-row groups)
__device__ __forceinline__ uint64_t ua_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// byte offset of element (row, k) in a K-major SW128 tile with `rows` rows: k-block (32 floats) major, 8-row groups of 1 KB
__device__ __forceinline__ uint32_t ua_off(int row, int k, int rows) {
  const int kb = k >> 5, kk = k & 31, r = row & 7;
  return (uint32_t)(kb * rows * 128 + (row >> 3) * 1024 + r * 128 + (((kk >> 2) ^ r) << 4) + (kk & 3) * 4);
}
__device__ __forceinline__ void ua_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ua_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
#define UA_LD32(taddr, v)                                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),   \
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),    \
                 "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),    \
                 "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                                       \
               : "r"(taddr))
#define UA_ST32(taddr, v)                                                                                                    \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),  \
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), \
                 "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), \
                 "r"(v[30]), "r"(v[31]) : "memory")

__global__ void __launch_bounds__(UA_THREADS, 1) k_sparse_attn_umma(PPAttnParams p) {
  extern __shared__ __align__(1024) uint8_t ua_raw[];
  // barriers: 0,1 kv_full[2]  2,3 kv_empty[2]  4,5 s_full[2]  6,7 p_full[2]  8 pv_done
  __shared__ __align__(8) unsigned long long bars[9];
  __shared__ uint32_t tmem_base_s;
  const int win = blockIdx.z, head = blockIdx.y;
  if (p.flags[win] == 0) return;
  uint8_t* base = (uint8_t*)(((uintptr_t)ua_raw + 1023) & ~(uintptr_t)1023);        // SWIZZLE_128B tiles need 1 KB alignment
  uint8_t* sQ = base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int* ktab = p.key_tok + (long)win * p.NKO;
  const int hoff = head * 128;
  const int q0 = blockIdx.x * UA_BM;
  const int nq = min(UA_BM, p.t * p.WN - q0);
  const int keys_per_frame = p.NKO + p.NP;
  const int nkeys = p.nkf * keys_per_frame;
  const int ntiles = (nkeys + UA_BN - 1) / UA_BN;
  const uint32_t b0 = ua_smem(&bars[0]);
  auto bar = [&](int i) { return b0 + 8u * (uint32_t)i; };

  if (tid == 0) {
    ua_bar_init(bar(0), 128); ua_bar_init(bar(1), 128);        // kv_full: every producer thread arrives
    ua_bar_init(bar(2), 1); ua_bar_init(bar(3), 1);            // kv_empty: tcgen05.commit
    ua_bar_init(bar(4), 1); ua_bar_init(bar(5), 1);            // s_full: tcgen05.commit
    ua_bar_init(bar(6), 128); ua_bar_init(bar(7), 128);        // p_full: every softmax thread arrives
    ua_bar_init(bar(8), 1);                                    // pv_done: tcgen05.commit
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ua_smem(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- query tile -> shared memory (K-major SW128), pre-scaled into the log2 domain
  for (int idx = tid; idx < UA_BM * 32; idx += UA_THREADS) {
    const int row = idx >> 5, c4 = idx & 31;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nq) {
      const int qi = q0 + row, fr = qi / p.WN, tok = ktab[qi - fr * p.WN];
      v = *reinterpret_cast<const float4*>(p.qkv + ((long)fr * p.NT + tok) * p.ld_qkv + hoff + c4 * 4);
    }
    v.x *= p.scale_log2; v.y *= p.scale_log2; v.z *= p.scale_log2; v.w *= p.scale_log2;
    *reinterpret_cast<float4*>(sQ + ua_off(row, c4 * 4, UA_BM)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t tO = tbase, tS0 = tbase + 128, tP0 = tbase + 256;      // O: 128 cols; S: 2 x 64; P: 2 x 64

  if (warp >= 4 && warp < 8) {
    // ================================================= producers: K tile (cp.async) + V^T tile (transposing stores)
    const int ptid = tid - 128;
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1, use = j >> 1;
      ua_bar_wait(bar(2 + s), (use & 1) ^ 1);                            // stage free (passes immediately on first use)
      uint8_t* sK = base + UA_Q_BYTES + s * UA_STAGE_BYTES;
      uint8_t* sV = sK + UA_K_BYTES;
      // K: 64 keys x 32 chunks of 16 B
      for (int idx = ptid; idx < UA_BN * 32; idx += 128) {
        const int key = idx >> 5, c4 = idx & 31, jk = j * UA_BN + key;
        uint8_t* dst = sK + ua_off(key, c4 * 4, UA_BN);
        if (jk < nkeys) {
          const int kfi = jk / keys_per_frame, slot = jk - kfi * keys_per_frame, fr = p.kf_start + kfi * p.kf_step;
          const float* src = slot < p.NKO ? p.qkv + ((long)fr * p.NT + ktab[slot]) * p.ld_qkv + p.C + hoff
                                          : p.pool + ((long)fr * p.NP + (slot - p.NKO)) * p.ld_pool + hoff;
          pp_cp_async16(dst, src + c4 * 4);
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      pp_cp_async_commit();
      // V^T: thread <-> (key, half of the head dims); for a fixed dim the 32 lanes (keys) hit 32 distinct banks
      {
        const int key = ptid & 63, half = ptid >> 6, jk = j * UA_BN + key;
        const float* src = nullptr;
        if (jk < nkeys) {
          const int kfi = jk / keys_per_frame, slot = jk - kfi * keys_per_frame, fr = p.kf_start + kfi * p.kf_step;
          src = slot < p.NKO ? p.qkv + ((long)fr * p.NT + ktab[slot]) * p.ld_qkv + 2 * p.C + hoff
                             : p.pool + ((long)fr * p.NP + (slot - p.NKO)) * p.ld_pool + p.C + hoff;
        }
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int d = half * 64 + i * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (src) v = *reinterpret_cast<const float4*>(src + d);
          *reinterpret_cast<float*>(sV + ua_off(d + 0, key, 128)) = v.x;
          *reinterpret_cast<float*>(sV + ua_off(d + 1, key, 128)) = v.y;
          *reinterpret_cast<float*>(sV + ua_off(d + 2, key, 128)) = v.z;
          *reinterpret_cast<float*>(sV + ua_off(d + 3, key, 128)) = v.w;
        }
      }
      pp_cp_async_wait<0>();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05.mma
      ua_bar_arrive(bar(0 + s));
    }
  } else if (warp == 8) {
    // ================================================= MMA issuer (one elected thread)
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UA_BN >> 3) << 17) | ((uint32_t)(UA_BM >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(UA_BM >> 4) << 24);
      const uint32_t qaddr = ua_smem(sQ);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        ua_bar_wait(bar(0 + s), (j >> 1) & 1);                           // K/V tile j landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t kaddr = ua_smem(base + UA_Q_BYTES + s * UA_STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)                                   // 128 head dims = 4 k-blocks x 4 k-steps of 8
          ua_mma_ss(tS0 + s * UA_BN, ua_desc(qaddr + (ks >> 2) * (UA_BM * 128) + (ks & 3) * 32),
                    ua_desc(kaddr + (ks >> 2) * (UA_BN * 128) + (ks & 3) * 32), idesc_s, ks > 0);
        ua_commit(bar(4 + s));                                           // S_j ready for the softmax warps
      };
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        if (j + 1 < ntiles) issue_s(j + 1);                              // S buffer (j+1)&1 is free: P_{j-1} was consumed by PV_{j-1}
        ua_bar_wait(bar(6 + s), (j >> 1) & 1);                           // P_j written (and O rescaled if needed)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t vaddr = ua_smem(base + UA_Q_BYTES + s * UA_STAGE_BYTES + UA_K_BYTES);
#pragma unroll
        for (int ks = 0; ks < UA_BN / 8; ++ks)                            // 64 keys = 2 k-blocks x 4 k-steps
          ua_mma_ts(tO, tP0 + s * UA_BN + ks * 8, ua_desc(vaddr + (ks >> 2) * (128 * 128) + (ks & 3) * 32), idesc_o,
                    (j > 0 || ks > 0));
        ua_commit(bar(2 + s));                                           // stage s may be refilled
        ua_commit(bar(8));                                               // O holds tiles 0..j
      }
    }
  } else {
    // ================================================= softmax / correction / epilogue: thread <-> query row <-> TMEM lane
    const int row = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      ua_bar_wait(bar(4 + s), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v0[32], v1[32];
      UA_LD32(tS0 + s * UA_BN + lane_off, v0);
      UA_LD32(tS0 + s * UA_BN + 32 + lane_off, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int kbase = j * UA_BN;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float a = kbase + c < nkeys ? __uint_as_float(v0[c]) : -INFINITY;
        float b = kbase + 32 + c < nkeys ? __uint_as_float(v1[c]) : -INFINITY;
        v0[c] = __float_as_uint(a); v1[c] = __float_as_uint(b);
        mx = fmaxf(mx, fmaxf(a, b));
      }
      // lazy rescaling: keep the running reference maximum unless some row grew by more than 2^8
      const bool grow = mx > m_used + 8.0f;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_used;
        const float alpha = exp2f(m_used - m_new);                        // m_used = -inf on the first tile -> alpha = 0
        if (j > 0) {
          ua_bar_wait(bar(8), (j - 1) & 1);                               // PV_{j-1} finished: O is quiescent
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t o[32];
            UA_LD32(tO + c0 + lane_off, o);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            UA_ST32(tO + c0 + lane_off, o);
          }
        }
        l *= alpha;
        m_used = m_new;
      }
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float a = exp2f(__uint_as_float(v0[c]) - m_used), b = exp2f(__uint_as_float(v1[c]) - m_used);
        rs += a + b;
        v0[c] = __float_as_uint(a); v1[c] = __float_as_uint(b);
      }
      l += rs;
      UA_ST32(tP0 + s * UA_BN + lane_off, v0);
      UA_ST32(tP0 + s * UA_BN + 32 + lane_off, v1);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ua_bar_arrive(bar(6 + s));
    }
    // ---- epilogue: O / l -> global
    ua_bar_wait(bar(8), (ntiles - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = 1.0f / l;
    float* orow = nullptr;
    if (row < nq) { const int qi = q0 + row, fr = qi / p.WN; orow = p.out + ((long)fr * p.NT + ktab[qi - fr * p.WN]) * p.ld_out + hoff; }
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t o[32];
      UA_LD32(tO + c0 + lane_off, o);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (orow) {
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          *reinterpret_cast<float4*>(orow + c0 + c) = make_float4(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv,
                                                                 __uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// masked windows on tcgen05; called by pp_sparse_window_attn (mma_kernels.cu)
int pp_launch_sparse_attn_umma(const PPAttnParams& p, int n_windows, cudaStream_t stream) {
  if (cudaFuncSetAttribute(k_sparse_attn_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, UA_SMEM_BYTES) != cudaSuccess)
    return PP_ERR_LAUNCH;
  dim3 grid((p.t * p.WN + UA_BM - 1) / UA_BM, p.C / 128, n_windows);
  k_sparse_attn_umma<<<grid, UA_THREADS, UA_SMEM_BYTES, stream>>>(p);
  return cudaPeekAtLastError() == cudaSuccess ? PP_OK : PP_ERR_LAUNCH;
}
