// Mask-guided sparse window attention on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Replaces SparseWindowAttention.forward (model/modules/sparse_transformer.py:177-275) for *masked*
// windows: 128 query rows x 128 head dims per CTA, keys streamed in tiles of 64 (own + rolled tokens by
// table lookup, pooled tokens), flash-style online softmax.  Unmasked windows (45x45 per frame) stay on
// the small mma.sync kernel in mma_kernels.cu.
//
//   S = Q K^T   tcgen05.mma kind::tf32, A = Q (TMEM, written once by the softmax warps), B = K tile (smem, K-major
//               SWIZZLE_128B), D = S in TMEM (two 64-column buffers)
//   P = exp2(S - m)   softmax warps: tcgen05.ld S -> registers -> tcgen05.st P (TMEM)
//   O += P V    tcgen05.mma with A = P straight from TMEM, B = V^T tile (smem, K-major SW128: the producer
//               warps transpose V while staging it; bank-conflict free), D = O in TMEM (128 columns)
// O is rescaled in TMEM only when a row maximum grows by more than 2^8 (lazy rescaling); the final 1/l is
// applied in the epilogue.  Roles: warps 0-3 softmax / correction / epilogue (one query row per thread =
// one TMEM lane), warps 4-11 K/V producers (cp.async + transposing stores; two groups on alternate tiles), warp 12 TMEM
// allocator + MMA issuer.  Q lives in TMEM too (A operand of S), which leaves shared memory for the K/V ring and the V
// staging buffers.  The tcgen05 / mbarrier primitives and descriptor encodings live in pp_umma.cuh.
#include "pp_elem.cuh"
#include "pp_mma.cuh"
#include "pp_umma.cuh"
#include "../../include/propainter_b200.h"

// UA_V_MN = 1: V tiles stay row-major (MN-major B operand in the SWIZZLE_128B_BASE32B layout, the only MN-major layout
// tcgen05 accepts for 32-bit operands -- profiles/probes/umma_probe.cu variants 5/6), copied by cp.async like K: no
// staging buffer, no transposing pass.  V then reaches the tensor core un-rounded (hardware truncation to TF32).
// Measured on B200 with three stages (gpurun_out exp_vmn3, round 2): 110 / 215 us vs 135 / 312 us for the transposing
// two-stage build (5/16, 16/16 windows masked), within 2.6e-4 / 2.3e-3 of the mma.sync kernel (the transposing build:
// 2.1e-4 / 1.9e-3) -- the default since round 2; UA_V_MN=0 keeps the transposing producer for comparison.
#ifndef UA_V_MN
#define UA_V_MN 1
#endif
#define UA_BM 128
#define UA_BN 64
#define UA_THREADS 416                             // 4 softmax + 8 producer (two groups on alternate tiles) + 1 MMA warp
#define UA_K_BYTES (UA_BN * 128 * 4)               // 32 KB: 4 k-blocks of [64 rows x 128 B]
#define UA_V_BYTES (128 * UA_BN * 4)               // 32 KB: 2 k-blocks of [128 rows x 128 B]  (V^T: rows = head dim)
#define UA_STAGE_BYTES (UA_K_BYTES + UA_V_BYTES)
#define UA_LDSTG 132                               // V staging rows: 128 floats + 4 pad (conflict-free LDS.128 by key)
#ifndef UA_STAGES
#define UA_STAGES (UA_V_MN ? 3 : 2)                // K/V stages; 3 fit only without the V staging buffers (UA_V_MN=1)
#endif
#if UA_V_MN
#define UA_STG_BYTES 0                             // no V staging buffers
#else
#define UA_STG_BYTES (UA_BN * UA_LDSTG * 4)
#endif
#define UA_TAIL_BYTES 3072                          // barriers, TMEM base (+128), token table (<= 224 ints at +256), key row pointers (+1152)
#define UA_SMEM_BYTES (UA_STAGES * UA_STAGE_BYTES + 2 * UA_STG_BYTES + UA_TAIL_BYTES + 1024)   // staging: one per producer group
static_assert(UA_STAGES >= 2 && UA_STAGES <= 3 && UA_SMEM_BYTES <= 227 * 1024, "stage count does not fit the 227 KB of shared memory");

// MN-major SWIZZLE_128B_BASE32B descriptor for the row-major V tile: LBO = stride between 32-dim blocks (UA_BN rows x 128 B),
// SBO = 512 B between 4-row swizzle atoms, layout type 1
__device__ __forceinline__ uint64_t ua_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((UA_BN * 128) >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of the 16-byte chunk c4 (dims 4*c4..) of key row `key` in that tile: 32-byte granule g stored at g ^ (key & 3)
__device__ __forceinline__ uint32_t ua_off_mn(int key, int c4) {
  const int nb = c4 >> 3, c8 = c4 & 7;
  return (uint32_t)(nb * (UA_BN * 128) + key * 128 + ((((c8 >> 1) ^ (key & 3))) << 5) + (c8 & 1) * 16);
}
__global__ void __launch_bounds__(UA_THREADS, 1) k_sparse_attn_umma(PPAttnParams p) {
  extern __shared__ __align__(1024) uint8_t ua_raw[];
  const int win = blockIdx.z, head = blockIdx.y;
  if (p.flags[win] == 0) return;
  // SWIZZLE_128B tiles need 1 KB alignment; the pad is applied as an offset so the pointer stays in the shared window
  uint8_t* base = ua_raw + ((1024u - (ua_smem(ua_raw) & 1023u)) & 1023u);
  uint8_t* stg_base = base + UA_STAGES * UA_STAGE_BYTES;             // per-group V staging (row-major, padded)
  uint8_t* tail = stg_base + 2 * UA_STG_BYTES;
  // barriers: 0-2 kv_full[3]  3-5 kv_empty[3]  6,7 s_full[2]  8,9 p_full[2]  10 pv_done  11 q_ready
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tail);
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(tail + 128);
  int* stab = reinterpret_cast<int*>(tail + 256);                      // this window's token table, staged once
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
    for (int i = 0; i < UA_STAGES; ++i) { ua_bar_init(bar(i), 128); ua_bar_init(bar(3 + i), 1); }   // kv_full: producer threads; kv_empty: commit
    ua_bar_init(bar(6), 1); ua_bar_init(bar(7), 1);            // s_full: tcgen05.commit
    ua_bar_init(bar(8), 128); ua_bar_init(bar(9), 128);        // p_full: every softmax thread arrives
    ua_bar_init(bar(10), 1);                                   // pv_done: tcgen05.commit
    ua_bar_init(bar(11), 128);                                 // q_ready: softmax threads wrote Q into TMEM
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ua_smem(tmem_base_p)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < p.NKO; i += UA_THREADS) stab[i] = ktab[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tmem_base_p;
  const uint32_t tO = tbase, tS0 = tbase + 128, tP0 = tbase + 256, tQ = tbase + 384;   // O 128 | S 2x64 | P 2x64 | Q 128 columns

  if (warp >= 4 && warp < 12) {
    // ================================================= producers: K tile (cp.async) + V^T tile (transposing stores)
    // two groups of 4 warps fill alternate tiles, so one group's global-load latency hides behind the other's stores
    const int grp = (warp - 4) >> 2, ptid = (tid - 128) & 127;
    for (int j = grp; j < ntiles; j += 2) {
      const int s = j % UA_STAGES, use = j / UA_STAGES;
      ua_bar_wait(bar(3 + s), (use & 1) ^ 1);                            // stage free (passes immediately on first use)
      uint8_t* sK = base + s * UA_STAGE_BYTES;
      uint8_t* sV = sK + UA_K_BYTES;
      // row pointers of the tile's 64 keys (K part; V follows at +C in token rows and pooled rows alike)
      const float** kp = reinterpret_cast<const float**>(tail + 1152) + s * UA_BN;        // per stage
      if (ptid < UA_BN) {
        const int jk = j * UA_BN + ptid;
        const float* src = nullptr;
        if (jk < nkeys) {
          const int kfi = jk / keys_per_frame, slot = jk - kfi * keys_per_frame, fr = p.kf_start + kfi * p.kf_step;
          src = slot < p.NKO ? p.qkv + ((long)fr * p.NT + stab[slot]) * p.ld_qkv + p.C + hoff
                             : p.pool + ((long)fr * p.NP + (slot - p.NKO)) * p.ld_pool + hoff;
        }
        kp[ptid] = src;
      }
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");       // this producer group only
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      // K and V rows: a warp copies whole 512-byte rows (lane <-> 16-byte chunk, 4 cache lines per instruction);
      // K goes straight into its swizzled tile, V into this group's row-major staging buffer
      float* stg = reinterpret_cast<float*>(stg_base + grp * UA_STG_BYTES);
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int kk = (ptid >> 5) + 4 * i, c4 = ptid & 31;
        const float* src = kp[kk];
        uint8_t* dk = sK + ua_off(kk, c4 * 4, UA_BN);
#if UA_V_MN
        float* dv = reinterpret_cast<float*>(sV + ua_off_mn(kk, c4));
#else
        float* dv = stg + kk * UA_LDSTG + c4 * 4;
#endif
        if (src) { pp_cp_async16(dk, src + c4 * 4); pp_cp_async16(dv, src + p.C + c4 * 4); }
        else {
          *reinterpret_cast<float4*>(dk) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(dv) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      pp_cp_async_commit();
      pp_cp_async_wait<0>();
#if !UA_V_MN
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");       // every thread's copies have landed
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      // V^T: thread <-> (key, half of the head dims).  LDS.128 by key is conflict-free (row stride 132 floats); for a
      // fixed dim the 32 lanes (= 32 consecutive keys) store to 32 distinct banks of the swizzled tile
      {
        const int key = ptid & 63, half = ptid >> 6;
        const float* vr = stg + key * UA_LDSTG + half * 64;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(vr + i * 4);
          const int d = half * 64 + i * 4;
          *reinterpret_cast<uint32_t*>(sV + ua_off(d + 0, key, 128)) = pp_tf32(v.x);
          *reinterpret_cast<uint32_t*>(sV + ua_off(d + 1, key, 128)) = pp_tf32(v.y);
          *reinterpret_cast<uint32_t*>(sV + ua_off(d + 2, key, 128)) = pp_tf32(v.z);
          *reinterpret_cast<uint32_t*>(sV + ua_off(d + 3, key, 128)) = pp_tf32(v.w);
        }
      }
#endif
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05.mma
      ua_bar_arrive(bar(s));
    }
  } else if (warp == 12) {
    // ================================================= MMA issuer: the warp stays converged, one elected lane issues.
    // (Inside a divergent `if (lane == 0)` region ptxas wraps every tcgen05 instruction in an ELECT / BRA.U.ANY loop and
    // rebuilds its descriptors through R2UR moves: ~100 issue cycles per MMA, measured with the conv kernel's cycle
    // counters -- more than the 32-64 cycles the MMA itself takes.  Descriptors are built once per tile and advanced by
    // adding to their 16-byte address field.)
    {
      const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UA_BN >> 3) << 17) | ((uint32_t)(UA_BM >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(UA_BM >> 4) << 24) |
                               (UA_V_MN ? (1u << 16) : 0u);               // bit 16: B operand MN-major
      auto issue_s = [&](int j) {                                        // caller has observed kv_full for tile j
        const int sb = j & 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t kd = ua_desc(ua_smem(base + (j % UA_STAGES) * UA_STAGE_BYTES));
        if (ua_elect()) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)                                 // 128 head dims = 4 k-blocks x 4 k-steps of 8; A = Q from TMEM
            ua_mma_ts(tS0 + sb * UA_BN, tQ + ks * 8, kd + (uint64_t)(((ks >> 2) * (UA_BN * 128) + (ks & 3) * 32) >> 4), idesc_s, ks > 0);
          ua_commit(bar(6 + sb));                                        // S_j ready for the softmax warps
        }
        __syncwarp();
      };
      auto issue_pv = [&](int j) {
        const int sb = j & 1, st = j % UA_STAGES;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t vaddr = ua_smem(base + st * UA_STAGE_BYTES + UA_K_BYTES);
#if UA_V_MN
        const uint64_t vd = ua_desc_mn(vaddr);
#else
        const uint64_t vd = ua_desc(vaddr);
#endif
        if (ua_elect()) {
#pragma unroll
          for (int ks = 0; ks < UA_BN / 8; ++ks)                          // 64 keys = 2 k-blocks x 4 k-steps
#if UA_V_MN
            ua_mma_ts(tO, tP0 + sb * UA_BN + ks * 8, vd + (uint64_t)((ks * 1024) >> 4), idesc_o, (j > 0 || ks > 0));   // 8 key rows x 128 B
#else
            ua_mma_ts(tO, tP0 + sb * UA_BN + ks * 8, vd + (uint64_t)(((ks >> 2) * (128 * 128) + (ks & 3) * 32) >> 4), idesc_o,
                      (j > 0 || ks > 0));
#endif
          ua_commit(bar(3 + st));                                        // stage may be refilled
          ua_commit(bar(10));                                            // O holds tiles 0..j
        }
        __syncwarp();
      };
      ua_bar_wait(bar(11), 0);                                           // Q is in TMEM
      ua_bar_wait(bar(0), 0);
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        // issue whichever is ready first: S_{j+1} (needs K/V tile j+1; its S buffer is free since P_{j-1} was consumed)
        // or P_j.V_j (needs the softmax warps' P_j, and O rescaled if that was required)
        bool s_done = j + 1 >= ntiles, pv_done = false;
        for (long spin = 0; !(s_done && pv_done); ++spin) {
          if (!s_done && __any_sync(0xffffffffu, ua_bar_test(bar((j + 1) % UA_STAGES), (((j + 1) / UA_STAGES) & 1)))) { issue_s(j + 1); s_done = true; }
          if (!pv_done && __any_sync(0xffffffffu, ua_bar_test(bar(8 + (j & 1)), (j >> 1) & 1))) { issue_pv(j); pv_done = true; }
          if (spin > (1L << 26)) __trap();
        }
      }
    }
  } else {
    // ================================================= softmax / correction / epilogue: thread <-> query row <-> TMEM lane
    const int row = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float m_used = -INFINITY, l = 0.f;
    {   // this thread's query row -> TMEM lane `row`, columns tQ..tQ+127 (A operand of S = Q K^T), pre-scaled into the log2 domain
      const float* qrow = nullptr;
      if (row < nq) { const int qi = q0 + row, fr = qi / p.WN; qrow = p.qkv + ((long)fr * p.NT + stab[qi - fr * p.WN]) * p.ld_qkv + hoff; }
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t qv[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 v = qrow ? *reinterpret_cast<const float4*>(qrow + c0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          // round-to-nearest TF32 here: the tensor core would otherwise truncate the fp32 mantissa
          qv[c] = pp_tf32(v.x * p.scale_log2); qv[c + 1] = pp_tf32(v.y * p.scale_log2);
          qv[c + 2] = pp_tf32(v.z * p.scale_log2); qv[c + 3] = pp_tf32(v.w * p.scale_log2);
        }
        UA_ST32(tQ + c0 + lane_off, qv);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ua_bar_arrive(bar(11));
    }
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      ua_bar_wait(bar(6 + s), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v0[32], v1[32];
      UA_LD32(tS0 + s * UA_BN + lane_off, v0);
      UA_LD32(tS0 + s * UA_BN + 32 + lane_off, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int kbase = j * UA_BN;
      float mx = -INFINITY;
      if (kbase + UA_BN > nkeys) {                                        // ragged last tile: mask the missing keys
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (kbase + c >= nkeys) v0[c] = 0xff800000u;
          if (kbase + 32 + c >= nkeys) v1[c] = 0xff800000u;
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[c]), __uint_as_float(v1[c])));
      // lazy rescaling: keep the running reference maximum unless some row grew by more than 2^8
      const bool grow = mx > m_used + 8.0f;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_used;
        const float alpha = ua_ex2(m_used - m_new);                       // m_used = -inf on the first tile -> alpha = 0
        if (j > 0) {
          ua_bar_wait(bar(10), (j - 1) & 1);                              // PV_{j-1} finished: O is quiescent
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
        const float a = ua_ex2(__uint_as_float(v0[c]) - m_used), b = ua_ex2(__uint_as_float(v1[c]) - m_used);
        rs += a + b;
        v0[c] = pp_tf32(a); v1[c] = pp_tf32(b);
      }
      l += rs;
      UA_ST32(tP0 + s * UA_BN + lane_off, v0);
      UA_ST32(tP0 + s * UA_BN + 32 + lane_off, v1);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ua_bar_arrive(bar(8 + s));
    }
    // ---- epilogue: O / l -> global
    ua_bar_wait(bar(10), (ntiles - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = 1.0f / l;
    float* orow = nullptr;
    if (row < nq) { const int qi = q0 + row, fr = qi / p.WN; orow = p.out + ((long)fr * p.NT + stab[qi - fr * p.WN]) * p.ld_out + hoff; }
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
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// masked windows on tcgen05; called by pp_sparse_window_attn (mma_kernels.cu)
int pp_launch_sparse_attn_umma(const PPAttnParams& p, int n_windows, cudaStream_t stream) {
  if (p.NKO > 224) return PP_ERR_SHAPE;
  if (cudaFuncSetAttribute(k_sparse_attn_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, UA_SMEM_BYTES) != cudaSuccess)
    return PP_ERR_LAUNCH;
  dim3 grid((p.t * p.WN + UA_BM - 1) / UA_BM, p.C / 128, n_windows);
  k_sparse_attn_umma<<<grid, UA_THREADS, UA_SMEM_BYTES, stream>>>(p);
  return cudaPeekAtLastError() == cudaSuccess ? PP_OK : PP_ERR_LAUNCH;
}
