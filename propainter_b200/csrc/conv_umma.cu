// Implicit-GEMM stride-1 "same" convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces the cuDNN convolutions of the two recurrent propagation scans -- the offset nets and backbones of
// BidirectionalPropagation (model/propainter.py:42-50,86-96,121-175; model/recurrent_flow_completion.py:17-29,60-116) --
// and, as a 1x1 conv over the sampled columns, the GEMM of torchvision.ops.deform_conv2d (model/propainter.py:67-69,
// model/recurrent_flow_completion.py:42-44), together with everything that followed each of them as separate launches:
// bias add, LeakyReLU / ReLU / sigmoid / tanh, residual add, final ReLU, placement into a channel slice of a concat
// buffer, and the torch.cat that built the conv input (inputs are given as up to 4 channel segments).
//
//   out[p][n] = post( act( sum_{seg,c,dy,dx} W[n][seg,c,dy,dx] * x_seg[p + (dy,dx)][c]  + bias[n] + pre[p][n] ) + res[p][n] )
//
// One CTA = one 128-pixel tile (BH x BW pixels of one map, BH*BW = 128 = UMMA M) x BN output channels.
//   * A operand: no im2col.  For every 32-channel block TMA lands KW shifted copies of the (BH+KH-1) x BW halo box
//     (4-D tiled tensor map, SWIZZLE_128B; out-of-range rows/columns/channels are zero-filled by the TMA unit = the conv's
//     zero padding and the channel padding to a multiple of 32).  A box is [(BH+KH-1)*BW rows][128 B] = exactly the K-major
//     SWIZZLE_128B layout tcgen05 wants, and tap row dy is the same box read from row dy*BW on: a shared-memory descriptor
//     whose start address moves by dy*BW*128 B (a multiple of the 1 KB swizzle atom), so one copy feeds KH taps.
//   * B operand: packed weights [Cout][K], K index = ((blk*KH + dy)*KW + dx)*32 + c, loaded as [BN x 32] K-major boxes.
//   * D: 128 lanes x BN fp32 columns in TMEM; kind::tf32 (weights are pre-rounded to TF32 at pack time; activations
//     written by this kernel are optionally rounded on store so the next conv's operands are round-to-nearest TF32 too).
// Warp roles: warps 0-3 epilogue (thread <-> TMEM lane <-> pixel), warp 4 TMA producer, warp 5 TMEM allocator + MMA issuer.
// Two rings: A (one slot per 32-channel block) and B (one slot per (block, dy) = KW taps).  mbarrier full/empty pairs,
// tcgen05.commit releases slots.  Descriptor / instruction encodings: pp_umma.cuh (validated on B200, round 1).
#include <cuda.h>
#include <stdlib.h>
#include "pp_elem.cuh"
#include "pp_mma.cuh"
#include "pp_umma.cuh"
#include "../../include/propainter_b200.h"

#define CV_THREADS 192
#define CV_MAX_A_SLOTS 6
#define CV_MAX_B_SLOTS 8
#define CV_SMEM_BUDGET (216 * 1024)

struct alignas(64) CVParams {
  CUtensorMap tmA[PP_CONV_MAX_SEG];
  CUtensorMap tmB;
  int seg_blocks[PP_CONV_MAX_SEG];   // pipeline blocks per segment (kgroup: groups of KW 32-channel blocks)
  int seg_kblocks[PP_CONV_MAX_SEG];  // real 32-channel blocks per segment (= K extent of the segment / 32 per tap)
  int nseg, nblk, kreal;      // nblk pipeline blocks; kreal real 32-channel blocks
  int n, H, W, KH, KW, BH, BW, BN, M;    // M = BH*BW = 128 or 64 (UMMA M)
  int kgroup;                 // 1x1 convs: the `KW` loop walks `KW` consecutive 32-channel blocks (one pipeline stage = KW blocks)
  int tiles_x, tiles_y;
  int na, nb;                 // ring depths
  int ring_bytes;             // A ring + B ring, at least what the epilogue staging tiles need (4 warps x BN/32 x 4.5 KB)
  int a_copy_bytes;           // (BH+KH-1)*BW*128
  int Cout;
  const float* bias; const float* pre; const float* res; float* out;
  int ld_pre, ld_res, ld_out;
  int act, post_relu, round_tf32;
  float slope;
#ifdef CV_PROFILE
  long long* prof;
#endif
};

#ifdef CV_PROFILE
// profiling variant (profiles/build_variant.py prof -DCV_PROFILE=1): per-CTA cycle attribution written to a caller buffer
static long long* g_cv_prof = nullptr;
extern "C" void pp_conv_profile_buffer(long long* p) { g_cv_prof = p; }
#define CV_CLK() clock64()
#define CV_PROF(i, v) do { if (p.prof) p.prof[(long)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (i)] = (v); } while (0)
#else
#define CV_CLK() 0LL
#define CV_PROF(i, v) do { } while (0)
#endif

__device__ __forceinline__ void cv_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cv_tma4(uint32_t dst, const CUtensorMap* tm, int c, int x, int y, int n, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tm), "r"(c), "r"(x), "r"(y), "r"(n), "r"(bar) : "memory");
}
__device__ __forceinline__ void cv_tma2(uint32_t dst, const CUtensorMap* tm, int k, int n, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(k), "r"(n), "r"(bar) : "memory");
}
__device__ __forceinline__ float cv_act(float v, int act, float slope) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : v * slope;
    case 3: return 1.0f / (1.0f + expf(-v));
    case 4: return tanhf(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(CV_THREADS, 1) k_conv_umma(const __grid_constant__ CVParams p) {
  extern __shared__ __align__(1024) uint8_t cv_raw[];
  uint8_t* base = cv_raw + ((1024u - (ua_smem(cv_raw) & 1023u)) & 1023u);
  const int a_slot_bytes = p.KW * p.a_copy_bytes;
  const int b_tap_bytes = p.BN * 128, b_slot_bytes = p.KW * b_tap_bytes;
  uint8_t* sA = base;
  uint8_t* sB = sA + p.na * a_slot_bytes;
  uint8_t* tail = base + p.ring_bytes;
  // barriers: [0,na) a_full  [8,8+na) a_empty  [16,16+nb) b_full  [24,24+nb) b_empty  32 acc_full
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tail);
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(tail + 40 * 8);
  const uint32_t b0 = ua_smem(bars);
  auto bar = [&](int i) { return b0 + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, n0 = blockIdx.y * p.BN;
  const int tpi = p.tiles_x * p.tiles_y;
  const int img = tile / tpi, trem = tile - img * tpi;
  const int y0 = (trem / p.tiles_x) * p.BH, x0 = (trem % p.tiles_x) * p.BW;
  const int tmem_cols = p.BN < 32 ? 32 : p.BN;

  if (tid == 0) {
    for (int i = 0; i < p.na; ++i) { ua_bar_init(bar(i), 1); ua_bar_init(bar(8 + i), 1); }
    for (int i = 0; i < p.nb; ++i) { ua_bar_init(bar(16 + i), 1); ua_bar_init(bar(24 + i), 1); }
    ua_bar_init(bar(32), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ua_smem(tmem_base_p)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tD = *tmem_base_p;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, parameter / descriptor fetch) touches no
  // global memory and overlaps the tail of the previous kernel in the stream; the next kernel may start its own prologue
  // now.  Global reads and writes below wait for the previous grid to have completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // The producer and MMA warps run their loops with all 32 lanes converged and hand single instructions to one elected
  // lane (elect.sync): TMA and tcgen05 instructions execute on the uniform datapath, and inside a divergent `if (lane == 0)`
  // region the compiler wraps every one of them in an ELECT / BRA.U.ANY loop and keeps the descriptor arithmetic in vector
  // registers (R2UR per operand) -- measured here at ~100 issue cycles per MMA against 32-64 cycles of tensor-pipe work.
  if (warp == 4) {
    // ================================================= TMA producer
    int seg = 0, cb = 0, kbase = 0;                                  // kbase: first real k-block of the current segment
    long long wa = 0, wb = 0, t1;
    (void)wa; (void)wb; (void)t1;
    CV_PROF(0, CV_CLK());
    for (int blk = 0; blk < p.nblk; ++blk) {
      const int sa = blk % p.na;
      t1 = CV_CLK();
      ua_bar_wait(bar(8 + sa), ((blk / p.na) & 1) ^ 1);
      wa += CV_CLK() - t1;
      if (ua_elect()) {
        cv_expect_tx(bar(sa), (uint32_t)a_slot_bytes);
        for (int dx = 0; dx < p.KW; ++dx)
          if (p.kgroup)
            cv_tma4(ua_smem(sA + sa * a_slot_bytes + dx * p.a_copy_bytes), &p.tmA[seg], (cb * p.KW + dx) * 32, x0, y0, img, bar(sa));
          else
            cv_tma4(ua_smem(sA + sa * a_slot_bytes + dx * p.a_copy_bytes), &p.tmA[seg], cb * 32, x0 + dx - p.KW / 2,
                    y0 - p.KH / 2, img, bar(sa));
      }
      __syncwarp();
      for (int dy = 0; dy < p.KH; ++dy) {
        const int ib = blk * p.KH + dy, sb = ib % p.nb;
        t1 = CV_CLK();
        ua_bar_wait(bar(24 + sb), ((ib / p.nb) & 1) ^ 1);
        wb += CV_CLK() - t1;
        if (ua_elect()) {
          cv_expect_tx(bar(16 + sb), (uint32_t)b_slot_bytes);
          for (int dx = 0; dx < p.KW; ++dx)
            cv_tma2(ua_smem(sB + sb * b_slot_bytes + dx * b_tap_bytes), &p.tmB,
                    (p.kgroup ? kbase + cb * p.KW + dx : ib * p.KW + dx) * 32, n0, bar(16 + sb));
        }
        __syncwarp();
      }
      if (++cb == p.seg_blocks[seg]) { cb = 0; kbase += p.seg_kblocks[seg]; ++seg; }
    }
    if (lane == 0) { CV_PROF(1, wa); CV_PROF(2, wb); CV_PROF(3, CV_CLK()); }
  } else if (warp == 5) {
    // ================================================= MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
    uint32_t acc = 0;
    long long wa = 0, wb = 0, t1, tfirst = 0;
    (void)wa; (void)wb; (void)t1; (void)tfirst;
    if (lane == 0) CV_PROF(4, CV_CLK());
    for (int blk = 0; blk < p.nblk; ++blk) {
      const int sa = blk % p.na;
      t1 = CV_CLK();
      ua_bar_wait(bar(sa), (blk / p.na) & 1);
      wa += CV_CLK() - t1;
      if (blk == 0) tfirst = CV_CLK();
      const uint64_t a_desc0 = ua_desc(ua_smem(sA + sa * a_slot_bytes));
      for (int dy = 0; dy < p.KH; ++dy) {
        const int ib = blk * p.KH + dy, sb = ib % p.nb;
        t1 = CV_CLK();
        ua_bar_wait(bar(16 + sb), (ib / p.nb) & 1);
        wb += CV_CLK() - t1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t b_desc0 = ua_desc(ua_smem(sB + sb * b_slot_bytes));
        if (ua_elect()) {
          // descriptors advance in their 16-byte address field: +2 per 8-float k-step, + tap / row offsets >> 4
          uint64_t ad = a_desc0 + (uint64_t)((dy * p.BW * 128) >> 4), bd = b_desc0;
          for (int dx = 0; dx < p.KW; ++dx) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              ua_mma_ss(tD, ad + 2 * ks, bd + 2 * ks, idesc, acc);
              acc = 1;
            }
            ad += (uint64_t)(p.a_copy_bytes >> 4); bd += (uint64_t)(b_tap_bytes >> 4);
          }
          ua_commit(bar(24 + sb));
          if (dy == p.KH - 1) ua_commit(bar(8 + sa));
          if (dy == p.KH - 1 && blk == p.nblk - 1) ua_commit(bar(32));
        }
        acc = 1;
        __syncwarp();
      }
    }
    if (lane == 0) { CV_PROF(5, wa); CV_PROF(6, wb); CV_PROF(7, tfirst); CV_PROF(8, CV_CLK()); }
  } else {
    // ================================================= epilogue.  tcgen05.ld hands every thread one accumulator row (TMEM lane =
    // pixel) x 32 columns; storing that way makes each warp instruction touch 32 different 128-byte lines (measured: ~2.2 k
    // cycles of LSU wavefronts per 32 columns, and the same again for each of pre / res).  The rows therefore go through a
    // 32 x 36-float staging tile per warp (the operand ring is idle by now) and the bias / pre / activation / residual math
    // runs in the transposed mapping lane <-> (row = 4i + lane/8, 4 columns = lane%8): 8 lanes cover one pixel's 128 bytes,
    // so every global load and store instruction moves four full lines.
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int rsub = lane >> 3, c4 = lane & 7;
    const float* __restrict__ bias = p.bias;
    const int act = p.act, post_relu = p.post_relu, round_tf32 = p.round_tf32;
    const float slope = p.slope;
    // M = 128: warp w's 32 TMEM lanes hold accumulator rows 32w .. 32w+31; M = 64: rows 16w .. 16w+15 in lanes 0-15 (the
    // "half subpartition" layout of cta_group::1 M=64 accumulators, cute/atom/mma_traits_sm100.hpp), lanes 16-31 unused.
    // The row pointers of this lane's 8 rows (row = 4i + lane/8) are built here, while the MMAs still run: the epilogue is
    // one warp per scheduler, so every instruction of its dependent address arithmetic costs ~4 cycles of latency, and
    // computing them per 32-column chunk (~1000 instructions with the 64-bit index math) was 2.4 k cycles per chunk.
    const int rpw = p.M >> 2, bw_shift = p.BW == 16 ? 4 : 3;
    const float* prow[8]; const float* rrow[8]; float* orow[8];
    unsigned okmask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = 4 * i + rsub, r = warp * rpw + rl;
      const int y = y0 + (r >> bw_shift), x = x0 + (r & (p.BW - 1));
      const bool ok = rl < rpw && y < p.H && x < p.W;
      const long pix = ok ? ((long)img * p.H + y) * p.W + x : 0;
      okmask |= ok ? (1u << i) : 0u;
      orow[i] = p.out + pix * p.ld_out + n0 + 4 * c4;
      prow[i] = p.pre ? p.pre + pix * p.ld_pre + n0 + 4 * c4 : nullptr;
      rrow[i] = p.res ? p.res + pix * p.ld_res + n0 + 4 * c4 : nullptr;
      asm volatile("" : "+l"(orow[i]), "+l"(prow[i]), "+l"(rrow[i]));        // keep them materialised here (no sinking into the loop)
    }
    const bool has_pre = p.pre != nullptr, has_res = p.res != nullptr;
    const int nchunk = p.BN >> 5;
    float* stgw = reinterpret_cast<float*>(sA) + warp * nchunk * (32 * 36);      // this warp's staging tiles, one per 32 columns
    ua_bar_wait(bar(32), 0);
    if (tid == 0) CV_PROF(9, CV_CLK());
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // Phase A: all TMEM loads first.  tcgen05.wait::ld also waits for the thread's outstanding global stores (measured with
    // profiles/probes/umma_rate_probe.cu: 55 cycles with nothing in flight, 150-900 right after a burst of STG), so a
    // load -> store -> load -> store sequence pays one store round trip per 32 columns.
    for (int ch = 0; ch < nchunk; ++ch) {
      uint32_t v[32];
      UA_LD32(tD + ch * 32 + lane_off, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* stg = stgw + ch * (32 * 36);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j) =
            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    }
    __syncwarp();
    if (tid == 0) CV_PROF(12, CV_CLK());
    // Phase B: bias / pre / activation / residual on the transposed mapping, coalesced loads and stores
    for (int ch = 0; ch < nchunk; ++ch) {
      const int n = n0 + ch * 32 + 4 * c4;
      const bool n_in = n < p.Cout;
      const float* stg = stgw + ch * (32 * 36);
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), pv[8], rv[8];
      if (bias && n_in) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
      const unsigned on = n_in ? okmask : 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        pv[i] = (has_pre && ((on >> i) & 1)) ? *reinterpret_cast<const float4*>(prow[i] + ch * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        rv[i] = (has_res && ((on >> i) & 1)) ? *reinterpret_cast<const float4*>(rrow[i] + ch * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // the activation is selected once per chunk (warp-uniform branch), not per element: with the switch inside the
      // element loop ptxas inlined the exp / tanh paths 32 times per chunk (~260 instructions between each shared-memory
      // load and its global store; 3.8 k cycles per chunk measured)
      // branch-free over the 8 rows (all shared-memory loads first, predicated stores last): with an `if (valid)` around
      // each row the compiler serialised load -> ~50 dependent instructions -> store eight times (2.4 k cycles per chunk
      // on four warps, measured), although the rows are independent
      auto finish = [&](auto actf) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * 36 + 4 * c4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i].x = actf(a[i].x + (bv.x + pv[i].x)) + rv[i].x; a[i].y = actf(a[i].y + (bv.y + pv[i].y)) + rv[i].y;
          a[i].z = actf(a[i].z + (bv.z + pv[i].z)) + rv[i].z; a[i].w = actf(a[i].w + (bv.w + pv[i].w)) + rv[i].w;
          if (post_relu) { a[i].x = fmaxf(a[i].x, 0.f); a[i].y = fmaxf(a[i].y, 0.f); a[i].z = fmaxf(a[i].z, 0.f); a[i].w = fmaxf(a[i].w, 0.f); }
          if (round_tf32) {
            a[i].x = __uint_as_float(pp_tf32(a[i].x)); a[i].y = __uint_as_float(pp_tf32(a[i].y));
            a[i].z = __uint_as_float(pp_tf32(a[i].z)); a[i].w = __uint_as_float(pp_tf32(a[i].w));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if ((on >> i) & 1) *reinterpret_cast<float4*>(orow[i] + ch * 32) = a[i];
      };
      if (act == 0) finish([](float v) { return v; });
      else if (act == 1) finish([](float v) { return fmaxf(v, 0.f); });
      else if (act == 2) finish([slope](float v) { return v > 0.f ? v : v * slope; });
      else if (act == 3) finish([](float v) { return 1.0f / (1.0f + expf(-v)); });
      else finish([](float v) { return tanhf(v); });
    }
    if (tid == 0) CV_PROF(14, CV_CLK());
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) CV_PROF(10, CV_CLK());
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tD), "r"(tmem_cols));
}

// launch with the programmatic-stream-serialization attribute (PDL); PP_PDL=0 in the environment falls back to plain launches
static bool cv_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("PP_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}
template <typename... KArgs, typename... Args>
static cudaError_t cv_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = cv_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

typedef CUresult (*PFN_cvEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_cvEncodeTiled cv_encoder() {
  static PFN_cvEncodeTiled cached = nullptr;      // idempotent lookup (same pointer every time): benign if raced
  if (cached) return cached;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  cached = (PFN_cvEncodeTiled)fn;
  return cached;
}

// tile / ring plan shared by the launcher and pp_conv2d_umma_plan (so callers and tests can see what will run)
static int cv_plan(const PPConvParams* q, CVParams* p, int* smem_bytes) {
  if (q->nseg < 1 || q->nseg > PP_CONV_MAX_SEG || q->n < 1 || q->H < 1 || q->W < 1) return PP_ERR_SHAPE;
  if (q->KH < 1 || q->KW < 1 || q->KH > 7 || q->KW > 7 || !(q->KH & 1) || !(q->KW & 1)) return PP_ERR_SHAPE;
  if (q->Cout < 4 || q->Cout % 4) return PP_ERR_SHAPE;
  int nblk = 0, kblocks = 0;
  for (int s = 0; s < q->nseg; ++s) {
    if (q->seg[s].C < 1) return PP_ERR_SHAPE;
    if (q->seg[s].ld % 4 || ((uintptr_t)q->seg[s].x & 15)) return PP_ERR_ALIGN;
    p->seg_kblocks[s] = (q->seg[s].C + 31) / 32;
    kblocks += p->seg_kblocks[s];
  }
  // 1x1 convs (plain GEMMs over the channels): a pipeline stage of one 32-channel block holds only 4 MMAs, and the fixed
  // cost of a stage (barrier round trip, tcgen05 fence, commits: ~300-900 cycles measured) then exceeds the MMAs' own
  // ~290 cycles.  Group 4 consecutive blocks per stage (the tap loop walks channels instead of x-shifts).
  const int kv = (q->KH == 1 && q->KW == 1 && kblocks >= 8) ? 4 : 0;
  p->kgroup = kv ? 1 : 0;
  for (int s = 0; s < q->nseg; ++s) {
    p->seg_blocks[s] = kv ? (p->seg_kblocks[s] + kv - 1) / kv : p->seg_kblocks[s];
    nblk += p->seg_blocks[s];
  }
  if (q->ld_out % 4 || ((uintptr_t)q->out & 15) || ((uintptr_t)q->w_packed & 15)) return PP_ERR_ALIGN;
  if (q->bias && ((uintptr_t)q->bias & 15)) return PP_ERR_ALIGN;
  if (q->pre && (q->ld_pre % 4 || ((uintptr_t)q->pre & 15))) return PP_ERR_ALIGN;
  if (q->res && (q->ld_res % 4 || ((uintptr_t)q->res & 15))) return PP_ERR_ALIGN;
  p->nseg = q->nseg; p->nblk = nblk;
  p->n = q->n; p->H = q->H; p->W = q->W; p->KH = q->KH; p->KW = kv ? kv : q->KW;
  p->kreal = kblocks;
  // tile shape: (M/8) x 8 or (M/16) x 16 pixels, whichever wastes fewer padded pixels (ties: 8 columns)
  int bw = q->tile_w;
  if (bw != 8 && bw != 16) {
    const int mm = q->tile_m == 64 ? 64 : 128;
    const long a8 = (long)((q->W + 7) / 8) * ((q->H + mm / 8 - 1) / (mm / 8)), a16 = (long)((q->W + 15) / 16) * ((q->H + mm / 16 - 1) / (mm / 16));
    bw = a16 < a8 ? 16 : 8;
  }
  const int M = q->tile_m == 64 ? 64 : 128;                      // UMMA M (pixels per CTA); 64 halves the per-MMA A traffic and the CTA's work
  p->M = M;
  p->BW = bw; p->BH = M / bw;
  p->tiles_x = (q->W + p->BW - 1) / p->BW; p->tiles_y = (q->H + p->BH - 1) / p->BH;
  const long tiles = (long)p->tiles_x * p->tiles_y * q->n;
  if (tiles > 0x7fffffffL) return PP_ERR_SHAPE;
  // N tile.  Measured with the CV_PROFILE counters (profiles/conv_prof.py): one M=128 kind::tf32 MMA with both operands in
  // shared memory costs ~85 + 0.23*N cycles (the 128 x 32-byte A slices are read at ~48 B/clk whatever N is), and a CTA
  // runs K/8 of them back to back; CTAs beyond one per SM run as further waves.  Pick the N that minimises
  // waves x (85 + 0.23 N); ties go to the larger tile (fewer re-reads of A from L2).
  int bn = q->bn;
  if (bn != 32 && bn != 64 && bn != 128) {
    long best = -1;
    for (int cand = 128; cand >= 32; cand >>= 1) {
      const long ctas = tiles * ((q->Cout + cand - 1) / cand), waves = (ctas + PP_NUM_SMS - 1) / PP_NUM_SMS;
      const long cost = waves * (850 + 23 * cand / 10);
      if (best < 0 || cost < best) { best = cost; bn = cand; }
    }
  }
  p->BN = bn;
  p->a_copy_bytes = (p->BH + p->KH - 1) * p->BW * 128;
  const int a_slot = p->KW * p->a_copy_bytes, b_slot = p->KW * bn * 128;
  int na = (q->KH * q->KW == 1 && !p->kgroup) ? 4 : 2;
  if (na > nblk) na = nblk;
  if (na * a_slot + b_slot > CV_SMEM_BUDGET) na = 1;
  if (na * a_slot + b_slot > CV_SMEM_BUDGET) return PP_ERR_SHAPE;
  int nb = (CV_SMEM_BUDGET - na * a_slot) / b_slot;
  if (nb > CV_MAX_B_SLOTS) nb = CV_MAX_B_SLOTS;
  if (nb > nblk * q->KH) nb = nblk * q->KH;
  // spend what is left on more A slots (1x1 convs: deeper prefetch of the only large operand)
  while (na < CV_MAX_A_SLOTS && na < nblk && (na + 1) * a_slot + nb * b_slot <= CV_SMEM_BUDGET) ++na;
  p->na = na; p->nb = nb;
  p->ring_bytes = na * a_slot + nb * b_slot;
  if (p->ring_bytes < 4 * (bn / 32) * 32 * 36 * 4) p->ring_bytes = 4 * (bn / 32) * 32 * 36 * 4;   // epilogue staging tiles
  *smem_bytes = p->ring_bytes + 512 + 1024;
  p->Cout = q->Cout;
  p->bias = q->bias; p->pre = q->pre; p->res = q->res; p->out = q->out;
  p->ld_pre = q->ld_pre; p->ld_res = q->ld_res; p->ld_out = q->ld_out;
  p->act = q->act; p->post_relu = q->post_relu; p->round_tf32 = q->round_tf32; p->slope = q->slope;
  return PP_OK;
}

extern "C" int pp_conv2d_umma_plan(const PPConvParams* q, int* tile_h, int* tile_w, int* bn, int* ctas, int* smem_bytes) {
  CVParams p;
  int smem = 0;
  const int rc = cv_plan(q, &p, &smem);
  if (rc != PP_OK) return rc;
  if (tile_h) *tile_h = p.BH;
  if (tile_w) *tile_w = p.BW;
  if (bn) *bn = p.BN;
  if (ctas) *ctas = p.tiles_x * p.tiles_y * p.n * ((p.Cout + p.BN - 1) / p.BN);
  if (smem_bytes) *smem_bytes = smem;
  return PP_OK;
}

extern "C" int pp_conv2d_umma(const PPConvParams* q, cudaStream_t stream) {
  CVParams p;
  int smem = 0;
  int rc = cv_plan(q, &p, &smem);
  if (rc != PP_OK) return rc;
  PFN_cvEncodeTiled enc = cv_encoder();
  if (!enc) return PP_ERR_LAUNCH;
  for (int s = 0; s < q->nseg; ++s) {
    const cuuint64_t ld = (cuuint64_t)q->seg[s].ld;
    cuuint64_t dims[4] = {(cuuint64_t)q->seg[s].C, (cuuint64_t)q->W, (cuuint64_t)q->H, (cuuint64_t)q->n};
    cuuint64_t strides[3] = {ld * 4, ld * 4 * (cuuint64_t)q->W, ld * 4 * (cuuint64_t)q->W * (cuuint64_t)q->H};
    cuuint32_t box[4] = {32, (cuuint32_t)p.BW, (cuuint32_t)(p.BH + p.KH - 1), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&p.tmA[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)q->seg[s].x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return PP_ERR_LAUNCH;
  }
  {
    const cuuint64_t ktot = (cuuint64_t)p.kreal * q->KH * q->KW * 32;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)q->Cout};
    cuuint64_t strides[1] = {ktot * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)p.BN};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)q->w_packed, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return PP_ERR_LAUNCH;
  }
  if (cudaFuncSetAttribute(k_conv_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_BUDGET + 2048) != cudaSuccess)
    return PP_ERR_LAUNCH;
#ifdef CV_PROFILE
  p.prof = g_cv_prof;
#endif
  dim3 grid((unsigned)(p.tiles_x * p.tiles_y * p.n), (unsigned)((p.Cout + p.BN - 1) / p.BN));
  if (cv_launch(k_conv_umma, grid, dim3(CV_THREADS), (size_t)smem, stream, p) != cudaSuccess) return PP_ERR_LAUNCH;
  return cudaPeekAtLastError() == cudaSuccess ? PP_OK : PP_ERR_LAUNCH;
}

// ================================================================ deformable sampling -> columns
// First half of torchvision.ops.deform_conv2d (3x3, stride 1, pad 1, 16 offset groups) as called from
// DeformableAlignment.forward / SecondOrderDeformableAlignment.forward (model/propainter.py:57-69,
// model/recurrent_flow_completion.py:31-44): decode the raw conv_offset output (max_res*tanh offsets (+ flow.flip),
// sigmoid modulation), sample x bilinearly at the 9 x 16 positions of every pixel and write the modulated samples as
// columns cols[p][k*Cin + c] (rounded to TF32: they are the A operand of the GEMM that follows = pp_conv2d_umma with a
// 1x1 kernel over `cols`).  One warp per pixel: lane <-> (group, half of the group's channels), so per tap a warp reads
// 16 positions x 4 corners x 32/64 B and writes one contiguous Cin*4-byte run.
template <int CPL>   // channels per lane: 4 (Cin = 128) or 8 (Cin = 256)
__global__ void __launch_bounds__(256) k_deform_gather(const float* __restrict__ x, int ld_x, const float* __restrict__ x2, int ld_x2,
    const float* __restrict__ o, int ld_o,
    const float* __restrict__ obias, const float* __restrict__ flow, float max_res, float* __restrict__ cols, long npix, int H, int W) {
  // One warp per pixel, lane <-> (offset group g, half of the group's channels).  The 27 offset-net outputs of (pixel, g)
  // are fetched up front, then the 9 taps run in batches of 3 with all 12 (24) corner loads of a batch in flight before
  // the first one is used: the kernel is latency-bound (two dependent memory round trips per tap), so what matters is
  // how many independent loads each warp keeps outstanding.
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");               // inputs come from the previous kernel in the stream (PDL)
  const long pix = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int lane = threadIdx.x & 31, g = lane >> 1, half = lane & 1;
  const long HW = (long)H * W, img = pix / HW, pim = pix - img * HW;
  const int y = (int)(pim / W), xx = (int)(pim - (long)y * W);
  const float* op = o + pix * ld_o;
  float2 off[9]; float ml[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { off[k] = *reinterpret_cast<const float2*>(op + g * 18 + 2 * k); ml[k] = op[288 + g * 9 + k]; }
  float fy = 0.f, fx = 0.f;
  if (flow) { fx = flow[2 * pix]; fy = flow[2 * pix + 1]; }
  if (obias) {
#pragma unroll
    for (int k = 0; k < 9; ++k) { off[k].x += obias[g * 18 + 2 * k]; off[k].y += obias[g * 18 + 2 * k + 1]; ml[k] += obias[288 + g * 9 + k]; }
  }
  constexpr int CIN = CPL * 32, NV = CPL / 4;
  const int c = g * (2 * CPL) + half * CPL;
  // x2 != NULL: the input channels are split over two maps of CIN/2 channels each (offset groups 0-7 | 8-15): the two
  // previous states of the second-order scan live in different slots of the history buffer (no torch.cat)
  const bool second = x2 != nullptr && c >= CIN / 2;
  if (second) { x = x2; ld_x = ld_x2; }
  const int cs = second ? c - CIN / 2 : c;
  const float* xi = x + img * HW * ld_x;
  float* dst = cols + pix * (9L * CIN) + c;
#pragma unroll
  for (int kb = 0; kb < 9; kb += 3) {
    float wts[3][4];
    float4 v[3][4][NV];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int k = kb + u;
      PPDTap t;
      t.py = (float)(y - 1 + k / 3) + (max_res * tanhf(off[k].x) + fy);      // same association as pp_deform_tap (pp_elem.cuh)
      t.px = (float)(xx - 1 + k % 3) + (max_res * tanhf(off[k].y) + fx);
      t.m = 1.0f / (1.0f + expf(-ml[k]));
      const PPDW d = pp_deform_weights(t, H, W);
      wts[u][0] = d.w00; wts[u][1] = d.w01; wts[u][2] = d.w10; wts[u][3] = d.w11;
      const float* p00 = xi + ((long)d.y0 * W + d.x0) * ld_x + cs;
      const float* q[4] = {p00, p00 + ld_x, p00 + (long)W * ld_x, p00 + (long)W * ld_x + ld_x};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* a = wts[u][j] != 0.f ? q[j] : xi + cs;                // never dereference an out-of-image corner
#pragma unroll
        for (int i = 0; i < NV; ++i) v[u][j][i] = *reinterpret_cast<const float4*>(a + 4 * i);
      }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float w = wts[u][j];
          acc.x += v[u][j][i].x * w; acc.y += v[u][j][i].y * w; acc.z += v[u][j][i].z * w; acc.w += v[u][j][i].w * w;
        }
        *reinterpret_cast<float4*>(dst + (long)(kb + u) * CIN + 4 * i) =
            make_float4(__uint_as_float(pp_tf32(acc.x)), __uint_as_float(pp_tf32(acc.y)), __uint_as_float(pp_tf32(acc.z)), __uint_as_float(pp_tf32(acc.w)));
      }
    }
  }
}

extern "C" int pp_deform_gather(const float* x, int ld_x, const float* x2, int ld_x2, const float* o, int ld_o, const float* o_bias,
                                const float* flow, float max_res, float* cols, int n, int H, int W, int Cin, cudaStream_t stream) {
  if ((Cin != 128 && Cin != 256) || n < 1 || H < 1 || W < 1) return PP_ERR_SHAPE;
  if (ld_x % 4 || ld_o < 432 || ((uintptr_t)x & 15) || ((uintptr_t)cols & 15)) return PP_ERR_ALIGN;
  if (x2 && (ld_x2 % 4 || ((uintptr_t)x2 & 15))) return PP_ERR_ALIGN;
  if (ld_o % 2 || ((uintptr_t)o & 7)) return PP_ERR_ALIGN;           // (dy,dx) pairs are fetched as float2
  const long npix = (long)n * H * W;
  const long blocks = (npix + 7) / 8;
  if (blocks > 0x7fffffffL) return PP_ERR_SHAPE;
  cudaError_t e;
  if (Cin == 128) e = cv_launch(k_deform_gather<4>, dim3((unsigned)blocks), dim3(256), 0, stream, x, ld_x, x2, ld_x2, o, ld_o, o_bias, flow, max_res, cols, npix, H, W);
  else e = cv_launch(k_deform_gather<8>, dim3((unsigned)blocks), dim3(256), 0, stream, x, ld_x, x2, ld_x2, o, ld_o, o_bias, flow, max_res, cols, npix, H, W);
  if (e != cudaSuccess) return PP_ERR_LAUNCH;
  return cudaPeekAtLastError() == cudaSuccess ? PP_OK : PP_ERR_LAUNCH;
}
