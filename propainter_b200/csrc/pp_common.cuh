// Common definitions for the propainter_b200 sm_100a kernels.
//
// Layout conventions:
//   * "planar"  : [n][c][H][W]   -- tensors that cross the reference API boundary (frames, flows, masks)
//   * "pixel-major" (NHWC) : [n][H][W][ld] with ld >= C -- every internal feature map; kernels take the
//     pixel stride `ld` explicitly so they can read / write channel slices of wider concat buffers.
//
// Element functions (per output element, no inter-thread cooperation) are PP_HD so that
// tests/hostsim can compile the very same index arithmetic for the CPU test-suite.  The product
// never runs them on the host.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(PP_HOSTSIM)
#define PP_HD static inline
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 r = {a, b, c, d}; return r; }
#else
#include <cuda_runtime.h>
#define PP_HD __host__ __device__ __forceinline__
#endif

// error codes of the C ABI (include/propainter_b200.h)
#define PP_OK 0
#define PP_ERR_SHAPE (-1)
#define PP_ERR_DTYPE (-2)
#define PP_ERR_WORKSPACE (-3)
#define PP_ERR_LAUNCH (-4)
#define PP_ERR_ALIGN (-5)

// Unfused fp32 arithmetic for the discontinuous paths (nearest rounding, thresholds): nvcc would
// otherwise contract a*b+c into FMA and move results across rounding boundaries relative to ATen.
#if defined(__CUDA_ARCH__)
#define PP_MUL(a, b) __fmul_rn((a), (b))
#define PP_ADD(a, b) __fadd_rn((a), (b))
#define PP_SUB(a, b) __fsub_rn((a), (b))
#define PP_DIV(a, b) __fdiv_rn((a), (b))
#else
#define PP_MUL(a, b) ((a) * (b))
#define PP_ADD(a, b) ((a) + (b))
#define PP_SUB(a, b) ((a) - (b))
#define PP_DIV(a, b) ((a) / (b))
#endif

#define PP_NUM_SMS 148
