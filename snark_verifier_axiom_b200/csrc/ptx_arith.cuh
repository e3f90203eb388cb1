// 32-bit add/sub/mad-with-carry primitives.
//
// Device: one PTX instruction each (`asm volatile` keeps the carry-flag order; ptxas fuses
// mad.lo.cc/madc.hi.cc pairs into IMAD.WIDE.U32.X with predicate carries -- checked with
// cuobjdump, see DESIGN.md "K0").
// Host:   bit-exact emulation with an explicit carry flag, so the very same limb algorithms can be
// unit-tested on a CPU-only box (tests/host_selftest.cpp).  This is NOT a CPU fallback of the
// product: no exported entry point ever runs field code on the host except constant folding in
// the protocol compiler.
#pragma once
#include <cstdint>

typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN inline __host__ __device__ __noinline__
#else
#define HD inline
#define HDN inline
#endif

#if !defined(__CUDACC__)
// host-only builds (tests/host): minimal stand-ins for the CUDA vector type used by the loaders
struct uint4 { u32 x, y, z, w; };
inline uint4 make_uint4(u32 x, u32 y, u32 z, u32 w) { return uint4{x, y, z, w}; }
#endif

namespace ptx {

#if defined(__CUDA_ARCH__)

HD u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 mul_lo(u32 a, u32 b) { u32 r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 mul_hi(u32 a, u32 b) { u32 r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD u32 madc_lo(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
HD u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

#else  // host emulation ---------------------------------------------------------------------

static thread_local u32 g_cy = 0;

HD u32 add_cc(u32 a, u32 b) { u64 t = (u64)a + b; g_cy = (u32)(t >> 32); return (u32)t; }
HD u32 addc_cc(u32 a, u32 b) { u64 t = (u64)a + b + g_cy; g_cy = (u32)(t >> 32); return (u32)t; }
HD u32 addc(u32 a, u32 b) { return a + b + g_cy; }
HD u32 sub_cc(u32 a, u32 b) { u64 t = (u64)a - b; g_cy = (u32)((t >> 32) & 1); return (u32)t; }  // g_cy = borrow
HD u32 subc_cc(u32 a, u32 b) { u64 t = (u64)a - b - g_cy; g_cy = (u32)((t >> 32) & 1); return (u32)t; }
HD u32 subc(u32 a, u32 b) { return a - b - g_cy; }
HD u32 mul_lo(u32 a, u32 b) { return a * b; }
HD u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
HD u32 mad_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)(a * b) + c; g_cy = (u32)(t >> 32); return (u32)t; }
HD u32 mad_hi_cc(u32 a, u32 b, u32 c) { u64 t = (u64)mul_hi(a, b) + c; g_cy = (u32)(t >> 32); return (u32)t; }
HD u32 madc_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)(a * b) + c + g_cy; g_cy = (u32)(t >> 32); return (u32)t; }
HD u32 madc_hi_cc(u32 a, u32 b, u32 c) { u64 t = (u64)mul_hi(a, b) + c + g_cy; g_cy = (u32)(t >> 32); return (u32)t; }
HD u32 madc_lo(u32 a, u32 b, u32 c) { return a * b + c + g_cy; }
HD u32 madc_hi(u32 a, u32 b, u32 c) { return mul_hi(a, b) + c + g_cy; }

#endif

}  // namespace ptx
