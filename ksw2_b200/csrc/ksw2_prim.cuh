// ksw2_prim.cuh -- packed-lane primitives of the B200 wavefront engine.
//
// Arithmetic model.  The reference computes in int8 with wrap-around and mixes signed and unsigned
// max/min (ksw2_extz2_sse.c:26-47).  sm_100a has single-instruction 16x2 SIMD integer ops
// (VIADD.16x2, VIMNMX.{S,U}16x2, VIMNMX3.S16x2, VIADDMNMX.S16x2) but only emulated 8x4 ones, so a
// 32-bit register ("pk") carries TWO DP lanes, each as   int8 value << 8   in a 16-bit half with a zero
// low byte.  In that fixed-point form 16-bit wrap-around IS int8 wrap-around, signed/unsigned 16-bit
// order IS signed/unsigned int8 order, so every reference operation maps to one native instruction
// and stays bit-exact for all inputs (no range assumptions, no guards).
//
// The same source builds for the device (intrinsics) and for the host (plain C++), the latter only
// for the test-side simulator (tests/sim/); the product never runs the host build.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KS_HD __host__ __device__ __forceinline__
#else
#define KS_HD static inline
#endif

typedef uint32_t pk;   // two lanes: lo half = block lane i, hi half = block lane i+8, each (int8 << 8)

#if defined(__CUDA_ARCH__)
KS_HD pk add2(pk a, pk b)            { return __vadd2(a, b); }
KS_HD pk maxs2(pk a, pk b)           { return __vmaxs2(a, b); }
KS_HD pk mins2(pk a, pk b)           { return __vmins2(a, b); }
KS_HD pk maxu2(pk a, pk b)           { return __vmaxu2(a, b); }
KS_HD pk minu2(pk a, pk b)           { return __vminu2(a, b); }
KS_HD pk max3s2(pk a, pk b, pk c)    { return __vimax3_s16x2(a, b, c); }
KS_HD pk addmaxs2(pk a, pk b, pk c)  { return __vmaxs2(__vadd2(a, b), c); }   // ptxas fuses: VIADDMNMX.S16x2
KS_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)   // raw PRMT: __byte_perm() masks the selector with 0x7777 and loses the sign-replicate bit
{ uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s)); return r; }
KS_HD uint32_t fshr16(uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); } // (hi:lo) >> 16
#else
KS_HD pk add2(pk a, pk b)  { return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16); }
KS_HD int16_t h_lo(pk a)   { return (int16_t)(a & 0xffffu); }
KS_HD int16_t h_hi(pk a)   { return (int16_t)(a >> 16); }
KS_HD pk mk2(uint32_t lo, uint32_t hi) { return (lo & 0xffffu) | (hi << 16); }
KS_HD pk maxs2(pk a, pk b) { return mk2((uint16_t)(h_lo(a) > h_lo(b) ? h_lo(a) : h_lo(b)), (uint16_t)(h_hi(a) > h_hi(b) ? h_hi(a) : h_hi(b))); }
KS_HD pk mins2(pk a, pk b) { return mk2((uint16_t)(h_lo(a) < h_lo(b) ? h_lo(a) : h_lo(b)), (uint16_t)(h_hi(a) < h_hi(b) ? h_hi(a) : h_hi(b))); }
KS_HD pk maxu2(pk a, pk b) { uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16; return mk2(al > bl ? al : bl, ah > bh ? ah : bh); }
KS_HD pk minu2(pk a, pk b) { uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16; return mk2(al < bl ? al : bl, ah < bh ? ah : bh); }
KS_HD pk max3s2(pk a, pk b, pk c)   { return maxs2(maxs2(a, b), c); }
KS_HD pk addmaxs2(pk a, pk b, pk c) { return maxs2(add2(a, b), c); }
KS_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{   // PTX prmt.b32, default mode: selector nibble n picks byte (n&7) of {b,a}; n&8 replicates its sign bit
	uint64_t v = ((uint64_t)b << 32) | a; uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		uint32_t n = (s >> (4 * i)) & 0xf, byte = (uint32_t)(v >> (8 * (n & 7))) & 0xff;
		if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
KS_HD uint32_t fshr16(uint32_t lo, uint32_t hi) { return (lo >> 16) | (hi << 16); }
#endif

// ---- helpers on the (int8<<8)x2 form -----------------------------------------------------------
KS_HD pk  rep2(int v)                 { uint32_t b = ((uint32_t)v & 0xffu) << 8; return b | (b << 16); }  // both lanes = int8(v)
#if defined(__CUDACC__)
__constant__ int ks_opaque_m1 = -1, ks_opaque_p1 = 1, ks_opaque_z = 0;
__constant__ uint32_t ks_opaque_c40 = 0x40404040u;
#endif
#if defined(__CUDA_ARCH__) && !defined(KS_NO_IMAD_TRICKS)
// The packed 16x2 ops, LOP3 and PRMT all issue on the ALU pipe (one warp-instruction per 2 cycles and scheduler), which is the
// busiest unit of the fill kernels; IMAD goes to the otherwise idle FMA pipe.  ~a == a * (-1) + (-1) in 32-bit arithmetic, and
// with the multiplier read from constant memory ptxas cannot fold it back into a LOP3.
KS_HD pk  not2(pk a)                  { const int m = ks_opaque_m1; return (pk)((int)a * m + m); }
#else
KS_HD pk  not2(pk a)                  { return ~a; }                         // per lane: -a - 1/256 (low byte 0xff)
#endif
#define KS_ONE1 0x00010001u                                                  // +1/256 per lane: completes a two's complement
// z + 1/256 per lane for a z with zero low bytes (completes the two's complement of a later "+ ~x")
#if defined(__CUDA_ARCH__) && !defined(KS_NO_IMAD_TRICKS)
KS_HD pk  plus_one2(pk z)             { const int o = ks_opaque_p1; return (pk)((int)z * o + (int)KS_ONE1); }   // IMAD: FMA pipe
#else
KS_HD pk  plus_one2(pk z)             { return z | KS_ONE1; }
#endif
// max(a + b, 0) per lane: the zero comes from constant memory (as an instruction operand); a literal 0 makes ptxas build a zero
// register with a PRMT in front of every VIADDMNMX
#if defined(__CUDA_ARCH__) && !defined(KS_NO_IMAD_TRICKS)
KS_HD pk  addmax0s2(pk a, pk b)       { return addmaxs2(a, b, (pk)ks_opaque_z); }
// (c & (t | q)) | (~c & (t ^ q)) in ONE LOP3 (c = 0x40 in every byte, kept opaque so that ptxas does not split it into three)
KS_HD uint32_t ks_class_bits(uint32_t t, uint32_t q)
{ uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xBC;" : "=r"(r) : "r"(t), "r"(q), "r"(ks_opaque_c40)); return r; }
#else
KS_HD pk  addmax0s2(pk a, pk b)       { return addmaxs2(a, b, 0u); }
KS_HD uint32_t ks_class_bits(uint32_t t, uint32_t q) { const uint32_t c = 0x40404040u; return ((t ^ q) & ~c) | ((t | q) & c); }
#endif
// a - b, exact per lane:  a + ~b + 1
KS_HD pk  sub2(pk a, pk b)            { return add2(add2(a, not2(b)), KS_ONE1); }
// extract lane value: half h (0 lo / 1 hi)
#if defined(__CUDA_ARCH__)
KS_HD int lane_s(pk a, int h)         { return h ? ((int32_t)a >> 24) : (int)prmt(a, 0u, 0x9991u); }       // signed int8 (PRMT sign-replicate)
KS_HD int lane_u(pk a, int h)         { return h ? (int)(a >> 24) : (int)prmt(a, 0u, 0x4441u); }           // unsigned byte
#else
KS_HD int lane_s(pk a, int h)         { return h ? ((int32_t)a >> 24) : ((int32_t)(a << 16) >> 24); }      // signed int8
KS_HD int lane_u(pk a, int h)         { return h ? (int)(a >> 24) : (int)((a >> 8) & 0xffu); }             // unsigned byte
#endif
KS_HD pk  set_lane(pk a, int h, int v){ uint32_t b = ((uint32_t)v & 0xffu) << 8; return h ? ((a & 0x0000ffffu) | (b << 16)) : ((a & 0xffff0000u) | b); }
// per-lane select: m has 0xffff in lanes taken from a, 0 in lanes taken from b
KS_HD pk  sel2(pk m, pk a, pk b)      { return (a & m) | (b & ~m); }
// 1 (i.e. 0x0100 per lane) where lane != 0
KS_HD pk  nz_one2(pk a)               { return minu2(a, 0x01000100u); }

// block lane j (0..15)  <->  (register j&7, half j>>3)
#define KS_REG(j)  ((j) & 7)
#define KS_HALF(j) ((j) >> 3)
