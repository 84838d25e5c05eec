// kgx_field.cuh -- secp256k1 field arithmetic for sm_100a, 8 x 32-bit limbs, written from scratch.
//
// Replaces the reference's GPU/GPUMath.h (ModSub256 :476-516, _ModMult :810-907, _ModSqr :909-1019).
// The reference works on 4 x u64 limbs with 64-bit PTX multiplies that ptxas lowers to chains of
// 32x32 IMAD.WIDE.U32; here the 32-bit structure is explicit: every carry chain is ONE asm block of
// mad.lo.cc / madc.hi.cc pairs on an even/odd column split, which ptxas fuses into
// IMAD.WIDE.U32.X with predicate carries (65 IMAD.WIDE + 16 IADD3 for the 256x256->512 product).
//
// Result conventions (bit-exact with IntMod.cpp:873-942):
//   fe_mul / fe_sqr : R1 = lo + hi*0x1000003D1 (exact, 290 bits); out = (R1 mod 2^256 + (R1>>256)*0x1000003D1) mod 2^256
//                     -- no final conditional subtraction, last carry dropped.
//   fe_sub          : a - b mod 2^256, plus p if the subtraction borrowed.
#pragma once
#include <cstdint>

typedef uint32_t u32;
typedef uint64_t u64;

#define KGX_C 0x3D1u  // 0x1000003D1 = 2^32 + 0x3D1

// acc[0..7] += {a0,a2,a4,a6} * b at 64-bit aligned column pairs, carry out added into acc[8].
__device__ __forceinline__ void kgx_mad_row(u32* acc, u32 a0, u32 a2, u32 a4, u32 a6, u32 b) {
  asm("mad.lo.cc.u32  %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32       %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(acc[8])
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}

// 512 -> 256 fold, see header comment. w[16] little-endian 32-bit words.
__device__ __forceinline__ void kgx_fold(u32* r, const u32* w) {
  u32 r0, r1, r2, r3, r4, r5, r6, r7, r8, r9;
  const u32 c = KGX_C;
  // r[0..8] = w[0..7] + {h0,h2,h4,h6} * c   (h = w[8..15])
  asm("mad.lo.cc.u32  %0, %9, %13, %14;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %15;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %16;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %17;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %18;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %19;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %20;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %21;\n\t"
      "addc.u32       %8, 0, 0;"
      : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3), "=&r"(r4), "=&r"(r5), "=&r"(r6), "=&r"(r7), "=&r"(r8)
      : "r"(w[8]), "r"(w[10]), "r"(w[12]), "r"(w[14]), "r"(c), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
        "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]));
  // r[1..9] += {h1,h3,h5,h7} * c
  asm("mad.lo.cc.u32  %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32       %8, 0, 0;"
      : "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "=r"(r9)
      : "r"(w[9]), "r"(w[11]), "r"(w[13]), "r"(w[15]), "r"(c));
  // r[1..9] += h   (the 2^32 part of 0x1000003D1)
  asm("add.cc.u32  %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.cc.u32 %7, %7, %16;\n\t"
      "addc.u32    %8, %8, 0;"
      : "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "+r"(r9)
      : "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]));
  // second fold: top = r8 + r9*2^32 ; V = top*(2^32+c) = r8*c + (r8 + r9*c)*2^32 + r9*2^64 ; r[0..7] += V (carry dropped)
  u32 v0, v1, v2;
  asm("{\n\t"
      ".reg .u32 ulo, uhi;\n\t"
      "mul.lo.u32     %0, %3, %5;\n\t"
      "mul.hi.u32     %1, %3, %5;\n\t"
      "mad.lo.cc.u32  ulo, %4, %5, %3;\n\t"
      "addc.u32       uhi, 0, 0;\n\t"
      "add.cc.u32     %1, %1, ulo;\n\t"
      "addc.u32       %2, %4, uhi;\n\t"
      "}"
      : "=&r"(v0), "=&r"(v1), "=&r"(v2)
      : "r"(r8), "r"(r9), "r"(c));
  asm("add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, 0;"
      : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
      : "r"(v0), "r"(v1), "r"(v2));
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3; r[4] = r4; r[5] = r5; r[6] = r6; r[7] = r7;
}

// w[0..15] = E + (O << 32) where E[i] sits at word i and O[i] at word i+1.
__device__ __forceinline__ void kgx_merge_eo(u32* w, const u32* E, const u32* O) {
  w[0] = E[0];
  asm("add.cc.u32  %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32    %14, %29, %44;"
      : "=&r"(w[1]), "=&r"(w[2]), "=&r"(w[3]), "=&r"(w[4]), "=&r"(w[5]), "=&r"(w[6]), "=&r"(w[7]), "=&r"(w[8]),
        "=&r"(w[9]), "=&r"(w[10]), "=&r"(w[11]), "=&r"(w[12]), "=&r"(w[13]), "=&r"(w[14]), "=&r"(w[15])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
        "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
        "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]),
        "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
}

// 256 x 256 -> 512 schoolbook on the even/odd split.
__device__ __forceinline__ void kgx_mul512(u32* w, const u32* a, const u32* b) {
  u32 E[17], O[15];
#pragma unroll
  for (int i = 0; i < 17; i++) E[i] = 0;
#pragma unroll
  for (int i = 0; i < 15; i++) O[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    kgx_mad_row(E + i, a[0], a[2], a[4], a[6], b[i]);          // row i,   even j -> even columns
    kgx_mad_row(O + i, a[1], a[3], a[5], a[7], b[i]);          // row i,   odd  j -> odd columns
    kgx_mad_row(E + i + 2, a[1], a[3], a[5], a[7], b[i + 1]);  // row i+1, odd  j -> even columns
    kgx_mad_row(O + i, a[0], a[2], a[4], a[6], b[i + 1]);      // row i+1, even j -> odd columns
  }
  kgx_merge_eo(w, E, O);
}

__device__ __forceinline__ void fe_mul(u32* r, const u32* a, const u32* b) {
  u32 w[16];
  kgx_mul512(w, a, b);
  kgx_fold(r, w);
}

// carry chains of 1..3 column pairs (kgx_mad_row is the 4-pair form): acc[0..2n-1] += {a...} * b, carry into acc[2n]
__device__ __forceinline__ void kgx_mad_c1(u32* acc, u32 b, u32 a0) {
  asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32       %2, %2, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(a0), "r"(b));
}
__device__ __forceinline__ void kgx_mad_c2(u32* acc, u32 b, u32 a0, u32 a1) {
  asm("mad.lo.cc.u32  %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32       %4, %4, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]) : "r"(a0), "r"(a1), "r"(b));
}
__device__ __forceinline__ void kgx_mad_c3(u32* acc, u32 b, u32 a0, u32 a1, u32 a2) {
  asm("mad.lo.cc.u32  %0, %7, %10, %0;\n\t"
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
      "addc.u32       %6, %6, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
      : "r"(a0), "r"(a1), "r"(a2), "r"(b));
}

// a^2 as 28 cross products (computed once, doubled) + 8 squares: 36 IMAD.WIDE instead of 64.
// Same exact 512-bit value as a*a, hence the same folded result (GPUMath.h:909-1019 / IntMod.cpp:1030-1234).
__device__ __forceinline__ void kgx_sqr512(u32* w, const u32* a) {
  u32 E[16], O[15];
#pragma unroll
  for (int i = 0; i < 16; i++) E[i] = 0;
#pragma unroll
  for (int i = 0; i < 15; i++) O[i] = 0;
  // E[i] sits at word i, O[i] at word i+1; product a_i*a_j lands at word i+j
  kgx_mad_row(O + 0, a[1], a[3], a[5], a[7], a[0]);   // 0x{1,3,5,7} -> words 1,3,5,7
  kgx_mad_c3(E + 2, a[0], a[2], a[4], a[6]);          // 0x{2,4,6}   -> words 2,4,6
  kgx_mad_c3(O + 2, a[1], a[2], a[4], a[6]);          // 1x{2,4,6}   -> words 3,5,7
  kgx_mad_c3(E + 4, a[1], a[3], a[5], a[7]);          // 1x{3,5,7}   -> words 4,6,8
  kgx_mad_c3(O + 4, a[2], a[3], a[5], a[7]);          // 2x{3,5,7}   -> words 5,7,9
  kgx_mad_c2(E + 6, a[2], a[4], a[6]);                // 2x{4,6}     -> words 6,8
  kgx_mad_c2(O + 6, a[3], a[4], a[6]);                // 3x{4,6}     -> words 7,9
  kgx_mad_c2(E + 8, a[3], a[5], a[7]);                // 3x{5,7}     -> words 8,10
  kgx_mad_c2(O + 8, a[4], a[5], a[7]);                // 4x{5,7}     -> words 9,11
  kgx_mad_c1(E + 10, a[4], a[6]);                     // 4x6         -> word 10
  kgx_mad_c1(O + 10, a[5], a[6]);                     // 5x6         -> word 11
  kgx_mad_c1(E + 12, a[5], a[7]);                     // 5x7         -> word 12
  kgx_mad_c1(O + 12, a[6], a[7]);                     // 6x7         -> word 13
  // C = E + (O << 32), words 1..15 (word 0 is zero)
  u32 c[16];
  c[0] = 0;
  asm("add.cc.u32  %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32    %14, %29, %44;"
      : "=&r"(c[1]), "=&r"(c[2]), "=&r"(c[3]), "=&r"(c[4]), "=&r"(c[5]), "=&r"(c[6]), "=&r"(c[7]), "=&r"(c[8]),
        "=&r"(c[9]), "=&r"(c[10]), "=&r"(c[11]), "=&r"(c[12]), "=&r"(c[13]), "=&r"(c[14]), "=&r"(c[15])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
        "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
        "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]),
        "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
  // 2C by funnel shifts (no carry chain)
  u32 c2[16];
  c2[0] = 0;
#pragma unroll
  for (int i = 15; i >= 1; i--) c2[i] = __funnelshift_l(c[i - 1], c[i], 1);
  // + squares a_i^2 at words (2i, 2i+1)
  asm("mad.lo.cc.u32  %0, %16, %16, %24;\n\t"
      "madc.hi.cc.u32 %1, %16, %16, %25;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %26;\n\t"
      "madc.hi.cc.u32 %3, %17, %17, %27;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %28;\n\t"
      "madc.hi.cc.u32 %5, %18, %18, %29;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %30;\n\t"
      "madc.hi.cc.u32 %7, %19, %19, %31;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %32;\n\t"
      "madc.hi.cc.u32 %9, %20, %20, %33;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %34;\n\t"
      "madc.hi.cc.u32 %11, %21, %21, %35;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %36;\n\t"
      "madc.hi.cc.u32 %13, %22, %22, %37;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %38;\n\t"
      "madc.hi.u32    %15, %23, %23, %39;"
      : "=&r"(w[0]), "=&r"(w[1]), "=&r"(w[2]), "=&r"(w[3]), "=&r"(w[4]), "=&r"(w[5]), "=&r"(w[6]), "=&r"(w[7]),
        "=&r"(w[8]), "=&r"(w[9]), "=&r"(w[10]), "=&r"(w[11]), "=&r"(w[12]), "=&r"(w[13]), "=&r"(w[14]), "=&r"(w[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(c2[0]), "r"(c2[1]), "r"(c2[2]), "r"(c2[3]), "r"(c2[4]), "r"(c2[5]), "r"(c2[6]), "r"(c2[7]),
        "r"(c2[8]), "r"(c2[9]), "r"(c2[10]), "r"(c2[11]), "r"(c2[12]), "r"(c2[13]), "r"(c2[14]), "r"(c2[15]));
}

__device__ __forceinline__ void fe_sqr(u32* r, const u32* a) {
  u32 w[16];
  kgx_sqr512(w, a);
  kgx_fold(r, w);
}

// r = a - b (mod 2^256), + p if borrow.   GPUMath.h:476-494 semantics.
__device__ __forceinline__ void fe_sub(u32* r, const u32* a, const u32* b) {
  u32 t0, t1, t2, t3, t4, t5, t6, t7, m;
  asm("sub.cc.u32  %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32    %8, 0, 0;"
      : "=&r"(t0), "=&r"(t1), "=&r"(t2), "=&r"(t3), "=&r"(t4), "=&r"(t5), "=&r"(t6), "=&r"(t7), "=&r"(m)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  // + p (mod 2^256)  ==  - 0x1000003D1
  u32 k0 = m & KGX_C, k1 = m & 1u;
  asm("sub.cc.u32  %0, %0, %8;\n\t"
      "subc.cc.u32 %1, %1, %9;\n\t"
      "subc.cc.u32 %2, %2, 0;\n\t"
      "subc.cc.u32 %3, %3, 0;\n\t"
      "subc.cc.u32 %4, %4, 0;\n\t"
      "subc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\t"
      "subc.u32    %7, %7, 0;"
      : "+r"(t0), "+r"(t1), "+r"(t2), "+r"(t3), "+r"(t4), "+r"(t5), "+r"(t6), "+r"(t7)
      : "r"(k0), "r"(k1));
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3; r[4] = t4; r[5] = t5; r[6] = t6; r[7] = t7;
}

// 128-bit distance accumulate (GPUMath.h:119-121 Add128): wraps silently.
__device__ __forceinline__ void d128_add(u32* d, u32 j0, u32 j1, u32 j2, u32 j3) {
  asm("add.cc.u32  %0, %0, %4;\n\t"
      "addc.cc.u32 %1, %1, %5;\n\t"
      "addc.cc.u32 %2, %2, %6;\n\t"
      "addc.u32    %3, %3, %7;"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(j0), "r"(j1), "r"(j2), "r"(j3));
}

__device__ __forceinline__ void fe_copy(u32* r, const u32* a) {
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = a[i];
}
__device__ __forceinline__ void fe_set_one(u32* r) {
  r[0] = 1;
#pragma unroll
  for (int i = 1; i < 8; i++) r[i] = 0;
}
