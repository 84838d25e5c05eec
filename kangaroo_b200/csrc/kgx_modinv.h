// kgx_modinv.h -- modular inverse mod p = 2^256 - 0x1000003D1, variable time, host+device.
//
// Replaces the reference's GPU/GPUMath.h:700-803 (_ModInv, DRS62 "delayed right shift" divsteps).  This
// is an independent implementation of the Bernstein-Yang "safegcd" divstep iteration in batches of 62
// steps on signed 62-bit limbs (2x2 transition matrices applied to (f,g) and (d,e)); the result
// contract is the reference's: the canonical inverse in [0,p), and inv(0) = 0
// (GPUMath.h:785-801, IntMod.cpp:560-569).
//
// In the jump kernel every lane of the inverting warp runs this on the SAME value (the group product
// after the butterfly), so the data-dependent control flow is warp-uniform: no divergence.
// Compiles as plain C++ for the CPU unit test (tests/test_modinv_host.py via csrc/kgx_hosttest.cpp).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define KGX_HD __host__ __device__ __forceinline__
#else
#define KGX_HD static inline
#endif

namespace kgx {

typedef __int128 i128;
struct s62 { int64_t v[5]; };
struct t2x2 { int64_t u, v, q, r; };

KGX_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}

// 62 divsteps on the low words; returns new eta and the transition matrix scaled by 2^62.
KGX_HD int64_t divsteps_62_var(int64_t eta, uint64_t f0, uint64_t g0, t2x2* t) {
  uint64_t u = 1, v = 0, q = 0, r = 1;
  uint64_t f = f0, g = g0, m;
  uint32_t w;
  int i = 62, limit, zeros;
  for (;;) {
    zeros = ctz64(g | (~0ULL << i));
    g >>= zeros; u <<= zeros; v <<= zeros; eta -= zeros; i -= zeros;
    if (i == 0) break;
    if (eta < 0) {
      uint64_t tmp;
      eta = -eta;
      tmp = f; f = g; g = 0 - tmp;
      tmp = u; u = q; q = 0 - tmp;
      tmp = v; v = r; r = 0 - tmp;
      limit = ((int)eta + 1) > i ? i : ((int)eta + 1);
      m = (~0ULL >> (64 - limit)) & 63U;
      w = (uint32_t)((f * g * (f * f - 2)) & m);
    } else {
      limit = ((int)eta + 1) > i ? i : ((int)eta + 1);
      m = (~0ULL >> (64 - limit)) & 15U;
      w = (uint32_t)(f + (((f + 1) & 4) << 1));
      w = (uint32_t)((0 - (uint64_t)w) * g & m);
    }
    g += f * w; q += u * w; r += v * w;
  }
  t->u = (int64_t)u; t->v = (int64_t)v; t->q = (int64_t)q; t->r = (int64_t)r;
  return eta;
}

// p in signed-62 form: p = 256*2^248 - 0x1000003D1
#define KGX_P0 (-0x1000003D1LL)
#define KGX_P4 (256LL)
#define KGX_PINV62 0x27C7F6E22DDACACFULL  // p^-1 mod 2^62

// (d,e) <- t * (d,e) / 2^62 mod p
KGX_HD void update_de_62(s62* d, s62* e, const t2x2* t) {
  const uint64_t M62 = ~0ULL >> 2;
  const int64_t d0 = d->v[0], d1 = d->v[1], d2 = d->v[2], d3 = d->v[3], d4 = d->v[4];
  const int64_t e0 = e->v[0], e1 = e->v[1], e2 = e->v[2], e3 = e->v[3], e4 = e->v[4];
  const int64_t u = t->u, v = t->v, q = t->q, r = t->r;
  int64_t md, me, sd, se;
  i128 cd, ce;
  sd = d4 >> 63; se = e4 >> 63;
  md = (u & sd) + (v & se);
  me = (q & sd) + (r & se);
  cd = (i128)u * d0 + (i128)v * e0;
  ce = (i128)q * d0 + (i128)r * e0;
  md -= (int64_t)((KGX_PINV62 * (uint64_t)cd + (uint64_t)md) & M62);
  me -= (int64_t)((KGX_PINV62 * (uint64_t)ce + (uint64_t)me) & M62);
  cd += (i128)KGX_P0 * md;
  ce += (i128)KGX_P0 * me;
  cd >>= 62; ce >>= 62;
  cd += (i128)u * d1 + (i128)v * e1;
  ce += (i128)q * d1 + (i128)r * e1;
  d->v[0] = (int64_t)((uint64_t)cd & M62); cd >>= 62;
  e->v[0] = (int64_t)((uint64_t)ce & M62); ce >>= 62;
  cd += (i128)u * d2 + (i128)v * e2;
  ce += (i128)q * d2 + (i128)r * e2;
  d->v[1] = (int64_t)((uint64_t)cd & M62); cd >>= 62;
  e->v[1] = (int64_t)((uint64_t)ce & M62); ce >>= 62;
  cd += (i128)u * d3 + (i128)v * e3;
  ce += (i128)q * d3 + (i128)r * e3;
  d->v[2] = (int64_t)((uint64_t)cd & M62); cd >>= 62;
  e->v[2] = (int64_t)((uint64_t)ce & M62); ce >>= 62;
  cd += (i128)u * d4 + (i128)v * e4;
  ce += (i128)q * d4 + (i128)r * e4;
  cd += (i128)KGX_P4 * md;
  ce += (i128)KGX_P4 * me;
  d->v[3] = (int64_t)((uint64_t)cd & M62); cd >>= 62;
  e->v[3] = (int64_t)((uint64_t)ce & M62); ce >>= 62;
  d->v[4] = (int64_t)cd;
  e->v[4] = (int64_t)ce;
}

// (f,g) <- t * (f,g) / 2^62 on the first len limbs
KGX_HD void update_fg_62_var(int len, s62* f, s62* g, const t2x2* t) {
  const uint64_t M62 = ~0ULL >> 2;
  const int64_t u = t->u, v = t->v, q = t->q, r = t->r;
  int64_t fi, gi;
  i128 cf, cg;
  fi = f->v[0]; gi = g->v[0];
  cf = (i128)u * fi + (i128)v * gi;
  cg = (i128)q * fi + (i128)r * gi;
  cf >>= 62; cg >>= 62;
  for (int i = 1; i < len; ++i) {
    fi = f->v[i]; gi = g->v[i];
    cf += (i128)u * fi + (i128)v * gi;
    cg += (i128)q * fi + (i128)r * gi;
    f->v[i - 1] = (int64_t)((uint64_t)cf & M62); cf >>= 62;
    g->v[i - 1] = (int64_t)((uint64_t)cg & M62); cg >>= 62;
  }
  f->v[len - 1] = (int64_t)cf;
  g->v[len - 1] = (int64_t)cg;
}

// r <- r * sign(f) normalised into [0,p)
KGX_HD void normalize_62(s62* r, int64_t sign) {
  const int64_t M62 = (int64_t)(~0ULL >> 2);
  int64_t r0 = r->v[0], r1 = r->v[1], r2 = r->v[2], r3 = r->v[3], r4 = r->v[4];
  int64_t cond_add, cond_negate;
  cond_add = r4 >> 63;
  r0 += KGX_P0 & cond_add;
  r4 += KGX_P4 & cond_add;
  cond_negate = sign >> 63;
  r0 = (r0 ^ cond_negate) - cond_negate;
  r1 = (r1 ^ cond_negate) - cond_negate;
  r2 = (r2 ^ cond_negate) - cond_negate;
  r3 = (r3 ^ cond_negate) - cond_negate;
  r4 = (r4 ^ cond_negate) - cond_negate;
  r1 += r0 >> 62; r0 &= M62;
  r2 += r1 >> 62; r1 &= M62;
  r3 += r2 >> 62; r2 &= M62;
  r4 += r3 >> 62; r3 &= M62;
  cond_add = r4 >> 63;
  r0 += KGX_P0 & cond_add;
  r4 += KGX_P4 & cond_add;
  r1 += r0 >> 62; r0 &= M62;
  r2 += r1 >> 62; r1 &= M62;
  r3 += r2 >> 62; r2 &= M62;
  r4 += r3 >> 62; r3 &= M62;
  r->v[0] = r0; r->v[1] = r1; r->v[2] = r2; r->v[3] = r3; r->v[4] = r4;
}

// x: 4 x u64 little-endian (any value < 2^256); out: canonical inverse, 0 -> 0.
KGX_HD void modinv256(uint64_t out[4], const uint64_t in[4]) {
  const uint64_t M62 = ~0ULL >> 2;
  s62 d = {{0, 0, 0, 0, 0}}, e = {{1, 0, 0, 0, 0}};
  s62 f = {{KGX_P0, 0, 0, 0, KGX_P4}};
  s62 g;
  g.v[0] = (int64_t)(in[0] & M62);
  g.v[1] = (int64_t)(((in[0] >> 62) | (in[1] << 2)) & M62);
  g.v[2] = (int64_t)(((in[1] >> 60) | (in[2] << 4)) & M62);
  g.v[3] = (int64_t)(((in[2] >> 58) | (in[3] << 6)) & M62);
  g.v[4] = (int64_t)(in[3] >> 56);
  int len = 5;
  int64_t eta = -1;
  for (int it = 0; it < 24; ++it) {   // 12 batches bound 256-bit inputs; 24 is a safety net
    t2x2 t;
    eta = divsteps_62_var(eta, (uint64_t)f.v[0], (uint64_t)g.v[0], &t);
    update_de_62(&d, &e, &t);
    update_fg_62_var(len, &f, &g, &t);
    if (g.v[0] == 0) {
      int64_t cond = 0;
      for (int j = 1; j < len; ++j) cond |= g.v[j];
      if (cond == 0) break;
    }
    int64_t fn = f.v[len - 1], gn = g.v[len - 1];
    int64_t cond = ((int64_t)len - 2) >> 63;
    cond |= fn ^ (fn >> 63);
    cond |= gn ^ (gn >> 63);
    if (cond == 0) {
      f.v[len - 2] |= (int64_t)((uint64_t)fn << 62);
      g.v[len - 2] |= (int64_t)((uint64_t)gn << 62);
      --len;
    }
  }
  // gcd is |f| = 1 for invertible input; for input 0 (or a multiple of p) f = +-p and d = 0.
  normalize_62(&d, f.v[len - 1]);
  out[0] = (uint64_t)d.v[0] | ((uint64_t)d.v[1] << 62);
  out[1] = ((uint64_t)d.v[1] >> 2) | ((uint64_t)d.v[2] << 60);
  out[2] = ((uint64_t)d.v[2] >> 4) | ((uint64_t)d.v[3] << 58);
  out[3] = ((uint64_t)d.v[3] >> 6) | ((uint64_t)d.v[4] << 56);
}

}  // namespace kgx
