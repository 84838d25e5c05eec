// kgx_hosttest.cpp -- CPU build of the host/device-portable pieces of the engine (currently the
// safegcd modular inverse) so they can be unit-tested without a GPU (tests/test_abi_cpu.py).
// Not part of the product data path.
#include "kgx_modinv.h"
extern "C" void kgx_host_modinv(uint64_t out[4], const uint64_t in[4]) {
  uint32_t i32[8], o32[8];
  for (int i = 0; i < 4; i++) { i32[2 * i] = (uint32_t)in[i]; i32[2 * i + 1] = (uint32_t)(in[i] >> 32); }
  kgx::modinv256(o32, i32);
  for (int i = 0; i < 4; i++) out[i] = (uint64_t)o32[2 * i] | ((uint64_t)o32[2 * i + 1] << 32);
}
// number of 30-step batches used for a given input (statistics for DESIGN.md)
extern "C" int kgx_host_modinv_batches(const uint64_t in[4]) {
  using namespace kgx;
  const uint32_t M30 = 0xFFFFFFFFu >> 2;
  uint32_t w[8];
  for (int i = 0; i < 4; i++) { w[2 * i] = (uint32_t)in[i]; w[2 * i + 1] = (uint32_t)(in[i] >> 32); }
  s30 d = {{0}}, e = {{1}}, f = {{KGX_P0, KGX_P1, 0, 0, 0, 0, 0, 0, KGX_P8}}, g;
  g.v[0] = w[0] & M30; g.v[1] = ((w[0] >> 30) | (w[1] << 2)) & M30; g.v[2] = ((w[1] >> 28) | (w[2] << 4)) & M30;
  g.v[3] = ((w[2] >> 26) | (w[3] << 6)) & M30; g.v[4] = ((w[3] >> 24) | (w[4] << 8)) & M30; g.v[5] = ((w[4] >> 22) | (w[5] << 10)) & M30;
  g.v[6] = ((w[5] >> 20) | (w[6] << 12)) & M30; g.v[7] = ((w[6] >> 18) | (w[7] << 14)) & M30; g.v[8] = w[7] >> 16;
  int32_t eta = -1; int n = 0;
  for (int it = 0; it < 40; ++it) {
    t2x2 t; eta = divsteps_30_var(eta, (uint32_t)f.v[0], (uint32_t)g.v[0], &t);
    update_de_30(&d, &e, &t); update_fg_30<9>(&f, &g, &t); n++;
    int32_t c = 0; for (int j = 0; j < 9; j++) c |= g.v[j];
    if (!c) break;
  }
  return n;
}

// FP64-pipe multiplier (kgx_field_fp64.cuh): 512-bit product / square through the 52-bit-limb DFMA scheme, on the host
// with the FPU in round-toward-zero (the device uses fma.rz).  out: 8 x u64 limbs.
#include <cfenv>
#include "kgx_field_fp64.cuh"
extern "C" void kgx_host_mul512_fp64(uint64_t out[8], const uint64_t a[4], const uint64_t b[4], int square) {
  uint32_t a32[8], b32[8], w[16];
  for (int i = 0; i < 4; i++) { a32[2 * i] = (uint32_t)a[i]; a32[2 * i + 1] = (uint32_t)(a[i] >> 32); b32[2 * i] = (uint32_t)b[i]; b32[2 * i + 1] = (uint32_t)(b[i] >> 32); }
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  double da[5], db[5];
  kgx::fe_to_d52(da, a32);
  kgx::fe_to_d52(db, b32);
  if (square) kgx::kgx_sqr512_d52(w, da); else kgx::kgx_mul512_d52(w, da, db);
  fesetround(old);
  for (int i = 0; i < 8; i++) out[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
}
