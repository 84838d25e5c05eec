// kgx_kernel.cuh -- the kangaroo jump kernel for sm_100a (B200).
//
// Replaces the reference's GPU/GPUCompute.h:22-117 (ComputeKangaroos) + comp_kangaroos (GPUEngine.cu:35-40).
// Same mathematics, different machine mapping:
//
//   reference : every thread owns 128 kangaroos in LOCAL memory (18.5 KB stack/thread), private Montgomery
//               chain per thread, one _ModInv per thread per jump -> ~416 B/jump of local-memory traffic.
//   here      : a CTA of 128 threads owns a TILE of 896 kangaroos whose whole state (x, y, d, running prefix
//               product = 112 B each) stays in SHARED MEMORY for all NB_RUN=64 jumps of a launch; HBM sees each
//               kangaroo once in and once out per launch (2.5 B/jump).  The Montgomery batch inverse spans
//               the whole tile: per-thread chains (7 kangaroos) -> per-lane chains across the 4 warps ->
//               XOR-butterfly product over the 32 lanes with warp shuffles -> ONE safegcd inverse per tile
//               per jump (warp-uniform, no divergence) -> back down the same tree.
//               Two CTAs are resident per SM so one tile's serial inverse overlaps the other's parallel phase.
//
//   The per-kangaroo pass fuses: back-substitution of this jump's inverse, the affine add, the distance
//   update, the DP test, and the *next* jump's dx / prefix product (accumulated in the opposite order), so
//   each jump is one sweep over shared memory: 5 ModMult + 1 ModSqr + 7 ModSub per kangaroo.
#pragma once
#include "kgx_field.cuh"
#include "kgx_modinv.h"

namespace kgx {

constexpr int T = 128;            // threads per CTA
constexpr int K = 7;              // kangaroos per thread
constexpr int TILE = T * K;       // kangaroos per tile
constexpr int CHUNKS = 5;         // 16-byte chunks per kangaroo in HBM: x0 x1 y0 y1 d
constexpr int JT_WORDS = 20 * 32; // jump table: jpx[8][32] jpy[8][32] jd[4][32], word-major (bank = jump index)

// shared memory carve-up (in uint4 units)
constexpr int S_X = 0;
constexpr int S_Y = S_X + K * 2 * T;
constexpr int S_P = S_Y + K * 2 * T;
constexpr int S_D = S_P + K * 2 * T;
constexpr int S_TOT = S_D + K * T;
constexpr int S_JT = S_TOT + 2 * T;                         // u32 view starts here
constexpr int SMEM_BYTES = S_JT * 16 + JT_WORDS * 4;        // 107,008 B -> 2 CTAs / SM

struct LaunchParams {
  uint4* state;          // [numTiles][K][CHUNKS][T]
  const u32* jtab;       // JT_WORDS words
  u32* out;              // DP slab: [count][maxFound * 14 words]
  u64 dpMask;
  u64 nKangaroos;        // real (unpadded) count: DPs of padding slots are dropped
  u32 numTiles;
  u32 maxFound;
  int nRun;
};

__device__ __forceinline__ void lds_fe(u32* r, const uint4* base, int idx0, int idx1) {
  uint4 a = base[idx0], b = base[idx1];
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void sts_fe(uint4* base, int idx0, int idx1, const u32* r) {
  base[idx0] = make_uint4(r[0], r[1], r[2], r[3]);
  base[idx1] = make_uint4(r[4], r[5], r[6], r[7]);
}
__device__ __forceinline__ void lds_jp(u32* r, const u32* tab, u32 j) {
#pragma unroll
  for (int w = 0; w < 8; w++) r[w] = tab[w * 32 + j];
}
__device__ __forceinline__ void shfl_xor_fe(u32* r, const u32* a, int mask) {
#pragma unroll
  for (int w = 0; w < 8; w++) r[w] = __shfl_xor_sync(0xffffffffu, a[w], mask);
}

// u32[8] <-> u64[4] for the inverse
__device__ __forceinline__ void fe_inv(u32* r, const u32* a) {
  uint64_t in[4], out[4];
#pragma unroll
  for (int i = 0; i < 4; i++) in[i] = (uint64_t)a[2 * i] | ((uint64_t)a[2 * i + 1] << 32);
  modinv256(out, in);
#pragma unroll
  for (int i = 0; i < 4; i++) { r[2 * i] = (u32)out[i]; r[2 * i + 1] = (u32)(out[i] >> 32); }
}

// Warp 0 only: turn the 128 per-thread products in sTot into their 128 inverses (in place).
__device__ __noinline__ void tile_inverse(uint4* sTot, int lane) {
  u32 t0[8], t1[8], t2[8], t3[8], c1[8], c2[8], v[8];
  lds_fe(t0, sTot, lane, T + lane);
  lds_fe(t1, sTot, 32 + lane, T + 32 + lane);
  lds_fe(t2, sTot, 64 + lane, T + 64 + lane);
  lds_fe(t3, sTot, 96 + lane, T + 96 + lane);
  fe_mul(c1, t0, t1);
  fe_mul(c2, c1, t2);
  fe_mul(v, c2, t3);
  // XOR butterfly: after level k every lane holds the product of its 2^(k+1)-lane block; keep the siblings.
  u32 sib[5][8];
#pragma unroll
  for (int l = 0; l < 5; l++) {
    shfl_xor_fe(sib[l], v, 1 << l);
    fe_mul(v, v, sib[l]);
  }
  u32 inv[8];
  fe_inv(inv, v);                 // identical in all 32 lanes
#pragma unroll
  for (int l = 4; l >= 0; l--) fe_mul(inv, inv, sib[l]);   // -> inverse of this lane's c3 = t0*t1*t2*t3
  u32 o[8];
  fe_mul(o, inv, c2);             // 1/t3
  sts_fe(sTot, 96 + lane, T + 96 + lane, o);
  fe_mul(inv, inv, t3);           // 1/(t0 t1 t2)
  fe_mul(o, inv, c1);             // 1/t2
  sts_fe(sTot, 64 + lane, T + 64 + lane, o);
  fe_mul(inv, inv, t2);           // 1/(t0 t1)
  fe_mul(o, inv, t0);             // 1/t1
  sts_fe(sTot, 32 + lane, T + 32 + lane, o);
  fe_mul(o, inv, t1);             // 1/t0
  sts_fe(sTot, lane, T + lane, o);
}

__global__ void __launch_bounds__(T, 2) jump_kernel(LaunchParams p) {
  extern __shared__ uint4 smem[];
  uint4* sX = smem + S_X;
  uint4* sY = smem + S_Y;
  uint4* sP = smem + S_P;
  uint4* sD = smem + S_D;
  uint4* sTot = smem + S_TOT;
  u32* sJ = reinterpret_cast<u32*>(smem + S_JT);
  const u32* jpx = sJ;
  const u32* jpy = sJ + 8 * 32;
  const u32* jd = sJ + 16 * 32;
  const int t = threadIdx.x;
  const int lane = t & 31;
  const u32 mlo = (u32)p.dpMask, mhi = (u32)(p.dpMask >> 32);

  for (int i = t; i < JT_WORDS; i += T) sJ[i] = p.jtab[i];

  for (u32 tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x) {
    __syncthreads();   // jump table visible; previous tile's shared state fully consumed
    uint4* gsrc = p.state + (size_t)tile * (K * CHUNKS * T);
#pragma unroll
    for (int g = 0; g < K; g++) {
      sX[(g * 2 + 0) * T + t] = gsrc[(g * CHUNKS + 0) * T + t];
      sX[(g * 2 + 1) * T + t] = gsrc[(g * CHUNKS + 1) * T + t];
      sY[(g * 2 + 0) * T + t] = gsrc[(g * CHUNKS + 2) * T + t];
      sY[(g * 2 + 1) * T + t] = gsrc[(g * CHUNKS + 3) * T + t];
      sD[g * T + t] = gsrc[(g * CHUNKS + 4) * T + t];
    }
    // prologue: forward chain of dx = x - jPx[x & 31]; sP[g] = product of the dx before g
    u32 P[8];
    {
      u32 x[8], jx[8], dx[8];
#pragma unroll 1
      for (int g = 0; g < K; g++) {
        lds_fe(x, sX, (g * 2) * T + t, (g * 2 + 1) * T + t);
        lds_jp(jx, jpx, x[0] & 31u);
        fe_sub(dx, x, jx);
        if (g == 0) {
          u32 one[8]; fe_set_one(one);
          sts_fe(sP, t, T + t, one);
          fe_copy(P, dx);
        } else {
          sts_fe(sP, (g * 2) * T + t, (g * 2 + 1) * T + t, P);
          fe_mul(P, P, dx);
        }
      }
      sts_fe(sTot, t, T + t, P);
    }
    int backward = 1;   // the pass after a forward accumulation consumes in reverse order

    for (int run = 0; run < p.nRun; run++) {
      __syncthreads();
      if (t < 32) tile_inverse(sTot, lane);
      __syncthreads();
      u32 I[8];
      lds_fe(I, sTot, t, T + t);
      const bool last = (run == p.nRun - 1);
#pragma unroll 1
      for (int i = 0; i < K; i++) {
        const int g = backward ? (K - 1 - i) : i;
        const int i0 = (g * 2) * T + t, i1 = (g * 2 + 1) * T + t;
        u32 x[8], y[8], jx[8], jy[8], dx[8], inv[8], s[8], rx[8], ry[8];
        lds_fe(x, sX, i0, i1);
        lds_fe(inv, sP, i0, i1);                 // prefix of this kangaroo
        const u32 j = x[0] & 31u;
        lds_jp(jx, jpx, j);
        fe_sub(dx, x, jx);
        fe_mul(inv, inv, I);                     // 1/dx
        if (i != K - 1) fe_mul(I, I, dx);        // strip this dx from the running inverse
        lds_fe(y, sY, i0, i1);
        lds_jp(jy, jpy, j);
        fe_sub(s, y, jy);                        // dy
        fe_mul(s, s, inv);                       // s = dy/dx
        fe_sqr(rx, s);                           // s^2
        fe_sub(rx, rx, jx);
        fe_sub(rx, rx, x);                       // rx = s^2 - jx - x
        fe_sub(ry, x, rx);
        fe_mul(ry, ry, s);
        fe_sub(ry, ry, y);                       // ry = s (x - rx) - y
        sts_fe(sX, i0, i1, rx);
        sts_fe(sY, i0, i1, ry);
        uint4 dv = sD[g * T + t];
        u32 d[4] = {dv.x, dv.y, dv.z, dv.w};
        d128_add(d, jd[j], jd[32 + j], jd[64 + j], jd[96 + j]);
        sD[g * T + t] = make_uint4(d[0], d[1], d[2], d[3]);
        if (((rx[7] & mhi) | (rx[6] & mlo)) == 0u) {          // GPUCompute.h:96
          const u64 kidx = (u64)tile * TILE + (u64)g * T + (u64)t;
          if (kidx < p.nKangaroos) {
            const u32 pos = atomicAdd(p.out, 1u);
            if (pos < p.maxFound) {                           // GPUMath.h:173-188 record layout
              u32* o = p.out + 1 + (size_t)pos * 14;
#pragma unroll
              for (int w = 0; w < 8; w++) o[w] = rx[w];
              o[8] = d[0]; o[9] = d[1]; o[10] = d[2]; o[11] = d[3];
              o[12] = (u32)kidx; o[13] = (u32)(kidx >> 32);
            }
          }
        }
        if (!last) {                             // next jump's dx and prefix, accumulated in THIS order
          lds_jp(jx, jpx, rx[0] & 31u);
          fe_sub(dx, rx, jx);
          if (i == 0) {
            u32 one[8]; fe_set_one(one);
            sts_fe(sP, i0, i1, one);
            fe_copy(P, dx);
          } else {
            sts_fe(sP, i0, i1, P);
            fe_mul(P, P, dx);
          }
        }
      }
      if (!last) sts_fe(sTot, t, T + t, P);
      backward ^= 1;
    }

    uint4* gdst = gsrc;
#pragma unroll
    for (int g = 0; g < K; g++) {
      gdst[(g * CHUNKS + 0) * T + t] = sX[(g * 2 + 0) * T + t];
      gdst[(g * CHUNKS + 1) * T + t] = sX[(g * 2 + 1) * T + t];
      gdst[(g * CHUNKS + 2) * T + t] = sY[(g * 2 + 0) * T + t];
      gdst[(g * CHUNKS + 3) * T + t] = sY[(g * 2 + 1) * T + t];
      gdst[(g * CHUNKS + 4) * T + t] = sD[g * T + t];
    }
  }
}

// ---- host AoS (kIdx order) <-> tile layout --------------------------------------------------------------
// slot s -> tile = s / TILE, g = (s % TILE) / T, t = s % T.  Padding slots (s >= n) replicate kangaroo s % n
// so that every tile is full of valid walkers; their DPs are dropped by the kidx < nKangaroos test.
__global__ void pack_kernel(uint4* state, const uint4* px, const uint4* py, const uint4* d, u64 n, u64 nPadded) {
  u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nPadded) return;
  u64 src = s < n ? s : s % n;
  u64 tile = s / TILE; int r = (int)(s % TILE); int g = r / T, t = r % T;
  uint4* dst = state + tile * (K * CHUNKS * T);
  dst[(g * CHUNKS + 0) * T + t] = px[2 * src];
  dst[(g * CHUNKS + 1) * T + t] = px[2 * src + 1];
  dst[(g * CHUNKS + 2) * T + t] = py[2 * src];
  dst[(g * CHUNKS + 3) * T + t] = py[2 * src + 1];
  dst[(g * CHUNKS + 4) * T + t] = d[src];
}
__global__ void unpack_kernel(const uint4* state, uint4* px, uint4* py, uint4* d, u64 n) {
  u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  u64 tile = s / TILE; int r = (int)(s % TILE); int g = r / T, t = r % T;
  const uint4* src = state + tile * (K * CHUNKS * T);
  px[2 * s] = src[(g * CHUNKS + 0) * T + t];
  px[2 * s + 1] = src[(g * CHUNKS + 1) * T + t];
  py[2 * s] = src[(g * CHUNKS + 2) * T + t];
  py[2 * s + 1] = src[(g * CHUNKS + 3) * T + t];
  d[s] = src[(g * CHUNKS + 4) * T + t];
}
struct PatchArgs { uint4 c[CHUNKS]; };
__global__ void patch_kernel(uint4* state, u64 s, PatchArgs a) {
  u64 tile = s / TILE; int r = (int)(s % TILE); int g = r / T, t = r % T;
  uint4* dst = state + tile * (K * CHUNKS * T);
  if (threadIdx.x < CHUNKS) dst[(g * CHUNKS + threadIdx.x) * T + t] = a.c[threadIdx.x];
}

// ---- unit-test / microbench kernels ------------------------------------------------------------------
__global__ void test_field_kernel(int op, int n, const u32* a, const u32* b, u32* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[8], y[8], r[8];
#pragma unroll
  for (int w = 0; w < 8; w++) { x[w] = a[i * 8 + w]; y[w] = b[i * 8 + w]; }
  if (op == 0) fe_mul(r, x, y);
  else if (op == 1) fe_sqr(r, x);
  else if (op == 2) fe_sub(r, x, y);
  else fe_inv(r, x);
#pragma unroll
  for (int w = 0; w < 8; w++) out[i * 8 + w] = r[w];
}

// kind 0: dependent-free IMAD.WIDE.U32 issue-rate probe; kind 1: fe_mul chains; kind 2: fe_sqr chains; kind 3: fe_inv
__global__ void bench_raw_kernel(int kind, int iters, u32* sink) {
  const u32 seed = blockIdx.x * blockDim.x + threadIdx.x + 1u;
  if (kind == 0) {
    u64 acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = seed * (k + 3u);
    const u32 m0 = seed | 1u, m1 = seed * 7u + 5u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(m0 + k), "r"(m1));
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567ull) sink[0] = (u32)s;
    return;
  }
  u32 a[8], b[8];
#pragma unroll
  for (int w = 0; w < 8; w++) { a[w] = seed * (2654435761u + w); b[w] = seed * (40503u + 7u * w) + w; }
  if (kind == 1) { for (int it = 0; it < iters; it++) { fe_mul(a, a, b); fe_mul(b, b, a); } }
  else if (kind == 2) { for (int it = 0; it < iters; it++) { fe_sqr(a, a); fe_sqr(b, b); } }
  else { for (int it = 0; it < iters; it++) { fe_inv(a, a); a[0] ^= b[0]; } }
  u32 s = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) s ^= a[w] ^ b[w];
  if (s == 0x12345u) sink[0] = s;
}

}  // namespace kgx
