// shim_bench.cpp -- end-to-end timing THROUGH the reference's own C++ interface: `class GPUEngine` (reference header
// GPU/GPUEngine.h, unchanged) as implemented by GPUEngine_b200.cpp over libkgx.so.  This is the loop of
// Kangaroo::SolveKeyGPU (Kangaroo.cpp:563-575): callKernel() once, then K x Launch(std::vector<ITEM>&), each of which waits
// for the kernel, relaunches, copies the DP records device -> host and marshals every record into reference `Int`s with the
// wild-offset correction (GPUEngine.cu:653-675) -- i.e. what a maintainer who links the shim actually gets per step.
//
//   build/kgx_shim_bench <gpuId> <gridX> <gridY> <dpBits> <warmup> <steps> <table.bin> <herd.bin>
//     table.bin : 32 x (jd[2], jpx[4], jpy[4]) u64 limbs, then wildOffset[4]          (written by bench.py)
//     herd.bin  : n x (px[4], py[4], d[4]) u64 limbs, kIdx order, d = true distance mod n (what CreateHerd produces)
//   prints one JSON line: {"steps":K,"seconds":S,"items":I,"kangaroos":n,"upload_s":U}
// Built by build_dropin.sh next to the drop-in binary (needs the reference sources at build time; travels prebuilt).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "GPU/GPUEngine.h"
#include "SECPK1/SECP256k1.h"

static bool read_all(const char* path, std::vector<uint64_t>& buf, size_t words) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  buf.resize(words);
  const bool ok = fread(buf.data(), 8, words, f) == words;
  fclose(f);
  return ok;
}

int main(int argc, char** argv) {
  if (argc != 9) { fprintf(stderr, "usage: %s gpuId gridX gridY dpBits warmup steps table.bin herd.bin\n", argv[0]); return 2; }
  const int gpuId = atoi(argv[1]), gx = atoi(argv[2]), gy = atoi(argv[3]), dp = atoi(argv[4]), warm = atoi(argv[5]), steps = atoi(argv[6]);
  Secp256K1 secp; secp.Init();                       // installs the group order used by Int::ModAddK1order / ModSubK1order
  std::vector<uint64_t> tab, herd;
  if (!read_all(argv[7], tab, 32 * 10 + 4)) { fprintf(stderr, "cannot read %s\n", argv[7]); return 2; }
  GPUEngine eng(gx, gy, gpuId, 65536 * 2);           // Kangaroo.cpp:523
  const size_t n = (size_t)eng.GetNbThread() * GPU_GRP_SIZE;
  if (!read_all(argv[8], herd, n * 12)) { fprintf(stderr, "cannot read %s (%zu kangaroos)\n", argv[8], n); return 2; }
  Int jd[NB_JUMP], jx[NB_JUMP], jy[NB_JUMP], wo;
  for (int j = 0; j < NB_JUMP; j++) {
    jd[j].SetInt32(0); jx[j].SetInt32(0); jy[j].SetInt32(0);
    for (int k = 0; k < 2; k++) jd[j].bits64[k] = tab[j * 10 + k];
    for (int k = 0; k < 4; k++) { jx[j].bits64[k] = tab[j * 10 + 2 + k]; jy[j].bits64[k] = tab[j * 10 + 6 + k]; }
  }
  wo.SetInt32(0);
  for (int k = 0; k < 4; k++) wo.bits64[k] = tab[320 + k];
  Int* px = new Int[n]; Int* py = new Int[n]; Int* d = new Int[n];
  for (size_t i = 0; i < n; i++) {
    px[i].SetInt32(0); py[i].SetInt32(0); d[i].SetInt32(0);
    for (int k = 0; k < 4; k++) { px[i].bits64[k] = herd[i * 12 + k]; py[i].bits64[k] = herd[i * 12 + 4 + k]; d[i].bits64[k] = herd[i * 12 + 8 + k]; }
  }
  const uint64_t dMask = dp == 0 ? 0 : (~0ULL << (64 - dp));      // Kangaroo::SetDP, Kangaroo.cpp:154-164
  auto t0 = std::chrono::steady_clock::now();
  eng.SetWildOffset(&wo);
  eng.SetParams(dMask, jd, jx, jy);
  eng.SetKangaroos(px, py, d);
  const double upload = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::vector<ITEM> found;
  if (!eng.callKernel()) { fprintf(stderr, "callKernel failed\n"); return 1; }
  for (int i = 0; i < warm; i++) eng.Launch(found);
  uint64_t items = 0, chk = 0;
  t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < steps; i++) {
    if (!eng.Launch(found)) { fprintf(stderr, "Launch failed\n"); return 1; }
    items += found.size();
    for (size_t k = 0; k < found.size(); k++) chk ^= found[k].x.bits64[0] ^ found[k].d.bits64[0] ^ found[k].kIdx;   // touch every record
  }
  // each Launch waited for one kernel (the one started by the previous call) and read its DPs back: exactly K kernels and K
  // readbacks are inside the timed region; the launch started by the last call is outside (the destructor waits for it)
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // checkpoint read-out (Kangaroo.cpp:618-626): GetKangaroos waits for the launch in flight, copies the herd out and marshals it into Int
  t0 = std::chrono::steady_clock::now();
  eng.GetKangaroos(px, py, d);
  const double getk = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"steps\": %d, \"seconds\": %.6f, \"items\": %llu, \"kangaroos\": %zu, \"upload_s\": %.3f, \"get_kangaroos_s\": %.3f, \"chk\": %llu}\n",
         steps, sec, (unsigned long long)items, n, upload, getk, (unsigned long long)chk);
  delete[] px; delete[] py; delete[] d;
  return 0;
}
