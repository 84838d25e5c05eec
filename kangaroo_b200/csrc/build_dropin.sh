#!/bin/bash
# build_dropin.sh -- link the reference's UNMODIFIED host program against the B200 engine.
#   build/kangaroo_b200 = reference main.cpp Kangaroo.cpp Check.cpp Thread.cpp Backup.cpp Network.cpp Merge.cpp
#                         PartMerge.cpp HashTable.cpp Timer.cpp SECPK1/*.cpp   (compiled where they lie, -DWITHGPU)
#                       + kangaroo_b200/csrc/GPUEngine_b200.cpp (replaces GPU/GPUEngine.cu) + libkgx.so
# Needs the reference sources at $REF (default /root/reference); nothing is copied into the repo.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=$ROOT/build
mkdir -p "$OUT"
[ -f "$HERE/libkgx.so" ] || { echo "build libkgx.so first (__graft_entry__.build())"; exit 1; }
FLAGS="-O2 -m64 -mssse3 -Wno-unused-result -Wno-write-strings -include cstdint -DWITHGPU -I$REF"
SRC="main.cpp Kangaroo.cpp Check.cpp Thread.cpp Backup.cpp Network.cpp Merge.cpp PartMerge.cpp HashTable.cpp Timer.cpp \
     SECPK1/Int.cpp SECPK1/IntMod.cpp SECPK1/IntGroup.cpp SECPK1/Point.cpp SECPK1/SECP256K1.cpp SECPK1/Random.cpp"
OBJS=""
for f in $SRC; do
  o="$OUT/obj_$(echo $f | tr '/' '_' | sed 's/\.cpp$/.o/')"
  g++ $FLAGS -c "$REF/$f" -o "$o" &
  OBJS="$OBJS $o"
done
g++ $FLAGS -c "$HERE/GPUEngine_b200.cpp" -o "$OUT/obj_GPUEngine_b200.o" &
wait
g++ -o "$OUT/kangaroo_b200" $OBJS "$OUT/obj_GPUEngine_b200.o" -L"$HERE" -lkgx -lpthread -Wl,-rpath,'$ORIGIN/../kangaroo_b200/csrc'
# end-to-end timing harness through the same class GPUEngine (bench.py's e2e leg): shim + reference SECPK1 only
SECP_OBJS=$(for f in Int IntMod IntGroup Point SECP256K1 Random; do echo "$OUT/obj_SECPK1_$f.o"; done)
g++ $FLAGS -o "$OUT/kgx_shim_bench" "$HERE/shim_bench.cpp" "$OUT/obj_GPUEngine_b200.o" $SECP_OBJS "$OUT/obj_Timer.o" \
    -L"$HERE" -lkgx -lpthread -Wl,-rpath,'$ORIGIN/../kangaroo_b200/csrc'
rm -f $OUT/obj_*.o
echo "built $OUT/kangaroo_b200"
