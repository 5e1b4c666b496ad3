#!/bin/bash
# Histogram of the Blackwell-specific SASS opcodes per object of libb2r.so (no GPU needed):
#   UTCHMMA/UTCQMMA  tcgen05.mma          LDTM / STTM      tcgen05.ld / st (TMEM)
#   UTCBAR           tcgen05.commit       UBLKCP           cp.async.bulk (TMA, 1-D)
#   UTMALDG          tensor-map TMA       LDGSTS           cp.async
#   UCGABAR_*        barrier.cluster      SYNCS            mbarrier ops
#   REDUX            redux.sync           ATOMG/RED        global atomics
cd "$(dirname "$0")/.." || exit 1
out=profiles/r02/sass_opcodes.txt
{
  echo "# cuobjdump -sass of backtoreality_b200/csrc/build/*.o (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), sm_100a)"
  printf "%-16s %8s %8s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" object instrs UTCHMMA LDTM UTCBAR UBLKCP UTMALDG LDGSTS UCGABAR SYNCS REDUX RED/ATOMG
  for o in backtoreality_b200/csrc/build/*.o; do
    s=$(cuobjdump -sass "$o" 2>/dev/null)
    c() { echo "$s" | grep -c -E "$1"; }
    printf "%-16s %8d %8d %8d %8d %8d %8d %8d %8d %8d %8d %8d\n" "$(basename "$o")" \
      "$(echo "$s" | grep -c -E '^\s+/\*[0-9a-f]{4}\*/')" "$(c 'UTC[HQ]MMA')" "$(c 'LDTM')" "$(c 'UTCBAR')" \
      "$(c 'UBLKCP')" "$(c 'UTMALDG')" "$(c 'LDGSTS')" "$(c 'UCGABAR')" "$(c 'SYNCS')" "$(c 'REDUX')" "$(c 'REDG|ATOMG|RED\.')"
  done
  echo
  echo "# kernels per object (cuobjdump -elf symbols of type FUNC in .text.*)"
  for o in backtoreality_b200/csrc/build/*.o; do
    echo "$(basename "$o"): $(cuobjdump -sass "$o" 2>/dev/null | grep -c 'Function :') kernels"
  done
} > "$out"
cat "$out"
