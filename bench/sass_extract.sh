#!/bin/sh
# Evidence for profiles/: which Blackwell-era instructions the built library holds, per kernel of the hot path
# (TMA bulk copies UBLKCP, mbarrier SYNCS.*TRANS64, 256-bit loads LDG.E.*.256, shared-memory atomics, 128-bit LDS).
# Usage: sh bench/sass_extract.sh > profiles/r2_sass_extract.txt      (needs cuobjdump; no GPU)
cd "$(dirname "$0")/.." || exit 1
cuobjdump -sass panagram_b200/libpkanchor.so 2>/dev/null |
  awk '/Function :/{fn=$3; next}
       {for(i=1;i<=NF;i++) if ($i ~ /^(UBLKCP|UBLKPF|SYNCS|LDG\.E.*(128|256)|ATOMS|ATOMG|LDS\.128|STG\.E\.128)/) {c[fn" "$i]++; break}}
       END{for(k in c) print c[k], k}' |
  grep -E "probe_g32l2_kernelILi256ELi4ELi8|probe_g32c_kernelILi384ELi4ELi4ELi1|reducew_kernelILi8|probe_win_kernelILi384ELi4ELi3ELi4ELi0|probe_win_kernelILi256ELi3ELi2ELi4ELi1|gather_slice_dst_kernelILi8ELi1|gather_slice_kernelILi8ELi1|unpermute_slice|partition_seq_roll_kernelILi4|partition_fine|bgzf_encode|reduce1" |
  sort -k2,2 -k1,1nr
