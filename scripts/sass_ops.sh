#!/bin/bash
# SASS opcode histogram of the hot kernels of libmobula_roialign_sm100.so (runs without a GPU).
#   bash scripts/sass_ops.sh > profiles/r2_sass_ops.txt
LIB=mobulaop_b200/lib/libmobula_roialign_sm100.so
for k in 'resident_fwd2r_kernelILi7ELi1ELb0ELb0ELb1' 'resident_fwd2r_kernelILi7ELi1ELb0ELb0ELb0' 'pull8_bwd_kernelILi4ELb0ELi4ELi3ELb0' 'resident_fwd2p_kernelILi7ELi2ELb0ELb0' 'pull_bwd_kernelILi4ELb0ELb0ELi8ELb0' \
         'pull_tables_kernelILb0ELb0ELi128' 'sweep_fwd_kernelILi7ELi0ELb0' 'sweep_fwd_kernelILi7ELi2ELb0' 'col2im_kernel' 'im2col_kernel'; do
  echo "== $k"
  cuobjdump -sass $LIB 2>/dev/null | awk -v pat="$k" '
    /Function :/ { on = index($0, pat) > 0 }
    on && /^ +\/\*[0-9a-f]+\*\/ / { op = $2; if (op ~ /^@/) op = $3; sub(/;$/, "", op); split(op, a, "."); n[a[1]]++; full[op]++; t++ }
    END { printf "  instructions %d\n", t;
          printf "  of interest: UBLKCP %d  SYNCS %d  LDGSTS %d  UTMALDG %d  FFMA2 %d  FMUL2 %d  LDS %d  LDG %d  STG %d  SHFL %d  ATOM/RED %d\n",
                 n["UBLKCP"], n["SYNCS"], n["LDGSTS"], n["UTMALDG"], n["FFMA2"], n["FMUL2"], n["LDS"], n["LDG"], n["STG"], n["SHFL"], n["ATOM"] + n["ATOMG"] + n["RED"] + n["ATOMS"];
          for (o in full) printf "  %6d %s\n", full[o], o | "sort -k1,1nr | head -28" }'
done
