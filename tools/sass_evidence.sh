#!/bin/bash
# SASS mnemonic histogram of the kernels whose mangled name matches $1 (default: the bulk-copy ring kernels):
# evidence of the Blackwell copy-engine path (UBLKCP = cp.async.bulk, SYNCS.* = mbarrier arrive / try_wait).
#   tools/sass_evidence.sh 'cic32_ringIdLb0ELi3ELi3E' > profiles/r2_sass_ring.txt
PAT=${1:-cic32_ring}
SO=$(dirname "$0")/../pmesh_b200/csrc/libpmesh_b200.so
cuobjdump -sass "$SO" 2>/dev/null | awk -v pat="$PAT" '
  /Function : /{ p = ($0 ~ pat); if (p) print $0 }
  p && /^[ \t]+\/\*[0-9a-f]+\*\// {
      line = $0; sub(/\/\* 0x[0-9a-f]+ \*\//, "", line);
      n = split(line, f, " "); op = f[2]; if (op ~ /^@/) op = f[3];
      sub(/;$/, "", op); cnt[op]++ }
  END { for (o in cnt) printf "%6d %s\n", cnt[o], o | "sort -rn" }'
