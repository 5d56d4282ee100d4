#!/bin/bash
# SASS evidence for profiles/: per kernel, how many tcgen05 / TMEM / TMA-engine instructions the shipped library contains.
# usage: scripts/sass_counts.sh > profiles/r02_sass_counts.txt
cd "$(dirname "$0")/.."
LIB=baybe_b200/_C/libbaybe_b200.so
echo "# cuobjdump -sass $LIB  ($(date -u +%F), nvcc $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (1-D, TMA engine),"
echo "# UTMASTG = cp.async.bulk.tensor store through a tensor map, SYNCS = mbarrier ops, FENCE.VIEW.ASYNC = fence.proxy.async"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
  /Function :/ { fn=$3 }
  /UTCHMMA/ {a[fn]++} /UTCBAR/ {b[fn]++} /LDTM/ {c[fn]++} /STTM/ {d[fn]++} /UBLKCP/ {e[fn]++} /UTMASTG/ {f[fn]++} /UTMALDG/ {g[fn]++} /SYNCS/ {h[fn]++} /ELECT/ {k[fn]++}
  /Function :/ { seen[fn]=1 }
  END { printf "%-8s %-7s %-5s %-5s %-7s %-8s %-8s %-6s %-6s %s\n","UTCHMMA","UTCBAR","LDTM","STTM","UBLKCP","UTMASTG","UTMALDG","SYNCS","ELECT","kernel";
        for (fn in seen) if (a[fn]+c[fn]+e[fn]+f[fn] > 0) printf "%-8d %-7d %-5d %-5d %-7d %-8d %-8d %-6d %-6d %s\n", a[fn],b[fn],c[fn],d[fn],e[fn],f[fn],g[fn],h[fn],k[fn],fn }' | (read -r hdr; echo "$hdr"; sort -k10)
echo
echo "# excerpt: the constant-folded MMA issue loop of k_fused_ts<matern52> (first V chunk: TS-form UTCHMMA with the A operand in tensor memory)"
cuobjdump -sass -fun '_ZN2bb10k_fused_tsILi2ELb0ELb0EEEvNS_11FusedParamsE' "$LIB" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | awk '/UTCHMMA/ {n++} n>=25 && n<=40' | head -60
