#!/bin/bash
# SASS evidence for profiles/: which Blackwell-specific instructions each
# hand-written kernel contains (run here, no GPU needed):
#   tools/sass_evidence.sh r02 > profiles/r02_sass.md
# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA),
# SYNCS.*TRANS64 = mbarrier transaction counts, UTCBAR = tcgen05.commit,
# LDGSTS = cp.async, REDG = red.global, STG.*.EF = evict-first streaming stores.
R=${1:-r02}
echo "# SASS evidence ($R): \`cuobjdump -sass build/csrc/*.o\`, instruction counts per kernel"
echo
echo "Built by \`make -C fb-bev_b200/csrc\` with \`-gencode arch=compute_100a,code=sm_100a -lineinfo\`."
echo "Mnemonic map (B200_PROFILING.md): \`tcgen05.mma\` -> UTCHMMA, \`tcgen05.ld\` / \`tcgen05.st\` -> LDTM / STTM, \`tcgen05.commit\` -> UTCBAR,"
echo "\`cp.async.bulk\` (TMA 1-D) -> UBLKCP, mbarrier tx -> SYNCS.*TRANS64, \`cp.async\` -> LDGSTS, \`red.global\` -> REDG."
echo
echo "Kernels with none of these instructions (plain LDG / STG SIMT code: the msda samplers with global gathers,"
echo "geometry, history_warp) are omitted."
echo
echo "| object | kernel | UTCHMMA | LDTM | STTM | UTCBAR | UBLKCP | SYNCS | LDGSTS | REDG/ATOMG | LDS | STG.EF |"
echo "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"
for f in linear_tf32 ffn_tf32 da_sca_smem lift_tail msda_fwd bev_pool_split bev_pool_fwd voxel_prepare geometry history_warp msda_bwd bev_pool_bwd; do
  cuobjdump -sass build/csrc/$f.o 2>/dev/null | awk -v obj=$f '
    /Function :/ { fn=$3 }
    { for (i=1;i<=NF;i++) {
        t=$i
        if (t ~ /^UTCHMMA/) c[fn,"a"]++
        else if (t ~ /^LDTM/) c[fn,"b"]++
        else if (t ~ /^STTM/) c[fn,"j"]++
        else if (t ~ /^UTCBAR/) c[fn,"c"]++
        else if (t ~ /^UBLKCP/) c[fn,"d"]++
        else if (t ~ /^SYNCS/) c[fn,"e"]++
        else if (t ~ /^LDGSTS/) c[fn,"f"]++
        else if (t ~ /^(REDG|ATOMG|RED\.)/) c[fn,"g"]++
        else if (t ~ /^LDS/) c[fn,"h"]++
        else if (t ~ /^STG.*\.EF/) c[fn,"i"]++
        seen[fn]=1 } }
    END { for (fn in seen) { if (fn=="") continue
            cmd="echo " fn " | c++filt | cut -c1-70"; cmd | getline nm; close(cmd)
            tot=c[fn,"a"]+c[fn,"b"]+c[fn,"c"]+c[fn,"d"]+c[fn,"e"]+c[fn,"f"]+c[fn,"g"]+c[fn,"h"]+c[fn,"i"]+c[fn,"j"]
            if (tot==0) continue
            printf "| %s.o | `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n", obj, nm,
              c[fn,"a"], c[fn,"b"], c[fn,"j"], c[fn,"c"], c[fn,"d"], c[fn,"e"], c[fn,"f"], c[fn,"g"], c[fn,"h"], c[fn,"i"] } }' | sort
done
