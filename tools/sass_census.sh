#!/bin/bash
# SASS census of the shipped library: proves the sm_100a-only instructions (tcgen05 = UTC*MMA / LDTM / STTM, TMA = UTMALDG / UTMASTG,
# bulk DSMEM copies = UBLKCP, PDL = no dedicated mnemonic (ACQBULK/PREEXIT)).  Usage: tools/sass_census.sh > profiles/r02_sass_census.txt
LIB=${1:-vitok-release_b200/vitok_b200/libvitok_b200.so}
echo "# SASS census of $LIB ($(stat -c %y "$LIB" | cut -d. -f1), $(stat -c %s "$LIB") bytes); cuobjdump -sass | grep -c <mnemonic>"
cuobjdump -sass "$LIB" > /tmp/vtk_sass.txt
for op in 'UTCHMMA ' 'UTCHMMA\.2CTA' 'UTCQMMA\.2CTA' 'UTCQMMA ' 'UTCBAR' 'LDTM' 'STTM' 'UTMALDG' 'UTMALDG\.2D\.2CTA' 'MULTICAST' 'UTMASTG' 'UBLKCP' 'UTMAPF' 'SYNCS' 'MUFU\.EX2' 'F2FP' 'FFMA2\|FMUL2\|FADD2' '[^C]HMMA' 'LDGSTS' 'ACQBULK' 'PREEXIT\|ACQUIRE.*GDC\|griddep'; do
  printf "%-28s %6d\n" "$op" "$(grep -c "$op" /tmp/vtk_sass.txt)"
done
echo "# kernels (Function :) in the library: $(grep -c 'Function :' /tmp/vtk_sass.txt)"
echo "# target arch lines: $(cuobjdump -lelf "$LIB" | sort | uniq -c | tr '\n' ';')"
