#!/bin/sh
# SASS evidence for profiles/: tools/sass_excerpts.sh > profiles/sass_rNN.md   (cuobjdump of the built objects, no GPU needed)
B=$(cd "$(dirname "$0")/.." && pwd)/gr-ieee80211_b200/csrc/build
demod() { cuobjdump -sass "$B/k_demod.o" | awk '/Function : .*k_demodEPK/{f=1} /Function : .*k_demod2/{f=0} f'; }
echo '# SASS excerpts (`cuobjdump -sass` of the built objects, sm_100a)'
echo
echo '## k_demod: the scattered LLR line leaves shared memory by the bulk-copy engine'
echo '`cp.async.bulk.global.shared::cta.bulk_group` -> `UBLKCP.G.S`; the writers fence towards the async proxy (`FENCE.VIEW.ASYNC.S`),'
echo 'one lane issues the copy (one per row when the rows are padded), `UTMACMDFLUSH` + `DEPBAR` = commit_group / wait_group.read 0'
echo 'before the CTA gives its shared memory back.'
echo '```'
demod | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | awk '/FENCE.VIEW.ASYNC/{p=1} p{print} /DEPBAR.LE SB0/{if(p){exit}}' | grep -v "NOP"
echo '```'
echo
echo '## k_demod: global loads'
echo '```'
demod | grep -E "LDG" | awk '{for(i=2;i<=NF;i++) if ($i ~ /^LDG/) {print $i; break}}' | sort | uniq -c | sort -rn
echo '```'
echo '(LDG.E.64: the 8 samples, requested back to back, 8 x 1/H, the item offset, the first twiddle pair; LDG.E.128: the other twiddle rows'
echo 'and the 8 demapTab entries of the thread; LDG.E: frame-record fields, pilot polarity)'
echo
echo '## k_viterbi_tp: soft-bit staging by LDGSTS (cp.async), groups retired with LDGDEPBAR / DEPBAR'
echo '```'
cuobjdump -sass "$B/k_viterbi_tp.o" | grep -E "LDGSTS|LDGDEPBAR|DEPBAR" | awk '{for(i=2;i<=NF;i++) if ($i ~ /LDGSTS|LDGDEPBAR|DEPBAR/) {print $i; break}}' | sort | uniq -c | sort -rn | head -8
echo '```'
