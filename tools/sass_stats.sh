#!/bin/bash
# usage: tools/sass_stats.sh <mangled-name-substring e.g. attention_kernelILi64> [lib.so]  -> instruction mix of one kernel (SASS left in /tmp/kernel.sass)
cuobjdump -sass ${2:-/root/repo/ullsam_b200/lib/libullsam_b200.so} > /tmp/all.sass
start=$(grep -n "Function :" /tmp/all.sass | grep "$1" | head -1 | cut -d: -f1)
end=$(grep -n "Function :" /tmp/all.sass | awk -F: -v s=$start '$1>s{print $1; exit}')
[ -z "$end" ] && end=$(wc -l < /tmp/all.sass)
sed -n "${start},${end}p" /tmp/all.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/kernel.sass
echo "instructions: $(wc -l < /tmp/kernel.sass)"
for op in FFMA2 FADD2 FMUL2 FMNMX3 "FMNMX " MUFU "MOV" STL LDL F2FP "FFMA " "FADD " "FMUL " "LDS" "STS" LDTM STTM UTCHMMA "LEA" "SHF" "IMAD" "IADD" "FSEL" "BAR"; do printf "%-8s %s\n" "$op" "$(grep -c "$op" /tmp/kernel.sass)"; done
