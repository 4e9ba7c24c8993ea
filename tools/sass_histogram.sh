#!/bin/bash
# SASS opcode histogram of the shipped library: the Blackwell-native claim in the tree (UTCHMMA = tcgen05.mma, LDTM / STTM =
# tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier).
# usage: tools/sass_histogram.sh > profiles/rN_sass_histogram.txt
set -e
LIB=${1:-alphazero-piskvorky_b200/libazp_b200.so}
echo "# cuobjdump -sass $LIB ($(date -u +%F), $(nvcc --version | tail -1))"
cuobjdump -sass "$LIB" > /tmp/azp_sass.txt
echo "# functions: $(grep -c 'Function :' /tmp/azp_sass.txt)"
echo "# tensor-core / TMA / tensor-memory opcodes, whole library:"
grep -oE '\b(UTCHMMA(\.2CTA)?|UTCQMMA|LDTM(\.[A-Za-z0-9x]+)*|STTM(\.[A-Za-z0-9x]+)*|UTMALDG(\.[0-9A-Z]+)*|UTMASTG(\.[0-9A-Z]+)*|UTCBAR(\.[A-Z0-9]+)*|UTCATOMSWS(\.[A-Z_0-9]+)*|SYNCS(\.[A-Z0-9_]+)*|ELECT|HMMA(\.[0-9A-Z]+)*|R2UR(\.[A-Z]+)*)\b' /tmp/azp_sass.txt | sort | uniq -c | sort -rn
echo "# per kernel: UTCHMMA count and R2UR.BROADCAST within 6 instructions before a UTCHMMA (the issue waterfall, 0 = uniform issue)"
awk '/Function :/{name=$3} /UTCHMMA/{c[name]++} END{for(n in c) print c[n], n}' /tmp/azp_sass.txt | sort -rn | while read cnt name; do
  w=$(awk -v n="$name" '/Function :/{f=($3==n)} f{print}' /tmp/azp_sass.txt | grep -B6 UTCHMMA | grep -c 'R2UR.BROADCAST' || true)
  echo "$cnt UTCHMMA, $w R2UR.BROADCAST near them: $(echo $name | sed -E "s/^__nv_static_[0-9]+__[0-9a-f]+_[0-9]+_[A-Za-z0-9_]+_cu_[0-9a-f]+__//" | c++filt | sed -E "s/\(anonymous namespace\):://g" | cut -c1-120)"
done
