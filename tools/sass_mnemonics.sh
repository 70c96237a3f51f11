#!/bin/bash
# Counts of the Blackwell-specific SASS mnemonics per kernel of the shipped library (tcgen05 = UTCHMMA / UTCBAR / LDTM / STTM /
# UTCATOMSWS (TMEM alloc), TMA = UTMALDG / UTMASTG (tensor maps) and UBLKCP (bulk copies), mbarrier = SYNCS, elect = ELECT).
cuobjdump -sass ${1:-retinaface_b200/librf_b200.so} 2>/dev/null | awk '/Function : /{fn=$3} /UTMALDG|UTMASTG|UTCHMMA|UTCIMMA|UTCBAR|LDTM|STTM|UTCATOMSWS|UBLKCP|ELECT/{n=split($0,a," "); for(i=1;i<=n;i++) if (a[i] ~ /^(UTMALDG|UTMASTG|UTCHMMA|UTCIMMA|UTCBAR|LDTM|STTM|UTCATOMSWS|UBLKCP|ELECT)/) {sub(/\..*/,"",a[i]); c[fn" "a[i]]++}} END{for(k in c) print k, c[k]}' | sort | c++filt | awk '{cnt=$NF; m=$(NF-1); $NF=""; $(NF-1)=""; printf "%-10s %5d  %s\n", m, cnt, $0}'
