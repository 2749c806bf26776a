#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA in the in-tree libfvy.so (sm_100a only).
# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld (TMEM -> registers),
# UTCBAR = tcgen05.commit, HMMA = warp-level mma.sync (stem), SYNCS = mbarrier ops.   usage: tools/sass_counts.sh > profiles/r02_sass_counts.txt
cd "$(dirname "$0")/.."
SO=face_vijnana_yolov3_b200/libfvy.so
echo "# $(sha256sum $SO | cut -c1-16)  $SO  ($(cuobjdump -lelf $SO | tr '\n' ' '))"
cuobjdump -sass $SO | awk '
  /Function :/ { fn=$3; next }
  { for (i=1;i<=NF;i++) { t=$i; sub(/\..*/,"",t);
      if (t=="UTCHMMA"||t=="UTMALDG"||t=="UTMASTG"||t=="LDTM"||t=="UTCBAR"||t=="HMMA"||t=="SYNCS"||t=="UTMAPF"||t=="UTCCP") c[fn" "t]++ } }
  END { for (k in c) print k, c[k] }' | sort | c++filt | awk '
  { fn=""; for (i=1;i<=NF-2;i++) fn=fn $i " "; key=$(NF-1); n=$NF; gsub(/\(.*/,"",fn); tot[fn]=tot[fn] " " key "=" n }
  END { for (f in tot) print f ":" tot[f] }' | sort
