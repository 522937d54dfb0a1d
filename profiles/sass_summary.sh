#!/bin/bash
# Counts the SASS mnemonics that prove tensor-core / TMEM / TMA use (B200_PROFILING.md) per kernel of the built library.
# Runs in the build container (no GPU needed).   usage: bash profiles/sass_summary.sh > profiles/r02/sass_mnemonics.txt
SO=multimodal-framework-for-speaker-emotion-recognition_b200/liblsthm_b200.so
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; sub(/^_ZN5lsthm[0-9]*/, "", fn) }
  { for (i = 1; i <= NF; i++) if ($i ~ /^(UTCHMMA|UTCBAR|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|SYNCS|USETMAXREG|FFMA2|HMMA|LDG|STG)(\.|;|$)/) { m=$i; sub(/[.;].*/, "", m); c[fn" "m]++ } }
  END { for (k in c) print k, c[k] }' | sort | grep -v " LDG \| STG \| FFMA2 [0-9]$" 
