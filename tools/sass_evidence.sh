#!/bin/bash
# Regenerates profiles/sass_*.txt: per-kernel opcode histograms and the Blackwell-specific instructions of the shipped
# libmst_b200.so (cuobjdump -sass).  usage: tools/sass_evidence.sh
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
lib=$root/ml_music_style_transfer_b200/libmst_b200.so
out=$root/profiles
cuobjdump -sass "$lib" > /tmp/mst_all.sass
kernel() {  # $1 = mangled-name regex, $2 = output file, $3 = grep pattern for the excerpt
  awk -v pat="$1" '/Function : /{f = ($0 ~ pat)} f' /tmp/mst_all.sass > /tmp/mst_k.sass
  {
    echo "# cuobjdump -sass ml_music_style_transfer_b200/libmst_b200.so, function(s) matching /$1/"
    grep "Function : " /tmp/mst_k.sass | sed 's/^\s*/# /'
    echo "# opcode histogram (static):"
    grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] +)?[A-Z0-9_.]+" /tmp/mst_k.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40 | sed 's/^/#   /'
    echo "# excerpt: lines matching /$3/ (first 60)"
    grep -E "$3" /tmp/mst_k.sass | sed 's#/\* 0x[0-9a-f]* \*/##' | sed 's/^\s*//' | head -60
  } > "$out/$2"
  echo "wrote $out/$2 ($(wc -l < "$out/$2") lines)"
}
kernel "mel_gemm_kernel" sass_mel_gemm.txt "UTC[A-Z]*MMA|UTMALDG|UBLKCP|LDTM|STTM|UTCBAR|SYNCS|UTCATOM|TCGEN|UGETNEXT|R2UR"
kernel "gl_kernelILb0ELb0" sass_gl_kernel.txt "FFMA2|FADD2|FMUL2|LDGSTS|REDG|RED\.|SHFL|MUFU"
kernel "mst11stft_kernelILi3" sass_stft_log1p.txt "FFMA2|FADD2|FMUL2|SHFL|MUFU|LDG\.E\.64"
kernel "upsample_kernelIaLi2" sass_upsample.txt "STG\.E\.128|PRMT|SHFL|VOTE|LOP3"
