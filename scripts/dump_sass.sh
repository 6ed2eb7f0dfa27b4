#!/bin/bash
# SASS listings of the hot kernels (cuobjdump, no GPU needed) -> profiles/r02_sass/<kernel>.sass.gz, plus the loop table
# (instruction count and opcode histogram of every loop) and the markers that say which data movers are in use.
set -e
cd "$(dirname "$0")/.."
B=alac_b200/csrc/build
OUT=profiles/r02_sass
mkdir -p $OUT
dump() {   # <object> <mangled name> <file stem>
  cuobjdump -sass -fun "$2" $B/$1 > /tmp/$3.sass
  gzip -9 -c /tmp/$3.sass > $OUT/$3.sass.gz
  { echo "== $3  ($2)"; cuobjdump -res-usage -fun "$2" $B/$1 2>/dev/null | grep -E "REG" | sed 's/^ */   /';
    python profiles/sass_loops.py /tmp/$3.sass 60;
    echo "   data movers: LDGSTS $(grep -c LDGSTS /tmp/$3.sass)  LDG $(grep -c 'LDG\.' /tmp/$3.sass)  LDS $(grep -c 'LDS' /tmp/$3.sass)  STS $(grep -c 'STS' /tmp/$3.sass)  STG $(grep -c 'STG' /tmp/$3.sass)  BAR $(grep -c 'BAR\.' /tmp/$3.sass)  UBLKCP $(grep -c UBLKCP /tmp/$3.sass)  UTMALDG $(grep -c UTMALDG /tmp/$3.sass)  SYNCS $(grep -c SYNCS /tmp/$3.sass)"; echo; } >> $OUT/loops.txt
}
rm -f $OUT/loops.txt
dump kernels_d16.o _ZN5alacb23enc_search_split_kernelILi16ELb1ELb1ELb0ELb0EEEvNS_7EncArgsEjjNS_8JobListsE enc_search_split_16_stereo
dump kernels_d16.o _ZN5alacb17enc_final2_kernelILi16ELb1ELb0ELb1EEEvNS_7EncArgsENS_8JobListsEj enc_final2_16_stereo_twowarp
dump kernels_d16.o _ZN5alacb19enc_assemble_kernelILi16EEEvNS_7AsmArgsE enc_assemble_16
dump kernels_d16.o _ZN5alacb16dec_fused_kernelILi16EEEvNS_7DecArgsE dec_fused_16
dump kernels_d24.o _ZN5alacb23enc_search_split_kernelILi24ELb1ELb1ELb0ELb1EEEvNS_7EncArgsEjjNS_8JobListsE enc_search_split_24_stereo_dense
dump kernels_d24.o _ZN5alacb17enc_final2_kernelILi24ELb1ELb0ELb0EEEvNS_7EncArgsENS_8JobListsEj enc_final2_24_stereo_onewarp
dump kernels_d24.o _ZN5alacb19enc_assemble_kernelILi24EEEvNS_7AsmArgsE enc_assemble_24
dump kernels_d24.o _ZN5alacb18dec_entropy_kernelILi24EEEvNS_7DecArgsE dec_entropy_24
dump kernels_d24.o _ZN5alacb17dec_finish_kernelILi24EEEvNS_7DecArgsE dec_finish_24
ls -la $OUT
