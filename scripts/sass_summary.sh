#!/bin/bash
# Opcode evidence per shipped library: counts of the sm_100a-specific SASS mnemonics (tcgen05 = UTC*, TMA = UTMA* / UBLKCP,
# DSMEM st.async = STAS, mbarrier = SYNCS, cluster barrier = UCGABAR, setmaxnreg = USETMAXREG, REDUX, strong L2 loads).
# Usage: scripts/sass_summary.sh > profiles/rNN_sass_opcodes.txt
cd "$(dirname "$0")/.."
for so in tacotron_wavenet_vocoder_korean_b200/*.so; do
    echo "== $so  ($(cuobjdump -lelf "$so" | grep -c sm_100a) sm_100a cubin(s))"
    cuobjdump -sass "$so" 2>/dev/null | awk '
        /Function : /{fn=$3}
        {
          for (i = 1; i <= NF; ++i) {
            t = $i
            if (t ~ /^(UTCHMMA|UTCMMA|UTCQMMA|UTCOMMA|UTCBAR|UTCCP|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTMACCTL|UBLKCP|UBLKRED|STAS|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|USETMAXREG|REDUX|HMMA|IMMA|DMMA|LDGSTS)/) {
              split(t, a, "."); c[a[1]]++; k[fn" "a[1]]++
            }
            if (t ~ /^LDG\.E\.64\.STRONG\.GPU/) { c["LDG.STRONG.GPU"]++ }
            if (t ~ /^ST\.E\.64\.STRONG\.GPU|^STG\.E\.64\.STRONG\.GPU/) { c["STG.STRONG.GPU"]++ }
          }
        }
        END { for (m in c) printf "  %-18s %6d\n", m, c[m]; print "  -- per kernel (tensor core / TMA / DSMEM only)";
              for (x in k) if (x ~ /UTC|LDTM|STTM|UTMA|UBLKCP|STAS|USETMAXREG/) printf "     %6d  %s\n", k[x], x }' | sort -k1,1 -s | cut -c1-200
done
