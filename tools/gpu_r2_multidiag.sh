#!/bin/bash
# which switch makes --gpus 2 differ from --gpus 1 on the small collection?
mkdir -p gpurun_out
D=data/small; X=centrifuger_b200/centrifuger-b200
BASE="-x $D/idx -1 $D/pe_150_1.fq -2 $D/pe_150_2.fq -k 5 --batch 1500"
$X $BASE > /tmp/one.tsv 2>/dev/null
for V in "" "CFR_B200_PACK_INPUT=0" "CFR_B200_BULK_INGEST=0" "CFR_B200_PACK_INPUT=0 CFR_B200_BULK_INGEST=0" "CFR_B200_DENSE_LOCATE=-1" "CFR_B200_DENSE16=0" "CFR_B200_PAIRS=0" "CFR_B200_WIDE_LOOKUP=0"; do
  env $V $X $BASE --gpus 2 > /tmp/two.tsv 2>/tmp/two.err
  echo "== [$V] rc=$? differing rows: $(diff /tmp/one.tsv /tmp/two.tsv | grep -c '^[<>]')"
  diff /tmp/one.tsv /tmp/two.tsv | head -6
done
env $X $BASE --gpus 2 > /tmp/two_b.tsv 2>/dev/null; echo "2-GPU run vs 2-GPU run: $(diff /tmp/two.tsv /tmp/two_b.tsv | grep -c '^[<>]')"
# which batch / GPU do the differing reads belong to?
env $X $BASE --gpus 2 > /tmp/two.tsv 2>/dev/null
diff /tmp/one.tsv /tmp/two.tsv | grep '^<' | cut -f1 | sed 's/^< //' | sort -u | head -40 > /tmp/ids.txt
grep -n -F -f /tmp/ids.txt <(awk 'NR%4==1' $D/pe_150_1.fq | sed 's/^@//; s#/1$##') | head -40
