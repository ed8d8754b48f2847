#!/bin/bash
# tools/diff_against_ref.sh — pins the oracle the day the true reference mapper can be built: runs
# oracle/_ref/bidir_index + bidir_mapping (built by oracle/ref_hook/build_ref.sh from the reference sources + SeqAn), the
# oracle restatement (oracle/oracle_bidir_mapping) and — on a GPU box — the CUDA drop-in on the same seeded FASTA inputs for
# k = 0..8 with and without -P AG, and diffs the SAM files: first the parity key (QNAME, strand, RNAME, POS, NM) as a set,
# then the files byte for byte (order, FLAG bit 256, MD style).  Exit code 0 = identical everywhere.
set -u
cd "$(dirname "$0")/.."
REFBIN=oracle/_ref
[ -x $REFBIN/bidir_mapping ] || { echo "no $REFBIN/bidir_mapping: run oracle/ref_hook/build_ref.sh with a SeqAn 2.4.0rc2 checkout (SEQAN_DIR=...)"; exit 2; }
W=$(mktemp -d)
python - "$W" <<'PY'
import sys
sys.path.insert(0, ".")
from tests.util import make_case, write_fasta, write_guides
w = sys.argv[1]
case = make_case(seed=2024, contig_lens=[120000, 45, 45, 45, 23, 22, 46, 30000] + [45] * 500, n_guides=12, k=8, pam=None)
write_fasta(f"{w}/genome.fa", [f"ctg{i}" for i in range(len(case.offsets) - 1)], case.ascii, case.offsets)
write_guides(f"{w}/guides.fa", [f"g{i}" for i in range(12)], case.guide_strs)
PY
key() { awk -F'\t' '{ s = (int($2 / 16) % 2); print $1, s, $3, $4, $12 }' "$1" | sort; }
rc=0
$REFBIN/bidir_index -G $W/genome.fa -I $W/refidx >/dev/null || exit 3
HAVE_GPU=0; [ -x build/read_mapping_build/bidir_mapping ] && nvidia-smi -L >/dev/null 2>&1 && HAVE_GPU=1
[ $HAVE_GPU = 1 ] && build/read_mapping_build/bidir_index -G $W/genome.fa -I $W/ouridx >/dev/null
for k in 0 1 2 3 4 5 6 7 8; do
  for pam in "" "-P AG"; do
    $REFBIN/bidir_mapping -G $W/genome.fa -I $W/refidx -R $W/guides.fa -M $k -T 4 $pam -O $W/ref.sam >/dev/null || { echo "reference failed k=$k"; rc=1; continue; }
    oracle/oracle_bidir_mapping -G $W/genome.fa -R $W/guides.fa -M $k $pam -O $W/oracle.sam >/dev/null
    if ! diff <(key $W/ref.sam) <(key $W/oracle.sam) >/dev/null; then echo "k=$k $pam: HIT SET differs (reference vs oracle)"; rc=1
    elif ! cmp -s $W/ref.sam $W/oracle.sam; then echo "k=$k $pam: hit set equal, bytes differ (order / FLAG / MD): $(diff $W/ref.sam $W/oracle.sam | head -4)"; rc=1
    else echo "k=$k $pam: oracle == reference ($(wc -l < $W/ref.sam) records)"; fi
    if [ $HAVE_GPU = 1 ]; then
      build/read_mapping_build/bidir_mapping -G $W/genome.fa -I $W/ouridx -R $W/guides.fa -M $k -T 4 $pam -O $W/ours.sam >/dev/null
      cmp -s $W/ref.sam $W/ours.sam && echo "k=$k $pam: CUDA drop-in == reference" || { echo "k=$k $pam: CUDA drop-in differs from the reference"; rc=1; }
    fi
  done
done
rm -rf "$W"
exit $rc
