#!/bin/bash
# Our bedmap against the reference bedmap, byte for byte (stdout, stderr, exit code), on a handful of complete command lines
# over small synthetic files: a quick end-to-end check of the command-line tool on a GPU box without pytest or torch.
# usage: profiles/tools/cli_spot_check.sh [out.log]
R=${GRAFT_REPO_ROOT:-/root/repo}
OURS=$R/bedops_b200/bin/bedmap; REF=$R/oracle/_ref/bin/bedmap
D=$(mktemp -d /dev/shm/cli_spot.XXXX); LOG=$(realpath -m ${1:-/dev/stdout})
$R/bedops_b200/bin/synth-bed 3000 2 7.0 1.0 5 $D/r.bed chr1 chr2 chrX > /dev/null
$R/bedops_b200/bin/synth-bed 30000 1 5.5 1.0 5 $D/m.bed chr1 chr2 chrX > /dev/null
cd $D; ok=0; bad=0
while IFS= read -r line; do
  [ -z "$line" ] && continue
  eval "timeout 30 $REF $line" > e.out 2> e.err; erc=$?
  eval "timeout 30 $OURS $line" > g.out 2> g.err; grc=$?
  if [ $erc -eq $grc ] && cmp -s e.out g.out && cmp -s e.err g.err; then ok=$((ok+1)); else bad=$((bad+1)); echo "DIFF: $line (rc $erc / $grc)" >> $LOG; head -c 300 g.err >> $LOG; head -c 300 e.err >> $LOG; diff e.out g.out | head -8 >> $LOG; fi
done <<'CASES'
--echo --count --mean --bases r.bed m.bed
--range 100 --echo --sum --max --min --prec 3 r.bed m.bed
--fraction-both 0.5 --faster --indicator --echo-map-id --multidelim , r.bed m.bed
--ec --delim '\t' --skip-unmapped --echo --median --kth 0.3 --mad --tmean 0 0.2 r.bed m.bed
--chrom chr2 --sci --stdev --variance --cv --wmean r.bed m.bed
--exact --echo-ref-row-id --echo-ref-size --echo-ref-name m.bed
--bp-ovr 25 --kth 1 --kth 0 --echo-map-range --echo-overlap-size --bases-uniq --bases-uniq-f r.bed m.bed
--prec '' --mean r.bed m.bed
--multidelim --x r.bed m.bed
--ec --count r.bed nofile.bed
--count --mad
CASES
echo "cli_spot_check: $ok identical, $bad different" >> $LOG
rm -rf $D
[ $bad -eq 0 ]
