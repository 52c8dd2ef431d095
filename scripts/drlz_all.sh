#!/usr/bin/env bash
# The DRLZ line of the reference's launch.sh:41 (`seq 1 22 | xargs -I{} -n 1 -P 4 ./drlz.sh {} $PGPATHHDFS $HDFSURI`)
# for one box of B200s: the chromosomes are spread over the visible GPUs, one drlz process per GPU at a time
# (each process builds that chromosome's index on its own GPU, parses all its haplotypes and exits, so only one
# index is resident per GPU).  Same arguments as the reference line: INPATH HDFSURI [first last].
#   DRLZ_GPUS   number of GPUs to use (default: all that nvidia-smi lists)
INPATH=$1
HDFSURI=$2
FIRST=${3:-1}
LAST=${4:-22}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NGPU="${DRLZ_GPUS:-$(nvidia-smi -L 2>/dev/null | wc -l)}"
[ "$NGPU" -ge 1 ] 2>/dev/null || NGPU=1

start=$(date +%s)
# xargs hands chromosome k to slot (k mod NGPU); -P NGPU keeps every GPU busy with one chromosome
seq "$FIRST" "$LAST" | xargs -I{} -n 1 -P "$NGPU" bash -c \
  'DRLZ_DEVICE=$(( {} % '"$NGPU"' )) DRLZ_GPUS=1 "'"$HERE"'/drlz.sh" {} "'"$INPATH"'" "'"$HDFSURI"'"'
rc=$?
end=$(date +%s)
echo "DRLZ: $((end - start))" >> runtime.log
exit $rc
