#!/usr/bin/env bash
# The DRLZ line of the reference's launch.sh:41 (`seq 1 22 | xargs -I{} -n 1 -P 4 ./drlz.sh {} $PGPATHHDFS $HDFSURI`)
# for one box of B200s.  One worker per GPU; every worker takes the next chromosome from a shared queue, runs
# drlz.sh for it on ITS GPU and comes back for more -- so a GPU never holds two chromosomes at a time (one index
# resident per GPU, no time-slicing between two drlz processes) and no GPU idles while chromosomes are left.
# Same arguments as the reference line: INPATH HDFSURI [first last].
#   DRLZ_GPUS   number of GPUs to use (default: all that nvidia-smi lists)
#   DRLZ_SH     the per-chromosome script (default: drlz.sh next to this file)
INPATH=$1
HDFSURI=$2
FIRST=${3:-1}
LAST=${4:-22}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DRLZ_SH="${DRLZ_SH:-$HERE/drlz.sh}"
NGPU="${DRLZ_GPUS:-$(nvidia-smi -L 2>/dev/null | wc -l)}"
[ "$NGPU" -ge 1 ] 2>/dev/null || NGPU=1

start=$(date +%s)
QUEUE=$(mktemp -d "${TMPDIR:-/tmp}/drlz_queue_XXXXXX")
trap 'rm -rf "$QUEUE"' EXIT
seq "$FIRST" "$LAST" > "$QUEUE/todo"
: > "$QUEUE/failed"

worker() {    # $1 = GPU of this worker
  local gpu=$1 chr
  while :; do
    # pop the first line of the queue under a lock
    chr=$(flock "$QUEUE/lock" bash -c 'head -n 1 "$0/todo"; sed -i 1d "$0/todo"' "$QUEUE")
    [ -n "$chr" ] || break
    DRLZ_DEVICE=$gpu DRLZ_GPUS=1 "$DRLZ_SH" "$chr" "$INPATH" "$HDFSURI" || echo "$chr" >> "$QUEUE/failed"
  done
}

for ((g = 0; g < NGPU; ++g)); do worker "$g" & done
wait
end=$(date +%s)
echo "DRLZ: $((end - start))" >> runtime.log
if [ -s "$QUEUE/failed" ]; then
  echo "drlz_all: failed chromosomes: $(tr '\n' ' ' < "$QUEUE/failed")" >&2
  exit 1
fi
exit 0
