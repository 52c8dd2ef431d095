#!/usr/bin/env bash
# Drop-in replacement of the reference's drlz.sh (same arguments: CHR INPATH HDFSURI).
# Only the launcher line differs: the Spark submit is replaced by the native B200 binary.  With an hdfs:// URI the
# binary lists and fetches the rows with `hdfs dfs -ls -C` / `-get` when "$INPATH/*.NN" matches nothing locally (Spark
# read them from HDFS), stages its outputs locally and uploads them with `hdfs dfs -put -f` (the mkdirs below are the
# reference's).  It exits non-zero when no row matches the glob, like Spark's "Path does not exist".
CHR=$1
INPATH=$2
DRLZOUT=$2drlz
GAPLESSOUT=$2gapless
HDFSURI=$3
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DRLZ_BIN="${DRLZ_BIN:-$HERE/../pangenspark_b200/bin/drlz}"
REFDIR="${DRLZ_REFDIR:-/data/grch37}"

case "$HDFSURI" in
  hdfs://*) hdfs dfs -mkdir -p "$DRLZOUT"; hdfs dfs -mkdir -p "$GAPLESSOUT" ;;
  *) mkdir -p "$DRLZOUT" "$GAPLESSOUT" ;;
esac

i=$(printf "%02d" "$CHR")
echo Submitting job with args $i "$INPATH/*.${i}" $DRLZOUT

# quoted glob: the binary expands it itself (Spark did the same for the HDFS path)
"$DRLZ_BIN" "$REFDIR/radix$1" "$REFDIR/chr${1}.fa" "$INPATH/*.${i}" "$HDFSURI" "$DRLZOUT" "$GAPLESSOUT/chr${i}.fa" > "log$1" 2>&1
