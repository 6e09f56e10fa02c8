#!/bin/bash
# Runs bench.py once per line of a config file ("label ENV=V ENV=V ... -- extra bench args") and appends one JSON object per
# run ({"label", "env", "line"}) to the output file.  Used under gpurun to sweep kernel variants in a single GPU call.
#   tools/sweep_bench.sh configs.txt gpurun_out/sweep.jsonl [common bench args...]
cfg="$1"; out="$2"; shift 2
: > "$out"
while IFS= read -r line; do
  [ -z "$line" ] && continue
  case "$line" in \#*) continue;; esac
  label="${line%% *}"; rest="${line#* }"
  envs="${rest%%--*}"; extra=""
  case "$rest" in *--*) extra="${rest#*--}";; esac
  res=$(env $envs timeout 600 python bench.py --no-cpu-baseline "$@" $extra 2> "${out%.jsonl}_${label}.err" | tail -1)
  [ -z "$res" ] && res=null
  echo "{\"label\": \"$label\", \"env\": \"$envs\", \"line\": $res}" >> "$out"
done < "$cfg"
