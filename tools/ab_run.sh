#!/bin/bash
# usage: tools/ab_run.sh <out.log> <preset> <batch> <variant> [<variant> ...]   (A/B timing of variants/lib_<variant>.so on the GPU box)
out=$1; preset=$2; batch=$3; shift 3
for v in "$@"; do
  timeout 300 python tools/prof_run.py --preset "$preset" --batch "$batch" --steps 2 --warmup 1 --check --lib variants/lib_$v.so --tag "$v" >> "$out" 2>> "$out.err" || echo "{\"tag\": \"$v\", \"failed\": $?}" >> "$out"
done
