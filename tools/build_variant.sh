#!/bin/bash
# usage: tools/build_variant.sh <name> [ENV=VAL ...]   -> risc0_b200/lib/libr0b200_<name>.so (experiment builds)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
for kv in "$@"; do export "$kv"; done
export R0B200_VARIANT=$name
[ -d risc0_b200/build_$name ] || cp -r risc0_b200/build risc0_b200/build_$name
rm -f risc0_b200/build_$name/*.cubin risc0_b200/build_$name/embed_cubins.* risc0_b200/build_$name/gen_eval_check_rv32im.o
python tools/gen_eval_check.py rv32im --from-ir > /dev/null
python -m risc0_b200.build -v 2>&1 | grep -E "spill" | awk '{s+=$6} END {print "variant '$name' total spill-store bytes:", s}'
