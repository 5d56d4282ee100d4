#!/bin/bash
# Build the library of another commit (or of the working tree with one source file swapped) into
# baybe_b200/_C/variants/<name>.so for same-box A/B timing (scripts/ab_kernel.py).
# usage: scripts/build_variant.sh <name> <commit> [replacement fused_ts.cu]
set -e
NAME=$1; COMMIT=$2; REPL=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WT=/tmp/wt_$NAME
rm -rf $WT; git -C $ROOT worktree prune; git -C $ROOT worktree add -f --detach $WT $COMMIT > /dev/null 2>&1
[ -n "$REPL" ] && cp "$REPL" $WT/baybe_b200/csrc/fused_ts.cu
if [ "$COMMIT" = "WORKTREE" ]; then :; fi
(cd $WT && python -m baybe_b200.build > /dev/null 2>&1)
mkdir -p $ROOT/baybe_b200/_C/variants
cp $WT/baybe_b200/_C/libbaybe_b200.so $ROOT/baybe_b200/_C/variants/$NAME.so
git -C $ROOT worktree remove --force $WT
echo "built variant $NAME from $COMMIT"
