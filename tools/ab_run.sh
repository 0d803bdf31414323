#!/bin/bash
# Run a command on the GPU box with an alternative build of the library swapped in (tools/ab_build.py), then restore.
#   bash tools/ab_run.sh <name> <command ...>        name = head: the current build
name=$1; shift
cd ${GRAFT_REPO_ROOT:-$(dirname $0)/..}
lib=aznet_b200/libaznet_b200.so
[ -f /tmp/azn_lib_head.so ] || cp $lib /tmp/azn_lib_head.so
if [ "$name" != head ]; then cp aznet_b200/build/ab/lib_$name.so $lib || exit 1; fi
touch $lib                                   # newer than the sources: _lib.build() must not rebuild it
"$@"; rc=$?
cp /tmp/azn_lib_head.so $lib; touch $lib
exit $rc
