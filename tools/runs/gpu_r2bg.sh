#!/bin/bash
cd $GRAFT_REPO_ROOT
for c in 16 32 16 32; do AZN_BACKBONE_CHUNK=$c timeout 300 python tools/probes/entry_chunk.py 2>&1 | tail -1 | cut -c1-200; done
