cd $GRAFT_REPO_ROOT
timeout 600 python tools/probes/entry_profile.py 2>&1 | tail -28
