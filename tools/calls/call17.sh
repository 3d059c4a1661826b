cd $GRAFT_REPO_ROOT; python tools/calls/dw.py
