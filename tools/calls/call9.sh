cd $GRAFT_REPO_ROOT
python tools/calls/wd.py
