cd $GRAFT_REPO_ROOT; timeout 60 tools/bin/poll_lat
