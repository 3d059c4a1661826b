cd $GRAFT_REPO_ROOT; timeout 60 tools/bin/lu_lab; timeout 60 tools/bin/lu_lab_prof | grep -v "^sweep"
