cd $GRAFT_REPO_ROOT
SWEEP_GPUS=4 SWEEP_VARIANTS="dist_nb=4" timeout 600 python tools/option_sweep.py lap3d 100
