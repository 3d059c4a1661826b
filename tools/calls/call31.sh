cd $GRAFT_REPO_ROOT
SWEEP_GPUS=2 SWEEP_VARIANTS="dist_nb=1;dist_nb=16;dist_nb=64;dist_nb=16,mirror_min=2;dist_nb=16,mirror_min=1000" timeout 600 python tools/option_sweep.py lap3d 64
