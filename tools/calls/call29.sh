cd $GRAFT_REPO_ROOT
SWEEP_GPUS=2 SWEEP_VARIANTS="dist_nb=4;dist_nb=2;dist_nb=8,order_alpha=75;dist_nb=4,order_alpha=75" timeout 900 python tools/option_sweep.py lap3d 100
