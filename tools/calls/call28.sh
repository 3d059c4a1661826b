cd $GRAFT_REPO_ROOT
SWEEP_GPUS=2 SWEEP_VARIANTS="order_alpha=25;order_alpha=75;dist_nb=8;dist_nb=32;mirror_min=2" timeout 900 python tools/option_sweep.py lap3d 100
