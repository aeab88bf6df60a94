bash tools/gpu_run_all.sh r01j k_tile
bash tools/gpu_variants.sh
