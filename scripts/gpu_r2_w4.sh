bash scripts/gpu_r2_w.sh 4
