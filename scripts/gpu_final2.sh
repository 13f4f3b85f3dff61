bash scripts/gpu_final.sh
bash scripts/gpu_sanitize.sh > gpurun_out/sanitizer_r02.txt 2>&1; tail -16 gpurun_out/sanitizer_r02.txt
