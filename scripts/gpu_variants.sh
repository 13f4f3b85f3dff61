mkdir -p gpurun_out
python scripts/b4_variants.py 50000 gpurun_out/b4_variants_100k.json 2>&1 | tail -30
python scripts/b4_variants.py 500000 gpurun_out/b4_variants_1m.json 2>&1 | tail -30
