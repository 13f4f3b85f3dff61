mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_4gpu_b.json 2> gpurun_out/bench_4gpu_b.err
tail -c 2400 gpurun_out/bench_4gpu_b.json | cut -c1-1500; tail -3 gpurun_out/bench_4gpu_b.err
for i in 0 1 2 3; do cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $i | tr 'A-Z' 'a-z' | sed 's/^0000//')/numa_node; done; nproc; lscpu | grep -i "numa\|socket\|model name" | head
