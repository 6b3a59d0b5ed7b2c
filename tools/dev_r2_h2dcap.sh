mkdir -p gpurun_out
: > gpurun_out/r2_h2d_cap.txt
for N in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N tools/dev_h2d_cap.py 2>/dev/null | grep "^N=" >> gpurun_out/r2_h2d_cap.txt
done
cat gpurun_out/r2_h2d_cap.txt
nvidia-smi topo -m | head -14
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" 
