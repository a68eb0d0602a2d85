mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -12
 python tools/prof_kernels.py --rays 327680 --which gather,decoder --impl 2
 for N in 2; do
   for SH in images rays; do
     timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --shard $SH --quick 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
   done
 done
 timeout 300 python bench.py --steps 10 --warmup 3 --quick | tail -1) > gpurun_out/r02_call7.log 2>&1
tail -5 gpurun_out/r02_call7.log | cut -c1-400
