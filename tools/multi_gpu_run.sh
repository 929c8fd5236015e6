N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -E "NUMA|Socket|^CPU\(s\)"
$TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; tail -c 2500 gpurun_out/r2_scale_n$N.json | head -c 2500; echo
python bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_ref_n$N.json 2>/dev/null; head -c 600 gpurun_out/r2_scale_ref_n$N.json; echo
$TR --master-port 29502 tools/ddp_train_bevtxt.py --steps 20 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/r2_ddp_bevtxt_n$N.json
$TR --master-port 29503 tools/ddp_train_bevtxt.py --steps 8 --stock 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/r2_ddp_bevtxt_stock_n$N.json
bash tools/bench_configs.sh r2 $N
