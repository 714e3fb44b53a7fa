TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
for n in 4 8; do
  timeout 200 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_r2_target_${n}gpu.json 2> gpurun_out/bt$n.err; tail -c 1500 gpurun_out/bench_r2_target_${n}gpu.json | head -c 700; echo; tail -2 gpurun_out/bt$n.err
done
timeout 200 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --workload config2 > gpurun_out/bench_r2_config2_8gpu.json 2> gpurun_out/bc8.err; head -c 300 gpurun_out/bench_r2_config2_8gpu.json; echo
timeout 200 $TR --nproc-per-node 4 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 3 --workload config4 --scaling strong > gpurun_out/bench_r2_config4_4gpu_strong.json 2> gpurun_out/bs4.err; head -c 300 gpurun_out/bench_r2_config4_4gpu_strong.json; echo; tail -2 gpurun_out/bs4.err
timeout 200 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --workload config5 --scaling strong > gpurun_out/bench_r2_config5_8gpu_strong.json 2> gpurun_out/bs8.err; head -c 300 gpurun_out/bench_r2_config5_8gpu_strong.json; echo; tail -2 gpurun_out/bs8.err
timeout 200 python bench.py --workload config4 --steps 20 --warmup 3 > gpurun_out/bench_r2_config4_1gpu.json 2> gpurun_out/b41.err & 
CUDA_VISIBLE_DEVICES=1 timeout 200 python bench.py --workload config5 --steps 20 --warmup 3 > gpurun_out/bench_r2_config5_1gpu.json 2> gpurun_out/b51.err &
wait
head -c 300 gpurun_out/bench_r2_config4_1gpu.json; echo; head -c 300 gpurun_out/bench_r2_config5_1gpu.json; echo
