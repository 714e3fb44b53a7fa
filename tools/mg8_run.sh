# the multi-GPU record of a round: weak scaling at 8 (target, config2), strong scaling (config4 x 4, config5 x 8) and the
# 1-GPU lines they are compared with.   gpurun --gpus 8 -- bash tools/mg8_run.sh
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc port args...
  name=$1; n=$2; port=$3; shift 3
  timeout 200 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  tail -c 400 gpurun_out/$name.err | grep -i "error\|Traceback" ; head -c 260 gpurun_out/$name.json; echo
}
run bench_r2_target_8gpu 8 29521
run bench_r2_config2_8gpu 8 29522 --workload config2
run bench_r2_config4_4gpu_strong 4 29523 --workload config4 --scaling strong
run bench_r2_config5_8gpu_strong 8 29524 --workload config5 --scaling strong
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --workload config2 --steps 20 --warmup 3 > gpurun_out/bench_r2_config2_1gpu.json 2> gpurun_out/b21.err &
CUDA_VISIBLE_DEVICES=1 timeout 200 python bench.py --workload config4 --steps 20 --warmup 3 > gpurun_out/bench_r2_config4_1gpu.json 2> gpurun_out/b41.err &
CUDA_VISIBLE_DEVICES=2 timeout 200 python bench.py --workload config5 --steps 20 --warmup 3 > gpurun_out/bench_r2_config5_1gpu.json 2> gpurun_out/b51.err &
wait
for f in config2 config4 config5; do head -c 200 gpurun_out/bench_r2_${f}_1gpu.json; echo; done
