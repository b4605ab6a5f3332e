mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/allreduce_check.py > gpurun_out/allreduce_n8.txt 2>&1
grep -E "world|ours|skipped|rror" gpurun_out/allreduce_n8.txt
B="bench.py --gpus 8 --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
timeout 300 $TR --master-port 29522 $B > gpurun_out/bench_n8_peer_auto.json 2> gpurun_out/bench_n8_peer_auto.err
ASR_ALLREDUCE_MULTICAST=0 timeout 300 $TR --master-port 29523 $B > gpurun_out/bench_n8_peer_p2p.json 2> gpurun_out/bench_n8_peer_p2p.err
timeout 300 $TR --master-port 29524 $B --allreduce nccl > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err
timeout 300 $TR --master-port 29525 $B --no-allreduce > gpurun_out/bench_n8_none.json 2> gpurun_out/bench_n8_none.err
for f in peer_auto peer_p2p nccl none; do tail -1 gpurun_out/bench_n8_$f.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['config']['grad_allreduce'][:120])" || tail -3 gpurun_out/bench_n8_$f.err; done
