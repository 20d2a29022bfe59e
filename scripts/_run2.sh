for n in 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/r12_err.log | tail -1 > gpurun_out/r12_bench_n$n.json
tail -3 gpurun_out/r12_err.log
python -c "import json;d=json.load(open('gpurun_out/r12_bench_n$n.json'));print('N=',d['n_gpus'],'ms',round(d['ms_per_step'],4),'value',d['value'],'e2e ms',round(d['e2e']['ms_per_step'],3),d['config']['launch'])"
done
