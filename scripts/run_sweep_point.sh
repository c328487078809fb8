N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N "$@"; }
summ() { python -c "
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], 'value %.4g ms/step %.4f e2e %.4g (%.3f ms) sweep %s' % (d['value'], d['ms_per_step'], d['e2e'].get('value',0), d['e2e'].get('ms_per_step',0), json.dumps(d.get('sweep'))[:400]))
print('   collective:', d['impl_notes']['collective'][:90], '| kernels', d['roofline']['kernel_ms'], 'eager', d['impl_notes']['launch'][-40:])
" $1; }
run --steps 20 --warmup 5 > gpurun_out/r2h_n${N}_peer_s20.json 2> gpurun_out/r2h_n${N}_peer_s20.err; echo rc=$?; summ gpurun_out/r2h_n${N}_peer_s20.json
run --steps 20 --warmup 5 --collective nccl > gpurun_out/r2h_n${N}_nccl_s20.json 2> gpurun_out/r2h_n${N}_nccl_s20.err; echo rc=$?; summ gpurun_out/r2h_n${N}_nccl_s20.json
run --steps 1000 --warmup 20 > gpurun_out/r2h_n${N}_peer_s1000.json 2> gpurun_out/r2h_n${N}_peer_s1000.err; echo rc=$?; summ gpurun_out/r2h_n${N}_peer_s1000.json
run --steps 1000 --warmup 20 --collective nccl > gpurun_out/r2h_n${N}_nccl_s1000.json 2> gpurun_out/r2h_n${N}_nccl_s1000.err; echo rc=$?; summ gpurun_out/r2h_n${N}_nccl_s1000.json
tail -3 gpurun_out/r2h_n${N}_peer_s20.err
