mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) "$@" 2>&1 | grep -E "^\{|CGDIST|Error|error|Traceback" ; }
for n in 2 4 8; do run $n bench.py --gpus $n --steps 500 --warmup 10 | tee -a gpurun_out/scale_spmv.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('BENCH', d['n_gpus'], round(d['value'],1), 'GF/s', round(d['ms_per_step']*1e3,1),'us/step kern', round(d['roofline']['kernel_ms']*1e3,1), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(l[:300])
"; done
for n in 1 2 4 8; do run $n tools/bench_cg_dist.py 4096 300 | tee -a gpurun_out/scale_cg.txt; done
for n in 2 8; do run $n tools/bench_cg_dist.py 4096 300 weak | tee -a gpurun_out/scale_cg.txt; done
