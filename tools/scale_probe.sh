#!/bin/bash
# gpurun --gpus 2 -- 'bash tools/scale_probe.sh': step time and host enqueue time at 1 and 2 GPUs
show() { grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d.get('ms_per_step_by_rank'), 'ms_per_step', round(d['ms_per_step'],4), 'host', round(d['host_enqueue_ms_per_step'],4), 'cpus', d['host_cpus'], 'value %.3e' % d['value'], {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | show "2gpu"
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | show "1gpu"
