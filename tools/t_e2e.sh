python bench.py --workload T --steps 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('T', d['ms_per_step'], 'host', d['host_enqueue_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
