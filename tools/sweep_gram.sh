# A/B sweep of the Gram-kernel tuning knobs at the bench size (development aid)
set -e
timeout 300 python -m pytest tests/test_gpu_vmlmb.py -m gpu -q -x 2>&1 | tail -3
run() { env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['kernels']
print('$*', 'it/s=%.2f gram=%.3f ms (%.3f) dir=%.3f (%.3f) trial=%.3f' % (d['value'], k['gram']['avg_ms'], k['gram']['frac'], k['direction']['avg_ms'], k['direction']['frac'], k.get('trial', {'avg_ms': 0})['avg_ms']))
"; }
for v in "$@"; do run $v; done
