for p in 2097152 4194304 8388608 16777216 33554432; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pool $p 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pool $p', 'Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'launches',d['gpu_launches'])"
done
