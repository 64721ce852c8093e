for e in 0 1 2 3 4 8 16 32 64 127; do
  FC_ACC_EXP=$e python bench.py --config $1 --no-extra --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exp', $e, 'cfg $1', d['detail']['merge_stages_us'], 'scan', round(d['detail']['scan_ms'],3))"
done
