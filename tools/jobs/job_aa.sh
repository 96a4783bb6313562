# GPU job aa: graph groups / copy-out parts of the one-shot call (C2, pageable out)
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
run1() { timeout 300 python bench.py --steps 5 --warmup 2 --extra 0 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1: e2e pageable ms', round(d['e2e']['ms_per_step'],2), d['e2e'].get('ms_each'), 'pinned out', round(d['e2e_pinned_out']['ms_per_step'],2))"; }
run1 default_32_groups_8_parts
WAE_AUTO_GROUPS=64 run1 64_groups
WAE_AUTO_GROUPS=48 run1 48_groups
WAE_COPY_PARTS=14 run1 14_parts
WAE_AUTO_GROUPS=64 WAE_COPY_PARTS=14 run1 64_groups_14_parts
