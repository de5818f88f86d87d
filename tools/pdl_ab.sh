set -x
timeout 300 python -m pytest tests/test_gpu_forward.py tests/test_gpu_conv_tc.py tests/test_gpu_ops.py -x -q 2>&1 | tail -4
SEDT_PDL=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=0', d['value'], d['ms_per_step'], d['e2e']['value'])"
SEDT_PDL=1 timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=1', d['value'], d['ms_per_step'], d['e2e']['value'])"
SEDT_PDL=0 timeout 300 python bench.py --mode train 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train PDL=0', d['value'], d['ms_per_step'], d['phases_ms'])"
SEDT_PDL=1 timeout 300 python bench.py --mode train 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train PDL=1', d['value'], d['ms_per_step'], d['phases_ms'])"
