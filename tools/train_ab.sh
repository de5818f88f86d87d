# A/B of the batched weight packing / data-gradient re-layout knobs on the training step (config 4)
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  SEDT_PACK_BATCH=$1 SEDT_DGRAD_BATCH=$2 timeout 300 python bench.py --mode train 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pack_batch=$1 dgrad_batch=$2', round(d['value']), round(d['ms_per_step'],3), d['phases_ms'], 'loss', d['loss'])"
done
