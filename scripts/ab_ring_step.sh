# whole-step A/B of the operand-ring depths (SEMB_AFF_RING = fwd,fwdb,red,redb,app,appb)
for cfg in "$@"; do
  export SEMB_AFF_RING=$cfg
  python bench.py --no-cpu-baseline --no-check --steps 40 > gpurun_out/r02_bench_ring_$cfg.json 2> gpurun_out/r02_bench_ring.err; tail -c 300 gpurun_out/r02_bench_ring.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r02_bench_ring_$cfg.json')); print('STEP $cfg', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k[:14]:round(v['ms'],3) for k,v in d['roofline_by_kernel'].items()})"
done
