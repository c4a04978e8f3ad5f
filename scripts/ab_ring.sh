# sweep of the cp.async operand-ring depth of the affine kernels (SEMB_AFF_RING = fwd,fwdb,red,redb,app,appb): per-kernel rates and the whole step
for cfg in 0,8,4,4,0,0 4,8,8,8,4,4 2,4,4,8,2,2 0,8,8,8,0,4 0,0,0,0,0,0; do
  export SEMB_AFF_RING=$cfg
  echo "== SEMB_AFF_RING=$cfg"
  python scripts/bench_layers.py --only affine 2>&1 | tail -27
  python bench.py --no-cpu-baseline --no-check > gpurun_out/r02_bench_ring_$cfg.json 2> gpurun_out/r02_bench_ring.err; tail -c 300 gpurun_out/r02_bench_ring.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r02_bench_ring_$cfg.json')); print('STEP', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['ms'],3) for k,v in d['roofline_by_kernel'].items()})"
done
