# whole-step A/B of environment switches: scripts/ab_env.sh "VAR=val VAR2=val" "..." (an empty string = defaults)
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --no-cpu-baseline --no-check --steps 40 > gpurun_out/ab_env_$i.json 2> gpurun_out/ab_env.err; tail -c 300 gpurun_out/ab_env.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab_env_$i.json')); print('STEP [$cfg]', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k[:14]:round(v['ms'],3) for k,v in d['roofline_by_kernel'].items()})"
done
