cd /root/repo
python -m pytest tests -m gpu -x -q -k "host or rollout or record or stats or c_abi or integration" > gpurun_out/r52_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r52_pytest.log
python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r52_quick.json 2> gpurun_out/r52_quick.err; echo "quick rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r52_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], {k:(v['ms_per_step'] if isinstance(v,dict) else v) for k,v in d['e2e_modes'].items()})
PY
