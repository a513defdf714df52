cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r30_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r30_pytest.log
python tools/ab_kernels.py run nopred base > gpurun_out/r30_ab.log 2>&1; echo "ab rc=$?"
python - <<'PY'
import json
r=json.load(open('gpurun_out/ab_results.json'))
for k,v in r.items():
    print(k, v['matches_first_variant'], 'roll %.3f ms %.4g greedy %.4g | step1M %.2f (%.3f) 8M %.2f (%.3f) 2^16 %.2f | after %.1f (%.3f) env %.2f (%.3f)' % (v['rollout_ms_2^24'], v['rollout_env_steps_per_s'], v['greedy_env_steps_per_s'], v['step_1M_us'], v['step_1M_frac_hbm'], v['step_8M_us'], v['step_8M_frac_hbm'], v['step_2^16_us'], v['afterstates_8M_us'], v['afterstates_frac_hbm'], v['env_step_1M_us'], v['env_step_frac_hbm']))
PY
