"""The JSON contract of bench.py: the committed line of the last GPU run (profiles/) carries every key the driver and the
judge read, and the CPU reference arm -- runnable here -- prints the same shape."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
             'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d['metric'] == 'train_steps_per_sec' and d['unit'] == 'steps/s' and d['higher_is_better'] is True
    assert d['scaling'] in ('weak', 'strong') and d['vs_baseline'] is None and d['data'] == 'synthetic' and d['dtype'] == 'f32'
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(d['e2e'])
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(d['cpu_baseline']) and d['cpu_baseline']['kind'] in ('port', 'reference')
    assert d['value'] > 0 and d['ms_per_step'] > 0


def test_committed_gpu_bench_line_has_every_contract_key():
    lines = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r0*_bench_line.json')))
    assert lines, 'no committed bench line under profiles/'
    d = json.load(open(lines[-1]))
    _check_common(d)
    assert d.get('impl', 'ours') == 'ours' and d['n_gpus'] == 1 and d['warmup'] >= 3
    assert d['gpu_launches'] > 0 and d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    r = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r) and r['bound'] in ('hbm', 'tensor')
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    c = d['clocks']
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(c)
    assert not set(c['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    assert abs(d['value'] - 1e3 / d['ms_per_step'] * d.get('batches_per_sync_step', 1)) < 1e-6 * d['value']


def test_reference_arm_prints_the_same_shape_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'lightgcn-gowalla',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, check=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    _check_common(d)
    assert d['impl'] == 'reference' and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['e2e']['value'] == d['value'] == d['cpu_baseline']['value'] and d['cpu_baseline']['cores'] >= 1
    # a rank other than 0 under torchrun prints nothing and exits 0
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''
