"""bench.py end to end on the CPU (tests/bench_dry_run.py: inert CUDA stand-ins, emulated kernels, a tiny clip): both
arms, the look-ahead A/B, the failing pre-flight child and its fallback, the CPU baseline and the JSON assembly all
execute, exactly one JSON line reaches stdout and it carries every key the contract names."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT

REQUIRED = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline']


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'bench_dry_run.py'), '--steps', '6', '--warmup', '3',
                        *extra], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0]), r.stderr


def test_bench_control_flow_runs_and_prints_one_json_line():
    line, err = _run()
    for k in REQUIRED:
        assert k in line, k
    assert line['steps'] == 6 and line['n_gpus'] == 1 and line['higher_is_better'] is True
    assert line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['d2h_bytes_per_step'] > 0
    assert line['config']['workload'].startswith('tiny')
    # no CUDA here: the pre-flight child must fail and switch the optional forms (and only them) off
    assert line['config']['optional_forms']['checked'] is True and 'pre-flight' in line['config']['optional_forms']['note']
    assert line['config']['optional_forms']['conv_epilogues'] is False
    assert line['config']['encoder_lookahead'] is False
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['value'] > 0
    assert line['warmup'] >= 11                         # graphs requested: warm-up covers every capture variant


def test_bench_with_lookahead_ab_and_no_preflight():
    line, err = _run('--skip-preflight', '--no-cpu-baseline')
    ab = line['config']['encoder_lookahead_ab_ms']
    assert ab and set(ab) == {'with', 'without', 'kept'}
    assert line['config']['encoder_lookahead'] == ab['kept']
    assert line['warmup'] == 11 + 20                    # warm-up + two A/B windows, all untimed
    assert line['config']['conv_epilogues'] is not None and line['config']['glue_ops'] is not None
    assert line['roofline'] is not None and line['roofline']['achieved'] > 0
