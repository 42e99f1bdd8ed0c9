"""bench.py runs the measurement in a child process and repeats it when the child stops making progress.
CPU test of that plumbing with a stand-in child (LDB200_BENCH_SELFTEST): the result line reaches stdout, a stalled first
attempt is killed and repeated, and the retry is recorded in the line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode, **env):
    e = dict(os.environ, LDB200_BENCH_SELFTEST=mode, **env)
    e.pop('RANK', None)
    e.pop('WORLD_SIZE', None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1'], capture_output=True, text=True,
                       env=e, timeout=120)
    return r


def test_supervisor_passes_the_result_line_through():
    r = _run('ok')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['metric'] == 'selftest' and 'supervisor' not in d


def test_supervisor_repeats_a_stalled_measurement():
    r = _run('stall-once', LDB200_BENCH_STALL_S='4')
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][-1])
    assert d['supervisor']['retries'] == 1 and d['supervisor']['reason'] == 'stalled'
    assert 'no progress' in r.stderr


def test_supervisor_on_a_non_zero_rank_expects_no_result_line():
    e = dict(os.environ, LDB200_BENCH_SELFTEST='ok', RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--gpus', '2', '--steps', '1'],
                       capture_output=True, text=True, env=e, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_phases_completed_before_a_stall_are_not_measured_again():
    r = _run('phases', LDB200_BENCH_STALL_S='3')
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][-1])
    # phase a ran in attempt 0 and was kept; b stalled twice and was measured in attempt 2; c followed it
    assert d['phases'] == dict(a=0, b=2, c=2), d
    assert d['supervisor']['retries'] == 2
    assert r.stderr.count('a: taken from an earlier attempt') == 2
