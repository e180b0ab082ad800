"""Host-side pieces of the measurement contract that need no GPU: the algorithmic-bytes model of
SURVEY.md section 8(d) that bench.py's roofline uses, and the bench command line."""
import os
import subprocess
import sys

from trimem_b200 import workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_model():
    # energy 52, gradient sweep 1 60, sweep 2 84 B/vertex; + (4P + 4) per sweep with repulsion
    assert workload.bytes_energy() == 52.0 and workload.bytes_grad_props() == 60.0 and workload.bytes_grad_sweep() == 84.0
    assert workload.bytes_eg() == 196.0
    P = 18.69
    assert abs(workload.bytes_eg(P, True) - (196.0 + 3 * (4 * P + 4))) < 1e-12
    assert 430.0 < workload.bytes_eg(P, True) < 434.0                 # "432 B/vertex" of the survey
    # 655,362 vertices without repulsion: 128.5 MB
    assert abs(workload.bytes_eg() * workload.N_VERTICES[8] / 1e6 - 128.45) < 0.1
    assert workload.N_VERTICES == {4: 2562, 6: 40962, 8: 655362, 10: 10485762}


def test_bench_command_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--subdiv", "--part-subdiv", "--replicas"):
        assert flag in out.stdout
