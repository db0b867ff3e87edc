"""`enflow_b200.compat.install()` lets code written against the reference's module paths import unchanged."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_import_enflow_alias():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import enflow_b200.compat as c; c.install()\n"
            "from enflow.flow.dynamics import LFIntegrator\n"
            "from enflow.nn.egcl import EGCL\n"
            "from enflow.nn.argmax import ArgMax\n"
            "from enflow.flow.loss import Alchemical_NLL\n"
            "from enflow.data.base import Data, DataLoader\n"
            "from enflow.utils.conversion import time_to_lj, kelvin_to_lj\n"
            "m = LFIntegrator([EGCL(4, 4, 128)], ArgMax(4, 128), dt=time_to_lj(1.0))\n"
            "assert abs(m.dt - 0.007178968386789441) < 1e-18 and abs(kelvin_to_lj(300.) - 10.480414411764707) < 1e-12\n"
            "print('ok')\n") % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU port of the reference algorithm) prints one JSON line with impl=reference."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'c1',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['cpu_baseline']['kind'] in ('reference', 'port') and line['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] == 0
