"""The host front-end runs its passes on all threads; its outputs must not depend on the thread count
(the op list is part of the bit-exact contract, and the task / operand order decides the rounding)."""
import json, os, subprocess, sys

import conftest  # noqa: F401  (puts tools/ on sys.path)

HERE = os.path.dirname(os.path.abspath(__file__))


def run(path, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    r = subprocess.run([sys.executable, os.path.join(HERE, "_hash_worker.py"), str(path)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_planner_and_compiler_do_not_depend_on_thread_count(sg, tmp_path):
    # 3D 36^3: 0.6 M ops, large enough for the multi-threaded counting sort and task construction
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap3d", 36)
    path = os.path.join(str(tmp_path), "lap3d_36.mtx")
    gen_mtx.write_mtx(path, n, r, c, v, False)
    one, many, odd = run(path, 1), run(path, 8), run(path, 3)
    assert all(v != 0 for v in one.values())
    assert one == many == odd
