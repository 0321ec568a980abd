import json, subprocess, sys, os
root = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for lib in sys.argv[1:]:
    code = f"""
import sys, json
sys.path.insert(0, {root!r})
from breakdancer_b200 import api
api.LIB_PATH = {os.path.join(root, 'breakdancer_b200', lib)!r}
sys.argv = ['bench.py', '--steps', '10', '--warmup', '3', '--no-cpu']
import runpy
runpy.run_path({os.path.join(root, 'bench.py')!r}, run_name='__main__')
"""
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root)
    try:
        d = json.loads(p.stdout.strip().split("\n")[-1])
        print(lib, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["kernel_ms_per_step"]["k1_classify"], flush=True)
    except Exception as e:
        print(lib, "failed", p.stderr[-400:], flush=True)
