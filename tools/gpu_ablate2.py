"""Ablation sweep of one conv shape with an explicit flag list (profiling experiment)."""
import os, sys, subprocess
shape = sys.argv[1].split("x")
for ab in sys.argv[2:]:
    env = dict(os.environ, CAL_DEBUG_ABLATE=ab)
    subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "gpu_ablate.py"), "child", *shape], env=env)
