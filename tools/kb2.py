import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# each variant in its own process with a hard timeout (a hung kernel must never eat the GPU budget)
name, spp, variants = sys.argv[1], sys.argv[2], sys.argv[3].split(",")
tmo = int(sys.argv[4]) if len(sys.argv) > 4 else 45
for v in variants:
    try:
        out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "kbench.py"), name, spp, v], capture_output=True, text=True, timeout=tmo)
        print((out.stdout.strip().splitlines() or [out.stderr.strip()[-200:]])[-1], flush=True)
    except subprocess.TimeoutExpired:
        print(f"variant={v} TIMEOUT", flush=True)
