import sys, subprocess
sys.path.insert(0, '.')
import adaprox_b200 as A
dev = A.default_device()
print(dev.info())
P = A.generate_planted_lasso(65536, 131072, power_iters=0)
M = P["A"]
for reps in (3, 40, 3, 80):
    for which in (0, 1):
        ms = M.time_kernel(which, reps=reps)
        print(f"which={which} reps={reps} ms={ms:.3f} GB/s={65536*131072*8/ms/1e6:.0f}", flush=True)
p = subprocess.run("nvidia-smi --query-gpu=power.draw,clocks.sm,clocks.mem,temperature.gpu --format=csv,noheader", shell=True, capture_output=True, text=True)
print(p.stdout)
