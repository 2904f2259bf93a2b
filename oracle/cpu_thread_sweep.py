import sys, os, time
sys.path.insert(0, os.getcwd())
import bench
for th in (8, 16, 32, 64, 128):
    r, s, c = bench.cpu_reference_rate(2, threads=th)
    print("threads", th, "steps/s", r, flush=True)
