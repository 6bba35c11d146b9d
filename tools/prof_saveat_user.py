"""Development aid: one saveat configuration with the NVRTC twin of the built-in Lorenz system (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quick_bench as Q
layout = int(sys.argv[1]); n = int(sys.argv[2]); dt = float(sys.argv[3])
Q.probe_saveat(n, dt, layout, reps=1, user=True)
