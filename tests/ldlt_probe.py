"""GPU probe (not a pytest): residual and phase timing of the damped solve for a list of sizes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxel_slam_b200 as v
c = v.Context()
for n in [int(a) for a in sys.argv[1:]] or [33, 40, 45, 60, 64, 65, 75, 96, 120, 300, 750, 1500]:
    o = c.ldlt_phases(n)
    print(n, "ms=%.3f load=%.2f strips=%.2f schur=%.2f barrier=%.2f la=%.2f resid=%.2e" % (o[0], o[1], o[2], o[3], o[4], o[7], o[8]), flush=True)
