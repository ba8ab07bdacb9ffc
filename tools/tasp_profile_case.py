import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, "ldpc-lib_b200"); sys.path.insert(0, ".")
from codes import load_code
import pyldpcb200 as L
hd, _ = load_code("ref32x16_b")
with L.Decoder(hd, 126, 7) as d:
    print(d.kernel_info())
    d.simulate(2.0, 2000, 50, seed=1)
    r = d.simulate(2.0, 40000, 50, seed=1, stream=1)
    print(r["frames"], r["iter_sum"] / r["frames"], d.last_kernel_ms())
