"""GPU debugging aid: per-leaf gradient error of the CUDA backward vs torch-CPU autograd of the oracle."""
import sys
sys.path.insert(0, "tests")
import numpy as np
import train_util as U

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tree, audio, labels = U.setup(B)
lref, gref, zref = U.oracle_grads(tree, audio, labels)
lcu, gcu, zcu, eng = U.cuda_grads(tree, audio, labels)
print(f"loss ref {lref:.6f} cuda {lcu:.6f}  max|dlogit| {np.abs(zref - zcu).max():.4f}  launches {eng.launch_count()}")
rows = U.compare(gref, gcu)
import re
def order(k):
    return k
for k, rel, nr, nc in sorted(rows, key=lambda r: r[0]):
    flag = " <<<" if (rel > 0.08 or not np.isfinite(rel)) else ""
    print(f"{rel:9.4f}  ref {nr:10.4e} cuda {nc:10.4e}  {k}{flag}")
bad = [r for r in rows if not (r[1] < 0.08)]
print(f"{len(bad)} of {len(rows)} leaves above 8% relative error; worst: {rows[0][0]} {rows[0][1]:.4f}")
