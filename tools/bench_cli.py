#!/usr/bin/env python
"""Time the `gkd fastaDist` command end to end on a synthetic FASTA (file in, report out)."""
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genome.distance_b200 as gkd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
path = "/tmp/gkd_cli_bench.fa"
t0 = time.perf_counter()
with open(path, "wb") as f:
    buf = np.empty(length, dtype=np.uint8)
    for g in range(n):
        gkd.synth(buf, 0x5EED0000, g // 10, g % 10, 0.0 if g % 10 == 0 else 0.01)
        f.write(b">g%06d synthetic len=%d\n" % (g, length))
        rows = buf[: (length // 80) * 80].reshape(-1, 80)
        out = np.empty((rows.shape[0], 81), dtype=np.uint8)
        out[:, :80] = rows
        out[:, 80] = 10
        f.write(out.tobytes())
        if length % 80:
            f.write(buf[(length // 80) * 80:].tobytes() + b"\n")
t1 = time.perf_counter()
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "genome", "distance_b200", "gkd")
p = subprocess.run([exe, "fastaDist", "-i", path, "-o", "/tmp/gkd_cli_bench.tbl"], capture_output=True, text=True)
t2 = time.perf_counter()
lines = sum(1 for _ in open("/tmp/gkd_cli_bench.tbl"))
print(json.dumps({"records": n, "bp": length, "fasta_bytes": os.path.getsize(path), "write_fasta_s": t1 - t0,
                  "cli_wall_s": t2 - t1, "pairs": n * (n - 1) // 2, "report_lines": lines, "rc": p.returncode,
                  "log_tail": p.stderr.strip().split("\n")[-3:]}))
