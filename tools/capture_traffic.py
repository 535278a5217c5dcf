#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of the fused kernel, stamped with the hash of the kernel's
device sources (gap_tv_ws.cuh, gap_tv_stream.cuh, ws_inst_r4.cu) as they were when it was measured.  bench.py reports roofline.traffic only while that hash equals the hash of the
sources in the tree (a stale capture reads as null).

    python tools/capture_traffic.py gpurun_out/<report>.ncu-rep [kernel-index]
"""
import csv
import glob
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


KERNEL_SOURCES = ("gap_tv_ws.cuh", "gap_tv_stream.cuh", "ws_inst_r4.cu")     # the device code of the measured kernel


def _code_only(text):
    """C++ source without comments and whitespace: the stamp follows the code, not its commentary."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return re.sub(r"\s+", "", text)


def source_hash():
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        p = os.path.join(ROOT, "sci-algorithms_b200", "csrc", name)
        h.update(name.encode())
        h.update(_code_only(open(p).read()).encode())
    return h.hexdigest()[:16]


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d = dict(zip(rows[0], zip(rows[1], rows[2 + idx])))

    def val(k):
        unit, v = d[k]
        scale = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.}
        return float(v.replace(",", "")) * scale.get(unit, 1.)
    rec = {"kernel": d["Kernel Name"][1][:120], "report": os.path.basename(rep),
           "dram_bytes_read_per_launch": val("dram__bytes_read.sum"),
           "dram_bytes_write_per_launch": val("dram__bytes_write.sum"),
           "gpu_time_s_under_ncu": val("gpu__time_duration.sum"), "source_hash": source_hash(),
           "note": "one ncu --set full capture of one launch on the 3840x2160x24 scene (profiles/prof_driver.py)"}
    rec["dram_bytes_per_launch"] = rec["dram_bytes_read_per_launch"] + rec["dram_bytes_write_per_launch"]
    json.dump(rec, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
