"""Measure roofline.traffic for bench.py: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of every
library kernel of one headline step, with ncu on the GPU box, stamped with the state of the kernel sources.

    python tools/measure_traffic.py            # on a GPU box (under gpurun); writes profiles/traffic.json

bench.py puts `traffic` into its JSON line only when the stamp (sha256 of deformablelka_b200/csrc + include) matches the
sources it is running: a number measured on other code is not reported.  One GPU only; never under torchrun."""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha() -> str:
    h = hashlib.sha256()
    for d in ("deformablelka_b200/csrc", "include"):
        for f in sorted(os.listdir(os.path.join(ROOT, d))):
            h.update(f.encode())
            h.update(open(os.path.join(ROOT, d, f), "rb").read())
    return h.hexdigest()[:16]


def launch_name(kernel: str):
    """ncu prints the C++ kernel name; map it to the name the library's profiler (dlka_profile_summary) uses."""
    k = kernel
    if "deform3d_tc_kernel" in k or "deform3d_ps_kernel" in k:
        return "tc_deform3d_chain"
    if "conv_tiled_kernel" in k:
        return "tc_conv_tiled"
    m = re.search(r"dwconv_smem_kernel<\(int\)(\d+), \(int\)(\d+)", k) or re.search(r"dwconv_smem_kernel<(\d+), (\d+)", k)
    if m:
        return {("5", "5"): "dwconv3d_smem_k5", ("7", "7"): "dwconv3d_smem_k7d3"}.get((m.group(1), m.group(2)), "dwconv3d_smem_aniso")
    if "tc_igemm_kernel" in k or "dense_stream_kernel" in k:
        return "tc_dense"
    return None


def main():
    out_csv = os.path.join(ROOT, "gpurun_out", "traffic_raw.csv")
    os.makedirs(os.path.dirname(out_csv), exist_ok=True)
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "--csv", "--log-file", out_csv, sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3",
           "--no-e2e", "--no-cpu-baseline", "--no-profile-pass", "--no-other-configs"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    rows = [r for r in csv.reader(open(out_csv)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = {}
    for r in rows[1:]:
        name = launch_name(r[ki])
        if name is None:
            continue
        val = float(r[vi].replace(",", ""))
        unit = r[ui].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        d = per.setdefault(name, {"launches": 0, "bytes": 0.0, "ns": 0.0})
        if r[mi] == "gpu__time_duration.sum":
            d["ns"] += val * mult
            d["launches"] += 1
        else:
            d["bytes"] += val * mult
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    res = {"_source": "tools/measure_traffic.py: ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, headline shape, "
                      "averaged over the launches of bench.py --steps 1 --warmup 3", "csrc_sha": csrc_sha(), "git_head": head,
           "kernels": {k: {"dram_bytes_per_launch": v["bytes"] / max(v["launches"], 1), "launches": v["launches"],
                           "ncu_ms_per_launch": v["ns"] / max(v["launches"], 1) / 1e6} for k, v in per.items()}}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
