#!/usr/bin/env python
"""Refresh profiles/inst_counts.json from ncu --set full captures (read here, on the CPU box).

usage: tools/update_inst_counts.py <key> <report.ncu-rep> <kernel-name-regex> <env_steps_per_launch> <profiles/summary.txt> [note]

<key> is one of invmgmt / newsvendor / netinv (fused rollout kernels: warp-instructions + DRAM bytes per env-step, tied
to the kernel sources by sha1) or invmgmt_step / newsvendor_step / netinv_step (one-period kernels: DRAM bytes only).
The FIRST profiled launch whose name matches the regex is used."""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    key, rep, pat, steps, summary = sys.argv[1:6]
    note = sys.argv[6] if len(sys.argv) > 6 else ""
    steps = float(eval(steps, {}))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw = list(csv.reader(io.StringIO(out)))
    hdr = raw[0]
    col = {h: i for i, h in enumerate(hdr)}
    # ORGYM_COUNT_SUM=n: one call of the API launches n kernels that belong together (e.g. the newsvendor's order-level
    # kernel + rollout kernel): counts and durations of the first n matching launches are added up
    nsum = int(os.environ.get("ORGYM_COUNT_SUM", "1"))
    rows = [r for r in raw[2:] if re.search(pat, r[col["Kernel Name"]])][:nsum]
    row = rows[0]
    units = raw[1]

    def val1(r, name):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
        return v * scale

    def val(name):
        return sum(val1(r, name) for r in rows)
    inst = val("smsp__inst_executed.sum")
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    dur_us = val("gpu__time_duration.sum")
    du = units[col["gpu__time_duration.sum"]].lower()
    dur_us *= {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(du, 1)
    issue = sum(val1(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") * val1(r, "gpu__time_duration.sum") for r in rows) / \
        sum(val1(r, "gpu__time_duration.sum") for r in rows)      # duration-weighted
    dthr = val1(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    path = os.path.join(ROOT, "profiles", "inst_counts.json")
    J = json.load(open(path))
    e = {"dram_bytes_per_env_step": dram / steps}
    names = " + ".join(r[col['Kernel Name']][:40] for r in rows)
    src = (f"{summary}: {names}, {steps:.0f} env-steps per launch: smsp__inst_executed.sum {inst:.0f} "
           f"({inst * 32 / steps:.1f} per warp-step), issue-slot utilisation {issue:.1f} %, dram read+write {dram / 1e9:.3f} GB, "
           f"{dthr:.1f} % DRAM throughput, {dur_us:.1f} us under ncu")
    if not key.endswith("_step"):
        import bench
        e["warp_inst_per_env_step"] = inst / steps
        e["src_sha1"] = bench.kernel_source_hash(key)
        e = {"warp_inst_per_env_step": e["warp_inst_per_env_step"], "dram_bytes_per_env_step": e["dram_bytes_per_env_step"],
             "src_sha1": e["src_sha1"]}
    e["source"] = src + (" -- " + note if note else "")
    J[key] = e
    json.dump(J, open(path, "w"), indent=1)
    print(key, json.dumps(e, indent=1))


if __name__ == "__main__":
    main()
