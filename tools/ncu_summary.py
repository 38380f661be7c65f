#!/usr/bin/env python
"""Turn a gpurun session directory (launch list + ncu full capture + bench json) into the tracked
summaries under profiles/.   python tools/ncu_summary.py gpurun_out/r01a r01"""
import collections
import csv
import json
import os
import re
import subprocess
import sys


def short(name):
    m = re.search(r"stag_kernel<[^>]*>", name)
    if m:
        return m.group(0)
    m = re.search(r"normal_kernel<[^>]*>", name)
    if m:
        return m.group(0)
    m = re.search(r"normal1_kernel<[^>]*>", name)
    if m:
        return m.group(0)
    m = re.search(r"coarse_ring_kernel<[^>]*>", name)
    if m:
        return m.group(0)
    m = re.search(r"ew_kernel<[^,]*, *(?:glb::)?(\w+)", name)
    if m:
        return "ew_kernel<" + m.group(1) + ">"
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("glb::", "")


def traffic_key(name):
    """bench.py's name for the kernels whose DRAM traffic it reports (profiles/ncu_traffic.json)"""
    m = re.match(r"normal1?_kernel<([01]),", name)
    if m:
        return "normal_kernel_fused" if m.group(1) == "1" else "normal_kernel"
    if name.startswith("cg_update_kernel"):
        return "cg_update_kernel"
    if name.startswith("stag_kernel"):
        return "stag_kernel"
    if name.startswith("coarse_ring_kernel"):
        return "coarse_ring_kernel"
    return None


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    out = []
    for r in data:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for w in want:
            if w in hdr:
                d[w] = r[hdr.index(w)] + " " + units[hdr.index(w)]
        out.append(d)
    return out


def gb(s):
    v, u = s.split()[:2]
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u]


def main():
    src, tag = sys.argv[1], sys.argv[2]
    os.makedirs("profiles", exist_ok=True)
    md = ["# %s -- ncu summaries (source: %s)\n" % (tag, src)]
    lp = os.path.join(src, "launches.csv")
    if os.path.exists(lp):
        agg = launches(lp)
        tot = sum(v[1] for v in agg.values())
        md.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised: compare SHARES)\n")
        md.append("command: `python bench.py --steps 1 --warmup 3 --no-cpu --apply-reps 5`\n")
        md.append("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append("| `%s` | %d | %.2f | %.1f | %.1f %% |" % (k, v[0], v[1] / 1e3, v[1] / v[0], 100 * v[1] / tot))
        md.append("")
    rp = os.path.join(src, "prof_top.ncu-rep")
    tpath = os.path.join("profiles", "ncu_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    lattice = int(sys.argv[3]) if len(sys.argv) > 3 else 4096   # lattice extent of the captured command
    if os.path.exists(rp):
        md.append("## `ncu --set full --clock-control none --import-source on` (per launch)\n")
        md.append("| kernel | time | DRAM read | DRAM write | traffic GB | GB/s | DRAM %% of ncu peak | warps active %% | regs | grid x block | L2 hit %% |\n|---|---|---|---|---:|---:|---|---|---|---|---|")
        seen = set()
        for d in full(rp):
            key = d["kernel"]
            if key in seen:
                continue
            seen.add(key)
            t_us = float(d["gpu__time_duration.sum"].split()[0].replace(",", ""))
            if d["gpu__time_duration.sum"].split()[1] == "ms":
                t_us *= 1e3
            tr = gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"])
            tkey = traffic_key(key)
            if tkey:
                traffic.setdefault(str(lattice), {})[tkey] = {"bytes": tr * 1e9, "kernel": key, "capture": tag,
                                                               "ncu_time_us": t_us}
            md.append("| `%s` | %s | %s | %s | %.3f | %.0f | %s | %s | %s | %s x %s | %s |" % (
                key, d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"], tr,
                tr / (t_us * 1e-6), d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "").split()[0],
                d.get("sm__warps_active.avg.pct_of_peak_sustained_active", "").split()[0],
                d.get("launch__registers_per_thread", "").split()[0], d.get("launch__grid_size", "").split()[0],
                d.get("launch__block_size", "").split()[0], d.get("lts__t_sector_hit_rate.pct", "").split()[0]))
        md.append("")
    for name in ("bench.json", "bench_sizes.json"):
        bp = os.path.join(src, name)
        if os.path.exists(bp):
            md.append("## %s (not under a profiler)\n" % name)
            for line in open(bp):
                line = line.strip()
                if not line.startswith("{"):
                    continue
                j = json.loads(line)
                r = j.get("roofline", {})
                md.append("* lattice %s: value %.0f GB/s (%.1f %% of measured HBM peak), %.2f ms/solve, %d iterations, "
                          "roofline kernel %.0f GB/s (frac %.3f, %.4f ms/launch), e2e %.0f GB/s, launches %d, clocks %s" % (
                              j["config"].get("lattice"), j["value"], 100 * j.get("frac_of_hbm_peak", 0), j["ms_per_step"],
                              j["config"].get("iterations", 0), r.get("achieved", 0), r.get("frac", 0),
                              r.get("ms_per_launch", 0), j.get("e2e", {}).get("value", 0), j.get("gpu_launches", 0),
                              json.dumps(j.get("clocks"))))
                for o in [r] + r.get("other_kernels", []):
                    if "kernel" in o:
                        md.append("  * `%s`: %.0f GB/s, frac %.3f, %.4f ms/launch%s" % (
                            o["kernel"].split(" (")[0], o.get("achieved", 0), o.get("frac", 0), o.get("ms_per_launch", 0),
                            (", share of step %.3f" % o["share_of_step"]) if "share_of_step" in o else ""))
                if "cpu_baseline" in j:
                    md.append("  * cpu_baseline: %s" % json.dumps(j["cpu_baseline"]))
            md.append("")
            with open(os.path.join("profiles", "%s_%s" % (tag, name)), "w") as f:
                f.write(open(bp).read())
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    open(os.path.join("profiles", "%s_summary.md" % tag), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
