"""Condense `ncu --page raw --csv` of one reverse step (every launch, --set full) into a small JSON for profiles/.

    python tools/ncu_step_summary.py gpurun_out/ncu_r1_final_step_raw.csv profiles/ncu_r1_final_step.json
"""
import collections
import csv
import json
import re
import sys

KEEP = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_mb",
    "dram__bytes_write.sum": "dram_write_mb",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max": "sm_cycles",
}
UNIT = {"time_us": 1e-3, "dram_read_mb": 1.0 / 2 ** 20, "dram_write_mb": 1.0 / 2 ** 20}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    cols = {h: i for i, h in enumerate(hdr)}
    scale = {}
    for h in KEEP:
        u = units[cols[h]] if h in cols else ""
        scale[h] = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "byte": 1.0,
                                                        "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    launches = []
    for r in rows[2:]:
        if len(r) <= ik:
            continue
        name = re.sub(r"^(void )?.*?unnamed>::", "", r[ik])
        name = re.sub(r"\((int|bool)\)", "", name).split("(")[0]
        d = {"kernel": name}
        for h, k in KEEP.items():
            if h in cols and r[cols[h]] not in ("", "n/a"):
                v = float(r[cols[h]].replace(",", "")) * scale[h]
                d[k] = round(v * UNIT.get(k, 1.0), 3)
        launches.append(d)
    agg = collections.OrderedDict()
    tot = sum(l.get("time_us", 0) for l in launches)
    for l in launches:
        a = agg.setdefault(l["kernel"], {"launches": 0, "time_us": 0.0, "dram_read_mb": 0.0, "dram_write_mb": 0.0, "tensor_pipe_pct_timeweighted": 0.0,
                                         "l1tex_pct_timeweighted": 0.0, "regs": l.get("regs")})
        a["launches"] += 1
        t = l.get("time_us", 0)
        a["time_us"] += t
        a["dram_read_mb"] += l.get("dram_read_mb", 0)
        a["dram_write_mb"] += l.get("dram_write_mb", 0)
        a["tensor_pipe_pct_timeweighted"] += t * l.get("tensor_pipe_pct", 0)
        a["l1tex_pct_timeweighted"] += t * l.get("l1tex_pct", 0)
    for a in agg.values():
        for k in ("tensor_pipe_pct_timeweighted", "l1tex_pct_timeweighted"):
            a[k] = round(a[k] / a["time_us"], 2) if a["time_us"] else 0.0
        a["share_of_step"] = round(a["time_us"] / tot, 4)
        for k in ("time_us", "dram_read_mb", "dram_write_mb"):
            a[k] = round(a[k], 1)
    conv = [l for l in launches if l["kernel"].startswith("conv_gemm_kernel")]
    out = {
        "note": "ncu --set full --clock-control none --launch-skip 144; python tools/profile_step.py 64 1 (64 patches of 4x256x256, one "
                "reverse step, eager launches). Per-launch times are serialised and cold-cache: compare shares, not absolutes. "
                "If step_launches < 120 the capture was cut by the box's time limit (the tail of the step — ups.3.1 onward, i.e. "
                "more of the same 64-channel kernels — is missing); profiles/launches_r1_step_mb64.csv is the complete launch list.",
        "step_launches": len(launches), "step_time_us": round(tot, 1),
        "conv_gemm": {"launches": len(conv), "time_us": round(sum(l["time_us"] for l in conv), 1),
                      "share_of_step": round(sum(l["time_us"] for l in conv) / tot, 4),
                      "dram_bytes_per_launch": int(sum((l.get("dram_read_mb", 0) + l.get("dram_write_mb", 0)) for l in conv) * 2 ** 20 / max(len(conv), 1))},
        "kernels": sorted(({"kernel": k, **v} for k, v in agg.items()), key=lambda x: -x["time_us"]),
        "launch_list": launches,
    }
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(launches)} launches, {tot:.1f} us; conv share {out['conv_gemm']['share_of_step']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
