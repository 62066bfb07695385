"""Condense an ncu launch list of ONE reverse step (long-format CSV, a few metrics per launch) into profiles/ncu_step_summary.json.

    python tools/ncu_launch_summary.py gpurun_out/<run>/launches.csv gpurun_out/<run>/layers.json profiles/ncu_step_summary.json

The summary carries the `plan_signature` of the layer plan it was captured on (taken from a `bench.py --dump-layers` file of the
same build); `bench.py` reports `roofline.traffic` from it only while the running plan has the same signature."""
import collections
import csv
import json
import re
import sys

WANT = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct"}
SCALE = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def main(src, layers, dst, first_kernel="chain_step_begin_kernel", last_kernel="chain_advance_kernel"):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    iid, ik, im, iu, iv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv or r[im] not in WANT:
            continue
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|ndiff::", "", r[ik])
        name = re.sub(r"^void ", "", name).split("(")[0]
        d = launches.setdefault(r[iid], {"kernel": name})
        d[WANT[r[im]]] = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1.0)
    L = list(launches.values())
    # keep the LAST complete step: from the last step prologue to the step counter after it
    starts = [i for i, l in enumerate(L) if l["kernel"].startswith(first_kernel)]
    ends = [i for i, l in enumerate(L) if l["kernel"].startswith(last_kernel)]
    if starts and ends and ends[-1] > starts[-1]:
        L = L[starts[-1]:ends[-1] + 1]
    tot = sum(l.get("time_us", 0.0) for l in L)
    agg = collections.OrderedDict()
    for l in L:
        base = l["kernel"].split("<")[0]
        a = agg.setdefault(base, collections.Counter())
        a["launches"] += 1
        for k in ("time_us", "dram_read", "dram_write"):
            a[k] += l.get(k, 0.0)
        a["tensor_pipe_pct_x_time"] += l.get("tensor_pipe_pct", 0.0) * l.get("time_us", 0.0)
    fam = {}
    for k, a in agg.items():
        fam[k] = {"launches": int(a["launches"]), "time_us": round(a["time_us"], 1), "share_of_step": round(a["time_us"] / tot, 4),
                  "dram_bytes_per_launch": int((a["dram_read"] + a["dram_write"]) / a["launches"]),
                  "achieved_dram_gbs": round((a["dram_read"] + a["dram_write"]) / (a["time_us"] * 1e-6) / 1e9, 1) if a["time_us"] else None,
                  "tensor_pipe_pct_timeweighted": round(a["tensor_pipe_pct_x_time"] / a["time_us"], 2) if a["time_us"] else None}
    sig = json.load(open(layers)).get("plan_signature")
    out = {"note": "ncu --metrics <time, dram bytes, tensor pipe, l1tex, issue> --clock-control none, python tools/profile_step.py 64 2 "
                   "(64 patches of 4x256x256, eager launches; the last complete reverse step). Per-launch times are serialised and "
                   "cold-cache: compare shares, not absolutes.",
           "plan_signature": sig, "step_launches": len(L), "step_time_us": round(tot, 1),
           "conv_gemm": fam.get("conv_gemm_kernel", {}), "kernels": fam,
           "launch_list": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in l.items()} for l in L]}
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(L)} launches, {tot:.1f} us; conv_gemm {fam.get('conv_gemm_kernel')}; signature {sig}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
