#!/usr/bin/env python
"""Per-kernel table from a metrics-only ncu launch list (`ncu --metrics ... --csv --log-file X.csv bench.py ...`):
   python profiles/summarize_step.py X.csv [out.md] [traffic.json] [--step N]
--step N keeps only the N-th step of the capture (steps end with fuse_nms_kernel; 0 = the warm-up step, 1 = the timed one).
Groups launches by kernel (and template arguments), prints launches, total/avg device time, DRAM bytes per launch,
achieved DRAM GB/s and the pipe counters; writes the DRAM bytes per launch of the tcgen05 GEMM to traffic.json."""
import collections, csv, json, re, sys

STEP = None
if "--step" in sys.argv:
    i = sys.argv.index("--step")
    STEP = int(sys.argv[i + 1])
    del sys.argv[i:i + 2]
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
ix = {n: H.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launch = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) != len(H):
        continue
    d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[ix["Metric Name"]]] = v * scale


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("cone::", "").replace("<unnamed>::", "")
    return n[:46]


if STEP is not None:  # launches after the (STEP-1)-th fuse_nms up to and including the STEP-th one
    ids = list(launch)
    ends = [i for i, k in enumerate(ids) if "fuse_nms_kernel" in launch[k]["name"]]
    first_step = [i for i, k in enumerate(ids) if "l2norm_rows_kernel" in launch[k]["name"]]
    lo = (ends[STEP - 1] + 1) if STEP > 0 else min(i for i in first_step if i > 4)  # skip the weight-handle set-up launches
    keep = ids[lo:ends[STEP] + 1]
    launch = collections.OrderedDict((k, launch[k]) for k in keep)
groups = collections.OrderedDict()
for d in launch.values():
    groups.setdefault(short(d["name"]), []).append(d)
tot = sum(d.get("gpu__time_duration.sum", 0) for d in launch.values())
lines = ["| kernel | launches | total us | share | avg us | DRAM MB/launch | DRAM GB/s | dram % | tensor % | issue % | warps % | regs |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for k, ds in sorted(groups.items(), key=lambda kv: -sum(d.get("gpu__time_duration.sum", 0) for d in kv[1])):
    t = sum(d.get("gpu__time_duration.sum", 0) for d in ds)
    by = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in ds)
    avg = lambda m: sum(d.get(m, 0) * d.get("gpu__time_duration.sum", 0) for d in ds) / max(t, 1e-9)
    lines.append(f"| {k} | {len(ds)} | {t:.0f} | {100 * t / tot:.1f}% | {t / len(ds):.1f} | {by / len(ds) / 1e6:.1f} | {by / max(t, 1e-9) / 1e3:.0f} | "
                 f"{avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
                 f"{avg('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
                 f"{avg('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.0f} | "
                 f"{avg('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {ds[0].get('launch__registers_per_thread', 0):.0f} |")
    if k.startswith("tc_gemm_kernel"):
        traffic.setdefault("gemm_tc_launches", 0)
        traffic["gemm_tc_launches"] += len(ds)
        traffic["gemm_tc_bytes"] = traffic.get("gemm_tc_bytes", 0) + by
out = "\n".join(lines)
print(out)
print(f"\ntotal device time of the captured launches: {tot / 1e3:.2f} ms over {len(launch)} launches")
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + f"\n\ntotal device time of the captured launches: {tot / 1e3:.2f} ms over {len(launch)} launches\n")
if len(sys.argv) > 3:
    # DRAM traffic of the step and per kernel category (bench.py's profile categories), for bench.py's `roofline.traffic`
    cats = {"enc_tail": "enc_tail_kernel", "gemm_tc": "tc_gemm_kernel", "enc_attention": "enc_attention_f16_kernel",
            "dec_attention": "dec_cross_attention_mem_kernel", "gemm_fp32": "sgemm_nt_kernel"}
    kern = {}
    for cat, pat in cats.items():
        ds = [d for d in launch.values() if pat in d["name"]]
        if ds:
            by = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in ds)
            kern[cat] = {"launches": len(ds), "dram_bytes_per_launch": by / len(ds),
                         "time_us_per_launch": sum(d.get("gpu__time_duration.sum", 0) for d in ds) / len(ds)}
    step_bytes = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in launch.values())
    json.dump({"step_dram_bytes": step_bytes, "kernels": kern, "launches": len(launch),
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of ONE bench.py step (default "
                         "workload: 640 queries x 30 windows; --step 1 of a `--steps 1 --warmup 1` capture)"},
              open(sys.argv[3], "w"), indent=1)
