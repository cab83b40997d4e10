"""Summarise an `ncu --set full` capture of the three passes of one Laplacian apply (tools/prof_lapl.py) as the JSON
file bench.py reads `roofline.traffic` and the kernel name from (profiles/r<round>_ncu_full_<n>.json).

usage: python tools/ncu_summary.py CAPTURE.ncu-rep OUT.json [--n 512] [--note "..."]
Runs `ncu -i CAPTURE --page raw --csv` (ncu is in the build container too: reading a capture needs no GPU)."""
import argparse
import csv
import io
import json
import subprocess

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"   # ..._<reason>_per_issue_active.ratio
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0,
              "nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("capture")
    ap.add_argument("out")
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.capture, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    names, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, nm in enumerate(names):
        col.setdefault(nm.split(".TriageCompute.")[-1] if ".TriageCompute." in nm else nm, i)
    kernels = []
    for r in data:
        kname = r[col["Kernel Name"]]
        if "x_tma_kernel" in kname or "xpass" in kname:
            which = "x"
        elif "yz_tma_kernel<0" in kname or "ypass" in kname:
            which = "y"
        elif "yz_tma_kernel<1" in kname or "zpass" in kname:
            which = "z"
        else:
            continue
        k = {"pass": which, "kernel": kname}
        for m in KEEP:
            if m in col and r[col[m]] not in ("", "no data"):
                v, u = float(r[col[m]].replace(",", "")), units[col[m]]
                if u in UNIT_SCALE and ("byte" in u):
                    v, u = v * UNIT_SCALE[u], "byte"
                elif u in UNIT_SCALE:
                    v, u = v * UNIT_SCALE[u], "us"
                k[m] = {"value": v, "unit": u}
        stalls = {}
        for nm, i in col.items():
            if nm.startswith(STALL) and nm.endswith("_per_issue_active.ratio") and r[i] not in ("", "no data"):
                stalls[nm[len(STALL):-len("_per_issue_active.ratio")]] = float(r[i])
        k["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        if "dram__bytes_read.sum" in k and "dram__bytes_write.sum" in k:
            k["dram_bytes_per_launch"] = k["dram__bytes_read.sum"]["value"] + k["dram__bytes_write.sum"]["value"]
        kernels.append(k)
    out = {"source": a.note or f"ncu -i {a.capture} --page raw --csv; ncu per-launch times are cold-cache and serialised",
           "grid": [a.n] * 3, "kernels": kernels}
    json.dump(out, open(a.out, "w"), indent=1)
    for k in kernels:
        print(k["pass"], k["kernel"][:90], k.get("gpu__time_duration.sum"), k.get("dram_bytes_per_launch"))


if __name__ == "__main__":
    main()
