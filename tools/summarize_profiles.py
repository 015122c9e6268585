#!/usr/bin/env python
"""Turn the raw evidence a `tools/make_profiles.sh` run leaves in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/summarize_profiles.py sass  [TAG]     # no GPU needed: cuobjdump -sass of the shipped .so, per kernel
    python tools/summarize_profiles.py ncu   [TAG]     # launch list -> per-kernel shares; GEMM DRAM traffic (stamped with the
                                                       # sha256 of csrc/gemm_tcgen05.cu); --set full reports -> one row per capture
"""
import collections
import csv
import glob
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
RAW = os.path.join(ROOT, "gpurun_out")
SO = os.path.join(ROOT, "speechclip_b200", "libspeechclip_b200.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU.EX2", "SYNCS"]


def sass(tag):
    """Per kernel of the shipped library: how many tcgen05 MMAs (UTCHMMA), TMEM loads / stores (LDTM / STTM), TMA loads / stores
    (UTMALDG / UTMASTG) and legacy tensor-core instructions (HMMA) its SASS holds (B200_PROFILING.md, "What proves a
    Blackwell-native kernel")."""
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = cur.split("(")[0][-110:]
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = counts[cur]
        c["instructions"] += 1
        for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "SYNCS"):
            if op.startswith(k):
                c[k] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            c["UTCHMMA.2CTA"] += 1
        if op.startswith("MUFU.EX2"):
            c["MUFU.EX2"] += 1
    path = os.path.join(OUT, f"{tag}_sass_summary.csv")
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "instructions"] + MNEMONICS)
        tot = collections.Counter()
        for k, c in counts.items():
            if c["instructions"] == 0:
                continue
            w.writerow([k, c["instructions"]] + [c[m] for m in MNEMONICS])
            tot.update(c)
        w.writerow(["TOTAL (libspeechclip_b200.so, sm_100a)", tot["instructions"]] + [tot[m] for m in MNEMONICS])
    print("wrote", path, {m: tot[m] for m in MNEMONICS})


def _ncu_csv_rows(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")))
    hdr = rows[0]
    return hdr, rows[1:]


def ncu(tag):
    # 1. launch list -> per-kernel totals and shares
    lp = os.path.join(RAW, f"{tag}_launches.csv")
    if os.path.exists(lp):
        hdr, rows = _ncu_csv_rows(lp)
        ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
        ui = hdr.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in rows:
            if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
                continue
            us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
            name = re.sub(r"\(anonymous namespace\)::|scb::|<unnamed>::", "", r[ki]).split("(")[0][:90]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += us
        total = sum(v[1] for v in agg.values())
        with open(os.path.join(OUT, f"{tag}_launches_summary.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "share"])
            for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, n, f"{us:.1f}", f"{us / total:.4f}"])
            w.writerow(["TOTAL", sum(v[0] for v in agg.values()), f"{total:.1f}", "1.0"])
        print("wrote", f"{tag}_launches_summary.csv", f"{total / 1e3:.2f} ms under ncu")
    # 2. GEMM DRAM traffic, stamped with the GEMM source it was captured on
    dp = os.path.join(RAW, f"{tag}_gemm_dram.csv")
    if os.path.exists(dp):
        hdr, rows = _ncu_csv_rows(dp)
        mi, vi, ui, ii = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
        by_id = collections.defaultdict(dict)
        for r in rows:
            if len(r) <= vi:
                continue
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
            by_id[r[ii]][r[mi]] = float(r[vi].replace(",", "")) * scale
        ids = sorted(by_id, key=int)
        n = 121 if len(ids) >= 121 else len(ids)   # the GEMM launches of the timed step are the last ones of the command
        last = ids[-n:]
        total = sum(by_id[i].get("dram__bytes_read.sum", 0.0) + by_id[i].get("dram__bytes_write.sum", 0.0) for i in last)
        sha = hashlib.sha256(open(os.path.join(ROOT, "speechclip_b200", "csrc", "gemm_tcgen05.cu"), "rb").read()).hexdigest()
        rec = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm on `python bench.py --steps 1 --warmup 3 "
                         "--graphs off` (tools/make_profiles.sh), the GEMM launches of the timed step",
               "captured": f"profiles set {tag}", "gemm_source_sha256": sha, "launches": n, "dram_bytes_per_step": total,
               "dram_bytes_per_launch": total / max(n, 1),
               "gemm_us_per_step_under_ncu": sum(by_id[i].get("gpu__time_duration.sum", 0.0) for i in last)}
        json.dump(rec, open(os.path.join(OUT, "gemm_dram_traffic.json"), "w"), indent=1)
        print("wrote gemm_dram_traffic.json", rec["dram_bytes_per_launch"])
    # 3. --set full reports -> one row each
    want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__cycles_elapsed.max"]
    reps = sorted(glob.glob(os.path.join(RAW, f"{tag}_full_*.ncu-rep")))
    if reps:
        with open(os.path.join(OUT, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["capture", "kernel"] + want)
            for rp in reps:
                txt = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
                rows = list(csv.reader(txt.splitlines()))
                if len(rows) < 3:
                    continue
                hdr, vals = rows[0], rows[-1]
                d = dict(zip(hdr, vals))
                w.writerow([os.path.basename(rp)[:-8], d.get("Kernel Name", "")[:80]] + [d.get(k, "") for k in want])
        print("wrote", f"{tag}_ncu_full_summary.csv", len(reps), "captures")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "sass"
    tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
    {"sass": sass, "ncu": ncu}[what](tag)
