"""Turn the ncu outputs of scripts/profile_round.sh (gpurun_out/<tag>_*) into the text summaries committed under
profiles/ and into profiles/traffic.json (the DRAM bytes bench.py reports as `roofline.traffic`).
usage: python scripts/summarize_profiles.py <tag> <out-prefix>   (needs ncu, no GPU)"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, prefix = sys.argv[1], sys.argv[2]
G = "gpurun_out"


def launch_share(kind):
    path = f"{G}/{tag}_launches_{kind}.csv"
    if not os.path.exists(path):
        return f"# {path} missing\n"
    rows = list(csv.reader(open(path)))
    hdr = None
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        us = v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)
        tot[d["Kernel Name"][:110]] += us
        cnt[d["Kernel Name"][:110]] += 1
    s = sum(tot.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --workload {kind} --steps 3 --warmup 3: "
           f"{sum(cnt.values())} launches, {s / 1e3:.2f} ms of kernel time (cold-cache, serialised: shares, not absolutes)",
           f"# {'share':>7s} {'avg us':>11s} {'count':>6s}  kernel"]
    for k, v in tot.most_common(18):
        out.append(f"  {v / s * 100:6.2f}% {v / cnt[k]:11.1f} {cnt[k]:6d}  {k}")
    return "\n".join(out) + "\n"


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__inst_executed.sum"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def full(name, traffic=None):
    rep = f"{G}/{tag}_prof_{name}.ncu-rep"
    if not os.path.exists(rep):
        return f"# {rep} missing\n"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none --import-source on, one launch, from {os.path.basename(rep)}"]
    for r in rows[2:]:
        kname = r[hdr.index("Kernel Name")]
        out.append(f"kernel: {kname[:140]}")
        for k in KEYS:
            if k in hdr:
                out.append(f"  {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        if traffic is not None:
            b = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                b += float(r[hdr.index(k)].replace(",", "")) * UNIT[units[hdr.index(k)]]
            traffic.append((kname, b, float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")), os.path.basename(rep)))
    return "\n".join(out) + "\n"


os.makedirs("profiles", exist_ok=True)
open(f"profiles/{prefix}_launch_share_bench_render.txt", "w").write(launch_share("render"))
open(f"profiles/{prefix}_launch_share_bench_train.txt", "w").write(launch_share("train"))
tr, fw = [], []
with open(f"profiles/{prefix}_train_kernels_ncu.txt", "w") as fh:
    for n in ("fwdsave", "dgrad", "wgrad"):
        fh.write(full(n, tr) + "\n")
open(f"profiles/{prefix}_fwd_chain_ncu.txt", "w").write(full("fwd", fw))
open(f"profiles/{prefix}_composite_ncu.txt", "w").write(full("composite"))
for kind in ("render", "train"):
    if os.path.exists(f"{G}/{tag}_launches_{kind}.csv"):
        os.system(f"cp {G}/{tag}_launches_{kind}.csv profiles/{prefix}_launches_bench_{kind}.csv")
traffic = {}
if fw:
    traffic["chain_kernel<FwdEpi<false>>"] = {"dram_bytes": fw[0][1], "kernel_us": fw[0][2],
                                             "source": f"dram__bytes_read.sum + dram__bytes_write.sum of one 800x800x64 launch, ncu --set full, profiles/{prefix}_fwd_chain_ncu.txt ({fw[0][3]})"}
bwd = [t for t in tr if "Dgrad" in t[0] or "wgrad" in t[0]]
if len(bwd) == 2:
    traffic["train_step_backward"] = {"dram_bytes": bwd[0][1] + bwd[1][1], "kernel_us": bwd[0][2] + bwd[1][2],
                                      "source": f"dram bytes of one chain_kernel<DgradEpi> + one mlp_wgrad_tc_kernel launch (4096 rays x 64), profiles/{prefix}_train_kernels_ncu.txt"}
    traffic["train_step_mlp_kernels"] = {"dram_bytes": sum(t[1] for t in tr), "kernel_us": sum(t[2] for t in tr),
                                         "source": f"forward-with-saved-tiles + delta chain + wgrad, profiles/{prefix}_train_kernels_ncu.txt"}
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
print("written", list(traffic))
