"""Turn ncu outputs into the small text summaries kept under profiles/.
  python tools/ncu_summary.py list  gpurun_out/launches.csv            -> per-kernel totals of a launch list
  python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep            -> key metrics per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launch_list(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0.0, 0])
    total = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else (v if unit.startswith("m") else v * 1e3))
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("svdpp::", "").replace("void ", "")
        agg[name][0] += ms
        agg[name][1] += 1
        total += ms
    print(f"# per-kernel device time of one eager SVD-XT UNet step under ncu (cold cache, serialised): total {total:.2f} ms")
    print(f"{'ms':>10} {'share':>7} {'launches':>9}  kernel")
    for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:10.3f} {100 * ms / total:6.1f}% {n:9d}  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[idx["Kernel Name"]][:110])
        for k in KEYS:
            if k in idx:
                print(f"  {k:75s} {r[idx[k]]:>16s} {units[idx[k]]}")
        print()


def traffic(path, out_json=None, note=""):
    """DRAM bytes of the tcgen05 GEMM / conv launches of one forward (ncu --metrics dram__bytes_read.sum,
    dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc): the `roofline.traffic` source of bench.py."""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    per_id = collections.defaultdict(dict)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"].lower()
        name = row["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        elif name.startswith("gpu__time"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)   # -> ms
        per_id[row["ID"]][name] = v
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in per_id.values())
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in per_id.values())
    ms = sum(d.get("gpu__time_duration.sum", 0.0) for d in per_id.values())
    res = dict(launches=len(per_id), dram_read_bytes=rd, dram_write_bytes=wr,
               traffic_bytes_per_launch=(rd + wr) / max(len(per_id), 1), ncu_time_ms=ms,
               source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc over one "
                      "eager SVD-XT UNet forward (25 frames, 72x128 latent)" + (", " + note if note else ""))
    print(json.dumps(res, indent=1))
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"list": launch_list, "full": full}[sys.argv[1]](sys.argv[2])
