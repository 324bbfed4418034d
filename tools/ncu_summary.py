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


if __name__ == "__main__":
    {"list": launch_list, "full": full}[sys.argv[1]](sys.argv[2])
