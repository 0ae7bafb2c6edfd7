"""Selected raw metrics of an .ncu-rep (per launch) -> CSV on stdout, plus per-kernel DRAM traffic averages on stderr."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, units = r[0], r[1]
KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "launch__occupancy_limit_shared_mem")
keep = [i for i, h in enumerate(hdr) if h in ("Kernel Name", "Grid Size", "Block Size") or h in KEYS]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] + (" [" + units[i] + "]" if units[i] else "") for i in keep])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


for row in r[2:]:
    w.writerow([row[i] for i in keep])
    k = row[ik].split("(")[0]
    a = agg[k]
    a[0] += 1
    a[1] += to_bytes(row[ir], units[ir])
    a[2] += to_bytes(row[iw], units[iw])
    a[3] += float(row[it].replace(",", ""))
for k, a in agg.items():
    print(f"{k}: launches {a[0]}, avg dram read {a[1] / a[0] / 1e6:.3f} MB, write {a[2] / a[0] / 1e6:.3f} MB, avg duration {a[3] / a[0]:.2f} {units[it]}", file=sys.stderr)
