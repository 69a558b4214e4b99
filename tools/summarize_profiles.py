"""Turn the ncu artefacts brought back in gpurun_out/ into the small tracked files under profiles/.

usage: python tools/summarize_profiles.py <tag> <report.ncu-rep> [envs]
writes  profiles/<tag>_metrics.csv   selected raw metrics of the step kernel (one row per metric)
        profiles/<tag>_hot.txt       opcode histogram, stall reasons, hottest SASS lines
        profiles/traffic.json        DRAM bytes per launch of the newest capture (read by bench.py)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep = sys.argv[1], sys.argv[2]
envs = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 24
KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed launch__registers_per_thread launch__grid_size launch__block_size
launch__occupancy_limit_registers sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active lts__t_sector_hit_rate.pct
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed dram__sectors_read.sum dram__sectors_write.sum""".split()

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", tag + "_metrics.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit"] + ["launch%d" % i for i in range(len(data))])
    kname = data[0][hdr.index("Kernel Name")]
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            w.writerow([kname, m, units[i]] + [d[i] for d in data])
d0 = data[-1]


def val(name):
    i = hdr.index(name)
    v = float(d0[i].replace(",", ""))
    u = units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)


traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
# traffic.json feeds bench.py's roofline of the DEFAULT env step kernel only (not the penalty instantiation, not the generic step)
STEP_KERNEL = ("::step_kernel<0" in kname or "::step_kernel<false" in kname) and "generic" not in kname
if STEP_KERNEL:
  json.dump({"envs": envs, "dram_bytes_per_launch": traffic, "dram_bytes_read": val("dram__bytes_read.sum"),
           "dram_bytes_write": val("dram__bytes_write.sum"), "algorithmic_bytes_per_launch": 93 * envs,
           "source": os.path.basename(rep), "kernel": kname}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tmp = os.path.join("/tmp", tag + "_src.csv")
open(tmp, "w").write(src)
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), tmp, "20"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", tag + "_hot.txt"), "w").write(
    "# %s  (ncu --set full --clock-control none --import-source on; %d envs; numbers are for %d profiled launch(es))\n%s" % (
        os.path.basename(rep), envs, len(data), hot))
print("wrote profiles/%s_metrics.csv, profiles/%s_hot.txt, profiles/traffic.json (traffic %.1f MB vs algorithmic %.1f MB)" % (
    tag, tag, traffic / 1e6, 93 * envs / 1e6))
