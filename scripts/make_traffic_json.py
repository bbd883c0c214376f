"""ncu launch list of bench.py's own first step (gpurun_out/r02_bench_launch_metrics.csv, scripts/r02_ncu3.sh) -> profiles/r02_bench_launch_traffic.json
(DRAM bytes and fp64 instruction counts of every chain_check launch of one step; read by bench.py for roofline.traffic / roofline_fp64) and a table."""
import csv, json, sys, collections, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_bench_launch_metrics.csv")
rows = list(csv.reader(open(src)))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i
        break
idx = {h: i for i, h in enumerate(hdr)}
d = collections.OrderedDict()
for r in rows[start + 2:]:
    if len(r) < len(hdr):
        continue
    d.setdefault((int(r[idx["ID"]]), r[idx["Kernel Name"]], r[idx["Grid Size"]], r[idx["Block Size"]]), {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
launches = []
first = None
for (i, name, grid, block), v in d.items():
    key = (name, grid, block)
    if first is None:
        first = key
    elif key == first:       # the next batch starts with the same shortest-window launch
        break
    fl = 2 * v["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + v["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] + v["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
    launches.append({"kernel": name.replace("void ", "").replace("(BatchArgs)", ""), "grid": grid, "block": block, "ms_under_ncu": v["gpu__time_duration.sum"] / 1e6,
                     "dram_bytes_read": v.get("dram__bytes_read.sum", 0.0), "dram_bytes_write": v.get("dram__bytes_write.sum", 0.0), "fp64_flops": fl,
                     "warp_instructions": v.get("smsp__inst_executed.sum", 0.0), "registers": int(v.get("launch__registers_per_thread", 0))})
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum "
                 "--clock-control none -k regex:chain_check -c 6: python bench.py --steps 1 --warmup 0 --no-cpu --no-full-retries (scripts/r02_ncu3.sh; "
                 "profiles/r02_bench_launch_metrics.csv)",
       "workload": "M3500 full matrix, 2 242 424 checks, every chain_check launch of the first step",
       "dram_bytes_read": sum(l["dram_bytes_read"] for l in launches), "dram_bytes_write": sum(l["dram_bytes_write"] for l in launches),
       "fp64_flops": sum(l["fp64_flops"] for l in launches), "ms_under_ncu": sum(l["ms_under_ncu"] for l in launches), "launches": launches}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_bench_launch_traffic.json"), "w"), indent=1)
for l in launches:
    print("%-40s %-14s %9.1f ms  DRAM rd %8.2f GB  wr %8.2f GB  fp64 %.3e flop  %d regs" % (l["kernel"], l["grid"], l["ms_under_ncu"], l["dram_bytes_read"] / 1e9, l["dram_bytes_write"] / 1e9, l["fp64_flops"], l["registers"]))
print("step: %.2f s under ncu, DRAM %.1f GB read + %.1f GB written, %.3e fp64 flop" % (out["ms_under_ncu"] / 1e3, out["dram_bytes_read"] / 1e9, out["dram_bytes_write"] / 1e9, out["fp64_flops"]))
